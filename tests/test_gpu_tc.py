"""-m gpu tests of the tcgen05 tensor-core path (csrc/sacx_tc.cuh): MLP GEMMs at large batch.

The path must agree with the FFMA tile-parallel kernel (SACX_TC=0) -- itself pinned to the reference's recorded vectors by
test_gpu_parity.py -- on every intermediate of the update: forward activations (EPI_FWD tiles), backward deltas
(EPI_DACT tiles), and parameters / Adam moments / targets after the step (EPI_DW tiles + the reduce/optimiser kernel).
Tolerances: 3xTF32 (hi/lo split of both operands, fp32 accumulate in TMEM) and a different summation order, not
different math -> rel-L2 2e-5 on activations / targets, 1e-4 on deltas and on parameters after two free-running steps
(the same bars as the row-parallel path's tests)."""
import numpy as np
import pytest
import torch

from gpu_helpers import assert_close, base_config, fill_ring, load_nets
from test_gpu_parity import _random_nets

pytestmark = pytest.mark.gpu

SNAP = ("batch.sa", "batch.s2a", "batch.spi", "batch.r", "batch.d", "out.y", "out.logpi", "out.logpi_next", "out.q1", "out.q2",
        "out.tq1", "out.tq2", "out.q1_pi", "out.q2_pi", "scr.dhead", "scr.dout1", "act.pia.h0", "act.q1.h1", "delta.q1.0",
        "delta.pi.0", "block.params", "block.targets", "block.m", "block.v")


def _engine(obs, act, hp, hq, B, actfn, monkeypatch, tc, min_batch=1024, cap=None, fill=None, scale=0.15):
    from sac.engine import UpdateEngine
    from sac.replay_buffer import ReplayBuffer
    monkeypatch.setenv("SACX_TC", "1" if tc else "0")
    monkeypatch.setenv("SACX_ROWPAR", "0")               # the comparison partner is the FFMA tile-parallel kernel
    monkeypatch.setenv("SACX_TC_MIN_BATCH", str(min_batch))
    cap = cap or 2 * B
    fill = fill or cap - 7
    cfg = base_config(hidden=hp, q_hidden=hq, act=actfn, batch=B, capacity=cap)
    eng = UpdateEngine(obs, act, cfg)
    load_nets(eng, _random_nets(obs, act, hp, hq, scale=scale))
    eng.reset_state()
    rb = ReplayBuffer(cap, obs, act)
    fill_ring(rb, fill, obs, act)
    eng.attach_ring(rb)
    return eng


def test_tc_path_selection(monkeypatch):
    eng = _engine(24, 4, (256, 256), (256, 256), 2048, "relu", monkeypatch, True)
    on, why, n0 = eng.tensor_core()
    assert on and why == "" and n0 == 0
    assert eng.path()[0] == "tiles"                    # the row-parallel latency kernel steps aside at large batch
    eng.update(None, None, None, 1)
    eng.sync()
    assert eng.tensor_core()[2] > 0                    # tcgen05 kernels actually launched
    small = _engine(24, 4, (256, 256), (256, 256), 256, "relu", monkeypatch, True)
    on, why, _ = small.tensor_core()
    assert not on and "batch" in why
    off = _engine(24, 4, (256, 256), (256, 256), 2048, "relu", monkeypatch, False)
    assert off.tensor_core()[:2] == (False, "disabled by SACX_TC=0")
    elu = _engine(24, 4, (256, 256), (256, 256), 2048, "elu", monkeypatch, True)
    assert not elu.tensor_core()[0]                    # saved pre-activations: FFMA tiles


CASES = [
    # obs, act, hidden_pi, hidden_q, B, activation
    (24, 4, (256, 256), (256, 256), 2048, "relu"),       # BASELINE config 5 shape (per-rank slice of the 65536 batch)
    (32, 2, (256, 256), (256, 256), 1024, "relu"),       # Donkey latent shape, batch 1024
    (17, 6, (64, 128), (128, 64), 1100, "tanh"),         # ragged batch (8 full row tiles + 76 rows), mixed widths < 256
    (11, 3, (128, 48, 256), (80, 256, 128), 1536, "leaky_relu"),   # three hidden layers; widths that are not multiples of 32
    (8, 1, (512, 256), (256, 512), 1280, "relu"),        # layers wider than 256 stay on FFMA tiles next to tensor-core layers; 2A = 2
    (3, 2, (256,), (128,), 1024, "tanh"),                # a single hidden layer; K = 3 / 5 first layers (small-K tiles)
    (24, 4, (64, 64, 64, 64), (256, 256, 256, 256), 1024, "relu"),   # four hidden layers
]


@pytest.mark.parametrize("obs,act,hp,hq,B,actfn", CASES)
def test_tc_matches_ffma_path(obs, act, hp, hq, B, actfn, monkeypatch):
    """Two free-running updates on identical inputs through (a) the tensor-core path, (b) the FFMA kernel with 64x64 tiles,
    (c) the FFMA kernel with 32x32 tiles. (b) vs (c) is the noise floor of fp32 summation order for this configuration:
    SAC's update is discontinuous in places (torch.min routing, relu', the sign of a near-zero gradient in Adam's first
    step: SURVEY F16), so a handful of elements may legitimately flip between two correct fp32 implementations. The
    tensor-core path has to sit at that floor: err(a,b) <= max(tol, 4 err(c,b)) for every tensor of the update."""
    rng = np.random.default_rng(5)
    K = 2
    cap = 2 * B
    idx = np.stack([rng.choice(cap - 7, B, replace=False) for _ in range(K)]).astype(np.int64)
    e1 = rng.standard_normal((K, B, act)).astype(np.float32)
    e2 = rng.standard_normal((K, B, act)).astype(np.float32)
    out = {}
    for mode in ("tc", "ffma", "ffma_small"):
        monkeypatch.setenv("SACX_TILE", "small" if mode == "ffma_small" else "large")
        eng = _engine(obs, act, hp, hq, B, actfn, monkeypatch, mode == "tc")
        assert eng.tensor_core()[0] == (mode == "tc"), eng.tensor_core()
        snaps = []
        for k in range(K):
            m = eng.update_host(idx[k], e1[k], e2[k], 1)
            assert m["nonfinite"] == 0 and m["updates"] == k + 1
            names = [n for n in SNAP if n in eng.layout]
            names += [n for n in eng.layout if n.startswith(("m.q1.", "m.q2.", "act.q", "act.pi"))]
            snap = {n: eng.view(n).cpu().numpy().copy() for n in names}
            snap["metrics"] = m
            snaps.append(snap)
        if mode == "tc":
            assert eng.tensor_core()[2] > 0
        out[mode] = snaps
    from helpers import rel_l2

    def batch_rows(x):
        if x.ndim == 2 and x.shape[0] == B:
            return x
        if x.ndim == 2 and x.shape == (1, B):
            return x.T
        return None

    flipped = False
    for k in range(K):
        a, b, c = out["tc"][k], out["ffma"][k], out["ffma_small"][k]
        for n in ("batch.sa", "batch.r", "batch.d"):
            assert np.array_equal(a[n], b[n]), n
        tol = 2e-5 * (3 ** k)
        names = [n for n in a if n not in ("metrics", "batch.sa", "batch.r", "batch.d") and np.any(b[n])]
        # tensors with one row per transition first: a flip shows up there as a few outlier rows
        for n in [n for n in names if batch_rows(b[n]) is not None]:
            lim = max(tol, 4 * rel_l2(c[n], b[n]), 2e-3 if flipped else 0.0)
            if rel_l2(a[n], b[n]) < lim:
                continue
            ra, rb = batch_rows(a[n]).astype(np.float64), batch_rows(b[n]).astype(np.float64)
            row_err = ((ra - rb) ** 2).sum(axis=1)
            keep = np.argsort(row_err)[: B - max(1, B // 200)]            # drop the worst 0.5% of the rows
            e = np.sqrt(row_err[keep].sum() / (rb[keep] ** 2).sum())
            assert e < lim, f"step{k} {n}: rel-L2 {e:.3e} >= {lim:.1e} even without the worst rows"
            flipped = True
        # weight-shaped tensors (parameters, targets, Adam moments = gradients): a flipped row moves a whole gradient by
        # ~|row| / sqrt(B), so they are only held to the tight bar while no flip has been seen
        for n in [n for n in names if batch_rows(b[n]) is None]:
            base = 1e-4 * (2 ** k) if n in ("block.params", "block.targets") else 5 * tol
            lim = max(base, 4 * rel_l2(c[n], b[n]), 2e-2 if flipped else 0.0)
            assert_close(f"step{k} {n}", a[n], b[n], lim)
        for key in ("q1_loss", "q2_loss", "policy_loss", "alpha_loss", "log_alpha"):
            fl = abs(c["metrics"][key] - b["metrics"][key])
            assert abs(a["metrics"][key] - b["metrics"][key]) <= max(1e-4 * abs(b["metrics"][key]) + 1e-6, 4 * fl), key
    # the first update's forward passes and targets come before anything that can flip: always tight
    a, b = out["tc"][0], out["ffma"][0]
    for n in a:
        if n.startswith(("act.pit", "act.qt", "out.tq", "out.y", "out.q1", "out.q2", "out.logpi", "scr.dout")) \
                and not n.endswith("_pi") and np.any(b[n]):
            assert_close(f"first update {n}", a[n], b[n], 2e-5)


def test_tc_gradient_entry_points_match_ffma(monkeypatch):
    """Data-parallel building blocks (critic/actor gradient plans + flat Adam apply) through the tensor-core path: the stored
    gradient blocks equal the FFMA path's."""
    obs, act, B = 24, 4, 2048
    rng = np.random.default_rng(9)
    idx = rng.choice(2 * B - 7, B, replace=False).astype(np.int64)
    e1 = rng.standard_normal((B, act)).astype(np.float32)
    e2 = rng.standard_normal((B, act)).astype(np.float32)
    res = {}
    for tc in (True, False):
        eng = _engine(obs, act, (256, 256), (256, 256), B, "relu", monkeypatch, tc)
        eng.sample_batch(torch.as_tensor(idx).cuda())
        eng.target(torch.as_tensor(e1).cuda())
        eng.critic_step(None, grads_only=True)
        eng.sync()
        gq = eng.view("block.g.critics").cpu().numpy().copy()
        eng.apply_grads(1, polyak=True)
        eng.actor_step(torch.as_tensor(e2).cuda(), None, grads_only=True)
        eng.sync()
        gp = eng.view("block.g.policy").cpu().numpy().copy()
        eng.apply_grads(2 | 4)
        eng.sync()
        res[tc] = (gq, gp, eng.view("block.params").cpu().numpy().copy())
    assert_close("critic grads", res[True][0], res[False][0], 1e-4)
    assert_close("policy grads", res[True][1], res[False][1], 1e-4)
    assert_close("params", res[True][2], res[False][2], 1e-5)


def test_tc_multi_step_launch_and_device_rng(monkeypatch):
    """n updates in one call == n calls of one update, bit for bit (same kernels, same order), with the device RNG."""
    a = _engine(24, 4, (256, 256), (256, 256), 2048, "relu", monkeypatch, True)
    b = _engine(24, 4, (256, 256), (256, 256), 2048, "relu", monkeypatch, True)
    a.update(None, None, None, 3)
    for _ in range(3):
        b.update(None, None, None, 1)
    a.sync(); b.sync()
    assert np.array_equal(a.view("block.params").cpu().numpy(), b.view("block.params").cpu().numpy())
    assert a.metrics()["updates"] == 3 and b.metrics()["updates"] == 3
    assert a.metrics()["nonfinite"] == 0


def test_tc_population_matches_ffma_population(monkeypatch):
    """Population mode (BASELINE config 3 shape: obs 4, act 1, 2x256, batch 256): 64 independent agents through the
    tensor-core path (3-D tensor maps: agent = third coordinate; one launch per phase covers every agent) against the
    one-CTA-per-agent FFMA kernel. Per-agent state after two device-RNG updates: every agent within the parameter bar
    except at most a few hit by an fp32-order flip; no agent may be far off."""
    from sac.population import SACPopulation
    obs, act, B, n = 4, 1, 256, 64
    cfg = base_config(hidden=(256, 256), batch=B, capacity=2000, rng="device")
    res = {}
    for tc in (True, False):
        monkeypatch.setenv("SACX_TC", "1" if tc else "0")
        monkeypatch.setenv("SACX_TC_MIN_BATCH", "4096")
        pop = SACPopulation(obs, act, cfg, n, reference_init=False)
        on, why, _ = pop.engine.tensor_core()
        assert on == tc, (on, why)
        s, a, r, s2, d = (torch.from_numpy(x).cuda() for x in __import__("helpers").synth_transitions(1500, obs, act, 3))
        pop.push_device_all(s, a, r, s2, d.float())
        pop.engine.update(None, None, None, 2)
        pop.engine.sync()
        if tc:
            assert pop.engine.tensor_core()[2] > 0
        res[tc] = {k: pop.engine.population_view(k).cpu().numpy().copy() for k in ("block.params", "block.targets", "block.m", "out.y")}
        ups = [pop.engine.metrics(ag)["updates"] for ag in (0, n - 1)]
        assert ups == [2, 2] and pop.engine.metrics(0)["nonfinite"] == 0
    from helpers import rel_l2
    for k, bar in (("out.y", 1e-4), ("block.params", 2e-4), ("block.targets", 2e-5)):
        errs = np.array([rel_l2(res[True][k][ag], res[False][k][ag]) for ag in range(n)])
        assert np.median(errs) < bar / 4, (k, np.median(errs))
        assert (errs > bar).sum() <= 3, (k, np.sort(errs)[-5:])
        assert errs.max() < 100 * bar, (k, errs.max())
    # agents are independent: different seeds -> different parameters
    assert not np.array_equal(res[True]["block.params"][0], res[True]["block.params"][1])


def test_tc_full_size_batch_65536(monkeypatch):
    """BASELINE config 5 at its full size (BipedalWalker shape, batch 65536, one rank): one update through the tensor-core
    path against the FFMA path on the same device-generated batch. 512 row tiles per layer, 64 batch splits per dW. Also
    the size-independent property of the update: the Polyak step is exactly tau * theta + (1 - tau) * target."""
    obs, act, B = 24, 4, 65536
    res = {}
    for tc in (True, False):
        eng = _engine(obs, act, (256, 256), (256, 256), B, "relu", monkeypatch, tc, min_batch=4096, cap=100_000, fill=100_000 - 3, scale=0.1)
        assert eng.tensor_core()[0] == tc
        t0 = eng.view("block.targets").clone()
        eng.update(None, None, None, 1)
        eng.sync()
        m = eng.metrics()
        assert m["nonfinite"] == 0 and m["updates"] == 1
        res[tc] = {n: eng.view(n).cpu().numpy().copy() for n in ("batch.idx", "out.y", "out.q1", "out.logpi", "block.params", "block.targets")}
        res[tc]["m"] = {n: eng.view(n).cpu().numpy().copy() for n in ("m.q1.W1", "m.q2.W0", "m.q1.b1", "m.q1.W2", "m.pi.W1")}
        res[tc]["metrics"] = m
        # Polyak exactness (agent.py:288-291): target' = tau * theta' + (1 - tau) * target, products rounded separately
        q_on = np.concatenate([eng.view(f"{c}.{t}").cpu().numpy().ravel() for c in ("q1", "q2") for t in ("W0", "b0", "W1", "b1", "W2", "b2")])
        q_tg = np.concatenate([eng.view(f"{c}t.{t}").cpu().numpy().ravel() for c in ("q1", "q2") for t in ("W0", "b0", "W1", "b1", "W2", "b2")])
        q_t0 = np.concatenate([t0.cpu().numpy().ravel()])[: q_tg.size]
        tau = np.float32(0.005)
        want = (tau * q_on).astype(np.float32) + (np.float32(1.0 - 0.005) * q_t0[: q_on.size]).astype(np.float32)
        if q_t0.size >= q_on.size and eng.view("block.targets").numel() == q_on.size:
            assert np.array_equal(q_tg, want)
    a, b = res[True], res[False]
    assert np.array_equal(a["batch.idx"], b["batch.idx"])                 # same device index stream (distinct indices)
    assert len(np.unique(a["batch.idx"])) == B
    for n in ("out.y", "out.q1", "out.logpi"):
        assert_close(n, a[n], b[n], 2e-5)
    for n in a["m"]:
        assert_close(n, a["m"][n], b["m"][n], 5e-4 if n.startswith("m.pi") else 1e-4)
    assert_close("block.params", a["block.params"], b["block.params"], 1e-4)
    for key in ("q1_loss", "q2_loss", "policy_loss", "log_alpha"):
        assert abs(a["metrics"][key] - b["metrics"][key]) <= 1e-4 * abs(b["metrics"][key]) + 1e-6, key


def test_tc_setup_failure_falls_back_to_ffma_plans(monkeypatch):
    """If the tensor-core setup fails after the plans were built for it (head GEMMs, tails, epilogue projections), the engine
    rebuilds plain FFMA plans: same results, bit for bit, as an engine created with SACX_TC=0."""
    obs, act, B = 24, 4, 2048
    rng = np.random.default_rng(3)
    idx = rng.choice(2 * B - 7, B, replace=False).astype(np.int64)
    e1 = rng.standard_normal((B, act)).astype(np.float32)
    e2 = rng.standard_normal((B, act)).astype(np.float32)
    res = {}
    for mode in ("fail", "off"):
        if mode == "fail":
            monkeypatch.setenv("SACX_TC_FAIL", "1")
        else:
            monkeypatch.delenv("SACX_TC_FAIL", raising=False)
        eng = _engine(obs, act, (256, 256), (256, 256), B, "relu", monkeypatch, mode == "fail")
        on, why, _ = eng.tensor_core()
        assert not on and (("forced" in why) if mode == "fail" else ("SACX_TC=0" in why))
        m = eng.update_host(idx, e1, e2, 1)
        assert m["nonfinite"] == 0 and eng.tensor_core()[2] == 0
        res[mode] = {n: eng.view(n).cpu().numpy().copy() for n in ("block.params", "block.targets", "out.y", "out.logpi", "scr.dhead")}
    for n in res["off"]:
        assert np.array_equal(res["fail"][n], res["off"][n]), n
