"""-m gpu tests of the row-parallel cluster kernel (csrc/sacx_rowpar.cuh): the single-agent fast path.

It must agree with the reference's recorded vectors (bipedal / pendulum128 goldens, also exercised by
test_gpu_parity.py::test_fused_update_free_running_vs_reference, which runs this path when eligible) and with the
tile-parallel kernel (SACX_ROWPAR=0) on every intermediate the two paths share. Tolerances: 3xTF32 tensor-core
tiles and a different summation order, not different math -> rel-L2 2e-5 on activations / targets, 1e-4 on
parameters after two free-running steps."""
import numpy as np
import pytest
import torch

from gpu_helpers import assert_close, base_config, dev, engine_from_golden, fill_ring, load_nets, read_net
from helpers import Golden
from test_gpu_parity import _random_nets

pytestmark = pytest.mark.gpu


def _engine(obs, act, hp, hq, B, actfn, monkeypatch, rowpar, out_act="identity", auto=True, cap=2000, fill=1500):
    from sac.engine import UpdateEngine
    from sac.replay_buffer import ReplayBuffer
    monkeypatch.setenv("SACX_ROWPAR", "1" if rowpar else "0")
    cfg = base_config(hidden=hp, q_hidden=hq, act=actfn, out_act=out_act, batch=B, capacity=cap, auto=auto)
    eng = UpdateEngine(obs, act, cfg)
    load_nets(eng, _random_nets(obs, act, hp, hq, scale=0.15))
    eng.reset_state()
    rb = ReplayBuffer(cap, obs, act)
    fill_ring(rb, fill, obs, act)
    eng.attach_ring(rb)
    return eng


def test_rowpar_is_the_default_single_agent_path(monkeypatch):
    g = Golden("bipedal")
    monkeypatch.delenv("SACX_ROWPAR", raising=False)
    eng = engine_from_golden(g)
    assert eng.path()[0] == "rowpar"
    gx, gy, smem = eng.grid()
    assert gx >= 8 and gy == 1 and smem <= 227 * 1024        # 8 CTAs per row group (+ CTAs that only serve the dW phases)
    tiny = engine_from_golden(Golden("tiny_auto"))
    kind, why = tiny.path()
    assert kind == "tiles" and why                                # narrow nets keep the generic tile-parallel kernel
    monkeypatch.setenv("SACX_ROWPAR", "0")
    assert engine_from_golden(g).path() == ("tiles", "disabled by SACX_ROWPAR=0")


CASES = [
    # obs, act, hidden_pi, hidden_q, B, activation
    (24, 4, (256, 256), (256, 256), 256, "relu"),        # BASELINE config 2
    (4, 1, (128, 128), (128, 128), 256, "relu"),         # pendulum notebook config: K0 = 4 / 5 (unaligned first layers)
    (17, 6, (64, 128), (128, 64), 100, "tanh"),          # ragged batch (100 = 6 row blocks + 4 rows), mixed widths
    (32, 2, (256, 256), (256, 256), 1024, "relu"),       # Donkey latent shape: 64 row blocks, 4 per cluster
    (11, 3, (128, 64, 256), (64, 256, 128), 48, "leaky_relu"),   # three hidden layers
    (8, 8, (64, 64), (64, 64), 16, "identity"),          # one row block, maximum action dimension of this path
]


@pytest.mark.parametrize("obs,act,hp,hq,B,actfn", CASES)
def test_rowpar_matches_tile_parallel_path(obs, act, hp, hq, B, actfn, monkeypatch):
    rng = np.random.default_rng(11)
    K = 2
    idx = np.stack([rng.choice(1500, B, replace=False) for _ in range(K)]).astype(np.int64)
    e1 = rng.standard_normal((K, B, act)).astype(np.float32)
    e2 = rng.standard_normal((K, B, act)).astype(np.float32)
    out = {}
    for rowpar in (True, False):
        eng = _engine(obs, act, hp, hq, B, actfn, monkeypatch, rowpar)
        assert eng.path()[0] == ("rowpar" if rowpar else "tiles"), eng.path()
        snaps = []
        for k in range(K):
            m = eng.update_host(idx[k], e1[k], e2[k], 1)
            assert m["nonfinite"] == 0 and m["updates"] == k + 1
            snap = {n: eng.view(n).cpu().numpy().copy() for n in
                    ("batch.sa", "batch.s2a", "batch.spi", "batch.r", "batch.d", "out.y", "out.logpi", "out.logpi_next",
                     "out.q1", "out.q2", "out.tq1", "out.tq2", "out.q1_pi", "out.q2_pi", "scr.dhead", "scr.dout1",
                     "delta.q1.0", "delta.pi.0", "act.pia.h0", "block.params", "block.targets", "block.m", "block.v")}
            snap["metrics"] = m
            snaps.append(snap)
        out[rowpar] = snaps
    for k in range(K):
        a, b = out[True][k], out[False][k]
        for n in ("batch.sa", "batch.r", "batch.d"):
            assert np.array_equal(a[n], b[n]), n                     # the gather is bit-exact
        tol = 2e-5 * (3 ** k)
        for n in ("batch.s2a", "batch.spi", "out.y", "out.logpi", "out.logpi_next", "out.q1", "out.q2", "out.tq1", "out.tq2",
                  "out.q1_pi", "out.q2_pi", "act.pia.h0"):
            assert_close(f"step{k} {n}", a[n], b[n], tol)
        for n in ("scr.dhead", "scr.dout1", "delta.q1.0", "delta.pi.0"):
            assert_close(f"step{k} {n}", a[n], b[n], 5 * tol)
        for n in ("block.params", "block.targets"):
            assert_close(f"step{k} {n}", a[n], b[n], 1e-4 * (2 ** k))
        for key in ("q1_loss", "q2_loss", "policy_loss", "alpha_loss", "log_alpha"):
            assert abs(a["metrics"][key] - b["metrics"][key]) <= 1e-4 * abs(b["metrics"][key]) + 1e-6, key


def test_rowpar_device_rng_and_multi_step_launch(monkeypatch):
    """Device RNG mode draws the same indices / normals as the tile-parallel kernel (same keys), and n updates inside one
    launch equal n launches of one update (bit for bit: same kernel, same order)."""
    res = {}
    for rowpar in (True, False):
        eng = _engine(24, 4, (256, 256), (256, 256), 256, "relu", monkeypatch, rowpar)
        eng.update(None, None, None, 1)
        eng.sync()
        res[rowpar] = {n: eng.view(n).cpu().numpy().copy() for n in ("batch.idx", "batch.eps1", "batch.eps2", "out.y")}
    assert np.array_equal(res[True]["batch.idx"], res[False]["batch.idx"])
    assert np.array_equal(res[True]["batch.eps1"], res[False]["batch.eps1"])
    assert np.array_equal(res[True]["batch.eps2"], res[False]["batch.eps2"])
    assert_close("y", res[True]["out.y"], res[False]["out.y"], 2e-5)
    a = _engine(24, 4, (256, 256), (256, 256), 256, "relu", monkeypatch, True)
    b = _engine(24, 4, (256, 256), (256, 256), 256, "relu", monkeypatch, True)
    a.update(None, None, None, 5)
    for _ in range(5):
        b.update(None, None, None, 1)
    a.sync(); b.sync()
    assert np.array_equal(a.view("block.params").cpu().numpy(), b.view("block.params").cpu().numpy())
    assert a.metrics()["updates"] == 5 and b.metrics()["updates"] == 5


def test_rowpar_teacher_forced_bipedal_vs_reference(monkeypatch):
    """One fused update from the reference's recorded pre-update state, for every recorded step: y / logpi / losses and
    the post-update parameters against the reference's own values (tolerances of SURVEY section 8c)."""
    from sac.replay_buffer import ReplayBuffer
    monkeypatch.delenv("SACX_ROWPAR", raising=False)
    g = Golden("bipedal")
    eng = engine_from_golden(g)
    assert eng.path()[0] == "rowpar"
    rb = ReplayBuffer(g.cfg["buffer"]["capacity"], g.obs, g.act)
    fill_ring(rb, g.n_fill, g.obs, g.act)
    eng.attach_ring(rb)
    m = eng.update_host(g["step0/idx"], g["step0/eps1"], g["step0/eps2"], 1)
    assert_close("y", eng.view("out.y").cpu().numpy().ravel(), g["step0/y"], 2e-5)
    assert_close("logpi", eng.view("out.logpi").cpu().numpy().ravel(), g["step0/lp"], 2e-5)
    assert_close("q1", eng.view("out.q1").cpu().numpy().ravel(), g["step0/q1"], 2e-5)
    assert abs(m["q1_loss"] - float(g["step0/q1_loss"])) <= 2e-5 * abs(float(g["step0/q1_loss"])) + 1e-7
    assert abs(m["log_alpha"] - float(g["step0/log_alpha"])) < 1e-6


@pytest.mark.parametrize("obs,act,hp,hq,B,actfn", [CASES[0], CASES[2], CASES[4]])
def test_rowpar_weight_gradient_tiles_tma_vs_ffma(obs, act, hp, hq, B, actfn, monkeypatch):
    """The weight-gradient tiles of the two tile-parallel phases have two implementations: 3xTF32 tensor-core tiles with
    TMA-staged operands (default) and the FFMA / cp.async tile the kernel falls back to when a tensor map cannot be encoded
    (SACX_RP_DW_TMA=0 forces it). Same gradients, same Adam / Polyak arithmetic: parameters, moments and targets agree to
    summation-order tolerance after two updates, everything in front of the dW phases bit for bit."""
    rng = np.random.default_rng(5)
    K = 2
    idx = np.stack([rng.choice(1500, B, replace=False) for _ in range(K)]).astype(np.int64)
    e1 = rng.standard_normal((K, B, act)).astype(np.float32)
    e2 = rng.standard_normal((K, B, act)).astype(np.float32)
    out = {}
    for tma in ("1", "0"):
        monkeypatch.setenv("SACX_RP_DW_TMA", tma)
        eng = _engine(obs, act, hp, hq, B, actfn, monkeypatch, True)
        assert eng.path()[0] == "rowpar"
        snaps = []
        for k in range(K):
            m = eng.update_host(idx[k], e1[k], e2[k], 1)
            assert m["nonfinite"] == 0
            snaps.append({n: eng.view(n).cpu().numpy().copy() for n in ("out.y", "out.q1", "out.tq1", "block.params", "block.targets",
                                                                         "block.m", "block.v")})
        out[tma] = snaps
    monkeypatch.delenv("SACX_RP_DW_TMA")
    for n in ("out.y", "out.q1", "out.tq1"):
        assert np.array_equal(out["1"][0][n], out["0"][0][n]), n            # first update, phase A: in front of any dW tile
    for k in range(K):
        for n in ("block.params", "block.targets"):
            assert_close(f"step{k} {n}", out["1"][k][n], out["0"][k][n], 2e-6 * (4 ** k))
        assert_close(f"step{k} block.m", out["1"][k]["block.m"], out["0"][k]["block.m"], 2e-5 * (4 ** k))


def test_rowpar_soak_is_deterministic_and_finite(monkeypatch):
    """2 000 free-running updates (device RNG) in two differently chunked launch sequences: same parameters bit for bit --
    the barrier counters, the alternating barrier sets, the TMA barriers' phase parities and the dW phases of a grid wider than
    the row groups all have to survive thousands of updates -- and nothing non-finite."""
    a = _engine(24, 4, (256, 256), (256, 256), 256, "relu", monkeypatch, True, cap=20000, fill=15000)
    b = _engine(24, 4, (256, 256), (256, 256), 256, "relu", monkeypatch, True, cap=20000, fill=15000)
    assert a.path()[0] == "rowpar"
    a.update(None, None, None, 2000)
    for n in (1, 7, 500, 1492):
        b.update(None, None, None, n)
    a.sync(); b.sync()
    pa, pb = a.view("block.params").cpu().numpy(), b.view("block.params").cpu().numpy()
    assert np.isfinite(pa).all() and np.array_equal(pa, pb)
    assert np.array_equal(a.view("block.targets").cpu().numpy(), b.view("block.targets").cpu().numpy())
    ma, mb = a.metrics(), b.metrics()
    assert ma["updates"] == 2000 and mb["updates"] == 2000 and ma["nonfinite"] == 0
    assert ma["log_alpha"] == mb["log_alpha"] and np.isfinite(ma["q1_loss"])
