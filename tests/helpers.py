"""Shared test helpers: golden loading, synthetic transitions, oracle construction."""
from __future__ import annotations

import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def synth_transitions(n, obs, act, seed=0):
    """s, s' ~ N(0,1); a ~ U(-1,1); r ~ N(0,1); d ~ Bernoulli(0.01); numpy default_rng(seed)."""
    rng = np.random.default_rng(seed)
    s = rng.standard_normal((n, obs)).astype(np.float32)
    s2 = rng.standard_normal((n, obs)).astype(np.float32)
    a = rng.uniform(-1, 1, (n, act)).astype(np.float32)
    r = rng.standard_normal(n).astype(np.float32)
    d = rng.random(n) < 0.01
    return s, a, r, s2, d


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
        self.meta = json.loads(bytes(self.z["meta"]).decode())
        self.cfg = self.meta["config"]
        self.obs, self.act, self.K = self.meta["obs"], self.meta["act"], self.meta["K"]
        self.n_fill = self.meta["n_fill"]
        self.full = self.meta["full_state"]
        self.big = bool(self.meta.get("big", False))
        self.stride = int(self.meta.get("stride", 1))
        self.store_init = bool(self.meta.get("store_init", True))
        self.start_ckpt = self.meta.get("start_ckpt")
        self.name = name
        self._init = None
        self._streams = None

    def has(self, k):
        return k in self.z.files

    def ckpt_path(self):
        return os.path.join(GOLDEN, self.start_ckpt) if self.start_ckpt else None

    def init_sd(self, tag):
        """Initial weights of one network: stored in the file (round-1 goldens), rebuilt from train.seed with the
        reference's init recipe (F10: nn.Linear default init consumes the stream, then xavier_uniform_ / zeros; policy
        and Q1 with `seed`, Q2 with `seed + 1`) and verified against the stored checksums + samples, or read from the
        reference-written checkpoint the run started from."""
        if self.store_init and not self.start_ckpt:
            return self.sd(f"init/{tag}")
        if self._init is None:
            import torch
            if self.start_ckpt:
                ck = torch.load(self.ckpt_path(), map_location="cpu", weights_only=False)
                self._init = {t: {k: v.detach().numpy().copy() for k, v in ck[key].items()} for t, key in
                              (("pi", "policy_net_state_dict"), ("q1", "q_net1_state_dict"), ("q2", "q_net2_state_dict"),
                               ("q1t", "q_net1_target_state_dict"), ("q2t", "q_net2_target_state_dict"))}
            else:
                c, seed = self.cfg, self.cfg["train"]["seed"]
                self._init = {}
                rng_state = torch.get_rng_state()             # nn.Linear's init draws from the global generator: leave it untouched
                for t, sizes, sd in (("pi", [self.obs] + list(c["policy_net"]["hidden_sizes"]) + [2 * self.act], seed),
                                     ("q1", [self.obs + self.act] + list(c["q_net"]["hidden_sizes"]) + [1], seed),
                                     ("q2", [self.obs + self.act] + list(c["q_net"]["hidden_sizes"]) + [1], seed + 1)):
                    torch.manual_seed(sd)
                    lins = [torch.nn.Linear(sizes[i], sizes[i + 1]) for i in range(len(sizes) - 1)]
                    out = {}
                    for l, lin in enumerate(lins):
                        torch.nn.init.xavier_uniform_(lin.weight)
                        torch.nn.init.zeros_(lin.bias)
                        out[f"net.{2 * l}.weight"] = lin.weight.detach().numpy().copy()
                        out[f"net.{2 * l}.bias"] = lin.bias.detach().numpy().copy()
                    for k, v in out.items():          # the rebuilt init IS the reference's: bit-equal samples and checksums
                        v64 = v.astype(np.float64).ravel()
                        assert np.array_equal(np.array([v64.sum(), (v64 * v64).sum()]), self.z[f"init/{t}/{k}#chk"]), (t, k)
                        assert np.array_equal(v.ravel()[::97], self.z[f"init/{t}/{k}#smp"]), (t, k)
                    self._init[t] = out
                torch.set_rng_state(rng_state)
        return self._init[tag]

    def streams(self, k):
        """(idx, eps1, eps2) of update k. Stored in the file, or -- `big` goldens -- regenerated exactly as the reference
        consumed them: Python's `random` and torch's CPU generator seeded by SAC._set_seed (agent.py:117-124), one
        random.sample(range(n), B) and two empty(B, A).normal_() per update; verified against the stored checksums."""
        if not self.big:
            return self.z[f"step{k}/idx"], self.z[f"step{k}/eps1"], self.z[f"step{k}/eps2"]
        if self._streams is None:
            import random
            import torch
            seed, B = self.cfg["train"]["seed"], self.cfg["train"]["batch_size"]
            n = min(self.n_fill, self.cfg["buffer"]["capacity"])
            rnd = random.Random(seed)
            gen = torch.Generator().manual_seed(seed)
            self._streams = []
            for kk in range(self.K):
                idx = np.asarray(rnd.sample(range(n), B), dtype=np.int64)
                e1 = torch.empty(B, self.act).normal_(generator=gen).numpy()
                e2 = torch.empty(B, self.act).normal_(generator=gen).numpy()
                for nm, v in (("idx", idx), ("eps1", e1), ("eps2", e2)):
                    v64 = v.astype(np.float64).ravel()
                    assert np.array_equal(np.array([v64.sum(), (v64 * v64).sum()]), self.z[f"step{kk}/{nm}#chk"]), nm
                self._streams.append((idx, e1, e2))
        return self._streams[k]

    def rows(self, k, key):
        """Per-row output of update k as (reference values, selector): full vector, or the strided sample of a `big` golden."""
        if self.big:
            return self.z[f"step{k}/{key}#smp"], slice(None, None, self.stride)
        return self.z[f"step{k}/{key}"], slice(None)

    def __getitem__(self, k):
        return self.z[k]

    def sd(self, prefix):
        p = prefix + "/"
        return {k[len(p):]: self.z[k] for k in self.z.files if k.startswith(p) and "#" not in k}

    def keys(self, prefix):
        p = prefix + "/"
        return [k for k in self.z.files if k.startswith(p)]


class VirtualGolden:
    """A configuration no golden file was recorded for (ragged batches, odd widths, deep nets): same interface as Golden
    with K = 0 recorded updates. ReferenceRun then IS the reference for it: oracle/torch_port.py executes the reference's
    torch op sequence (pinned bit-exact to /root/reference on every recorded configuration by test_oracle_golden.py)."""
    full, big, stride, store_init, start_ckpt, K = True, False, 1, True, None, 0

    def __init__(self, name, obs, act, cfg, n_fill):
        self.name, self.obs, self.act, self.cfg, self.n_fill = name, obs, act, cfg, n_fill
        self.meta = {"obs": obs, "act": act, "n_fill": n_fill, "K": 0, "config": cfg}
        self._init = None

    def has(self, k):
        return False

    def ckpt_path(self):
        return None

    def init_sd(self, tag):
        if self._init is None:
            import torch
            from oracle.torch_port import _init_net
            rng_state = torch.get_rng_state()
            c, seed = self.cfg, self.cfg["train"]["seed"]
            self._init = {}
            for t, sizes, sd in (("pi", [self.obs] + list(c["policy_net"]["hidden_sizes"]) + [2 * self.act], seed),
                                 ("q1", [self.obs + self.act] + list(c["q_net"]["hidden_sizes"]) + [1], seed),
                                 ("q2", [self.obs + self.act] + list(c["q_net"]["hidden_sizes"]) + [1], seed + 1)):
                ps = _init_net(sizes, sd)
                self._init[t] = {f"net.{2 * (i // 2)}.{'weight' if i % 2 == 0 else 'bias'}": p.detach().numpy().copy() for i, p in enumerate(ps)}
            torch.set_rng_state(rng_state)
        return self._init[tag]


def numpy_oracle_from_golden(g, dtype=np.float32):
    from oracle.sac_numpy import SACOracle, Hyper, mlp_from_state_dict

    c = g.cfg
    hp = Hyper(gamma=c["sac"]["gamma"], tau=c["sac"]["tau"], alpha=c["sac"]["alpha"],
               auto_entropy_tuning=c["sac"]["auto_entropy_tuning"], actor_lr=c["sac"]["actor_lr"],
               critic_lr=c["sac"]["critic_lr"], alpha_lr=c["sac"]["alpha_lr"],
               log_std_min=c["policy_net"]["log_std_min"], log_std_max=c["policy_net"]["log_std_max"],
               action_scale=c["policy_net"]["action_scale"])
    pi = mlp_from_state_dict(g.init_sd("pi"), c["policy_net"]["hidden_layers_act"], c["policy_net"]["output_activation"], dtype)
    q1 = mlp_from_state_dict(g.init_sd("q1"), c["q_net"]["hidden_layers_act"], c["q_net"]["output_activation"], dtype)
    q2 = mlp_from_state_dict(g.init_sd("q2"), c["q_net"]["hidden_layers_act"], c["q_net"]["output_activation"], dtype)
    return SACOracle(pi, q1, q2, hp, dtype=dtype)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def max_rel(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / den) if den > 0 else float(np.abs(a - b).max())


def chk64(v):
    v64 = np.asarray(v, dtype=np.float64).ravel()
    return np.array([v64.sum(), (v64 * v64).sum()])


def tensor_err(g: Golden, key, got):
    """rel-L2 error of a weight-shaped tensor against the golden entry `key`: the full tensor when the file has it, else the
    stride-97 sample; plus the relative error of the sum of squares against the stored float64 checksum (None when full)."""
    got = np.asarray(got)
    if g.has(key):
        return rel_l2(got, g[key]), None
    e = rel_l2(got.ravel()[::97], g[key + "#smp"])
    ref2 = float(g[key + "#chk"][1])
    e2 = abs(float(chk64(got)[1]) - ref2) / ref2 if ref2 > 0 else abs(float(chk64(got)[1]))
    return e, e2


# ---------------------------------------------------------------------------------------------------------------------
# The reference, step by step, with FULL tensors: oracle/torch_port.py replays a golden run bit for bit (proved against the
# golden's checksums / samples by tests/test_oracle_golden.py in the CPU suite, and re-checked here at every step), which gives
# the GPU tests the complete reference state before / inside / after every update -- what teacher forcing needs at shapes
# whose golden file only holds checksums and strided samples.
class ReferenceRun:
    NETS = ("pi", "q1", "q2", "q1t", "q2t")

    def __init__(self, g: Golden, fill=True, strict=False):
        """strict=True (the CPU suite, same machine as the recording): bit equality with the golden file. strict=False (the
        GPU box: another CPU, another GEMM blocking in torch's backward): the replay must sit within 1e-5 rel-L2 of every
        recorded value -- it is then used as the full-tensor stand-in for the reference, next to direct comparisons of the
        CUDA results with the golden file's own samples."""
        import torch
        from oracle.torch_port import TorchPortSAC

        self.g = g
        self.strict = strict
        self.port = TorchPortSAC(g.obs, g.act, g.cfg, capacity=g.cfg["buffer"]["capacity"])
        if g.start_ckpt:
            self.port.load_checkpoint(torch.load(g.ckpt_path(), map_location="cpu", weights_only=False))
        self.data = synth_transitions(g.n_fill, g.obs, g.act)
        if fill:
            s, a, r, s2, d = self.data
            rl, dl = r.tolist(), d.tolist()
            for i in range(g.n_fill):
                self.port.push(s[i], a[i], rl[i], s2[i], dl[i])
        self.k = 0
        self.port.hook_after_critics = self._mid
        self.mid = None
        # the reference consumes the GLOBAL generators (Python `random`, torch CPU); keep private copies of their states so
        # that other code seeding them between steps (network constructors, ...) cannot disturb the replay
        import random
        self._rng = (random.getstate(), torch.get_rng_state())

    # ---- snapshots (numpy copies) ---------------------------------------------------------------------------------
    @staticmethod
    def _sd(ps):
        out = {}
        for l in range(len(ps) // 2):
            out[f"net.{2 * l}.weight"] = ps[2 * l].detach().numpy().copy()
            out[f"net.{2 * l}.bias"] = ps[2 * l + 1].detach().numpy().copy()
        return out

    @staticmethod
    def _grads(ps):
        out = {}
        for l in range(len(ps) // 2):
            out[f"net.{2 * l}.weight"] = ps[2 * l].grad.detach().numpy().copy()
            out[f"net.{2 * l}.bias"] = ps[2 * l + 1].grad.detach().numpy().copy()
        return out

    @staticmethod
    def _adam(opt, ps):
        out = {}
        for l in range(len(ps) // 2):
            for kind, p in (("weight", ps[2 * l]), ("bias", ps[2 * l + 1])):
                st = opt.state.get(p, None)
                if st:
                    out[f"net.{2 * l}.{kind}"] = (st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy(), float(st["step"]))
                else:
                    z = np.zeros(tuple(p.shape), np.float32)
                    out[f"net.{2 * l}.{kind}"] = (z, z.copy(), 0.0)
        return out

    def state(self):
        p = self.port
        st = {t: self._sd(getattr(p, t)) for t in self.NETS}
        st["adam"] = {"pi": self._adam(p.opt_pi, p.pi), "q1": self._adam(p.opt_q1, p.q1), "q2": self._adam(p.opt_q2, p.q2)}
        if p.auto:
            st["log_alpha"] = float(p.log_alpha.detach().reshape(-1)[0])
            a = p.opt_alpha.state.get(p.opt_alpha.param_groups[0]["params"][0], None)
            st["adam_alpha"] = (float(a["exp_avg"]), float(a["exp_avg_sq"]), float(a["step"])) if a else (0.0, 0.0, 0.0)
        st["alpha"] = float(p.alpha.detach().reshape(-1)[0])
        return st

    def _mid(self, port):
        self.mid = {"q1": self._sd(port.q1), "q2": self._sd(port.q2), "gq1": self._grads(port.q1), "gq2": self._grads(port.q2)}

    def step(self):
        """One reference update on the reference's own RNG streams. Returns a dict with the inputs it consumed, the state
        before, the critics after their step (`mid`), every output and the state after; asserts bit-equality with the golden
        file wherever the file has the value."""
        import random
        import torch
        g, p, k = self.g, self.port, self.k
        B = g.cfg["train"]["batch_size"]
        outer = (random.getstate(), torch.get_rng_state())
        random.setstate(self._rng[0])
        torch.set_rng_state(self._rng[1])
        before = self.state()
        st = random.getstate()
        idx = np.asarray(random.sample(range(len(p.memory)), B), dtype=np.int64)
        random.setstate(st)
        ts = torch.get_rng_state()
        e1 = torch.empty(B, g.act).normal_().numpy()
        e2 = torch.empty(B, g.act).normal_().numpy()
        torch.set_rng_state(ts)
        info = p.training_step()
        self._rng = (random.getstate(), torch.get_rng_state())
        random.setstate(outer[0])
        torch.set_rng_state(outer[1])
        out = {"idx": idx, "eps1": e1, "eps2": e2, "before": before, "mid": self.mid, "after": self.state(), "info": info,
               "y": p.last["y"].numpy().copy(), "lp": p.last["lp"].numpy().copy(), "q1": p.last["q1"].numpy().copy(),
               "q2": p.last["q2"].numpy().copy(), "q1_loss": float(p.last["q1_loss"]), "q2_loss": float(p.last["q2_loss"]),
               "policy_loss": float(p.last["policy_loss"]), "gpi": self._grads(p.pi)}
        if k < g.K:                                     # inside the recorded range: the port IS the reference
            gi, ge1, ge2 = g.streams(k)
            assert np.array_equal(idx, gi) and np.array_equal(e1, ge1) and np.array_equal(e2, ge2)      # streams: always bits
            for key in ("y", "lp"):
                ref, sel = g.rows(k, key)
                if self.strict:
                    assert np.array_equal(out[key][sel], ref), key
                else:
                    assert rel_l2(out[key][sel], ref) < 1e-5, key
            if self.strict:
                assert out["q1_loss"] == float(g[f"step{k}/q1_loss"])
            for tag in ("pi", "q1", "q2", "q1t", "q2t"):
                for nm, v in out["after"][tag].items():
                    key = f"step{k}/{tag}/{nm}"
                    if self.strict:
                        assert np.array_equal(v, g[key]) if g.has(key) else np.array_equal(chk64(v), g[key + "#chk"]), key
                    else:
                        e, e2 = tensor_err(g, key, v)
                        assert e < 1e-5 and (e2 is None or e2 < 1e-5), (key, e, e2)
        self.k += 1
        return out

    def batch(self, idx):
        s, a, r, s2, d = self.data
        # logical position j of the deque = push number (pushes - len + j)
        n, cap = self.g.n_fill, self.g.cfg["buffer"]["capacity"]
        rows = idx + max(n - cap, 0)
        return s[rows], a[rows], r[rows], s2[rows], d[rows].astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------------
# Discontinuity-aware comparison. relu' (and torch.min routing) make SAC's backward discontinuous: a hidden unit whose
# pre-activation sits within rounding of zero gets derivative 0 in one correct fp32 implementation and 1 in another, and that
# ONE (row, unit) moves a whole layer gradient by ~1e-3 rel-L2 (measured: tools/flip_probe.py, profiles/r02_flip_probe.txt --
# the entire discrepancy of a 2048-row gradient sat in one row; without it 3e-6). The oracle knows which rows those are: the
# ones where it sees a hidden pre-activation below `thr`. Gradients are therefore compared EXACTLY on all other rows: the
# contribution of the flagged rows is subtracted on both sides (from the engine's per-row deltas / from the oracle's).
def oracle_forward_backward(mlp, x, d_out):
    """NumPy oracle MLP: per-layer inputs h_l (h_0 = x), pre-activations z_l and deltas (gradient w.r.t. z_l), l = 0..L-1."""
    from oracle.sac_numpy import act_bwd
    out, cache = mlp.forward(x)
    L = mlp.n_layers
    deltas = [None] * L
    delta = d_out * act_bwd(mlp.out_act, cache["z"][L - 1], cache["h"][L])
    for l in range(L - 1, -1, -1):
        deltas[l] = delta
        if l > 0:
            delta = (delta @ mlp.W[l]) * act_bwd(mlp.hidden_act, cache["z"][l - 1], cache["h"][l])
    dx = deltas[0] @ mlp.W[0]
    return out, cache, deltas, dx


def ambiguous_rows(mlp, cache, thr=2e-5):
    """Rows with a hidden pre-activation within `thr` of the kink of a piecewise-linear activation (none for smooth ones)."""
    B = cache["z"][0].shape[0]
    amb = np.zeros(B, dtype=bool)
    if mlp.hidden_act in ("relu", "leaky_relu"):
        for z in cache["z"][:-1]:
            amb |= (np.abs(z) < thr * max(1.0, float(np.sqrt(np.mean(z.astype(np.float64) ** 2))))).any(axis=1)
    return amb


def grads_without_rows(dW, db, deltas, inputs, rows):
    """(dW_l, db_l) minus the contribution of `rows` (boolean mask): dW_l -= delta_l[rows]^T inputs_l[rows]."""
    outW, outb = [], []
    for l in range(len(dW)):
        d = np.asarray(deltas[l], dtype=np.float64)[rows]
        h = np.asarray(inputs[l], dtype=np.float64)[rows]
        outW.append(np.asarray(dW[l], dtype=np.float64) - d.T @ h)
        outb.append(np.asarray(db[l], dtype=np.float64).ravel() - d.sum(axis=0))
    return outW, outb
