"""Diagnostic (GPU box): where does a gradient discrepancy against the reference come from? Per-row comparison of the actor
backward's intermediates (head gradient, deltas of the policy's hidden layers) with the NumPy oracle on the same teacher-forced
state: a discontinuity flip (relu' at a pre-activation within rounding of zero, torch.min routing) shows up as ONE outlier row,
a precision problem as a uniform error.   python tools/flip_probe.py <golden> <path>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "soft-actor-critic_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from helpers import Golden, ReferenceRun, numpy_oracle_from_golden, rel_l2
from gpu_helpers import dev, load_nets, set_engine_state, fill_ring
import test_gpu_baseline_configs as T
from oracle.sac_numpy import mlp_from_state_dict, squash_sample

name, path = sys.argv[1], sys.argv[2]
for k, v in T.PATHS[path].items():
    os.environ[k] = v
g = Golden(name)
ref = ReferenceRun(g)
from sac.engine import UpdateEngine
from sac.replay_buffer import ReplayBuffer
cfg = dict(g.cfg); cfg["train"] = dict(cfg["train"], device="cuda")
eng = UpdateEngine(g.obs, g.act, cfg)
rb = ReplayBuffer(g.cfg["buffer"]["capacity"], g.obs, g.act); fill_ring(rb, g.n_fill, g.obs, g.act); eng.attach_ring(rb)
print(name, path, eng.path(), eng.tensor_core())
r = ref.step()
set_engine_state(eng, r["before"])
eng.sample_batch(dev(r["idx"]))
B = g.cfg["train"]["batch_size"]
y = torch.empty(B, device="cuda"); eng.target(dev(r["eps1"]), y)
eng.critic_step(dev(r["y"]))
load_nets(eng, {"q1": r["mid"]["q1"], "q2": r["mid"]["q2"]})
lp = torch.empty(B, device="cuda"); eng.actor_step(dev(r["eps2"]), lp, grads_only=True); eng.sync()
s, a, rew, s2, d = ref.batch(r["idx"])
for dt in (np.float32, np.float64):
    o = numpy_oracle_from_golden(g, dt)
    c = lambda x: x.astype(dt)
    for tag in ("pi", "q1", "q2"):
        src = r["before"]["pi"] if tag == "pi" else r["mid"][tag]
        net = mlp_from_state_dict(src, g.cfg["policy_net" if tag == "pi" else "q_net"]["hidden_layers_act"], "identity", dt)
        for w_dst, w_src in zip(getattr(o, tag).tensors(), net.tensors()): w_dst[...] = w_src
    ag = o.actor_grads(c(s), c(r["eps2"]))
    head, pc = o.pi.forward(c(s))
    L = o.pi.n_layers
    deltas = {}
    delta = ag["d_head"]
    for l in range(L - 1, 0, -1):
        delta = (delta @ o.pi.W[l]) * (pc["h"][l] > 0)
        deltas[l - 1] = delta
    print(f"--- oracle {dt.__name__}")
    def rows(nm, got, want):
        got = got.astype(np.float64); want = want.astype(np.float64)
        e = np.sqrt(((got - want) ** 2).sum(1)); n = np.sqrt((want ** 2).sum(1)) + 1e-30
        tot = np.sqrt((e ** 2).sum() / (n ** 2).sum())
        bad = np.nonzero(e / n.mean() > 1e-3)[0]
        rest = np.sqrt((np.delete(e, bad) ** 2).sum() / (n ** 2).sum())
        print(f"{nm:14s} rel-L2 {tot:.2e}; rows off by > 1e-3 of the mean row norm: {len(bad)} {bad[:8].tolist()}; without them {rest:.2e}")
    rows("scr.dhead", eng.view("scr.dhead").cpu().numpy(), ag["d_head"])
    for l in sorted(deltas):
        rows(f"delta.pi.{l}", eng.view(f"delta.pi.{l}").cpu().numpy()[:, :deltas[l].shape[1]], deltas[l])
    for l in range(L):
        print(f"g.pi.W{l} vs oracle {rel_l2(eng.view(f'g.pi.W{l}').cpu().numpy(), ag['dW'][l]):.2e}  vs reference {rel_l2(eng.view(f'g.pi.W{l}').cpu().numpy(), r['gpi'][f'net.{2*l}.weight']):.2e}")
    # smallest |pre-activation| the oracle sees (flip candidates)
    zs = np.concatenate([np.abs(z).ravel() for z in pc["z"][:-1]])
    print("policy hidden pre-activations: min |z| %.2e, count(|z| < 1e-6) %d of %d" % (zs.min(), (zs < 1e-6).sum(), zs.size))
