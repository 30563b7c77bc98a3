"""Error table of the tensor-core path against the FFMA path (run on the GPU box): python tools/tc_debug.py [B]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "soft-actor-critic_b200"), os.path.join(ROOT, "tests")]
import torch
from gpu_helpers import base_config, fill_ring, load_nets
from helpers import rel_l2
from test_gpu_parity import _random_nets
from sac.engine import UpdateEngine
from sac.replay_buffer import ReplayBuffer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
CASES = {0: (24, 4, (256, 256), (256, 256), "relu"), 1: (11, 3, (128, 48, 256), (80, 256, 128), "leaky_relu"),
         2: (17, 6, (64, 128), (128, 64), "tanh")}
obs, act, hp, hq, actfn = CASES[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
os.environ["SACX_ROWPAR"] = "0"
rng = np.random.default_rng(5)
K = 2
cap = 2 * B
idx = np.stack([rng.choice(cap - 7, B, replace=False) for _ in range(K)]).astype(np.int64)
e1 = rng.standard_normal((K, B, act)).astype(np.float32)
e2 = rng.standard_normal((K, B, act)).astype(np.float32)
out = {}
for tc in (1, 0, 2):
    os.environ["SACX_TC"] = "0" if tc == 0 else "1"
    os.environ["SACX_TC_MIN_BATCH"] = "1024"
    os.environ["SACX_TILE"] = "small" if tc == 2 else "large"
    if tc == 2:
        os.environ["SACX_TC"] = "0"
    eng = UpdateEngine(obs, act, base_config(hidden=hp, q_hidden=hq, act=actfn, batch=B, capacity=cap))
    load_nets(eng, _random_nets(obs, act, hp, hq, scale=0.15))
    eng.reset_state()
    rb = ReplayBuffer(cap, obs, act)
    fill_ring(rb, cap - 7, obs, act)
    eng.attach_ring(rb)
    snaps = []
    for k in range(K):
        eng.update_host(idx[k], e1[k], e2[k], 1)
        snaps.append({n: eng.view(n).cpu().numpy().copy() for n in eng.layout if not n.startswith("scal") and n != "batch.idx"})
    out[tc] = snaps
    print("engine", tc, eng.tensor_core(), eng.path())
for k in range(K):
    print("step", k, " name: tc-vs-ffma(large)   ffma(small)-vs-ffma(large)")
    for n in sorted(out[1][k]):
        a, b, c = out[1][k][n], out[0][k][n], out[2][k][n]
        if b.size and np.linalg.norm(b) > 0:
            if rel_l2(a, b) > 1e-6:
                d = np.abs(a.astype(np.float64) - b)
                print(f"  {n:24s} {rel_l2(a, b):.3e}   {rel_l2(c, b):.3e}   max|d| {d.max():.3e} at {np.unravel_index(d.argmax(), d.shape)} n>1e-4: {(d > 1e-4).sum()}")
