"""Debug aid: run the row-parallel and the tile-parallel kernels on the same inputs and print rel-L2 of every shared
intermediate.  python tools/rp_debug.py [obs act B h1,h2 act]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "soft-actor-critic_b200"), os.path.join(ROOT, "tests")]
import torch
from gpu_helpers import base_config, fill_ring, load_nets
from helpers import rel_l2
from test_gpu_parity import _random_nets

obs = int(sys.argv[1]) if len(sys.argv) > 1 else 24
act = int(sys.argv[2]) if len(sys.argv) > 2 else 4
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
hid = tuple(int(x) for x in sys.argv[4].split(",")) if len(sys.argv) > 4 else (256, 256)
actfn = sys.argv[5] if len(sys.argv) > 5 else "relu"
K = 2
rng = np.random.default_rng(11)
idx = np.stack([rng.choice(1500, B, replace=False) for _ in range(K)]).astype(np.int64)
e1 = rng.standard_normal((K, B, act)).astype(np.float32)
e2 = rng.standard_normal((K, B, act)).astype(np.float32)
snaps = {}
for rowpar in (1, 0):
    os.environ["SACX_ROWPAR"] = str(rowpar)
    from sac.engine import UpdateEngine
    from sac.replay_buffer import ReplayBuffer
    eng = UpdateEngine(obs, act, base_config(hidden=hid, act=actfn, batch=B, capacity=2000))
    print("path", eng.path(), "grid", eng.grid())
    load_nets(eng, _random_nets(obs, act, hid, hid, scale=0.15))
    eng.reset_state()
    rb = ReplayBuffer(2000, obs, act)
    fill_ring(rb, 1500, obs, act)
    eng.attach_ring(rb)
    out = []
    for k in range(K):
        m = eng.update_host(idx[k], e1[k], e2[k], 1)
        d = {n: eng.view(n).cpu().numpy().copy() for n in eng.layout if not n.startswith(("scal", "part", "block.g", "g."))}
        d["_m"] = m
        out.append(d)
    snaps[rowpar] = out
for k in range(K):
    print(f"--- step {k}")
    a, b = snaps[1][k], snaps[0][k]
    for n in a:
        if n == "_m":
            print("metrics rp   ", {q: round(float(v), 6) for q, v in a[n].items()})
            print("metrics tiles", {q: round(float(v), 6) for q, v in b[n].items()})
            continue
        if a[n].dtype.kind != "f":
            print(f"{n:24s} equal={np.array_equal(a[n], b[n])}")
            continue
        e = rel_l2(a[n], b[n])
        flag = "" if e < 1e-4 else "   <<<<<<"
        if n.startswith(("m.", "v.")) and e < 1e-4:
            continue
        print(f"{n:24s} {e:.3e}{flag}")
