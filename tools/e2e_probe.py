"""Where does an end-to-end SAC.training_step() go? Host sections timed with perf_counter over N steps (run on the GPU box):
   python tools/e2e_probe.py [steps]
draw = replay index stream (random.sample semantics), normals = torch CPU normals, submit = C-ABI call (pinned copy + H2D + launch;
blocks when both I/O slots are still in flight, i.e. when the GPU is the bottleneck)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "soft-actor-critic_b200")]
import numpy as np
import bench
import torch
from sac.agent import SAC

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
w = dict(bench.WORKLOADS["bipedal"])
w["fill"] = min(w["fill"], 200_000)
agent = SAC(bench.FakeEnv(w["obs"], w["act"]), bench.make_config(w, "host"))
agent.replay_buffer.push_batch(*bench.synth(w["fill"], w["obs"], w["act"]))
B = agent.config["train"]["batch_size"]
for _ in range(200):
    agent.training_step()
agent.last_metrics()
t = {"draw": 0.0, "normals": 0.0, "submit": 0.0}
pc = time.perf_counter
t0 = pc()
for _ in range(steps):
    a = pc(); idx = np.asarray(agent.replay_buffer.draw_indices(B), dtype=np.int64)
    b = pc(); e1, e2 = agent._normal_pair(B)
    c = pc(); agent.engine.update_host_pipelined(idx, e1, e2, 1)
    d = pc()
    t["draw"] += b - a; t["normals"] += c - b; t["submit"] += d - c
agent._host_pending = True
agent.last_metrics()
tot = pc() - t0
print(f"{steps} steps: {tot / steps * 1e6:.1f} us/step total = {steps / tot:.0f}/s; " + ", ".join(f"{k} {v / steps * 1e6:.1f} us" for k, v in t.items()))
# the same host work without the GPU in the loop
t1 = pc()
for _ in range(steps):
    idx = np.asarray(agent.replay_buffer.draw_indices(B), dtype=np.int64)
    e1, e2 = agent._normal_pair(B)
print(f"host draw + normals alone: {(pc() - t1) / steps * 1e6:.1f} us/step")
# device-only pace for comparison: the same kernel fed from the device RNG, one launch per update
eng = agent.engine
torch.cuda.synchronize(); t2 = pc()
for _ in range(steps):
    eng.update(None, None, None, 1)
eng.sync(); print(f"one launch per update, device RNG, no host inputs: {(pc() - t2) / steps * 1e6:.1f} us/step")
