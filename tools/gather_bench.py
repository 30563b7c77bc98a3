"""Replay-ring gather alone ("gather-only GB/s"): uniform-index gather of (s, a, r, s', done)
from a 1M-transition BipedalWalker-shape ring (224 MB of packed 224-byte records) through the C ABI (sacx_ring_gather), device indices.
   python tools/gather_bench.py            -> one JSON line per batch size; run on the GPU box."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "soft-actor-critic_b200")]
import bench  # noqa: E402
import torch  # noqa: E402
from sac.replay_buffer import ReplayBuffer  # noqa: E402

O, A, N = 24, 4, 1_000_000
rb = ReplayBuffer(N, O, A)
s, a, r, s2, d = bench.synth(N, O, A)
rb.push_batch(s, a, r, s2, d)
torch.cuda.synchronize()
peak = bench.peaks()["hbm_gbs"]
row_bytes = (2 * O + A + 2) * 4
for B in (256, 4096, 65536, 1_000_000):
    reps = 200 if B <= 65536 else 20
    idx = [torch.randint(0, N, (B,), device="cuda", dtype=torch.int64) for _ in range(4)]     # with replacement: bandwidth test
    for i in range(3):
        rb.sample_tensors(B, idx[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        out = rb.sample_tensors(B, idx[i % 4])
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    # the same gather into PREALLOCATED outputs straight through the C ABI (what the fused update's consumers see; the call above
    # also allocates five fresh torch tensors per call)
    from sac import _engine as E
    kw = dict(dtype=torch.float32, device="cuda")
    so, ao, ro, s2o, do = torch.empty(B, O, **kw), torch.empty(B, A, **kw), torch.empty(B, **kw), torch.empty(B, O, **kw), torch.empty(B, **kw)
    call = lambda ix: E.check(rb._lib.sacx_ring_gather(rb.handle, 0, ix.data_ptr(), B, so.data_ptr(), ao.data_ptr(), ro.data_ptr(), s2o.data_ptr(), do.data_ptr()))
    for i in range(3):
        call(idx[i % 4])
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for i in range(reps):
        call(idx[i % 4])
    k1.record()
    torch.cuda.synchronize()
    us_k = k0.elapsed_time(k1) * 1e3 / reps
    assert np.array_equal(so.cpu().numpy(), s[idx[(reps - 1) % 4].cpu().numpy()])
    # check one batch against the host copy (bit-exact)
    j = idx[(reps - 1) % 4].cpu().numpy()
    assert np.array_equal(out.state.cpu().numpy(), s[j]) and np.array_equal(out.reward.cpu().numpy(), r[j])
    alg = B * row_bytes
    print(json.dumps({"what": "sacx_ring_gather (+ output allocation) per call", "batch": B, "us_per_call": round(us, 2),
                      "algorithmic_read_bytes": alg, "read_GBps": round(alg / us / 1e3, 1),
                      "read_plus_write_GBps": round((2 * alg + 8 * B) / us / 1e3, 1), "hbm_peak_GBps": peak,
                      "frac_of_peak_rw": round((2 * alg + 8 * B) / us / 1e3 / peak, 4),
                      "preallocated_us_per_call": round(us_k, 2), "preallocated_read_plus_write_GBps": round((2 * alg + 8 * B) / us_k / 1e3, 1),
                      "preallocated_frac_of_peak_rw": round((2 * alg + 8 * B) / us_k / 1e3 / peak, 4),
                      "note": "packed records [s | s2 | a | r | d | pad]: 224 B = seven 32-byte sectors per 216 B row"}))
