"""-m gpu: device-side observation assembly of the DonkeyVae producer against the NumPy restatement of
vae_env.py's bookkeeping (oracle/donkey_obs.py), bit for bit, including the transitions that land in the replay ring."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("z,n_cmd,n_hist,n_stack", [(32, 2, 20, 3), (32, 2, 20, 1), (8, 1, 3, 4), (16, 2, 0, 2)])
def test_obs_assembly_matches_the_environment_bookkeeping(z, n_cmd, n_hist, n_stack):
    from oracle.donkey_obs import DonkeyObsOracle
    from sac.donkey_obs import DeviceObservationAssembler
    from sac.replay_buffer import ReplayBuffer
    rng = np.random.default_rng(0)
    ring = ReplayBuffer(64)
    dev = DeviceObservationAssembler(z, n_cmd, n_hist, n_stack, ring=ring)
    ora = DonkeyObsOracle(z, n_cmd, n_hist, n_stack)
    assert dev.obs_dim == n_stack * (z + n_cmd * n_hist)          # 216 for the shipped setup
    want = []
    lat = rng.standard_normal(z).astype(np.float32)
    o_dev = dev.reset(torch.from_numpy(lat).cuda())
    o_ref = np.asarray(ora.reset(lat), np.float32).reshape(-1).copy()
    assert np.array_equal(o_dev.cpu().numpy(), o_ref)
    for t in range(40):
        lat = rng.standard_normal(z).astype(np.float32)
        act = rng.uniform(-1, 1, n_cmd).astype(np.float32)
        rew, done = float(rng.standard_normal()), bool(t % 11 == 10)
        prev = o_ref
        a_in = torch.from_numpy(act).cuda() if t % 2 else act           # device- and host-resident actions
        o_dev = dev.step(torch.from_numpy(lat).cuda(), a_in, rew, done)
        o_ref = np.asarray(ora.step(lat, act, done), np.float32).reshape(-1).copy()
        assert np.array_equal(o_dev.cpu().numpy(), o_ref), t
        want.append((prev, act, np.float32(rew), o_ref, np.float32(done)))
        if done:
            lat = rng.standard_normal(z).astype(np.float32)
            o_dev = dev.reset(torch.from_numpy(lat).cuda())
            o_ref = np.asarray(ora.reset(lat), np.float32).reshape(-1).copy()
            assert np.array_equal(o_dev.cpu().numpy(), o_ref)
    assert len(ring) == 40
    b = ring.sample_tensors(40, indices=np.arange(40))
    for i, (s, a, r, s2, d) in enumerate(want):
        assert np.array_equal(b.state[i].cpu().numpy(), s) and np.array_equal(b.action[i].cpu().numpy(), a)
        assert b.reward[i].item() == r and np.array_equal(b.next_state[i].cpu().numpy(), s2) and b.done[i].item() == d
