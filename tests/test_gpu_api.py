"""-m gpu tests of the drop-in surface: sac.agent.SAC / sac.replay_buffer.ReplayBuffer used exactly as the
reference's callers (main.py, notebooks) use them, compared with vectors recorded from the reference."""
import json
import os
import random

import numpy as np
import pytest
import torch

from gpu_helpers import FakeEnv, assert_close, base_config
from helpers import Golden, rel_l2, synth_transitions

pytestmark = pytest.mark.gpu


def _agent_from_golden(g, rng="host"):
    from sac.agent import SAC
    cfg = json.loads(json.dumps(g.cfg))
    cfg["train"]["device"] = "cuda"
    cfg["train"]["rng"] = rng
    agent = SAC(FakeEnv(g.obs, g.act), cfg)
    s, a, r, s2, d = synth_transitions(g.n_fill, g.obs, g.act)
    for i in range(g.n_fill):
        agent.store_transition(s[i], a[i], float(r[i]), s2[i], bool(d[i]))
    return agent


@pytest.mark.parametrize("name", ["tiny_auto", "tiny_fixed", "acts_selu", "bipedal"])
def test_sac_training_step_reproduces_the_reference_run(name):
    """Construct SAC like main.py does, push the same transitions, call training_step() K times: with
    train.rng=host the agent consumes Python's `random` and torch's CPU generator exactly like the reference,
    so it draws the same indices and normals and must land on the same numbers (fp32 tolerance)."""
    g = Golden(name)
    agent = _agent_from_golden(g)
    # F10: bit-identical initial weights through the same torch init calls
    for tag, net in (("pi", agent.policy_net), ("q1", agent.q_net1), ("q2", agent.q_net2)):
        for k, v in net.state_dict().items():
            assert np.array_equal(v.cpu().numpy(), g[f"init/{tag}/{k}"]), (tag, k)
    assert torch.equal(agent.q_net1_target.net[0].weight, agent.q_net1.net[0].weight)
    for k in range(g.K):
        agent.training_step()
        m = agent.last_metrics()
        eng = agent.engine
        assert np.array_equal(eng.view("batch.idx").cpu().numpy().ravel(), g[f"step{k}/idx"])      # same index stream
        assert np.array_equal(eng.view("batch.eps1").cpu().numpy(), g[f"step{k}/eps1"])
        assert np.array_equal(eng.view("batch.eps2").cpu().numpy(), g[f"step{k}/eps2"])
        tol = 3e-5 * 3 ** k
        assert_close("y", eng.view("out.y").cpu().numpy().ravel(), g[f"step{k}/y"], tol)
        assert_close("logpi", eng.view("out.logpi").cpu().numpy().ravel(), g[f"step{k}/lp"], tol)
        if g.cfg["sac"]["auto_entropy_tuning"]:
            assert abs(float(agent.log_alpha) - float(g[f"step{k}/log_alpha"])) < 2e-6
            assert agent.log_alpha.dtype == torch.float64 and agent.alpha.dtype == torch.float64     # F6
            assert abs(float(agent.alpha) - float(g[f"step{k}/alpha"])) < 2e-6
        else:
            assert agent.alpha.dtype == torch.float32
    if g.full:
        for k2, v in agent.policy_net.state_dict().items():
            assert_close(k2, v.cpu().numpy(), g[f"step{g.K - 1}/pi/{k2}"], 1e-4 * 2 ** g.K)
    # a13: select_action on the recorded state with the next normal of the same generator
    got_s = agent.select_action(g["act/state"])
    got_d = agent.select_action(g["act/state"], deterministic=True)
    assert got_s.shape == (g.act,) and got_s.dtype == np.float32
    assert_close("select_action stochastic", got_s, g["act/stochastic"], 2e-4)
    assert_close("select_action deterministic", got_d, g["act/deterministic"], 2e-4)


def test_sac_phase_methods_match_reference_signatures():
    """The five per-phase methods of the reference (agent.py:195-300) driven from Python like training_step does."""
    g = Golden("acts_leaky_relu")
    agent = _agent_from_golden(g)
    idx = g["step0/idx"]
    random.seed(g.cfg["train"]["seed"])
    torch.manual_seed(g.cfg["train"]["seed"])
    batch = agent.sample_batch()
    assert type(batch).__name__ == "Transition" and batch.state.is_cuda and batch.state.dtype == torch.float32
    S, A, R, S2, D = synth_transitions(g.n_fill, g.obs, g.act)
    assert np.array_equal(batch.reward.cpu().numpy(), R[idx])
    y = agent.compute_target_q_values(rewards=batch.reward, dones=batch.done, next_states=batch.next_state)
    assert_close("y", y.cpu().numpy(), g["step0/y"], 2e-5)
    agent.update_q_networks(states=batch.state, actions=batch.action, target_q_values=y)
    log_pi = agent.update_policy_network(states=batch.state)
    assert_close("logpi", log_pi.cpu().numpy(), g["step0/lp"], 2e-5)
    info = agent.update_entropy_temperature(log_pi=log_pi)
    assert set(info) == {"alpha_loss", "alpha"}
    assert abs(info["alpha"] - float(g["step0/alpha"])) < 1e-6
    agent.soft_update_target_networks()
    for k, v in agent.q_net1_target.state_dict().items():
        assert_close(k, v.cpu().numpy(), g[f"step0/q1t/{k}"], 1e-5)
    for k, v in agent.policy_net.state_dict().items():
        assert_close(k, v.cpu().numpy(), g[f"step0/pi/{k}"], 1e-4)


def test_checkpoint_schema_and_round_trip(tmp_path, golden_dir):
    """save_agent writes the reference's schema (agent.py:521-536); load_agent restores every tensor."""
    g = Golden("tiny_auto")
    agent = _agent_from_golden(g)
    for _ in range(3):
        agent.training_step()
    path = str(tmp_path / "sac_agent.pth")
    agent.save_agent(path)
    ck = torch.load(path, map_location="cpu")

    def describe(v):
        if torch.is_tensor(v):
            return {"tensor": list(v.shape), "dtype": str(v.dtype)}
        if isinstance(v, dict):
            return {str(k): describe(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return [describe(x) for x in v]
        return {"py": type(v).__name__}

    def strip(d):
        if isinstance(d, dict):
            return {k: strip(v) for k, v in d.items() if k != "value"}
        if isinstance(d, list):
            return [strip(x) for x in d]
        return d

    with open(os.path.join(golden_dir, "checkpoint_schema.json")) as f:
        ref = json.load(f)
    assert strip(describe(ck)) == strip(ref)
    assert float(ck["policy_optimizer_state_dict"]["state"][0]["step"]) == 3.0
    before = {k: v.clone() for k, v in agent.policy_net.state_dict().items()}
    m_before = agent.engine.view("m.q1.W0").clone()
    la = float(agent.log_alpha)
    for _ in range(2):
        agent.training_step()
    assert not torch.equal(before["net.0.weight"], agent.policy_net.state_dict()["net.0.weight"])
    agent.load_agent(path)
    for k, v in agent.policy_net.state_dict().items():
        assert torch.equal(v, before[k])
    assert torch.equal(agent.engine.view("m.q1.W0"), m_before)
    assert float(agent.log_alpha) == la and abs(float(agent.alpha) - np.exp(la)) < 1e-12
    assert int(agent.engine.view("scal.step")[0]) == 3
    agent.training_step()                                   # continues from the restored optimiser state
    assert agent.last_metrics()["nonfinite"] == 0


def test_underfilled_buffer_and_config_errors():
    from sac.agent import SAC
    cfg = base_config(hidden=(16, 16), batch=32)
    agent = SAC(FakeEnv(3, 1), cfg)
    assert not agent.can_update()
    with pytest.raises(ValueError, match="Not enough samples"):
        agent.training_step()
    bad = base_config(hidden=(16, 16), act="swish")
    with pytest.raises(KeyError):
        SAC(FakeEnv(3, 1), bad)
    cpu = base_config(hidden=(16, 16))
    cpu["train"]["device"] = "cpu"
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SAC(FakeEnv(3, 1), cpu)


class ConstantRewardEnv:
    """One-step episodes, reward 1, always terminal (the reference's probe env, sac/envs.py:15-46, restated):
    the target is exactly y = 1, so both critics must converge to Q = 1."""
    spec = None

    def __init__(self):
        from gpu_helpers import _Space
        self.observation_space, self.action_space = _Space(1), _Space(1)
        self.rng = np.random.default_rng(0)

    def reset(self, seed=None):
        return np.zeros(1, np.float32), {}

    def step(self, action):
        return np.zeros(1, np.float32), 1.0, True, False, {}


class PointMassEnv:
    """1-D point mass (sac/envs.py:161-222 restated): move to x=1 +- 0.05 with |a| <= 0.1, -0.01 per step, +1 at goal."""
    spec = None

    def __init__(self):
        from gpu_helpers import _Space
        self.observation_space, self.action_space = _Space(1), _Space(1)

    def reset(self, seed=None):
        self.pos, self.t = 0.0, 0
        return np.array([self.pos], np.float32), {}

    def step(self, action):
        self.t += 1
        self.pos += float(np.clip(action[0], -0.1, 0.1))
        reached = abs(self.pos - 1.0) <= 0.05
        return np.array([self.pos], np.float32), -0.01 + (1.0 if reached else 0.0), reached, self.t >= 50, {}


def test_known_answer_constant_reward_env():
    """Analytic pin: every transition terminal with r = 1  =>  Q1, Q2 -> 1."""
    from sac.agent import SAC
    cfg = base_config(hidden=(64, 64), batch=64, auto=False, alpha=0.1, rng="device")
    cfg["train"]["warming_steps"] = 100
    agent = SAC(ConstantRewardEnv(), cfg)
    out = agent.run_training_loop(num_episodes=900, tqdm_disable=True)
    assert set(out) == {"total_episodes", "best_avg_return", "final_avg_return"} and out["final_avg_return"] == 1.0
    q1, q2 = agent.engine.q_values_host(np.zeros((4, 1), np.float32), np.linspace(-1, 1, 4, dtype=np.float32)[:, None])
    assert np.all(np.abs(q1 - 1.0) < 0.05) and np.all(np.abs(q2 - 1.0) < 0.05), (q1, q2)
    m = agent.last_metrics()
    assert m["updates"] == 900 - 99 and abs(m["y_mean"] - 1.0) < 1e-6


def test_point_mass_learns_through_the_public_loop():
    """BASELINE config 1 (plumbing + behaviour): train through run_training_loop / eval_agent on the device engine."""
    from sac.agent import SAC
    cfg = base_config(hidden=(256, 256), batch=256, auto=False, alpha=0.02, rng="device")
    cfg["train"]["warming_steps"] = 1000
    agent = SAC(PointMassEnv(), cfg)
    out = agent.run_training_loop(num_episodes=260, tqdm_disable=True)
    ret = agent.eval_agent(num_episodes=5, tqdm_disable=True)
    # reference last-100 return 0.863 (optimum 0.90); random policy ~ -0.5
    assert out["final_avg_return"] > 0.5 and ret > 0.8, (out, ret)


def test_gradient_steps_burst_and_push_device():
    """UTD > 1 bursts collapse into one launch; device-side producer pushes without a host bounce (8f-4)."""
    from sac.agent import SAC
    cfg = base_config(hidden=(32, 32), batch=64, rng="device")
    agent = SAC(FakeEnv(6, 2), cfg)
    s = torch.randn(500, 6, device="cuda")
    agent.replay_buffer.push_device(s, torch.rand(500, 2, device="cuda") * 2 - 1, torch.randn(500, device="cuda"),
                                    torch.randn(500, 6, device="cuda"), torch.zeros(500, device="cuda"))
    assert len(agent.replay_buffer) == 500
    got = agent.replay_buffer.sample_tensors(10, indices=list(range(10)))
    assert torch.equal(got.state, s[:10])
    l0 = agent.engine.launch_count()
    agent.training_steps(5)
    assert agent.last_metrics()["updates"] == 5 and agent.engine.launch_count() - l0 == 1


def test_snapshot_resumes_bit_exactly(tmp_path):
    """save_snapshot / load_snapshot (arena head + ring image + RNG states) -- a fresh agent that loads the
    snapshot continues exactly like the original: same parameters after M more updates, in both RNG modes, and the same
    rollout actions."""
    import random
    from gpu_helpers import FakeEnv, base_config
    from helpers import synth_transitions
    from sac.agent import SAC
    obs, act = 6, 2
    for mode in ("device", "host"):
        cfg = base_config(hidden=(64, 64), batch=64, capacity=500, rng=mode)
        a = SAC(FakeEnv(obs, act), cfg)
        s, ac, r, s2, d = synth_transitions(700, obs, act, 4)          # wraps the 500-slot ring
        for i in range(700):
            a.store_transition(s[i], ac[i], float(r[i]), s2[i], bool(d[i]))
        random.seed(3); torch.manual_seed(3)
        for _ in range(5):
            a.training_step()
        a.select_action(s[0])                                            # advances the rollout-noise counter
        path = str(tmp_path / f"snap_{mode}.pt")
        a.save_snapshot(path)
        for _ in range(4):
            a.training_step()
        act_a = a.select_action(s[1])
        with pytest.raises(ValueError):
            SAC(FakeEnv(obs, act), base_config(hidden=(64, 64), batch=64, capacity=500, rng=mode, seed=9)).load_snapshot(path)
        b = SAC(FakeEnv(obs, act), cfg)
        b.engine.view("block.params").mul_(1.7)                          # a state that is nothing like the snapshot's
        for i in range(80):
            b.store_transition(s[i] * 3, ac[i], 1.0, s2[i], False)
        b.training_step()
        b.load_snapshot(path)
        assert len(b.replay_buffer) == 500
        for _ in range(4):
            b.training_step()
        act_b = b.select_action(s[1])
        for n in ("block.params", "block.targets", "block.m", "block.v"):
            assert torch.equal(a.engine.view(n), b.engine.view(n)), (mode, n)
        assert float(a.engine.view("scal.log_alpha")) == float(b.engine.view("scal.log_alpha"))
        assert int(a.engine.view("scal.updates")) == int(b.engine.view("scal.updates")) == 9
        assert np.array_equal(act_a, act_b)


def test_population_act_all_matches_per_agent_act():
    """one launch returns every agent's action on its own observation; equal to per-agent sacx_act."""
    from gpu_helpers import base_config
    from sac.population import SACPopulation
    obs, act, n = 5, 3, 7
    pop = SACPopulation(obs, act, base_config(hidden=(32, 48), batch=16, capacity=100), n, reference_init=True)
    rng = np.random.default_rng(0)
    states = rng.standard_normal((n, obs)).astype(np.float32)
    det = pop.act_all(states, deterministic=True).cpu().numpy()
    for ag in range(n):
        assert np.array_equal(det[ag], pop.act(ag, states[ag], deterministic=True))
    eps = torch.as_tensor(rng.standard_normal((n, 1, act)).astype(np.float32)).cuda()
    sto = pop.act_all(states, deterministic=False, eps=eps).cpu().numpy()
    for ag in range(n):
        one = pop.engine.act(torch.as_tensor(states[ag:ag + 1]).cuda(), eps[ag], deterministic=False, agent=ag).cpu().numpy()[0]
        assert np.array_equal(sto[ag], one)
    assert np.all(np.abs(sto) <= 1.0) and not np.array_equal(det, sto)


def test_gather_never_reads_outside_the_ring():
    """Advisor finding (round 1): shapes of batched pushes are validated; host index lists are range-checked; a DEVICE index
    tensor is not read back (no sync per gather) -- the kernels return a zero row for a position outside [0, len)."""
    from sac.replay_buffer import ReplayBuffer
    rb = ReplayBuffer(100, 8, 4)
    n = 60
    s = np.arange(n * 8, dtype=np.float32).reshape(n, 8) + 1
    a = np.ones((n, 4), np.float32)
    rb.push_batch(s, a, np.arange(n), s + 0.5, np.zeros(n))
    with pytest.raises(ValueError):
        rb.push_batch(s[:, :7], a, np.arange(n), s, np.zeros(n))              # wrong observation width
    with pytest.raises(ValueError):
        rb.push_batch(s, a, np.arange(n), s, np.zeros(n - 1))                 # one done flag short
    with pytest.raises(ValueError):
        rb.push_device(torch.zeros(4, 8).cuda(), torch.zeros(4, 3).cuda(), torch.zeros(4).cuda(), torch.zeros(4, 8).cuda(), torch.zeros(4).cuda())
    with pytest.raises(ValueError, match="out of range"):
        rb.sample_tensors(3, indices=[0, 5, 60])
    with pytest.raises(ValueError, match="out of range"):
        rb.sample_tensors(2, indices=[-1, 5])
    idx = torch.tensor([3, -7, 59, 60, 10**12], dtype=torch.int64, device="cuda")
    b = rb.sample_tensors(5, indices=idx)
    got = b.state.cpu().numpy()
    assert np.array_equal(got[0], s[3]) and np.array_equal(got[2], s[59])
    assert not got[1].any() and not got[3].any() and not got[4].any() and b.reward.cpu().numpy()[3] == 0.0
    odd = ReplayBuffer(50, 3, 2)                                             # dimensions that are not multiples of 4: scalar kernel
    odd.push_batch(np.ones((10, 3), np.float32), np.ones((10, 2), np.float32), np.ones(10), np.ones((10, 3)), np.zeros(10))
    bo = odd.sample_tensors(3, indices=torch.tensor([0, 10, -1], dtype=torch.int64, device="cuda"))
    assert bo.state.cpu().numpy()[0].all() and not bo.state.cpu().numpy()[1:].any()
