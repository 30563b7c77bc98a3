"""-m gpu tests of the multi-GPU modes, run on ONE GPU: ranks are emulated as separate engine instances driven in
lockstep (the exchange points are explicit host calls, so no kernel waits on another rank)."""
import numpy as np
import pytest
import torch

from gpu_helpers import assert_close, base_config, dev, fill_ring, load_nets
from helpers import Golden, synth_transitions

pytestmark = pytest.mark.gpu


def _dp_engines(cfg, obs, act, global_batch, world, n_fill):
    from sac.population import DataParallelSAC
    ranks = []
    for r in range(world):
        dp = DataParallelSAC(obs, act, cfg, global_batch, rank=r, world=world)
        fill_ring(dp.ring, n_fill, obs, act)          # ring replicated on every rank
        ranks.append(dp)
    return ranks


def _lockstep_update(ranks, idx=None, e1=None, e2=None):
    """What DataParallelSAC.update does, with the NCCL all-reduce replaced by an in-process sum over the ranks."""
    def allreduce(name):
        views = [getattr(dp, name) for dp in ranks]
        total = torch.stack(views).sum(0)
        for v in views:
            v.copy_(total)

    G, Bl = len(ranks), ranks[0].local_batch
    sl = lambda t, r: None if t is None else t[r * Bl:(r + 1) * Bl].contiguous()
    for r, dp in enumerate(ranks):
        dp.segment_critic_grads(sl(idx, r), sl(e1, r))
    allreduce("g_critics")
    for r, dp in enumerate(ranks):
        dp.segment_critic_apply_actor_grads(sl(e2, r))
    allreduce("g_policy")                       # policy gradients + the temperature-gradient share: one message
    for dp in ranks:
        dp.segment_actor_apply()


@pytest.mark.parametrize("world", [2, 4])
def test_data_parallel_matches_single_rank(world):
    """Global batch split over G ranks + gradient sums == the same global batch on one rank (summation order only)."""
    obs, act, B = 6, 2, 256
    cfg = base_config(hidden=(64, 64), batch=B, capacity=4000)
    rng = np.random.default_rng(1)
    K = 3
    idx = [torch.as_tensor(rng.integers(0, 3000, B).astype(np.int64)).cuda() for _ in range(K)]
    e1 = [torch.as_tensor(rng.standard_normal((B, act)).astype(np.float32)).cuda() for _ in range(K)]
    e2 = [torch.as_tensor(rng.standard_normal((B, act)).astype(np.float32)).cuda() for _ in range(K)]
    single = _dp_engines(cfg, obs, act, B, 1, 3000)
    multi = _dp_engines(cfg, obs, act, B, world, 3000)
    for k in range(K):
        _lockstep_update(single, idx[k], e1[k], e2[k])
        _lockstep_update(multi, idx[k], e1[k], e2[k])
    ref = single[0].engine
    for dp in multi:
        assert_close("params", dp.engine.view("block.params").cpu().numpy(), ref.view("block.params").cpu().numpy(), 2e-5)
        assert_close("targets", dp.engine.view("block.targets").cpu().numpy(), ref.view("block.targets").cpu().numpy(), 2e-6)
        assert abs(float(dp.engine.view("scal.log_alpha")) - float(ref.view("scal.log_alpha"))) < 1e-6
        assert int(dp.engine.view("scal.updates")) == K
    # every rank holds the same replicated parameters
    assert torch.equal(multi[0].engine.view("block.params"), multi[1].engine.view("block.params"))


def test_data_parallel_equals_fused_update_and_device_rng_is_rank_invariant():
    """(1) the segmented gradient/apply path lands on the fused kernel's result; (2) with device RNG the global batch
    (indices, normals keyed by global row) does not depend on G."""
    from sac.engine import UpdateEngine
    from sac.replay_buffer import ReplayBuffer
    g = Golden("tiny_auto")
    B = g.cfg["train"]["batch_size"]
    dp = _dp_engines(g.cfg, g.obs, g.act, B, 1, g.n_fill)[0]
    load_nets(dp.engine, {"pi": g.sd("init/pi"), "q1": g.sd("init/q1"), "q2": g.sd("init/q2")})
    dp.engine.reset_state()
    for k in range(g.K):
        _lockstep_update([dp], dev(g[f"step{k}/idx"]), dev(g[f"step{k}/eps1"]), dev(g[f"step{k}/eps2"]))
        assert_close("y", dp.engine.view("out.y").cpu().numpy().ravel(), g[f"step{k}/y"], 3e-5 * 3 ** k)
        assert abs(float(dp.engine.view("scal.log_alpha")) - float(g[f"step{k}/log_alpha"])) < 2e-6
    for tag in ("pi", "q1", "q1t"):
        for kk, v in g.sd(f"step{g.K - 1}/{tag}").items():
            l = int(kk.split(".")[1]) // 2
            name = f"{tag}.{'W' if kk.endswith('weight') else 'b'}{l}"
            assert_close(name, dp.engine.view(name).cpu().numpy().reshape(v.shape), v, 4e-4)
    # device RNG: G = 1 vs G = 2 draw the same global batch
    cfg = base_config(hidden=(32, 32), batch=64, capacity=2000, rng="device")
    one = _dp_engines(cfg, 5, 2, 64, 1, 1500)
    two = _dp_engines(cfg, 5, 2, 64, 2, 1500)
    _lockstep_update(one)
    _lockstep_update(two)
    idx1 = one[0].engine.view("batch.idx").cpu().numpy().ravel()
    idx2 = np.concatenate([dp.engine.view("batch.idx").cpu().numpy().ravel() for dp in two])
    assert np.array_equal(idx1, idx2) and len(np.unique(idx1)) == 64
    eps1 = one[0].engine.view("batch.eps2").cpu().numpy()
    eps2 = np.concatenate([dp.engine.view("batch.eps2").cpu().numpy() for dp in two])
    assert np.array_equal(eps1, eps2)
    assert_close("params G=2 vs G=1", two[0].engine.view("block.params").cpu().numpy(), one[0].engine.view("block.params").cpu().numpy(), 2e-5)


def test_large_batch_split_k_gradients():
    """Batch >= 2048 switches dW to split-K over CTAs with atomic accumulation: same gradients as the oracle."""
    from oracle.sac_numpy import Hyper, SACOracle, mlp_from_state_dict
    from sac.engine import UpdateEngine
    obs, act, B = 5, 2, 4096
    cfg = base_config(hidden=(32, 24), batch=B, capacity=8000)
    eng = UpdateEngine(obs, act, cfg)
    torch.manual_seed(0)
    sds = {}
    for tag, dims in (("pi", [obs, 32, 24, 2 * act]), ("q1", [obs + act, 32, 24, 1]), ("q2", [obs + act, 32, 24, 1])):
        sd = {}
        for l in range(3):
            sd[f"net.{2 * l}.weight"] = (torch.randn(dims[l + 1], dims[l]) * 0.3).numpy()
            sd[f"net.{2 * l}.bias"] = (torch.randn(dims[l + 1]) * 0.1).numpy()
        sds[tag] = sd
    load_nets(eng, sds)
    eng.reset_state()
    S, A, R, S2, D = synth_transitions(B, obs, act)
    D = D.astype(np.float32)
    rng = np.random.default_rng(2)
    e1, e2 = rng.standard_normal((B, act)).astype(np.float32), rng.standard_normal((B, act)).astype(np.float32)
    o = SACOracle(mlp_from_state_dict(sds["pi"], "relu", "identity"), mlp_from_state_dict(sds["q1"], "relu", "identity"),
                  mlp_from_state_dict(sds["q2"], "relu", "identity"), Hyper())
    eng.load_batch(dev(S), dev(A), dev(R), dev(S2), dev(D))
    y = torch.empty(B, device="cuda")
    eng.target(dev(e1), y)
    yo = o.target(R, D, S2, e1)
    assert_close("y", y.cpu().numpy(), yo, 2e-5)
    eng.critic_step(y, grads_only=True)
    cg = o.critic_grads(S, A, yo)
    for l in range(3):
        assert_close(f"dW{l}", eng.view(f"g.q1.W{l}").cpu().numpy(), cg["q1"]["dW"][l], 1e-4)
        assert_close(f"db{l}", eng.view(f"g.q2.b{l}").cpu().numpy().ravel(), cg["q2"]["db"][l], 1e-4)
    eng.actor_step(dev(e2), None, grads_only=True)
    ag = o.actor_grads(S, e2)
    for l in range(3):
        assert_close(f"pi dW{l}", eng.view(f"g.pi.W{l}").cpu().numpy(), ag["dW"][l], 2e-4)


def test_population_class_shards_and_matches_reference_init():
    from sac.population import SACPopulation, shard_agents
    g = Golden("tiny_auto")
    cfg = dict(g.cfg)
    cfg["train"] = dict(cfg["train"], device="cuda")
    pops = [SACPopulation(g.obs, g.act, cfg, n_agents=5, rank=r, world=2) for r in range(2)]
    assert [p.agent_ids for p in pops] == [[0, 1, 2], [3, 4]]
    # agent with seed 0 starts from the reference's weights for seed 0
    assert np.array_equal(pops[0].engine.view("pi.W0", 0).cpu().numpy(), g["init/pi/net.0.weight"])
    assert not torch.equal(pops[0].engine.view("pi.W0", 1), pops[0].engine.view("pi.W0", 0))
    S, A, R, S2, D = synth_transitions(60, g.obs, g.act)
    for p in pops:
        for a in range(p.n_local):
            p.push_batch(a, S, A, R, S2, D.astype(np.float32))
        p.update(4)
        assert all(p.metrics(a)["updates"] == 4 and p.metrics(a)["nonfinite"] == 0 for a in range(p.n_local))
    losses = pops[0].gather_metrics("q1_loss")
    assert losses.shape == (3,) and np.all(np.isfinite(losses)) and len(set(losses.tolist())) == 3
    sd = pops[1].agent_state_dict(0)
    assert set(sd) == {"policy_net_state_dict", "q_net1_state_dict", "q_net2_state_dict", "q_net1_target_state_dict", "q_net2_target_state_dict"}
    a = pops[0].act(0, S[0])
    assert a.shape == (g.act,) and np.all(np.abs(a) <= 1.0)


def test_population_trials_match_single_agents_with_those_hyperparameters():
    """the reference's Optuna study varies sac.alpha and sac.alpha_lr per trial
    (hparam_search/configs/search_space.yaml). A population whose agents carry per-agent (alpha, alpha_lr) must follow, agent
    by agent, a single-agent engine built from a config with those values (same weights, ring, indices and normals)."""
    from sac.engine import UpdateEngine
    from sac.population import SACPopulation
    from sac.replay_buffer import ReplayBuffer
    from test_gpu_parity import _random_nets
    obs, act, B, K = 5, 2, 64, 3
    trials = [{"alpha": 0.2, "alpha_lr": 3e-4}, {"alpha": 0.004, "alpha_lr": 2e-2}, {"alpha": 0.05, "alpha_lr": 1e-5}]
    cfg = base_config(hidden=(64, 64), batch=B, capacity=600, alpha=0.2)
    pop = SACPopulation(obs, act, cfg, len(trials), reference_init=False)
    nets = _random_nets(obs, act, (64, 64), (64, 64), scale=0.2)
    s, a, r, s2, d = synth_transitions(500, obs, act, 1)
    for ag in range(len(trials)):
        load_nets(pop.engine, nets, agent=ag)
    pop.engine.reset_state()
    pop.set_trials(trials)
    for ag in range(len(trials)):
        pop.ring.push_batch(s, a, r, s2, d.astype(np.float32), agent=ag)
    rng = np.random.default_rng(2)
    idx = rng.integers(0, 500, (K, len(trials), B)).astype(np.int64)
    e1 = rng.standard_normal((K, len(trials), B, act)).astype(np.float32)
    e2 = rng.standard_normal((K, len(trials), B, act)).astype(np.float32)
    pop.engine.update(dev(idx), dev(e1), dev(e2), K)
    pop.engine.sync()
    la = []
    for ag, t in enumerate(trials):
        c1 = base_config(hidden=(64, 64), batch=B, capacity=600, alpha=t["alpha"])
        c1["sac"]["alpha_lr"] = t["alpha_lr"]
        one = UpdateEngine(obs, act, c1)
        load_nets(one, nets)
        one.reset_state()
        rb = ReplayBuffer(600, obs, act)
        rb.push_batch(s, a, r, s2, d.astype(np.float32))
        one.attach_ring(rb)
        one.update(dev(idx[:, ag]), dev(e1[:, ag]), dev(e2[:, ag]), K)
        one.sync()
        assert_close(f"agent {ag} params", pop.engine.view("block.params", ag).cpu().numpy(), one.view("block.params").cpu().numpy(), 3e-5)
        got, want = float(pop.engine.view("scal.log_alpha", ag)), float(one.view("scal.log_alpha"))
        assert abs(got - want) < 1e-6 * max(1.0, abs(want)), (ag, got, want)
        la.append(got)
    assert len({round(x, 6) for x in la}) == len(trials)              # the trials really differ


@pytest.mark.parametrize("path", ["ffma", "tc"])
def test_population_trials_full_hyperparameter_vectors(path, monkeypatch):
    """Every scalar of the YAML's `sac` section per agent (run_search.py:24-39 is generic over section/param): actor_lr,
    critic_lr, tau, gamma next to alpha / alpha_lr. Each agent of the population must follow a single-agent engine built from a
    config with that trial's values -- through the one-CTA-per-agent FFMA kernel and through the population tensor-core path."""
    from sac.engine import UpdateEngine
    from sac.population import SACPopulation
    from sac.replay_buffer import ReplayBuffer
    from test_gpu_parity import _random_nets
    tc = path == "tc"
    obs, act = 5, 2
    B, K, n = (128, 2, 128) if tc else (64, 3, 4)                    # tensor-core population path: >= 16384 rows in total
    monkeypatch.setenv("SACX_TC_POP", "1" if tc else "0")
    base = [{"alpha": 0.2, "alpha_lr": 3e-4, "actor_lr": 1e-3, "critic_lr": 3e-5, "tau": 0.05, "gamma": 0.9},
            {"alpha": 0.004, "alpha_lr": 2e-2, "actor_lr": 1e-5, "critic_lr": 2e-3, "tau": 0.001, "gamma": 0.999},
            {"critic_lr": 5e-4, "gamma": 0.5},
            {}]
    trials = [dict(base[i % 4]) for i in range(n)]
    cfg = base_config(hidden=(64, 64), batch=B, capacity=600, alpha=0.1)
    pop = SACPopulation(obs, act, cfg, n, reference_init=False)
    assert pop.engine.tensor_core()[0] == tc, pop.engine.tensor_core()
    nets = _random_nets(obs, act, (64, 64), (64, 64), scale=0.2)
    s, a, r, s2, d = synth_transitions(500, obs, act, 1)
    for ag in range(n):
        load_nets(pop.engine, nets, agent=ag)
    pop.engine.reset_state()
    pop.set_trials(trials)
    with pytest.raises(KeyError):
        pop.set_trials([{"batch_size": 3}] * n)                      # structural keys are not per-agent
    for ag in range(n):
        pop.ring.push_batch(s, a, r, s2, d.astype(np.float32), agent=ag)
    rng = np.random.default_rng(2)
    idx = np.broadcast_to(rng.integers(0, 500, (K, 1, B)), (K, n, B)).astype(np.int64).copy()
    e1 = np.broadcast_to(rng.standard_normal((K, 1, B, act)), (K, n, B, act)).astype(np.float32).copy()
    e2 = np.broadcast_to(rng.standard_normal((K, 1, B, act)), (K, n, B, act)).astype(np.float32).copy()
    pop.engine.update(dev(idx), dev(e1), dev(e2), K)
    pop.engine.sync()
    finals = []
    for ag in range(4):
        t = trials[ag]
        c1 = base_config(hidden=(64, 64), batch=B, capacity=600, alpha=t.get("alpha", 0.1))
        for k in ("alpha_lr", "actor_lr", "critic_lr", "tau", "gamma"):
            if k in t:
                c1["sac"][k] = t[k]
        monkeypatch.setenv("SACX_ROWPAR", "0")
        one = UpdateEngine(obs, act, c1)
        load_nets(one, nets)
        one.reset_state()
        rb = ReplayBuffer(600, obs, act)
        rb.push_batch(s, a, r, s2, d.astype(np.float32))
        one.attach_ring(rb)
        one.update(dev(idx[:, ag]), dev(e1[:, ag]), dev(e2[:, ag]), K)
        one.sync()
        tol = 1e-4 if tc else 3e-5                                   # 3xTF32 tiles vs FFMA tiles
        for blk in ("block.params", "block.targets"):
            assert_close(f"agent {ag} {blk}", pop.engine.view(blk, ag).cpu().numpy(), one.view(blk).cpu().numpy(), tol)
        assert_close(f"agent {ag} y", pop.engine.view("out.y", ag).cpu().numpy(), one.view("out.y").cpu().numpy(), 2e-5)
        got, want = float(pop.engine.view("scal.log_alpha", ag)), float(one.view("scal.log_alpha"))
        assert abs(got - want) < 2e-6 * max(1.0, abs(want)), (ag, got, want)
        finals.append(pop.engine.view("block.params", ag).cpu().numpy().copy())
        if n > 4:                                                    # agents 4.. repeat the trials of agents 0..3
            assert_close(f"agent {ag + 4} == agent {ag}", pop.engine.view("block.params", ag + 4).cpu().numpy(), finals[ag], 1e-6)
    assert not np.array_equal(finals[0], finals[1]) and not np.array_equal(finals[2], finals[3])


def test_population_rng_streams_are_keyed_by_global_agent_and_seed():
    """Advisor finding (round 1): the device index / normal streams were keyed by the LOCAL agent index, so local agent 0 of
    two ranks drew identical permutations and noise. Now: (train.seed or the agent's own seed, GLOBAL agent id)."""
    from sac.population import SACPopulation
    obs, act, B = 4, 1, 64
    cfg = base_config(hidden=(32, 32), batch=B, capacity=2000, rng="device")
    s, a, r, s2, d = (torch.from_numpy(x).cuda() for x in synth_transitions(1500, obs, act, 3))

    def draws(pop):
        pop.push_device_all(s, a, r, s2, d.float())
        pop.update(1)
        pop.engine.sync()
        return [(pop.engine.view("batch.idx", ag).cpu().numpy().copy(), pop.engine.view("batch.eps1", ag).cpu().numpy().copy(),
                 pop.engine.act_population(torch.zeros(pop.n_local, 1, obs, device="cuda")).cpu().numpy()[ag].copy())
                for ag in range(pop.n_local)]

    r0 = draws(SACPopulation(obs, act, cfg, 4, rank=0, world=2, reference_init=False))
    r1 = draws(SACPopulation(obs, act, cfg, 4, rank=1, world=2, reference_init=False))
    whole = draws(SACPopulation(obs, act, cfg, 4, rank=0, world=1, reference_init=False))
    for x in r0 + r1:
        assert len(np.unique(x[0])) == B
    assert not np.array_equal(r0[0][0], r1[0][0]) and not np.array_equal(r0[0][1], r1[0][1])       # same local index, other rank
    for g, x in enumerate(r0 + r1):                    # sharding does not change an agent's streams: global agent g either way
        assert np.array_equal(x[0], whole[g][0]) and np.array_equal(x[1], whole[g][1])
    # an agent's own seed replaces train.seed as the first key
    seeded = draws(SACPopulation(obs, act, cfg, 4, seeds=[11, 12, 13, 14], rank=0, world=1, reference_init=False))
    again = draws(SACPopulation(obs, act, cfg, 4, seeds=[11, 12, 13, 14], rank=0, world=1, reference_init=False))
    assert not np.array_equal(seeded[1][0], whole[1][0])
    assert all(np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) for x, y in zip(seeded, again))
