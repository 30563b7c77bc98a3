"""Record golden vectors from the REAL reference (runs only in the build container).

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz, *.json

The reference (/root/reference, read-only, untouched) is imported with two stub
modules for packages that are not installed here (``gymnasium``, ``matplotlib``:
SURVEY.md F15) and driven through its own public API: ``SAC(env, config)``,
``store_transition``, ``training_step``, ``select_action``, ``save_agent``.  Before
every update the generator peeks (and then restores) the two RNG streams the
reference consumes -- Python's global ``random`` (index stream, replay_buffer.py:39)
and torch's CPU generator (two N(0,1) draws of shape [B, act], models.py:82-83) --
so that the recorded (indices, eps1, eps2) are exactly what the update used.

Nothing in this script imports the repo's own ``sac`` package or the oracle.
The outputs are the pin for oracle/ (tests/test_oracle_golden.py) and, through
it, for the CUDA path.
"""
from __future__ import annotations

import copy
import json
import os
import random
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    gym = types.ModuleType("gymnasium")

    class _Env:  # only used as a type annotation by the reference agent
        pass

    gym.Env = _Env
    gym.make = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no gymnasium here"))
    sys.modules["gymnasium"] = gym
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.style = types.SimpleNamespace(use=lambda *a, **k: None)
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    sys.path.insert(0, REF)
    import sac.agent as agent_mod  # noqa: E402
    import sac.replay_buffer as rb_mod  # noqa: E402

    assert agent_mod.__file__.startswith(REF)
    return agent_mod, rb_mod


class _Space:
    def __init__(self, n):
        self.shape = (n,)

    def seed(self, s):
        return [s]


class FakeEnv:
    spec = None

    def __init__(self, obs, act):
        self.observation_space = _Space(obs)
        self.action_space = _Space(act)

    def reset(self, seed=None):
        return np.zeros(self.observation_space.shape, np.float32), {}


def make_config(hidden, act="relu", out_act="identity", auto=True, batch=8, seed=0, capacity=1000,
                alpha=0.1, q_hidden=None, log_std_min=-20, log_std_max=2, action_scale=1.0):
    return {
        "sac": {"gamma": 0.99, "tau": 0.005, "alpha": alpha, "auto_entropy_tuning": auto,
                "actor_lr": 3e-4, "critic_lr": 3e-4, "alpha_lr": 3e-4},
        "q_net": {"hidden_sizes": list(q_hidden or hidden), "hidden_layers_act": act, "output_activation": out_act},
        "policy_net": {"hidden_sizes": list(hidden), "hidden_layers_act": act, "output_activation": out_act,
                       "log_std_min": log_std_min, "log_std_max": log_std_max, "action_scale": action_scale},
        "buffer": {"capacity": capacity},
        "train": {"gradient_steps_per_update": 1, "seed": seed, "batch_size": batch, "warming_steps": 10, "device": "cpu"},
        "logger": {"enabled": False, "log_dir": "runs", "env_name": "Synthetic", "agent_name": "SAC", "run_name": "g",
                   "use_timestamp": False, "timestamp_format": "%Y", "flush_secs": 10, "log_episode_stats": False,
                   "log_q_values": False, "save_model": {"enabled": False, "path": None}},
    }


def synth_transitions(n, obs, act, seed=0):
    """SURVEY section 8d synthetic inputs."""
    rng = np.random.default_rng(seed)
    s = rng.standard_normal((n, obs)).astype(np.float32)
    s2 = rng.standard_normal((n, obs)).astype(np.float32)
    a = rng.uniform(-1, 1, (n, act)).astype(np.float32)
    r = rng.standard_normal(n).astype(np.float32)
    d = rng.random(n) < 0.01
    return s, a, r, s2, d


def _sd(module):
    return {k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def _grads(module):
    return {k: p.grad.detach().numpy().copy() for k, p in module.named_parameters()}


def _adam(opt):
    st = opt.state_dict()["state"]
    return {f"{i}.{k}": np.asarray(v.detach().numpy() if hasattr(v, "detach") else v).copy()
            for i, d in st.items() for k, v in d.items()}


def record_run(agent_mod, name, obs, act, cfg, n_fill, K, full_state=True, start_ckpt=None, big=False, stride=61, store_init=True):
    """big=True (batch 65536): the index stream and the normals are NOT stored -- the test regenerates them from the
    seeds exactly as the reference consumed them (random.seed / torch.manual_seed of SAC._set_seed, agent.py:117-124;
    checksums are stored to prove it) -- and per-row outputs are stored as strided samples + float64 checksums.
    start_ckpt: a checkpoint loaded through the reference's own load_agent (agent.py:538-554) before the first update."""
    import torch

    env = FakeEnv(obs, act)
    agent = agent_mod.SAC(env, copy.deepcopy(cfg))
    if start_ckpt is not None:
        agent.load_agent(start_ckpt)
    B = cfg["train"]["batch_size"]
    s, a, r, s2, d = synth_transitions(n_fill, obs, act)
    for i in range(n_fill):
        agent.store_transition(s[i], a[i], float(r[i]), s2[i], bool(d[i]))
    out = {}

    def put(prefix, dct):
        for k, v in dct.items():
            out[f"{prefix}/{k}"] = np.asarray(v)

    def chk(v):
        v64 = np.asarray(v, dtype=np.float64).ravel()
        return np.array([v64.sum(), (v64 * v64).sum()])

    if start_ckpt is None:       # (a run started from a checkpoint reads its start state from the .pth fixture)
        for tag, net in (("pi", agent.policy_net), ("q1", agent.q_net1), ("q2", agent.q_net2)):
            if store_init:
                put(f"init/{tag}", _sd(net))
            else:                # large nets: the init recipe (F10) is reproducible from train.seed; keep its checksums + samples
                for key, v in _sd(net).items():
                    out[f"init/{tag}/{key}#chk"] = chk(v)
                    out[f"init/{tag}/{key}#smp"] = v.ravel()[::97].copy()

    cap = {}
    orig_target = agent.compute_target_q_values
    orig_q = agent.update_q_networks
    orig_pi = agent.update_policy_network
    orig_alpha = agent.update_entropy_temperature

    def w_target(rewards, dones, next_states):
        y = orig_target(rewards=rewards, dones=dones, next_states=next_states)
        cap["y"] = y.numpy().copy()
        return y

    def w_q(states, actions, target_q_values):
        with torch.no_grad():
            q1 = agent.q_net1(states, actions)
            q2 = agent.q_net2(states, actions)
            cap["q1"], cap["q2"] = q1.numpy().copy(), q2.numpy().copy()
            cap["q1_loss"] = torch.nn.functional.mse_loss(q1, target_q_values).numpy().copy()
            cap["q2_loss"] = torch.nn.functional.mse_loss(q2, target_q_values).numpy().copy()
        orig_q(states=states, actions=actions, target_q_values=target_q_values)
        cap["gq1"], cap["gq2"] = _grads(agent.q_net1), _grads(agent.q_net2)

    def w_pi(states):
        lp = orig_pi(states=states)
        cap["lp"] = lp.detach().numpy().copy()
        cap["gpi"] = _grads(agent.policy_net)
        return lp

    def w_alpha(log_pi):
        info = orig_alpha(log_pi=log_pi)
        cap["alpha_info"] = info
        return info

    agent.compute_target_q_values = w_target
    agent.update_q_networks = w_q
    agent.update_policy_network = w_pi
    agent.update_entropy_temperature = w_alpha

    for k in range(K):
        # peek the index stream and the two normal draws, then restore both generators
        st = random.getstate()
        idx = random.sample(range(len(agent.replay_buffer)), B)
        random.setstate(st)
        ts = torch.get_rng_state()
        e1 = torch.empty(B, act).normal_()
        e2 = torch.empty(B, act).normal_()
        torch.set_rng_state(ts)
        alpha_before = float(agent.alpha)
        agent.training_step()
        if big:
            out[f"step{k}/idx#chk"] = chk(idx)
            out[f"step{k}/idx#smp"] = np.asarray(idx, dtype=np.int64)[::stride].copy()
            out[f"step{k}/eps1#chk"] = chk(e1.numpy())
            out[f"step{k}/eps2#chk"] = chk(e2.numpy())
        else:
            out[f"step{k}/idx"] = np.asarray(idx, dtype=np.int64)
            out[f"step{k}/eps1"] = e1.numpy()
            out[f"step{k}/eps2"] = e2.numpy()
        out[f"step{k}/alpha_before"] = np.float64(alpha_before)
        for key in ("y", "q1", "q2", "q1_loss", "q2_loss", "lp"):
            if big and np.ndim(cap[key]) == 1:
                out[f"step{k}/{key}#chk"] = chk(cap[key])
                out[f"step{k}/{key}#smp"] = cap[key][::stride].copy()
            else:
                out[f"step{k}/{key}"] = cap[key]
        if cfg["sac"]["auto_entropy_tuning"]:
            out[f"step{k}/alpha_loss"] = np.float64(cap["alpha_info"]["alpha_loss"])
            out[f"step{k}/alpha"] = np.float64(cap["alpha_info"]["alpha"])
            out[f"step{k}/log_alpha"] = agent.log_alpha.detach().numpy().copy()
            if k == K - 1:
                st = agent.alpha_optimizer.state_dict()["state"]
                if st:      # (empty after load_agent: the reference's optimiser keeps stepping the PRE-load tensor)
                    out[f"step{k}/adam_alpha"] = np.array([float(st[0]["exp_avg"]), float(st[0]["exp_avg_sq"]), float(st[0]["step"])])
        nets = (("pi", agent.policy_net), ("q1", agent.q_net1), ("q2", agent.q_net2),
                ("q1t", agent.q_net1_target), ("q2t", agent.q_net2_target))
        if full_state:
            for tag, net in nets:
                put(f"step{k}/{tag}", _sd(net))
            put(f"step{k}/gq1", cap["gq1"])
            put(f"step{k}/gq2", cap["gq2"])
            put(f"step{k}/gpi", cap["gpi"])
            if k == K - 1:
                put(f"step{k}/adam_pi", _adam(agent.policy_optimizer))
                put(f"step{k}/adam_q1", _adam(agent.q1_optimizer))
        else:  # large shapes: float64 checksums per tensor (sum, sum of squares) + a strided sample
            if k == K - 1:
                for tag, opt in (("adam_pi", agent.policy_optimizer), ("adam_q1", agent.q1_optimizer)):
                    for key, v in _adam(opt).items():
                        out[f"step{k}/{tag}/{key}#chk"] = chk(v)
                        out[f"step{k}/{tag}/{key}#smp"] = np.asarray(v).ravel()[::97].copy()
            for tag, net in nets:
                for key, v in _sd(net).items():
                    v64 = v.astype(np.float64).ravel()
                    out[f"step{k}/{tag}/{key}#chk"] = np.array([v64.sum(), (v64 * v64).sum()])
                    out[f"step{k}/{tag}/{key}#smp"] = v.ravel()[::97].copy()
            for tag in ("gq1", "gq2", "gpi"):
                for key, v in cap[tag].items():
                    v64 = v.astype(np.float64).ravel()
                    out[f"step{k}/{tag}/{key}#chk"] = np.array([v64.sum(), (v64 * v64).sum()])
                    out[f"step{k}/{tag}/{key}#smp"] = v.ravel()[::97].copy()
    # one stochastic and one deterministic select_action on a fixed state (a13)
    st0 = s[0]
    ts = torch.get_rng_state()
    e = torch.empty(1, act).normal_()
    torch.set_rng_state(ts)
    out["act/state"] = st0
    out["act/eps"] = e.numpy()
    out["act/stochastic"] = agent.select_action(st0)
    out["act/deterministic"] = agent.select_action(st0, deterministic=True)
    meta = {"obs": obs, "act": act, "n_fill": n_fill, "K": K, "config": cfg, "full_state": full_state, "big": big,
            "stride": stride, "store_init": store_init, "start_ckpt": os.path.basename(start_ckpt) if start_ckpt else None}
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(f"wrote {name}.npz  ({len(out)} arrays)")
    return agent


def record_sampling(rb_mod):
    """F2/F3: reference ReplayBuffer.sample under a seed == random.sample(range(len), B) on
    logical deque positions; transitions carry their push number so the result is checkable."""
    cases = []
    for seed in (0, 1, 2):
        for n, k, cap in ((800, 256, 5000), (5000, 256, 5000), (5000, 1024, 5000), (3000, 1024, 5000),
                          (7000, 256, 5000), (1045, 256, 2000), (1046, 256, 2000), (300, 128, 300), (12, 12, 12)):
            rb = rb_mod.ReplayBuffer(cap)
            for p in range(n):
                rb.push(np.float32(p), p, float(p), np.float32(p), False)
            random.seed(seed)
            draws = []
            for _ in range(2):
                rows = rb.sample(k)
                draws.append([t.action for t in rows])        # push numbers returned by the reference
            cases.append({"seed": seed, "pushes": n, "k": k, "capacity": cap, "push_ids": draws})
    # under-filled buffer raises ValueError (replay_buffer.py:34-38)
    rb = rb_mod.ReplayBuffer(10)
    rb.push(0, 0, 0.0, 0, False)
    try:
        rb.sample(2)
        raised = None
    except Exception as e:  # noqa: BLE001
        raised = type(e).__name__
    with open(os.path.join(HERE, "sampling.json"), "w") as f:
        json.dump({"cases": cases, "underfilled_raises": raised}, f)
    print("wrote sampling.json")


def record_checkpoint_schema(agent, name="checkpoint_schema"):
    import tempfile

    import torch

    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "sac_agent.pth")
        agent.save_agent(p)
        ck = torch.load(p, map_location="cpu", weights_only=False)

    def describe(v):
        if hasattr(v, "shape") and hasattr(v, "dtype"):
            return {"tensor": list(v.shape), "dtype": str(v.dtype)}
        if isinstance(v, dict):
            return {str(k): describe(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return [describe(x) for x in v]
        return {"py": type(v).__name__, "value": v if isinstance(v, (int, float, bool, str, type(None))) else str(v)}

    with open(os.path.join(HERE, f"{name}.json"), "w") as f:
        json.dump(describe(ck), f, indent=1, sort_keys=True)
    print(f"wrote {name}.json")


def main():
    agent_mod, rb_mod = _import_reference()
    if "--round2-only" in sys.argv:      # keep the round-1 files byte-identical: tiny_auto is re-run into a scratch file
        a = record_run(agent_mod, "_scratch_tiny_auto", 3, 2, make_config([8, 8], batch=8, auto=True), n_fill=50, K=3)
        os.remove(os.path.join(HERE, "_scratch_tiny_auto.npz"))
        record_round2(agent_mod, a)
        return
    record_sampling(rb_mod)
    a = record_run(agent_mod, "tiny_auto", 3, 2, make_config([8, 8], batch=8, auto=True), n_fill=50, K=3)
    record_checkpoint_schema(a)
    record_run(agent_mod, "tiny_fixed", 3, 2, make_config([8, 8], batch=8, auto=False, alpha=0.2), n_fill=50, K=3)
    for act in ("relu", "tanh", "elu", "leaky_relu", "gelu", "selu", "identity"):
        record_run(agent_mod, f"acts_{act}", 5, 3,
                   make_config([12, 10, 6], act=act, batch=16, auto=True, q_hidden=[10, 12, 7],
                               log_std_min=-5, log_std_max=0.5, action_scale=2.0),
                   n_fill=64, K=2)
    record_run(agent_mod, "outact_tanh", 4, 1, make_config([8], act="elu", out_act="tanh", batch=8, auto=True),
               n_fill=40, K=2)
    record_run(agent_mod, "bipedal", 24, 4, make_config([256, 256], batch=256, auto=True, capacity=5000),
               n_fill=3000, K=3, full_state=False)
    record_run(agent_mod, "pendulum128", 4, 1, make_config([128, 128], batch=256, auto=True, capacity=2000),
               n_fill=2500, K=2, full_state=False)
    record_round2(agent_mod, a)


SHIPPED_CKPT = REF + "/notebooks/runs/InvertedPendulum-v5/SAC/sac-inverted-pendulum-2025_11_30-11_40_35/sac_agent.pth"


def record_round2(agent_mod, tiny_agent):
    """Round 2: the BASELINE configurations 3, 4, 5 themselves, a K = 10 free run, and runs that start
    from reference-WRITTEN checkpoints (agent.py:521-536) in both ``log_alpha`` forms (0-dim float64 as written today;
    (1,) float32 as in the shipped files)."""
    # cfg 3: InvertedPendulum shape of the Optuna study (hparam_search/configs/inverted_pendulum.yaml): 4/1/2x256, batch 256
    record_run(agent_mod, "cfg3_pendulum256", 4, 1, make_config([256, 256], batch=256, auto=True, capacity=2000),
               n_fill=2500, K=3, full_state=False, store_init=False)
    # cfg 4: Donkey latent 32/2/2x256, batch 1024, 50k ring; its shipped network [256,256,32] elu
    # (notebooks/configs/donkey_car_new.yaml:15-16,8-11: fixed alpha, tau 0.02, lr 4e-4); real observation width 216 (F11)
    record_run(agent_mod, "cfg4_donkey", 32, 2, make_config([256, 256], batch=1024, auto=True, capacity=50000),
               n_fill=3000, K=2, full_state=False, store_init=False)
    cfg = make_config([256, 256, 32], act="elu", batch=1024, auto=False, alpha=0.1, capacity=50000, seed=23)
    cfg["sac"].update(tau=0.02, actor_lr=4e-4, critic_lr=4e-4, alpha_lr=4e-4)
    record_run(agent_mod, "cfg4_donkey_elu", 32, 2, cfg, n_fill=3000, K=2, full_state=False, store_init=False)
    record_run(agent_mod, "cfg4_donkey_obs216", 216, 2, make_config([256, 256], batch=1024, auto=True, capacity=50000),
               n_fill=2000, K=1, full_state=False, store_init=False)
    # cfg 5: BipedalWalker shape at a per-rank slice (2048 rows) and at the full global batch (65536 rows, one update)
    record_run(agent_mod, "cfg5_b2048", 24, 4, make_config([256, 256], batch=2048, auto=True, capacity=5000),
               n_fill=5000, K=2, full_state=False, store_init=False)
    record_run(agent_mod, "cfg5_b65536", 24, 4, make_config([256, 256], batch=65536, auto=True, capacity=100000),
               n_fill=100000 - 3, K=1, full_state=False, big=True, store_init=False)
    # SURVEY 8c protocol: K = 10 free-running updates at BipedalWalker shape
    record_run(agent_mod, "bipedal_k10", 24, 4, make_config([256, 256], batch=256, auto=True, capacity=5000),
               n_fill=3000, K=10, full_state=False, store_init=False)
    # reference-written checkpoints and runs continued from them
    import torch
    p = os.path.join(HERE, "ref_ckpt_tiny_auto.pth")
    tiny_agent.save_agent(p)                                    # after the 3 recorded updates of tiny_auto
    print("wrote", os.path.basename(p))
    record_run(agent_mod, "ckpt_tiny_auto", 3, 2, make_config([8, 8], batch=8, auto=True), n_fill=50, K=2, start_ckpt=p)
    cfgp = make_config([128, 128], batch=256, auto=True, capacity=2000)
    env = FakeEnv(4, 1)
    ag = agent_mod.SAC(env, copy.deepcopy(cfgp))
    ag.load_agent(SHIPPED_CKPT)                                  # shipped: Adam step 39886, log_alpha (1,) float32
    p2 = os.path.join(HERE, "ref_ckpt_pendulum128_auto.pth")
    ag.save_agent(p2)                                            # re-written by the reference's own save_agent
    ck = torch.load(p2, map_location="cpu", weights_only=False)
    assert tuple(ck["log_alpha"].shape) == (1,) and ck["log_alpha"].dtype == torch.float32
    print("wrote", os.path.basename(p2))
    record_run(agent_mod, "ckpt_pendulum128", 4, 1, cfgp, n_fill=2500, K=3, full_state=False, start_ckpt=p2)


if __name__ == "__main__":
    main()
