"""Helpers for the -m gpu parity tests: build an engine from a golden file / config, move state in and out."""
from __future__ import annotations

import numpy as np
import torch

from helpers import Golden, rel_l2, synth_transitions


def base_config(hidden=(256, 256), q_hidden=None, act="relu", out_act="identity", auto=True, batch=256, seed=0,
                capacity=5000, alpha=0.1, log_std_min=-20, log_std_max=2, action_scale=1.0, rng="host"):
    return {
        "sac": {"gamma": 0.99, "tau": 0.005, "alpha": alpha, "auto_entropy_tuning": auto,
                "actor_lr": 3e-4, "critic_lr": 3e-4, "alpha_lr": 3e-4},
        "q_net": {"hidden_sizes": list(q_hidden or hidden), "hidden_layers_act": act, "output_activation": out_act},
        "policy_net": {"hidden_sizes": list(hidden), "hidden_layers_act": act, "output_activation": out_act,
                       "log_std_min": log_std_min, "log_std_max": log_std_max, "action_scale": action_scale},
        "buffer": {"capacity": capacity},
        "train": {"gradient_steps_per_update": 1, "seed": seed, "batch_size": batch, "warming_steps": 10,
                  "device": "cuda", "rng": rng},
        "logger": {"enabled": False, "log_dir": "runs", "env_name": "Synthetic", "agent_name": "SAC", "run_name": "t",
                   "use_timestamp": False, "timestamp_format": "%Y", "flush_secs": 10, "log_episode_stats": False,
                   "log_q_values": False, "save_model": {"enabled": False, "path": None}},
    }


class _Space:
    def __init__(self, n):
        self.shape = (n,)

    def seed(self, s):
        return [s]

    def sample(self):
        return np.random.uniform(-1, 1, self.shape).astype(np.float32)


class FakeEnv:
    """Duck-typed env exposing exactly what SAC.__init__ touches (agent.py:32-33,122-124)."""
    spec = None

    def __init__(self, obs, act):
        self.observation_space = _Space(obs)
        self.action_space = _Space(act)

    def reset(self, seed=None):
        return np.zeros(self.observation_space.shape, np.float32), {}


def engine_from_golden(g: Golden, n_agents=1, **kw):
    from sac.engine import UpdateEngine

    cfg = dict(g.cfg)
    cfg["train"] = dict(cfg["train"], device="cuda")
    eng = UpdateEngine(g.obs, g.act, cfg, device="cuda", n_agents=n_agents, **kw)
    for ag in range(n_agents):
        load_nets(eng, {"pi": g.sd("init/pi"), "q1": g.sd("init/q1"), "q2": g.sd("init/q2")}, agent=ag)
    eng.reset_state()
    return eng


def load_nets(eng, sds, agent=0):
    """sds: {tag: state_dict with keys net.{2l}.weight/bias} -> arena views."""
    for tag, sd in sds.items():
        n = len(sd) // 2
        for l in range(n):
            eng.view(f"{tag}.W{l}", agent).copy_(torch.as_tensor(np.asarray(sd[f"net.{2 * l}.weight"])))
            eng.view(f"{tag}.b{l}", agent).reshape(-1).copy_(torch.as_tensor(np.asarray(sd[f"net.{2 * l}.bias"])))


def read_net(eng, tag, n_lin, prefix="", agent=0):
    out = {}
    for l in range(n_lin):
        out[f"net.{2 * l}.weight"] = eng.view(f"{prefix}{tag}.W{l}", agent).cpu().numpy().copy()
        out[f"net.{2 * l}.bias"] = eng.view(f"{prefix}{tag}.b{l}", agent).reshape(-1).cpu().numpy().copy()
    return out


def fill_ring(ring, n, obs, act, seed=0, agent=0):
    s, a, r, s2, d = synth_transitions(n, obs, act, seed)
    ring.push_batch(s, a, r, s2, d.astype(np.float32), agent=agent)
    return s, a, r, s2, d.astype(np.float32)


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x)).cuda()


def assert_close(name, got, ref, tol):
    e = rel_l2(got, ref)
    assert e < tol, f"{name}: rel-L2 {e:.3e} >= {tol:.1e}"
    return e


def set_engine_state(eng, st, agent=0):
    """Put a ReferenceRun.state() snapshot into the engine: parameters, targets, Adam moments and step counters, temperature."""
    load_nets(eng, {t: st[t] for t in ("pi", "q1", "q2", "q1t", "q2t")}, agent=agent)
    steps = [0, 0, 0, 0]
    for oi, tag in enumerate(("pi", "q1", "q2")):
        for nm, (m, v, step) in st["adam"][tag].items():
            l = int(nm.split(".")[1]) // 2
            wb = "W" if nm.endswith("weight") else "b"
            eng.view(f"m.{tag}.{wb}{l}", agent).reshape(m.shape).copy_(torch.as_tensor(m))
            eng.view(f"v.{tag}.{wb}{l}", agent).reshape(v.shape).copy_(torch.as_tensor(v))
            steps[oi] = int(step)
    if "log_alpha" in st:
        eng.view("scal.log_alpha", agent).fill_(st["log_alpha"])
        eng.view("scal.alpha_m", agent).fill_(st["adam_alpha"][0])
        eng.view("scal.alpha_v", agent).fill_(st["adam_alpha"][1])
        steps[3] = int(st["adam_alpha"][2])
    eng.view("scal.step", agent).copy_(torch.as_tensor(steps, dtype=torch.int64))
    eng.refresh_alpha()


def net_errs(eng, tag, ref_sd, prefix="", agent=0):
    """{tensor name: rel-L2 error} of one network block of the engine against a reference state_dict."""
    got = read_net(eng, tag, len(ref_sd) // 2, prefix=prefix, agent=agent)
    return {nm: rel_l2(got[nm], ref_sd[nm]) for nm in ref_sd}


def assert_net(eng, tag, ref_sd, tol, what, prefix="", agent=0):
    for nm, e in net_errs(eng, tag, ref_sd, prefix, agent).items():
        assert e < tol, f"{what} {prefix}{tag}.{nm}: rel-L2 {e:.3e} >= {tol:.1e}"
