"""Shared test helpers: golden loading, synthetic transitions, oracle construction."""
from __future__ import annotations

import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def synth_transitions(n, obs, act, seed=0):
    """s, s' ~ N(0,1); a ~ U(-1,1); r ~ N(0,1); d ~ Bernoulli(0.01); numpy default_rng(seed)."""
    rng = np.random.default_rng(seed)
    s = rng.standard_normal((n, obs)).astype(np.float32)
    s2 = rng.standard_normal((n, obs)).astype(np.float32)
    a = rng.uniform(-1, 1, (n, act)).astype(np.float32)
    r = rng.standard_normal(n).astype(np.float32)
    d = rng.random(n) < 0.01
    return s, a, r, s2, d


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
        self.meta = json.loads(bytes(self.z["meta"]).decode())
        self.cfg = self.meta["config"]
        self.obs, self.act, self.K = self.meta["obs"], self.meta["act"], self.meta["K"]
        self.n_fill = self.meta["n_fill"]
        self.full = self.meta["full_state"]

    def __getitem__(self, k):
        return self.z[k]

    def sd(self, prefix):
        p = prefix + "/"
        return {k[len(p):]: self.z[k] for k in self.z.files if k.startswith(p) and "#" not in k}

    def keys(self, prefix):
        p = prefix + "/"
        return [k for k in self.z.files if k.startswith(p)]


def numpy_oracle_from_golden(g: Golden, dtype=np.float32):
    from oracle.sac_numpy import SACOracle, Hyper, mlp_from_state_dict

    c = g.cfg
    hp = Hyper(gamma=c["sac"]["gamma"], tau=c["sac"]["tau"], alpha=c["sac"]["alpha"],
               auto_entropy_tuning=c["sac"]["auto_entropy_tuning"], actor_lr=c["sac"]["actor_lr"],
               critic_lr=c["sac"]["critic_lr"], alpha_lr=c["sac"]["alpha_lr"],
               log_std_min=c["policy_net"]["log_std_min"], log_std_max=c["policy_net"]["log_std_max"],
               action_scale=c["policy_net"]["action_scale"])
    pi = mlp_from_state_dict(g.sd("init/pi"), c["policy_net"]["hidden_layers_act"], c["policy_net"]["output_activation"], dtype)
    q1 = mlp_from_state_dict(g.sd("init/q1"), c["q_net"]["hidden_layers_act"], c["q_net"]["output_activation"], dtype)
    q2 = mlp_from_state_dict(g.sd("init/q2"), c["q_net"]["hidden_layers_act"], c["q_net"]["output_activation"], dtype)
    return SACOracle(pi, q1, q2, hp, dtype=dtype)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def max_rel(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / den) if den > 0 else float(np.abs(a - b).max())
