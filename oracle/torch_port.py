"""Oracle (test infrastructure only): torch-eager CPU restatement of the reference update.

Purpose: (1) the CPU baseline that ``bench.py`` times on the GPU box's host
cores (``cpu_baseline.kind = "port"``; /root/reference cannot travel to the
box), doing the same work the reference does per ``training_step`` --
``random.sample`` on a ``deque`` of namedtuples, numpy stacking, eager torch
forward, autograd backward, ``torch.optim.Adam``, per-tensor Polyak
(/root/reference/sac/agent.py:166-327, sac/replay_buffer.py:30-39,
sac/models.py:30-33,73-87); (2) a second, autograd-based witness for the
hand-derived gradients of ``sac_numpy``.

It is pinned bit-for-bit against the real reference by
tests/test_oracle_golden.py using vectors recorded by
tests/golden/make_golden.py (same torch build, same op sequence => identical
bits).

Written functionally (flat tensor lists + F.linear) rather than as nn.Modules;
weights use the nn.Linear [out, in] layout and the reference's init recipe
 is reproduced in ``init_like_reference``.
"""
from __future__ import annotations

import math
import random
from collections import deque, namedtuple
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Row = namedtuple("Row", "s a r s2 d")

_ACT = {
    "relu": F.relu,
    "tanh": torch.tanh,
    "elu": F.elu,
    "leaky_relu": F.leaky_relu,
    "gelu": F.gelu,
    "selu": F.selu,
    "identity": lambda t: t,
}


def _init_net(sizes: Sequence[int], seed: int) -> List[torch.Tensor]:
    """nn.Linear default init (consumes the RNG stream) then xavier_uniform_ / zeros,
    after torch.manual_seed(seed) -- sac/models.py:19-28,38-42 (F10)."""
    torch.manual_seed(seed)
    layers = [torch.nn.Linear(sizes[i], sizes[i + 1]) for i in range(len(sizes) - 1)]
    out: List[torch.Tensor] = []
    for lin in layers:
        torch.nn.init.xavier_uniform_(lin.weight)
        torch.nn.init.zeros_(lin.bias)
        out += [lin.weight.detach().clone().requires_grad_(True), lin.bias.detach().clone().requires_grad_(True)]
    return out


def _run(params: List[torch.Tensor], x: torch.Tensor, hidden: str, out_act: str) -> torch.Tensor:
    n = len(params) // 2
    for l in range(n):
        x = F.linear(x, params[2 * l], params[2 * l + 1])
        x = _ACT[hidden if l < n - 1 else out_act](x)
    return x


class TorchPortSAC:
    """Same update, same order, same dtypes as the reference agent (F4-F7)."""

    def __init__(self, obs: int, act: int, cfg: dict, capacity: int = 1_000_000):
        self.obs, self.act, self.cfg = obs, act, cfg
        sac, tr = cfg["sac"], cfg["train"]
        qh, ph = list(cfg["q_net"]["hidden_sizes"]), list(cfg["policy_net"]["hidden_sizes"])
        seed = tr["seed"]
        self.pi = _init_net([obs] + ph + [2 * act], seed)
        self.q1 = _init_net([obs + act] + qh + [1], seed)
        self.q2 = _init_net([obs + act] + qh + [1], seed + 1)
        self.q1t = [p.detach().clone() for p in self.q1]
        self.q2t = [p.detach().clone() for p in self.q2]
        self.opt_pi = torch.optim.Adam(self.pi, lr=sac["actor_lr"])
        self.opt_q1 = torch.optim.Adam(self.q1, lr=sac["critic_lr"])
        self.opt_q2 = torch.optim.Adam(self.q2, lr=sac["critic_lr"])
        np.random.seed(seed)
        torch.manual_seed(seed)
        random.seed(seed)
        self.target_entropy = -float(act)
        self.auto = bool(sac["auto_entropy_tuning"])
        if self.auto:
            self.log_alpha = torch.tensor(np.log(sac["alpha"]), requires_grad=True)   # 0-dim float64 (F6)
            self.alpha = torch.exp(self.log_alpha).detach()
            self.opt_alpha = torch.optim.Adam([self.log_alpha], lr=sac["alpha_lr"])
        else:
            self.alpha = torch.tensor(sac["alpha"])
        self.memory: deque = deque(maxlen=capacity)
        self.last: Dict[str, torch.Tensor] = {}
        self.hook_after_critics = None          # tests: observe the state between the critic steps and the actor step

    # -- nets --------------------------------------------------------------------
    def _q(self, params, s, a):
        c = self.cfg["q_net"]
        return _run(params, torch.cat([s, a], dim=-1), c["hidden_layers_act"], c["output_activation"]).squeeze(-1)

    def sample_action(self, s: torch.Tensor, eps: torch.Tensor | None = None):
        c = self.cfg["policy_net"]
        head = _run(self.pi, s, c["hidden_layers_act"], c["output_activation"])
        mu, log_std = torch.chunk(head, 2, dim=-1)
        log_std = torch.clamp(log_std, c["log_std_min"], c["log_std_max"])
        std = log_std.exp()
        if eps is None:
            eps = torch.empty(mu.shape).normal_()                  # what Normal.rsample draws
        z = mu + eps * std
        action = torch.tanh(z) * c["action_scale"]
        var = std ** 2
        lp = (-((z - mu) ** 2) / (2 * var) - std.log() - math.log(math.sqrt(2 * math.pi))).sum(-1)
        lp = lp - (2 * (np.log(2) - z - F.softplus(-2 * z))).sum(-1)
        return action, lp

    # -- replay (host AoS deque, as in sac/replay_buffer.py) -----------------------
    def push(self, s, a, r, s2, d):
        self.memory.append(Row(s, a, r, s2, d))

    def sample_batch(self, batch_size: int):
        rows = random.sample(self.memory, batch_size)
        cols = Row(*zip(*rows))
        f32 = torch.float32
        return (
            torch.as_tensor(np.stack(cols.s), dtype=f32),
            torch.as_tensor(np.stack(cols.a), dtype=f32),
            torch.as_tensor(np.array(cols.r), dtype=f32),
            torch.as_tensor(np.stack(cols.s2), dtype=f32),
            torch.as_tensor(np.array(cols.d), dtype=f32),
        )

    # -- the update ----------------------------------------------------------------
    def update_from_batch(self, s, a, r, s2, d, eps1=None, eps2=None):
        gamma, tau = self.cfg["sac"]["gamma"], self.cfg["sac"]["tau"]
        with torch.no_grad():
            alpha = self.alpha.detach()
            a2, lp2 = self.sample_action(s2, eps1)
            minq = torch.min(self._q(self.q1t, s2, a2), self._q(self.q2t, s2, a2))
            y = r + gamma * (1 - d) * (minq - alpha * lp2)
        q1v, q2v = self._q(self.q1, s, a), self._q(self.q2, s, a)
        l1 = F.mse_loss(q1v, y)
        l2 = F.mse_loss(q2v, y)
        self.opt_q1.zero_grad()
        l1.backward()
        self.opt_q1.step()
        self.opt_q2.zero_grad()
        l2.backward()
        self.opt_q2.step()
        if self.hook_after_critics is not None:
            self.hook_after_critics(self)
        an, lp = self.sample_action(s, eps2)
        minq_pi = torch.min(self._q(self.q1, s, an), self._q(self.q2, s, an))
        lpi = (self.alpha.detach() * lp - minq_pi).mean()
        self.opt_pi.zero_grad()
        lpi.backward()
        self.opt_pi.step()
        info = {}
        if self.auto:
            la = -(self.log_alpha * (lp + self.target_entropy).detach()).mean()
            self.opt_alpha.zero_grad()
            la.backward()
            self.opt_alpha.step()
            self.alpha = self.log_alpha.exp()
            info = {"alpha_loss": la.item(), "alpha": self.alpha.item()}
        with torch.no_grad():
            for tgt, src in ((self.q1t, self.q1), (self.q2t, self.q2)):
                for t, p in zip(tgt, src):
                    t.copy_(tau * p.data + (1.0 - tau) * t)
        self.last = {"y": y, "q1_loss": l1.detach(), "q2_loss": l2.detach(), "policy_loss": lpi.detach(), "lp": lp.detach(),
                     "q1": q1v.detach(), "q2": q2v.detach(), "q1_pi": None, "q2_pi": None}
        return info

    def training_step(self):
        """Comparator (A): full step including the deque sample."""
        batch = self.sample_batch(self.cfg["train"]["batch_size"])
        return self.update_from_batch(*batch)

    # -- checkpoint (sac/agent.py:538-554), statement by statement ------------------------
    def load_checkpoint(self, ck: dict) -> None:
        """What the reference's ``load_agent`` does with a ``save_agent`` dict. Note the last branch: the reference REBINDS
        ``self.log_alpha`` to the loaded tensor while ``alpha_optimizer`` keeps the tensor created in ``__init__`` as its
        parameter, so after a load the temperature optimiser steps a tensor nobody reads and ``alpha`` stays at
        ``exp(loaded log_alpha)`` (in the loaded tensor's dtype/shape: ``(1,) float32`` in the shipped files). Restated
        literally so that the port stays bit-exact to runs the reference continued from a checkpoint."""
        with torch.no_grad():
            for ps, key in ((self.pi, "policy_net_state_dict"), (self.q1, "q_net1_state_dict"), (self.q2, "q_net2_state_dict"),
                            (self.q1t, "q_net1_target_state_dict"), (self.q2t, "q_net2_target_state_dict")):
                sd = ck[key]
                for l in range(len(ps) // 2):
                    ps[2 * l].copy_(sd[f"net.{2 * l}.weight"])
                    ps[2 * l + 1].copy_(sd[f"net.{2 * l}.bias"])
        self.opt_pi.load_state_dict(ck["policy_optimizer_state_dict"])
        self.opt_q1.load_state_dict(ck["q1_optimizer_state_dict"])
        self.opt_q2.load_state_dict(ck["q2_optimizer_state_dict"])
        if self.auto:
            self.log_alpha = ck["log_alpha"]
            self.opt_alpha.load_state_dict(ck["alpha_optimizer_state_dict"])
            self.alpha = self.log_alpha.exp()

    # -- state access for the pin tests ---------------------------------------------
    def flat_state(self) -> Dict[str, np.ndarray]:
        out = {}
        for name, ps in (("pi", self.pi), ("q1", self.q1), ("q2", self.q2), ("q1t", self.q1t), ("q2t", self.q2t)):
            for i, p in enumerate(ps):
                out[f"{name}.{i}"] = p.detach().numpy().copy()
        if self.auto:
            out["log_alpha"] = self.log_alpha.detach().numpy().copy()
        return out
