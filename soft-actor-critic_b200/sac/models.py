"""Network containers for the B200 SAC engine.

Reference: /root/reference/sac/models.py -- ``QNetwork`` (:8-42), ``PolicyNetwork`` (:45-101),
``_ACTIVATIONS`` (:104-112), ``build_mlp`` (:115-149).

In this package the modules are *containers*, not the compute path: they give the engine its
initial weights (same torch calls in the same order as the reference, so the starting point is
bit-identical: ``torch.manual_seed`` -> ``nn.Linear`` default init -> ``xavier_uniform_`` / zero bias,
SURVEY F10), they keep the ``state_dict`` layout ``net.{0,2,4,..}.{weight,bias}`` of the reference's
checkpoints, and once an agent adopts them (``bind_to_views``) their parameters are zero-copy views
of the engine's packed device arena, so ``state_dict()`` / ``load_state_dict()`` / printing keep
working on live device state.  ``forward`` is provided for convenience (notebooks call the
modules directly); the training update never goes through it.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

_ACTIVATIONS = {
    "relu": nn.ReLU,
    "tanh": nn.Tanh,
    "elu": nn.ELU,
    "leaky_relu": nn.LeakyReLU,
    "gelu": nn.GELU,
    "selu": nn.SELU,
    "identity": nn.Identity,
}


def build_mlp(obs_size: int, hidden_sizes: List[int], action_size: int, hidden_activations: str = "relu",
              output_activation: str = "identity") -> nn.Sequential:
    """[Linear, act] x (len(hidden_sizes) + 1); Linear modules sit at even indices."""
    if not hidden_sizes:
        raise ValueError("hidden_sizes cannot be empty")
    hidden_cls, out_cls = _ACTIVATIONS[hidden_activations], _ACTIVATIONS[output_activation]   # KeyError if unknown
    widths = [obs_size, *hidden_sizes, action_size]
    mods: List[nn.Module] = []
    last = len(widths) - 2
    for i, (fan_in, fan_out) in enumerate(zip(widths[:-1], widths[1:])):
        mods.append(nn.Linear(fan_in, fan_out))
        mods.append(hidden_cls() if i < last else out_cls())
    return nn.Sequential(*mods)


class _EngineBackedMLP(nn.Module):
    """Shared behaviour: seeded construction, xavier init, binding to arena views."""

    def _construct(self, in_dim: int, hidden_sizes: Sequence[int], out_dim: int, hidden_act: str, out_act: str, seed) -> None:
        if seed is not None:
            torch.manual_seed(seed)
        self.net = build_mlp(in_dim, list(hidden_sizes), out_dim, hidden_act, out_act)
        for mod in self.modules():
            if isinstance(mod, nn.Linear):
                nn.init.xavier_uniform_(mod.weight)
                nn.init.zeros_(mod.bias)

    def linears(self) -> List[nn.Linear]:
        return [m for m in self.net if isinstance(m, nn.Linear)]

    def bind_to_views(self, views: Dict[str, torch.Tensor], tag: str, copy_in: bool = True) -> None:
        """Re-point every parameter at the engine's arena view ``{tag}.W{l}`` / ``{tag}.b{l}``."""
        with torch.no_grad():
            for l, lin in enumerate(self.linears()):
                w, b = views[f"{tag}.W{l}"], views[f"{tag}.b{l}"].reshape(-1)
                if copy_in:
                    w.copy_(lin.weight.detach().to(w.device))
                    b.copy_(lin.bias.detach().to(b.device))
                lin.weight.data = w
                lin.bias.data = b

    def save_weights(self, filepath) -> None:
        torch.save({k: v.detach().clone().cpu() for k, v in self.state_dict().items()}, filepath)


class QNetwork(_EngineBackedMLP):
    def __init__(self, obs_size, action_size, hidden_sizes, hidden_activations="relu", output_activation="identity", seed=None):
        super().__init__()
        self._construct(obs_size + action_size, hidden_sizes, 1, hidden_activations, output_activation, seed)

    def forward(self, state, action):
        return self.net(torch.cat([state, action], dim=-1)).squeeze(-1)


class PolicyNetwork(_EngineBackedMLP):
    def __init__(self, obs_size, action_size, hidden_sizes, log_std_min=-20, log_std_max=2, seed=None, action_scale=1.0,
                 hidden_activations="relu", output_activation="identity"):
        super().__init__()
        self.log_std_min, self.log_std_max, self.action_scale = log_std_min, log_std_max, action_scale
        self._construct(obs_size, hidden_sizes, 2 * action_size, hidden_activations, output_activation, seed)

    def forward(self, state):
        mu, log_std = self.net(state).chunk(2, dim=-1)
        return mu, log_std.clamp(self.log_std_min, self.log_std_max)

    def sample_action(self, state):
        """tanh-squashed Gaussian sample and its log-density (no -log(action_scale) term, as in the reference)."""
        mu, log_std = self.forward(state)
        std = log_std.exp()
        z = mu + torch.randn_like(mu) * std
        gauss = -((z - mu) ** 2) / (2 * std ** 2) - std.log() - math.log(math.sqrt(2 * math.pi))
        squash = 2 * (math.log(2.0) - z - F.softplus(-2 * z))
        return torch.tanh(z) * self.action_scale, (gauss - squash).sum(-1)

    def deterministic_action(self, state):
        mu, _ = self.forward(state)
        return torch.tanh(mu) * self.action_scale
