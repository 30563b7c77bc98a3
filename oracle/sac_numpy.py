"""Oracle (test infrastructure only): NumPy restatement of one SAC gradient update.

Follows, function by function, the reference's hot path
(/root/reference/sac/agent.py:195-327, /root/reference/sac/models.py:30-33,73-149)
with the gradients written out by hand instead of autograd, so that the CUDA
kernels (which implement the same closed forms) can be compared phase by phase.
The third-party arithmetic it restates is PyTorch's (pinned torch==2.7.1 in the
reference's requirements.txt:16; 2.11.0 in this image): ``nn.Linear``,
the seven activations of ``_ACTIVATIONS`` (models.py:104-112),
``torch.distributions.Normal.rsample/log_prob``, ``F.softplus(beta=1,
threshold=20)``, ``F.mse_loss``, ``torch.min`` (ties split the gradient 1/2-1/2),
``torch.clamp`` (gradient passes on the closed interval) and
``torch.optim.Adam`` (``_single_tensor_adam``: lerp, addcmul, python-double bias
corrections, ``sqrt(v)/sqrt(bc2) + eps``).

``dtype=np.float32`` reproduces the reference's precision up to summation
order; ``dtype=np.float64`` gives the "truth" twin used to show the noise floor.
Pinned by tests/test_oracle_golden.py against vectors recorded from the real
reference (tests/golden/make_golden.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

try:  # erf for GELU; scipy is in the image, math.erf is the fallback
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)

ACTIVATIONS = ("relu", "tanh", "elu", "leaky_relu", "gelu", "selu", "identity")
_SELU_ALPHA = 1.6732632423543772848170429916717
_SELU_SCALE = 1.0507009873554804934193349852946
_LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


# ----------------------------------------------------------------------------
# activations: value from pre-activation z, derivative from (z, h)
# reference: sac/models.py:104-112 (-> torch.nn.{ReLU,Tanh,ELU,LeakyReLU,GELU,SELU,Identity})
# ----------------------------------------------------------------------------
def act_fwd(name: str, z: np.ndarray) -> np.ndarray:
    dt = z.dtype.type
    if name == "relu":
        return np.maximum(z, dt(0))
    if name == "tanh":
        return np.tanh(z)
    if name == "elu":
        return np.where(z > 0, z, np.expm1(np.minimum(z, dt(0))))
    if name == "leaky_relu":
        return np.where(z > 0, z, dt(0.01) * z)
    if name == "gelu":
        return (dt(0.5) * z * (dt(1) + _erf(z * dt(math.sqrt(0.5))))).astype(z.dtype)
    if name == "selu":
        return np.where(z > 0, dt(_SELU_SCALE) * z, dt(_SELU_SCALE * _SELU_ALPHA) * np.expm1(np.minimum(z, dt(0))))
    if name == "identity":
        return z
    raise KeyError(name)  # same error class as models.py:138-139


def act_bwd(name: str, z: np.ndarray, h: np.ndarray) -> np.ndarray:
    """d act / d z (torch backward formulas: threshold_backward, tanh_backward,
    elu_backward(is_result=False), leaky_relu_backward, gelu_backward('none'))."""
    dt = z.dtype.type
    if name == "relu":
        return (h > 0).astype(z.dtype)
    if name == "tanh":
        return dt(1) - h * h
    if name == "elu":
        return np.where(z > 0, dt(1), np.exp(np.minimum(z, dt(0))))
    if name == "leaky_relu":
        return np.where(z > 0, dt(1), dt(0.01))
    if name == "gelu":
        cdf = dt(0.5) * (dt(1) + _erf(z * dt(math.sqrt(0.5))))
        pdf = np.exp(dt(-0.5) * z * z) * dt(1.0 / math.sqrt(2 * math.pi))
        return (cdf + z * pdf).astype(z.dtype)
    if name == "selu":
        return np.where(z > 0, dt(_SELU_SCALE), dt(_SELU_SCALE * _SELU_ALPHA) * np.exp(np.minimum(z, dt(0))))
    if name == "identity":
        return np.ones_like(z)
    raise KeyError(name)


# ----------------------------------------------------------------------------
# MLP (reference: build_mlp, sac/models.py:115-149) -- weights in nn.Linear layout [out, in]
# ----------------------------------------------------------------------------
@dataclass
class MLP:
    W: List[np.ndarray]
    b: List[np.ndarray]
    hidden_act: str = "relu"
    out_act: str = "identity"

    @property
    def n_layers(self) -> int:
        return len(self.W)

    def copy(self) -> "MLP":
        return MLP([w.copy() for w in self.W], [x.copy() for x in self.b], self.hidden_act, self.out_act)

    def tensors(self) -> List[np.ndarray]:
        """Parameter order of ``nn.Module.parameters()``: W0, b0, W1, b1, ..."""
        out = []
        for w, x in zip(self.W, self.b):
            out += [w, x]
        return out

    def forward(self, x: np.ndarray) -> Tuple[np.ndarray, dict]:
        zs, hs = [], [x]
        h = x
        for l, (w, bias) in enumerate(zip(self.W, self.b)):
            z = h @ w.T + bias
            name = self.hidden_act if l < self.n_layers - 1 else self.out_act
            h = act_fwd(name, z)
            zs.append(z)
            hs.append(h)
        return h, {"z": zs, "h": hs}

    def backward(self, cache: dict, d_out: np.ndarray, need_dx: bool = False, need_dw: bool = True):
        """Return (dW list, db list, dx or None) for upstream gradient d_out [B, n_out]."""
        L = self.n_layers
        dW = [None] * L
        db = [None] * L
        delta = d_out * act_bwd(self.out_act, cache["z"][L - 1], cache["h"][L])
        dx = None
        for l in range(L - 1, -1, -1):
            if need_dw:
                dW[l] = delta.T @ cache["h"][l]
                db[l] = delta.sum(axis=0)
            if l > 0:
                delta = (delta @ self.W[l]) * act_bwd(self.hidden_act, cache["z"][l - 1], cache["h"][l])
            elif need_dx:
                dx = delta @ self.W[0]
        return dW, db, dx


def mlp_from_state_dict(sd: Dict[str, np.ndarray], hidden_act: str, out_act: str, dtype=np.float32) -> MLP:
    """state_dict keys ``net.{0,2,4,...}.{weight,bias}`` (sac/models.py:148-149: Linear, act interleaved)."""
    idx = sorted({int(k.split(".")[1]) for k in sd})
    return MLP(
        [np.asarray(sd[f"net.{i}.weight"], dtype=dtype).copy() for i in idx],
        [np.asarray(sd[f"net.{i}.bias"], dtype=dtype).copy() for i in idx],
        hidden_act,
        out_act,
    )


def mlp_to_state_dict(m: MLP) -> Dict[str, np.ndarray]:
    out = {}
    for l, (w, x) in enumerate(zip(m.W, m.b)):
        out[f"net.{2 * l}.weight"] = w
        out[f"net.{2 * l}.bias"] = x
    return out


# ----------------------------------------------------------------------------
# Adam (torch.optim.Adam defaults, _single_tensor_adam) -- sac/agent.py:105-115,51-53
# ----------------------------------------------------------------------------
@dataclass
class AdamState:
    lr: float
    m: List[np.ndarray]
    v: List[np.ndarray]
    step: int = 0
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8

    @staticmethod
    def zeros_like(tensors: List[np.ndarray], lr: float) -> "AdamState":
        return AdamState(lr, [np.zeros_like(t) for t in tensors], [np.zeros_like(t) for t in tensors])

    def apply(self, params: List[np.ndarray], grads: List[np.ndarray]) -> None:
        self.step += 1
        bc1 = 1.0 - self.beta1 ** self.step          # python doubles, as in torch
        bc2 = 1.0 - self.beta2 ** self.step
        step_size = self.lr / bc1
        bc2_sqrt = bc2 ** 0.5
        for p, g, m, v in zip(params, grads, self.m, self.v):
            dt = p.dtype.type
            m += dt(1.0 - self.beta1) * (g - m)                      # exp_avg.lerp_(grad, 1-beta1)
            v *= dt(self.beta2)
            v += dt(1.0 - self.beta2) * g * g                        # mul_(beta2).addcmul_(g, g, 1-beta2)
            denom = np.sqrt(v) / dt(bc2_sqrt) + dt(self.eps)
            p -= dt(step_size) * (m / denom)                         # addcdiv_(m, denom, value=-step_size)


# ----------------------------------------------------------------------------
# tanh-squashed Gaussian head (reference: PolicyNetwork.forward/sample_action, sac/models.py:73-87)
# ----------------------------------------------------------------------------
def softplus(x: np.ndarray) -> np.ndarray:
    """F.softplus(beta=1, threshold=20): linear above the threshold."""
    return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, x.dtype.type(20)))))


def squash_sample(head: np.ndarray, eps: np.ndarray, lo: float, hi: float, scale: float):
    dt = head.dtype.type
    A = head.shape[-1] // 2
    mu, ls_raw = head[..., :A], head[..., A:]
    ls = np.clip(ls_raw, dt(lo), dt(hi))
    std = np.exp(ls)
    z = mu + eps * std                                               # Normal.rsample: loc + eps*scale
    tz = np.tanh(z)
    action = tz * dt(scale)
    var = std * std
    lp = (-((z - mu) ** 2) / (dt(2) * var) - np.log(std) - dt(_LOG_SQRT_2PI)).sum(axis=-1)
    lp = lp - (dt(2) * (dt(math.log(2.0)) - z - softplus(dt(-2) * z))).sum(axis=-1)
    aux = {"tz": tz, "se": std * eps, "mask": ((ls_raw >= dt(lo)) & (ls_raw <= dt(hi))).astype(head.dtype), "z": z}
    return action, lp, aux


# ----------------------------------------------------------------------------
# The agent state and the update (reference: SAC, sac/agent.py)
# ----------------------------------------------------------------------------
@dataclass
class Hyper:
    gamma: float = 0.99
    tau: float = 0.005
    alpha: float = 0.1
    auto_entropy_tuning: bool = True
    actor_lr: float = 3e-4
    critic_lr: float = 3e-4
    alpha_lr: float = 3e-4
    log_std_min: float = -20.0
    log_std_max: float = 2.0
    action_scale: float = 1.0


@dataclass
class SACOracle:
    pi: MLP
    q1: MLP
    q2: MLP
    hp: Hyper
    q1t: Optional[MLP] = None
    q2t: Optional[MLP] = None
    dtype: type = np.float32
    opt_pi: AdamState = field(init=False)
    opt_q1: AdamState = field(init=False)
    opt_q2: AdamState = field(init=False)
    last: dict = field(default_factory=dict)

    def __post_init__(self):
        if self.q1t is None:
            self.q1t = self.q1.copy()                # deepcopy targets, sac/agent.py:88-89
        if self.q2t is None:
            self.q2t = self.q2.copy()
        self.opt_pi = AdamState.zeros_like(self.pi.tensors(), self.hp.actor_lr)
        self.opt_q1 = AdamState.zeros_like(self.q1.tensors(), self.hp.critic_lr)
        self.opt_q2 = AdamState.zeros_like(self.q2.tensors(), self.hp.critic_lr)
        self.act_dim = self.pi.W[-1].shape[0] // 2
        self.target_entropy = -float(self.act_dim)               # sac/agent.py:43
        # log_alpha is a 0-dim float64 tensor in the reference
        self.log_alpha = np.float64(np.log(self.hp.alpha))
        self.alpha = np.float64(np.exp(self.log_alpha)) if self.hp.auto_entropy_tuning else np.float64(np.float32(self.hp.alpha))
        self.a_m = np.float64(0.0)
        self.a_v = np.float64(0.0)
        self.a_step = 0

    # -- a6: compute_target_q_values, sac/agent.py:195-211 ---------------------
    def target(self, r, d, s2, eps1):
        dt = self.dtype
        head, _ = self.pi.forward(s2)
        a2, lp2, _ = squash_sample(head, eps1, self.hp.log_std_min, self.hp.log_std_max, self.hp.action_scale)
        x = np.concatenate([s2, a2], axis=-1)
        tq1 = self.q1t.forward(x)[0][:, 0]
        tq2 = self.q2t.forward(x)[0][:, 0]
        minq = np.minimum(tq1, tq2)
        alpha = dt(self.alpha)
        y = r + dt(self.hp.gamma) * (dt(1) - d) * (minq - alpha * lp2)
        self.last.update(a2=a2, lp2=lp2, tq1=tq1, tq2=tq2, y=y)
        return y

    # -- a7: update_q_networks, sac/agent.py:213-236 ----------------------------
    def critic_grads(self, s, a, y):
        dt = self.dtype
        B = s.shape[0]
        x = np.concatenate([s, a], axis=-1)
        out = {}
        for name, net in (("q1", self.q1), ("q2", self.q2)):
            q, cache = net.forward(x)
            q = q[:, 0]
            diff = q - y
            loss = np.mean(diff * diff, dtype=dt)                  # F.mse_loss
            d_out = (dt(2) * diff / dt(B))[:, None]
            dW, db, _ = net.backward(cache, d_out)
            out[name] = {"q": q, "loss": loss, "dW": dW, "db": db}
        return out

    def critic_step(self, s, a, y):
        g = self.critic_grads(s, a, y)
        for name, net, opt in (("q1", self.q1, self.opt_q1), ("q2", self.q2, self.opt_q2)):
            grads = []
            for w, x in zip(g[name]["dW"], g[name]["db"]):
                grads += [w, x]
            opt.apply(net.tensors(), grads)
        self.last.update(q1=g["q1"]["q"], q2=g["q2"]["q"], q1_loss=g["q1"]["loss"], q2_loss=g["q2"]["loss"], critic_grads=g)
        return g

    # -- a8: update_policy_network, sac/agent.py:238-260 -----------------------
    def actor_grads(self, s, eps2):
        dt = self.dtype
        B = s.shape[0]
        A = self.act_dim
        hp = self.hp
        head, pcache = self.pi.forward(s)
        a, lp, aux = squash_sample(head, eps2, hp.log_std_min, hp.log_std_max, hp.action_scale)
        x = np.concatenate([s, a], axis=-1)
        q1, c1 = self.q1.forward(x)
        q2, c2 = self.q2.forward(x)
        q1, q2 = q1[:, 0], q2[:, 0]
        minq = np.minimum(q1, q2)
        alpha = dt(self.alpha)
        loss = np.mean(alpha * lp - minq, dtype=dt)
        # d loss / d q_i : -1/B routed to the smaller critic, ties split 1/2-1/2 (torch.minimum backward)
        w1 = np.where(q1 < q2, dt(1), np.where(q1 == q2, dt(0.5), dt(0)))
        g1 = (-(w1) / dt(B))[:, None]
        g2 = (-(dt(1) - w1) / dt(B))[:, None]
        _, _, dx1 = self.q1.backward(c1, g1, need_dx=True, need_dw=False)
        _, _, dx2 = self.q2.backward(c2, g2, need_dx=True, need_dw=False)
        obs = s.shape[1]
        d_a = dx1[:, obs:] + dx2[:, obs:]                          # d loss / d action (through -minQ)
        tz, se, mask = aux["tz"], aux["se"], aux["mask"]
        # d loss/d z = (alpha/B) * dlogpi/dz + d_a * c (1 - tanh^2);  dlogpi/dz = 2 tanh z
        dz = (alpha / dt(B)) * (dt(2) * tz) + d_a * (dt(hp.action_scale) * (dt(1) - tz * tz))
        dmu = dz
        # d logpi / d logsigma = -1 (Gaussian term nets to -1) plus the z path: sigma*eps * dz
        dls = (se * dz - alpha / dt(B)) * mask
        d_head = np.concatenate([dmu, dls], axis=-1)
        dW, db, _ = self.pi.backward(pcache, d_head)
        return {"a": a, "lp": lp, "loss": loss, "q1": q1, "q2": q2, "d_a": d_a, "d_head": d_head, "dW": dW, "db": db}

    def actor_step(self, s, eps2):
        g = self.actor_grads(s, eps2)
        grads = []
        for w, x in zip(g["dW"], g["db"]):
            grads += [w, x]
        self.opt_pi.apply(self.pi.tensors(), grads)
        self.last.update(lp=g["lp"], policy_loss=g["loss"], actor_grads=g)
        return g["lp"]

    # -- a9: update_entropy_temperature, sac/agent.py:263-280 --------------------
    def alpha_step(self, lp):
        if not self.hp.auto_entropy_tuning:
            return {}
        dt = self.dtype
        t = lp + dt(self.target_entropy)
        alpha_loss = -np.mean(dt(self.log_alpha) * t, dtype=dt)
        grad = np.float64(-np.mean(t, dtype=dt))                   # f32 mean, f64 parameter
        self.a_step += 1
        b1, b2 = 0.9, 0.999
        self.a_m = self.a_m + (1.0 - b1) * (grad - self.a_m)
        self.a_v = self.a_v * b2 + (1.0 - b2) * grad * grad
        bc1 = 1.0 - b1 ** self.a_step
        bc2 = 1.0 - b2 ** self.a_step
        denom = math.sqrt(self.a_v) / (bc2 ** 0.5) + 1e-8
        self.log_alpha = np.float64(self.log_alpha - (self.hp.alpha_lr / bc1) * (self.a_m / denom))
        self.alpha = np.float64(np.exp(self.log_alpha))
        self.last.update(alpha_loss=float(alpha_loss), alpha=float(self.alpha))
        return {"alpha_loss": float(alpha_loss), "alpha": float(self.alpha)}

    # -- a10: soft_update_target_networks, sac/agent.py:282-300 ------------------
    def polyak(self):
        dt = self.dtype
        tau, omt = dt(self.hp.tau), dt(1.0 - self.hp.tau)
        for net, tgt in ((self.q1, self.q1t), (self.q2, self.q2t)):
            for p, t in zip(net.tensors(), tgt.tensors()):
                t[...] = tau * p + omt * t

    # -- a12: training_step order, sac/agent.py:302-327 --------------------------
    def update(self, s, a, r, s2, d, eps1, eps2):
        y = self.target(r, d, s2, eps1)
        self.critic_step(s, a, y)
        lp = self.actor_step(s, eps2)
        info = self.alpha_step(lp)
        self.polyak()
        return info

    # -- a13: select_action, sac/agent.py:149-156; models.py:89-92 ---------------
    def act(self, s, eps=None, deterministic=False):
        head, _ = self.pi.forward(s)
        A = self.act_dim
        if deterministic:
            return np.tanh(head[:, :A]) * self.dtype(self.hp.action_scale)
        a, _, _ = squash_sample(head, eps, self.hp.log_std_min, self.hp.log_std_max, self.hp.action_scale)
        return a

    def q_values(self, s, a):
        x = np.concatenate([s, a], axis=-1)
        return self.q1.forward(x)[0][:, 0], self.q2.forward(x)[0][:, 0]


def gather(ring: Dict[str, np.ndarray], logical_idx, pushes: int, capacity: int):
    """a2/a3: the sampled batch for a logical index stream (sac/agent.py:166-193)."""
    from .mt_sample import logical_to_slot

    slot = logical_to_slot(np.asarray(logical_idx), pushes, capacity)
    return tuple(ring[k][slot] for k in ("s", "a", "r", "s2", "d"))
