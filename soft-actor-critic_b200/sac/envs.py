"""The four probe environments ``main.py`` selects by name (``from sac.envs import *``; reference: sac/envs.py).

They are producers, not part of the update path; they are here so that the reference's ``main.py``, notebooks and Optuna
driver find the names they import. Same constructor arguments, spaces, dynamics, rewards and ``info`` keys as the
reference classes (file:line given per class); written once over a small common base instead of four times.
Needs ``gymnasium`` (the reference's dependency), like the callers that import this module.
"""
from typing import Optional

import gymnasium as gym
import numpy as np
from gymnasium import spaces

__all__ = ["ConstantRewardEnv", "QuadraticActionRewardEnv", "RandomObsBinaryRewardEnv", "OneDPointMassReachEnv"]


def _box(low, high, n=1):
    return spaces.Box(low=low, high=high, shape=(n,), dtype=np.float32)


class _ProbeEnv(gym.Env):
    """Episode bookkeeping shared by the probes: step counter, running return, ``info["episode"]`` on the last step."""

    def __init__(self, obs_dim: int, action_low: float, action_high: float, max_steps: int):
        super().__init__()
        self.max_steps = int(max_steps)
        self.action_space = _box(action_low, action_high)
        self.observation_space = _box(-np.inf, np.inf, obs_dim)
        self.current_step = 0
        self.episode_reward = 0.0

    def _observe(self) -> np.ndarray:
        return np.zeros(self.observation_space.shape, dtype=np.float32)

    def _begin(self) -> None:
        pass

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        super().reset(seed=seed)
        self.current_step = 0
        self.episode_reward = 0.0
        self._begin()
        return self._observe(), {}

    def _finish(self, reward: float, terminated: bool, truncated: bool, info: dict, count_reward: bool = True):
        if count_reward:
            self.episode_reward += reward
        if terminated or truncated:
            info["episode"] = {"r": self.episode_reward, "l": self.current_step}
        return self._observe(), reward, terminated, truncated, info


class ConstantRewardEnv(_ProbeEnv):
    """reference: sac/envs.py:15-46 -- the same reward at every step whatever the action; observation 0."""

    def __init__(self, reward: float = 1.0, max_steps: int = 1):
        super().__init__(1, -1.0, 1.0, max_steps)
        self.constant_reward = float(reward)

    def step(self, action):
        self.current_step += 1
        return self._finish(self.constant_reward, self.current_step >= self.max_steps, False, {})


class QuadraticActionRewardEnv(_ProbeEnv):
    """reference: sac/envs.py:57-99 -- continuous bandit, reward -(clip(a) - target)^2; observation 0."""

    def __init__(self, target: float = 0.5, action_low: float = -1.0, action_high: float = 1.0, max_steps: int = 1):
        super().__init__(1, action_low, action_high, max_steps)
        self.target = float(target)

    def step(self, action):
        self.current_step += 1
        a = np.clip(action[0], self.action_space.low[0], self.action_space.high[0])
        return self._finish(-((a - self.target) ** 2), self.current_step >= self.max_steps, False, {"action": a})


class RandomObsBinaryRewardEnv(_ProbeEnv):
    """reference: sac/envs.py:110-150 -- observations are uniform noise from the env's own generator; reward +1 inside
    |a| <= threshold, -1 outside. (As in the reference, ``episode_reward`` is reset but never accumulated here.)"""

    def __init__(self, obs_dim: int = 4, threshold: float = 0.2, max_steps: int = 1):
        super().__init__(int(obs_dim), -1.0, 1.0, max_steps)
        self.obs_dim = int(obs_dim)
        self.threshold = float(threshold)

    def _observe(self) -> np.ndarray:
        return self.np_random.uniform(low=-1.0, high=1.0, size=self.obs_dim).astype(np.float32)

    def step(self, action):
        self.current_step += 1
        a = float(action[0])
        return self._finish(1.0 if abs(a) <= self.threshold else -1.0, self.current_step >= self.max_steps, False,
                            {"action": a}, count_reward=False)


class OneDPointMassReachEnv(_ProbeEnv):
    """reference: sac/envs.py:161-222 -- x += clip(a) * dt; step penalty, bonus and termination within goal_tolerance of the
    goal; truncation after max_steps."""

    def __init__(self, start_pos: float = 0.0, goal_pos: float = 1.0, max_steps: int = 50, dt: float = 1.0,
                 action_low: float = -0.1, action_high: float = 0.1, step_penalty: float = -0.01, goal_reward: float = 1.0,
                 goal_tolerance: float = 0.05):
        super().__init__(1, action_low, action_high, max_steps)
        self.start_pos, self.goal_pos, self.dt = float(start_pos), float(goal_pos), float(dt)
        self.step_penalty, self.goal_reward, self.goal_tolerance = float(step_penalty), float(goal_reward), float(goal_tolerance)
        self.pos = 0.0

    def _begin(self) -> None:
        self.pos = self.start_pos

    def _observe(self) -> np.ndarray:
        return np.array([self.pos], dtype=np.float32)

    def step(self, action):
        self.current_step += 1
        a = float(np.clip(action[0], self.action_space.low[0], self.action_space.high[0]))
        self.pos += a * self.dt
        reached = abs(self.pos - self.goal_pos) <= self.goal_tolerance
        reward = self.step_penalty + (self.goal_reward if reached else 0.0)
        return self._finish(reward, reached, self.current_step >= self.max_steps, {"action": a})
