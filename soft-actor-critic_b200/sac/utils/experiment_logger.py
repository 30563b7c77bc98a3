"""Minimal logger with the interface ``SAC`` expects (reference: sac/utils/experiment_logger.py:16-148):
``run_dir``, ``episode_rewards`` / ``episode_lengths``, ``log_episode_metrics``, ``log_q_values``,
``log_hparams``, ``flush`` / ``close``.  TensorBoard writers are used when the package is importable and
silently skipped otherwise; nothing here touches the update path."""
from __future__ import annotations

from datetime import datetime
from pathlib import Path
from typing import Any, Dict, Optional


def _writer(path: str, flush_secs: int, suffix: str):
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(path, flush_secs=flush_secs, filename_suffix=suffix)
    except Exception:
        return None


class ExperimentLogger:
    def __init__(self, cfg: Dict[str, Any], run_name: Optional[str] = None, env_name: Optional[str] = None,
                 agent_name: Optional[str] = None):
        self.cfg = cfg
        self.env_name = env_name or cfg["env_name"] or "Environment"
        self.agent_name = agent_name or cfg["agent_name"] or "Agent"
        name = run_name or cfg["run_name"] or "sac"
        if cfg["use_timestamp"]:
            name = f"{name}-{datetime.now().strftime(cfg['timestamp_format'])}"
        self.run_id = name
        self.run_dir = Path(cfg["log_dir"]) / self.env_name / self.agent_name / name
        self.run_dir.mkdir(parents=True, exist_ok=True)
        self.metrics_writer = _writer(self.run_dir.as_posix(), cfg["flush_secs"], "_metrics")
        self.hparams_writer = _writer(self.run_dir.as_posix(), cfg["flush_secs"], "_hparams")
        self._hparams_logged = False
        self.episode_rewards, self.episode_lengths = [], []
        self.q1_values, self.q2_values = [], []

    def log_episode_metrics(self, episode_idx: int, reward: float, length: int) -> None:
        if not self.cfg["log_episode_stats"]:
            return
        if self.metrics_writer is not None:
            self.metrics_writer.add_scalar("Episode/Reward", reward, episode_idx)
            self.metrics_writer.add_scalar("Episode/Length", length, episode_idx)
        self.episode_rewards.append(reward)
        self.episode_lengths.append(length)

    def log_q_values(self, q1_value: float, q2_value: float, step: int) -> None:
        if not self.cfg["log_q_values"]:
            return
        if self.metrics_writer is not None:
            self.metrics_writer.add_scalar("QValues/Q1", q1_value, step)
            self.metrics_writer.add_scalar("QValues/Q2", q2_value, step)
        self.q1_values.append(q1_value)
        self.q2_values.append(q2_value)

    def log_hparams(self, hparams: Dict[str, Any], metrics: Dict[str, float]) -> None:
        if self._hparams_logged:
            return
        flat: Dict[str, Any] = {}

        def walk(prefix, node):
            if isinstance(node, dict):
                for k, v in node.items():
                    walk(f"{prefix}/{k}" if prefix else k, v)
            else:
                flat[prefix] = node if isinstance(node, (int, float, bool)) else str(node)

        walk("", hparams)
        if self.hparams_writer is not None:
            self.hparams_writer.add_hparams(flat, {k: float(v) for k, v in metrics.items()} or {"placeholder_metric": 0.0})
        self._hparams_logged = True

    def flush(self) -> None:
        for w in (self.metrics_writer, self.hparams_writer):
            if w is not None:
                w.flush()

    def close(self) -> None:
        self.flush()
        for w in (self.metrics_writer, self.hparams_writer):
            if w is not None:
                w.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc) -> None:
        self.close()
