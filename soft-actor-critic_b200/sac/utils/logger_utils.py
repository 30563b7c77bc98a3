"""Episode-curve persistence used by ``SAC.run_training_loop`` (reference: sac/utils/logger_utils.py:7-38).
Observability side-channel, host-only; the plotting helper of the reference is not part of the hot path and
is only provided when matplotlib is installed."""
from __future__ import annotations

from pathlib import Path
from typing import List

import numpy as np


def _dump(run_dir, name: str, values, dtype) -> None:
    d = Path(run_dir)
    d.mkdir(parents=True, exist_ok=True)
    np.save(d / name, np.asarray(values, dtype=dtype))


def save_rewards(run_dir, rewards: List[float]) -> None:
    _dump(run_dir, "episode_rewards.npy", rewards, np.float32)


def save_lengths(run_dir, lengths: List[int]) -> None:
    _dump(run_dir, "episode_lengths.npy", lengths, np.int32)


def load_rewards(run_dir) -> List[float]:
    return np.load(Path(run_dir) / "episode_rewards.npy").astype(float).tolist()


def load_lengths(run_dir) -> List[int]:
    return np.load(Path(run_dir) / "episode_lengths.npy").astype(int).tolist()


def make_and_save_graph(number_of_curves, data, title, xlabel, ylabel, filename, run_dir, legend=None) -> None:
    import matplotlib.pyplot as plt  # optional dependency

    plt.figure()
    for i in range(number_of_curves):
        plt.plot(data[i])
    plt.title(title)
    plt.xlabel(xlabel)
    plt.ylabel(ylabel)
    if legend:
        plt.legend(legend)
    plt.savefig(str(Path(run_dir) / filename))
    plt.close()
