"""B200-native drop-in for the reference's ``sac`` package (ignaschuemer7/soft-actor-critic).

Same import paths and class surface (``sac.agent.SAC``, ``sac.replay_buffer.ReplayBuffer`` /
``Transition``, ``sac.models.QNetwork`` / ``PolicyNetwork``); the update hot path runs in
hand-written sm_100a CUDA kernels behind the C ABI of ``include/sacx.h`` (libsacx.so).
"""
__all__ = ["agent", "models", "replay_buffer", "population"]
