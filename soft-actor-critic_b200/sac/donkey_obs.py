"""Device-side observation assembly for the DonkeyVae producer.

The environment itself (simulator, VAE) stays the reference's. What moves is the NumPy bookkeeping between the VAE's latent and
the agent -- ``DonkeyVAEEnv.postprocessing_step`` / ``reset`` (DonkeyCarEnv/donkey_gym/envs/vae_env.py:175-210,253-266): the
command-history roll, the ``[latent | history]`` frame and the frame stack -- for a latent that is already a CUDA tensor
(ae/autoencoder.py:64-89 encodes on the device), together with the push of the transition into the replay ring, so that a step
of the producer costs two small launches and no host round trip of the 216-float observation.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _engine as E
from .replay_buffer import ReplayBuffer


class DeviceObservationAssembler:
    def __init__(self, z_size: int = 32, n_commands: int = 2, n_command_history: int = 20, n_stack: int = 3,
                 ring: Optional[ReplayBuffer] = None, agent: int = 0, device=None):
        E.require_cuda()
        self.lib = E.load()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.z_size, self.n_commands = int(z_size), int(n_commands)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            E.check(self.lib.sacx_obs_create(z_size, n_commands, n_command_history, n_stack, C.byref(h)))
        self.h = h
        self.obs_dim = int(self.lib.sacx_obs_dim(h))
        self.ring, self.agent = ring, agent
        if ring is not None and ring.handle is None:
            ring._allocate(self.obs_dim, self.n_commands)

    def __del__(self):
        try:
            if getattr(self, "h", None) is not None:
                self.lib.sacx_obs_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _latent(self, latent) -> torch.Tensor:
        t = latent if torch.is_tensor(latent) else torch.as_tensor(np.asarray(latent, np.float32))
        t = t.to(device=self.device, dtype=torch.float32).contiguous().reshape(-1)
        if t.numel() != self.z_size:
            raise ValueError(f"latent has {t.numel()} elements, expected {self.z_size}")
        return t

    def reset(self, latent) -> torch.Tensor:
        """env.reset(): zero history and stack, newest frame = [latent | 0]. Returns the stacked observation (device)."""
        z = self._latent(latent)
        out = torch.empty(self.obs_dim, dtype=torch.float32, device=self.device)
        E.check(self.lib.sacx_obs_reset(self.h, z.data_ptr(), out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream))
        return out

    def step(self, latent, action, reward: float, done: bool, push: bool = True) -> torch.Tensor:
        """The post-processing of env.step(): roll the command history and append `action`, stack the new frame (the stack is
        zeroed first when `done`), and -- with a ring -- store (previous stack, action, reward, new stack, done)."""
        z = self._latent(latent)
        out = torch.empty(self.obs_dim, dtype=torch.float32, device=self.device)
        ring = self.ring.handle if (push and self.ring is not None) else None
        if torch.is_tensor(action) and action.is_cuda:
            a = action.to(dtype=torch.float32).contiguous().reshape(-1)
            if a.numel() != self.n_commands:
                raise ValueError("action width does not match n_commands")
            a_dev, a_host, keep = a.data_ptr(), None, a
        else:
            keep = np.ascontiguousarray(np.asarray(action, np.float32).reshape(-1))
            if keep.size != self.n_commands:
                raise ValueError("action width does not match n_commands")
            a_dev, a_host = None, keep.ctypes.data
        E.check(self.lib.sacx_obs_step(self.h, z.data_ptr(), a_dev, a_host, float(reward), 1 if done else 0, ring, self.agent,
                                       out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream))
        return out
