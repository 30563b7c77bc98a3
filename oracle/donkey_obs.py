"""Oracle (test infrastructure only): NumPy restatement of the observation bookkeeping of the reference's DonkeyVae
environment -- DonkeyCarEnv/donkey_gym/envs/vae_env.py:66-71 (state), :187-193 (command-history roll + concatenation),
:201-208 (frame stack, zeroed at episode end), :253-266 (reset). The simulator and the VAE are not involved: `latent` stands for
``self.vae.encode(observation)`` (shape [1, z_size]). Checked against the real class by tests/test_oracle_donkey.py when the
reference checkout is present."""
import numpy as np


class DonkeyObsOracle:
    def __init__(self, z_size=32, n_commands=2, n_command_history=20, n_stack=3):
        self.z_size, self.n_commands, self.n_command_history, self.n_stack = z_size, n_commands, n_command_history, n_stack
        self.command_history = np.zeros((1, n_commands * n_command_history))
        self.stacked_obs = np.zeros((1, n_stack * (z_size + n_commands * n_command_history)), np.float32) if n_stack > 1 else None

    def reset(self, latent):
        self.command_history = np.zeros((1, self.n_commands * self.n_command_history))
        observation = np.asarray(latent).reshape(1, -1)
        if self.n_command_history > 0:
            observation = np.concatenate((observation, self.command_history), axis=-1)
        if self.n_stack > 1:
            self.stacked_obs[...] = 0
            self.stacked_obs[..., -observation.shape[-1]:] = observation
            return self.stacked_obs
        return observation

    def step(self, latent, action, done):
        observation = np.asarray(latent).reshape(1, -1)
        if self.n_command_history > 0:
            self.command_history = np.roll(self.command_history, shift=-self.n_commands, axis=-1)
            self.command_history[..., -self.n_commands:] = action
            observation = np.concatenate((observation, self.command_history), axis=-1)
        if self.n_stack > 1:
            self.stacked_obs = np.roll(self.stacked_obs, shift=-observation.shape[-1], axis=-1)
            if done:
                self.stacked_obs[...] = 0
            self.stacked_obs[..., -observation.shape[-1]:] = observation
            return self.stacked_obs
        return observation
