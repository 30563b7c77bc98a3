"""Episode logger callback for Stable-Baselines3 runs (reference: sac/utils/stable_baseline_logger.py:7-73; used by the
comparison notebooks). Needs stable_baselines3, like the notebooks that import it."""
import os

from stable_baselines3.common.callbacks import BaseCallback

from sac.utils.logger_utils import save_lengths, save_rewards


class EpisodeLoggerSB3(BaseCallback):
    """Accumulates reward / length per episode from SB3's ``locals``; writes ``Episode/Reward`` and ``Episode/Length``; stops
    training (returns False) after ``max_episodes`` and dumps the two curves as .npy."""

    def __init__(self, writer, max_episodes: int, save_dir: str = "", save_npy: bool = True, verbose: int = 0):
        super().__init__(verbose)
        self.writer, self.max_episodes, self.save_dir, self.save_npy = writer, max_episodes, save_dir, save_npy
        self.current_episode_reward, self.current_episode_length = 0.0, 0
        self.episode_rewards, self.episode_lengths = [], []
        self.episode_count = 0

    def _on_step(self):
        for reward, done in zip(self.locals["rewards"], self.locals["dones"]):
            self.current_episode_reward += reward
            self.current_episode_length += 1
            if not done:
                continue
            ep_r, ep_l = float(self.current_episode_reward), int(self.current_episode_length)
            self.writer.add_scalar("Episode/Reward", ep_r, self.episode_count)
            self.writer.add_scalar("Episode/Length", ep_l, self.episode_count)
            if self.verbose:
                print(f"[Episode {self.episode_count}] Reward={ep_r}, Length={ep_l}")
            self.episode_rewards.append(ep_r)
            self.episode_lengths.append(ep_l)
            self.current_episode_reward, self.current_episode_length = 0.0, 0
            self.episode_count += 1
            if self.episode_count >= self.max_episodes:
                if self.save_npy:
                    os.makedirs(self.save_dir, exist_ok=True)
                    save_rewards(self.save_dir, self.episode_rewards)
                    save_lengths(self.save_dir, self.episode_lengths)
                if self.verbose:
                    print(f"Saved episode rewards and lengths to {self.save_dir}")
                print("Reached max episodes → early stopping training.")
                return False
        return True
