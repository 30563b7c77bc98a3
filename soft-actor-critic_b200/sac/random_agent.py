"""Uniform-random baseline rollouts (reference: sac/random_agent.py:5-28; imported by the notebooks). A caller of the
environments only -- nothing here touches the update engine."""
import numpy as np

try:
    from tqdm import tqdm
except Exception:  # pragma: no cover
    def tqdm(it, **kw):
        return it


def random_agent_loop(env, num_episodes, writer, seed):
    """``num_episodes`` episodes of ``env.action_space.sample()``; per-episode return to ``writer`` (a SummaryWriter or None)
    under ``RandomAgent/Reward``; every tenth episode printed, as the reference does."""
    if seed is not None:
        np.random.seed(seed)
        env.reset(seed=seed)
        env.action_space.seed(seed)
    for episode in tqdm(range(num_episodes)):
        env.reset()
        done, episode_reward = False, 0
        while not done:
            _, reward, terminated, truncated, _ = env.step(env.action_space.sample())
            done = terminated or truncated
            episode_reward += reward
        if writer is not None:
            writer.add_scalar("RandomAgent/Reward", episode_reward, episode)
        if episode % 10 == 0:
            print(f"Episode {episode}: avg batch reward = {episode_reward:.3f}")
