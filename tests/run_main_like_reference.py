"""Harness restatement of the reference's CLI (main.py:10-51), for the GPU box where /root/reference does not exist: load the
YAML, decode string-encoded hidden_sizes, build the environment named by logger.env_name, construct SAC(env, config), run
run_training_loop(num_episodes) and print the line the Optuna driver scrapes (run_search.py:74-80)."""
import argparse
import json

import yaml

import gymnasium as gym
from sac.agent import SAC
from sac.envs import *  # noqa: F401,F403


def main(args):
    with open(args.config) as f:
        config = yaml.safe_load(f)
    for net in ("q_net", "policy_net"):
        if net in config and isinstance(config[net].get("hidden_sizes"), str):
            config[net]["hidden_sizes"] = json.loads(config[net]["hidden_sizes"])
    print("Configuration loaded:")
    print(config)
    probes = {"ConstantRewardEnv": ConstantRewardEnv, "QuadraticActionRewardEnv": QuadraticActionRewardEnv,      # noqa: F405
              "RandomObsBinaryRewardEnv": RandomObsBinaryRewardEnv, "OneDPointMassReachEnv": OneDPointMassReachEnv}   # noqa: F405
    name = config["logger"]["env_name"]
    env = probes[name]() if name in probes else gym.make(name, max_episode_steps=config["train"].get("max_episode_steps", 1000))
    agent = SAC(env, config)
    print("Agent initialized. Starting training...")
    metrics = agent.run_training_loop(num_episodes=config["train"].get("num_episodes", 1000))
    print(f"Final average return: {metrics['final_avg_return']}")
    print("Training finished.")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=str, default="configs/example_config_env.yaml")
    main(ap.parse_args())
