"""YAML config -> keyword arguments of Stable-Baselines3's SAC (reference: sac/utils/stable_baseline_params.py:16-60; used by
the comparison notebooks). Pure dictionary work: stable_baselines3 itself is only needed by the caller."""
from datetime import datetime
from typing import Any, Dict

import torch

activation_lookup = {"relu": torch.nn.ReLU, "tanh": torch.nn.Tanh, "elu": torch.nn.ELU, "leaky_relu": torch.nn.LeakyReLU,
                     "gelu": torch.nn.GELU, "selu": torch.nn.SELU, "identity": torch.nn.Identity}


def get_sb3_sac_params(env, config: Dict[str, Any], seed: int, env_id: str = "") -> Dict[str, Any]:
    sac, tr, lg = config["sac"], config["train"], config["logger"]
    stamp = datetime.now().strftime(lg["timestamp_format"])
    # (computed as the reference does; SB3's own tensorboard logging stays off there as well)
    _tensorboard_log = f"{lg.get('log_dir', 'runs')}/{env_id}/{lg['agent_name']}_sb3/{lg['run_name']}-{stamp}"
    return {
        "policy": "MlpPolicy",
        "env": env,
        "learning_rate": sac["actor_lr"],                 # SB3 has one learning rate for all networks
        "buffer_size": config["buffer"]["capacity"],
        "learning_starts": tr["warming_steps"],
        "batch_size": tr["batch_size"],
        "tau": sac["tau"],
        "gamma": sac["gamma"],
        "train_freq": (1, "step"),
        "gradient_steps": tr["gradient_steps_per_update"],
        "ent_coef": "auto" if sac["auto_entropy_tuning"] else sac["alpha"],
        "target_entropy": -env.action_space.shape[0],
        "policy_kwargs": {
            "net_arch": {"pi": config["policy_net"]["hidden_sizes"], "qf": config["q_net"]["hidden_sizes"]},
            "activation_fn": activation_lookup.get(config["policy_net"].get("hidden_layers_act", "relu"), torch.nn.ReLU),
        },
        "device": tr["device"],
        "seed": seed,
    }
