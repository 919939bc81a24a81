"""Experiment helpers (reference: /root/reference/fsrl/utils/exp_util.py): seeding, run names,
config/model loading.  Host-only."""
from __future__ import annotations

import os
import random
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch


def seed_all(seed=1029, others: Optional[list] = None) -> None:
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    if others is not None:
        for item in others:
            if hasattr(item, "seed"):
                item.seed(seed)


def load_config_and_model(path: str, best: bool = False) -> Tuple[dict, dict]:
    """``<path>/config.yaml`` + ``<path>/checkpoint/model[_best].pt`` (exp_util.py:60-84)."""
    import yaml
    if not os.path.exists(path):
        raise ValueError(f"{path} doesn't exist!")
    with open(os.path.join(path, "config.yaml")) as f:
        config = yaml.load(f.read(), Loader=yaml.FullLoader)
    name = "model_best.pt" if best else "model.pt"
    model = torch.load(os.path.join(path, "checkpoint", name), map_location="cpu", weights_only=False)
    return config, model


def to_string(values):
    if isinstance(values, (list, tuple)):
        return "-".join(to_string(v) for v in values)
    if isinstance(values, float):
        return f"{values:g}"
    return str(values)


DEFAULT_SKIP_KEY = ["task", "reward_threshold", "logdir", "worker", "project", "group", "name",
                    "prefix", "suffix", "save_interval", "render", "verbose", "save_ckpt",
                    "training_num", "testing_num", "epoch", "device", "thread"]
DEFAULT_KEY_ABBRE = {"cost_limit": "cost", "mstep_iter_num": "mnum", "estep_iter_num": "enum",
                     "estep_kl": "ekl", "mstep_kl_mu": "kl_mu", "mstep_kl_std": "kl_std",
                     "mstep_dual_lr": "mlr", "estep_dual_lr": "elr", "update_per_step": "update"}


def auto_name(default_cfg: dict, current_cfg: dict, prefix: str = "", suffix: str = "",
              skip_keys: Sequence[str] = DEFAULT_SKIP_KEY, key_abbre: Dict = DEFAULT_KEY_ABBRE) -> str:
    """Run name = the keys that differ from the default config (exp_util.py:131-169)."""
    parts = [prefix] if prefix else []
    for k in sorted(default_cfg.keys()):
        if k in skip_keys or k not in current_cfg:
            continue
        if default_cfg[k] != current_cfg[k]:
            parts.append(f"{key_abbre.get(k, k)}_{to_string(current_cfg[k])}")
    if suffix:
        parts.append(suffix)
    name = "-".join(parts) if parts else "default"
    return f"{name}-{str(hash(name))[-4:]}" if False else name
