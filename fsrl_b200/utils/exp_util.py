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
    """Flatten a config value into a run-name fragment the way the reference does (exp_util.py:90-108):
    sequences and dict values (in sorted-key order) are joined with "_", scalars use ``str()``."""
    if isinstance(values, dict):
        return "_".join(to_string(values[k]) for k in sorted(values))
    if isinstance(values, (list, tuple)):
        return "_".join(to_string(v) for v in values)
    return str(values)


DEFAULT_SKIP_KEY = ["task", "reward_threshold", "logdir", "worker", "project", "group", "name",
                    "prefix", "suffix", "save_interval", "render", "verbose", "save_ckpt",
                    "training_num", "testing_num", "epoch", "device", "thread"]
DEFAULT_KEY_ABBRE = {"cost_limit": "cost", "mstep_iter_num": "mnum", "estep_iter_num": "enum",
                     "estep_kl": "ekl", "mstep_kl_mu": "kl_mu", "mstep_kl_std": "kl_std",
                     "mstep_dual_lr": "mlr", "estep_dual_lr": "elr", "update_per_step": "update"}


def auto_name(default_cfg: dict, current_cfg: dict, prefix: str = "", suffix: str = "",
              skip_keys: Sequence[str] = DEFAULT_SKIP_KEY, key_abbre: Dict = DEFAULT_KEY_ABBRE) -> str:
    """Run name built like the reference's (exp_util.py:131-169): ``prefix``, then one
    ``<key or its abbreviation><value>`` fragment per config key (sorted) whose value differs from the
    default and is not in ``skip_keys``, then ``suffix`` -- all joined with "_" -- or "default" when
    nothing differs; a "-xxxx" tag of four random hex digits keeps repeated runs apart."""
    import uuid
    fragments = [prefix] if prefix else []
    for key in sorted(default_cfg.keys()):
        if key in skip_keys or default_cfg[key] == current_cfg[key]:
            continue
        fragments.append(key_abbre.get(key, key) + to_string(current_cfg[key]))
    if suffix:
        fragments.append(suffix)
    name = "_".join(fragments) if fragments else "default"
    return f"{name}-{str(uuid.uuid4())[:4]}"
