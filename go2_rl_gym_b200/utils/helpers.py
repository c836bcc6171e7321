"""helpers (drop-in for legged_gym/utils/helpers.py:12-170): class_to_dict, set_seed, get_args, cfg overrides, checkpoint paths.
`get_args` no longer goes through isaacgym.gymutil.parse_arguments; it accepts the same flags the reference's scripts use."""
import argparse
import copy
import os
import random

import numpy as np
import torch

from .cfg_dict import class_to_dict  # noqa: F401


def update_class_from_dict(obj, dict_):
    for key, val in dict_.items():
        attr = getattr(obj, key, None)
        if isinstance(attr, type) or (attr is not None and hasattr(attr, "__dict__") and isinstance(val, dict)):
            update_class_from_dict(attr, val)
        else:
            setattr(obj, key, val)


def set_seed(seed):
    if seed == -1:
        seed = np.random.randint(0, 10000)
    print("Setting seed: {}".format(seed))
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    os.environ['PYTHONHASHSEED'] = str(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)


def parse_sim_params(args, cfg):
    """The reference returns a gymapi.SimParams (helpers.py:50-72); the B200 env reads cfg.sim directly, so this is a namespace."""
    ns = argparse.Namespace(**{k: v for k, v in cfg.get("sim", {}).items() if not isinstance(v, dict)})
    ns.use_gpu_pipeline = True
    return ns


def get_load_path(root, load_run=-1, checkpoint=-1):
    try:
        runs = sorted(os.listdir(root))
        if 'exported' in runs:
            runs.remove('exported')
        last_run = os.path.join(root, runs[-1])
    except Exception:
        raise ValueError("No runs in this directory: " + root)
    load_run = last_run if load_run == -1 else os.path.join(root, load_run)
    if checkpoint == -1:
        models = [f for f in os.listdir(load_run) if 'model' in f]
        models.sort(key=lambda m: '{0:0>15}'.format(m))
        model = models[-1]
    else:
        model = "model_{}.pt".format(checkpoint)
    return os.path.join(load_run, model)


def update_cfg_from_args(env_cfg, cfg_train, args):
    if env_cfg is not None and args.num_envs is not None:
        env_cfg.env.num_envs = args.num_envs
    if cfg_train is not None:
        if args.seed is not None:
            cfg_train.seed = args.seed
        if args.max_iterations is not None:
            cfg_train.runner.max_iterations = args.max_iterations
        if args.resume:
            cfg_train.runner.resume = args.resume
        if args.experiment_name is not None:
            cfg_train.runner.experiment_name = args.experiment_name
        if args.run_name is not None:
            cfg_train.runner.run_name = args.run_name
        if args.load_run is not None:
            cfg_train.runner.load_run = args.load_run
        if args.checkpoint is not None:
            cfg_train.runner.checkpoint = args.checkpoint
        # the RoboGauge evaluation client itself is out of scope; its switches are still carried in the config like the reference does
        if getattr(args, "robogauge", None) is not None and hasattr(cfg_train, "robogauge"):
            cfg_train.robogauge.enabled = args.robogauge
        if getattr(args, "robogauge_port", None) is not None and hasattr(cfg_train, "robogauge"):
            cfg_train.robogauge.port = args.robogauge_port
    return env_cfg, cfg_train


def get_args(argv=None):
    p = argparse.ArgumentParser(description="RL Policy")
    p.add_argument("--task", type=str, default="go2")
    p.add_argument("--resume", action="store_true", default=False)
    p.add_argument("--experiment_name", type=str)
    p.add_argument("--run_name", type=str)
    p.add_argument("--load_run", type=str)
    p.add_argument("--checkpoint", type=int)
    p.add_argument("--headless", action="store_true", default=False)
    p.add_argument("--horovod", action="store_true", default=False)
    p.add_argument("--rl_device", type=str, default="cuda:0")
    p.add_argument("--sim_device", type=str, default="cuda:0")
    p.add_argument("--pipeline", type=str, default="gpu")
    p.add_argument("--num_envs", type=int)
    p.add_argument("--seed", type=int)
    p.add_argument("--max_iterations", type=int)
    p.add_argument("--robogauge", action="store_true", default=False)
    p.add_argument("--robogauge_port", type=int, default=9973)
    args = p.parse_args(argv)
    args.physics_engine = "b200"
    args.sim_device_id = int(args.sim_device.split(":")[1]) if ":" in args.sim_device else 0
    args.sim_device_type = args.sim_device.split(":")[0]
    if args.sim_device_type == 'cuda':
        args.sim_device = f"cuda:{args.sim_device_id}"
    return args
