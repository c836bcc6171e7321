from go2_rl_gym_b200.utils import (class_to_dict, get_load_path, get_args, set_seed, update_class_from_dict, task_registry, Terrain, Logger,  # noqa: F401
                                  quat_apply_yaw, wrap_to_pi, torch_rand_sqrt_float)
