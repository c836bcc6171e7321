from go2_rl_gym_b200.utils import class_to_dict, get_load_path, get_args, set_seed, update_class_from_dict, task_registry, Terrain  # noqa: F401
