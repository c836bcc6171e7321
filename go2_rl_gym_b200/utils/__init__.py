from .helpers import class_to_dict, get_load_path, get_args, set_seed, update_class_from_dict  # noqa: F401
from .task_registry import task_registry  # noqa: F401
from .terrain import Terrain  # noqa: F401
from .logger import Logger  # noqa: F401
from .math import quat_apply_yaw, wrap_to_pi, torch_rand_sqrt_float  # noqa: F401
