"""Nested-class configuration tree (public API; reference: legged_gym/envs/base/base_config.py:4-24).

Instantiating a config turns every nested class attribute into an instance, recursively, so that
`cfg.terrain.num_rows = 7` on one config object never leaks into another."""
import inspect


class BaseConfig:
    def __init__(self):
        _instantiate_members(self)


def _instantiate_members(node):
    for name in dir(node):
        if name == "__class__":
            continue
        member = getattr(node, name)
        if inspect.isclass(member):
            child = member()
            setattr(node, name, child)
            _instantiate_members(child)


BaseConfig.init_member_classes = staticmethod(_instantiate_members)
