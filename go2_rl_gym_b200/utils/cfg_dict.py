"""class_to_dict (legged_gym/utils/helpers.py:12-27): nested config object -> plain dict (public attributes only)."""


def class_to_dict(obj):
    if not hasattr(obj, "__dict__"):
        return obj
    result = {}
    for key in dir(obj):
        if key.startswith("_"):
            continue
        val = getattr(obj, key)
        if isinstance(val, list):
            result[key] = [class_to_dict(v) for v in val]
        else:
            result[key] = class_to_dict(val)
    return result
