def parse_device_str(s):
    if ':' in s:
        a, b = s.split(':')
        return a, int(b)
    return s, 0


def parse_arguments(*a, **k):
    raise RuntimeError("stub")


def parse_sim_config(*a, **k):
    pass
