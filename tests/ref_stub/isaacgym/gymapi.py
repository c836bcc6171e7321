class _Any:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


SIM_PHYSX = 1
Vec3 = Transform = AssetOptions = PlaneParams = HeightFieldParams = TriangleMeshParams = SimParams = _Any
KEY_ESCAPE = KEY_V = 0


def acquire_gym():
    return _Any()
