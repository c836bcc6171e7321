from go2_rl_gym_b200.utils.exporter import export_policy_as_jit, export_policy_as_onnx, export_policy_as_pkl, build_export_module  # noqa: F401
