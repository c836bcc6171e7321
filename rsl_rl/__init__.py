"""Import shim: `rsl_rl.*` resolves to the B200-native trainer (go2_rl_gym_b200.rl)."""
