"""Minimal stand-in for NVIDIA Isaac Gym so that the REFERENCE's Python (legged_gym / rsl_rl.runners) can be imported
from /root/reference in this container (Isaac Gym itself is closed source and absent: SURVEY F1/F2).
Test tooling only — used by tests/golden/make_golden_*.py to generate fixtures; never imported by the product."""
