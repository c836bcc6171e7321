from .ppo import PPO
from .cts import CTS, MoECTS, MoENGCTS, ACMoECTS, DualMoECTS, MCPCTS
