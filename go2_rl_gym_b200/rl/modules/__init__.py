from .actor_critic import ActorCritic
from .actor_critic_cts import (ActorCriticCTS, ActorCriticMoECTS, ActorCriticMoENGCTS, ActorCriticACMoECTS, ActorCriticDualMoECTS,
                               ActorCriticMCPCTS)
