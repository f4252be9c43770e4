from .consumer import GRPOConsumer
from .launch import launch_distributed
from .producer import EngineRolloutBackend, ModelRolloutBackend, Producer
from .reward import boxed_math_reward, extract_boxed, format_reward, make_reward_fn

__all__ = ["Producer", "ModelRolloutBackend", "EngineRolloutBackend", "GRPOConsumer", "launch_distributed",
           "boxed_math_reward", "format_reward", "extract_boxed", "make_reward_fn"]
