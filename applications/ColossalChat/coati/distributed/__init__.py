from .comm import RolloutBuffer, WeightMailbox, deserialize_rollout, merge_rollouts, serialize_rollout
from .consumer import GRPOConsumer
from .launch import launch_distributed
from .launch_zero_bubble import launch_zero_bubble
from .producer import EngineRolloutBackend, ModelRolloutBackend, Producer
from .profiling_utils import StepProfiler
from .reward import (boxed_math_reward, code_reward, combine_rewards, extract_boxed, format_reward, length_penalty,
                     make_reward_fn)

__all__ = ["Producer", "ModelRolloutBackend", "EngineRolloutBackend", "GRPOConsumer", "launch_distributed",
           "launch_zero_bubble", "RolloutBuffer", "WeightMailbox", "serialize_rollout", "deserialize_rollout", "merge_rollouts",
           "StepProfiler", "boxed_math_reward", "format_reward", "extract_boxed", "make_reward_fn", "length_penalty",
           "code_reward", "combine_rewards"]
