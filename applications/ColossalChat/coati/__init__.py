"""coati: RLHF / preference-optimisation toolkit on top of colossalai_b200 (SFT, reward modelling, DPO / SimPO, ORPO,
KTO, PPO, GRPO / DAPO, rollout producer - trainer consumer loop).

Parity: reference `applications/ColossalChat/coati` (models, experience_maker, experience_buffer, dataset, trainer,
distributed).  The reference's distributed RL runs producers (vLLM / SGLang) and consumers as ray actors; here both
sides are ordinary ranks of one torchrun job and the rollout backend is our own paged-KV inference engine.
"""
