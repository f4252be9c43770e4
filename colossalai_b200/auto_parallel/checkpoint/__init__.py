from .rotor import Chain, CheckpointSolverRotor, Sequence, apply_rotor_checkpointing, profile_chain

__all__ = ["Chain", "Sequence", "CheckpointSolverRotor", "apply_rotor_checkpointing", "profile_chain"]
