from enum import Enum


class TensorType(Enum):
    MODEL = 0
    NONMODEL = 1      # mainly activations
