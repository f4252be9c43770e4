from .parallel_2d import Linear2D, split_2d
from .parallel_2p5d import Linear2p5D, split_2p5d
from .parallel_3d import Linear3D, split_3d_input

__all__ = ["Linear2D", "split_2d", "Linear2p5D", "split_2p5d", "Linear3D", "split_3d_input"]
