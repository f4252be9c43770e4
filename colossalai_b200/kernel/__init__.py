from .loader import KernelLoader, build_all, launch_counter, load, native_available

__all__ = ["KernelLoader", "build_all", "launch_counter", "load", "native_available"]
