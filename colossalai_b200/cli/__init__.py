from .cli import cli

__all__ = ["cli"]
