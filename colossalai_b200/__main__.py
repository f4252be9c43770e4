import sys

from .cli import cli

sys.exit(cli())
