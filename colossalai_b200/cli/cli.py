"""`colossalai_b200` command line: `run` (launcher) and `check` (installation report).
Parity: reference `colossalai/cli/cli.py` (click group with `run`, `check`); argparse here (click is not required)."""
from __future__ import annotations

import argparse
import sys
from typing import List, Optional


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(prog="colossalai_b200", description="colossalai_b200 command line")
    sub = parser.add_subparsers(dest="command")
    from .check import add_check_parser
    from .launcher import add_run_parser

    add_run_parser(sub)
    add_check_parser(sub)
    return parser


def cli(argv: Optional[List[str]] = None) -> int:
    parser = build_parser()
    args, extra = parser.parse_known_args(argv)
    if args.command is None:
        parser.print_help()
        return 0
    return int(args.func(args, extra) or 0)


if __name__ == "__main__":
    sys.exit(cli())
