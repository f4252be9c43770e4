"""`colossalai_b200 run`: launch a training script on one or several nodes.

    colossalai_b200 run --nproc_per_node 8 train.py --config cfg.yaml
    colossalai_b200 run --hostfile hosts --nproc_per_node 8 --master_addr node0 train.py

Parity: reference `colossalai/cli/launcher/__init__.py` (`colossalai run` options)."""
from __future__ import annotations

import argparse

from .run import launch_multi_processes

__all__ = ["add_run_parser"]


def _run(args: argparse.Namespace, extra) -> int:
    if args.user_script is None:
        print("Error: missing script argument. Try colossalai_b200 run --help")
        return 2
    if not args.m and not args.user_script.endswith(".py"):
        print(f"Error: invalid Python file {args.user_script}. Did you use a wrong option? Try colossalai_b200 run --help")
        return 2
    args.user_args = list(args.user_args or []) + list(extra or [])
    return launch_multi_processes(args)


def add_run_parser(sub) -> None:
    p = sub.add_parser("run", help="Launch distributed training on a single node or multiple nodes")
    p.add_argument("-H", "-host", "--host", type=str, default=None, help="list of hostnames <host1>,<host2>")
    p.add_argument("--hostfile", type=str, default=None, help="file with one hostname per line")
    p.add_argument("--include", type=str, default=None, help="hosts of the hostfile to use: <host1>,<host2>")
    p.add_argument("--exclude", type=str, default=None, help="hosts of the hostfile NOT to use")
    p.add_argument("--num_nodes", type=int, default=-1, help="number of nodes to use (with --hostfile)")
    p.add_argument("--nproc_per_node", type=int, default=None, help="GPUs to use on each node")
    p.add_argument("--master_port", type=int, default=29500)
    p.add_argument("--master_addr", type=str, default="127.0.0.1")
    p.add_argument("--extra_launch_args", type=str, default=None,
                   help="extra torch.distributed.run arguments: arg1=1,arg2=2 (flag-only args without '=')")
    p.add_argument("--ssh-port", dest="ssh_port", type=int, default=None)
    p.add_argument("-m", action="store_true", help="run a library module as a script")
    p.add_argument("user_script", type=str, nargs="?", default=None)
    p.add_argument("user_args", nargs=argparse.REMAINDER)
    p.set_defaults(func=_run)
