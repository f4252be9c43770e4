"""Build and dispatch the `torch.distributed.run` command for every node.
Parity: reference `colossalai/cli/launcher/run.py` (`fetch_hostfile`, `parse_device_filter`, `get_launch_command`,
`launch_multi_processes`)."""
from __future__ import annotations

import os
import sys
from typing import Dict, List, Optional

from .hostinfo import HostInfo, HostInfoList
from .multinode_runner import MultiNodeRunner

__all__ = ["fetch_hostfile", "parse_device_filter", "get_launch_command", "launch_multi_processes"]


def fetch_hostfile(hostfile_path: str, ssh_port: Optional[int]) -> HostInfoList:
    if not os.path.isfile(hostfile_path):
        raise FileNotFoundError(f"Unable to find the hostfile, no such file: {hostfile_path}")
    pool = HostInfoList()
    with open(hostfile_path) as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            host = line.split()[0]
            if pool.has(host):
                raise ValueError(f"Hostfile contains multiple entries for {host}")
            pool.append(HostInfo(hostname=host, port=ssh_port))
    return pool


def parse_device_filter(device_pool: HostInfoList, include_str: Optional[str] = None,
                        exclude_str: Optional[str] = None) -> HostInfoList:
    if include_str and exclude_str:
        raise ValueError("include_str and exclude_str are mutually exclusive, only one can be used")
    if not include_str and not exclude_str:
        return device_pool
    names = (include_str or exclude_str).split(",")
    for n in names:
        if not device_pool.has(n):
            raise ValueError(f"Hostname {n} is not defined in the hostfile")
    out = HostInfoList()
    for h in device_pool:
        if (include_str and h.hostname in names) or (exclude_str and h.hostname not in names):
            out.append(h)
    return out


def get_launch_command(master_addr: str, master_port: int, nproc_per_node: int, user_script: str,
                       user_args: List[str], node_rank: int = 0, num_nodes: int = 1, run_as_module: bool = False,
                       extra_launch_args: Optional[str] = None) -> str:
    extra: Dict[str, Optional[str]] = {}
    if extra_launch_args:
        for item in extra_launch_args.split(","):
            k, _, v = item.partition("=")
            extra[k.strip()] = v.strip() if v else None
    args = {"nproc_per_node": nproc_per_node, "nnodes": num_nodes}
    if num_nodes == 1 and "standalone" not in extra:
        args.update(master_addr=master_addr, master_port=master_port, node_rank=0)
    else:
        args.update(node_rank=node_rank, rdzv_backend="c10d", rdzv_endpoint=f"{master_addr}:{master_port}",
                    rdzv_id="colossalai_b200-default-job")
    parts = [sys.executable, "-m", "torch.distributed.run"]
    for k, v in {**args, **extra}.items():
        parts.append(f"--{k}" if v is None else f"--{k}={v}")
    if run_as_module:
        parts.append("-m")
    parts.append(user_script)
    parts += list(user_args)
    return " ".join(parts)


def launch_multi_processes(args) -> int:
    if args.nproc_per_node is None:
        try:
            import torch

            args.nproc_per_node = max(torch.cuda.device_count(), 1)
        except Exception:
            args.nproc_per_node = 1
    if args.hostfile:
        pool = fetch_hostfile(args.hostfile, args.ssh_port)
        active = parse_device_filter(pool, args.include, args.exclude)
        if args.num_nodes > 0:
            limited = HostInfoList()
            for i, h in enumerate(active):
                if i < args.num_nodes:
                    limited.append(h)
            active = limited
    elif args.host:
        active = HostInfoList()
        for h in args.host.strip().split(","):
            active.append(HostInfo(hostname=h, port=args.ssh_port))
    else:
        active = HostInfoList()
        active.append(HostInfo(hostname="127.0.0.1", port=args.ssh_port))
    runner = MultiNodeRunner()
    runner.connect(active, os.getcwd(), dict(os.environ))
    master = args.master_addr or "127.0.0.1"
    for rank, host in enumerate(active):
        cmd = get_launch_command(master, args.master_port, args.nproc_per_node, args.user_script, args.user_args,
                                 node_rank=rank, num_nodes=len(active), run_as_module=args.m,
                                 extra_launch_args=args.extra_launch_args)
        runner.send(host, cmd)
    results = runner.recv_from_all()
    ok = all(v == "success" for v in results.values())
    if not ok:
        print("\n".join(f"{h}: {v}" for h, v in results.items()), file=sys.stderr)
    return 0 if ok else 1
