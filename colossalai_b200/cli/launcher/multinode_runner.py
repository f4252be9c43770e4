"""Run one launcher command per host: subprocess locally, `ssh` remotely.
Parity: reference `colossalai/cli/launcher/multinode_runner.py` (fabric connections + one process per host); fabric is
optional here, plain `ssh` is used so the launcher has no hard dependency."""
from __future__ import annotations

import shlex
import subprocess
from typing import Dict, List

from .hostinfo import HostInfo, HostInfoList

__all__ = ["MultiNodeRunner"]


class MultiNodeRunner:
    def __init__(self) -> None:
        self.procs: Dict[str, subprocess.Popen] = {}
        self.workdir = "."
        self.env: Dict[str, str] = {}

    def connect(self, host_info_list: HostInfoList, workdir: str, env: dict) -> None:
        self.workdir, self.env = workdir, dict(env)
        self.hosts = host_info_list

    def send(self, hostinfo: HostInfo, cmd: str) -> None:
        exports = " ".join(f"export {k}={shlex.quote(str(v))};" for k, v in self.env.items()
                           if k.startswith(("NCCL_", "CUDA_", "CB200_", "PYTHONPATH", "PATH", "LD_LIBRARY_PATH", "OMP_")))
        full = f"cd {shlex.quote(self.workdir)}; {exports} {cmd}"
        if hostinfo.is_local_host:
            proc = subprocess.Popen(["bash", "-c", full])
        else:
            ssh = ["ssh", "-o", "StrictHostKeyChecking=no"]
            if hostinfo.port:
                ssh += ["-p", str(hostinfo.port)]
            proc = subprocess.Popen(ssh + [hostinfo.hostname, full])
        self.procs[hostinfo.hostname] = proc

    def stop_all(self) -> None:
        for p in self.procs.values():     # exact processes we started, never by pattern
            if p.poll() is None:
                p.terminate()

    def recv_from_all(self) -> Dict[str, str]:
        out = {}
        for host, p in self.procs.items():
            rc = p.wait()
            out[host] = "success" if rc == 0 else f"failure (exit code {rc})"
        return out
