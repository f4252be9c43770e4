"""Host bookkeeping for multi-node launches.  Parity: reference `colossalai/cli/launcher/hostinfo.py`."""
from __future__ import annotations

import socket
from typing import Iterator, List, Optional

__all__ = ["HostInfo", "HostInfoList"]


class HostInfo:
    def __init__(self, hostname: str, port: Optional[int] = None) -> None:
        self.hostname = hostname
        self.port = port
        self.is_local_host = HostInfo.is_host_localhost(hostname, port)

    @staticmethod
    def is_host_localhost(hostname: str, port: Optional[int] = None) -> bool:
        if port is None:
            port = 22
        if hostname in ("localhost", "127.0.0.1", socket.gethostname()):
            return True
        try:
            target = socket.getaddrinfo(hostname, port)[0][4][0]
            local = {a[4][0] for a in socket.getaddrinfo(socket.gethostname(), port)}
            return target in local or target.startswith("127.")
        except OSError:
            return False

    def __repr__(self) -> str:
        return f"hostname: {self.hostname}, port: {self.port}"


class HostInfoList:
    def __init__(self) -> None:
        self.hostinfo_list: List[HostInfo] = []

    def append(self, hostinfo: HostInfo) -> None:
        self.hostinfo_list.append(hostinfo)

    def remove(self, hostname: str) -> None:
        self.hostinfo_list.remove(self.get_hostinfo(hostname))

    def get_hostinfo(self, hostname: str) -> HostInfo:
        for h in self.hostinfo_list:
            if h.hostname == hostname:
                return h
        raise Exception(f"Hostname {hostname} is not found")

    def has(self, hostname: str) -> bool:
        return any(h.hostname == hostname for h in self.hostinfo_list)

    def __iter__(self) -> Iterator[HostInfo]:
        return iter(self.hostinfo_list)

    def __len__(self) -> int:
        return len(self.hostinfo_list)
