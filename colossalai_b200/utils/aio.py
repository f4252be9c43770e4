"""Python face of the native async file IO (`kernel/csrc/async_file_io.cpp`): safetensors writer running in
background threads and a disk offloader for optimizer states.

Parity: reference `colossalai/utils/safetensors.py:162-205` (tensornvme AsyncFileWriter glue) and the `DiskOffloader`
used by `nn/optimizer/nvme_optimizer.py`.
"""
from __future__ import annotations

import ctypes
import json
import os
import struct
from typing import Dict, List, Optional

import torch

from ..kernel import loader

__all__ = ["AsyncFile", "AsyncSafetensorsWriter", "DiskOffloader"]

_lib = None
_ST_DTYPES = {torch.float32: "F32", torch.float16: "F16", torch.bfloat16: "BF16", torch.float64: "F64",
              torch.int64: "I64", torch.int32: "I32", torch.int16: "I16", torch.int8: "I8", torch.uint8: "U8",
              torch.bool: "BOOL"}
if hasattr(torch, "float8_e4m3fn"):
    _ST_DTYPES[torch.float8_e4m3fn] = "F8_E4M3"
    _ST_DTYPES[torch.float8_e5m2] = "F8_E5M2"


def _get_lib():
    global _lib
    if _lib is None:
        lib = loader.load("cb200_aio")
        lib.cb_aio_open.restype = ctypes.c_void_p
        lib.cb_aio_open.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.cb_aio_write.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64]
        lib.cb_aio_read.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64]
        lib.cb_aio_synchronize.argtypes = [ctypes.c_void_p]
        lib.cb_aio_fsync.argtypes = [ctypes.c_void_p]
        lib.cb_aio_close.argtypes = [ctypes.c_void_p]
        _lib = lib
    return _lib


class AsyncFile:
    def __init__(self, path: str, write: bool = True, truncate: bool = True, n_threads: int = 4) -> None:
        self._lib = _get_lib()
        self.path = path
        self._h = self._lib.cb_aio_open(path.encode(), int(write), int(truncate), n_threads)
        if not self._h:
            raise OSError(f"cannot open {path}")
        self._keep: List = []

    def write(self, buf, nbytes: int, offset: int, keepalive=None) -> None:
        addr = buf if isinstance(buf, int) else buf.data_ptr()
        if keepalive is not None:
            self._keep.append(keepalive)
        self._lib.cb_aio_write(self._h, ctypes.c_void_p(addr), nbytes, offset)

    def write_bytes(self, data: bytes, offset: int) -> None:
        b = ctypes.create_string_buffer(data, len(data))
        self._keep.append(b)
        self._lib.cb_aio_write(self._h, ctypes.cast(b, ctypes.c_void_p), len(data), offset)

    def read(self, tensor: torch.Tensor, offset: int) -> None:
        self._keep.append(tensor)
        self._lib.cb_aio_read(self._h, ctypes.c_void_p(tensor.data_ptr()), tensor.numel() * tensor.element_size(),
                              offset)

    def synchronize(self) -> None:
        if self._h:
            err = self._lib.cb_aio_synchronize(self._h)
            self._keep.clear()
            if err:
                raise OSError(err, f"async IO on {self.path} failed: {os.strerror(err)}")

    def close(self) -> None:
        if self._h:
            self.synchronize()
            self._lib.cb_aio_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class AsyncSafetensorsWriter:
    """Writes a `.safetensors` file from host tensors in background threads; `synchronize()` joins and closes."""

    def __init__(self, path: str, n_threads: int = 4) -> None:
        self.path = path
        self._file: Optional[AsyncFile] = AsyncFile(path, write=True, truncate=True, n_threads=n_threads)

    def write(self, tensors: Dict[str, torch.Tensor], metadata: Optional[Dict[str, str]] = None) -> None:
        header: Dict[str, dict] = {"__metadata__": dict(metadata or {"format": "pt"})}
        off = 0
        items = []
        for name in sorted(tensors.keys()):
            t = tensors[name]
            assert t.device.type == "cpu", "stage tensors on the host before the async write"
            t = t if t.is_contiguous() else t.contiguous()
            n = t.numel() * t.element_size()
            header[name] = {"dtype": _ST_DTYPES[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + n]}
            items.append((t, n, off))
            off += n
        hb = json.dumps(header, separators=(",", ":")).encode()
        pad = (8 - len(hb) % 8) % 8
        hb += b" " * pad
        self._file.write_bytes(struct.pack("<Q", len(hb)) + hb, 0)
        base = 8 + len(hb)
        for t, n, o in items:
            if n:
                self._file.write(t, n, base + o, keepalive=t)

    def synchronize(self) -> None:
        if self._file is not None:
            self._file.close()
            self._file = None


class DiskOffloader:
    """Keeps tensors' storage on disk between uses (optimizer states on NVMe)."""

    def __init__(self, dir_name: str, n_entries: int = 8, backend: str = "pthread") -> None:
        os.makedirs(dir_name, exist_ok=True)
        self.dir = dir_name
        self._files: Dict[int, AsyncFile] = {}
        self._shapes: Dict[int, tuple] = {}
        self._reading: List[AsyncFile] = []
        self._writing: List[AsyncFile] = []

    def _file(self, t: torch.Tensor) -> AsyncFile:
        k = id(t)
        if k not in self._files:
            self._files[k] = AsyncFile(os.path.join(self.dir, f"offload_{k}.bin"), write=True, truncate=True,
                                       n_threads=2)
        return self._files[k]

    def async_write(self, t: torch.Tensor) -> None:
        """Persist `t` and release its host storage."""
        f = self._file(t)
        self._shapes[id(t)] = (t.numel(), t.element_size())
        f.write(t, t.numel() * t.element_size(), 0, keepalive=t.data)
        self._writing.append((f, t))

    def sync_write_events(self) -> None:
        for f, t in self._writing:
            f.synchronize()
            t.data.untyped_storage().resize_(0)
        self._writing.clear()

    def async_read(self, t: torch.Tensor) -> None:
        if id(t) not in self._shapes:
            return
        numel, esize = self._shapes[id(t)]
        if t.untyped_storage().size() == 0:
            t.untyped_storage().resize_(numel * esize)
        f = self._file(t)
        f.read(t, 0)
        self._reading.append(f)

    def sync_read_events(self) -> None:
        for f in self._reading:
            f.synchronize()
        self._reading.clear()

    def synchronize(self) -> None:
        self.sync_read_events()
        self.sync_write_events()
