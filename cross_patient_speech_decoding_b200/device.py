"""Device plumbing: torch supplies HBM allocations, streams and pinned staging only.

All arithmetic on the product path happens in the kernels of ``libcpsd_b200.so``; this
module fails loudly when CUDA or the library is unavailable (there is no CPU fallback).
"""
import ctypes
import os

# see bench.py: streams that alias onto one hardware queue serialise; set before CUDA initialises
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')

import numpy as np  # noqa: E402
import torch  # noqa: E402

from . import _lib

_SYNC_DEBUG = os.environ.get('CPSD_SYNC', '0') == '1'
# CPSD_POISON=1: every fresh device allocation is filled with NaN (0x7f.. for ints) so that a
# kernel that reads memory nobody wrote shows up in the parity tests instead of depending on
# whatever the caching allocator hands back
_POISON = os.environ.get('CPSD_POISON', '0') == '1'


class Context:
    """One per process/GPU: library handle, device, stream."""

    _instances = {}

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise _lib.CpsdError('CUDA device required: the B200 path has no CPU fallback')
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else
                                   'cuda:%d' % torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        arch = self.lib.cpsd_device_arch()
        if arch < 100:
            raise _lib.CpsdError('libcpsd_b200 is built for sm_100a only (device reports sm_%d)'
                                 % arch)

    @classmethod
    def get(cls, device=None):
        key = str(device)
        if key not in cls._instances:
            cls._instances[key] = cls(device)
        return cls._instances[key]

    @property
    def stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ allocation
    def empty(self, shape, dtype=torch.float32):
        t = torch.empty(shape, dtype=dtype, device=self.device)
        if _POISON:
            if t.dtype.is_floating_point:
                t.fill_(float('nan'))
            elif t.dtype == torch.uint8:
                t.fill_(0x7f)
            else:
                t.fill_(0x7f7f7f7f if t.dtype == torch.int32 else 0x7f)
        return t

    def zeros(self, shape, dtype=torch.float32):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def upload(self, arr, dtype=None):
        """Host numpy -> new device tensor (through pinned staging, async on the stream)."""
        a = np.ascontiguousarray(arr, dtype=dtype)
        t = torch.from_numpy(a)
        if a.size:
            t = t.pin_memory()
        return t.to(self.device, non_blocking=True)

    def launches(self):
        return int(self.lib.cpsd_launch_count())

    def call(self, name, *args):
        fn = getattr(self.lib, name)
        _lib.check(fn(*args, self.stream), name)
        if _SYNC_DEBUG:      # CPSD_SYNC=1: localise an asynchronous fault to the call that made it
            try:
                torch.cuda.synchronize(self.device)
            except Exception as e:
                raise _lib.CpsdError('%s faulted: %s' % (name, str(e).splitlines()[0]))


class LaneContext:
    """A Context bound to one CUDA stream (an engine lane): ``call`` passes the cached stream
    handle instead of asking torch for the current stream on every launch (2 000 launches per
    streamed job made that lookup 10 % of the host time of the end-to-end path).  The owner
    keeps torch's current stream equal to ``stream`` while it uses the lane."""

    def __init__(self, base, stream):
        self.base = base
        self.lib = base.lib
        self.device = base.device
        self.torch_stream = stream
        self.stream = ctypes.c_void_p(stream.cuda_stream)
        self.empty, self.zeros, self.upload, self.launches = base.empty, base.zeros, base.upload, base.launches

    def call(self, name, *args):
        st = getattr(self.lib, name)(*args, self.stream)
        if st:
            _lib.check(st, name)
        if _SYNC_DEBUG:
            try:
                torch.cuda.synchronize(self.device)
            except Exception as e:
                raise _lib.CpsdError('%s faulted: %s' % (name, str(e).splitlines()[0]))


def ptr(t, offset_elems=0):
    """Device address of a tensor element as c_void_p (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr() + offset_elems * t.element_size())


def addr(t, offset_elems=0):
    """Device address as a python int (for descriptor fields)."""
    if t is None:
        return 0
    return t.data_ptr() + offset_elems * t.element_size()


class HostPack:
    """Packs the per-batch integer tables and descriptor records on the host so that one
    batch needs exactly two host->device copies."""

    def __init__(self, ctx):
        self.ctx = ctx
        self._ints = []
        self._n_int = 0
        self._descs = []
        self._n_desc = 0
        self.ibuf = None
        self.dbuf = None
        self._icap = 0
        self._dcap = 0
        self._ih = None
        self._dh = None

    def reset(self):
        self._ints, self._n_int, self._descs, self._n_desc = [], 0, [], 0

    def add_ints(self, arr):
        """Queues an int32 array; returns its element offset inside the int table."""
        a = np.ascontiguousarray(arr, dtype=np.int32).ravel()
        off = self._n_int
        self._ints.append(a)
        self._n_int += (a.size + 3) & ~3   # keep 16-byte alignment between tables
        return off

    def reserve_ints(self):
        """Allocates the device int table (after all add_ints calls); returns base address."""
        need = max(self._n_int, 4)
        if need > self._icap:
            self._icap = int(need * 1.5) + 64
            self.ibuf = self.ctx.empty((self._icap,), torch.int32)
            self._ih = torch.empty((self._icap,), dtype=torch.int32).pin_memory()
        return self.ibuf.data_ptr()

    def iaddr(self, off):
        return self.ibuf.data_ptr() + 4 * off

    def add_descs(self, recs):
        """Queues a structured array of descriptor records; returns byte offset."""
        b = np.ascontiguousarray(recs).view(np.uint8).ravel()
        off = self._n_desc
        self._descs.append(b)
        self._n_desc += (b.size + 15) & ~15
        return off

    def upload(self):
        """Copies both tables; returns nothing (use iaddr / daddr afterwards)."""
        need = max(self._n_desc, 16)
        if need > self._dcap:
            self._dcap = int(need * 1.5) + 256
            self.dbuf = self.ctx.empty((self._dcap,), torch.uint8)
            self._dh = torch.empty((self._dcap,), dtype=torch.uint8).pin_memory()
        hi = self._ih.numpy()
        o = 0
        for a in self._ints:
            hi[o:o + a.size] = a
            o += (a.size + 3) & ~3
        self.ibuf[:max(o, 1)].copy_(self._ih[:max(o, 1)], non_blocking=True)
        hd = self._dh.numpy()
        o = 0
        for b in self._descs:
            hd[o:o + b.size] = b
            o += (b.size + 15) & ~15
        self.dbuf[:max(o, 1)].copy_(self._dh[:max(o, 1)], non_blocking=True)
        self.h2d_bytes = 4 * self._n_int + self._n_desc

    def daddr(self, off):
        return ctypes.c_void_p(self.dbuf.data_ptr() + off)
