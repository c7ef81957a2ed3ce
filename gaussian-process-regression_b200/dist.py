"""Multi-GPU Cholesky + solve (BASELINE config 5): host glue around gprc_dist_* of libgprc.
One process per GPU (torchrun); torch.distributed is only used to ship the NCCL unique id and for barriers."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .kernels import KernelSpec, as_matrix


def _loaded_nccl_path():
    """The libnccl.so.2 this process already holds (torch's bundled one), so that libgprc binds the same library."""
    try:
        with open("/proc/self/maps") as f:
            for line in f:
                if "libnccl.so" in line:
                    return line.split()[-1]
    except OSError:
        pass
    return None


class DistGPR:
    """Collective GPR "train" for matrices beyond one GPU: K + noise I factored across the ranks of ``group``."""

    def __init__(self, ctx=None, group=None):
        import torch.distributed as dist
        self.ctx = ctx or _lib.default_context()
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        path = _loaded_nccl_path()
        bpath = path.encode() if path else None
        ident = [None]
        if self.rank == 0:
            buf = C.create_string_buffer(128)
            _lib.check(self.ctx.lib.gprc_dist_unique_id(buf, bpath))
            ident[0] = buf.raw
        dist.broadcast_object_list(ident, src=0, group=group)
        h = _lib._P()
        _lib.check(self.ctx.lib.gprc_dist_create(self.ctx.handle, ident[0], self.rank, self.world, bpath, C.byref(h)))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.gprc_dist_free(self.handle)
            self.handle = None

    def fit(self, X, y, noise, spec: KernelSpec, want_alpha=True):
        """-> dict(logp, alpha, info, phase_ms=dict(build, factor, solve, total)); identical on every rank."""
        X = as_matrix(X)
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64))
        n = X.shape[1]
        xp = _lib.points(X)
        kc, keep = spec.to_c()
        logp, info = C.c_double(0.0), C.c_long(0)
        alpha = np.empty(n) if want_alpha else None
        phases = (C.c_double * 4)()
        _lib.check(self.ctx.lib.gprc_dist_gpr_fit(self.handle, kc, _lib.dptr(xp), X.shape[0], n, _lib.dptr(y),
                                                  float(noise), C.byref(logp), _lib.dptr(alpha) if want_alpha else None,
                                                  C.byref(info), phases))
        return dict(logp=logp.value, alpha=alpha, info=info.value,
                    phase_ms=dict(build=phases[0], factor=phases[1], solve=phases[2], total=phases[3]))

    def fit_replicated(self, xp, d, n, y, noise, spec: KernelSpec):
        """Collective train; every rank gets an ordinary model handle (gprc_gpr*) holding the complete factor.
        xp: the ABI's point-major buffer (n x d C-contiguous).  -> (handle, logp, info, phase_ms)"""
        kc, keep = spec.to_c()
        h = _lib._P()
        logp, info = C.c_double(0.0), C.c_long(0)
        phases = (C.c_double * 4)()
        _lib.check(self.ctx.lib.gprc_dist_gpr_fit_replicated(self.handle, kc, _lib.dptr(xp), d, n, _lib.dptr(y),
                                                             float(noise), C.byref(h), C.byref(logp), C.byref(info),
                                                             phases))
        return h, logp.value, info.value, dict(build=phases[0], factor=phases[1], solve=phases[2], total=phases[3])
