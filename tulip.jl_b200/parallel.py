"""Multi-GPU KKT backend: one process per GPU (torch.distributed / NCCL over NVLink), independent
elimination-tree subtrees sharded across the ranks, the separator ("top") part replicated.

The reference has no distributed code (its block-angular linear algebra was moved out of the package,
/root/reference/NEWS.md:31); this is the B200-native equivalent described in SURVEY.md 8e:

    update!:  every rank assembles + factors its own subtrees; their Schur-complement contributions land
              in the rank-local copy of the top panels  ->  ONE all-reduce(sum) of that contiguous buffer
              ->  every rank factors the (small) top part redundantly.
    solve!:   forward sweep on own subtrees -> all-reduce(sum) of the work vector -> top forward+backward
              (redundant) -> backward sweep on own subtrees -> all-reduce(sum) -> dx, dy on every rank.

The same `update(θinv, regP, regD)` / `solve(dx, dy, ξp, ξd)` interface as B200KKTSolver, so the IPM
driver (hsd.py) runs unchanged on every rank (SPMD, identical host state).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .kkt import B200KKTSolver, Backend, DimensionMismatch, K1, K2, PosDefException, _dp, _raise  # noqa: F401


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can view a device buffer owned by the library."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class DistB200KKT:
    def __init__(self, A, system, backend: Backend | None = None, group=None):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("DistB200KKT needs an initialised torch.distributed process group (backend nccl)")
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        be = backend or Backend()
        be = Backend(**{**be.__dict__, "rank": self.rank, "nranks": self.world, "use_graph": False})
        self.local = B200KKTSolver(A, system, be)
        self.m, self.n = self.local.m, self.local.n
        self.device = torch.device(f"cuda:{be.device}")
        lib = _lib.load()
        p = C.c_void_p(); cnt = C.c_int64(0)
        lib.tlpb200_top_panels(self.local._h, C.byref(p), C.byref(cnt))
        self._top = torch.as_tensor(_DevArray(p.value, cnt.value), device=self.device) if cnt.value > 0 else None
        lib.tlpb200_work_vector(self.local._h, C.byref(p), C.byref(cnt))
        self._wk = torch.as_tensor(_DevArray(p.value, cnt.value), device=self.device)
        self._code = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.bytes_allreduce_update = 8 * (self._top.numel() if self._top is not None else 0)
        self.bytes_allreduce_solve = 2 * 8 * self._wk.numel()

    def update(self, theta_inv, regP, regD):
        theta_inv = np.ascontiguousarray(theta_inv, dtype=np.float64)
        regP = np.ascontiguousarray(regP, dtype=np.float64)
        regD = np.ascontiguousarray(regD, dtype=np.float64)
        if theta_inv.shape[0] != self.n or regP.shape[0] != self.n or regD.shape[0] != self.m:
            raise DimensionMismatch("update!: vector lengths do not match the KKT solver")
        lib, h = _lib.load(), self.local._h
        rc = lib.tlpb200_update_begin(h, _dp(theta_inv), _dp(regP), _dp(regD))
        if rc != _lib.OK:
            _raise(rc, h)
        if self._top is not None:
            self.dist.all_reduce(self._top, op=self.dist.ReduceOp.SUM, group=self.group)   # separator reduce over NVLink
            self.torch.cuda.current_stream(self.device).synchronize()
        bad = C.c_int64(-1)
        rc = lib.tlpb200_update_end(h, C.byref(bad))
        # a breakdown inside one rank's subtree must raise PosDefException on EVERY rank (step.jl:34-51)
        self._code[0] = rc
        self.dist.all_reduce(self._code, op=self.dist.ReduceOp.MAX, group=self.group)
        worst = int(self._code.item())
        if worst != _lib.OK:
            if rc == _lib.OK and worst == _lib.NOT_POSDEF:
                raise PosDefException("factorisation breakdown on another rank")
            _raise(rc if rc != _lib.OK else worst, h)

    def solve(self, dx, dy, xi_p, xi_d):
        xi_p = np.ascontiguousarray(xi_p, dtype=np.float64)
        xi_d = np.ascontiguousarray(xi_d, dtype=np.float64)
        lib, h = _lib.load(), self.local._h
        rc = lib.tlpb200_solve_begin(h, _dp(xi_p), _dp(xi_d))
        if rc != _lib.OK:
            _raise(rc, h)
        self.dist.all_reduce(self._wk, op=self.dist.ReduceOp.SUM, group=self.group)
        self.torch.cuda.current_stream(self.device).synchronize()
        rc = lib.tlpb200_solve_mid(h)
        if rc != _lib.OK:
            _raise(rc, h)
        self.dist.all_reduce(self._wk, op=self.dist.ReduceOp.SUM, group=self.group)
        self.torch.cuda.current_stream(self.device).synchronize()
        rc = lib.tlpb200_solve_end(h, _dp(dx), _dp(dy))
        if rc != _lib.OK:
            _raise(rc, h)

    def stats(self):
        return self.local.stats()

    def dist_info(self):
        return self.local.dist_info()
