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
from .kkt import B200KKTSolver, Backend, DimensionMismatch, K1, K2, PosDefException, TlpB200Error, _dp, _raise  # noqa: F401


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can view a device buffer owned by the library."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class DistB200KKT:
    """Sharded KKT solver.  ``mode="library"`` (default): the collectives are NCCL calls made by libtlpb200.so itself on
    the solver's stream (tlpb200_comm_init), so update!/solve! are the ordinary C-ABI calls and each is one CUDA-graph
    replay -- no host synchronisation between the phases, only the separator entries are reduced in a solve.
    ``mode="phases"``: the round-1 choreography (phase API of the C ABI + torch.distributed all-reduces between the
    phases), kept as the cross-check of the library path."""

    def __init__(self, A, system, backend: Backend | None = None, group=None, mode="library"):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("DistB200KKT needs an initialised torch.distributed process group (backend nccl)")
        if mode not in ("library", "phases"):
            raise ValueError("mode must be 'library' or 'phases'")
        self.torch, self.dist, self.group, self.mode = torch, dist, group, mode
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        be = backend or Backend()
        be = Backend(**{**be.__dict__, "rank": self.rank, "nranks": self.world,
                        "use_graph": be.use_graph if mode == "library" else False})
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
        if mode == "library":
            # one rank creates the ncclUniqueId, torch.distributed (the plumbing) broadcasts it, every rank joins
            idbuf = (C.c_char * 128)()
            if self.rank == 0:
                rc = lib.tlpb200_comm_unique_id(C.cast(idbuf, C.c_void_p))
                if rc != _lib.OK:
                    raise TlpB200Error("tlpb200_comm_unique_id failed (libnccl.so.2 not loadable?)")
            t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).to(self.device)
            dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            raw = bytes(t.cpu().numpy().tobytes())
            idbuf2 = C.create_string_buffer(raw, 128)
            rc = lib.tlpb200_comm_init(self.local._h, C.cast(idbuf2, C.c_void_p))
            if rc != _lib.OK:
                _raise(rc, self.local._h)

    # -- KKT.update! --------------------------------------------------------------------------
    def update(self, theta_inv, regP, regD):
        if self.mode == "library":
            return self.local.update(theta_inv, regP, regD)     # identical return code on every rank (all-reduced status)
        theta_inv = np.ascontiguousarray(theta_inv, dtype=np.float64)
        regP = np.ascontiguousarray(regP, dtype=np.float64)
        regD = np.ascontiguousarray(regD, dtype=np.float64)
        if theta_inv.shape[0] != self.n or regP.shape[0] != self.n or regD.shape[0] != self.m:
            raise DimensionMismatch("update!: vector lengths do not match the KKT solver")
        lib, h = _lib.load(), self.local._h
        rc = self._agree(lib.tlpb200_update_begin(h, _dp(theta_inv), _dp(regP), _dp(regD)))
        if self._top is not None:
            self.dist.all_reduce(self._top, op=self.dist.ReduceOp.SUM, group=self.group)   # separator reduce over NVLink
            self.torch.cuda.current_stream(self.device).synchronize()
        bad = C.c_int64(-1)
        self._agree(lib.tlpb200_update_end(h, C.byref(bad)))

    def _agree(self, rc):
        """all-reduce(MAX) of a return code BEFORE the next data collective, so that no rank is left waiting in a
        collective its peers never enter, and every rank raises the same exception class."""
        self._code[0] = int(rc)
        self.dist.all_reduce(self._code, op=self.dist.ReduceOp.MAX, group=self.group)
        worst = int(self._code.item())
        if worst != _lib.OK:
            if worst == _lib.NOT_POSDEF:
                raise PosDefException("factorisation breakdown" + ("" if rc == worst else " on another rank"))
            if rc == worst:
                _raise(rc, self.local._h)
            raise TlpB200Error(f"tlpb200 error {worst} on another rank")
        return rc

    # -- KKT.solve! ---------------------------------------------------------------------------
    def solve(self, dx, dy, xi_p, xi_d):
        if self.mode == "library":
            return self.local.solve(dx, dy, xi_p, xi_d)
        xi_p = np.ascontiguousarray(xi_p, dtype=np.float64)
        xi_d = np.ascontiguousarray(xi_d, dtype=np.float64)
        if dx.shape[0] != self.n or xi_d.shape[0] != self.n or dy.shape[0] != self.m or xi_p.shape[0] != self.m:
            raise DimensionMismatch("solve!: vector lengths do not match the KKT solver")
        if not (dx.flags.c_contiguous and dy.flags.c_contiguous and dx.dtype == np.float64 and dy.dtype == np.float64):
            raise TypeError("dx, dy must be contiguous float64 arrays (they are overwritten in place)")
        lib, h = _lib.load(), self.local._h
        self._agree(lib.tlpb200_solve_begin(h, _dp(xi_p), _dp(xi_d)))
        self.dist.all_reduce(self._wk, op=self.dist.ReduceOp.SUM, group=self.group)
        self.torch.cuda.current_stream(self.device).synchronize()
        self._agree(lib.tlpb200_solve_mid(h))
        self.dist.all_reduce(self._wk, op=self.dist.ReduceOp.SUM, group=self.group)
        self.torch.cuda.current_stream(self.device).synchronize()
        self._agree(lib.tlpb200_solve_end(h, _dp(dx), _dp(dy)))

    def stats(self):
        return self.local.stats()

    def dist_info(self):
        return self.local.dist_info()

    def close(self):
        self.local.close()

    def describe(self):
        owner, off, cnt = self.dist_info()
        ntop = int(np.sum(np.diff(self.local.symbolic()["sn_first"])[owner == -1]))
        if self.mode == "library":
            return (f"collectives = NCCL all-reduce issued by the library on the solver stream inside the CUDA graph: top panels "
                    f"{cnt * 8 / 1e6:.2f} MB per update!, {ntop} separator entries ({ntop * 8 / 1e3:.1f} KB) + the solution vector "
                    f"({self._wk.numel() * 8 / 1e6:.2f} MB) per solve!; no host synchronisation between the phases")
        return (f"phase API + torch.distributed NCCL all-reduce between the phases ({cnt * 8 / 1e6:.2f} MB per update!, "
                f"{2 * 8 * self._wk.numel() / 1e6:.2f} MB per solve!), host-synchronised")

    def comm_profile(self, reps=50):
        """per-collective times of one sharded update!/solve! (timed alone, CUDA events); library mode only"""
        if self.mode != "library":
            return None
        ms = (C.c_float * 4)(); nb = (C.c_int64 * 4)()
        rc = _lib.load().tlpb200_comm_profile(self.local._h, reps, ms, nb)
        if rc != _lib.OK:
            _raise(rc, self.local._h)
        names = ("update_top_panels", "solve_separator_entries", "solve_solution_vector", "update_status_words")
        return {"unit": "ms per collective, timed alone (NCCL over NVLink, CUDA events on the solver stream)",
                **{k: {"ms": round(float(ms[i]), 4), "bytes": int(nb[i])} for i, k in enumerate(names)},
                "per_update_ms": round(float(ms[0] + ms[3]), 4), "per_solve_ms": round(float(ms[1] + ms[2]), 4)}
