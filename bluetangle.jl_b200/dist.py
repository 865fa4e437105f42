"""Multi-GPU plumbing: one process per GPU, torch.distributed for the control plane only.

The data path never goes through torch or NCCL: every rank maps its peers' shard buffers with CUDA IPC
(bt_sv_ipc_export / bt_sv_ipc_attach) and the library's remap kernel pulls amplitudes straight over NVLink.
torch.distributed supplies (i) the all-gather of the three 64-byte IPC handles, (ii) the barrier that brackets a remap,
(iii) the all-reduce of a handful of doubles for reductions -- exactly the three callbacks of include/bluetangle_cuda.h.
A Julia deployment would use MPI.jl for the same three calls.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import _lib as L


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


class ShardedState:
    """One shard (2^(N - log2 P) amplitudes) of an N-qubit state distributed over P ranks."""

    def __init__(self, n_qubits: int, group=None, device_index: Optional[int] = None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.N = n_qubits
        self.n_batch = 1
        self.lib = L.load()
        if device_index is None:
            device_index = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = torch.device("cuda", device_index)
        torch.cuda.set_device(self.device)
        L.check(self.lib.bt_set_device(device_index))
        h = C.c_void_p()
        L.check(self.lib.bt_sv_create_shard(n_qubits, self.rank, self.world, C.byref(h)))
        self.h = h
        self._cb_barrier = L.BARRIER_FN(self._barrier)
        self._cb_allreduce = L.ALLREDUCE_FN(self._allreduce)
        L.check(self.lib.bt_sv_set_barrier(self.h, self._cb_barrier, None))
        L.check(self.lib.bt_sv_set_allreduce(self.h, self._cb_allreduce, None))
        if self.world > 1:
            mine = np.zeros(3 * 64, dtype=np.uint8)  # BT_IPC_HANDLES_PER_SHARD x BT_IPC_HANDLE_BYTES: both buffers + the flag page
            L.check(self.lib.bt_sv_ipc_export(self.h, L.ptr(mine)))
            t = torch.from_numpy(mine).to(self.device)
            out = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(out, t, group=group)
            allh = np.concatenate([o.cpu().numpy() for o in out]).astype(np.uint8)
            L.check(self.lib.bt_sv_ipc_attach(self.h, L.ptr(allh)))
            dist.barrier(group=group)

    # callbacks (called from C on the host thread) -------------------------------------------------------------
    def _barrier(self, ctx):
        if self.world > 1:
            self.dist.barrier(group=self.group)

    def _allreduce(self, ctx, buf, n):
        if self.world == 1:
            return
        a = np.ctypeslib.as_array(buf, shape=(n,))
        t = self.torch.from_numpy(a.copy()).to(self.device)
        self.dist.all_reduce(t, group=self.group)
        a[:] = t.cpu().numpy()

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.bt_sv_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def sync(self):
        L.check(self.lib.bt_sv_sync(self.h))

    def launch_count(self) -> int:
        n = C.c_uint64()
        L.check(self.lib.bt_sv_launch_count(self.h, C.byref(n)))
        return int(n.value)

    def remap_stats(self):
        n, b, ms = C.c_uint64(), C.c_uint64(), C.c_float()
        L.check(self.lib.bt_sv_remap_stats(self.h, C.byref(n), C.byref(b), C.byref(ms)))
        return int(n.value), int(b.value), float(ms.value)

    def local_numpy(self) -> np.ndarray:
        out = np.empty(1 << (self.N - (self.world.bit_length() - 1)), dtype=np.complex128)
        L.check(self.lib.bt_sv_download(self.h, L.ptr(out), out.size))
        return out

    def layout(self):
        arr = (C.c_int * self.N)()
        L.check(self.lib.bt_sv_layout(self.h, arr))
        return list(arr)

    def gather_logical(self) -> np.ndarray:
        """Full state in logical index order on every rank (tests only: small N)."""
        loc = self.local_numpy()
        t = self.torch.from_numpy(np.stack([loc.real, loc.imag])).to(self.device)
        outs = [self.torch.empty_like(t) for _ in range(self.world)]
        if self.world > 1:
            self.dist.all_gather(outs, t, group=self.group)
        else:
            outs = [t]
        phys = np.concatenate([(o[0] + 1j * o[1]).cpu().numpy() for o in outs])
        lay = self.layout()
        idx = np.arange(1 << self.N, dtype=np.uint64)
        pidx = np.zeros_like(idx)
        for lb in range(self.N):
            pidx |= ((idx >> np.uint64(lb)) & np.uint64(1)) << np.uint64(lay[lb])
        return phys[pidx]


class LocalShards:
    """All P shards of a state inside ONE process (possibly on one device): the single-process counterpart of
    ShardedState, used by a single-threaded host (e.g. one Julia process driving several GPUs) and by the 1-GPU tests of
    the remap logic.  Every call is issued shard by shard in rank order (SPMD executed sequentially)."""

    def __init__(self, n_qubits: int, world: int, devices=None):
        self.lib = L.load()
        self.N, self.world = n_qubits, world
        self.g = world.bit_length() - 1
        self.hs = []
        for r in range(world):
            if devices is not None:
                L.check(self.lib.bt_set_device(devices[r]))
            h = C.c_void_p()
            L.check(self.lib.bt_sv_create_shard(n_qubits, r, world, C.byref(h)))
            self.hs.append(h)
        arr = (C.c_void_p * world)(*[h.value for h in self.hs])
        L.check(self.lib.bt_sv_attach_local_peers(arr, world))

    def __del__(self):
        try:
            for h in getattr(self, "hs", []):
                self.lib.bt_sv_destroy(h)
            self.hs = []
        except Exception:
            pass

    def each(self, fn):
        for h in self.hs:
            L.check(fn(h))

    def apply(self, op):
        if op.q == 1:
            m = L.cmat(op.mat, 2)
            self.each(lambda h: self.lib.bt_sv_apply_1q(h, op.qubit, L.ptr(m), op.control))
        else:
            m = L.cmat(op.mat, 4)
            self.each(lambda h: self.lib.bt_sv_apply_2q(h, op.qubit, op.target_qubit, L.ptr(m), op.control))

    def apply_circuit(self, arr: np.ndarray, fuse: int = 1):
        hs = (C.c_void_p * self.world)(*[h.value for h in self.hs])
        L.check(self.lib.bt_group_apply_circuit(hs, self.world, L.ptr(arr), len(arr), fuse))

    def set_basis(self, index: int):
        self.each(lambda h: self.lib.bt_sv_set_basis(h, index))

    def layout(self, r: int = 0):
        arr = (C.c_int * self.N)()
        L.check(self.lib.bt_sv_layout(self.hs[r], arr))
        return list(arr)

    def remap_stats(self, r: int = 0):
        n, b, ms = C.c_uint64(), C.c_uint64(), C.c_float()
        L.check(self.lib.bt_sv_remap_stats(self.hs[r], C.byref(n), C.byref(b), C.byref(ms)))
        return int(n.value), int(b.value), float(ms.value)

    def upload_logical(self, vec: np.ndarray):
        """Distribute a full logical state (identity layout)."""
        nl = self.N - self.g
        v = np.ascontiguousarray(vec, dtype=np.complex128)
        for r, h in enumerate(self.hs):
            part = np.ascontiguousarray(v[r << nl:(r + 1) << nl])
            L.check(self.lib.bt_sv_upload(h, L.ptr(part), part.size))

    def gather_logical(self) -> np.ndarray:
        nl = self.N - self.g
        parts = []
        for h in self.hs:
            out = np.empty(1 << nl, dtype=np.complex128)
            L.check(self.lib.bt_sv_download(h, L.ptr(out), out.size))
            parts.append(out)
        phys = np.concatenate(parts)
        lay = self.layout(0)
        for r in range(1, self.world):
            assert self.layout(r) == lay, "shards disagree on the layout"
        idx = np.arange(1 << self.N, dtype=np.uint64)
        pidx = np.zeros_like(idx)
        for lb in range(self.N):
            pidx |= ((idx >> np.uint64(lb)) & np.uint64(1)) << np.uint64(lay[lb])
        return phys[pidx]

    def partial(self, fn_name: str, *args, n_out: int):
        """Sum of the per-shard partial results of a reduction (no all-reduce callback in single-process mode)."""
        tot = np.zeros(n_out)
        for h in self.hs:
            out = np.zeros(n_out)
            L.check(getattr(self.lib, fn_name)(h, *args, L.pdouble(out)))
            tot += out
        return tot


def plan_circuit(n_qubits: int, world: int, gate_array: np.ndarray, cap: int = 4096):
    """The library's segment plan for `world` shards (pure host logic, no device): list of dicts with the layout
    (physical position of every logical bit), whether a remap precedes the segment, and the gate indices it runs."""
    lib = L.load()
    nseg = C.c_int()
    seg_g = (C.c_int * cap)()
    seg_r = (C.c_int * cap)()
    lay = (C.c_int * (cap * n_qubits))()
    order = (C.c_int * max(1, len(gate_array)))()
    L.check(lib.bt_plan_circuit_host(n_qubits, world, L.ptr(gate_array), len(gate_array), C.byref(nseg), seg_g, seg_r, lay, order, cap))
    out, pos = [], 0
    for i in range(nseg.value):
        n = seg_g[i]
        out.append({"remap": bool(seg_r[i]), "layout": [lay[i * n_qubits + b] for b in range(n_qubits)], "gates": [order[pos + j] for j in range(n)]})
        pos += n
    return out
