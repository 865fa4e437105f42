#!/usr/bin/env python
"""Remap diagnostics on N >= 2 GPUs (run under torchrun; every rank writes gpurun_out/remap_diag_rank<r>.txt):
  1. pull microbenchmark: the state is remapped back and forth between two layouts with nothing in between, for every
     iteration order / synchronisation variant -> per-remap (wait, pull, done) milliseconds per rank;
  2. the C5 step free-running (no host synchronisation between steps, as in bench.py's timed loop) for the same variants,
     with the per-launch profile on and off, and with a barrier + synchronise before every step.
Knobs are environment variables the library reads at every call, so one process can sweep them."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

bt = ge.load_package()
L = bt._lib
from importlib import import_module  # noqa: E402

wl = import_module(ge.PKG_NAME + ".workloads")
D = import_module(ge.PKG_NAME + ".dist")

KNOBS = ("BT_REMAP_ORDER", "BT_REMAP_SEL_LO", "BT_REMAP_ROT", "BT_REMAP_DEVICE_SYNC", "BT_REMAP_CTAS_PER_SM")


def set_knobs(kv):
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in kv.items():
        os.environ[k] = str(v)


def remap_log(st):
    buf = (C.c_float * (3 * 256))()
    n = C.c_int()
    L.check(st.lib.bt_sv_remap_log(st.h, 256, buf, C.byref(n)))
    return [(round(buf[3 * i], 2), round(buf[3 * i + 1], 2), round(buf[3 * i + 2], 2)) for i in range(n.value)]


def main():
    rank, world, local = D.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = world.bit_length() - 1
    n_local = int(os.environ.get("DIAG_SHARD_QUBITS", "31"))
    steps = int(os.environ.get("DIAG_STEPS", "4"))
    N = n_local + g
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", f"remap_diag_rank{rank}.txt"), "w")

    def say(msg):
        out.write(msg + "\n")
        out.flush()
        if rank == 0:
            print(msg, flush=True)

    st = D.ShardedState(N)
    lib = st.lib
    shard_gb = 16.0 * (1 << n_local) / 1e9
    say(f"# world={world} N={N} n_local={n_local} shard={shard_gb:.1f} GB rank={rank}")

    # ---- 1. pull microbenchmark --------------------------------------------------------------------------------------
    ident = list(range(N))
    swapped = ident[:]
    for j in range(g):  # global bit j <-> local bit n_local-1-j (the planner's choice of victims)
        swapped[n_local + j], swapped[n_local - 1 - j] = swapped[n_local - 1 - j], swapped[n_local + j]
    lay_a = (C.c_int * N)(*swapped)
    lay_b = (C.c_int * N)(*ident)
    variants = [
        ("linear walk, device flags", {"BT_REMAP_ORDER": 0}),
        ("interleaved sel_lo=8 rot, device flags", {"BT_REMAP_ORDER": 1, "BT_REMAP_SEL_LO": 8}),
        ("interleaved sel_lo=16 rot (default)", {}),
        ("interleaved sel_lo=16 rot, host barriers", {"BT_REMAP_DEVICE_SYNC": 0}),
        ("interleaved sel_lo=16 no rot", {"BT_REMAP_ROT": 0}),
        ("interleaved sel_lo=20 rot", {"BT_REMAP_SEL_LO": 20}),
        ("interleaved sel_lo=16 rot, 4 CTAs/SM", {"BT_REMAP_CTAS_PER_SM": 4}),
    ]
    if os.environ.get("DIAG_QUICK", "") not in ("", "0"):
        variants = variants[:3]
    remote_gb = shard_gb * (1.0 - 2.0 ** -g)
    for name, kv in variants:
        set_knobs(kv)
        dist.barrier()
        torch.cuda.synchronize()
        remap_log(st)
        t0 = time.perf_counter()
        for k in range(4):
            L.check(lib.bt_sv_remap(st.h, lay_a if k % 2 == 0 else lay_b))
        st.sync()
        dt = time.perf_counter() - t0
        lg = remap_log(st)
        pulls = [p for (_, p, _) in lg]
        say(f"[pull] {name:45s} wall {dt * 1e3 / 4:7.1f} ms/remap  pull ms {pulls}  -> {remote_gb / (np.median(pulls) / 1e3):6.0f} GB/s remote per rank (median)  (wait,pull,done) {lg}")

    # ---- 2. the C5 step ------------------------------------------------------------------------------------------------
    specs = wl.c5_random(N, 20, 31)
    arr = bt.pack_gates(wl.to_ops(bt, specs))
    ngates = len(arr)
    os.environ.setdefault("BT_TILE_JIT_AFTER", "1")

    def step():
        L.check(lib.bt_sv_set_basis(st.h, 0))
        L.check(lib.bt_sv_apply_circuit(st.h, L.ptr(arr), ngates, 1))

    set_knobs({})
    t0 = time.perf_counter()
    step()
    st.sync()
    t1 = time.perf_counter()
    L.check(lib.bt_jit_wait(None))  # the first step ran on the interpreter while the workers compiled
    t2 = time.perf_counter()
    for _ in range(2):
        step()
    st.sync()
    say(f"[step] warm-up: first step (interpreter, compile in the background) {t1 - t0:.2f} s, waiting for the compile workers {t2 - t1:.2f} s, 2 more steps {time.perf_counter() - t2:.2f} s")
    step_variants = [
        ("linear walk, device flags, profile on", {"BT_REMAP_ORDER": 0}, True, False),
        ("interleaved sel_lo=8, device flags, profile on", {"BT_REMAP_SEL_LO": 8}, True, False),
        ("interleaved sel_lo=16 (default), device flags, profile on", {}, True, False),
        ("interleaved sel_lo=16 (default), device flags, profile off", {}, False, False),
        ("interleaved sel_lo=16, host barriers", {"BT_REMAP_DEVICE_SYNC": 0}, False, False),
    ]
    if os.environ.get("DIAG_QUICK", "") not in ("", "0"):
        step_variants = step_variants[:3]
    for name, kv, prof, barrier in step_variants:
        set_knobs(kv)
        dist.barrier()
        torch.cuda.synchronize()
        remap_log(st)
        L.check(lib.bt_sv_profile_enable(st.h, 1 if prof else 0))
        ms = C.c_float()
        th = time.perf_counter()
        L.check(lib.bt_sv_timer_start(st.h))
        enq = []
        for _ in range(steps):
            if barrier:
                dist.barrier()
                torch.cuda.synchronize()
            te = time.perf_counter()
            step()
            enq.append(round((time.perf_counter() - te) * 1e3, 1))
        L.check(lib.bt_sv_timer_stop(st.h, C.byref(ms)))
        host = time.perf_counter() - th
        counts = (C.c_uint64 * 4)()
        cms = (C.c_double * 4)()
        if prof:
            L.check(lib.bt_sv_profile_read(st.h, counts, cms))
        L.check(lib.bt_sv_profile_enable(st.h, 0))
        lg = remap_log(st)
        say(f"[step] {name:55s} {ms.value / steps:7.1f} ms/step (device)  host {host * 1e3 / steps:7.1f} ms/step  enqueue {enq}  fused {cms[0] / steps:6.1f} ms/step  (wait,pull,done) {lg}")
    nrm = np.empty(1)
    L.check(lib.bt_sv_norm2(st.h, L.pdouble(nrm)))
    say(f"[step] norm2 {nrm[0]:.12f}")
    out.close()
    dist.barrier()
    del st
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
