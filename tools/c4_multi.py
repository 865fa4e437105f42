#!/usr/bin/env python
"""Config C4 (BASELINE.json configs[3]) for real on N GPUs: 4096 trajectories of the 20-qubit monitored brickwork circuit,
4096/N per rank as ONE batched device state, the shared draw matrix U[4096, M] sliced by rank, outcomes gathered on rank 0.
Replicas only -- no data-path communication; torch.distributed carries the final all_gather of the outcomes and the
barrier / max-over-ranks of the timing.  Launch: torchrun --nproc-per-node N tools/c4_multi.py   (N = 1 works without torchrun).

Prints one JSON line (trajectories/s over all ranks, device-timed with CUDA events on each rank's stream, max over ranks;
e2e = host wall clock from the host-resident op list and draw matrix to the gathered outcome table) and checks a sample of
trajectories per rank against the oracle's sequential per-shot loop (src/ops.jl:671-676) fed the same draws."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

bt = ge.load_package()
L = bt._lib
from importlib import import_module  # noqa: E402

wl = import_module(ge.PKG_NAME + ".workloads")


def main():
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    L.check(L.load().bt_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = 20
    total = int(os.environ.get("C4_TRAJECTORIES", "4096"))
    reps = int(os.environ.get("C4_REPS", "3"))
    check = int(os.environ.get("C4_CHECK", "8"))
    T = total // world
    specs, M = wl.c4_monitored(N, 20, 20)
    ops = wl.to_ops(bt, specs)
    U = np.random.Generator(np.random.PCG64(20)).random((total, M))  # every rank generates the same matrix and takes its rows
    mine = U[rank * T:(rank + 1) * T]
    dev_ms, host_s = [], []
    out = None
    for rep in range(reps + 1):  # first repetition = warm-up
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = bt.zero_state(N, T)
        ms = C.c_float()
        L.check(st.lib.bt_sv_timer_start(st.h))
        _, mids = bt.apply(ops, st, rng=bt.BatchDraws(mine), track_measurements=True)
        L.check(st.lib.bt_sv_timer_stop(st.h, C.byref(ms)))
        out = np.stack([np.asarray(m) for m in mids], axis=1).astype(np.int32)
        tt = torch.from_numpy(out).cuda()
        if world > 1:
            parts = [torch.empty_like(tt) for _ in range(world)]
            dist.all_gather(parts, tt)
            table = torch.cat(parts).cpu().numpy()
        else:
            table = tt.cpu().numpy()
        dt = time.perf_counter() - t0
        tm = torch.tensor([ms.value, dt * 1e3], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        if rep > 0:
            dev_ms.append(float(tm[0]))
            host_s.append(float(tm[1]) / 1e3)
        nrm = bt.norm2(st)
        ez = bt.expect(st, "Z") if rep == reps else None
        del st
    # parity: `check` trajectories of this rank against the sequential loop on the CPU oracle
    from oracle import bt_oracle as O
    from oracle import strided as S

    oops = wl.to_ops(O, specs)
    bad = 0
    for t in np.linspace(0, T - 1, check).astype(int):
        sv, mo = S.SV(N).apply_ops(oops, draws=O.ListDraws(mine[t]), track_measurements=True)
        if list(out[t]) != mo or np.max(np.abs(ez[t] - sv.expect_z_all())) > 1e-10:
            bad += 1
    tb = torch.tensor([bad, int(np.max(np.abs(nrm - 1)) > 1e-9)], device="cuda")
    if world > 1:
        dist.all_reduce(tb)
    if rank == 0:
        d, h = float(np.median(dev_ms)), float(np.median(host_s))
        line = {"metric": "trajectories/s", "value": total / (d / 1e3), "unit": "trajectories/s", "n_gpus": world, "steps": reps, "warmup": 1, "ms_per_step": d,
                "higher_is_better": True, "scaling": "strong", "dtype": "f64 (ComplexF64 amplitudes)", "data": "synthetic",
                "config": {"workload": f"C4: {total} trajectories of a {N}-qubit monitored brickwork circuit (depth 20, {len(ops) - M} gates + {M} mid-circuit measurements), "
                                       f"{T} trajectories per GPU in one batched state ({16 * T * 2 ** N / 2 ** 30:.1f} GiB)", "parallelism": f"{world} replicas, no data-path communication",
                           "l2": "inputs larger than L2"},
                "e2e": {"value": total / h, "unit": "trajectories/s", "seconds_per_step": h, "h2d_bytes_per_step": int(mine.nbytes * world), "d2h_bytes_per_step": int(table.nbytes),
                        "api": "zero_state + apply(ops, state; track_measurements) with per-trajectory draws + all_gather of the outcome table"},
                "outcome_table": list(table.shape), "mean_outcome": float(table.mean()),
                "parity": {"trajectories_checked_per_rank": check, "mismatches_all_ranks": int(tb[0]), "ranks_with_norm_error": int(tb[1]),
                           "against": "oracle sequential per-shot loop (strided C port), same uniform draws: outcome vectors identical, <Z_q> within 1e-10"}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0 if int(tb[0]) == 0 and int(tb[1]) == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
