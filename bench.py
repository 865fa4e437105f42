#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 backend (contract: see task prompt / DESIGN.md section "Measurement").

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # the reference path's CPU port on the host cores

N = 1: workload C2 of BASELINE.json configs[1] -- 28-qubit state vector, QFT(28) + 100 random layers (4,556 gates),
       ComplexF64.  One step = reset to |0..0> + the whole circuit.
N > 1: workload C5 -- (31 + log2 N)-qubit random circuit sharded over N ranks (torchrun, one rank per GPU); value is
       reported in 28-qubit-equivalent gates/s (gates x local amplitudes / 2^28, summed over ranks) so that the
       numbers at different N measure the same per-GPU work ("weak" scaling).  `unit` is "gates/s" at every N
       (the normalisation is spelled out in `unit_note`).
metric: gates/s.  `value` is device-resident throughput (CUDA events on the launching stream); `e2e` goes through
the public host API with host buffers (gate list in, expectation values and samples out).
"""
from __future__ import annotations

import argparse
import ctypes as C
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region.  The process is started BEFORE the warm-up
    (nvidia-smi's own start-up initialises NVML over every GPU of the box and can hold up CUDA calls of the ranks for
    hundreds of milliseconds -- that belongs into the warm-up, not into a timed step); only the samples that arrive
    between begin() and stop() are reported."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int = 0):
        self.index = index
        self.rows = []
        self.p = None
        self.t_begin = None

    def start(self):
        try:
            if os.environ.get("BENCH_NO_CLOCKS", "") not in ("", "0"):
                raise RuntimeError("clock sampling switched off")
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", os.environ.get("BENCH_CLOCKS_MS", "100")],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def begin(self):
        """the timed region starts now"""
        self.t_begin = time.perf_counter()

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.perf_counter()
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            pass
        t0 = self.t_begin if self.t_begin is not None else 0.0
        rows = [r for (t, r) in self.rows if t0 <= t <= t_end + 0.12]
        if not rows:  # region shorter than one sampling period: take the nearest samples
            rows = [r for (_, r) in self.rows[-2:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
def _timed_sv(L, lib, h, fn, reps, warm=2):
    for _ in range(warm):
        fn()
    t = C.c_float()
    L.check(lib.bt_sv_timer_start(h))
    for _ in range(reps):
        fn()
    L.check(lib.bt_sv_timer_stop(h, C.byref(t)))
    return t.value / reps


def kraus_block(bt, L, s, N, peak):
    """SV Kraus trajectory steps (__QuantumChannel_new_apply, src/struct.jl:31-41) at N qubits: one step = RDM read
    (16 B/amp) + decision + scaled-Kraus apply (32 B/amp) = 48 B x 2^N algorithmic bytes."""
    lib = s.lib
    K1 = np.stack([np.asfortranarray(k).reshape(-1, order="F") for k in bt.noise_model("depolarizing", 0.01)])
    K2 = np.stack([np.asfortranarray(k).reshape(-1, order="F") for k in bt.noise_model("depolarizing", 0.01, True)])
    A1 = np.stack([np.asfortranarray(k).reshape(-1, order="F") for k in bt.noise_model("amplitude_damping", 0.02)])
    ua = np.array([0.5])
    out = {}
    byts = 48.0 * (1 << N)
    L.check(lib.bt_sv_set_plus(s.h))
    cases = [("sv_1q_depolarizing_q%d" % (N // 2), lambda: L.check(lib.bt_sv_kraus(s.h, 1, N // 2, -1, L.ptr(K1), 4, L.pdouble(ua), None))),
             ("sv_1q_depolarizing_q%d" % N, lambda: L.check(lib.bt_sv_kraus(s.h, 1, N, -1, L.ptr(K1), 4, L.pdouble(ua), None))),
             ("sv_1q_amplitude_damping_q3", lambda: L.check(lib.bt_sv_kraus(s.h, 1, 3, -1, L.ptr(A1), 2, L.pdouble(ua), None))),
             ("sv_2q_depolarizing_q3_q17", lambda: L.check(lib.bt_sv_kraus(s.h, 2, 3, min(17, N), L.ptr(K2), 16, L.pdouble(ua), None))),
             ("sv_2q_depolarizing_q%d_q%d" % (N - 1, N), lambda: L.check(lib.bt_sv_kraus(s.h, 2, N - 1, N, L.ptr(K2), 16, L.pdouble(ua), None)))]
    for name, fn in cases:
        ms = _timed_sv(L, lib, s.h, fn, 5)
        g = byts / (ms / 1e3) / 1e9
        out[name] = {"ms": ms, "GBps": g, "frac_of_measured": g / peak, "frac_of_8TBs_spec": g / 8000.0}
    out["algorithmic_bytes"] = "48 B x 2^N per step: RDM read 16 B/amp + scaled-Kraus apply 32 B/amp"
    return out


def c4_block(bt, L, wl, n=20, trajectories=512, reps=3):
    """Config C4 (BASELINE.json configs[3]) on this GPU: `trajectories` monitored 20-qubit brickwork trajectories as one batched state
    (the per-shot loop of src/ops.jl:616-631, 671-676), per-trajectory draws from a host matrix, mid-circuit outcomes returned to
    the host.  Replicas only: N GPUs run N such batches (tools/c4_multi.py is the multi-rank driver with the oracle comparison)."""
    specs, M = wl.c4_monitored(n, 20, 20)
    ops = wl.to_ops(bt, specs)
    U = np.random.Generator(np.random.PCG64(20)).random((trajectories, M))
    dev, host, table = [], [], None
    for rep in range(reps + 1):  # first repetition = warm-up (specialised passes compile in the background)
        t0 = time.perf_counter()
        st = bt.zero_state(n, trajectories)
        ms = C.c_float()
        L.check(st.lib.bt_sv_timer_start(st.h))
        _, mids = bt.apply(ops, st, rng=bt.BatchDraws(U), track_measurements=True)
        L.check(st.lib.bt_sv_timer_stop(st.h, C.byref(ms)))
        table = np.stack([np.asarray(m) for m in mids], axis=1)
        dt = time.perf_counter() - t0
        if rep == 0:
            L.check(st.lib.bt_jit_wait(None))
        else:
            dev.append(ms.value)
            host.append(dt)
        nrm = bt.norm2(st)
        del st
    d, h = float(np.median(dev)), float(np.median(host))
    return {"workload": f"C4: {trajectories} trajectories of a {n}-qubit monitored brickwork circuit (depth 20, {len(ops) - M} gates + {M} mid-circuit measurements) in one batched state "
                        f"({16 * trajectories * 2 ** n / 2 ** 30:.1f} GiB)", "ms": d, "trajectories_per_s": trajectories / (d / 1e3), "e2e_trajectories_per_s": trajectories / h,
            "outcome_table": list(table.shape), "mean_outcome": float(table.mean()), "max_norm_error": float(np.max(np.abs(nrm - 1))),
            "note": "runs of measurements share one read + one collapse pass; outcomes are logged on the device and read once"}


def dm14_block(bt, L, wl, peak, n=14, depth=20):
    """Config C3 (BASELINE.json configs[2]): n-qubit density matrix, depolarizing + amplitude damping after every gate
    (to_rho's loop, src/ops.jl:813-841 with apply(rho,op) src/hilbert.jl:655-656 and the channels of src/struct.jl:58-76),
    fused superoperators vs op by op, plus the single-kernel rooflines (32 B x 4^n per pass)."""
    ops3 = []
    for e in wl.c3_noisy_dm(n, depth, 14):
        if e[0] == "gate":
            name, q, t, c = e[1]
            ops3.append(bt.Op(name, q, t, control=c))
        else:
            _, model, p, q, t = e
            ops3.append(bt.OpQC(model, p, q, t))
    ngate = sum(1 for o in ops3 if isinstance(o, bt.Op))
    out = {"workload": f"C3: {n}-qubit density matrix, C1-style brickwork depth {depth}, depolarizing(0.01) + amplitude_damping(0.02) after every gate: {ngate} gates + {len(ops3) - ngate} Kraus channels"}
    pass_bytes = 32.0 * 4 ** n
    for fused in (True, False):
        best = None
        for _ in range(2):
            rho = bt.CuRho(n)
            rho.sync()
            ms = C.c_float()
            n0 = rho.launch_count()
            L.check(rho.lib.bt_dm_timer_start(rho.h))
            if fused:
                bt.apply(ops3, rho)
            else:
                for o in ops3:
                    bt.apply(rho, o)
            L.check(rho.lib.bt_dm_timer_stop(rho.h, C.byref(ms)))
            nl = rho.launch_count() - n0
            tr = L.bt_c64()
            L.check(rho.lib.bt_dm_trace(rho.h, C.byref(tr)))
            if best is None or ms.value < best["ms"]:
                g = pass_bytes * nl / (ms.value / 1e3) / 1e9
                best = {"ms": ms.value, "passes": int(nl), "ops_per_s": len(ops3) / (ms.value / 1e3), "gates_per_s": ngate / (ms.value / 1e3),
                        "GBps": g, "frac_of_measured": g / peak, "frac_of_8TBs_spec": g / 8000.0, "trace": tr.re}
            del rho
        out["fused_superoperators" if fused else "op_by_op"] = best
    # single kernels on rho
    d = bt.CuRho(n)
    dl = d.lib

    def timed_dm(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        ms = C.c_float()
        L.check(dl.bt_dm_timer_start(d.h))
        for _ in range(reps):
            fn()
        L.check(dl.bt_dm_timer_stop(d.h, C.byref(ms)))
        return ms.value / reps

    H = L.cmat(bt.gate["H"], 2)
    U4 = L.cmat(bt.gates("FSIM(0.3,0.2)"), 4)
    K1 = np.stack([np.asfortranarray(k).reshape(-1, order="F") for k in bt.noise_model("depolarizing", 0.01)])
    A1 = np.stack([np.asfortranarray(k).reshape(-1, order="F") for k in bt.noise_model("amplitude_damping", 0.02)])
    K2 = np.stack([np.asfortranarray(k).reshape(-1, order="F") for k in bt.noise_model("depolarizing", 0.01, True)])
    kern = {}
    for name, fn in (("dm_1q_unitary_q%d" % (n // 2), lambda: L.check(dl.bt_dm_apply_1q(d.h, n // 2, L.ptr(H), -2))),
                     ("dm_2q_unitary_q2_q9", lambda: L.check(dl.bt_dm_apply_2q(d.h, 2, min(9, n), L.ptr(U4), -2))),
                     ("dm_1q_depolarizing_q%d" % n, lambda: L.check(dl.bt_dm_kraus(d.h, 1, n, -1, L.ptr(K1), 4))),
                     ("dm_1q_amplitude_damping_q1", lambda: L.check(dl.bt_dm_kraus(d.h, 1, 1, -1, L.ptr(A1), 2))),
                     ("dm_2q_depolarizing_16_kraus_q%d_q%d" % (n - 1, n), lambda: L.check(dl.bt_dm_kraus(d.h, 2, n - 1, n, L.ptr(K2), 16)))):
        ms = timed_dm(fn)
        g = pass_bytes / (ms / 1e3) / 1e9
        kern[name] = {"ms": ms, "GBps": g, "frac_of_measured": g / peak, "frac_of_8TBs_spec": g / 8000.0}
    out["kernels"] = kern
    out["algorithmic_bytes"] = f"32 B x 4^{n} = {pass_bytes / 1e9:.2f} GB per unitary or channel (one pass whatever the number of Kraus operators)"
    del d
    return out


def fp64_peak(lib, L, clocks_index=0):
    """FP64 FMA peak measured in this run (bt_fp64_peak: DFMA loop on every SM, best of 5), with the SM clock nvidia-smi reports
    right after it -- a measured, clock-stamped denominator instead of a constant in the source."""
    tf, ms = C.c_double(), C.c_double()
    L.check(lib.bt_fp64_peak(C.byref(tf), C.byref(ms), 5))
    sm = None
    try:
        r = subprocess.run(["nvidia-smi", "-i", str(clocks_index), "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10)
        f = [x.strip() for x in r.stdout.strip().split(",")]
        sm = {"sm_mhz_after": float(f[0]), "sm_max_mhz": float(f[1])}
    except Exception:
        pass
    return {"tflops": tf.value, "ms_per_launch": ms.value, "clock": sm, "how": "bt_fp64_peak: 8 independent DFMA chains per thread, 8 x 256 threads per SM, register operands, CUDA events, best of 5 (measured in this run)"}


def bench_single(args):
    import __graft_entry__ as ge

    bt = ge.load_package()
    L = bt._lib
    from importlib import import_module

    wl = import_module(ge.PKG_NAME + ".workloads")
    N, depth = args.qubits, args.depth
    specs = wl.c2_qft_layered(N, depth, 28)
    ops = wl.to_ops(bt, specs)
    arr = bt.pack_gates(ops)  # host buffer: what crosses the C ABI every step
    ngates = len(arr)
    s = bt.zero_state(N)
    lib = s.lib

    def step():
        L.check(lib.bt_sv_set_basis(s.h, 0))
        L.check(lib.bt_sv_apply_circuit(s.h, L.ptr(arr), ngates, 1))

    clocks = ClockSampler(0)
    clocks.start()
    # ---- cold start: the very first execution of the circuit in this process.  The library hands every fused pass to its
    # compile workers (NVRTC, or the on-disk cubin cache of an earlier process) and runs the pass on the interpreter meanwhile.
    s.sync()
    t_w0 = time.perf_counter()
    step()
    s.sync()
    cold_first_step_s = time.perf_counter() - t_w0
    pend = C.c_uint64()
    L.check(lib.bt_jit_wait(C.byref(pend)))
    jit_ready_s = time.perf_counter() - t_w0
    for _ in range(args.warmup):
        step()
    s.sync()
    warmup_seconds = time.perf_counter() - t_w0
    time.sleep(max(0.0, 1.5 - warmup_seconds))  # nvidia-smi has finished starting up before the timed region
    gc.collect()
    gc.disable()  # no cyclic-GC pause inside the timed region (a full collection with torch imported takes ~0.5 s)
    clocks.begin()
    L.check(lib.bt_sv_profile_enable(s.h, 1))
    fl0 = C.c_double()
    L.check(lib.bt_fusion_flops(C.byref(fl0)))
    n0 = s.launch_count()
    jit0 = jit_stats(lib)["specialised_launches"]
    ms = C.c_float()
    L.check(lib.bt_sv_timer_start(s.h))
    for _ in range(args.steps):
        step()
    L.check(lib.bt_sv_timer_stop(s.h, C.byref(ms)))
    n1 = s.launch_count()
    jit_timed = jit_stats(lib)["specialised_launches"] - jit0
    fl1 = C.c_double()
    L.check(lib.bt_fusion_flops(C.byref(fl1)))
    counts = (C.c_uint64 * 4)()
    cms = (C.c_double * 4)()
    L.check(lib.bt_sv_profile_read(s.h, counts, cms))
    L.check(lib.bt_sv_profile_enable(s.h, 0))
    clk = clocks.stop()
    gc.enable()
    ms_per_step = ms.value / args.steps
    value = ngates / (ms_per_step / 1e3)
    norm = bt.norm2(s)

    # roofline of the dominant kernel (fused tile kernel): algorithmic bytes = 32 B x 2^N per launch
    peaks, peak_src = load_peaks()
    cls_names = ["tile", "dense", "diag", "other"]
    per_cls = {cls_names[i]: {"launches": int(counts[i]), "ms": float(cms[i])} for i in range(4)}
    dom = max(range(4), key=lambda i: cms[i])
    bytes_per_launch = 32.0 * (1 << N)
    fp64 = fp64_peak(lib, L)
    roof = None
    if counts[dom] > 0:
        avg_ms = cms[dom] / counts[dom]
        ach = bytes_per_launch / (avg_ms / 1e3) / 1e9
        if dom == 0 and 2 * jit_timed >= counts[0]:
            kname = f"bt_jit_pass (fused multi-gate pass specialised by NVRTC: TMA tile in/out + straight-line register programs; {jit_timed} of {int(counts[0])} fused launches, the rest k_tile_tma)"
        elif dom == 0:
            kname = "k_tile_tma (fused multi-gate pass: TMA tile in/out + interpreted register programs)"
        else:
            kname = cls_names[dom]
        roof = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src, "avg_launch_ms": avg_ms,
                "share_of_step": cms[dom] / (ms.value), "bytes_per_launch": bytes_per_launch, "frac_of_8TBs_spec": ach / 8000.0}
        if dom == 0 and cms[0] > 0:
            tf = (fl1.value - fl0.value) / (cms[0] / 1e3) / 1e12
            roof["fp64"] = {"achieved_tflops": tf, "peak_tflops_measured": fp64["tflops"], "frac": tf / fp64["tflops"], "peak": fp64,
                            "note": "useful FP64 flops of the structured micro-ops (real / RX-like / diagonal gates cost half of a dense 2x2, CX none)"}
            # issued FP64 instructions: counted from the text the specialiser generates for exactly these passes (planner dry run in
            # a subprocess with this process's code shape / reshaping flags; tools/jit_fp64_count.py, checked against cuobjdump -sass
            # in profiles/r2_jit_fp64_counts.txt).  One DFMA / DADD / DMUL each takes one issue slot of the FP64 pipe: peak = TF/s / 2.
            try:
                if 2 * jit_timed >= counts[0]:
                    sys.path.insert(0, os.path.join(ROOT, "tools"))
                    import jit_fp64_count as J
                    jc = [C.c_int() for _ in range(4)]
                    L.check(lib.bt_jit_config(*[C.byref(x) for x in jc]))
                    est = J.source_estimate(f"{N}, wl.c2_qft_layered({N}, {depth}, 28)", {"BT_JIT_VARIANT": str(jc[2].value), "BT_JIT_OPT": str(jc[3].value)})
                    issued = est["executed_per_amplitude"] * float(1 << N) * args.steps / (cms[0] / 1e3) / 1e12
                    roof["fp64"]["issued"] = {"instr_per_amplitude_per_step": est["executed_per_amplitude"], "static_instr_per_amplitude_per_step": est["static_per_amplitude"],
                                              "tera_instr_per_s": issued, "peak_tera_instr_per_s": fp64["tflops"] / 2.0, "frac": issued / (fp64["tflops"] / 2.0),
                                              "note": "DFMA + DADD + DMUL the specialised passes execute (each one FP64-pipe issue slot) / fused-kernel time / (measured FMA peak / 2)"}
            except Exception as e:  # a diagnostic: never at the expense of the line
                roof["fp64"]["issued"] = {"error": repr(e)[:300]}
            roof["gates_per_launch"] = ngates * args.steps / max(1, int(counts[0]))
            # a fused pass carries tens of gates per HBM round trip, so it sits between the two roofs; the look-ahead scheduler (round 2)
            # deliberately trades HBM fraction for fewer passes: the step gets faster while bytes / launch-time drops
            roof["binding_roof"] = "fp64" if roof["fp64"]["frac"] > roof["frac"] else "hbm"
            roof["note"] = ("frac is the HBM fraction of the dominant kernel (contract: algorithmic 32 B x 2^N per launch / launch time / measured copy peak); "
                            "the kernel is bound by the roof named in binding_roof -- see fp64.frac for the FP64 side")
        for tname in ("traffic_r2.json", "traffic_r1.json"):
            tfile = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tfile):
                try:
                    tj = json.load(open(tfile))
                    roof["traffic"] = tj.get("bt_jit_pass_dram_bytes_per_launch" if 2 * jit_timed >= counts[0] else "k_tile_dram_bytes_per_launch", tj.get("k_tile_dram_bytes_per_launch"))
                    roof["traffic_source"] = "profiles/" + tname
                    break
                except Exception:
                    pass

    # the same circuit on the interpreter alone (what a circuit that is executed once costs): specialiser off
    os.environ["BT_TILE_JIT"] = "0"
    step()
    mi = C.c_float()
    L.check(lib.bt_sv_timer_start(s.h))
    for _ in range(2):
        step()
    L.check(lib.bt_sv_timer_stop(s.h, C.byref(mi)))
    del os.environ["BT_TILE_JIT"]
    interp = {"ms_per_step": mi.value / 2, "gates_per_s": ngates / (mi.value / 2 / 1e3),
              "hbm_equivalent_GBps": None, "note": "BT_TILE_JIT=0: every fused pass on the pre-compiled interpreter kernel (k_tile_tma)"}
    if counts[0] > 0:
        passes_per_step = counts[0] / args.steps
        interp["hbm_equivalent_GBps"] = bytes_per_launch * passes_per_step / (mi.value / 2 / 1e3) / 1e9
        interp["frac_of_measured"] = interp["hbm_equivalent_GBps"] / peaks["hbm_gbs"]

    # unfused single-gate kernels on the same state: the per-gate HBM roofline the north star quotes
    micro = {}
    for name, fn in (("1q_dense_H_q14", lambda: L.check(lib.bt_sv_apply_1q(s.h, N // 2, L.ptr(L.cmat(bt.gate["H"], 2)), -2))),
                     ("2q_dense_q3_q17", lambda: L.check(lib.bt_sv_apply_2q(s.h, 3, min(17, N), L.ptr(L.cmat(bt.gates("FSIM(0.3,0.2)"), 4)), -2)))):
        t = _timed_sv(L, lib, s.h, fn, 10, warm=3)
        g = bytes_per_launch / (t / 1e3) / 1e9
        micro[name] = {"ms": t, "GBps": g, "frac_of_measured": g / peaks["hbm_gbs"], "frac_of_8TBs_spec": g / 8000.0}

    # end to end through the host API: gate list in (host), <Z_q> for every qubit and 4096 samples out (host)
    us = np.random.Generator(np.random.PCG64(7)).random(4096)
    e2e_t = []
    ez = None
    for _ in range(max(1, min(args.steps, 3))):
        t0 = time.perf_counter()
        st = s
        L.check(lib.bt_sv_set_basis(st.h, 0))
        L.check(lib.bt_sv_apply_circuit(st.h, L.ptr(arr), ngates, 1))
        ez = bt.expect(st, "Z")
        smp = bt.sample(st, 4096, uniforms=us)
        e2e_t.append(time.perf_counter() - t0)
    e2e_s = float(np.median(e2e_t))
    e2e = {"value": ngates / e2e_s, "unit": "gates/s", "h2d_bytes_per_step": int(arr.nbytes + us.nbytes), "d2h_bytes_per_step": int(ez.nbytes + smp.nbytes + 8),
           "seconds_per_step": e2e_s, "api": "bt_sv_set_basis + bt_sv_apply_circuit(host gate list) + bt_sv_expect_1q_all + bt_sv_sample"}

    # the other half of BASELINE's metric, driver-run: Kraus trajectory steps at N qubits, the 14-qubit density matrix (C3)
    extra_err = {}
    try:
        kraus = kraus_block(bt, L, s, N, peaks["hbm_gbs"]) if not args.no_blocks else None
    except Exception as e:  # the headline must not be lost to a side block
        kraus, extra_err["kraus"] = None, repr(e)
    jitst = jit_stats(lib, warmup_seconds)
    hits, cdir = C.c_uint64(), C.create_string_buffer(512)
    L.check(lib.bt_jit_cache_info(C.byref(hits), cdir, 512))
    cfg = [C.c_int() for _ in range(4)]
    L.check(lib.bt_jit_config(*[C.byref(x) for x in cfg]))
    jitst.update({"nvrtc": f"{cfg[0].value}.{cfg[1].value}", "code_shape_variant": cfg[2].value, "arith_opt": cfg[3].value})
    jitst.update({"cold_first_step_s": cold_first_step_s, "all_modules_ready_s": jit_ready_s, "structures_compiling_after_first_step": int(pend.value),
                  "disk_cache_hits": int(hits.value), "disk_cache_dir": cdir.value.decode(),
                  "note": "first execution runs on the interpreter while worker threads compile (or read the on-disk cubin cache); steady-state steps use the specialised kernels"})
    del s
    try:
        dm14 = dm14_block(bt, L, wl, peaks["hbm_gbs"]) if not args.no_blocks else None
    except Exception as e:
        dm14, extra_err["dm14"] = None, repr(e)

    try:
        c4 = c4_block(bt, L, wl) if not args.no_blocks else None
    except Exception as e:
        c4, extra_err["c4"] = None, repr(e)

    cpu = cpu_baseline_port(N, specs, budget_s=args.cpu_budget) if not args.no_cpu else None
    out = {"metric": "gates/s", "value": value, "unit": "gates/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (ComplexF64 amplitudes)", "data": "synthetic",
           "config": {"workload": f"C2: {N}-qubit state vector, QFT({N}) + {depth} random layers (H/RX/RY/RZ/T + CNOT/CZ/CP brickwork), {ngates} gates, seed 28",
                      "fusion": "host fusion pass + shared-memory tile kernel", "l2": f"inputs larger than L2 ({(16 << N) / 2**30:.1f} GiB state)", "parallelism": "1 GPU"},
           "clocks": clk, "e2e": e2e, "gpu_launches": int(n1 - n0), "roofline": roof, "kernels": per_cls, "unfused_gate_kernels": micro,
           "interpreter": interp, "cold_first_step_s": cold_first_step_s, "kraus": kraus, "dm14": dm14, "c4": c4,
           "cpu_baseline": cpu, "state_norm2": norm, "jit": jitst}
    if extra_err:
        out["block_errors"] = extra_err
    print(json.dumps(out))


def jit_stats(lib, warmup_seconds=None):
    """pass specialiser statistics: modules compiled (all during warm-up), specialised launches, compile time"""
    jc, jl, jf, jt = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_double()
    lib.bt_jit_stats(C.byref(jc), C.byref(jl), C.byref(jf), C.byref(jt))
    out = {"modules_compiled": int(jc.value), "specialised_launches": int(jl.value), "fell_back": int(jf.value), "compile_seconds": float(jt.value),
           "note": "recurring fused passes are compiled once by NVRTC (worker threads, on-disk cubin cache) into straight-line kernels; bench.py waits for the workers in the untimed warm-up"}
    if warmup_seconds is not None:
        out["warmup_seconds"] = warmup_seconds
    return out


def cpu_baseline_port(N, specs, budget_s=15.0):
    """The oracle's strided C port (oracle/strided_cpu.c, OpenMP over all host cores) timed on a bounded sample of the
    same workload: consecutive gates of the layered section at the full size, until the time budget is used."""
    from oracle import bt_oracle as O
    from oracle import strided as S

    try:
        S.use_all_cores()  # the launcher's OMP_NUM_THREADS (torchrun: 1) does not decide the baseline's core count
        layered = specs[N * (N + 1) // 2:]  # skip the QFT prefix
        sv = S.SV(N)
        t0 = time.perf_counter()
        n = 0
        for name, q, t, c in layered:
            sv.apply(O.Op(name, q, t, control=c))
            n += 1
            if time.perf_counter() - t0 > budget_s and n >= 4:
                break
        dt = time.perf_counter() - t0
        return {"value": n / dt, "unit": "gates/s", "cores": S.num_threads(), "kind": "port",
                "sample": f"first {n} gates of the layered section of the same circuit at {N} qubits, in-place strided C + OpenMP ({dt:.1f} s)"}
    except MemoryError as e:
        return {"value": None, "unit": "gates/s", "cores": 0, "kind": "port", "sample": f"not enough host memory: {e}"}


# ------------------------------------------------------------------------------------------------------------------
def remap_log(L, st):
    """per-remap (ms waiting for the peers, ms in the pull kernel, ms until every peer has finished reading) since the last call"""
    buf = (C.c_float * (3 * 512))()
    n = C.c_int()
    L.check(st.lib.bt_sv_remap_log(st.h, 512, buf, C.byref(n)))
    return [(round(buf[3 * i], 3), round(buf[3 * i + 1], 3), round(buf[3 * i + 2], 3)) for i in range(n.value)]


def bench_sharded(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    bt = ge.load_package()
    L = bt._lib
    from importlib import import_module

    wl = import_module(ge.PKG_NAME + ".workloads")
    D = import_module(ge.PKG_NAME + ".dist")
    rank, world, local = D.env_rank_world()
    torch.cuda.set_device(local)
    # rank 0 prints exactly ONE JSON line on stdout: NCCL writes its version banner to the process's stdout from C, so file
    # descriptor 1 points at stderr while the job runs and the line goes out through a saved copy of the real stdout
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    debug = os.environ.get("BENCH_DEBUG", "") not in ("", "0")
    if debug:  # per-rank stderr file: every rank's own timings (and the library's BT_TILE_DEBUG lines, if switched on)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        os.dup2(os.open(os.path.join(ROOT, "gpurun_out", f"bench_rank{rank}.err"), os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644), 2)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = world.bit_length() - 1
    if args.total_qubits > 0:  # strong scaling: fixed total size
        N = args.total_qubits
        n_local = N - g
    else:                      # weak scaling: fixed shard size
        n_local = args.shard_qubits
        N = n_local + g
    specs = wl.c5_random(N, args.c5_depth, 31)
    arr = bt.pack_gates(wl.to_ops(bt, specs))
    ngates = len(arr)
    st = D.ShardedState(N)
    lib = st.lib
    t_start = time.perf_counter()

    def step():
        L.check(lib.bt_sv_set_basis(st.h, 0))
        L.check(lib.bt_sv_apply_circuit(st.h, L.ptr(arr), ngates, 1))

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # first execution: interpreter, the library's compile workers specialise the fused passes meanwhile; then wait for them
    st.sync()
    t_c0 = time.perf_counter()
    step()
    st.sync()
    cold_first_step_s = time.perf_counter() - t_c0
    L.check(lib.bt_jit_wait(None))
    jit_ready_s = time.perf_counter() - t_c0
    for _ in range(args.warmup):
        step()
    st.sync()
    if rank == 0:
        time.sleep(max(0.0, 1.5 - (time.perf_counter() - t_start)))  # nvidia-smi has finished starting up before the timed region
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    gc.collect()
    gc.disable()  # no cyclic-GC pause inside the timed region (a full collection with torch imported takes ~0.5 s)
    clocks.begin()
    r0 = st.remap_stats()
    remap_log(L, st)  # drop the warm-up's records
    n0 = st.launch_count()
    ms = C.c_float()
    L.check(lib.bt_sv_profile_enable(st.h, 1))
    jit0 = jit_stats(lib)["specialised_launches"]
    t_host0 = time.perf_counter()
    L.check(lib.bt_sv_timer_start(st.h))
    step_host = []
    for _ in range(args.steps):
        t_s = time.perf_counter()
        step()
        step_host.append(time.perf_counter() - t_s)
    L.check(lib.bt_sv_timer_stop(st.h, C.byref(ms)))
    t_host = time.perf_counter() - t_host0
    jit_timed = jit_stats(lib)["specialised_launches"] - jit0
    counts = (C.c_uint64 * 4)()
    cms = (C.c_double * 4)()
    L.check(lib.bt_sv_profile_read(st.h, counts, cms))
    L.check(lib.bt_sv_profile_enable(st.h, 0))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n1 = st.launch_count()
    r1 = st.remap_stats()
    rlog = remap_log(L, st)
    if debug:
        sys.stderr.write(f"[bench rank {rank}] per remap (ms waiting for peers, ms pulling, ms until all peers done): {rlog}\n")
        sys.stderr.write(f"[bench rank {rank}] timed region {ms.value:.1f} ms (device events), host {t_host * 1e3:.1f} ms, host time to enqueue each step {[round(x * 1e3, 1) for x in step_host]} ms, "
                         f"launches {n1 - n0}, classes {[(int(counts[i]), round(float(cms[i]), 1)) for i in range(4)]}, remaps {r1[0] - r0[0]} in {r1[2] - r0[2]:.1f} ms, jit {jit_stats(lib)}\n")
        sys.stderr.flush()
    t = torch.tensor([ms.value], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    clk = clocks.stop() if rank == 0 else None
    gc.enable()
    # end to end through the host API: gate list in host memory -> <Z_q> for every qubit and the norm back in host memory,
    # host wall clock between barriers, max over ranks, median of up to 3 steps
    ez = np.empty(N)
    nrm = np.empty(1)
    zmat = L.cmat(bt.gate["Z"], 2)
    e2e_t = []
    for _ in range(max(1, min(args.steps, 3))):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        L.check(lib.bt_sv_expect_1q_all(st.h, L.ptr(zmat), L.pdouble(ez)))
        L.check(lib.bt_sv_norm2(st.h, L.pdouble(nrm)))
        te = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_t.append(float(te.item()))
    e2e_s = float(np.median(e2e_t))
    if rank == 0:
        ms_per_step = ms_max / args.steps
        equiv = ngates * world * (2.0 ** (n_local - 28))
        value = equiv / (ms_per_step / 1e3)
        remaps = (r1[0] - r0[0]) / args.steps
        rbytes = (r1[1] - r0[1]) / args.steps
        rms = (r1[2] - r0[2]) / args.steps
        pulls = [p for (_, p, _) in rlog]
        out = {"metric": "gates/s", "value": value, "unit": "gates/s", "unit_note": "28-qubit-equivalent gates/s: circuit gates x shard amplitudes / 2^28, summed over ranks (equal per-GPU work at every N)", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if args.total_qubits > 0 else "weak", "vs_baseline": None,
               "dtype": "f64 (ComplexF64 amplitudes)", "data": "synthetic",
               "config": {"workload": f"C5: {N}-qubit state vector random circuit (depth {args.c5_depth}: random 1q gate per qubit + CNOT/CZ brickwork, seed 31), {ngates} gates, "
                                      f"2^{n_local} amplitudes per GPU", "parallelism": f"{world} shards, top {g} index bits global, qubit remap by peer-memory pull over NVLink",
                          "l2": "inputs larger than L2"},
               "circuit_gates_per_s": ngates / (ms_per_step / 1e3), "clocks": clk, "gpu_launches": int(n1 - n0),
               "kernels_rank0": {n: {"launches": int(counts[i]), "ms": float(cms[i])} for i, n in enumerate(["tile", "dense", "diag", "other"])},
               "host_seconds_rank0": t_host, "jit_rank0": jit_stats(lib),
               "roofline": ({"bound": "hbm", "kernel": ("bt_jit_pass (fused multi-gate pass specialised by NVRTC)" if 2 * jit_timed >= int(counts[0]) else "k_tile_tma (fused multi-gate pass)") + " on rank 0's shard", "achieved": 32.0 * (1 << n_local) / (cms[0] / counts[0] / 1e3) / 1e9,
                             "peak": load_peaks()[0]["hbm_gbs"], "unit": "GB/s", "frac": 32.0 * (1 << n_local) / (cms[0] / counts[0] / 1e3) / 1e9 / load_peaks()[0]["hbm_gbs"],
                             "traffic": None, "avg_launch_ms": cms[0] / counts[0], "bytes_per_launch": 32.0 * (1 << n_local)} if counts[0] else None),
               "remap": {"per_step": remaps, "nvlink_bytes_per_rank_per_step": rbytes, "ms_per_step": rms, "GBps_per_rank": (rbytes / (rms / 1e3) / 1e9) if rms > 0 else None,
                         "nvlink_peak_GBps": 900.0, "pull_ms_rank0": {"min": min(pulls), "median": float(np.median(pulls)), "max": max(pulls)} if pulls else None,
                         "wait_for_peers_ms_rank0_median": float(np.median([w for (w, _, _) in rlog])) if rlog else None,
                         "sync": "device-side epoch flags in peer memory" if os.environ.get("BT_REMAP_DEVICE_SYNC", "1") != "0" else "host barriers"},
               "cold_first_step_s": cold_first_step_s, "all_modules_ready_s": jit_ready_s,
               "e2e": {"value": equiv / e2e_s, "unit": "gates/s", "h2d_bytes_per_step": int(arr.nbytes), "d2h_bytes_per_step": int(ez.nbytes + 8), "seconds_per_step": e2e_s,
                       "api": "host wall clock, max over ranks: bt_sv_set_basis + bt_sv_apply_circuit(host gate list) + bt_sv_expect_1q_all + bt_sv_norm2 (results in host memory)"},
               "checksum": {"norm2": float(nrm[0]), "sum_expect_Z": float(np.sum(ez))}}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.barrier()
    del st
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------------
def bench_reference(args):
    """Reference arm: the reference path's CPU port on the host cores (Julia itself is not installable here).  Value =
    the oracle's strided C port, all the cores this process may use (set explicitly -- torchrun exports OMP_NUM_THREADS=1),
    on a bounded sample of the repo arm's workload at the same --gpus: C2 (28 qubits) for N = 1, C5 for N > 1 (consecutive
    gates of the same circuit on a state of the per-GPU shard size, in the same 28-qubit-equivalent gates/s).  The kron-chain
    restatement of the reference ALGORITHM (2^N x 2^N sparse operator per gate) is timed beside it on what it can reach."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as ge
    from importlib import import_module

    ge.load_package()
    wl = import_module(ge.PKG_NAME + ".workloads")
    from oracle import bt_oracle as O
    from oracle import strided as S

    cores = S.use_all_cores()
    world = max(int(os.environ.get("WORLD_SIZE", "1")), args.gpus)
    sharded = world > 1 or args.workload == "c5"
    if sharded:
        g = world.bit_length() - 1
        n_total = args.total_qubits if args.total_qubits > 0 else args.shard_qubits + g
        n_cpu = min(n_total, args.shard_qubits if args.total_qubits <= 0 else args.total_qubits)
        try:  # a 31-qubit state is 32 GiB of host memory: fall back to what fits
            import psutil
            while n_cpu > 24 and (24 << n_cpu) > psutil.virtual_memory().available:
                n_cpu -= 1
        except Exception:
            pass
        specs_full = wl.c5_random(n_total, args.c5_depth, 31)
        # the same layers restricted to the qubits the CPU state holds (labels above n_cpu dropped): same gate mix per qubit
        layered = [(nm, q, t, c) for (nm, q, t, c) in wl.c5_random(n_cpu, args.c5_depth, 31)]
        scale = 2.0 ** (n_cpu - 28)
        N = n_cpu
        workload = (f"C5: {n_total}-qubit state vector random circuit (depth {args.c5_depth}, {len(specs_full)} gates) -- bounded sample: consecutive gates of the same "
                    f"generator at {n_cpu} qubits (the per-GPU shard size), reported in 28-qubit-equivalent gates/s")
    else:
        N, depth = args.qubits, args.depth
        specs = wl.c2_qft_layered(N, depth, 28)
        layered = specs[N * (N + 1) // 2:]
        scale = 1.0
        workload = f"C2: {N}-qubit state vector, QFT({N}) + {depth} random layers, bounded sample of the layered section"
    budget = max(4.0, args.cpu_budget)
    sv = S.SV(N)
    for name, q, t, c in layered[:2]:  # warm-up (pages the state in)
        sv.apply(O.Op(name, q, t, control=c))
    tot_g, tot_t = 0, 0.0
    pos = 2
    for _ in range(max(1, args.steps)):
        t0 = time.perf_counter()
        n = 0
        while True:
            name, q, t, c = layered[pos % len(layered)]
            sv.apply(O.Op(name, q, t, control=c))
            pos += 1
            n += 1
            if time.perf_counter() - t0 > budget / max(1, args.steps) and n >= 2:
                break
        tot_g += n
        tot_t += time.perf_counter() - t0
    value = tot_g / tot_t * scale
    # the reference ALGORITHM (kron chain + sparse mat-vec), on a size it can hold
    nk = 16
    ks = wl.layered(nk, 1, 28)[:12]
    st = O.zero_state(nk)
    t0 = time.perf_counter()
    for name, q, t, c in ks:
        st = O.Op(name, q, t, control=c).expand(nk) @ st
    kdt = time.perf_counter() - t0
    kron = {"qubits": nk, "gates_per_s": len(ks) / kdt, "gates_per_s_28q_equivalent": len(ks) / kdt * 2.0 ** (nk - 28),
            "note": "restatement of hilbert()+SpMV (src/hilbert.jl:18-159,505) in scipy, single thread like the reference; cannot reach 28 qubits"}
    out = {"impl": "reference", "metric": "gates/s", "value": value, "unit": "gates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * tot_t / max(1, args.steps), "higher_is_better": True, "scaling": "strong" if args.total_qubits > 0 else "weak", "vs_baseline": None,
           "dtype": "f64 (ComplexF64 amplitudes)", "data": "synthetic",
           "config": {"workload": workload, "parallelism": "host CPU"},
           "cpu_baseline": {"value": value, "unit": "gates/s", "cores": cores, "kind": "port",
                            "sample": f"{tot_g} consecutive gates at {N} qubits, in-place strided C + OpenMP on {cores} threads ({tot_t:.1f} s)"},
           "reference_algorithm": kron,
           "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    if sharded:
        out["unit_note"] = "28-qubit-equivalent gates/s: gates x state amplitudes / 2^28 (the repo arm's unit at N > 1)"
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--qubits", type=int, default=28)
    ap.add_argument("--depth", type=int, default=100)
    ap.add_argument("--shard-qubits", type=int, default=31, help="local index bits per GPU for the sharded workload")
    ap.add_argument("--total-qubits", type=int, default=0, help="> 0: strong scaling of the sharded workload at this total size")
    ap.add_argument("--c5-depth", type=int, default=20)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-blocks", action="store_true", help="skip the kraus / dm14 side blocks of the N = 1 line")
    ap.add_argument("--workload", default="auto", choices=["auto", "c2", "c5"], help="auto: C2 (28q) on 1 GPU, C5 (31q per GPU, sharded) on N > 1")
    args = ap.parse_args()
    if args.impl == "reference":
        bench_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 or args.gpus > 1 or args.workload == "c5":
        if world == 1 and args.gpus > 1:
            # launched without torchrun: re-exec under torch.distributed.run
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port", "29533", __file__] + sys.argv[1:]
            os.execv(sys.executable, cmd)
        bench_sharded(args)
    else:
        bench_single(args)


if __name__ == "__main__":
    main()
