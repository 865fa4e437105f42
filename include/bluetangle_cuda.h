/*
 * bluetangle_cuda.h -- C ABI of libbluetangle_cuda.so, the B200 (sm_100a) simulation backend for
 * BlueTangle.jl's apply -> noise -> measure/sample -> expect path.
 *
 * The reference (pure Julia) has no FFI; its backend seam is multiple dispatch on the state type
 * (src/hilbert.jl:469 / :639 / :566).  Every entry point below names the reference method it replaces for a
 * device-resident state; the Julia `ccall` stubs a maintainer would add are in INTEGRATION.md and
 * julia/BlueTangleCUDA.jl.
 *
 * Conventions:
 *  - qubit arguments are the reference's 1-based labels; qubit 1 is the most significant bit of the basis
 *    index (src/bit.jl:9-15), i.e. qubit q <-> bit N-q.  target = -1 means "1-qubit op", control = -2 means
 *    "no control" (src/struct.jl:377,389,456).
 *  - matrices are column-major ComplexF64 (Julia native), indexed 2*b_qubit + b_target exactly like Op.mat.
 *  - every call returns 0 on success, <0 on error; bt_last_error() returns a thread-local message.
 *    Nothing unwinds across the ABI.  There is no CPU fallback: no device => BT_ERR_CUDA.
 *  - random numbers are never generated inside the library: callers pass uniform draws in [0,1).
 *  - one CUDA stream per handle; gate calls only enqueue, calls returning data to the host synchronise.
 */
#ifndef BLUETANGLE_CUDA_H
#define BLUETANGLE_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BT_OK 0
#define BT_ERR_ARG (-1)      /* invalid argument (mirrors the reference's throws) */
#define BT_ERR_CUDA (-2)     /* CUDA runtime failure / no device */
#define BT_ERR_ALLOC (-3)    /* out of device memory */
#define BT_ERR_SAMPLE (-4)   /* _weighted_sample found no index (src/hilbert.jl:810-819 returns nothing) */
#define BT_ERR_UNSUPPORTED (-5)

typedef struct bt_sv bt_sv; /* state vector: n_batch trajectories of 2^n amplitudes, optionally one shard of P */
typedef struct bt_dm bt_dm; /* density matrix stored as a 2n-qubit vector (row qubit q <-> bit n-q, column qubit q <-> bit 2n-q) */
typedef struct { double re, im; } bt_c64; /* == Julia ComplexF64 */

/* One op of a circuit handed to bt_sv_apply_circuit / bt_dm_apply_circuit (replaces the per-op loop of
 * src/hilbert.jl:517-553 for plain gates).  nq in {1,2}; m column-major, first 4^nq entries used. */
typedef struct {
  int32_t nq, qubit, target, control;
  bt_c64 m[16];
} bt_gate;

/* ---- library ---------------------------------------------------------------------------------------- */
const char* bt_last_error(void);
int bt_version(void);
int bt_device_count(int* n);
int bt_set_device(int device);
/* measured FP64 FMA peak of the current device (register-operand DFMA loop on every SM, best of `reps` launches, CUDA events):
 * the FP64 roofline denominator bench.py reports against.  No reference analogue. */
int bt_fp64_peak(double* tflops, double* ms_per_launch, int reps);

/* ---- state vector life cycle: zero_state/one_state/plus_state/product_state src/hilbert.jl:835-882 -- */
int bt_sv_create(int n_qubits, int64_t n_batch, bt_sv** out);     /* on the current device, |0..0> in every trajectory */
int bt_sv_destroy(bt_sv* s);
int bt_pool_release(void);                                       /* free the handles bt_sv_destroy parked for re-use (small unsharded states) */
int bt_sv_n_qubits(const bt_sv* s, int* n);                       /* get_N  src/ops.jl:8 */
int bt_sv_set_basis(bt_sv* s, uint64_t index);                    /* every trajectory := |index> */
int bt_sv_set_plus(bt_sv* s);                                     /* plus_state src/hilbert.jl:869 */
int bt_sv_upload(bt_sv* s, const bt_c64* host, uint64_t len);     /* len = n_batch * 2^n (local shard if sharded) */
int bt_sv_download(const bt_sv* s, bt_c64* host, uint64_t len);
int bt_sv_copy(bt_sv* dst, const bt_sv* src);
int bt_sv_sync(const bt_sv* s);
/* CUDA-event timing on the handle's stream (used by bench.py; events see exactly the launching stream) */
int bt_sv_timer_start(bt_sv* s);
int bt_sv_timer_stop(bt_sv* s, float* ms);
int bt_sv_launch_count(const bt_sv* s, uint64_t* n);              /* kernels launched on this handle so far */
/* per-launch CUDA-event profile by kernel class {0: fused tile kernel, 1: dense gate, 2: diagonal gate, 3: other} */
int bt_sv_profile_enable(bt_sv* s, int on);
int bt_sv_profile_read(bt_sv* s, uint64_t counts[4], double ms[4]);

/* ---- gates: op.expand(N)*state  src/hilbert.jl:505 with hilbert() src/hilbert.jl:18-159 --------------- */
int bt_sv_apply_1q(bt_sv* s, int qubit, const bt_c64 m[4], int control);               /* hilbert.jl:143-159 */
int bt_sv_apply_2q(bt_sv* s, int qubit, int target, const bt_c64 m[16], int control);  /* hilbert.jl:18-70 (+CCZX :73-103; any matrix may be controlled unless strict) */
int bt_sv_apply_3q(bt_sv* s, int first_qubit, const bt_c64 m[64]);                     /* hilbert3 hilbert.jl:106-128 */
/* whole op list in one call; fuse != 0 enables the host fusion pass + multi-gate shared-memory kernel */
int bt_sv_apply_circuit(bt_sv* s, const bt_gate* g, uint64_t n, int fuse);
/* gate applied only to trajectories whose outcome[t] == want (ifOp branches, src/struct.jl:587-590);
 * the outcomes are the handle's device-side record of the last bt_sv_measure_z call) */
int bt_sv_apply_1q_if(bt_sv* s, int qubit, const bt_c64 m[4], int control, int want);
int bt_sv_apply_2q_if(bt_sv* s, int qubit, int target, const bt_c64 m[16], int control, int want);
/* pure host (no device needed): the fused passes bt_sv_apply_circuit(fuse != 0) forms for this gate list on one unsharded GPU:
 * number of passes and fused blocks, and per pass (up to cap) the original gates it carries, the index bits its tile takes beyond the
 * fixed low bits (16 ints per pass, -1 padded), its cost units and why it ended (0 no eligible block left, 1 cost cap, 2 block cap);
 * launch_counts (optional, 4 ints): kernel launches the passes need (a pass that overflows the slots of one launch is split),
 * register programs, other items, single-gate passes */
int bt_fusion_plan_host(int n_qubits, const bt_gate* g, uint64_t n, int* n_passes, int* n_blocks, int* gates_in_pass, int* tile_bits /* cap x 16 */,
                        double* cost_in_pass, int* end_reason, int cap, int* launch_counts);
int bt_fusion_stats(uint64_t* passes, uint64_t* blocks); /* cumulative: fused tile-kernel launches and blocks they carried */
int bt_fusion_flops(double* flops); /* cumulative FP64 flops issued by the fused passes (FMA = 2 flops) */
/* Pass specialiser (csrc/bt_jit.cu): fused passes that recur are compiled once (NVRTC) into straight-line kernels.
 * bt_jit_stats: modules compiled, specialised launches, structures that fell back to the interpreter, seconds spent compiling.
 * bt_jit_selftest: host-only check (no device): generate and compile a synthetic pass using every micro-op; 0 = ok;
 * `source` (optional, `cap` bytes) receives the generated CUDA text; unless the on-disk cubin cache is off (BT_JIT_CACHE_DIR="")
 * it also writes the cubin there and reads it back (-5 = it did not come back).  No reference analogue. */
int bt_jit_stats(uint64_t* compiled, uint64_t* launches, uint64_t* failed, double* compile_seconds);
int bt_jit_selftest(char* source, uint64_t cap);
/* Compilation runs on worker threads while the interpreter keeps executing the pass; bt_jit_wait blocks until they are idle
 * (*pending_before, optional: structures still compiling at the call).  bt_jit_cache_info: modules that came from the on-disk
 * cubin cache instead of NVRTC, and the cache directory in use (BT_JIT_CACHE_DIR; default ~/.cache/bluetangle_cuda; "" = off). */
int bt_jit_wait(uint64_t* pending_before);
int bt_jit_selftest_workers(int jobs); /* host-only: `jobs` synthetic passes through the compile workers; 0 = ok, -2 = no libnvrtc */
int bt_jit_cache_info(uint64_t* disk_hits, char* dir, uint64_t cap);
/* BT_JIT_VERIFY=1 (environment): every specialised launch is cross-checked against the interpreter kernel on a copy of the state; a
 * disagreement is counted and reported on stderr and the interpreter's result is kept.  Returns launches checked / disagreements. */
int bt_jit_verify_stats(uint64_t* checked, uint64_t* failed);
/* debugging aid: only the k-th eligible fused pass since this call (0-based) runs specialised, the others stay on the interpreter
 * (k < 0 removes the filter); dump != 0 prints the generated CUDA text of that pass to stderr */
int bt_jit_debug_only(int k, int dump);
/* what the specialiser will use in this process: version of the run-time compiler it loaded (0.0 = none), the code shape
 * (BT_JIT_VARIANT, chosen by compiler version unless set) and the arithmetic reshaping flags (BT_JIT_OPT).  Host only. */
int bt_jit_config(int* nvrtc_major, int* nvrtc_minor, int* variant, int* opt);
int bt_set_strict(int strict); /* strict != 0: controlled non-adjacent 2q gates other than CX/CZ are rejected like hilbert.jl:58-64 */

/* ---- reductions: partial_trace src/linalg.jl:167-230, :83-140 ---------------------------------------- */
int bt_sv_rdm1(const bt_sv* s, int qubit, bt_c64* out /* n_batch x 4, column-major 2x2 each */);
int bt_sv_rdm2(const bt_sv* s, int qubit_a, int qubit_b, bt_c64* out /* n_batch x 16; index 2*b_min + b_max */);
int bt_sv_rdm3(const bt_sv* s, int first_qubit, bt_c64* out /* n_batch x 64; qubits first..first+2 */);
int bt_sv_rdm(const bt_sv* s, int k, const int* qubits, bt_c64* out /* n_batch x 4^k */); /* partial_trace(state, keep) linalg.jl:83-86, k <= 12 arbitrary qubits, ascending-label order (k >= 4: tiled Gram kernel, unsharded states) */
/* Squared singular values (descending) of the 2^n_low x 2^(N - n_low) column-major reshape of the state: the spectrum
 * entanglement_entropy(psi) src/func.jl:299-312 sums over (n_low = N / 2).  One-sided Jacobi iteration on a scratch copy, on the
 * device; spec: n_batch x 2^min(n_low, N - n_low); sweeps (optional): Jacobi sweeps used.  The long side may have up to 2^13 entries. */
int bt_sv_schmidt_spectrum(const bt_sv* s, int n_low, double* spec, int* sweeps);
/* pure host (no device needed): the disjoint vector pairs the Jacobi kernel rotates in round `round` (0..nvec-2) of a sweep */
int bt_jacobi_pairs_host(int nvec, int round, int* pairs /* nvec/2 x 2 */);
/* expect(state, op) src/func.jl:91 for any Op: 1- or 2-qubit matrix (column-major), optional control (-2 = none):
 * real(state' * op.expand(N) * state) as a trace against the reduced density matrix of the qubits the op touches; out[n_batch] */
int bt_sv_expect_op(const bt_sv* s, int nq, int qubit, int target, int control, const bt_c64* m, double* out);
int bt_sv_norm2(const bt_sv* s, double* out /* n_batch: sum |a|^2 */);
int bt_sv_inner(const bt_sv* a, const bt_sv* b, bt_c64* out /* n_batch: <a|b>, src/tensor.jl:199 */);
int bt_sv_normalize(bt_sv* s);
int bt_sv_probs(const bt_sv* s, double* host /* n_batch * 2^n : abs2.(state) src/ops.jl:51,98 */);

/* ---- measurement: born_measure_Z src/hilbert.jl:682-696, _reset_Z :752-759 ----------------------------
 * basis rotations (MX/MY/MR, src/hilbert.jl:669-679) are applied by the caller with bt_sv_apply_1q.
 * u: n_batch uniforms; outcome = (u < p0) ? 0 : 1; the state is projected and renormalised by its own norm.
 * outcome/p0 may be NULL (no host synchronisation then; outcomes stay in the device outcome buffer). */
int bt_sv_measure_z(bt_sv* s, int qubit, const double* u, int32_t* outcome, double* p0, int reset);
/* k <= 4 consecutive born_measure_Z calls (src/hilbert.jl:682-696; _reset_Z :752-759 where reset[j] != 0) on DISTINCT qubits in one
 * read pass + one collapse pass instead of k of each: the joint distribution of the k measured bits is reduced per trajectory, then
 * the outcomes are decided in order -- measurement j sees the state collapsed and renormalised by measurements 0..j-1, exactly like
 * the sequential calls (the per-shot loop of a monitored circuit, src/ops.jl:616-631).  u[t*k + j], outcomes[t*k + j] (may be NULL:
 * no synchronisation); the handle's outcome record (bt_sv_outcomes) then holds the pattern sum_j outcome_j << j. */
int bt_sv_measure_z_multi(bt_sv* s, int k, const int* qubits, const double* u, int32_t* outcomes, const int* reset /* k flags or NULL */);
/* Deferred outcomes: while the log is on, bt_sv_measure_z_multi calls with outcomes == NULL append their n_batch x k outcomes to a
 * device-side log (call after call) instead of returning them; bt_sv_measure_log_read copies the log to the host (one
 * synchronisation for a whole monitored circuit, src/ops.jl:616-631) and empties it.  on != 0 also empties the log. */
int bt_sv_measure_log(bt_sv* s, int on);
int bt_sv_measure_log_read(bt_sv* s, int32_t* out, uint64_t cap, uint64_t* n);
int bt_sv_outcomes(const bt_sv* s, int32_t* outcome /* n_batch: results of the last measure/kraus call */);
/* Trajectory mask for batched states: while set, bt_sv_apply_1q/2q/3q, bt_sv_apply_circuit (gate by gate then), bt_sv_kraus and
 * bt_sv_measure_z act only on trajectories with mask[t] != 0; the others are untouched, consume no draw and report outcome /
 * chosen = -1.  This is how the branch op lists of an ifOp (src/struct.jl:587-590: apply(state, ifop; noise)) run on a batch whose
 * trajectories measured different outcomes, incl. the branch's own noise draws and nested measurements.  mask = NULL clears it. */
int bt_sv_set_mask(bt_sv* s, const int32_t* mask /* n_batch */);

/* ---- Kraus trajectory step: __QuantumChannel_new_apply src/struct.jl:31-55, __calc_prob :9-29,
 * _weighted_sample src/hilbert.jl:810-819 (k = first i with u <= cumsum(p)[i]).
 * nq in {1,2,3}; K = nK column-major (2^nq x 2^nq) matrices; target ignored unless nq == 2.
 * Reproduces the reference's quirk for nq == 2 and qubit > target.
 * chosen may be NULL (no synchronisation). */
int bt_sv_kraus(bt_sv* s, int nq, int qubit, int target, const bt_c64* K, int nK, const double* u, int32_t* chosen);
int bt_sv_kraus_probs(const bt_sv* s, int nq, int qubit, int target, const bt_c64* K, int nK, double* probs /* n_batch x nK */);

/* ---- observables: expect src/func.jl:91-101, correlation :139-147 ------------------------------------ */
int bt_sv_expect_pauli(const bt_sv* s, const char* paulis /* n chars from IXYZ, qubit 1 first */, double* out /* n_batch */);
int bt_sv_expect_1q_all(const bt_sv* s, const bt_c64 m[4], double* out /* n_batch x n : Re<psi|m_q|psi> */);
int bt_sv_expect_product(const bt_sv* s, int n_ops, const int* qubits, const bt_c64* mats /* n_ops x 4 */, double* out /* n_batch */);
int bt_sv_expect_matrix2q(const bt_sv* s, int qubit, int target, const bt_c64 m[16], double* out);
/* Pauli-sum (Hamiltonian) expectation value sum_k coefs[k] * <P_k>: the scalar real(state' * hamiltonian(...) * state) of
 * src/vqa.jl:36-67 + src/func.jl:91 (the VQE loss, src/vqa.jl:282-283) without building the 2^N x 2^N operator.  paulis =
 * n_terms strings of n characters (I/X/Y/Z, qubit 1 first, no separators).  Terms diagonal in a common product basis are
 * evaluated together in one read of the state (a scratch copy is rotated into that basis first); out[n_batch]. */
int bt_sv_expect_pauli_sum(const bt_sv* s, int n_terms, const char* paulis, const double* coefs, double* out /* n_batch */);

/* ---- sampling: sample src/ops.jl:46-62 as inverse CDF on caller-supplied uniforms:
 * t = u*sum(p); index = first i with cumsum_i >= t.  n_batch must be 1 unless per_traj != 0, in which case
 * shots uniforms are consumed per trajectory (out is n_batch x shots). */
int bt_sv_sample(const bt_sv* s, const double* u, uint64_t shots, int64_t* out);
int bt_sv_sample_batched(const bt_sv* s, const double* u, uint64_t shots_per_traj, int64_t* out);

/* ---- density matrix: apply(rho,op) src/hilbert.jl:639-666; channels src/struct.jl:58-76 --------------- */
int bt_dm_create(int n_qubits, bt_dm** out);                      /* |0..0><0..0| */
int bt_dm_destroy(bt_dm* d);
int bt_dm_n_qubits(const bt_dm* d, int* n);
int bt_dm_from_sv(bt_dm* d, const bt_sv* s);                      /* state*state' src/ops.jl:810 */
int bt_dm_upload(bt_dm* d, const bt_c64* host, uint64_t len);     /* column-major 2^n x 2^n */
int bt_dm_download(const bt_dm* d, bt_c64* host, uint64_t len);
int bt_dm_sync(const bt_dm* d);
int bt_dm_timer_start(bt_dm* d);
int bt_dm_timer_stop(bt_dm* d, float* ms);
int bt_dm_launch_count(const bt_dm* d, uint64_t* n);
int bt_dm_apply_1q(bt_dm* d, int qubit, const bt_c64 m[4], int control);               /* e_op*rho*e_op' hilbert.jl:655-656 */
int bt_dm_apply_2q(bt_dm* d, int qubit, int target, const bt_c64 m[16], int control);
int bt_dm_kraus(bt_dm* d, int nq, int qubit, int target, const bt_c64* K, int nK);      /* sum_k E_k rho E_k'  struct.jl:58-76: one HBM pass */
int bt_dm_dephase(bt_dm* d, int qubit);                                                 /* born_measure_Z(N,rho,q) hilbert.jl:784-796 */
int bt_dm_apply_circuit(bt_dm* d, const bt_gate* g, uint64_t n, int fuse);
/* mixed list of unitaries and Kraus channels (to_rho's loop src/ops.jl:813-841 incl. apply_noise); fuse != 0 multiplies
 * consecutive ops on the same qubit (pair) into one 4x4 / 16x16 superoperator => one pass over rho for gate + noise */
typedef struct {
  int32_t kind;            /* 0 = unitary (1 matrix), 1 = Kraus channel (nK matrices) */
  int32_t nq;              /* 1 or 2 */
  int32_t qubit, target, control; /* target = -1 for nq == 1; control = -2 for none (kind 0 only) */
  int32_t nK;
  const bt_c64* mats;      /* column-major 2^nq x 2^nq each */
} bt_dm_op;
int bt_dm_apply_ops(bt_dm* d, const bt_dm_op* ops, uint64_t n, int fuse);
int bt_dm_diag(const bt_dm* d, double* host /* 2^n : real(diag(rho)) src/ops.jl:130 */);
int bt_dm_trace(const bt_dm* d, bt_c64* out);
int bt_dm_expect_pauli(const bt_dm* d, const char* paulis, double* out);               /* real(tr(rho*O)) func.jl:92,146 */
int bt_dm_expect_1q_all(const bt_dm* d, const bt_c64 m[4], double* out /* n */);        /* func.jl:98 */
int bt_dm_expect_product(const bt_dm* d, int n_ops, const int* qubits, const bt_c64* mats, double* out);
int bt_dm_sample(const bt_dm* d, const double* u, uint64_t shots, int64_t* out);
/* partial_trace(rho, dims, trace_out) src/linalg.jl:88-140: reduced density matrix of the k kept qubits (ascending-label order,
 * first = most significant index bit), column-major 2^k x 2^k; k <= 12 */
int bt_dm_rdm(const bt_dm* d, int k, const int* qubits, bt_c64* out);
/* expect(rho, op) src/func.jl:92 for any Op (2-qubit and controlled operators included): real(tr(rho * op.expand(N))) */
int bt_dm_expect_op(const bt_dm* d, int nq, int qubit, int target, int control, const bt_c64* m, double* out);
/* singular values (descending) of bipartition_trace(rho) src/linalg.jl:151-161 generalised to the last n_keep qubits: the spectrum
 * entanglement_entropy(rho) src/func.jl:323-328 sums over (n_keep = N / 2); spec: 2^n_keep */
int bt_dm_bipartition_spectrum(const bt_dm* d, int n_keep, double* spec, int* sweeps);
/* fidelity(rho, sigma) src/tensor.jl:222-229: real(tr(sqrt(sqrt(rho) * sigma * sqrt(rho)))^2), as dense linear algebra on the device
 * (eigen-decomposition of rho by a Jacobi iteration with accumulated eigenvectors, two complex GEMMs, a second Jacobi iteration);
 * O(8^N) like the reference's sqrt(Matrix(rho)): registers of up to 11 qubits. */
int bt_dm_fidelity(const bt_dm* rho, const bt_dm* sigma, double* out);

/* ---- multi-GPU shards (no reference analogue).  One process per GPU: each rank creates its
 * shard, the host plumbing (torch.distributed / MPI) all-gathers the IPC handles and supplies a barrier.
 * The top log2(world) index bits are global (= the lowest-numbered qubits).  Gates on global qubits that
 * are diagonal or controls need no communication; others trigger a qubit-remap exchange in which every rank
 * pulls its new shard from its peers' memory over NVLink in one kernel (bit permutation fused in). */
#define BT_IPC_HANDLE_BYTES 64
#define BT_IPC_HANDLES_PER_SHARD 3 /* both amplitude buffers + the flag page of the device-side remap synchronisation */
typedef void (*bt_barrier_fn)(void* ctx);
int bt_sv_create_shard(int n_qubits_total, int rank, int world, bt_sv** out);
int bt_sv_ipc_export(bt_sv* s, void* handles /* BT_IPC_HANDLES_PER_SHARD * BT_IPC_HANDLE_BYTES */);
int bt_sv_ipc_attach(bt_sv* s, const void* all_handles /* world x BT_IPC_HANDLES_PER_SHARD * BT_IPC_HANDLE_BYTES, rank order */);
int bt_sv_attach_local_peers(bt_sv** shards, int world); /* single-process variant: all shards in this process */
int bt_sv_set_barrier(bt_sv* s, bt_barrier_fn fn, void* ctx);
int bt_sv_remap(bt_sv* s, const int* new_phys_of_logical_bit /* n_qubits_total entries */);
/* all shards of one state living in THIS process (single-threaded host driving several GPUs): segments in lockstep */
int bt_group_apply_circuit(bt_sv** shards, int world, const bt_gate* g, uint64_t n, int fuse);
/* pure host (no device needed): the segment plan bt_sv_apply_circuit follows for `world` shards from the identity layout */
int bt_plan_circuit_host(int n_qubits, int world, const bt_gate* g, uint64_t n, int* n_segments, int* seg_gates, int* seg_remap,
                         int* layouts /* cap x n_qubits */, int* order /* n gate indices in execution order */, int cap);
int bt_sv_layout(const bt_sv* s, int* phys_of_logical_bit /* n_qubits_total */);
/* pure host (no device needed): for the given loop indices of rank `rank`'s pull kernel, the local destination index it writes and
 * the physical source index ((source rank << n_local) | local) it reads -- the remap's index arithmetic, testable on a CPU */
int bt_remap_walk_host(int n_qubits, int n_local, int rank, const int* cur_phys, const int* new_phys, uint64_t n_idx, const uint64_t* loop_idx,
                       uint64_t* dest, uint64_t* src);
int bt_sv_remap_stats(const bt_sv* s, uint64_t* n_remaps, uint64_t* bytes_pulled_remote, float* ms_total);
/* per-remap timing record since the last call (at most `cap` most recent remaps; cleared on return): out[3*i + 0] = ms this rank
 * waited for its peers to reach the remap, [1] = ms in the pull kernel, [2] = ms until every peer had finished reading */
int bt_sv_remap_log(bt_sv* s, int cap, float* out /* cap x 3 */, int* n);
/* scalars returned by reductions on a shard are LOCAL partial sums unless an all-reduce callback is set */
typedef void (*bt_allreduce_fn)(void* ctx, double* buf, int n);
int bt_sv_set_allreduce(bt_sv* s, bt_allreduce_fn fn, void* ctx);

#ifdef __cplusplus
}
#endif
#endif /* BLUETANGLE_CUDA_H */
