// bt_internal.cuh -- shared declarations for libbluetangle_cuda.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <complex>
#include <vector>
#include <string>
#include <algorithm>

#include "../../include/bluetangle_cuda.h"

typedef std::complex<double> cplx;

// ---- error plumbing -----------------------------------------------------------------------------------
void bt_set_error(const char* fmt, ...);
#define BT_FAIL(code, ...)          \
  do {                              \
    bt_set_error(__VA_ARGS__);      \
    return (code);                  \
  } while (0)
#define BT_CUDA(call)                                                                             \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      bt_set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #call); \
      return (e__ == cudaErrorMemoryAllocation) ? BT_ERR_ALLOC : BT_ERR_CUDA;                     \
    }                                                                                             \
  } while (0)
#define BT_TRY(call)           \
  do {                         \
    int rc__ = (call);         \
    if (rc__ != BT_OK) return rc__; \
  } while (0)
#define BT_CHECK_LAUNCH(s)                    \
  do {                                        \
    (s)->launches++;                          \
    BT_CUDA(cudaPeekAtLastError());           \
  } while (0)

// ---- handle -------------------------------------------------------------------------------------------
#define BT_FLAG_PAGE_BYTES 4096

struct bt_sv {
  int n_qubits;      // qubits per trajectory, whole (possibly sharded) register
  int n_local;       // index bits held locally per trajectory
  int64_t n_batch;   // trajectories
  uint64_t len;      // local amplitudes = n_batch << n_local
  int device;
  cudaStream_t stream;
  double2* amp;      // device amplitudes (interleaved re,im == Julia ComplexF64)
  double2* alt;      // second buffer: remap target / scratch (lazily allocated unless sharded)
  // reductions
  double* d_part;    // per-block partials
  size_t part_cap;   // doubles
  double* d_res;     // final results (device)
  size_t res_cap;
  double* h_res;     // pinned host mirror of d_res
  // per-trajectory decision buffers (device)
  double* d_u;        // uniforms
  int32_t* d_outcome; // measurement outcomes / chosen Kraus indices
  double2* d_mats;    // per-trajectory selected (scaled) matrices, 64 entries each
  double* d_scale;    // per-trajectory scale factors
  int32_t* d_err;     // device error flag
  int32_t* h_flag;    // pinned
  size_t traj_cap;
  // trajectory mask (bt_sv_set_mask): gates, Kraus steps and measurements act only on trajectories with d_mask[t] != 0
  int32_t* d_mask;
  bool mask_on;
  // deferred outcome log (bt_sv_measure_log): multi-measurement calls append their outcomes here instead of returning them, so a
  // monitored circuit runs without a host synchronisation per measurement layer
  int32_t* d_mlog; size_t mlog_cap, mlog_len; bool mlog_on;
  // grow-only device scratch (sampling prefix sums, uniforms, results)
  void* d_scratch; size_t scratch_cap;
  // timing / accounting
  cudaEvent_t ev0, ev1;
  mutable uint64_t launches;
  // optional per-launch CUDA-event profile (bt_sv_profile_*): kernel class -> count / summed duration
  bool prof_on;
  std::vector<cudaEvent_t>* prof_ev;   // pairs (start, stop)
  std::vector<int>* prof_cls;
  size_t prof_used;
  // sharding
  int rank, world, g;        // g = log2(world)
  int phys_of_bit[64];       // logical bit -> physical bit (local bits 0..n_local-1, then rank bits)
  double2* peer_amp[16];     // peers' current buffers (IPC mapped or in-process), index = rank
  double2* peer_alt[16];
  bool peers_attached;
  bool ipc_opened;
  bt_barrier_fn barrier; void* barrier_ctx;
  bt_allreduce_fn allreduce; void* allreduce_ctx;
  uint64_t n_remaps, remap_bytes; float remap_ms;
  // device-side remap synchronisation (multi-process shards): a 4 KB flag page per shard (its own allocation and IPC handle),
  // mapped by the peers; peer r writes its epoch into slot r (ready: [0..15], done: [16..31])
  double2* buf0;             // the first buffer as allocated (amp and alt swap roles at every remap, buf0 does not)
  uint32_t* flags;           // own flag page, or nullptr
  uint32_t* peer_flags[16];  // every rank's flag page as seen from this device
  uint32_t remap_epoch;
  uint64_t* d_remap_tab;     // digit tables of the remap in flight (built on the device)
  std::vector<cudaEvent_t>* remap_ev;  // 4 events per remap (before the ready flags, before the pull, after it, after the done flags) not read back yet
  std::vector<float>* remap_log;       // per remap: ms waiting for the peers, ms pulling, ms until everybody has finished reading
  bt_sv* local_peers[16];              // single-process shards (bt_sv_attach_local_peers): the other handles, for stream ordering
  // density-matrix view
  bool is_dm; int dm_n;
};

struct bt_dm {
  bt_sv* v;   // 2n-qubit vector
  int n;
};

// logical bit of a 1-based qubit label (qubit 1 = MSB): src/bit.jl:9-15
static inline int bt_bit_of_qubit(const bt_sv* s, int q) { return s->n_qubits - q; }

// ---- canonical gate description handed to the kernels ---------------------------------------------------
// k target bits (matrix index bit t <-> physical bit tb[t]), nc control bits (all must be 1).
struct GateDesc {
  int k;            // 0..4
  int nc;           // number of controls
  int tb[4];
  int cb[4];
  bool diag;        // matrix is diagonal
  cplx m[256];      // row-major (1<<k) x (1<<k); for diag only the diagonal in m[0..(1<<k)-1]
};

// Reduce a dense (1<<k)x(1<<k) row-major matrix acting on physical bits tb[] (plus explicit controls) to
// canonical form: index bits on which the matrix is the identity unless the bit is 1 become controls.
void bt_canonicalize(int k, const int* tb, const cplx* m_rowmajor, int nc, const int* cb, GateDesc* out);

// launchers (bt_gates.cu).  cond_outcome != nullptr => only trajectories with cond_outcome[t] == want.
int bt_launch_gate(bt_sv* s, const GateDesc& g, const int32_t* cond_outcome = nullptr, int want = 0);
// per-trajectory matrices (dense, k targets, row-major, stride 64 entries) read from device memory
int bt_launch_gate_devmat(bt_sv* s, int k, const int* tb, const double2* d_mats);

// reductions (bt_reduce.cu): results land in s->d_res (n_batch x nvals doubles); sync_to_host copies to h_res
int bt_reduce_rdm(const bt_sv* s, int k, const int* tb /* matrix index bit t <-> physical bit tb[t] */);
int bt_reduce_norm2(const bt_sv* s);
int bt_results_to_host(const bt_sv* s, size_t n_doubles);
int bt_ensure_traj(bt_sv* s);
int bt_ensure_partials(bt_sv* s, size_t doubles);
int bt_ensure_alt(bt_sv* s);
int bt_ensure_scratch(bt_sv* s, size_t bytes);

// fused multi-gate pass (bt_tile.cu)
int bt_apply_fused(bt_sv* s, const std::vector<GateDesc>& gates);

// kernel classes for the per-launch profile
#define BT_CLS_TILE 0
#define BT_CLS_DENSE 1
#define BT_CLS_DIAG 2
#define BT_CLS_OTHER 3
void bt_prof_begin(bt_sv* s, int cls);
void bt_prof_end(bt_sv* s);

// helpers
static inline cplx c64(const bt_c64& z) { return cplx(z.re, z.im); }
int bt_check_sv(const bt_sv* s);
int bt_sv_create_internal(int n_qubits, int n_local, int64_t n_batch, bool want_alt, bt_sv** out);

// local physical bit for a logical bit, or -1-(rank bit index) if global
static inline int bt_phys(const bt_sv* s, int logical_bit) { return s->phys_of_bit[logical_bit]; }
