// bt_state.cu -- handle life cycle, host<->device transfers, error plumbing.
// Replaces the state constructors of src/hilbert.jl:835-882 for a device-resident state.
#include "bt_internal.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";
static int g_strict = 0;

void bt_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* bt_last_error(void) { return g_err; }
extern "C" int bt_version(void) { return 100; }
extern "C" int bt_set_strict(int strict) { g_strict = strict; return BT_OK; }
int bt_is_strict() { return g_strict; }

extern "C" int bt_device_count(int* n) {
  if (!n) BT_FAIL(BT_ERR_ARG, "bt_device_count: null output");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) { *n = 0; BT_FAIL(BT_ERR_CUDA, "no CUDA device: %s", cudaGetErrorString(e)); }
  *n = c;
  return BT_OK;
}

extern "C" int bt_set_device(int device) {
  BT_CUDA(cudaSetDevice(device));
  return BT_OK;
}

int bt_check_sv(const bt_sv* s) {
  if (!s) BT_FAIL(BT_ERR_ARG, "null state handle");
  BT_CUDA(cudaSetDevice(s->device));
  return BT_OK;
}

// ---- kernels --------------------------------------------------------------------------------------------
__global__ void k_set_basis(double2* a, uint64_t per_traj, uint64_t len, uint64_t index, int hit) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < len; i += stride) {
    uint64_t r = i & (per_traj - 1);
    a[i] = make_double2((hit && r == index) ? 1.0 : 0.0, 0.0);
  }
}

__global__ void k_fill(double2* a, uint64_t len, double2 v) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < len; i += stride) a[i] = v;
}

static int grid_for(uint64_t n, int block) {
  uint64_t g = (n + block - 1) / block;
  uint64_t cap = 148ull * 16;
  return (int)std::min<uint64_t>(std::max<uint64_t>(g, 1), cap);
}

// ---- handle pool -------------------------------------------------------------------------------------------------
// The reference's shot loops (src/ops.jl:619-630, :671-676) create a fresh zero_state per shot.  Creating a handle costs a
// stream, events, pinned staging and several cudaMallocs (hundreds of microseconds), which dominates small circuits, so
// destroyed unsharded handles of up to 64 MiB are parked here and re-used by the next create of the same shape.
#include <mutex>
static std::mutex g_pool_mu;
static std::vector<bt_sv*> g_pool;
static const size_t POOL_MAX_HANDLES = 16;
static const uint64_t POOL_MAX_LEN = 1ull << 22;

static bt_sv* pool_take(int n_qubits, int n_local, int64_t n_batch, int device) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  for (size_t i = 0; i < g_pool.size(); ++i) {
    bt_sv* s = g_pool[i];
    if (s->n_qubits == n_qubits && s->n_local == n_local && s->n_batch == n_batch && s->device == device) {
      g_pool.erase(g_pool.begin() + i);
      return s;
    }
  }
  return nullptr;
}

static void really_destroy(bt_sv* s);

static bool pool_put(bt_sv* s) {
  if (s->world != 1 || s->is_dm || s->alt || s->len > POOL_MAX_LEN || s->prof_ev) return false;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (g_pool.size() >= POOL_MAX_HANDLES) return false;
  g_pool.push_back(s);
  return true;
}

int bt_sv_create_internal(int n_qubits, int n_local, int64_t n_batch, bool want_alt, bt_sv** out) {
  if (!out) BT_FAIL(BT_ERR_ARG, "null output handle");
  *out = nullptr;
  if (n_qubits < 1 || n_qubits > 62) BT_FAIL(BT_ERR_ARG, "n_qubits=%d out of range", n_qubits);
  if (n_local < 1 || n_local > n_qubits || n_local > 40) BT_FAIL(BT_ERR_ARG, "n_local=%d out of range", n_local);
  if (n_batch < 1) BT_FAIL(BT_ERR_ARG, "n_batch must be >= 1");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) BT_FAIL(BT_ERR_CUDA, "no CUDA device available (%s); this backend has no CPU fallback", cudaGetErrorString(e));
  if (!want_alt && n_local == n_qubits) {
    int dev = 0;
    BT_CUDA(cudaGetDevice(&dev));
    if (bt_sv* r = pool_take(n_qubits, n_local, n_batch, dev)) {
      r->launches = 0;
      r->mask_on = false;
      r->mlog_on = false; r->mlog_len = 0;
      for (int b = 0; b < 64; ++b) r->phys_of_bit[b] = b;
      k_set_basis<<<grid_for(r->len, 256), 256, 0, r->stream>>>(r->amp, 1ull << n_local, r->len, 0, 1);
      BT_CHECK_LAUNCH(r);
      *out = r;
      return BT_OK;
    }
  }
  bt_sv* s = new bt_sv();
  memset(s, 0, sizeof(*s));
  BT_CUDA(cudaGetDevice(&s->device));
  s->n_qubits = n_qubits;
  s->n_local = n_local;
  s->n_batch = n_batch;
  s->len = (uint64_t)n_batch << n_local;
  s->rank = 0; s->world = 1; s->g = 0;
  for (int b = 0; b < 64; ++b) s->phys_of_bit[b] = b;
  // Both shard buffers are allocated with exactly the same power-of-two size.  (Round 1 put the 4 KB flag page of the device-side
  // remap synchronisation behind the amplitudes of the first buffer; a 32 GiB + 4 KB allocation is mapped differently from a
  // 32 GiB one, and every remap that read the odd-sized buffer ran 4-12x slower at 4 and 8 GPUs -- profiles/r2_remap_diag_n4.txt.)
  cudaError_t ea = cudaMalloc(&s->amp, s->len * sizeof(double2));
  if (ea != cudaSuccess) { delete s; cudaGetLastError(); BT_FAIL(BT_ERR_ALLOC, "cudaMalloc of %llu bytes failed: %s", (unsigned long long)(s->len * sizeof(double2)), cudaGetErrorString(ea)); }
  if (want_alt) {
    ea = cudaMalloc(&s->alt, s->len * sizeof(double2));
    if (ea != cudaSuccess) { cudaFree(s->amp); delete s; cudaGetLastError(); BT_FAIL(BT_ERR_ALLOC, "cudaMalloc (second buffer) failed: %s", cudaGetErrorString(ea)); }
  }
  BT_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  s->buf0 = s->amp;
  if (want_alt) {  // flag page of the device-side remap synchronisation (bt_dist.cu): its own small allocation, exported by its own IPC handle
    // a whole 2 MiB granule: the IPC handle then names an allocation of its own, not a slice of a page shared with other small buffers
    BT_CUDA(cudaMalloc(&s->flags, (size_t)2 << 20));
    BT_CUDA(cudaMemset(s->flags, 0, BT_FLAG_PAGE_BYTES));
  }
  BT_CUDA(cudaEventCreate(&s->ev0));
  BT_CUDA(cudaEventCreate(&s->ev1));
  s->res_cap = std::max<size_t>(4096, (size_t)n_batch * 128);
  BT_CUDA(cudaMalloc(&s->d_res, s->res_cap * sizeof(double)));
  BT_CUDA(cudaMallocHost(&s->h_res, s->res_cap * sizeof(double)));
  BT_CUDA(cudaMalloc(&s->d_err, sizeof(int32_t)));
  BT_CUDA(cudaMemsetAsync(s->d_err, 0, sizeof(int32_t), s->stream));
  BT_CUDA(cudaMallocHost(&s->h_flag, 4 * sizeof(int32_t)));
  k_set_basis<<<grid_for(s->len, 256), 256, 0, s->stream>>>(s->amp, 1ull << n_local, s->len, 0, 1);
  BT_CHECK_LAUNCH(s);
  *out = s;
  return BT_OK;
}

extern "C" int bt_sv_create(int n_qubits, int64_t n_batch, bt_sv** out) {
  return bt_sv_create_internal(n_qubits, n_qubits, n_batch, false, out);
}

extern "C" int bt_sv_destroy(bt_sv* s) {
  if (!s) return BT_OK;
  cudaSetDevice(s->device);
  cudaStreamSynchronize(s->stream);
  if (pool_put(s)) return BT_OK;
  really_destroy(s);
  return BT_OK;
}

extern "C" int bt_pool_release(void) {
  std::vector<bt_sv*> v;
  { std::lock_guard<std::mutex> lk(g_pool_mu); v.swap(g_pool); }
  for (bt_sv* s : v) { cudaSetDevice(s->device); really_destroy(s); }
  return BT_OK;
}

static void really_destroy(bt_sv* s) {
  if (s->ipc_opened) {
    for (int r = 0; r < s->world; ++r) {
      if (r == s->rank) continue;
      if (s->peer_amp[r]) cudaIpcCloseMemHandle(s->peer_amp[r]);
      if (s->peer_alt[r]) cudaIpcCloseMemHandle(s->peer_alt[r]);
      if (s->peer_flags[r]) cudaIpcCloseMemHandle(s->peer_flags[r]);
    }
  }
  if (s->flags) cudaFree(s->flags);
  if (s->remap_ev) { for (cudaEvent_t e : *s->remap_ev) cudaEventDestroy(e); delete s->remap_ev; }
  delete s->remap_log;
  if (s->d_remap_tab) cudaFree(s->d_remap_tab);
  cudaFree(s->amp);
  if (s->alt) cudaFree(s->alt);
  if (s->d_part) cudaFree(s->d_part);
  if (s->d_scratch) cudaFree(s->d_scratch);
  cudaFree(s->d_res);
  cudaFreeHost(s->h_res);
  if (s->d_u) cudaFree(s->d_u);
  if (s->d_outcome) cudaFree(s->d_outcome);
  if (s->d_mats) cudaFree(s->d_mats);
  if (s->d_scale) cudaFree(s->d_scale);
  if (s->d_mask) cudaFree(s->d_mask);
  if (s->d_mlog) cudaFree(s->d_mlog);
  cudaFree(s->d_err);
  cudaFreeHost(s->h_flag);
  cudaEventDestroy(s->ev0);
  cudaEventDestroy(s->ev1);
  if (s->prof_ev) { for (cudaEvent_t e : *s->prof_ev) cudaEventDestroy(e); delete s->prof_ev; delete s->prof_cls; }
  cudaStreamDestroy(s->stream);
  delete s;
}

int bt_ensure_traj(bt_sv* s) {
  if (s->traj_cap >= (size_t)s->n_batch) return BT_OK;
  size_t nb = (size_t)s->n_batch;
  BT_CUDA(cudaMalloc(&s->d_u, nb * sizeof(double)));
  BT_CUDA(cudaMalloc(&s->d_outcome, nb * sizeof(int32_t)));
  BT_CUDA(cudaMemsetAsync(s->d_outcome, 0, nb * sizeof(int32_t), s->stream));
  BT_CUDA(cudaMalloc(&s->d_mats, nb * 64 * sizeof(double2)));
  BT_CUDA(cudaMalloc(&s->d_scale, nb * 2 * sizeof(double)));
  BT_CUDA(cudaMalloc(&s->d_mask, nb * sizeof(int32_t)));
  s->traj_cap = nb;
  return BT_OK;
}

int bt_ensure_partials(bt_sv* s, size_t doubles) {
  if (s->part_cap >= doubles) return BT_OK;
  if (s->d_part) { BT_CUDA(cudaStreamSynchronize(s->stream)); BT_CUDA(cudaFree(s->d_part)); s->d_part = nullptr; }
  BT_CUDA(cudaMalloc(&s->d_part, doubles * sizeof(double)));
  s->part_cap = doubles;
  return BT_OK;
}

int bt_ensure_scratch(bt_sv* s, size_t bytes) {
  if (s->scratch_cap >= bytes) return BT_OK;
  if (s->d_scratch) { BT_CUDA(cudaStreamSynchronize(s->stream)); BT_CUDA(cudaFree(s->d_scratch)); s->d_scratch = nullptr; s->scratch_cap = 0; }
  size_t cap = std::max<size_t>(bytes, 1 << 20);
  BT_CUDA(cudaMalloc(&s->d_scratch, cap));
  s->scratch_cap = cap;
  return BT_OK;
}

int bt_ensure_alt(bt_sv* s) {
  if (s->alt) return BT_OK;
  cudaError_t e = cudaMalloc(&s->alt, s->len * sizeof(double2));
  if (e != cudaSuccess) { cudaGetLastError(); BT_FAIL(BT_ERR_ALLOC, "scratch buffer of %llu bytes: %s", (unsigned long long)(s->len * 16), cudaGetErrorString(e)); }
  return BT_OK;
}

extern "C" int bt_sv_n_qubits(const bt_sv* s, int* n) {
  if (!s || !n) BT_FAIL(BT_ERR_ARG, "null argument");
  *n = s->n_qubits;
  return BT_OK;
}

extern "C" int bt_sv_set_basis(bt_sv* s, uint64_t index) {
  BT_TRY(bt_check_sv(s));
  if (s->n_qubits < 64 && index >= (1ull << s->n_qubits)) BT_FAIL(BT_ERR_ARG, "basis index out of range");
  // sharded: physical layout is reset to identity, the owning rank holds the 1
  for (int b = 0; b < 64; ++b) s->phys_of_bit[b] = b;
  uint64_t local = index & ((1ull << s->n_local) - 1);
  int hit = (int)((index >> s->n_local) == (uint64_t)s->rank);
  k_set_basis<<<grid_for(s->len, 256), 256, 0, s->stream>>>(s->amp, 1ull << s->n_local, s->len, local, hit);
  BT_CHECK_LAUNCH(s);
  return BT_OK;
}

// Trajectory mask: while set, bt_sv_apply_1q/2q/3q, bt_sv_apply_circuit (unfused then), bt_sv_kraus and bt_sv_measure_z touch
// only the trajectories with mask[t] != 0 -- the branch of an ifOp (src/struct.jl:587-590) on a batch whose trajectories
// measured different outcomes.  Masked-out trajectories report outcome / chosen = -1.  NULL clears the mask.
extern "C" int bt_sv_set_mask(bt_sv* s, const int32_t* mask) {
  BT_TRY(bt_check_sv(s));
  if (!mask) { s->mask_on = false; return BT_OK; }
  BT_TRY(bt_ensure_traj(s));
  BT_CUDA(cudaMemcpyAsync(s->d_mask, mask, s->n_batch * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
  BT_CUDA(cudaStreamSynchronize(s->stream));  // the caller's buffer may be pageable and short-lived
  s->mask_on = true;
  return BT_OK;
}

extern "C" int bt_sv_set_plus(bt_sv* s) {
  BT_TRY(bt_check_sv(s));
  double v = 1.0 / sqrt(ldexp(1.0, s->n_qubits));  // (1+0im)/sqrt(2^N)  src/hilbert.jl:869
  k_fill<<<grid_for(s->len, 256), 256, 0, s->stream>>>(s->amp, s->len, make_double2(v, 0.0));
  BT_CHECK_LAUNCH(s);
  return BT_OK;
}

extern "C" int bt_sv_upload(bt_sv* s, const bt_c64* host, uint64_t len) {
  BT_TRY(bt_check_sv(s));
  if (!host || len != s->len) BT_FAIL(BT_ERR_ARG, "upload length %llu != %llu", (unsigned long long)len, (unsigned long long)s->len);
  BT_CUDA(cudaMemcpyAsync(s->amp, host, len * sizeof(double2), cudaMemcpyHostToDevice, s->stream));
  BT_CUDA(cudaStreamSynchronize(s->stream));
  return BT_OK;
}

extern "C" int bt_sv_download(const bt_sv* s, bt_c64* host, uint64_t len) {
  BT_TRY(bt_check_sv(s));
  if (!host || len != s->len) BT_FAIL(BT_ERR_ARG, "download length %llu != %llu", (unsigned long long)len, (unsigned long long)s->len);
  BT_CUDA(cudaMemcpyAsync(host, s->amp, len * sizeof(double2), cudaMemcpyDeviceToHost, s->stream));
  BT_CUDA(cudaStreamSynchronize(s->stream));
  return BT_OK;
}

extern "C" int bt_sv_copy(bt_sv* dst, const bt_sv* src) {
  BT_TRY(bt_check_sv(dst));
  if (!src || src->len != dst->len || src->n_qubits != dst->n_qubits) BT_FAIL(BT_ERR_ARG, "bt_sv_copy: shape mismatch");
  BT_CUDA(cudaStreamSynchronize(src->stream));
  BT_CUDA(cudaMemcpyAsync(dst->amp, src->amp, dst->len * sizeof(double2), cudaMemcpyDeviceToDevice, dst->stream));
  // the copy runs on dst's stream: later work enqueued on src's stream (which may overwrite the source) must not overtake it
  BT_CUDA(cudaStreamSynchronize(dst->stream));
  memcpy(dst->phys_of_bit, src->phys_of_bit, sizeof(dst->phys_of_bit));
  return BT_OK;
}

extern "C" int bt_sv_sync(const bt_sv* s) {
  BT_TRY(bt_check_sv(s));
  BT_CUDA(cudaStreamSynchronize(s->stream));
  return BT_OK;
}

extern "C" int bt_sv_timer_start(bt_sv* s) {
  BT_TRY(bt_check_sv(s));
  BT_CUDA(cudaEventRecord(s->ev0, s->stream));
  return BT_OK;
}

extern "C" int bt_sv_timer_stop(bt_sv* s, float* ms) {
  BT_TRY(bt_check_sv(s));
  if (!ms) BT_FAIL(BT_ERR_ARG, "null output");
  BT_CUDA(cudaEventRecord(s->ev1, s->stream));
  BT_CUDA(cudaEventSynchronize(s->ev1));
  BT_CUDA(cudaEventElapsedTime(ms, s->ev0, s->ev1));
  return BT_OK;
}

extern "C" int bt_sv_launch_count(const bt_sv* s, uint64_t* n) {
  if (!s || !n) BT_FAIL(BT_ERR_ARG, "null argument");
  *n = s->launches;
  return BT_OK;
}

int bt_results_to_host(const bt_sv* s, size_t n_doubles) {
  if (n_doubles > s->res_cap) BT_FAIL(BT_ERR_ARG, "result buffer too small");
  BT_CUDA(cudaMemcpyAsync(s->h_res, s->d_res, n_doubles * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  BT_CUDA(cudaStreamSynchronize(s->stream));
  return BT_OK;
}

// ---- per-launch profile -------------------------------------------------------------------------------------------
void bt_prof_begin(bt_sv* s, int cls) {
  if (!s->prof_on) return;
  if (s->prof_used + 2 > s->prof_ev->size()) {
    for (int i = 0; i < 512; ++i) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) { s->prof_on = false; return; } s->prof_ev->push_back(e); }
  }
  cudaEventRecord((*s->prof_ev)[s->prof_used], s->stream);
  s->prof_cls->push_back(cls);
}
void bt_prof_end(bt_sv* s) {
  if (!s->prof_on) return;
  cudaEventRecord((*s->prof_ev)[s->prof_used + 1], s->stream);
  s->prof_used += 2;
}

extern "C" int bt_sv_profile_enable(bt_sv* s, int on) {
  BT_TRY(bt_check_sv(s));
  if (!s->prof_ev) { s->prof_ev = new std::vector<cudaEvent_t>(); s->prof_cls = new std::vector<int>(); }
  BT_CUDA(cudaStreamSynchronize(s->stream));
  s->prof_used = 0;
  s->prof_cls->clear();
  s->prof_on = on != 0;
  return BT_OK;
}

// counts[4], ms[4]: launches and summed CUDA-event durations per kernel class since the last enable
extern "C" int bt_sv_profile_read(bt_sv* s, uint64_t* counts, double* ms) {
  BT_TRY(bt_check_sv(s));
  if (!counts || !ms) BT_FAIL(BT_ERR_ARG, "null output");
  for (int c = 0; c < 4; ++c) { counts[c] = 0; ms[c] = 0.0; }
  if (!s->prof_ev) return BT_OK;
  BT_CUDA(cudaStreamSynchronize(s->stream));
  for (size_t i = 0; i + 1 < s->prof_used; i += 2) {
    float t = 0.f;
    BT_CUDA(cudaEventElapsedTime(&t, (*s->prof_ev)[i], (*s->prof_ev)[i + 1]));
    int c = (*s->prof_cls)[i / 2];
    counts[c]++; ms[c] += t;
  }
  return BT_OK;
}

// ---- FP64 FMA peak of the current device, measured (bench.py's FP64 roofline denominator) ------------------------------
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 0.5;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// Dependent-chain-free DFMA loop on every SM (8 independent chains per thread, 8 CTAs of 256 threads per SM), best of `reps`
// launches of ~`ms_target` ms each, timed with CUDA events: the FP64 vector peak bench.py divides by (no tensor-core FP64 path is used).
extern "C" int bt_fp64_peak(double* tflops, double* ms_per_launch, int reps) {
  if (!tflops) BT_FAIL(BT_ERR_ARG, "null output");
  int dev = 0, sms = 0;
  BT_CUDA(cudaGetDevice(&dev));
  BT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8, threads = 256, iters = 40000;
  double* d = nullptr;
  BT_CUDA(cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  BT_CUDA(cudaEventCreate(&e0));
  BT_CUDA(cudaEventCreate(&e1));
  k_fp64_peak<<<blocks, threads>>>(d, 1000);
  float best = 1e30f;
  for (int r = 0; r < std::max(1, reps); ++r) {
    cudaEventRecord(e0);
    k_fp64_peak<<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  BT_CUDA(cudaGetLastError());
  *tflops = 2.0 * 8.0 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
  if (ms_per_launch) *ms_per_launch = best;
  return BT_OK;
}
