// bt_dist.cu -- multi-GPU shards.  No reference analogue (the reference is single-threaded Julia); the closest
// idea is its label-permutation helper relabel_swap (src/hilbert.jl:266-317).
//
// Layout: physical index = (rank << n_local) | local.  A logical->physical bit map is kept per handle.  Gates whose
// non-diagonal targets are all local need no communication; diagonal gates and controls on rank bits are resolved
// from the rank id (bt_gates.cu: localize).  Otherwise the qubits are re-mapped: every rank PULLS its new shard
// straight out of its peers' HBM over NVLink (CUDA IPC / peer mappings) with an arbitrary bit permutation fused into
// the same kernel -- exchange + local transposition in one pass, no pack/unpack buffers, no NCCL staging.
#include "bt_internal.cuh"

#define REMAP_KEEP 5        // the 5 lowest physical bits never move: every lane of a warp reads one contiguous 512 B run
#define REMAP_TBITS 10      // permutation tables of 2^10 entries per 10-bit digit of the destination index
#define REMAP_NT 4          // 4 digits cover 40 bits

struct RemapParams {
  const double2* src[16];
  int n_local;
  int rank;
  // Iteration order.  The kernel walks a LOOP index; sel_n of its bits, starting at bit sel_lo, are the destination bits
  // that decide which rank an amplitude comes from (sel_pos[], ascending), the other loop bits fill the remaining destination
  // bits in order.  Consecutive 2^sel_lo-amplitude chunks of the walk therefore come from different ranks: every rank reads
  // from all of its peers all the time, and (rot = own rank bits) no two ranks start on the same peer.  A linear walk of the
  // destination (sel_n = 0) makes all ranks read one peer's HBM at a time -- an incast on that peer's NVLink egress.
  int sel_n, sel_lo;
  int sel_pos[4];
  uint32_t rot;
};

// loop index (local bits, rotation already applied) -> destination index
__host__ __device__ static inline uint64_t remap_dest(uint64_t l, int sel_n, int sel_lo, const int* sel_pos) {
  if (sel_n == 0) return l;
  const uint64_t low = l & ((1ull << sel_lo) - 1);
  const uint64_t selv = (l >> sel_lo) & ((1ull << sel_n) - 1);
  uint64_t x = low | ((l >> (sel_lo + sel_n)) << sel_lo);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j < sel_n) {
      const int p = sel_pos[j];
      x = ((x >> p) << (p + 1)) | (x & ((1ull << p) - 1)) | (((selv >> j) & 1ull) << p);
    }
  return x;
}

// dst[D(l)] = src_rank(S)[local(S)] with S = T0[d0] | T1[d1] | T2[d2] | T3[d3], d* = digits of ((rank << n_local) | l), l = loop index
__global__ void __launch_bounds__(256) k_remap_pull(double2* __restrict__ dst, uint64_t n, int64_t n_batch,
                                                     const __grid_constant__ RemapParams P, const uint64_t* __restrict__ tables) {
  __shared__ uint64_t T[REMAP_NT << REMAP_TBITS];
  for (int i = threadIdx.x; i < (REMAP_NT << REMAP_TBITS); i += blockDim.x) T[i] = tables[i];
  __syncthreads();
  const uint64_t lmask = (1ull << P.n_local) - 1;
  const uint64_t rank_hi = (uint64_t)P.rank << P.n_local;
  const uint64_t rotm = (uint64_t)P.rot << P.sel_lo;
  const uint64_t total = n * (uint64_t)n_batch;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t dm = (1ull << REMAP_TBITS) - 1;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // 4 independent 16-byte loads in flight per thread: NVLink round trips are ~2-4 us
  for (; i + 3 * stride < total; i += 4 * stride) {
    double2 v[4];
    uint64_t o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint64_t idx = i + u * stride;
      uint64_t t = idx >> P.n_local;  // trajectory
      uint64_t l = (idx & lmask) ^ rotm;
      uint64_t d = rank_hi | l;
      uint64_t s = T[d & dm] | T[(1 << REMAP_TBITS) + ((d >> REMAP_TBITS) & dm)] | T[(2 << REMAP_TBITS) + ((d >> (2 * REMAP_TBITS)) & dm)] |
                   T[(3 << REMAP_TBITS) + ((d >> (3 * REMAP_TBITS)) & dm)];
      o[u] = (t << P.n_local) | remap_dest(l, P.sel_n, P.sel_lo, P.sel_pos);
      v[u] = P.src[s >> P.n_local][(t << P.n_local) | (s & lmask)];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) dst[o[u]] = v[u];
  }
  for (; i < total; i += stride) {
    uint64_t t = i >> P.n_local;
    uint64_t l = (i & lmask) ^ rotm;
    uint64_t d = rank_hi | l;
    uint64_t s = T[d & dm] | T[(1 << REMAP_TBITS) + ((d >> REMAP_TBITS) & dm)] | T[(2 << REMAP_TBITS) + ((d >> (2 * REMAP_TBITS)) & dm)] |
                 T[(3 << REMAP_TBITS) + ((d >> (3 * REMAP_TBITS)) & dm)];
    dst[(t << P.n_local) | remap_dest(l, P.sel_n, P.sel_lo, P.sel_pos)] = P.src[s >> P.n_local][(t << P.n_local) | (s & lmask)];
  }
}

// T[t][v] = OR over the set bits j of v of (1 << src_of_loop_bit[t * REMAP_TBITS + j]): where the t-th digit of a loop index comes from
struct RemapSigma { int s[64]; };
__global__ void k_remap_tables(uint64_t* __restrict__ tab, const __grid_constant__ RemapSigma SG) {
  const int t = blockIdx.x;
  const uint32_t v = threadIdx.x;
  uint64_t o = 0;
#pragma unroll
  for (int j = 0; j < REMAP_TBITS; ++j) {
    const int src = SG.s[t * REMAP_TBITS + j];
    if (src >= 0 && ((v >> j) & 1u)) o |= 1ull << src;
  }
  tab[((size_t)t << REMAP_TBITS) + v] = o;
}

// ---- device-side synchronisation of a remap (multi-process shards, one GPU per process) --------------------------------
// A remap needs two agreements between the ranks: "everyone's previous kernels have finished writing the buffer I am about to
// read" and "everyone has finished reading the buffer I am about to overwrite".  Made on the host (stream synchronise + barrier,
// twice per remap) they stop the host from running ahead of the GPU, and every delay of a host thread at one of those points idles
// all GPUs.  Here both are epoch flags in peer memory: after its segment's kernels a rank stores the remap's epoch into its slot of
// every peer's flag page (k_flag_signal, st.release.sys over NVLink); the pull is preceded by k_flag_wait, which polls the own page
// (ld.acquire.sys) until every slot has reached the epoch.  Everything is stream-ordered; the host never blocks in a remap.
struct FlagPeers { uint32_t* page[16]; };

__device__ __forceinline__ unsigned long long bt_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// which = 0: "ready" slots [0..15], which = 1: "done" slots [16..31]
__global__ void k_flag_signal(const __grid_constant__ FlagPeers P, int world, int rank, int which, uint32_t epoch) {
  const int r = threadIdx.x;
  if (r < world) {
    __threadfence_system();
    uint32_t* p = P.page[r] + which * 16 + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(epoch) : "memory");
  }
}

__global__ void k_flag_wait(const uint32_t* __restrict__ page, int world, int which, uint32_t epoch, int32_t* __restrict__ err, unsigned long long timeout_ns) {
  const int r = threadIdx.x;
  if (r < world) {
    const uint32_t* p = page + which * 16 + r;
    const unsigned long long t0 = bt_globaltimer();
    for (;;) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
      if ((int32_t)(v - epoch) >= 0) break;
      __nanosleep(256);
      if (bt_globaltimer() - t0 > timeout_ns) {  // a peer died or never reached this remap: fail loudly instead of hanging the box
        *err = 7;
        __threadfence_system();
        __trap();
      }
    }
  }
  __syncwarp();
  __threadfence_system();
}

extern "C" int bt_sv_create_shard(int n_qubits_total, int rank, int world, bt_sv** out) {
  if (world < 1 || world > 16 || (world & (world - 1))) BT_FAIL(BT_ERR_ARG, "world must be a power of two <= 16");
  if (rank < 0 || rank >= world) BT_FAIL(BT_ERR_ARG, "rank out of range");
  int g = 0;
  while ((1 << g) < world) ++g;
  if (n_qubits_total - g < REMAP_KEEP + g) BT_FAIL(BT_ERR_ARG, "too few qubits (%d) for %d shards", n_qubits_total, world);
  BT_TRY(bt_sv_create_internal(n_qubits_total, n_qubits_total - g, 1, world > 1, out));
  bt_sv* s = *out;
  s->rank = rank; s->world = world; s->g = g;
  for (int r = 0; r < 16; ++r) { s->peer_amp[r] = nullptr; s->peer_alt[r] = nullptr; }
  s->peer_amp[rank] = s->amp;
  s->peer_alt[rank] = s->alt;
  for (int r = 0; r < 16; ++r) s->peer_flags[r] = nullptr;
  s->peer_flags[rank] = s->flags;
  if (world == 1) s->peers_attached = true;
  // |0..0> lives on rank 0 only
  return bt_sv_set_basis(s, 0);
}

extern "C" int bt_sv_ipc_export(bt_sv* s, void* handles) {
  BT_TRY(bt_check_sv(s));
  if (!handles) BT_FAIL(BT_ERR_ARG, "null output");
  if (!s->alt) BT_FAIL(BT_ERR_ARG, "not a shard handle");
  static_assert(sizeof(cudaIpcMemHandle_t) <= BT_IPC_HANDLE_BYTES, "IPC handle size");
  if (s->amp != s->buf0) BT_FAIL(BT_ERR_ARG, "export the IPC handles before the first remap");
  if (!s->flags) BT_FAIL(BT_ERR_ARG, "not a shard handle (no flag page)");
  cudaIpcMemHandle_t h0, h1, h2;
  BT_CUDA(cudaIpcGetMemHandle(&h0, s->amp));
  BT_CUDA(cudaIpcGetMemHandle(&h1, s->alt));
  BT_CUDA(cudaIpcGetMemHandle(&h2, s->flags));
  memset(handles, 0, BT_IPC_HANDLES_PER_SHARD * BT_IPC_HANDLE_BYTES);
  memcpy(handles, &h0, sizeof(h0));
  memcpy((char*)handles + BT_IPC_HANDLE_BYTES, &h1, sizeof(h1));
  memcpy((char*)handles + 2 * BT_IPC_HANDLE_BYTES, &h2, sizeof(h2));
  return BT_OK;
}

extern "C" int bt_sv_ipc_attach(bt_sv* s, const void* all_handles) {
  BT_TRY(bt_check_sv(s));
  if (!all_handles) BT_FAIL(BT_ERR_ARG, "null handles");
  for (int r = 0; r < s->world; ++r) {
    if (r == s->rank) continue;
    cudaIpcMemHandle_t h0, h1, h2;
    const char* base = (const char*)all_handles + (size_t)r * BT_IPC_HANDLES_PER_SHARD * BT_IPC_HANDLE_BYTES;
    memcpy(&h0, base, sizeof(h0));
    memcpy(&h1, base + BT_IPC_HANDLE_BYTES, sizeof(h1));
    memcpy(&h2, base + 2 * BT_IPC_HANDLE_BYTES, sizeof(h2));
    void *p0 = nullptr, *p1 = nullptr, *p2 = nullptr;
    BT_CUDA(cudaIpcOpenMemHandle(&p0, h0, cudaIpcMemLazyEnablePeerAccess));
    BT_CUDA(cudaIpcOpenMemHandle(&p1, h1, cudaIpcMemLazyEnablePeerAccess));
    BT_CUDA(cudaIpcOpenMemHandle(&p2, h2, cudaIpcMemLazyEnablePeerAccess));
    s->peer_amp[r] = (double2*)p0;
    s->peer_alt[r] = (double2*)p1;
    s->peer_flags[r] = (uint32_t*)p2;
  }
  s->ipc_opened = true;
  s->peers_attached = true;
  return BT_OK;
}

extern "C" int bt_sv_attach_local_peers(bt_sv** shards, int world) {
  if (!shards || world < 1) BT_FAIL(BT_ERR_ARG, "invalid shard list");
  for (int r = 0; r < world; ++r)
    if (!shards[r] || shards[r]->world != world || shards[r]->rank != r) BT_FAIL(BT_ERR_ARG, "shard %d does not match (rank/world)", r);
  for (int r = 0; r < world; ++r) {
    bt_sv* s = shards[r];
    BT_CUDA(cudaSetDevice(s->device));
    for (int q = 0; q < world; ++q) {
      if (shards[q]->device != s->device) {
        int can = 0;
        BT_CUDA(cudaDeviceCanAccessPeer(&can, s->device, shards[q]->device));
        if (!can) BT_FAIL(BT_ERR_CUDA, "device %d cannot access device %d", s->device, shards[q]->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(shards[q]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) BT_CUDA(e);
        cudaGetLastError();
      }
      s->peer_amp[q] = shards[q]->amp;
      s->peer_alt[q] = shards[q]->alt;
      s->local_peers[q] = shards[q];
    }
    s->peers_attached = true;
  }
  return BT_OK;
}

extern "C" int bt_sv_set_barrier(bt_sv* s, bt_barrier_fn fn, void* ctx) {
  if (!s) BT_FAIL(BT_ERR_ARG, "null handle");
  s->barrier = fn; s->barrier_ctx = ctx;
  return BT_OK;
}

extern "C" int bt_sv_set_allreduce(bt_sv* s, bt_allreduce_fn fn, void* ctx) {
  if (!s) BT_FAIL(BT_ERR_ARG, "null handle");
  s->allreduce = fn; s->allreduce_ctx = ctx;
  return BT_OK;
}

extern "C" int bt_sv_layout(const bt_sv* s, int* phys) {
  if (!s || !phys) BT_FAIL(BT_ERR_ARG, "null argument");
  for (int b = 0; b < s->n_qubits; ++b) phys[b] = s->phys_of_bit[b];
  return BT_OK;
}

// durations of the remaps enqueued without a host synchronisation: read back once their last event has completed
static void remap_log_push(bt_sv* s, float wait_ms, float pull_ms, float done_ms) {
  if (!s->remap_log) s->remap_log = new std::vector<float>();
  if (s->remap_log->size() >= 3 * 4096) s->remap_log->erase(s->remap_log->begin(), s->remap_log->begin() + 3 * 2048);
  s->remap_log->push_back(wait_ms); s->remap_log->push_back(pull_ms); s->remap_log->push_back(done_ms);
  s->remap_ms += pull_ms;
}

static void drain_remap_events(bt_sv* s, bool wait) {
  if (!s->remap_ev) return;
  std::vector<cudaEvent_t>& v = *s->remap_ev;
  size_t keep = 0;
  for (size_t i = 0; i + 3 < v.size(); i += 4) {
    const bool done = wait ? (cudaEventSynchronize(v[i + 3]) == cudaSuccess) : (cudaEventQuery(v[i + 3]) == cudaSuccess);
    if (done) {
      float w = 0, p = 0, d = 0;
      cudaEventElapsedTime(&w, v[i], v[i + 1]);
      cudaEventElapsedTime(&p, v[i + 1], v[i + 2]);
      cudaEventElapsedTime(&d, v[i + 2], v[i + 3]);
      remap_log_push(s, w, p, d);
      for (int k = 0; k < 4; ++k) cudaEventDestroy(v[i + k]);
    } else {
      for (int k = 0; k < 4; ++k) v[keep++] = v[i + k];
    }
  }
  v.resize(keep);
  cudaGetLastError();
}

extern "C" int bt_sv_remap_stats(const bt_sv* s, uint64_t* n_remaps, uint64_t* bytes_remote, float* ms_total) {
  if (!s) BT_FAIL(BT_ERR_ARG, "null handle");
  drain_remap_events(const_cast<bt_sv*>(s), true);
  if (n_remaps) *n_remaps = s->n_remaps;
  if (bytes_remote) *bytes_remote = s->remap_bytes;
  if (ms_total) *ms_total = s->remap_ms;
  return BT_OK;
}

// per-remap record since the last call: (ms waiting for the peers' "ready", ms in the pull, ms until all peers are "done"); cleared on return
extern "C" int bt_sv_remap_log(bt_sv* s, int cap, float* out, int* n) {
  if (!s || !n) BT_FAIL(BT_ERR_ARG, "null argument");
  drain_remap_events(s, true);
  int have = s->remap_log ? (int)(s->remap_log->size() / 3) : 0;
  int take = std::min(have, std::max(cap, 0));
  if (out) for (int i = 0; i < 3 * take; ++i) out[i] = (*s->remap_log)[3 * (have - take) + i];
  *n = take;
  if (s->remap_log) s->remap_log->clear();
  return BT_OK;
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// Host half of a remap: validates the new layout and derives what the pull kernel needs -- sigma[dst phys bit] = src phys bit,
// the iteration order (RemapParams) and, per LOOP bit, the source bit it feeds (RemapSigma, turned into digit tables on the device).
static int remap_build(int n, int n_local, int rank, int world, const int* cur_phys, const int* new_phys, RemapParams* P, RemapSigma* SG,
                       int* sigma, bool* identity) {
  bool seen[64] = {false};
  *identity = true;
  for (int lb = 0; lb < n; ++lb) {
    int d = new_phys[lb];
    if (d < 0 || d >= n || seen[d]) BT_FAIL(BT_ERR_ARG, "layout is not a permutation");
    seen[d] = true;
    sigma[d] = cur_phys[lb];
    if (sigma[d] != d) *identity = false;
  }
  if (*identity) return BT_OK;
  for (int d = 0; d < REMAP_KEEP; ++d)
    if (sigma[d] != d) BT_FAIL(BT_ERR_ARG, "the %d lowest physical bits must stay in place (coalescing)", REMAP_KEEP);
  P->n_local = n_local;
  P->rank = rank;
  // iteration order of the pull (see RemapParams): the destination bits that select the source rank cycle fastest above sel_lo
  P->sel_n = 0; P->sel_lo = REMAP_KEEP; P->rot = 0;
  for (int j = 0; j < 4; ++j) P->sel_pos[j] = 0;
  if (world > 1 && env_int("BT_REMAP_ORDER", 1) != 0) {
    for (int d = REMAP_KEEP; d < n_local && P->sel_n < 4; ++d)
      if (sigma[d] >= n_local) P->sel_pos[P->sel_n++] = d;
    int lo = env_int("BT_REMAP_SEL_LO", 16);
    lo = std::max(lo, REMAP_KEEP);
    lo = std::min(lo, n_local - P->sel_n);
    P->sel_lo = lo;
    if (env_int("BT_REMAP_ROT", 1) != 0)
      for (int j = 0; j < P->sel_n; ++j) P->rot |= (uint32_t)((rank >> (sigma[P->sel_pos[j]] - n_local)) & 1) << j;
  }
  for (int lb = 0; lb < 64; ++lb) {
    if (lb >= n) { SG->s[lb] = -1; continue; }
    int dbit = lb;
    if (lb < n_local) {
      uint64_t img = remap_dest(1ull << lb, P->sel_n, P->sel_lo, P->sel_pos);
      dbit = 0;
      while (!((img >> dbit) & 1ull)) ++dbit;
    }
    SG->s[lb] = sigma[dbit];
  }
  return BT_OK;
}

// Pure host (no device): where the pull kernel of rank `rank` would write (dest: local index) and read (src: physical index =
// (source rank << n_local) | local) for the given loop indices.  tests/test_host_logic.py checks that the walk is a bijection
// of the shard and that src = sigma(dest) for every iteration order.
extern "C" int bt_remap_walk_host(int n_qubits, int n_local, int rank, const int* cur_phys, const int* new_phys, uint64_t n_idx, const uint64_t* loop_idx,
                                  uint64_t* dest, uint64_t* src) {
  if (!cur_phys || !new_phys || (n_idx && (!loop_idx || !dest || !src))) BT_FAIL(BT_ERR_ARG, "null argument");
  if (n_qubits < 1 || n_qubits > 62 || n_local < REMAP_KEEP || n_local > n_qubits || n_qubits - n_local > 4) BT_FAIL(BT_ERR_ARG, "invalid shard shape");
  RemapParams P; RemapSigma SG; int sigma[64]; bool identity;
  BT_TRY(remap_build(n_qubits, n_local, rank, 1 << (n_qubits - n_local), cur_phys, new_phys, &P, &SG, sigma, &identity));
  const uint64_t lmask = (1ull << n_local) - 1;
  for (uint64_t k = 0; k < n_idx; ++k) {
    if (identity) { dest[k] = loop_idx[k] & lmask; src[k] = ((uint64_t)rank << n_local) | dest[k]; continue; }
    const uint64_t l = (loop_idx[k] & lmask) ^ ((uint64_t)P.rot << P.sel_lo);
    const uint64_t d = ((uint64_t)rank << n_local) | l;
    uint64_t sidx = 0;
    for (int lb = 0; lb < n_qubits; ++lb) if ((d >> lb) & 1ull) sidx |= 1ull << SG.s[lb];
    dest[k] = remap_dest(l, P.sel_n, P.sel_lo, P.sel_pos);
    src[k] = sidx;
  }
  return BT_OK;
}

// new_phys[lb] = physical bit position of logical bit lb after the remap (a permutation of 0..n-1)
extern "C" int bt_sv_remap(bt_sv* s, const int* new_phys) {
  BT_TRY(bt_check_sv(s));
  if (!new_phys) BT_FAIL(BT_ERR_ARG, "null layout");
  int n = s->n_qubits;
  if (s->world == 1 && !s->alt) BT_TRY(bt_ensure_alt(s));
  if (!s->peers_attached) BT_FAIL(BT_ERR_ARG, "shard peers are not attached (bt_sv_ipc_attach / bt_sv_attach_local_peers)");
  RemapParams P; RemapSigma SG; int sigma[64]; bool identity;
  BT_TRY(remap_build(n, s->n_local, s->rank, s->world, s->phys_of_bit, new_phys, &P, &SG, sigma, &identity));
  if (identity) return BT_OK;
  for (int r = 0; r < 16; ++r) P.src[r] = (r < s->world) ? s->peer_amp[r] : nullptr;
  // digit tables, built on the device from the bit permutation (no host buffer has to outlive the call, no copy to wait for);
  // one table per handle: the next remap's build is stream-ordered behind this remap's pull
  if (!s->d_remap_tab) BT_CUDA(cudaMalloc(&s->d_remap_tab, ((size_t)REMAP_NT << REMAP_TBITS) * sizeof(uint64_t)));
  uint64_t* d_tab = s->d_remap_tab;
  k_remap_tables<<<REMAP_NT, 1 << REMAP_TBITS, 0, s->stream>>>(d_tab, SG);
  BT_CUDA(cudaGetLastError());
  // fraction of the new shard that comes from other ranks (for the NVLink traffic figure)
  int moved_global = 0;
  for (int d = s->n_local; d < n; ++d) if (sigma[d] < s->n_local) moved_global++;
  uint64_t nloc = 1ull << s->n_local;
  unsigned grid = (unsigned)std::min<uint64_t>((s->len / 4 + 255) / 256 + 1, 148ull * (unsigned)std::max(1, env_int("BT_REMAP_CTAS_PER_SM", 8)));
  bool dev_sync = s->world > 1 && s->ipc_opened && s->flags != nullptr;
  if (dev_sync && env_int("BT_REMAP_DEVICE_SYNC", 1) == 0) dev_sync = false;
  if (dev_sync) {
    // flags in peer memory: nothing here blocks the host (see k_flag_signal / k_flag_wait)
    cudaEvent_t ev[4];
    for (int k = 0; k < 4; ++k) BT_CUDA(cudaEventCreate(&ev[k]));
    FlagPeers F;
    for (int r = 0; r < 16; ++r) F.page[r] = (r < s->world) ? s->peer_flags[r] : nullptr;
    const uint32_t epoch = ++s->remap_epoch;
    // a peer that has not reached the remap by then has died: trap instead of hanging (BT_REMAP_TIMEOUT_S)
    const unsigned long long timeout_ns = (unsigned long long)std::max(1, env_int("BT_REMAP_TIMEOUT_S", 120)) * 1000000000ull;
    BT_CUDA(cudaEventRecord(ev[0], s->stream));
    k_flag_signal<<<1, 32, 0, s->stream>>>(F, s->world, s->rank, 0, epoch);   // my previous kernels are done: my amp may be read
    k_flag_wait<<<1, 32, 0, s->stream>>>(s->flags, s->world, 0, epoch, s->d_err, timeout_ns);
    BT_CUDA(cudaEventRecord(ev[1], s->stream));
    k_remap_pull<<<grid, 256, 0, s->stream>>>(s->alt, nloc, s->n_batch, P, d_tab);
    BT_CHECK_LAUNCH(s);
    BT_CUDA(cudaEventRecord(ev[2], s->stream));
    k_flag_signal<<<1, 32, 0, s->stream>>>(F, s->world, s->rank, 1, epoch);   // I have read everything I need from the peers
    k_flag_wait<<<1, 32, 0, s->stream>>>(s->flags, s->world, 1, epoch, s->d_err, timeout_ns);  // nobody reads my old buffer any more
    BT_CUDA(cudaEventRecord(ev[3], s->stream));
    BT_CUDA(cudaGetLastError());
    if (!s->remap_ev) s->remap_ev = new std::vector<cudaEvent_t>();
    for (int k = 0; k < 4; ++k) s->remap_ev->push_back(ev[k]);
    if (s->remap_ev->size() >= 128) drain_remap_events(s, false);
  } else {
    // everyone's previous kernels must have finished writing `amp`: shards of this process by their streams (they may have been
    // driven shard by shard, e.g. LocalShards.apply), shards of other processes through the host barrier
    cudaEvent_t t0, t1;
    BT_CUDA(cudaEventCreate(&t0));
    BT_CUDA(cudaEventCreate(&t1));
    BT_CUDA(cudaStreamSynchronize(s->stream));
    for (int r = 0; r < s->world; ++r)
      if (s->local_peers[r] && s->local_peers[r] != s) BT_CUDA(cudaStreamSynchronize(s->local_peers[r]->stream));
    if (s->barrier) s->barrier(s->barrier_ctx);
    BT_CUDA(cudaEventRecord(t0, s->stream));
    k_remap_pull<<<grid, 256, 0, s->stream>>>(s->alt, nloc, s->n_batch, P, d_tab);
    BT_CHECK_LAUNCH(s);
    BT_CUDA(cudaEventRecord(t1, s->stream));
    BT_CUDA(cudaStreamSynchronize(s->stream));
    float ms = 0;
    BT_CUDA(cudaEventElapsedTime(&ms, t0, t1));
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    remap_log_push(s, 0.f, ms, 0.f);
    // nobody may overwrite the buffer others are still reading
    if (s->barrier) s->barrier(s->barrier_ctx);
  }
  std::swap(s->amp, s->alt);
  for (int r = 0; r < s->world; ++r) std::swap(s->peer_amp[r], s->peer_alt[r]);
  for (int lb = 0; lb < n; ++lb) s->phys_of_bit[lb] = new_phys[lb];
  s->n_remaps++;
  // bytes pulled from other ranks: a fraction (1 - 2^-moved_global) of the shard
  double frac = 1.0 - ldexp(1.0, -moved_global);
  s->remap_bytes += (uint64_t)(frac * (double)s->len * 16.0);
  return BT_OK;
}

// Make the given logical bits local.  Victims are the highest local physical positions not in `need`.
int bt_prepare_local_bits(bt_sv* s, int n, const int* logical_bits) {
  if (s->world == 1) return BT_OK;
  int nq = s->n_qubits, nl = s->n_local;
  std::vector<int> need_global;
  bool is_needed[64] = {false};
  for (int i = 0; i < n; ++i) {
    is_needed[logical_bits[i]] = true;
    if (s->phys_of_bit[logical_bits[i]] >= nl) need_global.push_back(logical_bits[i]);
  }
  if (need_global.empty()) return BT_OK;
  int logical_at[64];
  for (int lb = 0; lb < nq; ++lb) logical_at[s->phys_of_bit[lb]] = lb;
  int new_phys[64];
  for (int lb = 0; lb < nq; ++lb) new_phys[lb] = s->phys_of_bit[lb];
  int pos = nl - 1;
  for (size_t i = 0; i < need_global.size(); ++i) {
    while (pos >= REMAP_KEEP && is_needed[logical_at[pos]]) --pos;
    if (pos < REMAP_KEEP) BT_FAIL(BT_ERR_UNSUPPORTED, "cannot make all requested qubits local");
    int victim = logical_at[pos];
    int gl = need_global[i];
    std::swap(new_phys[victim], new_phys[gl]);
    --pos;
  }
  return bt_sv_remap(s, new_phys);
}

// Gate-level preparation: non-diagonal targets must be local (physical bits in g refer to the current layout).
int bt_prepare_local(bt_sv* s, const GateDesc& g) {
  if (s->world == 1) return BT_OK;
  if (g.diag) return BT_OK;
  int logical_at[64];
  for (int lb = 0; lb < s->n_qubits; ++lb) logical_at[s->phys_of_bit[lb]] = lb;
  int lbs[4], n = 0;
  bool any_global = false;
  for (int t = 0; t < g.k; ++t) {
    lbs[n++] = logical_at[g.tb[t]];
    if (g.tb[t] >= s->n_local) any_global = true;
  }
  if (!any_global) return BT_OK;
  return bt_prepare_local_bits(s, n, lbs);
}
