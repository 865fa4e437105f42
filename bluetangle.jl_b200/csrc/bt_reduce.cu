// bt_reduce.cu -- read-only reductions over the amplitudes: reduced density matrices, norms, inner products,
// per-bit probabilities and Pauli-string expectation values.
//
// Replaces partial_trace (src/linalg.jl:167-192 one qubit, :198-230 adjacent pair, :83-140 general via
// state*state'), expect (src/func.jl:91-101) and correlation (:139-147).  One pass reads 16 B per amplitude;
// per-thread accumulators -> warp shuffles -> shared memory -> one partial per block, then a second tiny
// kernel adds the partials in a fixed order, so every result is bit-reproducible for a given shape.
#include "bt_internal.cuh"

static const int RB = 256;  // reduction block size

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double* __restrict__ out) {
  __shared__ double sm[(RB / 32) * NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double x = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) sm[warp * NV + i] = x;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NV; i += RB) {
    double x = 0.0;
#pragma unroll
    for (int w = 0; w < RB / 32; ++w) x += sm[w * NV + i];
    out[i] = x;
  }
  __syncthreads();
}

// final stage: res[t][v] = sum_b part[(t*nblk + b)*nv + v]; one 128-thread block per (t, v), fixed-order tree
__global__ void __launch_bounds__(128) k_sum_partials(const double* __restrict__ part, double* __restrict__ res, int nblk, int nv, int64_t n_batch) {
  __shared__ double sm[128];
  const int64_t i = blockIdx.x;
  const int64_t t = i / nv;
  const int v = (int)(i % nv);
  double x = 0.0;
  for (int b = threadIdx.x; b < nblk; b += 128) x += part[((size_t)t * nblk + b) * nv + v];
  sm[threadIdx.x] = x;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) res[i] = sm[0];
}

struct RdmPlan {
  int ni;
  int ins[4];
  uint64_t off[8];
};

__device__ __forceinline__ uint64_t rdm_expand(uint64_t g, const RdmPlan& p) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < p.ni) {
      int b = p.ins[i];
      g = ((g >> b) << (b + 1)) | (g & ((1ull << b) - 1));
    }
  return g;
}

// rho[a][b] = sum x_a conj(x_b).  Stored as D*D doubles: [a*D+a] = diag (real); for a<b: [a*D+b] = Re, [b*D+a] = Im of rho[a][b].
__device__ __forceinline__ void ld256(const double2* p, double2& a, double2& b) {  // see bt_gates.cu
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
}

// T0 >= 0: matrix index bit T0 sits on physical bit 0 -> entries j and j | (1 << T0) are one 256-bit load
template <int K, int T0>
__global__ void __launch_bounds__(RB) k_rdm(const double2* __restrict__ a, int n_local, const __grid_constant__ RdmPlan P,
                                             double* __restrict__ part, int nblk) {
  constexpr int D = 1 << K;
  const int64_t traj = blockIdx.y;
  const uint64_t ngroups = 1ull << (n_local - K);
  const double2* base = a + ((uint64_t)traj << n_local);
  double acc[D * D];
#pragma unroll
  for (int i = 0; i < D * D; ++i) acc[i] = 0.0;
#pragma unroll 4
  for (uint64_t g = (uint64_t)blockIdx.x * RB + threadIdx.x; g < ngroups; g += (uint64_t)nblk * RB) {
    uint64_t i0 = rdm_expand(g, P);
    double2 x[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      if (T0 >= 0) { if (!((j >> (T0 >= 0 ? T0 : 0)) & 1)) ld256(base + i0 + P.off[j], x[j], x[j | (1 << (T0 >= 0 ? T0 : 0))]); }
      else x[j] = base[i0 + P.off[j]];
    }
#pragma unroll
    for (int r = 0; r < D; ++r) {
      acc[r * D + r] += x[r].x * x[r].x + x[r].y * x[r].y;
#pragma unroll
      for (int c = r + 1; c < D; ++c) {
        // x_r * conj(x_c)
        acc[r * D + c] += x[r].x * x[c].x + x[r].y * x[c].y;
        acc[c * D + r] += x[r].y * x[c].x - x[r].x * x[c].y;
      }
    }
  }
  block_reduce_store<D * D>(acc, part + ((size_t)traj * nblk + blockIdx.x) * (D * D));
}

__global__ void __launch_bounds__(RB) k_norm2(const double2* __restrict__ a, int n_local, double* __restrict__ part, int nblk) {
  const int64_t traj = blockIdx.y;
  const uint64_t n = 1ull << n_local;
  const double2* base = a + ((uint64_t)traj << n_local);
  double acc[1] = {0.0};
#pragma unroll 8
  for (uint64_t i = (uint64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (uint64_t)nblk * RB) {
    double2 x = base[i];
    acc[0] += x.x * x.x + x.y * x.y;
  }
  block_reduce_store<1>(acc, part + ((size_t)traj * nblk + blockIdx.x));
}

__global__ void __launch_bounds__(RB) k_inner(const double2* __restrict__ a, const double2* __restrict__ b, int n_local,
                                               double* __restrict__ part, int nblk) {
  const int64_t traj = blockIdx.y;
  const uint64_t n = 1ull << n_local;
  const double2* pa = a + ((uint64_t)traj << n_local);
  const double2* pb = b + ((uint64_t)traj << n_local);
  double acc[2] = {0.0, 0.0};
#pragma unroll 4
  for (uint64_t i = (uint64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (uint64_t)nblk * RB) {
    double2 x = pa[i], y = pb[i];
    acc[0] += x.x * y.x + x.y * y.y;  // conj(x)*y
    acc[1] += x.x * y.y - x.y * y.x;
  }
  block_reduce_store<2>(acc, part + ((size_t)traj * nblk + blockIdx.x) * 2);
}

// Per-bit probabilities in one pass: out[0] = sum p, out[1+b] = sum of p over amplitudes with bit b set.
// Block j of a trajectory owns the contiguous range of 2^(cbits+lcpb) amplitudes starting at j << (cbits+lcpb):
// bits 0..7 are thread bits, 8..cbits-1 iteration bits, cbits..cbits+lcpb-1 chunk-in-block bits (per-thread
// accumulators), and the remaining high bits are constant per block and resolved from the block index in the final stage.
#define BP_MAXBITS 40
#define BP_MAXLC 10
__global__ void __launch_bounds__(RB) k_bitprobs(const double2* __restrict__ a, int n_local, int cbits, int lcpb, double* __restrict__ part, int nblk) {
  const int64_t traj = blockIdx.y;
  const double2* base = a + ((uint64_t)traj << n_local) + ((uint64_t)blockIdx.x << (cbits + lcpb));
  const uint32_t chunk_len = 1u << cbits;
  const uint32_t cpb = 1u << lcpb;
  const int iters = (int)((chunk_len + RB - 1) / RB);  // <= 16
  double tot = 0.0;
  double it_acc[4] = {0, 0, 0, 0};
  double hi[BP_MAXLC];
#pragma unroll
  for (int b = 0; b < BP_MAXLC; ++b) hi[b] = 0.0;
  for (uint32_t ch = 0; ch < cpb; ++ch) {
    const double2* cp = base + (uint64_t)ch * chunk_len;
    double p[16];
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      uint32_t k = (uint32_t)it * RB + threadIdx.x;
      double2 x = (it < iters && k < chunk_len) ? cp[k] : make_double2(0.0, 0.0);
      p[it] = x.x * x.x + x.y * x.y;
    }
    double ct = 0.0;
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      ct += p[it];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if ((it >> j) & 1) it_acc[j] += p[it];
    }
    tot += ct;
#pragma unroll
    for (int b = 0; b < BP_MAXLC; ++b)
      if (b < lcpb && ((ch >> b) & 1)) hi[b] += ct;
  }
  double* out = part + ((size_t)traj * nblk + blockIdx.x) * (BP_MAXBITS + 1);
  __shared__ double sm[RB / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = 1 + cbits + lcpb;
  for (int v = 0; v < nv; ++v) {
    double x;
    if (v == 0) x = tot;
    else {
      int b = v - 1;
      if (b < 8 && b < cbits) x = ((threadIdx.x >> b) & 1) ? tot : 0.0;
      else if (b < cbits) {
        int j = b - 8;
        x = (j == 0) ? it_acc[0] : (j == 1) ? it_acc[1] : (j == 2) ? it_acc[2] : it_acc[3];
      } else {
        int j = b - cbits;
        x = 0.0;
#pragma unroll
        for (int q = 0; q < BP_MAXLC; ++q)
          if (q == j) x = hi[q];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) sm[warp] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < RB / 32; ++w) s += sm[w];
      out[v] = s;
    }
    __syncthreads();
  }
}

// final stage for k_bitprobs: value v <= inblock bits: plain sum over blocks; higher bits: sum of block totals over the
// blocks whose index has that bit set.  One 128-thread block per (t, v), fixed-order tree.
__global__ void __launch_bounds__(128) k_bitprobs_final(const double* __restrict__ part, double* __restrict__ res, int nblk, int inblock_bits, int n_local) {
  __shared__ double sm[128];
  const int nvs = BP_MAXBITS + 1;
  const int64_t t = blockIdx.x / nvs;
  const int v = (int)(blockIdx.x % nvs);
  double x = 0.0;
  if (v <= n_local) {
    if (v <= inblock_bits) {
      for (int b = threadIdx.x; b < nblk; b += 128) x += part[((size_t)t * nblk + b) * nvs + v];
    } else {
      int hb = v - 1 - inblock_bits;  // bit of the block index
      for (int b = threadIdx.x; b < nblk; b += 128)
        if ((b >> hb) & 1) x += part[((size_t)t * nblk + b) * nvs];
    }
  }
  sm[threadIdx.x] = x;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) res[blockIdx.x] = sm[0];
}

// <P> for a Pauli string: xm = X|Y positions, zm = Z|Y positions, ny = number of Y.
// Pairs (i, i^xm) are visited once (i has the top bit of xm clear) so every amplitude is read exactly once.
__global__ void __launch_bounds__(RB) k_pauli(const double2* __restrict__ a, int n_local, uint64_t xm, uint64_t zm, int ny, int topx,
                                               double* __restrict__ part, int nblk) {
  const int64_t traj = blockIdx.y;
  const double2* base = a + ((uint64_t)traj << n_local);
  double acc[1] = {0.0};
  if (xm == 0) {
    const uint64_t n = 1ull << n_local;
#pragma unroll 8
    for (uint64_t i = (uint64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (uint64_t)nblk * RB) {
      double2 x = base[i];
      double p = x.x * x.x + x.y * x.y;
      acc[0] += (__popcll(i & zm) & 1) ? -p : p;
    }
  } else {
    // (-i)^ny as a complex constant
    double cr, ci;
    switch (ny & 3) { case 0: cr = 1; ci = 0; break; case 1: cr = 0; ci = -1; break; case 2: cr = -1; ci = 0; break; default: cr = 0; ci = 1; }
    const uint64_t n = 1ull << (n_local - 1);
#pragma unroll 4
    for (uint64_t g = (uint64_t)blockIdx.x * RB + threadIdx.x; g < n; g += (uint64_t)nblk * RB) {
      uint64_t i = ((g >> topx) << (topx + 1)) | (g & ((1ull << topx) - 1));
      uint64_t j = i ^ xm;
      double2 x = base[i], y = base[j];
      // term_i = c * s_i * conj(x) * y ; term_j = c * s_j * conj(y) * x
      double tr = x.x * y.x + x.y * y.y, ti = x.x * y.y - x.y * y.x;  // conj(x)*y
      double si = (__popcll(i & zm) & 1) ? -1.0 : 1.0;
      double sj = (__popcll(j & zm) & 1) ? -1.0 : 1.0;
      // real part of c*(si*(tr + i ti) + sj*(tr - i ti))
      double re = (si + sj) * tr, im = (si - sj) * ti;
      acc[0] += cr * re - ci * im;
    }
  }
  block_reduce_store<1>(acc, part + ((size_t)traj * nblk + blockIdx.x));
}

__global__ void k_abs2(const double2* __restrict__ a, double* __restrict__ p, uint64_t len) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < len; i += stride) {
    double2 x = a[i];
    p[i] = x.x * x.x + x.y * x.y;
  }
}

__global__ void k_scale_by_norm(double2* __restrict__ a, int n_local, uint64_t len, const double* __restrict__ norm2) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < len; i += stride) {
    double s = 1.0 / sqrt(norm2[i >> n_local]);
    double2 x = a[i];
    a[i] = make_double2(x.x * s, x.y * s);
  }
}

// ---- host side --------------------------------------------------------------------------------------------
static int pick_nblk(const bt_sv* s, uint64_t work_items_per_traj) {
  uint64_t want = (work_items_per_traj + RB - 1) / RB;
  uint64_t cap = std::max<uint64_t>(1, (148ull * 8) / (uint64_t)s->n_batch);
  return (int)std::max<uint64_t>(1, std::min<uint64_t>(want, cap));
}

static int check_batch_grid(const bt_sv* s) {
  if (s->n_batch > 65535) BT_FAIL(BT_ERR_UNSUPPORTED, "n_batch > 65535 not supported by the reduction kernels");
  return BT_OK;
}

static int finish(const bt_sv* s, int nblk, int nv, size_t res_off) {
  int64_t tot = s->n_batch * nv;
  if (res_off + (size_t)tot > s->res_cap) BT_FAIL(BT_ERR_ARG, "internal: result buffer overflow");
  k_sum_partials<<<(unsigned)tot, 128, 0, s->stream>>>(s->d_part, s->d_res + res_off, nblk, nv, s->n_batch);
  BT_CHECK_LAUNCH(s);
  return BT_OK;
}

int bt_reduce_rdm_at(const bt_sv* cs, int k, const int* tb, size_t res_off) {
  bt_sv* s = const_cast<bt_sv*>(cs);
  BT_TRY(check_batch_grid(s));
  if (k < 1 || k > 3) BT_FAIL(BT_ERR_ARG, "rdm arity %d unsupported", k);
  if (s->n_local < k) BT_FAIL(BT_ERR_ARG, "state too small for a %d-qubit RDM", k);
  RdmPlan P;
  int sorted[4];
  for (int i = 0; i < k; ++i) {
    if (tb[i] < 0 || tb[i] >= s->n_local) BT_FAIL(BT_ERR_UNSUPPORTED, "reduced density matrix over a global qubit needs a remap first");
    sorted[i] = tb[i];
  }
  std::sort(sorted, sorted + k);
  for (int i = 0; i + 1 < k; ++i)
    if (sorted[i] == sorted[i + 1]) BT_FAIL(BT_ERR_ARG, "repeated qubit");
  P.ni = k;
  for (int i = 0; i < 4; ++i) P.ins[i] = i < k ? sorted[i] : 0;
  for (int j = 0; j < (1 << k); ++j) {
    uint64_t o = 0;
    for (int t = 0; t < k; ++t)
      if ((j >> t) & 1) o |= 1ull << tb[t];
    P.off[j] = o;
  }
  int D = 1 << k, nv = D * D;
  int nblk = pick_nblk(s, 1ull << (s->n_local - k));
  BT_TRY(bt_ensure_partials(s, (size_t)s->n_batch * nblk * nv));
  dim3 grid(nblk, (unsigned)s->n_batch);
  int t0 = -1;  // matrix bit on index bit 0: its pairs are read as 256-bit loads
  for (int t = 0; t < k; ++t)
    if (tb[t] == 0) t0 = t;
  static const bool wide = []() { const char* v = getenv("BT_WIDE_LOADS"); return !(v && *v == '0'); }();
  if (!wide) t0 = -1;
#define RDM_LAUNCH(K, T0) k_rdm<K, T0><<<grid, RB, 0, s->stream>>>(s->amp, s->n_local, P, s->d_part, nblk)
  if (k == 1) { if (t0 == 0) RDM_LAUNCH(1, 0); else RDM_LAUNCH(1, -1); }
  else if (k == 2) { if (t0 == 0) RDM_LAUNCH(2, 0); else if (t0 == 1) RDM_LAUNCH(2, 1); else RDM_LAUNCH(2, -1); }
  else { if (t0 == 0) RDM_LAUNCH(3, 0); else if (t0 == 1) RDM_LAUNCH(3, 1); else if (t0 == 2) RDM_LAUNCH(3, 2); else RDM_LAUNCH(3, -1); }
#undef RDM_LAUNCH
  BT_CHECK_LAUNCH(s);
  return finish(s, nblk, nv, res_off);
}

int bt_reduce_rdm(const bt_sv* s, int k, const int* tb) { return bt_reduce_rdm_at(s, k, tb, 0); }

// joint distribution of K index bits: acc[j] = sum over the other bits of |a[.., bits = j, ..]|^2 (the diagonal of the K-bit RDM)
template <int K>
__global__ void __launch_bounds__(RB) k_joint_probs(const double2* __restrict__ a, int n_local, const __grid_constant__ RdmPlan P, const uint64_t* __restrict__ off16,
                                                     double* __restrict__ part, int nblk) {
  constexpr int D = 1 << K;
  const int64_t traj = blockIdx.y;
  const uint64_t ngroups = 1ull << (n_local - K);
  const double2* base = a + ((uint64_t)traj << n_local);
  uint64_t off[D];
#pragma unroll
  for (int j = 0; j < D; ++j) off[j] = K <= 3 ? P.off[j] : off16[j];
  double acc[D];
#pragma unroll
  for (int i = 0; i < D; ++i) acc[i] = 0.0;
#pragma unroll 2
  for (uint64_t g = (uint64_t)blockIdx.x * RB + threadIdx.x; g < ngroups; g += (uint64_t)nblk * RB) {
    const uint64_t i0 = rdm_expand(g, P);
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double2 x = base[i0 + off[j]];
      acc[j] = fma(x.x, x.x, fma(x.y, x.y, acc[j]));
    }
  }
  block_reduce_store<D>(acc, part + ((size_t)traj * nblk + blockIdx.x) * D);
}

// results: s->d_res[t * 2^k + j], j's bit i <-> physical bit tb[i]; k <= 4
int bt_reduce_joint_probs(const bt_sv* cs, int k, const int* tb) {
  bt_sv* s = const_cast<bt_sv*>(cs);
  BT_TRY(check_batch_grid(s));
  if (k < 1 || k > 4 || s->n_local < k) BT_FAIL(BT_ERR_ARG, "joint distribution of %d bits unsupported", k);
  RdmPlan P;
  int sorted[4];
  for (int i = 0; i < k; ++i) {
    if (tb[i] < 0 || tb[i] >= s->n_local) BT_FAIL(BT_ERR_UNSUPPORTED, "measured qubit on a global bit needs a remap first");
    sorted[i] = tb[i];
  }
  std::sort(sorted, sorted + k);
  for (int i = 0; i + 1 < k; ++i)
    if (sorted[i] == sorted[i + 1]) BT_FAIL(BT_ERR_ARG, "repeated qubit");
  P.ni = k;
  for (int i = 0; i < 4; ++i) P.ins[i] = i < k ? sorted[i] : 0;
  uint64_t off[16];
  for (int j = 0; j < (1 << k); ++j) {
    uint64_t o = 0;
    for (int t = 0; t < k; ++t)
      if ((j >> t) & 1) o |= 1ull << tb[t];
    off[j] = o;
    if (j < 8) P.off[j] = o;
  }
  const int D = 1 << k;
  const int nblk = pick_nblk(s, 1ull << (s->n_local - k));
  BT_TRY(bt_ensure_partials(s, (size_t)s->n_batch * nblk * D));
  if ((size_t)s->n_batch * D > s->res_cap) BT_FAIL(BT_ERR_ARG, "internal: result buffer overflow");
  uint64_t* d_off = nullptr;
  if (k == 4) {
    BT_TRY(bt_ensure_scratch(s, sizeof(off)));
    d_off = (uint64_t*)s->d_scratch;
    BT_CUDA(cudaMemcpyAsync(d_off, off, sizeof(off), cudaMemcpyHostToDevice, s->stream));
  }
  dim3 grid(nblk, (unsigned)s->n_batch);
  if (k == 1) k_joint_probs<1><<<grid, RB, 0, s->stream>>>(s->amp, s->n_local, P, d_off, s->d_part, nblk);
  else if (k == 2) k_joint_probs<2><<<grid, RB, 0, s->stream>>>(s->amp, s->n_local, P, d_off, s->d_part, nblk);
  else if (k == 3) k_joint_probs<3><<<grid, RB, 0, s->stream>>>(s->amp, s->n_local, P, d_off, s->d_part, nblk);
  else k_joint_probs<4><<<grid, RB, 0, s->stream>>>(s->amp, s->n_local, P, d_off, s->d_part, nblk);
  BT_CHECK_LAUNCH(s);
  return finish(s, nblk, D, 0);
}

int bt_reduce_norm2(const bt_sv* cs) {
  bt_sv* s = const_cast<bt_sv*>(cs);
  BT_TRY(check_batch_grid(s));
  int nblk = pick_nblk(s, 1ull << s->n_local);
  BT_TRY(bt_ensure_partials(s, (size_t)s->n_batch * nblk));
  dim3 grid(nblk, (unsigned)s->n_batch);
  k_norm2<<<grid, RB, 0, s->stream>>>(s->amp, s->n_local, s->d_part, nblk);
  BT_CHECK_LAUNCH(s);
  return finish(s, nblk, 1, 0);
}

// unpack the packed D*D doubles of one trajectory into a column-major complex matrix (rho[a][b] at a + b*D)
static void unpack_rdm(const double* v, int D, bt_c64* out) {
  for (int a = 0; a < D; ++a) {
    out[a + a * D].re = v[a * D + a];
    out[a + a * D].im = 0.0;
    for (int b = a + 1; b < D; ++b) {
      double re = v[a * D + b], im = v[b * D + a];
      out[a + b * D].re = re; out[a + b * D].im = im;     // rho[a][b]
      out[b + a * D].re = re; out[b + a * D].im = -im;    // rho[b][a] = conj
    }
  }
}

static int allreduce_host(const bt_sv* s, double* buf, int n) {
  if (s->world > 1) {
    // without a callback the caller gets this shard's partial sums (single-process hosts add them up themselves)
    if (s->allreduce) s->allreduce(s->allreduce_ctx, buf, n);
  }
  return BT_OK;
}

int bt_prepare_local_bits(bt_sv* s, int n, const int* logical_bits);  // bt_dist.cu
int bt_reduce_rdm_gram(bt_sv* s, int k, const int* tb, bt_c64* out_host);  // bt_linalg.cu

extern "C" int bt_sv_rdm1(const bt_sv* s, int qubit, bt_c64* out) {
  BT_TRY(bt_check_sv(s));
  if (!out) BT_FAIL(BT_ERR_ARG, "null output");
  if (qubit < 1 || qubit > s->n_qubits) BT_FAIL(BT_ERR_ARG, "qubit %d out of range", qubit);
  int lb = s->n_qubits - qubit;
  if (s->world > 1) BT_TRY(bt_prepare_local_bits(const_cast<bt_sv*>(s), 1, &lb));
  int tb[1] = {s->phys_of_bit[lb]};
  BT_TRY(bt_reduce_rdm(s, 1, tb));
  BT_TRY(bt_results_to_host(s, (size_t)s->n_batch * 4));
  BT_TRY(allreduce_host(s, s->h_res, (int)(s->n_batch * 4)));
  for (int64_t t = 0; t < s->n_batch; ++t) unpack_rdm(s->h_res + t * 4, 2, out + t * 4);
  return BT_OK;
}

extern "C" int bt_sv_rdm2(const bt_sv* s, int qa, int qb, bt_c64* out) {
  BT_TRY(bt_check_sv(s));
  if (!out) BT_FAIL(BT_ERR_ARG, "null output");
  if (qa < 1 || qb < 1 || qa > s->n_qubits || qb > s->n_qubits || qa == qb) BT_FAIL(BT_ERR_ARG, "invalid qubit pair (%d,%d)", qa, qb);
  int lo = std::min(qa, qb), hi = std::max(qa, qb);
  // index = 2*b_min + b_max (src/linalg.jl:206, :83-86): matrix bit 1 <-> smaller label, bit 0 <-> larger label
  int lbs[2] = {s->n_qubits - hi, s->n_qubits - lo};
  if (s->world > 1) BT_TRY(bt_prepare_local_bits(const_cast<bt_sv*>(s), 2, lbs));
  int tb[2] = {s->phys_of_bit[lbs[0]], s->phys_of_bit[lbs[1]]};
  BT_TRY(bt_reduce_rdm(s, 2, tb));
  BT_TRY(bt_results_to_host(s, (size_t)s->n_batch * 16));
  BT_TRY(allreduce_host(s, s->h_res, (int)(s->n_batch * 16)));
  for (int64_t t = 0; t < s->n_batch; ++t) unpack_rdm(s->h_res + t * 16, 4, out + t * 16);
  return BT_OK;
}

extern "C" int bt_sv_rdm3(const bt_sv* s, int first, bt_c64* out) {
  BT_TRY(bt_check_sv(s));
  if (!out) BT_FAIL(BT_ERR_ARG, "null output");
  if (first < 1 || first + 2 > s->n_qubits) BT_FAIL(BT_ERR_ARG, "N must be larger than all three qubits");
  int lbs[3] = {s->n_qubits - (first + 2), s->n_qubits - (first + 1), s->n_qubits - first};
  if (s->world > 1) BT_TRY(bt_prepare_local_bits(const_cast<bt_sv*>(s), 3, lbs));
  int tb[3] = {s->phys_of_bit[lbs[0]], s->phys_of_bit[lbs[1]], s->phys_of_bit[lbs[2]]};
  BT_TRY(bt_reduce_rdm(s, 3, tb));
  BT_TRY(bt_results_to_host(s, (size_t)s->n_batch * 64));
  BT_TRY(allreduce_host(s, s->h_res, (int)(s->n_batch * 64)));
  for (int64_t t = 0; t < s->n_batch; ++t) unpack_rdm(s->h_res + t * 64, 8, out + t * 64);
  return BT_OK;
}

// general form of partial_trace(state, keep_qubits) (src/linalg.jl:83-86) for up to three arbitrary qubits: the kept
// qubits are ordered by ascending label (first = most significant index bit of the reduced matrix), whatever order they are
// listed in -- exactly what setdiff(1:N, keep) + the tensor reshape of the reference produce.
extern "C" int bt_sv_rdm(const bt_sv* s, int k, const int* qubits, bt_c64* out) {
  BT_TRY(bt_check_sv(s));
  if (!qubits || !out) BT_FAIL(BT_ERR_ARG, "null argument");
  if (k < 1 || k > 12 || k > s->n_qubits) BT_FAIL(BT_ERR_UNSUPPORTED, "device partial_trace keeps 1..12 qubits (asked for %d)", k);
  if (k > 3) {  // tiled Gram kernel over the gathered 2^k x 2^(N-k) matrix (bt_linalg.cu)
    if (s->world > 1) BT_FAIL(BT_ERR_UNSUPPORTED, "partial_trace keeping more than 3 qubits of a sharded state");
    int qs[12], tbs[12];
    for (int i = 0; i < k; ++i) {
      qs[i] = qubits[i];
      if (qs[i] < 1 || qs[i] > s->n_qubits) BT_FAIL(BT_ERR_ARG, "qubit %d out of range", qs[i]);
    }
    std::sort(qs, qs + k);
    for (int i = 0; i + 1 < k; ++i) if (qs[i] == qs[i + 1]) BT_FAIL(BT_ERR_ARG, "repeated qubit %d", qs[i]);
    for (int t = 0; t < k; ++t) tbs[t] = s->phys_of_bit[s->n_qubits - qs[k - 1 - t]];
    return bt_reduce_rdm_gram(const_cast<bt_sv*>(s), k, tbs, out);
  }
  int q[3];
  for (int i = 0; i < k; ++i) {
    q[i] = qubits[i];
    if (q[i] < 1 || q[i] > s->n_qubits) BT_FAIL(BT_ERR_ARG, "qubit %d out of range", q[i]);
  }
  std::sort(q, q + k);
  for (int i = 0; i + 1 < k; ++i) if (q[i] == q[i + 1]) BT_FAIL(BT_ERR_ARG, "repeated qubit %d", q[i]);
  int lbs[3];
  for (int t = 0; t < k; ++t) lbs[t] = s->n_qubits - q[k - 1 - t];  // matrix bit 0 <-> largest label
  if (s->world > 1) BT_TRY(bt_prepare_local_bits(const_cast<bt_sv*>(s), k, lbs));
  int tb[3];
  for (int t = 0; t < k; ++t) tb[t] = s->phys_of_bit[lbs[t]];
  BT_TRY(bt_reduce_rdm(s, k, tb));
  int D = 1 << k;
  BT_TRY(bt_results_to_host(s, (size_t)s->n_batch * D * D));
  BT_TRY(allreduce_host(s, s->h_res, (int)(s->n_batch * D * D)));
  for (int64_t t = 0; t < s->n_batch; ++t) unpack_rdm(s->h_res + t * D * D, D, out + t * D * D);
  return BT_OK;
}

extern "C" int bt_sv_norm2(const bt_sv* s, double* out) {
  BT_TRY(bt_check_sv(s));
  if (!out) BT_FAIL(BT_ERR_ARG, "null output");
  BT_TRY(bt_reduce_norm2(s));
  BT_TRY(bt_results_to_host(s, (size_t)s->n_batch));
  BT_TRY(allreduce_host(s, s->h_res, (int)s->n_batch));
  for (int64_t t = 0; t < s->n_batch; ++t) out[t] = s->h_res[t];
  return BT_OK;
}

extern "C" int bt_sv_inner(const bt_sv* a, const bt_sv* b, bt_c64* out) {
  BT_TRY(bt_check_sv(a));
  if (!b || !out) BT_FAIL(BT_ERR_ARG, "null argument");
  if (a->len != b->len || a->n_qubits != b->n_qubits || a->device != b->device) BT_FAIL(BT_ERR_ARG, "bt_sv_inner: shape/device mismatch");
  if (a->world > 1 && memcmp(a->phys_of_bit, b->phys_of_bit, sizeof(a->phys_of_bit)) != 0) BT_FAIL(BT_ERR_UNSUPPORTED, "bt_sv_inner: shard layouts differ");
  bt_sv* s = const_cast<bt_sv*>(a);
  BT_TRY(check_batch_grid(s));
  BT_CUDA(cudaStreamSynchronize(b->stream));
  int nblk = pick_nblk(s, 1ull << s->n_local);
  BT_TRY(bt_ensure_partials(s, (size_t)s->n_batch * nblk * 2));
  dim3 grid(nblk, (unsigned)s->n_batch);
  k_inner<<<grid, RB, 0, s->stream>>>(a->amp, b->amp, s->n_local, s->d_part, nblk);
  BT_CHECK_LAUNCH(s);
  BT_TRY(finish(s, nblk, 2, 0));
  BT_TRY(bt_results_to_host(s, (size_t)s->n_batch * 2));
  BT_TRY(allreduce_host(s, s->h_res, (int)(s->n_batch * 2)));
  for (int64_t t = 0; t < s->n_batch; ++t) { out[t].re = s->h_res[2 * t]; out[t].im = s->h_res[2 * t + 1]; }
  return BT_OK;
}

extern "C" int bt_sv_normalize(bt_sv* s) {
  BT_TRY(bt_check_sv(s));
  BT_TRY(bt_reduce_norm2(s));
  if (s->world > 1) {
    BT_TRY(bt_results_to_host(s, (size_t)s->n_batch));
    BT_TRY(allreduce_host(s, s->h_res, (int)s->n_batch));
    BT_CUDA(cudaMemcpyAsync(s->d_res, s->h_res, s->n_batch * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  }
  unsigned grid = (unsigned)std::min<uint64_t>((s->len + 255) / 256, 148ull * 32);
  k_scale_by_norm<<<grid, 256, 0, s->stream>>>(s->amp, s->n_local, s->len, s->d_res);
  BT_CHECK_LAUNCH(s);
  return BT_OK;
}

extern "C" int bt_sv_probs(const bt_sv* cs, double* host) {
  BT_TRY(bt_check_sv(cs));
  if (!host) BT_FAIL(BT_ERR_ARG, "null output");
  bt_sv* s = const_cast<bt_sv*>(cs);
  if (s->world > 1 && !s->alt) BT_FAIL(BT_ERR_ARG, "internal: shard without second buffer");
  BT_TRY(bt_ensure_alt(s));
  double* p = reinterpret_cast<double*>(s->alt);
  unsigned grid = (unsigned)std::min<uint64_t>((s->len + 255) / 256, 148ull * 32);
  k_abs2<<<grid, 256, 0, s->stream>>>(s->amp, p, s->len);
  BT_CHECK_LAUNCH(s);
  BT_CUDA(cudaMemcpyAsync(host, p, s->len * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  BT_CUDA(cudaStreamSynchronize(s->stream));
  return BT_OK;
}

// ---- expectation values -------------------------------------------------------------------------------------
int bt_bitprobs(const bt_sv* cs) {
  // leaves n_batch x (BP_MAXBITS+1) doubles in d_res: [0] = total, [1+b] = P(physical bit b = 1) (unnormalised)
  bt_sv* s = const_cast<bt_sv*>(cs);
  BT_TRY(check_batch_grid(s));
  int cbits = std::min(12, s->n_local);
  int nchunk_bits = s->n_local - cbits;
  // blocks per trajectory: a power of two, about 8 CTAs per SM over the whole batch
  int lblk = 0;
  while (lblk < nchunk_bits && ((uint64_t)2 << lblk) * (uint64_t)s->n_batch <= 148ull * 8) ++lblk;
  int lcpb = nchunk_bits - lblk;
  while (lcpb > BP_MAXLC) { --lcpb; ++lblk; }
  int nblk = 1 << lblk;
  int nv = BP_MAXBITS + 1;
  BT_TRY(bt_ensure_partials(s, (size_t)s->n_batch * nblk * nv));
  dim3 grid(nblk, (unsigned)s->n_batch);
  k_bitprobs<<<grid, RB, 0, s->stream>>>(s->amp, s->n_local, cbits, lcpb, s->d_part, nblk);
  BT_CHECK_LAUNCH(s);
  if ((size_t)s->n_batch * nv > s->res_cap) BT_FAIL(BT_ERR_UNSUPPORTED, "too many trajectories for expect_1q_all");
  k_bitprobs_final<<<(unsigned)(s->n_batch * nv), 128, 0, s->stream>>>(s->d_part, s->d_res, nblk, cbits + lcpb, s->n_local);
  BT_CHECK_LAUNCH(s);
  return BT_OK;
}

extern "C" int bt_sv_expect_1q_all(const bt_sv* cs, const bt_c64 m[4], double* out) {
  BT_TRY(bt_check_sv(cs));
  if (!m || !out) BT_FAIL(BT_ERR_ARG, "null argument");
  bt_sv* s = const_cast<bt_sv*>(cs);
  int N = s->n_qubits;
  bool diag = (m[1].re == 0 && m[1].im == 0 && m[2].re == 0 && m[2].im == 0);
  if (diag) {
    BT_TRY(bt_bitprobs(s));
    int nv = BP_MAXBITS + 1;
    BT_TRY(bt_results_to_host(s, (size_t)s->n_batch * nv));
    BT_TRY(allreduce_host(s, s->h_res, (int)(s->n_batch * nv)));
    for (int64_t t = 0; t < s->n_batch; ++t) {
      const double* v = s->h_res + t * nv;
      for (int q = 1; q <= N; ++q) {
        int pb = s->phys_of_bit[N - q];
        double p1, p0;
        if (pb < s->n_local) { p1 = v[1 + pb]; p0 = v[0] - p1; }
        else {
          // global bit: after the all-reduce v[0] is the global total; the per-rank totals are gone, so
          // global-bit probabilities are accumulated by the callback-free path below
          p1 = 0; p0 = 0;
        }
        out[t * N + (q - 1)] = m[0].re * p0 + m[3].re * p1;
      }
    }
    if (s->world > 1) {
      // global bits: P(bit=1) = sum over ranks with that rank bit set of the local total -> second all-reduce
      BT_TRY(bt_reduce_norm2(s));
      BT_TRY(bt_results_to_host(s, (size_t)s->n_batch));
      std::vector<double> buf((size_t)s->n_batch * (s->g + 1), 0.0);
      for (int64_t t = 0; t < s->n_batch; ++t) {
        buf[t * (s->g + 1)] = s->h_res[t];
        for (int j = 0; j < s->g; ++j)
          if ((s->rank >> j) & 1) buf[t * (s->g + 1) + 1 + j] = s->h_res[t];
      }
      BT_TRY(allreduce_host(s, buf.data(), (int)buf.size()));
      for (int64_t t = 0; t < s->n_batch; ++t)
        for (int q = 1; q <= N; ++q) {
          int pb = s->phys_of_bit[N - q];
          if (pb >= s->n_local) {
            double p1 = buf[t * (s->g + 1) + 1 + (pb - s->n_local)];
            double p0 = buf[t * (s->g + 1)] - p1;
            out[t * N + (q - 1)] = m[0].re * p0 + m[3].re * p1;
          }
        }
    }
    return BT_OK;
  }
  // general 2x2: N reduced density matrices, one host synchronisation at the end
  if ((size_t)s->n_batch * 4 * N > s->res_cap) BT_FAIL(BT_ERR_UNSUPPORTED, "too many trajectories for expect_1q_all");
  int bt_reduce_rdm_at(const bt_sv*, int, const int*, size_t);
  for (int q = 1; q <= N; ++q) {
    int lb = N - q;
    if (s->world > 1) BT_TRY(bt_prepare_local_bits(s, 1, &lb));
    int tb[1] = {s->phys_of_bit[lb]};
    BT_TRY(bt_reduce_rdm_at(s, 1, tb, (size_t)(q - 1) * s->n_batch * 4));
  }
  BT_TRY(bt_results_to_host(s, (size_t)s->n_batch * 4 * N));
  BT_TRY(allreduce_host(s, s->h_res, (int)(s->n_batch * 4 * N)));
  for (int q = 1; q <= N; ++q)
    for (int64_t t = 0; t < s->n_batch; ++t) {
      const double* v = s->h_res + ((size_t)(q - 1) * s->n_batch + t) * 4;
      // <O> = sum_ab O[a,b] rho[b][a];  packed: v[0]=rho00, v[3]=rho11, v[1]=Re rho01, v[2]=Im rho01
      // O column-major: O[a,b] = m[a + 2b]
      double r01 = v[1], i01 = v[2];
      double e = m[0].re * v[0] + m[3].re * v[3];
      // O[0,1]*rho[1][0] + O[1,0]*rho[0][1];  rho10 = conj(rho01)
      e += (m[2].re * r01 + m[2].im * i01);   // Re(O01 * conj(rho01))
      e += (m[1].re * r01 - m[1].im * i01);   // Re(O10 * rho01)
      out[t * N + (q - 1)] = e;
    }
  return BT_OK;
}

extern "C" int bt_sv_expect_pauli(const bt_sv* cs, const char* paulis, double* out) {
  BT_TRY(bt_check_sv(cs));
  if (!paulis || !out) BT_FAIL(BT_ERR_ARG, "null argument");
  bt_sv* s = const_cast<bt_sv*>(cs);
  BT_TRY(check_batch_grid(s));
  int N = s->n_qubits;
  if ((int)strlen(paulis) != N) BT_FAIL(BT_ERR_ARG, "Pauli string must have %d characters", N);
  // X/Y on a global qubit needs the partner shard: bring those qubits local first
  if (s->world > 1) {
    int lbs[64], n = 0;
    for (int q = 1; q <= N; ++q) {
      char c = paulis[q - 1];
      if (c == 'X' || c == 'x' || c == 'Y' || c == 'y') lbs[n++] = N - q;
    }
    if (n > s->n_local) BT_FAIL(BT_ERR_UNSUPPORTED, "too many X/Y factors for a sharded state");
    if (n) BT_TRY(bt_prepare_local_bits(s, n, lbs));
  }
  uint64_t xm = 0, zm = 0, zglobal = 0;
  int ny = 0;
  for (int q = 1; q <= N; ++q) {
    char c = paulis[q - 1];
    int pb = s->phys_of_bit[N - q];
    uint64_t bit = 1ull << pb;
    switch (c) {
      case 'I': case 'i': break;
      case 'X': case 'x': xm |= bit; break;
      case 'Y': case 'y': xm |= bit; zm |= bit; ny++; break;
      case 'Z': case 'z': if (pb < s->n_local) zm |= bit; else zglobal |= 1ull << (pb - s->n_local); break;
      default: BT_FAIL(BT_ERR_ARG, "Pauli string may only contain I, X, Y, Z");
    }
  }
  int topx = 0;
  for (int b = 0; b < 64; ++b) if ((xm >> b) & 1) topx = b;
  uint64_t items = xm ? (1ull << (s->n_local - 1)) : (1ull << s->n_local);
  int nblk = pick_nblk(s, items);
  BT_TRY(bt_ensure_partials(s, (size_t)s->n_batch * nblk));
  dim3 grid(nblk, (unsigned)s->n_batch);
  k_pauli<<<grid, RB, 0, s->stream>>>(s->amp, s->n_local, xm, zm, ny, topx, s->d_part, nblk);
  BT_CHECK_LAUNCH(s);
  BT_TRY(finish(s, nblk, 1, 0));
  BT_TRY(bt_results_to_host(s, (size_t)s->n_batch));
  if (s->world > 1 && (__builtin_popcountll((uint64_t)s->rank & zglobal) & 1))
    for (int64_t t = 0; t < s->n_batch; ++t) s->h_res[t] = -s->h_res[t];
  BT_TRY(allreduce_host(s, s->h_res, (int)s->n_batch));
  for (int64_t t = 0; t < s->n_batch; ++t) out[t] = s->h_res[t];
  return BT_OK;
}

// ---- Pauli-sum expectation values (Hamiltonians): sum_k c_k <P_k> ---------------------------------------------------
// The reference builds H = sum_k c_k expand_multi_op(P_k) as one 2^N x 2^N sparse matrix (hamiltonian, src/vqa.jl:36-67)
// and takes real(state' * H * state) (src/func.jl:91; the VQE loss, src/vqa.jl:282-283).  Here the terms are never
// expanded.  Terms that are diagonal in a common product basis form a group (greedy qubit-wise grouping: a TFIM chain is
// 2 groups, a Heisenberg chain 3); a group costs ONE read of the state: rotate a scratch copy into the group's basis
// (H for X, H*S' for Y: one fused pass per <= 12 qubits, nothing for an all-Z group), then k_zsum accumulates
//   sum_i |a_i|^2 f(i),  f(i) = sum_k c_k (-1)^popc(i & z_k)
// for all terms of the group at once.  f is evaluated with a Walsh-Hadamard step: a thread holds the 16 probabilities
// of index bits 8..11, transforms them in registers (4 additions per amplitude), and a term then costs one sign and one
// addition per 16 amplitudes (terms pre-sorted by their pattern on bits 8..11), so a chain Hamiltonian of ~2N terms stays
// HBM-bound (16 B per amplitude per group) instead of one pass per term.
#define ZS_MAXT 160
struct ZsumParams {
  int32_t start[17];      // terms sorted by m = (z >> 8) & 15; start[m]..start[m+1]
  int32_t n;
  uint64_t z[ZS_MAXT];    // physical-bit parity mask
  double c[ZS_MAXT];
};

__global__ void __launch_bounds__(RB) k_zsum(const double2* __restrict__ a, int n_local, const __grid_constant__ ZsumParams P, double* __restrict__ part, int nblk) {
  const int64_t traj = blockIdx.y;
  const double2* base = a + ((uint64_t)traj << n_local);
  const uint64_t nsg = 1ull << (n_local - 12);  // supergroups of 2^12 amplitudes: bits 0..7 thread, 8..11 registers
  double acc[1] = {0.0};
  for (uint64_t sg = blockIdx.x; sg < nsg; sg += (uint64_t)nblk) {
    const uint64_t i0 = (sg << 12) | threadIdx.x;
    double w[16];
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      const double2 x = base[i0 | ((uint64_t)it << 8)];
      w[it] = x.x * x.x + x.y * x.y;
    }
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (!((j >> b) & 1)) {
          const double u = w[j], v = w[j | (1 << b)];
          w[j] = u + v;
          w[j | (1 << b)] = u - v;
        }
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      double g = 0.0;
      for (int k = P.start[m]; k < P.start[m + 1]; ++k) g += (__popcll(i0 & P.z[k]) & 1) ? -P.c[k] : P.c[k];
      acc[0] = fma(w[m], g, acc[0]);
    }
  }
  block_reduce_store<1>(acc, part + ((size_t)traj * nblk + blockIdx.x));
}

// states of fewer than 2^12 amplitudes: plain loop over the terms
__global__ void __launch_bounds__(RB) k_zsum_small(const double2* __restrict__ a, int n_local, const __grid_constant__ ZsumParams P, double* __restrict__ part, int nblk) {
  const int64_t traj = blockIdx.y;
  const double2* base = a + ((uint64_t)traj << n_local);
  const uint64_t n = 1ull << n_local;
  double acc[1] = {0.0};
  for (uint64_t i = (uint64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (uint64_t)nblk * RB) {
    const double2 x = base[i];
    double f = 0.0;
    for (int k = 0; k < P.n; ++k) f += (__popcll(i & P.z[k]) & 1) ? -P.c[k] : P.c[k];
    acc[0] = fma(x.x * x.x + x.y * x.y, f, acc[0]);
  }
  block_reduce_store<1>(acc, part + ((size_t)traj * nblk + blockIdx.x));
}

// sum over the terms (z, c) of c * sum_i |amp_i|^2 (-1)^popc(i & z) on the handle's CURRENT amp buffer, added to acc[n_batch]
static int zsum_accumulate(bt_sv* s, const std::vector<uint64_t>& z, const std::vector<double>& c, double* acc) {
  for (size_t k0 = 0; k0 < z.size(); k0 += ZS_MAXT) {
    const size_t k1 = std::min(z.size(), k0 + (size_t)ZS_MAXT);
    ZsumParams P;
    memset(&P, 0, sizeof(P));
    P.n = (int)(k1 - k0);
    int fill = 0;
    for (int m = 0; m < 16; ++m) {  // bucket by the pattern on index bits 8..11 (small states: one bucket, see k_zsum_small)
      P.start[m] = fill;
      for (size_t k = k0; k < k1; ++k) {
        const int mk = s->n_local >= 12 ? (int)((z[k] >> 8) & 15) : 0;
        if (mk != m) continue;
        P.z[fill] = z[k];
        P.c[fill] = c[k];
        fill++;
      }
    }
    P.start[16] = fill;
    if (fill != P.n) BT_FAIL(BT_ERR_ARG, "internal: term bucketing lost a term");
    int nblk;
    if (s->n_local >= 12) {
      nblk = pick_nblk(s, (uint64_t)RB << (s->n_local - 12));
      BT_TRY(bt_ensure_partials(s, (size_t)s->n_batch * nblk));
      dim3 grid(nblk, (unsigned)s->n_batch);
      k_zsum<<<grid, RB, 0, s->stream>>>(s->amp, s->n_local, P, s->d_part, nblk);
    } else {
      nblk = pick_nblk(s, 1ull << s->n_local);
      BT_TRY(bt_ensure_partials(s, (size_t)s->n_batch * nblk));
      dim3 grid(nblk, (unsigned)s->n_batch);
      k_zsum_small<<<grid, RB, 0, s->stream>>>(s->amp, s->n_local, P, s->d_part, nblk);
    }
    BT_CHECK_LAUNCH(s);
    BT_TRY(finish(s, nblk, 1, 0));
    BT_TRY(bt_results_to_host(s, (size_t)s->n_batch));
    for (int64_t t = 0; t < s->n_batch; ++t) acc[t] += s->h_res[t];
  }
  return BT_OK;
}

extern "C" int bt_sv_expect_pauli_sum(const bt_sv* cs, int n_terms, const char* paulis, const double* coefs, double* out) {
  BT_TRY(bt_check_sv(cs));
  if (n_terms < 0 || (n_terms > 0 && (!paulis || !coefs)) || !out) BT_FAIL(BT_ERR_ARG, "null argument");
  bt_sv* s = const_cast<bt_sv*>(cs);
  BT_TRY(check_batch_grid(s));
  const int N = s->n_qubits;
  for (int k = 0; k < n_terms; ++k)
    for (int q = 0; q < N; ++q) {
      const char ch = paulis[(size_t)k * N + q];
      if (!(ch == 'I' || ch == 'X' || ch == 'Y' || ch == 'Z')) BT_FAIL(BT_ERR_ARG, "Pauli strings may only contain I, X, Y, Z (term %d)", k);
    }
  for (int64_t t = 0; t < s->n_batch; ++t) out[t] = 0.0;
  if (s->world > 1) {
    // sharded state: term by term (the scratch-copy rotation below would have to remap both buffers)
    std::vector<double> one((size_t)s->n_batch);
    std::string str((size_t)N, 'I');
    for (int k = 0; k < n_terms; ++k) {
      str.assign(paulis + (size_t)k * N, (size_t)N);
      BT_TRY(bt_sv_expect_pauli(s, str.c_str(), one.data()));
      for (int64_t t = 0; t < s->n_batch; ++t) out[t] += coefs[k] * one[t];
    }
    return BT_OK;
  }
  // greedy qubit-wise grouping: a term joins the first group whose basis assignment it does not contradict
  struct Group { std::string basis; std::vector<int> terms; };
  std::vector<Group> groups;
  for (int k = 0; k < n_terms; ++k) {
    const char* tk = paulis + (size_t)k * N;
    Group* home = nullptr;
    for (Group& g : groups) {
      bool ok = true;
      for (int q = 0; q < N && ok; ++q)
        if (tk[q] != 'I' && g.basis[q] != 'I' && g.basis[q] != tk[q]) ok = false;
      if (ok) { home = &g; break; }
    }
    if (!home) { groups.push_back(Group{std::string((size_t)N, 'I'), {}}); home = &groups.back(); }
    for (int q = 0; q < N; ++q) if (tk[q] != 'I') home->basis[q] = tk[q];
    home->terms.push_back(k);
  }
  const double r = 0.70710678118654752440;
  for (const Group& g : groups) {
    std::vector<uint64_t> z;
    std::vector<double> c;
    for (int k : g.terms) {
      uint64_t m = 0;
      for (int q = 0; q < N; ++q)
        if (paulis[(size_t)k * N + q] != 'I') m |= 1ull << s->phys_of_bit[N - 1 - q];
      z.push_back(m);
      c.push_back(coefs[k]);
    }
    std::vector<bt_gate> rot;
    for (int q = 0; q < N; ++q) {
      if (g.basis[q] != 'X' && g.basis[q] != 'Y') continue;
      bt_gate u;
      memset(&u, 0, sizeof(u));
      u.nq = 1; u.qubit = q + 1; u.target = -1; u.control = -2;
      // column-major 2x2: X -> H;  Y -> H * S' = [[1, -i], [1, i]] / sqrt(2)   (U' Z U = P, so <psi|P|psi> = <U psi|Z|U psi>)
      if (g.basis[q] == 'X') { u.m[0].re = r; u.m[1].re = r; u.m[2].re = r; u.m[3].re = -r; }
      else { u.m[0].re = r; u.m[1].re = r; u.m[2].im = -r; u.m[3].im = r; }
      rot.push_back(u);
    }
    if (rot.empty()) {
      BT_TRY(zsum_accumulate(s, z, c, out));
      continue;
    }
    BT_TRY(bt_ensure_alt(s));
    BT_CUDA(cudaMemcpyAsync(s->alt, s->amp, s->len * sizeof(double2), cudaMemcpyDeviceToDevice, s->stream));
    std::swap(s->amp, s->alt);  // the rotations and the reduction act on the copy
    int rc = bt_sv_apply_circuit(s, rot.data(), rot.size(), 1);
    if (rc == BT_OK) rc = zsum_accumulate(s, z, c, out);
    std::swap(s->amp, s->alt);
    if (rc != BT_OK) return rc;
  }
  return BT_OK;
}

int bt_build_gate(const bt_sv* s, int nq, int qubit, int target, int control, const bt_c64* m, GateDesc* out);

extern "C" int bt_sv_expect_product(const bt_sv* cs, int n_ops, const int* qubits, const bt_c64* mats, double* out) {
  // Re <psi| (x)_j O_j |psi>  (src/func.jl:139-142 with expand_multi_op src/ops.jl:928-944): phi = (x)O psi on
  // the scratch buffer with the ordinary gate kernels, then one fused inner-product pass.
  BT_TRY(bt_check_sv(cs));
  if (n_ops < 0 || (n_ops > 0 && (!qubits || !mats)) || !out) BT_FAIL(BT_ERR_ARG, "null argument");
  bt_sv* s = const_cast<bt_sv*>(cs);
  BT_TRY(check_batch_grid(s));
  for (int i = 0; i < n_ops; ++i) {
    if (qubits[i] < 1 || qubits[i] > s->n_qubits) BT_FAIL(BT_ERR_ARG, "qubit %d out of range", qubits[i]);
    for (int j = 0; j < i; ++j)
      if (qubits[i] == qubits[j]) BT_FAIL(BT_ERR_ARG, "repeated qubit %d in operator list", qubits[i]);
  }
  if (s->world > 1) {
    int lbs[64], n = 0;
    for (int i = 0; i < n_ops; ++i) {
      const bt_c64* m = mats + 4 * i;
      bool diag = (m[1].re == 0 && m[1].im == 0 && m[2].re == 0 && m[2].im == 0);
      if (!diag) lbs[n++] = s->n_qubits - qubits[i];
    }
    if (n) BT_TRY(bt_prepare_local_bits(s, n, lbs));
  }
  BT_TRY(bt_ensure_alt(s));
  BT_CUDA(cudaMemcpyAsync(s->alt, s->amp, s->len * sizeof(double2), cudaMemcpyDeviceToDevice, s->stream));
  std::swap(s->amp, s->alt);  // gates now act on the copy
  int rc = BT_OK;
  for (int i = 0; i < n_ops && rc == BT_OK; ++i) {
    GateDesc g;
    rc = bt_build_gate(s, 1, qubits[i], -1, -2, mats + 4 * i, &g);
    if (rc == BT_OK) rc = bt_launch_gate(s, g);
  }
  std::swap(s->amp, s->alt);  // amp = psi, alt = phi
  if (rc != BT_OK) return rc;
  int nblk = pick_nblk(s, 1ull << s->n_local);
  BT_TRY(bt_ensure_partials(s, (size_t)s->n_batch * nblk * 2));
  dim3 grid(nblk, (unsigned)s->n_batch);
  k_inner<<<grid, RB, 0, s->stream>>>(s->amp, s->alt, s->n_local, s->d_part, nblk);
  BT_CHECK_LAUNCH(s);
  BT_TRY(finish(s, nblk, 2, 0));
  BT_TRY(bt_results_to_host(s, (size_t)s->n_batch * 2));
  BT_TRY(allreduce_host(s, s->h_res, (int)(s->n_batch * 2)));
  for (int64_t t = 0; t < s->n_batch; ++t) out[t] = s->h_res[2 * t];
  return BT_OK;
}

extern "C" int bt_sv_expect_matrix2q(const bt_sv* s, int qubit, int target, const bt_c64 m[16], double* out) {
  // Re <psi| O_{qubit,target} |psi> with O indexed 2*b_qubit + b_target, from the 4x4 RDM
  BT_TRY(bt_check_sv(s));
  if (!m || !out) BT_FAIL(BT_ERR_ARG, "null argument");
  if (qubit < 1 || target < 1 || qubit > s->n_qubits || target > s->n_qubits || qubit == target) BT_FAIL(BT_ERR_ARG, "invalid qubit pair");
  int lbs[2] = {s->n_qubits - target, s->n_qubits - qubit};
  if (s->world > 1) BT_TRY(bt_prepare_local_bits(const_cast<bt_sv*>(s), 2, lbs));
  int tb[2] = {s->phys_of_bit[lbs[0]], s->phys_of_bit[lbs[1]]};
  BT_TRY(bt_reduce_rdm(s, 2, tb));
  BT_TRY(bt_results_to_host(s, (size_t)s->n_batch * 16));
  BT_TRY(allreduce_host(s, s->h_res, (int)(s->n_batch * 16)));
  for (int64_t t = 0; t < s->n_batch; ++t) {
    bt_c64 rho[16];
    unpack_rdm(s->h_res + t * 16, 4, rho);
    double e = 0.0;
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) {
        // O[a,b] * rho[b][a]
        const bt_c64& o = m[a + 4 * b];
        const bt_c64& r = rho[b + 4 * a];
        e += o.re * r.re - o.im * r.im;
      }
    out[t] = e;
  }
  return BT_OK;
}
