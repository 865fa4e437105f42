// bt_gates.cu -- single-gate kernels: dense k-target pair/quad/oct updates, diagonal-phase kernels and
// controlled variants, all on ComplexF64 (double2) amplitudes with 16-byte loads/stores.
//
// Replaces `op.expand(N)*state` (src/hilbert.jl:505) where expand builds a 2^N x 2^N sparse matrix through
// hilbert() (src/hilbert.jl:18-70, :143-159), CCZX (:73-103) or hilbert3 (:106-128).  Here a gate is applied
// in place: one thread owns one group of 2^k amplitudes (all combinations of the k target bits, with every
// control bit fixed to 1), consecutive threads own consecutive groups so that each of the 2^k loads of a warp
// is a contiguous 512-byte run whenever the inserted bits are >= 5.
//
// HBM roofline: one pass = 32 B per touched amplitude (16 B read + 16 B write); a control halves the touched
// set, a diagonal entry equal to 1 removes its half as well (canonicalisation below).
#include "bt_internal.cuh"

int bt_is_strict();

// ---- canonicalisation (host) --------------------------------------------------------------------------
void bt_canonicalize(int k, const int* tb, const cplx* m, int nc, const int* cb, GateDesc* out) {
  int tbits[4];
  int ctrl[8];
  int nctrl = 0;
  for (int i = 0; i < nc; ++i) ctrl[nctrl++] = cb[i];
  for (int i = 0; i < k; ++i) tbits[i] = tb[i];
  std::vector<cplx> cur(m, m + ((size_t)1 << (2 * k)));
  int kk = k;
  bool changed = true;
  while (changed && kk > 0) {
    changed = false;
    for (int t = 0; t < kk; ++t) {
      int D = 1 << kk;
      bool is_ctrl = true;
      for (int r = 0; r < D && is_ctrl; ++r)
        for (int c = 0; c < D; ++c) {
          if (((r >> t) & 1) && ((c >> t) & 1)) continue;
          cplx want = (r == c) ? cplx(1, 0) : cplx(0, 0);
          if (cur[(size_t)r * D + c] != want) { is_ctrl = false; break; }
        }
      if (!is_ctrl) continue;
      if (nctrl >= 4) continue;  // kernels take at most 4 controls
      // extract sub-matrix with bit t == 1 in both indices
      int D2 = D >> 1;
      std::vector<cplx> sub((size_t)D2 * D2);
      for (int r2 = 0; r2 < D2; ++r2)
        for (int c2 = 0; c2 < D2; ++c2) {
          int lo_mask = (1 << t) - 1;
          int r = ((r2 >> t) << (t + 1)) | (1 << t) | (r2 & lo_mask);
          int c = ((c2 >> t) << (t + 1)) | (1 << t) | (c2 & lo_mask);
          sub[(size_t)r2 * D2 + c2] = cur[(size_t)r * D + c];
        }
      ctrl[nctrl++] = tbits[t];
      for (int i = t; i < kk - 1; ++i) tbits[i] = tbits[i + 1];
      kk--;
      cur.swap(sub);
      changed = true;
      break;
    }
  }
  out->k = kk;
  out->nc = nctrl;
  for (int i = 0; i < kk; ++i) out->tb[i] = tbits[i];
  for (int i = 0; i < nctrl; ++i) out->cb[i] = ctrl[i];
  int D = 1 << kk;
  bool diag = true;
  for (int r = 0; r < D && diag; ++r)
    for (int c = 0; c < D; ++c)
      if (r != c && cur[(size_t)r * D + c] != cplx(0, 0)) { diag = false; break; }
  out->diag = diag;
  if (diag) {
    for (int r = 0; r < D; ++r) out->m[r] = cur[(size_t)r * D + r];
  } else {
    for (int i = 0; i < D * D; ++i) out->m[i] = cur[i];
  }
}

// ---- device side ----------------------------------------------------------------------------------------
struct IdxPlan {
  int ni;            // number of inserted bit positions (targets + controls)
  int ins[8];        // ascending
  uint64_t cmask;    // control bits forced to 1
};

__device__ __forceinline__ uint64_t bt_expand(uint64_t g, const IdxPlan& p) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < p.ni) {
      int b = p.ins[i];
      g = ((g >> b) << (b + 1)) | (g & ((1ull << b) - 1));
    }
  }
  return g | p.cmask;
}

template <int K>
struct DenseParams {
  IdxPlan plan;
  uint64_t off[1 << K];
  double2 m[(1 << K) * (1 << K)];  // row-major
};

template <int K>
struct DiagParams {
  IdxPlan plan;
  uint64_t off[1 << K];
  double2 d[1 << K];
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ void cfma(double2& acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): two neighbouring amplitudes in one request.  Used when index bit 0 is a
// target: the two amplitudes of a pair are adjacent, and a warp-wide 16-byte load at a 32- or 64-byte stride would otherwise
// fetch every sector twice through L1.
__device__ __forceinline__ void ld256(const double2* p, double2& a, double2& b) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
}
__device__ __forceinline__ void st256(double2* p, double2 a, double2 b) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}

// gshift = log2(groups per trajectory): G >> gshift is the trajectory a group belongs to.
// T0 >= 0: matrix index bit T0 sits on physical bit 0 -> entries j and j | (1 << T0) are adjacent and move as one 256-bit access.
template <int K, int U, bool DEVMAT, bool COND, int T0>
__global__ void __launch_bounds__(256) k_dense(double2* __restrict__ a, uint64_t ngroups, int gshift,
                                                const __grid_constant__ DenseParams<K> P,
                                                const double2* __restrict__ dmats,
                                                const int32_t* __restrict__ cond, int want) {
  constexpr int D = 1 << K;
  const uint64_t g0 = ((uint64_t)blockIdx.x * U) * blockDim.x + threadIdx.x;
  double2 x[U][D];
  uint64_t base[U];
  bool ok[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    uint64_t G = g0 + (uint64_t)u * blockDim.x;
    ok[u] = G < ngroups;
    if (COND) { if (ok[u]) ok[u] = (cond[G >> gshift] == want); }
    base[u] = bt_expand(G, P.plan);
    if (ok[u]) {
#pragma unroll
      for (int j = 0; j < D; ++j) {
        if (T0 >= 0) { if (!((j >> (T0 >= 0 ? T0 : 0)) & 1)) ld256(a + base[u] + P.off[j], x[u][j], x[u][j | (1 << (T0 >= 0 ? T0 : 0))]); }
        else x[u][j] = a[base[u] + P.off[j]];
      }
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (!ok[u]) continue;
    const double2* mm = nullptr;
    if (DEVMAT) mm = dmats + (((g0 + (uint64_t)u * blockDim.x) >> gshift) << 6);
    double2 y[D];
#pragma unroll
    for (int r = 0; r < D; ++r) {
      double2 acc = make_double2(0.0, 0.0);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        double2 mv = DEVMAT ? __ldg(&mm[r * D + c]) : P.m[r * D + c];
        cfma(acc, mv, x[u][c]);
      }
      // plain form: every row is stored as soon as it is complete, so the stores drain while the FP64 pipe works on the next rows
      // (16 x 16 superoperators sit at the FP64 ridge: holding all rows back for one burst of stores cost them 19 %, 1.24 -> 1.47 ms)
      // (the masked per-trajectory-matrix form keeps the buffered stores it was validated with on the device)
      if (T0 >= 0 || (DEVMAT && COND)) y[r] = acc;
      else a[base[u] + P.off[r]] = acc;
    }
#pragma unroll
    for (int r = 0; r < D; ++r) {
      if (T0 >= 0) { if (!((r >> (T0 >= 0 ? T0 : 0)) & 1)) st256(a + base[u] + P.off[r], y[r], y[r | (1 << (T0 >= 0 ? T0 : 0))]); }
      else if (DEVMAT && COND) a[base[u] + P.off[r]] = y[r];
    }
  }
}

template <int K, int U, bool COND>
__global__ void __launch_bounds__(256) k_diag(double2* __restrict__ a, uint64_t ngroups, int gshift,
                                               const __grid_constant__ DiagParams<K> P,
                                               const int32_t* __restrict__ cond, int want) {
  constexpr int D = 1 << K;
  const uint64_t g0 = ((uint64_t)blockIdx.x * U) * blockDim.x + threadIdx.x;
  double2 x[U][D];
  uint64_t base[U];
  bool ok[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    uint64_t G = g0 + (uint64_t)u * blockDim.x;
    ok[u] = G < ngroups;
    if (COND) { if (ok[u]) ok[u] = (cond[G >> gshift] == want); }
    base[u] = bt_expand(G, P.plan);
    if (ok[u]) {
#pragma unroll
      for (int j = 0; j < D; ++j) x[u][j] = a[base[u] + P.off[j]];
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (!ok[u]) continue;
#pragma unroll
    for (int j = 0; j < D; ++j) a[base[u] + P.off[j]] = cmul(P.d[j], x[u][j]);
  }
}

// ---- host launch ------------------------------------------------------------------------------------------
static int make_plan(const bt_sv* s, const GateDesc& g, IdxPlan* plan, uint64_t* off) {
  int all[8];
  int n = 0;
  for (int i = 0; i < g.k; ++i) all[n++] = g.tb[i];
  for (int i = 0; i < g.nc; ++i) all[n++] = g.cb[i];
  for (int i = 0; i < n; ++i) {
    if (all[i] < 0 || all[i] >= s->n_local) BT_FAIL(BT_ERR_ARG, "internal: bit %d is not local (n_local=%d)", all[i], s->n_local);
    for (int j = 0; j < i; ++j)
      if (all[i] == all[j]) BT_FAIL(BT_ERR_ARG, "gate acts twice on the same qubit");
  }
  std::sort(all, all + n);
  plan->ni = n;
  for (int i = 0; i < 8; ++i) plan->ins[i] = (i < n) ? all[i] : 0;
  plan->cmask = 0;
  for (int i = 0; i < g.nc; ++i) plan->cmask |= 1ull << g.cb[i];
  for (int j = 0; j < (1 << g.k); ++j) {
    uint64_t o = 0;
    for (int t = 0; t < g.k; ++t)
      if ((j >> t) & 1) o |= 1ull << g.tb[t];
    off[j] = o;
  }
  return BT_OK;
}

template <int K, int U, int T0>
static void launch_dense_t0(bt_sv* s, unsigned grid, uint64_t ngroups, int gshift, const DenseParams<K>& P, const double2* dmats, const int32_t* cond, int want) {
  if (dmats && cond) k_dense<K, U, true, true, T0><<<grid, 256, 0, s->stream>>>(s->amp, ngroups, gshift, P, dmats, cond, want);
  else if (dmats) k_dense<K, U, true, false, T0><<<grid, 256, 0, s->stream>>>(s->amp, ngroups, gshift, P, dmats, nullptr, 0);
  else if (cond) k_dense<K, U, false, true, T0><<<grid, 256, 0, s->stream>>>(s->amp, ngroups, gshift, P, nullptr, cond, want);
  else k_dense<K, U, false, false, T0><<<grid, 256, 0, s->stream>>>(s->amp, ngroups, gshift, P, nullptr, nullptr, 0);
}

template <int K, int U>
static int launch_dense(bt_sv* s, const GateDesc& g, const double2* dmats, const int32_t* cond, int want) {
  DenseParams<K> P;
  BT_TRY(make_plan(s, g, &P.plan, P.off));
  constexpr int D = 1 << K;
  if (!dmats)
    for (int i = 0; i < D * D; ++i) P.m[i] = make_double2(g.m[i].real(), g.m[i].imag());
  int gshift = s->n_local - P.plan.ni;
  uint64_t ngroups = (uint64_t)s->n_batch << gshift;
  uint64_t per_block = 256ull * U;
  unsigned grid = (unsigned)((ngroups + per_block - 1) / per_block);
  // matrix bit that sits on index bit 0 (1- and 2-qubit gates): its pairs move as 256-bit accesses
  int t0 = -1;
  for (int t = 0; t < K; ++t)
    if (g.tb[t] == 0) t0 = t;
  static const bool wide = []() { const char* v = getenv("BT_WIDE_LOADS"); return !(v && *v == '0'); }();
  if (!wide) t0 = -1;
  if (K >= 3 && (dmats || cond)) t0 = -1;  // 3- and 4-bit blocks (superoperators of 2-qubit gates / channels on rho): plain form only
  bt_prof_begin(s, BT_CLS_DENSE);
  if (K <= 2) {
    if (t0 == 0) launch_dense_t0<K, U, 0>(s, grid, ngroups, gshift, P, dmats, cond, want);
    else if (t0 == 1 && K == 2) launch_dense_t0<K, U, (K == 2 ? 1 : -1)>(s, grid, ngroups, gshift, P, dmats, cond, want);
    else launch_dense_t0<K, U, -1>(s, grid, ngroups, gshift, P, dmats, cond, want);
  } else if (t0 < 0) {
    launch_dense_t0<K, U, -1>(s, grid, ngroups, gshift, P, dmats, cond, want);
  } else {
    switch (t0) {
      case 0: k_dense<K, U, false, false, 0><<<grid, 256, 0, s->stream>>>(s->amp, ngroups, gshift, P, nullptr, nullptr, 0); break;
      case 1: k_dense<K, U, false, false, (K >= 3 ? 1 : -1)><<<grid, 256, 0, s->stream>>>(s->amp, ngroups, gshift, P, nullptr, nullptr, 0); break;
      case 2: k_dense<K, U, false, false, (K >= 3 ? 2 : -1)><<<grid, 256, 0, s->stream>>>(s->amp, ngroups, gshift, P, nullptr, nullptr, 0); break;
      default: k_dense<K, U, false, false, (K >= 4 ? 3 : -1)><<<grid, 256, 0, s->stream>>>(s->amp, ngroups, gshift, P, nullptr, nullptr, 0); break;
    }
  }
  bt_prof_end(s);
  BT_CHECK_LAUNCH(s);
  return BT_OK;
}

template <int K, int U>
static int launch_diag(bt_sv* s, const GateDesc& g, const int32_t* cond, int want) {
  DiagParams<K> P;
  BT_TRY(make_plan(s, g, &P.plan, P.off));
  for (int i = 0; i < (1 << K); ++i) P.d[i] = make_double2(g.m[i].real(), g.m[i].imag());
  int gshift = s->n_local - P.plan.ni;
  uint64_t ngroups = (uint64_t)s->n_batch << gshift;
  uint64_t per_block = 256ull * U;
  unsigned grid = (unsigned)((ngroups + per_block - 1) / per_block);
  bt_prof_begin(s, BT_CLS_DIAG);
  if (cond)
    k_diag<K, U, true><<<grid, 256, 0, s->stream>>>(s->amp, ngroups, gshift, P, cond, want);
  else
    k_diag<K, U, false><<<grid, 256, 0, s->stream>>>(s->amp, ngroups, gshift, P, nullptr, 0);
  bt_prof_end(s);
  BT_CHECK_LAUNCH(s);
  return BT_OK;
}

// Resolve bits that live on the rank index of a sharded state: a control on a global bit is a
// rank-conditional skip, a diagonal target on a global bit selects one half of the diagonal.  Returns 1 when
// nothing is left to do on this rank.
int bt_localize(const bt_sv* s, GateDesc* g) {
  int nl = s->n_local;
  // controls
  int nc2 = 0;
  for (int i = 0; i < g->nc; ++i) {
    int b = g->cb[i];
    if (b >= nl) {
      if (!((s->rank >> (b - nl)) & 1)) return 1;
    } else {
      g->cb[nc2++] = b;
    }
  }
  g->nc = nc2;
  for (int t = 0; t < g->k;) {
    int b = g->tb[t];
    if (b < nl) { ++t; continue; }
    if (!g->diag) BT_FAIL(BT_ERR_UNSUPPORTED, "non-diagonal gate on a global qubit: remap required first");
    int v = (s->rank >> (b - nl)) & 1;
    int D = 1 << g->k;
    cplx nd[16];
    int n2 = 0;
    for (int j = 0; j < D; ++j)
      if (((j >> t) & 1) == v) nd[n2++] = g->m[j];
    for (int j = 0; j < n2; ++j) g->m[j] = nd[j];
    for (int i = t; i < g->k - 1; ++i) g->tb[i] = g->tb[i + 1];
    g->k--;
  }
  return 0;
}

int bt_launch_gate(bt_sv* s, const GateDesc& g_in, const int32_t* cond, int want) {
  GateDesc g = g_in;
  if (s->mask_on) {  // trajectory mask (bt_sv_set_mask): the same per-trajectory predicate the ifOp branches use
    if (cond) BT_FAIL(BT_ERR_UNSUPPORTED, "outcome-conditional gates cannot be combined with a trajectory mask: fold the outcome into the mask");
    cond = s->d_mask;
    want = 1;
  }
  if (s->world > 1) {
    int r = bt_localize(s, &g);
    if (r < 0) return r;
    if (r == 1) return BT_OK;
  }
  if (g.diag) {
    bool all_one = true;
    for (int j = 0; j < (1 << g.k); ++j)
      if (g.m[j] != cplx(1, 0)) all_one = false;
    if (all_one) return BT_OK;  // identity: nothing to move through HBM
    switch (g.k) {
      case 0: return launch_diag<0, 4>(s, g, cond, want);
      case 1: return launch_diag<1, 4>(s, g, cond, want);
      case 2: return launch_diag<2, 2>(s, g, cond, want);
      case 3: return launch_diag<3, 1>(s, g, cond, want);
      case 4: return launch_diag<4, 1>(s, g, cond, want);
    }
  } else {
    switch (g.k) {
      case 1: return launch_dense<1, 4>(s, g, nullptr, cond, want);
      case 2: return launch_dense<2, 2>(s, g, nullptr, cond, want);
      case 3: return launch_dense<3, 1>(s, g, nullptr, cond, want);
      case 4: return launch_dense<4, 1>(s, g, nullptr, cond, want);
    }
  }
  BT_FAIL(BT_ERR_ARG, "unsupported gate arity k=%d", g.k);
}

int bt_launch_gate_devmat(bt_sv* s, int k, const int* tb, const double2* d_mats) {
  GateDesc g;
  g.k = k; g.nc = 0; g.diag = false;
  for (int i = 0; i < k; ++i) g.tb[i] = tb[i];
  const int32_t* cond = s->mask_on ? s->d_mask : nullptr;
  switch (k) {
    case 1: return launch_dense<1, 4>(s, g, d_mats, cond, 1);
    case 2: return launch_dense<2, 2>(s, g, d_mats, cond, 1);
    case 3: return launch_dense<3, 1>(s, g, d_mats, cond, 1);
  }
  BT_FAIL(BT_ERR_ARG, "unsupported arity for per-trajectory matrices");
}

// ---- public entry points ----------------------------------------------------------------------------------
// column-major (Julia) -> row-major
static void colmajor_to_rowmajor(const bt_c64* m, int D, cplx* out) {
  for (int r = 0; r < D; ++r)
    for (int c = 0; c < D; ++c) out[r * D + c] = c64(m[r + c * D]);
}

int bt_build_gate(const bt_sv* s, int nq, int qubit, int target, int control, const bt_c64* m, GateDesc* out) {
  int N = s->n_qubits;
  if (!m) BT_FAIL(BT_ERR_ARG, "null matrix");
  if (nq == 1) {
    // src/hilbert.jl:145-149, src/struct.jl:455-462
    if (qubit < 1 || N < qubit || N < control) BT_FAIL(BT_ERR_ARG, "N must be larger than qubit");
    if (control != -2 && control < 1) BT_FAIL(BT_ERR_ARG, "invalid control qubit %d", control);
    if (qubit == control) BT_FAIL(BT_ERR_ARG, "`qubit` must differ from `control` qubit");
    cplx rm[4];
    colmajor_to_rowmajor(m, 2, rm);
    int tb[1] = {s->phys_of_bit[N - qubit]};
    int cb[1] = {control != -2 ? s->phys_of_bit[N - control] : 0};
    bt_canonicalize(1, tb, rm, control != -2 ? 1 : 0, cb, out);
    return BT_OK;
  }
  if (nq == 2) {
    // src/hilbert.jl:22-26, src/struct.jl:463-471
    if (qubit < 1 || target < 1 || N < qubit || N < target || N < control) BT_FAIL(BT_ERR_ARG, "N must be larger than qubits");
    if (control != -2 && control < 1) BT_FAIL(BT_ERR_ARG, "invalid control qubit %d", control);
    if (qubit == target) BT_FAIL(BT_ERR_ARG, "`qubit` and `target_qubit` must differ");
    if (control == qubit || control == target) BT_FAIL(BT_ERR_ARG, "either `qubit` or `target_qubit` must differ from `control` qubit");
    cplx rm[16];
    colmajor_to_rowmajor(m, 4, rm);
    if (control != -2 && bt_is_strict() && abs(qubit - target) != 1) {
      // src/hilbert.jl:58-64: only CX / CZ may be controlled across a distance
      static const double cx[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0};
      static const double cz[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1};
      bool is_cx = true, is_cz = true;
      for (int i = 0; i < 16; ++i) {
        if (rm[i] != cplx(cx[i], 0)) is_cx = false;
        if (rm[i] != cplx(cz[i], 0)) is_cz = false;
      }
      if (!is_cx && !is_cz) BT_FAIL(BT_ERR_UNSUPPORTED, "Unsupported operation. Use 'CCZ' or 'CCX'.");
    }
    // matrix index = 2*b_qubit + b_target  (src/hilbert.jl:30 canonicalises by swapping, same operator)
    int tb[2] = {s->phys_of_bit[N - target], s->phys_of_bit[N - qubit]};
    int cb[1] = {control != -2 ? s->phys_of_bit[N - control] : 0};
    bt_canonicalize(2, tb, rm, control != -2 ? 1 : 0, cb, out);
    return BT_OK;
  }
  if (nq == 3) {
    // src/hilbert.jl:106-115: qubits first, first+1, first+2; first is the MSB of the 8x8 index
    if (qubit < 1 || N < qubit + 2) BT_FAIL(BT_ERR_ARG, "N must be larger than all three qubits");
    cplx rm[64];
    colmajor_to_rowmajor(m, 8, rm);
    int tb[3] = {s->phys_of_bit[N - (qubit + 2)], s->phys_of_bit[N - (qubit + 1)], s->phys_of_bit[N - qubit]};
    bt_canonicalize(3, tb, rm, 0, nullptr, out);
    return BT_OK;
  }
  BT_FAIL(BT_ERR_ARG, "only 1-, 2- and 3-qubit operations are supported");
}

int bt_prepare_local(bt_sv* s, const GateDesc& g);  // bt_dist.cu: remap if a non-diagonal target is global

static int apply_built(bt_sv* s, GateDesc& g, const int32_t* cond, int want) {
  if (s->world > 1) BT_TRY(bt_prepare_local(s, g));
  return bt_launch_gate(s, g, cond, want);
}

extern "C" int bt_sv_apply_1q(bt_sv* s, int qubit, const bt_c64 m[4], int control) {
  BT_TRY(bt_check_sv(s));
  GateDesc g;
  if (s->world > 1) {
    // physical bits may move during the remap: build after preparing, from logical bits
    BT_TRY(bt_build_gate(s, 1, qubit, -1, control, m, &g));
    BT_TRY(bt_prepare_local(s, g));
  }
  BT_TRY(bt_build_gate(s, 1, qubit, -1, control, m, &g));
  return bt_launch_gate(s, g);
}

extern "C" int bt_sv_apply_2q(bt_sv* s, int qubit, int target, const bt_c64 m[16], int control) {
  BT_TRY(bt_check_sv(s));
  GateDesc g;
  if (s->world > 1) {
    BT_TRY(bt_build_gate(s, 2, qubit, target, control, m, &g));
    BT_TRY(bt_prepare_local(s, g));
  }
  BT_TRY(bt_build_gate(s, 2, qubit, target, control, m, &g));
  return bt_launch_gate(s, g);
}

extern "C" int bt_sv_apply_3q(bt_sv* s, int first_qubit, const bt_c64 m[64]) {
  BT_TRY(bt_check_sv(s));
  GateDesc g;
  if (s->world > 1) {
    BT_TRY(bt_build_gate(s, 3, first_qubit, -1, -2, m, &g));
    BT_TRY(bt_prepare_local(s, g));
  }
  BT_TRY(bt_build_gate(s, 3, first_qubit, -1, -2, m, &g));
  return bt_launch_gate(s, g);
}

extern "C" int bt_sv_apply_1q_if(bt_sv* s, int qubit, const bt_c64 m[4], int control, int want) {
  BT_TRY(bt_check_sv(s));
  BT_TRY(bt_ensure_traj(s));
  GateDesc g;
  BT_TRY(bt_build_gate(s, 1, qubit, -1, control, m, &g));
  return apply_built(s, g, s->d_outcome, want);
}

extern "C" int bt_sv_apply_2q_if(bt_sv* s, int qubit, int target, const bt_c64 m[16], int control, int want) {
  BT_TRY(bt_check_sv(s));
  BT_TRY(bt_ensure_traj(s));
  GateDesc g;
  BT_TRY(bt_build_gate(s, 2, qubit, target, control, m, &g));
  return apply_built(s, g, s->d_outcome, want);
}
