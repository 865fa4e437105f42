// bt_tile.cu -- host-side gate fusion + the multi-gate shared-memory tile kernels.
//
// No reference analogue: the reference applies one 2^N x 2^N sparse matrix per op (src/hilbert.jl:505); the
// nearest ideas are the full-circuit product sa.sparse(circ) (src/struct.jl:877-888) and the duplicate-cancelling
// optimize_simple (src/ops.jl:335-367).  Here:
//   1. fusion (fuse_blocks): gates are grouped into blocks on one or two qubits.  A block keeps its dense product
//      (re-canonicalised: a lone CX comes back out as control + X, a CZ/CP chain as a diagonal) AND the list of its gates
//      as structured micro-ops (real / RX-like / diagonal / phase 1-qubit ops, CX, CPHASE), which costs about half the
//      FP64 work of the dense 4x4 for the named gates;
//   2. scheduling (bt_fuse_and_run): blocks are packed greedily (dependency order preserved) into passes whose
//      non-diagonally-acted bits fit a tile of T index bits (low bits, for coalescing, plus chosen high bits); controls
//      and diagonal factors on other bits ride along: they only depend on the thread's group or on the tile's base index;
//   3. each pass is ONE kernel (launch_pass): a CTA brings a 2^T-amplitude tile into shared memory -- one TMA tensor
//      copy (k_tile_tma; 128-byte hardware swizzle) or cp.async (k_tile, when the tile's bits do not form a tensor box)
//      -- runs the pass's items in place between __syncthreads, and writes the tile back.  Items: register programs (a
//      thread holds the 16 amplitudes of 4 tile bits and interprets a micro-op list: run_prog), dense register clusters
//      and single dense gates with per-slot code (run_cluster, run_gate), diagonal sweeps (run_diag).
// HBM roofline per pass: 32 B per amplitude.  A pass carries ~30 gates, so the kernel sits between the HBM and FP64 roofs
// and is tuned as a compute kernel (measurements in profiles/r1_fusion_sweep.txt).
// Build: tools/nvcc_brx.py rewrites the micro-op switch of run_prog into one indirect branch between cicc and ptxas.
#include "bt_internal.cuh"
#include <cuda_pipeline_primitives.h>
#include <cuda.h>   // CUtensorMap (driver types only; the encode function is fetched through cudaGetDriverEntryPoint)
#include <stdlib.h>
#include <atomic>
#include <algorithm>

#include "bt_tile_types.cuh"

// XOR swizzle of 16-byte slots: linear over GF(2), so sw(a | b) = sw(a) ^ sw(b) for disjoint a, b
// (slot index bits 0..2 select the 16-byte bank group; they become the XOR of ALL 3-bit digits of the tile index, so a run
// of 8 lanes is conflict-free whenever the three index bits it varies have distinct residues mod 3)
__host__ __device__ __forceinline__ uint32_t sw(uint32_t c) { return c ^ ((c >> 3) & 7u) ^ ((c >> 6) & 7u) ^ ((c >> 9) & 7u); }
// mode 0: bank group = digit0 ^ digit1 only -- exactly the TMA hardware swizzle CU_TENSOR_MAP_SWIZZLE_128B (16-byte chunk index
// XOR 128-byte row index mod 8), used by the tensor-copy variant of the kernel; mode 1: all digits (cp.async variant)
__host__ __device__ __forceinline__ uint32_t swz(uint32_t c, int mode) { return mode ? sw(c) : (c ^ ((c >> 3) & 7u)); }

__device__ __forceinline__ void cfma2(double2& acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ double2 cmul2(double2 d, double2 x) { return make_double2(d.x * x.x - d.y * x.y, d.x * x.y + d.y * x.x); }

__device__ __forceinline__ uint32_t thread_slot(const uint32_t* bit_sw, uint32_t tid) {
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if ((tid >> k) & 1u) r ^= bit_sw[k];
  return r;
}

template <int GI, int NT>
__device__ __forceinline__ void run_gate(const TileParams& P, double2* __restrict__ sm, uint64_t base, uint32_t tid, uint32_t nloc) {
  const TileGate& G = P.g[GI];
  const int P_swz = P.swz_mode;
  if ((base & G.ext_cmask) != G.ext_cmask) return;  // uniform per CTA
  // local target bits and the fixed (external) part of the matrix index
  int kl = 0;
  uint32_t jext = 0;
  uint32_t so[2] = {0u, 0u};
  int tpos[2] = {0, 0};  // matrix bit of the i-th local target
#pragma unroll
  for (int t = 0; t < 2; ++t)
    if (t < G.k) {
      if (G.tloc[t] >= 0) { so[kl] = swz(1u << G.tloc[t], P_swz); tpos[kl] = t; kl++; }
      else if ((base >> G.text[t]) & 1ull) jext |= 1u << t;
    }
  const uint32_t ng = nloc >> G.ni;
  const bool active = tid < ng;
  const uint32_t niter = G.niter;
  // slot of this thread's first group; further groups and the partner amplitudes are XOR offsets (sw is linear)
  const uint32_t s0 = thread_slot(G.bit_sw, tid) ^ swz(G.lcmask, P_swz);
  if (G.kind == 0) {
    if (G.k == 2) {
      const double2* m = G.m;
      const uint32_t o0 = so[0], o1 = so[1], o01 = so[0] ^ so[1];
      if (active)
        for (uint32_t it = 0; it < niter; ++it) {
          const uint32_t i0 = s0 ^ G.iter_sw[it], i1 = i0 ^ o0, i2 = i0 ^ o1, i3 = i0 ^ o01;
          double2 x0 = sm[i0], x1 = sm[i1], x2 = sm[i2], x3 = sm[i3];
          double2 y0 = make_double2(0, 0), y1 = y0, y2 = y0, y3 = y0;
          cfma2(y0, m[0], x0); cfma2(y0, m[1], x1); cfma2(y0, m[2], x2); cfma2(y0, m[3], x3);
          cfma2(y1, m[4], x0); cfma2(y1, m[5], x1); cfma2(y1, m[6], x2); cfma2(y1, m[7], x3);
          cfma2(y2, m[8], x0); cfma2(y2, m[9], x1); cfma2(y2, m[10], x2); cfma2(y2, m[11], x3);
          cfma2(y3, m[12], x0); cfma2(y3, m[13], x1); cfma2(y3, m[14], x2); cfma2(y3, m[15], x3);
          sm[i0] = y0; sm[i1] = y1; sm[i2] = y2; sm[i3] = y3;
        }
    } else {  // k == 1
      const double2 m0 = G.m[0], m1 = G.m[1], m2 = G.m[2], m3 = G.m[3];
      const uint32_t o0 = so[0];
      if (active)
        for (uint32_t it = 0; it < niter; ++it) {
          const uint32_t i0 = s0 ^ G.iter_sw[it], i1 = i0 ^ o0;
          double2 x0 = sm[i0], x1 = sm[i1];
          double2 y0 = make_double2(0, 0), y1 = y0;
          cfma2(y0, m0, x0); cfma2(y0, m1, x1);
          cfma2(y1, m2, x0); cfma2(y1, m3, x1);
          sm[i0] = y0; sm[i1] = y1;
        }
    }
  } else {
    // diagonal: entries with matrix index j; local bits enumerate, external bits fixed by the tile base
    if (kl == 0) {
      const double2 d = G.m[jext];
      if (active && !(d.x == 1.0 && d.y == 0.0))
        for (uint32_t it = 0; it < niter; ++it) {
          const uint32_t i0 = s0 ^ G.iter_sw[it];
          sm[i0] = cmul2(d, sm[i0]);
        }
    } else if (kl == 1) {
      const double2 d0 = G.m[jext], d1 = G.m[jext | (1u << tpos[0])];
      if (active)
        for (uint32_t it = 0; it < niter; ++it) {
          const uint32_t i0 = s0 ^ G.iter_sw[it], i1 = i0 ^ so[0];
          double2 x0 = sm[i0], x1 = sm[i1];
          sm[i0] = cmul2(d0, x0);
          sm[i1] = cmul2(d1, x1);
        }
    } else {
      const double2 d0 = G.m[0], d1 = G.m[1u << tpos[0]], d2 = G.m[1u << tpos[1]], d3 = G.m[3];
      if (active)
        for (uint32_t it = 0; it < niter; ++it) {
          const uint32_t i0 = s0 ^ G.iter_sw[it], i1 = i0 ^ so[0], i2 = i0 ^ so[1], i3 = i1 ^ so[1];
          double2 x0 = sm[i0], x1 = sm[i1], x2 = sm[i2], x3 = sm[i3];
          sm[i0] = cmul2(d0, x0);
          sm[i1] = cmul2(d1, x1);
          sm[i2] = cmul2(d2, x2);
          sm[i3] = cmul2(d3, x3);
        }
    }
  }
}

template <int NT>
__device__ __noinline__ void run_diag(const TileDiag& G, double2* __restrict__ sm, uint64_t base, uint32_t tid, uint32_t nloc, int P_swz) {
  if ((base & G.ext_cmask) != G.ext_cmask) return;  // uniform per CTA
  int kl = 0;
  uint32_t jext = 0;
  uint32_t so[2] = {0u, 0u};
  int tpos[2] = {0, 0};
#pragma unroll
  for (int t = 0; t < 2; ++t)
    if (t < G.k) {
      if (G.tloc[t] >= 0) { so[kl] = swz(1u << G.tloc[t], P_swz); tpos[kl] = t; kl++; }
      else if ((base >> G.text[t]) & 1ull) jext |= 1u << t;
    }
  const uint32_t ng = nloc >> G.ni;
  if (tid >= ng) return;
  const uint32_t niter = G.niter;
  const uint32_t s0 = thread_slot(G.bit_sw, tid) ^ swz(G.lcmask, P_swz);
  if (kl == 0) {
    const double2 d = G.m[jext];
    if (!(d.x == 1.0 && d.y == 0.0))
      for (uint32_t it = 0; it < niter; ++it) {
        const uint32_t i0 = s0 ^ G.iter_sw[it];
        sm[i0] = cmul2(d, sm[i0]);
      }
  } else if (kl == 1) {
    const double2 d0 = G.m[jext], d1 = G.m[jext | (1u << tpos[0])];
    for (uint32_t it = 0; it < niter; ++it) {
      const uint32_t i0 = s0 ^ G.iter_sw[it], i1 = i0 ^ so[0];
      double2 x0 = sm[i0], x1 = sm[i1];
      sm[i0] = cmul2(d0, x0);
      sm[i1] = cmul2(d1, x1);
    }
  } else {
    const double2 d0 = G.m[0], d1 = G.m[1u << tpos[0]], d2 = G.m[1u << tpos[1]], d3 = G.m[3];
    for (uint32_t it = 0; it < niter; ++it) {
      const uint32_t i0 = s0 ^ G.iter_sw[it], i1 = i0 ^ so[0], i2 = i0 ^ so[1], i3 = i1 ^ so[1];
      double2 x0 = sm[i0], x1 = sm[i1], x2 = sm[i2], x3 = sm[i3];
      sm[i0] = cmul2(d0, x0);
      sm[i1] = cmul2(d1, x1);
      sm[i2] = cmul2(d2, x2);
      sm[i3] = cmul2(d3, x3);
    }
  }
}

// dense 4x4 on cluster positions PL < PH of the CL_AMPS register-resident amplitudes x[] (index bit p <-> position p)
template <int PL, int PH>
__device__ __forceinline__ void cl_apply(double2 (&x)[CL_AMPS], const double2* __restrict__ m) {
#pragma unroll
  for (int r = 0; r < CL_AMPS / 4; ++r) {
    // enumerate the positions other than PL, PH
    int base = 0, rb = r;
#pragma unroll
    for (int p = 0; p < CL_BITS; ++p)
      if (p != PL && p != PH) { base |= (rb & 1) << p; rb >>= 1; }
    const int i0 = base, i1 = base | (1 << PL), i2 = base | (1 << PH), i3 = base | (1 << PL) | (1 << PH);
    const double2 x0 = x[i0], x1 = x[i1], x2 = x[i2], x3 = x[i3];
    double2 y0 = make_double2(0, 0), y1 = y0, y2 = y0, y3 = y0;
    cfma2(y0, m[0], x0); cfma2(y0, m[1], x1); cfma2(y0, m[2], x2); cfma2(y0, m[3], x3);
    cfma2(y1, m[4], x0); cfma2(y1, m[5], x1); cfma2(y1, m[6], x2); cfma2(y1, m[7], x3);
    cfma2(y2, m[8], x0); cfma2(y2, m[9], x1); cfma2(y2, m[10], x2); cfma2(y2, m[11], x3);
    cfma2(y3, m[12], x0); cfma2(y3, m[13], x1); cfma2(y3, m[14], x2); cfma2(y3, m[15], x3);
    x[i0] = y0; x[i1] = y1; x[i2] = y2; x[i3] = y3;
  }
}

template <int CI, int NT>
__device__ __forceinline__ void run_cluster(const TileParams& P, double2* __restrict__ sm, uint32_t tid, uint32_t nloc) {
  const TileCluster& Cl = P.cl[CI];
  const int P_swz = P.swz_mode;
  const uint32_t ng = nloc >> CL_BITS;
  if (tid >= ng) return;
  const uint32_t s0 = thread_slot(Cl.bit_sw, tid);
  uint32_t o[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) o[p] = (p < CL_BITS) ? swz(1u << Cl.lp[p], P_swz) : 0u;
  const uint32_t use = Cl.use;
  uint32_t off[CL_AMPS];
#pragma unroll
  for (int j = 0; j < CL_AMPS; ++j) off[j] = ((j & 1) ? o[0] : 0u) ^ ((j & 2) ? o[1] : 0u) ^ ((j & 4) ? o[2] : 0u) ^ ((j & 8) ? o[3] : 0u);
  uint32_t it = 0;
#if CL_PIPE
  // two cluster groups in flight per thread: the shared-memory loads of one overlap the DFMA chains of the other
  for (; it + 1 < Cl.niter; it += 2) {
    const uint32_t ba = s0 ^ Cl.iter_sw[it], bb = s0 ^ Cl.iter_sw[it + 1];
    double2 xa[CL_AMPS], xb[CL_AMPS];
#pragma unroll
    for (int j = 0; j < CL_AMPS; ++j) xa[j] = sm[ba ^ off[j]];
#pragma unroll
    for (int j = 0; j < CL_AMPS; ++j) xb[j] = sm[bb ^ off[j]];
#if CL_BITS == 4
    if (use & 1u) { cl_apply<0, 1>(xa, Cl.m[0]); cl_apply<0, 1>(xb, Cl.m[0]); }
    if (use & 2u) { cl_apply<2, 3>(xa, Cl.m[1]); cl_apply<2, 3>(xb, Cl.m[1]); }
    if (use & 4u) { cl_apply<1, 2>(xa, Cl.m[2]); cl_apply<1, 2>(xb, Cl.m[2]); }
#else
    if (use & 1u) { cl_apply<0, 1>(xa, Cl.m[0]); cl_apply<0, 1>(xb, Cl.m[0]); }
    if (use & 2u) { cl_apply<1, 2>(xa, Cl.m[1]); cl_apply<1, 2>(xb, Cl.m[1]); }
    if (use & 4u) { cl_apply<0, 1>(xa, Cl.m[2]); cl_apply<0, 1>(xb, Cl.m[2]); }
#endif
#pragma unroll
    for (int j = 0; j < CL_AMPS; ++j) sm[ba ^ off[j]] = xa[j];
#pragma unroll
    for (int j = 0; j < CL_AMPS; ++j) sm[bb ^ off[j]] = xb[j];
  }
#endif
  for (; it < Cl.niter; ++it) {
    const uint32_t b = s0 ^ Cl.iter_sw[it];
    double2 x[CL_AMPS];
#pragma unroll
    for (int j = 0; j < CL_AMPS; ++j) x[j] = sm[b ^ off[j]];
#if CL_BITS == 4
    if (use & 1u) cl_apply<0, 1>(x, Cl.m[0]);
    if (use & 2u) cl_apply<2, 3>(x, Cl.m[1]);
    if (use & 4u) cl_apply<1, 2>(x, Cl.m[2]);
#if CL_SLOTS > 3
    if (use & 8u) cl_apply<0, 1>(x, Cl.m[3]);
    if (use & 16u) cl_apply<2, 3>(x, Cl.m[4]);
#endif
#else
    if (use & 1u) cl_apply<0, 1>(x, Cl.m[0]);
    if (use & 2u) cl_apply<1, 2>(x, Cl.m[1]);
    if (use & 4u) cl_apply<0, 1>(x, Cl.m[2]);
#endif
#pragma unroll
    for (int j = 0; j < CL_AMPS; ++j) sm[b ^ off[j]] = x[j];
  }
}

// ---- register programs ------------------------------------------------------------------------------------------------
#include "bt_prog_ops.cuh"

// markers for tools/ptx_brx.py (PTX comments: no code)
#define BT_CASE_MARK(site) asm volatile("// BT_CASE %0;" ::"n"(site))
#if PROG_BITS >= 5
#define PROG_CASE_U1_4(kind) case PROG_SITE_U1(kind, 4): BT_CASE_MARK(PROG_SITE_U1(kind, 4)); prog_u1<kind, 4>(x, c, cp); break;
#else
#define PROG_CASE_U1_4(kind)
#endif
#define PROG_CASE_U1(kind)                                                   \
  case PROG_SITE_U1(kind, 0): BT_CASE_MARK(PROG_SITE_U1(kind, 0)); prog_u1<kind, 0>(x, c, cp); break;                 \
  case PROG_SITE_U1(kind, 1): BT_CASE_MARK(PROG_SITE_U1(kind, 1)); prog_u1<kind, 1>(x, c, cp); break;                 \
  case PROG_SITE_U1(kind, 2): BT_CASE_MARK(PROG_SITE_U1(kind, 2)); prog_u1<kind, 2>(x, c, cp); break;                 \
  case PROG_SITE_U1(kind, 3): BT_CASE_MARK(PROG_SITE_U1(kind, 3)); prog_u1<kind, 3>(x, c, cp); break;                 \
  PROG_CASE_U1_4(kind)
#define PROG_CASE_CX(a, b) case PROG_SITE_CX(a, b): BT_CASE_MARK(PROG_SITE_CX(a, b)); prog_cx<a, b>(x); break;
#define PROG_CASE_CP(a, b) case PROG_SITE_CPHASE(a, b): BT_CASE_MARK(PROG_SITE_CPHASE(a, b)); prog_cphase<a, b>(x, c); break;
#define PROG_CASE_CPH1(p)                                                                                           \
  case PROG_SITE_CPH1(p): {                                                                                         \
    BT_CASE_MARK(PROG_SITE_CPH1(p));                                                                               \
    const uint64_t em = (uint64_t)__double_as_longlong(c[2]), lm = (uint64_t)__double_as_longlong(c[3]);            \
    if ((base & em) == em && (gl & lm) == lm) prog_u1<PK_PHASE, p>(x, c, cp);                                           \
  } break;
#define PROG_CASE_CCX1(p)                                                                                           \
  case PROG_SITE_CCX1(p): {                                                                                         \
    BT_CASE_MARK(PROG_SITE_CCX1(p));                                                                               \
    const uint64_t em = (uint64_t)__double_as_longlong(c[0]), lm = (uint64_t)__double_as_longlong(c[1]);            \
    if ((base & em) == em && (gl & lm) == lm) prog_x1<p>(x);                                                        \
  } break;

template <int NT>
__device__ __forceinline__ void run_prog(const TileParams& P, int pi, double2* __restrict__ sm, uint64_t base, uint32_t tid, uint32_t nloc) {
  const TileProg& G = P.pr[pi];
  const int P_swz = P.swz_mode;
  const uint32_t ng = nloc >> PROG_BITS;
  if (tid >= ng) return;
  const uint32_t s0 = thread_slot(G.bit_sw, tid);
  uint32_t o[PROG_BITS];
#pragma unroll
  for (int q = 0; q < PROG_BITS; ++q) o[q] = swz(1u << G.lp[q], P_swz);
  const uint32_t nops = G.nops;
  const uint32_t g0 = thread_slot(G.bit_lin, tid);
  for (uint32_t it = 0; it < G.niter; ++it) {
    const uint32_t b = s0 ^ G.iter_sw[it];
    const uint64_t gl = g0 ^ G.iter_lin[it];
    double2 x[PROG_AMPS];
#pragma unroll
    for (int j = 0; j < PROG_AMPS; ++j) {
      uint32_t a = b;
#pragma unroll
      for (int q = 0; q < PROG_BITS; ++q) if ((j >> q) & 1) a ^= o[q];
      x[j] = sm[a];
    }
    // Three dependent constant loads would sit on the critical path of every micro-op (op word -> jump-table entry ->
    // coefficients) with only three warps per scheduler to hide them.  So: the op word of op k+1 is fetched during op k
    // (op[nops] is a padding entry), and the first four coefficients are fetched BEFORE the indirect branch (the marker
    // takes them as inputs), in the shadow of the jump-table load.
    uint32_t opw = G.op[0];
    for (uint32_t k = 0; k < nops; ++k) {
      const uint32_t op = opw;
      opw = G.op[k + 1];
      const double* __restrict__ cp = G.coef + (op >> 8);
      const double c[4] = {cp[0], cp[1], cp[2], cp[3]};
      const uint32_t site = op & 0xffu;
      // tools/ptx_brx.py turns the switch below into one brx.idx
      asm volatile("// BT_DISPATCH %0;" ::"r"(site), "r"(opw), "d"(c[0]), "d"(c[1]), "d"(c[2]), "d"(c[3]));
      switch (site) {
        PROG_CASE_U1(PK_GEN)
        PROG_CASE_U1(PK_REAL)
        PROG_CASE_U1(PK_RXL)
        PROG_CASE_U1(PK_DIAG)
        PROG_CASE_U1(PK_PHASE)
        PROG_CASE_CX(0, 1) PROG_CASE_CX(0, 2) PROG_CASE_CX(0, 3)
        PROG_CASE_CX(1, 0) PROG_CASE_CX(1, 2) PROG_CASE_CX(1, 3)
        PROG_CASE_CX(2, 0) PROG_CASE_CX(2, 1) PROG_CASE_CX(2, 3)
        PROG_CASE_CX(3, 0) PROG_CASE_CX(3, 1) PROG_CASE_CX(3, 2)
        PROG_CASE_CP(0, 1) PROG_CASE_CP(0, 2) PROG_CASE_CP(0, 3)
        PROG_CASE_CP(1, 2) PROG_CASE_CP(1, 3) PROG_CASE_CP(2, 3)
#if PROG_BITS >= 5
        PROG_CASE_CX(0, 4) PROG_CASE_CX(1, 4) PROG_CASE_CX(2, 4) PROG_CASE_CX(3, 4)
        PROG_CASE_CX(4, 0) PROG_CASE_CX(4, 1) PROG_CASE_CX(4, 2) PROG_CASE_CX(4, 3)
        PROG_CASE_CP(0, 4) PROG_CASE_CP(1, 4) PROG_CASE_CP(2, 4) PROG_CASE_CP(3, 4)
        PROG_CASE_CPH1(4) PROG_CASE_CCX1(4)
#endif
        case PROG_SITE_CSCALE: {
          BT_CASE_MARK(PROG_SITE_CSCALE);
          const uint64_t em = (uint64_t)__double_as_longlong(c[2]), lm = (uint64_t)__double_as_longlong(c[3]);
          if ((base & em) == em && (gl & lm) == lm) {
            const double dr = c[0], di = c[1], ndi = -di;
#pragma unroll
            for (int i = 0; i < PROG_AMPS; ++i) ip_cmul(x[i].x, x[i].y, dr, di, ndi);
          }
        } break;
        PROG_CASE_CPH1(0) PROG_CASE_CPH1(1) PROG_CASE_CPH1(2) PROG_CASE_CPH1(3)
        PROG_CASE_CCX1(0) PROG_CASE_CCX1(1) PROG_CASE_CCX1(2) PROG_CASE_CCX1(3)
        default: BT_CASE_MARK(255); break;
      }
    }
#pragma unroll
    for (int j = 0; j < PROG_AMPS; ++j) {
      uint32_t a = b;
#pragma unroll
      for (int q = 0; q < PROG_BITS; ++q) if ((j >> q) & 1) a ^= o[q];
      sm[a] = x[j];
    }
  }
}

// one code copy per slot: the slot index is a compile-time constant inside, so every matrix element is a uniform-register /
// constant-bank operand of its DFMA instead of a live register.  (A single run-time-indexed copy was measured 20 % slower.)
template <int NT, bool FULL = true>
__device__ __forceinline__ void run_item(int item, const TileParams& P, double2* __restrict__ sm, uint64_t base, uint32_t tid, uint32_t nloc) {
  if (item >= TILE_PBASE) { run_prog<NT>(P, item - TILE_PBASE, sm, base, tid, nloc); return; }
  if (item >= TILE_DBASE) { run_diag<NT>(P.d[item - TILE_DBASE], sm, base, tid, nloc, P.swz_mode); return; }
  if (!FULL) return;  // the lite kernel (register programs + diagonal slots only) never sees dense slots
  switch (item) {
    case 0: run_gate<0, NT>(P, sm, base, tid, nloc); break;
    case 1: run_gate<1, NT>(P, sm, base, tid, nloc); break;
    case 2: run_gate<2, NT>(P, sm, base, tid, nloc); break;
#if TILE_MAXG > 4
    case 3: run_gate<3, NT>(P, sm, base, tid, nloc); break;
    case 4: run_gate<4, NT>(P, sm, base, tid, nloc); break;
    case 5: run_gate<5, NT>(P, sm, base, tid, nloc); break;
    case 6: run_gate<6, NT>(P, sm, base, tid, nloc); break;
    case 7: run_gate<7, NT>(P, sm, base, tid, nloc); break;
#else
    case 3: run_gate<3, NT>(P, sm, base, tid, nloc); break;
#endif
    case TILE_MAXG + 0: run_cluster<0, NT>(P, sm, tid, nloc); break;
    case TILE_MAXG + 1: run_cluster<1, NT>(P, sm, tid, nloc); break;
#if TILE_MAXC > 3
    case TILE_MAXG + 2: run_cluster<2, NT>(P, sm, tid, nloc); break;
    default: run_cluster<3, NT>(P, sm, tid, nloc); break;
#else
    default: run_cluster<2, NT>(P, sm, tid, nloc); break;
#endif
  }
}

__global__ void __launch_bounds__(TILE_THREADS, TILE_MINB) k_tile(double2* __restrict__ a, const __grid_constant__ TileParams P) {
  extern __shared__ double2 sm[];
  __shared__ uint64_t hi_off[1 << (TILE_TMAX - TILE_LOWB_MIN)];
  const int T = P.T, lowb = P.lowb;
  const uint32_t tid = threadIdx.x;
  const uint32_t nloc = 1u << T;
  // base index of this tile: blockIdx with zeros inserted at every tile bit position
  uint64_t base = (uint64_t)blockIdx.x << lowb;
  for (int j = lowb; j < T; ++j) {
    int b = P.tbits[j];
    base = ((base >> b) << (b + 1)) | (base & ((1ull << b) - 1ull));
  }
  const uint32_t nhi = 1u << (T - lowb);
  for (uint32_t h = tid; h < nhi; h += TILE_THREADS) {
    uint64_t o = 0;
    for (int j = lowb; j < T; ++j)
      if ((h >> (j - lowb)) & 1u) o |= 1ull << P.tbits[j];
    hi_off[h] = o;
  }
  __syncthreads();
  const uint32_t lmask = (1u << lowb) - 1u;
  const int dbg = P.stagger_ns < 0 ? -P.stagger_ns : 0;  // measurement aid: 1 = no gates, 2 = no HBM traffic (results invalid)
  if (dbg != 2) {
    for (uint32_t c = tid; c < nloc; c += TILE_THREADS) {
      const double2* src = a + (base + hi_off[c >> lowb] + (c & lmask));
      __pipeline_memcpy_async(&sm[sw(c)], src, sizeof(double2));
    }
    __pipeline_commit();
    __pipeline_wait_prior(0);
  }
  __syncthreads();

  if (dbg != 1)
    for (int i = 0; i < P.nitems; ++i) {
      run_item<TILE_THREADS>(P.item[i], P, sm, base, tid, nloc);
      __syncthreads();
    }

  if (dbg != 2)
    for (uint32_t c = tid; c < nloc; c += TILE_THREADS) a[base + hi_off[c >> lowb] + (c & lmask)] = sm[sw(c)];
}

// ---- tensor-copy (TMA) variant ------------------------------------------------------------------------------------------
// The tile moves between HBM and shared memory with ONE cp.async.bulk.tensor each way instead of 32 cp.async + 32 LDS/STG per
// thread.  The tile's bit set {0..4} + runs of consecutive higher bits is described as a <= 5-D box of a tensor map over the state:
// dim 0 = bits 0..2 (128 B, hardware 128-B swizzle == swz mode 0), dim k = the index bits from the start of run k up to the next
// run, box = the run.  This takes the global traffic off the LSU / L1TEX data pipe that the gates' shared-memory round trips need.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// FULL = false: passes made of register programs and diagonal slots only -- a third fewer registers (no dense 4x4 code), so
// four to five CTAs fit an SM at T = 11 and more of them are in their compute phase at any time.
#ifndef TILE_LITE_MINB
#define TILE_LITE_MINB (PROG_BITS >= 5 ? 2 : 3)
#endif
template <bool FULL>
__global__ void __launch_bounds__(TILE_THREADS, FULL ? TILE_MINB : TILE_LITE_MINB) k_tile_tma(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TileParams P) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte aligned tile (required by the 128-B swizzle), then the mbarrier
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
  double2* sm = reinterpret_cast<double2*>(smem_raw + pad);
  const int T = P.T, lowb = P.lowb;
  const uint32_t nloc = 1u << T;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + pad + ((size_t)sizeof(double2) << T));
  const uint32_t tid = threadIdx.x;
  uint64_t base = (uint64_t)blockIdx.x << lowb;
  for (int j = lowb; j < T; ++j) {
    int b = P.tbits[j];
    base = ((base >> b) << (b + 1)) | (base & ((1ull << b) - 1ull));
  }
  int32_t c1 = (int32_t)((base >> P.tma_coord_shift[1]) & P.tma_coord_mask[1]);
  int32_t c2 = (int32_t)((base >> P.tma_coord_shift[2]) & P.tma_coord_mask[2]);
  int32_t c3 = (int32_t)((base >> P.tma_coord_shift[3]) & P.tma_coord_mask[3]);
  int32_t c4 = (int32_t)((base >> P.tma_coord_shift[4]) & P.tma_coord_mask[4]);
  const uint32_t mb = smem_u32(mbar), dst = smem_u32(sm);
  const uint64_t tm = reinterpret_cast<uint64_t>(&tmap);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int dbg = P.stagger_ns < 0 ? -P.stagger_ns : 0;  // measurement aid: 1 = no gates, 2 = no HBM traffic (results invalid)
  if (tid == 0 && dbg != 2) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((uint32_t)(sizeof(double2) << T)) : "memory");
    const uint32_t chunk = (uint32_t)(sizeof(double2) << T) / (uint32_t)P.tma_ncopy;
    for (int e = 0; e < P.tma_ncopy; ++e)
      asm volatile(
          "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst + e * chunk),
          "l"(tm), "r"(0), "r"(c1), "r"(c2), "r"(c3), "r"(c4 + P.tma_c4add[e]), "r"(mb)
          : "memory");
  }
  // wait for the tile (phase 0); bounded spin so that a descriptor mistake traps instead of hanging the GPU
  if (dbg != 2) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mb) : "memory");
      if (spin > (1u << 22)) __trap();
    }
  }

  for (int i = 0; i < (dbg == 1 ? 0 : P.nitems); ++i) {
    run_item<TILE_THREADS, FULL>(P.item[i], P, sm, base, tid, nloc);
    __syncthreads();
  }

  // generic-proxy writes -> visible to the async proxy, then one bulk tensor store
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0 && dbg != 2) {
    const uint32_t chunk = (uint32_t)(sizeof(double2) << T) / (uint32_t)P.tma_ncopy;
    for (int e = 0; e < P.tma_ncopy; ++e)
      asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];" ::"l"(tm), "r"(0), "r"(c1), "r"(c2),
                   "r"(c3), "r"(c4 + P.tma_c4add[e]), "r"(dst + e * chunk)
                   : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must stay valid until it has been read
  }
}

// Pipelined tensor-copy variant (BT_TILE_PIPE = K > 1, tiles of <= 2^11 amplitudes): a CTA owns K consecutive tiles and two
// tile buffers.  The load of tile i+1 (and i+2 once the store of tile i has left its buffer) is in flight while the items of
// tile i run, and the store of tile i drains while tile i+1 computes: the only exposed copies are the first load and the
// last store of the CTA.
template <bool FULL>
__global__ void __launch_bounds__(TILE_THREADS, FULL ? TILE_MINB : TILE_LITE_MINB) k_tile_pipe(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TileParams P,
                                                                                               int K) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
  const int T = P.T, lowb = P.lowb;
  const uint32_t nloc = 1u << T;
  const uint32_t tile_bytes = (uint32_t)(sizeof(double2) << T);
  double2* buf0 = reinterpret_cast<double2*>(smem_raw + pad);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + pad + 2 * (size_t)tile_bytes);
  const uint32_t tid = threadIdx.x;
  const uint32_t mb0 = smem_u32(mbar), dst0 = smem_u32(buf0);
  const uint64_t tm = reinterpret_cast<uint64_t>(&tmap);
  const uint64_t t0 = (uint64_t)blockIdx.x * (uint64_t)K;
  const uint32_t chunk = tile_bytes / (uint32_t)P.tma_ncopy;

  auto tile_base = [&](uint64_t t) {
    uint64_t base = t << lowb;
    for (int j = lowb; j < T; ++j) {
      int b = P.tbits[j];
      base = ((base >> b) << (b + 1)) | (base & ((1ull << b) - 1ull));
    }
    return base;
  };
  auto issue_load = [&](int i) {  // thread 0 only
    const uint64_t base = tile_base(t0 + i);
    const int32_t c1 = (int32_t)((base >> P.tma_coord_shift[1]) & P.tma_coord_mask[1]), c2 = (int32_t)((base >> P.tma_coord_shift[2]) & P.tma_coord_mask[2]);
    const int32_t c3 = (int32_t)((base >> P.tma_coord_shift[3]) & P.tma_coord_mask[3]), c4 = (int32_t)((base >> P.tma_coord_shift[4]) & P.tma_coord_mask[4]);
    const uint32_t mb = mb0 + 8u * (i & 1), dst = dst0 + (i & 1) * tile_bytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(tile_bytes) : "memory");
    for (int e = 0; e < P.tma_ncopy; ++e)
      asm volatile(
          "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst + e * chunk),
          "l"(tm), "r"(0), "r"(c1), "r"(c2), "r"(c3), "r"(c4 + P.tma_c4add[e]), "r"(mb)
          : "memory");
  };

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb0 + 8u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    issue_load(0);
    if (K > 1) issue_load(1);
  }
  for (int i = 0; i < K; ++i) {
    const uint32_t mb = mb0 + 8u * (i & 1), dst = dst0 + (i & 1) * tile_bytes;
    double2* sm = buf0 + (size_t)(i & 1) * nloc;
    const uint64_t base = tile_base(t0 + i);
    {
      const uint32_t parity = (uint32_t)(i >> 1) & 1u;
      uint32_t done = 0;
      for (uint32_t spin = 0; !done; ++spin) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mb), "r"(parity) : "memory");
        if (spin > (1u << 22)) __trap();
      }
    }
    for (int it = 0; it < P.nitems; ++it) {
      run_item<TILE_THREADS, FULL>(P.item[it], P, sm, base, tid, nloc);
      __syncthreads();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      const int32_t c1 = (int32_t)((base >> P.tma_coord_shift[1]) & P.tma_coord_mask[1]), c2 = (int32_t)((base >> P.tma_coord_shift[2]) & P.tma_coord_mask[2]);
      const int32_t c3 = (int32_t)((base >> P.tma_coord_shift[3]) & P.tma_coord_mask[3]), c4 = (int32_t)((base >> P.tma_coord_shift[4]) & P.tma_coord_mask[4]);
      for (int e = 0; e < P.tma_ncopy; ++e)
        asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];" ::"l"(tm), "r"(0), "r"(c1), "r"(c2),
                     "r"(c3), "r"(c4 + P.tma_c4add[e]), "r"(dst + e * chunk)
                     : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (i + 2 < K) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the store has left this buffer: refill it
        issue_load(i + 2);
      }
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- host: fusion ------------------------------------------------------------------------------------------------------
namespace {

// structured micro-op (see TileProg): the gate as written, before it is multiplied into a dense block
struct MOp {
  int kind;          // PK_*
  int b[2];          // physical bits: U1 -> b[0]; CX -> control b[0], target b[1]; CPHASE -> the pair
  int ncoef;
  double c[8];
  double cost;       // FP64 instructions per amplitude
  double flops;      // FP64 flops per amplitude (FMA = 2)
};

static void mop_finish(MOp& o) {
  static const int ncoef[7] = {8, 4, 4, 4, 2, 0, 2};
  static const double cost[7] = {8, 4, 4, 4, 2, 0.5, 1};
  static const double flops[7] = {14, 6, 6, 6, 3, 0, 1.5};
  o.ncoef = ncoef[o.kind]; o.cost = cost[o.kind]; o.flops = flops[o.kind];
}

// classify a 2x2 (row-major) on physical bit `bit`; returns false for the identity (nothing to do)
static bool mop_from_2x2(const cplx* m, int bit, MOp& o) {
  o.b[0] = bit; o.b[1] = -1;
  const cplx z(0, 0);
  if (m[1] == z && m[2] == z) {
    if (m[0] == cplx(1, 0)) {
      if (m[3] == cplx(1, 0)) return false;
      o.kind = PK_PHASE; o.c[0] = m[3].real(); o.c[1] = m[3].imag();
    } else {
      o.kind = PK_DIAG; o.c[0] = m[0].real(); o.c[1] = m[0].imag(); o.c[2] = m[3].real(); o.c[3] = m[3].imag();
    }
  } else if (m[0].imag() == 0 && m[1].imag() == 0 && m[2].imag() == 0 && m[3].imag() == 0) {
    o.kind = PK_REAL; for (int i = 0; i < 4; ++i) o.c[i] = m[i].real();
  } else if (m[0].imag() == 0 && m[3].imag() == 0 && m[1].real() == 0 && m[2].real() == 0) {
    o.kind = PK_RXL; o.c[0] = m[0].real(); o.c[1] = m[1].imag(); o.c[2] = m[2].imag(); o.c[3] = m[3].real();
  } else {
    o.kind = PK_GEN; for (int i = 0; i < 4; ++i) { o.c[2 * i] = m[i].real(); o.c[2 * i + 1] = m[i].imag(); }
  }
  mop_finish(o);
  return true;
}

static void mop_to_2x2(const MOp& o, cplx* m) {
  const cplx z(0, 0);
  switch (o.kind) {
    case PK_GEN: for (int i = 0; i < 4; ++i) m[i] = cplx(o.c[2 * i], o.c[2 * i + 1]); break;
    case PK_REAL: for (int i = 0; i < 4; ++i) m[i] = cplx(o.c[i], 0); break;
    case PK_RXL: m[0] = cplx(o.c[0], 0); m[1] = cplx(0, o.c[1]); m[2] = cplx(0, o.c[2]); m[3] = cplx(o.c[3], 0); break;
    case PK_DIAG: m[0] = cplx(o.c[0], o.c[1]); m[1] = z; m[2] = z; m[3] = cplx(o.c[2], o.c[3]); break;
    default: m[0] = cplx(1, 0); m[1] = z; m[2] = z; m[3] = cplx(o.c[0], o.c[1]); break;  // PK_PHASE
  }
}

// structured form of one canonical gate; `none` = identity.  false: no structured form (the block goes dense).
static bool mop_classify(const GateDesc& g, MOp& o, bool* none) {
  *none = false;
  if (g.k == 1 && g.nc == 0) {
    cplx m[4];
    if (g.diag) { m[0] = g.m[0]; m[1] = m[2] = cplx(0, 0); m[3] = g.m[1]; }
    else for (int i = 0; i < 4; ++i) m[i] = g.m[i];
    if (!mop_from_2x2(m, g.tb[0], o)) *none = true;
    return true;
  }
  if (g.k == 0 && g.nc == 1) {
    cplx m[4] = {cplx(1, 0), cplx(0, 0), cplx(0, 0), g.m[0]};
    if (!mop_from_2x2(m, g.cb[0], o)) *none = true;
    return true;
  }
  if (g.k == 0 && g.nc == 2) {
    if (g.m[0] == cplx(1, 0)) { *none = true; return true; }
    o.kind = PK_CPHASE; o.b[0] = g.cb[0]; o.b[1] = g.cb[1]; o.c[0] = g.m[0].real(); o.c[1] = g.m[0].imag();
    mop_finish(o);
    return true;
  }
  if (g.k == 1 && g.nc == 1 && !g.diag && g.m[0] == cplx(0, 0) && g.m[3] == cplx(0, 0) && g.m[1] == cplx(1, 0) && g.m[2] == cplx(1, 0)) {
    o.kind = PK_CX; o.b[0] = g.cb[0]; o.b[1] = g.tb[0];
    mop_finish(o);
    return true;
  }
  return false;
}

static void matmul2(const cplx* A, const cplx* B, cplx* C) {
  cplx t[4] = {A[0] * B[0] + A[1] * B[2], A[0] * B[1] + A[1] * B[3], A[2] * B[0] + A[3] * B[2], A[2] * B[1] + A[3] * B[3]};
  for (int i = 0; i < 4; ++i) C[i] = t[i];
}

// append `o` to a block's micro-op list, merging it into the last op on its bit(s) when that keeps the structure
static void mop_append(std::vector<MOp>& prog, const MOp& o) {
  const bool one = o.kind <= PK_PHASE;
  for (int i = (int)prog.size() - 1; i >= 0; --i) {
    MOp& q = prog[i];
    const bool q_one = q.kind <= PK_PHASE;
    bool touches = false;
    for (int a = 0; a < (one ? 1 : 2); ++a)
      for (int b = 0; b < (q_one ? 1 : 2); ++b)
        if (o.b[a] == q.b[b]) touches = true;
    if (!touches) continue;
    if (one && q_one) {
      cplx A[4], B[4], C[4];
      mop_to_2x2(o, A); mop_to_2x2(q, B);
      matmul2(A, B, C);
      MOp r;
      if (mop_from_2x2(C, o.b[0], r)) q = r; else prog.erase(prog.begin() + i);
      return;
    }
    if (o.kind == PK_CPHASE && q.kind == PK_CPHASE && ((o.b[0] == q.b[0] && o.b[1] == q.b[1]) || (o.b[0] == q.b[1] && o.b[1] == q.b[0]))) {
      cplx d = cplx(o.c[0], o.c[1]) * cplx(q.c[0], q.c[1]);
      q.c[0] = d.real(); q.c[1] = d.imag();
      return;
    }
    break;
  }
  prog.push_back(o);
}

// the bit an op acts on non-diagonally (it must be one of the program's positions), or -1
static int mop_acting_bit(const MOp& o) {
  if (o.kind == PK_GEN || o.kind == PK_REAL || o.kind == PK_RXL) return o.b[0];
  if (o.kind == PK_CX) return o.b[1];
  if (o.kind == PK_DIAG && o.c[0] == 0.0 && o.c[1] == 0.0) return o.b[0];  // d0 = 0 cannot be written as d0 * phase
  return -1;
}

struct Block {
  std::vector<MOp> prog;  // structured form (valid when sok)
  std::vector<int> abits; // bits the structured form acts on non-diagonally: they must be program positions (hence tile bits)
  bool sok = false;       // the structured form exists and is cheaper than the dense block
  double scost = 0.0, sflops = 0.0;
  int nb;            // number of bits (1 or 2 for fusable blocks; >2 => opaque)
  int bits[2];       // physical bits; matrix index bit t <-> bits[t]
  cplx m[16];        // row-major dense (1<<nb)^2
  bool opaque;       // not fusable (>= 3 bits): executed by the direct kernels
  GateDesc desc;     // opaque: original; fusable: filled by finalize()
  int ngates;        // original gates folded into this block
  std::vector<int> touch;  // all physical bits the block touches (ordering)
};

// dense matrix of a canonical GateDesc over the ordered bit list `bits` (targets + controls), nb <= 2
bool desc_to_dense(const GateDesc& g, int* nb, int* bits, cplx* m) {
  int n = g.k + g.nc;
  if (n > 2 || n == 0) return false;
  *nb = n;
  for (int i = 0; i < g.k; ++i) bits[i] = g.tb[i];
  for (int i = 0; i < g.nc; ++i) bits[g.k + i] = g.cb[i];
  int D = 1 << n, Dk = 1 << g.k;
  uint32_t cm = ((1u << g.nc) - 1u) << g.k;
  for (int r = 0; r < D; ++r)
    for (int c = 0; c < D; ++c) {
      cplx v;
      bool rc = (r & cm) == cm, cc = (c & cm) == cm;
      if (rc && cc) {
        int rt = r & (Dk - 1), ct = c & (Dk - 1);
        v = g.diag ? (rt == ct ? g.m[rt] : cplx(0, 0)) : g.m[rt * Dk + ct];
      } else {
        v = (r == c) ? cplx(1, 0) : cplx(0, 0);
      }
      m[r * D + c] = v;
    }
  return true;
}

void matmul(int D, const cplx* A, const cplx* B, cplx* Cout) {
  cplx tmp[16];
  for (int r = 0; r < D; ++r)
    for (int c = 0; c < D; ++c) {
      cplx s(0, 0);
      for (int k = 0; k < D; ++k) s += A[r * D + k] * B[k * D + c];
      tmp[r * D + c] = s;
    }
  for (int i = 0; i < D * D; ++i) Cout[i] = tmp[i];
}

// embed a 1-bit matrix u acting on matrix index bit `pos` of a 2-bit block
void embed1(const cplx* u, int pos, cplx* out) {
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      int ro = (r >> (1 - pos)) & 1, co = (c >> (1 - pos)) & 1;  // the other bit
      int rp = (r >> pos) & 1, cp = (c >> pos) & 1;
      out[r * 4 + c] = (ro == co) ? u[rp * 2 + cp] : cplx(0, 0);
    }
}

// reorder a 2-bit matrix given for bit order (b1,b0) into order (b0,b1)
void swap_bits(cplx* m) {
  static const int p[4] = {0, 2, 1, 3};
  cplx t[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) t[r * 4 + c] = m[p[r] * 4 + p[c]];
  for (int i = 0; i < 16; ++i) m[i] = t[i];
}

}  // namespace

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// fixed low bits of every tile: 3 (one 128-byte row of the tensor copy) with the look-ahead scheduler -- the 9 free bits buy deeper
// passes (C2: 66 instead of 93 launches, 165 instead of 184 ms) --, 5 with first fit (round-1 behaviour); BT_TILE_LOWB overrides
static int tile_lowb() {
  const int dflt = env_int("BT_FUSE_SCHED", 1) != 0 ? TILE_LOWB_MIN : TILE_LOWB;
  return std::max(TILE_LOWB_MIN, std::min(TILE_LOWB, env_int("BT_TILE_LOWB", dflt)));
}

static void fuse_blocks(const std::vector<GateDesc>& gates, std::vector<Block>& blocks) {
  std::vector<int> last(64, -1);  // last block index touching a physical bit
  const bool use_prog = env_int("BT_TILE_PROGS", 1) != 0;
  for (size_t gi = 0; gi < gates.size(); ++gi) {
    const GateDesc& g = gates[gi];
    int nb, bits[2];
    cplx m[16];
    if (g.k == 0 && g.nc == 0) {
      // pure scalar (global factor) -- keep it: non-unitary scalars matter (Kraus), unit ones are skipped at launch
      Block b; b.nb = 0; b.opaque = true; b.desc = g; b.ngates = 1;
      blocks.push_back(b);
      continue;
    }
    if (!desc_to_dense(g, &nb, bits, m)) {
      Block b; b.nb = g.k + g.nc; b.opaque = true; b.desc = g; b.ngates = 1;
      for (int i = 0; i < g.k; ++i) b.touch.push_back(g.tb[i]);
      for (int i = 0; i < g.nc; ++i) b.touch.push_back(g.cb[i]);
      blocks.push_back(b);
      for (int t : b.touch) last[t] = (int)blocks.size() - 1;
      continue;
    }
    MOp mop;
    bool mnone = false;
    const bool mok = mop_classify(g, mop, &mnone);
    if (nb == 1) {
      int lb = last[bits[0]];
      if (lb >= 0 && !blocks[lb].opaque) {
        Block& B = blocks[lb];
        if (B.nb == 1) { matmul(2, m, B.m, B.m); }
        else {
          cplx e[16];
          embed1(m, B.bits[0] == bits[0] ? 0 : 1, e);
          matmul(4, e, B.m, B.m);
        }
        B.ngates++;
        if (B.sok && mok) { if (!mnone) mop_append(B.prog, mop); } else B.sok = false;
        continue;
      }
      Block b; b.nb = 1; b.bits[0] = bits[0]; b.opaque = false; b.ngates = 1;
      b.sok = mok;
      if (mok && !mnone) b.prog.push_back(mop);
      for (int i = 0; i < 4; ++i) b.m[i] = m[i];
      b.touch.push_back(bits[0]);
      blocks.push_back(b);
      last[bits[0]] = (int)blocks.size() - 1;
      continue;
    }
    // nb == 2
    int l0 = last[bits[0]], l1 = last[bits[1]];
    if (l0 >= 0 && l0 == l1 && !blocks[l0].opaque && blocks[l0].nb == 2) {
      Block& B = blocks[l0];
      if (B.bits[0] != bits[0]) swap_bits(m);
      matmul(4, m, B.m, B.m);
      B.ngates++;
      if (B.sok && mok) { if (!mnone) mop_append(B.prog, mop); } else B.sok = false;
      continue;
    }
    Block b; b.nb = 2; b.bits[0] = bits[0]; b.bits[1] = bits[1]; b.opaque = false; b.ngates = 1;
    b.sok = mok;
    if (mok && !mnone) b.prog.push_back(mop);
    for (int i = 0; i < 16; ++i) b.m[i] = m[i];
    // absorb pending pure 1-bit blocks that are the last thing on either bit
    for (int t = 0; t < 2; ++t) {
      int lb = last[bits[t]];
      if (lb >= 0 && !blocks[lb].opaque && blocks[lb].nb == 1) {
        cplx e[16];
        embed1(blocks[lb].m, t, e);
        matmul(4, b.m, e, b.m);
        b.ngates += blocks[lb].ngates;
        if (b.sok && blocks[lb].sok) b.prog.insert(b.prog.begin(), blocks[lb].prog.begin(), blocks[lb].prog.end());
        else b.sok = false;
        blocks[lb].nb = -1;  // tombstone
      }
    }
    b.touch.push_back(bits[0]);
    b.touch.push_back(bits[1]);
    blocks.push_back(b);
    last[bits[0]] = last[bits[1]] = (int)blocks.size() - 1;
  }
  // finalize
  std::vector<Block> out;
  out.reserve(blocks.size());
  for (Block& b : blocks) {
    if (b.nb == -1) continue;
    if (!b.opaque) {
      bt_canonicalize(b.nb, b.bits, b.m, 0, nullptr, &b.desc);
      b.scost = b.sflops = 0.0;
      for (const MOp& o : b.prog) { b.scost += o.cost; b.sflops += o.flops; }
      // keep the structured form only when it is cheaper than the dense block and the dense form is not already diagonal
      // (a diagonal block always does: its alternative is a shared-memory sweep + barrier of its own)
      // With the pass specialiser on (bt_jit.cu) a pass made of structured blocks only becomes a straight-line kernel (1.6 instead
      // of 2.6 ms at 28 qubits) in which real / RX-like ops cost about half of the interpreter's table: the structured form is
      // kept up to BT_FUSE_STRUCT_SLACK_PCT (150) per cent of the dense cost, so the 4-op blocks at the head of a circuit
      // (1-qubit gates on both sides of the first CX) and lone general 1-qubit blocks no longer force an interpreted pass.
      const double slack = env_int("BT_TILE_JIT", 1) != 0 ? 0.01 * (double)std::max(100, env_int("BT_FUSE_STRUCT_SLACK_PCT", 150)) : 1.0;
      const double dense_cost = (b.desc.k == 2 ? 16.0 : 8.0) * slack;
      if (!use_prog || b.prog.empty() || (!b.desc.diag && b.scost >= dense_cost) || (int)b.prog.size() > 12) b.sok = false;
      if (b.sok)
        for (const MOp& o : b.prog) {
          int a = mop_acting_bit(o);
          if (a >= 0 && std::find(b.abits.begin(), b.abits.end(), a) == b.abits.end()) b.abits.push_back(a);
        }
    }
    out.push_back(b);
  }
  blocks.swap(out);
}

static std::atomic<uint64_t> g_fused_passes{0}, g_fused_blocks{0};
static std::atomic<double> g_fused_flops{0.0};  // FP64 flops issued by the fused passes (FMA = 2)

// bits of `d` that must be inside the tile: non-diagonal targets
static void needed_bits(const GateDesc& d, std::vector<int>& out) {
  out.clear();
  if (!d.diag)
    for (int i = 0; i < d.k; ++i) out.push_back(d.tb[i]);
}

// Group enumeration for an item whose fixed tile positions are `fixed` (targets + local controls): group-index bit k walks
// free tile position order[k].  The first three are chosen with distinct residues mod 3 so that 8 consecutive lanes hit 8
// distinct bank groups under sw(); thread bits come first, loop-iteration bits after.
static int build_group_walk(int T, uint32_t fixed_mask, int nfixed, uint32_t* bit_sw, uint32_t* iter_sw, int iter_cap, uint32_t* niter_out, int mode,
                            uint32_t* bit_lin = nullptr, uint32_t* iter_lin = nullptr) {
  int order[TILE_TMAX], n = 0;
  bool used[TILE_TMAX] = {false};
  bool res_taken[3] = {false, false, false};
  const int plim = mode ? T : std::min(T, 6);  // mode 0: only positions below 6 reach the bank-group bits
  for (int round = 0; round < 3; ++round)
    for (int p = 0; p < plim; ++p)
      if (!((fixed_mask >> p) & 1u) && !used[p] && !res_taken[p % 3]) { order[n++] = p; used[p] = true; res_taken[p % 3] = true; break; }
  for (int p = 0; p < T; ++p)
    if (!((fixed_mask >> p) & 1u) && !used[p]) { order[n++] = p; used[p] = true; }
  if (n != T - nfixed) return -1;
  int tbits = 0;
  while ((1 << tbits) < TILE_THREADS) ++tbits;
  for (int k = 0; k < 8; ++k) bit_sw[k] = (k < n && k < tbits) ? swz(1u << order[k], mode) : 0u;
  if (bit_lin) for (int k = 0; k < 8; ++k) bit_lin[k] = (k < n && k < tbits) ? (1u << order[k]) : 0u;
  uint32_t ngroups = 1u << n;
  uint32_t niter = ngroups > (uint32_t)TILE_THREADS ? ngroups / TILE_THREADS : 1u;
  if ((int)niter > iter_cap) return -1;
  for (uint32_t it = 0; it < niter; ++it) {
    uint32_t c = 0;
    for (int k = tbits; k < n; ++k)
      if ((it >> (k - tbits)) & 1u) c |= 1u << order[k];
    iter_sw[it] = swz(c, mode);
    if (iter_lin) iter_lin[it] = c;
  }
  *niter_out = niter;
  return 0;
}

static inline bool cluster_eligible(const GateDesc& d) { return !d.diag && d.k == 2 && d.nc == 0; }

static int fill_gate_slot(TileGate& G, const GateDesc& d, const int* local_pos, int T, int mode) {
  G.kind = d.diag ? 1 : 0;
  G.k = d.k;
  G.lcmask = 0; G.ext_cmask = 0;
  std::vector<int> ins;
  for (int t = 0; t < d.k; ++t) {
    int lp = local_pos[d.tb[t]];
    G.tloc[t] = lp; G.text[t] = d.tb[t];
    if (lp >= 0) ins.push_back(lp);
    else if (!d.diag) BT_FAIL(BT_ERR_ARG, "internal: dense target outside the tile");
  }
  for (int c = 0; c < d.nc; ++c) {
    int lp = local_pos[d.cb[c]];
    if (lp >= 0) { G.lcmask |= 1u << lp; ins.push_back(lp); }
    else G.ext_cmask |= 1ull << d.cb[c];
  }
  if (ins.size() > 6) BT_FAIL(BT_ERR_ARG, "internal: too many local bits in one gate");
  G.ni = (int)ins.size();
  uint32_t fixed = 0;
  for (int lp : ins) fixed |= 1u << lp;
  int cntm = d.diag ? (1 << d.k) : (1 << (2 * d.k));
  for (int i = 0; i < cntm; ++i) G.m[i] = make_double2(d.m[i].real(), d.m[i].imag());
  if (build_group_walk(T, fixed, G.ni, G.bit_sw, G.iter_sw, 32, &G.niter, mode) != 0) BT_FAIL(BT_ERR_ARG, "internal: tile gate loop too long");
  return BT_OK;
}

// diagonal gate -> run-time-indexed slot; returns 1 if it does not fit a TileDiag (then it takes a specialised gate slot)
static int fill_diag_slot(TileDiag& G, const GateDesc& d, const int* local_pos, int T, int mode) {
  if (!d.diag || d.k > 2) return 1;
  G.k = d.k;
  G.lcmask = 0; G.ext_cmask = 0;
  uint32_t fixed = 0;
  int ni = 0;
  for (int t = 0; t < 2; ++t) { G.tloc[t] = -1; G.text[t] = 0; }
  for (int t = 0; t < d.k; ++t) {
    int lp = local_pos[d.tb[t]];
    G.tloc[t] = lp; G.text[t] = d.tb[t];
    if (lp >= 0) { fixed |= 1u << lp; ni++; }
  }
  for (int c = 0; c < d.nc; ++c) {
    int lp = local_pos[d.cb[c]];
    if (lp >= 0) { G.lcmask |= 1u << lp; fixed |= 1u << lp; ni++; }
    else G.ext_cmask |= 1ull << d.cb[c];
  }
  G.ni = ni;
  for (int i = 0; i < (1 << d.k); ++i) G.m[i] = make_double2(d.m[i].real(), d.m[i].imag());
  if (build_group_walk(T, fixed, ni, G.bit_sw, G.iter_sw, 16, &G.niter, mode) != 0) return 1;
  return 0;
}

#if CL_BITS == 4
static const int CL_PAIR[5][2] = {{0, 1}, {2, 3}, {1, 2}, {0, 1}, {2, 3}};
static const int CL_SEED2 = 2;
#else
static const int CL_PAIR[5][2] = {{0, 1}, {1, 2}, {0, 1}, {0, 1}, {0, 1}};
static const int CL_SEED2 = 1;
#endif

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (PFN_encodeTiled)p;
    else cudaGetLastError();
  }
  return fn;
}

// Describe the tile {in[b]} as a box of a <= 5-D tensor over the local amplitudes.  Returns false when the bit set needs more
// than four runs (the cp.async kernel handles those).
static bool build_tensor_map(bt_sv* s, const bool* in, int T, CUtensorMap* map, TileParams& P) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc || T < 8) return false;
  for (int b = 0; b < 3; ++b) if (!in[b]) return false;
  int total_bits = 0;
  while ((1ull << total_bits) < s->len) ++total_bits;
  if ((1ull << total_bits) != s->len) return false;  // ragged batches stay on the cp.async path
  // runs of consecutive tile bits from bit 3 up, each at most 8 bits long
  int run_start[8], run_len[8], nr = 0;
  for (int b = 3; b < s->n_local;) {
    if (!in[b]) { ++b; continue; }
    int e = b;
    while (e < s->n_local && in[e] && e - b < 8) ++e;
    if (nr >= 7) return false;
    run_start[nr] = b; run_len[nr] = e - b; nr++;
    b = e;
  }
  if (nr == 0) return false;
  if (run_start[0] != 3) {
    // bits 3 .. run_start[0]-1 are not tile bits: they take a dimension of their own with a box of 1
    for (int k = nr; k > 0; --k) { run_start[k] = run_start[k - 1]; run_len[k] = run_len[k - 1]; }
    run_start[0] = 3; run_len[0] = 0; nr++;
  }
  // more than four runs: the tile bits of the runs above the fourth are enumerated -- one box per value (<= 16 boxes of >= 4 KB)
  P.tma_ncopy = 1;
  P.tma_c4add[0] = 0;
  if (nr > 4) {
    int ebits[8], ne = 0;  // at most 4 used
    for (int k = 4; k < nr; ++k)
      for (int j = 0; j < run_len[k]; ++j) {
        if (ne >= 4 || T - ne < 6) return false;  // every box stays a multiple of the 1 KB swizzle pattern
        ebits[ne++] = run_start[k] + j;
      }
    P.tma_ncopy = 1 << ne;
    for (int e = 0; e < (1 << ne); ++e) {
      int64_t add = 0;
      for (int j = 0; j < ne; ++j) if ((e >> j) & 1) add += (int64_t)1 << (ebits[j] - run_start[3]);
      if (add > 0x7fffffff) return false;
      P.tma_c4add[e] = (int32_t)add;
    }
    nr = 4;
  }
  cuuint64_t gdim[5], gstride[4];
  cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  gdim[0] = 16; box[0] = 16;  // 8 amplitudes = 16 doubles = 128 B
  for (int k = 0; k < 5; ++k) { P.tma_coord_shift[k] = 0; P.tma_coord_mask[k] = 0; }
  for (int k = 1; k <= 4; ++k) {
    if (k <= nr) {
      int st = run_start[k - 1];
      int en = (k < nr) ? run_start[k] : total_bits;
      if (en - st > 32) return false;
      gdim[k] = 1ull << (en - st);
      gstride[k - 1] = (cuuint64_t)16 << st;
      box[k] = 1u << run_len[k - 1];
      P.tma_coord_shift[k] = st;
      P.tma_coord_mask[k] = (en - st >= 32) ? 0xffffffffu : ((1u << (en - st)) - 1u);
      if (gdim[k] > (1ull << 31)) return false;  // coordinates are signed 32-bit
    } else {
      gdim[k] = 1; box[k] = 1;
      gstride[k - 1] = (cuuint64_t)16 << total_bits;  // never stepped
    }
  }
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, (void*)s->amp, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

int bt_jit_try_launch(bt_sv* s, const TileParams& P, const CUtensorMap& tmap, uint64_t ntiles, size_t tile_bytes, int np);  // bt_jit.cu
bool bt_jit_source_for(const TileParams& P, int np, std::string& src);  // bt_jit.cu

// ---- BT_JIT_VERIFY: max |amp - alt|^2 and max |alt|^2 over the state (positive doubles order like their bit patterns) ----------------
__global__ void __launch_bounds__(256) k_verify_diff(const double2* __restrict__ a, const double2* __restrict__ b, uint64_t len, unsigned long long* __restrict__ out) {
  double md = 0.0, mn = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (uint64_t)gridDim.x * blockDim.x) {
    const double2 x = a[i], y = b[i];
    const double dx = x.x - y.x, dy = x.y - y.y;
    const double d = dx * dx + dy * dy;
    md = (d > md || d != d) ? (d != d ? 1e300 : d) : md;  // NaN counts as a disagreement
    mn = fmax(mn, y.x * y.x + y.y * y.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { md = fmax(md, __shfl_xor_sync(0xffffffffu, md, o)); mn = fmax(mn, __shfl_xor_sync(0xffffffffu, mn, o)); }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out, (unsigned long long)__double_as_longlong(md));
    atomicMax(out + 1, (unsigned long long)__double_as_longlong(mn));
  }
}

static std::atomic<uint64_t> g_verify_checked{0}, g_verify_failed{0};

// amp = specialised result, alt = interpreter result of the same pass
static int verify_compare(bt_sv* s) {
  BT_TRY(bt_ensure_scratch(s, 16));
  unsigned long long* d = (unsigned long long*)s->d_scratch;
  BT_CUDA(cudaMemsetAsync(d, 0, 16, s->stream));
  k_verify_diff<<<148 * 8, 256, 0, s->stream>>>(s->amp, s->alt, s->len, d);
  BT_CUDA(cudaPeekAtLastError());
  unsigned long long h[2];
  BT_CUDA(cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, s->stream));
  BT_CUDA(cudaStreamSynchronize(s->stream));
  double md, mn;
  memcpy(&md, &h[0], 8); memcpy(&mn, &h[1], 8);
  g_verify_checked++;
  if (sqrt(md) > 1e-12 * std::max(1e-30, sqrt(mn)) + 1e-300) {
    g_verify_failed++;
    fprintf(stderr, "[bluetangle_cuda] BT_JIT_VERIFY: a specialised pass differs from the interpreter (max |diff| %.3e, max |amp| %.3e): keeping the interpreter's result\n", sqrt(md),
            sqrt(mn));
    BT_CUDA(cudaMemcpyAsync(s->amp, s->alt, s->len * sizeof(double2), cudaMemcpyDeviceToDevice, s->stream));
  }
  return BT_OK;
}

extern "C" int bt_jit_verify_stats(uint64_t* checked, uint64_t* failed) {
  if (checked) *checked = g_verify_checked.load();
  if (failed) *failed = g_verify_failed.load();
  return BT_OK;
}

// dry run (host-only planning, bt_fusion_plan): launch_pass forms slots and programs exactly as for a real launch but touches no
// device; t_dry[0] counts kernel launches, [1] register programs, [2] other items, [3] single-gate passes
static thread_local int* t_dry = nullptr;

static int launch_pass(bt_sv* s, const std::vector<const Block*>& pass_in, const std::vector<int>& tile_bits_in) {
  const bool dry = t_dry != nullptr;
  if (pass_in.size() == 1) {
    if (dry) { t_dry[0]++; t_dry[3]++; return BT_OK; }
    return bt_launch_gate(s, pass_in[0]->desc);
  }
  int T = std::min(s->n_local, env_int("BT_TILE_BITS", TILE_TDEF));
  int lowb = std::min(tile_lowb(), T);
  const bool use_clusters = env_int("BT_TILE_CLUSTERS", 1) != 0 && T >= CL_BITS;
  // tile bits: low bits + requested + padding with the lowest free bits
  bool in[64] = {false};
  for (int j = 0; j < lowb; ++j) in[j] = true;
  int cnt = lowb;
  for (int b : tile_bits_in)
    if (!in[b]) { in[b] = true; cnt++; }
  for (int b = lowb; b < s->n_local && cnt < T; ++b)
    if (!in[b]) { in[b] = true; cnt++; }
  if (cnt != T) BT_FAIL(BT_ERR_ARG, "internal: tile has %d bits, expected %d", cnt, T);
  TileParams P;
  memset(&P, 0, sizeof(P));
  P.T = T; P.lowb = lowb;
  alignas(64) CUtensorMap tmap;
  const bool use_tma = dry || (env_int("BT_TILE_TMA", 1) != 0 && build_tensor_map(s, in, T, &tmap, P));
  P.swz_mode = use_tma ? 0 : 1;
  if (dry) P.tma_ncopy = 1;
  int local_pos[64];
  for (int b = 0; b < 64; ++b) local_pos[b] = -1;
  int j = 0;
  for (int b = 0; b < s->n_local; ++b)
    if (in[b]) { P.tbits[j] = b; local_pos[b] = j; j++; }
  // drop identity diagonals
  std::vector<const Block*> pass;
  for (const Block* blk : pass_in) {
    const GateDesc& d = blk->desc;
    if (d.k > 2 || d.nc > 4) BT_FAIL(BT_ERR_ARG, "internal: gate not tileable");
    if (d.diag) {
      bool all_one = true;
      for (int i = 0; i < (1 << d.k); ++i) if (d.m[i] != cplx(1, 0)) all_one = false;
      if (all_one) continue;
    }
    pass.push_back(blk);
  }
  const size_t n = pass.size();
  std::vector<char> used(n, 0);
  int ng = 0, nc = 0, nd = 0, np = 0, nitems = 0;
  uint64_t ntiles = s->len >> T;
  size_t smem = sizeof(double2) << T;
  static bool attr_set[64] = {false};  // the opt-in shared-memory size is a per-device function attribute
  if (!dry && !attr_set[s->device & 63]) {
    BT_CUDA(cudaFuncSetAttribute(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double2) << TILE_TMAX)));
    BT_CUDA(cudaFuncSetAttribute(k_tile_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((sizeof(double2) << TILE_TMAX) + 1024 + 64)));
    BT_CUDA(cudaFuncSetAttribute(k_tile_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((sizeof(double2) << TILE_TMAX) + 1024 + 64)));
    BT_CUDA(cudaFuncSetAttribute(k_tile_pipe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((sizeof(double2) << TILE_TMAX) + 1024 + 64)));
    BT_CUDA(cudaFuncSetAttribute(k_tile_pipe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((sizeof(double2) << TILE_TMAX) + 1024 + 64)));
    attr_set[s->device & 63] = true;
  }
  int nsm = 148;
  if (!dry) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, s->device);
  auto flush = [&]() -> int {
    if (nitems == 0) return BT_OK;
    if (dry) {
      if (env_int("BT_JIT_DUMP", 0) && ng == 0 && nc == 0 && nd == 0 && np > 0) {
        P.nitems = nitems;
        std::string src;
        if (bt_jit_source_for(P, np, src)) fprintf(stderr, "// ===== pass %d =====\n%s\n", t_dry[0], src.c_str());
      }
      t_dry[0]++; t_dry[1] += np; t_dry[2] += ng + nc + nd;
      ng = nc = nd = np = nitems = 0;
      return BT_OK;
    }
      P.nitems = nitems;
    P.stagger_ns = env_int("BT_TILE_STAGGER_NS", 0);
    P.n_sm = nsm;
    if (env_int("BT_TILE_DEBUG", 0)) {
      if (!use_tma) { fprintf(stderr, "[tile] no tensor map for bits:"); for (int b = 0; b < 64; ++b) if (in[b]) fprintf(stderr, " %d", b); fprintf(stderr, "\n"); }
      fprintf(stderr, "[tile] T=%d tma=%d items=%d gates=%d clusters=%d diags=%d progs=%d ops:", T, (int)use_tma, nitems, ng, nc, nd, np);
      for (int q = 0; q < np; ++q) fprintf(stderr, " %u(%d,%d,%d,%d)", P.pr[q].nops, P.pr[q].lp[0], P.pr[q].lp[1], P.pr[q].lp[2], P.pr[q].lp[3]);
      fprintf(stderr, "\n");
    }
    // BT_JIT_VERIFY=1: every specialised launch is cross-checked against the interpreter kernel on a copy of the state (results
    // of a specialised pass must never depend on the run-time compiler); a disagreement is counted, reported on stderr, and the
    // interpreter's result is kept.  bt_jit_verify_stats() returns the counts.
    const bool jit_eligible = use_tma && ng == 0 && nc == 0 && nd == 0 && np > 0;
    const bool verify = jit_eligible && env_int("BT_JIT_VERIFY", 0) != 0;
    if (verify) {
      BT_TRY(bt_ensure_alt(s));
      BT_CUDA(cudaMemcpyAsync(s->alt, s->amp, s->len * sizeof(double2), cudaMemcpyDeviceToDevice, s->stream));
    }
    bt_prof_begin(s, BT_CLS_TILE);
    if (jit_eligible && bt_jit_try_launch(s, P, tmap, ntiles, smem, np) == 1) {
      // launched as a specialised straight-line kernel (bt_jit.cu)
      if (verify) {
        std::swap(s->amp, s->alt);  // the interpreter runs on the copy
        alignas(64) CUtensorMap tmap2;
        TileParams P2 = P;
        const bool ok2 = build_tensor_map(s, in, T, &tmap2, P2);
        if (ok2) k_tile_tma<false><<<(unsigned)ntiles, TILE_THREADS, smem + 1024 + 64, s->stream>>>(tmap2, P2);
        std::swap(s->amp, s->alt);
        if (ok2) BT_TRY(verify_compare(s));
      }
    } else if (use_tma) {
      const bool lite = ng == 0 && nc == 0 && env_int("BT_TILE_LITE", 1);
      int K = env_int("BT_TILE_PIPE", 0);
      while (K > 1 && (ntiles % (uint64_t)K != 0 || ntiles / (uint64_t)K < (uint64_t)(6 * nsm))) K >>= 1;
      if (K > 1 && T <= 11 && (K & (K - 1)) == 0) {
        if (lite) k_tile_pipe<false><<<(unsigned)(ntiles / K), TILE_THREADS, 2 * smem + 1024 + 64, s->stream>>>(tmap, P, K);
        else k_tile_pipe<true><<<(unsigned)(ntiles / K), TILE_THREADS, 2 * smem + 1024 + 64, s->stream>>>(tmap, P, K);
      } else if (lite) k_tile_tma<false><<<(unsigned)ntiles, TILE_THREADS, smem + 1024 + 64, s->stream>>>(tmap, P);
      else k_tile_tma<true><<<(unsigned)ntiles, TILE_THREADS, smem + 1024 + 64, s->stream>>>(tmap, P);
    } else {
      k_tile<<<(unsigned)ntiles, TILE_THREADS, smem, s->stream>>>(s->amp, P);
    }
    bt_prof_end(s);
    BT_CHECK_LAUNCH(s);
    g_fused_passes++;
    ng = nc = nd = np = nitems = 0;
    return BT_OK;
  };

  // a block can run inside a register program when it kept its structured form and the bits it acts on are tile bits
  const bool use_progs = env_int("BT_TILE_PROGS", 1) != 0 && T >= PROG_BITS;
  auto prog_eligible = [&](const Block* b) -> bool {
    if (!use_progs || !b->sok || b->opaque) return false;
    for (int t : b->abits) if (local_pos[t] < 0) return false;
    return true;
  };

  struct Try {
    int nm;
    uint32_t use;
    int bit_at_pos[4];
    int slot_of_member[CL_SLOTS];
    size_t member[CL_SLOTS];
    bool swapped[CL_SLOTS];
  };
  // greedy cluster formation over the 5-slot pattern, seeded at (seed_slot, seed_flip); gates that cannot join keep their
  // place (after the cluster), so a later gate may only be pulled in if it shares no bit with any skipped gate
  auto try_cluster = [&](size_t i, int seed_slot, int seed_flip, Try& R) {
    int pos_of_bit[64];
    for (int b = 0; b < 64; ++b) pos_of_bit[b] = -1;
    for (int p = 0; p < 4; ++p) R.bit_at_pos[p] = -1;
    R.use = 0; R.nm = 0;
    int last_slot_on_pos[4] = {-1, -1, -1, -1};
    bool blocked[64] = {false};
    for (size_t jj = i; jj < n && jj < i + 24 && R.nm < CL_SLOTS; ++jj) {
      if (used[jj]) continue;
      const Block* hb = pass[jj];
      const GateDesc& h = hb->desc;
      bool dep = false;
      for (int t : hb->touch) if (blocked[t]) dep = true;
      bool placed = false;
      // (a structured block only asked for the bits it acts on: its dense form may have a target outside the tile)
      if (!dep && cluster_eligible(h) && !prog_eligible(hb) && local_pos[h.tb[0]] >= 0 && local_pos[h.tb[1]] >= 0) {
        int a = h.tb[0], b = h.tb[1];  // matrix bit 0 <-> a, bit 1 <-> b
        for (int sl = (jj == i ? seed_slot : 0); sl < CL_SLOTS && !placed; ++sl) {
          if (R.use & (1u << sl)) continue;
          int p = CL_PAIR[sl][0], q = CL_PAIR[sl][1];
          if (sl <= last_slot_on_pos[p] || sl <= last_slot_on_pos[q]) continue;  // after earlier blocks on these positions
          for (int f = 0; f < 2 && !placed; ++f) {
            int flip = (jj == i) ? (seed_flip ^ f) : f;
            int pa = flip ? q : p, pb = flip ? p : q;  // a -> pa, b -> pb
            bool oka = (pos_of_bit[a] == pa) || (pos_of_bit[a] == -1 && R.bit_at_pos[pa] == -1);
            bool okb = (pos_of_bit[b] == pb) || (pos_of_bit[b] == -1 && R.bit_at_pos[pb] == -1);
            if (!oka || !okb) continue;
            pos_of_bit[a] = pa; R.bit_at_pos[pa] = a;
            pos_of_bit[b] = pb; R.bit_at_pos[pb] = b;
            R.use |= 1u << sl;
            last_slot_on_pos[p] = sl; last_slot_on_pos[q] = sl;
            R.slot_of_member[R.nm] = sl; R.member[R.nm] = jj; R.swapped[R.nm] = (pa > pb);
            R.nm++;
            placed = true;
          }
          if (jj == i && !placed) break;  // the seed is only tried at its seed slot
        }
      }
      if (jj == i && !placed) return;
      if (!placed)
        for (int t : hb->touch) blocked[t] = true;
    }
  };

  // ops / coefficients a block needs once its ops are resolved against a program (upper bounds: an out-of-program DIAG splits)
  auto prog_demand = [&](const Block* b, int* nops, int* ncoef) {
    *nops = 0; *ncoef = 0;
    for (const MOp& o : b->prog) {
      if (o.kind == PK_DIAG) { *nops += 2; *ncoef += 8; }
      else { *nops += 1; *ncoef += std::max(o.ncoef, 4); }
    }
  };

  for (size_t i = 0; i < n; ++i) {
    if (used[i]) continue;
    const GateDesc& d0 = pass[i]->desc;
    if (prog_eligible(pass[i])) {
      // greedy register program seeded here: later blocks join while the union of ACTING bits stays within PROG_BITS and no
      // skipped block shares a bit with them (order preserved); control / phase bits may lie anywhere
      int sbits[PROG_BITS], nsb = 0, nops = 0, ncoef = 0;
      std::vector<size_t> members;
      bool blocked[64] = {false};
      for (size_t jj = i; jj < n && jj < i + 64; ++jj) {
        if (used[jj]) continue;
        const Block* hb = pass[jj];
        bool dep = false;
        for (int t : hb->touch) if (blocked[t]) dep = true;
        bool fits = !dep && prog_eligible(hb);
        if (fits) {
          int extra = 0;
          for (int t : hb->abits) {
            bool have = false;
            for (int q = 0; q < nsb; ++q) if (sbits[q] == t) have = true;
            if (!have) extra++;
          }
          int ho, hc;
          prog_demand(hb, &ho, &hc);
          if (nsb + extra > PROG_BITS || nops + ho > PROG_MAXOPS || ncoef + hc > PROG_MAXCOEF) fits = false;
          if (fits) {
            for (int t : hb->abits) {
              bool have = false;
              for (int q = 0; q < nsb; ++q) if (sbits[q] == t) have = true;
              if (!have) sbits[nsb++] = t;
            }
            nops += ho; ncoef += hc;
            members.push_back(jj);
          }
        }
        if (!fits) {
          if (hb->touch.empty()) break;
          for (int t : hb->touch) blocked[t] = true;
        }
      }
      if (!members.empty()) {
        if (np >= TILE_MAXP) BT_TRY(flush());
        TileProg& G = P.pr[np];
        // spare positions go to the tile bits the members' diagonal ops use most (their ops become unconditional), then to
        // free tile bits from the top
        {
          int freq[64] = {0};
          for (size_t mj : members)
            for (const MOp& o : pass[mj]->prog)
              for (int e = 0; e < 2; ++e) {
                int t = o.b[e];
                if (t >= 0 && local_pos[t] >= 0) freq[t]++;
              }
          while (nsb < PROG_BITS) {
            int bestb = -1;
            for (int t = 0; t < 64; ++t) {
              if (freq[t] == 0) continue;
              bool have = false;
              for (int q = 0; q < nsb; ++q) if (sbits[q] == t) have = true;
              if (!have && (bestb < 0 || freq[t] > freq[bestb])) bestb = t;
            }
            if (bestb < 0) break;
            sbits[nsb++] = bestb;
          }
        }
        std::sort(sbits, sbits + nsb, [&](int a, int b) { return local_pos[a] < local_pos[b]; });
        bool taken[32] = {false};
        int pos_of[64];
        for (int b = 0; b < 64; ++b) pos_of[b] = -1;
        for (int q = 0; q < nsb; ++q) { G.lp[q] = local_pos[sbits[q]]; taken[G.lp[q]] = true; pos_of[sbits[q]] = q; }
        int freeb = T - 1;
        for (int q = nsb; q < PROG_BITS; ++q) {
          while (freeb >= 0 && taken[freeb]) --freeb;
          if (freeb < 0) BT_FAIL(BT_ERR_ARG, "internal: no free tile bit for a register program");
          G.lp[q] = freeb; taken[freeb] = true;
        }
        uint32_t fixed = 0;
        for (int q = 0; q < PROG_BITS; ++q) fixed |= 1u << G.lp[q];
        if (build_group_walk(T, fixed, PROG_BITS, G.bit_sw, G.iter_sw, 8, &G.niter, P.swz_mode, G.bit_lin, G.iter_lin) != 0)
          BT_FAIL(BT_ERR_ARG, "internal: program loop too long");
        int ko = 0, kc = 0;
        double fl = 0.0;
        auto emit = [&](int site, const double* c, int ncf) {
          G.op[ko++] = (uint16_t)(site | (kc << 8));
          for (int e = 0; e < ncf; ++e) G.coef[kc++] = c[e];
        };
        // condition masks of a bit that is not a program position: tile-local -> lm, outside the tile -> em
        auto add_cond = [&](int bit, uint64_t* em, uint64_t* lm) {
          if (local_pos[bit] >= 0) *lm |= 1ull << local_pos[bit]; else *em |= 1ull << bit;
        };
        auto as_double = [](uint64_t v) { double d; memcpy(&d, &v, sizeof(d)); return d; };
        for (size_t mj : members) {
          for (const MOp& o : pass[mj]->prog) {
            fl += o.flops;
            const int pa = pos_of[o.b[0]], pb = (o.kind == PK_CX || o.kind == PK_CPHASE) ? pos_of[o.b[1]] : -1;
            uint64_t em = 0, lm = 0;
            if (o.kind == PK_GEN || o.kind == PK_REAL || o.kind == PK_RXL) {
              emit(PROG_SITE_U1(o.kind, pa), o.c, o.ncoef);
            } else if (o.kind == PK_PHASE) {
              if (pa >= 0) emit(PROG_SITE_U1(PK_PHASE, pa), o.c, 2);
              else { add_cond(o.b[0], &em, &lm); double c[4] = {o.c[0], o.c[1], as_double(em), as_double(lm)}; emit(PROG_SITE_CSCALE, c, 4); }
            } else if (o.kind == PK_DIAG) {
              if (pa >= 0) emit(PROG_SITE_U1(PK_DIAG, pa), o.c, 4);
              else {
                // diag(d0, d1) on a bit outside the program = d0 * diag(1, d1 / d0)
                const cplx d0(o.c[0], o.c[1]), r = cplx(o.c[2], o.c[3]) / d0;
                double c0[4] = {o.c[0], o.c[1], as_double(0), as_double(0)};
                emit(PROG_SITE_CSCALE, c0, 4);
                add_cond(o.b[0], &em, &lm);
                double c1[4] = {r.real(), r.imag(), as_double(em), as_double(lm)};
                emit(PROG_SITE_CSCALE, c1, 4);
              }
            } else if (o.kind == PK_CX) {
              if (pa >= 0) emit(PROG_SITE_CX(pa, pb), o.c, 0);
              else { add_cond(o.b[0], &em, &lm); double c[2] = {as_double(em), as_double(lm)}; emit(PROG_SITE_CCX1(pb), c, 2); }
            } else {  // PK_CPHASE
              if (pa >= 0 && pb >= 0) emit(PROG_SITE_CPHASE(std::min(pa, pb), std::max(pa, pb)), o.c, 2);
              else if (pa >= 0 || pb >= 0) {
                add_cond(pa >= 0 ? o.b[1] : o.b[0], &em, &lm);
                double c[4] = {o.c[0], o.c[1], as_double(em), as_double(lm)};
                emit(PROG_SITE_CPH1(pa >= 0 ? pa : pb), c, 4);
              } else {
                add_cond(o.b[0], &em, &lm); add_cond(o.b[1], &em, &lm);
                double c[4] = {o.c[0], o.c[1], as_double(em), as_double(lm)};
                emit(PROG_SITE_CSCALE, c, 4);
              }
            }
          }
          used[mj] = 1;
        }
        if (ko > PROG_MAXOPS || kc > PROG_MAXCOEF) BT_FAIL(BT_ERR_ARG, "internal: register program overflow");
        G.nops = (uint32_t)ko;
        G.op[ko] = (uint16_t)0xff;  // padding (fetched, never executed)
        P.item[nitems++] = (uint8_t)(TILE_PBASE + np);
        np++;
        g_fused_blocks += members.size();
        g_fused_flops.store(g_fused_flops.load() + fl * (double)s->len);
        continue;
      }
    }
    bool made_cluster = false;
    if (use_clusters && cluster_eligible(d0)) {
      Try best; best.nm = 0;
      for (int seed_slot = 0; seed_slot <= CL_SEED2; seed_slot += CL_SEED2)
        for (int seed_flip = 0; seed_flip < 2; ++seed_flip) {
          Try R;
          try_cluster(i, seed_slot, seed_flip, R);
          if (R.nm > best.nm) best = R;
        }
      if (best.nm >= 2) {
        if (nc >= TILE_MAXC) BT_TRY(flush());
        TileCluster& Cl = P.cl[nc];
        // unused positions take free tile bits from the top (keeps the low bits as lane bits)
        bool taken[32] = {false};
        for (int p = 0; p < CL_BITS; ++p) if (best.bit_at_pos[p] >= 0) taken[local_pos[best.bit_at_pos[p]]] = true;
        int freeb = T - 1;
        for (int p = 0; p < CL_BITS; ++p) {
          if (best.bit_at_pos[p] >= 0) { Cl.lp[p] = local_pos[best.bit_at_pos[p]]; continue; }
          while (freeb >= 0 && taken[freeb]) --freeb;
          if (freeb < 0) BT_FAIL(BT_ERR_ARG, "internal: no free tile bit for a cluster");
          Cl.lp[p] = freeb; taken[freeb] = true;
        }
        uint32_t fixed = 0;
        for (int p = 0; p < CL_BITS; ++p) fixed |= 1u << Cl.lp[p];
        Cl.use = best.use;
        if (build_group_walk(T, fixed, CL_BITS, Cl.bit_sw, Cl.iter_sw, 8, &Cl.niter, P.swz_mode) != 0)
          BT_FAIL(BT_ERR_ARG, "internal: cluster loop too long (T=%d fixed=0x%x lp=%d,%d,%d bits=%d,%d,%d,%d nm=%d)", T, fixed, Cl.lp[0], Cl.lp[1], Cl.lp[2],
                  best.bit_at_pos[0], best.bit_at_pos[1], best.bit_at_pos[2], best.bit_at_pos[3], best.nm);
        for (int mI = 0; mI < best.nm; ++mI) {
          const GateDesc& h = pass[best.member[mI]]->desc;
          cplx mm[16];
          for (int e = 0; e < 16; ++e) mm[e] = h.m[e];
          if (best.swapped[mI]) swap_bits(mm);  // matrix bit 0 must be the lower cluster position
          for (int e = 0; e < 16; ++e) Cl.m[best.slot_of_member[mI]][e] = make_double2(mm[e].real(), mm[e].imag());
          used[best.member[mI]] = 1;
        }
        P.item[nitems++] = (uint8_t)(TILE_MAXG + nc);
        nc++;
        g_fused_blocks += best.nm;
        g_fused_flops.store(g_fused_flops.load() + 32.0 * (double)best.nm * (double)s->len);
        made_cluster = true;
      }
    }
    if (made_cluster) continue;
    if (d0.diag && d0.k <= 2) {
      if (nd >= TILE_MAXD) BT_TRY(flush());
      if (fill_diag_slot(P.d[nd], d0, local_pos, T, P.swz_mode) == 0) {
        P.item[nitems++] = (uint8_t)(TILE_DBASE + nd);
        nd++;
        used[i] = 1;
        g_fused_blocks++;
        g_fused_flops.store(g_fused_flops.load() + ldexp((double)s->len, -d0.nc) * 6.0);
        continue;
      }
    }
    if (ng >= TILE_MAXG) BT_TRY(flush());
    BT_TRY(fill_gate_slot(P.g[ng], d0, local_pos, T, P.swz_mode));
    P.item[nitems++] = (uint8_t)ng;
    ng++;
    used[i] = 1;
    g_fused_blocks++;
    {
      double touched = ldexp((double)s->len, -d0.nc);
      g_fused_flops.store(g_fused_flops.load() + touched * (d0.diag ? 6.0 : (d0.k == 2 ? 32.0 : 16.0)));
    }
  }
  return flush();
}

// ---- pass scheduler ---------------------------------------------------------------------------------------------------------
// A pass = the blocks one launch of the tile kernel carries + the index bits its tile needs beyond the fixed low bits.
// Dependencies: a block may run once no earlier unfinished block touches one of its bits ("blocked" set while scanning).
// Two policies (BT_FUSE_SCHED):
//   0  first fit: scan in program order, a block that does not depend on a skipped one joins the pass if its non-diagonal bits
//      still fit the tile -- the tile bits are whatever the first few blocks happen to need;
//   1  (default) look-ahead: the tile bits are CHOSEN.  Starting from the fixed low bits, the scheduler repeatedly evaluates, for
//      every candidate bit set (the missing bits of a block on the dependency frontier), how much work the pass would carry with
//      the tile enlarged by it -- a full dry run of the scan restricted to that tile -- and keeps the best gain per added bit;
//      brickwork / layered circuits then advance a deep light-cone on 12 chosen qubits instead of one layer on the first 12.
struct PassPlan {
  std::vector<int> blocks;     // indices into the block list, execution order
  std::vector<int> tile_bits;  // beyond the fixed low bits
  double cost = 0.0;
  int end_reason = 0;          // statistics: 1 = cost cap reached, 2 = block cap, 0 = ran out of eligible blocks
};

struct SchedCfg { int n_local, T, lowb, window; double maxg; int policy; bool split_classes; };

static inline double block_cost(const Block& b) {
  // cost units: a dense 4x4 block = 2.8 (the cap of 28 keeps ~10 of them per pass, the measured optimum for dense passes);
  // a structured block pays per micro-op (dispatch) and per FP64 instruction
  if (b.sok) return 0.15 * (double)b.prog.size() + 0.04 * b.scost;
  return b.desc.diag ? 0.7 : (b.desc.k == 2 ? 2.8 : 1.7);
}

struct SchedBlock { uint64_t touch, need; double cost; int ngates; bool solo, scalar, sok; };

// dry run (out == nullptr) or commit of one pass restricted to the tile bit set `tile`; returns the gain (cost units carried)
static double scan_pass(const std::vector<SchedBlock>& sb, const std::vector<char>& done, size_t first, const SchedCfg& cfg, uint64_t tile, PassPlan* out,
                        uint64_t* frontier_need /* optional: per frontier block, the missing bits (appended) */, std::vector<uint64_t>* frontier) {
  uint64_t blocked = 0;
  const uint64_t all = cfg.n_local >= 64 ? ~0ull : ((1ull << cfg.n_local) - 1);
  double cost = 0.0;
  int cnt = 0, reason = 0;
  int cls = -1;  // split_classes: the first block taken decides whether the pass carries structured (program) blocks or dense ones --
                 // only a program-only pass can run as a specialised straight-line kernel, one dense block would put the whole pass
                 // back on the interpreter
  size_t scanned = 0;
  (void)frontier_need;
  for (size_t i = first; i < sb.size() && scanned < (size_t)cfg.window && (blocked & all) != all; ++i) {
    if (done[i]) continue;
    scanned++;
    const SchedBlock& b = sb[i];
    bool ok = !(b.touch & blocked);
    if (ok && b.solo) {
      if (cnt == 0) { if (out) { out->blocks.push_back((int)i); out->cost = 0.0; } return 1e9; }  // runs alone, in order
      ok = false;
    }
    if (ok && cfg.split_classes && cls >= 0 && cls != (int)b.sok) ok = false;
    if (ok) {
      const uint64_t missing = b.need & ~tile;
      if (missing) { ok = false; if (frontier) frontier->push_back(missing); }
      else if (cnt >= 44) { ok = false; reason = 2; }
      else if (cost + b.cost > cfg.maxg && cnt > 0) { ok = false; reason = 1; }
      if (ok) {
        cost += b.cost; cnt++;
        cls = (int)b.sok;
        if (out) out->blocks.push_back((int)i);
        continue;
      }
    }
    blocked |= b.touch;
    if (b.scalar) break;  // a global scalar orders everything
  }
  if (out) { out->cost = cost; out->end_reason = reason; }
  return cost;
}

static int plan_next_pass(const std::vector<SchedBlock>& sb, const std::vector<char>& done, size_t first, const SchedCfg& cfg, PassPlan& out) {
  out = PassPlan();
  const uint64_t low = (1ull << cfg.lowb) - 1;
  uint64_t tile = low;
  if (cfg.policy == 0) {
    // first fit: tile bits are taken as the scan meets them
    uint64_t blocked = 0;
    const uint64_t all = cfg.n_local >= 64 ? ~0ull : ((1ull << cfg.n_local) - 1);
    double cost = 0.0;
    size_t scanned = 0;
    for (size_t i = first; i < sb.size() && scanned < (size_t)cfg.window && (blocked & all) != all; ++i) {
      if (done[i]) continue;
      scanned++;
      const SchedBlock& b = sb[i];
      bool ok = !(b.touch & blocked);
      if (ok && b.solo) {
        if (out.blocks.empty()) { out.blocks.push_back((int)i); return BT_OK; }
        ok = false;
      }
      if (ok) {
        const int extra = __builtin_popcountll(b.need & ~tile);
        if (__builtin_popcountll(tile) + extra > cfg.T || (int)out.blocks.size() >= 44 || (cost + b.cost > cfg.maxg && !out.blocks.empty())) ok = false;
        if (ok) { tile |= b.need; out.blocks.push_back((int)i); cost += b.cost; continue; }
      }
      blocked |= b.touch;
      if (b.scalar) break;
    }
    out.cost = cost;
  } else {
    std::vector<uint64_t> frontier;
    double cur = scan_pass(sb, done, first, cfg, tile, nullptr, nullptr, &frontier);
    if (cur < 1e8) {
      while (__builtin_popcountll(tile) < cfg.T) {
        // candidates: the distinct missing-bit sets of the blocks that were refused only for their bits
        std::sort(frontier.begin(), frontier.end());
        frontier.erase(std::unique(frontier.begin(), frontier.end()), frontier.end());
        double best_rate = 0.0, best_gain = 0.0;
        uint64_t best = 0;
        const int room = cfg.T - __builtin_popcountll(tile);
        for (uint64_t cand : frontier) {
          const int add = __builtin_popcountll(cand);
          if (add > room) continue;
          const double g = scan_pass(sb, done, first, cfg, tile | cand, nullptr, nullptr, nullptr);
          const double rate = (g - cur) / (double)add;
          if (rate > best_rate + 1e-12 || (rate > best_rate - 1e-12 && best && cand < best && rate > 0)) { best_rate = rate; best = cand; best_gain = g; }
        }
        if (!best) break;
        tile |= best;
        cur = best_gain;
        if (cur >= cfg.maxg - 0.15) break;  // the cost cap is reached: more bits cannot add work
        frontier.clear();
        scan_pass(sb, done, first, cfg, tile, nullptr, nullptr, &frontier);
      }
    }
    scan_pass(sb, done, first, cfg, tile, &out, nullptr, nullptr);
  }
  if (out.blocks.empty()) BT_FAIL(BT_ERR_ARG, "internal: fusion scheduler made no progress");
  uint64_t used = 0;
  for (int bi : out.blocks) used |= sb[bi].need;
  for (int b = cfg.lowb; b < cfg.n_local; ++b)
    if (used >> b & 1) out.tile_bits.push_back(b);
  return BT_OK;
}

static void sched_blocks(const std::vector<Block>& blocks, std::vector<SchedBlock>& sb) {
  std::vector<int> need;
  sb.resize(blocks.size());
  for (size_t i = 0; i < blocks.size(); ++i) {
    const Block& b = blocks[i];
    SchedBlock& o = sb[i];
    o.touch = o.need = 0;
    for (int t : b.touch) o.touch |= 1ull << t;
    o.solo = b.opaque || b.desc.k > 2;
    o.scalar = b.touch.empty();
    o.ngates = b.ngates;
    o.sok = b.sok;
    o.cost = 0.0;
    if (!o.solo) {
      needed_bits(b.desc, need);
      if (b.sok) need = b.abits;  // structured blocks: the bits they act on non-diagonally must be program positions
      for (int t : need) o.need |= 1ull << t;
      o.cost = block_cost(b);
    }
  }
}

static SchedCfg sched_cfg(int n_local) {
  SchedCfg c;
  c.n_local = n_local;
  c.T = std::min(n_local, env_int("BT_TILE_BITS", TILE_TDEF));
  c.lowb = std::min(tile_lowb(), c.T);
  c.policy = env_int("BT_FUSE_SCHED", 1);
  // cost units per pass (measured sweeps: profiles/r1_fusion_sweep.txt for first fit, profiles/r2_sched_sweep.txt for look-ahead)
  c.maxg = (double)std::max(1, std::min(80, env_int("BT_FUSE_MAX_GATES", c.policy != 0 ? 40 : 28)));
  c.window = env_int("BT_FUSE_WINDOW", 256);
  c.split_classes = env_int("BT_FUSE_SPLIT_CLASSES", env_int("BT_TILE_JIT", 1) != 0 ? 1 : 0) != 0;
  return c;
}

int bt_fuse_and_run(bt_sv* s, const std::vector<GateDesc>& gates) {
  std::vector<Block> blocks;
  fuse_blocks(gates, blocks);
  std::vector<SchedBlock> sb;
  sched_blocks(blocks, sb);
  const SchedCfg cfg = sched_cfg(s->n_local);
  const size_t n = blocks.size();
  std::vector<char> done(n, 0);
  size_t first = 0;
  PassPlan plan;
  std::vector<const Block*> pass;
  while (first < n) {
    if (done[first]) { first++; continue; }
    BT_TRY(plan_next_pass(sb, done, first, cfg, plan));
    pass.clear();
    for (int bi : plan.blocks) { pass.push_back(&blocks[bi]); done[bi] = 1; }
    BT_TRY(launch_pass(s, pass, plan.tile_bits));
  }
  return BT_OK;
}

// pure host (no device): the passes the scheduler forms for a gate list in physical-bit space on a register of n_local bits
int bt_fusion_plan(const std::vector<GateDesc>& gates, int n_local, int* n_passes, int* n_blocks, int* gates_in_pass, int* tile_bits /* cap x 16 */, double* cost_in_pass,
                   int* end_reason, int cap, int* dry_counts /* optional, 4 ints: launches, programs, other items, single-gate passes */) {
  std::vector<Block> blocks;
  fuse_blocks(gates, blocks);
  std::vector<SchedBlock> sb;
  sched_blocks(blocks, sb);
  const SchedCfg cfg = sched_cfg(n_local);
  std::vector<char> done(blocks.size(), 0);
  size_t first = 0;
  int np = 0;
  PassPlan plan;
  while (first < blocks.size()) {
    if (done[first]) { first++; continue; }
    BT_TRY(plan_next_pass(sb, done, first, cfg, plan));
    int ng = 0;
    for (int bi : plan.blocks) { done[bi] = 1; ng += blocks[bi].ngates; }
    if (dry_counts) {
      std::vector<const Block*> pass;
      for (int bi : plan.blocks) pass.push_back(&blocks[bi]);
      bt_sv fake;
      memset(&fake, 0, sizeof(fake));
      fake.n_qubits = fake.n_local = n_local;
      fake.n_batch = 1;
      fake.len = 1ull << n_local;
      t_dry = dry_counts;
      int rc = launch_pass(&fake, pass, plan.tile_bits);
      t_dry = nullptr;
      BT_TRY(rc);
    }
    if (np < cap) {
      if (gates_in_pass) gates_in_pass[np] = ng;
      if (cost_in_pass) cost_in_pass[np] = plan.cost;
      if (end_reason) end_reason[np] = plan.end_reason;
      if (tile_bits) {
        for (int j = 0; j < 16; ++j) tile_bits[np * 16 + j] = -1;
        for (size_t j = 0; j < plan.tile_bits.size() && j < 16; ++j) tile_bits[np * 16 + j] = plan.tile_bits[j];
      }
    }
    np++;
  }
  if (n_passes) *n_passes = np;
  if (n_blocks) *n_blocks = (int)blocks.size();
  return BT_OK;
}

extern "C" int bt_fusion_flops(double* flops) {
  if (!flops) BT_FAIL(BT_ERR_ARG, "null output");
  *flops = g_fused_flops.load();
  return BT_OK;
}

extern "C" int bt_fusion_stats(uint64_t* passes, uint64_t* blocks) {
  if (passes) *passes = g_fused_passes.load();
  if (blocks) *blocks = g_fused_blocks.load();
  return BT_OK;
}
