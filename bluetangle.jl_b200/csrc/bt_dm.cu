// bt_dm.cu -- density-matrix path.  rho (2^n x 2^n, Julia column-major: entry (r,c) at r + c*2^n) is stored as
// a 2n-qubit vector: row qubit q <-> bit n-q, column qubit q <-> bit 2n-q.
//
// Replaces apply(rho,op) = e_op*rho*e_op' (src/hilbert.jl:639-666: two sparse GEMMs per gate) and the Kraus sum
// sum_k E_k rho E_k' (src/struct.jl:58-76: 2*nK sparse GEMMs + nK sparse adds).  Here a unitary is the
// superoperator conj(U) (x) U and a channel is S = sum_k conj(K_k) (x) K_k on the (column, row) bit pair(s), so
// either costs ONE pass over rho (32 B per entry) whatever the number of Kraus operators.
#include "bt_internal.cuh"

int bt_build_gate(const bt_sv* s, int nq, int qubit, int target, int control, const bt_c64* m, GateDesc* out);
int bt_sample_diag(bt_sv* v, int n, const double* u, uint64_t shots, int64_t* out);

static int check_dm(const bt_dm* d) {
  if (!d || !d->v) BT_FAIL(BT_ERR_ARG, "null density-matrix handle");
  return bt_check_sv(d->v);
}

extern "C" int bt_dm_create(int n_qubits, bt_dm** out) {
  if (!out) BT_FAIL(BT_ERR_ARG, "null output handle");
  *out = nullptr;
  if (n_qubits < 1 || n_qubits > 20) BT_FAIL(BT_ERR_ARG, "density matrix supports 1..20 qubits");
  bt_sv* v = nullptr;
  BT_TRY(bt_sv_create_internal(2 * n_qubits, 2 * n_qubits, 1, false, &v));  // |0><0| is basis vector 0
  v->is_dm = true;
  v->dm_n = n_qubits;
  bt_dm* d = new bt_dm();
  d->v = v;
  d->n = n_qubits;
  *out = d;
  return BT_OK;
}

extern "C" int bt_dm_destroy(bt_dm* d) {
  if (!d) return BT_OK;
  bt_sv_destroy(d->v);
  delete d;
  return BT_OK;
}

extern "C" int bt_dm_n_qubits(const bt_dm* d, int* n) {
  if (!d || !n) BT_FAIL(BT_ERR_ARG, "null argument");
  *n = d->n;
  return BT_OK;
}

extern "C" int bt_dm_upload(bt_dm* d, const bt_c64* host, uint64_t len) { BT_TRY(check_dm(d)); return bt_sv_upload(d->v, host, len); }
extern "C" int bt_dm_download(const bt_dm* d, bt_c64* host, uint64_t len) { BT_TRY(check_dm(d)); return bt_sv_download(d->v, host, len); }
extern "C" int bt_dm_sync(const bt_dm* d) { BT_TRY(check_dm(d)); return bt_sv_sync(d->v); }
extern "C" int bt_dm_timer_start(bt_dm* d) { BT_TRY(check_dm(d)); return bt_sv_timer_start(d->v); }
extern "C" int bt_dm_timer_stop(bt_dm* d, float* ms) { BT_TRY(check_dm(d)); return bt_sv_timer_stop(d->v, ms); }
extern "C" int bt_dm_launch_count(const bt_dm* d, uint64_t* n) { BT_TRY(check_dm(d)); return bt_sv_launch_count(d->v, n); }

__global__ void k_outer(double2* __restrict__ rho, const double2* __restrict__ psi, int n) {
  // rho[r + c*2^n] = psi_r * conj(psi_c)
  uint64_t len = 1ull << (2 * n);
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t mask = (1ull << n) - 1;
  for (; i < len; i += stride) {
    double2 a = psi[i & mask], b = psi[i >> n];
    rho[i] = make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
  }
}

extern "C" int bt_dm_from_sv(bt_dm* d, const bt_sv* s) {
  BT_TRY(check_dm(d));
  if (!s || s->n_qubits != d->n || s->n_batch != 1 || s->world != 1) BT_FAIL(BT_ERR_ARG, "bt_dm_from_sv: need an unsharded single state vector on the same number of qubits");
  BT_CUDA(cudaStreamSynchronize(s->stream));
  bt_sv* v = d->v;
  unsigned grid = (unsigned)std::min<uint64_t>((v->len + 255) / 256, 148ull * 32);
  k_outer<<<grid, 256, 0, v->stream>>>(v->amp, s->amp, d->n);
  BT_CHECK_LAUNCH(v);
  return BT_OK;
}

// ---- superoperators -----------------------------------------------------------------------------------------------
// V acts on k qubits (row-major DxD).  S = conj(V) (x) V on 2k bits: matrix index bits [0..k-1] = row bits,
// [k..2k-1] = column bits;  S[(c,r),(c',r')] = conj(V[c][c']) * V[r][r'].
static void superop_add(int D, const cplx* V, std::vector<cplx>& S) {
  int DD = D * D;
  for (int c = 0; c < D; ++c)
    for (int r = 0; r < D; ++r)
      for (int c2 = 0; c2 < D; ++c2)
        for (int r2 = 0; r2 < D; ++r2)
          S[(size_t)(c * D + r) * DD + (c2 * D + r2)] += std::conj(V[c * D + c2]) * V[r * D + r2];
}

static void colmajor_to_rowmajor_c(const bt_c64* m, int D, cplx* out) {
  for (int r = 0; r < D; ++r)
    for (int c = 0; c < D; ++c) out[r * D + c] = c64(m[r + c * D]);
}

// Apply a general superoperator S (row-major, (D*D)x(D*D)) on row bits rb[] and column bits.
static int apply_superop(bt_dm* d, int k, const int* rb /* matrix bit t <-> row bit rb[t] */, const std::vector<cplx>& S) {
  bt_sv* v = d->v;
  int tb[4];
  for (int t = 0; t < k; ++t) { tb[t] = rb[t]; tb[k + t] = rb[t] + d->n; }
  GateDesc g;
  bt_canonicalize(2 * k, tb, S.data(), 0, nullptr, &g);
  return bt_launch_gate(v, g);
}

// apply the same operator V (any k, any controls) to the row bits, then conj(V) to the column bits: two launches,
// each touching what V touches (a control halves it) -- used for controlled gates where conj(V)(x)V is not itself a
// controlled operator.
static int apply_two_sided(bt_dm* d, const GateDesc& row) {
  bt_sv* v = d->v;
  BT_TRY(bt_launch_gate(v, row));
  GateDesc col = row;
  for (int i = 0; i < col.k; ++i) col.tb[i] += d->n;
  for (int i = 0; i < col.nc; ++i) col.cb[i] += d->n;
  int cnt = col.diag ? (1 << col.k) : (1 << (2 * col.k));
  for (int i = 0; i < cnt; ++i) col.m[i] = std::conj(col.m[i]);
  return bt_launch_gate(v, col);
}

// Build the row-side gate description using the n-qubit labelling (bits 0..n-1 of the 2n-bit vector).
static int build_row_gate(const bt_dm* d, int nq, int qubit, int target, int control, const bt_c64* m, GateDesc* g) {
  bt_sv tmp;
  memset(&tmp, 0, sizeof(tmp));
  tmp.n_qubits = d->n;
  tmp.n_local = d->n;
  for (int b = 0; b < 64; ++b) tmp.phys_of_bit[b] = b;
  return bt_build_gate(&tmp, nq, qubit, target, control, m, g);
}

extern "C" int bt_dm_apply_1q(bt_dm* d, int qubit, const bt_c64 m[4], int control) {
  BT_TRY(check_dm(d));
  GateDesc row;
  BT_TRY(build_row_gate(d, 1, qubit, -1, control, m, &row));
  if (row.nc > 0 || row.diag || row.k == 0) return apply_two_sided(d, row);
  // dense uncontrolled 1-qubit gate: one pass with the 4x4 superoperator
  std::vector<cplx> S(16, cplx(0, 0));
  superop_add(2, row.m, S);
  int rb[1] = {row.tb[0]};
  return apply_superop(d, 1, rb, S);
}

extern "C" int bt_dm_apply_2q(bt_dm* d, int qubit, int target, const bt_c64 m[16], int control) {
  BT_TRY(check_dm(d));
  GateDesc row;
  BT_TRY(build_row_gate(d, 2, qubit, target, control, m, &row));
  if (row.nc > 0 || row.diag || row.k < 2) return apply_two_sided(d, row);
  // dense 2-qubit unitary: one pass with the 16x16 superoperator conj(U)(x)U.  Measured on B200 (profiles/):
  // the 16x16 block (4 flop/B) still runs at the HBM roofline, so one pass beats two 4x4 passes.
  std::vector<cplx> S(256, cplx(0, 0));
  superop_add(4, row.m, S);
  int rb[2] = {row.tb[0], row.tb[1]};
  return apply_superop(d, 2, rb, S);
}

extern "C" int bt_dm_kraus(bt_dm* d, int nq, int qubit, int target, const bt_c64* K, int nK) {
  BT_TRY(check_dm(d));
  if (!K || nK < 1) BT_FAIL(BT_ERR_ARG, "invalid Kraus list");
  if (nq != 1 && nq != 2) BT_FAIL(BT_ERR_ARG, "density-matrix channels are available for 1 and 2 qubits");
  int n = d->n;
  if (nq == 1) {
    if (qubit < 1 || qubit > n) BT_FAIL(BT_ERR_ARG, "N must be larger than qubit");
    std::vector<cplx> S(16, cplx(0, 0));
    for (int k = 0; k < nK; ++k) {
      cplx V[4];
      colmajor_to_rowmajor_c(K + 4 * k, 2, V);
      superop_add(2, V, S);
    }
    int rb[1] = {n - qubit};
    return apply_superop(d, 1, rb, S);
  }
  if (qubit < 1 || target < 1 || qubit > n || target > n) BT_FAIL(BT_ERR_ARG, "N must be larger than qubits");
  if (qubit == target) BT_FAIL(BT_ERR_ARG, "`qubit` and `target_qubit` must differ");
  std::vector<cplx> S(256, cplx(0, 0));
  for (int k = 0; k < nK; ++k) {
    cplx V[16];
    colmajor_to_rowmajor_c(K + 16 * k, 4, V);
    superop_add(4, V, S);
  }
  // K indexed 2*b_qubit + b_target: matrix bit 0 <-> target, bit 1 <-> qubit
  int rb[2] = {n - target, n - qubit};
  return apply_superop(d, 2, rb, S);
}

// ---- op lists with superoperator fusion -----------------------------------------------------------------------------
// to_rho's loop (src/ops.jl:813-841) applies gate, then noise channel(s), op by op: 2 sparse GEMMs per unitary and 2*nK per
// channel.  Here consecutive unitaries and channels on the same qubit (pair) are multiplied on the host into ONE 4x4 /
// 16x16 superoperator, so "gate + depolarizing + amplitude damping" costs one pass over rho instead of three.
namespace {
struct SBlock {
  int nb;           // qubits: 1 or 2
  int bits[2];      // row bits; superoperator index bits: [0..nb-1] row, [nb..2nb-1] column
  std::vector<cplx> S;
  bool dead;
};

void smatmul(int D, const std::vector<cplx>& A, const std::vector<cplx>& B, std::vector<cplx>& C) {
  std::vector<cplx> T((size_t)D * D, cplx(0, 0));
  for (int r = 0; r < D; ++r)
    for (int k = 0; k < D; ++k) {
      cplx a = A[(size_t)r * D + k];
      if (a == cplx(0, 0)) continue;
      for (int c = 0; c < D; ++c) T[(size_t)r * D + c] += a * B[(size_t)k * D + c];
    }
  C.swap(T);
}

// 4x4 superoperator (index: bit0 row, bit1 column) of the qubit at block position p -> 16x16 (r0 r1 c0 c1)
void sembed1(const std::vector<cplx>& S1, int p, std::vector<cplx>& out) {
  out.assign(256, cplx(0, 0));
  int q = 1 - p;
  for (int I = 0; I < 16; ++I)
    for (int J = 0; J < 16; ++J) {
      int rI[2] = {I & 1, (I >> 1) & 1}, cI[2] = {(I >> 2) & 1, (I >> 3) & 1};
      int rJ[2] = {J & 1, (J >> 1) & 1}, cJ[2] = {(J >> 2) & 1, (J >> 3) & 1};
      if (rI[q] != rJ[q] || cI[q] != cJ[q]) continue;
      out[(size_t)I * 16 + J] = S1[(size_t)(rI[p] | (cI[p] << 1)) * 4 + (rJ[p] | (cJ[p] << 1))];
    }
}

// swap the two qubits of a 16x16 superoperator: index bits 0<->1 and 2<->3
void sswap(std::vector<cplx>& S) {
  auto perm = [](int I) { return ((I & 1) << 1) | ((I >> 1) & 1) | (((I >> 2) & 1) << 3) | (((I >> 3) & 1) << 2); };
  std::vector<cplx> T(256);
  for (int I = 0; I < 16; ++I)
    for (int J = 0; J < 16; ++J) T[(size_t)perm(I) * 16 + perm(J)] = S[(size_t)I * 16 + J];
  S.swap(T);
}
}  // namespace

extern "C" int bt_dm_apply_ops(bt_dm* d, const bt_dm_op* ops, uint64_t n, int fuse) {
  BT_TRY(check_dm(d));
  if (n && !ops) BT_FAIL(BT_ERR_ARG, "null op list");
  const int N = d->n;
  std::vector<SBlock> blocks;
  int last[64];
  for (int b = 0; b < 64; ++b) last[b] = -1;
  auto flush = [&]() -> int {
    for (SBlock& B : blocks) {
      if (B.dead) continue;
      BT_TRY(apply_superop(d, B.nb, B.bits, B.S));
    }
    blocks.clear();
    for (int b = 0; b < 64; ++b) last[b] = -1;
    return BT_OK;
  };
  for (uint64_t i = 0; i < n; ++i) {
    const bt_dm_op& o = ops[i];
    if (!o.mats) BT_FAIL(BT_ERR_ARG, "op %llu: null matrices", (unsigned long long)i);
    int nq = o.nq;
    if (nq != 1 && nq != 2) BT_FAIL(BT_ERR_ARG, "density-matrix ops act on 1 or 2 qubits");
    if (o.qubit < 1 || o.qubit > N || (nq == 2 && (o.target < 1 || o.target > N))) BT_FAIL(BT_ERR_ARG, "N must be larger than qubits");
    if (nq == 2 && o.qubit == o.target) BT_FAIL(BT_ERR_ARG, "`qubit` and `target_qubit` must differ");
    bool controlled = (o.kind == 0 && o.control != -2);
    if (!fuse || (controlled && nq == 2) || (o.kind == 1 && o.nK < 1)) {
      BT_TRY(flush());
      if (o.kind == 0) {
        if (nq == 1) BT_TRY(bt_dm_apply_1q(d, o.qubit, o.mats, o.control));
        else BT_TRY(bt_dm_apply_2q(d, o.qubit, o.target, o.mats, o.control));
      } else {
        BT_TRY(bt_dm_kraus(d, nq, o.qubit, o.target, o.mats, o.nK));
      }
      continue;
    }
    // superoperator of this op on its own bits
    int nb = nq, bits[2];
    std::vector<cplx> S;
    if (controlled) {
      // controlled 1-qubit unitary = dense 4x4 V on (bit0 = qubit, bit1 = control): P1 (x) U + P0 (x) I
      if (o.control < 1 || o.control > N || o.control == o.qubit) BT_FAIL(BT_ERR_ARG, "invalid control qubit");
      cplx U[4], V[16];
      colmajor_to_rowmajor_c(o.mats, 2, U);
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
          int rc = r >> 1, cc = c >> 1, rt = r & 1, ct = c & 1;
          V[r * 4 + c] = (rc != cc) ? cplx(0, 0) : (rc == 1 ? U[rt * 2 + ct] : (rt == ct ? cplx(1, 0) : cplx(0, 0)));
        }
      nb = 2; bits[0] = N - o.qubit; bits[1] = N - o.control;
      S.assign(256, cplx(0, 0));
      superop_add(4, V, S);
    } else {
      int D = 1 << nq;
      int cnt = (o.kind == 0) ? 1 : o.nK;
      S.assign((size_t)D * D * D * D, cplx(0, 0));
      for (int k = 0; k < cnt; ++k) {
        cplx V[16];
        colmajor_to_rowmajor_c(o.mats + (size_t)k * D * D, D, V);
        superop_add(D, V, S);
      }
      if (nq == 1) bits[0] = N - o.qubit;
      else { bits[0] = N - o.target; bits[1] = N - o.qubit; }  // matrix bit 0 <-> target, bit 1 <-> qubit
    }
    if (nb == 1) {
      int lb = last[bits[0]];
      if (lb >= 0) {
        SBlock& B = blocks[lb];
        if (B.nb == 1) smatmul(4, S, B.S, B.S);
        else {
          std::vector<cplx> E;
          sembed1(S, B.bits[0] == bits[0] ? 0 : 1, E);
          smatmul(16, E, B.S, B.S);
        }
        continue;
      }
      SBlock B; B.nb = 1; B.bits[0] = bits[0]; B.bits[1] = -1; B.S = S; B.dead = false;
      blocks.push_back(B);
      last[bits[0]] = (int)blocks.size() - 1;
      continue;
    }
    int l0 = last[bits[0]], l1 = last[bits[1]];
    if (l0 >= 0 && l0 == l1 && blocks[l0].nb == 2) {
      SBlock& B = blocks[l0];
      if (B.bits[0] != bits[0]) sswap(S);
      smatmul(16, S, B.S, B.S);
      continue;
    }
    SBlock B; B.nb = 2; B.bits[0] = bits[0]; B.bits[1] = bits[1]; B.S = S; B.dead = false;
    for (int t = 0; t < 2; ++t) {
      int lb = last[bits[t]];
      if (lb >= 0 && blocks[lb].nb == 1 && !blocks[lb].dead) {
        std::vector<cplx> E;
        sembed1(blocks[lb].S, t, E);
        smatmul(16, B.S, E, B.S);
        blocks[lb].dead = true;
      }
    }
    blocks.push_back(B);
    last[bits[0]] = last[bits[1]] = (int)blocks.size() - 1;
  }
  return flush();
}

__global__ void k_dephase(double2* __restrict__ a, uint64_t len, int rbit, int cbit) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < len; i += stride)
    if (((i >> rbit) ^ (i >> cbit)) & 1) a[i] = make_double2(0.0, 0.0);
}

extern "C" int bt_dm_dephase(bt_dm* d, int qubit) {
  BT_TRY(check_dm(d));
  if (qubit < 1 || qubit > d->n) BT_FAIL(BT_ERR_ARG, "N must be larger than qubit");
  bt_sv* v = d->v;
  unsigned grid = (unsigned)std::min<uint64_t>((v->len + 255) / 256, 148ull * 32);
  k_dephase<<<grid, 256, 0, v->stream>>>(v->amp, v->len, d->n - qubit, 2 * d->n - qubit);
  BT_CHECK_LAUNCH(v);
  return BT_OK;
}

__global__ void k_diag_extract(const double2* __restrict__ a, int n, double2* __restrict__ out) {
  uint64_t dim = 1ull << n;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < dim) out[i] = a[i * (dim + 1)];
}

static int dm_diag_to_host(const bt_dm* d, std::vector<double2>& h) {
  bt_sv* v = d->v;
  uint64_t dim = 1ull << d->n;
  double2* dd = nullptr;
  BT_CUDA(cudaMallocAsync(&dd, dim * sizeof(double2), v->stream));
  k_diag_extract<<<(unsigned)((dim + 255) / 256), 256, 0, v->stream>>>(v->amp, d->n, dd);
  BT_CHECK_LAUNCH(v);
  h.resize(dim);
  BT_CUDA(cudaMemcpyAsync(h.data(), dd, dim * sizeof(double2), cudaMemcpyDeviceToHost, v->stream));
  BT_CUDA(cudaFreeAsync(dd, v->stream));
  BT_CUDA(cudaStreamSynchronize(v->stream));
  return BT_OK;
}

extern "C" int bt_dm_diag(const bt_dm* d, double* host) {
  BT_TRY(check_dm(d));
  if (!host) BT_FAIL(BT_ERR_ARG, "null output");
  std::vector<double2> h;
  BT_TRY(dm_diag_to_host(d, h));
  for (size_t i = 0; i < h.size(); ++i) host[i] = h[i].x;
  return BT_OK;
}

extern "C" int bt_dm_trace(const bt_dm* d, bt_c64* out) {
  BT_TRY(check_dm(d));
  if (!out) BT_FAIL(BT_ERR_ARG, "null output");
  std::vector<double2> h;
  BT_TRY(dm_diag_to_host(d, h));
  double re = 0, im = 0;
  for (size_t i = 0; i < h.size(); ++i) { re += h[i].x; im += h[i].y; }
  out->re = re; out->im = im;
  return BT_OK;
}

// tr(rho * P) for a Pauli string: sum_r rho[r, r^xm] * (P)[r^xm, r]  -- a gather of 2^n entries.
__global__ void __launch_bounds__(256) k_dm_pauli(const double2* __restrict__ a, int n, uint64_t xm, uint64_t zm, int ny, double* __restrict__ part) {
  __shared__ double sm[8];
  uint64_t dim = 1ull << n;
  double cr, ci;
  switch (ny & 3) { case 0: cr = 1; ci = 0; break; case 1: cr = 0; ci = -1; break; case 2: cr = -1; ci = 0; break; default: cr = 0; ci = 1; }
  double acc = 0.0;
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < dim; c += (uint64_t)gridDim.x * blockDim.x) {
    // tr(rho P) = sum_{r,c} rho[r,c] P[c,r];  P[c,r] != 0 iff c = r^xm:  P[r^xm, r] = (-i)^ny (-1)^{popc((r^xm) & zm)}
    uint64_t r = c ^ xm;
    double2 x = a[r + c * dim];
    double sgn = (__popcll(c & zm) & 1) ? -1.0 : 1.0;
    acc += sgn * (cr * x.x - ci * x.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0; for (int w = 0; w < 8; ++w) s += sm[w]; part[blockIdx.x] = s; }
}

extern "C" int bt_dm_expect_pauli(const bt_dm* d, const char* paulis, double* out) {
  BT_TRY(check_dm(d));
  if (!paulis || !out) BT_FAIL(BT_ERR_ARG, "null argument");
  int n = d->n;
  if ((int)strlen(paulis) != n) BT_FAIL(BT_ERR_ARG, "Pauli string must have %d characters", n);
  uint64_t xm = 0, zm = 0; int ny = 0;
  for (int q = 1; q <= n; ++q) {
    uint64_t bit = 1ull << (n - q);
    switch (paulis[q - 1]) {
      case 'I': case 'i': break;
      case 'X': case 'x': xm |= bit; break;
      case 'Y': case 'y': xm |= bit; zm |= bit; ny++; break;
      case 'Z': case 'z': zm |= bit; break;
      default: BT_FAIL(BT_ERR_ARG, "Pauli string may only contain I, X, Y, Z");
    }
  }
  bt_sv* v = d->v;
  uint64_t dim = 1ull << n;
  int nblk = (int)std::min<uint64_t>((dim + 255) / 256, 1024);
  BT_TRY(bt_ensure_partials(v, nblk));
  k_dm_pauli<<<nblk, 256, 0, v->stream>>>(v->amp, n, xm, zm, ny, v->d_part);
  BT_CHECK_LAUNCH(v);
  BT_CUDA(cudaMemcpyAsync(v->h_res, v->d_part, nblk * sizeof(double), cudaMemcpyDeviceToHost, v->stream));
  BT_CUDA(cudaStreamSynchronize(v->stream));
  double s = 0; for (int b = 0; b < nblk; ++b) s += v->h_res[b];
  *out = s;
  return BT_OK;
}

// the 2x2 blocks of rho needed for <O_q>: for each q, sum over the other bits of rho restricted to bit q
__global__ void __launch_bounds__(256) k_dm_rdm1(const double2* __restrict__ a, int n, int bit, double* __restrict__ part) {
  __shared__ double sm[8 * 4];
  uint64_t dim = 1ull << n;
  uint64_t half = dim >> 1;
  double acc[4] = {0, 0, 0, 0};  // rho00, rho11, Re rho01, Im rho01
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < half; g += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r0 = ((g >> bit) << (bit + 1)) | (g & ((1ull << bit) - 1));
    uint64_t r1 = r0 | (1ull << bit);
    acc[0] += a[r0 + r0 * dim].x;
    acc[1] += a[r1 + r1 * dim].x;
    double2 x = a[r0 + r1 * dim];
    acc[2] += x.x; acc[3] += x.y;
  }
  for (int i = 0; i < 4; ++i) {
    double x = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) sm[(threadIdx.x >> 5) * 4 + i] = x;
  }
  __syncthreads();
  if (threadIdx.x < 4) { double s = 0; for (int w = 0; w < 8; ++w) s += sm[w * 4 + threadIdx.x]; part[blockIdx.x * 4 + threadIdx.x] = s; }
}

extern "C" int bt_dm_expect_1q_all(const bt_dm* d, const bt_c64 m[4], double* out) {
  BT_TRY(check_dm(d));
  if (!m || !out) BT_FAIL(BT_ERR_ARG, "null argument");
  bt_sv* v = d->v;
  int n = d->n;
  uint64_t half = 1ull << (n - 1);
  int nblk = (int)std::max<uint64_t>(1, std::min<uint64_t>((half + 255) / 256, 64));
  if ((size_t)n * nblk * 4 > v->res_cap) nblk = 1;
  BT_TRY(bt_ensure_partials(v, (size_t)n * nblk * 4));
  for (int q = 1; q <= n; ++q) {
    k_dm_rdm1<<<nblk, 256, 0, v->stream>>>(v->amp, n, n - q, v->d_part + (size_t)(q - 1) * nblk * 4);
    BT_CHECK_LAUNCH(v);
  }
  BT_CUDA(cudaMemcpyAsync(v->h_res, v->d_part, (size_t)n * nblk * 4 * sizeof(double), cudaMemcpyDeviceToHost, v->stream));
  BT_CUDA(cudaStreamSynchronize(v->stream));
  for (int q = 1; q <= n; ++q) {
    double r[4] = {0, 0, 0, 0};
    for (int b = 0; b < nblk; ++b)
      for (int i = 0; i < 4; ++i) r[i] += v->h_res[((size_t)(q - 1) * nblk + b) * 4 + i];
    // tr(rho_q O) = sum_ab rho[a][b] O[b][a]; rho01 = r[2] + i r[3], rho10 = conj
    double e = m[0].re * r[0] + m[3].re * r[1];
    e += m[1].re * r[2] - m[1].im * r[3];  // rho01 * O[1,0]
    e += m[2].re * r[2] + m[2].im * r[3];  // rho10 * O[0,1]
    out[q - 1] = e;
  }
  return BT_OK;
}

extern "C" int bt_dm_expect_product(const bt_dm* d, int n_ops, const int* qubits, const bt_c64* mats, double* out) {
  // real(tr(rho * (x)O))  (src/func.jl:144-146): apply (x)O to the row index of a copy, then take the trace.
  BT_TRY(check_dm(d));
  if (n_ops < 0 || (n_ops > 0 && (!qubits || !mats)) || !out) BT_FAIL(BT_ERR_ARG, "null argument");
  bt_sv* v = d->v;
  for (int i = 0; i < n_ops; ++i) {
    if (qubits[i] < 1 || qubits[i] > d->n) BT_FAIL(BT_ERR_ARG, "qubit %d out of range", qubits[i]);
    for (int j = 0; j < i; ++j)
      if (qubits[i] == qubits[j]) BT_FAIL(BT_ERR_ARG, "repeated qubit %d in operator list", qubits[i]);
  }
  BT_TRY(bt_ensure_alt(v));
  BT_CUDA(cudaMemcpyAsync(v->alt, v->amp, v->len * sizeof(double2), cudaMemcpyDeviceToDevice, v->stream));
  std::swap(v->amp, v->alt);
  int rc = BT_OK;
  for (int i = 0; i < n_ops && rc == BT_OK; ++i) {
    GateDesc g;
    rc = build_row_gate(d, 1, qubits[i], -1, -2, mats + 4 * i, &g);
    if (rc == BT_OK) rc = bt_launch_gate(v, g);   // (O rho): O acts on the row index; tr(O rho) = tr(rho O)
  }
  bt_c64 tr;
  if (rc == BT_OK) rc = bt_dm_trace(d, &tr);
  std::swap(v->amp, v->alt);
  if (rc != BT_OK) return rc;
  *out = tr.re;
  return BT_OK;
}

extern "C" int bt_dm_sample(const bt_dm* d, const double* u, uint64_t shots, int64_t* out) {
  BT_TRY(check_dm(d));
  return bt_sample_diag(d->v, d->n, u, shots, out);
}
