// bt_linalg.cu -- the dense linear algebra the reference's observables need beyond 1-3 qubit reductions, on the device:
//
//   * Schmidt spectrum of a bipartition (entanglement_entropy, src/func.jl:299-312: svdvals of the 2^(N/2) x 2^(N-N/2)
//     reshape) by a one-sided (Hestenes) Jacobi iteration on a scratch copy of the state: the vectors of the short side
//     are rotated pairwise until mutually orthogonal; their squared norms are the squared singular values.  One CTA owns
//     one pair per round (both vectors live in registers between the dot products and the rotation: one read and one
//     write of the pair), a round-robin tournament gives nvec/2 disjoint pairs per launch, batches of trajectories add a
//     grid dimension.  No Gram matrix (no squaring of the condition number), no library, no host LAPACK.
//   * partial_trace(state, keep) for any number of kept qubits (src/linalg.jl:83-140; the reference forms state*state' and
//     loops over 4^k x 2^(N-k) entries): a tiled Gram kernel rho = A A^dagger over the gathered 2^k x 2^(N-k) matrix.
//   * partial_trace / bipartition_trace of a density matrix (src/linalg.jl:88-161) and expect(x, op) for every Op the
//     reference accepts (src/func.jl:91-92: 2-qubit and controlled operators, on states and density matrices) as a trace
//     against the reduced density matrix of the qubits the operator touches.
#include "bt_internal.cuh"

int bt_prepare_local_bits(bt_sv* s, int n, const int* logical_bits);  // bt_dist.cu

// ---- scratch copy with the vectors of the short side contiguous ---------------------------------------------------------
// out[t][r * m + c] = in[t][c * n + r]  (n = 2^lb rows = the low lb index bits, m = 2^(n_local - lb))
__global__ void __launch_bounds__(256) k_rows_from_low_bits(const double2* __restrict__ in, double2* __restrict__ out, int lb, int n_local) {
  __shared__ double2 tile[32][33];
  const uint64_t n = 1ull << lb, m = 1ull << (n_local - lb);
  const double2* src = in + ((uint64_t)blockIdx.z << n_local);
  double2* dst = out + ((uint64_t)blockIdx.z << n_local);
  const uint64_t r0 = (uint64_t)blockIdx.x * 32, c0 = (uint64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    uint64_t c = c0 + j, r = r0 + tx;
    if (c < m && r < n) tile[j][tx] = src[c * n + r];
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    uint64_t r = r0 + j, c = c0 + tx;
    if (r < n && c < m) dst[r * m + c] = tile[tx][j];
  }
}

// fixed-order block sum of NV doubles per thread; every thread returns the totals
template <int NV>
__device__ __forceinline__ void block_allreduce(double (&v)[NV], double* sm /* 32 * NV */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double x = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sm[warp * NV + i] = x;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double x = 0.0;
    for (int w = 0; w < nw; ++w) x += sm[w * NV + i];
    v[i] = x;
  }
}

// round-robin tournament (circle method): in round r = 0..nvec-2, slot k = 0..nvec/2-1 plays the pair (p, q); every unordered
// pair of the nvec (even) vectors meets exactly once per sweep and the pairs of one round are disjoint
__host__ __device__ __forceinline__ void tournament_pair(int nvec, int round, int k, int* p, int* q) {
  const int n1 = nvec - 1;
  if (k == 0) { *p = n1; *q = round; }
  else { *p = (round + k) % n1; *q = (round - k + n1) % n1; }
}

extern "C" int bt_jacobi_pairs_host(int nvec, int round, int* pairs /* nvec/2 x 2 */) {
  if (nvec < 2 || (nvec & 1) || round < 0 || round >= nvec - 1 || !pairs) BT_FAIL(BT_ERR_ARG, "bt_jacobi_pairs_host: nvec even >= 2, round in 0..nvec-2");
  for (int k = 0; k < nvec / 2; ++k) tournament_pair(nvec, round, k, pairs + 2 * k, pairs + 2 * k + 1);
  return BT_OK;
}

// One round of the tournament: CTA k rotates the pair (p, q) of vectors of length len (EPT elements per thread).
// A pair is rotated iff |<y,x>| > tol * |x| |y|; the rotation makes the two vectors orthogonal (complex Hestenes step:
// a phase on y makes the inner product real, then a real plane rotation).
// `acc` (optional): a companion matrix of nvec vectors of length acc_len that receives the same column rotations -- started as the
// identity it accumulates the right singular vectors (eigenvectors, for a Hermitian input).
template <int EPT>
__global__ void __launch_bounds__(512) k_jacobi_round(double2* __restrict__ R, int nvec, int len, int round, double tol, const double* __restrict__ frob2,
                                                      unsigned int* __restrict__ rotations, double2* __restrict__ accm = nullptr, int acc_len = 0) {
  __shared__ double sm[32 * 4];
  int p, q;
  tournament_pair(nvec, round, blockIdx.x, &p, &q);
  double2* __restrict__ x = R + ((uint64_t)blockIdx.y * nvec + p) * (uint64_t)len;
  double2* __restrict__ y = R + ((uint64_t)blockIdx.y * nvec + q) * (uint64_t)len;
  double2 xv[EPT], yv[EPT];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};  // |x|^2, |y|^2, Re/Im of sum x conj(y)
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    const int i = threadIdx.x + j * blockDim.x;
    if (i < len) {
      xv[j] = x[i]; yv[j] = y[i];
      acc[0] = fma(xv[j].x, xv[j].x, fma(xv[j].y, xv[j].y, acc[0]));
      acc[1] = fma(yv[j].x, yv[j].x, fma(yv[j].y, yv[j].y, acc[1]));
      acc[2] = fma(xv[j].x, yv[j].x, fma(xv[j].y, yv[j].y, acc[2]));
      acc[3] = fma(xv[j].y, yv[j].x, fma(-xv[j].x, yv[j].y, acc[3]));
    }
  }
  block_allreduce<4>(acc, sm);
  const double a = acc[0], b = acc[1], g2 = acc[2] * acc[2] + acc[3] * acc[3];
  // relative criterion, plus an absolute floor at rounding level of the whole matrix (Frobenius norm^2 F of this trajectory): vectors that
  // have been rotated down to rounding noise keep an O(1) relative overlap with each other for ever -- their inner products are
  // below 1e-16 F and change no singular value at the 1e-16 level
  const double fl = 1e-16 * frob2[blockIdx.y];
  if (!(g2 > tol * tol * a * b && g2 > fl * fl)) return;  // uniform across the CTA (also covers zero vectors and NaN)
  if (threadIdx.x == 0) atomicAdd(rotations, 1u);
  const double gabs = sqrt(g2);
  const double pr = acc[2] / gabs, pi = acc[3] / gabs;  // e^{i phi} = g / |g|
  const double zeta = (b - a) / (2.0 * gabs);
  const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    const int i = threadIdx.x + j * blockDim.x;
    if (i < len) {
      const double yr = pr * yv[j].x - pi * yv[j].y, yi = pr * yv[j].y + pi * yv[j].x;  // y' = e^{i phi} y
      x[i] = make_double2(c * xv[j].x - sn * yr, c * xv[j].y - sn * yi);
      y[i] = make_double2(sn * xv[j].x + c * yr, sn * xv[j].y + c * yi);
    }
  }
  if (accm) {
    double2* __restrict__ vx = accm + ((uint64_t)blockIdx.y * nvec + p) * (uint64_t)acc_len;
    double2* __restrict__ vy = accm + ((uint64_t)blockIdx.y * nvec + q) * (uint64_t)acc_len;
    for (int i = threadIdx.x; i < acc_len; i += blockDim.x) {
      const double2 a2 = vx[i], b2 = vy[i];
      const double yr = pr * b2.x - pi * b2.y, yi = pr * b2.y + pi * b2.x;
      vx[i] = make_double2(c * a2.x - sn * yr, c * a2.y - sn * yi);
      vy[i] = make_double2(sn * a2.x + c * yr, sn * a2.y + c * yi);
    }
  }
}

// squared norm of every vector: one CTA per (vector, trajectory)
__global__ void __launch_bounds__(256) k_vec_norm2(const double2* __restrict__ R, int nvec, int len, double* __restrict__ out) {
  __shared__ double sm[32];
  const double2* x = R + ((uint64_t)blockIdx.y * nvec + blockIdx.x) * (uint64_t)len;
  double acc[1] = {0.0};
  for (int i = threadIdx.x; i < len; i += blockDim.x) acc[0] = fma(x[i].x, x[i].x, fma(x[i].y, x[i].y, acc[0]));
  block_allreduce<1>(acc, sm);
  if (threadIdx.x == 0) out[(uint64_t)blockIdx.y * nvec + blockIdx.x] = acc[0];
}

// F[t] = sum of the squared norms of trajectory t's vectors (one thread per trajectory: nvec <= 8192 additions, once per call)
__global__ void k_frob2(const double* __restrict__ norm2, int nvec, int64_t n_batch, double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_batch) return;
  double f = 0.0;
  for (int i = 0; i < nvec; ++i) f += norm2[t * nvec + i];
  out[t] = f;
}

static int launch_round(bt_sv* s, double2* R, int nvec, int len, int round, double tol, const double* d_frob, unsigned int* d_rot, double2* accm = nullptr,
                        int acc_len = 0) {
  dim3 grid(nvec / 2, (unsigned)s->n_batch);
  int threads = len <= 256 ? std::max(32, len) : (len <= 4096 ? 256 : 512);
  int ept = (len + threads - 1) / threads;
  switch (ept) {
    case 1: k_jacobi_round<1><<<grid, threads, 0, s->stream>>>(R, nvec, len, round, tol, d_frob, d_rot, accm, acc_len); break;
    case 2: k_jacobi_round<2><<<grid, threads, 0, s->stream>>>(R, nvec, len, round, tol, d_frob, d_rot, accm, acc_len); break;
    case 4: k_jacobi_round<4><<<grid, threads, 0, s->stream>>>(R, nvec, len, round, tol, d_frob, d_rot, accm, acc_len); break;
    case 8: k_jacobi_round<8><<<grid, threads, 0, s->stream>>>(R, nvec, len, round, tol, d_frob, d_rot, accm, acc_len); break;
    case 16: k_jacobi_round<16><<<grid, threads, 0, s->stream>>>(R, nvec, len, round, tol, d_frob, d_rot, accm, acc_len); break;
    default: BT_FAIL(BT_ERR_UNSUPPORTED, "internal: vector length %d", len);
  }
  BT_CHECK_LAUNCH(s);
  return BT_OK;
}

// Jacobi sweeps over nvec vectors of length len per trajectory stored contiguously in R; the squared norms of the
// converged vectors (descending) go to spec[n_batch][nvec].
static int jacobi_spectrum(bt_sv* s, double2* R, int nvec, int len, double* spec, int* sweeps_out, double2* accm = nullptr, int acc_len = 0, bool sorted = true) {
  if (len > 8192) BT_FAIL(BT_ERR_UNSUPPORTED, "Schmidt spectrum: the long side of the cut has 2^%d > 2^13 entries", (int)log2((double)len));
  if (s->n_batch > 65535) BT_FAIL(BT_ERR_UNSUPPORTED, "n_batch > 65535");
  size_t nvals = (size_t)s->n_batch * nvec;
  BT_TRY(bt_ensure_scratch(s, (nvals + (size_t)s->n_batch) * sizeof(double) + 64));
  double* d_norm = (double*)s->d_scratch;
  double* d_frob = d_norm + nvals;
  unsigned int* d_rot = (unsigned int*)(d_frob + s->n_batch);
  {
    dim3 g0(nvec, (unsigned)s->n_batch);
    k_vec_norm2<<<g0, std::min(256, std::max(32, len)), 0, s->stream>>>(R, nvec, len, d_norm);
    BT_CHECK_LAUNCH(s);
    k_frob2<<<(unsigned)((s->n_batch + 127) / 128), 128, 0, s->stream>>>(d_norm, nvec, s->n_batch, d_frob);
    BT_CHECK_LAUNCH(s);
  }
  const double tol = 1e-15 * sqrt((double)len) + 1e-15;
  int sweeps = 0;
  const int max_sweeps = 60;
  if (nvec >= 2) {
    for (; sweeps < max_sweeps;) {
      BT_CUDA(cudaMemsetAsync(d_rot, 0, sizeof(unsigned int), s->stream));
      for (int r = 0; r < nvec - 1; ++r) BT_TRY(launch_round(s, R, nvec, len, r, tol, d_frob, d_rot, accm, acc_len));
      ++sweeps;
      BT_CUDA(cudaMemcpyAsync(s->h_flag, d_rot, sizeof(unsigned int), cudaMemcpyDeviceToHost, s->stream));
      BT_CUDA(cudaStreamSynchronize(s->stream));
      if (*(unsigned int*)s->h_flag == 0) break;
    }
    if (*(unsigned int*)s->h_flag != 0) BT_FAIL(BT_ERR_UNSUPPORTED, "Schmidt spectrum: Jacobi iteration did not converge in %d sweeps", max_sweeps);
  }
  dim3 grid(nvec, (unsigned)s->n_batch);
  k_vec_norm2<<<grid, std::min(256, std::max(32, len)), 0, s->stream>>>(R, nvec, len, d_norm);
  BT_CHECK_LAUNCH(s);
  BT_CUDA(cudaMemcpyAsync(spec, d_norm, nvals * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  BT_CUDA(cudaStreamSynchronize(s->stream));
  if (sorted)
    for (int64_t t = 0; t < s->n_batch; ++t) std::sort(spec + t * nvec, spec + (t + 1) * nvec, [](double u, double v) { return u > v; });
  if (sweeps_out) *sweeps_out = sweeps;
  return BT_OK;
}

extern "C" int bt_sv_schmidt_spectrum(const bt_sv* cs, int n_low, double* spec, int* sweeps) {
  BT_TRY(bt_check_sv(cs));
  bt_sv* s = const_cast<bt_sv*>(cs);
  if (!spec) BT_FAIL(BT_ERR_ARG, "null output");
  if (s->world > 1) BT_FAIL(BT_ERR_UNSUPPORTED, "Schmidt spectrum of a sharded state");
  if (s->is_dm) BT_FAIL(BT_ERR_ARG, "Schmidt spectrum takes a state vector");
  const int N = s->n_local;
  if (n_low < 0 || n_low > N) BT_FAIL(BT_ERR_ARG, "cut position %d outside 0..%d", n_low, N);
  if (std::max(n_low, N - n_low) > 13)
    BT_FAIL(BT_ERR_UNSUPPORTED, "Schmidt spectrum: the long side of the cut has 2^%d > 2^13 entries (a vector pair must fit one CTA's registers)", std::max(n_low, N - n_low));
  BT_TRY(bt_ensure_alt(s));
  int nvec, len;
  if (n_low <= N - n_low) {  // the low side is the short one: its 2^n_low rows become contiguous vectors
    nvec = 1 << n_low; len = 1 << (N - n_low);
    if (n_low == 0) BT_CUDA(cudaMemcpyAsync(s->alt, s->amp, s->len * sizeof(double2), cudaMemcpyDeviceToDevice, s->stream));
    else {
      dim3 grid((nvec + 31) / 32, (len + 31) / 32, (unsigned)s->n_batch);
      if (s->n_batch > 65535) BT_FAIL(BT_ERR_UNSUPPORTED, "n_batch > 65535");
      k_rows_from_low_bits<<<grid, 256, 0, s->stream>>>(s->amp, s->alt, n_low, N);
      BT_CHECK_LAUNCH(s);
    }
  } else {  // the columns (one per value of the high bits) are contiguous already
    nvec = 1 << (N - n_low); len = 1 << n_low;
    BT_CUDA(cudaMemcpyAsync(s->alt, s->amp, s->len * sizeof(double2), cudaMemcpyDeviceToDevice, s->stream));
  }
  return jacobi_spectrum(s, s->alt, nvec, len, spec, sweeps);
}

// ---- partial_trace(state, keep) for any k: rho = A A^dagger, A = 2^k x 2^(N-k) gathered view of the state ---------------
struct GramPlan {
  int k, nenv;
  int kb[16];  // matrix index bit t -> physical bit
  int eb[48];  // environment index bit j -> physical bit (ascending)
};

__device__ __forceinline__ uint64_t deposit_bits(uint64_t v, const int* __restrict__ pos, int n) {
  uint64_t o = 0;
  for (int j = 0; j < n; ++j) o |= ((v >> j) & 1ull) << pos[j];
  return o;
}

// CTA = one T x T tile of rho over one slice of the environment; 256 threads as 16 x 16, (T/16)^2 outputs each.
template <int T>
__global__ void __launch_bounds__(256) k_gram(const double2* __restrict__ a, int n_local, const __grid_constant__ GramPlan P, double2* __restrict__ part, int S) {
  constexpr int PER = T / 16;
  __shared__ double2 As[T][33], Bs[T][33];
  __shared__ uint64_t rowoff[2][T], elow[32];
  const int D = 1 << P.k, tiles = D / T;
  const int ti = blockIdx.x % tiles, tj = blockIdx.x / tiles, sl = blockIdx.y;
  const uint64_t E = 1ull << P.nenv;
  const uint64_t stages = (E + 31) / 32, per = (stages + S - 1) / S;
  const uint64_t st0 = sl * per, st1 = min(stages, st0 + per);
  const double2* __restrict__ src = a + ((uint64_t)blockIdx.z << n_local);
  for (int i = threadIdx.x; i < T; i += 256) {
    rowoff[0][i] = deposit_bits((uint64_t)(ti * T + i), P.kb, P.k);
    rowoff[1][i] = deposit_bits((uint64_t)(tj * T + i), P.kb, P.k);
  }
  if (threadIdx.x < 32) elow[threadIdx.x] = deposit_bits((uint64_t)threadIdx.x, P.eb, min(P.nenv, 5));
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int le = threadIdx.x & 31, lr = threadIdx.x >> 5;
  double2 acc[PER][PER];
#pragma unroll
  for (int u = 0; u < PER; ++u)
#pragma unroll
    for (int v = 0; v < PER; ++v) acc[u][v] = make_double2(0.0, 0.0);
  for (uint64_t st = st0; st < st1; ++st) {
    const uint64_t ehigh = P.nenv > 5 ? deposit_bits(st, P.eb + 5, P.nenv - 5) : 0ull;
    const bool live = (st * 32 + le) < E;
#pragma unroll
    for (int r = lr; r < T; r += 8) {
      As[r][le] = live ? src[rowoff[0][r] | ehigh | elow[le]] : make_double2(0.0, 0.0);
      Bs[r][le] = live ? src[rowoff[1][r] | ehigh | elow[le]] : make_double2(0.0, 0.0);
    }
    __syncthreads();
#pragma unroll 8
    for (int e = 0; e < 32; ++e) {
      double2 av[PER], bv[PER];
#pragma unroll
      for (int u = 0; u < PER; ++u) { av[u] = As[ty + 16 * u][e]; bv[u] = Bs[tx + 16 * u][e]; }
#pragma unroll
      for (int u = 0; u < PER; ++u)
#pragma unroll
        for (int v = 0; v < PER; ++v) {  // acc += a * conj(b)
          acc[u][v].x = fma(av[u].x, bv[v].x, fma(av[u].y, bv[v].y, acc[u][v].x));
          acc[u][v].y = fma(av[u].y, bv[v].x, fma(-av[u].x, bv[v].y, acc[u][v].y));
        }
    }
    __syncthreads();
  }
  double2* out = part + ((uint64_t)blockIdx.z * S + sl) * (uint64_t)D * D;
#pragma unroll
  for (int u = 0; u < PER; ++u)
#pragma unroll
    for (int v = 0; v < PER; ++v) {
      const int i = ti * T + ty + 16 * u, j = tj * T + tx + 16 * v;
      out[(uint64_t)i + (uint64_t)j * D] = acc[u][v];  // column-major rho[i][j]
    }
}

// res[t][v] = sum_sl part[(t * S + sl) * nv + v], fixed order
__global__ void __launch_bounds__(256) k_sum_slices(const double* __restrict__ part, double* __restrict__ res, int S, uint64_t nv, uint64_t total) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t t = i / nv, v = i % nv;
    double x = 0.0;
    for (int sl = 0; sl < S; ++sl) x += part[(t * S + sl) * nv + v];
    res[i] = x;
  }
}

// general reduced density matrix of a state vector: kept physical bits tb[0..k) (matrix bit t <-> tb[t]), k >= 4
int bt_reduce_rdm_gram(bt_sv* s, int k, const int* tb, bt_c64* out_host) {
  if (k < 4 || k > 12) BT_FAIL(BT_ERR_UNSUPPORTED, "general partial_trace keeps up to 12 qubits (asked for %d)", k);
  if (s->n_batch > 65535) BT_FAIL(BT_ERR_UNSUPPORTED, "n_batch > 65535");
  GramPlan P;
  memset(&P, 0, sizeof(P));
  P.k = k; P.nenv = s->n_local - k;
  uint64_t used = 0;
  for (int t = 0; t < k; ++t) {
    if (tb[t] < 0 || tb[t] >= s->n_local) BT_FAIL(BT_ERR_UNSUPPORTED, "reduced density matrix over a global qubit needs a remap first");
    if (used >> tb[t] & 1) BT_FAIL(BT_ERR_ARG, "repeated qubit");
    used |= 1ull << tb[t];
    P.kb[t] = tb[t];
  }
  for (int b = 0, j = 0; b < s->n_local; ++b)
    if (!(used >> b & 1)) P.eb[j++] = b;
  const int D = 1 << k, T = k == 4 ? 16 : 32, tiles = D / T;
  const uint64_t stages = ((1ull << P.nenv) + 31) / 32;
  uint64_t S = 1;
  while ((uint64_t)tiles * tiles * S * (uint64_t)s->n_batch < 592 && S * 2 <= stages) S *= 2;
  const size_t mat = (size_t)D * D * sizeof(double2);
  const size_t need = mat * (size_t)s->n_batch * (S + 1);
  BT_TRY(bt_ensure_scratch(s, need));
  double2* res = (double2*)s->d_scratch;
  double2* part = S > 1 ? res + (size_t)D * D * s->n_batch : res;
  dim3 grid(tiles * tiles, (unsigned)S, (unsigned)s->n_batch);
  if (T == 16) k_gram<16><<<grid, 256, 0, s->stream>>>(s->amp, s->n_local, P, part, (int)S);
  else k_gram<32><<<grid, 256, 0, s->stream>>>(s->amp, s->n_local, P, part, (int)S);
  BT_CHECK_LAUNCH(s);
  if (S > 1) {
    const uint64_t nv = 2ull * D * D, total = nv * (uint64_t)s->n_batch;
    k_sum_slices<<<(unsigned)std::min<uint64_t>((total + 255) / 256, 148 * 8), 256, 0, s->stream>>>((const double*)part, (double*)res, (int)S, nv, total);
    BT_CHECK_LAUNCH(s);
  }
  BT_CUDA(cudaMemcpyAsync(out_host, res, mat * s->n_batch, cudaMemcpyDeviceToHost, s->stream));
  BT_CUDA(cudaStreamSynchronize(s->stream));
  return BT_OK;
}

// ---- reduced density matrix of a density matrix: rho_K[i][j] = sum_e rho[(i,e)][(j,e)] -------------------------------------
struct DmRdmPlan {
  int k, nenv;
  int kb[16];
  int eb[32];
};

// one warp per entry of rho_K
__global__ void __launch_bounds__(256) k_dm_rdm(const double2* __restrict__ a, int n, const __grid_constant__ DmRdmPlan P, double2* __restrict__ out) {
  const uint64_t o = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const uint64_t D = 1ull << P.k, dim = 1ull << n;
  if (o >= D * D) return;
  const uint64_t i = o & (D - 1), j = o >> P.k;
  const uint64_t roff = deposit_bits(i, P.kb, P.k), coff = deposit_bits(j, P.kb, P.k);
  double re = 0.0, im = 0.0;
  for (uint64_t e = lane; e < (1ull << P.nenv); e += 32) {
    const uint64_t eo = deposit_bits(e, P.eb, P.nenv);
    const double2 x = a[(roff | eo) + (coff | eo) * dim];
    re += x.x; im += x.y;
  }
#pragma unroll
  for (int of = 16; of > 0; of >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, of); im += __shfl_xor_sync(0xffffffffu, im, of); }
  if (lane == 0) out[o] = make_double2(re, im);
}

static int sorted_labels(int n, int k, const int* qubits, int* q) {
  if (!qubits) BT_FAIL(BT_ERR_ARG, "null argument");
  for (int i = 0; i < k; ++i) {
    q[i] = qubits[i];
    if (q[i] < 1 || q[i] > n) BT_FAIL(BT_ERR_ARG, "qubit %d out of range", q[i]);
  }
  std::sort(q, q + k);
  for (int i = 0; i + 1 < k; ++i)
    if (q[i] == q[i + 1]) BT_FAIL(BT_ERR_ARG, "repeated qubit %d", q[i]);
  return BT_OK;
}

extern "C" int bt_dm_rdm(const bt_dm* d, int k, const int* qubits, bt_c64* out) {
  if (!d || !d->v) BT_FAIL(BT_ERR_ARG, "null density matrix");
  if (!out) BT_FAIL(BT_ERR_ARG, "null output");
  const int n = d->n;
  if (k < 0 || k > n || k > 12) BT_FAIL(BT_ERR_UNSUPPORTED, "partial trace of a density matrix keeps up to 12 qubits (asked for %d)", k);
  int q[16];
  BT_TRY(sorted_labels(n, k, qubits, q));
  bt_sv* v = d->v;
  DmRdmPlan P;
  memset(&P, 0, sizeof(P));
  P.k = k; P.nenv = n - k;
  uint64_t used = 0;
  for (int t = 0; t < k; ++t) { P.kb[t] = n - q[k - 1 - t]; used |= 1ull << P.kb[t]; }  // matrix bit 0 <-> largest label
  for (int b = 0, j = 0; b < n; ++b)
    if (!(used >> b & 1)) P.eb[j++] = b;
  const uint64_t DD = 1ull << (2 * k);
  BT_TRY(bt_ensure_scratch(v, DD * sizeof(double2)));
  k_dm_rdm<<<(unsigned)((DD + 7) / 8), 256, 0, v->stream>>>(v->amp, n, P, (double2*)v->d_scratch);
  BT_CHECK_LAUNCH(v);
  BT_CUDA(cudaMemcpyAsync(out, v->d_scratch, DD * sizeof(double2), cudaMemcpyDeviceToHost, v->stream));
  BT_CUDA(cudaStreamSynchronize(v->stream));
  return BT_OK;
}

// singular values of the reduced density matrix of the LAST n_keep qubits (bipartition_trace src/linalg.jl:151-161 keeps the low
// half of the index; entanglement_entropy(rho) src/func.jl:323-328 takes svdvals of it), descending
extern "C" int bt_dm_bipartition_spectrum(const bt_dm* d, int n_keep, double* spec, int* sweeps) {
  if (!d || !d->v) BT_FAIL(BT_ERR_ARG, "null density matrix");
  if (!spec) BT_FAIL(BT_ERR_ARG, "null output");
  const int n = d->n;
  if (n_keep < 1 || n_keep > n || n_keep > 12) BT_FAIL(BT_ERR_UNSUPPORTED, "bipartition spectrum keeps 1..12 qubits");
  bt_sv* v = d->v;
  DmRdmPlan P;
  memset(&P, 0, sizeof(P));
  P.k = n_keep; P.nenv = n - n_keep;
  for (int t = 0; t < n_keep; ++t) P.kb[t] = t;
  for (int j = 0; j < P.nenv; ++j) P.eb[j] = n_keep + j;
  const uint64_t DD = 1ull << (2 * n_keep);
  // the reduced matrix lives behind the spectrum scratch (jacobi_spectrum uses the head of d_scratch)
  const size_t head = ((size_t)((1 << n_keep) + 1) * sizeof(double) + 64 + 255) & ~(size_t)255;  // norms + Frobenius norm + rotation counter
  BT_TRY(bt_ensure_scratch(v, head + DD * sizeof(double2)));
  double2* M = (double2*)((char*)v->d_scratch + head);
  k_dm_rdm<<<(unsigned)((DD + 7) / 8), 256, 0, v->stream>>>(v->amp, n, P, M);
  BT_CHECK_LAUNCH(v);
  // singular values of the (Hermitian) matrix: Jacobi on its columns; their norms (not squared) are the singular values
  const int D = 1 << n_keep;
  int64_t nb = v->n_batch;
  v->n_batch = 1;
  int rc = jacobi_spectrum(v, M, D, D, spec, sweeps);
  v->n_batch = nb;
  if (rc != BT_OK) return rc;
  for (int i = 0; i < D; ++i) spec[i] = sqrt(spec[i]);
  return BT_OK;
}

// ---- expect(x, op) for every Op (src/func.jl:91-92): tr(rho_K O_K) over the qubits K the operator touches -------------------
static int op_labels(int n, int nq, int qubit, int target, int control, int* labels, int* k_out) {
  if (nq != 1 && nq != 2) BT_FAIL(BT_ERR_ARG, "operator must act on 1 or 2 qubits");
  int k = 0;
  labels[k++] = qubit;
  if (nq == 2) labels[k++] = target;
  if (control != -2) labels[k++] = control;
  for (int i = 0; i < k; ++i) {
    if (labels[i] < 1 || labels[i] > n) BT_FAIL(BT_ERR_ARG, "qubit %d out of range", labels[i]);
    for (int j = 0; j < i; ++j)
      if (labels[i] == labels[j]) BT_FAIL(BT_ERR_ARG, "qubit, target and control must differ");
  }
  *k_out = k;
  return BT_OK;
}

// O_K in the ordering of the reduced density matrix (labels ascending, first = most significant bit):
// control == 0 -> identity on the rest, control == 1 -> U (hilbert() src/hilbert.jl:18-159: P0 (x) I + P1 (x) U)
static void build_op_matrix(int k, const int* sorted, int nq, int qubit, int target, int control, const bt_c64* m, cplx* O /* D x D, O[a + b*D] = O[a][b] */) {
  const int D = 1 << k, dim = 1 << nq;
  auto pos = [&](int label) { for (int i = 0; i < k; ++i) if (sorted[i] == label) return k - 1 - i; return -1; };
  const int pq = pos(qubit), pt = nq == 2 ? pos(target) : -1, pc = control != -2 ? pos(control) : -1;
  for (int a = 0; a < D; ++a)
    for (int b = 0; b < D; ++b) {
      cplx v(0.0, 0.0);
      const int ca = pc >= 0 ? (a >> pc) & 1 : 1, cb = pc >= 0 ? (b >> pc) & 1 : 1;
      if (ca == cb) {
        if (ca == 0) v = (a == b) ? cplx(1.0, 0.0) : cplx(0.0, 0.0);
        else {
          const int ua = nq == 2 ? 2 * ((a >> pq) & 1) + ((a >> pt) & 1) : (a >> pq) & 1;
          const int ub = nq == 2 ? 2 * ((b >> pq) & 1) + ((b >> pt) & 1) : (b >> pq) & 1;
          v = c64(m[ua + ub * dim]);
        }
      }
      O[a + b * D] = v;
    }
}

static double trace_against(const bt_c64* rho, const cplx* O, int D) {
  cplx s(0.0, 0.0);
  for (int a = 0; a < D; ++a)
    for (int b = 0; b < D; ++b) s += c64(rho[a + b * D]) * O[b + a * D];  // tr(rho O) = sum rho[a][b] O[b][a]
  return s.real();
}

extern "C" int bt_sv_expect_op(const bt_sv* s, int nq, int qubit, int target, int control, const bt_c64* m, double* out) {
  BT_TRY(bt_check_sv(s));
  if (!m || !out) BT_FAIL(BT_ERR_ARG, "null argument");
  int labels[3], k;
  BT_TRY(op_labels(s->n_qubits, nq, qubit, target, control, labels, &k));
  int sorted[3];
  BT_TRY(sorted_labels(s->n_qubits, k, labels, sorted));
  const int D = 1 << k;
  std::vector<bt_c64> rho((size_t)s->n_batch * D * D);
  BT_TRY(bt_sv_rdm(s, k, labels, rho.data()));
  cplx O[64];
  build_op_matrix(k, sorted, nq, qubit, target, control, m, O);
  for (int64_t t = 0; t < s->n_batch; ++t) out[t] = trace_against(rho.data() + (size_t)t * D * D, O, D);
  return BT_OK;
}

extern "C" int bt_dm_expect_op(const bt_dm* d, int nq, int qubit, int target, int control, const bt_c64* m, double* out) {
  if (!d || !d->v) BT_FAIL(BT_ERR_ARG, "null density matrix");
  if (!m || !out) BT_FAIL(BT_ERR_ARG, "null argument");
  int labels[3], k;
  BT_TRY(op_labels(d->n, nq, qubit, target, control, labels, &k));
  int sorted[3];
  BT_TRY(sorted_labels(d->n, k, labels, sorted));
  const int D = 1 << k;
  bt_c64 rho[64];
  BT_TRY(bt_dm_rdm(d, k, labels, rho));
  cplx O[64];
  build_op_matrix(k, sorted, nq, qubit, target, control, m, O);
  *out = trace_against(rho, O, D);
  return BT_OK;
}

// ---- fidelity(rho, sigma) = (tr sqrt(sqrt(rho) sigma sqrt(rho)))^2  (src/tensor.jl:222-229) -------------------------------------------
// rho = V L V^dagger by one-sided Jacobi on the columns of rho with the rotations accumulated in V (rho V = V L: the converged columns
// have norms L and directions V); the eigenvalues of sqrt(rho) sigma sqrt(rho) are those of C = sqrt(L) (V^dagger sigma V) sqrt(L),
// again by Jacobi.  Two tiled complex GEMMs in between.  Dense O(n^3) like the reference's sqrt(Matrix): registers of <= 11 qubits.

// C = op(A) * B, column-major n x n; CONJA: op(A) = A^dagger.  16 x 16 threads, 2 x 2 outputs each, k-chunks of 32 through shared memory.
template <bool CONJA>
__global__ void __launch_bounds__(256) k_zgemm(const double2* __restrict__ A, const double2* __restrict__ B, double2* __restrict__ Cm, int n) {
  __shared__ double2 As[32][33], Bs[32][33];  // As[i][k] = op(A)(i0+i, k0+k), Bs[k][j] = B(k0+k, j0+j)
  const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double2 acc[2][2] = {{make_double2(0, 0), make_double2(0, 0)}, {make_double2(0, 0), make_double2(0, 0)}};
  for (int k0 = 0; k0 < n; k0 += 32) {
    for (int e = threadIdx.x; e < 1024; e += 256) {
      const int a = e & 31, b = e >> 5;
      double2 va = make_double2(0, 0), vb = make_double2(0, 0);
      if (CONJA) {  // op(A)(i, k) = conj(A(k, i)): read A(k0+a, i0+b) (coalesced in a)
        if (k0 + a < n && i0 + b < n) { va = A[(size_t)(k0 + a) + (size_t)(i0 + b) * n]; va.y = -va.y; }
        As[b][a] = va;
      } else {      // op(A)(i, k) = A(i0+a, k0+b)
        if (i0 + a < n && k0 + b < n) va = A[(size_t)(i0 + a) + (size_t)(k0 + b) * n];
        As[a][b] = va;
      }
      if (k0 + a < n && j0 + b < n) vb = B[(size_t)(k0 + a) + (size_t)(j0 + b) * n];
      Bs[a][b] = vb;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      double2 av[2] = {As[ty][k], As[ty + 16][k]}, bv[2] = {Bs[k][tx], Bs[k][tx + 16]};
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          acc[u][v].x = fma(av[u].x, bv[v].x, fma(-av[u].y, bv[v].y, acc[u][v].x));
          acc[u][v].y = fma(av[u].x, bv[v].y, fma(av[u].y, bv[v].x, acc[u][v].y));
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int i = i0 + ty + 16 * u, j = j0 + tx + 16 * v;
      if (i < n && j < n) Cm[(size_t)i + (size_t)j * n] = acc[u][v];
    }
}

__global__ void k_set_identity(double2* __restrict__ V, int n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n * n; i += (size_t)gridDim.x * blockDim.x)
    V[i] = make_double2((i % n) == (i / n) ? 1.0 : 0.0, 0.0);
}

// C(i, j) *= sqrt(max(l_i, 0)) * sqrt(max(l_j, 0))
__global__ void k_scale_sqrt(double2* __restrict__ Cm, const double* __restrict__ lam, int n) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)n * n; e += (size_t)gridDim.x * blockDim.x) {
    const double f = sqrt(fmax(lam[e % n], 0.0)) * sqrt(fmax(lam[e / n], 0.0));
    Cm[e].x *= f; Cm[e].y *= f;
  }
}

// squared column norms -> norms (the eigenvalues of a positive semi-definite matrix after the Jacobi iteration on its columns)
__global__ void k_sqrt_inplace(double* __restrict__ x, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = sqrt(fmax(x[i], 0.0));
}

extern "C" int bt_dm_fidelity(const bt_dm* rho, const bt_dm* sigma, double* out) {
  if (!rho || !rho->v || !sigma || !sigma->v) BT_FAIL(BT_ERR_ARG, "null density matrix");
  if (!out) BT_FAIL(BT_ERR_ARG, "null output");
  if (rho->n != sigma->n) BT_FAIL(BT_ERR_ARG, "fidelity: registers of %d and %d qubits", rho->n, sigma->n);
  if (rho->v->device != sigma->v->device) BT_FAIL(BT_ERR_ARG, "fidelity: the two density matrices live on different devices");
  if (rho->n > 11) BT_FAIL(BT_ERR_UNSUPPORTED, "fidelity(rho, sigma) is dense O(8^N) linear algebra: up to 11 qubits on the device (asked for %d)", rho->n);
  bt_sv* v = rho->v;
  const int n = 1 << rho->n;
  const size_t nn = (size_t)n * n;
  BT_CUDA(cudaStreamSynchronize(sigma->v->stream));
  double2* W = nullptr;  // A | V | T | C
  BT_CUDA(cudaMalloc(&W, 4 * nn * sizeof(double2)));
  double2 *A = W, *V = W + nn, *T = W + 2 * nn, *Cm = W + 3 * nn;
  int rc = BT_OK;
  int64_t nb = v->n_batch;
  v->n_batch = 1;
  std::vector<double> lam((size_t)n), mu((size_t)n);
  do {
    if (cudaMemcpyAsync(A, v->amp, nn * sizeof(double2), cudaMemcpyDeviceToDevice, v->stream) != cudaSuccess) { rc = BT_ERR_CUDA; break; }
    k_set_identity<<<148 * 4, 256, 0, v->stream>>>(V, n);
    // eigen-decomposition of rho: columns of A converge to l_p v_p, V accumulates the eigenvectors; spectrum unsorted (index-aligned with V)
    if (n >= 2) rc = jacobi_spectrum(v, A, n, n, lam.data(), nullptr, V, n, false);
    else { bt_c64 r00; rc = bt_dm_trace(rho, &r00); lam[0] = r00.re * r00.re; }
    if (rc != BT_OK) break;
    // lam holds squared column norms = l^2: the device copy in d_scratch (head) is reused as l after a square root
    double* d_lam = (double*)v->d_scratch;
    k_sqrt_inplace<<<(n + 127) / 128, 128, 0, v->stream>>>(d_lam, n);
    const dim3 g((n + 31) / 32, (n + 31) / 32);
    k_zgemm<false><<<g, 256, 0, v->stream>>>(sigma->v->amp, V, T, n);   // T = sigma V
    k_zgemm<true><<<g, 256, 0, v->stream>>>(V, T, Cm, n);               // C = V^dagger sigma V
    k_scale_sqrt<<<148 * 4, 256, 0, v->stream>>>(Cm, d_lam, n);         // C = sqrt(L) C sqrt(L)
    if (cudaPeekAtLastError() != cudaSuccess) { rc = BT_ERR_CUDA; break; }
    v->launches += 4;
    if (n >= 2) rc = jacobi_spectrum(v, Cm, n, n, mu.data(), nullptr);
    else {
      double2 c00;
      if (cudaMemcpyAsync(&c00, Cm, sizeof(double2), cudaMemcpyDeviceToHost, v->stream) != cudaSuccess || cudaStreamSynchronize(v->stream) != cudaSuccess) { rc = BT_ERR_CUDA; break; }
      mu[0] = c00.x * c00.x;
    }
    if (rc != BT_OK) break;
    double tr = 0.0;
    for (int i = 0; i < n; ++i) tr += sqrt(sqrt(std::max(mu[(size_t)i], 0.0)));  // mu = squared singular values of C = eigenvalues^2; sqrt(eigenvalue)
    *out = tr * tr;
  } while (0);
  v->n_batch = nb;
  cudaStreamSynchronize(v->stream);
  cudaFree(W);
  if (rc == BT_ERR_CUDA) BT_FAIL(BT_ERR_CUDA, "CUDA error in bt_dm_fidelity: %s", cudaGetErrorString(cudaGetLastError()));
  return rc;
}
