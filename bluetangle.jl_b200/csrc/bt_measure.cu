// bt_measure.cu -- Born measurement + collapse, Kraus trajectory step, inverse-CDF sampling.
//
// Replaces born_measure_Z (src/hilbert.jl:682-696), _reset_Z (:752-759), __calc_prob (src/struct.jl:9-29),
// __QuantumChannel_new_apply (src/struct.jl:31-55), _weighted_sample (src/hilbert.jl:810-819) and sample
// (src/ops.jl:46-62).  All decisions (outcome, Kraus index, normalisation factor) are taken on the device from
// caller-supplied uniforms, per trajectory, so a batch of trajectories advances without host round trips.
#include "bt_internal.cuh"

int bt_reduce_rdm_at(const bt_sv* cs, int k, const int* tb, size_t res_off);
int bt_prepare_local_bits(bt_sv* s, int n, const int* logical_bits);

// ---- measurement ----------------------------------------------------------------------------------------------
// res: n_batch x 4 packed 2x2 RDMs ([0] = p0, [3] = p1).  outcome = (u < p0) ? 0 : 1  (src/hilbert.jl:693)
__global__ void k_decide_measure(const double* __restrict__ res, const double* __restrict__ u, int32_t* __restrict__ outcome,
                                 double* __restrict__ scale, int64_t n_batch, const int32_t* __restrict__ mask) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_batch) return;
  double p0 = res[4 * t], p1 = res[4 * t + 3];
  if (mask && !mask[t]) {  // trajectory not in the active branch: untouched, outcome -1
    outcome[t] = -1;
    scale[2 * t] = 1.0;
    scale[2 * t + 1] = p0;
    return;
  }
  int ind = (u[t] < p0) ? 0 : 1;
  outcome[t] = ind;
  scale[2 * t] = 1.0 / sqrt(ind == 0 ? p0 : p1);  // normalize(P_ind * state): divide by the actual norm
  scale[2 * t + 1] = p0;
}

// pair (i0 = bit clear, i1 = bit set): keep the measured half scaled, zero the other; reset moves |1> to |0>.
__global__ void __launch_bounds__(256) k_collapse(double2* __restrict__ a, int n_local, int bit, uint64_t npairs,
                                                   const int32_t* __restrict__ outcome, const double* __restrict__ scale, int reset) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const double2 zero = make_double2(0.0, 0.0);
  for (; g < npairs; g += stride) {
    uint64_t t = g >> (n_local - 1);
    uint64_t i0 = ((g >> bit) << (bit + 1)) | (g & ((1ull << bit) - 1));
    uint64_t i1 = i0 | (1ull << bit);
    int ind = outcome[t];
    double sc = scale[2 * t];
    if (ind < 0) continue;
    if (ind == 0) {
      double2 x = a[i0];
      a[i0] = make_double2(x.x * sc, x.y * sc);
      a[i1] = zero;
    } else {
      double2 x = a[i1];
      x = make_double2(x.x * sc, x.y * sc);
      if (reset) { a[i0] = x; a[i1] = zero; }
      else { a[i1] = x; a[i0] = zero; }
    }
  }
}

extern "C" int bt_sv_measure_z(bt_sv* s, int qubit, const double* u, int32_t* outcome, double* p0, int reset) {
  BT_TRY(bt_check_sv(s));
  if (!u) BT_FAIL(BT_ERR_ARG, "null uniforms");
  if (qubit < 1 || qubit > s->n_qubits) BT_FAIL(BT_ERR_ARG, "N must be larger than qubit");
  BT_TRY(bt_ensure_traj(s));
  int lb = s->n_qubits - qubit;
  if (s->world > 1) BT_TRY(bt_prepare_local_bits(s, 1, &lb));
  int bit = s->phys_of_bit[lb];
  BT_CUDA(cudaMemcpyAsync(s->d_u, u, s->n_batch * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  int tb[1] = {bit};
  BT_TRY(bt_reduce_rdm(s, 1, tb));
  if (s->world > 1) {
    BT_TRY(bt_results_to_host(s, (size_t)s->n_batch * 4));
    if (!s->allreduce) BT_FAIL(BT_ERR_ARG, "sharded state: set an all-reduce callback first");
    s->allreduce(s->allreduce_ctx, s->h_res, (int)(s->n_batch * 4));
    BT_CUDA(cudaMemcpyAsync(s->d_res, s->h_res, s->n_batch * 4 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  }
  k_decide_measure<<<(unsigned)((s->n_batch + 127) / 128), 128, 0, s->stream>>>(s->d_res, s->d_u, s->d_outcome, s->d_scale, s->n_batch, s->mask_on ? s->d_mask : nullptr);
  BT_CHECK_LAUNCH(s);
  uint64_t npairs = s->len >> 1;
  unsigned grid = (unsigned)std::min<uint64_t>((npairs + 255) / 256, 148ull * 32);
  k_collapse<<<grid, 256, 0, s->stream>>>(s->amp, s->n_local, bit, npairs, s->d_outcome, s->d_scale, reset);
  BT_CHECK_LAUNCH(s);
  if (outcome || p0) {
    if (outcome) BT_CUDA(cudaMemcpyAsync(outcome, s->d_outcome, s->n_batch * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
    if (p0) {
      BT_CUDA(cudaMemcpyAsync(s->h_res, s->d_scale, s->n_batch * 2 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    }
    BT_CUDA(cudaStreamSynchronize(s->stream));
    if (p0) for (int64_t t = 0; t < s->n_batch; ++t) p0[t] = s->h_res[2 * t + 1];
  }
  return BT_OK;
}

// ---- several measurements in one read pass + one collapse pass -----------------------------------------------------------------
int bt_reduce_joint_probs(const bt_sv* cs, int k, const int* tb);  // bt_reduce.cu

// One thread per trajectory: the k outcomes in order.  Measurement j sees the state collapsed and renormalised by measurements
// 0..j-1 (born_measure_Z normalises after every projection, src/hilbert.jl:694): p0 = mass(pattern so far, bit j = 0) / mass(pattern
// so far); the first one uses the raw mass of the state as it is, like the single call.
__global__ void k_decide_multi(const double* __restrict__ res, int k, const double* __restrict__ u, int32_t* __restrict__ pattern_out, int32_t* __restrict__ outcomes,
                               double* __restrict__ scale, int64_t n_batch, const int32_t* __restrict__ mask) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_batch) return;
  const int D = 1 << k;
  const double* p = res + t * D;
  if (mask && !mask[t]) {
    pattern_out[t] = -1;
    scale[2 * t] = 1.0; scale[2 * t + 1] = 0.0;
    for (int j = 0; j < k; ++j) outcomes[t * k + j] = -1;
    return;
  }
  int pattern = 0, decided = 0;
  for (int j = 0; j < k; ++j) {
    double tot = 0.0, m0 = 0.0;
    for (int b = 0; b < D; ++b)
      if ((b & decided) == pattern) {
        tot += p[b];
        if (!((b >> j) & 1)) m0 += p[b];
      }
    const double p0 = j == 0 ? m0 : m0 / tot;
    const int ind = (u[t * k + j] < p0) ? 0 : 1;
    outcomes[t * k + j] = ind;
    pattern |= ind << j;
    decided |= 1 << j;
  }
  pattern_out[t] = pattern;
  scale[2 * t] = 1.0 / sqrt(p[pattern]);
  scale[2 * t + 1] = p[pattern];
}

struct MultiPlan {
  int k, ni;
  int ins[4];
  uint64_t off[16];
  uint32_t reset_mask;  // bit j set: measurement j resets its qubit (|1> -> |0>)
};

// one thread per group of 2^k amplitudes: the member selected by the trajectory's outcome pattern survives (scaled), the others
// become zero; reset bits move the survivor to the member with those bits cleared
__global__ void __launch_bounds__(256) k_collapse_multi(double2* __restrict__ a, int n_local, uint64_t ngroups_total, const __grid_constant__ MultiPlan P,
                                                         const int32_t* __restrict__ pattern, const double* __restrict__ scale) {
  const int D = 1 << P.k;
  const double2 zero = make_double2(0.0, 0.0);
  for (uint64_t G = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; G < ngroups_total; G += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t t = G >> (n_local - P.k);
    const int pat = pattern[t];
    if (pat < 0) continue;
    uint64_t g = G;
    for (int i = 0; i < P.ni; ++i) {
      const int b = P.ins[i];
      g = ((g >> b) << (b + 1)) | (g & ((1ull << b) - 1));
    }
    const double sc = scale[2 * t];
    double2 x = a[g + P.off[pat]];
    x = make_double2(x.x * sc, x.y * sc);
    const int dst = pat & ~(int)P.reset_mask;
    for (int j = 0; j < D; ++j) a[g + P.off[j]] = (j == dst) ? x : zero;
  }
}

extern "C" int bt_sv_measure_z_multi(bt_sv* s, int k, const int* qubits, const double* u, int32_t* outcomes, const int* reset) {
  BT_TRY(bt_check_sv(s));
  if (!u || !qubits) BT_FAIL(BT_ERR_ARG, "null argument");
  if (k < 1 || k > 4) BT_FAIL(BT_ERR_ARG, "bt_sv_measure_z_multi takes 1..4 qubits per call");
  int lbs[4];
  for (int j = 0; j < k; ++j) {
    if (qubits[j] < 1 || qubits[j] > s->n_qubits) BT_FAIL(BT_ERR_ARG, "N must be larger than qubit");
    for (int i = 0; i < j; ++i)
      if (qubits[i] == qubits[j]) BT_FAIL(BT_ERR_ARG, "repeated qubit %d: measure it in separate calls", qubits[j]);
    lbs[j] = s->n_qubits - qubits[j];
  }
  BT_TRY(bt_ensure_traj(s));
  if (s->world > 1) BT_TRY(bt_prepare_local_bits(s, k, lbs));
  int tb[4];
  for (int j = 0; j < k; ++j) tb[j] = s->phys_of_bit[lbs[j]];
  const int D = 1 << k;
  BT_TRY(bt_reduce_joint_probs(s, k, tb));  // may use the head of d_scratch (k == 4): the buffers below come after it
  if (s->world > 1) {
    BT_TRY(bt_results_to_host(s, (size_t)s->n_batch * D));
    if (!s->allreduce) BT_FAIL(BT_ERR_ARG, "sharded state: set an all-reduce callback first");
    s->allreduce(s->allreduce_ctx, s->h_res, (int)(s->n_batch * D));
    BT_CUDA(cudaMemcpyAsync(s->d_res, s->h_res, s->n_batch * D * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  }
  const size_t off_u = 256, off_o = off_u + (((size_t)s->n_batch * k * sizeof(double) + 255) / 256) * 256;
  BT_TRY(bt_ensure_scratch(s, off_o + (size_t)s->n_batch * k * sizeof(int32_t)));
  double* d_u = (double*)((char*)s->d_scratch + off_u);
  int32_t* d_out = (int32_t*)((char*)s->d_scratch + off_o);
  const bool to_log = s->mlog_on && !outcomes;
  if (to_log) {  // the outcomes of this call are appended to the device-side log (read once at the end: bt_sv_measure_log_read)
    const size_t need = s->mlog_len + (size_t)s->n_batch * k;
    if (need > s->mlog_cap) {
      size_t cap = std::max<size_t>(need * 2, (size_t)s->n_batch * 256);
      int32_t* nl = nullptr;
      BT_CUDA(cudaMalloc(&nl, cap * sizeof(int32_t)));
      if (s->d_mlog) {
        BT_CUDA(cudaMemcpyAsync(nl, s->d_mlog, s->mlog_len * sizeof(int32_t), cudaMemcpyDeviceToDevice, s->stream));
        BT_CUDA(cudaStreamSynchronize(s->stream));
        BT_CUDA(cudaFree(s->d_mlog));
      }
      s->d_mlog = nl; s->mlog_cap = cap;
    }
    d_out = s->d_mlog + s->mlog_len;
    s->mlog_len = need;
  }
  BT_CUDA(cudaMemcpyAsync(d_u, u, (size_t)s->n_batch * k * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  k_decide_multi<<<(unsigned)((s->n_batch + 127) / 128), 128, 0, s->stream>>>(s->d_res, k, d_u, s->d_outcome, d_out, s->d_scale, s->n_batch, s->mask_on ? s->d_mask : nullptr);
  BT_CHECK_LAUNCH(s);
  MultiPlan P;
  memset(&P, 0, sizeof(P));
  P.k = k; P.ni = k;
  int sorted[4];
  for (int j = 0; j < k; ++j) sorted[j] = tb[j];
  std::sort(sorted, sorted + k);
  for (int j = 0; j < k; ++j) P.ins[j] = sorted[j];
  for (int j = 0; j < D; ++j) {
    uint64_t o = 0;
    for (int i = 0; i < k; ++i)
      if ((j >> i) & 1) o |= 1ull << tb[i];
    P.off[j] = o;
  }
  for (int j = 0; j < k; ++j)
    if (reset && reset[j]) P.reset_mask |= 1u << j;
  const uint64_t ngroups = s->len >> k;
  const unsigned grid = (unsigned)std::min<uint64_t>((ngroups + 255) / 256, 148ull * 32);
  k_collapse_multi<<<grid, 256, 0, s->stream>>>(s->amp, s->n_local, ngroups, P, s->d_outcome, s->d_scale);
  BT_CHECK_LAUNCH(s);
  if (outcomes) {
    BT_CUDA(cudaMemcpyAsync(outcomes, d_out, (size_t)s->n_batch * k * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
    BT_CUDA(cudaStreamSynchronize(s->stream));
  }
  return BT_OK;
}

extern "C" int bt_sv_measure_log(bt_sv* s, int on) {
  BT_TRY(bt_check_sv(s));
  s->mlog_on = on != 0;
  s->mlog_len = 0;
  return BT_OK;
}

extern "C" int bt_sv_measure_log_read(bt_sv* s, int32_t* out, uint64_t cap, uint64_t* n) {
  BT_TRY(bt_check_sv(s));
  if (n) *n = (uint64_t)s->mlog_len;
  if (s->mlog_len == 0) return BT_OK;
  if (!out || cap < s->mlog_len) BT_FAIL(BT_ERR_ARG, "outcome log holds %llu entries", (unsigned long long)s->mlog_len);
  BT_CUDA(cudaMemcpyAsync(out, s->d_mlog, s->mlog_len * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
  BT_CUDA(cudaStreamSynchronize(s->stream));
  s->mlog_len = 0;
  return BT_OK;
}

extern "C" int bt_sv_outcomes(const bt_sv* s, int32_t* outcome) {
  BT_TRY(bt_check_sv(s));
  if (!outcome) BT_FAIL(BT_ERR_ARG, "null output");
  if (!s->d_outcome) BT_FAIL(BT_ERR_ARG, "no measurement has been made on this handle");
  BT_CUDA(cudaMemcpyAsync(outcome, s->d_outcome, s->n_batch * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
  BT_CUDA(cudaStreamSynchronize(s->stream));
  return BT_OK;
}

// ---- Kraus trajectory step ---------------------------------------------------------------------------------------
// One block per trajectory.  rho (packed, D*D doubles) is the RDM in the reference's *probability* ordering;
// `swap` != 0 means the operator is applied in the other ordering (2-qubit channel with qubit > target,
// SURVEY App. A.5 #2): probabilities use rho as is, the normalisation uses the bit-swapped rho.
// K: nK row-major DxD matrices.  mats out: row-major DxD (stride 64), scaled by 1/||K_ind psi||.
template <int D>
__device__ double kraus_weight(const double2* __restrict__ K, const double2* rho) {
  // Re tr(K rho K') = sum_a sum_b sum_c Re( K[a][b] rho[b][c] conj(K[a][c]) )
  double tot = 0.0;
  for (int a = 0; a < D; ++a)
    for (int b = 0; b < D; ++b) {
      double2 kab = K[a * D + b];
      if (kab.x == 0.0 && kab.y == 0.0) continue;
      for (int c = 0; c < D; ++c) {
        double2 kac = K[a * D + c];
        double2 r = rho[b * D + c];
        // kab * r
        double xr = kab.x * r.x - kab.y * r.y, xi = kab.x * r.y + kab.y * r.x;
        // * conj(kac), real part
        tot += xr * kac.x + xi * kac.y;
      }
    }
  return tot;
}

template <int D>
__global__ void __launch_bounds__(64) k_decide_kraus(const double* __restrict__ res, const double* __restrict__ u,
                                                      const double2* __restrict__ K, int nK, int swap,
                                                      int32_t* __restrict__ chosen, double2* __restrict__ mats,
                                                      double* __restrict__ probs_out, int32_t* __restrict__ err, const int32_t* __restrict__ mask) {
  if (mask && u && !mask[blockIdx.x]) {  // trajectory not in the active branch: no draw, no operator (the apply kernel skips it too)
    if (threadIdx.x == 0) chosen[blockIdx.x] = -1;
    return;
  }
  __shared__ double2 rho[D * D];
  __shared__ double2 rho_s[D * D];
  __shared__ double probs[64];
  __shared__ int s_ind;
  __shared__ double s_scale;
  const int64_t t = blockIdx.x;
  const double* v = res + t * (D * D);
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
    int a = i / D, b = i % D;
    double2 r;
    if (a == b) r = make_double2(v[a * D + a], 0.0);
    else if (a < b) r = make_double2(v[a * D + b], v[b * D + a]);
    else r = make_double2(v[b * D + a], -v[a * D + b]);
    rho[i] = r;
  }
  __syncthreads();
  if (D == 4) {
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
      int a = i / D, b = i % D;
      int a2 = ((a & 1) << 1) | (a >> 1), b2 = ((b & 1) << 1) | (b >> 1);
      rho_s[i] = swap ? rho[a2 * D + b2] : rho[i];
    }
  } else {
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) rho_s[i] = rho[i];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nK; k += blockDim.x) {
    probs[k] = kraus_weight<D>(K + (size_t)k * D * D, rho);
    if (probs_out) probs_out[t * nK + k] = probs[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int ind = -1;
    if (u) {
      double rval = u[t], cum = 0.0;
      for (int k = 0; k < nK; ++k) {  // src/hilbert.jl:814-817: first i with rval <= cumsum[i]
        cum += probs[k];
        if (rval <= cum) { ind = k; break; }
      }
      if (ind < 0) { atomicExch(err, 1); ind = nK - 1; }
      double w = kraus_weight<D>(K + (size_t)ind * D * D, rho_s);
      s_scale = 1.0 / sqrt(w);
      chosen[t] = ind;
    }
    s_ind = ind;
  }
  __syncthreads();
  if (u) {
    const double2* Ks = K + (size_t)s_ind * D * D;
    for (int i = threadIdx.x; i < D * D; i += blockDim.x)
      mats[t * 64 + i] = make_double2(Ks[i].x * s_scale, Ks[i].y * s_scale);
  }
}

static int kraus_common(bt_sv* s, int nq, int qubit, int target, const bt_c64* K, int nK, const double* u,
                        int32_t* chosen, double* probs_host) {
  if (!K) BT_FAIL(BT_ERR_ARG, "null Kraus operators");
  if (nK < 1 || nK > 64) BT_FAIL(BT_ERR_ARG, "number of Kraus operators must be 1..64");
  int N = s->n_qubits;
  int D = 1 << nq;
  int lb_apply[3], lb_prob[3];
  int swap = 0;
  if (nq == 1) {
    if (qubit < 1 || qubit > N) BT_FAIL(BT_ERR_ARG, "N must be larger than qubit");
    lb_apply[0] = lb_prob[0] = N - qubit;
  } else if (nq == 2) {
    if (qubit < 1 || target < 1 || qubit > N || target > N) BT_FAIL(BT_ERR_ARG, "N must be larger than qubits");
    if (qubit == target) BT_FAIL(BT_ERR_ARG, "`qubit` and `target_qubit` must differ");
    lb_apply[0] = N - target; lb_apply[1] = N - qubit;            // K indexed 2*b_qubit + b_target
    int lo = std::min(qubit, target), hi = std::max(qubit, target);
    lb_prob[0] = N - hi; lb_prob[1] = N - lo;                      // rho_A indexed 2*b_min + b_max
    swap = qubit > target;
  } else if (nq == 3) {
    if (qubit < 1 || qubit + 2 > N) BT_FAIL(BT_ERR_ARG, "N must be larger than all three qubits");
    for (int i = 0; i < 3; ++i) lb_apply[i] = lb_prob[i] = N - (qubit + 2 - i);
  } else {
    BT_FAIL(BT_ERR_ARG, "Noise models are available only for up to 3 qubits!");
  }
  BT_TRY(bt_ensure_traj(s));
  if (s->world > 1) BT_TRY(bt_prepare_local_bits(s, nq, lb_apply));
  int tb_apply[3], tb_prob[3];
  for (int i = 0; i < nq; ++i) { tb_apply[i] = s->phys_of_bit[lb_apply[i]]; tb_prob[i] = s->phys_of_bit[lb_prob[i]]; }
  // Kraus table -> device (row-major), staged through the scratch area behind d_mats' sibling buffer
  std::vector<double2> hk((size_t)nK * D * D);
  for (int k = 0; k < nK; ++k)
    for (int r = 0; r < D; ++r)
      for (int c = 0; c < D; ++c) {
        const bt_c64& z = K[(size_t)k * D * D + r + c * D];
        hk[(size_t)k * D * D + r * D + c] = make_double2(z.re, z.im);
      }
  double2* d_K = nullptr;
  BT_CUDA(cudaMallocAsync(&d_K, hk.size() * sizeof(double2), s->stream));
  BT_CUDA(cudaMemcpyAsync(d_K, hk.data(), hk.size() * sizeof(double2), cudaMemcpyHostToDevice, s->stream));
  BT_CUDA(cudaStreamSynchronize(s->stream));  // hk is pageable host memory: finish the copy before it goes away
  if (u) BT_CUDA(cudaMemcpyAsync(s->d_u, u, s->n_batch * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  BT_TRY(bt_reduce_rdm(s, nq, tb_prob));
  if (s->world > 1) {
    BT_TRY(bt_results_to_host(s, (size_t)s->n_batch * D * D));
    if (!s->allreduce) BT_FAIL(BT_ERR_ARG, "sharded state: set an all-reduce callback first");
    s->allreduce(s->allreduce_ctx, s->h_res, (int)(s->n_batch * D * D));
    BT_CUDA(cudaMemcpyAsync(s->d_res, s->h_res, s->n_batch * D * D * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  }
  double* d_probs = nullptr;
  if (probs_host) BT_CUDA(cudaMallocAsync(&d_probs, (size_t)s->n_batch * nK * sizeof(double), s->stream));
  const double* du = u ? s->d_u : nullptr;
  const int32_t* dmask = s->mask_on ? s->d_mask : nullptr;
  if (nq == 1) k_decide_kraus<2><<<(unsigned)s->n_batch, 64, 0, s->stream>>>(s->d_res, du, d_K, nK, 0, s->d_outcome, s->d_mats, d_probs, s->d_err, dmask);
  else if (nq == 2) k_decide_kraus<4><<<(unsigned)s->n_batch, 64, 0, s->stream>>>(s->d_res, du, d_K, nK, swap, s->d_outcome, s->d_mats, d_probs, s->d_err, dmask);
  else k_decide_kraus<8><<<(unsigned)s->n_batch, 64, 0, s->stream>>>(s->d_res, du, d_K, nK, 0, s->d_outcome, s->d_mats, d_probs, s->d_err, dmask);
  BT_CHECK_LAUNCH(s);
  BT_CUDA(cudaFreeAsync(d_K, s->stream));
  if (u) BT_TRY(bt_launch_gate_devmat(s, nq, tb_apply, s->d_mats));
  if (probs_host) {
    BT_CUDA(cudaMemcpyAsync(probs_host, d_probs, (size_t)s->n_batch * nK * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    BT_CUDA(cudaFreeAsync(d_probs, s->stream));
    BT_CUDA(cudaStreamSynchronize(s->stream));
  }
  if (chosen) {
    BT_CUDA(cudaMemcpyAsync(chosen, s->d_outcome, s->n_batch * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
    BT_CUDA(cudaMemcpyAsync(s->h_flag, s->d_err, sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
    BT_CUDA(cudaStreamSynchronize(s->stream));
    if (s->h_flag[0]) {
      BT_CUDA(cudaMemsetAsync(s->d_err, 0, sizeof(int32_t), s->stream));
      BT_FAIL(BT_ERR_SAMPLE, "_weighted_sample: uniform draw exceeds the cumulative Kraus probabilities (reference returns nothing)");
    }
  }
  return BT_OK;
}

extern "C" int bt_sv_kraus(bt_sv* s, int nq, int qubit, int target, const bt_c64* K, int nK, const double* u, int32_t* chosen) {
  BT_TRY(bt_check_sv(s));
  if (!u) BT_FAIL(BT_ERR_ARG, "null uniforms");
  return kraus_common(s, nq, qubit, target, K, nK, u, chosen, nullptr);
}

extern "C" int bt_sv_kraus_probs(const bt_sv* s, int nq, int qubit, int target, const bt_c64* K, int nK, double* probs) {
  BT_TRY(bt_check_sv(s));
  if (!probs) BT_FAIL(BT_ERR_ARG, "null output");
  return kraus_common(const_cast<bt_sv*>(s), nq, qubit, target, K, nK, nullptr, nullptr, probs);
}

// ---- sampling ------------------------------------------------------------------------------------------------------
// Two-level inverse CDF: (1) sums of p over blocks of 2^SB entries, (2) inclusive scan of the block sums,
// (3) one warp per shot: binary search over the block prefix, then a warp scan inside the block.
// p(i) = |a[i]|^2 for a state vector, Re a[i*(dim+1)] for the diagonal of a density matrix.
#define SB 12

__device__ __forceinline__ double prob_at(const double2* __restrict__ a, uint64_t i, uint64_t dm_stride) {
  if (dm_stride) return a[i * dm_stride].x;
  double2 x = a[i];
  return x.x * x.x + x.y * x.y;
}

// All three kernels take a batch of `ntraj` independent distributions of n entries each (trajectory t at a + t * n * stride-free
// layout: consecutive state vectors); a single state is ntraj == 1.
#define SUBS (1 << (SB - 8))  // sub-blocks of 256 entries per block
__global__ void __launch_bounds__(256) k_block_sums(const double2* __restrict__ a, uint64_t n, uint64_t dm_stride, double* __restrict__ bsum, double* __restrict__ ssum,
                                                     uint64_t nb) {
  // one CTA per block of 2^SB entries; also the sums of its 16 sub-blocks of 256 entries (the in-block search of k_sample starts
  // from them).  Warp w owns the 512 consecutive entries of sub-blocks 2w and 2w+1: 8 independent loads per lane and one shuffle
  // tree per sub-block; fixed order => reproducible
  __shared__ double sub[SUBS];
  const uint64_t traj = blockIdx.x / nb, blk = blockIdx.x % nb;
  const double2* __restrict__ base = a + traj * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t w0 = (blk << SB) + (uint64_t)warp * 512 + lane;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    double v[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const uint64_t i = w0 + (uint64_t)h * 256 + (uint64_t)it * 32;
      v[it] = (i < n) ? prob_at(base, i, dm_stride) : 0.0;
    }
    double acc = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) sub[warp * 2 + h] = acc;
  }
  __syncthreads();
  if (threadIdx.x < SUBS) ssum[(uint64_t)blockIdx.x * SUBS + threadIdx.x] = sub[threadIdx.x];
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < SUBS; ++k) t += sub[k];
    bsum[blockIdx.x] = t;
  }
}

// inclusive scan of each trajectory's block sums: one CTA per trajectory (sequential carry over tiles of 1024)
__global__ void __launch_bounds__(1024) k_scan_inclusive(double* __restrict__ xs, uint64_t n) {
  __shared__ double warp_tot[32];
  __shared__ double carry_s;
  double* __restrict__ x = xs + (uint64_t)blockIdx.x * n;
  if (threadIdx.x == 0) carry_s = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint64_t base = 0; base < n; base += 4096) {  // a thread owns 4 consecutive entries
    const uint64_t i0 = base + (uint64_t)threadIdx.x * 4;
    double e[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) e[k] = (i0 + k < n) ? x[i0 + k] : 0.0;
    e[1] += e[0]; e[2] += e[1]; e[3] += e[2];
    double v = e[3];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double y = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += y;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    if (warp == 0) {
      double w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        double y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_tot[lane] = w;
    }
    __syncthreads();
    const double off = carry_s + (warp > 0 ? warp_tot[warp - 1] : 0.0) + (v - e[3]);  // everything before this thread's 4 entries
#pragma unroll
    for (int k = 0; k < 4; ++k) if (i0 + k < n) x[i0 + k] = off + e[k];
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = off + e[3];
    __syncthreads();
  }
}

// Sharded sampling: every rank holds the same inclusive prefix of the per-rank totals (same doubles, same order), so ownership of a
// shot is decided identically everywhere: owner = first rank r with rank_prefix[r] >= target (the last rank takes what rounding
// pushes past the end).  world == 0: unsharded, the target is u * local total.
struct ShardCdf {
  int world, rank;
  double prefix[16];  // inclusive prefix of the rank totals, rank order = physical index order
};

// pre: prefix = inclusive scan of the block sums of each trajectory (nb entries each); one warp per shot: binary search over the
// block prefix, then a warp scan inside the block.  out = -1 for a shot another rank owns.
__global__ void __launch_bounds__(128) k_sample(const double2* __restrict__ a, uint64_t n, uint64_t dm_stride,
                                                 const double* __restrict__ prefixes, const double* __restrict__ ssum, uint64_t nb,
                                                 const double* __restrict__ u, uint64_t shots_per_traj, uint64_t shots, const __grid_constant__ ShardCdf sh,
                                                 int64_t* __restrict__ out) {
  const uint64_t shot = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (shot >= shots) return;
  const uint64_t traj = shot / shots_per_traj;
  const double* __restrict__ prefix = prefixes + traj * nb;
  const double2* __restrict__ base_a = a + traj * n;
  const double local_total = prefix[nb - 1];
  double t;
  if (sh.world > 1) {
    const double target = u[shot] * sh.prefix[sh.world - 1];
    int owner = sh.world - 1;
    for (int r = 0; r < sh.world - 1; ++r)
      if (sh.prefix[r] >= target) { owner = r; break; }
    if (owner != sh.rank) { if (lane == 0) out[shot] = -1; return; }
    t = target - (owner > 0 ? sh.prefix[owner - 1] : 0.0);
  } else {
    t = u[shot] * local_total;
  }
  // binary search: first block j with prefix[j] >= t
  uint64_t lo = 0, hi = nb - 1;
  while (lo < hi) {
    uint64_t mid = (lo + hi) >> 1;
    if (prefix[mid] >= t) hi = mid; else lo = mid + 1;
  }
  const uint64_t j = lo;
  double carry = (j > 0) ? prefix[j - 1] : 0.0;
  const uint64_t b0 = j << SB;
  const uint64_t bend = (b0 + (1ull << SB) < n) ? b0 + (1ull << SB) : n;
  // sub-block (256 entries) that holds the target: warp scan over the block's 16 sub-block sums; the entries before it are skipped
  uint64_t start = b0;
  {
    double v = (lane < SUBS) ? ssum[(traj * nb + j) * SUBS + lane] : 0.0;
#pragma unroll
    for (int o = 1; o < SUBS; o <<= 1) {
      double y = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += y;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, lane < SUBS && carry + v >= t);
    // the first sub-block whose inclusive sum reaches the target; none (rounding: the entry-level sums differ in the last bits): start at the last one
    const int sb = hit ? (__ffs(hit) - 1) : (SUBS - 1);
    const double before = __shfl_sync(0xffffffffu, v, sb > 0 ? sb - 1 : 0);
    if (sb > 0) { carry += before; start = b0 + (uint64_t)sb * 256; }
  }
  int64_t found = -1;
  int64_t last_nz = -1;
  for (uint64_t base = start; base < bend && found < 0; base += 32) {
    uint64_t i = base + lane;
    double p = (i < bend) ? prob_at(base_a, i, dm_stride) : 0.0;
    double v = p;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double y = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += y;
    }
    double cum = carry + v;
    unsigned hit = __ballot_sync(0xffffffffu, (i < bend) && (cum >= t));
    unsigned nz = __ballot_sync(0xffffffffu, p != 0.0);
    if (hit) found = (int64_t)(base + (__ffs(hit) - 1));
    else {
      if (nz) last_nz = (int64_t)(base + (31 - __clz(nz)));
      carry = __shfl_sync(0xffffffffu, cum, 31);
    }
  }
  if (found < 0) found = (last_nz >= 0) ? last_nz : (int64_t)(bend - 1);  // rounding fell off the end of the block
  if (lane == 0) out[shot] = found;
}

// ntraj distributions of n entries (consecutive in memory), shots_per_traj uniforms each: three launches for the whole batch
static int sample_impl(bt_sv* s, const double2* base, uint64_t n, uint64_t dm_stride, const double* u, uint64_t shots_per_traj, uint64_t ntraj, int64_t* out) {
  if (!u || !out) BT_FAIL(BT_ERR_ARG, "null argument");
  const uint64_t shots = shots_per_traj * ntraj;
  if (shots == 0) return BT_OK;
  const uint64_t nb = (n + (1ull << SB) - 1) >> SB;
  if (nb * ntraj >= (1ull << 31)) BT_FAIL(BT_ERR_UNSUPPORTED, "sampling: batch too large");
  size_t off_s = ((nb * ntraj * sizeof(double) + 255) / 256) * 256;
  size_t off_u = off_s + ((nb * ntraj * SUBS * sizeof(double) + 255) / 256) * 256;
  size_t off_o = off_u + ((shots * sizeof(double) + 255) / 256) * 256;
  BT_TRY(bt_ensure_scratch(s, off_o + shots * sizeof(int64_t)));
  double* d_prefix = (double*)s->d_scratch;
  double* d_sub = (double*)((char*)s->d_scratch + off_s);
  double* d_u = (double*)((char*)s->d_scratch + off_u);
  int64_t* d_out = (int64_t*)((char*)s->d_scratch + off_o);
  BT_CUDA(cudaMemcpyAsync(d_u, u, shots * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  k_block_sums<<<(unsigned)(nb * ntraj), 256, 0, s->stream>>>(base, n, dm_stride, d_prefix, d_sub, nb);
  BT_CHECK_LAUNCH(s);
  k_scan_inclusive<<<(unsigned)ntraj, 1024, 0, s->stream>>>(d_prefix, nb);
  BT_CHECK_LAUNCH(s);
  ShardCdf sh;
  memset(&sh, 0, sizeof(sh));
  if (s->world > 1) {
    if (ntraj != 1) BT_FAIL(BT_ERR_UNSUPPORTED, "batched sampling of a sharded state");
    if (s->world > 16) BT_FAIL(BT_ERR_UNSUPPORTED, "sharded sampling supports up to 16 ranks");
    BT_CUDA(cudaMemcpyAsync(s->h_res, d_prefix + (nb - 1), sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    BT_CUDA(cudaStreamSynchronize(s->stream));
    if (!s->allreduce) BT_FAIL(BT_ERR_ARG, "sharded state: set an all-reduce callback first");
    // the totals are all-gathered through a sum of one-hot vectors: every rank then holds the same doubles and forms the same prefix
    std::vector<double> tots(s->world, 0.0);
    tots[s->rank] = s->h_res[0];
    s->allreduce(s->allreduce_ctx, tots.data(), s->world);
    sh.world = s->world; sh.rank = s->rank;
    double run = 0.0;
    for (int r = 0; r < s->world; ++r) { run += tots[r]; sh.prefix[r] = run; }
  }
  uint64_t threads = shots * 32;
  k_sample<<<(unsigned)((threads + 127) / 128), 128, 0, s->stream>>>(base, n, dm_stride, d_prefix, d_sub, nb, d_u, shots_per_traj, shots, sh, d_out);
  BT_CHECK_LAUNCH(s);
  BT_CUDA(cudaMemcpyAsync(out, d_out, shots * sizeof(int64_t), cudaMemcpyDeviceToHost, s->stream));
  BT_CUDA(cudaStreamSynchronize(s->stream));
  return BT_OK;
}

// physical index -> logical index (identity unless a sharded state has been remapped)
static inline int64_t phys_to_logical(const bt_sv* s, uint64_t phys) {
  uint64_t out = 0;
  for (int lb = 0; lb < s->n_qubits; ++lb)
    if ((phys >> s->phys_of_bit[lb]) & 1) out |= 1ull << lb;
  return (int64_t)out;
}

extern "C" int bt_sv_sample(const bt_sv* cs, const double* u, uint64_t shots, int64_t* out) {
  BT_TRY(bt_check_sv(cs));
  bt_sv* s = const_cast<bt_sv*>(cs);
  if (s->n_batch != 1) BT_FAIL(BT_ERR_ARG, "bt_sv_sample needs n_batch == 1 (use bt_sv_sample_batched)");
  if (s->world > 1) {
    // the inverse CDF runs over LOGICAL basis indices: bring the shards back to the identity layout first
    int ident[64];
    for (int b = 0; b < 64; ++b) ident[b] = b;
    BT_TRY(bt_sv_remap(s, ident));
  }
  BT_TRY(sample_impl(s, s->amp, 1ull << s->n_local, 0, u, shots, 1, out));
  if (s->world > 1) {
    // exactly one rank owns a shot (ownership is computed from identical numbers everywhere): the owner contributes index + 1,
    // the others 0; a sum of 0 means nobody claimed the shot -- an error, not index 0
    std::vector<double> buf(shots);
    for (uint64_t i = 0; i < shots; ++i)
      buf[i] = out[i] >= 0 ? (double)(phys_to_logical(s, ((uint64_t)s->rank << s->n_local) | (uint64_t)out[i]) + 1) : 0.0;
    s->allreduce(s->allreduce_ctx, buf.data(), (int)shots);
    for (uint64_t i = 0; i < shots; ++i) {
      if (buf[i] < 1.0) BT_FAIL(BT_ERR_SAMPLE, "sharded sampling: no rank owned shot %llu", (unsigned long long)i);
      out[i] = (int64_t)buf[i] - 1;
    }
  }
  return BT_OK;
}

extern "C" int bt_sv_sample_batched(const bt_sv* cs, const double* u, uint64_t shots_per_traj, int64_t* out) {
  BT_TRY(bt_check_sv(cs));
  bt_sv* s = const_cast<bt_sv*>(cs);
  if (s->world > 1) BT_FAIL(BT_ERR_UNSUPPORTED, "batched sampling of a sharded state");
  // one segmented launch set for the whole batch (block sums, per-trajectory scan, one warp per (trajectory, shot)): no host loop
  return sample_impl(s, s->amp, 1ull << s->n_local, 0, u, shots_per_traj, (uint64_t)s->n_batch, out);
}

int bt_sample_diag(bt_sv* v, int n, const double* u, uint64_t shots, int64_t* out) {
  return sample_impl(v, v->amp, 1ull << n, (1ull << n) + 1, u, shots, 1, out);
}
