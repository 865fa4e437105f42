/*
 * oracle/strided_cpu.c -- TEST INFRASTRUCTURE ONLY (see oracle/bt_oracle.py header).
 *
 * Plain-C, in-place, strided restatement of the arithmetic the reference performs with
 * `op.expand(N)*state` (src/hilbert.jl:505, operator built by src/hilbert.jl:18-159), `partial_trace`
 * (src/linalg.jl:167-230) and the projector/normalise step of born_measure_Z (src/hilbert.jl:682-696), for sizes
 * the kron-chain restatement (oracle/bt_oracle.py) cannot reach.  It is validated against that restatement at
 * N <= 12 in tests/test_oracle_strided.py and is then the oracle for N up to ~26 and the "port" CPU baseline of
 * bench.py.  Index convention: amplitude index bit b <-> qubit N-b (src/bit.jl:9-15); matrices row-major here.
 *
 * Built by `make oracle` (gcc -O3 -fopenmp) into oracle/_build/libbt_oracle_c.so.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cd;

static inline uint64_t expand_idx(uint64_t g, int ni, const int* ins) {
  for (int i = 0; i < ni; ++i) {
    int b = ins[i];
    g = ((g >> b) << (b + 1)) | (g & ((1ull << b) - 1ull));
  }
  return g;
}

static int cmp_int(const void* a, const void* b) { return *(const int*)a - *(const int*)b; }

int bto_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* launchers such as torchrun export OMP_NUM_THREADS=1: the benchmark states its thread count explicitly */
void bto_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* y = M x on the k target bits tb[] (matrix index bit t <-> tb[t]), only where all control bits cb[] are 1.
 * m: row-major (1<<k)x(1<<k), interleaved re/im.  n_bits: total index bits.  k <= 4. */
int bto_apply(double* amp_, int n_bits, int k, const int* tb, const double* m_, int nc, const int* cb) {
  if (k < 0 || k > 4 || nc < 0 || nc > 4) return -1;
  cd* amp = (cd*)amp_;
  const cd* m = (const cd*)m_;
  int D = 1 << k;
  int ins[8], ni = 0;
  for (int i = 0; i < k; ++i) ins[ni++] = tb[i];
  for (int i = 0; i < nc; ++i) ins[ni++] = cb[i];
  qsort(ins, ni, sizeof(int), cmp_int);
  uint64_t cmask = 0;
  for (int i = 0; i < nc; ++i) cmask |= 1ull << cb[i];
  uint64_t off[16];
  for (int j = 0; j < D; ++j) {
    uint64_t o = 0;
    for (int t = 0; t < k; ++t)
      if ((j >> t) & 1) o |= 1ull << tb[t];
    off[j] = o;
  }
  uint64_t ngroups = 1ull << (n_bits - ni);
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < (int64_t)ngroups; ++g) {
    uint64_t base = expand_idx((uint64_t)g, ni, ins) | cmask;
    cd x[16], y[16];
    for (int j = 0; j < D; ++j) x[j] = amp[base + off[j]];
    for (int r = 0; r < D; ++r) {
      cd acc = 0;
      for (int c = 0; c < D; ++c) acc += m[r * D + c] * x[c];
      y[r] = acc;
    }
    for (int j = 0; j < D; ++j) amp[base + off[j]] = y[j];
  }
  return 0;
}

/* rho[a][b] = sum x_a conj(x_b) over the k bits tb[] (row-major out, (1<<k)^2 complex). */
int bto_rdm(const double* amp_, int n_bits, int k, const int* tb, double* out_) {
  if (k < 1 || k > 3) return -1;
  const cd* amp = (const cd*)amp_;
  cd* out = (cd*)out_;
  int D = 1 << k;
  int ins[4];
  for (int i = 0; i < k; ++i) ins[i] = tb[i];
  qsort(ins, k, sizeof(int), cmp_int);
  uint64_t off[8];
  for (int j = 0; j < D; ++j) {
    uint64_t o = 0;
    for (int t = 0; t < k; ++t)
      if ((j >> t) & 1) o |= 1ull << tb[t];
    off[j] = o;
  }
  uint64_t ngroups = 1ull << (n_bits - k);
  double re[64], im[64];
  memset(re, 0, sizeof(re));
  memset(im, 0, sizeof(im));
#pragma omp parallel
  {
    double lre[64], lim[64];
    memset(lre, 0, sizeof(lre));
    memset(lim, 0, sizeof(lim));
#pragma omp for schedule(static)
    for (int64_t g = 0; g < (int64_t)ngroups; ++g) {
      uint64_t base = expand_idx((uint64_t)g, k, ins);
      cd x[8];
      for (int j = 0; j < D; ++j) x[j] = amp[base + off[j]];
      for (int a = 0; a < D; ++a)
        for (int b = 0; b < D; ++b) {
          cd v = x[a] * conj(x[b]);
          lre[a * D + b] += creal(v);
          lim[a * D + b] += cimag(v);
        }
    }
#pragma omp critical
    for (int i = 0; i < D * D; ++i) { re[i] += lre[i]; im[i] += lim[i]; }
  }
  for (int i = 0; i < D * D; ++i) out[i] = re[i] + im[i] * I;
  return 0;
}

double bto_norm2(const double* amp_, uint64_t len) {
  const cd* amp = (const cd*)amp_;
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < (int64_t)len; ++i) s += creal(amp[i]) * creal(amp[i]) + cimag(amp[i]) * cimag(amp[i]);
  return s;
}

void bto_scale(double* amp_, uint64_t len, double f) {
  cd* amp = (cd*)amp_;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)len; ++i) amp[i] *= f;
}

/* |a|^2 into p; returns the total */
double bto_probs(const double* amp_, uint64_t len, double* p) {
  const cd* amp = (const cd*)amp_;
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < (int64_t)len; ++i) {
    double v = creal(amp[i]) * creal(amp[i]) + cimag(amp[i]) * cimag(amp[i]);
    p[i] = v;
    s += v;
  }
  return s;
}

/* out += a (complex vectors) */
void bto_axpy(double* out_, const double* a_, uint64_t len) {
  cd* out = (cd*)out_;
  const cd* a = (const cd*)a_;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)len; ++i) out[i] += a[i];
}
