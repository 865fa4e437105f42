// bt_prog_ops.cuh -- the micro-ops of the register programs: in-place updates of the PROG_AMPS amplitudes a thread holds.
// Included by bt_tile.cu (interpreter: one case per site) and, as text, by the pass specialiser (bt_jit.cu hands it to NVRTC),
// so it must stay self-contained: only double2, PROG_AMPS and the PK_* kinds are assumed.
// In-place primitives written as PTX with tied operands: every micro-op leaves amplitude j in the registers it found it in,
// so the run-time op loop carries x[] without the ~64 register moves per op that the compiler's phi resolution otherwise adds.
// (a, b) <- (c0 a + c1 b, c2 a + c3 b)
__device__ __forceinline__ void ip_real(double& a, double& b, double c0, double c1, double c2, double c3) {
  asm("{\n\t.reg .f64 t;\n\tmul.f64 t, %0, %4;\n\tmul.f64 %0, %0, %2;\n\tfma.rn.f64 %0, %1, %3, %0;\n\tfma.rn.f64 %1, %1, %5, t;\n\t}"
      : "+d"(a), "+d"(b)
      : "d"(c0), "d"(c1), "d"(c2), "d"(c3));
}
// (x + i y) <- (dr + i di)(x + i y); ndi = -di
__device__ __forceinline__ void ip_cmul(double& x, double& y, double dr, double di, double ndi) {
  asm("{\n\t.reg .f64 t;\n\tmul.f64 t, %0, %3;\n\tmul.f64 %0, %0, %2;\n\tfma.rn.f64 %0, %1, %4, %0;\n\tfma.rn.f64 %1, %1, %2, t;\n\t}"
      : "+d"(x), "+d"(y)
      : "d"(dr), "d"(di), "d"(ndi));
}
// general complex 2x2 on the pair (a, b); c = re/im of m00, m01, m10, m11; n = -im of the same
__device__ __forceinline__ void ip_gen(double2& a, double2& b, const double (&c)[8], const double (&n)[4]) {
  asm("{\n\t.reg .f64 t1, t2, u;\n\t"
      "mul.f64 t1, %0, %8;\n\tfma.rn.f64 t1, %1, %14, t1;\n\tfma.rn.f64 t1, %2, %10, t1;\n\tfma.rn.f64 t1, %3, %15, t1;\n\t"
      "mul.f64 t2, %1, %8;\n\tfma.rn.f64 t2, %0, %9, t2;\n\tfma.rn.f64 t2, %3, %10, t2;\n\tfma.rn.f64 t2, %2, %11, t2;\n\t"
      "mul.f64 u, %0, %4;\n\tfma.rn.f64 u, %1, %12, u;\n\tfma.rn.f64 u, %2, %6, u;\n\tfma.rn.f64 u, %3, %13, u;\n\t"
      "mul.f64 %1, %1, %4;\n\tfma.rn.f64 %1, %0, %5, %1;\n\tfma.rn.f64 %1, %3, %6, %1;\n\tfma.rn.f64 %1, %2, %7, %1;\n\t"
      "mov.f64 %0, u;\n\tmov.f64 %2, t1;\n\tmov.f64 %3, t2;\n\t}"
      : "+d"(a.x), "+d"(a.y), "+d"(b.x), "+d"(b.y)
      : "d"(c[0]), "d"(c[1]), "d"(c[2]), "d"(c[3]), "d"(c[4]), "d"(c[5]), "d"(c[6]), "d"(c[7]), "d"(n[0]), "d"(n[1]), "d"(n[2]), "d"(n[3]));
}

// in-place exchange (XOR swap: a register-renaming swap would make the op loop shuffle all 64 data registers every iteration)
__device__ __forceinline__ void ip_swap(double& a, double& b) {
  asm("{\n\t.reg .b64 p, q;\n\tmov.b64 p, %0;\n\tmov.b64 q, %1;\n\txor.b64 p, p, q;\n\txor.b64 q, q, p;\n\txor.b64 p, p, q;\n\tmov.b64 %0, p;\n\tmov.b64 %1, q;\n\t}"
      : "+d"(a), "+d"(b));
}

template <int KIND, int PQ>
__device__ __forceinline__ void prog_u1(double2 (&x)[PROG_AMPS], const double (&cc)[4], const double* __restrict__ cp) {
  if (KIND == PK_GEN) {  // cc = re/im of m00, m01 (preloaded), cp[4..7] = re/im of m10, m11
    const double c[8] = {cc[0], cc[1], cc[2], cc[3], cp[4], cp[5], cp[6], cp[7]};
    const double n[4] = {-c[1], -c[3], -c[5], -c[7]};
#pragma unroll
    for (int r = 0; r < PROG_AMPS / 2; ++r) {
      const int i0 = ((r >> PQ) << (PQ + 1)) | (r & ((1 << PQ) - 1)), i1 = i0 | (1 << PQ);
      ip_gen(x[i0], x[i1], c, n);
    }
  } else if (KIND == PK_REAL) {  // cc = m00, m01, m10, m11 (real)
    const double c0 = cc[0], c1 = cc[1], c2 = cc[2], c3 = cc[3];
#pragma unroll
    for (int r = 0; r < PROG_AMPS / 2; ++r) {
      const int i0 = ((r >> PQ) << (PQ + 1)) | (r & ((1 << PQ) - 1)), i1 = i0 | (1 << PQ);
      ip_real(x[i0].x, x[i1].x, c0, c1, c2, c3);
      ip_real(x[i0].y, x[i1].y, c0, c1, c2, c3);
    }
  } else if (KIND == PK_RXL) {  // cc = m00, Im m01, Im m10, m11: a' = m00 a + i s01 b, b' = i s10 a + m11 b
    const double c0 = cc[0], s01 = cc[1], s10 = cc[2], c3 = cc[3], ns01 = -s01, ns10 = -s10;
#pragma unroll
    for (int r = 0; r < PROG_AMPS / 2; ++r) {
      const int i0 = ((r >> PQ) << (PQ + 1)) | (r & ((1 << PQ) - 1)), i1 = i0 | (1 << PQ);
      ip_real(x[i0].x, x[i1].y, c0, ns01, s10, c3);
      ip_real(x[i0].y, x[i1].x, c0, s01, ns10, c3);
    }
  } else if (KIND == PK_DIAG) {  // cc = re/im of d0, d1
    const double d0r = cc[0], d0i = cc[1], d1r = cc[2], d1i = cc[3], nd0i = -d0i, nd1i = -d1i;
#pragma unroll
    for (int r = 0; r < PROG_AMPS / 2; ++r) {
      const int i0 = ((r >> PQ) << (PQ + 1)) | (r & ((1 << PQ) - 1)), i1 = i0 | (1 << PQ);
      ip_cmul(x[i0].x, x[i0].y, d0r, d0i, nd0i);
      ip_cmul(x[i1].x, x[i1].y, d1r, d1i, nd1i);
    }
  } else {  // PK_PHASE: cc = re/im of d
    const double dr = cc[0], di = cc[1], ndi = -di;
#pragma unroll
    for (int r = 0; r < PROG_AMPS / 2; ++r) {
      const int i1 = ((r >> PQ) << (PQ + 1)) | (r & ((1 << PQ) - 1)) | (1 << PQ);
      ip_cmul(x[i1].x, x[i1].y, dr, di, ndi);
    }
  }
}

template <int PC, int PT>
__device__ __forceinline__ void prog_cx(double2 (&x)[PROG_AMPS]) {
#pragma unroll
  for (int i = 0; i < PROG_AMPS; ++i)
    if (((i >> PC) & 1) && !((i >> PT) & 1)) {
      ip_swap(x[i].x, x[i | (1 << PT)].x);
      ip_swap(x[i].y, x[i | (1 << PT)].y);
    }
}

template <int PT>
__device__ __forceinline__ void prog_x1(double2 (&x)[PROG_AMPS]) {
#pragma unroll
  for (int i = 0; i < PROG_AMPS; ++i)
    if (!((i >> PT) & 1)) {
      ip_swap(x[i].x, x[i | (1 << PT)].x);
      ip_swap(x[i].y, x[i | (1 << PT)].y);
    }
}

template <int PA, int PB>
__device__ __forceinline__ void prog_cphase(double2 (&x)[PROG_AMPS], const double (&c)[4]) {
#pragma unroll
  for (int i = 0; i < PROG_AMPS; ++i)
    if (((i >> PA) & 1) && ((i >> PB) & 1)) ip_cmul(x[i].x, x[i].y, c[0], c[1], -c[1]);
}

