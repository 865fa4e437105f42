/* abi_smoke.c -- a plain C caller of include/bluetangle_cuda.h, compiled with -Wall -Werror: signature drift between the
 * header and the library shows up as a compile or link error, and on a GPU box the program runs the minimal path
 * create -> apply_1q -> apply_2q -> rdm1 -> measure_z -> sample -> destroy and checks the numbers.
 * Build (tests/test_abi.py does this): gcc -std=c11 -Wall -Werror -Iinclude tests/abi_smoke.c -o abi_smoke -Lbluetangle.jl_b200/lib -l:libbluetangle_cuda.so
 * Exit code: 0 ok, 2 no device (link-only check passed), 1 wrong result. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "bluetangle_cuda.h"

#define CHECK(call)                                                          \
  do {                                                                       \
    int rc_ = (call);                                                        \
    if (rc_ != BT_OK) {                                                      \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, bt_last_error());        \
      return 1;                                                              \
    }                                                                        \
  } while (0)

int main(void) {
  int ndev = 0;
  if (bt_device_count(&ndev) != BT_OK || ndev == 0) {
    printf("abi_smoke: no device (%s); header and library link\n", bt_last_error());
    return 2;
  }
  const int n = 10;
  const double r = 0.70710678118654752440;
  const bt_c64 H[4] = {{r, 0}, {r, 0}, {r, 0}, {-r, 0}};          /* column-major */
  bt_c64 CX[16];
  memset(CX, 0, sizeof(CX));
  CX[0 + 4 * 0].re = 1; CX[1 + 4 * 1].re = 1; CX[3 + 4 * 2].re = 1; CX[2 + 4 * 3].re = 1; /* index 2*b_qubit + b_target, qubit = control */
  bt_sv* s = NULL;
  CHECK(bt_sv_create(n, 1, &s));
  int nq = 0;
  CHECK(bt_sv_n_qubits(s, &nq));
  if (nq != n) return 1;
  CHECK(bt_sv_apply_1q(s, 1, H, -2));            /* H on qubit 1 (MSB) */
  CHECK(bt_sv_apply_2q(s, 1, n, CX, -2));        /* CX control 1 -> target n: Bell pair on (1, n) */
  bt_c64 rdm[4];
  CHECK(bt_sv_rdm1(s, n, rdm));
  if (fabs(rdm[0].re - 0.5) > 1e-12 || fabs(rdm[3].re - 0.5) > 1e-12 || fabs(rdm[1].re) > 1e-12) { fprintf(stderr, "rdm1 wrong\n"); return 1; }
  double nrm = 0;
  CHECK(bt_sv_norm2(s, &nrm));
  if (fabs(nrm - 1.0) > 1e-12) return 1;
  const double u = 0.75;                          /* u >= p0 = 0.5 -> outcome 1 */
  int32_t outcome = -1;
  double p0 = 0;
  CHECK(bt_sv_measure_z(s, 1, &u, &outcome, &p0, 0));
  if (outcome != 1 || fabs(p0 - 0.5) > 1e-12) { fprintf(stderr, "measure_z wrong: %d %g\n", outcome, p0); return 1; }
  const double us[3] = {0.1, 0.5, 0.9};
  int64_t idx[3];
  CHECK(bt_sv_sample(s, us, 3, idx));
  const int64_t want = (1ll << (n - 1)) | 1;      /* |1 0..0 1> : qubit 1 and qubit n set */
  for (int i = 0; i < 3; ++i)
    if (idx[i] != want) { fprintf(stderr, "sample wrong: %lld\n", (long long)idx[i]); return 1; }
  double ez = 0;
  char pauli[16];
  memset(pauli, 'I', n); pauli[n] = 0; pauli[0] = 'Z';
  CHECK(bt_sv_expect_pauli(s, pauli, &ez));
  if (fabs(ez + 1.0) > 1e-12) return 1;
  uint64_t launches = 0;
  CHECK(bt_sv_launch_count(s, &launches));
  CHECK(bt_sv_destroy(s));
  bt_dm* d = NULL;
  CHECK(bt_dm_create(4, &d));
  CHECK(bt_dm_apply_1q(d, 2, H, -2));
  bt_c64 tr;
  CHECK(bt_dm_trace(d, &tr));
  if (fabs(tr.re - 1.0) > 1e-12) return 1;
  CHECK(bt_dm_destroy(d));
  printf("abi_smoke ok: %llu kernel launches, outcome %d, p0 %.3f, sample %lld\n", (unsigned long long)launches, outcome, p0, (long long)idx[0]);
  return 0;
}
