// bt_circuit.cu -- whole-circuit entry points: replaces the per-op loop of apply(ops, state)
// (src/hilbert.jl:517-553) and to_rho's loop (src/ops.jl:813-841) for plain gates.
#include "bt_internal.cuh"

int bt_build_gate(const bt_sv* s, int nq, int qubit, int target, int control, const bt_c64* m, GateDesc* out);
int bt_prepare_local(bt_sv* s, const GateDesc& g);
int bt_fuse_and_run(bt_sv* s, const std::vector<GateDesc>& gates);  // bt_tile.cu

extern "C" int bt_sv_apply_circuit(bt_sv* s, const bt_gate* g, uint64_t n, int fuse) {
  BT_TRY(bt_check_sv(s));
  if (n && !g) BT_FAIL(BT_ERR_ARG, "null gate list");
  if (fuse && s->world == 1) {
    std::vector<GateDesc> descs(n);
    for (uint64_t i = 0; i < n; ++i) BT_TRY(bt_build_gate(s, g[i].nq, g[i].qubit, g[i].target, g[i].control, g[i].m, &descs[i]));
    return bt_fuse_and_run(s, descs);
  }
  for (uint64_t i = 0; i < n; ++i) {
    GateDesc d;
    BT_TRY(bt_build_gate(s, g[i].nq, g[i].qubit, g[i].target, g[i].control, g[i].m, &d));
    if (s->world > 1) {
      BT_TRY(bt_prepare_local(s, d));
      BT_TRY(bt_build_gate(s, g[i].nq, g[i].qubit, g[i].target, g[i].control, g[i].m, &d));
    }
    BT_TRY(bt_launch_gate(s, d));
  }
  return BT_OK;
}

extern "C" int bt_dm_apply_circuit(bt_dm* d, const bt_gate* g, uint64_t n, int fuse) {
  (void)fuse;
  if (!d) BT_FAIL(BT_ERR_ARG, "null handle");
  if (n && !g) BT_FAIL(BT_ERR_ARG, "null gate list");
  for (uint64_t i = 0; i < n; ++i) {
    if (g[i].nq == 1) BT_TRY(bt_dm_apply_1q(d, g[i].qubit, g[i].m, g[i].control));
    else if (g[i].nq == 2) BT_TRY(bt_dm_apply_2q(d, g[i].qubit, g[i].target, g[i].m, g[i].control));
    else BT_FAIL(BT_ERR_ARG, "density-matrix circuits take 1- and 2-qubit gates");
  }
  return BT_OK;
}
