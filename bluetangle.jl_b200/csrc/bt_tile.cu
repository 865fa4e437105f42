// bt_tile.cu -- placeholder: fused multi-gate pass (replaced by the shared-memory tile kernel).
#include "bt_internal.cuh"
int bt_fuse_and_run(bt_sv* s, const std::vector<GateDesc>& gates) {
  for (size_t i = 0; i < gates.size(); ++i) BT_TRY(bt_launch_gate(s, gates[i]));
  return BT_OK;
}
