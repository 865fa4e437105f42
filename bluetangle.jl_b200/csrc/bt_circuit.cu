// bt_circuit.cu -- whole-circuit entry points: replaces the per-op loop of apply(ops, state)
// (src/hilbert.jl:517-553) and to_rho's loop (src/ops.jl:813-841) for plain gates, and holds the multi-GPU circuit
// planner (no reference analogue; the label-permutation idea is relabel_swap, src/hilbert.jl:266-317).
//
// Planner: the top g = log2(P) physical index bits are the rank id.  A gate needs no communication unless one of its
// NON-DIAGONAL targets sits on a rank bit (controls and diagonal factors there are resolved from the rank id).  The
// circuit is cut into segments; inside a segment every executable gate (dependency order preserved, gates on disjoint
// qubits commute) runs locally through the fused tile kernel; between segments one remap exchanges the g global
// logical qubits for the g qubits whose pending gates unlock the most work (greedy over contiguous label windows).
#include "bt_internal.cuh"

int bt_build_gate(const bt_sv* s, int nq, int qubit, int target, int control, const bt_c64* m, GateDesc* out);
int bt_fuse_and_run(bt_sv* s, const std::vector<GateDesc>& gates);  // bt_tile.cu
int bt_localize(const bt_sv* s, GateDesc* g);                       // bt_gates.cu

#define PLAN_KEEP 5  // must match REMAP_KEEP in bt_dist.cu: the 5 lowest physical bits never move

struct Segment {
  bool remap;
  int new_phys[64];
  std::vector<int> gates;  // indices into the logical gate list, execution order
};

// gates in LOGICAL bit space (tb/cb = logical bit = n_qubits - qubit label)
static int build_logical(int n_qubits, const bt_gate* g, uint64_t n, std::vector<GateDesc>& out) {
  bt_sv tmp;
  memset(&tmp, 0, sizeof(tmp));
  tmp.n_qubits = n_qubits;
  tmp.n_local = n_qubits;
  for (int b = 0; b < 64; ++b) tmp.phys_of_bit[b] = b;
  out.resize(n);
  for (uint64_t i = 0; i < n; ++i) BT_TRY(bt_build_gate(&tmp, g[i].nq, g[i].qubit, g[i].target, g[i].control, g[i].m, &out[i]));
  return BT_OK;
}

static inline void gate_bits(const GateDesc& d, int* bits, int* nb) {
  int n = 0;
  for (int i = 0; i < d.k; ++i) bits[n++] = d.tb[i];
  for (int i = 0; i < d.nc; ++i) bits[n++] = d.cb[i];
  *nb = n;
}

// gates executable without communication when the logical bits flagged in is_global are on the rank index
static size_t executable(const std::vector<GateDesc>& L, const std::vector<int>& remaining, const bool* is_global, int n_qubits, std::vector<int>* out_exec,
                         std::vector<int>* out_rest) {
  bool blocked[64] = {false};
  int nblocked = 0;
  size_t cnt = 0;
  for (size_t r = 0; r < remaining.size(); ++r) {
    int gi = remaining[r];
    if (nblocked >= n_qubits) {
      if (out_rest) out_rest->insert(out_rest->end(), remaining.begin() + r, remaining.end());
      break;
    }
    const GateDesc& d = L[gi];
    int bits[8], nb;
    gate_bits(d, bits, &nb);
    bool ok = true;
    for (int i = 0; i < nb; ++i) if (blocked[bits[i]]) ok = false;
    if (ok && !d.diag)
      for (int i = 0; i < d.k; ++i) if (is_global[d.tb[i]]) ok = false;
    if (ok) {
      cnt++;
      if (out_exec) out_exec->push_back(gi);
    } else {
      for (int i = 0; i < nb; ++i) if (!blocked[bits[i]]) { blocked[bits[i]] = true; nblocked++; }
      if (out_rest) out_rest->push_back(gi);
    }
  }
  return cnt;
}

static int plan_circuit(int n_qubits, int n_local, const int* cur_phys_in, const std::vector<GateDesc>& L, std::vector<Segment>& plan) {
  const int g = n_qubits - n_local;
  int phys[64];
  for (int b = 0; b < n_qubits; ++b) phys[b] = cur_phys_in[b];
  std::vector<int> remaining(L.size());
  for (size_t i = 0; i < L.size(); ++i) remaining[i] = (int)i;
  if (g == 0) {
    Segment s; s.remap = false;
    for (int b = 0; b < 64; ++b) s.new_phys[b] = b < n_qubits ? phys[b] : b;
    s.gates = remaining;
    plan.push_back(s);
    return BT_OK;
  }
  int guard = 0;
  while (!remaining.empty()) {
    if (++guard > 100000) BT_FAIL(BT_ERR_ARG, "internal: circuit planner did not converge");
    bool cur_global[64] = {false};
    for (int b = 0; b < n_qubits; ++b) if (phys[b] >= n_local) cur_global[b] = true;
    size_t best_cnt = executable(L, remaining, cur_global, n_qubits, nullptr, nullptr);
    bool best_is_cur = true;
    bool best_set[64];
    memcpy(best_set, cur_global, sizeof(best_set));
    // a remap costs about as much as several full passes: only move when it at least doubles the segment, or nothing runs
    size_t cur_cnt = best_cnt;
    for (int w = 0; w + g <= n_qubits; ++w) {
      bool cand[64] = {false};
      bool valid = true;
      for (int b = w; b < w + g; ++b) {
        cand[b] = true;
        if (!cur_global[b] && phys[b] < PLAN_KEEP) valid = false;  // pinned low physical bits cannot become global
      }
      if (!valid) continue;
      size_t c = executable(L, remaining, cand, n_qubits, nullptr, nullptr);
      if (c > best_cnt) { best_cnt = c; memcpy(best_set, cand, sizeof(best_set)); best_is_cur = false; }
    }
    if (!best_is_cur && cur_cnt > 0 && best_cnt < 2 * cur_cnt && cur_cnt * 4 >= remaining.size()) {
      // the current layout already runs a large part of what is left: take it first
      memcpy(best_set, cur_global, sizeof(best_set));
      best_is_cur = true;
      best_cnt = cur_cnt;
    }
    if (best_cnt == 0) {
      // fall back: make the non-diagonal targets of the first pending gate local, evicting bits it does not touch
      const GateDesc& d = L[remaining[0]];
      bool need[64] = {false};
      for (int i = 0; i < d.k; ++i) need[d.tb[i]] = true;
      bool cand[64] = {false};
      int chosen = 0;
      for (int b = n_qubits - 1; b >= 0 && chosen < g; --b)
        if (!need[b] && (cur_global[b] || phys[b] >= PLAN_KEEP)) { cand[b] = true; chosen++; }
      if (chosen < g) BT_FAIL(BT_ERR_UNSUPPORTED, "cannot place the circuit's qubits locally");
      memcpy(best_set, cand, sizeof(best_set));
      best_is_cur = false;
      best_cnt = executable(L, remaining, best_set, n_qubits, nullptr, nullptr);
      if (best_cnt == 0) BT_FAIL(BT_ERR_ARG, "internal: planner found no executable gate");
    }
    Segment seg;
    seg.remap = !best_is_cur;
    if (seg.remap) {
      // pair leaving-global with entering-global bits and swap their physical positions
      std::vector<int> leaving, entering;
      for (int b = 0; b < n_qubits; ++b) {
        if (cur_global[b] && !best_set[b]) leaving.push_back(b);
        if (!cur_global[b] && best_set[b]) entering.push_back(b);
      }
      if (leaving.size() != entering.size()) BT_FAIL(BT_ERR_ARG, "internal: planner set sizes differ");
      for (size_t i = 0; i < leaving.size(); ++i) std::swap(phys[leaving[i]], phys[entering[i]]);
    }
    for (int b = 0; b < 64; ++b) seg.new_phys[b] = b < n_qubits ? phys[b] : b;
    std::vector<int> rest;
    executable(L, remaining, best_set, n_qubits, &seg.gates, &rest);
    remaining.swap(rest);
    plan.push_back(seg);
  }
  return BT_OK;
}

// run the gates of one segment on one shard under its current layout
static int run_segment_gates(bt_sv* s, const std::vector<GateDesc>& L, const Segment& seg, int fuse) {
  std::vector<GateDesc> phys;
  phys.reserve(seg.gates.size());
  for (int gi : seg.gates) {
    GateDesc d = L[gi];
    for (int i = 0; i < d.k; ++i) d.tb[i] = s->phys_of_bit[d.tb[i]];
    for (int i = 0; i < d.nc; ++i) d.cb[i] = s->phys_of_bit[d.cb[i]];
    if (s->world > 1) {
      int r = bt_localize(s, &d);
      if (r < 0) return r;
      if (r == 1) continue;  // a control on a rank bit is 0 on this rank
    }
    phys.push_back(d);
  }
  if (fuse) return bt_fuse_and_run(s, phys);
  for (const GateDesc& d : phys) BT_TRY(bt_launch_gate(s, d));
  return BT_OK;
}

extern "C" int bt_sv_apply_circuit(bt_sv* s, const bt_gate* g, uint64_t n, int fuse) {
  BT_TRY(bt_check_sv(s));
  if (n && !g) BT_FAIL(BT_ERR_ARG, "null gate list");
  if (s->mask_on) fuse = 0;  // the fused tile kernel has no per-trajectory predicate: masked circuits run gate by gate
  std::vector<GateDesc> L;
  BT_TRY(build_logical(s->n_qubits, g, n, L));
  std::vector<Segment> plan;
  BT_TRY(plan_circuit(s->n_qubits, s->n_local, s->phys_of_bit, L, plan));
  for (const Segment& seg : plan) {
    if (seg.remap) BT_TRY(bt_sv_remap(s, seg.new_phys));
    BT_TRY(run_segment_gates(s, L, seg, fuse));
  }
  return BT_OK;
}

// All shards of one state inside this process (bt_sv_attach_local_peers): segments run in lockstep --
// every shard remaps, then every shard runs the segment's gates.
extern "C" int bt_group_apply_circuit(bt_sv** shards, int world, const bt_gate* g, uint64_t n, int fuse) {
  if (!shards || world < 1) BT_FAIL(BT_ERR_ARG, "invalid shard list");
  for (int r = 0; r < world; ++r)
    if (!shards[r] || shards[r]->world != world || shards[r]->rank != r) BT_FAIL(BT_ERR_ARG, "shard %d does not match (rank/world)", r);
  if (n && !g) BT_FAIL(BT_ERR_ARG, "null gate list");
  bt_sv* s0 = shards[0];
  std::vector<GateDesc> L;
  BT_TRY(build_logical(s0->n_qubits, g, n, L));
  std::vector<Segment> plan;
  BT_TRY(plan_circuit(s0->n_qubits, s0->n_local, s0->phys_of_bit, L, plan));
  for (const Segment& seg : plan) {
    if (seg.remap) {
      for (int r = 0; r < world; ++r) { BT_TRY(bt_check_sv(shards[r])); BT_CUDA(cudaStreamSynchronize(shards[r]->stream)); }
      for (int r = 0; r < world; ++r) BT_TRY(bt_sv_remap(shards[r], seg.new_phys));
    }
    for (int r = 0; r < world; ++r) { BT_TRY(bt_check_sv(shards[r])); BT_TRY(run_segment_gates(shards[r], L, seg, fuse)); }
  }
  return BT_OK;
}

// Pure host: the plan the library would follow for `world` shards starting from the identity layout.
// seg_gates[i] = number of gates in segment i, seg_remap[i] = 1 if a remap precedes it, layouts = cap x n_qubits.
extern "C" int bt_plan_circuit_host(int n_qubits, int world, const bt_gate* g, uint64_t n, int* n_segments, int* seg_gates, int* seg_remap, int* layouts,
                                    int* order, int cap) {
  if (!n_segments) BT_FAIL(BT_ERR_ARG, "null output");
  if (world < 1 || (world & (world - 1))) BT_FAIL(BT_ERR_ARG, "world must be a power of two");
  int gg = 0;
  while ((1 << gg) < world) ++gg;
  if (n_qubits - gg < PLAN_KEEP + gg && world > 1) BT_FAIL(BT_ERR_ARG, "too few qubits for %d shards", world);
  std::vector<GateDesc> L;
  BT_TRY(build_logical(n_qubits, g, n, L));
  int phys[64];
  for (int b = 0; b < 64; ++b) phys[b] = b;
  std::vector<Segment> plan;
  BT_TRY(plan_circuit(n_qubits, n_qubits - gg, phys, L, plan));
  *n_segments = (int)plan.size();
  size_t pos = 0;
  for (size_t i = 0; i < plan.size() && (int)i < cap; ++i) {
    if (seg_gates) seg_gates[i] = (int)plan[i].gates.size();
    if (seg_remap) seg_remap[i] = plan[i].remap ? 1 : 0;
    if (layouts) for (int b = 0; b < n_qubits; ++b) layouts[i * n_qubits + b] = plan[i].new_phys[b];
    if (order) for (int gi : plan[i].gates) order[pos++] = gi;
  }
  return BT_OK;
}

int bt_fusion_plan(const std::vector<GateDesc>& gates, int n_local, int* n_passes, int* n_blocks, int* gates_in_pass, int* tile_bits, double* cost_in_pass,
                   int* end_reason, int cap, int* dry_counts);  // bt_tile.cu

// Pure host: the fused passes bt_sv_apply_circuit(fuse = 1) forms for this gate list on one unsharded GPU.
extern "C" int bt_fusion_plan_host(int n_qubits, const bt_gate* g, uint64_t n, int* n_passes, int* n_blocks, int* gates_in_pass, int* tile_bits, double* cost_in_pass,
                                   int* end_reason, int cap, int* launch_counts) {
  if (n && !g) BT_FAIL(BT_ERR_ARG, "null gate list");
  if (n_qubits < 1 || n_qubits > 40) BT_FAIL(BT_ERR_ARG, "n_qubits out of range");
  std::vector<GateDesc> L;
  BT_TRY(build_logical(n_qubits, g, n, L));
  if (launch_counts) memset(launch_counts, 0, 4 * sizeof(int));
  return bt_fusion_plan(L, n_qubits, n_passes, n_blocks, gates_in_pass, tile_bits, cost_in_pass, end_reason, cap, launch_counts);
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
