// planner_harness.cc — the host planner (viyadb_b200/csrc/planner.h: predicate lowering, segment pruning, key-domain
// tightening) and the device's leaf evaluators (csrc/device_arith.h: leaf_mask16, gen_compare, post_compare), compiled
// with plain g++ for the CPU tests (tests/test_planner_fuzz.py). TEST INFRASTRUCTURE: the very source libvgpu.so
// compiles; only the two stack machines (scan_kernel.cuh eval_predicate, kernels.cuh group_passes — both entangled with
// device loads) are restated here, instruction for instruction. Not a CPU path of the product: it answers no query.
#include "../viyadb_b200/csrc/planner.h"
#include "../viyadb_b200/csrc/device_arith.h"

#include <memory>

namespace {

struct FakeTable {
  std::vector<ColInfo> cols;
};
struct FakeSeg {
  uint64_t nrows = 0;
  std::vector<uint64_t> omin, omax;
};

struct Handle {
  FakeTable table;
  std::vector<vgpu_pred_node> nodes;
  std::vector<uint64_t> args;
  std::vector<vgpu_key> keys;
  std::vector<uint32_t> metric_cols;
  vgpu_plan plan{};
  std::unique_ptr<PlannerT<FakeTable>> pl;
  std::string error;
};

// eval_predicate (scan_kernel.cuh) for the 16 rows whose raw cells are given per slot: raw[slot][i] = the zero-extended
// stored value for the vector classes (what load_vec16 produces), wide[slot][i] = the widened value / cardinality for C_GEN
uint32_t eval16(const ScanParams &P, const uint32_t (*raw)[kRowsPerThread], const uint64_t (*wide)[kRowsPerThread], bool force_interp) {
  if (P.conj && !force_interp) {
    uint32_t m = 0xffffu;
    for (uint32_t pc = 0; pc < 4; ++pc) {
      if (pc < P.nprog) {
        const PInstr &in = P.prog[pc];
        uint32_t lm = leaf_mask16(in, raw[in.slot]);
        if (in.neg) lm ^= 0xffffu;
        m &= lm;
      }
    }
    return m;
  }
  uint32_t stk[kStackDepth];
  for (int i = 0; i < kStackDepth; ++i) stk[i] = 0;
  for (uint32_t pc = 0; pc < P.nprog; ++pc) {
    const PInstr &in = P.prog[pc];
    const uint32_t kind = in.kind;
    if (kind <= P_OR_LEAF) {
      uint32_t m = 0;
      const uint32_t cls = in.cls;
      if (cls == C_TRUE) {
        m = 0xffffu;
      } else if (cls == C_FALSE) {
        m = 0;
      } else if (cls == C_GEN) {
        for (int i = 0; i < kRowsPerThread; ++i)
          if (gen_compare(in.gcls, in.gop, wide[in.slot][i], in.arg)) m |= 1u << i;
      } else {
        m = leaf_mask16(in, raw[in.slot]);
      }
      if (in.neg) m ^= 0xffffu;
      if (kind == P_PUSH) {
        for (int i = kStackDepth - 1; i > 0; --i) stk[i] = stk[i - 1];
        stk[0] = m;
      } else if (kind == P_AND_LEAF) {
        stk[0] &= m;
      } else {
        stk[0] |= m;
      }
    } else {
      uint32_t r = (kind == P_AND) ? (stk[1] & stk[0]) : (stk[1] | stk[0]);
      stk[0] = r;
      for (int i = 1; i < kStackDepth - 1; ++i) stk[i] = stk[i + 1];
    }
  }
  return stk[0];
}

}  // namespace

extern "C" {

// cols: per schema column {kind, type, agg, lit_type (0: none)}; bitset columns have kind VGPU_METRIC_BITSET.
// having != 0: lower plan->hnodes-style program over (keys, metrics) sources instead of the row filter.
void *h_planner_new(uint32_t ncols, const uint32_t *kinds, const uint32_t *types, const uint32_t *aggs, const uint32_t *lit_types,
                    uint32_t nnodes, const vgpu_pred_node *nodes, uint32_t nargs, const uint64_t *args,
                    uint32_t nkeys, const uint32_t *key_cols, uint32_t nmetrics, const uint32_t *metric_cols, int having,
                    uint32_t tune) {
  auto *h = new Handle();
  uint64_t off = 0;
  uint32_t nb = 0;
  for (uint32_t c = 0; c < ncols; ++c) {
    ColInfo ci{};
    // the rules of vgpu_table_create (vgpu.cu)
    ci.kind = kinds[c];
    ci.type = types[c];
    ci.agg = aggs[c];
    ci.lit_type = lit_types[c] ? lit_types[c] - 1 : types[c];
    ci.width = type_width(types[c]);
    ci.sext = type_signed(types[c]);
    if (kinds[c] == VGPU_METRIC_BITSET) {
      ci.bitset = true;
      ci.bitset_idx = nb++;
      ci.agg = VGPU_AGG_BITSET;
    } else {
      ci.off_per_row = off;
      off += ci.width;
    }
    h->table.cols.push_back(ci);
  }
  h->nodes.assign(nodes, nodes + nnodes);
  h->args.assign(args, args + nargs);
  for (uint32_t k = 0; k < nkeys; ++k) {
    vgpu_key key{};
    key.col = key_cols[k];
    key.query_granularity = VGPU_TU_NONE;
    h->keys.push_back(key);
  }
  h->metric_cols.assign(metric_cols, metric_cols + nmetrics);
  h->plan.nkeys = nkeys;
  h->plan.keys = h->keys.data();
  h->plan.nmetrics = nmetrics;
  h->plan.metric_cols = h->metric_cols.data();
  if (having) {
    h->plan.nhnodes = nnodes; h->plan.hnodes = h->nodes.data();
    h->plan.nhargs = nargs; h->plan.hargs = h->args.data();
  } else {
    h->plan.nnodes = nnodes; h->plan.nodes = h->nodes.data();
    h->plan.nargs = nargs; h->plan.args = h->args.data();
  }
  try {
    if (having) h->pl.reset(new PlannerT<FakeTable>(&h->table, &h->plan, true));
    else h->pl.reset(new PlannerT<FakeTable>(&h->table, &h->plan));
    h->pl->build_predicate();
    if (!having) h->pl->finish_predicate(tune);
  } catch (const Err &e) {
    h->error = e.msg;
  }
  return h;
}
void h_planner_free(void *p) { delete static_cast<Handle *>(p); }
const char *h_planner_error(void *p) {
  auto *h = static_cast<Handle *>(p);
  return h->error.empty() ? nullptr : h->error.c_str();
}
uint32_t h_planner_nprog(void *p) { return static_cast<Handle *>(p)->pl->P.nprog; }
uint32_t h_planner_conj(void *p) { return static_cast<Handle *>(p)->pl->P.conj; }
uint32_t h_planner_max_depth(void *p) { return (uint32_t)static_cast<Handle *>(p)->pl->max_depth; }
// instruction i as {kind, cls, slot, neg, gcls, gop}
void h_planner_instr(void *p, uint32_t i, uint32_t *out) {
  const PInstr &in = static_cast<Handle *>(p)->pl->P.prog[i];
  out[0] = in.kind; out[1] = in.cls; out[2] = in.slot; out[3] = in.neg; out[4] = in.gcls; out[5] = in.gop;
}
// schema column behind predicate slot s
uint32_t h_planner_slot_col(void *p, uint32_t s) { return static_cast<Handle *>(p)->pl->slot_cols[s]; }
uint32_t h_planner_nslots(void *p) { return static_cast<Handle *>(p)->pl->P.nslots; }

// Row predicate over n rows. cols[c] = the column's cells widened to 64 bits the way the kernel widens them
// (zero-/sign-extended integers, raw IEEE bits, a bitset column's cardinality per row). out[r] = 0/1.
void h_planner_eval_rows(void *p, uint64_t n, const uint64_t *const *cols, int force_interp, uint8_t *out) {
  auto *h = static_cast<Handle *>(p);
  const ScanParams &P = h->pl->P;
  uint32_t raw[kMaxSlots][kRowsPerThread];
  uint64_t wide[kMaxSlots][kRowsPerThread];
  for (uint64_t r0 = 0; r0 < n; r0 += kRowsPerThread) {
    for (uint32_t s = 0; s < P.nslots; ++s) {
      const ColInfo &ci = h->table.cols[h->pl->slot_cols[s]];
      const uint64_t *col = cols[h->pl->slot_cols[s]];
      for (int i = 0; i < kRowsPerThread; ++i) {
        const uint64_t v = r0 + i < n ? col[r0 + i] : 0;   // the slab's tail is zero
        wide[s][i] = v;
        // load_vec16: the stored cell, zero-extended to 32 bits (1/2/4-byte columns)
        raw[s][i] = ci.width == 4 ? (uint32_t)v : (uint32_t)(v & (ci.width == 2 ? 0xffffu : 0xffu));
      }
    }
    const uint32_t m = eval16(P, raw, wide, force_interp != 0);
    for (int i = 0; i < kRowsPerThread && r0 + i < n; ++i) out[r0 + i] = (m >> i) & 1u;
  }
}

// Planner::process_segment for one segment with the given ordered (to_ordered) min / max per schema column
int h_planner_process_segment(void *p, uint64_t nrows, const uint64_t *omin, const uint64_t *omax) {
  auto *h = static_cast<Handle *>(p);
  FakeSeg sd;
  sd.nrows = nrows;
  sd.omin.assign(omin, omin + h->table.cols.size());
  sd.omax.assign(omax, omax + h->table.cols.size());
  return h->pl->process_segment(sd, h->pl->root) ? 1 : 0;
}

// HAVING program over ngroups groups: sources[src] = widened values of selected key k (src = k) or metric m
// (src = nkeys + m: the raw accumulator, a count-distinct metric's cardinality); group_passes (kernels.cuh)
void h_planner_eval_groups(void *p, uint64_t ngroups, const uint64_t *const *sources, uint8_t *out) {
  auto *h = static_cast<Handle *>(p);
  const ScanParams &P = h->pl->P;
  for (uint64_t g = 0; g < ngroups; ++g) {
    uint32_t stk = 0;
    int sp = 0;
    for (uint32_t pc = 0; pc < P.nprog; ++pc) {
      const PInstr &in = P.prog[pc];
      if (in.kind <= P_OR_LEAF) {
        bool m = in.cls == C_TRUE ? true : in.cls == C_FALSE ? false : post_compare(in.gcls, in.gop, sources[in.slot][g], in.arg);
        if (in.neg) m = !m;
        if (in.kind == P_PUSH) { stk = (stk & ~(1u << sp)) | ((m ? 1u : 0u) << sp); ++sp; }
        else if (in.kind == P_AND_LEAF) { if (!m) stk &= ~(1u << (sp - 1)); }
        else { if (m) stk |= 1u << (sp - 1); }
      } else {
        const bool b = (stk >> (sp - 1)) & 1u, a = (stk >> (sp - 2)) & 1u;
        const bool r = in.kind == P_AND ? (a && b) : (a || b);
        --sp;
        stk = (stk & ~(1u << (sp - 1))) | ((r ? 1u : 0u) << (sp - 1));
      }
    }
    out[g] = sp == 0 ? 1 : (stk & 1u);
  }
}

// tighten_key_domain for the key on schema column `col` (must be one of the predicate's slots); returns the lookup mask
uint64_t h_planner_tighten(void *p, uint32_t col, uint64_t *lo, uint64_t *hi, int *applies) {
  auto *h = static_cast<Handle *>(p);
  const ScanParams &P = h->pl->P;
  *applies = 0;
  const ColInfo &ci = h->table.cols[col];
  if (!P.conj || type_signed(ci.type) || ci.width > 4 || ci.bitset) return 0;   // the caller's condition (query_agg.inl)
  for (uint32_t s = 0; s < P.nslots; ++s) {
    if (h->pl->slot_cols[s] != col) continue;
    *applies = 1;
    return tighten_key_domain(P, s, *lo, *hi);
  }
  return 0;
}

uint64_t h_to_ordered_host(uint64_t v, uint32_t type) { return to_ordered_host(v, type); }

}  // extern "C"
