// planner.h — host side of the predicate: vgpu_plan (include/vgpu.h) -> the device's predicate program (PInstr,
// scan_params.h) and the per-segment pruning decision. Plain C++ (no CUDA): vgpu.cu includes it for the product,
// tests/planner_harness.cc compiles the very same source with g++ so that the lowering (NOT-pushed trees, IN lists as
// lookup masks, fused ranges, signed / 64-bit / floating-point / cardinality compares) and SegmentSkip
// (src/codegen/query/filter.cc:263-349, NOT-IN quirk included) are tested on the CPU against the reference's own runs.
#ifndef VGPU_PLANNER_H_
#define VGPU_PLANNER_H_

#include "../../include/vgpu.h"
#include "scan_params.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cstring>
#include <stdint.h>
#include <string>
#include <vector>

namespace {

using namespace vgpu;

struct Err {
  int code;
  std::string msg;
};
[[noreturn]] void fail(int code, const std::string &m) { throw Err{code, m}; }


// ---------------------------------------------------------------------------------------------
// element types
// ---------------------------------------------------------------------------------------------
uint32_t type_width(uint32_t t) {
  switch (t) {
    case VGPU_U8: case VGPU_I8: return 1;
    case VGPU_U16: case VGPU_I16: return 2;
    case VGPU_U32: case VGPU_I32: case VGPU_F32: return 4;
    case VGPU_U64: case VGPU_I64: case VGPU_F64: return 8;
  }
  fail(VGPU_ERR_INVALID, "unknown element type " + std::to_string(t));
}
bool type_signed(uint32_t t) { return t >= VGPU_I8 && t <= VGPU_I64; }
bool type_float(uint32_t t) { return t == VGPU_F32 || t == VGPU_F64; }

// raw 8-byte AnyNum image -> value widened the way the kernel widens a cell (low `width` bytes,
// sign-extended for signed ints). The upper bytes of an AnyNum are uninitialised in the reference
// (src/db/column.h:98-121), so they are never looked at.
uint64_t widen_arg(uint64_t raw, uint32_t type) {
  switch (type) {
    case VGPU_U8: return raw & 0xffull;
    case VGPU_U16: return raw & 0xffffull;
    case VGPU_U32: case VGPU_F32: return raw & 0xffffffffull;
    case VGPU_I8: return (uint64_t)(int64_t)(int8_t)(raw & 0xff);
    case VGPU_I16: return (uint64_t)(int64_t)(int16_t)(raw & 0xffff);
    case VGPU_I32: return (uint64_t)(int64_t)(int32_t)(raw & 0xffffffffull);
    default: return raw;
  }
}

// Same mapping as the device-side to_ordered(): order-preserving image in uint64.
uint64_t to_ordered_host(uint64_t v, uint32_t type) {
  switch (type) {
    case VGPU_I8: case VGPU_I16: case VGPU_I32: case VGPU_I64:
      return v ^ 0x8000000000000000ull;
    case VGPU_F32: {
      uint32_t b = (uint32_t)v;
      if (b == 0x80000000u) b = 0;
      b = (b >> 31) ? ~b : (b | 0x80000000u);
      return b;
    }
    case VGPU_F64: {
      uint64_t b = v;
      if (b == 0x8000000000000000ull) b = 0;
      return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    }
    default: return v;
  }
}
uint64_t from_ordered_int(uint64_t o, uint32_t type) {
  return type_signed(type) ? (o ^ 0x8000000000000000ull) : o;
}

// NumericType::cpp_min_value / cpp_max_value (src/db/column.cc:187-219): the initial values of MAX /
// MIN accumulators (store.cc:106-117) and of SegmentStats dmax / dmin (store.cc:171-184). Note
// FLT_MIN / DBL_MIN (smallest positive) for floating point — reproduced on purpose.
uint64_t type_min_value(uint32_t t) {
  switch (t) {
    case VGPU_I8: return (uint64_t)(int64_t)INT8_MIN;
    case VGPU_I16: return (uint64_t)(int64_t)INT16_MIN;
    case VGPU_I32: return (uint64_t)(int64_t)INT32_MIN;
    case VGPU_I64: return (uint64_t)INT64_MIN;
    case VGPU_F32: { float f = FLT_MIN; uint32_t b; memcpy(&b, &f, 4); return b; }
    case VGPU_F64: { double d = DBL_MIN; uint64_t b; memcpy(&b, &d, 8); return b; }
    default: return 0;
  }
}
uint64_t type_max_value(uint32_t t) {
  switch (t) {
    case VGPU_U8: return UINT8_MAX;
    case VGPU_U16: return UINT16_MAX;
    case VGPU_U32: return UINT32_MAX;
    case VGPU_U64: return UINT64_MAX;
    case VGPU_I8: return INT8_MAX;
    case VGPU_I16: return INT16_MAX;
    case VGPU_I32: return INT32_MAX;
    case VGPU_I64: return INT64_MAX;
    case VGPU_F32: { float f = FLT_MAX; uint32_t b; memcpy(&b, &f, 4); return b; }
    default: { double d = DBL_MAX; uint64_t b; memcpy(&b, &d, 8); return b; }
  }
}

struct ColInfo {
  uint32_t kind, type, agg;
  uint32_t lit_type;  // type of a filter literal on this column (BITSET: the reference's id type, which hosts may have widened)
  uint32_t width;
  bool sext;
  bool bitset;
  uint32_t bitset_idx;
  uint64_t off_per_row;  // bytes per row of the preceding fixed-width columns
  uint32_t row_off;      // byte offset inside a row of the row-major mirror
};

// ---------------------------------------------------------------------------------------------
// planner: vgpu_plan -> ScanParams
// ---------------------------------------------------------------------------------------------
struct TreeNode {
  uint32_t kind, op, col, arg, n;
  std::vector<int> kids;
};

template <class TableT>
struct PlannerT {
  const TableT *t;
  const vgpu_plan *plan;
  // the predicate this planner lowers: the row filter, or (having = true) the HAVING filter over selected columns
  const vgpu_pred_node *nodes_;
  const uint64_t *args_;
  uint32_t nnodes_, nargs_;
  bool having = false;
  ScanParams P{};
  std::vector<TreeNode> tree;
  int root = -1;
  int depth = 0, max_depth = 0;

  PlannerT(const TableT *table, const vgpu_plan *p)
      : t(table), plan(p), nodes_(p->nodes), args_(p->args), nnodes_(p->nnodes), nargs_(p->nargs) {}
  PlannerT(const TableT *table, const vgpu_plan *p, bool having_filter)
      : t(table), plan(p), nodes_(p->hnodes), args_(p->hargs), nnodes_(p->nhnodes), nargs_(p->nhargs), having(having_filter) {}

  // HAVING: a leaf reads a selected key (source k) or a selected metric (source nkeys + m) of the group
  uint32_t source_of(uint32_t col) const {
    for (uint32_t k = 0; k < plan->nkeys; ++k)
      if (plan->keys[k].col == col) return k;
    for (uint32_t m = 0; m < plan->nmetrics; ++m)
      if (plan->metric_cols[m] == col) return plan->nkeys + m;
    fail(VGPU_ERR_INVALID, "HAVING on a column that is not selected");
  }

  uint32_t slot_of(uint32_t col) {
    if (col >= t->cols.size()) fail(VGPU_ERR_INVALID, "column index out of range");
    const ColInfo &ci = t->cols[col];
    for (uint32_t s = 0; s < P.nslots; ++s)
      if (slot_cols[s] == col) return s;
    if (P.nslots >= kMaxSlots) fail(VGPU_ERR_UNSUPPORTED, "query touches too many columns");
    Slot &sl = P.slots[P.nslots];
    sl.off = ci.off_per_row;
    sl.width = ci.width;
    sl.sext = ci.sext;
    sl.bitset = ci.bitset;
    sl.bitset_idx = ci.bitset_idx;
    sl.row_off = ci.row_off;
    if (ci.bitset) {  // ids / CSR offsets are uint32
      sl.width = 4;
      sl.sext = 0;
    }
    sl.vmask = sl.width == 8 ? ~0ull : ((1ull << (8 * sl.width)) - 1);
    sl.signbit = sl.sext ? (1ull << (8 * sl.width - 1)) : 0;
    slot_cols[P.nslots] = col;
    return P.nslots++;
  }
  uint32_t slot_cols[kMaxSlots];

  void build_tree() {
    if (nnodes_ == 0) {  // no filter at all == EmptyFilter
      TreeNode e{};
      e.kind = VGPU_NODE_EMPTY;
      tree.push_back(e);
      root = 0;
      return;
    }
    std::vector<int> stack;
    for (uint32_t i = 0; i < nnodes_; ++i) {
      const vgpu_pred_node &pn = nodes_[i];
      TreeNode tn{};
      tn.kind = pn.kind; tn.op = pn.op; tn.col = pn.col; tn.arg = pn.arg; tn.n = pn.n;
      switch (pn.kind) {
        case VGPU_NODE_RELOP:
          if (pn.op > VGPU_OP_GE) fail(VGPU_ERR_INVALID, "bad relational operator");
          if (pn.arg >= nargs_) fail(VGPU_ERR_INVALID, "predicate argument out of range");
          if (pn.col >= t->cols.size()) fail(VGPU_ERR_INVALID, "predicate column out of range");
          break;
        case VGPU_NODE_IN:
          if (pn.n == 0) fail(VGPU_ERR_INVALID, "IN filter without values");
          if ((uint64_t)pn.arg + pn.n > nargs_) fail(VGPU_ERR_INVALID, "predicate argument out of range");
          if (pn.col >= t->cols.size()) fail(VGPU_ERR_INVALID, "predicate column out of range");
          break;
        case VGPU_NODE_AND:
        case VGPU_NODE_OR:
          if (pn.n == 0) fail(VGPU_ERR_INVALID, "composite filter without children");
          if (pn.n > stack.size()) fail(VGPU_ERR_INVALID, "malformed predicate program");
          tn.kids.assign(stack.end() - pn.n, stack.end());
          stack.resize(stack.size() - pn.n);
          break;
        case VGPU_NODE_EMPTY:
          break;
        default:
          fail(VGPU_ERR_INVALID, "unknown predicate node kind");
      }
      tree.push_back(tn);
      stack.push_back((int)tree.size() - 1);
    }
    if (stack.size() != 1) fail(VGPU_ERR_INVALID, "predicate program does not reduce to one expression");
    root = stack[0];
  }

  // ---- leaf classification ----
  static uint32_t ord32(uint64_t widened, uint32_t type) {
    switch (type) {
      case VGPU_I8: return (uint32_t)(widened & 0xff) ^ 0x80u;
      case VGPU_I16: return (uint32_t)(widened & 0xffff) ^ 0x8000u;
      case VGPU_I32: return (uint32_t)widened ^ 0x80000000u;
      default: return (uint32_t)widened;
    }
  }
  static uint32_t bias32(uint32_t type) {
    switch (type) {
      case VGPU_I8: return 0x80u;
      case VGPU_I16: return 0x8000u;
      case VGPU_I32: return 0x80000000u;
      default: return 0;
    }
  }
  static uint32_t ordmax32(uint32_t type) {
    switch (type_width(type)) {
      case 1: return 0xffu;
      case 2: return 0xffffu;
      default: return 0xffffffffu;
    }
  }

  void emit(PInstr in) {
    if (P.nprog >= kMaxProg) fail(VGPU_ERR_UNSUPPORTED, "predicate too long for the device program");
    P.prog[P.nprog++] = in;
  }
  void push_depth() { max_depth = std::max(max_depth, ++depth); }

  static uint8_t leaf_kind(int mode) { return mode == 0 ? P_PUSH : (mode == 1 ? P_AND_LEAF : P_OR_LEAF); }

  // one comparison `col OP arg`; mode 0 push, 1 and-into-top, 2 or-into-top
  void emit_compare(uint32_t col, uint32_t op, uint64_t raw_arg, int mode) {
    const ColInfo &ci = t->cols[col];
    PInstr in{};
    in.kind = leaf_kind(mode);
    if (having) {
      // per group, on the device: generic compare of the widened value (post_compare, kernels.cuh)
      if (mode == 0) push_depth();
      in.slot = (uint8_t)source_of(col);
      in.cls = C_GEN;
      in.gop = (uint8_t)op;
      if (ci.bitset) { in.gcls = G_CARD; in.arg = widen_arg(raw_arg, ci.lit_type); }
      else {
        in.gcls = ci.type == VGPU_F32 ? G_F32 : ci.type == VGPU_F64 ? G_F64 : type_signed(ci.type) ? G_I64 : G_U64;
        in.arg = widen_arg(raw_arg, ci.type);
      }
      emit(in);
      return;
    }
    in.slot = (uint8_t)slot_of(col);
    if (mode == 0) push_depth();
    if (ci.bitset) {  // compares cardinality() (filter.cc:215-217)
      in.cls = C_GEN; in.gcls = G_CARD; in.gop = (uint8_t)op;
      in.arg = widen_arg(raw_arg, ci.lit_type);  // only the literal's own bytes of the AnyNum image are defined
      emit(in);
      return;
    }
    uint64_t a = widen_arg(raw_arg, ci.type);
    if (ci.width == 8 || type_float(ci.type)) {
      in.cls = C_GEN;
      in.gcls = ci.type == VGPU_F32 ? G_F32 : ci.type == VGPU_F64 ? G_F64 : ci.type == VGPU_I64 ? G_I64 : G_U64;
      in.gop = (uint8_t)op;
      in.arg = a;
      emit(in);
      return;
    }
    const uint32_t ao = ord32(a, ci.type), omax = ordmax32(ci.type);
    in.bias = bias32(ci.type);
    switch (op) {
      case VGPU_OP_EQ: case VGPU_OP_NE:
        in.cls = C_EQ32;
        in.arg = (uint32_t)(a & (ci.width == 4 ? 0xffffffffull : ((1ull << (8 * ci.width)) - 1)));
        in.neg = op == VGPU_OP_NE;
        break;
      case VGPU_OP_LT: case VGPU_OP_GE:
        in.cls = C_LT32; in.arg = ao; in.neg = op == VGPU_OP_GE;
        break;
      default:  // LE / GT:  x <= a  <=>  x < a+1 unless a is the largest value of the type
        if (ao == omax) { in.cls = C_TRUE; } else { in.cls = C_LT32; in.arg = ao + 1; }
        in.neg = op == VGPU_OP_GT;
        break;
    }
    emit(in);
  }

  bool is_small_int_col(uint32_t col) const {
    if (having) return false;  // no vector leaf classes, lookup masks or range fusion per group
    const ColInfo &ci = t->cols[col];
    return !ci.bitset && ci.width <= 4 && !type_float(ci.type);
  }

  // mode: 0 push, 1 and-into-top, 2 or-into-top
  void emit_node(int idx, int mode) {
    const TreeNode &n = tree[idx];
    switch (n.kind) {
      case VGPU_NODE_EMPTY: {
        PInstr in{};
        in.kind = leaf_kind(mode);
        in.cls = C_TRUE;
        if (mode == 0) push_depth();
        emit(in);
      } break;
      case VGPU_NODE_RELOP:
        emit_compare(n.col, n.op, args_[n.arg], mode);
        break;
      case VGPU_NODE_IN: {
        // IN = OR chain of ==, NOT IN = AND chain of != (filter.cc:222-241)
        const bool eq = n.op != 0;
        // small dictionary / numeric domains: one 64-bit lookup mask instead of a compare chain
        if (n.n >= 2 && is_small_int_col(n.col) && !type_signed(t->cols[n.col].type)) {
          const ColInfo &ci = t->cols[n.col];
          uint64_t lut = 0;
          bool ok = true;
          for (uint32_t i = 0; i < n.n && ok; ++i) {
            uint64_t a = widen_arg(args_[n.arg + i], ci.type);
            // a literal that is missing from the dictionary decodes to UINTn_MAX: it matches no stored
            // code (dictionary.cc:46-75 hands out codes from 0 upwards), so it adds nothing to the mask
            if (ci.kind == VGPU_DIM_STRING && a == type_max_value(ci.type)) continue;
            if (a >= 64) ok = false; else lut |= 1ull << a;
          }
          if (ok) {
            PInstr in{};
            in.kind = leaf_kind(mode);
            if (mode == 0) push_depth();
            in.cls = C_LUT64;
            in.slot = (uint8_t)slot_of(n.col);
            in.arg = lut;
            in.neg = !eq;
            emit(in);
            break;
          }
        }
        const int chain = eq ? 2 : 1;
        const uint32_t op = eq ? VGPU_OP_EQ : VGPU_OP_NE;
        if (mode == chain) {
          for (uint32_t i = 0; i < n.n; ++i) emit_compare(n.col, op, args_[n.arg + i], chain);
        } else {
          emit_compare(n.col, op, args_[n.arg], 0);
          for (uint32_t i = 1; i < n.n; ++i) emit_compare(n.col, op, args_[n.arg + i], chain);
          if (mode != 0) combine(mode);
        }
      } break;
      default: {  // AND / OR
        const int chain = n.kind == VGPU_NODE_AND ? 1 : 2;
        std::vector<int> kids = n.kids;
        std::vector<char> done(kids.size(), 0);
        bool first = mode != chain;  // the first emitted child must PUSH unless we chain into the top
        auto child_mode = [&]() { int m = first ? 0 : chain; first = false; return m; };
        if (chain == 1) {
          // peephole: lo <= x < hi on one small integer column -> one subtract-and-compare
          for (size_t i = 0; i < kids.size(); ++i) {
            if (done[i]) continue;
            const TreeNode &a = tree[kids[i]];
            if (a.kind != VGPU_NODE_RELOP || !is_small_int_col(a.col)) continue;
            const bool a_lo = a.op == VGPU_OP_GE || a.op == VGPU_OP_GT;
            const bool a_hi = a.op == VGPU_OP_LT || a.op == VGPU_OP_LE;
            if (!a_lo && !a_hi) continue;
            for (size_t j = i + 1; j < kids.size(); ++j) {
              if (done[j]) continue;
              const TreeNode &b = tree[kids[j]];
              if (b.kind != VGPU_NODE_RELOP || b.col != a.col) continue;
              const bool b_lo = b.op == VGPU_OP_GE || b.op == VGPU_OP_GT;
              const bool b_hi = b.op == VGPU_OP_LT || b.op == VGPU_OP_LE;
              if (!((a_lo && b_hi) || (a_hi && b_lo))) continue;
              const TreeNode &lo = a_lo ? a : b, &hi = a_lo ? b : a;
              const ColInfo &ci = t->cols[a.col];
              uint64_t lo_o = ord32(widen_arg(args_[lo.arg], ci.type), ci.type);
              uint64_t hi_o = ord32(widen_arg(args_[hi.arg], ci.type), ci.type);
              if (lo.op == VGPU_OP_GT) lo_o += 1;
              if (hi.op == VGPU_OP_LE) hi_o += 1;  // exclusive upper bound, may be ordmax+1
              PInstr in{};
              in.kind = leaf_kind(child_mode());
              if (in.kind == P_PUSH) push_depth();
              in.slot = (uint8_t)slot_of(a.col);
              if (hi_o <= lo_o || lo_o > ordmax32(ci.type)) {
                in.cls = C_FALSE;
              } else {
                in.cls = C_RNG32;
                in.bias = bias32(ci.type);
                in.arg = (uint32_t)lo_o;
                uint64_t len = hi_o - lo_o;  // <= 2^32
                if (len > 0xffffffffull) { in.cls = C_TRUE; } else in.arg2 = (uint32_t)len;
              }
              emit(in);
              done[i] = done[j] = 1;
              break;
            }
          }
        }
        for (size_t i = 0; i < kids.size(); ++i) {
          if (done[i]) continue;
          emit_node(kids[i], child_mode());
        }
        if (mode != 0 && mode != chain) combine(mode);
      } break;
    }
  }
  void combine(int mode) {
    PInstr in{};
    in.kind = mode == 1 ? P_AND : P_OR;
    emit(in);
    --depth;
  }

  void build_predicate() {
    build_tree();
    emit_node(root, 0);
    if (max_depth > kStackDepth) fail(VGPU_ERR_UNSUPPORTED, "predicate nesting too deep");
  }

  // predicate program -> the columns it streams (bulk L2 prefetch table) and the unrolled-conjunction flag
  void finish_predicate(uint32_t tune) {
    for (uint32_t i = 0; i < P.nprog; ++i) {
      const PInstr &in = P.prog[i];
      if (in.kind > P_OR_LEAF || !(in.cls == C_EQ32 || in.cls == C_LT32 || in.cls == C_RNG32 || in.cls == C_LUT64)) continue;
      bool seen = false;
      for (uint32_t f = 0; f < P.nfilter_slots; ++f) seen = seen || P.filter_slots[f] == in.slot;
      if (!seen) P.filter_slots[P.nfilter_slots++] = in.slot;
    }
    for (uint32_t f = 0; f < P.nfilter_slots; ++f) {
      P.pf_width[f] = (uint8_t)P.slots[P.filter_slots[f]].width;
      P.pf_off[f] = P.slots[P.filter_slots[f]].off;
    }
    // conjunction of at most 4 vectorisable leaves: the kernel's unrolled fast path
    P.conj = P.nprog >= 1 && P.nprog <= 4 && !(tune & 4096u);
    for (uint32_t i = 0; i < P.nprog && P.conj; ++i) {
      const PInstr &in = P.prog[i];
      const bool vec = in.cls == C_EQ32 || in.cls == C_LT32 || in.cls == C_RNG32 || in.cls == C_LUT64;
      if (!vec || in.kind != (i == 0 ? P_PUSH : P_AND_LEAF)) P.conj = 0;
    }
  }

  // ---- segment pruning: SegmentSkipBuilder (filter.cc:263-335), evaluated on the host ----
  template <class SegT>
  bool skip_leaf(const SegT &sd, uint32_t col, uint32_t op, uint64_t raw_arg) const {
    const ColInfo &ci = t->cols[col];
    if (!(ci.kind == VGPU_DIM_NUMERIC || ci.kind == VGPU_DIM_TIME || ci.kind == VGPU_DIM_MICROTIME))
      return true;
    // SegmentStats: dmax starts at cpp_min_value, dmin at cpp_max_value (store.cc:171-184)
    uint64_t dmax = to_ordered_host(type_min_value(ci.type), ci.type);
    uint64_t dmin = to_ordered_host(type_max_value(ci.type), ci.type);
    if (sd.nrows > 0) {
      dmax = std::max(dmax, sd.omax[col]);
      dmin = std::min(dmin, sd.omin[col]);
    }
    uint64_t v = to_ordered_host(widen_arg(raw_arg, ci.type), ci.type);
    switch (op) {
      case VGPU_OP_EQ: return dmin <= v && dmax >= v;
      case VGPU_OP_LT: case VGPU_OP_LE: return dmin <= v;
      case VGPU_OP_GT: case VGPU_OP_GE: return dmax >= v;
      default: return true;
    }
  }
  template <class SegT>
  bool process_segment(const SegT &sd, int idx) const {
    const TreeNode &n = tree[idx];
    switch (n.kind) {
      case VGPU_NODE_EMPTY: return true;
      case VGPU_NODE_RELOP: return skip_leaf(sd, n.col, n.op, args_[n.arg]);
      case VGPU_NODE_IN: {
        const ColInfo &ci = t->cols[n.col];
        if (!(ci.kind == VGPU_DIM_NUMERIC || ci.kind == VGPU_DIM_TIME || ci.kind == VGPU_DIM_MICROTIME))
          return true;
        // NOT IN is pruned with the same "some value inside [min,max]" test (filter.cc:303-327
        // ignores equal()) — reproduced as is
        bool r = false;
        for (uint32_t i = 0; i < n.n; ++i) r = r || skip_leaf(sd, n.col, VGPU_OP_EQ, args_[n.arg + i]);
        return r;
      }
      case VGPU_NODE_AND: {
        bool r = true;
        for (int k : n.kids) r = process_segment(sd, k) && r;
        return r;
      }
      default: {
        bool r = false;
        for (int k : n.kids) r = process_segment(sd, k) || r;
        return r;
      }
    }
  }
};

// A top-level conjunction (ScanParams::conj) restricts what a key column can hold in passing rows: tighten the key's
// domain [lo, hi] from the leaves on that column (unsigned keys of at most 4 bytes, no rollup; leaf arguments are raw
// zero-extended values). Returns the lookup mask when an IN list restricts the key to codes < 64 (the key is then
// numbered by its rank in the mask, KeySpec::lut), else 0.
inline uint64_t tighten_key_domain(const ScanParams &P, uint32_t slot, uint64_t &lo, uint64_t &hi) {
  uint64_t lut = 0;
  for (uint32_t i = 0; i < P.nprog; ++i) {
    const PInstr &in = P.prog[i];
    if (in.slot != slot || in.neg) continue;
    const uint64_t a = (uint32_t)in.arg;
    if (in.cls == C_EQ32) { lo = std::max(lo, a); hi = std::min(hi, a); }
    else if (in.cls == C_RNG32 && in.bias == 0) { lo = std::max(lo, a); hi = std::min(hi, a + in.arg2 - 1); }
    else if (in.cls == C_LT32 && in.bias == 0 && a > 0) { hi = std::min(hi, a - 1); }
    else if (in.cls == C_LUT64) lut = lut ? (lut & in.arg) : in.arg;
  }
  if (lo > hi) hi = lo;  // nothing can pass: any one-value domain will do
  if (lut) {
    for (uint32_t b = 0; b < 64; ++b)
      if (b < lo || b > hi) lut &= ~(1ull << b);
  }
  return lut;
}

}  // namespace

#endif  // VGPU_PLANNER_H_
