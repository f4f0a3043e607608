// mock_vgpu.h — TEST INFRASTRUCTURE: a mock of the C ABI of include/vgpu.h, DEFINED here and linked instead of
// libvgpu.so by the CPU harnesses (tests/adapter_mock_harness.cc, tests/viyadb_patched_harness.cc). It computes nothing:
//   vgpu_table_create        records the schema
//   vgpu_segment_put_async / vgpu_segment_update   keep a shadow of what would sit in HBM (sync tests)
//   compare_with_live                              that shadow against the reference's live segments, cell by cell
//   vgpu_query_agg / _select / _search             record the plan the adapter lowered and answer with the canned
//                                                  result the harness placed in g_canned / g_sel_* / g_srch_*
// Include it in exactly one translation unit, after gpu_query_runner.h (it needs include/vgpu.h and nlohmann::json).
#ifndef VGPU_TESTS_MOCK_VGPU_H_
#define VGPU_TESTS_MOCK_VGPU_H_

#include <cstring>
#include <map>
#include <mutex>
#include <nlohmann/json.hpp>
#include <vector>

using json = nlohmann::json;

// ---------------------------------------------------------------------------------------------
// the mock device
// ---------------------------------------------------------------------------------------------
struct vgpu_ctx { int dummy; };
// the mock's "HBM": what put / update calls have written, per segment and schema column (fixed-width cells as bytes; a
// BITSET column as CSR offsets + id bytes)
struct ShadowSeg {
  uint64_t nrows = 0;
  std::vector<std::vector<uint8_t>> cols;
  std::vector<std::vector<uint64_t>> offsets;
};
struct vgpu_table {
  json schema;
  std::vector<uint32_t> kind, width;   // per schema column
  std::map<uint32_t, ShadowSeg> segs;
  json calls = json::array();          // the put / update calls seen, in order
};
struct vgpu_result { vgpu_result_view view; };
struct vgpu_rows {
  std::vector<std::vector<char>> bufs;
  std::vector<const void *> ptrs;
  vgpu_rows_view view;
};
struct vgpu_search { vgpu_search_view view; };

namespace {
vgpu_ctx g_ctx;
// per thread: several queries may run at once (the concurrent mode of adapter_mock_harness.cc), each with its own
// canned answer, like `query_threads` pool threads of the server
thread_local json g_plan;                 // the last plan vgpu_query_agg saw on this thread
thread_local vgpu_result_view g_canned;   // what it answers
std::mutex g_mock_mu;                     // call log + shadow store of the tables
json g_schema;
vgpu_table *g_table = nullptr;   // the last table created (the sync mode looks into its shadow store)
thread_local std::vector<std::vector<uint64_t>> g_sel_cells;   // select: per schema column, the widened cells of the rows to send
thread_local uint64_t g_sel_nrows = 0;
thread_local std::vector<uint64_t> g_srch_offsets{0}, g_srch_codes;   // search: what the device hands back
thread_local std::vector<uint32_t> g_srch_rows;

json nodes_json(const vgpu_pred_node *nodes, uint32_t n) {
  json a = json::array();
  for (uint32_t i = 0; i < n; ++i) a.push_back({nodes[i].kind, nodes[i].op, nodes[i].col, nodes[i].arg, nodes[i].n});
  return a;
}
}  // namespace

extern "C" {
const char *vgpu_last_error(void) { return "mock device"; }
int vgpu_init(int, vgpu_ctx **out) { *out = &g_ctx; return VGPU_OK; }
void vgpu_shutdown(vgpu_ctx *) {}
int vgpu_table_create(vgpu_ctx *, const vgpu_schema *schema, vgpu_table **out) {
  auto *t = new vgpu_table();
  json cols = json::array();
  for (uint32_t c = 0; c < schema->ncols; ++c)
    cols.push_back({schema->cols[c].kind, schema->cols[c].type, schema->cols[c].agg, schema->cols[c].lit_type});
  t->schema = {{"ncols", schema->ncols}, {"ndims", schema->ndims}, {"segment_size", schema->segment_size}, {"cols", cols}};
  static const uint32_t widths[10] = {1, 2, 4, 8, 1, 2, 4, 8, 4, 8};
  for (uint32_t c = 0; c < schema->ncols; ++c) {
    t->kind.push_back(schema->cols[c].kind);
    t->width.push_back(widths[schema->cols[c].type]);
  }
  g_schema = t->schema;
  g_table = t;
  *out = t;
  return VGPU_OK;
}
void vgpu_table_free(vgpu_table *t) { delete t; }
int vgpu_segment_put_async(vgpu_table *t, uint32_t seg, uint64_t nrows, const void *const *ptrs) {
  std::lock_guard<std::mutex> lk(g_mock_mu);
  t->calls.push_back({"put", seg, nrows});
  ShadowSeg &sh = t->segs[seg];
  sh = ShadowSeg();
  sh.nrows = nrows;
  sh.cols.resize(t->kind.size());
  sh.offsets.resize(t->kind.size());
  for (size_t c = 0; c < t->kind.size(); ++c) {
    if (t->kind[c] == VGPU_METRIC_BITSET) {
      auto *csr = static_cast<const vgpu_bitset_csr *>(ptrs[c]);
      if (csr->offsets) sh.offsets[c].assign(csr->offsets, csr->offsets + nrows + 1);
      const uint8_t *v = static_cast<const uint8_t *>(csr->values);
      sh.cols[c].assign(v, v + csr->nvalues * t->width[c]);
    } else {
      const uint8_t *v = static_cast<const uint8_t *>(ptrs[c]);
      sh.cols[c].assign(v, v + nrows * t->width[c]);
    }
  }
  return VGPU_OK;
}
int vgpu_segment_update(vgpu_table *t, uint32_t seg, uint64_t row_begin, uint64_t nrows, const void *const *ptrs) {
  std::lock_guard<std::mutex> lk(g_mock_mu);
  t->calls.push_back({"update", seg, row_begin, nrows});
  auto it = t->segs.find(seg);
  if (it == t->segs.end() || row_begin > it->second.nrows) return VGPU_ERR_STATE;
  ShadowSeg &sh = it->second;
  for (size_t c = 0; c < t->kind.size(); ++c) {
    if (t->kind[c] == VGPU_METRIC_BITSET) return VGPU_ERR_UNSUPPORTED;
    const uint64_t w = t->width[c];
    if (sh.cols[c].size() < (row_begin + nrows) * w) sh.cols[c].resize((row_begin + nrows) * w);
    std::memcpy(sh.cols[c].data() + row_begin * w, ptrs[c], nrows * w);
  }
  sh.nrows = std::max(sh.nrows, row_begin + nrows);
  return VGPU_OK;
}
int vgpu_table_sync(vgpu_table *) { return VGPU_OK; }
int vgpu_table_invalidate(vgpu_table *, uint32_t) { return VGPU_OK; }
int vgpu_host_pin(vgpu_ctx *, const void *, size_t) { return VGPU_OK; }
int vgpu_host_unpin(vgpu_ctx *, const void *) { return VGPU_OK; }
int vgpu_query_agg(vgpu_table *, const vgpu_plan *p, vgpu_result **out) {
  json keys = json::array();
  for (uint32_t k = 0; k < p->nkeys; ++k) {
    const vgpu_key &key = p->keys[k];
    json b = json::array(), g = json::array();
    for (uint32_t r = 0; r < key.nrules; ++r) { b.push_back(key.rule_boundary[r]); g.push_back(key.rule_granularity[r]); }
    keys.push_back({{"col", key.col}, {"nrules", key.nrules}, {"query_granularity", key.query_granularity},
                    {"rule_boundary", b}, {"rule_granularity", g}});
  }
  json args = json::array(), hargs = json::array(), mcols = json::array();
  for (uint32_t i = 0; i < p->nargs; ++i) args.push_back(p->args[i]);
  for (uint32_t i = 0; i < p->nhargs; ++i) hargs.push_back(p->hargs[i]);
  for (uint32_t i = 0; i < p->nmetrics; ++i) mcols.push_back(p->metric_cols[i]);
  g_plan = {{"nodes", nodes_json(p->nodes, p->nnodes)}, {"args", args}, {"keys", keys}, {"metric_cols", mcols},
            {"need_hidden_count", p->need_hidden_count}, {"flags", p->flags},
            {"hnodes", nodes_json(p->hnodes, p->nhnodes)}, {"hargs", hargs},
            {"sort_col", p->sort_col}, {"sort_descending", p->sort_descending}, {"top_k", p->top_k}};
  auto *r = new vgpu_result();
  r->view = g_canned;
  *out = r;
  return VGPU_OK;
}
int vgpu_result_get(const vgpu_result *r, vgpu_result_view *view) { *view = r->view; return VGPU_OK; }
void vgpu_result_free(vgpu_result *r) { delete r; }
// select: the rows the reference would send, as raw cells of the requested schema columns (vgpu.h: BITSET cells are
// cardinalities, uint64; the hidden count is uint64)
int vgpu_query_select(vgpu_table *, const vgpu_rows_plan *p, vgpu_rows **out) {
  json args = json::array(), cols = json::array();
  for (uint32_t i = 0; i < p->nargs; ++i) args.push_back(p->args[i]);
  for (uint32_t i = 0; i < p->ncols; ++i) cols.push_back(p->cols[i]);
  g_plan = {{"nodes", nodes_json(p->nodes, p->nnodes)}, {"args", args}, {"cols", cols}, {"skip", p->skip}, {"limit", p->limit}};
  auto *r = new vgpu_rows();
  r->bufs.resize(p->ncols);
  r->ptrs.resize(p->ncols);
  for (uint32_t i = 0; i < p->ncols; ++i) {
    const uint32_t c = p->cols[i];
    const uint32_t kind = g_schema["cols"][c][0].get<uint32_t>(), type = g_schema["cols"][c][1].get<uint32_t>();
    static const uint32_t widths[10] = {1, 2, 4, 8, 1, 2, 4, 8, 4, 8};
    const uint32_t w = (kind == VGPU_METRIC_BITSET || kind == VGPU_METRIC_HIDDEN_COUNT) ? 8u : widths[type];
    r->bufs[i].resize(g_sel_nrows * w + 8);
    for (uint64_t row = 0; row < g_sel_nrows; ++row) std::memcpy(r->bufs[i].data() + row * w, &g_sel_cells[c][row], w);
    r->ptrs[i] = r->bufs[i].data();
  }
  r->view = vgpu_rows_view{};
  r->view.nrows = g_sel_nrows;
  r->view.ncols = p->ncols;
  r->view.cells = r->ptrs.data();
  r->view.scanned_recs = g_canned.scanned_recs;
  r->view.scanned_segments = g_canned.scanned_segments;
  *out = r;
  return VGPU_OK;
}
int vgpu_rows_get(const vgpu_rows *r, vgpu_rows_view *view) { *view = r->view; return VGPU_OK; }
void vgpu_rows_free(vgpu_rows *r) { delete r; }
// search: per processed segment the distinct values of the dimension among the passing rows with their first rows
int vgpu_query_search(vgpu_table *, const vgpu_search_plan *p, vgpu_search **out) {
  json args = json::array();
  for (uint32_t i = 0; i < p->nargs; ++i) args.push_back(p->args[i]);
  g_plan = {{"nodes", nodes_json(p->nodes, p->nnodes)}, {"args", args}, {"col", p->col}};
  auto *r = new vgpu_search();
  r->view = vgpu_search_view{};
  r->view.nsegments = (uint32_t)(g_srch_offsets.size() - 1);
  r->view.seg_offsets = g_srch_offsets.data();
  r->view.codes = g_srch_codes.data();
  r->view.first_row = g_srch_rows.data();
  r->view.scanned_recs = g_canned.scanned_recs;
  r->view.scanned_segments = g_canned.scanned_segments;
  *out = r;
  return VGPU_OK;
}
int vgpu_search_get(const vgpu_search *r, vgpu_search_view *view) { *view = r->view; return VGPU_OK; }
void vgpu_search_free(vgpu_search *r) { delete r; }
}  // extern "C"

namespace {
// cells of the live store that differ from the shadow (0 == the resident copy is current)
json compare_with_live(viya::db::Table *table, vgpu_host::SegmentAccess &access) {
  namespace db = viya::db;
  uint64_t diff_cells = 0, rows = 0;
  bool missing = false;
  const size_t ndims = table->dimensions().size(), nmetrics = table->metrics().size();
  auto segments = table->store()->segments_copy();
  for (size_t si = 0; si < segments.size(); ++si) {
    const size_t size = segments[si]->size();
    rows += size;
    auto it = g_table->segs.find((uint32_t)si);
    if (it == g_table->segs.end() || it->second.nrows != size) { missing = true; continue; }
    const ShadowSeg &sh = it->second;
    std::vector<const void *> dims(ndims), metrics(nmetrics);
    std::vector<uint64_t> stats(2 * ndims + 2);
    const void *hidden = nullptr;
    access.columns()(segments[si], dims.data(), metrics.data(), &hidden, stats.data());
    size_t c = 0;
    auto cmp = [&](const void *live, size_t col) {
      const uint64_t w = g_table->width[col];
      const uint8_t *a = static_cast<const uint8_t *>(live);
      for (size_t r = 0; r < size; ++r)
        if (std::memcmp(a + r * w, sh.cols[col].data() + r * w, w) != 0) ++diff_cells;
    };
    for (size_t d = 0; d < ndims; ++d, ++c) cmp(dims[d], c);
    for (auto *m : table->metrics()) {
      if (m->agg_type() != db::Metric::AggregationType::BITSET) { cmp(metrics[m->index()], c++); continue; }
      std::vector<uint64_t> offsets(size + 1);
      const uint64_t total = access.bitset()(segments[si], m->index(), size, offsets.data(), nullptr);
      std::vector<uint64_t> wide(total + 1);
      access.bitset()(segments[si], m->index(), size, offsets.data(), wide.data());
      const uint64_t w = g_table->width[c];
      if (sh.offsets[c] != offsets || sh.cols[c].size() != total * w) { ++diff_cells; ++c; continue; }
      for (uint64_t i = 0; i < total; ++i) {
        uint64_t v = 0;
        std::memcpy(&v, sh.cols[c].data() + i * w, w);
        if (v != wide[i]) ++diff_cells;
      }
      ++c;
    }
    if (access.has_hidden_count()) cmp(hidden, c);
  }
  return {{"segments", segments.size()}, {"rows", rows}, {"differing_cells", diff_cells}, {"missing_or_short_segments", missing}};
}
}  // namespace

#endif  // VGPU_TESTS_MOCK_VGPU_H_
