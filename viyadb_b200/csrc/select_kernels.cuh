// select_kernels.cuh — select and search queries on the scan core.
//
// Reference semantics (paths relative to the viyadb/viyadb tree):
//   select   src/codegen/query/scan.cc:75-166   every passing row, in (segment, tuple) order, skip / limit
//   search   src/codegen/query/scan.cc:249-299  distinct values of one dimension over the passing rows, in
//                                               first-seen order
// Both share IterationStart (scan.cc:40-69) — segment pruning and the row predicate — with the aggregate
// query, so the kernels below reuse eval_predicate and the ScanParams predicate program unchanged. Output
// ORDER is defined for these queries, which the fused aggregate kernel never needed:
//   rows_count_kernel   passing rows per 512-row chunk
//   rows_write_kernel   row numbers of the passing rows whose per-segment ordinal falls in a wanted range,
//                       written at their ordinal (the host turns chunk counts into ordinals and applies the
//                       reference's skip / limit rules, including "limit only breaks the tuple loop")
//   rows_gather_kernel  cells of the selected columns for those rows, SoA
//   search_first_kernel per segment and code: the first row that holds it (atomicMin into a dense array)
//   search_emit_kernel  (code, first row) pairs of one batch of segments, compacted
#ifndef VGPU_SELECT_KERNELS_CUH_
#define VGPU_SELECT_KERNELS_CUH_

#include "scan_kernel.cuh"

namespace vgpu {

struct RowsParams2 {                // on top of the ScanParams predicate
  uint32_t *chunk_counts;           // [nactive * tiles_per_seg]
  const uint64_t *chunk_ord;        // [nactive * tiles_per_seg] ordinal of the chunk's first passing row in its segment
  const uint64_t *want_lo;          // [nactive] first wanted ordinal of the segment
  const uint64_t *want_n;           // [nactive] wanted rows of the segment (0: none)
  const uint64_t *out_base;         // [nactive] position of the segment's first wanted row in the output
  uint32_t *out_row;                // [total wanted]
  uint32_t *out_seg;                // [total wanted] table segment index
};

__device__ __forceinline__ void load_cur(const ScanParams &P, uint32_t si, CurSeg &seg) {
  seg.index = P.active[si];
  const SegDesc &sd = P.segs[seg.index];
  seg.slab = sd.slab;
  seg.rows = sd.rows;
  seg.cap = (uint32_t)sd.cap;
  seg.nrows = (uint32_t)sd.nrows;
}

// mask of the chunk's passing rows (bit s*4+j of lane l = row s*128 + l*4 + j), rows past the end cleared
__device__ __forceinline__ uint32_t chunk_mask(const ScanParams &P, const CurSeg &seg, uint32_t chunk_row, uint32_t lane,
                                               uint64_t pol) {
  const uint32_t row0 = chunk_row + lane * kVec;
  uint32_t mask = eval_predicate(P, seg, row0, pol);
  if (chunk_row + kChunkRows > seg.nrows) {
#pragma unroll
    for (int s = 0; s < kSub; ++s)
#pragma unroll
      for (int j = 0; j < kVec; ++j)
        if (row0 + s * kSubChunk + j >= seg.nrows) mask &= ~(1u << (s * 4 + j));
  }
  return mask;
}

// one warp per chunk, chunks interleaved over the grid (order does not matter here)
__global__ void __launch_bounds__(kThreads) rows_count_kernel(const __grid_constant__ ScanParams P, RowsParams2 R) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t nwarps = (uint64_t)gridDim.x * kWarps;
  const uint64_t pol = make_stream_policy(true);
  for (uint64_t g = (uint64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); g < P.total_tiles; g += nwarps) {
    const uint32_t si = (uint32_t)(g / P.tiles_per_seg), ci = (uint32_t)(g - (uint64_t)si * P.tiles_per_seg);
    CurSeg seg;
    load_cur(P, si, seg);
    uint32_t n = 0;
    if (ci * kChunkRows < seg.nrows) n = __popc(chunk_mask(P, seg, ci * kChunkRows, lane, pol));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_down_sync(0xffffffffu, n, o);
    if (lane == 0) R.chunk_counts[g] = n;
  }
}

__global__ void __launch_bounds__(kThreads) rows_write_kernel(const __grid_constant__ ScanParams P, RowsParams2 R) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t nwarps = (uint64_t)gridDim.x * kWarps;
  const uint64_t pol = make_stream_policy(true);
  for (uint64_t g = (uint64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); g < P.total_tiles; g += nwarps) {
    const uint32_t si = (uint32_t)(g / P.tiles_per_seg), ci = (uint32_t)(g - (uint64_t)si * P.tiles_per_seg);
    const uint64_t want_n = R.want_n[si];
    const uint32_t cnt = R.chunk_counts[g];
    if (want_n == 0 || cnt == 0) continue;
    const uint64_t lo = R.want_lo[si], hi = lo + want_n, ord0 = R.chunk_ord[g];
    if (ord0 + cnt <= lo || ord0 >= hi) continue;  // uniform per warp
    CurSeg seg;
    load_cur(P, si, seg);
    const uint32_t chunk_row = ci * kChunkRows;
    const uint32_t mask = chunk_mask(P, seg, chunk_row, lane, pol);
    // ordinals in ROW order: sub-chunk s before s+1, inside a sub-chunk lane l before l+1, inside a lane bit j
    uint32_t packed = __popc(mask & 0xfu) | (__popc(mask & 0xf0u) << 8) | (__popc(mask & 0xf00u) << 16) |
                      (__popc(mask & 0xf000u) << 24);  // per-sub-chunk counts, 8 bits each (totals <= 128)
    uint32_t incl = packed;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t excl = incl - packed;
    uint32_t base = 0;
#pragma unroll
    for (int s = 0; s < kSub; ++s) {
      uint32_t ord_in_chunk = base + ((excl >> (8 * s)) & 0xffu);
#pragma unroll
      for (int j = 0; j < kVec; ++j) {
        if (mask & (1u << (s * 4 + j))) {
          const uint64_t ord = ord0 + ord_in_chunk++;
          if (ord >= lo && ord < hi) {
            const uint64_t pos = R.out_base[si] + (ord - lo);
            R.out_row[pos] = chunk_row + s * kSubChunk + lane * kVec + j;
            R.out_seg[pos] = seg.index;
          }
        }
      }
      base += (tot >> (8 * s)) & 0xffu;
    }
  }
}

struct GatherCol {
  uint64_t col_off;     // bytes per row of the preceding fixed-width columns (column base = slab + col_off * cap)
  uint32_t width;       // 1, 2, 4, 8
  uint32_t bitset;      // BITSET column: the output is the cell's cardinality (uint64)
  uint32_t bitset_idx;
  uint32_t pad;
  void *out;            // [nrows] elements of `width` bytes (uint64 for BITSET)
};
struct GatherParams {
  const SegDesc *segs;
  const uint32_t *out_row, *out_seg;
  uint64_t nrows;
  uint32_t ncols;
  GatherCol cols[32];
};

__global__ void __launch_bounds__(256) rows_gather_kernel(const __grid_constant__ GatherParams G) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < G.nrows; i += (uint64_t)gridDim.x * blockDim.x) {
    const SegDesc &sd = G.segs[G.out_seg[i]];
    const uint32_t row = G.out_row[i];
    for (uint32_t c = 0; c < G.ncols; ++c) {
      const GatherCol &gc = G.cols[c];
      if (gc.bitset) {  // ids of a cell are a set: its cardinality is the CSR row length
        const uint32_t *off = sd.bs_offsets[gc.bitset_idx];
        reinterpret_cast<uint64_t *>(gc.out)[i] = off == nullptr ? 1ull : (uint64_t)(off[row + 1] - off[row]);
        continue;
      }
      const uint8_t *src = sd.slab + gc.col_off * sd.cap + (uint64_t)row * gc.width;
      switch (gc.width) {
        case 1: reinterpret_cast<uint8_t *>(gc.out)[i] = *src; break;
        case 2: reinterpret_cast<uint16_t *>(gc.out)[i] = *reinterpret_cast<const uint16_t *>(src); break;
        case 4: reinterpret_cast<uint32_t *>(gc.out)[i] = *reinterpret_cast<const uint32_t *>(src); break;
        default: reinterpret_cast<uint64_t *>(gc.out)[i] = *reinterpret_cast<const uint64_t *>(src); break;
      }
    }
  }
}

// ---- search ----------------------------------------------------------------------------------------
struct SearchParams {
  uint32_t seg_begin, seg_end;  // active-segment slots of this batch
  uint64_t col_off;             // the dimension's column
  uint32_t width, sext;
  uint64_t lo;                  // smallest value of the dimension over the batch (sign-extended domain)
  uint64_t range;               // values lo .. lo + range - 1
  uint32_t *first;              // [(seg_end - seg_begin) * range], initialised to 0xffffffff
  // emit
  unsigned long long *cursor;   // entries written
  uint64_t cap;
  uint32_t *out_slot;           // active-segment slot
  uint64_t *out_code;           // raw value, widened
  uint32_t *out_first;          // first row
};

__global__ void __launch_bounds__(kThreads) search_first_kernel(const __grid_constant__ ScanParams P, SearchParams S) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t nwarps = (uint64_t)gridDim.x * kWarps;
  const uint64_t pol = make_stream_policy(true);
  const uint64_t g_begin = (uint64_t)S.seg_begin * P.tiles_per_seg, g_end = (uint64_t)S.seg_end * P.tiles_per_seg;
  for (uint64_t g = g_begin + (uint64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); g < g_end; g += nwarps) {
    const uint32_t si = (uint32_t)(g / P.tiles_per_seg), ci = (uint32_t)(g - (uint64_t)si * P.tiles_per_seg);
    CurSeg seg;
    load_cur(P, si, seg);
    if (ci * kChunkRows >= seg.nrows) continue;
    const uint32_t chunk_row = ci * kChunkRows;
    uint32_t mask = chunk_mask(P, seg, chunk_row, lane, pol);
    uint32_t *first = S.first + (uint64_t)(si - S.seg_begin) * S.range;
    while (mask) {
      const uint32_t b = __ffs(mask) - 1;
      mask &= mask - 1;
      const uint32_t row = chunk_row + (b >> 2) * kSubChunk + lane * kVec + (b & 3u);
      const uint64_t v = load_elem(seg.slab + S.col_off * seg.cap + (uint64_t)row * S.width, S.width, S.sext);
      atomicMin(first + (v - S.lo), row);
    }
  }
}

__global__ void __launch_bounds__(256) search_emit_kernel(const SearchParams S) {
  const uint64_t n = (uint64_t)(S.seg_end - S.seg_begin) * S.range;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t f = S.first[i];
    if (f == 0xffffffffu) continue;
    const unsigned long long p = atomicAdd(S.cursor, 1ull);
    if (p < S.cap) {
      S.out_slot[p] = S.seg_begin + (uint32_t)(i / S.range);
      S.out_code[p] = S.lo + i % S.range;
      S.out_first[p] = f;
    }
  }
}

}  // namespace vgpu

#endif  // VGPU_SELECT_KERNELS_CUH_
