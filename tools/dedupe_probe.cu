// tools/dedupe_probe.cu — design probe for the count-distinct dedupe (C2 shape: 2.5e7 (cell,id) pairs, 5e5 cells,
// ids < 1e6). Times, on the real chip:
//   red_only        one red.global.add.u32 per pair into distinct[cell]            (L2 RED throughput, 2 MB table)
//   cas_global      one atomicCAS per pair into a 2^26-slot global set             (L2 CAS throughput, DRAM-resident set)
//   partition       ragged regions -> NB cell-range buckets (tile histogram in shared memory, global cursors)
//   dedupe_cas      one bucket per CTA iteration: shared-memory open-addressing set, atomicCAS(64) + global RED
//   dedupe_plain    same, but the set is filled with plain stores + CTA barriers (no shared-memory atomics)
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/dedupe_probe tools/dedupe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
constexpr uint64_t kEmpty = ~0ull;

__global__ void gen_kernel(uint64_t *regions, uint32_t *counts, uint32_t nreg, uint32_t cap, uint32_t per, uint32_t cells, uint32_t ids) {
  for (uint32_t r = blockIdx.x; r < nreg; r += gridDim.x) {
    for (uint32_t i = threadIdx.x; i < per; i += blockDim.x) {
      const uint64_t g = (uint64_t)r * per + i;
      const uint64_t cell = mix64(g * 2 + 1) % cells, id = mix64(g * 2 + 2) % ids;
      regions[(uint64_t)r * cap + i] = (cell << 32) | id;
    }
    if (threadIdx.x == 0) counts[r] = per;
  }
}

__global__ void red_only_kernel(const uint64_t *regions, const uint32_t *counts, uint32_t nreg, uint32_t cap, uint32_t *distinct) {
  for (uint32_t r = blockIdx.x; r < nreg; r += gridDim.x) {
    const uint64_t *src = regions + (uint64_t)r * cap;
    const uint32_t n = counts[r];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(distinct + (src[i] >> 32), 1u);
  }
}

__global__ void cas_global_kernel(const uint64_t *regions, const uint32_t *counts, uint32_t nreg, uint32_t cap, uint64_t *set, uint64_t mask,
                                  uint32_t *distinct, int do_red) {
  for (uint32_t r = blockIdx.x; r < nreg; r += gridDim.x) {
    const uint64_t *src = regions + (uint64_t)r * cap;
    const uint32_t n = counts[r];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      const uint64_t key = src[i];
      uint64_t slot = mix64(key) & mask;
      while (true) {
        unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(set + slot), (unsigned long long)kEmpty, (unsigned long long)key);
        if (old == kEmpty) { if (do_red) atomicAdd(distinct + (key >> 32), 1u); break; }
        if (old == key) break;
        slot = (slot + 1) & mask;
      }
    }
  }
}

// ---- partition: tile of kTile pairs per CTA iteration, histogram + ranks in shared memory, one global atomic per
// non-empty bucket per tile ----
constexpr int kPThreads = 1024;
template <int kPer>
__global__ void __launch_bounds__(kPThreads) partition_kernel(const uint64_t *regions, const uint32_t *counts, uint32_t nreg, uint32_t cap,
                                                              uint32_t nb, uint32_t cpb, uint32_t bucket_cap, uint32_t *cursors, uint64_t *out,
                                                              uint32_t *overflow) {
  extern __shared__ uint32_t s_cnt[];  // [nb] counts, then [nb] bases
  uint32_t *s_base = s_cnt + nb;
  // a CTA walks a contiguous range of regions as one virtual stream of tiles
  for (uint32_t r = blockIdx.x; r < nreg; r += gridDim.x) {
    const uint64_t *src = regions + (uint64_t)r * cap;
    const uint32_t n = counts[r];
    for (uint32_t t0 = 0; t0 < n; t0 += kPThreads * kPer) {
      for (uint32_t i = threadIdx.x; i < nb; i += kPThreads) s_cnt[i] = 0;
      __syncthreads();
      uint64_t key[kPer];
      uint32_t bkt[kPer], pos[kPer];
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const uint32_t i = t0 + j * kPThreads + threadIdx.x;
        bkt[j] = 0xffffffffu;
        if (i < n) {
          key[j] = src[i];
          bkt[j] = (uint32_t)(key[j] >> 32) / cpb;
          pos[j] = atomicAdd(&s_cnt[bkt[j]], 1u);
        }
      }
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < nb; i += kPThreads)
        if (s_cnt[i]) s_base[i] = atomicAdd(&cursors[i], s_cnt[i]);
      __syncthreads();
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        if (bkt[j] == 0xffffffffu) continue;
        const uint32_t p = s_base[bkt[j]] + pos[j];
        if (p < bucket_cap) out[(uint64_t)bkt[j] * bucket_cap + p] = key[j];
        else *overflow = 1;
      }
      __syncthreads();
    }
  }
}

// ---- shared-memory set dedupe, atomicCAS flavour ----
constexpr int kDThreads = 1024;
__global__ void __launch_bounds__(kDThreads) dedupe_cas_kernel(const uint64_t *buckets, const uint32_t *cursors, uint32_t nb, uint32_t bucket_cap,
                                                               uint32_t nslots, uint32_t *distinct, uint32_t *overflow) {
  extern __shared__ __align__(16) uint64_t s_set[];
  for (uint32_t b = blockIdx.x; b < nb; b += gridDim.x) {
    const uint32_t n = min(cursors[b], bucket_cap);
    for (uint32_t i = threadIdx.x; i < nslots; i += kDThreads) s_set[i] = kEmpty;
    __syncthreads();
    if (n > nslots - nslots / 4) { if (threadIdx.x == 0) *overflow = 1; __syncthreads(); continue; }
    const uint64_t *src = buckets + (uint64_t)b * bucket_cap;
    for (uint32_t i = threadIdx.x; i < n; i += kDThreads) {
      const uint64_t key = src[i];
      uint32_t slot = (uint32_t)(((mix64(key) >> 32) * nslots) >> 32);
      while (true) {
        unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(s_set + slot), (unsigned long long)kEmpty, (unsigned long long)key);
        if (old == kEmpty) { atomicAdd(distinct + (key >> 32), 1u); break; }
        if (old == key) break;
        slot = slot + 1 == nslots ? 0 : slot + 1;
      }
    }
    __syncthreads();
  }
}

// local counts: the bucket owns its cells, distinct[cell] is a plain store of a shared-memory counter
template <int kT>
__global__ void __launch_bounds__(kT) dedupe_cas_local_kernel(const uint64_t *buckets, const uint32_t *cursors, uint32_t nb, uint32_t bucket_cap,
                                                               uint32_t nslots, uint32_t cpb, uint32_t cells, uint32_t *distinct, uint32_t *overflow) {
  extern __shared__ __align__(16) uint64_t s_set[];
  uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_set + nslots);
  for (uint32_t b = blockIdx.x; b < nb; b += gridDim.x) {
    const uint32_t n = min(cursors[b], bucket_cap);
    for (uint32_t i = threadIdx.x; i < nslots; i += kT) s_set[i] = kEmpty;
    for (uint32_t i = threadIdx.x; i < cpb; i += kT) s_cnt[i] = 0;
    __syncthreads();
    if (n > nslots - nslots / 4) { if (threadIdx.x == 0) *overflow = 1; __syncthreads(); continue; }
    const uint64_t *src = buckets + (uint64_t)b * bucket_cap;
    const uint32_t cell0 = b * cpb;
    for (uint32_t i = threadIdx.x; i < n; i += kT) {
      const uint64_t key = src[i];
      uint32_t slot = (uint32_t)(((mix64(key) >> 32) * nslots) >> 32);
      while (true) {
        unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(s_set + slot), (unsigned long long)kEmpty, (unsigned long long)key);
        if (old == kEmpty) { atomicAdd(s_cnt + ((uint32_t)(key >> 32) - cell0), 1u); break; }
        if (old == key) break;
        slot = slot + 1 == nslots ? 0 : slot + 1;
      }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < cpb; i += kT) if (cell0 + i < cells) distinct[cell0 + i] = s_cnt[i];
    __syncthreads();
  }
}

// ---- shared-memory set dedupe, plain loads / stores + CTA barriers ----
// Per probe round and key: (1) read the slot, nothing is being written: equal -> duplicate, done; occupied by another
// key (stable for ever) -> next slot; empty -> candidate. (2) candidates store their key (some store wins). (3) read
// back: another key won -> next slot; own key there -> contest for the slot's owner word (plain store of the thread
// id). (4) the owner counts the key. A thread carries kPer keys through the rounds at once.
template <int kPer>
__global__ void __launch_bounds__(kDThreads) dedupe_plain_kernel(const uint64_t *buckets, const uint32_t *cursors, uint32_t nb, uint32_t bucket_cap,
                                                                 uint32_t nslots, uint32_t *distinct, uint32_t *overflow) {
  extern __shared__ __align__(16) uint64_t s_set[];
  uint16_t *s_owner = reinterpret_cast<uint16_t *>(s_set + nslots);
  for (uint32_t b = blockIdx.x; b < nb; b += gridDim.x) {
    const uint32_t n = min(cursors[b], bucket_cap);
    for (uint32_t i = threadIdx.x; i < nslots; i += kDThreads) s_set[i] = kEmpty;
    __syncthreads();
    if (n > nslots - nslots / 4) { if (threadIdx.x == 0) *overflow = 1; __syncthreads(); continue; }
    const uint64_t *src = buckets + (uint64_t)b * bucket_cap;
    for (uint32_t t0 = 0; t0 < n; t0 += kDThreads * kPer) {
      uint64_t key[kPer];
      uint32_t slot[kPer];
      uint32_t live = 0;   // bit j: key j still looking for its slot
      uint32_t cand = 0;
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const uint32_t i = t0 + j * kDThreads + threadIdx.x;
        if (i < n) {
          key[j] = src[i];
          slot[j] = (uint32_t)(((mix64(key[j]) >> 32) * nslots) >> 32);
          live |= 1u << j;
        }
      }
      while (__syncthreads_or(live != 0)) {   // the barrier also closes the previous round's stores
        cand = 0;
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
          if (!(live & (1u << j))) continue;
          const uint64_t v = s_set[slot[j]];
          if (v == key[j]) live &= ~(1u << j);
          else if (v == kEmpty) cand |= 1u << j;
          else slot[j] = slot[j] + 1 == nslots ? 0 : slot[j] + 1;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kPer; ++j)
          if (cand & (1u << j)) s_set[slot[j]] = key[j];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
          if (!(cand & (1u << j))) continue;
          if (s_set[slot[j]] == key[j]) s_owner[slot[j]] = (uint16_t)threadIdx.x;
          else { cand &= ~(1u << j); slot[j] = slot[j] + 1 == nslots ? 0 : slot[j] + 1; }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
          if (!(cand & (1u << j))) continue;
          // same thread may hold the same key twice: the lower j counts
          bool first = true;
#pragma unroll
          for (int q = 0; q < j; ++q) first = first && !((cand & (1u << q)) && key[q] == key[j]);
          if (s_owner[slot[j]] == (uint16_t)threadIdx.x && first) atomicAdd(distinct + (key[j] >> 32), 1u);
          live &= ~(1u << j);
        }
      }
    }
    __syncthreads();
  }
}

int main(int argc, char **argv) {
  const uint32_t cells = 500000, ids = argc > 1 ? atoi(argv[1]) : 1000000;
  const uint32_t nreg = 444, per = 56306;  // 2.5e7 pairs
  const uint32_t cap = per + per / 4 + 1024;
  const uint64_t N = (uint64_t)nreg * per;
  uint64_t *regions; uint32_t *counts, *d1, *d2, *d3, *d4, *cursors, *overflow;
  CK(cudaMalloc(&regions, (uint64_t)nreg * cap * 8));
  CK(cudaMalloc(&counts, nreg * 4));
  CK(cudaMalloc(&d1, cells * 4)); CK(cudaMalloc(&d2, cells * 4)); CK(cudaMalloc(&d3, cells * 4)); CK(cudaMalloc(&d4, cells * 4));
  uint64_t *buckets = nullptr;
  CK(cudaMalloc(&cursors, 8192 * 4)); CK(cudaMalloc(&overflow, 4));
  const uint64_t gset = 1ull << 26;
  uint64_t *set; CK(cudaMalloc(&set, gset * 8));
  gen_kernel<<<444, 1024>>>(regions, counts, nreg, cap, per, cells, ids);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char *name, auto &&fn, auto &&prep) {
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      prep();
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      fn();
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
    printf("%-36s %8.3f ms  (%.3g pairs/s)\n", name, best, N / (best * 1e-3));
  };
  timeit("red_only", [&] { red_only_kernel<<<444, 1024>>>(regions, counts, nreg, cap, d1); }, [&] { cudaMemset(d1, 0, cells * 4); });
  uint32_t h_over = 0;
  char nm[96];
  for (uint32_t NB : {512u, 1024u, 2048u, 4096u}) {
    const uint32_t cpb = (cells + NB - 1) / NB;
    const uint32_t bucket_cap = (uint32_t)(N / NB + N / NB / 4 + 1024);
    if (buckets) cudaFree(buckets);
    CK(cudaMalloc(&buckets, (uint64_t)NB * bucket_cap * 8));
    CK(cudaFuncSetAttribute(partition_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * NB * 4));
    auto prep_part = [&] { cudaMemset(cursors, 0, NB * 4); cudaMemset(overflow, 0, 4); };
    snprintf(nm, sizeof nm, "partition<8> NB=%u", NB);
    timeit(nm, [&] { partition_kernel<8><<<148, kPThreads, 2 * NB * 4>>>(regions, counts, nreg, cap, NB, cpb, bucket_cap, cursors, buckets, overflow); }, prep_part);
    if (NB < 2048) continue;
    for (uint32_t nslots : {16384u, 24576u}) {
      if (NB == 2048 && nslots == 16384u) continue;
      const size_t smem_c = (size_t)nslots * 8, smem_l = smem_c + cpb * 4;
      CK(cudaFuncSetAttribute(dedupe_cas_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
      CK(cudaFuncSetAttribute(dedupe_cas_local_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));
      snprintf(nm, sizeof nm, "dedupe_cas NB=%u slots=%u", NB, nslots);
      timeit(nm, [&] { dedupe_cas_kernel<<<148, kDThreads, smem_c>>>(buckets, cursors, NB, bucket_cap, nslots, d2, overflow); }, [&] { cudaMemset(d2, 0, cells * 4); });
      snprintf(nm, sizeof nm, "dedupe_cas_local NB=%u slots=%u", NB, nslots);
      timeit(nm, [&] { dedupe_cas_local_kernel<1024><<<148, 1024, smem_l>>>(buckets, cursors, NB, bucket_cap, nslots, cpb, cells, d3, overflow); }, [&] { cudaMemset(d3, 0, cells * 4); });
    }
    if (NB == 4096) {   // two CTAs per SM, half-size sets
      const uint32_t nslots = 12288;
      const size_t smem_l = (size_t)nslots * 8 + cpb * 4;
      CK(cudaFuncSetAttribute(dedupe_cas_local_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));
      snprintf(nm, sizeof nm, "dedupe_cas_local 2x512 slots=%u", nslots);
      timeit(nm, [&] { dedupe_cas_local_kernel<512><<<296, 512, smem_l>>>(buckets, cursors, NB, bucket_cap, nslots, cpb, cells, d4, overflow); }, [&] { cudaMemset(d4, 0, cells * 4); });
      CK(cudaFuncSetAttribute(dedupe_cas_local_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));
      snprintf(nm, sizeof nm, "dedupe_cas_local 2x1024 slots=%u", nslots);
      timeit(nm, [&] { dedupe_cas_local_kernel<1024><<<296, 1024, smem_l>>>(buckets, cursors, NB, bucket_cap, nslots, cpb, cells, d4, overflow); }, [&] { cudaMemset(d4, 0, cells * 4); });
    }
  }
  CK(cudaMemcpy(&h_over, overflow, 4, cudaMemcpyDeviceToHost));
  printf("overflow=%u\n", h_over);
  // verify: cas_global+red (d1) vs smem variants
  cudaMemset(set, 0xff, gset * 8); cudaMemset(d1, 0, cells * 4);
  cas_global_kernel<<<444, 1024>>>(regions, counts, nreg, cap, set, gset - 1, d1, 1);
  std::vector<uint32_t> h1(cells), h2(cells), h3(cells), h4(cells);
  CK(cudaMemcpy(h1.data(), d1, cells * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h2.data(), d2, cells * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h3.data(), d3, cells * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h4.data(), d4, cells * 4, cudaMemcpyDeviceToHost));
  uint64_t tot = 0, bad2 = 0, bad3 = 0, bad4 = 0;
  for (uint32_t c = 0; c < cells; ++c) { tot += h1[c]; bad2 += h1[c] != h2[c]; bad3 += h1[c] != h3[c]; bad4 += h1[c] != h4[c]; }
  printf("distinct total %llu of %llu pairs; mismatching cells: cas %llu plain8 %llu plain4 %llu\n", (unsigned long long)tot, (unsigned long long)N,
         (unsigned long long)bad2, (unsigned long long)bad3, (unsigned long long)bad4);
  return 0;
}
