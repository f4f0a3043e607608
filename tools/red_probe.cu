// tools/red_probe.cu — do REDs on a small table stay in L2 while a large column streams through it?
//   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_red_lookup_miss.sum ./red_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
// MODE bit0: stream the big buffer; bit1: do REDs (1 per 40 streamed rows); bit2: stream with evict_first policy
template <int MODE>
__global__ void probe(const uint4 *big, uint64_t n16, int *table, uint32_t cells, unsigned long long *out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  uint32_t acc = 0;
  for (; i < n16; i += stride) {
    if (MODE & 1) {
      uint4 v;
      if (MODE & 4) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(big + i), "l"(pol));
      else asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(big + i));
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (MODE & 2) {
      uint64_t h = mix(i);
      if (h % 10 == 0) atomicMin(table + (h >> 8) % cells, (int)(h >> 40));   // 1 RED per 10 x 16 B = per 40 u32 rows
    }
  }
  if (acc == 0x12345678) atomicAdd(out, 1ull);
}
int main() {
  const uint64_t n16 = 1ull << 28;  // 4 GiB
  uint4 *big; int *table; unsigned long long *out;
  cudaMalloc(&big, n16 * 16); cudaMalloc(&out, 8);
  const uint32_t cells = 1u << 20;  // 4 MB table
  cudaMalloc(&table, cells * 4);
  cudaMemset(big, 1, n16 * 16); cudaMemset(table, 0x7f, cells * 4);
  probe<1><<<148 * 8, 256>>>(big, n16, table, cells, out);   // stream only
  probe<2><<<148 * 8, 256>>>(big, n16, table, cells, out);   // REDs only
  probe<3><<<148 * 8, 256>>>(big, n16, table, cells, out);   // both
  probe<7><<<148 * 8, 256>>>(big, n16, table, cells, out);   // both, stream evict_first
  cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
