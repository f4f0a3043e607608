// tools/gather_probe2.cu — DRAM bytes per gather for a SPARSE-SEQUENTIAL pattern (row r is read when
// hash(r) % 40 == 0, rows visited in order by consecutive threads), for several load flavours.
//   ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./gather_probe2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
template <int MODE>
__global__ void sparse(const uint32_t *a, uint64_t n, unsigned long long *out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint32_t acc = 0; unsigned long long cnt = 0;
  for (; i < n; i += stride) {
    if (mix(i) % 40 != 0) continue;
    const uint32_t *p = a + i;
    uint32_t v;
    if (MODE == 0) v = __ldg(p);
    else if (MODE == 1) asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (MODE == 2) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (MODE == 3) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (MODE == 4) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (MODE == 5) asm volatile("ld.global.cg.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (MODE == 6) asm volatile("ld.global.L1::no_allocate.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else { v = atomicAdd((unsigned int *)p, 0u); }
    acc += v; ++cnt;
  }
  if (acc == 0x12345678) atomicAdd(out, 1ull);
  atomicAdd(out + 1, cnt);
}
int main() {
  const uint64_t n = 1ull << 29;  // 2 GiB of uint32
  uint32_t *a; unsigned long long *out;
  cudaMalloc(&a, n * 4); cudaMalloc(&out, 16);
  cudaMemset(a, 1, n * 4); cudaMemset(out, 0, 16);
  sparse<0><<<148 * 8, 256>>>(a, n, out);
  sparse<1><<<148 * 8, 256>>>(a, n, out);
  sparse<2><<<148 * 8, 256>>>(a, n, out);
  sparse<3><<<148 * 8, 256>>>(a, n, out);
  sparse<4><<<148 * 8, 256>>>(a, n, out);
  sparse<5><<<148 * 8, 256>>>(a, n, out);
  sparse<6><<<148 * 8, 256>>>(a, n, out);
  sparse<7><<<148 * 8, 256>>>(a, n, out);
  unsigned long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  printf("gathers per kernel: %llu  %s\n", h[1] / 8, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
