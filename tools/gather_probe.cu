// tools/gather_probe.cu — microbenchmark: DRAM bytes fetched per random 4-byte gather on B200 as a
// function of cudaLimitMaxL2FetchGranularity and of the load flavour. Run under
//   ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./gather_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
template <int MODE>
__global__ void gather(const uint32_t *a, uint64_t n, uint64_t ngather, unsigned long long *out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (; i < ngather; i += stride) {
    uint64_t idx = mix(i) % n;
    const uint32_t *p = a + idx;
    uint32_t v;
    if (MODE == 0) v = __ldg(p);
    else if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (MODE == 2) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    acc += v;
  }
  if (acc == 0x12345678) atomicAdd(out, 1ull);
}
int main() {
  const uint64_t n = 1ull << 30;  // 4 GiB of uint32
  uint32_t *a; unsigned long long *out;
  cudaMalloc(&a, n * 4); cudaMalloc(&out, 8);
  cudaMemset(a, 1, n * 4);
  const uint64_t ng = 1ull << 24;
  int grans[3] = {32, 64, 128};
  for (int g = 0; g < 3; ++g) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, grans[g]);
    size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    printf("set %d -> %s, get %zu\n", grans[g], cudaGetErrorString(e), got);
    gather<0><<<148 * 8, 256>>>(a, n, ng, out);
    gather<1><<<148 * 8, 256>>>(a, n, ng, out);
    gather<2><<<148 * 8, 256>>>(a, n, ng, out);
    gather<3><<<148 * 8, 256>>>(a, n, ng, out);
    cudaDeviceSynchronize();
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
