// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N and operand layout.
//   umma_rate <mode: 0 nosw | 1 sw128> <N> <pattern: 0 same operands | 1 conv pattern (3 shifted taps, 3 tiles)>
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "../modulationdetectioncnn_b200/csrc/sm100.cuh"
using namespace sm100;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s: %s\n", #x, cudaGetErrorString(e)); return 2; } } while (0)

constexpr int REP = 4096;

__global__ void __launch_bounds__(128) rate(int mode, int N, int pattern, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = uniform_warp_idx();
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tmem_base;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem) + 131072;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      const uint32_t hiA = mode ? smem_desc_hi(1024, 2) : smem_desc_hi(128, 0);
      const uint32_t a_lo = mode ? smem_desc_lo(a_base, 16) : smem_desc_lo(a_base, 384 * 16);
      const uint32_t b_lo = mode ? smem_desc_lo(b_base, 16) : smem_desc_lo(b_base, N * 16);
      t0 = clock64();
      for (int r = 0; r < REP / 18; ++r) {
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              uint32_t ao = 0, bo = 0;
              if (pattern) {
                if (mode == 0) { ao = ((2 * ks) * 384 * 16 + (128 * t + j) * 16) >> 4; bo = ((j * 4 + 2 * ks) * N * 16) >> 4; }
                else { ao = (t * 16384 + ks * 32) >> 4; bo = (j * 32768 / 4 + ks * 32) >> 4; }
              }
              mma_bf16_ss(tb + (pattern ? t * N % 256 : 0), desc64(a_lo + ao, hiA), desc64(b_lo + bo, hiA), idesc, 1);
            }
      }
      mma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (elect_one()) {
      t1 = clock64();
      out[blockIdx.x] = t1 - t0;
    }
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tb);
}

int main(int argc, char** argv) {
  const int mode = atoi(argv[1]), N = atoi(argv[2]), pattern = atoi(argv[3]);
  const int grid = argc > 4 ? atoi(argv[4]) : 1;
  long long* d;
  CK(cudaMalloc(&d, grid * 8));
  CK(cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int it = 0; it < 2; ++it) {
    rate<<<grid, 128, 200 * 1024>>>(mode, N, pattern, d);
    CK(cudaDeviceSynchronize());
  }
  long long h[148];
  CK(cudaMemcpy(h, d, (grid < 148 ? grid : 148) * 8, cudaMemcpyDeviceToHost));
  const int issued = REP / 18 * 18;
  printf("mode=%s N=%3d pattern=%d grid=%3d : %.1f cycles/MMA (floor %d)\n", mode ? "sw128" : "nosw ", N, pattern, grid,
         (double)h[0] / issued, N / 2);
  return 0;
}
