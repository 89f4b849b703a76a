// Probe of tcgen05.mma.cta_group::2 (CTA pair, M = 256) with the no-swizzle K-major layout:
// correctness of the operand split (which half of B each CTA provides) and cycles per MMA.
//   umma2_probe check <N> <shift>     D = A[shift : shift+256] * B^T against the CPU
//   umma2_probe rate  <N>             cycles per MMA
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "../modulationdetectioncnn_b200/csrc/sm100.cuh"
using namespace sm100;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 2; } } while (0)

constexpr int K = 64;
constexpr int RA = 136;     // A rows staged per CTA (128 + halo)
constexpr int REP = 4096;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
probe2(const __nv_bfloat16* __restrict__ Ag, const __nv_bfloat16* __restrict__ Bg, float* __restrict__ D, int N, int shift,
       int rate, long long* cyc, int commit_every) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t bar2;       // absorbs the extra commits of the commit-cost experiment
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = uniform_warp_idx();
  const uint32_t rank = cluster_rank();
  const int NH = N / 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 32768;
  // A: this CTA's rows [128 rank, 128 rank + RA) as [kc][RA][8];  B: rows [NH rank, NH rank + NH) as [kc][NH][8]
  for (int i = tid; i < RA * K; i += 128) {
    int row = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sA + ((k / 8) * RA + row) * 16 + (k % 8) * 2) = Ag[(128 * rank + row) * K + k];
  }
  for (int i = tid; i < NH * K; i += 128) {
    int row = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sB + ((k / 8) * NH + row) * 16 + (k % 8) * 2) = Bg[(NH * rank + row) * K + k];
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1 << 19); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tb = tmem_base;
  long long t0 = 0;
  if (rank == 0 && warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(256, N);
      const uint32_t hi = smem_desc_hi(128, 0);
      const uint32_t a_lo = smem_desc_lo(smem_u32(sA) + shift * 16, RA * 16);
      const uint32_t b_lo = smem_desc_lo(smem_u32(sB), NH * 16);
      t0 = clock64();
      const int reps = rate ? REP / 4 : 1;
      int since = 0;
      for (int r = 0; r < reps; ++r) {
        if (commit_every > 0 && (since += K / 16) >= commit_every) {
          since = 0;
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                           smem_u32(&bar2)),
                       "h"((uint16_t)3)
                       : "memory");
        }
#pragma unroll
        for (int s = 0; s < K / 16; ++s) {
          const uint64_t ad = desc64(a_lo + ((2 * s * RA * 16) >> 4), hi), bd = desc64(b_lo + ((2 * s * NH * 16) >> 4), hi);
          uint32_t acc = (r | s) != 0;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb), "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
              : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                       smem_u32(&bar)),
                   "h"((uint16_t)3)
                   : "memory");
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  if (rank == 0 && warp == 1 && (tid & 31) == 0 && rate) cyc[blockIdx.x / 2] = clock64() - t0;
  if (!rate) {
    for (int c = 0; c < N; c += 16) {
      uint32_t v[16];
      tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c, v);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) D[(128 * rank + warp * 32 + (tid & 31)) * N + c + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "n"(512) : "memory");
}

int main(int argc, char** argv) {
  if (argc < 3) { printf("usage\n"); return 1; }
  const bool rate = !strcmp(argv[1], "rate");
  const int N = atoi(argv[2]);
  const int shift = (!rate && argc > 3) ? atoi(argv[3]) : 0;
  const int grid = (rate && argc > 3) ? atoi(argv[3]) : 2;
  const int commit_every = (rate && argc > 4) ? atoi(argv[4]) : 0;   // extra tcgen05.commit after every this many MMAs
  const int RT = 128 + RA;   // A rows in global
  std::vector<__nv_bfloat16> A(RT * K), B(N * K);
  std::vector<float> Af(RT * K), Bf(N * K);
  srand(1);
  for (int i = 0; i < RT * K; ++i) { float v = (rand() % 17 - 8) / 8.0f; A[i] = __float2bfloat16(v); Af[i] = v; }
  for (int i = 0; i < N * K; ++i) { float v = (rand() % 13 - 6) / 4.0f; B[i] = __float2bfloat16(v); Bf[i] = v; }
  __nv_bfloat16 *dA, *dB; float* dD; long long* dC;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, 256 * N * 4)); CK(cudaMalloc(&dC, 8 * 256));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, 256 * N * 4));
  CK(cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  for (int it = 0; it < (rate ? 2 : 1); ++it) {
    probe2<<<grid, 128, 65536>>>(dA, dB, dD, N, shift, rate, dC, commit_every);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
  }
  if (rate) {
    long long c;
    CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
    printf("cta_group::2 M=256 N=%d grid=%d commit_every=%d: %.1f cycles/MMA (1-CTA equivalent work: 2 x (128 x %d))\n", N, grid,
           commit_every, (double)c / REP, N);
    return 0;
  }
  std::vector<float> D(256 * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)Af[(m + shift) * K + k] * Bf[n * K + k];
      maxerr = fmax(maxerr, fabs(s - D[m * N + n]));
    }
  printf("%s cta_group::2 N=%d shift=%d max_err=%g\n", maxerr < 1e-3 ? "PASS" : "FAIL", N, shift, maxerr);
  return 0;
}
