// Probe of tcgen05.mma kind::tf32 in CTA-pair mode (M = 256, K = 8) with the no-swizzle K-major layout
// (4 fp32 per 16-B group), including the 16-B row shift used by the implicit-GEMM conv.
//   umma_tf32_probe <N> <shift>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../modulationdetectioncnn_b200/csrc/sm100.cuh"
using namespace sm100;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 2; } } while (0)
constexpr int K = 32, RA = 136;

__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
probe(const float* __restrict__ Ag, const float* __restrict__ Bg, float* __restrict__ D, int N, int shift) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = uniform_warp_idx();
  const uint32_t rank = cluster_ctarank();
  const int NH = N / 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 32768;
  for (int i = tid; i < RA * K; i += 128) {
    int row = i / K, k = i % K;
    *reinterpret_cast<float*>(sA + ((k / 4) * RA + row) * 16 + (k % 4) * 4) = Ag[(128 * rank + row) * K + k];
  }
  for (int i = tid; i < NH * K; i += 128) {
    int row = i / K, k = i % K;
    *reinterpret_cast<float*>(sB + ((k / 4) * NH + row) * 16 + (k % 4) * 4) = Bg[(NH * rank + row) * K + k];
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc_pair<512>(&tmem_base);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tb = tmem_base;
  if (rank == 0 && warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(256, N);
      const uint32_t hi = smem_desc_hi(128, 0);
      const uint32_t a_lo = smem_desc_lo(smem_u32(sA) + shift * 16, RA * 16);
      const uint32_t b_lo = smem_desc_lo(smem_u32(sB), NH * 16);
#pragma unroll
      for (int s = 0; s < K / 8; ++s) {
        const uint64_t ad = desc64(a_lo + ((2 * s * RA * 16) >> 4), hi), bd = desc64(b_lo + ((2 * s * NH * 16) >> 4), hi);
        uint32_t acc = s != 0;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb), "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
            : "memory");
      }
      mma_commit_pair(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  for (int c = 0; c < N; c += 16) {
    uint32_t v[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(128 * rank + warp * 32 + (tid & 31)) * N + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair<512>(tb);
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 80, shift = argc > 2 ? atoi(argv[2]) : 0;
  const int RT = 128 + RA;
  std::vector<float> A(RT * K), B(N * K);
  srand(1);
  for (auto& v : A) v = (rand() % 17 - 8) / 8.0f;
  for (auto& v : B) v = (rand() % 13 - 6) / 4.0f;
  // one value that needs more than 10 mantissa bits: the MMA must TRUNCATE it (tf32), not round the fp32 product
  A[5] = 1.0f + 1.0f / 4096.0f;
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, 256 * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, 256 * N * 4));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  probe<<<2, 128, 65536>>>(dA, dB, dD, N, shift);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D(256 * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, row0err = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[(m + shift) * K + k] * B[n * K + k];
      const double e = fabs(s - D[m * N + n]);
      if (m + shift == 0) row0err = fmax(row0err, e); else maxerr = fmax(maxerr, e);
    }
  printf("%s tf32 cta_group::2 N=%d shift=%d max_err(exact rows)=%g  err(row with 13-bit mantissa)=%g\n",
         maxerr < 1e-4 ? "PASS" : "FAIL", N, shift, maxerr, row0err);
  return 0;
}
