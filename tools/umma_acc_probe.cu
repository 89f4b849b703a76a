// How does tcgen05.mma kind::tf32 round when it adds a product into the fp32 TMEM accumulator?
// Row m of A carries BIG[m] in the first MMA (accumulate = 0) and DELTA[m] in each of the L following
// MMAs (accumulate = 1); B[n][0] = 1, so D[m][*] = BIG[m] + L x DELTA[m] under exact arithmetic.
// DELTA is a fraction of one ulp of BIG, so the printed result shows the rounding rule of every step:
//   RN: 0.75 ulp steps each round up, 0.25 ulp steps are lost;   RZ: every sub-ulp step toward zero is lost ...
//   umma_acc_probe [L]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../modulationdetectioncnn_b200/csrc/sm100.cuh"
using namespace sm100;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 2; } } while (0)
constexpr int N = 16, R = 128;

__global__ void __launch_bounds__(128) probe(const float* __restrict__ A0g, const float* __restrict__ A1g, float* __restrict__ D, int L) {
  __shared__ __align__(1024) uint8_t sA0[2 * R * 16], sA1[2 * R * 16], sB[2 * N * 16];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = uniform_warp_idx();
  for (int i = tid; i < R * 8; i += 128) {
    const int row = i / 8, k = i % 8;
    *reinterpret_cast<float*>(sA0 + ((k / 4) * R + row) * 16 + (k % 4) * 4) = A0g[row * 8 + k];
    *reinterpret_cast<float*>(sA1 + ((k / 4) * R + row) * 16 + (k % 4) * 4) = A1g[row * 8 + k];
  }
  for (int i = tid; i < N * 8; i += 128) {
    const int row = i / 8, k = i % 8;
    *reinterpret_cast<float*>(sB + ((k / 4) * N + row) * 16 + (k % 4) * 4) = 1.0f;   // every k contributes A[m][k]
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<32>(&tmem_base);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tmem_base;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(128, N);
      const uint32_t hi = smem_desc_hi(128, 0);
      const uint64_t a0 = desc64(smem_desc_lo(smem_u32(sA0), R * 16), hi);
      const uint64_t a1 = desc64(smem_desc_lo(smem_u32(sA1), R * 16), hi);
      const uint64_t bd = desc64(smem_desc_lo(smem_u32(sB), N * 16), hi);
      mma_tf32_ss(tb, a0, bd, idesc, 0);
      for (int s = 0; s < L; ++s) mma_tf32_ss(tb, a1, bd, idesc, 1);
      mma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  uint32_t v[16];
  tmem_ld16(tb + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  D[warp * 32 + (tid & 31)] = __uint_as_float(v[3]);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<32>(tb);
}

int main(int argc, char** argv) {
  const int L = argc > 1 ? atoi(argv[1]) : 16;
  std::vector<float> A0(R * 8, 0.f), A1(R * 8, 0.f);
  const float ulp = ldexpf(1.f, -23);             // ulp of 1.0
  struct Case { float big, frac; int spread; const char* what; };
  const Case cases[] = {
      {1.f, 0.25f, 0, "+1 + 0.25ulp"}, {1.f, 0.5f, 0, "+1 + 0.50ulp"}, {1.f, 0.75f, 0, "+1 + 0.75ulp"},
      {1.f, -0.25f, 0, "+1 - 0.25ulp"}, {1.f, -0.5f, 0, "+1 - 0.50ulp"}, {1.f, -0.75f, 0, "+1 - 0.75ulp"},
      {-1.f, 0.25f, 0, "-1 + 0.25ulp"}, {-1.f, 0.5f, 0, "-1 + 0.50ulp"}, {-1.f, 0.75f, 0, "-1 + 0.75ulp"},
      {-1.f, -0.25f, 0, "-1 - 0.25ulp"}, {-1.f, -0.5f, 0, "-1 - 0.50ulp"}, {-1.f, -0.75f, 0, "-1 - 0.75ulp"},
      {1.f, 1.25f, 0, "+1 + 1.25ulp"}, {1.f, 1.75f, 0, "+1 + 1.75ulp"},
      {1.f, 0.75f, 1, "+1 + 8 x (0.75/8)ulp in one MMA"}, {1.f, 0.25f, 1, "+1 + 8 x (0.25/8)ulp in one MMA"},
      {1.f, 0.0625f, 0, "+1 + 1/16 ulp"}, {1.f, 0.9375f, 0, "+1 + 15/16 ulp"},
      {1.5f, 0.75f, 0, "+1.5 + 0.75ulp"}, {1.5f, -0.75f, 0, "+1.5 - 0.75ulp"},
  };
  const int nc = sizeof(cases) / sizeof(cases[0]);
  for (int m = 0; m < nc; ++m) {
    A0[m * 8] = cases[m].big;
    if (cases[m].spread) for (int k = 0; k < 8; ++k) A1[m * 8 + k] = cases[m].frac * ulp / 8;
    else A1[m * 8] = cases[m].frac * ulp;
  }
  float *dA0, *dA1, *dD;
  CK(cudaMalloc(&dA0, R * 32)); CK(cudaMalloc(&dA1, R * 32)); CK(cudaMalloc(&dD, R * 4));
  CK(cudaMemcpy(dA0, A0.data(), R * 32, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dA1, A1.data(), R * 32, cudaMemcpyHostToDevice));
  probe<<<1, 128>>>(dA0, dA1, dD, L);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D(R);
  CK(cudaMemcpy(D.data(), dD, R * 4, cudaMemcpyDeviceToHost));
  printf("L = %d accumulating MMAs; result and exact value in ulps of 1.0 away from BIG\n", L);
  for (int m = 0; m < nc; ++m)
    printf("  %-34s got %+9.3f ulp   exact %+9.3f ulp\n", cases[m].what, (double)(D[m] - cases[m].big) / ulp,
           (double)L * cases[m].frac);
  return 0;
}
