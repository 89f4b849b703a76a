// Probe of tcgen05.mma shared-memory descriptor conventions on real hardware.
//   umma_probe nosw  <shift> <tmem_col> <swap_lbo_sbo> <N>
//   umma_probe sw128 <N>
// Prints "PASS"/"FAIL max_err" comparing D = A[shift:shift+128] * B^T with a CPU reference.
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
constexpr int RA = 136;

struct Args { int mode, shift, tcol, swap, N; };

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* __restrict__ Ag, const __nv_bfloat16* __restrict__ Bg,
                                             float* __restrict__ D, Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = a.N;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 32768;
  if (a.mode == 0) {
    // A: [kc][RA][8]   B: [kc][N][8]
    for (int i = tid; i < RA * K; i += 128) {
      int row = i / K, k = i % K;
      *reinterpret_cast<__nv_bfloat16*>(sA + ((k / 8) * RA + row) * 16 + (k % 8) * 2) = Ag[row * K + k];
    }
    for (int i = tid; i < N * K; i += 128) {
      int row = i / K, k = i % K;
      *reinterpret_cast<__nv_bfloat16*>(sB + ((k / 8) * N + row) * 16 + (k % 8) * 2) = Bg[row * K + k];
    }
  } else {
    // 128B-swizzled K-major: row = 128 B; 16-B chunk c stored at chunk (c ^ (row & 7))
    for (int i = tid; i < 128 * K; i += 128) {
      int row = i / K, k = i % K;
      int chunk = (k / 8) ^ (row & 7);
      *reinterpret_cast<__nv_bfloat16*>(sA + row * 128 + chunk * 16 + (k % 8) * 2) = Ag[row * K + k];
    }
    for (int i = tid; i < N * K; i += 128) {
      int row = i / K, k = i % K;
      int chunk = (k / 8) ^ (row & 7);
      *reinterpret_cast<__nv_bfloat16*>(sB + row * 128 + chunk * 16 + (k % 8) * 2) = Bg[row * K + k];
    }
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = tmem_base + a.tcol;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    for (int s = 0; s < K / 16; ++s) {
      uint64_t ad, bd;
      if (a.mode == 0) {
        uint32_t lboA = RA * 16, sbo = 128, lboB = N * 16;
        uint32_t aaddr = smem_u32(sA) + (2 * s) * RA * 16 + a.shift * 16;
        uint32_t baddr = smem_u32(sB) + (2 * s) * N * 16;
        ad = a.swap ? make_smem_desc(aaddr, sbo, lboA, 0) : make_smem_desc(aaddr, lboA, sbo, 0);
        bd = a.swap ? make_smem_desc(baddr, sbo, lboB, 0) : make_smem_desc(baddr, lboB, sbo, 0);
      } else {
        ad = make_smem_desc(smem_u32(sA) + s * 32, 16, 1024, 2);
        bd = make_smem_desc(smem_u32(sB) + s * 32, 16, 1024, 2);
      }
      mma_bf16_ss(tbase, ad, bd, idesc, s > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  for (int c = 0; c < N; c += 16) {
    uint32_t v[16];
    tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(warp * 32 + (tid & 31)) * N + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

int main(int argc, char** argv) {
  Args a{0, 0, 0, 0, 80};
  if (argc < 2) { printf("usage\n"); return 1; }
  if (!strcmp(argv[1], "nosw")) {
    a.mode = 0; a.shift = atoi(argv[2]); a.tcol = atoi(argv[3]); a.swap = atoi(argv[4]); a.N = atoi(argv[5]);
  } else { a.mode = 1; a.N = atoi(argv[2]); }
  const int N = a.N;
  std::vector<__nv_bfloat16> A(RA * K), B(N * K);
  std::vector<float> Af(RA * K), Bf(N * K);
  srand(1);
  for (int i = 0; i < RA * K; ++i) { float v = (rand() % 17 - 8) / 8.0f; A[i] = __float2bfloat16(v); Af[i] = v; }
  for (int i = 0; i < N * K; ++i) { float v = (rand() % 13 - 6) / 4.0f; B[i] = __float2bfloat16(v); Bf[i] = v; }
  __nv_bfloat16 *dA, *dB; float* dD;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, 128 * N * 4));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 65536));
  probe<<<1, 128, 32768 + 65536>>>(dA, dB, dD, a);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)Af[(m + (a.mode == 0 ? a.shift : 0)) * K + k] * Bf[n * K + k];
      maxerr = fmax(maxerr, fabs(s - D[m * N + n]));
    }
  printf("%s mode=%s shift=%d tcol=%d swap=%d N=%d max_err=%g\n", maxerr < 1e-3 ? "PASS" : "FAIL", a.mode ? "sw128" : "nosw",
         a.shift, a.tcol, a.swap, N, maxerr);
  return 0;
}
