// Microbenchmark: issue rate of the CUDA-core instructions the hot paths lean on (per SM sub-partition).
// Each warp runs ILP independent dependency chains of one instruction; prints cycles per warp-instruction
// per SMSP for 1, 2, 4 and 8 warps per SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s: %s\n", #x, cudaGetErrorString(e)); return 2; } } while (0)
constexpr int ITER = 2048, ILP = 8;

template <int OP>
__global__ void k(unsigned long long* out, long long* cyc, float seed) {
  uint64_t a[ILP];
  float f[ILP];
  int i32[ILP];
  long long w64[ILP];
  uint32_t u[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { f[i] = seed + i + threadIdx.x; a[i] = (uint64_t)__float_as_uint(f[i]) << 32 | __float_as_uint(f[i] * 0.5f); i32[i] = (int)f[i]; w64[i] = i32[i]; u[i] = i32[i]; }
  const uint64_t b = (uint64_t)__float_as_uint(seed * 0.999f) << 32 | __float_as_uint(seed * 1.001f);
  uint64_t xx[ILP], yy[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { xx[i] = a[i] ^ 0x0000100000001000ull * (i + 1); yy[i] = a[i] ^ 0x0000020000000200ull * (i + 3); }
  const float fb = seed * 0.999f;
  const int ib = (int)(seed * 77.f) | 1;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a[i]) : "l"(b));
      if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(fb));
      if (OP == 2) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w64[i]) : "r"(i32[i]), "r"(ib));
      if (OP == 3) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(i32[i]) : "r"(ib));
      if (OP == 4) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fb));
      if (OP == 5) asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(f[i]), "f"(__uint_as_float(u[i])));
      if (OP == 6) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(u[i]) : "r"(ib));
      if (OP == 7) asm volatile("lop3.b32 %0, %0, %1, 0x20000, 0xE4;" : "+r"(u[i]) : "r"(ib));
      if (OP == 8) asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, %0;" : "+f"(f[i]));
      if (OP == 9) asm volatile("add.s32 %0, %0, %1;" : "+r"(i32[i]) : "r"(ib));
      if (OP == 10) asm volatile("fma.rm.f32x2 %0, %0, %1, %1;" : "+l"(a[i]) : "l"(b));
      if (OP == 11) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(b));
      if (OP == 12) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(b));
      if (OP == 13) asm volatile("fma.rm.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(fb));
      if (OP == 20) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(xx[i]), "l"(yy[i]));
      if (OP == 21) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(xx[0]), "l"(yy[i]));
      if (OP == 22) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(xx[i & ~3]), "l"(yy[i]));
      if (OP == 23) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[i]) : "f"(__uint_as_float((unsigned)xx[i])), "f"(__uint_as_float((unsigned)yy[i])));
      if (OP == 15) { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a[i]) : "l"(b)); asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fb)); }
      if (OP == 16) { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a[i]) : "l"(b)); asm volatile("add.s32 %0, %0, %1;" : "+r"(i32[i]) : "r"(ib)); }
      if (OP == 17) { asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(i32[i]) : "r"(ib)); asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fb)); }
      if (OP == 18) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(fb)); asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(u[i]) : "r"(ib)); }
      if (OP == 19) { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a[i]) : "l"(b)); asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fb)); asm volatile("add.s32 %0, %0, %1;" : "+r"(i32[i]) : "r"(ib)); }
      if (OP == 14) asm volatile("{.reg .f32 lo, hi; mov.b64 {lo, hi}, %0; max.f32 lo, lo, 0f00000000; max.f32 hi, hi, 0f00000000; mov.b64 %0, {lo, hi}; fma.rn.f32x2 %0, %0, %1, %1;}" : "+l"(a[i]) : "l"(b));
    }
  }
  const long long t1 = clock64();
  unsigned long long acc = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc += a[i] + __float_as_uint(f[i]) + i32[i] + w64[i] + u[i] + xx[i] + yy[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
int run(const char* name) {
  unsigned long long* out; long long* cyc;
  CK(cudaMalloc(&out, 148 * 1024 * 8)); CK(cudaMalloc(&cyc, 148 * 8));
  printf("%-28s", name);
  for (int wps : {1, 2, 4, 8}) {
    const int threads = wps * 4 * 32;
    k<OP><<<148, threads>>>(out, cyc, 1.0f); CK(cudaDeviceSynchronize());
    k<OP><<<148, threads>>>(out, cyc, 1.0f); CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
    printf("  %dw/SMSP: %.2f", wps, (double)c / ((double)ITER * ILP * wps));
  }
  printf("   cycles per warp-instruction per SMSP\n");
  cudaFree(out); cudaFree(cyc);
  return 0;
}
int main() {
  run<0>("FFMA2 (fma.rn.f32x2)");
  run<1>("FFMA  (3-register)");
  run<8>("FFMA  (immediate operand)");
  run<2>("IMAD.WIDE");
  run<3>("IMAD");
  run<4>("FMNMX");
  run<5>("F2FP.RELU.BF16.PACK_AB");
  run<6>("SHF (funnel)");
  run<7>("LOP3");
  run<9>("IADD");
  run<10>("FFMA2.RM");
  run<11>("FADD2");
  run<12>("FMUL2");
  run<13>("FFMA.RM");
  run<14>("2 FMNMX + FFMA2 (per 3 instr)");
  run<20>("FFMA2 d=a*b+d, 3 distinct register pairs");
  run<21>("FFMA2 d=X*b+d, X shared by all (reuse)");
  run<22>("FFMA2 d=X*b+d, X shared by 4 in a row");
  run<23>("FFMA d=a*b+d, 3 distinct registers");
  run<15>("FFMA2 + FMNMX independent (pair)");
  run<16>("FFMA2 + IADD independent (pair)");
  run<17>("IMAD + FMNMX independent (pair)");
  run<18>("FFMA + SHF independent (pair)");
  run<19>("FFMA2 + FMNMX + IADD (triple)");
  return 0;
}
