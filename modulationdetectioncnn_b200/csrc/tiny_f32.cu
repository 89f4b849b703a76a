// TinyCNN2(F,C) fp32 forward: the nets the reference actually ships.
//
//   Reshape(2,128,1) -> ZeroPadding2D((0,0),(1,1)) -> Conv2D(F,(1,2),relu) -> Flatten
//   -> Dense(C, relu) -> softmax            (/root/reference/CNN.ipynb:1 cell 6; h5 model_config)
//
//   y[r][p][f] = relu(xp[r][p]*k0[f] + xp[r][p+1]*k1[f] + b[f]),  p = 0..128, xp[0]=xp[129]=0
//   z[c]       = relu(sum_{r,p,f} y[r][p][f] * D[(r*129 + p)*F + f][c] + d[c])
//
// 1,036 B per frame.  A warp maps lane l to positions p = 4l..4l+3 of both rows; position 128 (which only needs
// x[127]) is spread over lanes 0..2F-1, one (row,filter) pair each.  The math is packed FFMA2 (two positions per
// instruction), the Dense rows a lane needs live in shared memory, and the epilogue - cross-lane sums by recursive
// halving, softmax, argmax, class histogram - is fused.
#include "mdc_internal.cuh"
#include "sm100.cuh"

namespace mdc {
using namespace sm100;

struct TinyParams {
  float conv[3 * kMaxFilters];   // k0,k1,b per filter
  float bias[kMaxClasses];
  int F, C;
};

__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_relu(uint64_t v) {
  float lo, hi;
  f2_unpack(v, lo, hi);
  return f2_pack(fmaxf(lo, 0.f), fmaxf(hi, 0.f));
}

// dense image (packed on host):
//   main [r][f][c][128]  entry p = D[(r*129 + p)*F + f][c]            -> float4 per lane
//   tail [r][f][c]       = D[(r*129 + 128)*F + f][c]                  (position 128)
//
// FOUR frames per warp pass.  (The first formulation - two frames per pass, frames through a CTA-wide TMA ring or
// register prefetch, conv constants packed with MOVs - measured 483 warp-instructions and 131 shared-memory wavefronts
// per frame for F = 10, profiles/r01_ncu_small.md; it was removed when this one replaced it.)  Every lane needs its own
// 2 F C float4 of Dense weights, so a warp pass streams the WHOLE 30 KB weight image through the LSU: with four frames
// per pass that is 60 wavefronts per frame.  The conv constants arrive pre-duplicated as 64-bit kernel parameters (FFMA2 reads them straight from the constant bank),
// and the frames come through a per-warp ring of three 4 KB buffers filled by 1-D bulk copies (one elected lane, one
// mbarrier per buffer, no register prefetch, no coupling between warps: a warp refills the buffer it has just read).
struct TinyParams4 {
  unsigned long long k0[kMaxFilters], k1[kMaxFilters], b[kMaxFilters];   // {v, v} pairs
  float bias[kMaxClasses];
  float zero;                                                            // 0.0f (see the shifted pairs in the kernel)
};

// FMT: frame format in global memory and in the ring (MDC_IN_F32 1,024 B, MDC_IN_U8IQ 256 B, MDC_IN_I16 512 B per
// frame); raw formats are converted when a lane picks its samples out of the ring - exact, so the results are
// bit-identical to mdc_sdr_ingest_u8 (or s / 4096) followed by the f32 call.
template <int F, int C, int W, int B, int FMT = MDC_IN_F32>
struct Tiny4 {
  static constexpr int kWarps = W, kThreads = kWarps * 32, R = 4, kBufs = B;
  static constexpr int kFB = FMT == MDC_IN_U8IQ ? 256 : (FMT == MDC_IN_I16 ? 512 : 1024);   // bytes per frame
  static constexpr int kWBytes = 2 * F * C * 128 * 4;
  static constexpr int ring = kWBytes;
  static constexpr int bars = ring + kWarps * kBufs * R * kFB;
  static constexpr int total = bars + kWarps * kBufs * 8;
  static_assert(C <= 4 && 2 * F <= 32, "shape outside the specialisation");
};

// the lane's samples of one staged frame: positions 4 lane .. 4 lane + 3 of both rows, and the sample left of them
template <int FMT>
__device__ __forceinline__ void tiny_lane_samples(const uint8_t* fr, int lane, float4& xi, float4& xq, float& pi, float& pq) {
  if constexpr (FMT == MDC_IN_F32) {
    const float* f = reinterpret_cast<const float*>(fr);
    xi = reinterpret_cast<const float4*>(f)[lane];
    xq = reinterpret_cast<const float4*>(f + 128)[lane];
    pi = lane ? f[4 * lane - 1] : 0.f;
    pq = lane ? f[128 + 4 * lane - 1] : 0.f;
  } else if constexpr (FMT == MDC_IN_U8IQ) {
    // I0 Q0 I1 Q1 ...: value (u - 127.5) / 128 = (2u - 255) / 256, exact (== sdr_ingest_kernel)
    auto cv = [](unsigned u) { return (float)(2 * (int)u - 255) * (1.f / 256.f); };
    const uint2 b = reinterpret_cast<const uint2*>(fr)[lane];
    xi = make_float4(cv(b.x & 255u), cv((b.x >> 16) & 255u), cv(b.y & 255u), cv((b.y >> 16) & 255u));
    xq = make_float4(cv((b.x >> 8) & 255u), cv(b.x >> 24), cv((b.y >> 8) & 255u), cv(b.y >> 24));
    const unsigned pb = lane ? reinterpret_cast<const uint16_t*>(fr)[4 * lane - 1] : 0u;
    pi = lane ? cv(pb & 255u) : 0.f;
    pq = lane ? cv(pb >> 8) : 0.f;
  } else {
    // int16 [2][128] Q6.12 in the test_table address map: value s / 4096, exact
    const int16_t* h = reinterpret_cast<const int16_t*>(fr);
    auto cv = [](int v) { return (float)v * (1.f / 4096.f); };
    const uint2 a = reinterpret_cast<const uint2*>(h)[lane], b = reinterpret_cast<const uint2*>(h + 128)[lane];
    xi = make_float4(cv((short)(a.x & 0xFFFFu)), cv((int)a.x >> 16), cv((short)(a.y & 0xFFFFu)), cv((int)a.y >> 16));
    xq = make_float4(cv((short)(b.x & 0xFFFFu)), cv((int)b.x >> 16), cv((short)(b.y & 0xFFFFu)), cv((int)b.y >> 16));
    pi = lane ? cv(h[4 * lane - 1]) : 0.f;
    pq = lane ? cv(h[128 + 4 * lane - 1]) : 0.f;
  }
}
// sample 127 of row r (position 128 of the padded row only needs that one)
template <int FMT>
__device__ __forceinline__ float tiny_last_sample(const uint8_t* fr, int r) {
  if constexpr (FMT == MDC_IN_F32) return reinterpret_cast<const float*>(fr)[r * 128 + 127];
  else if constexpr (FMT == MDC_IN_U8IQ) return (float)(2 * (int)fr[254 + r] - 255) * (1.f / 256.f);
  else return (float)reinterpret_cast<const int16_t*>(fr)[r * 128 + 127] * (1.f / 4096.f);
}

template <int F, int C, int W, int B, int FMT>
__global__ void __launch_bounds__(Tiny4<F, C, W, B, FMT>::kThreads, 1)
tiny_f32_kernel4(const __grid_constant__ TinyParams4 p, const float4* __restrict__ dmain, const float* __restrict__ dtail,
                 const uint8_t* __restrict__ x, long long n, float* __restrict__ probs, float* __restrict__ dense,
                 int* __restrict__ cls, unsigned long long* __restrict__ hist) {
  using Cfg = Tiny4<F, C, W, B, FMT>;
  constexpr int R = Cfg::R, kBufs = Cfg::kBufs, kFB = Cfg::kFB;
  extern __shared__ __align__(128) uint8_t smem[];
  // let the next launch on the stream start its own prologue (weight image to shared memory, barriers) on SMs as this
  // grid's CTAs retire: back-to-back launches of 65,536 frames last 20-40 us, a 5 us gap + prologue is 15-25 % of that
  asm volatile("griddepcontrol.launch_dependents;");
  const float4* wsm = reinterpret_cast<const float4*>(smem);
  const int lane = threadIdx.x & 31, warp = uniform_warp_idx();
  uint8_t* myring = smem + Cfg::ring + warp * (kBufs * R * kFB);
  uint64_t* mybar = reinterpret_cast<uint64_t*>(smem + Cfg::bars) + warp * kBufs;

  {
    float4* wdst = reinterpret_cast<float4*>(smem);
    for (int i = threadIdx.x; i < 2 * F * C * 32; i += Cfg::kThreads) wdst[i] = __ldg(dmain + i);
  }
  if (lane == 0) {
    for (int b = 0; b < kBufs; ++b) mbar_init(&mybar[b], 1);
    fence_barrier_init();
  }
  __syncthreads();

  // position 128: lane j < 2F handles (r = j / F, f = j % F)
  float tw[C];
  float tk0 = 0.f, tb = 0.f;
  const int tr = lane / F;
#pragma unroll
  for (int c = 0; c < C; ++c) tw[c] = 0.f;
  if (lane < 2 * F) {
    const int tf = lane % F;
    tk0 = __uint_as_float((unsigned)(p.k0[tf] & 0xFFFFFFFFull));
    tb = __uint_as_float((unsigned)(p.b[tf] & 0xFFFFFFFFull));
#pragma unroll
    for (int c = 0; c < C; ++c) tw[c] = __ldg(dtail + (tr * F + tf) * C + c);
  }

  // groups of R consecutive frames, dealt round-robin to the warps of the grid
  const long long ngroups = (n + R - 1) / R;
  const long long stride = (long long)gridDim.x * Cfg::kWarps;
  long long g = (long long)blockIdx.x * Cfg::kWarps + warp;
  auto issue = [&](long long grp, int b) {        // this warp's frames of group grp -> buffer b
    if (grp < ngroups && elect_one()) {
      const long long f0 = grp * R, left = n - f0;
      const uint32_t bytes = (uint32_t)(left < R ? left : R) * (uint32_t)kFB;
      mbar_arrive_expect_tx(&mybar[b], bytes);
      bulk_g2s(myring + b * (R * kFB), x + f0 * kFB, bytes, &mybar[b]);
    }
    __syncwarp();
  };
  // Programmatic dependent launch: everything above touched only the handle's weight images, which no kernel on the
  // stream writes; from here on the kernel reads frames and writes outputs that the preceding kernel may own.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // kBufs - 1 passes in flight
  issue(g, 0);
  if (kBufs > 2) issue(g + stride, 1);
  unsigned cnt = 0;
  for (uint32_t it = 0; g < ngroups; g += stride, ++it) {
    const uint32_t b = it % kBufs;
    // the buffer read in the previous pass is free: every value loaded from it has been consumed by then
    issue(g + (kBufs - 1) * stride, (it + kBufs - 1) % kBufs);
    mbar_wait(&mybar[b], (it / kBufs) & 1);
    const uint8_t* fr = myring + b * (R * kFB);
    const long long f0 = g * R;

    uint64_t acc[R][C];       // {even positions, odd positions} partial sums
    uint64_t PI01[R], PI23[R], XI01[R], XI23[R], PQ01[R], PQ23[R], XQ01[R], XQ23[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float4 xi, xq;
      float pi, pq;
      tiny_lane_samples<FMT>(fr + r * kFB, lane, xi, xq, pi, pq);
      // y[i] = relu(x[i-1] k0 + x[i] k1 + b): pairs (y0,y1) and (y2,y3)
      // The shifted pairs (x[4l-1], x[4l]) and (x[4l+1], x[4l+2]) straddle the register pairs the 16-B loads fill.
      // Built with plain moves, ptxas re-copies them next to every use (170 MOVs per pass for F = 10); an add of a
      // zero it cannot see through (a kernel parameter) gives each shifted pair registers of its own, once per pass.
      PI01[r] = f2_pack(pi, xi.x + p.zero);           PI23[r] = f2_pack(xi.y + p.zero, xi.z + p.zero);
      XI01[r] = f2_pack(xi.x, xi.y);                  XI23[r] = f2_pack(xi.z, xi.w);
      PQ01[r] = f2_pack(pq, xq.x + p.zero);           PQ23[r] = f2_pack(xq.y + p.zero, xq.z + p.zero);
      XQ01[r] = f2_pack(xq.x, xq.y);                  XQ23[r] = f2_pack(xq.z, xq.w);
#pragma unroll
      for (int c = 0; c < C; ++c) acc[r][c] = 0ull;
    }
#pragma unroll
    for (int k = 0; k < F; ++k) {
      float4 wi[C], wq[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        wi[c] = wsm[((0 * F + k) * C + c) * 32 + lane];
        wq[c] = wsm[((1 * F + k) * C + c) * 32 + lane];
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const uint64_t yi01 = f2_relu(f2_fma(p.k0[k], PI01[r], f2_fma(p.k1[k], XI01[r], p.b[k])));
        const uint64_t yi23 = f2_relu(f2_fma(p.k0[k], PI23[r], f2_fma(p.k1[k], XI23[r], p.b[k])));
        const uint64_t yq01 = f2_relu(f2_fma(p.k0[k], PQ01[r], f2_fma(p.k1[k], XQ01[r], p.b[k])));
        const uint64_t yq23 = f2_relu(f2_fma(p.k0[k], PQ23[r], f2_fma(p.k1[k], XQ23[r], p.b[k])));
#pragma unroll
        for (int c = 0; c < C; ++c) {
          uint64_t a = acc[r][c];
          a = f2_fma(yi01, f2_pack(wi[c].x, wi[c].y), a);
          a = f2_fma(yi23, f2_pack(wi[c].z, wi[c].w), a);
          a = f2_fma(yq01, f2_pack(wq[c].x, wq[c].y), a);
          a = f2_fma(yq23, f2_pack(wq[c].z, wq[c].w), a);
          acc[r][c] = a;
        }
      }
    }
    // Cross-lane sums of the 4 x 4 (frame, class) partials by recursive halving: every step halves the values a lane
    // carries (8 + 4 + 2 + 1 + 1 shuffles instead of 5 per value) and leaves the sum for (frame r, class c) in the
    // lanes with bits 4..3 = r, bits 2..1 = c.
    float v[16];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < C) {
          float lo, hi;
          f2_unpack(acc[r][c], lo, hi);
          v[r * 4 + c] = lo + hi;
        } else {
          v[r * 4 + c] = 0.f;
        }
      }
    // position 128: xp[128] = x[127], xp[129] = 0 (the frames are still in this pass's buffer)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float yt = fmaxf(fmaf(tiny_last_sample<FMT>(fr + r * kFB, lane < 2 * F ? tr : 0), tk0, tb), 0.f);
#pragma unroll
      for (int c = 0; c < C; ++c) v[r * 4 + c] = fmaf(yt, tw[c], v[r * 4 + c]);   // tw = 0 in lanes without a (row, filter)
    }
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
    float w8[8], w4[4], w2[2];
#pragma unroll
    for (int j = 0; j < 8; ++j) w8[j] = (b4 ? v[8 + j] : v[j]) + __shfl_xor_sync(0xffffffffu, b4 ? v[j] : v[8 + j], 16);
#pragma unroll
    for (int j = 0; j < 4; ++j) w4[j] = (b3 ? w8[4 + j] : w8[j]) + __shfl_xor_sync(0xffffffffu, b3 ? w8[j] : w8[4 + j], 8);
#pragma unroll
    for (int j = 0; j < 2; ++j) w2[j] = (b2 ? w4[2 + j] : w4[j]) + __shfl_xor_sync(0xffffffffu, b2 ? w4[j] : w4[2 + j], 4);
    float t = (b1 ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, b1 ? w2[0] : w2[1], 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    const int myc = (lane >> 1) & 3, grp8 = lane & 24;
    float bsel = p.bias[0];
#pragma unroll
    for (int c = 1; c < C; ++c) if (myc == c) bsel = p.bias[c];
    t = fmaxf(t + bsel, 0.f);
    float z[C];
#pragma unroll
    for (int c = 0; c < C; ++c) z[c] = __shfl_sync(0xffffffffu, t, grp8 + 2 * c);
    const long long f = f0 + (lane >> 3);
    if (f < n) {
      int best = 0;
      float m = z[0];
#pragma unroll
      for (int c = 1; c < C; ++c) if (z[c] > m) { m = z[c]; best = c; }
      float e[C], sum = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) { e[c] = expf(z[c] - m); sum += e[c]; }
      const float inv = 1.0f / sum;
      const int l8 = lane & 7;                   // lane c of each 8-lane group writes class c of its frame
      float zsel = z[0], psel = e[0] * inv;
#pragma unroll
      for (int c = 1; c < C; ++c) if (l8 == c) { zsel = z[c]; psel = e[c] * inv; }
      if (l8 < C) {
        if (dense) dense[f * C + l8] = zsel;
        if (probs) probs[f * C + l8] = psel;
      }
      if (l8 == 0 && cls) cls[f] = best;
      cnt += (l8 == best);
    }
    __syncwarp();                                // every lane has read this pass's buffer before it is refilled
  }
  if (hist) {
    // lanes 8 j + c counted class c for the frames j of this warp's passes
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 8);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 16);
    if (lane < C && cnt) atomicAdd(hist + lane, (unsigned long long)cnt);
  }
}

__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// Any F<=16, C<=16 (runtime loops; slow path for shapes without a specialisation).
__global__ void __launch_bounds__(256)
tiny_f32_generic_kernel(const TinyParams p, const float* __restrict__ dmain,
                        const float* __restrict__ dtail, const float4* __restrict__ x, long long n,
                        float* __restrict__ probs, float* __restrict__ dense, int* __restrict__ cls,
                        unsigned long long* __restrict__ hist) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int F = p.F, C = p.C;
  unsigned cnt = 0;
  for (long long f = warp; f < n; f += nwarps) {
    const float4 xi = ldg_stream_f4(x + f * 64 + lane), xq = ldg_stream_f4(x + f * 64 + 32 + lane);
    float pi = __shfl_up_sync(0xffffffffu, xi.w, 1), pq = __shfl_up_sync(0xffffffffu, xq.w, 1);
    if (lane == 0) { pi = 0.f; pq = 0.f; }
    const float I[5] = {pi, xi.x, xi.y, xi.z, xi.w}, Q[5] = {pq, xq.x, xq.y, xq.z, xq.w};
    float acc[kMaxClasses];
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) acc[c] = 0.f;
    for (int k = 0; k < F; ++k) {
      const float k0 = p.conv[3 * k], k1 = p.conv[3 * k + 1], b = p.conv[3 * k + 2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float yi = fmaxf(fmaf(I[i], k0, fmaf(I[i + 1], k1, b)), 0.f);
        const float yq = fmaxf(fmaf(Q[i], k0, fmaf(Q[i + 1], k1, b)), 0.f);
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c) {
          if (c < C) {
            acc[c] = fmaf(yi, __ldg(dmain + ((0 * F + k) * C + c) * 128 + 4 * lane + i), acc[c]);
            acc[c] = fmaf(yq, __ldg(dmain + ((1 * F + k) * C + c) * 128 + 4 * lane + i), acc[c]);
          }
        }
      }
      if (lane == 31) {  // position 128
        const float yi = fmaxf(fmaf(xi.w, k0, b), 0.f), yq = fmaxf(fmaf(xq.w, k0, b), 0.f);
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c) {
          if (c < C) {
            acc[c] = fmaf(yi, __ldg(dtail + (0 * F + k) * C + c), acc[c]);
            acc[c] = fmaf(yq, __ldg(dtail + (1 * F + k) * C + c), acc[c]);
          }
        }
      }
    }
    float z[kMaxClasses];
    float m = -1.f;
    int best = 0;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) {
      if (c < C) {
        float a = acc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        z[c] = fmaxf(a + p.bias[c], 0.f);
        if (z[c] > m) { m = z[c]; best = c; }
      }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) if (c < C) s += expf(z[c] - m);
    const float inv = 1.0f / s;
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < kMaxClasses; ++c) {
        if (c < C) {
          if (dense) dense[f * C + c] = z[c];
          if (probs) probs[f * C + c] = expf(z[c] - m) * inv;
        }
      }
      if (cls) cls[f] = best;
    }
    cnt += (lane == best);
  }
  if (hist && lane < C && cnt) atomicAdd(hist + lane, (unsigned long long)cnt);
}

int pack_tiny(mdc_handle_s* h) {
  const int F = h->F, C = h->C;
  const std::vector<float>& D = h->w[MDC_T_DENSE1_K];   // (2*129*F, C)
  std::vector<float> main_img((size_t)2 * F * C * 128), tail((size_t)2 * F * C);
  for (int r = 0; r < 2; ++r)
    for (int f = 0; f < F; ++f)
      for (int c = 0; c < C; ++c) {
        for (int p = 0; p < 128; ++p)
          main_img[(((size_t)r * F + f) * C + c) * 128 + p] = D[((size_t)(r * 129 + p) * F + f) * C + c];
        tail[((size_t)r * F + f) * C + c] = D[((size_t)(r * 129 + 128) * F + f) * C + c];
      }
  if (int e = h->tiny_dense.reserve(main_img.size() * 4)) return e;
  if (int e = h->tiny_bias.reserve(tail.size() * 4)) return e;
  MDC_CUDA(cudaMemcpy(h->tiny_dense.ptr, main_img.data(), main_img.size() * 4, cudaMemcpyHostToDevice));
  MDC_CUDA(cudaMemcpy(h->tiny_bias.ptr, tail.data(), tail.size() * 4, cudaMemcpyHostToDevice));
  return MDC_OK;
}

int launch_tiny_f32(mdc_handle_s* h, const void* xv, int in_fmt, int64_t n, float* probs, float* dense,
                    int32_t* cls, unsigned long long* hist, cudaStream_t stream) {
  if (n == 0) return MDC_OK;
  const float* x = reinterpret_cast<const float*>(xv);
  const uint8_t* xb = reinterpret_cast<const uint8_t*>(xv);
  const bool special = (h->F == 3 || h->F == 10) && h->C == 3;
  MDC_REQUIRE(in_fmt == MDC_IN_F32 || special, MDC_ERR_UNSUPPORTED,
              "raw u8 / int16 frames: TinyCNN2 shapes F in {3, 10}, C = 3 only (this handle: F=%d, C=%d)", h->F, h->C);
  TinyParams p;
  const int F = h->F, C = h->C;
  const float* K = h->w[MDC_T_CONV1_K].data();   // (1,2,1,F): [tap][f]
  const float* B = h->w[MDC_T_CONV1_B].data();
  for (int f = 0; f < kMaxFilters; ++f) {
    p.conv[3 * f] = f < F ? K[f] : 0.f;
    p.conv[3 * f + 1] = f < F ? K[F + f] : 0.f;
    p.conv[3 * f + 2] = f < F ? B[f] : 0.f;
  }
  for (int c = 0; c < kMaxClasses; ++c) p.bias[c] = c < C ? h->w[MDC_T_DENSE1_B][c] : 0.f;
  p.F = F;
  p.C = C;
  const int threads = 256;
  const float4* dm = reinterpret_cast<const float4*>(h->tiny_dense.ptr);
  const float* dt = reinterpret_cast<const float*>(h->tiny_bias.ptr);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  auto grid = [&](int R) {
    long long warps = (n + R - 1) / R;
    long long blocks = (warps * 32 + threads - 1) / threads;
    long long max_blocks = (long long)h->num_sms * 2 * 4;
    return (unsigned)(blocks > max_blocks ? max_blocks : blocks);
  };
#define MDC_TINY4_LAUNCH(F_, C_, W_, B_, FMT_)                                                                \
  do {                                                                                                        \
    using Cfg_ = Tiny4<F_, C_, W_, B_, FMT_>;                                                                 \
    static bool attr_ = false;                                                                                \
    if (!attr_) {                                                                                             \
      MDC_CUDA(cudaFuncSetAttribute(tiny_f32_kernel4<F_, C_, W_, B_, FMT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg_::total)); \
      attr_ = true;                                                                                           \
    }                                                                                                         \
    const long long ngroups_ = (n + Cfg_::R - 1) / Cfg_::R;                                                   \
    const long long blocks_ = (ngroups_ + Cfg_::kWarps - 1) / Cfg_::kWarps;                                   \
    const unsigned grid_ = (unsigned)(blocks_ < h->num_sms ? blocks_ : h->num_sms);                           \
    cudaLaunchConfig_t cfg_ = {};                                                                             \
    cfg_.gridDim = dim3(grid_);                                                                               \
    cfg_.blockDim = dim3(Cfg_::kThreads);                                                                     \
    cfg_.dynamicSmemBytes = Cfg_::total;                                                                      \
    cfg_.stream = stream;                                                                                     \
    cudaLaunchAttribute pdl_[1];                                                                              \
    pdl_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                          \
    pdl_[0].val.programmaticStreamSerializationAllowed = 1;                                                   \
    cfg_.attrs = pdl_;                                                                                        \
    cfg_.numAttrs = 1;                                                                                        \
    MDC_CUDA(cudaLaunchKernelEx(&cfg_, tiny_f32_kernel4<F_, C_, W_, B_, FMT_>, p4, dm, dt, xb, (long long)n, probs, dense, cls, hist)); \
  } while (0)
  TinyParams4 p4;
  for (int f = 0; f < kMaxFilters; ++f) {
    auto dup = [](float v) {
      unsigned u;
      memcpy(&u, &v, 4);
      return ((unsigned long long)u << 32) | u;
    };
    p4.k0[f] = dup(p.conv[3 * f]);
    p4.k1[f] = dup(p.conv[3 * f + 1]);
    p4.b[f] = dup(p.conv[3 * f + 2]);
  }
  for (int c = 0; c < kMaxClasses; ++c) p4.bias[c] = p.bias[c];
  p4.zero = 0.f;
  prof_begin(h, stream);
  // 16 warps x 3 buffers: measured against 14 warps (F = 10: 1.60e9 -> 1.84e9 frames/s - the kernel wants warps to
  // fill the FMA pipe's off-cycles); 18+ warps would leave < 128 registers per thread and spill
#define MDC_TINY4_FORMATS(F_)                                                 \
  do {                                                                         \
    if (in_fmt == MDC_IN_U8IQ) MDC_TINY4_LAUNCH(F_, 3, 16, 3, MDC_IN_U8IQ);    \
    else if (in_fmt == MDC_IN_I16) MDC_TINY4_LAUNCH(F_, 3, 16, 3, MDC_IN_I16); \
    else MDC_TINY4_LAUNCH(F_, 3, 16, 3, MDC_IN_F32);                           \
  } while (0)
  if (F == 3 && C == 3) {
    MDC_TINY4_FORMATS(3);
  } else if (F == 10 && C == 3) {
    MDC_TINY4_FORMATS(10);
  } else {
    tiny_f32_generic_kernel<<<grid(1), threads, 0, stream>>>(
        p, reinterpret_cast<const float*>(h->tiny_dense.ptr), dt, x4, n, probs, dense, cls, hist);
  }
  prof_end(h, stream);
  h->launches++;
  MDC_CUDA(cudaGetLastError());
  return MDC_OK;
}

}  // namespace mdc
