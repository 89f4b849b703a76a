// TinyCNN2(F,C) fp32 forward: the nets the reference actually ships.
//
//   Reshape(2,128,1) -> ZeroPadding2D((0,0),(1,1)) -> Conv2D(F,(1,2),relu) -> Flatten
//   -> Dense(C, relu) -> softmax            (/root/reference/CNN.ipynb:1 cell 6; h5 model_config)
//
//   y[r][p][f] = relu(xp[r][p]*k0[f] + xp[r][p+1]*k1[f] + b[f]),  p = 0..128, xp[0]=xp[129]=0
//   z[c]       = relu(sum_{r,p,f} y[r][p][f] * D[(r*129 + p)*F + f][c] + d[c])
//
// Mapping: one warp per frame (R frames per pass).  Lane l owns positions p = 4l..4l+3 of
// both rows -> two coalesced 512 B float4 loads per frame per warp; position 128 (which
// only needs x[127]) is spread over lanes 0..2F-1, one (row,filter) pair each.  For F*C
// small enough (F=3,C=3: 72 registers) the Dense rows a lane needs stay in registers for
// the whole persistent kernel; otherwise they are re-read through L1 once per pass and
// applied to R frames.  Softmax, argmax and the class histogram are fused in the epilogue.
#include "mdc_internal.cuh"

namespace mdc {

struct TinyParams {
  float conv[3 * kMaxFilters];   // k0,k1,b per filter
  float bias[kMaxClasses];
  int F, C;
};

__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// dense image (packed on host):
//   main [r][f][c][128]  entry p = D[(r*129 + p)*F + f][c]            -> float4 per lane
//   tail [r][f][c]       = D[(r*129 + 128)*F + f][c]                  (position 128)
template <int F, int C, int R, bool WREG>
__global__ void __launch_bounds__(256, 2)
tiny_f32_kernel(const TinyParams p, const float4* __restrict__ dmain, const float* __restrict__ dtail,
                const float4* __restrict__ x, long long n, float* __restrict__ probs,
                float* __restrict__ dense, int* __restrict__ cls,
                unsigned long long* __restrict__ hist) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;

  float4 w[WREG ? 2 * F * C : 1];
  if (WREG) {
#pragma unroll
    for (int i = 0; i < 2 * F * C; ++i) w[i] = __ldg(dmain + i * 32 + lane);
  }
  // position 128: lane j < 2F handles (r = j / F, f = j % F)
  float tw[C];
  float tk0 = 0.f, tb = 0.f;
  const int tr = lane / F;
#pragma unroll
  for (int c = 0; c < C; ++c) tw[c] = 0.f;
  if (lane < 2 * F) {
    const int tf = lane % F;
    tk0 = p.conv[3 * tf];
    tb = p.conv[3 * tf + 2];
#pragma unroll
    for (int c = 0; c < C; ++c) tw[c] = __ldg(dtail + (tr * F + tf) * C + c);
  }
  unsigned cnt = 0;

  for (long long f0 = warp * R; f0 < n; f0 += nwarps * R) {
    float4 xi[R], xq[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long f = f0 + r < n ? f0 + r : n - 1;
      xi[r] = ldg_stream_f4(x + f * 64 + lane);
      xq[r] = ldg_stream_f4(x + f * 64 + 32 + lane);
    }
    float acc[R][C];
    float pi[R], pq[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      pi[r] = __shfl_up_sync(0xffffffffu, xi[r].w, 1);
      pq[r] = __shfl_up_sync(0xffffffffu, xq[r].w, 1);
      if (lane == 0) { pi[r] = 0.f; pq[r] = 0.f; }
      // position 128: xp[128] = x[127] (lane 31 .w), xp[129] = 0
      const float li = __shfl_sync(0xffffffffu, xi[r].w, 31);
      const float lq = __shfl_sync(0xffffffffu, xq[r].w, 31);
      const float yt = fmaxf(fmaf(tr ? lq : li, tk0, tb), 0.f);
#pragma unroll
      for (int c = 0; c < C; ++c) acc[r][c] = (lane < 2 * F) ? yt * tw[c] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < F; ++k) {
      const float k0 = p.conv[3 * k], k1 = p.conv[3 * k + 1], b = p.conv[3 * k + 2];
      float4 wi[C], wq[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        if (WREG) {
          wi[c] = w[(0 * F + k) * C + c];
          wq[c] = w[(1 * F + k) * C + c];
        } else {
          wi[c] = __ldg(dmain + ((0 * F + k) * C + c) * 32 + lane);
          wq[c] = __ldg(dmain + ((1 * F + k) * C + c) * 32 + lane);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float yi[4], yq[4];
        yi[0] = fmaxf(fmaf(pi[r], k0, fmaf(xi[r].x, k1, b)), 0.f);
        yi[1] = fmaxf(fmaf(xi[r].x, k0, fmaf(xi[r].y, k1, b)), 0.f);
        yi[2] = fmaxf(fmaf(xi[r].y, k0, fmaf(xi[r].z, k1, b)), 0.f);
        yi[3] = fmaxf(fmaf(xi[r].z, k0, fmaf(xi[r].w, k1, b)), 0.f);
        yq[0] = fmaxf(fmaf(pq[r], k0, fmaf(xq[r].x, k1, b)), 0.f);
        yq[1] = fmaxf(fmaf(xq[r].x, k0, fmaf(xq[r].y, k1, b)), 0.f);
        yq[2] = fmaxf(fmaf(xq[r].y, k0, fmaf(xq[r].z, k1, b)), 0.f);
        yq[3] = fmaxf(fmaf(xq[r].z, k0, fmaf(xq[r].w, k1, b)), 0.f);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float a = acc[r][c];
          a = fmaf(yi[0], wi[c].x, a); a = fmaf(yi[1], wi[c].y, a);
          a = fmaf(yi[2], wi[c].z, a); a = fmaf(yi[3], wi[c].w, a);
          a = fmaf(yq[0], wq[c].x, a); a = fmaf(yq[1], wq[c].y, a);
          a = fmaf(yq[2], wq[c].z, a); a = fmaf(yq[3], wq[c].w, a);
          acc[r][c] = a;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float z[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float a = acc[r][c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        z[c] = fmaxf(a + p.bias[c], 0.f);
      }
      const long long f = f0 + r;
      if (f < n) {
        int best = 0;
        float m = z[0];
#pragma unroll
        for (int c = 1; c < C; ++c) if (z[c] > m) { m = z[c]; best = c; }
        float e[C], s = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) { e[c] = expf(z[c] - m); s += e[c]; }
        const float inv = 1.0f / s;
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            if (dense) dense[f * C + c] = z[c];
            if (probs) probs[f * C + c] = e[c] * inv;
          }
          if (cls) cls[f] = best;
        }
        cnt += (lane == best);
      }
    }
  }
  if (hist && lane < C && cnt) atomicAdd(hist + lane, (unsigned long long)cnt);
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

int launch_tiny_f32(mdc_handle_s* h, const float* x, int64_t n, float* probs, float* dense,
                    int32_t* cls, unsigned long long* hist, cudaStream_t stream) {
  if (n == 0) return MDC_OK;
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
  prof_begin(h, stream);
  if (F == 3 && C == 3) {
    tiny_f32_kernel<3, 3, 1, true><<<grid(1), threads, 0, stream>>>(p, dm, dt, x4, n, probs, dense, cls, hist);
  } else if (F == 10 && C == 3) {
    tiny_f32_kernel<10, 3, 4, false><<<grid(4), threads, 0, stream>>>(p, dm, dt, x4, n, probs, dense, cls, hist);
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
