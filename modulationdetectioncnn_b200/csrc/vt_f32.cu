// VT-CNN2 forward in fp32 on CUDA cores: the parity mode (MDC_MODE_FP32).
//
// Layer stack: /root/reference/examples-master/modulation_recognition/
// RML2016.10a_VTCNN2_example.ipynb:231-243 (shapes :194-216); Dropout = identity.
//   xp1 = pad(x,2)                     (2,132)
//   a   = relu(conv1x3(xp1; W1,b1))    (2,130,256)   -> stored padded by 2: (2,134,256)
//   c   = relu(conv2x3(pad(a,2); W2))  (132,80)      implicit GEMM, K = 2*3*256 = 1536
//   h   = relu(flat(c) W3 + b3)        (256)         GEMM, K = 10560
//   y   = softmax(h W4 + b4)           (C)
// This path trades speed for fp32 accumulate everywhere; the tensor-core path is vt_bf16.cu.
#include "mdc_internal.cuh"

namespace mdc {

constexpr int kA1Rows = 2 * 134;   // padded conv1 rows per frame (r, q)

// a1p[f][r][q][ch], q in [0,134): zero for q<2 or q>=132, else conv1 at q-2.
__global__ void __launch_bounds__(256)
vt_conv1_f32_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                    const float* __restrict__ b1, float* __restrict__ a1p, long long n) {
  const int ch = threadIdx.x;
  const float k0 = w1[ch], k1 = w1[256 + ch], k2 = w1[512 + ch], b = b1[ch];
  __shared__ float xs[2][128 + 6];
  for (long long f = blockIdx.x; f < n; f += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * 134; i += 256) {
      const int r = i / 134, q = i % 134;   // xs[r][q] = xp[q-1] with xp1 index = q + 1 ... see below
      const int xi = q - 3;                 // xs[r][q] = x[r][q-3] (zero outside)
      xs[r][q] = (xi >= 0 && xi < 128) ? x[f * 256 + r * 128 + xi] : 0.f;
    }
    __syncthreads();
    float* dst = a1p + (size_t)f * kA1Rows * 256;
    for (int rq = 0; rq < kA1Rows; ++rq) {
      const int r = rq / 134, q = rq % 134;
      float v = 0.f;
      if (q >= 2 && q < 132) {
        // conv1 position q' = q-2 uses xp1[q'+t] = x[q'+t-2] = x[q+t-4] = xs[r][q+t-1]
        v = fmaxf(fmaf(xs[r][q - 1], k0, fmaf(xs[r][q], k1, fmaf(xs[r][q + 1], k2, b))), 0.f);
      }
      dst[(size_t)rq * 256 + ch] = v;
    }
  }
}

struct Conv2Rows {
  const float* a1p;
  __device__ __forceinline__ const float* operator()(long long m, int k0) const {
    const long long f = m / 132;
    const int u = (int)(m - f * 132);
    const int r = k0 / 768, kk = k0 - r * 768;
    return a1p + ((size_t)f * kA1Rows + r * 134 + u) * 256 + kk;
  }
};
struct PlainRows {
  const float* a;
  long long lda;
  __device__ __forceinline__ const float* operator()(long long m, int k0) const {
    return a + (size_t)m * lda + k0;
  }
};

// C[m][n] = relu?(sum_k A(m,k) B[k][n] + bias[n]);  B row-major [K][N]
template <int BM, int BN, int TM, int TN, class ARow>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_bias_act_kernel(ARow arow, const float* __restrict__ B, const float* __restrict__ bias,
                      float* __restrict__ C, long long M, int N, int K, int relu) {
  constexpr int BK = 16;
  constexpr int NT = (BM / TM) * (BN / TN);
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    for (int i = tid; i < BM * (BK / 4); i += NT) {
      const int row = i / (BK / 4), q = i % (BK / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + row < M) v = *reinterpret_cast<const float4*>(arow(m0 + row, k0) + 4 * q);
      As[4 * q][row] = v.x; As[4 * q + 1][row] = v.y; As[4 * q + 2][row] = v.z; As[4 * q + 3][row] = v.w;
    }
    for (int i = tid; i < BK * BN; i += NT) {
      const int kk = i / BN, nn = i % BN;
      Bs[kk][nn] = (n0 + nn < N) ? B[(size_t)(k0 + kk) * N + n0 + nn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int nn = n0 + tx * TN + j;
      if (nn < N) {
        float v = acc[i][j] + bias[nn];
        C[(size_t)m * N + nn] = relu ? fmaxf(v, 0.f) : v;
      }
    }
  }
}

constexpr int kHeadThreads = 384;   // 12 warps x <= 170 registers: one block per SM
// logits = h W4 + b4, softmax, argmax, histogram.  One warp per frame; lane l owns h[8l .. 8l+7] (two 16-B
// loads) and keeps its 8 x C slice of W4 in registers for the whole (persistent) kernel.  Shared with
// vt_tensor.cu.
template <int C>
__global__ void __launch_bounds__(kHeadThreads, 1)
vt_head_kernel(const float* __restrict__ hbuf, const float* __restrict__ w4, const float* __restrict__ b4,
               long long n, float* __restrict__ probs, float* __restrict__ logits_out,
               int* __restrict__ cls, unsigned long long* __restrict__ hist) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float w[8][C], bias[C];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < C; ++c) w[i][c] = __ldg(w4 + (lane * 8 + i) * C + c);
#pragma unroll
  for (int c = 0; c < C; ++c) bias[c] = __ldg(b4 + c);
  unsigned cnt = 0;
  long long f = warp;
  float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0;
  if (f < n) {
    h0 = __ldcs(reinterpret_cast<const float4*>(hbuf + f * 256) + lane * 2);
    h1 = __ldcs(reinterpret_cast<const float4*>(hbuf + f * 256) + lane * 2 + 1);
  }
  while (f < n) {
    const long long fn = f + nwarps;
    float4 n0 = h0, n1 = h1;
    if (fn < n) {   // next frame of this warp
      n0 = __ldcs(reinterpret_cast<const float4*>(hbuf + fn * 256) + lane * 2);
      n1 = __ldcs(reinterpret_cast<const float4*>(hbuf + fn * 256) + lane * 2 + 1);
    }
    const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
    float z[C];
    float m = -3.4e38f;
    int best = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) a = fmaf(hv[i], w[i][c], a);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      z[c] = a + bias[c];
      if (z[c] > m) { m = z[c]; best = c; }
    }
    float e[C], s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { e[c] = expf(z[c] - m); s += e[c]; }
    const float inv = 1.0f / s;
    // lane c writes class c: one coalesced 4C-byte store per output instead of C stores from lane 0
    float zl = z[0], pl = e[0];
#pragma unroll
    for (int c = 1; c < C; ++c) if (lane == c) { zl = z[c]; pl = e[c]; }
    if (lane < C) {
      if (logits_out) logits_out[f * C + lane] = zl;
      if (probs) probs[f * C + lane] = pl * inv;
    }
    if (lane == 0 && cls) cls[f] = best;
    cnt += (lane == best);
    h0 = n0; h1 = n1;
    f = fn;
  }
  if (hist && lane < C && cnt) atomicAdd(hist + lane, (unsigned long long)cnt);
}

template <int C>
static void head_launch(unsigned blocks, cudaStream_t stream, const float* hbuf, const float* w4, const float* b4,
                        long long n, float* probs, float* dense, int* cls, unsigned long long* hist) {
  vt_head_kernel<C><<<blocks, kHeadThreads, 0, stream>>>(hbuf, w4, b4, n, probs, dense, cls, hist);
}

int launch_vt_head(mdc_handle_s* h, const float* hbuf, int64_t n, float* probs, float* dense,
                   int32_t* cls, unsigned long long* hist, cudaStream_t stream) {
  long long blocks = (n * 32 + kHeadThreads - 1) / kHeadThreads;
  const long long maxb = (long long)h->num_sms;
  if (blocks > maxb) blocks = maxb;
  const float* w4 = (const float*)h->vt_w4.ptr;
  const float* b4 = (const float*)h->vt_b4.ptr;
  const unsigned g = (unsigned)blocks;
  switch (h->C) {
#define MDC_HEAD_CASE(CC) case CC: head_launch<CC>(g, stream, hbuf, w4, b4, n, probs, dense, cls, hist); break;
    MDC_HEAD_CASE(1) MDC_HEAD_CASE(2) MDC_HEAD_CASE(3) MDC_HEAD_CASE(4) MDC_HEAD_CASE(5) MDC_HEAD_CASE(6)
    MDC_HEAD_CASE(7) MDC_HEAD_CASE(8) MDC_HEAD_CASE(9) MDC_HEAD_CASE(10) MDC_HEAD_CASE(11) MDC_HEAD_CASE(12)
    MDC_HEAD_CASE(13) MDC_HEAD_CASE(14) MDC_HEAD_CASE(15) MDC_HEAD_CASE(16)
#undef MDC_HEAD_CASE
    default:
      set_error("classes=%d outside 1..16", h->C);
      return MDC_ERR_UNSUPPORTED;
  }
  h->launches++;
  MDC_CUDA(cudaGetLastError());
  return MDC_OK;
}

static int upload(DeviceBuffer& b, const float* src, size_t count) {
  if (int e = b.reserve(count * sizeof(float))) return e;
  MDC_CUDA(cudaMemcpy(b.ptr, src, count * sizeof(float), cudaMemcpyHostToDevice));
  return MDC_OK;
}

// dense1 kernel rows re-ordered to this library's activation order (pos*80 + ch)
void vt_permute_w3(const mdc_handle_s* h, std::vector<float>& out) {
  const std::vector<float>& w3 = h->w[MDC_T_DENSE1_K];
  out.resize(w3.size());
  if (h->flatten_order == 0) { out = w3; return; }
  for (int pos = 0; pos < kVtPos2; ++pos)
    for (int ch = 0; ch < kVtC2; ++ch)
      memcpy(&out[((size_t)pos * kVtC2 + ch) * kVtH], &w3[((size_t)ch * kVtPos2 + pos) * kVtH],
             kVtH * sizeof(float));
}

int pack_vt_small(mdc_handle_s* h) {   // tensors both VT paths use in fp32
  if (int e = upload(h->vt_w1, h->w[MDC_T_CONV1_K].data(), 768)) return e;
  if (int e = upload(h->vt_b1, h->w[MDC_T_CONV1_B].data(), 256)) return e;
  if (int e = upload(h->vt_b2, h->w[MDC_T_CONV2_B].data(), 80)) return e;
  if (int e = upload(h->vt_b3, h->w[MDC_T_DENSE1_B].data(), 256)) return e;
  if (int e = upload(h->vt_w4, h->w[MDC_T_DENSE2_K].data(), (size_t)256 * h->C)) return e;
  if (int e = upload(h->vt_b4, h->w[MDC_T_DENSE2_B].data(), h->C)) return e;
  return MDC_OK;
}

int pack_vt_f32(mdc_handle_s* h) {
  if (int e = pack_vt_small(h)) return e;
  if (int e = upload(h->vt_w2, h->w[MDC_T_CONV2_K].data(), (size_t)1536 * 80)) return e;
  std::vector<float> w3p;
  vt_permute_w3(h, w3p);
  return upload(h->vt_w3, w3p.data(), w3p.size());
}

int launch_vt_f32(mdc_handle_s* h, const float* x, int64_t n, float* probs, float* dense,
                  int32_t* cls, unsigned long long* hist, cudaStream_t stream) {
  constexpr int64_t CH = 1024;
  const int64_t ch = n < CH ? n : CH;
  if (int e = h->ws_a1.reserve((size_t)ch * kA1Rows * 256 * 4)) return e;
  if (int e = h->ws_act.reserve((size_t)ch * kVtFlat * 4)) return e;
  if (int e = h->ws_h.reserve((size_t)ch * kVtH * 4)) return e;
  float* a1p = (float*)h->ws_a1.ptr;
  float* act = (float*)h->ws_act.ptr;
  float* hb = (float*)h->ws_h.ptr;
  for (int64_t s = 0; s < n; s += CH) {
    const int64_t m = (n - s) < CH ? (n - s) : CH;
    vt_conv1_f32_kernel<<<(unsigned)m, 256, 0, stream>>>(x + s * 256, (const float*)h->vt_w1.ptr,
                                                        (const float*)h->vt_b1.ptr, a1p, m);
    prof_begin(h, stream);
    {
      const long long M = m * 132;
      dim3 grid((unsigned)((M + 127) / 128), 1);
      sgemm_bias_act_kernel<128, 80, 4, 10, Conv2Rows><<<grid, 256, 0, stream>>>(
          Conv2Rows{a1p}, (const float*)h->vt_w2.ptr, (const float*)h->vt_b2.ptr, act, M, 80, 1536, 1);
    }
    prof_end(h, stream);
    {
      dim3 grid((unsigned)((m + 63) / 64), 256 / 64);
      sgemm_bias_act_kernel<64, 64, 4, 4, PlainRows><<<grid, 256, 0, stream>>>(
          PlainRows{act, kVtFlat}, (const float*)h->vt_w3.ptr, (const float*)h->vt_b3.ptr, hb, m, 256,
          kVtFlat, 1);
    }
    h->launches += 3;
    MDC_CUDA(cudaGetLastError());
    if (int e = launch_vt_head(h, hb, m, probs ? probs + s * h->C : nullptr,
                               dense ? dense + s * h->C : nullptr, cls ? cls + s : nullptr, hist, stream))
      return e;
  }
  return MDC_OK;
}

}  // namespace mdc
