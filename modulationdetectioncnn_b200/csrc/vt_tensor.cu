// VT-CNN2 forward on the sm_100a tensor cores, three arithmetic modes:
//   MDC_MODE_BF16    bf16 operands, fp32 accumulate (fast mode, ~6e-3 of the largest logit)
//   MDC_MODE_F16X3   every fp32 operand split into fp16 hi + fp16 lo * 2^-11, three kind::f16 MMAs per product
//                    (hi*hi + hi*lo + lo*hi) at the FULL 16-bit tensor rate: fp32-level accuracy
//                    (<= 1e-5 of the fp64 oracle) at ~1/3 of the bf16 rate
//   MDC_MODE_TF32X3  the same split in tf32 (three kind::tf32 MMAs at half rate, ~1/6 of the bf16 rate); no
//                    range restriction - the fallback when activations leave the fp16 range
//
// Layer stack: /root/reference/examples-master/modulation_recognition/
// RML2016.10a_VTCNN2_example.ipynb:231-243 (shapes :194-216); Dropout = identity.
//
// Persistent, warp-specialised tcgen05 kernels:
//
//   vt_conv_kernel<MODE>      conv1 (1x3, 256 ch, fp32 FMA on CUDA cores, produced straight into the
//                             shared-memory A operand) -> conv2 (2x3, 80 ch) as an implicit GEMM
//                             M = frames*132, N = 80, K = 3 taps x 512 (row,channel) -> +bias, ReLU
//                             -> activations act[frames*132][80]  (== Keras channels_last flatten):
//                             bf16, fp16 hi / lo matrices (F16X3) or fp32 hi / lo matrices (TF32X3).
//                             Frames arrive as f32 [n,2,128] (what predict receives), raw interleaved u8 I/Q
//                             (RTL-SDR bytes, README.md:5) or int16 Q6.12 - converted in the frame load.
//   vt_dense_bf16_kernel<C>   act[frames][10560] x W3 -> +bias, ReLU -> Dense(C) -> softmax, argmax,
//                             histogram in the epilogue (TMA 128B-swizzled tiles, M = 256 per CTA,
//                             N = 256, K = 10560); h never goes to HBM
//   vt_dense_f16x3_kernel     dense1 with fp16 hi/lo operands: hi*hi runs and the cross terms in separate TMEM
//   vt_dense_tf32x3_kernel    accumulators / the same in 3xTF32, fp32 master sums in registers, then
//   vt_head_kernel            Dense(C) + softmax + argmax + histogram (fp32, vt_f32.cu)
//
// The implicit GEMM keeps conv1's padded output positions as GEMM rows: frame f owns rows
// [132 f, 132 f + 134) of one long activation "tape" whose rows 132 f and 132 f + 1 are the zero
// padding shared by frame f-1 (right pad) and frame f (left pad).  conv2 output row R needs tape
// rows R, R+1, R+2, so with the no-swizzle K-major operand layout (8-channel groups, rows 16 B
// apart) tap j is the SAME shared-memory image with the descriptor start address moved by 16 j
// bytes - conv1 activations are produced once and read by three MMAs.
#include <cudaTypedefs.h>
#include <cuda_fp16.h>

#include <algorithm>

#include "mdc_internal.cuh"
#include "sm100.cuh"

namespace mdc {
using namespace sm100;

int launch_vt_head(mdc_handle_s* h, const float* hbuf, int64_t n, float* probs, float* dense,
                   int32_t* cls, unsigned long long* hist, cudaStream_t stream);
int pack_vt_small(mdc_handle_s* h);
void vt_permute_w3(const mdc_handle_s* h, std::vector<float>& out);

enum { kConvBF16 = 0, kConvTF32 = 1, kConvF16 = 2 };

// ------------------------------------------------------------------------------------------
// conv kernel geometry.  CTAs work in pairs (cluster of 2, tcgen05 cta_group::2): each CTA owns
// its own super-tile (tape rows, conv1 producers, accumulators, epilogue) and HALF of every W2
// chunk; one M = 256 MMA issued by the pair's leader covers a 128-row tile of each CTA, so the
// B operand is fetched from shared memory once per pair (measured: 45 cycles per MMA against 52
// for the single-CTA M = 128 x N = 80 shape, tools/umma_rate.cu / tools/umma2_probe.cu).
constexpr int kGroups = 4;                // 16-B K groups per chunk image: g = 2 * (channel block) + input row
constexpr int kBHalf = 40;                // W2 output channels held by each CTA of the pair
constexpr int kXFrames = 4;               // frames a super-tile's tape rows can touch
constexpr int kOutTile = 128 * 160;       // one 128 x 80 tile of 16-bit outputs
constexpr int kProdWarp0 = 6;

// A chunk is kCC conv1 channels x {I row, Q row} = two UMMA K steps of two 16-B groups:
//   bf16:   16 channels, 8 per group, one image         (K step = 16 values)
//   f16x3:  16 channels, 8 per group, hi and lo images   (K step = 16 values)
//   tf32x3:  8 channels, 4 per group, hi and lo images   (K step =  8 values)
//
// Split-mode accumulation.  tcgen05.mma rounds every accumulate TOWARD ZERO (tools/umma_acc_probe.cu: 1 + 0.75 ulp
// stays 1), so a chain of n MMAs into one accumulator comes out low by about n x 1.5e-8 relative (measured
// -8.8e-6 for the 576-MMA tf32 conv2 chain, -6e-5 for the 3,960-MMA dense1 chain).  The split modes therefore
// work on ONE 128-row tile per super-tile and spread its MMAs over several TMEM accumulators that the epilogue
// adds in fp32 round-to-nearest:
//   tf32x3: accumulator 0 takes the two small cross terms (lo*hi, hi*lo) of every K step, accumulators 1..5 the
//           hi*hi terms of chunks c = a - 1 (mod 5) - at most 42 truncating adds each; single-buffered
//           (6 x 80 = 480 of 512 columns: the MMA warp waits while the epilogue reads, ~4 % of a tile's 576 MMAs)
//   f16x3:  the B image of a CTA holds, per tap and K group, the hi rows of its 40 output channels followed by their
//           lo rows (which carry a factor 2^11, undone in the epilogue), so ONE N = 160 MMA A_hi x [B_hi | B_lo]
//           yields hi*hi and hi*lo - A_hi is fetched from shared memory once for both, the N = 80 shape is
//           operand-fetch-bound - into accumulator D1 (160 columns: per CTA half 40 hi*hi then 40 hi*lo), and an
//           N = 80 MMA A_lo x B_hi the lo*hi term into D2 (80 columns).  The hi*hi chain is 96 truncating adds
//           (~1.4e-6 low); 240 columns, double-buffered, so the next tile's MMAs run under the epilogue
template <int MODE>
struct ConvCfg {
  static constexpr bool kSplit = MODE != kConvBF16;        // hi / lo operand images, three MMAs per product
  static constexpr bool kWide = MODE == kConvTF32;         // 32-bit operand elements
  static constexpr int kNT = kSplit ? 1 : 3;               // accumulator tiles (128 rows x 80 cols) per super-tile
  static constexpr int kTapeRows = 128 * kNT;              // tape rows staged per super-tile
  static constexpr int kOutRows = kTapeRows - 2;           // conv2 rows produced per super-tile (2-row halo)
  static constexpr int kALbo = kTapeRows * 16;             // bytes between K groups of the A image
  // conv1 producers: one tape row per thread; f16x3 puts TWO threads on a row (8 of the chunk's 16 channels each):
  // the hi/lo split triples the per-value work while a chunk's MMAs take no longer than in bf16 mode, and with one
  // warp per scheduler the producers, not the tensor pipe, set the pace (measured 1,175 cycles per chunk)
  static constexpr int kProdSplit = MODE == kConvF16 ? 2 : 1;
  static constexpr int kProdWarps = kTapeRows / 32 * kProdSplit;
  static constexpr int kProdUnroll = MODE == kConvF16 ? 1 : 256 / (kWide ? 8 : 16);   // producers' chunk loop
  static constexpr int kThreads = (kProdWarp0 + kProdWarps) * 32;   // TMA, MMA, 4 epilogue, producers
  static constexpr int kAccBufs = MODE == kConvTF32 ? 1 : 2;        // accumulator buffers in TMEM
  static constexpr int kAccSplit = MODE == kConvBF16 ? 1 : (MODE == kConvTF32 ? 6 : 3);   // accumulators per tile
  static constexpr int kAccCols = kNT * kAccSplit * 80;    // TMEM columns per buffer
  static constexpr int kCC = kWide ? 8 : 16;
  static constexpr int kImgs = kSplit ? 2 : 1;
  static constexpr int kChunks = 256 / kCC;
  static constexpr int kStages = MODE == kConvF16 ? 4 : 5;
  static constexpr int kAImg = kGroups * kALbo;            // bf16 24,576; split modes 8,192
  static constexpr int kASlot = kImgs * kAImg;
  static constexpr int kBRows = MODE == kConvF16 ? 2 * kBHalf : kBHalf;   // f16x3: 40 hi rows then 40 lo rows
  static constexpr int kBLbo = kBRows * 16;                // bytes between K groups of the B image
  static constexpr int kBImgs = MODE == kConvTF32 ? 2 : 1; // tf32x3: separate hi and lo images
  static constexpr int kBImg = 3 * kGroups * kBLbo;        // [tap][group][rows][16 B]: 7,680 (f16x3 15,360)
  static constexpr int kBSlot = kBImgs * kBImg;
  static constexpr int kOutImgs = MODE == kConvTF32 ? 0 : kImgs;    // staged 16-bit output tiles per 128 rows
  // shared memory map
  static constexpr int a = 0;
  static constexpr int b = a + kStages * kASlot;
  static constexpr int xs = b + kStages * kBSlot;
  static constexpr int out = xs + 2 * kXFrames * 1024;     // two frame buffers before it
  static constexpr int b2 = out + 2 * kOutImgs * kOutTile; // two staged output tile sets (tf32x3 stores from registers)
  static constexpr int bars = b2 + 320;
  // full[S], empty[S], x_full[2], x_empty[2], tmem_full[2], tmem_empty[2]
  static constexpr int nbars = 2 * kStages + 4 + 4;
  static constexpr int tmem_slot = bars + nbars * 8;
  static constexpr int total = tmem_slot + 16;
  static_assert(kAccBufs * kAccCols <= 512, "accumulators exceed TMEM");
};
static_assert(ConvCfg<kConvBF16>::total <= 232448 && ConvCfg<kConvTF32>::total <= 232448 &&
              ConvCfg<kConvF16>::total <= 232448, "conv kernel shared memory exceeds 227 KB");

__device__ __forceinline__ uint64_t pack_dup(float v) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %1};" : "=l"(d) : "f"(v));
  return d;
}
__device__ __forceinline__ uint64_t fma2_u(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint32_t relu_pack(uint64_t v) {   // {lo, hi} fp32 -> bf16x2, ReLU
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return cvt_relu_bf16x2(hi, lo);
}

// fp16 hi/lo split of two non-negative fp32 values: hi = fp16(r), lo = fp16((r - hi) * 2^11); r - hi is exact
// in fp32, so hi + lo * 2^-11 carries 22 significant bits of r.  {second -> upper half}.
__device__ __forceinline__ void split_f16x2(float r0, float r1, uint32_t& hi, uint32_t& lo) {
  hi = cvt_f16x2_sat(r1, r0);
  float f0, f1;
  unpack_f16x2(hi, f0, f1);
  // (r - f) * 2^11 as two packed ops (exact: a power-of-two scale and an exact difference)
  uint64_t rr, ff, k, t;
  asm("mov.b64 %0, {%1, %2};" : "=l"(rr) : "f"(r0), "f"(r1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ff) : "f"(f0), "f"(f1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(k) : "f"(2048.f));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(rr), "l"(k));
  asm("mov.b64 %0, {%1, %1};" : "=l"(k) : "f"(-2048.f));
  t = fma2_u(ff, k, t);
  float t0, t1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(t));
  lo = cvt_f16x2_sat(t1, t0);
}

// conv1 weights travel as a kernel parameter (constant bank): with the chunk loop unrolled every
// weight is an immediate-offset uniform load feeding FFMA2 directly - no shared-memory broadcast
// loads (each LDS.128 broadcast costs 4 LSU wavefronts on the pipe the MMA operands also use) and
// no vector registers.  Layout: 32 channel groups x {w0[8], w1[8], w2[8], bias[8]} fp32.
struct ConvW1 {
  unsigned long long v[32 * 16];
};

// 8 conv1 channels of one tape row: relu(x0 w0 + x1 w1 + x2 w2 + b) as four packed fp32 pairs
__device__ __forceinline__ void conv1_fma(uint64_t x0, uint64_t x1, uint64_t x2, const unsigned long long* w,
                                          uint64_t (&a)[4]) {
  a[0] = fma2_u(x0, w[0], w[12]); a[1] = fma2_u(x0, w[1], w[13]);
  a[2] = fma2_u(x0, w[2], w[14]); a[3] = fma2_u(x0, w[3], w[15]);
  a[0] = fma2_u(x1, w[4], a[0]); a[1] = fma2_u(x1, w[5], a[1]);
  a[2] = fma2_u(x1, w[6], a[2]); a[3] = fma2_u(x1, w[7], a[3]);
  a[0] = fma2_u(x2, w[8], a[0]); a[1] = fma2_u(x2, w[9], a[1]);
  a[2] = fma2_u(x2, w[10], a[2]); a[3] = fma2_u(x2, w[11], a[3]);
}
// -> 8 x bf16 (16 B)
__device__ __forceinline__ uint4 conv1_item(uint64_t x0, uint64_t x1, uint64_t x2, const unsigned long long* w,
                                            uint32_t m) {
  uint64_t a[4];
  conv1_fma(x0, x1, x2, w, a);
  return make_uint4(relu_pack(a[0]) & m, relu_pack(a[1]) & m, relu_pack(a[2]) & m, relu_pack(a[3]) & m);
}
// -> 8 x fp16 hi (16 B) and 8 x fp16 lo (16 B)
__device__ __forceinline__ void conv1_item_f16(uint64_t x0, uint64_t x1, uint64_t x2, const unsigned long long* w,
                                               uint32_t m, uint4& hi, uint4& lo) {
  uint64_t a[4];
  conv1_fma(x0, x1, x2, w, a);
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v0, v1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v0), "=f"(v1) : "l"(a[i]));
    split_f16x2(fmaxf(v0, 0.f), fmaxf(v1, 0.f), h[i], l[i]);
  }
  hi = make_uint4(h[0] & m, h[1] & m, h[2] & m, h[3] & m);
  lo = make_uint4(l[0] & m, l[1] & m, l[2] & m, l[3] & m);
}

// 4 conv1 channels of one tape row in fp32, split for 3xTF32: hi = value truncated to tf32 (what the
// tensor core reads of an fp32 operand), lo = value - hi (exact)
__device__ __forceinline__ void conv1_quad(uint64_t x0, uint64_t x1, uint64_t x2, const unsigned long long* w, int q,
                                           uint32_t m, uint4& hi, uint4& lo) {
  uint64_t a0 = fma2_u(x0, w[2 * q], w[12 + 2 * q]), a1 = fma2_u(x0, w[2 * q + 1], w[13 + 2 * q]);
  a0 = fma2_u(x1, w[4 + 2 * q], a0); a1 = fma2_u(x1, w[5 + 2 * q], a1);
  a0 = fma2_u(x2, w[8 + 2 * q], a0); a1 = fma2_u(x2, w[9 + 2 * q], a1);
  float v[4];
  asm("mov.b64 {%0, %1}, %2;" : "=f"(v[0]), "=f"(v[1]) : "l"(a0));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(v[2]), "=f"(v[3]) : "l"(a1));
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float r = __uint_as_float(__float_as_uint(fmaxf(v[i], 0.f)) & m);
    h[i] = __float_as_uint(r) & 0xFFFFE000u;
    l[i] = __float_as_uint(r - __uint_as_float(h[i]));
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// one sample of a staged frame.  MDC_IN_F32: f32 [2][128];  MDC_IN_U8IQ: interleaved unsigned bytes I0 Q0 I1 Q1 ...,
// value (u - 127.5) / 128 = (2u - 255) / 256 exactly (== mdc_sdr_ingest_u8);  MDC_IN_I16: int16 [2][128] Q6.12
__device__ __forceinline__ float frame_sample(const uint8_t* frames, int fmt, int fi, int r, int xi) {
  if (fmt == MDC_IN_U8IQ) return (float)(2 * (int)frames[fi * 256 + 2 * xi + r] - 255) * (1.f / 256.f);
  if (fmt == MDC_IN_I16) return (float)reinterpret_cast<const int16_t*>(frames)[fi * 256 + r * 128 + xi] * (1.f / 4096.f);
  return reinterpret_cast<const float*>(frames)[fi * 256 + r * 128 + xi];
}

// act0/act1: bf16 mode -> act0 = bf16 [rows][80]; f16x3 -> act0 = hi, act1 = lo, fp16 [rows][80];
// tf32x3 -> act0 = hi, act1 = lo, fp32 [rows][80].  x: frames in format in_fmt (frame_sample).
// flags (f16x3): bit 0 is set when an input or activation leaves the range the fp16 split represents
// (|x| > x_limit - host-derived so that conv1 cannot exceed 65504 - or a conv2 activation > 65504).
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ConvCfg<MODE>::kThreads, 1)
vt_conv_kernel(const __grid_constant__ ConvW1 w1c, const uint8_t* __restrict__ x, int in_fmt, long long n,
               const float* __restrict__ b2g, const uint8_t* __restrict__ w2img,
               void* __restrict__ act0, void* __restrict__ act1, long long num_st, float x_limit,
               unsigned int* __restrict__ flags) {
  using ConvSmem = ConvCfg<MODE>;
  constexpr bool kSplit = ConvSmem::kSplit;
  constexpr int kStages = ConvSmem::kStages, kChunks = ConvSmem::kChunks;
  constexpr int kASlot = ConvSmem::kASlot, kBSlot = ConvSmem::kBSlot, kAImg = ConvSmem::kAImg, kBImg = ConvSmem::kBImg;
  constexpr int kNT = ConvSmem::kNT, kOutRows = ConvSmem::kOutRows, kALbo = ConvSmem::kALbo, kBLbo = ConvSmem::kBLbo;
  constexpr int kProdWarps = ConvSmem::kProdWarps, kConvThreads = ConvSmem::kThreads, kAccCols = ConvSmem::kAccCols;
  constexpr int kAccBufs = ConvSmem::kAccBufs;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ConvSmem::bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* x_full = bars + 2 * kStages;   // [2] frame buffers
  uint64_t* x_empty = x_full + 2;          // [2]
  uint64_t* tmem_full = x_empty + 2;       // [2] accumulator buffers
  uint64_t* tmem_empty = tmem_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + ConvSmem::tmem_slot);

  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const long long total_rows = n * 132;
  const uint32_t rank = cluster_ctarank();                 // 0 = leader (issues the pair's MMAs)
  // the pair walks super-tiles 2 i and 2 i + 1 in lockstep; a trailing odd one is an all-masked no-op
  const long long st_first = 2ll * cluster_id_x(), st_step = 2ll * cluster_count_x();
  const uint32_t frame_bytes = in_fmt == MDC_IN_U8IQ ? 256u : (in_fmt == MDC_IN_I16 ? 512u : 1024u);

  // ---- one-time setup
  for (int i = tid; i < 80; i += kConvThreads) reinterpret_cast<float*>(smem + ConvSmem::b2)[i] = b2g[i];
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      // own producers + own TMA (+ on the leader: the peer's relay once ITS stage is full)
      mbar_init(&full[s], kProdWarps + 1 + (rank == 0 ? 1 : 0));
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&x_full[b], 1);
      mbar_init(&x_empty[b], kProdWarps);
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 8);          // (leader's only) 4 epilogue warps of each CTA
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc_pair<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= TMA: frames of the super-tile + the W2 chunk stream (whole warp loops,
    // one elected lane issues, so every operand stays in uniform registers)
    uint32_t it = 0, k = 0;
    const uint8_t* w2half = w2img + (size_t)rank * kBSlot;   // image = [chunk][rank][hi/lo][tap][group][40][16 B]
    // frames of super-tile j go to buffer j & 1, one super-tile ahead of the producers
    auto load_frames = [&](uint32_t j, long long st) {
      const long long f0 = (st * kOutRows) / 132;
      const long long left = n - f0;
      const uint32_t nf = left <= 0 ? 0u : (left < kXFrames ? (uint32_t)left : (uint32_t)kXFrames);
      const uint32_t b = j & 1;
      mbar_wait(&x_empty[b], ((j >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&x_full[b], nf * frame_bytes);
        if (nf) bulk_g2s(smem + ConvSmem::xs + b * (kXFrames * 1024), x + f0 * frame_bytes, nf * frame_bytes, &x_full[b]);
      }
      __syncwarp();
    };
    if (st_first < num_st) load_frames(0, st_first + rank);
    for (long long base = st_first; base < num_st; base += st_step, ++k) {
      if (base + st_step < num_st) load_frames(k + 1, base + st_step + rank);
      for (int c = 0; c < kChunks; ++c, ++it) {
        const uint32_t s = it % kStages, ph = (it / kStages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[s], kBSlot);
          bulk_g2s(smem + ConvSmem::b + s * kBSlot, w2half + (size_t)c * 2 * kBSlot, kBSlot, &full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA; whole warp loops, one elected lane issues) /
    // relay (peer CTA: forwards "my stage is full" to the leader's barrier)
    uint32_t it = 0, k = 0;
    if (rank == 0) {
      const uint32_t idesc = MODE == kConvTF32 ? make_idesc_tf32(256, 80)
                                               : (MODE == kConvF16 ? make_idesc_f16(256, 80) : make_idesc_bf16(256, 80));
      const uint32_t idesc160 = make_idesc_f16(256, 160);    // f16x3: A_hi x [B_hi | B_lo]
      const uint32_t a_base = smem_u32(smem + ConvSmem::a), b_base = smem_u32(smem + ConvSmem::b);
      constexpr uint32_t hi = smem_desc_hi(128, 0);
      for (long long base = st_first; base < num_st; base += st_step, ++k) {
        const uint32_t buf = k % kAccBufs, use = k / kAccBufs;
        const uint32_t acc = tmem + buf * kAccCols;
        mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);          // both epilogues drained this buffer
        tc_fence_after_sync();
        for (int c = 0; c < kChunks; ++c, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          mbar_wait(&full[s], ph);               // own producers, own TMA and the peer's relay
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t a_lo = smem_desc_lo(a_base + s * kASlot, kALbo);
            const uint32_t b_lo = smem_desc_lo(b_base + s * kBSlot, kBLbo);
            // tf32x3: the hi*hi terms of chunk c go to accumulator 1 + c % 5, the cross terms to accumulator 0
            const uint32_t acc_hh = acc + (1 + c % 5) * 80;
            const bool hh_first_chunk = c < 5;
#pragma unroll
            for (int t = 0; t < kNT; ++t) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                  const uint32_t ao = ((2 * ks) * kALbo + (128 * t + j) * 16) >> 4;
                  const uint32_t bo = ((j * kGroups + 2 * ks) * kBLbo) >> 4;
                  if (!kSplit) {
                    mma_f16_ss_pair(acc + t * 80, desc64(a_lo + ao, hi), desc64(b_lo + bo, hi), idesc,
                                    (c | ks | j) != 0);
                  } else {
                    constexpr uint32_t al = kAImg >> 4, bl = kBImg >> 4;   // offsets of the lo images
                    const uint64_t ah_d = desc64(a_lo + ao, hi), al_d = desc64(a_lo + ao + al, hi);
                    const uint64_t bh_d = desc64(b_lo + bo, hi);
                    if (MODE == kConvTF32) {
                      const uint64_t bl_d = desc64(b_lo + bo + bl, hi);
                      mma_tf32_ss_pair(acc, al_d, bh_d, idesc, (c | ks | j) != 0);
                      mma_tf32_ss_pair(acc, ah_d, bl_d, idesc, 1);
                      mma_tf32_ss_pair(acc_hh, ah_d, bh_d, idesc, (!hh_first_chunk) || (ks | j) != 0);
                    } else {
                      // D1 (160 columns) += A_hi x [B_hi | B_lo]; D2 (80 columns) += A_lo x B_hi: the same B start
                      // address, 80 rows per CTA for the first, the leading 40 (hi) rows for the second
                      mma_f16_ss_pair(acc, ah_d, bh_d, idesc160, (c | ks | j) != 0);
                      mma_f16_ss_pair(acc + 160, al_d, bh_d, idesc, (c | ks | j) != 0);
                    }
                  }
                }
              }
            }
            mma_commit_pair(&empty[s]);
            if (c == kChunks - 1) mma_commit_pair(&tmem_full[buf]);
          }
          __syncwarp();
        }
      }
    } else {
      for (long long base = st_first; base < num_st; base += st_step) {
        for (int c = 0; c < kChunks; ++c, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          mbar_wait(&full[s], ph);
          if (elect_one()) mbar_arrive_remote(&full[s], 0);
          __syncwarp();
        }
      }
    }
  } else if (warp < kProdWarp0) {
    // ================= epilogue: TMEM -> +bias, ReLU -> 16-bit tile(s) in smem -> bulk store (bf16, f16x3)
    // or -> tf32 hi / lo fp32 rows straight to global (3xTF32 mode: the mainloop is 6x longer)
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const bool leader = (warp == 2 && lane == 0);
    const float4* b2s = reinterpret_cast<const float4*>(smem + ConvSmem::b2);
    uint32_t k = 0, tile_ctr = 0;
    float amax = 0.f;                            // f16x3: largest activation this thread has produced
    for (long long base = st_first; base < num_st; base += st_step, ++k) {
      const long long r0 = (base + rank) * kOutRows;
      const uint32_t buf = k % kAccBufs, use = k / kAccBufs;
      mbar_wait(&tmem_full[buf], use & 1);
      tc_fence_after_sync();
#pragma unroll 1
      for (int t = 0; t < kNT; ++t, ++tile_ctr) {
        uint8_t* obuf = smem + ConvSmem::out + (tile_ctr & 1) * (ConvSmem::kOutImgs * kOutTile);
        if (MODE != kConvTF32) {
          if (leader) bulk_wait_read<1>();        // the stores issued two tiles ago have drained obuf
          named_bar_sync(1, 128);
        }
        uint32_t v[80];
        const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + buf * kAccCols;
        if (MODE == kConvBF16) {
#pragma unroll
          for (int cc = 0; cc < 5; ++cc) {
            uint32_t(&vv)[16] = *reinterpret_cast<uint32_t(*)[16]>(&v[cc * 16]);
            tmem_ld16(tbase + t * 80 + cc * 16, vv);
          }
          tmem_ld_wait();
        } else if (MODE == kConvTF32) {
          // six partial accumulators -> one fp32 sum, round-to-nearest: ((h1 + h2) + (h3 + h4)) + h5, then the
          // small cross-term sum
#pragma unroll
          for (int cc = 0; cc < 5; ++cc) {
            uint32_t p[6][16];
#pragma unroll
            for (int a = 0; a < 6; ++a) tmem_ld16(tbase + a * 80 + cc * 16, p[a]);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float hh = ((__uint_as_float(p[1][e]) + __uint_as_float(p[2][e])) +
                                (__uint_as_float(p[3][e]) + __uint_as_float(p[4][e]))) + __uint_as_float(p[5][e]);
              v[cc * 16 + e] = __float_as_uint(hh + __uint_as_float(p[0][e]));
            }
          }
        } else {
          // hi*hi + 2^-11 x (hi*lo + lo*hi), fp32 round-to-nearest.  Output channels 40 hf + i, i < 40, sit at D1
          // columns 80 hf + i (hi*hi) and 80 hf + 40 + i (hi*lo), and at D2 column 160 + 40 hf + i (lo*hi)
#pragma unroll
          for (int o8 = 0; o8 < 10; ++o8) {
            const int d1 = 80 * (o8 / 5) + 8 * (o8 % 5);
            uint32_t hh[8], hl[8], lh[8];
            tmem_ld8(tbase + d1, hh);
            tmem_ld8(tbase + d1 + 40, hl);
            tmem_ld8(tbase + 160 + 8 * o8, lh);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 8; ++e)
              v[o8 * 8 + e] = __float_as_uint(fmaf(__uint_as_float(hl[e]) + __uint_as_float(lh[e]), 1.f / 2048.f,
                                                    __uint_as_float(hh[e])));
          }
        }
        if (t == kNT - 1) {                       // whole buffer read: hand it back to the MMA warp
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(&tmem_empty[buf]);
            else mbar_arrive_remote(&tmem_empty[buf], 0);
          }
        }
        const long long row_lo = r0 + 128 * t;
        long long rows = kOutRows - 128 * t;
        if (rows > 128) rows = 128;
        if (row_lo + rows > total_rows) rows = total_rows - row_lo;
        if (MODE == kConvTF32) {
          const int rr = q * 32 + lane;
          if (rr < rows) {
            float4* dh = reinterpret_cast<float4*>(reinterpret_cast<float*>(act0) + (row_lo + rr) * 80);
            float4* dl = reinterpret_cast<float4*>(reinterpret_cast<float*>(act1) + (row_lo + rr) * 80);
#pragma unroll
            for (int c4 = 0; c4 < 20; ++c4) {
              const float4 bb = b2s[c4];
              const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
              float h[4], l[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float r = fmaxf(__uint_as_float(v[c4 * 4 + e]) + bv[e], 0.f);
                h[e] = __uint_as_float(__float_as_uint(r) & 0xFFFFE000u);
                l[e] = r - h[e];
              }
              dh[c4] = make_float4(h[0], h[1], h[2], h[3]);
              dl[c4] = make_float4(l[0], l[1], l[2], l[3]);
            }
          }
          continue;
        }
        // rows are 160 B apart, so the 16-B chunks of lanes l and l + 4 fall into the same banks: lanes with
        // bit 2 set store their chunks rotated by one, which makes every quarter-warp store conflict-free
        uint8_t* orow = obuf + (q * 32 + lane) * 160;
        const bool rot = (lane & 4) != 0;
        if (MODE == kConvBF16) {
          uint4 o[10];
#pragma unroll
          for (int c8 = 0; c8 < 10; ++c8) {
            const float4 ba = b2s[2 * c8], bb = b2s[2 * c8 + 1];
            const uint32_t* vv = &v[c8 * 8];
            o[c8] = make_uint4(
                cvt_relu_bf16x2(__uint_as_float(vv[1]) + ba.y, __uint_as_float(vv[0]) + ba.x),
                cvt_relu_bf16x2(__uint_as_float(vv[3]) + ba.w, __uint_as_float(vv[2]) + ba.z),
                cvt_relu_bf16x2(__uint_as_float(vv[5]) + bb.y, __uint_as_float(vv[4]) + bb.x),
                cvt_relu_bf16x2(__uint_as_float(vv[7]) + bb.w, __uint_as_float(vv[6]) + bb.z));
          }
#pragma unroll
          for (int j = 0; j < 10; ++j) {
            const uint4 a = o[j], b = o[(j + 1) % 10];
            const uint4 val = make_uint4(rot ? b.x : a.x, rot ? b.y : a.y, rot ? b.z : a.z, rot ? b.w : a.w);
            *reinterpret_cast<uint4*>(orow + (rot ? ((j + 1) % 10) : j) * 16) = val;
          }
        } else {
          // fp16 hi tile, then the lo tile kOutTile bytes further.  Only rows that are stored count for the range
          // check: the tile's two halo rows (and rows past the last frame) are sums over whatever follows the A image
          const bool row_live = q * 32 + lane < rows;
          uint4 oh[10], ol[10];
#pragma unroll
          for (int c8 = 0; c8 < 10; ++c8) {
            const float4 ba = b2s[2 * c8], bb = b2s[2 * c8 + 1];
            const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float ra = fmaxf(__uint_as_float(v[c8 * 8 + 2 * e]) + bv[2 * e], 0.f);
              const float rb = fmaxf(__uint_as_float(v[c8 * 8 + 2 * e + 1]) + bv[2 * e + 1], 0.f);
              amax = fmaxf(amax, row_live ? fmaxf(ra, rb) : 0.f);
              split_f16x2(ra, rb, hw[e], lw[e]);
            }
            oh[c8] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            ol[c8] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
#pragma unroll
          for (int j = 0; j < 10; ++j) {
            const int jj = (j + 1) % 10;
            const uint4 a = oh[j], b = oh[jj], c = ol[j], d = ol[jj];
            *reinterpret_cast<uint4*>(orow + (rot ? jj : j) * 16) =
                make_uint4(rot ? b.x : a.x, rot ? b.y : a.y, rot ? b.z : a.z, rot ? b.w : a.w);
            *reinterpret_cast<uint4*>(orow + kOutTile + (rot ? jj : j) * 16) =
                make_uint4(rot ? d.x : c.x, rot ? d.y : c.y, rot ? d.z : c.z, rot ? d.w : c.w);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (leader) {
          if (rows > 0) {
            // (16-bit elements of either format: 160 B per row)
            bulk_s2g(reinterpret_cast<uint16_t*>(act0) + row_lo * 80, obuf, (uint32_t)rows * 160);
            if (MODE == kConvF16)
              bulk_s2g(reinterpret_cast<uint16_t*>(act1) + row_lo * 80, obuf + kOutTile, (uint32_t)rows * 160);
          }
          bulk_commit();
        }
      }
    }
    if (MODE != kConvTF32 && leader) bulk_wait<0>();
    if (MODE == kConvF16 && !(amax <= 65504.f)) atomicOr(flags, 1u);
  } else {
    // ================= conv1 producers: fp32 FMA -> ReLU -> operand format -> A operand image.
    // One tape row per thread; a chunk is kCC channels x {I row, Q row}; weights come from the
    // constant bank (uniform registers), inputs stay in registers for the whole super-tile.
    const int pw = warp - kProdWarp0;
    const int row = (pw % (kProdWarps / ConvSmem::kProdSplit)) * 32 + lane;
    const int phalf = pw / (kProdWarps / ConvSmem::kProdSplit);      // f16x3: which 8 of the chunk's 16 channels
    uint32_t it = 0, k = 0;
    float xmax = 0.f;
    for (long long base = st_first; base < num_st; base += st_step, ++k) {
      const long long t0 = (base + rank) * kOutRows;   // first tape row of this super-tile
      const long long f0 = t0 / 132;
      const long long tp = t0 + row;
      const long long f = tp / 132;
      const int p = (int)(tp - f * 132);
      const bool valid = (p >= 2) && (f < n);
      const uint32_t m = valid ? 0xFFFFFFFFu : 0u;
      const uint32_t xb = k & 1;
      mbar_wait(&x_full[xb], (k >> 1) & 1);
      uint64_t xd[2][3];
      {
        const uint8_t* frames = smem + ConvSmem::xs + xb * (kXFrames * 1024);
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const int xi = p - 4 + j;               // conv1 position p-2 reads x[p-4 .. p-2]
            const float xv = (valid && xi >= 0 && xi < 128) ? frame_sample(frames, in_fmt, (int)(f - f0), r, xi) : 0.f;
            if (MODE == kConvF16) xmax = fmaxf(xmax, fabsf(xv));
            xd[r][j] = pack_dup(xv);
          }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&x_empty[xb]);
      // bf16 / tf32x3: fully unrolled, every conv1 weight an immediate constant-bank operand.  f16x3: NOT unrolled -
      // with the split's conversions the unrolled loop is ~65 KB of code executed by 8 warps next to the epilogue's,
      // and a third of the producers' samples sat in instruction-fetch stalls (ncu: stall_no_inst); the weights are
      // fetched through uniform loads at (c, phalf)-relative offsets instead
#pragma unroll (ConvSmem::kProdUnroll)
      for (int c = 0; c < kChunks; ++c, ++it) {
        // compute the chunk into registers first: nothing here depends on the stage being free
        uint4 o[ConvSmem::kImgs * kGroups];
        if (MODE == kConvBF16) {
#pragma unroll
          for (int hc = 0; hc < 2; ++hc) {
            const unsigned long long* w = &w1c.v[(c * 2 + hc) * 16];
            o[2 * hc] = conv1_item(xd[0][0], xd[0][1], xd[0][2], w, m);
            o[2 * hc + 1] = conv1_item(xd[1][0], xd[1][1], xd[1][2], w, m);
          }
        } else if (MODE == kConvF16) {
          // this thread's channel block only: o[0], o[1] = hi of (block, I row), (block, Q row); o[2], o[3] = lo
          const unsigned long long* w = &w1c.v[(c * 2 + phalf) * 16];
          conv1_item_f16(xd[0][0], xd[0][1], xd[0][2], w, m, o[0], o[2]);
          conv1_item_f16(xd[1][0], xd[1][1], xd[1][2], w, m, o[1], o[3]);
        } else {
          const unsigned long long* w = &w1c.v[c * 16];       // chunk c = channels 8c .. 8c+7
#pragma unroll
          for (int qd = 0; qd < 2; ++qd) {
            conv1_quad(xd[0][0], xd[0][1], xd[0][2], w, qd, m, o[2 * qd], o[kGroups + 2 * qd]);
            conv1_quad(xd[1][0], xd[1][1], xd[1][2], w, qd, m, o[2 * qd + 1], o[kGroups + 2 * qd + 1]);
          }
        }
        // publish the PREVIOUS chunk now: its stores were issued a whole chunk of math ago, so the
        // generic->async proxy fence no longer waits on them
        if (it > 0) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[(it - 1) % kStages]);
        }
        const uint32_t s = it % kStages, ph = (it / kStages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* arow = smem + ConvSmem::a + s * kASlot + row * 16;
        if (MODE == kConvF16) {
          // groups 2 phalf, 2 phalf + 1 of the hi image and of the lo image
          uint8_t* ab = arow + 2 * phalf * kALbo;
          *reinterpret_cast<uint4*>(ab) = o[0];
          *reinterpret_cast<uint4*>(ab + kALbo) = o[1];
          *reinterpret_cast<uint4*>(ab + kAImg) = o[2];
          *reinterpret_cast<uint4*>(ab + kAImg + kALbo) = o[3];
        } else {
#pragma unroll
          for (int g = 0; g < ConvSmem::kImgs * kGroups; ++g) *reinterpret_cast<uint4*>(arow + g * kALbo) = o[g];
        }
      }
    }
    if (it > 0) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[(it - 1) % kStages]);
    }
    if (MODE == kConvF16 && !(xmax <= x_limit)) atomicOr(flags, 1u);
  }

  // ---- teardown (the pair leaves together: the leader's MMAs read the peer's shared memory)
  __syncwarp();
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair<512>(tmem);
}

// ------------------------------------------------------------------------------------------
// dense1: h = relu(act W3 + b3).  A = act [frames][10560] bf16 (K-major), B = W3^T [256][10560].
constexpr int kDM = 256;                  // frames per CTA tile (two M = 128 accumulators)
constexpr int kDK = 64;                   // K elements per stage (128 B swizzled rows)
constexpr int kDStages = 3;
constexpr int kDKBlocks = kVtFlat / kDK;  // 165
constexpr int kDTileBytes = kDM * 128;    // 32 KB (A and B tiles are the same size)
constexpr int kDenseThreads = 10 * 32;    // TMA, MMA, 8 epilogue warps
static_assert(kVtFlat % kDK == 0, "K must tile");

struct DenseSmem {
  static constexpr int a = 0;
  static constexpr int b = a + kDStages * kDTileBytes;
  static constexpr int b3 = b + kDStages * kDTileBytes;
  static constexpr int bars = b3 + 1024;
  static constexpr int nbars = 2 * kDStages + 2;
  static constexpr int tmem_slot = bars + nbars * 8;
  static constexpr int total = tmem_slot + 16 + 1024;   // + slack for the 1024 B alignment
};
static_assert(DenseSmem::total <= 232448, "dense kernel shared memory exceeds 227 KB");

// Dense(C) weights as a kernel parameter (constant bank): in the fused epilogue every W4 element is an
// immediate-offset constant operand of an FFMA, no shared-memory or global loads.  [256][C] fp32 + bias.
template <int C>
struct HeadW {
  float w[256 * C];
  float b[C];
};

// The epilogue is the rest of the network: thread = frame, so +b3, ReLU, Dense(C), softmax, argmax and the
// class histogram need no cross-lane traffic and h never goes to HBM (hbuf != NULL keeps a copy for debugging).
template <int C>
__global__ void __launch_bounds__(kDenseThreads, 1)
vt_dense_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ HeadW<C> hw, const float* __restrict__ b3g, float* __restrict__ hbuf,
                     long long n, int num_tiles, float* __restrict__ probs, float* __restrict__ logits_out,
                     int* __restrict__ cls, unsigned long long* __restrict__ hist) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DenseSmem::bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + kDStages;
  uint64_t* tmem_full = bars + 2 * kDStages;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + DenseSmem::tmem_slot);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;

  for (int i = tid; i < 256; i += kDenseThreads) reinterpret_cast<float*>(smem + DenseSmem::b3)[i] = b3g[i];
  if (tid == 0) {
    for (int s = 0; s < kDStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 8);
    fence_barrier_init();
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < kDKBlocks; ++kb, ++it) {
        const uint32_t s = it % kDStages, ph = (it / kDStages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[s], 2 * kDTileBytes);
          tma_load_2d(smem + DenseSmem::a + s * kDTileBytes, &map_a, kb * kDK, tile * kDM, &full[s]);
          tma_load_2d(smem + DenseSmem::b + s * kDTileBytes, &map_b, kb * kDK, 0, &full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, 256);
    const uint32_t a_base = smem_u32(smem + DenseSmem::a), b_base = smem_u32(smem + DenseSmem::b);
    constexpr uint32_t hi = smem_desc_hi(1024, 2);
    uint32_t it = 0, k = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++k) {
      mbar_wait(tmem_empty, (k & 1) ^ 1);
      for (int kb = 0; kb < kDKBlocks; ++kb, ++it) {
        const uint32_t s = it % kDStages, ph = (it / kDStages) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t a_lo = smem_desc_lo(a_base + s * kDTileBytes, 16);
          const uint32_t b_lo = smem_desc_lo(b_base + s * kDTileBytes, 16);
#pragma unroll
          for (int ks = 0; ks < kDK / 16; ++ks) {
#pragma unroll
            for (int m = 0; m < 2; ++m)
              mma_f16_ss(tmem + m * 256, desc64(a_lo + ((m * 16384 + ks * 32) >> 4), hi),
                          desc64(b_lo + ((ks * 32) >> 4), hi), idesc, (kb | ks) != 0);
          }
          mma_commit(&empty[s]);
          if (kb == kDKBlocks - 1) mma_commit(tmem_full);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3, m = (warp - 2) >> 2;
    const float* b3s = reinterpret_cast<const float*>(smem + DenseSmem::b3);
    uint32_t k = 0;
    unsigned cnt = 0;                             // lane c counts class c
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++k) {
      mbar_wait(tmem_full, k & 1);
      tc_fence_after_sync();
      const long long row = (long long)tile * kDM + m * 128 + q * 32 + lane;
      float z[C];
#pragma unroll
      for (int c = 0; c < C; ++c) z[c] = hw.b[c];
#pragma unroll
      for (int cc = 0; cc < 16; cc += 2) {
        uint32_t v0[16], v1[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + m * 256 + cc * 16, v0);
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + m * 256 + cc * 16 + 16, v1);
        tmem_ld_wait();
        float hv[32];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          hv[e] = fmaxf(__uint_as_float(v0[e]) + b3s[cc * 16 + e], 0.f);
          hv[16 + e] = fmaxf(__uint_as_float(v1[e]) + b3s[cc * 16 + 16 + e], 0.f);
        }
#pragma unroll
        for (int e = 0; e < 32; ++e)
#pragma unroll
          for (int c = 0; c < C; ++c) z[c] = fmaf(hv[e], hw.w[(cc * 16 + e) * C + c], z[c]);
        if (hbuf != nullptr && row < n) {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<float4*>(hbuf + row * 256 + cc * 16 + e) = make_float4(hv[e], hv[e + 1], hv[e + 2], hv[e + 3]);
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
      // softmax / argmax of this thread's frame
      float mx = z[0];
      int best = 0;
#pragma unroll
      for (int c = 1; c < C; ++c) if (z[c] > mx) { mx = z[c]; best = c; }
      if (row >= n) best = -1;
      if (row < n) {
        if (logits_out) {
#pragma unroll
          for (int c = 0; c < C; ++c) logits_out[row * C + c] = z[c];
        }
        if (probs) {
          float e[C], sum = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) { e[c] = expf(z[c] - mx); sum += e[c]; }
          const float inv = 1.0f / sum;
#pragma unroll
          for (int c = 0; c < C; ++c) probs[row * C + c] = e[c] * inv;
        }
        if (cls) cls[row] = best;
      }
      if (hist) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const unsigned votes = __popc(__ballot_sync(0xffffffffu, best == c));
          if (lane == c) cnt += votes;
        }
      }
    }
    if (hist && lane < C && cnt) atomicAdd(hist + lane, (unsigned long long)cnt);
  }

  __syncwarp();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------
// dense1 in 3xTF32: h = relu(act W3 + b3) with act = act_hi + act_lo, W3 = W3_hi + W3_lo (fp32 words,
// tf32 hi/lo split), three kind::tf32 MMAs per K step.  128 frames x 256 outputs per tile, K blocks
// of 32 values (128-B swizzled rows); 96 KB per stage (A hi/lo 16 KB each, B hi/lo 32 KB each), two
// stages.  CTAs run in clusters of two that walk the K blocks in lockstep on different frame tiles: each loads
// HALF of every W3 block and multicasts it into both CTAs' stages, which halves the L2 traffic of re-streaming
// the 21.6 MB of W3 hi/lo per tile (the kernel was L2-bound on exactly that).
//
// K = 10,560 would be a chain of 3,960 truncating accumulates (see ConvCfg), so the tensor core only
// ever sums a RUN of two K blocks: every run starts a fresh accumulator (24 MMAs) in one of two TMEM
// buffers, and eight epilogue warps fold the finished buffer into fp32 master sums held in registers
// (128 per thread, round-to-nearest FADD) while the MMAs of the next run fill the other buffer.  Two
// blocks, not one: tcgen05.ld moves 64 B/clk per SM, so folding a 128 x 256 fp32 buffer takes 2,048
// cycles - more than the 1,536 MMA cycles of one block, less than the 3,072 of two.
constexpr int kTM = 128;
constexpr int kTK = 32;
constexpr int kTStages = 2;
constexpr int kTKBlocks = kVtFlat / kTK;           // 330
constexpr int kTABytes = kTM * 128;                // 16 KB
constexpr int kTBBytes = 256 * 128;                // 32 KB
constexpr int kTStageBytes = 2 * kTABytes + 2 * kTBBytes;
constexpr int kTRunBlocks = 2;                     // K blocks the tensor core sums before the fold (24 MMAs)
constexpr int kTRuns = kTKBlocks / kTRunBlocks;    // 165
static_assert(kTKBlocks % kTRunBlocks == 0, "runs must tile K");
constexpr int kTEpiWarps = 8;                      // (TMEM lane quarter) x (column half)
constexpr int kDenseT32Threads = (2 + kTEpiWarps) * 32;   // TMA, MMA, 8 epilogue warps
static_assert(kVtFlat % kTK == 0, "K must tile");

struct DenseT32Smem {
  static constexpr int stages = 0;
  static constexpr int b3 = kTStages * kTStageBytes;
  static constexpr int bars = b3 + 1024;
  static constexpr int nbars = 2 * kTStages + 4;
  static constexpr int tmem_slot = bars + nbars * 8;
  static constexpr int total = tmem_slot + 16 + 1024;
};
static_assert(DenseT32Smem::total <= 232448, "tf32 dense kernel shared memory exceeds 227 KB");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kDenseT32Threads, 1)
vt_dense_tf32x3_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                       const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                       const float* __restrict__ b3g, float* __restrict__ hbuf, long long n, int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DenseT32Smem::bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + kTStages;
  uint64_t* acc_full = bars + 2 * kTStages;     // [2] one K block summed into this TMEM buffer
  uint64_t* acc_empty = acc_full + 2;           // [2] the epilogue warps have folded it into their registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + DenseT32Smem::tmem_slot);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  // the pair walks tiles 2 i and 2 i + 1 in lockstep (a trailing odd tile is all out-of-range rows: zero-filled
  // loads, no stores)
  const int tile_first = 2 * (int)cluster_id_x() + (int)rank, tile_step = 2 * (int)cluster_count_x();
  const int pair_iters = (num_tiles + 1) / 2;       // iterations of pair p: tiles 2 p, 2 p + 1

  for (int i = tid; i < 256; i += kDenseT32Threads) reinterpret_cast<float*>(smem + DenseT32Smem::b3)[i] = b3g[i];
  if (tid == 0) {
    for (int s = 0; s < kTStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2);                   // this CTA's MMAs and the peer's (its multicast writes land here too)
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kTEpiWarps);
    }
    fence_barrier_init();
    prefetch_tensormap(&map_ah);
    prefetch_tensormap(&map_al);
    prefetch_tensormap(&map_bh);
    prefetch_tensormap(&map_bl);
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();                            // the peer's barriers exist before anything is multicast at them
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    uint32_t it = 0;
    for (int tile = tile_first, pit = (int)cluster_id_x(); pit < pair_iters; tile += tile_step, pit += (int)cluster_count_x()) {
      for (int kb = 0; kb < kTKBlocks; ++kb, ++it) {
        const uint32_t s = it % kTStages, ph = (it / kTStages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one()) {
          uint8_t* st = smem + s * kTStageBytes;
          mbar_arrive_expect_tx(&full[s], kTStageBytes);
          tma_load_2d(st, &map_ah, kb * kTK, tile * kTM, &full[s]);
          tma_load_2d(st + kTABytes, &map_al, kb * kTK, tile * kTM, &full[s]);
          // my half of the W3 block (128 of its 256 rows), to both CTAs
          tma_load_2d_multicast(st + 2 * kTABytes + rank * (kTBBytes / 2), &map_bh, kb * kTK, (int)rank * 128, &full[s], 3);
          tma_load_2d_multicast(st + 2 * kTABytes + kTBBytes + rank * (kTBBytes / 2), &map_bl, kb * kTK, (int)rank * 128, &full[s], 3);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_tf32(128, 256);
    const uint32_t base = smem_u32(smem);
    constexpr uint32_t hi = smem_desc_hi(1024, 2);
    uint32_t it = 0, run = 0;
    for (int tile = tile_first, pit = (int)cluster_id_x(); pit < pair_iters; tile += tile_step, pit += (int)cluster_count_x()) {
      for (int r = 0; r < kTRuns; ++r, ++run) {
        const uint32_t buf = run & 1;
        mbar_wait(&acc_empty[buf], ((run >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        for (int kk = 0; kk < kTRunBlocks; ++kk, ++it) {
          const uint32_t s = it % kTStages, ph = (it / kTStages) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t st = base + s * kTStageBytes;
            const uint32_t ah = smem_desc_lo(st, 16), al = smem_desc_lo(st + kTABytes, 16);
            const uint32_t bh = smem_desc_lo(st + 2 * kTABytes, 16), bl = smem_desc_lo(st + 2 * kTABytes + kTBBytes, 16);
#pragma unroll
            for (int ks = 0; ks < kTK / 8; ++ks) {
              const uint32_t o = (ks * 32) >> 4;
              mma_tf32_ss(tmem + buf * 256, desc64(al + o, hi), desc64(bh + o, hi), idesc, (kk | ks) != 0);
              mma_tf32_ss(tmem + buf * 256, desc64(ah + o, hi), desc64(bl + o, hi), idesc, 1);
              mma_tf32_ss(tmem + buf * 256, desc64(ah + o, hi), desc64(bh + o, hi), idesc, 1);
            }
            mma_commit_multicast(&empty[s], 3);
            if (kk == kTRunBlocks - 1) mma_commit(&acc_full[buf]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;          // TMEM lane quarter, column half
    const float* b3s = reinterpret_cast<const float*>(smem + DenseT32Smem::b3) + half * 128;
    uint32_t it = 0;
    for (int tile = tile_first, pit = (int)cluster_id_x(); pit < pair_iters; tile += tile_step, pit += (int)cluster_count_x()) {
      float acc[128];
#pragma unroll
      for (int i = 0; i < 128; ++i) acc[i] = 0.f;
#pragma unroll 1
      for (int r = 0; r < kTRuns; ++r, ++it) {
        const uint32_t buf = it & 1;
        mbar_wait(&acc_full[buf], (it >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + buf * 256 + half * 128;
#pragma unroll
        for (int g = 0; g < 4; g += 2) {
          uint32_t v0[32], v1[32];
          tmem_ld32(tb + g * 32, v0);
          tmem_ld32(tb + g * 32 + 32, v1);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            acc[g * 32 + e] += __uint_as_float(v0[e]);
            acc[g * 32 + 32 + e] += __uint_as_float(v1[e]);
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      }
      const long long row = (long long)tile * kTM + q * 32 + lane;
      if (row < n) {
        float* dst = hbuf + row * 256 + half * 128;
#pragma unroll
        for (int e = 0; e < 128; e += 4) {
          float4 o;
          o.x = fmaxf(acc[e] + b3s[e], 0.f);
          o.y = fmaxf(acc[e + 1] + b3s[e + 1], 0.f);
          o.z = fmaxf(acc[e + 2] + b3s[e + 2], 0.f);
          o.w = fmaxf(acc[e + 3] + b3s[e + 3], 0.f);
          *reinterpret_cast<float4*>(dst + e) = o;
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();                            // no multicast may target a CTA that has already exited
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------
// dense1 in fp16 hi/lo split (MDC_MODE_F16X3): act = act_hi + 2^-11 act_lo, W3 = W3_hi + 2^-11 W3_lo (fp16 words),
// three kind::f16 MMAs per K step at the full 16-bit rate.  128 frames x 256 outputs per CTA, K blocks of 64 values.
//
// CTA pairs (cluster of 2, cta_group::2, M = 256): each CTA owns a frame tile - its A stages, its accumulators, its
// epilogue - and HALF of every W3 block (128 of the 256 output rows, hi and lo); the leader issues one M = 256 x
// N = 256 MMA per product for the pair.  Against one-CTA MMAs with multicast W3 halves (the 3xTF32 kernel below) a
// K block costs each SM 64 KB of TMA writes and 96 KB of operand reads instead of 96 + 144 KB: the one-CTA form was
// shared-memory-bound (1,875 cycles of traffic per block against 1,536 of math, tensor pipe 64 % active).
//
// TMEM: columns 0..255 take the hi*hi products, columns 256..511 the two cross terms (which carry a factor 2^11).
// The cross accumulator runs over the whole K range - it is 2^-11 of the result, its truncation does not matter -
// while the hi*hi accumulator is restarted every kHRunBlocks K blocks (44 truncating adds) and folded into fp32
// master sums in registers by the eight epilogue warps.  There is no room for a second hi*hi buffer, so within a
// K block the MMA warp issues the 8 cross MMAs first: at a run boundary the fold (2,048 cycles of tcgen05.ld) has
// those 1,024 MMA cycles plus the wait before the next run's first hi*hi MMA overwrites the accumulator - about
// 6 % of a run.
constexpr int kHK = 64;                            // fp16 values per K block (128-B swizzled rows)
constexpr int kHKBlocks = kVtFlat / kHK;           // 165
constexpr int kHRunBlocks = 11;                    // K blocks per hi*hi run: 44 MMAs
constexpr int kHRuns = kHKBlocks / kHRunBlocks;    // 15
constexpr int kHStages = 3;
constexpr int kHTile = 128 * 128;                  // 128 rows x 128 B: one A (hi or lo) or half-B (hi or lo) tile
constexpr int kHStageBytes = 4 * kHTile;           // A hi, A lo, B-half hi, B-half lo = 64 KB
static_assert(kVtFlat % kHK == 0 && kHKBlocks % kHRunBlocks == 0, "K blocks / runs must tile K");

struct DenseF16Smem {
  static constexpr int stages = 0;
  static constexpr int b3 = kHStages * kHStageBytes;
  static constexpr int zx = b3 + 1024;                 // fused head: partial logits of the upper column half, 2 tiles
  static constexpr int bars = zx + 2 * kTM * kMaxClasses * 4;
  // full[S], empty[S], hh_full, hh_empty, cross_empty
  static constexpr int nbars = 2 * kHStages + 3;
  static constexpr int tmem_slot = bars + nbars * 8;
  static constexpr int total = tmem_slot + 16 + 1024;   // + slack for the 1024 B alignment
};
static_assert(DenseF16Smem::total <= 232448, "f16x3 dense kernel shared memory exceeds 227 KB");

// softmax / argmax / outputs of one frame from its logits (shared by the fused epilogue and vt_head_ordered_kernel so
// that both produce the same bits); returns the class, -1 for a row past the end
template <int C>
__device__ __forceinline__ int head_finish(const float (&z)[C], long long row, long long n, float* __restrict__ probs,
                                           float* __restrict__ logits_out, int* __restrict__ cls) {
  float mx = z[0];
  int best = 0;
#pragma unroll
  for (int c = 1; c < C; ++c) if (z[c] > mx) { mx = z[c]; best = c; }
  if (row >= n) return -1;
  if (logits_out) {
#pragma unroll
    for (int c = 0; c < C; ++c) logits_out[row * C + c] = z[c];
  }
  if (probs) {
    float e[C], sum = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { e[c] = expf(z[c] - mx); sum += e[c]; }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int c = 0; c < C; ++c) probs[row * C + c] = e[c] * inv;
  }
  if (cls) cls[row] = best;
  return best;
}

// Partial logits of one frame over this thread's 128 dense1 outputs (column half H): +b3, ReLU, Dense(C) with every
// W4 element an immediate-offset constant-bank operand (H is a template parameter for exactly that reason).
template <int C, int H>
__device__ __forceinline__ void head_partial(const float (&acc)[128], const float* b3s, const HeadW<C>& hw, float (&z)[C]) {
#pragma unroll
  for (int e = 0; e < 128; ++e) {
    const float hv = fmaxf(acc[e] + b3s[e], 0.f);
#pragma unroll
    for (int c = 0; c < C; ++c) z[c] = fmaf(hv, hw.w[(H * 128 + e) * C + c], z[c]);
  }
}

// C > 0: the rest of the network is fused into the epilogue (Dense(C), softmax, argmax, histogram; h never goes to HBM
// and the separate head launch is gone).  C == 0: h is stored to hbuf and vt_head_kernel follows (any class count).
template <int C>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kDenseT32Threads, 1)
vt_dense_f16x3_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                      const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                      const __grid_constant__ HeadW<(C > 0 ? C : 1)> hw, const float* __restrict__ b3g,
                      float* __restrict__ hbuf, long long n, int num_tiles, int split_pairs, float* __restrict__ probs,
                      float* __restrict__ logits_out, int* __restrict__ cls, unsigned long long* __restrict__ hist) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DenseF16Smem::bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + kHStages;
  uint64_t* hh_full = bars + 2 * kHStages;      // a run of hi*hi MMAs (and, at the last run, the cross sums) is complete
  uint64_t* hh_empty = hh_full + 1;             // (leader's) both CTAs' epilogue warps have folded it into their registers
  uint64_t* cross_empty = hh_full + 2;          // (leader's) both CTAs' epilogue warps have read the tile's cross sums
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + DenseF16Smem::tmem_slot);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const uint32_t rank = cluster_ctarank();      // 0 = leader (issues the pair's MMAs)
  // Work items of this CTA pair.  A pair-tile p is tiles 2 p and 2 p + 1, one per CTA, walked in lockstep (a trailing
  // odd tile is all out-of-range rows: zero-filled loads, no stores).  The full pair-tiles c, c + G, ... below
  // `full_pairs` come first; then, when the last round would leave more than half of the pairs idle, ONE OUTPUT HALF
  // (128 of the 256 dense1 outputs, an N = 128 MMA: half the cycles) of a left-over pair-tile - split_pairs of them,
  // two CTA pairs each.  Every output is still summed in exactly the order of an unsplit tile, so results do not
  // depend on where a frame sits in the batch; the halves' h goes to hbuf and vt_head_ordered_kernel finishes them.
  const int pair_iters = (num_tiles + 1) / 2;
  const int full_pairs = pair_iters - split_pairs;
  const int cid = (int)cluster_id_x(), G = (int)cluster_count_x();
  const int n_full = full_pairs > cid ? (full_pairs - cid + G - 1) / G : 0;
  const int n_items = n_full + (cid < 2 * split_pairs ? 1 : 0);
  auto item = [&](int i, int& tile, int& nh) {       // nh: -1 = all 256 outputs, 0 / 1 = outputs [128 nh, 128 nh + 128)
    if (i < n_full) {
      tile = 2 * (cid + i * G) + (int)rank;
      nh = -1;
    } else {
      tile = 2 * (full_pairs + (cid >> 1)) + (int)rank;
      nh = cid & 1;
    }
  };

  for (int i = tid; i < 256; i += kDenseT32Threads) reinterpret_cast<float*>(smem + DenseF16Smem::b3)[i] = b3g[i];
  if (tid == 0) {
    for (int s = 0; s < kHStages; ++s) {
      mbar_init(&full[s], rank == 0 ? 2 : 1);    // own TMA (+ on the leader: the peer's relay once ITS stage is full)
      mbar_init(&empty[s], 1);
    }
    mbar_init(hh_full, 1);
    mbar_init(hh_empty, 2 * kTEpiWarps);
    mbar_init(cross_empty, 2 * kTEpiWarps);
    fence_barrier_init();
    prefetch_tensormap(&map_ah);
    prefetch_tensormap(&map_al);
    prefetch_tensormap(&map_bh);
    prefetch_tensormap(&map_bl);
  }
  if (warp == 0) tmem_alloc_pair<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= TMA: this CTA's frame tile (hi, lo) and its half of the W3 block (hi, lo)
    uint32_t it = 0;
    for (int i = 0; i < n_items; ++i) {
      int tile, nh;
      item(i, tile, nh);
      // this CTA's rows of W3^T: its half of the 256 outputs, or its quarter (64 rows) of an output half - the box
      // stays 128 rows (rows past 255 are zero-filled), the N = 128 MMA reads the first 64 of them
      const int brow = nh < 0 ? (int)rank * 128 : nh * 128 + (int)rank * 64;
      for (int kb = 0; kb < kHKBlocks; ++kb, ++it) {
        const uint32_t s = it % kHStages, ph = (it / kHStages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one()) {
          uint8_t* st = smem + s * kHStageBytes;
          mbar_arrive_expect_tx(&full[s], kHStageBytes);
          tma_load_2d(st, &map_ah, kb * kHK, tile * kTM, &full[s]);
          tma_load_2d(st + kHTile, &map_al, kb * kHK, tile * kTM, &full[s]);
          tma_load_2d(st + 2 * kHTile, &map_bh, kb * kHK, brow, &full[s]);
          tma_load_2d(st + 3 * kHTile, &map_bl, kb * kHK, brow, &full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    uint32_t it = 0, run = 0, tcount = 0;
    if (rank == 0) {
      // ================= MMA issuer (whole warp loops, one elected lane issues)
      const uint32_t idesc_full = make_idesc_f16(256, 256), idesc_half = make_idesc_f16(256, 128);
      const uint32_t base = smem_u32(smem);
      constexpr uint32_t hi = smem_desc_hi(1024, 2);
      for (int i = 0; i < n_items; ++i, ++tcount) {
        int tile, nh;
        item(i, tile, nh);
        const uint32_t idesc = nh < 0 ? idesc_full : idesc_half;
        mbar_wait(cross_empty, (tcount & 1) ^ 1);   // the previous tiles' cross sums have been read (both CTAs)
        tc_fence_after_sync();
        for (int r = 0; r < kHRuns; ++r, ++run) {
          for (int kk = 0; kk < kHRunBlocks; ++kk, ++it) {
            const uint32_t s = it % kHStages, ph = (it / kHStages) & 1;
            mbar_wait(&full[s], ph);                // own TMA and the peer's relay
            tc_fence_after_sync();
            const uint32_t st = base + s * kHStageBytes;
            const uint32_t ah = smem_desc_lo(st, 16), al = smem_desc_lo(st + kHTile, 16);
            const uint32_t bh = smem_desc_lo(st + 2 * kHTile, 16), bl = smem_desc_lo(st + 3 * kHTile, 16);
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < kHK / 16; ++ks) {
                const uint32_t o = (ks * 32) >> 4;
                mma_f16_ss_pair(tmem + 256, desc64(al + o, hi), desc64(bh + o, hi), idesc, (r | kk | ks) != 0);
                mma_f16_ss_pair(tmem + 256, desc64(ah + o, hi), desc64(bl + o, hi), idesc, 1);
              }
            }
            __syncwarp();
            if (kk == 0) {                          // the previous run has been folded: its accumulator may be restarted
              mbar_wait(hh_empty, (run & 1) ^ 1);
              tc_fence_after_sync();
            }
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < kHK / 16; ++ks) {
                const uint32_t o = (ks * 32) >> 4;
                mma_f16_ss_pair(tmem, desc64(ah + o, hi), desc64(bh + o, hi), idesc, (kk | ks) != 0);
              }
              mma_commit_pair(&empty[s]);
              if (kk == kHRunBlocks - 1) mma_commit_pair(hh_full);
            }
            __syncwarp();
          }
        }
      }
    } else {
      // ================= relay (peer CTA): forwards "my stage is full" to the leader's barrier
      for (int i = 0; i < n_items; ++i) {
        for (int kb = 0; kb < kHKBlocks; ++kb, ++it) {
          const uint32_t s = it % kHStages, ph = (it / kHStages) & 1;
          mbar_wait(&full[s], ph);
          if (elect_one()) mbar_arrive_remote(&full[s], 0);
          __syncwarp();
        }
      }
    }
  } else {
    // ================= epilogue: fold the hi*hi runs into fp32 master sums, add the cross sums, +b3, ReLU -> h
    const int q = warp & 3, half = (warp - 2) >> 2;          // TMEM lane quarter, column half
    uint32_t run = 0, titer = 0;
    unsigned cnt = 0;
    auto release = [&](uint64_t* bar) {           // one arrival per epilogue warp on the LEADER's barrier
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(bar);
        else mbar_arrive_remote(bar, 0);
      }
    };
    for (int i = 0; i < n_items; ++i) {
      int tile, nh;
      item(i, tile, nh);
      // a thread folds 128 columns of a full tile, 64 of an output half (ng = column groups of 64)
      const int ng = nh < 0 ? 2 : 1;
      const int col0 = nh < 0 ? half * 128 : half * 64;                  // first accumulator column of this thread
      const float* b3s = reinterpret_cast<const float*>(smem + DenseF16Smem::b3) + (nh < 0 ? 0 : nh * 128) + col0;
      float acc[128];
#pragma unroll
      for (int i = 0; i < 128; ++i) acc[i] = 0.f;
      const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + col0;
#pragma unroll 1
      for (int r = 0; r < kHRuns; ++r, ++run) {
        mbar_wait(hh_full, run & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int g = 0; g < 4; g += 2) {
          if (g < 2 * ng) {
            uint32_t v0[32], v1[32];
            tmem_ld32(tb + g * 32, v0);
            tmem_ld32(tb + g * 32 + 32, v1);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              acc[g * 32 + e] += __uint_as_float(v0[e]);
              acc[g * 32 + 32 + e] += __uint_as_float(v1[e]);
            }
          }
        }
        release(hh_empty);
      }
      // the last run's commit also covers every cross MMA of the tile
#pragma unroll
      for (int g = 0; g < 4; g += 2) {
        if (g < 2 * ng) {
          uint32_t v0[32], v1[32];
          tmem_ld32(tb + 256 + g * 32, v0);
          tmem_ld32(tb + 256 + g * 32 + 32, v1);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            acc[g * 32 + e] = fmaf(__uint_as_float(v0[e]), 1.f / 2048.f, acc[g * 32 + e]);
            acc[g * 32 + 32 + e] = fmaf(__uint_as_float(v1[e]), 1.f / 2048.f, acc[g * 32 + 32 + e]);
          }
        }
      }
      release(cross_empty);
      const long long row = (long long)tile * kTM + q * 32 + lane;
      if (nh >= 0) {
        // output half: this thread's 64 values of h = relu(. + b3); Dense(C) + softmax follow in vt_head_ordered_kernel
        if (row < n) {
          float* dst = hbuf + row * 256 + nh * 128 + col0;
#pragma unroll
          for (int e = 0; e < 64; e += 4) {
            float4 o;
            o.x = fmaxf(acc[e] + b3s[e], 0.f);
            o.y = fmaxf(acc[e + 1] + b3s[e + 1], 0.f);
            o.z = fmaxf(acc[e + 2] + b3s[e + 2], 0.f);
            o.w = fmaxf(acc[e + 3] + b3s[e + 3], 0.f);
            *reinterpret_cast<float4*>(dst + e) = o;
          }
        }
        continue;
      }
      if constexpr (C == 0) {
        if (row < n) {
          float* dst = hbuf + row * 256 + half * 128;
#pragma unroll
          for (int e = 0; e < 128; e += 4) {
            float4 o;
            o.x = fmaxf(acc[e] + b3s[e], 0.f);
            o.y = fmaxf(acc[e + 1] + b3s[e + 1], 0.f);
            o.z = fmaxf(acc[e + 2] + b3s[e + 2], 0.f);
            o.w = fmaxf(acc[e + 3] + b3s[e + 3], 0.f);
            *reinterpret_cast<float4*>(dst + e) = o;
          }
        }
      } else {
        // the rest of the network, under the next tile's MMAs (the accumulators were released above): each thread has
        // 128 of its frame's 256 dense1 outputs; the upper half hands its partial logits over through shared memory
        float z[C];
#pragma unroll
        for (int c = 0; c < C; ++c) z[c] = half == 0 ? hw.b[c] : 0.f;
        if (half == 0) head_partial<C, 0>(acc, b3s, hw, z);
        else head_partial<C, 1>(acc, b3s, hw, z);
        float* zx = reinterpret_cast<float*>(smem + DenseF16Smem::zx) + (titer & 1) * (kTM * C) + (q * 32 + lane) * C;
        if (half == 1) {
#pragma unroll
          for (int c = 0; c < C; ++c) zx[c] = z[c];
        }
        // the eight epilogue warps; the buffer alternates per tile, so one barrier per tile orders its reuse as well
        asm volatile("bar.sync 1, %0;" ::"n"(kTEpiWarps * 32) : "memory");
        if (half == 0) {
#pragma unroll
          for (int c = 0; c < C; ++c) z[c] += zx[c];
          const int best = head_finish<C>(z, row, n, probs, logits_out, cls);
          if (hist) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const unsigned votes = __popc(__ballot_sync(0xffffffffu, best == c));
              if (lane == c) cnt += votes;
            }
          }
        }
      }
      ++titer;
    }
    if (C > 0 && hist && half == 0 && lane < C && cnt) atomicAdd(hist + lane, (unsigned long long)cnt);
  }

  // ---- teardown (the pair leaves together: the leader's MMAs read the peer's shared memory)
  __syncwarp();
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair<512>(tmem);
}

// Dense(C) + softmax for the frames whose dense1 outputs came from two output-half items (hbuf holds their h): one
// thread per frame, the SAME order of operations as the fused epilogue above - the lower 128 outputs accumulate from
// the bias, the upper 128 from zero, the two partial logits are added last - so these frames get the same bits as
// frames of unsplit tiles.
template <int C>
__global__ void __launch_bounds__(128)
vt_head_ordered_kernel(const __grid_constant__ HeadW<C> hw, const float* __restrict__ hbuf, long long n,
                       float* __restrict__ probs, float* __restrict__ logits_out, int* __restrict__ cls,
                       unsigned long long* __restrict__ hist) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const float4* h4 = reinterpret_cast<const float4*>(hbuf + (row < n ? row : n - 1) * 256);
  float z0[C], z1[C];
#pragma unroll
  for (int c = 0; c < C; ++c) { z0[c] = hw.b[c]; z1[c] = 0.f; }
#pragma unroll 4
  for (int e4 = 0; e4 < 32; ++e4) {
    const float4 a = h4[e4], b = h4[32 + e4];
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < C; ++c) {
        z0[c] = fmaf(av[j], hw.w[(4 * e4 + j) * C + c], z0[c]);
        z1[c] = fmaf(bv[j], hw.w[(128 + 4 * e4 + j) * C + c], z1[c]);
      }
  }
  float z[C];
#pragma unroll
  for (int c = 0; c < C; ++c) z[c] = z0[c] + z1[c];
  const int best = head_finish<C>(z, row, n, probs, logits_out, cls);
  if (hist) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const unsigned votes = __popc(__ballot_sync(0xffffffffu, best == c));
      if (lane == c && votes) atomicAdd(hist + c, (unsigned long long)votes);
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side
static uint16_t f2bf(float f) {   // round to nearest even, like cvt.rn.bf16.f32
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

static PFN_cuTensorMapEncodeTiled get_encode() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  }
  return fn;
}

// [rows][kVtFlat] K-major matrix of bf16 / fp16 / fp32 elements: box = 128 B of K (64 / 64 / 32 values) x box_rows,
// 128B swizzle, out-of-range rows read as 0
enum { kElemBF16 = 0, kElemF16 = 1, kElemF32 = 2 };
static int make_kmajor_map(CUtensorMap* m, const void* base, uint64_t rows, int elem, uint32_t box_rows) {
  PFN_cuTensorMapEncodeTiled enc = get_encode();
  MDC_REQUIRE(enc != nullptr, MDC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const bool f32 = elem == kElemF32;
  const cuuint64_t dims[2] = {(cuuint64_t)kVtFlat, rows};
  const cuuint64_t strides[1] = {(cuuint64_t)kVtFlat * (f32 ? 4 : 2)};
  const cuuint32_t box[2] = {(cuuint32_t)(f32 ? kTK : kDK), box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                     : (elem == kElemF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  const CUresult r = enc(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MDC_REQUIRE(r == CUDA_SUCCESS, MDC_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return MDC_OK;
}
static_assert(kHK == kDK, "fp16 and bf16 K blocks share the tensor-map box");

static void split_tf32(float v, float& hi, float& lo) {   // hi = what the tensor core reads of v; lo exact
  uint32_t u;
  memcpy(&u, &v, 4);
  u &= 0xFFFFE000u;
  memcpy(&hi, &u, 4);
  lo = v - hi;
}

// v = hi + 2^-11 lo with hi = fp16(v), lo = fp16((v - hi) * 2^11): the same split split_f16x2 makes on the device
static void split_f16(float v, uint16_t& hi, uint16_t& lo) {
  const __half hh = __float2half_rn(v);
  const float hf = __half2float(hh);
  const __half lh = __float2half_rn((v - hf) * 2048.f);
  hi = __half_as_ushort(hh);
  lo = __half_as_ushort(lh);
}

static int conv_mode_of(const mdc_handle_s* h) {
  return h->mode == MDC_MODE_TF32X3 ? kConvTF32 : (h->mode == MDC_MODE_F16X3 ? kConvF16 : kConvBF16);
}

// conv2 image of one mode: [chunk][pair rank][image][tap][group][rows][16 B]; chunk c = conv1 channels
// kCC c .. kCC c + kCC - 1, group g = 2*(channel block) + input row, rank h holds output channels 40h..40h+39
// (Keras (2,3,256,80) = [r][j][ch][o]).  16-bit modes: 8 values per group; tf32x3: 4 fp32 per group.
// bf16: one image of 40 rows; tf32x3: a hi image and a lo image of 40 rows; f16x3: one image of 80 rows, the 40 hi rows
// then the 40 lo rows.  put(value, hi, lo) stores the split value.
template <int MODE, class Elem, class Put>
static void build_w2_image(const float* w2, std::vector<Elem>& img, Put put) {
  using Cfg = ConvCfg<MODE>;
  constexpr int per = 16 / (int)sizeof(Elem);
  constexpr size_t lo_off = MODE == kConvTF32 ? Cfg::kBImg / sizeof(Elem) : (size_t)kBHalf * per;
  img.assign((size_t)Cfg::kChunks * 2 * Cfg::kBSlot / sizeof(Elem), Elem());
  for (int c = 0; c < Cfg::kChunks; ++c)
    for (int hf = 0; hf < 2; ++hf)
      for (int j = 0; j < 3; ++j)
        for (int g = 0; g < kGroups; ++g)
          for (int oo = 0; oo < kBHalf; ++oo)
            for (int e = 0; e < per; ++e) {
              const int r = g & 1, ch = c * Cfg::kCC + (g >> 1) * per + e, o = hf * kBHalf + oo;
              const size_t base = (size_t)(c * 2 + hf) * (Cfg::kBSlot / sizeof(Elem)) +
                                  ((size_t)(j * kGroups + g) * Cfg::kBRows + oo) * per + e;
              put(w2[((size_t)(r * 3 + j) * 256 + ch) * 80 + o], img[base], img.data() + base + lo_off);
            }
}

int pack_vt_bf16(mdc_handle_s* h) {      // the tensor-core modes (MDC_MODE_BF16, MDC_MODE_F16X3, MDC_MODE_TF32X3)
  const int cm = conv_mode_of(h);
  if (int e = pack_vt_small(h)) return e;
  // conv1 image: 32 channel groups x {w0[8], w1[8], w2[8], bias[8]} fp32 (Keras (1,3,1,256) = [tap][ch])
  {
    std::vector<float> img(32 * 32);
    const float* w1 = h->w[MDC_T_CONV1_K].data();
    const float* b1 = h->w[MDC_T_CONV1_B].data();
    float xlim = 3.0e38f;
    for (int g = 0; g < 32; ++g)
      for (int e = 0; e < 8; ++e) {
        const int ch = g * 8 + e;
        img[g * 32 + e] = w1[ch];
        img[g * 32 + 8 + e] = w1[256 + ch];
        img[g * 32 + 16 + e] = w1[512 + ch];
        img[g * 32 + 24 + e] = b1[ch];
        // |x| <= xlim keeps every conv1 activation |x|(|w0|+|w1|+|w2|) + |b| inside the fp16 range
        const float s = fabsf(w1[ch]) + fabsf(w1[256 + ch]) + fabsf(w1[512 + ch]);
        if (s > 0.f) xlim = fminf(xlim, (65000.f - fabsf(b1[ch])) / s);
        if (fabsf(b1[ch]) > 65000.f) xlim = 0.f;
      }
    h->vt_w1_img = img;      // passed by value as a kernel parameter
    h->vt_xlimit = xlim;
  }
  if (cm == kConvF16) {
    // fp16 operands: every weight must fit the fp16 range (activations are checked by the kernels)
    for (int t : {MDC_T_CONV2_K, MDC_T_DENSE1_K})
      for (float v : h->w[t])
        MDC_REQUIRE(fabsf(v) <= 65000.f, MDC_ERR_UNSUPPORTED,
                    "weights tensor %d holds %g, outside the fp16 range of MDC_MODE_F16X3 (use MDC_MODE_TF32X3)", t, v);
    if (int e = h->vt_flags.reserve(256)) return e;
    MDC_CUDA(cudaMemset(h->vt_flags.ptr, 0, 256));
  }
  const float* w2 = h->w[MDC_T_CONV2_K].data();
  if (cm == kConvBF16) {
    std::vector<uint16_t> img;
    build_w2_image<kConvBF16>(w2, img, [](float v, uint16_t& hi, uint16_t*) { hi = f2bf(v); });
    if (int e = h->vt_w2_bf16.reserve(img.size() * 2)) return e;
    MDC_CUDA(cudaMemcpy(h->vt_w2_bf16.ptr, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  } else if (cm == kConvF16) {
    std::vector<uint16_t> img;
    build_w2_image<kConvF16>(w2, img, [](float v, uint16_t& hi, uint16_t* lo) { split_f16(v, hi, *lo); });
    if (int e = h->vt_w2_bf16.reserve(img.size() * 2)) return e;
    MDC_CUDA(cudaMemcpy(h->vt_w2_bf16.ptr, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  } else {
    std::vector<float> img;
    build_w2_image<kConvTF32>(w2, img, [](float v, float& hi, float* lo) { split_tf32(v, hi, *lo); });
    if (int e = h->vt_w2_bf16.reserve(img.size() * 4)) return e;
    MDC_CUDA(cudaMemcpy(h->vt_w2_bf16.ptr, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
  }
  // dense1 image: W3^T [256][10560] in this library's activation order (pos*80 + ch);
  // bf16: one bf16 matrix; f16x3: fp16 hi matrix then the lo matrix; tf32x3: fp32 hi matrix then the lo matrix
  {
    std::vector<float> w3p;
    vt_permute_w3(h, w3p);
    const size_t cnt = (size_t)kVtH * kVtFlat;
    if (!h->tmap_w3) h->tmap_w3 = aligned_alloc(64, 2 * sizeof(CUtensorMap));
    CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(h->tmap_w3);
    if (cm == kConvBF16) {
      std::vector<uint16_t> img(cnt);
      for (int kx = 0; kx < kVtFlat; ++kx)
        for (int o = 0; o < kVtH; ++o) img[(size_t)o * kVtFlat + kx] = f2bf(w3p[(size_t)kx * kVtH + o]);
      if (int e = h->vt_w3_bf16.reserve(cnt * 2)) return e;
      MDC_CUDA(cudaMemcpy(h->vt_w3_bf16.ptr, img.data(), cnt * 2, cudaMemcpyHostToDevice));
      if (int e = make_kmajor_map(&maps[0], h->vt_w3_bf16.ptr, kVtH, kElemBF16, kDM)) return e;
    } else if (cm == kConvF16) {
      std::vector<uint16_t> img(2 * cnt);
      for (int kx = 0; kx < kVtFlat; ++kx)
        for (int o = 0; o < kVtH; ++o)
          split_f16(w3p[(size_t)kx * kVtH + o], img[(size_t)o * kVtFlat + kx], img[cnt + (size_t)o * kVtFlat + kx]);
      if (int e = h->vt_w3_bf16.reserve(2 * cnt * 2)) return e;
      MDC_CUDA(cudaMemcpy(h->vt_w3_bf16.ptr, img.data(), 2 * cnt * 2, cudaMemcpyHostToDevice));
      const uint16_t* base = reinterpret_cast<const uint16_t*>(h->vt_w3_bf16.ptr);
      if (int e = make_kmajor_map(&maps[0], base, kVtH, kElemF16, 128)) return e;       // half a block per CTA of the pair
      if (int e = make_kmajor_map(&maps[1], base + cnt, kVtH, kElemF16, 128)) return e;
    } else {
      std::vector<float> img(2 * cnt);
      for (int kx = 0; kx < kVtFlat; ++kx)
        for (int o = 0; o < kVtH; ++o)
          split_tf32(w3p[(size_t)kx * kVtH + o], img[(size_t)o * kVtFlat + kx], img[cnt + (size_t)o * kVtFlat + kx]);
      if (int e = h->vt_w3_bf16.reserve(2 * cnt * 4)) return e;
      MDC_CUDA(cudaMemcpy(h->vt_w3_bf16.ptr, img.data(), 2 * cnt * 4, cudaMemcpyHostToDevice));
      const float* base = reinterpret_cast<const float*>(h->vt_w3_bf16.ptr);
      if (int e = make_kmajor_map(&maps[0], base, kVtH, kElemF32, 128)) return e;      // half a block per CTA of the pair
      if (int e = make_kmajor_map(&maps[1], base + cnt, kVtH, kElemF32, 128)) return e;
    }
  }
  // every kernel attribute is set here, so that the predict calls only enqueue
  MDC_CUDA(cudaFuncSetAttribute(vt_conv_kernel<kConvBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<kConvBF16>::total));
  MDC_CUDA(cudaFuncSetAttribute(vt_conv_kernel<kConvTF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<kConvTF32>::total));
  MDC_CUDA(cudaFuncSetAttribute(vt_conv_kernel<kConvF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<kConvF16>::total));
  MDC_CUDA(cudaFuncSetAttribute(vt_dense_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DenseT32Smem::total));
  MDC_CUDA(cudaFuncSetAttribute(vt_dense_f16x3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, DenseF16Smem::total));
  MDC_CUDA(cudaFuncSetAttribute(vt_dense_f16x3_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, DenseF16Smem::total));
  if (cm == kConvBF16) {
    switch (h->C) {
#define MDC_DENSE_ATTR(CC) \
  case CC: MDC_CUDA(cudaFuncSetAttribute(vt_dense_bf16_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, DenseSmem::total)); break;
      MDC_DENSE_ATTR(1) MDC_DENSE_ATTR(2) MDC_DENSE_ATTR(3) MDC_DENSE_ATTR(4) MDC_DENSE_ATTR(5) MDC_DENSE_ATTR(6)
      MDC_DENSE_ATTR(7) MDC_DENSE_ATTR(8) MDC_DENSE_ATTR(9) MDC_DENSE_ATTR(10) MDC_DENSE_ATTR(11) MDC_DENSE_ATTR(12)
      MDC_DENSE_ATTR(13) MDC_DENSE_ATTR(14) MDC_DENSE_ATTR(15) MDC_DENSE_ATTR(16)
#undef MDC_DENSE_ATTR
    }
  }
  return MDC_OK;
}

// frames per pass.  bf16: act = 21 KB/frame -> 1.38 GB; f16x3: hi + lo fp16 = 42 KB/frame -> 2.77 GB; tf32x3: hi + lo
// fp32 = 84 KB/frame, and 148 x 128 frames is exactly one wave of dense tiles -> 1.6 GB
int64_t vt_pass_frames(const mdc_handle_s* h) {
  return h->mode == MDC_MODE_TF32X3 ? (int64_t)h->num_sms * kTM : 65536;
}

static size_t act_bytes_per_elem(const mdc_handle_s* h) {
  return h->mode == MDC_MODE_TF32X3 ? 8 : (h->mode == MDC_MODE_F16X3 ? 4 : 2);
}

// work space for passes of up to `frames` frames (grow-only; never called while a stream is capturing)
int vt_reserve(mdc_handle_s* h, int64_t frames) {
  const size_t act_elems = (size_t)frames * kVtFlat;
  if (act_elems > h->vt_act_elems) {
    if (int e = h->ws_act.reserve(act_elems * act_bytes_per_elem(h))) return e;
    if (int e = h->ws_h.reserve((size_t)frames * kVtH * 4)) return e;
    h->vt_act_elems = act_elems;
  }
  return MDC_OK;
}

// conv1 + conv2 of m frames at x (format in_fmt) -> activations of frames [frame_offset, frame_offset + m) of the pass
int launch_vt_conv(mdc_handle_s* h, const void* x, int in_fmt, int64_t m, int64_t frame_offset, cudaStream_t stream) {
  const int cm = conv_mode_of(h);
  if (m == 0) return MDC_OK;
  const size_t off = (size_t)frame_offset * kVtFlat;
  uint8_t* ws = reinterpret_cast<uint8_t*>(h->ws_act.ptr);
  const size_t eb = cm == kConvTF32 ? 4 : 2;                 // bytes per element of one activation matrix
  void* act0 = ws + off * eb;
  void* act1 = cm == kConvBF16 ? nullptr : (void*)(ws + (h->vt_act_elems + off) * eb);
  ConvW1 w1c;
  static_assert(sizeof(ConvW1) == 32 * 32 * sizeof(float), "conv1 image size");
  memcpy(&w1c, h->vt_w1_img.data(), sizeof(w1c));
  const long long out_rows = cm == kConvBF16 ? ConvCfg<kConvBF16>::kOutRows : ConvCfg<kConvTF32>::kOutRows;
  static_assert(ConvCfg<kConvTF32>::kOutRows == ConvCfg<kConvF16>::kOutRows, "split modes share the tile height");
  const long long num_st = (m * 132 + out_rows - 1) / out_rows;
  const long long pairs_needed = (num_st + 1) / 2, pairs_max = h->num_sms / 2;
  const unsigned grid_c = 2u * (unsigned)(pairs_needed < pairs_max ? pairs_needed : pairs_max);
  const float* b2 = reinterpret_cast<const float*>(h->vt_b2.ptr);
  const uint8_t* w2 = reinterpret_cast<const uint8_t*>(h->vt_w2_bf16.ptr);
  const uint8_t* xb = reinterpret_cast<const uint8_t*>(x);
  unsigned int* flags = reinterpret_cast<unsigned int*>(h->vt_flags.ptr);
  prof_begin(h, stream);
  if (cm == kConvTF32)
    vt_conv_kernel<kConvTF32><<<grid_c, ConvCfg<kConvTF32>::kThreads, ConvCfg<kConvTF32>::total, stream>>>(
        w1c, xb, in_fmt, m, b2, w2, act0, act1, num_st, h->vt_xlimit, flags);
  else if (cm == kConvF16)
    vt_conv_kernel<kConvF16><<<grid_c, ConvCfg<kConvF16>::kThreads, ConvCfg<kConvF16>::total, stream>>>(
        w1c, xb, in_fmt, m, b2, w2, act0, act1, num_st, h->vt_xlimit, flags);
  else
    vt_conv_kernel<kConvBF16><<<grid_c, ConvCfg<kConvBF16>::kThreads, ConvCfg<kConvBF16>::total, stream>>>(
        w1c, xb, in_fmt, m, b2, w2, act0, act1, num_st, h->vt_xlimit, flags);
  prof_end(h, stream);
  h->launches += 1;
  MDC_CUDA(cudaGetLastError());
  return MDC_OK;
}

template <int C>
static int dense_bf16_launch(mdc_handle_s* h, unsigned grid, cudaStream_t stream, const CUtensorMap& map_a,
                             const CUtensorMap& map_b, const float* b3, float* hb, int64_t m, int tiles, float* probs,
                             float* dense, int32_t* cls, unsigned long long* hist) {
  HeadW<C> hw;
  memcpy(hw.w, h->w[MDC_T_DENSE2_K].data(), sizeof(hw.w));     // Keras (256, C) row-major
  memcpy(hw.b, h->w[MDC_T_DENSE2_B].data(), sizeof(hw.b));
  vt_dense_bf16_kernel<C><<<grid, kDenseThreads, DenseSmem::total, stream>>>(map_a, map_b, hw, b3, hb, m, tiles, probs,
                                                                            dense, cls, hist);
  return MDC_OK;
}

static int dense_bf16_dispatch(mdc_handle_s* h, unsigned grid, cudaStream_t stream, const CUtensorMap& map_a,
                               const CUtensorMap& map_b, const float* b3, float* hb, int64_t m, int tiles, float* probs,
                               float* dense, int32_t* cls, unsigned long long* hist) {
  switch (h->C) {
#define MDC_DENSE_CASE(CC) \
  case CC: return dense_bf16_launch<CC>(h, grid, stream, map_a, map_b, b3, hb, m, tiles, probs, dense, cls, hist);
    MDC_DENSE_CASE(1) MDC_DENSE_CASE(2) MDC_DENSE_CASE(3) MDC_DENSE_CASE(4) MDC_DENSE_CASE(5) MDC_DENSE_CASE(6)
    MDC_DENSE_CASE(7) MDC_DENSE_CASE(8) MDC_DENSE_CASE(9) MDC_DENSE_CASE(10) MDC_DENSE_CASE(11) MDC_DENSE_CASE(12)
    MDC_DENSE_CASE(13) MDC_DENSE_CASE(14) MDC_DENSE_CASE(15) MDC_DENSE_CASE(16)
#undef MDC_DENSE_CASE
  }
  set_error("classes=%d outside 1..16", h->C);
  return MDC_ERR_UNSUPPORTED;
}

// dense1 + Dense(C) + softmax over the first m frames of the pass
int launch_vt_dense_head(mdc_handle_s* h, int64_t m, float* probs, float* dense, int32_t* cls,
                         unsigned long long* hist, cudaStream_t stream) {
  const int cm = conv_mode_of(h);
  if (m == 0) return MDC_OK;
  float* hb = reinterpret_cast<float*>(h->ws_h.ptr);
  const CUtensorMap* wmaps = reinterpret_cast<const CUtensorMap*>(h->tmap_w3);
  const float* b3 = reinterpret_cast<const float*>(h->vt_b3.ptr);
  if (cm == kConvBF16) {
    CUtensorMap map_a;
    if (int e = make_kmajor_map(&map_a, h->ws_act.ptr, (uint64_t)m, kElemBF16, kDM)) return e;
    const int tiles = (int)((m + kDM - 1) / kDM);
    const unsigned grid_d = (unsigned)(tiles < h->num_sms ? tiles : h->num_sms);
    static const bool keep_h = getenv("MDC_VT_KEEP_H") != nullptr;       // debugging: also store dense1 activations
    if (int e = dense_bf16_dispatch(h, grid_d, stream, map_a, wmaps[0], b3, keep_h ? hb : nullptr, m, tiles, probs, dense,
                                    cls, hist))
      return e;
    h->launches += 1;
    MDC_CUDA(cudaGetLastError());
    return MDC_OK;
  }
  CUtensorMap map_ah, map_al;
  const int tiles = (int)((m + kTM - 1) / kTM);
  const int pairs = (tiles + 1) / 2, pairs_max = h->num_sms / 2;
  const unsigned grid_d = 2u * (unsigned)(pairs < pairs_max ? pairs : pairs_max);
  if (cm == kConvF16) {
    const uint16_t* act = reinterpret_cast<const uint16_t*>(h->ws_act.ptr);
    if (int e = make_kmajor_map(&map_ah, act, (uint64_t)m, kElemF16, kTM)) return e;
    if (int e = make_kmajor_map(&map_al, act + h->vt_act_elems, (uint64_t)m, kElemF16, kTM)) return e;
    // Last round: with P pair-tiles on G CTA pairs, P mod G pairs would work while the others idle for a whole tile
    // (65,536 frames: 256 = 3 x 74 + 34).  When at most half of the pairs are busy in that round, each of its
    // pair-tiles is given to TWO CTA pairs, 128 of the 256 outputs each (the round then costs half a tile); the order
    // of every sum is unchanged, so the results are bit-identical to unsplit tiles.  Small batches (P <= G / 2) are
    // split entirely: twice the SMs, half the latency.
    const int G = pairs_max;
    int split = pairs < G ? pairs : pairs % G;
    if (2 * split > G) split = 0;
    static const bool no_split = getenv("MDC_VT_NO_SPLIT") != nullptr;       // tuning aid
    if (no_split) split = 0;
    const unsigned grid_s = 2u * (unsigned)std::max(std::min(pairs - split, G), 2 * split);   // CTAs launched
    const int64_t m_full = std::min<int64_t>(m, (int64_t)(pairs - split) * 2 * kTM);          // frames of unsplit tiles
    if (h->C == 11) {     // the VT-CNN2 class count: Dense(11) + softmax fused into the epilogue, no head launch
      HeadW<11> hw;
      memcpy(hw.w, h->w[MDC_T_DENSE2_K].data(), sizeof(hw.w));     // Keras (256, C) row-major
      memcpy(hw.b, h->w[MDC_T_DENSE2_B].data(), sizeof(hw.b));
      vt_dense_f16x3_kernel<11><<<grid_s, kDenseT32Threads, DenseF16Smem::total, stream>>>(
          map_ah, map_al, wmaps[0], wmaps[1], hw, b3, hb, m, tiles, split, probs, dense, cls, hist);
      h->launches += 1;
      MDC_CUDA(cudaGetLastError());
      if (split) {   // the frames of the split tiles: same arithmetic order as the fused epilogue
        const int64_t rows = m - m_full;
        vt_head_ordered_kernel<11><<<(unsigned)((rows + 127) / 128), 128, 0, stream>>>(
            hw, hb + m_full * 256, rows, probs ? probs + m_full * 11 : nullptr, dense ? dense + m_full * 11 : nullptr,
            cls ? cls + m_full : nullptr, hist);
        h->launches += 1;
        MDC_CUDA(cudaGetLastError());
      }
      return MDC_OK;
    }
    vt_dense_f16x3_kernel<0><<<grid_s, kDenseT32Threads, DenseF16Smem::total, stream>>>(
        map_ah, map_al, wmaps[0], wmaps[1], HeadW<1>{}, b3, hb, m, tiles, split, nullptr, nullptr, nullptr, nullptr);
  } else {
    const float* act = reinterpret_cast<const float*>(h->ws_act.ptr);
    if (int e = make_kmajor_map(&map_ah, act, (uint64_t)m, kElemF32, kTM)) return e;
    if (int e = make_kmajor_map(&map_al, act + h->vt_act_elems, (uint64_t)m, kElemF32, kTM)) return e;
    vt_dense_tf32x3_kernel<<<grid_d, kDenseT32Threads, DenseT32Smem::total, stream>>>(map_ah, map_al, wmaps[0], wmaps[1],
                                                                                       b3, hb, m, tiles);
  }
  h->launches += 1;
  MDC_CUDA(cudaGetLastError());
  return launch_vt_head(h, hb, m, probs, dense, cls, hist, stream);
}

// x: n frames in format in_fmt (MDC_IN_*).  The work space is sized by mdc_reserve or by the first call; a call made
// while the stream is being captured into a CUDA graph must find it large enough already.
int launch_vt_bf16(mdc_handle_s* h, const void* x, int in_fmt, int64_t n, float* probs, float* dense,
                   int32_t* cls, unsigned long long* hist, cudaStream_t stream) {
  if (n == 0) return MDC_OK;
  const int64_t CH = vt_pass_frames(h);
  const int64_t need = n < CH ? n : CH;
  if ((size_t)need * kVtFlat > h->vt_act_elems) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) {
      set_error("mdc_predict_*: the work space holds %lld frames per pass, this call needs %lld - call mdc_reserve before "
                "capturing into a CUDA graph", (long long)(h->vt_act_elems / kVtFlat), (long long)need);
      return MDC_ERR_NOT_READY;
    }
    if (int e = vt_reserve(h, need)) return e;
  }
  const size_t fb = in_fmt == MDC_IN_U8IQ ? 256 : (in_fmt == MDC_IN_I16 ? 512 : 1024);
  for (int64_t s = 0; s < n; s += CH) {
    const int64_t m = (n - s) < CH ? (n - s) : CH;
    if (int e = launch_vt_conv(h, reinterpret_cast<const uint8_t*>(x) + (size_t)s * fb, in_fmt, m, 0, stream)) return e;
    if (int e = launch_vt_dense_head(h, m, probs ? probs + s * h->C : nullptr, dense ? dense + s * h->C : nullptr,
                                     cls ? cls + s : nullptr, hist, stream))
      return e;
  }
  return MDC_OK;
}

}  // namespace mdc
