// VT-CNN2 forward on the sm_100a tensor cores: MDC_MODE_BF16 (bf16 operands, fp32 accumulate) and
// MDC_MODE_TF32X3 (every fp32 operand split into tf32 hi + lo, three kind::tf32 MMAs per product:
// hi*hi + hi*lo + lo*hi - fp32-level accuracy, <= 1e-5 of the fp64 oracle, at ~1/6 of the bf16 rate).
//
// Layer stack: /root/reference/examples-master/modulation_recognition/
// RML2016.10a_VTCNN2_example.ipynb:231-243 (shapes :194-216); Dropout = identity.
//
// Persistent, warp-specialised tcgen05 kernels:
//
//   vt_conv_kernel<TF32>      conv1 (1x3, 256 ch, fp32 FMA on CUDA cores, produced straight into the
//                             shared-memory A operand) -> conv2 (2x3, 80 ch) as an implicit GEMM
//                             M = frames*132, N = 80, K = 3 taps x 512 (row,channel) -> +bias, ReLU
//                             -> activations act[frames*132][80]  (== Keras channels_last flatten):
//                             bf16, or fp32 hi / lo matrices in 3xTF32 mode
//   vt_dense_bf16_kernel<C>   act[frames][10560] x W3 -> +bias, ReLU -> Dense(C) -> softmax, argmax,
//                             histogram in the epilogue (TMA 128B-swizzled tiles, M = 256 per CTA,
//                             N = 256, K = 10560); h never goes to HBM
//   vt_dense_tf32x3_kernel    the same dense1 in 3xTF32 with fp32 master sums in registers, then
//   vt_head_kernel            Dense(C) + softmax + argmax + histogram (fp32, vt_f32.cu)
//   vt_conv240_kernel         experimental second bf16 conv formulation (taps as N), MDC_VT_CONV=n240
//
// The implicit GEMM keeps conv1's padded output positions as GEMM rows: frame f owns rows
// [132 f, 132 f + 134) of one long activation "tape" whose rows 132 f and 132 f + 1 are the zero
// padding shared by frame f-1 (right pad) and frame f (left pad).  conv2 output row R needs tape
// rows R, R+1, R+2, so with the no-swizzle K-major operand layout (8-channel groups, rows 16 B
// apart) tap j is the SAME shared-memory image with the descriptor start address moved by 16 j
// bytes - conv1 activations are produced once and read by three MMAs.
#include <cudaTypedefs.h>

#include "mdc_internal.cuh"
#include "sm100.cuh"

namespace mdc {
using namespace sm100;

int launch_vt_head(mdc_handle_s* h, const float* hbuf, int64_t n, float* probs, float* dense,
                   int32_t* cls, unsigned long long* hist, cudaStream_t stream);
int pack_vt_small(mdc_handle_s* h);
void vt_permute_w3(const mdc_handle_s* h, std::vector<float>& out);

// ------------------------------------------------------------------------------------------
// conv kernel geometry.  CTAs work in pairs (cluster of 2, tcgen05 cta_group::2): each CTA owns
// its own super-tile (tape rows, conv1 producers, accumulators, epilogue) and HALF of every W2
// chunk; one M = 256 MMA issued by the pair's leader covers a 128-row tile of each CTA, so the
// B operand is fetched from shared memory once per pair (measured: 45 cycles per MMA against 52
// for the single-CTA M = 128 x N = 80 shape, tools/umma_rate.cu / tools/umma2_probe.cu).
constexpr int kGroups = 4;                // 16-B K groups per chunk image: g = 2 * (channel block) + input row
constexpr int kBHalf = 40;                // W2 output channels held by each CTA of the pair
constexpr int kBLbo = kBHalf * 16;        // bytes between K groups of the B image
constexpr int kXFrames = 4;               // frames a super-tile's tape rows can touch
constexpr int kOutTile = 128 * 160;       // one 128 x 80 bf16 output tile
constexpr int kProdWarp0 = 6;

// A chunk is kCC conv1 channels x {I row, Q row} = two UMMA K steps of two 16-B groups:
//   bf16:   16 channels, 8 per group, one image        (K step = 16 values)
//   tf32x3:  8 channels, 4 per group, hi and lo images  (K step =  8 values)
//
// 3xTF32 accumulation.  tcgen05.mma rounds every accumulate TOWARD ZERO (tools/umma_acc_probe.cu: 1 + 0.75 ulp
// stays 1), so a chain of n MMAs into one accumulator comes out low by about n x 1.5e-8 relative (measured
// -8.8e-6 for the 576-MMA conv2 chain, -6e-5 for the 3,960-MMA dense1 chain).  The tf32x3 conv kernel
// therefore works on ONE 128-row tile per super-tile and spreads its MMAs over six TMEM accumulators:
// accumulator 0 takes the two small cross terms (lo*hi, hi*lo) of every K step, accumulators 1..5 take the
// hi*hi terms of chunks c = a - 1 (mod 5) - at most 42 truncating adds each - and the epilogue adds the six
// in fp32 round-to-nearest.  The accumulators are single-buffered (6 x 80 = 480 of 512 columns): the MMA
// warp waits while the epilogue reads them (about 4 % of a tile's 576 MMAs).
template <bool TF32>
struct ConvCfg {
  static constexpr int kNT = TF32 ? 1 : 3;                 // accumulator tiles (128 rows x 80 cols) per super-tile
  static constexpr int kTapeRows = 128 * kNT;              // tape rows staged per super-tile
  static constexpr int kOutRows = kTapeRows - 2;           // conv2 rows produced per super-tile (2-row halo)
  static constexpr int kALbo = kTapeRows * 16;             // bytes between K groups of the A image
  static constexpr int kProdWarps = kTapeRows / 32;        // one tape row per producer thread
  static constexpr int kThreads = (kProdWarp0 + kProdWarps) * 32;   // TMA, MMA, 4 epilogue, producers
  static constexpr int kAccBufs = TF32 ? 1 : 2;            // accumulator buffers in TMEM
  static constexpr int kAccSplit = TF32 ? 6 : 1;           // accumulators per tile (see above)
  static constexpr int kAccCols = kNT * kAccSplit * 80;    // TMEM columns per buffer
  static constexpr int kCC = TF32 ? 8 : 16;
  static constexpr int kPerGroup = TF32 ? 4 : 8;
  static constexpr int kImgs = TF32 ? 2 : 1;
  static constexpr int kChunks = 256 / kCC;
  static constexpr int kStages = 5;
  static constexpr int kAImg = kGroups * kALbo;            // bf16 24,576; tf32 8,192
  static constexpr int kASlot = kImgs * kAImg;
  static constexpr int kBImg = 3 * kGroups * kBLbo;        // 7,680: [tap][group][40][16 B]
  static constexpr int kBSlot = kImgs * kBImg;
  // shared memory map
  static constexpr int a = 0;
  static constexpr int b = a + kStages * kASlot;
  static constexpr int xs = b + kStages * kBSlot;
  static constexpr int out = xs + 2 * kXFrames * 1024;     // two frame buffers before it
  static constexpr int b2 = out + (TF32 ? 0 : 2 * kOutTile);   // bf16 only: two staged output tiles
  static constexpr int bars = b2 + 320;
  // full[S], empty[S], x_full[2], x_empty[2], tmem_full[2], tmem_empty[2]
  static constexpr int nbars = 2 * kStages + 4 + 4;
  static constexpr int tmem_slot = bars + nbars * 8;
  static constexpr int total = tmem_slot + 16;
  static_assert(kAccBufs * kAccCols <= 512, "accumulators exceed TMEM");
};
static_assert(ConvCfg<false>::total <= 232448 && ConvCfg<true>::total <= 232448, "conv kernel shared memory exceeds 227 KB");

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

// conv1 weights travel as a kernel parameter (constant bank): with the chunk loop unrolled every
// weight is an immediate-offset uniform load feeding FFMA2 directly - no shared-memory broadcast
// loads (each LDS.128 broadcast costs 4 LSU wavefronts on the pipe the MMA operands also use) and
// no vector registers.  Layout: 32 channel groups x {w0[8], w1[8], w2[8], bias[8]} fp32.
struct ConvW1 {
  unsigned long long v[32 * 16];
};

// 8 conv1 channels of one tape row: relu(x0 w0 + x1 w1 + x2 w2 + b) -> 8 x bf16 (16 B)
__device__ __forceinline__ uint4 conv1_item(uint64_t x0, uint64_t x1, uint64_t x2, const unsigned long long* w,
                                            uint32_t m) {
  uint64_t a0 = fma2_u(x0, w[0], w[12]), a1 = fma2_u(x0, w[1], w[13]);
  uint64_t a2 = fma2_u(x0, w[2], w[14]), a3 = fma2_u(x0, w[3], w[15]);
  a0 = fma2_u(x1, w[4], a0); a1 = fma2_u(x1, w[5], a1);
  a2 = fma2_u(x1, w[6], a2); a3 = fma2_u(x1, w[7], a3);
  a0 = fma2_u(x2, w[8], a0); a1 = fma2_u(x2, w[9], a1);
  a2 = fma2_u(x2, w[10], a2); a3 = fma2_u(x2, w[11], a3);
  return make_uint4(relu_pack(a0) & m, relu_pack(a1) & m, relu_pack(a2) & m, relu_pack(a3) & m);
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

// act0/act1: bf16 mode -> act0 = bf16 [rows][80]; tf32x3 mode -> act0 = hi, act1 = lo, fp32 [rows][80]
template <bool TF32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ConvCfg<TF32>::kThreads, 1)
vt_conv_kernel(const __grid_constant__ ConvW1 w1c, const float* __restrict__ x, long long n,
               const float* __restrict__ b2g, const uint8_t* __restrict__ w2img,
               void* __restrict__ act0, void* __restrict__ act1, long long num_st, int dbg_rt,
               long long* __restrict__ trace) {
  using ConvSmem = ConvCfg<TF32>;
  constexpr int kStages = ConvSmem::kStages, kChunks = ConvSmem::kChunks;
  constexpr int kASlot = ConvSmem::kASlot, kBSlot = ConvSmem::kBSlot, kAImg = ConvSmem::kAImg, kBImg = ConvSmem::kBImg;
  constexpr int kNT = ConvSmem::kNT, kTapeRows = ConvSmem::kTapeRows, kOutRows = ConvSmem::kOutRows, kALbo = ConvSmem::kALbo;
  constexpr int kProdWarps = ConvSmem::kProdWarps, kConvThreads = ConvSmem::kThreads, kAccCols = ConvSmem::kAccCols;
  constexpr int kAccBufs = ConvSmem::kAccBufs;
  (void)kTapeRows;
  // role-ablation flags for timing experiments (build with -DMDC_VT_ABLATE, set MDC_VT_DEBUG; the
  // results are garbage): 1 = producers skip conv1 math and stores, 2 = MMA warp skips the MMAs,
  // 4 = epilogue skips TMEM loads / math / stores.  Compiled out of the product build.
#ifdef MDC_VT_ABLATE
  const int dbg = dbg_rt;
  // clock64 trace of CTA 0 (ablate builds, MDC_VT_TRACE=file): trace[role * 512 + i]
#define MDC_TRACE3(role, i) do { if (trace && blockIdx.x == 0 && (i) < 512) trace[(role) * 512 + (i)] = clock64(); } while (0)
#else
  constexpr int dbg = 0;
#define MDC_TRACE3(role, i) do { } while (0)
#endif
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ConvSmem::bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* x_full = bars + 2 * kStages;   // [2] frame buffers
  uint64_t* x_empty = x_full + 2;          // [2]
  uint64_t* tmem_full = x_empty + 2;       // [2] accumulator buffers
  uint64_t* tmem_empty = tmem_full + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + ConvSmem::tmem_slot);

  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const long long total_rows = n * 132;
  const uint32_t rank = cluster_ctarank();                 // 0 = leader (issues the pair's MMAs)
  // the pair walks super-tiles 2 i and 2 i + 1 in lockstep; a trailing odd one is an all-masked no-op
  const long long st_first = 2ll * cluster_id_x(), st_step = 2ll * cluster_count_x();

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
        mbar_arrive_expect_tx(&x_full[b], nf * 1024);
        if (nf) bulk_g2s(smem + ConvSmem::xs + b * (kXFrames * 1024), x + f0 * 256, nf * 1024, &x_full[b]);
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
      const uint32_t idesc = TF32 ? make_idesc_tf32(256, 80) : make_idesc_bf16(256, 80);
      const uint32_t a_base = smem_u32(smem + ConvSmem::a), b_base = smem_u32(smem + ConvSmem::b);
      constexpr uint32_t hi = smem_desc_hi(128, 0);
      for (long long base = st_first; base < num_st; base += st_step, ++k) {
        const uint32_t buf = k % kAccBufs, use = k / kAccBufs;
        const uint32_t acc = tmem + buf * kAccCols;
        if (lane == 0) MDC_TRACE3(1, 2 * k);
        mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);          // both epilogues drained this buffer
        tc_fence_after_sync();
        if (lane == 0) MDC_TRACE3(1, 2 * k + 1);
        for (int c = 0; c < kChunks; ++c, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          mbar_wait(&full[s], ph);               // own producers, own TMA and the peer's relay
          tc_fence_after_sync();
          if (lane == 0) MDC_TRACE3(0, it);
          if (elect_one()) {
            const uint32_t a_lo = smem_desc_lo(a_base + s * kASlot, kALbo);
            const uint32_t b_lo = smem_desc_lo(b_base + s * kBSlot, kBLbo);
            // tf32x3: hi*hi terms of chunk c go to accumulator 1 + c % 5, the cross terms to accumulator 0
            const uint32_t acc_hh = acc + (1 + c % 5) * 80;
#pragma unroll
            for (int t = 0; t < kNT; ++t) {
              if (dbg & 2) continue;
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                  const uint32_t ao = ((2 * ks) * kALbo + (128 * t + j) * 16) >> 4;
                  const uint32_t bo = ((j * kGroups + 2 * ks) * kBLbo) >> 4;
                  if (!TF32) {
                    mma_bf16_ss_pair(acc + t * 80, desc64(a_lo + ao, hi), desc64(b_lo + bo, hi), idesc,
                                     (c | ks | j) != 0);
                  } else {
                    constexpr uint32_t al = kAImg >> 4, bl = kBImg >> 4;   // offsets of the lo images
                    mma_tf32_ss_pair(acc, desc64(a_lo + ao + al, hi), desc64(b_lo + bo, hi), idesc, (c | ks | j) != 0);
                    mma_tf32_ss_pair(acc, desc64(a_lo + ao, hi), desc64(b_lo + bo + bl, hi), idesc, 1);
                    mma_tf32_ss_pair(acc_hh, desc64(a_lo + ao, hi), desc64(b_lo + bo, hi), idesc,
                                     (c >= 5) || (ks | j) != 0);
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
    // ================= epilogue: TMEM -> +bias, ReLU -> bf16 tile in smem -> bulk store (bf16 mode)
    // or -> tf32 hi / lo fp32 rows straight to global (3xTF32 mode: the mainloop is 6x longer)
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const bool leader = (warp == 2 && lane == 0);
    const float4* b2s = reinterpret_cast<const float4*>(smem + ConvSmem::b2);
    uint32_t k = 0, tile_ctr = 0;
    for (long long base = st_first; base < num_st; base += st_step, ++k) {
      const long long r0 = (base + rank) * kOutRows;
      const uint32_t buf = k % kAccBufs, use = k / kAccBufs;
      mbar_wait(&tmem_full[buf], use & 1);
      tc_fence_after_sync();
#pragma unroll 1
      for (int t = 0; t < kNT; ++t, ++tile_ctr) {
        uint8_t* obuf = smem + ConvSmem::out + (tile_ctr & 1) * kOutTile;
        if (!TF32) {
          if (leader) bulk_wait_read<1>();        // the store issued two tiles ago has drained obuf
          named_bar_sync(1, 128);
        }
        uint32_t v[80];
        const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + buf * kAccCols;
        if (dbg & 4) {
#pragma unroll
          for (int i = 0; i < 80; ++i) v[i] = 0;
        } else if (!TF32) {
#pragma unroll
          for (int cc = 0; cc < 5; ++cc) {
            uint32_t(&vv)[16] = *reinterpret_cast<uint32_t(*)[16]>(&v[cc * 16]);
            tmem_ld16(tbase + t * 80 + cc * 16, vv);
          }
          tmem_ld_wait();
        } else {
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
        if (TF32) {
          const int rr = q * 32 + lane;
          if (rr < rows && !(dbg & 4)) {
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
        uint8_t* orow = obuf + (q * 32 + lane) * 160;
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
        // rows are 160 B apart, so the 16-B chunks of lanes l and l + 4 fall into the same banks: lanes with
        // bit 2 set store their chunks rotated by one, which makes every quarter-warp store conflict-free
        const bool rot = (lane & 4) != 0;
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          if (dbg & 4) break;
          const uint4 a = o[j], b = o[(j + 1) % 10];
          const uint4 val = make_uint4(rot ? b.x : a.x, rot ? b.y : a.y, rot ? b.z : a.z, rot ? b.w : a.w);
          *reinterpret_cast<uint4*>(orow + (rot ? ((j + 1) % 10) : j) * 16) = val;
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (leader) {
          if (rows > 0 && !(dbg & 4))
            bulk_s2g(reinterpret_cast<__nv_bfloat16*>(act0) + row_lo * 80, obuf, (uint32_t)rows * 160);
          bulk_commit();
        }
      }
    }
    if (!TF32 && leader) bulk_wait<0>();
  } else {
    // ================= conv1 producers: fp32 FMA -> ReLU -> bf16 -> A operand image.
    // One tape row per thread; a chunk is 16 channels x {I row, Q row}; weights come from the
    // constant bank (uniform registers), inputs stay in registers for the whole super-tile.
    const int pw = warp - kProdWarp0;
    const int row = pw * 32 + lane;
    uint32_t it = 0, k = 0;
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
        const float* xf = reinterpret_cast<const float*>(smem + ConvSmem::xs + xb * (kXFrames * 1024)) + (f - f0) * 256;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const int xi = p - 4 + j;               // conv1 position p-2 reads x[p-4 .. p-2]
            const float xv = (valid && xi >= 0 && xi < 128) ? xf[r * 128 + xi] : 0.f;
            xd[r][j] = pack_dup(xv);
          }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&x_empty[xb]);
#pragma unroll
      for (int c = 0; c < kChunks; ++c, ++it) {
        // compute the chunk into registers first: nothing here depends on the stage being free
        uint4 o[ConvSmem::kImgs * kGroups];
        if (!TF32) {
#pragma unroll
          for (int hc = 0; hc < 2; ++hc) {
            if (dbg & 1) break;
            const unsigned long long* w = &w1c.v[(c * 2 + hc) * 16];
            o[2 * hc] = conv1_item(xd[0][0], xd[0][1], xd[0][2], w, m);
            o[2 * hc + 1] = conv1_item(xd[1][0], xd[1][1], xd[1][2], w, m);
          }
        } else {
          const unsigned long long* w = &w1c.v[c * 16];       // chunk c = channels 8c .. 8c+7
#pragma unroll
          for (int qd = 0; qd < 2; ++qd) {
            if (dbg & 1) break;
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
        if (pw == 0 && lane == 0) MDC_TRACE3(2, 2 * it);
        mbar_wait(&empty[s], ph ^ 1);
        if (pw == 0 && lane == 0) MDC_TRACE3(2, 2 * it + 1);
        uint8_t* arow = smem + ConvSmem::a + s * kASlot + row * 16;
        if (!(dbg & 1)) {
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
  }

  // ---- teardown (the pair leaves together: the leader's MMAs read the peer's shared memory)
  __syncwarp();
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair<512>(tmem);
}

// ------------------------------------------------------------------------------------------
// bf16 conv kernel, second formulation ("N240"): the three conv2 taps become N instead of K.
//
// The kernel above fetches every conv1 activation tile from shared memory three times (once per tap, the
// descriptor shifted by one row) and its N = 80 MMAs are operand-fetch-bound: 45 cycles per M=256 x N=80 x
// K=16 MMA against a 40-cycle math floor, and with the producers' stores and the epilogue staging the
// kernel runs at 90 % of the 128 B/clk shared-memory bandwidth.  Here one MMA computes all three taps'
// partial products for a K step, P[row][tap*80 + o] = A[row][:] . W2[tap][:, o]   (N = 240: 120 cycles of
// math, 62 of operand fetch - math-bound, measured 120.0, tools/umma2_probe.cu), and the epilogue adds the
// three partials of an output row from three consecutive accumulator rows:
//     out[r][o] = P[r][o] + P[r+1][80 + o] + P[r+2][160 + o]
// (TMEM row = lane, so row r+1 is a warp shuffle; the two rows a warp needs from the next lane quarter go
// through a 960-B shared-memory exchange).  A is read once per K step, W2 (123 KB per CTA of the pair) stays
// resident in shared memory for the whole kernel, and the MMA count drops 3x.
//
// Geometry: CTA pair (cta_group::2, M = 256), one 128-row tape tile per CTA per iteration (126 outputs,
// 2-row halo), accumulators 2 x 240 TMEM columns (double-buffered), 16 K chunks of 16 channels x {I,Q}.
// 18 warps: TMA, MMA, 8 epilogue (lane quarter x column half: a tile is only 3,840 MMA cycles, one warp per
// quarter needs ~10,000 for the 240-column shift-add), 8 conv1 producers (32-row block x chunk phase).
struct Conv240 {
  static constexpr int kRows = 128;
  static constexpr int kOutRows = kRows - 2;
  static constexpr int kALbo = kRows * 16;                 // bytes between K groups of the A image
  static constexpr int kChunks = 16;
  static constexpr int kStages = 6;
  static constexpr int kASlot = kGroups * kALbo;           // 8,192
  static constexpr int kBRows = 120;                       // accumulator columns (tap*80 + o) held per CTA
  static constexpr int kBLbo = kBRows * 16;
  static constexpr int kBChunk = kGroups * kBLbo;          // 7,680
  static constexpr int kBBytes = kChunks * kBChunk;        // 122,880 resident
  static constexpr int kXF = 2;                            // frames a 128-row tile can touch
  static constexpr int kEpiWarps = 8;                      // (TMEM lane quarter) x (column half)
  static constexpr int kProd0 = 2 + kEpiWarps;             // first producer warp
  static constexpr int kPhases = 2;                        // a producer warp handles every kPhases-th chunk
  static constexpr int kProdWarps = 4 * kPhases;           // (32-row block) x (chunk phase)
  static constexpr int kThreads = (kProd0 + kProdWarps) * 32;
  static constexpr int kAccCols = 240;
  static constexpr int kOutBytes = kOutRows * 160;         // 20,160: one staged bf16 output tile
  static constexpr int kXchFloats = 240;                   // per warp: P1 of lane 0, P2 of lane 0, P2 of lane 1
  // shared memory map
  static constexpr int b = 0;
  static constexpr int a = b + kBBytes;
  static constexpr int xs = a + kStages * kASlot;
  static constexpr int out = xs + 2 * kXF * 1024;
  static constexpr int xch = out + 2 * 128 * 160;
  static constexpr int b2 = xch + 2 * 4 * kXchFloats * 4;
  static constexpr int bars = b2 + 320;
  // full[S], empty[S], b_full, b_peer, x_full[2], x_empty[2], tmem_full[2], tmem_empty[2]
  static constexpr int nbars = 2 * kStages + 2 + 4 + 4;
  static constexpr int tmem_slot = bars + nbars * 8;
  static constexpr int total = tmem_slot + 16;
};
static_assert(Conv240::total <= 232448, "N240 conv kernel shared memory exceeds 227 KB");

// producer main loop of the warps with chunk phase CP: chunks CP, CP + kPhases, ... of every tile, all four K
// groups of one tape row per thread (CP is compile-time, so conv1 weights are immediate constant-bank
// operands).
template <int CP>
__device__ __forceinline__ void conv240_produce(const ConvW1& w1c, const uint64_t (&xd)[2][3], uint32_t m, uint8_t* arow0,
                                                uint64_t* full, uint64_t* empty, uint32_t tile_k, uint32_t& prev, int lane,
                                                uint32_t rank, int dbg, long long* trace) {
#pragma unroll
  for (int i = 0; i < Conv240::kChunks / Conv240::kPhases; ++i) {
    const int c = Conv240::kPhases * i + CP;
    uint4 o[kGroups];
#pragma unroll
    for (int hc = 0; hc < 2; ++hc) {
      if (dbg & 1) break;
      const unsigned long long* w = &w1c.v[(c * 2 + hc) * 16];
      o[2 * hc] = conv1_item(xd[0][0], xd[0][1], xd[0][2], w, m);
      o[2 * hc + 1] = conv1_item(xd[1][0], xd[1][1], xd[1][2], w, m);
    }
    if (prev != 0xFFFFFFFFu) {       // publish this warp's previous chunk: its stores were issued a chunk of math ago
      if (!(dbg & 8)) fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&full[prev]);
        else mbar_arrive_remote(&full[prev], 0);
      }
    }
    const uint32_t it = tile_k * Conv240::kChunks + c;
    const uint32_t s = it % Conv240::kStages, ph = (it / Conv240::kStages) & 1;
    if (trace && lane == 0 && (tile_k * 8 + i) < 128) trace[2 * 512 + (tile_k * 8 + i) * 4] = clock64();
    mbar_wait(&empty[s], ph ^ 1);
    if (trace && lane == 0 && (tile_k * 8 + i) < 128) trace[2 * 512 + (tile_k * 8 + i) * 4 + 1] = clock64();
    if (!(dbg & 1)) {
#pragma unroll
      for (int g = 0; g < kGroups; ++g) *reinterpret_cast<uint4*>(arow0 + s * Conv240::kASlot + g * Conv240::kALbo) = o[g];
    }
    prev = s;
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Conv240::kThreads, 1)
vt_conv240_kernel(const __grid_constant__ ConvW1 w1c, const float* __restrict__ x, long long n,
                  const float* __restrict__ b2g, const uint8_t* __restrict__ w2img, __nv_bfloat16* __restrict__ act,
                  long long num_tiles, int dbg_rt, long long* __restrict__ trace) {
  using Cfg = Conv240;
#ifdef MDC_VT_ABLATE   // role-ablation flags for timing experiments (results are garbage): see vt_conv_kernel
  const int dbg = dbg_rt;
  // clock64 trace of CTA 0 (ablate builds): trace[role * 512 + i]
#define MDC_TRACE(role, i) do { if (trace && blockIdx.x == 0 && (i) < 512) trace[(role) * 512 + (i)] = clock64(); } while (0)
#else
  constexpr int dbg = 0;
#define MDC_TRACE(role, i) do { } while (0)
#endif
  constexpr int kStages = Cfg::kStages, kChunks = Cfg::kChunks, kProdWarps = Cfg::kProdWarps;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* b_full = bars + 2 * kStages;
  uint64_t* b_peer = b_full + 1;           // (leader's only) the peer CTA's W2 half is resident
  uint64_t* x_full = b_peer + 1;           // [2] frame buffers
  uint64_t* x_empty = x_full + 2;          // [2]
  uint64_t* tmem_full = x_empty + 2;       // [2] accumulator buffers
  uint64_t* tmem_empty = tmem_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::tmem_slot);

  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const long long total_rows = n * 132;
  const uint32_t rank = cluster_ctarank();                 // 0 = leader (issues the pair's MMAs)
  const long long t_first = 2ll * cluster_id_x(), t_step = 2ll * cluster_count_x();

  for (int i = tid; i < 80; i += Cfg::kThreads) reinterpret_cast<float*>(smem + Cfg::b2)[i] = b2g[i];
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 8);                 // (leader's only) the four row-block warps of the chunk's phase in each CTA
      mbar_init(&empty[s], 1);
    }
    mbar_init(b_full, 1);
    mbar_init(b_peer, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&x_full[b], 1);
      mbar_init(&x_empty[b], kProdWarps);
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 2 * Cfg::kEpiWarps);   // (leader's only) the epilogue warps of both CTAs
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
    // ================= TMA: this CTA's half of W2 (once), then the frames of every tile
    if (elect_one()) {
      mbar_arrive_expect_tx(b_full, Cfg::kBBytes);
      const uint8_t* src = w2img + (size_t)rank * Cfg::kBBytes;
      for (int c = 0; c < kChunks; ++c)
        bulk_g2s(smem + Cfg::b + c * Cfg::kBChunk, src + (size_t)c * Cfg::kBChunk, Cfg::kBChunk, b_full);
    }
    __syncwarp();
    uint32_t k = 0;
    for (long long base = t_first; base < num_tiles; base += t_step, ++k) {
      const long long f0 = ((base + rank) * Cfg::kOutRows) / 132;
      const long long left = n - f0;
      const uint32_t nf = left <= 0 ? 0u : (left < Cfg::kXF ? (uint32_t)left : (uint32_t)Cfg::kXF);
      const uint32_t b = k & 1;
      mbar_wait(&x_empty[b], ((k >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&x_full[b], nf * 1024);
        if (nf) bulk_g2s(smem + Cfg::xs + b * (Cfg::kXF * 1024), x + f0 * 256, nf * 1024, &x_full[b]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only; the peer's warp 1 just reports its W2 half resident)
    uint32_t it = 0, k = 0;
    mbar_wait(b_full, 0);                        // this CTA's W2 half has landed
    if (rank == 0) {
      mbar_wait(b_peer, 0);                      // ... and the peer's
      const uint32_t idesc = make_idesc_bf16(256, 240);
      const uint32_t a_base = smem_u32(smem + Cfg::a), b_base = smem_u32(smem + Cfg::b);
      constexpr uint32_t hi = smem_desc_hi(128, 0);
      for (long long base = t_first; base < num_tiles; base += t_step, ++k) {
        const uint32_t buf = k & 1;
        const uint32_t acc = tmem + buf * Cfg::kAccCols;
        mbar_wait(&tmem_empty[buf], ((k >> 1) & 1) ^ 1);     // both epilogues drained this buffer
        tc_fence_after_sync();
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          mbar_wait(&full[s], ph);               // the producers of both CTAs have published this stage
          tc_fence_after_sync();
          if (lane == 0) MDC_TRACE(0, it);
          if (elect_one()) {
            const uint32_t a_lo = smem_desc_lo(a_base + s * Cfg::kASlot, Cfg::kALbo);
            const uint32_t b_lo = smem_desc_lo(b_base + c * Cfg::kBChunk, Cfg::kBLbo);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              if (!(dbg & 2)) mma_bf16_ss_pair(acc, desc64(a_lo + ((2 * ks * Cfg::kALbo) >> 4), hi),
                               desc64(b_lo + ((2 * ks * Cfg::kBLbo) >> 4), hi), idesc, (c | ks) != 0);
            mma_commit_pair(&empty[s]);
            if (c == kChunks - 1) mma_commit_pair(&tmem_full[buf]);
          }
          __syncwarp();
        }
      }
    } else {
      if (elect_one()) mbar_arrive_remote(b_peer, 0);
      __syncwarp();
    }
  } else if (warp < Cfg::kProd0) {
    // ================= epilogue: shift-add the three taps, +bias, ReLU -> bf16 tile in smem -> bulk store.
    // warp = (TMEM lane quarter q, column half hf): output channels 40 hf .. 40 hf + 39 of rows 32 q .. 32 q + 31
    const int q = warp & 3, hf = (warp - 2) >> 2;
    const bool leader = (warp == 2 && lane == 0);
    constexpr int kEpiThreads = Cfg::kEpiWarps * 32;
    const float* b2s = reinterpret_cast<const float*>(smem + Cfg::b2) + hf * 40;
    float* xch = reinterpret_cast<float*>(smem + Cfg::xch);
    uint32_t k = 0;
    for (long long base = t_first; base < num_tiles; base += t_step, ++k) {
      const long long row_lo = (base + rank) * Cfg::kOutRows;
      const uint32_t buf = k & 1;
      uint8_t* obuf = smem + Cfg::out + buf * (128 * 160);
      float* xw = xch + ((k & 1) * 4 + q) * Cfg::kXchFloats + hf * 120;   // [P1 lane 0 | P2 lane 0 | P2 lane 1] x 40
      mbar_wait(&tmem_full[buf], (k >> 1) & 1);
      tc_fence_after_sync();
      if (warp == 2 && lane == 0) MDC_TRACE(1, 4 * k);
      const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + buf * Cfg::kAccCols + hf * 40;
      // pass 1: lanes 0 and 1 publish the tap-1 / tap-2 partials the previous lane quarter needs
      if (!(dbg & 4)) {
        uint32_t p1[5][8], p2[5][8];
#pragma unroll
        for (int g = 0; g < 5; ++g) {
          tmem_ld8(tbase + 80 + g * 8, p1[g]);
          tmem_ld8(tbase + 160 + g * 8, p2[g]);
        }
        tmem_ld_wait();
        if (lane == 0) {
#pragma unroll
          for (int g = 0; g < 5; ++g) {
            *reinterpret_cast<uint4*>(xw + g * 8) = make_uint4(p1[g][0], p1[g][1], p1[g][2], p1[g][3]);
            *reinterpret_cast<uint4*>(xw + g * 8 + 4) = make_uint4(p1[g][4], p1[g][5], p1[g][6], p1[g][7]);
            *reinterpret_cast<uint4*>(xw + 40 + g * 8) = make_uint4(p2[g][0], p2[g][1], p2[g][2], p2[g][3]);
            *reinterpret_cast<uint4*>(xw + 40 + g * 8 + 4) = make_uint4(p2[g][4], p2[g][5], p2[g][6], p2[g][7]);
          }
        }
        if (lane == 1) {
#pragma unroll
          for (int g = 0; g < 5; ++g) {
            *reinterpret_cast<uint4*>(xw + 80 + g * 8) = make_uint4(p2[g][0], p2[g][1], p2[g][2], p2[g][3]);
            *reinterpret_cast<uint4*>(xw + 80 + g * 8 + 4) = make_uint4(p2[g][4], p2[g][5], p2[g][6], p2[g][7]);
          }
        }
      }
      if (leader) bulk_wait_read<1>();            // the store issued two tiles ago has drained obuf
      named_bar_sync(1, kEpiThreads);             // exchange rows published; obuf free
      if (warp == 2 && lane == 0) MDC_TRACE(1, 4 * k + 1);
      // pass 2: out[r] = P0[r] + P1[r+1] + P2[r+2], +bias, ReLU, bf16 -> staged row
      const float* xn = xw + Cfg::kXchFloats;     // rows published by the next lane quarter
      const bool edge = (lane >= 30) && (q < 3);
      uint8_t* orow = obuf + (q * 32 + lane) * 160 + hf * 80;
      uint32_t p0[2][8], p1[2][8], p2[2][8];      // two groups in flight: the loads of g + 1 fly under the math of g
      if (!(dbg & 4)) {
        tmem_ld8(tbase, p0[0]);
        tmem_ld8(tbase + 80, p1[0]);
        tmem_ld8(tbase + 160, p2[0]);
      }
#pragma unroll
      for (int g = 0; g < 5; ++g) {
        const int cur = g & 1, nxt = cur ^ 1;
        if (!(dbg & 4)) tmem_ld_wait();
        if (g < 4 && !(dbg & 4)) {
          tmem_ld8(tbase + (g + 1) * 8, p0[nxt]);
          tmem_ld8(tbase + 80 + (g + 1) * 8, p1[nxt]);
          tmem_ld8(tbase + 160 + (g + 1) * 8, p2[nxt]);
        }
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float a1 = __shfl_down_sync(0xffffffffu, __uint_as_float(p1[cur][e]), 1);
          const float a2 = __shfl_down_sync(0xffffffffu, __uint_as_float(p2[cur][e]), 2);
          v[e] = __uint_as_float(p0[cur][e]) + (lane < 31 ? a1 : 0.f) + (lane < 30 ? a2 : 0.f);
        }
        if (edge) {
#pragma unroll
          for (int e = 0; e < 8; e += 4) {
            const int o = g * 8 + e;
            if (lane == 30) {
              const float4 p2a = *reinterpret_cast<const float4*>(xn + 40 + o);
              v[e] += p2a.x; v[e + 1] += p2a.y; v[e + 2] += p2a.z; v[e + 3] += p2a.w;
            } else {
              const float4 p1n = *reinterpret_cast<const float4*>(xn + o);
              const float4 p2b = *reinterpret_cast<const float4*>(xn + 80 + o);
              v[e] += p1n.x + p2b.x; v[e + 1] += p1n.y + p2b.y; v[e + 2] += p1n.z + p2b.z; v[e + 3] += p1n.w + p2b.w;
            }
          }
        }
        const float4 ba = *reinterpret_cast<const float4*>(b2s + g * 8), bb = *reinterpret_cast<const float4*>(b2s + g * 8 + 4);
        const uint4 o = make_uint4(cvt_relu_bf16x2(v[1] + ba.y, v[0] + ba.x), cvt_relu_bf16x2(v[3] + ba.w, v[2] + ba.z),
                                   cvt_relu_bf16x2(v[5] + bb.y, v[4] + bb.x), cvt_relu_bf16x2(v[7] + bb.w, v[6] + bb.z));
        *reinterpret_cast<uint4*>(orow + g * 16) = o;
        if (g == 3) {
          // (the loads of the last group were issued above; its wait is the next iteration's)
        }
      }
      // whole buffer read: hand it back to the MMA warp
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&tmem_empty[buf]);
        else mbar_arrive_remote(&tmem_empty[buf], 0);
      }
      long long rows = Cfg::kOutRows;
      if (row_lo + rows > total_rows) rows = total_rows - row_lo;
      if (warp == 2 && lane == 0) MDC_TRACE(1, 4 * k + 2);
      fence_proxy_async_smem();
      named_bar_sync(2, kEpiThreads);
      if (leader) {
        if (rows > 0) bulk_s2g(act + row_lo * 80, obuf, (uint32_t)rows * 160);
        bulk_commit();
      }
      if (warp == 2 && lane == 0) MDC_TRACE(1, 4 * k + 3);
    }
    if (leader) bulk_wait<0>();
  } else {
    // ================= conv1 producers: warp = (32-row block, chunk phase); fp32 FMA -> ReLU -> bf16 -> A image
    const int pw = warp - Cfg::kProd0;
    const int cp = pw & (Cfg::kPhases - 1);       // chunks cp, cp + kPhases, ... of every tile
    const int row = (pw / Cfg::kPhases) * 32 + lane;
    uint32_t k = 0, prev = 0xFFFFFFFFu;
    for (long long base = t_first; base < num_tiles; base += t_step, ++k) {
      const long long t0 = (base + rank) * Cfg::kOutRows;   // first tape row of this tile
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
        const float* xf = reinterpret_cast<const float*>(smem + Cfg::xs + xb * (Cfg::kXF * 1024)) + (f - f0) * 256;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const int xi = p - 4 + j;             // conv1 position p-2 reads x[p-4 .. p-2]
            xd[r][j] = pack_dup((valid && xi >= 0 && xi < 128) ? xf[r * 128 + xi] : 0.f);
          }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&x_empty[xb]);
      uint8_t* arow0 = smem + Cfg::a + row * 16;
#ifdef MDC_VT_ABLATE
      long long* ptrace = (blockIdx.x == 0 && pw == 0) ? trace : nullptr;
#else
      constexpr long long* ptrace = nullptr;
#endif
      static_assert(Cfg::kPhases == 2, "dispatch below");
      if (cp == 0) conv240_produce<0>(w1c, xd, m, arow0, full, empty, k, prev, lane, rank, dbg, ptrace);
      else conv240_produce<1>(w1c, xd, m, arow0, full, empty, k, prev, lane, rank, dbg, ptrace);
    }
    if (prev != 0xFFFFFFFFu) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&full[prev]);
        else mbar_arrive_remote(&full[prev], 0);
      }
    }
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
              mma_bf16_ss(tmem + m * 256, desc64(a_lo + ((m * 16384 + ks * 32) >> 4), hi),
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

// [rows][kVtFlat] K-major matrix: box = 128 B of K (64 bf16 / 32 fp32) x box_rows, 128B swizzle, OOB rows -> 0
static int make_kmajor_map(CUtensorMap* m, const void* base, uint64_t rows, bool f32, uint32_t box_rows) {
  PFN_cuTensorMapEncodeTiled enc = get_encode();
  MDC_REQUIRE(enc != nullptr, MDC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t dims[2] = {(cuuint64_t)kVtFlat, rows};
  const cuuint64_t strides[1] = {(cuuint64_t)kVtFlat * (f32 ? 4 : 2)};
  const cuuint32_t box[2] = {(cuuint32_t)(f32 ? kTK : kDK), box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                         const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MDC_REQUIRE(r == CUDA_SUCCESS, MDC_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return MDC_OK;
}

static void split_tf32(float v, float& hi, float& lo) {   // hi = what the tensor core reads of v; lo exact
  uint32_t u;
  memcpy(&u, &v, 4);
  u &= 0xFFFFE000u;
  memcpy(&hi, &u, 4);
  lo = v - hi;
}

int pack_vt_bf16(mdc_handle_s* h) {      // both tensor-core modes (MDC_MODE_BF16, MDC_MODE_TF32X3)
  const bool tf32 = h->mode == MDC_MODE_TF32X3;
  if (int e = pack_vt_small(h)) return e;
  // conv1 image: 32 channel groups x {w0[8], w1[8], w2[8], bias[8]} fp32 (Keras (1,3,1,256) = [tap][ch])
  {
    std::vector<float> img(32 * 32);
    const float* w1 = h->w[MDC_T_CONV1_K].data();
    const float* b1 = h->w[MDC_T_CONV1_B].data();
    for (int g = 0; g < 32; ++g)
      for (int e = 0; e < 8; ++e) {
        const int ch = g * 8 + e;
        img[g * 32 + e] = w1[ch];
        img[g * 32 + 8 + e] = w1[256 + ch];
        img[g * 32 + 16 + e] = w1[512 + ch];
        img[g * 32 + 24 + e] = b1[ch];
      }
    h->vt_w1_img = img;      // passed by value as a kernel parameter
  }
  // conv2 image: [chunk][pair rank][hi/lo][tap][group][40 out][16 B]; chunk c = conv1 channels
  // kCC c .. kCC c + kCC - 1, group g = 2*(channel block) + input row, rank h holds output channels
  // 40h..40h+39 (Keras (2,3,256,80) = [r][j][ch][o]).  bf16: 8 values per group, one image;
  // tf32x3: 4 fp32 per group, hi and lo images.
  const float* w2 = h->w[MDC_T_CONV2_K].data();
  if (!tf32) {
    using Cfg = ConvCfg<false>;
    std::vector<uint16_t> img((size_t)Cfg::kChunks * 2 * Cfg::kBSlot / 2);
    for (int c = 0; c < Cfg::kChunks; ++c)
      for (int hf = 0; hf < 2; ++hf)
        for (int j = 0; j < 3; ++j)
          for (int g = 0; g < kGroups; ++g)
            for (int oo = 0; oo < kBHalf; ++oo)
              for (int e = 0; e < 8; ++e) {
                const int r = g & 1, ch = c * Cfg::kCC + (g >> 1) * 8 + e, o = hf * kBHalf + oo;
                img[((size_t)(c * 2 + hf) * (Cfg::kBSlot / 2)) + ((size_t)(j * kGroups + g) * kBHalf + oo) * 8 + e] =
                    f2bf(w2[((size_t)(r * 3 + j) * 256 + ch) * 80 + o]);
              }
    if (int e = h->vt_w2_bf16.reserve(img.size() * 2)) return e;
    MDC_CUDA(cudaMemcpy(h->vt_w2_bf16.ptr, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
    // N240 image: [pair rank][chunk][group][120 accumulator columns][8 k]; column n = 120 rank + row = tap*80 + o
    std::vector<uint16_t> img2((size_t)2 * Conv240::kBBytes / 2);
    for (int hf = 0; hf < 2; ++hf)
      for (int c = 0; c < Conv240::kChunks; ++c)
        for (int g = 0; g < kGroups; ++g)
          for (int nr = 0; nr < Conv240::kBRows; ++nr)
            for (int e = 0; e < 8; ++e) {
              const int ncol = hf * Conv240::kBRows + nr, j = ncol / 80, o = ncol % 80;
              const int r = g & 1, ch = c * 16 + (g >> 1) * 8 + e;
              img2[(((size_t)(hf * Conv240::kChunks + c) * kGroups + g) * Conv240::kBRows + nr) * 8 + e] =
                  f2bf(w2[((size_t)(r * 3 + j) * 256 + ch) * 80 + o]);
            }
    if (int e = h->vt_w2_n240.reserve(img2.size() * 2)) return e;
    MDC_CUDA(cudaMemcpy(h->vt_w2_n240.ptr, img2.data(), img2.size() * 2, cudaMemcpyHostToDevice));
  } else {
    using Cfg = ConvCfg<true>;
    std::vector<float> img((size_t)Cfg::kChunks * 2 * Cfg::kBSlot / 4);
    for (int c = 0; c < Cfg::kChunks; ++c)
      for (int hf = 0; hf < 2; ++hf)
        for (int j = 0; j < 3; ++j)
          for (int g = 0; g < kGroups; ++g)
            for (int oo = 0; oo < kBHalf; ++oo)
              for (int e = 0; e < 4; ++e) {
                const int r = g & 1, ch = c * Cfg::kCC + (g >> 1) * 4 + e, o = hf * kBHalf + oo;
                float hi, lo;
                split_tf32(w2[((size_t)(r * 3 + j) * 256 + ch) * 80 + o], hi, lo);
                const size_t base = (size_t)(c * 2 + hf) * (Cfg::kBSlot / 4) + ((size_t)(j * kGroups + g) * kBHalf + oo) * 4 + e;
                img[base] = hi;
                img[base + Cfg::kBImg / 4] = lo;
              }
    if (int e = h->vt_w2_bf16.reserve(img.size() * 4)) return e;
    MDC_CUDA(cudaMemcpy(h->vt_w2_bf16.ptr, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
  }
  // dense1 image: W3^T [256][10560] in this library's activation order (pos*80 + ch);
  // bf16: one bf16 matrix; tf32x3: fp32 hi matrix followed by the lo matrix
  {
    std::vector<float> w3p;
    vt_permute_w3(h, w3p);
    const size_t cnt = (size_t)kVtH * kVtFlat;
    if (!h->tmap_w3) h->tmap_w3 = aligned_alloc(64, 2 * sizeof(CUtensorMap));
    CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(h->tmap_w3);
    if (!tf32) {
      std::vector<uint16_t> img(cnt);
      for (int kx = 0; kx < kVtFlat; ++kx)
        for (int o = 0; o < kVtH; ++o) img[(size_t)o * kVtFlat + kx] = f2bf(w3p[(size_t)kx * kVtH + o]);
      if (int e = h->vt_w3_bf16.reserve(cnt * 2)) return e;
      MDC_CUDA(cudaMemcpy(h->vt_w3_bf16.ptr, img.data(), cnt * 2, cudaMemcpyHostToDevice));
      if (int e = make_kmajor_map(&maps[0], h->vt_w3_bf16.ptr, kVtH, false, kDM)) return e;
    } else {
      std::vector<float> img(2 * cnt);
      for (int kx = 0; kx < kVtFlat; ++kx)
        for (int o = 0; o < kVtH; ++o)
          split_tf32(w3p[(size_t)kx * kVtH + o], img[(size_t)o * kVtFlat + kx], img[cnt + (size_t)o * kVtFlat + kx]);
      if (int e = h->vt_w3_bf16.reserve(2 * cnt * 4)) return e;
      MDC_CUDA(cudaMemcpy(h->vt_w3_bf16.ptr, img.data(), 2 * cnt * 4, cudaMemcpyHostToDevice));
      const float* base = reinterpret_cast<const float*>(h->vt_w3_bf16.ptr);
      if (int e = make_kmajor_map(&maps[0], base, kVtH, true, 128)) return e;      // half a block per CTA of the pair
      if (int e = make_kmajor_map(&maps[1], base + cnt, kVtH, true, 128)) return e;
    }
  }
  MDC_CUDA(cudaFuncSetAttribute(vt_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<false>::total));
  MDC_CUDA(cudaFuncSetAttribute(vt_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<true>::total));
  MDC_CUDA(cudaFuncSetAttribute(vt_conv240_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv240::total));
  MDC_CUDA(cudaFuncSetAttribute(vt_dense_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DenseT32Smem::total));
  return MDC_OK;
}

// frames per pass.  bf16: act = 21 KB/frame -> 1.38 GB; tf32x3: hi + lo fp32 = 84 KB/frame, and
// 148 x 128 frames is exactly one wave of dense tiles -> 1.6 GB
int64_t vt_pass_frames(const mdc_handle_s* h) {
  return h->mode == MDC_MODE_TF32X3 ? (int64_t)h->num_sms * kTM : 65536;
}

int vt_reserve(mdc_handle_s* h, int64_t frames) {
  const bool tf32 = h->mode == MDC_MODE_TF32X3;
  const size_t act_elems = (size_t)frames * kVtFlat;
  if (act_elems > h->vt_act_elems) {
    if (int e = h->ws_act.reserve(tf32 ? act_elems * 8 : act_elems * 2)) return e;
    if (int e = h->ws_h.reserve((size_t)frames * kVtH * 4)) return e;
    h->vt_act_elems = act_elems;
  }
  return MDC_OK;
}

// conv1 + conv2 of m frames at x -> activations of frames [frame_offset, frame_offset + m) of the pass
int launch_vt_conv(mdc_handle_s* h, const float* x, int64_t m, int64_t frame_offset, cudaStream_t stream) {
  const bool tf32 = h->mode == MDC_MODE_TF32X3;
  if (m == 0) return MDC_OK;
  const size_t off = (size_t)frame_offset * kVtFlat;
  void* act0 = tf32 ? (void*)(reinterpret_cast<float*>(h->ws_act.ptr) + off)
                    : (void*)(reinterpret_cast<__nv_bfloat16*>(h->ws_act.ptr) + off);
  void* act1 = tf32 ? (void*)(reinterpret_cast<float*>(h->ws_act.ptr) + h->vt_act_elems + off) : nullptr;
  ConvW1 w1c;
  static_assert(sizeof(ConvW1) == 32 * 32 * sizeof(float), "conv1 image size");
  memcpy(&w1c, h->vt_w1_img.data(), sizeof(w1c));
  static const int dbg = getenv("MDC_VT_DEBUG") ? atoi(getenv("MDC_VT_DEBUG")) : 0;   // timing experiments
  // MDC_VT_CONV=n240 selects the experimental second bf16 formulation (taps as N) for A/B timing: it is
  // numerically identical but slower (2.25 ms against 1.35 ms per 65,536 frames), see DESIGN.md section 5.1
  static const bool use_n240 = getenv("MDC_VT_CONV") && !strcmp(getenv("MDC_VT_CONV"), "n240");
  if (!tf32 && use_n240) {
    const long long num_tiles = (m * 132 + Conv240::kOutRows - 1) / Conv240::kOutRows;
    const long long pairs_needed = (num_tiles + 1) / 2, pairs_max = h->num_sms / 2;
    const unsigned grid_c = 2u * (unsigned)(pairs_needed < pairs_max ? pairs_needed : pairs_max);
    long long* trace = nullptr;
#ifdef MDC_VT_ABLATE
    static long long* trace_buf = nullptr;
    if (!trace_buf) { cudaMalloc(&trace_buf, 4 * 512 * 8); }
    cudaMemsetAsync(trace_buf, 0, 4 * 512 * 8, stream);
    trace = trace_buf;
#endif
    prof_begin(h, stream);
    vt_conv240_kernel<<<grid_c, Conv240::kThreads, Conv240::total, stream>>>(
        w1c, x, m, reinterpret_cast<const float*>(h->vt_b2.ptr), reinterpret_cast<const uint8_t*>(h->vt_w2_n240.ptr),
        reinterpret_cast<__nv_bfloat16*>(act0), num_tiles, dbg, trace);
    prof_end(h, stream);
#ifdef MDC_VT_ABLATE
    if (trace && getenv("MDC_VT_TRACE")) {
      std::vector<long long> t(4 * 512);
      cudaStreamSynchronize(stream);
      cudaMemcpy(t.data(), trace, t.size() * 8, cudaMemcpyDeviceToHost);
      if (FILE* f = fopen(getenv("MDC_VT_TRACE"), "w")) {
        for (size_t i = 0; i < t.size(); ++i) fprintf(f, "%zu %lld\n", i, t[i]);
        fclose(f);
      }
    }
#endif
    h->launches += 1;
    MDC_CUDA(cudaGetLastError());
    return MDC_OK;
  }
  const long long out_rows = tf32 ? ConvCfg<true>::kOutRows : ConvCfg<false>::kOutRows;
  const long long num_st = (m * 132 + out_rows - 1) / out_rows;
  const long long pairs_needed = (num_st + 1) / 2, pairs_max = h->num_sms / 2;
  const unsigned grid_c = 2u * (unsigned)(pairs_needed < pairs_max ? pairs_needed : pairs_max);
  const float* b2 = reinterpret_cast<const float*>(h->vt_b2.ptr);
  const uint8_t* w2 = reinterpret_cast<const uint8_t*>(h->vt_w2_bf16.ptr);
  long long* trace3 = nullptr;
#ifdef MDC_VT_ABLATE
  static long long* trace3_buf = nullptr;
  if (!trace3_buf) cudaMalloc(&trace3_buf, 4 * 512 * 8);
  cudaMemsetAsync(trace3_buf, 0, 4 * 512 * 8, stream);
  trace3 = trace3_buf;
#endif
  prof_begin(h, stream);
  if (tf32)
    vt_conv_kernel<true><<<grid_c, ConvCfg<true>::kThreads, ConvCfg<true>::total, stream>>>(w1c, x, m, b2, w2, act0, act1, num_st, dbg, trace3);
  else
    vt_conv_kernel<false><<<grid_c, ConvCfg<false>::kThreads, ConvCfg<false>::total, stream>>>(w1c, x, m, b2, w2, act0, act1, num_st, dbg, trace3);
  prof_end(h, stream);
#ifdef MDC_VT_ABLATE
  if (trace3 && getenv("MDC_VT_TRACE")) {
    std::vector<long long> t(4 * 512);
    cudaStreamSynchronize(stream);
    cudaMemcpy(t.data(), trace3, t.size() * 8, cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(getenv("MDC_VT_TRACE"), "w")) {
      for (size_t i = 0; i < t.size(); ++i) fprintf(f, "%zu %lld\n", i, t[i]);
      fclose(f);
    }
  }
#endif
  h->launches += 1;
  MDC_CUDA(cudaGetLastError());
  return MDC_OK;
}

template <int C>
static int dense_bf16_launch(mdc_handle_s* h, unsigned grid, cudaStream_t stream, const CUtensorMap& map_a,
                             const CUtensorMap& map_b, const float* b3, float* hb, int64_t m, int tiles, float* probs,
                             float* dense, int32_t* cls, unsigned long long* hist) {
  static bool attr_set[16] = {};                 // per device
  if (!attr_set[h->device & 15]) {
    MDC_CUDA(cudaFuncSetAttribute(vt_dense_bf16_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, DenseSmem::total));
    attr_set[h->device & 15] = true;
  }
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
  const bool tf32 = h->mode == MDC_MODE_TF32X3;
  if (m == 0) return MDC_OK;
  float* hb = reinterpret_cast<float*>(h->ws_h.ptr);
  const CUtensorMap* wmaps = reinterpret_cast<const CUtensorMap*>(h->tmap_w3);
  const float* b3 = reinterpret_cast<const float*>(h->vt_b3.ptr);
  if (!tf32) {
    CUtensorMap map_a;
    if (int e = make_kmajor_map(&map_a, h->ws_act.ptr, (uint64_t)m, false, kDM)) return e;
    const int tiles = (int)((m + kDM - 1) / kDM);
    const unsigned grid_d = (unsigned)(tiles < h->num_sms ? tiles : h->num_sms);
    static const bool keep_h = getenv("MDC_VT_KEEP_H") != nullptr;       // debugging: also store dense1 activations
    if (int e = dense_bf16_dispatch(h, grid_d, stream, map_a, wmaps[0], b3, keep_h ? hb : nullptr, m, tiles, probs, dense,
                                    cls, hist))
      return e;
    h->launches += 1;
    MDC_CUDA(cudaGetLastError());
    return MDC_OK;
  } else {
    CUtensorMap map_ah, map_al;
    const float* act = reinterpret_cast<const float*>(h->ws_act.ptr);
    if (int e = make_kmajor_map(&map_ah, act, (uint64_t)m, true, kTM)) return e;
    if (int e = make_kmajor_map(&map_al, act + h->vt_act_elems, (uint64_t)m, true, kTM)) return e;
    const int tiles = (int)((m + kTM - 1) / kTM);
    const int pairs = (tiles + 1) / 2, pairs_max = h->num_sms / 2;
    const unsigned grid_d = 2u * (unsigned)(pairs < pairs_max ? pairs : pairs_max);
    vt_dense_tf32x3_kernel<<<grid_d, kDenseT32Threads, DenseT32Smem::total, stream>>>(map_ah, map_al, wmaps[0], wmaps[1],
                                                                                       b3, hb, m, tiles);
  }
  h->launches += 1;
  MDC_CUDA(cudaGetLastError());
  return launch_vt_head(h, hb, m, probs, dense, cls, hist, stream);
}

int launch_vt_bf16(mdc_handle_s* h, const float* x, int64_t n, float* probs, float* dense,
                   int32_t* cls, unsigned long long* hist, cudaStream_t stream) {
  if (n == 0) return MDC_OK;
  const int64_t CH = vt_pass_frames(h);
  if (int e = vt_reserve(h, n < CH ? n : CH)) return e;
  for (int64_t s = 0; s < n; s += CH) {
    const int64_t m = (n - s) < CH ? (n - s) : CH;
    if (int e = launch_vt_conv(h, x + s * 256, m, 0, stream)) return e;
    if (int e = launch_vt_dense_head(h, m, probs ? probs + s * h->C : nullptr, dense ? dense + s * h->C : nullptr,
                                     cls ? cls + s : nullptr, hist, stream))
      return e;
  }
  return MDC_OK;
}

}  // namespace mdc
