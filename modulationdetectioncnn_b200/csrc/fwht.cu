// Unnormalised Walsh-Hadamard transform, int32, wrap-around arithmetic (exact).
//
// The reference has no FWHT code (README.md:5 is the only mention); the definition is
// X = H_N x, Sylvester/natural order, optional sequency order.
//
// Warp kernel (N = 128..2048): one warp per spectrum, N/32 values per lane in registers.
// Lane l loads int4 number j at element (32 j + l)*4, so natural index bits are
//   [1:0] = int4 component, [6:2] = lane, [log2N-1:7] = j
// -> the stages over bits 0,1 and >=7 are in-register butterflies, the five stages over
// bits 2..6 are __shfl_xor_sync exchanges.  Loads and stores are fully coalesced 512 B
// per warp instruction.  Sequency order is produced by scattering through shared memory
// and copying out coalesced.
// Block kernel (any N = 32..8192): one CTA per spectrum, butterflies in shared memory.
#include "mdc_internal.cuh"

namespace mdc {

__device__ __forceinline__ int4 ldg_stream_i4(const int4* p) {
  int4 r;
  asm volatile("ld.global.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_i4(int4* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// k such that sequency_perm[k] == j, i.e. k = gray^-1(bitrev(j))
__device__ __forceinline__ unsigned seq_slot(unsigned j, int log2n) {
  unsigned g = __brev(j) >> (32 - log2n);
  g ^= g >> 1; g ^= g >> 2; g ^= g >> 4; g ^= g >> 8; g ^= g >> 16;
  return g;
}

template <int LOG2N, bool SEQ>
__global__ void __launch_bounds__(256)
fwht_warp_kernel(const int4* in, int4* out, long long n) {
  constexpr int N = 1 << LOG2N;
  constexpr int J = N / 128;              // int4 per lane
  extern __shared__ int smem[];           // SEQ only: 8 warps x N ints
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long s = warp; s < n; s += nwarps) {
    int v[J * 4];
    const int4* src = in + s * (N / 4);
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int4 t = ldg_stream_i4(src + j * 32 + lane);
      v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
    // in-register stages: register index bits 0..log2(4J)-1  == natural bits 0,1,7,8,...
#pragma unroll
    for (int h = 1; h < J * 4; h <<= 1) {
#pragma unroll
      for (int i = 0; i < J * 4; ++i) {
        if ((i & h) == 0) {
          const int a = v[i], b = v[i + h];
          v[i] = a + b;
          v[i + h] = a - b;
        }
      }
    }
    // cross-lane stages: natural bits 2..6
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
      const int sgn = (lane & m) ? -1 : 1;
#pragma unroll
      for (int i = 0; i < J * 4; ++i) {
        const int o = __shfl_xor_sync(0xffffffffu, v[i], m);
        v[i] = v[i] * sgn + o;
      }
    }
    int4* dst = out + s * (N / 4);
    if (!SEQ) {
#pragma unroll
      for (int j = 0; j < J; ++j)
        stg_stream_i4(dst + j * 32 + lane, make_int4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
    } else {
      int* sm = smem + wib * N;
#pragma unroll
      for (int j = 0; j < J; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          sm[seq_slot((unsigned)((j * 32 + lane) * 4 + c), LOG2N)] = v[4 * j + c];
      __syncwarp();
#pragma unroll
      for (int j = 0; j < J; ++j)
        stg_stream_i4(dst + j * 32 + lane, reinterpret_cast<const int4*>(sm)[j * 32 + lane]);
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(256)
fwht_block_kernel(const int* in, int* out, long long n, int log2n, int seq) {
  extern __shared__ int smem[];
  const int N = 1 << log2n;
  for (long long s = blockIdx.x; s < n; s += gridDim.x) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) smem[i] = in[s * N + i];
    __syncthreads();
    for (int h = 1; h < N; h <<= 1) {
      for (int t = threadIdx.x; t < N / 2; t += blockDim.x) {
        const int i = ((t & ~(h - 1)) << 1) | (t & (h - 1));
        const int a = smem[i], b = smem[i + h];
        smem[i] = a + b;
        smem[i + h] = a - b;
      }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x)
      out[s * N + (seq ? seq_slot((unsigned)i, log2n) : (unsigned)i)] = smem[i];
    __syncthreads();
  }
}

template <int LOG2N>
static int launch_warp(const int32_t* in, int32_t* out, int64_t n, int ordering, int sms,
                       cudaStream_t stream) {
  const int threads = 256;
  long long blocks = (n * 32 + threads - 1) / threads;
  const long long max_blocks = (long long)sms * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  const int4* i4 = reinterpret_cast<const int4*>(in);
  int4* o4 = reinterpret_cast<int4*>(out);
  if (ordering == MDC_FWHT_SEQUENCY) {
    const size_t smem = (size_t)8 * (1 << LOG2N) * sizeof(int);
    if (smem > 48 * 1024)
      MDC_CUDA(cudaFuncSetAttribute(fwht_warp_kernel<LOG2N, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fwht_warp_kernel<LOG2N, true><<<(unsigned)blocks, threads, smem, stream>>>(i4, o4, n);
  } else {
    fwht_warp_kernel<LOG2N, false><<<(unsigned)blocks, threads, 0, stream>>>(i4, o4, n);
  }
  MDC_CUDA(cudaGetLastError());
  return MDC_OK;
}

int launch_fwht(const int32_t* in, int32_t* out, int64_t n, int log2_npt, int ordering,
                cudaStream_t stream) {
  if (n == 0) return MDC_OK;
  int dev = 0, sms = 148;
  MDC_CUDA(cudaGetDevice(&dev));
  MDC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  switch (log2_npt) {
    case 7: return launch_warp<7>(in, out, n, ordering, sms, stream);
    case 8: return launch_warp<8>(in, out, n, ordering, sms, stream);
    case 9: return launch_warp<9>(in, out, n, ordering, sms, stream);
    case 10: return launch_warp<10>(in, out, n, ordering, sms, stream);
    case 11: return launch_warp<11>(in, out, n, ordering, sms, stream);
    default: break;
  }
  const size_t smem = ((size_t)1 << log2_npt) * sizeof(int);
  long long blocks = n < (long long)sms * 8 ? n : (long long)sms * 8;
  fwht_block_kernel<<<(unsigned)blocks, 256, smem, stream>>>(in, out, n, log2_npt,
                                                             ordering == MDC_FWHT_SEQUENCY);
  MDC_CUDA(cudaGetLastError());
  return MDC_OK;
}

}  // namespace mdc
