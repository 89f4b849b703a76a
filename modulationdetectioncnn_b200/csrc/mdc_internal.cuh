// Internal declarations shared by the translation units of libmdc.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/mdc.h"

namespace mdc {

void set_error(const char* fmt, ...);

#define MDC_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      mdc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return MDC_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define MDC_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      mdc::set_error(__VA_ARGS__);    \
      return (code);                  \
    }                                 \
  } while (0)

constexpr int kMaxClasses = 16;
constexpr int kMaxFilters = 16;
constexpr int kFrameElems = 256;   // 2 x 128

// VT-CNN2 fixed geometry (example notebook :194-216)
constexpr int kVtC1 = 256;         // conv1 channels
constexpr int kVtC2 = 80;          // conv2 channels
constexpr int kVtPos1 = 130;       // conv1 output positions
constexpr int kVtPos2 = 132;       // conv2 output positions
constexpr int kVtFlat = kVtPos2 * kVtC2;  // 10560
constexpr int kVtH = 256;          // dense1 width

struct DeviceBuffer {
  void* ptr = nullptr;
  size_t bytes = 0;
  int reserve(size_t n);   // grow-only
  void release();
};

struct Profile {
  bool on = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
  double ms = 0.0;
  int64_t launches = 0;
};

struct HostPipe;  // host_pipeline.cu

}  // namespace mdc

struct mdc_handle_s {
  int model = 0, F = 0, C = 0, mode = 0, device = 0;
  int flatten_order = 0;
  int num_sms = 148;
  int64_t launches = 0;
  mdc::Profile prof;
  const char* dominant_kernel = "";

  // ---- float weights (host copies in Keras layout + packed device images)
  std::vector<float> w[8];
  bool have[8] = {false, false, false, false, false, false, false, false};
  bool packed = false;
  // TinyCNN2 fp32: conv [3F] = {k0,k1,b} per filter; dense [2*129*F*C] Keras order; bias [C]
  mdc::DeviceBuffer tiny_conv, tiny_dense, tiny_bias;
  // VT fp32 path
  mdc::DeviceBuffer vt_w1, vt_b1, vt_w2, vt_b2, vt_w3, vt_b3, vt_w4, vt_b4;
  // VT bf16 path (packed operand images)
  mdc::DeviceBuffer vt_w2_bf16, vt_w3_bf16;   // conv2 / dense1 operand images of the handle's tensor-core mode
  std::vector<float> vt_w1_img;  // conv1 weights as the conv kernel's by-value parameter (4 KB)
  float vt_xlimit = 3.0e38f;     // F16X3: largest |x| for which no conv1 activation can leave the fp16 range
  mdc::DeviceBuffer vt_flags;    // F16X3: u32 range flags the kernels OR into
  void* tmap_w3 = nullptr;      // CUtensorMap (host copy, 128 B)
  // work space
  mdc::DeviceBuffer ws_a1;      // padded conv1 activations (fp32 path)
  mdc::DeviceBuffer ws_act;     // conv2 activations
  size_t vt_act_elems = 0;      // tensor-core paths: activation elements the work space holds per pass
  mdc::DeviceBuffer ws_h;       // dense1 activations (fp32 path)

  // ---- Q6.12 ROM images
  bool have_q = false;
  std::vector<int> q_conv_host, q_bias_host;
  mdc::DeviceBuffer q_dense;    // pre-skewed [f][c][iq][128]
  int q_xfast = -1;             // input bound under which every 36-bit sum provably fits 29 bits (q612.cu)

  // ---- host pipeline (lazy)
  mdc::HostPipe* pipe = nullptr;
};

namespace mdc {

// kernels' launchers: all enqueue on `stream`, bump h->launches
// x holds frames in format in_fmt: MDC_IN_I32 (the 18-bit words of test_table), MDC_IN_I16 or MDC_IN_U8IQ
int launch_q612(mdc_handle_s* h, const void* x, int in_fmt, int64_t n, int32_t* out, int32_t* pre,
                int32_t* cls, unsigned long long* hist, cudaStream_t stream);
// x holds frames in format in_fmt (MDC_IN_*; the raw formats need one of the specialised shapes)
int launch_tiny_f32(mdc_handle_s* h, const void* x, int in_fmt, int64_t n, float* probs, float* dense,
                    int32_t* cls, unsigned long long* hist, cudaStream_t stream);
int launch_vt_f32(mdc_handle_s* h, const float* x, int64_t n, float* probs, float* dense,
                  int32_t* cls, unsigned long long* hist, cudaStream_t stream);
// the tensor-core VT-CNN2 modes; x holds frames in format in_fmt (MDC_IN_*)
int launch_vt_bf16(mdc_handle_s* h, const void* x, int in_fmt, int64_t n, float* probs, float* dense,
                   int32_t* cls, unsigned long long* hist, cudaStream_t stream);
// tensor-core VT-CNN2 path in two stages (the host pipeline copies and convolves chunk by chunk, then runs
// dense1 + head once per pass)
int64_t vt_pass_frames(const mdc_handle_s* h);
int vt_reserve(mdc_handle_s* h, int64_t frames);
int launch_vt_conv(mdc_handle_s* h, const void* x, int in_fmt, int64_t m, int64_t frame_offset, cudaStream_t stream);
int launch_vt_dense_head(mdc_handle_s* h, int64_t m, float* probs, float* dense, int32_t* cls,
                         unsigned long long* hist, cudaStream_t stream);
int pack_tiny(mdc_handle_s* h);
int pack_vt_f32(mdc_handle_s* h);
int pack_vt_bf16(mdc_handle_s* h);
int launch_fwht(const int32_t* in, int32_t* out, int64_t n, int log2_npt, int ordering,
                cudaStream_t stream);
int launch_sdr_ingest(const uint8_t* iq, int64_t n_samples, float* f32, int32_t* q612, int32_t* fwht,
                      cudaStream_t stream);

void prof_begin(mdc_handle_s* h, cudaStream_t s);
void prof_end(mdc_handle_s* h, cudaStream_t s);
void destroy_pipe(mdc_handle_s* h);

}  // namespace mdc
