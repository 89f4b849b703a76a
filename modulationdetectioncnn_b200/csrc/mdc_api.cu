// C-ABI entry points of libmdc.so (see include/mdc.h) and the host-buffer pipeline.
#include <string.h>

#include <emmintrin.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "mdc_internal.cuh"

namespace mdc {

// Makes the handle's device current for the duration of an entry point and restores the caller's device afterwards
// (a library call must not move the calling thread's later torch / CUDA allocations to another GPU).
struct DeviceGuard {
  int prev = -1, dev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) : dev(device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
};

struct PinnedBuffer {
  void* ptr = nullptr;
  size_t bytes = 0;
  int reserve(size_t n) {          // grow-only
    if (n <= bytes) return MDC_OK;
    if (ptr) {
      MDC_CUDA(cudaFreeHost(ptr));
      ptr = nullptr;
      bytes = 0;
    }
    MDC_CUDA(cudaHostAlloc(&ptr, n, cudaHostAllocDefault));
    bytes = n;
    return MDC_OK;
  }
  void release() {
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    bytes = 0;
  }
};

// Copy into the pinned staging ring with non-temporal stores: the CPU never reads the staging buffer again (the copy
// engine does), so write-allocating its cache lines only costs memory bandwidth - a read-for-ownership per line, a third
// of the traffic - which is what eight ranks staging at once run out of first.  (glibc's memcpy switches to streaming
// stores only far above the 4-16 MiB slices copied here.)
static void stream_copy(void* dst, const void* src, size_t n) {
  char* d = static_cast<char*>(dst);
  const char* s = static_cast<const char*>(src);
  if ((reinterpret_cast<uintptr_t>(d) & 15) != 0 || n < 4096) {
    memcpy(dst, src, n);
    return;
  }
  size_t i = 0;
  for (; i + 64 <= n; i += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 32));
    const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 48), e);
  }
  _mm_sfence();
  if (i < n) memcpy(d + i, s + i, n - i);
}

// A few persistent threads that copy a caller's pageable buffer into the pinned staging ring, one slice each (a single
// core's memcpy is slower than the PCIe link it feeds).  One pool per process, started on first use.
class CopyPool {
 public:
  static CopyPool& get() {
    static CopyPool p;
    return p;
  }
  // streaming: dst is a staging buffer only the copy engine will read (non-temporal stores)
  void copy(void* dst, const void* src, size_t bytes, bool streaming) {
    const size_t parts = bytes < ((size_t)1 << 20) ? 1 : threads_.size() + 1;
    if (parts == 1) {
      if (streaming) stream_copy(dst, src, bytes);
      else memcpy(dst, src, bytes);
      return;
    }
    const size_t step = ((bytes / parts) + 4095) & ~(size_t)4095;
    size_t off = step;
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (; off < bytes; off += step)
        jobs_.push_back({(char*)dst + off, (const char*)src + off, std::min(step, bytes - off), streaming});
      pending_ += jobs_.size();
    }
    cv_.notify_all();
    if (streaming) stream_copy(dst, src, std::min(step, bytes));  // the caller's own slice
    else memcpy(dst, src, std::min(step, bytes));
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  struct Job {
    char* dst;
    const char* src;
    size_t n;
    bool streaming;
  };
  CopyPool() {
    unsigned hw = std::thread::hardware_concurrency();
    int n = getenv("MDC_COPY_THREADS") ? atoi(getenv("MDC_COPY_THREADS")) - 1 : (int)std::min(3u, hw / 4);
    for (int i = 0; i < n; ++i) threads_.emplace_back([this] { work(); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  void work() {
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
      cv_.wait(lk, [this] { return stop_ || !jobs_.empty(); });
      if (stop_) return;
      Job j = jobs_.back();
      jobs_.pop_back();
      lk.unlock();
      if (j.streaming) stream_copy(j.dst, j.src, j.n);
      else memcpy(j.dst, j.src, j.n);
      lk.lock();
      if (--pending_ == 0) done_.notify_all();
    }
  }
  std::vector<std::thread> threads_;
  std::vector<Job> jobs_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  size_t pending_ = 0;
  bool stop_ = false;
};

// true for ordinary (pageable, unregistered) host memory - what numpy hands over
static bool is_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int DeviceBuffer::reserve(size_t n) {
  if (n <= bytes) return MDC_OK;
  if (ptr) {
    MDC_CUDA(cudaFree(ptr));
    ptr = nullptr;
    bytes = 0;
  }
  n = (n + 255) & ~(size_t)255;
  MDC_CUDA(cudaMalloc(&ptr, n));
  bytes = n;
  return MDC_OK;
}
void DeviceBuffer::release() {
  if (ptr) cudaFree(ptr);
  ptr = nullptr;
  bytes = 0;
}

void prof_begin(mdc_handle_s* h, cudaStream_t s) {
  if (!h->prof.on) return;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a, s);
  h->prof.pending.emplace_back(a, b);
}
void prof_end(mdc_handle_s* h, cudaStream_t s) {
  if (!h->prof.on || h->prof.pending.empty()) return;
  cudaEventRecord(h->prof.pending.back().second, s);
}

// ------------------------------------------------------------------ host pipeline
// Three streams (H2D, compute, D2H) over a ring of device slots, so that the copy of chunk
// i+1, the kernels of chunk i and the read-back of chunk i-1 overlap.
struct HostPipe {
  static constexpr int kSlots = 3;
  cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
  cudaEvent_t e_h2d[kSlots] = {}, e_comp[kSlots] = {}, e_d2h[kSlots] = {};
  cudaEvent_t e_pass[2] = {}, e_pass_d2h[2] = {};      // tensor-core VT path: per-pass output buffers
  // asynchronous calls: slot / pass counters run on across calls (a later call's first chunks must not overwrite
  // slots an earlier call's tail still reads), one completion event per call
  static constexpr int kTickets = 16;
  int64_t chunk_seq = 0, pass_seq = 0, ticket_seq = 0;
  cudaEvent_t e_ticket[kTickets] = {}, e_hist = nullptr;
  DeviceBuffer x[kSlots], o0[kSlots], o1[kSlots], o2[kSlots];
  DeviceBuffer hist;
  PinnedBuffer pin[kSlots];      // staging for pageable caller buffers (slot k is free again once e_h2d[k] has fired)
  PinnedBuffer pout[3];          // blocking calls: pinned staging for pageable output arrays
  PinnedBuffer range;            // u32 [kTickets]: F16X3 range flags as read back at the end of each call
  int init() {
    if (s_h2d) return MDC_OK;
    if (int e = range.reserve(kTickets * sizeof(unsigned int))) return e;
    memset(range.ptr, 0, kTickets * sizeof(unsigned int));
    MDC_CUDA(cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking));
    MDC_CUDA(cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking));
    MDC_CUDA(cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < kSlots; ++i) {
      MDC_CUDA(cudaEventCreateWithFlags(&e_h2d[i], cudaEventDisableTiming));
      MDC_CUDA(cudaEventCreateWithFlags(&e_comp[i], cudaEventDisableTiming));
      MDC_CUDA(cudaEventCreateWithFlags(&e_d2h[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 2; ++i) {
      MDC_CUDA(cudaEventCreateWithFlags(&e_pass[i], cudaEventDisableTiming));
      MDC_CUDA(cudaEventCreateWithFlags(&e_pass_d2h[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < kTickets; ++i) MDC_CUDA(cudaEventCreateWithFlags(&e_ticket[i], cudaEventDisableTiming));
    MDC_CUDA(cudaEventCreateWithFlags(&e_hist, cudaEventDisableTiming));
    return hist.reserve(kMaxClasses * kMaxClasses * sizeof(unsigned long long));
  }
  void destroy() {
    if (!s_h2d) return;
    cudaStreamSynchronize(s_h2d); cudaStreamSynchronize(s_comp); cudaStreamSynchronize(s_d2h);
    for (int i = 0; i < kSlots; ++i) {
      cudaEventDestroy(e_h2d[i]); cudaEventDestroy(e_comp[i]); cudaEventDestroy(e_d2h[i]);
      x[i].release(); o0[i].release(); o1[i].release(); o2[i].release(); pin[i].release();
    }
    for (auto& b : pout) b.release();
    range.release();
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(e_pass[i]); cudaEventDestroy(e_pass_d2h[i]); }
    for (int i = 0; i < kTickets; ++i) cudaEventDestroy(e_ticket[i]);
    cudaEventDestroy(e_hist);
    hist.release();
    cudaStreamDestroy(s_h2d); cudaStreamDestroy(s_comp); cudaStreamDestroy(s_d2h);
    s_h2d = nullptr;
  }
};

void destroy_pipe(mdc_handle_s* h) {
  if (h->pipe) {
    h->pipe->destroy();
    delete h->pipe;
    h->pipe = nullptr;
  }
}

static int ensure_packed(mdc_handle_s* h) {
  if (h->packed) return MDC_OK;
  // the packed images are rewritten with blocking copies on the legacy stream, which does not order against the
  // handle's non-blocking pipeline streams or the caller's: let every earlier prediction finish with the old weights
  MDC_CUDA(cudaDeviceSynchronize());
  const int need_tiny[] = {MDC_T_CONV1_K, MDC_T_CONV1_B, MDC_T_DENSE1_K, MDC_T_DENSE1_B};
  if (h->model == MDC_MODEL_TINY) {
    for (int t : need_tiny)
      MDC_REQUIRE(h->have[t], MDC_ERR_NOT_READY, "weights tensor %d not set (call mdc_set_weights_f32)", t);
    if (int e = pack_tiny(h)) return e;
  } else {
    for (int t = 0; t < 8; ++t)
      MDC_REQUIRE(h->have[t], MDC_ERR_NOT_READY, "weights tensor %d not set (call mdc_set_weights_f32)", t);
    if (h->mode == MDC_MODE_FP32) {
      if (int e = pack_vt_f32(h)) return e;
    } else {
      if (int e = pack_vt_bf16(h)) return e;
    }
  }
  h->packed = true;
  return MDC_OK;
}

static size_t tensor_count(const mdc_handle_s* h, int id) {
  if (h->model == MDC_MODEL_TINY) {
    switch (id) {
      case MDC_T_CONV1_K: return (size_t)2 * h->F;
      case MDC_T_CONV1_B: return (size_t)h->F;
      case MDC_T_DENSE1_K: return (size_t)2 * 129 * h->F * h->C;
      case MDC_T_DENSE1_B: return (size_t)h->C;
      default: return 0;
    }
  }
  switch (id) {
    case MDC_T_CONV1_K: return 3 * 256;
    case MDC_T_CONV1_B: return 256;
    case MDC_T_CONV2_K: return (size_t)2 * 3 * 256 * 80;
    case MDC_T_CONV2_B: return 80;
    case MDC_T_DENSE1_K: return (size_t)kVtFlat * kVtH;
    case MDC_T_DENSE1_B: return kVtH;
    case MDC_T_DENSE2_K: return (size_t)kVtH * h->C;
    case MDC_T_DENSE2_B: return (size_t)h->C;
    default: return 0;
  }
}

static bool vt_tensor_mode(const mdc_handle_s* h) {
  return h->model == MDC_MODEL_VT && h->mode != MDC_MODE_FP32;
}

static size_t frame_bytes(int in_fmt) { return in_fmt == MDC_IN_U8IQ ? 256 : (in_fmt == MDC_IN_I16 ? 512 : 1024); }

static int check_format(const mdc_handle_s* h, int in_fmt) {
  MDC_REQUIRE(in_fmt == MDC_IN_F32 || in_fmt == MDC_IN_U8IQ || in_fmt == MDC_IN_I16, MDC_ERR_INVALID,
              "unknown frame format %d", in_fmt);
  const bool tiny_special = h->model == MDC_MODEL_TINY && (h->F == 3 || h->F == 10) && h->C == 3;
  MDC_REQUIRE(in_fmt == MDC_IN_F32 || vt_tensor_mode(h) || tiny_special, MDC_ERR_UNSUPPORTED,
              "raw u8 / int16 frames are read by the VT-CNN2 tensor-core kernels (BF16, F16X3, TF32X3) and the "
              "specialised TinyCNN2 kernels (F in {3, 10}, C = 3) only; use mdc_sdr_ingest_u8 in front of this model");
  return MDC_OK;
}

static int check_q612_format(int in_fmt) {
  MDC_REQUIRE(in_fmt == MDC_IN_I32 || in_fmt == MDC_IN_I16 || in_fmt == MDC_IN_U8IQ, MDC_ERR_INVALID,
              "frame format %d: the integer model takes MDC_IN_I32, MDC_IN_I16 or MDC_IN_U8IQ", in_fmt);
  return MDC_OK;
}

static int predict_f32_dev(mdc_handle_s* h, const void* x, int in_fmt, int64_t n, float* probs, float* dense,
                           int32_t* cls, unsigned long long* hist, cudaStream_t stream) {
  if (h->model == MDC_MODEL_TINY) return launch_tiny_f32(h, x, in_fmt, n, probs, dense, cls, hist, stream);
  if (h->mode == MDC_MODE_FP32) return launch_vt_f32(h, (const float*)x, n, probs, dense, cls, hist, stream);
  return launch_vt_bf16(h, x, in_fmt, n, probs, dense, cls, hist, stream);
}

}  // namespace mdc

using namespace mdc;

// ---- host-buffer variants --------------------------------------------------------------
// One H2D chunk.  Pinned caller memory is copied directly; pageable memory (what numpy hands over, cnn.py:198) goes
// through pinned slot k first - cudaMemcpyAsync from pageable memory would stage it on ONE driver thread and block.
static int h2d_chunk(HostPipe& P, int k, void* dst_dev, const void* src, size_t bytes, bool pageable) {
  if (pageable) {
    if (int e = P.pin[k].reserve(bytes)) return e;
    MDC_CUDA(cudaEventSynchronize(P.e_h2d[k]));       // the previous copy out of this pinned slot has finished
    CopyPool::get().copy(P.pin[k].ptr, src, bytes, true);
    src = P.pin[k].ptr;
  }
  MDC_CUDA(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, P.s_h2d));
  MDC_CUDA(cudaEventRecord(P.e_h2d[k], P.s_h2d));
  return MDC_OK;
}

// End of a host call: histogram (and, F16X3, the range flags) back on the compute stream - after this call's last
// kernel, before the next call's memset - then the completion event on the D2H stream.
// ticket == NULL: synchronous (returns with the outputs in host memory); otherwise returns after enqueueing and
// *ticket names the completion event (mdc_host_wait).
static int finish_host_call(mdc_handle_s* h, HostPipe& P, unsigned long long* hist, int64_t* ticket) {
  const int64_t t = ++P.ticket_seq;
  unsigned int* range_word = reinterpret_cast<unsigned int*>(P.range.ptr) + (t % HostPipe::kTickets);
  if (hist)
    MDC_CUDA(cudaMemcpyAsync(hist, P.hist.ptr, h->C * sizeof(unsigned long long), cudaMemcpyDeviceToHost, P.s_comp));
  if (h->mode == MDC_MODE_F16X3) {
    MDC_CUDA(cudaMemcpyAsync(range_word, h->vt_flags.ptr, sizeof(unsigned int), cudaMemcpyDeviceToHost, P.s_comp));
    MDC_CUDA(cudaMemsetAsync(h->vt_flags.ptr, 0, sizeof(unsigned int), P.s_comp));
  } else {
    *range_word = 0;
  }
  MDC_CUDA(cudaEventRecord(P.e_hist, P.s_comp));
  MDC_CUDA(cudaStreamWaitEvent(P.s_d2h, P.e_hist, 0));
  MDC_CUDA(cudaEventRecord(P.e_ticket[t % HostPipe::kTickets], P.s_d2h));
  if (ticket) {
    *ticket = t;
    return MDC_OK;
  }
  MDC_CUDA(cudaEventSynchronize(P.e_ticket[t % HostPipe::kTickets]));
  MDC_REQUIRE(*range_word == 0, MDC_ERR_RANGE,
              "MDC_MODE_F16X3: an input or activation left the fp16 range; rerun this batch on an MDC_MODE_TF32X3 handle");
  return MDC_OK;
}

// Blocking calls with pageable OUTPUT arrays: a device-to-pageable cudaMemcpyAsync blocks the calling thread until the
// stream gets there, which would stop the enqueue loop at every chunk.  The results go to pinned staging instead and
// are copied to the caller's arrays after the final synchronisation.
struct OutStage {
  struct Item { void* user; void* pinned; size_t bytes; };
  Item items[3];
  int count = 0;
  template <class T>
  int redirect(HostPipe& P, T*& ptr, size_t elems) {
    if (!ptr || !is_pageable(ptr)) return MDC_OK;
    const size_t bytes = elems * sizeof(T);
    if (int e = P.pout[count].reserve(bytes)) return e;
    items[count] = {ptr, P.pout[count].ptr, bytes};
    ptr = reinterpret_cast<T*>(P.pout[count].ptr);
    ++count;
    return MDC_OK;
  }
  void deliver() {
    for (int i = 0; i < count; ++i) CopyPool::get().copy(items[i].user, items[i].pinned, items[i].bytes, false);
  }
};

// Slot counters run on across calls, so a later call's first chunks wait for the slots an earlier call's tail
// still uses.
template <class O0, class O1, class Launch>
static int run_host_pipeline(mdc_handle_s* h, const void* xv, size_t fb, int64_t n, O0* o0, O1* o1, int32_t* cls,
                             unsigned long long* hist, int64_t chunk, int64_t* ticket, Launch launch) {
  const uint8_t* x = reinterpret_cast<const uint8_t*>(xv);     // fb bytes per frame
  if (!h->pipe) h->pipe = new HostPipe();
  HostPipe& P = *h->pipe;
  if (int e = P.init()) return e;
  const int C = h->C;
  constexpr int S = HostPipe::kSlots;
  const bool pageable = is_pageable(x);
  OutStage stage;
  if (!ticket) {
    if (int e = stage.redirect(P, o0, (size_t)n * C)) return e;
    if (int e = stage.redirect(P, o1, (size_t)n * C)) return e;
    if (int e = stage.redirect(P, cls, (size_t)n)) return e;
  }
  if (hist) MDC_CUDA(cudaMemsetAsync(P.hist.ptr, 0, C * sizeof(unsigned long long), P.s_comp));
  for (int k = 0; k < S; ++k) {
    if (int e = P.x[k].reserve((size_t)chunk * fb)) return e;
    if (o0) if (int e = P.o0[k].reserve((size_t)chunk * C * sizeof(O0))) return e;
    if (o1) if (int e = P.o1[k].reserve((size_t)chunk * C * sizeof(O1))) return e;
    if (cls) if (int e = P.o2[k].reserve((size_t)chunk * sizeof(int32_t))) return e;
  }
  int64_t& i = P.chunk_seq;
  for (int64_t s = 0; s < n; s += chunk, ++i) {
    const int k = (int)(i % S);
    const int64_t m = (n - s) < chunk ? (n - s) : chunk;
    MDC_CUDA(cudaStreamWaitEvent(P.s_h2d, P.e_comp[k], 0));        // (a never-recorded event does not block)
    if (int e = h2d_chunk(P, k, P.x[k].ptr, x + (size_t)s * fb, (size_t)m * fb, pageable)) return e;
    MDC_CUDA(cudaStreamWaitEvent(P.s_comp, P.e_h2d[k], 0));
    MDC_CUDA(cudaStreamWaitEvent(P.s_comp, P.e_d2h[k], 0));
    if (int e = launch((const void*)P.x[k].ptr, m, o0 ? (O0*)P.o0[k].ptr : nullptr,
                       o1 ? (O1*)P.o1[k].ptr : nullptr, cls ? (int32_t*)P.o2[k].ptr : nullptr,
                       hist ? (unsigned long long*)P.hist.ptr : nullptr, P.s_comp))
      return e;
    MDC_CUDA(cudaEventRecord(P.e_comp[k], P.s_comp));
    MDC_CUDA(cudaStreamWaitEvent(P.s_d2h, P.e_comp[k], 0));
    if (o0) MDC_CUDA(cudaMemcpyAsync(o0 + s * C, P.o0[k].ptr, (size_t)m * C * sizeof(O0), cudaMemcpyDeviceToHost, P.s_d2h));
    if (o1) MDC_CUDA(cudaMemcpyAsync(o1 + s * C, P.o1[k].ptr, (size_t)m * C * sizeof(O1), cudaMemcpyDeviceToHost, P.s_d2h));
    if (cls) MDC_CUDA(cudaMemcpyAsync(cls + s, P.o2[k].ptr, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, P.s_d2h));
    MDC_CUDA(cudaEventRecord(P.e_d2h[k], P.s_d2h));
  }
  const int rc = finish_host_call(h, P, hist, ticket);
  if (rc == MDC_OK || rc == MDC_ERR_RANGE) stage.deliver();
  return rc;
}


// Tensor-core VT-CNN2 (bf16 / fp16x3 / 3xTF32) from host buffers, frames in any MDC_IN_* format.  The frames travel in
// small chunks (8 MiB of f32) so the first convolution starts early and every later copy hides under the previous
// chunk's convolution; dense1 and the head run ONCE per pass over all the activations (a dense tile is 128 / 256
// frames per SM: per-chunk launches would leave most SMs idle), and the pass's results are copied back under the
// next pass.
static int run_vt_host_pipeline(mdc_handle_s* h, const void* xv, int in_fmt, int64_t n, float* probs, float* dense,
                                int32_t* cls, unsigned long long* hist, int64_t* ticket) {
  if (!h->pipe) h->pipe = new HostPipe();
  HostPipe& P = *h->pipe;
  if (int e = P.init()) return e;
  const int C = h->C;
  constexpr int S = HostPipe::kSlots;
  const uint8_t* x = reinterpret_cast<const uint8_t*>(xv);
  const size_t fb = frame_bytes(in_fmt);
  const bool pageable = is_pageable(xv);
  OutStage stage;
  if (!ticket) {
    if (int e = stage.redirect(P, probs, (size_t)n * C)) return e;
    if (int e = stage.redirect(P, dense, (size_t)n * C)) return e;
    if (int e = stage.redirect(P, cls, (size_t)n)) return e;
  }
  // bf16, blocking call: passes of 32,768 frames (128 dense tiles, one wave) so that dense1 of the first half runs
  // under the second half's copies and convolutions, 8 MiB chunks after a 2 MiB first one (nothing overlaps the very
  // first copy).  Streaming call: the tail of this call runs under the next call's copies anyway, so one dense1
  // launch per 65,536 frames and 16 MiB chunks (fewer ~30 us launch prologues) win: measured per 65,536 frames,
  // blocking / streaming: 1.92 / 1.74 ms with the first setting, 2.04 / 1.70 ms with the second.
  // fp16x3: the kernels take three times as long as the copies, one 65,536-frame pass either way.
  // (MDC_VT_PASS / MDC_VT_CHUNK / MDC_VT_FIRST: tuning aids, override both)
  const bool streaming = ticket != nullptr;
  static const int64_t ov_pass = getenv("MDC_VT_PASS") ? atoll(getenv("MDC_VT_PASS")) : 0;
  static const int64_t ov_chunk = getenv("MDC_VT_CHUNK") ? atoll(getenv("MDC_VT_CHUNK")) : 0;
  static const int64_t ov_first = getenv("MDC_VT_FIRST") ? atoll(getenv("MDC_VT_FIRST")) : 0;
  const bool one_pass = streaming || h->mode == MDC_MODE_F16X3;
  const int64_t env_pass = ov_pass ? ov_pass : (one_pass ? 65536 : 32768);
  const int64_t env_chunk = ov_chunk ? ov_chunk : (streaming ? 16384 : 8192);
  const int64_t env_first = ov_first ? ov_first : (streaming ? 16384 : 2048);
  const bool tf32 = h->mode == MDC_MODE_TF32X3;
  const int64_t pass = tf32 ? vt_pass_frames(h) : env_pass;
  const int64_t chunk = tf32 ? pass : (env_chunk < pass ? env_chunk : pass);
  const int64_t cap = n < pass ? n : pass;
  if (int e = vt_reserve(h, cap)) return e;
  if (hist) MDC_CUDA(cudaMemsetAsync(P.hist.ptr, 0, C * sizeof(unsigned long long), P.s_comp));
  for (int k = 0; k < S; ++k)
    if (int e = P.x[k].reserve((size_t)(cap < chunk ? cap : chunk) * fb)) return e;
  for (int b = 0; b < 2; ++b) {
    if (probs) if (int e = P.o0[b].reserve((size_t)cap * C * sizeof(float))) return e;
    if (dense) if (int e = P.o1[b].reserve((size_t)cap * C * sizeof(float))) return e;
    if (cls) if (int e = P.o2[b].reserve((size_t)cap * sizeof(int32_t))) return e;
  }
  int64_t& i = P.chunk_seq;
  int64_t& pi = P.pass_seq;
  for (int64_t p0 = 0; p0 < n; p0 += pass, ++pi) {
    const int64_t pm = (n - p0) < pass ? (n - p0) : pass;
    const int ob = (int)(pi & 1);
    for (int64_t c0 = 0, step = 0; c0 < pm; c0 += step, ++i) {
      const int k = (int)(i % S);
      // the very first chunk is small (2 MiB): nothing can overlap its copy
      step = (p0 == 0 && c0 == 0 && !tf32 && pm > 2 * env_first) ? env_first : chunk;
      const int64_t m = (pm - c0) < step ? (pm - c0) : step;
      MDC_CUDA(cudaStreamWaitEvent(P.s_h2d, P.e_comp[k], 0));     // (a never-recorded event does not block)
      if (int e = h2d_chunk(P, k, P.x[k].ptr, x + (size_t)(p0 + c0) * fb, (size_t)m * fb, pageable)) return e;
      MDC_CUDA(cudaStreamWaitEvent(P.s_comp, P.e_h2d[k], 0));
      if (int e = launch_vt_conv(h, P.x[k].ptr, in_fmt, m, c0, P.s_comp)) return e;
      MDC_CUDA(cudaEventRecord(P.e_comp[k], P.s_comp));
    }
    MDC_CUDA(cudaStreamWaitEvent(P.s_comp, P.e_pass_d2h[ob], 0));
    if (int e = launch_vt_dense_head(h, pm, probs ? (float*)P.o0[ob].ptr : nullptr, dense ? (float*)P.o1[ob].ptr : nullptr,
                                     cls ? (int32_t*)P.o2[ob].ptr : nullptr,
                                     hist ? (unsigned long long*)P.hist.ptr : nullptr, P.s_comp))
      return e;
    MDC_CUDA(cudaEventRecord(P.e_pass[ob], P.s_comp));
    MDC_CUDA(cudaStreamWaitEvent(P.s_d2h, P.e_pass[ob], 0));
    if (probs) MDC_CUDA(cudaMemcpyAsync(probs + p0 * C, P.o0[ob].ptr, (size_t)pm * C * sizeof(float), cudaMemcpyDeviceToHost, P.s_d2h));
    if (dense) MDC_CUDA(cudaMemcpyAsync(dense + p0 * C, P.o1[ob].ptr, (size_t)pm * C * sizeof(float), cudaMemcpyDeviceToHost, P.s_d2h));
    if (cls) MDC_CUDA(cudaMemcpyAsync(cls + p0, P.o2[ob].ptr, (size_t)pm * sizeof(int32_t), cudaMemcpyDeviceToHost, P.s_d2h));
    MDC_CUDA(cudaEventRecord(P.e_pass_d2h[ob], P.s_d2h));
  }
  const int rc = finish_host_call(h, P, hist, ticket);
  if (rc == MDC_OK || rc == MDC_ERR_RANGE) stage.deliver();
  return rc;
}

// ---- confusion matrix (optionally one per group, e.g. per SNR) -----------------------------
// conf[g][t][p] += 1 for every frame; counters are block-local in shared memory when they fit, then added
// to the global matrix with one atomic per non-zero cell.
__global__ void confusion_kernel(const int* __restrict__ t, const int* __restrict__ p, const int* __restrict__ grp,
                                 long long n, int C, int G, int use_smem, unsigned long long* __restrict__ conf) {
  extern __shared__ unsigned int sm[];
  const int cells = G * C * C;
  if (use_smem) {
    for (int i = threadIdx.x; i < cells; i += blockDim.x) sm[i] = 0;
    __syncthreads();
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int a = t[i], b = p[i], g = grp ? grp[i] : 0;
    if (a >= 0 && a < C && b >= 0 && b < C && g >= 0 && g < G) {
      const int cell = (g * C + a) * C + b;
      if (use_smem) atomicAdd(&sm[cell], 1u);
      else atomicAdd(conf + cell, 1ull);
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < cells; i += blockDim.x)
      if (sm[i]) atomicAdd(conf + i, (unsigned long long)sm[i]);
  }
}

#define MDC_CHECK_HANDLE(h)                                                   \
  MDC_REQUIRE((h) != nullptr, MDC_ERR_INVALID, "null handle");                \
  mdc::DeviceGuard mdc_device_guard_((h)->device);                            \
  MDC_CUDA(mdc_device_guard_.err)

extern "C" {

const char* mdc_last_error(void) { return g_err; }
const char* mdc_version(void) { return "libmdc 0.2.0 (sm_100a)"; }

int mdc_create(int model_kind, int filters, int classes, int mode, int device, mdc_handle_t* out) {
  MDC_REQUIRE(out != nullptr, MDC_ERR_INVALID, "out is NULL");
  *out = nullptr;
  MDC_REQUIRE(model_kind == MDC_MODEL_TINY || model_kind == MDC_MODEL_VT, MDC_ERR_INVALID,
              "unknown model kind %d", model_kind);
  MDC_REQUIRE(classes >= 1 && classes <= kMaxClasses, MDC_ERR_UNSUPPORTED, "classes=%d outside 1..%d",
              classes, kMaxClasses);
  if (model_kind == MDC_MODEL_TINY) {
    MDC_REQUIRE(filters >= 1 && filters <= kMaxFilters, MDC_ERR_UNSUPPORTED, "filters=%d outside 1..%d",
                filters, kMaxFilters);
    MDC_REQUIRE(mode == MDC_MODE_FP32 || mode == MDC_MODE_Q612, MDC_ERR_INVALID,
                "TinyCNN2 supports MDC_MODE_FP32 and MDC_MODE_Q612, not mode %d", mode);
  } else {
    MDC_REQUIRE(mode == MDC_MODE_FP32 || mode == MDC_MODE_BF16 || mode == MDC_MODE_TF32X3 || mode == MDC_MODE_F16X3,
                MDC_ERR_INVALID,
                "VT-CNN2 supports MDC_MODE_FP32, MDC_MODE_BF16, MDC_MODE_F16X3 and MDC_MODE_TF32X3, not mode %d", mode);
  }
  int ndev = 0;
  MDC_CUDA(cudaGetDeviceCount(&ndev));
  MDC_REQUIRE(device >= 0 && device < ndev, MDC_ERR_INVALID, "device %d not in [0,%d)", device, ndev);
  DeviceGuard guard(device);
  MDC_CUDA(guard.err);
  cudaDeviceProp prop;
  MDC_CUDA(cudaGetDeviceProperties(&prop, device));
  MDC_REQUIRE(prop.major == 10, MDC_ERR_UNSUPPORTED,
              "device %d is sm_%d%d; libmdc is built for sm_100a only", device, prop.major, prop.minor);
  mdc_handle_s* h = new mdc_handle_s();
  h->model = model_kind;
  h->F = model_kind == MDC_MODEL_TINY ? filters : 0;
  h->C = classes;
  h->mode = mode;
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->dominant_kernel = model_kind == MDC_MODEL_TINY
                           ? (mode == MDC_MODE_Q612 ? "q612_kernel" : "tiny_f32_kernel")
                           : (mode == MDC_MODE_FP32 ? "sgemm_bias_act_kernel(conv2)"
                                                   : (mode == MDC_MODE_BF16 ? "vt_conv_kernel<bf16>"
                                                      : (mode == MDC_MODE_F16X3 ? "vt_conv_kernel<f16x3>" : "vt_conv_kernel<tf32x3>")));
  *out = h;
  return MDC_OK;
}

int mdc_destroy(mdc_handle_t h) {
  if (!h) return MDC_OK;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  destroy_pipe(h);
  for (auto& pr : h->prof.pending) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  DeviceBuffer* bufs[] = {&h->tiny_conv, &h->tiny_dense, &h->tiny_bias, &h->vt_w1, &h->vt_b1, &h->vt_w2,
                          &h->vt_b2, &h->vt_w3, &h->vt_b3, &h->vt_w4, &h->vt_b4, &h->vt_w2_bf16,
                          &h->vt_w3_bf16, &h->vt_flags, &h->ws_a1, &h->ws_act, &h->ws_h, &h->q_dense};
  for (DeviceBuffer* b : bufs) b->release();
  free(h->tmap_w3);
  delete h;
  return MDC_OK;
}

int mdc_set_option(mdc_handle_t h, int option, int value) {
  MDC_REQUIRE(h != nullptr, MDC_ERR_INVALID, "null handle");
  if (option == MDC_OPT_FLATTEN_ORDER) {
    MDC_REQUIRE(value == 0 || value == 1, MDC_ERR_INVALID, "flatten order must be 0 or 1");
    h->flatten_order = value;
    h->packed = false;
    return MDC_OK;
  }
  set_error("unknown option %d", option);
  return MDC_ERR_INVALID;
}

int mdc_set_weights_f32(mdc_handle_t h, int tensor_id, const float* host_ptr, size_t count) {
  MDC_REQUIRE(h != nullptr && host_ptr != nullptr, MDC_ERR_INVALID, "null argument");
  MDC_REQUIRE(h->mode != MDC_MODE_Q612, MDC_ERR_INVALID,
              "this handle is in Q6.12 mode: use mdc_set_weights_q612");
  MDC_REQUIRE(tensor_id >= 0 && tensor_id < 8, MDC_ERR_INVALID, "tensor id %d", tensor_id);
  const size_t want = tensor_count(h, tensor_id);
  MDC_REQUIRE(want != 0, MDC_ERR_INVALID, "tensor id %d does not exist for this model", tensor_id);
  MDC_REQUIRE(count == want, MDC_ERR_INVALID, "tensor %d: expected %zu elements, got %zu", tensor_id, want,
              count);
  h->w[tensor_id].assign(host_ptr, host_ptr + count);
  h->have[tensor_id] = true;
  h->packed = false;
  return MDC_OK;
}

int mdc_set_weights_q612(mdc_handle_t h, const int32_t* conv_tab, const int32_t* dense_bias,
                         const int32_t* dense_tabs) {
  MDC_CHECK_HANDLE(h);
  MDC_REQUIRE(h->model == MDC_MODEL_TINY && h->mode == MDC_MODE_Q612, MDC_ERR_INVALID,
              "handle is not a TinyCNN2 Q6.12 handle");
  MDC_REQUIRE(conv_tab && dense_bias && dense_tabs, MDC_ERR_INVALID, "null argument");
  const int F = h->F, C = h->C;
  auto in18 = [](int v) { return v >= -(1 << 17) && v < (1 << 17); };
  for (int i = 0; i < 3 * F; ++i)
    MDC_REQUIRE(in18(conv_tab[i]), MDC_ERR_INVALID, "conv_tab[%d]=%d is not an 18-bit signed value", i, conv_tab[i]);
  for (int i = 0; i < C; ++i)
    MDC_REQUIRE(in18(dense_bias[i]), MDC_ERR_INVALID, "dense_bias[%d]=%d is not an 18-bit signed value", i, dense_bias[i]);
  const size_t tab = (size_t)129 * F;
  for (size_t i = 0; i < 2 * C * tab; ++i)
    MDC_REQUIRE(in18(dense_tabs[i]), MDC_ERR_INVALID, "dense_tabs[%zu]=%d is not an 18-bit signed value", i, dense_tabs[i]);
  h->q_conv_host.assign(conv_tab, conv_tab + 3 * F);
  h->q_bias_host.assign(dense_bias, dense_bias + C);
  // pre-skew: image[f][c][iq][s] = 8 * tab[2c+iq][128 f + max(s-1,0)]   (dense_layer, sv:336,351-378;
  // the factor 8 positions the product for the kernel's slice36, see q612.cu)
  std::vector<int> img((size_t)F * C * 2 * 128);
  for (int f = 0; f < F; ++f)
    for (int c = 0; c < C; ++c)
      for (int iq = 0; iq < 2; ++iq)
        for (int s = 0; s < 128; ++s)
          img[(((size_t)f * C + c) * 2 + iq) * 128 + s] =
              8 * dense_tabs[(size_t)(2 * c + iq) * tab + 128 * f + (s > 0 ? s - 1 : 0)];
  // Input bound for the kernel's 32-bit path.  With |x| <= X: conv sums |m| <= X*csum, so |m| < 2^28 needs
  // X <= (2^28-1)/csum; then |slice| <= X*csum/4096 + 1 and, as long as that plus the largest |bias| stays below
  // ylim <= 2^16-2, the bias add cannot wrap and every conv output obeys y <= ylim, where ylim = (2^28-1)/dsum
  // keeps the dense sums |yI*wI + yQ*wQ| <= ylim*dsum < 2^28 too.
  {
    long long csum = 0, bmax = 0, dsum = 0;
    for (int f = 0; f < F; ++f) {
      csum = std::max(csum, (long long)std::abs(conv_tab[3 * f]) + std::abs(conv_tab[3 * f + 1]));
      bmax = std::max(bmax, (long long)std::abs(conv_tab[3 * f + 2]));
    }
    for (int c = 0; c < C; ++c)
      for (size_t a = 0; a < tab; ++a)
        dsum = std::max(dsum, (long long)std::abs(dense_tabs[(size_t)(2 * c) * tab + a]) + std::abs(dense_tabs[(size_t)(2 * c + 1) * tab + a]));
    const long long lim = (1ll << 28) - 1, top = (1ll << 17) - 1;
    const long long xlim = csum ? lim / csum : top;
    // (2^16 - 2: the conv bias is folded into the 32-bit accumulator as bias * 2^15, see conv18_sel)
    const long long ylim = std::min(dsum ? lim / dsum : top, (1ll << 16) - 2);
    long long xfast = -1;
    if (ylim > bmax + 1) xfast = std::min(xlim, csum ? ((ylim - bmax - 1) * 4096) / csum : top);
    h->q_xfast = (int)std::min(xfast, 1ll << 17);
  }
  // same ordering rule as ensure_packed: predictions already enqueued (the pipeline streams are non-blocking)
  // finish with the old ROM image before it is overwritten
  MDC_CUDA(cudaDeviceSynchronize());
  if (int e = h->q_dense.reserve(img.size() * sizeof(int))) return e;
  MDC_CUDA(cudaMemcpy(h->q_dense.ptr, img.data(), img.size() * sizeof(int), cudaMemcpyHostToDevice));
  h->have_q = true;
  return MDC_OK;
}

int mdc_predict_raw(mdc_handle_t h, const void* x_dev, int in_format, int64_t n, float* probs_dev, float* dense_dev,
                    int32_t* cls_dev, unsigned long long* hist_dev, void* stream) {
  MDC_CHECK_HANDLE(h);
  MDC_REQUIRE(h->mode != MDC_MODE_Q612, MDC_ERR_INVALID, "Q6.12 handle: use mdc_predict_q612");
  MDC_REQUIRE(n >= 0, MDC_ERR_INVALID, "n=%lld < 0", (long long)n);
  MDC_REQUIRE(n == 0 || x_dev != nullptr, MDC_ERR_INVALID, "x_dev is NULL");
  MDC_REQUIRE(((uintptr_t)x_dev & 15) == 0, MDC_ERR_INVALID, "x_dev must be 16-byte aligned");
  if (int e = check_format(h, in_format)) return e;
  if (int e = ensure_packed(h)) return e;
  return predict_f32_dev(h, x_dev, in_format, n, probs_dev, dense_dev, cls_dev, hist_dev, (cudaStream_t)stream);
}

int mdc_predict_f32(mdc_handle_t h, const float* x_dev, int64_t n, float* probs_dev, float* dense_dev,
                    int32_t* cls_dev, unsigned long long* hist_dev, void* stream) {
  return mdc_predict_raw(h, x_dev, MDC_IN_F32, n, probs_dev, dense_dev, cls_dev, hist_dev, stream);
}

int mdc_reserve(mdc_handle_t h, int64_t max_frames) {
  MDC_CHECK_HANDLE(h);
  MDC_REQUIRE(max_frames >= 0, MDC_ERR_INVALID, "max_frames=%lld < 0", (long long)max_frames);
  if (h->mode == MDC_MODE_Q612) return MDC_OK;            // the integer kernel needs no work space
  if (int e = ensure_packed(h)) return e;
  if (vt_tensor_mode(h)) {
    const int64_t pass = vt_pass_frames(h);
    MDC_CUDA(cudaDeviceSynchronize());                    // growing the work space frees the old one
    return vt_reserve(h, max_frames < pass ? max_frames : pass);
  }
  return MDC_OK;
}

int mdc_range_flags(mdc_handle_t h, unsigned int* flags, int reset) {
  MDC_CHECK_HANDLE(h);
  MDC_REQUIRE(flags != nullptr, MDC_ERR_INVALID, "flags is NULL");
  *flags = 0;
  if (h->mode != MDC_MODE_F16X3 || !h->vt_flags.ptr) return MDC_OK;
  MDC_CUDA(cudaDeviceSynchronize());
  MDC_CUDA(cudaMemcpy(flags, h->vt_flags.ptr, sizeof(unsigned int), cudaMemcpyDeviceToHost));
  if (reset) MDC_CUDA(cudaMemset(h->vt_flags.ptr, 0, sizeof(unsigned int)));
  return MDC_OK;
}

int mdc_predict_q612_raw(mdc_handle_t h, const void* x_dev, int in_format, int64_t n, int32_t* out_dev, int32_t* pre_dev,
                         int32_t* cls_dev, unsigned long long* hist_dev, void* stream) {
  MDC_CHECK_HANDLE(h);
  MDC_REQUIRE(h->mode == MDC_MODE_Q612, MDC_ERR_INVALID, "handle is not in Q6.12 mode");
  MDC_REQUIRE(h->have_q, MDC_ERR_NOT_READY, "ROM tables not set (call mdc_set_weights_q612)");
  MDC_REQUIRE(n >= 0, MDC_ERR_INVALID, "n=%lld < 0", (long long)n);
  MDC_REQUIRE(n == 0 || x_dev != nullptr, MDC_ERR_INVALID, "x_dev is NULL");
  MDC_REQUIRE(((uintptr_t)x_dev & 15) == 0, MDC_ERR_INVALID, "x_dev must be 16-byte aligned");
  if (int e = check_q612_format(in_format)) return e;
  return launch_q612(h, x_dev, in_format, n, out_dev, pre_dev, cls_dev, hist_dev, (cudaStream_t)stream);
}

int mdc_predict_q612(mdc_handle_t h, const int32_t* x_dev, int64_t n, int32_t* out_dev, int32_t* pre_dev,
                     int32_t* cls_dev, unsigned long long* hist_dev, void* stream) {
  return mdc_predict_q612_raw(h, x_dev, MDC_IN_I32, n, out_dev, pre_dev, cls_dev, hist_dev, stream);
}

static int predict_f32_host_impl(mdc_handle_t h, const void* x_host, int in_fmt, int64_t n, float* probs_host,
                                 float* dense_host, int32_t* cls_host, unsigned long long* hist_host, int64_t* ticket) {
  MDC_CHECK_HANDLE(h);
  MDC_REQUIRE(h->mode != MDC_MODE_Q612, MDC_ERR_INVALID, "Q6.12 handle: use mdc_predict_q612_host");
  MDC_REQUIRE(n >= 0 && (n == 0 || x_host), MDC_ERR_INVALID, "bad x_host / n");
  if (int e = check_format(h, in_fmt)) return e;
  if (int e = ensure_packed(h)) return e;
  if (n == 0) {
    if (hist_host) memset(hist_host, 0, h->C * sizeof(unsigned long long));
    return MDC_OK;
  }
  if (vt_tensor_mode(h))
    return run_vt_host_pipeline(h, x_host, in_fmt, n, probs_host, dense_host, cls_host, hist_host, ticket);
  static const int64_t chunk = getenv("MDC_HOST_CHUNK") ? atoll(getenv("MDC_HOST_CHUNK")) : 16384;   // frames per chunk
  return run_host_pipeline<float, float>(
      h, x_host, frame_bytes(in_fmt), n, probs_host, dense_host, cls_host, hist_host, chunk, ticket,
      [h, in_fmt](const void* x, int64_t m, float* p, float* d, int32_t* c, unsigned long long* hs, cudaStream_t s) {
        return predict_f32_dev(h, x, in_fmt, m, p, d, c, hs, s);
      });
}

static int predict_q612_host_impl(mdc_handle_t h, const void* x_host, int in_fmt, int64_t n, int32_t* out_host,
                                  int32_t* pre_host, int32_t* cls_host, unsigned long long* hist_host, int64_t* ticket) {
  MDC_CHECK_HANDLE(h);
  MDC_REQUIRE(h->mode == MDC_MODE_Q612, MDC_ERR_INVALID, "handle is not in Q6.12 mode");
  MDC_REQUIRE(h->have_q, MDC_ERR_NOT_READY, "ROM tables not set (call mdc_set_weights_q612)");
  MDC_REQUIRE(n >= 0 && (n == 0 || x_host), MDC_ERR_INVALID, "bad x_host / n");
  if (int e = check_q612_format(in_fmt)) return e;
  if (n == 0) {
    if (hist_host) memset(hist_host, 0, h->C * sizeof(unsigned long long));
    return MDC_OK;
  }
  static const int64_t chunk = getenv("MDC_HOST_CHUNK") ? atoll(getenv("MDC_HOST_CHUNK")) : 16384;   // frames per chunk
  return run_host_pipeline<int32_t, int32_t>(
      h, x_host, frame_bytes(in_fmt), n, out_host, pre_host, cls_host, hist_host, chunk, ticket,
      [h, in_fmt](const void* x, int64_t m, int32_t* o, int32_t* p, int32_t* c, unsigned long long* hs, cudaStream_t s) {
        return launch_q612(h, x, in_fmt, m, o, p, c, hs, s);
      });
}

int mdc_predict_f32_host(mdc_handle_t h, const float* x_host, int64_t n, float* probs_host, float* dense_host,
                         int32_t* cls_host, unsigned long long* hist_host) {
  return predict_f32_host_impl(h, x_host, MDC_IN_F32, n, probs_host, dense_host, cls_host, hist_host, nullptr);
}

int mdc_predict_raw_host(mdc_handle_t h, const void* x_host, int in_format, int64_t n, float* probs_host,
                         float* dense_host, int32_t* cls_host, unsigned long long* hist_host) {
  return predict_f32_host_impl(h, x_host, in_format, n, probs_host, dense_host, cls_host, hist_host, nullptr);
}

int mdc_predict_f32_host_async(mdc_handle_t h, const float* x_host, int64_t n, float* probs_host, float* dense_host,
                               int32_t* cls_host, unsigned long long* hist_host, int64_t* ticket) {
  MDC_REQUIRE(ticket != nullptr, MDC_ERR_INVALID, "ticket is NULL");
  *ticket = 0;                                   // 0: nothing pending (e.g. n == 0)
  return predict_f32_host_impl(h, x_host, MDC_IN_F32, n, probs_host, dense_host, cls_host, hist_host, ticket);
}

int mdc_predict_raw_host_async(mdc_handle_t h, const void* x_host, int in_format, int64_t n, float* probs_host,
                               float* dense_host, int32_t* cls_host, unsigned long long* hist_host, int64_t* ticket) {
  MDC_REQUIRE(ticket != nullptr, MDC_ERR_INVALID, "ticket is NULL");
  *ticket = 0;
  return predict_f32_host_impl(h, x_host, in_format, n, probs_host, dense_host, cls_host, hist_host, ticket);
}

int mdc_predict_q612_raw_host_async(mdc_handle_t h, const void* x_host, int in_format, int64_t n, int32_t* out_host,
                                    int32_t* pre_host, int32_t* cls_host, unsigned long long* hist_host, int64_t* ticket) {
  MDC_REQUIRE(ticket != nullptr, MDC_ERR_INVALID, "ticket is NULL");
  *ticket = 0;
  return predict_q612_host_impl(h, x_host, in_format, n, out_host, pre_host, cls_host, hist_host, ticket);
}

int mdc_predict_q612_host_async(mdc_handle_t h, const int32_t* x_host, int64_t n, int32_t* out_host, int32_t* pre_host,
                                int32_t* cls_host, unsigned long long* hist_host, int64_t* ticket) {
  return mdc_predict_q612_raw_host_async(h, x_host, MDC_IN_I32, n, out_host, pre_host, cls_host, hist_host, ticket);
}

int mdc_host_wait(mdc_handle_t h, int64_t ticket) {
  MDC_CHECK_HANDLE(h);
  if (ticket <= 0 || !h->pipe) return MDC_OK;
  HostPipe& P = *h->pipe;
  MDC_REQUIRE(ticket <= P.ticket_seq, MDC_ERR_INVALID, "ticket %lld was never issued", (long long)ticket);
  // completion events fire in issue order; a ticket whose event slot has been reused is covered by the newest one
  const int64_t t = (P.ticket_seq - ticket >= HostPipe::kTickets) ? P.ticket_seq : ticket;
  MDC_CUDA(cudaEventSynchronize(P.e_ticket[t % HostPipe::kTickets]));
  MDC_REQUIRE(reinterpret_cast<unsigned int*>(P.range.ptr)[t % HostPipe::kTickets] == 0, MDC_ERR_RANGE,
              "MDC_MODE_F16X3: an input or activation left the fp16 range; rerun this batch on an MDC_MODE_TF32X3 handle");
  return MDC_OK;
}

int mdc_predict_q612_raw_host(mdc_handle_t h, const void* x_host, int in_format, int64_t n, int32_t* out_host,
                              int32_t* pre_host, int32_t* cls_host, unsigned long long* hist_host) {
  return predict_q612_host_impl(h, x_host, in_format, n, out_host, pre_host, cls_host, hist_host, nullptr);
}

int mdc_predict_q612_host(mdc_handle_t h, const int32_t* x_host, int64_t n, int32_t* out_host, int32_t* pre_host,
                          int32_t* cls_host, unsigned long long* hist_host) {
  return predict_q612_host_impl(h, x_host, MDC_IN_I32, n, out_host, pre_host, cls_host, hist_host, nullptr);
}

// ---- FWHT ------------------------------------------------------------------------------
int mdc_fwht_i32(const int32_t* in_dev, int32_t* out_dev, int64_t n_spectra, int log2_npt, int ordering,
                 void* stream) {
  MDC_REQUIRE(log2_npt >= 5 && log2_npt <= 13, MDC_ERR_UNSUPPORTED, "log2_npt=%d outside 5..13", log2_npt);
  MDC_REQUIRE(ordering == MDC_FWHT_NATURAL || ordering == MDC_FWHT_SEQUENCY, MDC_ERR_INVALID, "ordering %d", ordering);
  MDC_REQUIRE(n_spectra >= 0, MDC_ERR_INVALID, "n_spectra < 0");
  MDC_REQUIRE(n_spectra == 0 || (in_dev && out_dev), MDC_ERR_INVALID, "null buffer");
  MDC_REQUIRE((((uintptr_t)in_dev | (uintptr_t)out_dev) & 15) == 0, MDC_ERR_INVALID, "buffers must be 16-byte aligned");
  {
    // in place (in == out) is fine: a spectrum is read completely by the warp / block that then writes it.  Any
    // other overlap would let one spectrum's stores land in another's unread input.
    const uintptr_t a = (uintptr_t)in_dev, b = (uintptr_t)out_dev;
    const uintptr_t bytes = (uintptr_t)n_spectra << (log2_npt + 2);
    MDC_REQUIRE(a == b || a + bytes <= b || b + bytes <= a, MDC_ERR_INVALID,
                "in_dev and out_dev overlap partially (identical or disjoint buffers only)");
  }
  return launch_fwht(in_dev, out_dev, n_spectra, log2_npt, ordering, (cudaStream_t)stream);
}

// streams and device slots of the FWHT host pipeline, created once per device and reused (a
// cudaMalloc/cudaFree pair per call costs more than the transform)
namespace {
struct FwhtPipe {
  static constexpr int S = 3;
  cudaStream_t st[S] = {};
  void* buf[S] = {};
  size_t bytes = 0;
  bool ready = false;
};
std::mutex g_fwht_mu;
FwhtPipe g_fwht_pipe[16];
}  // namespace

int mdc_fwht_i32_host(const int32_t* in_host, int32_t* out_host, int64_t n_spectra, int log2_npt, int ordering,
                      int device) {
  MDC_REQUIRE(log2_npt >= 5 && log2_npt <= 13, MDC_ERR_UNSUPPORTED, "log2_npt=%d outside 5..13", log2_npt);
  MDC_REQUIRE(ordering == MDC_FWHT_NATURAL || ordering == MDC_FWHT_SEQUENCY, MDC_ERR_INVALID, "ordering %d", ordering);
  MDC_REQUIRE(n_spectra >= 0 && (n_spectra == 0 || (in_host && out_host)), MDC_ERR_INVALID, "bad buffers");
  MDC_REQUIRE(device >= 0 && device < 16, MDC_ERR_INVALID, "device %d", device);
  if (n_spectra == 0) return MDC_OK;
  DeviceGuard guard(device);
  MDC_CUDA(guard.err);
  std::lock_guard<std::mutex> lock(g_fwht_mu);
  FwhtPipe& P = g_fwht_pipe[device];
  constexpr int S = FwhtPipe::S;
  const size_t N = (size_t)1 << log2_npt;
  const size_t slot_bytes = (size_t)16 << 20;                       // 16 MiB per slot
  const int64_t chunk = (int64_t)(slot_bytes / (N * 4));
  if (!P.ready) {
    for (int k = 0; k < S; ++k) {
      MDC_CUDA(cudaStreamCreateWithFlags(&P.st[k], cudaStreamNonBlocking));
      MDC_CUDA(cudaMalloc(&P.buf[k], slot_bytes));
    }
    P.bytes = slot_bytes;
    P.ready = true;
  }
  int rc = MDC_OK;
  int64_t i = 0;
  for (int64_t s = 0; s < n_spectra && rc == MDC_OK; s += chunk, ++i) {
    const int k = (int)(i % S);
    const int64_t m = (n_spectra - s) < chunk ? (n_spectra - s) : chunk;
    // a slot's stream is in-order: H2D -> kernel -> D2H; the three slots overlap each other
    MDC_CUDA(cudaMemcpyAsync(P.buf[k], in_host + s * N, (size_t)m * N * 4, cudaMemcpyHostToDevice, P.st[k]));
    rc = launch_fwht((const int32_t*)P.buf[k], (int32_t*)P.buf[k], m, log2_npt, ordering, P.st[k]);
    MDC_CUDA(cudaMemcpyAsync(out_host + s * N, P.buf[k], (size_t)m * N * 4, cudaMemcpyDeviceToHost, P.st[k]));
  }
  for (int k = 0; k < S; ++k) {
    cudaError_t e = cudaStreamSynchronize(P.st[k]);
    if (e != cudaSuccess && rc == MDC_OK) {
      set_error("fwht host pipeline: %s", cudaGetErrorString(e));
      rc = MDC_ERR_CUDA;
    }
  }
  return rc;
}

int mdc_sdr_ingest_u8(const uint8_t* iq_dev, int64_t n_samples, float* frames_f32_dev, int32_t* frames_q612_dev,
                      int32_t* fwht_dev, void* stream) {
  MDC_REQUIRE(n_samples >= 0 && n_samples % 128 == 0, MDC_ERR_INVALID, "n_samples=%lld must be a multiple of 128",
              (long long)n_samples);
  MDC_REQUIRE(fwht_dev == nullptr || n_samples % 1024 == 0, MDC_ERR_INVALID,
              "n_samples=%lld must be a multiple of 1024 for FWHT blocks", (long long)n_samples);
  MDC_REQUIRE(n_samples == 0 || iq_dev != nullptr, MDC_ERR_INVALID, "iq_dev is NULL");
  MDC_REQUIRE((((uintptr_t)iq_dev | (uintptr_t)frames_f32_dev | (uintptr_t)frames_q612_dev | (uintptr_t)fwht_dev) & 15) == 0,
              MDC_ERR_INVALID, "buffers must be 16-byte aligned");
  return launch_sdr_ingest(iq_dev, n_samples, frames_f32_dev, frames_q612_dev, fwht_dev, (cudaStream_t)stream);
}

int mdc_confusion_grouped_i32(const int32_t* true_dev, const int32_t* pred_dev, const int32_t* group_dev, int64_t n,
                              int classes, int groups, unsigned long long* conf_dev, void* stream) {
  MDC_REQUIRE(classes >= 1 && classes <= kMaxClasses, MDC_ERR_UNSUPPORTED, "classes=%d", classes);
  MDC_REQUIRE(groups >= 1 && groups <= 4096, MDC_ERR_UNSUPPORTED, "groups=%d outside 1..4096", groups);
  MDC_REQUIRE(n >= 0 && (n == 0 || (true_dev && pred_dev && conf_dev)), MDC_ERR_INVALID, "bad arguments");
  if (n == 0) return MDC_OK;
  const int cells = groups * classes * classes;
  const int use_smem = cells <= 8192;                   // 32 KB of 32-bit block-local counters
  // at most 2^20 increments of a 32-bit shared counter per block
  long long blocks = (n + 256LL * 4096 - 1) / (256LL * 4096);
  if (blocks < 148) blocks = n < 148 * 256 ? 1 : 148;
  confusion_kernel<<<(unsigned)blocks, 256, use_smem ? cells * sizeof(unsigned) : 0, (cudaStream_t)stream>>>(
      true_dev, pred_dev, group_dev, n, classes, groups, use_smem, conf_dev);
  MDC_CUDA(cudaGetLastError());
  return MDC_OK;
}

int mdc_confusion_i32(const int32_t* true_dev, const int32_t* pred_dev, int64_t n, int classes,
                      unsigned long long* conf_dev, void* stream) {
  return mdc_confusion_grouped_i32(true_dev, pred_dev, nullptr, n, classes, 1, conf_dev, stream);
}

// ---- introspection ---------------------------------------------------------------------
int64_t mdc_launch_count(mdc_handle_t h) { return h ? h->launches : 0; }

int mdc_profile_enable(mdc_handle_t h, int on) {
  MDC_REQUIRE(h != nullptr, MDC_ERR_INVALID, "null handle");
  h->prof.on = on != 0;
  return MDC_OK;
}

int mdc_profile_read(mdc_handle_t h, double* ms_total, int64_t* launches, const char** kernel_name) {
  MDC_CHECK_HANDLE(h);
  for (auto& pr : h->prof.pending) {
    MDC_CUDA(cudaEventSynchronize(pr.second));
    float ms = 0.f;
    MDC_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
    h->prof.ms += ms;
    h->prof.launches++;
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  h->prof.pending.clear();
  if (ms_total) *ms_total = h->prof.ms;
  if (launches) *launches = h->prof.launches;
  if (kernel_name) *kernel_name = h->dominant_kernel;
  h->prof.ms = 0.0;
  h->prof.launches = 0;
  return MDC_OK;
}

int mdc_debug_read(mdc_handle_t h, int what, void* host_dst, size_t bytes, size_t* copied) {
  MDC_CHECK_HANDLE(h);
  MDC_REQUIRE(host_dst != nullptr, MDC_ERR_INVALID, "null destination");
  MDC_REQUIRE(what == 0 || what == 1, MDC_ERR_INVALID, "unknown intermediate %d", what);
  const DeviceBuffer& b = what == 0 ? h->ws_act : h->ws_h;
  const size_t nb = bytes < b.bytes ? bytes : b.bytes;
  MDC_CUDA(cudaDeviceSynchronize());
  if (nb) MDC_CUDA(cudaMemcpy(host_dst, b.ptr, nb, cudaMemcpyDeviceToHost));
  if (copied) *copied = nb;
  return MDC_OK;
}

}  // extern "C"
