/*
 * mdc.h - C ABI of libmdc.so, the B200 (sm_100a) CNN2 / SV-datapath / FWHT hot path.
 *
 * The reference (peteroh23/ModulationDetectionCNN) has no FFI or plugin seam: its
 * only inference entry points are the Keras calls in cnn.py and CNN.ipynb and the
 * `layers_top` module port list in cnn_test_latest1.sv.  Each entry point below
 * names the reference interface it stands in for; INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add to cnn.py.
 *
 * Conventions
 *   - plain C, no exceptions across the boundary; every function returns an int
 *     status: 0 = MDC_OK, negative = error.  mdc_last_error() returns a
 *     thread-local, NUL-terminated description of the last failure.
 *   - pointers named *_dev are CUDA device pointers owned by the caller (e.g. a
 *     torch tensor's data_ptr()); pointers named *_host are host pointers.  The
 *     library never frees caller memory.
 *   - every device entry point takes an explicit stream (cudaStream_t passed as
 *     void*), only enqueues work on it and does not synchronise.
 *   - a handle belongs to one device; it is thread-compatible (one handle per
 *     thread, or external locking).
 *   - there is no CPU fallback: without a CUDA device every compute call fails
 *     with MDC_ERR_CUDA.
 */
#ifndef MDC_H_
#define MDC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDC_API __attribute__((visibility("default")))

typedef struct mdc_handle_s* mdc_handle_t;

/* status codes */
enum {
  MDC_OK = 0,
  MDC_ERR_INVALID = -1,     /* bad argument / wrong model kind or mode for this call */
  MDC_ERR_CUDA = -2,        /* CUDA runtime/driver error (text in mdc_last_error) */
  MDC_ERR_NOT_READY = -3,   /* weights not (completely) set */
  MDC_ERR_UNSUPPORTED = -4, /* shape outside what the kernels implement */
  MDC_ERR_RANGE = -5        /* MDC_MODE_F16X3 only: an input or activation left the range the fp16 hi/lo split
                               represents (|value| > 65504); the results of the call are not to be trusted -
                               rerun it on an MDC_MODE_TF32X3 handle (the Python facade does so by itself) */
};

/* model kinds */
enum {
  MDC_MODEL_TINY = 0, /* TinyCNN2(F,C): CNN.ipynb cell 6 / the *.wts.h5 model_config */
  MDC_MODEL_VT = 1    /* VT-CNN2(C):   examples-master/.../RML2016.10a_VTCNN2_example.ipynb:231-243 */
};

/* arithmetic modes */
enum {
  MDC_MODE_FP32 = 0,   /* fp32 FMA on CUDA cores (parity mode, <=1e-5 of the fp64 oracle) */
  MDC_MODE_BF16 = 1,   /* VT-CNN2 only: bf16 operands, fp32 accumulate, tcgen05 tensor cores */
  MDC_MODE_TF32X3 = 2, /* VT-CNN2 only: every fp32 operand split into tf32 hi + lo, three tcgen05 kind::tf32
                          MMAs per product; fp32-level accuracy (<=1e-5 of the fp64 oracle) on tensor cores */
  MDC_MODE_Q612 = 3,   /* TinyCNN2 only: bit-exact 18-bit Q6.12 SystemVerilog datapath */
  MDC_MODE_F16X3 = 4   /* VT-CNN2 only: every fp32 operand split into fp16 hi + 2^-11 fp16 lo, three tcgen05 kind::f16
                          MMAs per product at the full 16-bit rate; fp32-level accuracy (<=1e-5 of the fp64 oracle)
                          at ~2x the TF32X3 rate.  Values must stay inside the fp16 range (MDC_ERR_RANGE otherwise) */
};

/* frame formats of the mdc_predict_raw* / mdc_predict_q612_raw* calls (converted inside the kernels' frame load:
 * VT-CNN2 tensor-core modes, the specialised TinyCNN2 fp32 kernels F in {3,10}, C = 3, and the integer kernel) */
enum {
  MDC_IN_F32 = 0,  /* f32 [n,2,128], row 0 = I, row 1 = Q: what model.predict receives (cnn.py:198)            */
  MDC_IN_U8IQ = 1, /* u8 [n,128,2]: raw RTL-SDR bytes I0 Q0 I1 Q1 ... (README.md:5); value = (u - 127.5)/128,
                      identical to mdc_sdr_ingest_u8 followed by mdc_predict_f32.  256 B per frame               */
  MDC_IN_I16 = 2,  /* i16 [n,256]: Q6.12 samples in the test_table address map (0-127 I, 128-255 Q,
                      cnn_test_latest1.sv:88-102), value = s / 4096.  512 B per frame                             */
  MDC_IN_I32 = 3   /* i32 [n,256]: the 18-bit words of test_table as mdc_predict_q612 takes them (integer model
                      only).  1,024 B per frame                                                                   */
};

/* tensor ids for mdc_set_weights_f32 (Keras layouts: conv (kh,kw,cin,cout), dense (in,out)) */
enum {
  MDC_T_CONV1_K = 0, MDC_T_CONV1_B = 1,
  MDC_T_CONV2_K = 2, MDC_T_CONV2_B = 3,   /* VT only */
  MDC_T_DENSE1_K = 4, MDC_T_DENSE1_B = 5,
  MDC_T_DENSE2_K = 6, MDC_T_DENSE2_B = 7  /* VT only */
};

/* options for mdc_set_option */
enum {
  MDC_OPT_FLATTEN_ORDER = 0 /* VT only. 0 = channels_last (row = pos*80+ch, Keras 2 / TF,
                               default), 1 = channels_first (row = ch*132+pos, Keras 1 / Theano) */
};

/* FWHT output orderings */
enum { MDC_FWHT_NATURAL = 0, MDC_FWHT_SEQUENCY = 1 };

/* ---- lifetime --------------------------------------------------------------------------
 * Replaces: building the Keras `Sequential` (cnn.py:104-115, CNN.ipynb cell 6).
 * filters: TinyCNN2 F (3 or 10 in the shipped checkpoints; 1..16 accepted); ignored for VT.
 * classes: C (3 for every shipped checkpoint, 11 for VT-CNN2; 1..16 accepted).
 * device:  CUDA ordinal.                                                                  */
MDC_API int mdc_create(int model_kind, int filters, int classes, int mode, int device,
                       mdc_handle_t* out);
MDC_API int mdc_destroy(mdc_handle_t h);
MDC_API int mdc_set_option(mdc_handle_t h, int option, int value);

/* ---- weights ---------------------------------------------------------------------------
 * Replaces: model.load_weights(filepath) (cnn.py:147, CNN.ipynb cell 8) - the Python side
 * parses the .h5 (h5lite.py) and hands each tensor over in its Keras layout.
 * count must equal the tensor's element count for (model, F, C).  Synchronous.            */
MDC_API int mdc_set_weights_f32(mdc_handle_t h, int tensor_id, const float* host_ptr, size_t count);

/* Replaces: the ROM modules of cnn_test_latest1.sv (rom_cov :685-707, dense_bias :260-262,
 * rom_dense_{i,q}_class{1,2,3} :719-3132) == the *.Weights.txt tables.  Literal contents,
 * 18-bit signed values in int32:  conv_tab[3F] = {w0,w1,bias} per filter;  dense_bias[C];
 * dense_tabs[2C][129F] ordered class0-I, class0-Q, class1-I, ...   Synchronous.          */
MDC_API int mdc_set_weights_q612(mdc_handle_t h, const int32_t* conv_tab_host,
                                 const int32_t* dense_bias_host, const int32_t* dense_tabs_host);

/* ---- float inference -------------------------------------------------------------------
 * Replaces: model.predict(X, batch_size) (cnn.py:198,237; CNN.ipynb cells 12,17,18).
 * x_dev      f32 [n,2,128], row 0 = I, row 1 = Q, C-contiguous.
 * probs_dev  f32 [n,C]  softmax probabilities                     (may be NULL)
 * dense_dev  f32 [n,C]  last Dense output before softmax: TinyCNN2 Dense+ReLU
 *                       (`model2` of CNN.ipynb cell 17), VT-CNN2 logits     (may be NULL)
 * cls_dev    i32 [n]    argmax class, first maximum wins like np.argmax (may be NULL)
 * hist_dev   u64 [C]    class histogram, ACCUMULATED into (caller zeroes)  (may be NULL)
 * The fused argmax/histogram replace the per-row Python loops of cnn.py:205-211,241-247. */
MDC_API int mdc_predict_f32(mdc_handle_t h, const float* x_dev, int64_t n, float* probs_dev,
                            float* dense_dev, int32_t* cls_dev, unsigned long long* hist_dev,
                            void* stream);

/* Size the handle's work space for calls of up to max_frames frames (clamped to the pass size the mode uses), and
 * pack the weights: after it, mdc_predict_f32 / mdc_predict_raw only enqueue kernels - no allocation, copy or
 * synchronisation - from the very first call, which is what capturing them into a CUDA graph requires.  A graph
 * captured earlier stays valid as long as no later call or mdc_reserve asks for MORE frames (growing the work space
 * moves it).  Weights must be set.                                                                               */
MDC_API int mdc_reserve(mdc_handle_t h, int64_t max_frames);

/* mdc_predict_f32 for frames in one of the MDC_IN_* formats.  MDC_IN_U8IQ / MDC_IN_I16 need a VT-CNN2 handle in a
 * tensor-core mode (BF16, F16X3, TF32X3) or a TinyCNN2 fp32 handle of a shipped shape (F = 3 or 10, C = 3); other
 * handles answer MDC_ERR_UNSUPPORTED.  Results are bit-identical to mdc_sdr_ingest_u8 (or s / 4096) followed by
 * mdc_predict_f32.  x_dev must be 16-byte aligned.                                                                */
MDC_API int mdc_predict_raw(mdc_handle_t h, const void* x_dev, int in_format, int64_t n, float* probs_dev,
                            float* dense_dev, int32_t* cls_dev, unsigned long long* hist_dev, void* stream);

/* MDC_MODE_F16X3: *flags = bit 0 set when some call since the last reset saw an input or activation outside the
 * fp16 range (see MDC_ERR_RANGE).  Synchronises the device.  The host-buffer calls below check it themselves.    */
MDC_API int mdc_range_flags(mdc_handle_t h, unsigned int* flags, int reset);

/* Same call with HOST buffers (what a numpy caller has): copies x in chunks on internal
 * streams, overlapping H2D / kernels / D2H, and returns when the outputs are in host memory.
 * x_host may be ordinary pageable memory (what cnn.py:198,237 pass): the library then stages it through its own
 * pinned ring with a few copy threads; pinned (cudaHostAlloc / cudaHostRegister'ed) input is copied directly.
 * hist_host u64[C] is overwritten (not accumulated).                                      */
MDC_API int mdc_predict_f32_host(mdc_handle_t h, const float* x_host, int64_t n, float* probs_host,
                                 float* dense_host, int32_t* cls_host,
                                 unsigned long long* hist_host);

/* Streaming form of the call above: returns as soon as the copies and kernels are enqueued, so the next batch's
 * transfer runs under this batch's kernels.  All host buffers must stay valid (and should be pinned) until
 * mdc_host_wait(h, *ticket) returns; results of successive calls complete in issue order (ticket 0 = nothing
 * pending, e.g. n == 0).  mdc_predict_q612_host_async below is the integer counterpart.                          */
MDC_API int mdc_predict_f32_host_async(mdc_handle_t h, const float* x_host, int64_t n, float* probs_host,
                                       float* dense_host, int32_t* cls_host,
                                       unsigned long long* hist_host, int64_t* ticket);
MDC_API int mdc_host_wait(mdc_handle_t h, int64_t ticket);

/* The two host-buffer calls for frames in one of the MDC_IN_* formats (a quarter / half of the bytes over PCIe).   */
MDC_API int mdc_predict_raw_host(mdc_handle_t h, const void* x_host, int in_format, int64_t n, float* probs_host,
                                 float* dense_host, int32_t* cls_host, unsigned long long* hist_host);
MDC_API int mdc_predict_raw_host_async(mdc_handle_t h, const void* x_host, int in_format, int64_t n,
                                       float* probs_host, float* dense_host, int32_t* cls_host,
                                       unsigned long long* hist_host, int64_t* ticket);

/* ---- integer (SystemVerilog-exact) inference ------------------------------------------
 * Replaces: one reset-to-done run of `layers_top` (cnn_test_latest1.sv:144-209) per frame,
 * fed by `test_input`/`test_table` (:71-142).
 * x_dev    i32 [n,256]: entries 0..127 = I, 128..255 = Q, 18-bit signed, sign-extended
 *                       (the address map of test_table, sv:88-89,102).
 * out_dev  i32 [n,C]    `out_data`: ReLU'd 32-bit accumulators, Q.12        (may be NULL)
 * pre_dev  i32 [n,C]    `pre_out_data`: accumulators before the final ReLU  (may be NULL)
 * cls_dev / hist_dev as above (argmax over out_data).                                    */
MDC_API int mdc_predict_q612(mdc_handle_t h, const int32_t* x_dev, int64_t n, int32_t* out_dev,
                             int32_t* pre_dev, int32_t* cls_dev, unsigned long long* hist_dev,
                             void* stream);
MDC_API int mdc_predict_q612_host(mdc_handle_t h, const int32_t* x_host, int64_t n,
                                  int32_t* out_host, int32_t* pre_host, int32_t* cls_host,
                                  unsigned long long* hist_host);
MDC_API int mdc_predict_q612_host_async(mdc_handle_t h, const int32_t* x_host, int64_t n,
                                        int32_t* out_host, int32_t* pre_host, int32_t* cls_host,
                                        unsigned long long* hist_host, int64_t* ticket);
/* The same three calls with the frame format as an argument (what feeds test_table in a deployment is an ADC, not
 * 32-bit words): MDC_IN_I32 (as above), MDC_IN_I16 (int16 [n,256], same address map, sv:88-102: half the bytes) or
 * MDC_IN_U8IQ (raw RTL-SDR bytes, README.md:5; sample = (2u - 255) * 16, the Q6.12 value of (u - 127.5) / 128 -
 * identical to mdc_sdr_ingest_u8 followed by mdc_predict_q612: a quarter of the bytes).  Converted in the frame load. */
MDC_API int mdc_predict_q612_raw(mdc_handle_t h, const void* x_dev, int in_format, int64_t n, int32_t* out_dev,
                                 int32_t* pre_dev, int32_t* cls_dev, unsigned long long* hist_dev, void* stream);
MDC_API int mdc_predict_q612_raw_host(mdc_handle_t h, const void* x_host, int in_format, int64_t n,
                                      int32_t* out_host, int32_t* pre_host, int32_t* cls_host,
                                      unsigned long long* hist_host);
MDC_API int mdc_predict_q612_raw_host_async(mdc_handle_t h, const void* x_host, int in_format, int64_t n,
                                            int32_t* out_host, int32_t* pre_host, int32_t* cls_host,
                                            unsigned long long* hist_host, int64_t* ticket);

/* ---- Walsh-Hadamard transform ----------------------------------------------------------
 * Replaces: the FWHT spectrogram stage named in README.md:5 (no code in the reference).
 * Unnormalised WHT of each row: in/out i32 [n_spectra, 2^log2_npt], wrap mod 2^32.
 * log2_npt in [5,13]; in_dev == out_dev (in place) is allowed.                          */
MDC_API int mdc_fwht_i32(const int32_t* in_dev, int32_t* out_dev, int64_t n_spectra, int log2_npt,
                         int ordering, void* stream);
MDC_API int mdc_fwht_i32_host(const int32_t* in_host, int32_t* out_host, int64_t n_spectra,
                              int log2_npt, int ordering, int device);

/* ---- raw RTL-SDR ingest ------------------------------------------------------------------
 * The step before the path (README.md:5: samples arrive from an RTL-SDR through the HPS; no code in
 * the reference).  iq_dev u8 [2*n_samples]: interleaved unsigned 8-bit I0 Q0 I1 Q1 ..., centred on
 * 127.5.  value = (u - 127.5)/128; its Q6.12 integer (2u - 255)*16 is exact.  Any subset of:
 * frames_f32_dev  f32 [n/128,2,128]   (row 0 = I, row 1 = Q: the mdc_predict_f32 input)
 * frames_q612_dev i32 [n/128,256]     (0-127 I, 128-255 Q: the mdc_predict_q612 input)
 * fwht_dev        i32 [n/1024,2,1024] (Q6.12 I block then Q block: 2 rows for mdc_fwht_i32)
 * n_samples must be a multiple of 128 (of 1024 when fwht_dev is given); NULL outputs are skipped. */
MDC_API int mdc_sdr_ingest_u8(const uint8_t* iq_dev, int64_t n_samples, float* frames_f32_dev,
                              int32_t* frames_q612_dev, int32_t* fwht_dev, void* stream);

/* ---- caller-side epilogue --------------------------------------------------------------
 * Replaces: the confusion-matrix loops of cnn.py:200-211,239-247 / CNN.ipynb cell 12.
 * conf_dev u64 [C,C] accumulated: conf[true[i]][pred[i]] += 1.                           */
MDC_API int mdc_confusion_i32(const int32_t* true_dev, const int32_t* pred_dev, int64_t n,
                              int classes, unsigned long long* conf_dev, void* stream);
/* One matrix per group in a single pass - replaces the per-SNR loop of cnn.py:227-255 (select the
 * frames of one SNR, predict, count, accuracy = trace / sum).  group_dev i32 [n] in [0, groups)
 * (NULL = one group); conf_dev u64 [groups,C,C] accumulated.  Out-of-range labels are skipped.  */
MDC_API int mdc_confusion_grouped_i32(const int32_t* true_dev, const int32_t* pred_dev,
                                      const int32_t* group_dev, int64_t n, int classes, int groups,
                                      unsigned long long* conf_dev, void* stream);

/* ---- introspection ---------------------------------------------------------------------*/
MDC_API const char* mdc_last_error(void);
MDC_API const char* mdc_version(void);
/* number of kernel launches this handle has enqueued so far (bench.py's gpu_launches)     */
MDC_API int64_t mdc_launch_count(mdc_handle_t h);
/* name of the dominant kernel of the handle's predict path and the CUDA-event time of its
 * launches: mdc_profile_enable(h,1) makes every predict call bracket that kernel with
 * events on the caller's stream; mdc_profile_read returns the accumulated milliseconds and
 * launch count (synchronises those events) and resets the accumulators.                   */
MDC_API int mdc_profile_enable(mdc_handle_t h, int on);
MDC_API int mdc_profile_read(mdc_handle_t h, double* ms_total, int64_t* launches,
                             const char** kernel_name);

/* diagnostics: copy an intermediate of the handle's LAST predict pass to the host (synchronises
 * the device).  what = 0: VT-CNN2 conv2 activations (bf16 [frames*132][80] in BF16 mode, f32 in
 * FP32 mode, f32 hi matrix then lo matrix in TF32X3 mode) - `model3`-style layer taps of CNN.ipynb cell 17;
 * what = 1: dense1 activations f32 [frames][256] (BF16 mode fuses the rest of the network into the dense1
 * kernel and keeps no copy unless the process runs with MDC_VT_KEEP_H=1).  bytes is clamped to the workspace
 * size; returns the bytes copied in *copied.
 * (F16X3 mode: what = 0 gives the fp16 hi matrix then the lo matrix, [frames*132][80] each, of the pass size
 * the work space was reserved for.) */
MDC_API int mdc_debug_read(mdc_handle_t h, int what, void* host_dst, size_t bytes, size_t* copied);

#ifdef __cplusplus
}
#endif
#endif /* MDC_H_ */
