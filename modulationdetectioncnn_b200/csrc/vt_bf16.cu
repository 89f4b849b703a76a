// placeholder until the tcgen05 path lands (replaced in a later commit)
#include "mdc_internal.cuh"
namespace mdc {
int pack_vt_bf16(mdc_handle_s*) { set_error("MDC_MODE_BF16 not built yet"); return MDC_ERR_UNSUPPORTED; }
int launch_vt_bf16(mdc_handle_s*, const float*, int64_t, float*, float*, int32_t*, unsigned long long*, cudaStream_t) {
  set_error("MDC_MODE_BF16 not built yet"); return MDC_ERR_UNSUPPORTED; }
}
