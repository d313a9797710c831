// spconv_tc.cu -- tcgen05 (TF32, fp32 accumulate in TMEM) sparse convolution.  Placeholder until
// the tensor-core path lands: reports "unsupported" so callers fall through loudly, never silently.
#include "common.cuh"

namespace s2d {

int spconv_fwd_tf32(const float*, int, const float*, const int*, int, int, int, int, int, const float*, const float*,
                    const float*, int, float*, int, cudaStream_t) {
  set_error("s2d_spconv_fwd: TF32 tensor-core path not built in this version");
  return S2D_ERR_UNSUPPORTED;
}

}  // namespace s2d
