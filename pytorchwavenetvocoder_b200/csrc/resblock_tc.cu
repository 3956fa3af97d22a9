// resblock_tc.cu -- tcgen05/TMA fused residual block (WNB_MATH_TF32).  Placeholder until the kernel lands.
#include "common.cuh"

namespace wnb {
struct FwdParams;
bool resblock_fwd_tc_supported(int R, int S, int Ap, int ks) { (void)R; (void)S; (void)Ap; (void)ks; return false; }
int resblock_fwd_tc(const FwdParams& p, cudaStream_t st) {
  (void)p; (void)st;
  set_error("resblock_fwd_tc: not built");
  return WNB_ERR_UNSUPPORTED;
}
}  // namespace wnb
