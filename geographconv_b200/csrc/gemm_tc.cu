// gemm_tc.cu -- tcgen05 (5th-gen tensor core) GEMM path.  Placeholder until the kernel lands:
// reports "not supported" so the dispatchers in capi.cu use the CUDA-core kernels.
#include "common.cuh"

bool gcnb_highway_tc_supported(const gcnb_ctx*, int, int, int, int, int, int) { return false; }
size_t gcnb_highway_tc_workspace_bytes(int) { return 0; }
int gcnb_highway_tc(gcnb_ctx* ctx, int, int, const float*, int, const float*, int, const float*, int, const float*,
                    const float*, int, const float*, int, float*, int, float*, int, float*, int) {
  return gcnb_fail(ctx, GCNB_E_UNSUPPORTED, "tcgen05 highway kernel not built%s", "");
}
bool gcnb_gemm_tc_supported(const gcnb_ctx*, int, int, int, int, int, int, int, int, int) { return false; }
size_t gcnb_gemm_tc_workspace_bytes(int, int) { return 0; }
int gcnb_gemm_tc(gcnb_ctx* ctx, int, int, int, int, const float*, int, const float*, int, float*, int, const float*,
                 int) {
  return gcnb_fail(ctx, GCNB_E_UNSUPPORTED, "tcgen05 gemm kernel not built%s", "");
}
