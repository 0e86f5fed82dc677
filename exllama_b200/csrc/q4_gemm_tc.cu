// q4_gemm_tc.cu -- prefill hot path: tcgen05 fused-dequant GEMM (placeholder until the UMMA kernel lands;
// exl_q4_matmul routes M > EXL_SKINNY_MAX_M through reconstruct + cuBLAS while exl_tc_gemm_supported() is false).
#include "exl_common.cuh"

bool exl_tc_gemm_supported(const exl_q4_matrix* w, int M) { (void)w; (void)M; return false; }

int exl_tc_gemm_launch(ExlDevice* ds, const half* x, int M, const exl_q4_matrix* w, half* out, bool no_zero, cudaStream_t stream)
{
    (void)ds; (void)x; (void)M; (void)w; (void)out; (void)no_zero; (void)stream;
    return exl_set_err(EXL_ERR_ARG, "tc_gemm: not available in this build");
}
