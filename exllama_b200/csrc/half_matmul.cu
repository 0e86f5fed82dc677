// half_matmul.cu -- fp16 x fp16 matmuls on the operator surface (LoRA adapters, half_matmul / half_matmul_cublas).
// Replaces exllama_ext/cuda_func/half_matmul.cu.  These are plain library GEMMs in the reference
// (cublasHgemm, half_matmul.cu:119) and stay library GEMMs here; the custom entry point keeps the
// reference's "accumulate into a pre-zeroed out" contract (half_matmul.cu:15-54, cuda_ext.py:124).
#include "exl_common.cuh"

namespace {

// out[m, n] += sum_k x[m,k] w[k,n], fp32 accumulation, one warp per (m, 64-column strip), split along K by
// blockIdx.z with a deterministic two-pass is overkill for the tiny LoRA shapes this serves: each CTA owns its
// outputs completely (no atomics), looping over all of K.
__global__ void __launch_bounds__(128) half_matmul_kernel(const half* __restrict__ x, const half* __restrict__ w,
                                                          half* __restrict__ out, int M, int K, int N)
{
    const int n2 = (blockIdx.x * 128 + threadIdx.x) * 2;
    const int m = blockIdx.y;
    if (n2 >= N) return;
    const half* xr = x + (size_t)m * K;
    float a0 = 0.f, a1 = 0.f;
    if (n2 + 1 < N) {
        for (int k = 0; k < K; k++) {
            const float xv = __half2float(xr[k]);
            const float2 wv = __half22float2(*reinterpret_cast<const half2*>(w + (size_t)k * N + n2));
            a0 = fmaf(xv, wv.x, a0); a1 = fmaf(xv, wv.y, a1);
        }
        half2* o = reinterpret_cast<half2*>(out + (size_t)m * N + n2);
        const float2 prev = __half22float2(*o);
        *o = __floats2half2_rn(prev.x + a0, prev.y + a1);
    } else {
        for (int k = 0; k < K; k++) a0 = fmaf(__half2float(xr[k]), __half2float(w[(size_t)k * N + n2]), a0);
        out[(size_t)m * N + n2] = __float2half_rn(__half2float(out[(size_t)m * N + n2]) + a0);
    }
}

} // namespace

int exl_half_matmul_custom_launch(const half* x, const half* w, half* out, int M, int K, int N, cudaStream_t stream)
{
    if (M <= 0 || N <= 0) return EXL_OK;
    if (N % 2 != 0) return exl_set_err(EXL_ERR_ARG, "half_matmul: N=%d must be even", N);
    dim3 grid((N / 2 + 127) / 128, M);
    half_matmul_kernel<<<grid, 128, 0, stream>>>(x, w, out, M, K, N);
    EXL_CHECK_LAUNCH("half_matmul_kernel");
    return EXL_OK;
}

int exl_half_matmul_cublas_launch(ExlDevice* ds, const half* x, const half* w, half* out, int M, int K, int N,
                                  bool no_zero, cudaStream_t stream)
{
    if (M <= 0 || N <= 0) return EXL_OK;
    // row-major out[M,N] = x[M,K] . w[K,N]  ==  column-major out^T[N,M] = w^T[N,K] . x^T[K,M]
    const float alpha = 1.0f, beta = no_zero ? 1.0f : 0.0f;
    cublasStatus_t st = cublasSetStream(ds->blas, stream);
    if (st != CUBLAS_STATUS_SUCCESS) return exl_set_err(EXL_ERR_CUDA, "cublasSetStream failed (%d)", (int)st);
    st = cublasGemmEx(ds->blas, CUBLAS_OP_N, CUBLAS_OP_N, N, M, K, &alpha, w, CUDA_R_16F, N, x, CUDA_R_16F, K, &beta,
                      out, CUDA_R_16F, N, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT_TENSOR_OP);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (st != CUBLAS_STATUS_SUCCESS) return exl_set_err(EXL_ERR_CUDA, "cublasGemmEx failed (%d)", (int)st);
    return EXL_OK;
}
