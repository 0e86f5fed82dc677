// q4_matrix.cu -- Q4 matrix handle support kernels: act-order row re-ordering and full dequantisation.
// Replaces exllama_ext/cuda_func/q4_matrix.cu (make_sequential_kernel :61-102, reconstruct_kernel :170-210).
#include "exl_common.cuh"

namespace {

// Gather the 8 source rows of each new packed row.  One thread = one uint4 (4 columns) of one new row:
// coalesced 16-byte loads from 8 different source rows, one 16-byte store.  (The reference moves 64 bits
// per thread, q4_matrix.cu:70-101.)
__global__ void __launch_bounds__(256) make_sequential_kernel(const uint32_t* __restrict__ w, uint32_t* __restrict__ w_new,
                                                              const uint32_t* __restrict__ x_map, int N4)
{
    const int c4 = blockIdx.x * blockDim.x + threadIdx.x;
    if (c4 >= N4) return;
    const int r8 = blockIdx.y;
    uint4 dst = make_uint4(0, 0, 0, 0);
    #pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t src_row = x_map[r8 * 8 + i];
        const uint4 s = reinterpret_cast<const uint4*>(w)[(size_t)(src_row >> 3) * N4 + c4];
        const int sh = (src_row & 7) * 4;
        dst.x |= ((s.x >> sh) & 0xfu) << (4 * i);
        dst.y |= ((s.y >> sh) & 0xfu) << (4 * i);
        dst.z |= ((s.z >> sh) & 0xfu) << (4 * i);
        dst.w |= ((s.w >> sh) & 0xfu) << (4 * i);
    }
    reinterpret_cast<uint4*>(w_new)[(size_t)r8 * N4 + c4] = dst;
}

// out[k, n] = hmul(int2half(q - (z + 1)), scale)  -- bit-exact with reconstruct_kernel (q4_matrix.cu:196-208).
// One thread = 8 rows x 2 columns; half2 stores are coalesced along n.
__global__ void __launch_bounds__(256) reconstruct_kernel(const uint32_t* __restrict__ w, half* __restrict__ out,
                                                          const half* __restrict__ scales, const uint32_t* __restrict__ zeros,
                                                          int K8, int N, int groupsize)
{
    const int c2 = blockIdx.x * blockDim.x + threadIdx.x;     // column pair
    if (c2 * 2 >= N) return;
    const int r8 = blockIdx.y;
    const int col = c2 * 2;
    const int group = (r8 * 8) / groupsize;
    const half2 sc = *reinterpret_cast<const half2*>(scales + (size_t)group * N + col);
    const uint32_t zw = zeros[(size_t)group * (N >> 3) + (col >> 3)];
    const int z0 = (int)((zw >> ((col & 7) * 4)) & 0xfu) + 1;
    const int z1 = (int)((zw >> (((col + 1) & 7) * 4)) & 0xfu) + 1;
    const uint2 q = *reinterpret_cast<const uint2*>(w + (size_t)r8 * N + col);
    half2* o = reinterpret_cast<half2*>(out + (size_t)r8 * 8 * N + col);
    #pragma unroll
    for (int i = 0; i < 8; i++) {
        const int q0 = (int)((q.x >> (4 * i)) & 0xfu) - z0;
        const int q1 = (int)((q.y >> (4 * i)) & 0xfu) - z1;
        const half2 v = __halves2half2(__int2half_rn(q0), __int2half_rn(q1));
        o[(size_t)i * (N >> 1)] = __hmul2(v, sc);
    }
}

} // namespace

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda).
int exl_encode_weight_tmap(exl_q4_matrix* w)
{
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            return exl_set_err(EXL_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
        encode = (EncodeFn)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)w->N, (cuuint64_t)(w->K / 8)};
    const cuuint64_t strides[1] = {(cuuint64_t)w->N * 4};              // bytes between k8-rows
    const cuuint32_t box[2] = {32, 16};                                // 32 columns (128 B) x 16 k8-rows
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&w->tmap_w, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*)w->qweight, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return exl_set_err(EXL_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for K=%d N=%d", (int)r, w->K, w->N);
    const cuuint32_t box_p[2] = {128, 8};                              // 128 columns x 8 k8-rows (one 64-wide k-block), dense
    r = encode(&w->tmap_wp, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*)w->qweight, dims, strides, box_p, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return exl_set_err(EXL_ERR_CUDA, "cuTensorMapEncodeTiled (prefill box) failed (%d) for K=%d N=%d", (int)r, w->K, w->N);
    w->valid3 = 0;
    if (w->N % 128 == 0 && w->K % 128 == 0) {
        const cuuint64_t d3[3] = {32, (cuuint64_t)(w->K / 8), (cuuint64_t)(w->N / 32)};
        const cuuint64_t s3[2] = {(cuuint64_t)w->N * 4, 128};           // bytes: next k8-row, next column group
        const cuuint32_t b3[3] = {32, 16, 4};
        const cuuint32_t e3[3] = {1, 1, 1};
        r = encode(&w->tmap_w3, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void*)w->qweight, d3, s3, b3, e3,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return exl_set_err(EXL_ERR_CUDA, "cuTensorMapEncodeTiled (3-D unit box) failed (%d) for K=%d N=%d", (int)r, w->K, w->N);
        const cuuint32_t R = (w->groups > 1 && w->groupsize < 128 && 128 % w->groupsize == 0) ? (cuuint32_t)(128 / w->groupsize) : 1u;
        const cuuint64_t dsc[2] = {(cuuint64_t)w->N, (cuuint64_t)w->groups};
        const cuuint64_t ssc[1] = {(cuuint64_t)w->N * 2};
        const cuuint32_t bsc[2] = {128, R};
        r = encode(&w->tmap_sc, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)w->scales, dsc, ssc, bsc, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return exl_set_err(EXL_ERR_CUDA, "cuTensorMapEncodeTiled (scales) failed (%d)", (int)r);
        const cuuint64_t dqz[2] = {(cuuint64_t)(w->N / 8), (cuuint64_t)w->groups};
        const cuuint64_t sqz[1] = {(cuuint64_t)(w->N / 8) * 4};
        const cuuint32_t bqz[2] = {16, R};
        r = encode(&w->tmap_qz, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*)w->qzeros, dqz, sqz, bqz, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return exl_set_err(EXL_ERR_CUDA, "cuTensorMapEncodeTiled (qzeros) failed (%d)", (int)r);
        w->valid3 = 1;
    }
    return EXL_OK;
}

int exl_make_sequential_launch(uint32_t* qweight, uint32_t* tmp, const uint32_t* x_map, int K, int N, cudaStream_t stream)
{
    dim3 block(256), grid((N / 4 + 255) / 256, K / 8);
    make_sequential_kernel<<<grid, block, 0, stream>>>(qweight, tmp, x_map, N / 4);
    EXL_CHECK_LAUNCH("make_sequential_kernel");
    EXL_CUDA_TRY(cudaMemcpyAsync(qweight, tmp, (size_t)(K / 8) * N * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
    return EXL_OK;
}

int exl_reconstruct_launch(const exl_q4_matrix* w, half* out, cudaStream_t stream)
{
    dim3 block(256), grid((w->N / 2 + 255) / 256, w->K / 8);
    reconstruct_kernel<<<grid, block, 0, stream>>>(w->qweight, out, w->scales, w->qzeros, w->K / 8, w->N, w->groupsize);
    EXL_CHECK_LAUNCH("reconstruct_kernel");
    return EXL_OK;
}
