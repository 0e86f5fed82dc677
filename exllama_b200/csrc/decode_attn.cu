// decode_attn.cu -- single-query attention over the KV cache for the decode step ("next" row).
//
// Replaces the five torch launches between q4_attn and q4_attn_2 in ExLlamaAttention.fused (model.py:372-409:
// repeat_kv, matmul(q, k^T), /= sqrt(d), softmax(dtype=fp16), matmul(p, v), transpose) with one kernel that reads the
// K/V cache [kv_heads, max_seq, head_dim] exactly once.  HBM-bound: 2 * heads * S * 128 * 2 bytes per call.
//
//   grid  = (heads, nsplit), cluster (1, nsplit): the CTAs of a cluster split the sequence of one head
//   phase 1  16 lanes x 16 B = one K row (a warp load covers two whole rows): s_p = (q . K[p]) / sqrt(d) in fp32,
//            block max / sum (online-softmax partial)
//   phase 2  16 threads x 16 B = one V row, 16 position groups: o += p_p * V[p]
//   combine  partial (m, l, o[128]) deposited in the leader's shared memory through DSMEM, cluster barrier, leader
//            rescales and writes fp16 out[head*128 ..]  (the layout q4_attn_2 consumes, model.py:409-411)
// Everything in fp32 (the reference rounds scores and probabilities to fp16); head_dim == 128, q_len == 1, bsz == 1.
#include "exl_common.cuh"

namespace {

constexpr int DA_THREADS = 256;
constexpr int HD = 128;
#ifndef DA_VPRE
#define DA_VPRE 4          // V rows per thread requested together with the 16 K rows (register budget: 128 at 2 CTAs/SM)
#endif

__device__ __forceinline__ uint32_t da_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(DA_THREADS, 2) decode_attn_kernel(const half* __restrict__ q, const half* __restrict__ kc,
                                                                 const half* __restrict__ vc, half* __restrict__ out,
                                                                 int heads, int kv_heads, int seq, int max_seq, float scale, int nsplit)
{
    __shared__ float s_q[HD];
    __shared__ float s_p[DA_THREADS];
    __shared__ float s_red[DA_THREADS / 32];
    __shared__ float s_o[16][HD];
    __shared__ float s_part[8][HD + 2];        // leader: per split {o[128], m, l}

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int head = blockIdx.x, split = blockIdx.y;
    const int kvh = head / (heads / kv_heads);
    const int chunk = (seq + nsplit - 1) / nsplit;
    const int p0 = split * chunk, p1 = min(seq, p0 + chunk);
    const half* kbase = kc + (size_t)kvh * max_seq * HD;
    const half* vbase = vc + (size_t)kvh * max_seq * HD;

    if (nsplit > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");

    // online softmax over sub-blocks of DA_THREADS positions (one sub-block when chunk <= DA_THREADS)
    const int l16 = lane & 15, sub = lane >> 4;         // phase 1: 16 lanes x 16 B = one K row; a warp load covers 2 rows
    const int r16 = tid & 15, pg = tid >> 4;            // phase 2: 16 threads x 16 B = one V row; 16 position groups
    float qf[8];
    bool have_q = false;
    float acc[8];
    #pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
    float mx = -INFINITY, lsum = 0.f;
    for (int b0 = p0; b0 < p1; b0 += DA_THREADS) {
        const int b1 = min(p1, b0 + DA_THREADS);
        // every K and V byte of the sub-block is requested up front (32 x 16 B per thread in flight): one HBM round trip
        uint4 kv[16], vv[16];
        #pragma unroll
        for (int it = 0; it < 16; it++) {
            const int p = b0 + warp * 32 + it * 2 + sub;
            kv[it] = (p < b1 && (have_q || p != seq - 1)) ? __ldg(reinterpret_cast<const uint4*>(kbase + (size_t)p * HD) + l16) : make_uint4(0, 0, 0, 0);
        }
        #pragma unroll
        for (int it = 0; it < DA_VPRE; it++) {                      // part of V now, the rest once kv[] is consumed
            const int p = b0 + pg + it * 16;
            vv[it] = (p < b1 && (have_q || p != seq - 1)) ? __ldg(reinterpret_cast<const uint4*>(vbase + (size_t)p * HD) + r16) : make_uint4(0, 0, 0, 0);
        }
        if (!have_q) {
            // Programmatic dependent launch: the rows of the cache older than this token do not depend on the preceding
            // kernel (q4_attn writes only row seq-1), so they are already in flight; q and the newest row are read after
            // the wait.
            asm volatile("griddepcontrol.wait;" ::: "memory");
            asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
            if (tid < HD) s_q[tid] = __half2float(q[(size_t)head * HD + tid]) * scale;
            #pragma unroll
            for (int it = 0; it < 16; it++) {
                const int pk = b0 + warp * 32 + it * 2 + sub;
                if (pk == seq - 1) kv[it] = *(reinterpret_cast<const uint4*>(kbase + (size_t)pk * HD) + l16);
            }
            #pragma unroll
            for (int it = 0; it < DA_VPRE; it++) {
                const int pv = b0 + pg + it * 16;
                if (pv == seq - 1) vv[it] = *(reinterpret_cast<const uint4*>(vbase + (size_t)pv * HD) + r16);
            }
            __syncthreads();
            #pragma unroll
            for (int j = 0; j < 8; j++) qf[j] = s_q[l16 * 8 + j];
            have_q = true;
        }
        // ---- phase 1: scores; warp w owns positions b0 + 32 w .. + 31, two per load instruction ----
        #pragma unroll
        for (int it = 0; it < 16; it++) {
            const half2* h = reinterpret_cast<const half2*>(&kv[it]);
            float a = 0.f;
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                const float2 f = __half22float2(h[j]);
                a = fmaf(f.x, qf[2 * j], a);
                a = fmaf(f.y, qf[2 * j + 1], a);
            }
            a += __shfl_xor_sync(0xffffffffu, a, 8);
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            const int pl = warp * 32 + it * 2 + sub;
            if (l16 == 0) s_p[pl] = (b0 + pl < b1) ? a : -INFINITY;
        }
        #pragma unroll
        for (int it = DA_VPRE; it < 16; it++) {                     // after the wait: plain loads, the newest row included
            const int p = b0 + pg + it * 16;
            vv[it] = (p < b1) ? *(reinterpret_cast<const uint4*>(vbase + (size_t)p * HD) + r16) : make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
        const float s = s_p[tid];
        float m = s;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) s_red[warp] = m;
        __syncthreads();
        float bm = s_red[0];
        #pragma unroll
        for (int i = 1; i < DA_THREADS / 32; i++) bm = fmaxf(bm, s_red[i]);
        const float mnew = fmaxf(mx, bm);
        const float alpha = __expf(mx - mnew);                      // 0 on the first sub-block (mx = -inf)
        const float e = __expf(s - mnew);                           // exp(-inf) = 0 for the positions past b1
        float l = e;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
        __syncthreads();                                            // everyone has read s_p / s_red
        s_p[tid] = e;
        if (lane == 0) s_red[warp] = l;
        __syncthreads();
        float bl = 0.f;
        #pragma unroll
        for (int i = 0; i < DA_THREADS / 32; i++) bl += s_red[i];
        // ---- phase 2: o = alpha o + sum_p e_p V[p] ----
        #pragma unroll
        for (int j = 0; j < 8; j++) acc[j] *= alpha;
        #pragma unroll
        for (int it = 0; it < 16; it++) {
            const float w = s_p[pg + it * 16];                      // 0 past b1
            const half2* h = reinterpret_cast<const half2*>(&vv[it]);
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                const float2 f = __half22float2(h[j]);
                acc[2 * j] = fmaf(w, f.x, acc[2 * j]); acc[2 * j + 1] = fmaf(w, f.y, acc[2 * j + 1]);
            }
        }
        __syncthreads();                                            // s_p / s_red are rewritten by the next sub-block
        lsum = lsum * alpha + bl;
        mx = mnew;
    }
    #pragma unroll
    for (int j = 0; j < 8; j++) s_o[pg][r16 * 8 + j] = acc[j];
    __syncthreads();
    float osum = 0.f;
    if (tid < HD) {
        #pragma unroll
        for (int i = 0; i < 16; i++) osum += s_o[i][tid];
    }

    if (nsplit == 1) {
        if (tid < HD) out[(size_t)head * HD + tid] = __float2half_rn(osum / lsum);
        return;
    }
    // ---- combine the splits in the leader CTA (rank 0 of the cluster) through DSMEM ----
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");          // every CTA of the cluster is running
    uint32_t dst; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(da_smem(&s_part[rank][0])), "r"(0));
    if (tid < HD) {
        asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(dst + tid * 4), "f"(osum) : "memory");
    } else if (tid == HD) {
        asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(dst + HD * 4), "f"(mx) : "memory");
    } else if (tid == HD + 1) {
        asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(dst + (HD + 1) * 4), "f"(lsum) : "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (rank != 0) return;
    if (tid < HD) {
        float M = -INFINITY;
        for (int i = 0; i < nsplit; i++) M = fmaxf(M, s_part[i][HD]);
        float L = 0.f, o = 0.f;
        for (int i = 0; i < nsplit; i++) {
            const float w = __expf(s_part[i][HD] - M);           // empty splits carry m = -inf, l = 0 -> weight 0
            L = fmaf(s_part[i][HD + 1], w, L);
            o = fmaf(s_part[i][tid], w, o);
        }
        out[(size_t)head * HD + tid] = __float2half_rn(o / L);
    }
}

} // namespace

int exl_decode_attn_launch(const half* q, const half* kc, const half* vc, half* out, int heads, int kv_heads, int head_dim,
                           int seq, int max_seq, float scale, cudaStream_t stream)
{
    if (head_dim != HD) return exl_set_err(EXL_ERR_ARG, "decode_attn: head_dim %d != 128", head_dim);
    if (heads < 1 || kv_heads < 1 || heads % kv_heads != 0) return exl_set_err(EXL_ERR_ARG, "decode_attn: bad head counts %d / %d", heads, kv_heads);
    if (seq < 1 || seq > max_seq) return exl_set_err(EXL_ERR_ARG, "decode_attn: seq %d outside [1, %d]", seq, max_seq);
    int nsplit = (seq + DA_THREADS - 1) / DA_THREADS;
    if (nsplit > 8) nsplit = 8;                       // longer sequences: each split walks several sub-blocks
    // short sequences: still split a little so more than `heads` CTAs stream the cache
    if (nsplit < 4 && seq >= 256) nsplit = 4;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)heads, (unsigned)nsplit); cfg.blockDim = dim3(DA_THREADS); cfg.stream = stream;
    cudaLaunchAttribute at[2];
    int na = 0;
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1; na++;
    if (nsplit > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = 1; at[na].val.clusterDim.y = (unsigned)nsplit; at[na].val.clusterDim.z = 1; na++;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, decode_attn_kernel, q, kc, vc, out, heads, kv_heads, seq, max_seq, scale, nsplit);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return exl_set_err(EXL_ERR_CUDA, "launch of decode_attn_kernel failed: %s", cudaGetErrorString(e));
    return EXL_OK;
}
