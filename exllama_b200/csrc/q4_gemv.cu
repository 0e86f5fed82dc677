// q4_gemv.cu -- decode hot path: fused 4-bit unpack + scale/zero + skinny GEMM (M <= 8) for sm_100a.
//
// Replaces the reference's q4_matmul_kernel + dot_product_8* (exllama_ext/cuda_func/q4_matmul.cu:34-212,
// exllama_ext/matrix.cuh:87-286) and, through the fused prologue/epilogues, rms_norm (rms_norm.cu),
// rope (rope.cu), update_cache_kernel (q4_attn.cu:19-72) and silu_mul (q4_mlp.cu:46-88) on the decode path.
//
// Design (HBM-bound; see DESIGN.md "decode kernel"):
//  * packed qweight [K/8, N] is streamed exactly once with 128-bit loads (ld.global.nc.L1::no_allocate.v4):
//    lane (g = lane/4, t = lane%4) of a warp loads the 4 columns 4g..4g+3 of k8-row t, so one warp
//    request covers 4 rows x 128 B and a CTA step (8 warps) covers 128 columns x 64 k.
//  * nibbles are expanded two at a time with the 0x6400 fp16 magic (q | 0x6400 == 1024 + q), the zero point is
//    folded into the bias removal, and the products are formed by mma.sync m16n8k16 with the *weights as the
//    A operand straight from registers* (rows = 16 output columns) and the <= 8 activation rows as the n = 8
//    operand: exact fp16 products, fp32 accumulation, and the cost is independent of M for M <= 8.
//    K is traversed in the permuted order (0,4,1,5,2,6,3,7) inside each 8-block so no nibble shuffling is needed;
//    x is staged into shared memory in that order (with the act-order x_map gather and, for the fused decoder
//    ops, the RMS norm folded into the staging).
//  * group scales are applied once per group to an fp32 group accumulator.
//  * work is a flat list of (128-column tile, 64-k step) items split evenly over a persistent grid
//    (stream-K): every CTA streams the same number of bytes whatever N and K are.  Tiles that span several
//    CTAs are finished by the last CTA to arrive (partials in an L2-resident workspace, fixed summation
//    order => deterministic; no fp16 atomics unlike the reference, q4_matmul.cu:203-211).
#include "exl_common.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int WN = 4;              // warps across columns (4 x 32 = 128 columns)
constexpr int WK = 2;              // warps across k
constexpr int STEP_K = 64;         // k per CTA step
constexpr int U = 4;               // register prefetch depth (steps)
constexpr int RED_LD = GV_TILE_N + 4;

struct GemvMatDev
{
    const uint32_t* qw; const uint32_t* qz; const half* sc; half* out;
    int N; int tile0;
};

struct GemvArgs
{
    const half* x; const uint32_t* x_map;
    int M, K, groups, gs_shift32;      // gs_shift32: log2(groupsize / 32)
    int spt;                           // steps per tile = K / 64 (rounded up)
    int total_tiles;
    long long total_steps;
    int num_mats;
    GemvMatDev mats[3];
    int no_zero;
    int chunk_steps;                   // x staging chunk (steps)
    int xs_stride;                     // bytes, == 64 (mod 128)
    float* partials; unsigned* counters;
    // fused prologue / epilogue
    const half* norm_w; float eps; float r_dim;
    const half* sin; const half* cos;
    int head_dim, num_heads, num_kv_heads, past_len, max_seq_len;
    half* key_cache; half* value_cache;
    float* pair_stage; unsigned* pair_counters;   // GV_EPI_SILU_MUL
};

__device__ __forceinline__ uint4 ldg_stream_v4(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_v2(const void* p)
{
    uint2 r;
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_u32(const void* p)
{
    uint32_t r;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldcg_f32(const float* p)
{
    float r;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t h2_sub(uint32_t a, uint32_t b)
{
    uint32_t r; asm("sub.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t h2_fma(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t lop_and_or(uint32_t a, uint32_t m, uint32_t o)
{
    uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(m), "r"(o)); return r;   // (a & m) | o
}

// one packed word (8 k-values of one column) -> 4 half2 registers holding (q - zp) for the nibble pairs
// (0,4) (1,5) (2,6) (3,7).  zs = half2(1024 + zp), zf = half2(-(64 + zp)).
__device__ __forceinline__ void dequant_word(uint32_t w, uint32_t zs, uint32_t zf,
                                             uint32_t& p04, uint32_t& p15, uint32_t& p26, uint32_t& p37)
{
    const uint32_t MLO = 0x000f000fu, MHI = 0x00f000f0u, EX = 0x64006400u, R16 = 0x2c002c00u;
    uint32_t w8 = w >> 8;
    p04 = h2_sub(lop_and_or(w, MLO, EX), zs);
    p15 = h2_fma(lop_and_or(w, MHI, EX), R16, zf);
    p26 = h2_sub(lop_and_or(w8, MLO, EX), zs);
    p37 = h2_fma(lop_and_or(w8, MHI, EX), R16, zf);
}

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ half silu_h(half x)
{
    // same fp16 sequence as the reference (q4_mlp.cu:27-36)
    half one = __float2half(1.0f);
    half e = hexp(__hneg(x));
    half r = hrcp(__hadd(one, e));
    return __hmul(x, r);
}

template <int PRO, int EPI>
__global__ void __launch_bounds__(THREADS, 3) q4_gemv_kernel(const GemvArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned char* xs = smem;
    float* red = reinterpret_cast<float*>(smem + (size_t)a.M * a.xs_stride);
    __shared__ unsigned s_old;
    __shared__ float s_rm[GV_MAXM];
    __shared__ float s_wsum[THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wn = warp & (WN - 1), wk = warp >> 2;
    const int g = lane >> 2, t = lane & 3;
    const int M = a.M, K = a.K, spt = a.spt;
    const long long G = gridDim.x;
    const long long S0 = (long long)blockIdx.x * a.total_steps / G;
    const long long S1 = (long long)(blockIdx.x + 1) * a.total_steps / G;

    if (PRO == GV_PRO_RMSNORM) {
        // row factor rm = half(rsqrt(mean(x^2) + eps))  (rms_norm.cu:20-79,113-116), one block reduction per row
        for (int m = 0; m < M; m++) {
            float ss = 0.f;
            const uint4* xr = reinterpret_cast<const uint4*>(a.x + (size_t)m * K);
            for (int i = tid; i < K / 8; i += THREADS) {
                uint4 v = xr[i];
                const half2* h = reinterpret_cast<const half2*>(&v);
                #pragma unroll
                for (int j = 0; j < 4; j++) { float2 f = __half22float2(h[j]); ss = fmaf(f.x, f.x, ss); ss = fmaf(f.y, f.y, ss); }
            }
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if (lane == 0) s_wsum[warp] = ss;
            __syncthreads();
            if (tid == 0) {
                float tot = 0.f;
                for (int w = 0; w < THREADS / 32; w++) tot += s_wsum[w];
                s_rm[m] = __half2float(__float2half_rn(rsqrtf(tot * a.r_dim + a.eps)));
            }
            __syncthreads();
        }
    }

    long long s = S0;
    while (s < S1) {
        const int tile = (int)(s / spt);
        const int st0 = (int)(s - (long long)tile * spt);
        const int st1 = (int)min((long long)spt, (long long)st0 + (S1 - s));
        int mi = 0;
        if (EPI == GV_EPI_SILU_MUL) {
            mi = tile & 1;                       // tiles alternate gate_j, up_j
        } else {
            if (a.num_mats > 1 && tile >= a.mats[1].tile0) mi = 1;
            if (a.num_mats > 2 && tile >= a.mats[2].tile0) mi = 2;
        }
        const uint32_t* qw = mi == 0 ? a.mats[0].qw : (mi == 1 ? a.mats[1].qw : a.mats[2].qw);
        const uint32_t* qz = mi == 0 ? a.mats[0].qz : (mi == 1 ? a.mats[1].qz : a.mats[2].qz);
        const half* scp = mi == 0 ? a.mats[0].sc : (mi == 1 ? a.mats[1].sc : a.mats[2].sc);
        half* outp = mi == 0 ? a.mats[0].out : (mi == 1 ? a.mats[1].out : a.mats[2].out);
        const int N = mi == 0 ? a.mats[0].N : (mi == 1 ? a.mats[1].N : a.mats[2].N);
        const int tile0 = mi == 0 ? a.mats[0].tile0 : (mi == 1 ? a.mats[1].tile0 : a.mats[2].tile0);
        const int ctile = (EPI == GV_EPI_SILU_MUL) ? (tile >> 1) : (tile - tile0);
        const int col_tile0 = ctile * GV_TILE_N;
        const int colbase = col_tile0 + wn * 32 + 4 * g;
        const bool col_ok = (col_tile0 + wn * 32) < N;         // warp-uniform (N % 32 == 0)
        const int zshift = (g & 1) * 16;

        float acc[8], cg[8];
        #pragma unroll
        for (int j = 0; j < 8; j++) { acc[j] = 0.f; cg[j] = 0.f; }
        int cur_grp = -1;
        float cs[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t zs[4] = {0, 0, 0, 0}, zf[4] = {0, 0, 0, 0};
        const int xrow = min(g, M - 1);

        for (int c0 = st0; c0 < st1; c0 += a.chunk_steps) {
            const int c1 = min(st1, c0 + a.chunk_steps);
            __syncthreads();
            // ---- stage x[:, c0*64 .. c1*64) into smem, permuted (0,4,1,5,2,6,3,7) inside each 8-block ----
            {
                const int nk8 = (c1 - c0) * (STEP_K / 8);
                const int k8_0 = c0 * (STEP_K / 8);
                const int k8_lim = K / 8;
                for (int idx = tid; idx < M * nk8; idx += THREADS) {
                    const int m = idx / nk8, j = idx - m * nk8;
                    const int k8 = k8_0 + j;
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (k8 < k8_lim) {
                        const half* xr = a.x + (size_t)m * K;
                        if (a.x_map) {
                            const uint32_t* mp = a.x_map + (size_t)k8 * 8;
                            unsigned short h[8];
                            #pragma unroll
                            for (int i = 0; i < 8; i++) h[i] = __half_as_ushort(xr[mp[i]]);
                            v.x = h[0] | ((uint32_t)h[1] << 16); v.y = h[2] | ((uint32_t)h[3] << 16);
                            v.z = h[4] | ((uint32_t)h[5] << 16); v.w = h[6] | ((uint32_t)h[7] << 16);
                        } else {
                            v = *reinterpret_cast<const uint4*>(xr + (size_t)k8 * 8);
                        }
                        if (PRO == GV_PRO_RMSNORM) {
                            // (x * rm) * w with two fp16 multiplies, as rms_norm_kernel (rms_norm.cu:118-131)
                            const half2 rm2 = __float2half2_rn(s_rm[m]);
                            half2* hv = reinterpret_cast<half2*>(&v);
                            if (a.x_map) {
                                const uint32_t* mp = a.x_map + (size_t)k8 * 8;
                                #pragma unroll
                                for (int i = 0; i < 4; i++) {
                                    half2 w2 = __halves2half2(a.norm_w[mp[2 * i]], a.norm_w[mp[2 * i + 1]]);
                                    hv[i] = __hmul2(__hmul2(hv[i], rm2), w2);
                                }
                            } else {
                                uint4 wv = *reinterpret_cast<const uint4*>(a.norm_w + (size_t)k8 * 8);
                                const half2* w2 = reinterpret_cast<const half2*>(&wv);
                                #pragma unroll
                                for (int i = 0; i < 4; i++) hv[i] = __hmul2(__hmul2(hv[i], rm2), w2[i]);
                            }
                        }
                    }
                    uint4 o;
                    o.x = __byte_perm(v.x, v.z, 0x5410);   // {h0, h4}
                    o.y = __byte_perm(v.x, v.z, 0x7632);   // {h1, h5}
                    o.z = __byte_perm(v.y, v.w, 0x5410);   // {h2, h6}
                    o.w = __byte_perm(v.y, v.w, 0x7632);   // {h3, h7}
                    *reinterpret_cast<uint4*>(xs + (size_t)m * a.xs_stride + (size_t)j * 16) = o;
                }
            }
            __syncthreads();

            if (col_ok) {
                uint4 wb[U]; uint2 sb[U]; uint32_t zb[U];
                const int k8_lim = K / 8;
                auto load = [&](int st, uint4& w, uint2& sc, uint32_t& zw) {
                    const int k8row = st * 8 + wk * 4 + t;
                    const int grp = a.groups == 1 ? 0 : ((st * 2 + wk) >> a.gs_shift32);
                    if (k8row < k8_lim) w = ldg_stream_v4(qw + (size_t)k8row * N + colbase);
                    else w = make_uint4(0, 0, 0, 0);
                    const int gq = min(grp, a.groups - 1);
                    sc = ldg_v2(scp + (size_t)gq * N + colbase);
                    zw = ldg_u32(qz + (size_t)gq * (N >> 3) + (colbase >> 3));
                };
                #pragma unroll
                for (int u = 0; u < U; u++) if (c0 + u < c1) load(c0 + u, wb[u], sb[u], zb[u]);

                for (int sb0 = c0; sb0 < c1; sb0 += U) {
                    #pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int st = sb0 + u;
                        if (st < c1) {
                            const uint4 w = wb[u]; const uint2 sc = sb[u]; const uint32_t zw = zb[u];
                            if (st + U < c1) load(st + U, wb[u], sb[u], zb[u]);

                            const int grp = a.groups == 1 ? 0 : ((st * 2 + wk) >> a.gs_shift32);
                            if (grp != cur_grp) {
                                if (cur_grp >= 0) {
                                    #pragma unroll
                                    for (int j = 0; j < 8; j++) { acc[j] = fmaf(cs[j >> 1], cg[j], acc[j]); cg[j] = 0.f; }
                                }
                                const half2 s01 = *reinterpret_cast<const half2*>(&sc.x);
                                const half2 s23 = *reinterpret_cast<const half2*>(&sc.y);
                                cs[0] = __low2float(s01); cs[1] = __high2float(s01);
                                cs[2] = __low2float(s23); cs[3] = __high2float(s23);
                                const uint32_t z4 = zw >> zshift;
                                #pragma unroll
                                for (int j = 0; j < 4; j++) {
                                    const uint32_t zp = ((z4 >> (4 * j)) & 0xfu) + 1u;
                                    zs[j] = (0x6400u + zp) * 0x00010001u;            // half2(1024 + zp)
                                    zf[j] = (0xd400u + (zp << 4)) * 0x00010001u;     // half2(-(64 + zp))
                                }
                                cur_grp = grp;
                            }

                            const uint4 xb = *reinterpret_cast<const uint4*>(
                                xs + (size_t)xrow * a.xs_stride + (size_t)((st - c0) * 8 + wk * 4 + t) * 16);
                            uint32_t p0[4], p1[4], p2[4], p3[4];
                            dequant_word(w.x, zs[0], zf[0], p0[0], p0[1], p0[2], p0[3]);
                            dequant_word(w.y, zs[1], zf[1], p1[0], p1[1], p1[2], p1[3]);
                            dequant_word(w.z, zs[2], zf[2], p2[0], p2[1], p2[2], p2[3]);
                            dequant_word(w.w, zs[3], zf[3], p3[0], p3[1], p3[2], p3[3]);
                            float (&cA)[4] = *reinterpret_cast<float (*)[4]>(&cg[0]);
                            float (&cB)[4] = *reinterpret_cast<float (*)[4]>(&cg[4]);
                            // columns (c0, c1): rows g / g+8 of A;  k-steps (0,4,1,5) then (2,6,3,7)
                            mma16816(cA, p0[0], p1[0], p0[1], p1[1], xb.x, xb.y);
                            mma16816(cA, p0[2], p1[2], p0[3], p1[3], xb.z, xb.w);
                            mma16816(cB, p2[0], p3[0], p2[1], p3[1], xb.x, xb.y);
                            mma16816(cB, p2[2], p3[2], p2[3], p3[3], xb.z, xb.w);
                        }
                    }
                }
            }
        }
        if (cur_grp >= 0) {
            #pragma unroll
            for (int j = 0; j < 8; j++) acc[j] = fmaf(cs[j >> 1], cg[j], acc[j]);
        }

        // ---- reduce the WK k-warps through shared memory: red[wk][m][col] ----
        __syncthreads();
        {
            float* r = red + (size_t)wk * GV_MAXM * RED_LD;
            const int cl = wn * 32 + 4 * g;
            // acc[0..1]: col cl, tokens 2t, 2t+1; acc[2..3]: col cl+1; acc[4..5]: col cl+2; acc[6..7]: col cl+3
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                r[(2 * t) * RED_LD + cl + j] = acc[2 * j];
                r[(2 * t + 1) * RED_LD + cl + j] = acc[2 * j + 1];
            }
        }
        __syncthreads();

        const int ecol = tid & (GV_TILE_N - 1);
        const int em0 = tid >> 7;                 // 0 or 1; this thread owns rows em0, em0+2, em0+4, em0+6
        float v[4];
        #pragma unroll
        for (int i = 0; i < 4; i++) {
            const int m = em0 + 2 * i;
            float sum = 0.f;
            #pragma unroll
            for (int w = 0; w < WK; w++) sum += red[(size_t)w * GV_MAXM * RED_LD + m * RED_LD + ecol];
            v[i] = sum;
        }

        bool finalize = true;
        const bool full = (st0 == 0 && st1 == spt);
        if (!full) {
            // stream-K fix-up: publish the partial, last CTA to arrive sums all partials of the tile in CTA order
            const long long tstart = (long long)tile * spt, tend = tstart + spt - 1;
            const int c_first = (int)(((tstart + 1) * G - 1) / a.total_steps);
            const int c_last = (int)(((tend + 1) * G - 1) / a.total_steps);
            const int my_first_tile = (int)(S0 / spt);
            const int slot = blockIdx.x * 2 + (tile != my_first_tile ? 1 : 0);
            float* pp = a.partials + (size_t)slot * GV_MAXM * GV_TILE_N;
            #pragma unroll
            for (int i = 0; i < 4; i++) { const int m = em0 + 2 * i; if (m < M) pp[m * GV_TILE_N + ecol] = v[i]; }
            __threadfence();
            __syncthreads();
            if (tid == 0) s_old = atomicAdd(&a.counters[tile], 1u);
            __syncthreads();
            finalize = (s_old == (unsigned)(c_last - c_first));
            if (finalize) {
                __threadfence();
                #pragma unroll
                for (int i = 0; i < 4; i++) v[i] = 0.f;
                for (int c = c_first; c <= c_last; c++) {
                    const int cft = (int)(((long long)c * a.total_steps / G) / spt);
                    const float* qp = a.partials + (size_t)(c * 2 + (tile != cft ? 1 : 0)) * GV_MAXM * GV_TILE_N;
                    #pragma unroll
                    for (int i = 0; i < 4; i++) { const int m = em0 + 2 * i; if (m < M) v[i] += ldcg_f32(qp + m * GV_TILE_N + ecol); }
                }
                if (tid == 0) a.counters[tile] = 0u;     // ready for the next launch (stream-ordered)
            }
        }

        if (finalize) {
            const int col = col_tile0 + ecol;
            if (EPI == GV_EPI_STORE) {
                if (col < N) {
                    #pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int m = em0 + 2 * i;
                        if (m < M) {
                            float r = v[i];
                            half* o = outp + (size_t)m * N + col;
                            if (a.no_zero) r += __half2float(*o);
                            *o = __float2half_rn(r);
                        }
                    }
                }
            } else if (EPI == GV_EPI_ROPE_CACHE) {
                // tile == one head (head_dim == 128): rope on q / k (rope.cu:48-67), k/v written to the cache
                // (q4_attn.cu:32-51).  Rows are q_len * bsz tokens; decode path has bsz == 1 (model.py:528).
                __syncthreads();
                half* hs = reinterpret_cast<half*>(red);          // [m][128]
                #pragma unroll
                for (int i = 0; i < 4; i++) { const int m = em0 + 2 * i; if (m < M) hs[m * GV_TILE_N + ecol] = __float2half_rn(v[i]); }
                __syncthreads();
                const int head = col_tile0 / GV_TILE_N;
                #pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int m = em0 + 2 * i;
                    if (m < M && col < N) {
                        half val = hs[m * GV_TILE_N + ecol];
                        if (mi < 2) {
                            const int pos = a.past_len + m;            // row m of q_len rows, one batch
                            const half* sr = a.sin + (size_t)pos * GV_TILE_N;
                            const half* cr = a.cos + (size_t)pos * GV_TILE_N;
                            const half other = hs[m * GV_TILE_N + (ecol ^ 64)];
                            if (ecol < 64) val = __hfma(val, cr[ecol], __hmul(other, __hneg(sr[ecol])));
                            else           val = __hfma(val, cr[ecol], __hmul(other, sr[ecol]));
                        }
                        outp[(size_t)m * N + col] = val;
                        if (mi >= 1) {
                            half* cache = mi == 1 ? a.key_cache : a.value_cache;
                            cache[((size_t)head * a.max_seq_len + a.past_len + m) * GV_TILE_N + ecol] = val;
                        }
                    }
                }
                __syncthreads();
            } else if (EPI == GV_EPI_SILU_MUL) {
                // mats = {gate, up}: tiles alternate gate_j, up_j.  Whichever of the pair finishes second
                // combines silu(gate) * up (q4_mlp.cu:27-36,46-88) and writes mats[0].out.
                const int pair = ctile;
                float* stg = a.pair_stage + ((size_t)pair * 2 + mi) * GV_MAXM * GV_TILE_N;
                #pragma unroll
                for (int i = 0; i < 4; i++) { const int m = em0 + 2 * i; if (m < M) stg[m * GV_TILE_N + ecol] = v[i]; }
                __threadfence();
                __syncthreads();
                if (tid == 0) s_old = atomicAdd(&a.pair_counters[pair], 1u);
                __syncthreads();
                if (s_old == 1u) {
                    __threadfence();
                    const float* og = a.pair_stage + ((size_t)pair * 2 + 0) * GV_MAXM * GV_TILE_N;
                    const float* ou = a.pair_stage + ((size_t)pair * 2 + 1) * GV_MAXM * GV_TILE_N;
                    if (col < N) {
                        #pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int m = em0 + 2 * i;
                            if (m < M) {
                                const half gt = __float2half_rn(ldcg_f32(og + m * GV_TILE_N + ecol));
                                const half up = __float2half_rn(ldcg_f32(ou + m * GV_TILE_N + ecol));
                                a.mats[0].out[(size_t)m * N + col] = __hmul(silu_h(gt), up);
                            }
                        }
                    }
                    if (tid == 0) a.pair_counters[pair] = 0u;
                }
            }
        }
        s += st1 - st0;
    }
}

template <int PRO, int EPI>
int launch_cfg(ExlDevice* ds, const GemvArgs& a, size_t smem, cudaStream_t stream)
{
    auto kern = q4_gemv_kernel<PRO, EPI>;
    static int ctas_per_sm_cache[GV_MAXM + 1] = {0};
    static size_t smem_cache[GV_MAXM + 1] = {0};
    static int attr_device_done[EXL_MAX_DEVICES] = {0};
    if (!attr_device_done[ds->device]) {
        EXL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_device_done[ds->device] = 1;
    }
    if (smem_cache[a.M] != smem || ctas_per_sm_cache[a.M] == 0) {
        int nb = 0;
        EXL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, THREADS, smem));
        if (nb < 1) return exl_set_err(EXL_ERR_CUDA, "q4_gemv: kernel does not fit (smem %zu)", smem);
        ctas_per_sm_cache[a.M] = nb; smem_cache[a.M] = smem;
    }
    long long grid = (long long)ds->num_sms * ctas_per_sm_cache[a.M];
    if (grid > a.total_steps) grid = a.total_steps;
    if (grid > GV_MAX_CTAS) grid = GV_MAX_CTAS;
    kern<<<(unsigned)grid, THREADS, smem, stream>>>(a);
    EXL_CHECK_LAUNCH("q4_gemv_kernel");
    return EXL_OK;
}

} // namespace

int exl_gemv_launch(ExlDevice* ds, const half* x, int M, const exl_q4_matrix* const* mats, half* const* outs,
                    int num_mats, bool no_zero, int prologue, int epilogue, const GemvFused* fused, cudaStream_t stream)
{
    if (M < 1 || M > GV_MAXM) return exl_set_err(EXL_ERR_ARG, "q4_gemv: M=%d out of range [1,%d]", M, GV_MAXM);
    if (num_mats < 1 || num_mats > 3) return exl_set_err(EXL_ERR_ARG, "q4_gemv: num_mats=%d", num_mats);
    const exl_q4_matrix* w0 = mats[0];
    GemvArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.x_map = w0->x_map; a.M = M; a.K = w0->K; a.groups = w0->groups;
    if (w0->K % 32 != 0) return exl_set_err(EXL_ERR_ARG, "q4_gemv: K=%d must be a multiple of 32", w0->K);
    int gs32 = w0->groupsize / 32;
    if (w0->groups > 1 && (w0->groupsize % 32 != 0 || (gs32 & (gs32 - 1)) != 0))
        return exl_set_err(EXL_ERR_ARG, "q4_gemv: groupsize=%d must be 32 * 2^n", w0->groupsize);
    int sh = 0; while ((1 << sh) < gs32) sh++;
    a.gs_shift32 = sh;
    a.spt = (w0->K + STEP_K - 1) / STEP_K;
    int tiles = 0;
    for (int i = 0; i < num_mats; i++) {
        const exl_q4_matrix* w = mats[i];
        if (w->K != w0->K || w->groups != w0->groups || w->x_map != w0->x_map)
            return exl_set_err(EXL_ERR_ARG, "q4_gemv: fused matrices must share K, groups and x_map");
        if (w->N % 32 != 0) return exl_set_err(EXL_ERR_ARG, "q4_gemv: N=%d must be a multiple of 32", w->N);
        a.mats[i].qw = w->qweight; a.mats[i].qz = w->qzeros; a.mats[i].sc = w->scales; a.mats[i].out = outs[i];
        a.mats[i].N = w->N; a.mats[i].tile0 = tiles;
        tiles += (w->N + GV_TILE_N - 1) / GV_TILE_N;
    }
    if (epilogue == GV_EPI_SILU_MUL) {
        // gate/up tiles are interleaved (g0,u0,g1,u1,...) so the two halves of a pair finish close together
        if (num_mats != 2 || mats[0]->N != mats[1]->N)
            return exl_set_err(EXL_ERR_ARG, "q4_gemv: SILU_MUL epilogue needs {gate, up} of equal width");
        if (tiles / 2 > GV_MAX_PAIRS) return exl_set_err(EXL_ERR_ARG, "q4_gemv: too many gate/up tiles");
        a.pair_stage = ds->gemv_pair_stage; a.pair_counters = ds->gemv_pair_counters;
    }
    if (tiles > GV_MAX_TILES) return exl_set_err(EXL_ERR_ARG, "q4_gemv: too many tiles (%d)", tiles);
    a.num_mats = num_mats; a.total_tiles = tiles; a.total_steps = (long long)tiles * a.spt;
    a.no_zero = no_zero ? 1 : 0;
    // x staging chunk: keep M * chunk_k * 2 bytes <= 64 KB
    int chunk_k = (64 * 1024) / (2 * M);
    chunk_k = (chunk_k / STEP_K) * STEP_K;
    int chunk_steps = chunk_k / STEP_K;
    if (chunk_steps > a.spt) chunk_steps = a.spt;
    a.chunk_steps = chunk_steps;
    a.xs_stride = chunk_steps * STEP_K * 2 + 64;
    a.partials = ds->gemv_partials; a.counters = ds->gemv_counters;
    if (fused) {
        a.norm_w = fused->norm_w; a.eps = fused->eps; a.r_dim = 1.0f / (float)w0->K;
        a.sin = fused->sin; a.cos = fused->cos; a.head_dim = fused->head_dim; a.num_heads = fused->num_heads;
        a.num_kv_heads = fused->num_kv_heads; a.past_len = fused->past_len; a.max_seq_len = fused->max_seq_len;
        a.key_cache = fused->key_cache; a.value_cache = fused->value_cache;
    }
    size_t smem = (size_t)M * a.xs_stride + (size_t)WK * GV_MAXM * RED_LD * sizeof(float);

    if (prologue == GV_PRO_PLAIN && epilogue == GV_EPI_STORE) return launch_cfg<GV_PRO_PLAIN, GV_EPI_STORE>(ds, a, smem, stream);
    if (prologue == GV_PRO_RMSNORM && epilogue == GV_EPI_STORE) return launch_cfg<GV_PRO_RMSNORM, GV_EPI_STORE>(ds, a, smem, stream);
    if (prologue == GV_PRO_RMSNORM && epilogue == GV_EPI_ROPE_CACHE) {
        if (!fused || fused->head_dim != GV_TILE_N) return exl_set_err(EXL_ERR_ARG, "q4_gemv: rope epilogue needs head_dim == 128");
        return launch_cfg<GV_PRO_RMSNORM, GV_EPI_ROPE_CACHE>(ds, a, smem, stream);
    }
    if (prologue == GV_PRO_RMSNORM && epilogue == GV_EPI_SILU_MUL) return launch_cfg<GV_PRO_RMSNORM, GV_EPI_SILU_MUL>(ds, a, smem, stream);
    if (prologue == GV_PRO_PLAIN && epilogue == GV_EPI_SILU_MUL) return launch_cfg<GV_PRO_PLAIN, GV_EPI_SILU_MUL>(ds, a, smem, stream);
    return exl_set_err(EXL_ERR_ARG, "q4_gemv: unsupported prologue/epilogue combination %d/%d", prologue, epilogue);
}
