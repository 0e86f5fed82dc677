// q4_gemv.cu -- decode hot path: fused 4-bit unpack + scale/zero + skinny GEMM (M <= 8) for sm_100a.
//
// Replaces the reference's q4_matmul_kernel + dot_product_8* (exllama_ext/cuda_func/q4_matmul.cu:34-212,
// exllama_ext/matrix.cuh:87-286) and, through the fused prologue/epilogues, rms_norm (rms_norm.cu),
// rope (rope.cu), update_cache_kernel (q4_attn.cu:19-72) and silu_mul (q4_mlp.cu:46-88) on the decode path.
//
// Design (HBM-bound; see DESIGN.md "decode kernel"):
//  * packed qweight [K/8, N] is streamed exactly once with 128-bit loads (ld.global.nc.L1::no_allocate.v4):
//    lane (g = lane/4, t = lane%4) of a warp loads the 4 columns 4g..4g+3 of k8-row t, so one warp
//    request covers 4 rows x 128 B and a CTA step (8 warps) covers 128 columns x 64 k.  GV_U steps are kept
//    in flight per thread in registers.
//  * group scales / zeros of the CTA's k-range are staged into shared memory with 1-D TMA bulk copies
//    (cp.async.bulk + mbarrier); x is staged by the threads because it is permuted (and gathered through the
//    act-order x_map, and RMS-normalised for the fused decoder ops) on the way in.
//  * nibbles are expanded two at a time with the 0x6400 fp16 magic (q | 0x6400 == 1024 + q), the zero point is
//    folded into the bias removal, and the products are formed by mma.sync m16n8k16 with the *weights as the
//    A operand straight from registers* (rows = 16 output columns) and the <= 8 activation rows as the n = 8
//    operand: exact fp16 products, fp32 accumulation, cost independent of M for M <= 8.
//    K is traversed in the permuted order (0,4,1,5,2,6,3,7) inside each 8-block so no nibble shuffling is needed.
//  * group scales are applied once per group to an fp32 group accumulator.
//  * work is a flat list of (128-column tile, 64-k step) items split evenly over a persistent grid
//    (stream-K): every CTA streams the same number of bytes whatever N and K are.  Tiles that span several
//    CTAs are finished by the last CTA to arrive (partials in an L2-resident workspace, fixed summation
//    order => deterministic; no fp16 atomics unlike the reference, q4_matmul.cu:203-211).
//  * programmatic dependent launch: weights do not depend on the previous kernel, so a CTA prefetches its first
//    GV_U steps of weights and its scales/zeros BEFORE griddepcontrol.wait, and signals launch_dependents as soon
//    as its main loop is done -- back-to-back GEMVs keep HBM busy across the launch boundary.
#include "exl_common.cuh"
#include <cstdlib>
#include <cstring>

#ifndef GV_U
#define GV_U 4                      // register prefetch depth (steps)
#endif
#ifndef GV_MINB
#define GV_MINB 2                   // min CTAs per SM for __launch_bounds__
#endif

namespace {

constexpr int THREADS = 256;
constexpr int WN = 4;              // warps across columns (4 x 32 = 128 columns)
constexpr int WK = 2;              // warps across k
constexpr int STEP_K = 64;         // k per CTA step
constexpr int U = GV_U;
constexpr int RED_LD = GV_TILE_N + 4;
constexpr int GMAXC = 64;          // max quantisation groups staged per chunk
constexpr int SC_ROW = GV_TILE_N * 2;      // bytes of scales per group row in smem
constexpr int ZQ_ROW = GV_TILE_N / 2;      // bytes of packed zeros per group row in smem

struct GemvMatDev
{
    const uint32_t* qw; const uint32_t* qz; const half* sc; half* out;
    int N; int tile0;
};

struct GemvArgs
{
    const half* x; const uint32_t* x_map;
    int M, K, groups, gs_shift32;      // gs_shift32: log2(groupsize / 32) (30 when there is a single group)
    int spt;                           // steps per tile = ceil(K / 64)
    int total_tiles;
    long long total_steps;
    int num_mats;
    GemvMatDev mats[3];
    int no_zero;
    int chunk_steps;                   // staging chunk (steps)
    int xs_stride;                     // bytes, == 64 (mod 128)
    float* partials; unsigned* counters;
    // fused prologue / epilogue
    const half* norm_w; float eps; float r_dim;
    const half* sin; const half* cos;
    int head_dim, num_heads, num_kv_heads, past_len, max_seq_len;
    half* key_cache; half* value_cache;
    float* pair_stage; unsigned* pair_counters;   // GV_EPI_SILU_MUL
    int debug;                         // EXL_GV_DEBUG bitmask (profiling experiments only)
};

// predicated streaming load: keeps the old register contents when pred is false
__device__ __forceinline__ void ldg_stream_v4_pred(uint4& r, const void* p, bool pred)
{
    asm volatile("{\n\t.reg .pred pp;\n\tsetp.ne.b32 pp, %5, 0;\n\t"
                 "@pp ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
                 : "+r"(r.x), "+r"(r.y), "+r"(r.z), "+r"(r.w) : "l"(p), "r"((int)pred));
}
__device__ __forceinline__ float ldcg_f32(const float* p)
{
    float r;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned atom_add_acq_rel(unsigned* p, unsigned v)
{
    unsigned old;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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

// Per-thread state of the inner product loop for one tile segment.
struct Accum
{
    float acc[8];      // scaled totals
    float cg[8];       // current group, unscaled
    float cs[4];       // current group's scales (4 columns)
    uint32_t zs[4], zf[4];
    int cur_grp;
};

__device__ __forceinline__ void group_switch(Accum& A, int grp, int g_lo, const unsigned char* sc_s, const unsigned char* zq_s,
                                             int lane_col /* wn*32 + 4g */)
{
    if (A.cur_grp >= 0) {
        #pragma unroll
        for (int j = 0; j < 8; j++) { A.acc[j] = fmaf(A.cs[j >> 1], A.cg[j], A.acc[j]); A.cg[j] = 0.f; }
    }
    const int gl = grp - g_lo;
    const uint2 sc = *reinterpret_cast<const uint2*>(sc_s + gl * SC_ROW + lane_col * 2);
    const uint32_t zw = *reinterpret_cast<const uint32_t*>(zq_s + gl * ZQ_ROW + (lane_col >> 3) * 4);
    const half2 s01 = *reinterpret_cast<const half2*>(&sc.x);
    const half2 s23 = *reinterpret_cast<const half2*>(&sc.y);
    A.cs[0] = __low2float(s01); A.cs[1] = __high2float(s01);
    A.cs[2] = __low2float(s23); A.cs[3] = __high2float(s23);
    const uint32_t z4 = zw >> ((lane_col & 4) * 4);
    #pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t zp = ((z4 >> (4 * j)) & 0xfu) + 1u;
        A.zs[j] = (0x6400u + zp) * 0x00010001u;            // half2(1024 + zp)
        A.zf[j] = (0xd400u + (zp << 4)) * 0x00010001u;     // half2(-(64 + zp))
    }
    A.cur_grp = grp;
}

__device__ __forceinline__ void unit_mma(Accum& A, const uint4& w, const uint4& xb)
{
    uint32_t p0[4], p1[4], p2[4], p3[4];
    dequant_word(w.x, A.zs[0], A.zf[0], p0[0], p0[1], p0[2], p0[3]);
    dequant_word(w.y, A.zs[1], A.zf[1], p1[0], p1[1], p1[2], p1[3]);
    dequant_word(w.z, A.zs[2], A.zf[2], p2[0], p2[1], p2[2], p2[3]);
    dequant_word(w.w, A.zs[3], A.zf[3], p3[0], p3[1], p3[2], p3[3]);
    float (&cA)[4] = *reinterpret_cast<float (*)[4]>(&A.cg[0]);
    float (&cB)[4] = *reinterpret_cast<float (*)[4]>(&A.cg[4]);
    // columns (c0, c1): rows g / g+8 of A;  k order (0,4,1,5) then (2,6,3,7)
    mma16816(cA, p0[0], p1[0], p0[1], p1[1], xb.x, xb.y);
    mma16816(cA, p0[2], p1[2], p0[3], p1[3], xb.z, xb.w);
    mma16816(cB, p2[0], p3[0], p2[1], p3[1], xb.x, xb.y);
    mma16816(cB, p2[2], p3[2], p2[3], p3[3], xb.z, xb.w);
}

template <int PRO, int EPI>
__global__ void __launch_bounds__(THREADS, GV_MINB) q4_gemv_kernel(const GemvArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    // layout: [sc_s: GMAXC * 256][zq_s: GMAXC * 64][red: WK * 8 * RED_LD * 4][xs: M * xs_stride]
    unsigned char* sc_s = smem;
    unsigned char* zq_s = smem + GMAXC * SC_ROW;
    float* red = reinterpret_cast<float*>(smem + GMAXC * (SC_ROW + ZQ_ROW));
    unsigned char* xs = smem + GMAXC * (SC_ROW + ZQ_ROW) + WK * GV_MAXM * RED_LD * sizeof(float);
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ unsigned s_old;
    __shared__ float s_rm[GV_MAXM];
    __shared__ float s_wsum[THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wn = warp & (WN - 1), wk = warp >> 2;
    const int g = lane >> 2, t = lane & 3;
    const int lane_col = wn * 32 + 4 * g;
    const int M = a.M, K = a.K, spt = a.spt;
    const int k8_lim = K >> 3;
    const long long G = gridDim.x;
    const long long S0 = (long long)blockIdx.x * a.total_steps / G;
    const long long S1 = (long long)(blockIdx.x + 1) * a.total_steps / G;

    if (tid == 0) mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t bar_parity = 0;
    bool waited_dep = false;       // griddepcontrol.wait executed (x / out / workspace may be touched after it)

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
        const int tile_cols = min(GV_TILE_N, N - col_tile0);
        const bool col_ok = (wn * 32) < tile_cols;             // warp-uniform (N % 32 == 0)

        Accum A;
        #pragma unroll
        for (int j = 0; j < 8; j++) { A.acc[j] = 0.f; A.cg[j] = 0.f; }
        #pragma unroll
        for (int j = 0; j < 4; j++) { A.cs[j] = 0.f; A.zs[j] = 0; A.zf[j] = 0; }
        A.cur_grp = -1;
        const int xrow = min(g, M - 1);

        for (int c0 = st0; c0 < st1; c0 += a.chunk_steps) {
            const int c1 = min(st1, c0 + a.chunk_steps);
            const int n = c1 - c0;
            // this warp's unit of step c is k8-rows c * 8 + wk * 4 + [0, 4); drop the unit past K (K % 64 == 32)
            int nw = n;
            if ((c1 - 1) * 8 + wk * 4 + 4 > k8_lim) nw = n - 1;
            const int g_lo = a.groups == 1 ? 0 : ((c0 * 2) >> a.gs_shift32);
            const int g_hi = a.groups == 1 ? 0 : min(a.groups - 1, ((c1 * 2 - 1) >> a.gs_shift32));

            __syncthreads();                 // previous chunk / segment finished with sc_s, zq_s, xs, red
            // ---- scales / zeros of groups [g_lo, g_hi] for this tile's columns: TMA bulk copies onto s_bar ----
            if (warp == 0) {
                const int ng = g_hi - g_lo + 1;
                const uint32_t sc_bytes = (uint32_t)tile_cols * 2, zq_bytes = (uint32_t)tile_cols / 2;
                if (lane == 0) mbar_expect_tx(&s_bar, (uint32_t)ng * (sc_bytes + zq_bytes));
                __syncwarp();
                for (int gi = lane; gi < ng; gi += 32) {
                    bulk_g2s(sc_s + gi * SC_ROW, scp + (size_t)(g_lo + gi) * N + col_tile0, sc_bytes, &s_bar);
                    bulk_g2s(zq_s + gi * ZQ_ROW, qz + (size_t)(g_lo + gi) * (N >> 3) + (col_tile0 >> 3), zq_bytes, &s_bar);
                }
            }
            // ---- weight prefetch: the first U steps of this chunk (independent of the previous kernel) ----
            const uint4* wp = reinterpret_cast<const uint4*>(qw + (size_t)(c0 * 8 + wk * 4 + t) * N + col_tile0 + lane_col);
            const size_t wstep = (size_t)2 * N;          // uint4 per step (8 k8-rows)
            uint4 wb[U];
            #pragma unroll
            for (int u = 0; u < U; u++) {
                wb[u] = make_uint4(0, 0, 0, 0);
                ldg_stream_v4_pred(wb[u], wp + (size_t)u * wstep, col_ok && u < nw && !(a.debug & 1));
            }

            if (!waited_dep) {
                pdl_wait();                  // everything below may read x / write out and the shared workspace
                waited_dep = true;
                if (PRO == GV_PRO_RMSNORM) {
                    // row factor rm = half(rsqrt(mean(x^2) + eps))  (rms_norm.cu:20-79,113-116)
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
            }

            // ---- stage x[:, c0*64 .. c1*64) into smem, permuted (0,4,1,5,2,6,3,7) inside each 8-block ----
            {
                const int nk8 = n * (STEP_K / 8);
                const int k8_0 = c0 * (STEP_K / 8);
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
            mbar_wait(&s_bar, bar_parity);   // scales / zeros landed
            bar_parity ^= 1;
            __syncthreads();                 // x staged

            if (col_ok && !(a.debug & 1)) {
                const unsigned char* xq = xs + (size_t)xrow * a.xs_stride + (size_t)(wk * 4 + t) * 16;   // + step * 128
                const int ubase = c0 * 2 + wk;
                const uint4* wq = wp + (size_t)U * wstep;
                int i = 0;
                const int nfull = (nw / U) * U;
                for (; i < nfull; i += U) {
                    #pragma unroll
                    for (int u = 0; u < U; u++) {
                        const uint4 w = wb[u];
                        ldg_stream_v4_pred(wb[u], wq + (size_t)(i + u) * wstep, (i + u + U) < nw);
                        const int grp = (ubase + 2 * (i + u)) >> a.gs_shift32;
                        if (grp != A.cur_grp) group_switch(A, grp, g_lo, sc_s, zq_s, lane_col);
                        const uint4 xb = *reinterpret_cast<const uint4*>(xq + (size_t)(i + u) * 128);
                        unit_mma(A, w, xb);
                    }
                }
                #pragma unroll
                for (int u = 0; u < U; u++) {
                    if (i + u < nw) {
                        const int grp = (ubase + 2 * (i + u)) >> a.gs_shift32;
                        if (grp != A.cur_grp) group_switch(A, grp, g_lo, sc_s, zq_s, lane_col);
                        const uint4 xb = *reinterpret_cast<const uint4*>(xq + (size_t)(i + u) * 128);
                        unit_mma(A, wb[u], xb);
                    }
                }
            }
        }
        if (A.cur_grp >= 0) {
            #pragma unroll
            for (int j = 0; j < 8; j++) A.acc[j] = fmaf(A.cs[j >> 1], A.cg[j], A.acc[j]);
        }
        const bool last_segment = (s + (st1 - st0)) >= S1;
        if (last_segment) pdl_launch_dependents();       // this CTA has issued all its weight loads

        // ---- reduce the WK k-warps through shared memory: red[wk][m][col] ----
        {
            float* r = red + (size_t)wk * GV_MAXM * RED_LD;
            // acc[0..1]: col lane_col, tokens 2t, 2t+1; acc[2..3]: col +1; acc[4..5]: col +2; acc[6..7]: col +3
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                r[(2 * t) * RED_LD + lane_col + j] = A.acc[2 * j];
                r[(2 * t + 1) * RED_LD + lane_col + j] = A.acc[2 * j + 1];
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
        const bool full = (st0 == 0 && st1 == spt) || (a.debug & 2);
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
            __syncthreads();
            if (tid == 0) s_old = atom_add_acq_rel(&a.counters[tile], 1u);    // release our partial / acquire the others'
            __syncthreads();
            finalize = (s_old == (unsigned)(c_last - c_first));
            if (finalize) {
                #pragma unroll
                for (int i = 0; i < 4; i++) v[i] = 0.f;
                // only c_first can contribute its *second* slot (when it started in an earlier tile)
                const long long start_first = (long long)c_first * a.total_steps / G;
                const int first_par = (start_first < tstart) ? 1 : 0;
                for (int cb = c_first; cb <= c_last; cb += 8) {
                    float pv[8][4];
                    #pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int c = cb + j;
                        const float* qp = a.partials + (size_t)(c * 2 + (c == c_first ? first_par : 0)) * GV_MAXM * GV_TILE_N;
                        #pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int m = em0 + 2 * i;
                            pv[j][i] = (c <= c_last && m < M) ? ldcg_f32(qp + m * GV_TILE_N + ecol) : 0.f;
                        }
                    }
                    #pragma unroll
                    for (int j = 0; j < 8; j++)
                        #pragma unroll
                        for (int i = 0; i < 4; i++) v[i] += pv[j][i];
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
                // (q4_attn.cu:32-51).  Rows are q_len tokens of one sequence (decode path, model.py:528).
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
                            const int pos = a.past_len + m;
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
            } else if (EPI == GV_EPI_SILU_MUL) {
                // mats = {gate, up}: tiles alternate gate_j, up_j.  Whichever of the pair finishes second
                // combines silu(gate) * up (q4_mlp.cu:27-36,46-88) and writes mats[0].out.
                const int pair = ctile;
                float* stg = a.pair_stage + ((size_t)pair * 2 + mi) * GV_MAXM * GV_TILE_N;
                #pragma unroll
                for (int i = 0; i < 4; i++) { const int m = em0 + 2 * i; if (m < M) stg[m * GV_TILE_N + ecol] = v[i]; }
                __syncthreads();
                if (tid == 0) s_old = atom_add_acq_rel(&a.pair_counters[pair], 1u);
                __syncthreads();
                if (s_old == 1u) {
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
    static int use_pdl = -1;
    if (use_pdl < 0) { const char* e = getenv("EXL_GV_PDL"); use_pdl = e ? atoi(e) : 1; }
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
    int cps = ctas_per_sm_cache[a.M];
    if (const char* e = getenv("EXL_GV_CPS")) { int v = atoi(e); if (v >= 1 && v < cps) cps = v; }
    long long grid = (long long)ds->num_sms * cps;
    if (grid > a.total_steps) grid = a.total_steps;
    if (grid > GV_MAX_CTAS) grid = GV_MAX_CTAS;

    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return exl_set_err(EXL_ERR_CUDA, "launch of q4_gemv_kernel failed: %s", cudaGetErrorString(e));
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
    a.gs_shift32 = w0->groups == 1 ? 30 : sh;
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
    // staging chunk: M * chunk_k * 2 bytes of x <= 32 KB, and at most GMAXC quantisation groups
    int chunk_k = (32 * 1024) / (2 * M);
    if (w0->groups > 1) { int gk = (GMAXC - 1) * w0->groupsize; if (gk < chunk_k) chunk_k = gk; }
    chunk_k = (chunk_k / STEP_K) * STEP_K;
    int chunk_steps = chunk_k / STEP_K;
    if (chunk_steps < 1) return exl_set_err(EXL_ERR_ARG, "q4_gemv: cannot stage a chunk (groupsize %d)", w0->groupsize);
    if (chunk_steps > a.spt) chunk_steps = a.spt;
    a.chunk_steps = chunk_steps;
    a.xs_stride = chunk_steps * STEP_K * 2 + 64;
    a.partials = ds->gemv_partials; a.counters = ds->gemv_counters;
    if (const char* e = getenv("EXL_GV_DEBUG")) a.debug = atoi(e);
    if (fused) {
        a.norm_w = fused->norm_w; a.eps = fused->eps; a.r_dim = 1.0f / (float)w0->K;
        a.sin = fused->sin; a.cos = fused->cos; a.head_dim = fused->head_dim; a.num_heads = fused->num_heads;
        a.num_kv_heads = fused->num_kv_heads; a.past_len = fused->past_len; a.max_seq_len = fused->max_seq_len;
        a.key_cache = fused->key_cache; a.value_cache = fused->value_cache;
    }
    size_t smem = (size_t)GMAXC * (SC_ROW + ZQ_ROW) + (size_t)WK * GV_MAXM * RED_LD * sizeof(float) + (size_t)M * a.xs_stride;

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
