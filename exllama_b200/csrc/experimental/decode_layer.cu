// decode_layer.cu -- EXPERIMENTAL round-2 draft.  NOT part of the default build (exllama_b200/_build.py does not list it)
// and NOT yet run on hardware: it compiles for sm_100a (see the command at the bottom) and is kept so that the next round
// starts from code instead of a plan.  Nothing in the product, the tests or the bench depends on it.
//
// One persistent kernel per decoder layer for the decode step (M = 1), instead of five launches.
//
// Why: the five per-layer launches are latency-bound -- each pays launch + setup + an x-dependent
// prologue + reductions + epilogue (~4 us) around 1.3-9 us of streaming, and nothing overlaps across launch boundaries
// because two ~100 KB CTAs per SM leave no room for the next kernel.  Weights, scales and zeros do not depend on the
// activations, so a persistent kernel whose producer warps walk a static schedule of (phase, tile, K-slice) units can
// keep the TMA ring full across phase boundaries; only the small x-dependent steps wait for a grid barrier.
//
//   phase QKV   x --rmsnorm(ln1)--> q,k,v GEMV (96 tiles) --> rope(q,k); k,v rows -> cache; q -> scratch
//   -- grid barrier --
//   phase ATT   softmax(q K^T / sqrt(d)) V per head, the 8 CTAs of a cluster split the sequence  --> attn_out
//   -- grid barrier --
//   phase O     attn_out --> o_proj GEMV (32 tiles), x += result, per-tile sum of squares of the new x
//   -- grid barrier --
//   phase GU    x --rmsnorm(ln2, from the per-tile sums)--> gate / up GEMV (86 tile pairs) --> silu(gate)*up -> act
//   -- grid barrier --
//   phase DOWN  act --> down GEMV (32 tiles), x += result, per-tile sum of squares (for the next layer's ln1)
//
// Work split: clusters of CS = 8 CTAs; a cluster takes column tile u = cluster_id, cluster_id + n_clusters, ...; the CTA
// of rank r takes the r-th K slice; the 8 partials are exchanged through DSMEM (st.async + mbarrier, as q4_gemv.cu) and the
// leader runs the epilogue.  Restrictions of this draft: M = 1, groupsize 128 (group == ring stage), no act-order,
// head_dim 128, kv_heads == heads, hidden % 1024 == 0, single GPU.
//
// Every wait is bounded and traps on expiry (a wrong schedule must abort the launch, not hang the GPU).
#include "../exl_common.cuh"
#include <cstring>

namespace {

constexpr int CS = 8;                        // cluster size == K split factor
constexpr int CONSUMERS = 256, THREADS = CONSUMERS + 32;
constexpr int WN = 4, WK = 2;
constexpr int STAGE_ROWS = 16, STAGE_K = 128;
constexpr int BOX_COLS = 32, BOX_BYTES = STAGE_ROWS * BOX_COLS * 4;          // 2 KB
constexpr int W_BYTES = WN * BOX_BYTES;                                      // 8 KB of packed weights per stage
constexpr int META_SC = W_BYTES, META_ZQ = W_BYTES + 256;                     // the stage's group row: 128 scales, 128 zero nibbles
constexpr int STAGE_STRIDE = W_BYTES + 1024;                                  // keeps every stage 1 KB aligned (swizzle atom)
constexpr int NST = 8;
constexpr int TILE_N = 128;
constexpr int MAX_SLICE_STAGES = 16;          // K slice of one CTA, in stages (down_proj of 65B: 172 / 8 = 22 -> raise when needed)
constexpr unsigned SPIN_LIMIT = 1u << 26;
constexpr uint32_t XCH_BYTES = (CS - 1) * 130 * 4;   // every exchange moves 130 floats per peer (GEMV tiles pad with two zeros): one arming size

enum { PH_QKV = 0, PH_ATT = 1, PH_O = 2, PH_GU = 3, PH_DOWN = 4 };

struct LayerArgs
{
    CUtensorMap tm[7];                       // q k v o gate up down: qweight [K/8, N], box 16 x 32, SWIZZLE_128B (exl_q4_matrix::tmap_w)
    const uint32_t* qz[7]; const half* sc[7];
    int hidden, inter, heads;
    half* x;                                 // [hidden] residual stream, updated in place
    const half* ln1; const half* ln2; float eps;
    const half* sin; const half* cos; int past_len, max_seq;
    half* kc; half* vc;                      // this layer's caches [heads, max_seq, 128]
    half* q_buf;                             // [hidden] q after rope
    half* attn_out;                          // [hidden]
    half* act;                               // [inter]
    float* ssq_in;                           // [hidden/128] per-tile sum of squares of x (written by the previous layer), or nullptr
    float* ssq_mid;                          // [hidden/128] after o_proj (this launch)
    float* ssq_out;                          // [hidden/128] after down (for the next layer)
    unsigned* grid_bar;                      // {count, generation}
    int n_clusters;
};

// ------------------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar_addr) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar_addr) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar_addr, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar_addr), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar_addr, uint32_t parity)
{
    for (unsigned i = 0; !mbar_try(bar_addr, parity); i++) if (i > SPIN_LIMIT) __trap();
}
// remote arrive on an mbarrier of another CTA of the cluster (release at cluster scope: orders this thread's earlier reads
// of the exchange slots before the peer's next deposit)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_cluster(uint32_t bar_addr, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar_addr), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank)
{
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_async_f32(uint32_t addr, float v, uint32_t mbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" :: "r"(addr), "f"(v), "r"(mbar) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void imma_u8s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void imma_u8u8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ half silu_h(half x)
{
    // same fp16 sequence as the reference (q4_mlp.cu:27-36)
    half one = __float2half(1.0f);
    half e = hexp(__hneg(x));
    half r = hrcp(__hadd(one, e));
    return __hmul(x, r);
}

// Quantise one k8-row (8 halves) with scale 1/inv: sum of x_q and the byte planes {a even, a odd, b even, b odd}, x_q = 256 a + b
// (identical to q4_gemv.cu quantise_row)
__device__ __forceinline__ int quantise_row(const uint4& hv, float inv, uint4& o)
{
    const half2* h = reinterpret_cast<const half2*>(&hv);
    int sum = 0;
    o = make_uint4(0, 0, 0, 0);
    #pragma unroll
    for (int i = 0; i < 4; i++) {
        const float2 f = __half22float2(h[i]);
        const int q0 = __float2int_rn(f.x * inv), q1 = __float2int_rn(f.y * inv);
        sum += q0 + q1;
        o.x |= (uint32_t)((q0 >> 8) & 0xff) << (8 * i);
        o.y |= (uint32_t)((q1 >> 8) & 0xff) << (8 * i);
        o.z |= (uint32_t)(q0 & 0xff) << (8 * i);
        o.w |= (uint32_t)(q1 & 0xff) << (8 * i);
    }
    return sum;
}
__device__ __forceinline__ float row_absmax(const uint4& hv)
{
    const half2* h = reinterpret_cast<const half2*>(&hv);
    float mx = 0.f;
    #pragma unroll
    for (int i = 0; i < 4; i++) { const float2 f = __half22float2(__habs2(h[i])); mx = fmaxf(mx, fmaxf(f.x, f.y)); }
    return mx;
}

// Grid-wide barrier for the consumer threads of all CTAs (sense-reversing counter in global memory).
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nctas, unsigned& gen, int tid)
{
    consumer_sync();
    if (tid == 0) {
        __threadfence();
        const unsigned want = gen + 1u;
        const unsigned prev = atomicAdd(bar, 1u);
        if (prev == nctas - 1u) {
            bar[0] = 0u;
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(bar + 1), "r"(want) : "memory");
        } else {
            unsigned v, spins = 0;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar + 1) : "memory");
                if (++spins > SPIN_LIMIT) __trap();
            } while (v != want);
        }
        __threadfence();
    }
    gen += 1u;
    consumer_sync();
}

struct PhaseGeom { int K, spt, sg0, sg1, units, subs, mat0, tiles_per_mat; };

__device__ __forceinline__ PhaseGeom phase_geom(const LayerArgs& a, int ph, int rank)
{
    PhaseGeom g;
    if (ph == PH_QKV)       { g.K = a.hidden; g.tiles_per_mat = a.hidden / TILE_N; g.units = 3 * g.tiles_per_mat; g.subs = 1; g.mat0 = 0; }
    else if (ph == PH_O)    { g.K = a.hidden; g.tiles_per_mat = a.hidden / TILE_N; g.units = g.tiles_per_mat; g.subs = 1; g.mat0 = 3; }
    else if (ph == PH_GU)   { g.K = a.hidden; g.tiles_per_mat = (a.inter + TILE_N - 1) / TILE_N; g.units = g.tiles_per_mat; g.subs = 2; g.mat0 = 4; }
    else                    { g.K = a.inter;  g.tiles_per_mat = a.hidden / TILE_N; g.units = g.tiles_per_mat; g.subs = 1; g.mat0 = 6; }
    g.spt = g.K / STAGE_K;
    g.sg0 = rank * g.spt / CS; g.sg1 = (rank + 1) * g.spt / CS;
    return g;
}
// (phase, unit, sub) -> matrix index, column tile, N of that matrix
__device__ __forceinline__ void resolve_unit(const LayerArgs& a, int ph, const PhaseGeom& g, int u, int sub, int& mi, int& ctile, int& N)
{
    if (ph == PH_QKV) { mi = u / g.tiles_per_mat; ctile = u - mi * g.tiles_per_mat; N = a.hidden; }
    else if (ph == PH_GU) { mi = 4 + sub; ctile = u; N = a.inter; }
    else { mi = g.mat0; ctile = u; N = a.hidden; }
}

// Per-thread inner-product state for one sub-tile (see q4_gemv.cu Accum): M = 1, group == ring stage.
struct Accum
{
    float acc[4];        // fp32 totals of this lane's 4 columns (token 0)
    int ia[8], ib[8];    // integer dot products of the current stage: plane a / plane b  [tile A: 0..3, tile B: 4..7]
};

__global__ void __launch_bounds__(THREADS, 2) decode_layer_kernel(const __grid_constant__ LayerArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* ring = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);        // NST x STAGE_STRIDE
    unsigned char* xs = ring + NST * STAGE_STRIDE;                                      // MAX_SLICE_STAGES x 16 rows x 16 B quantised x
    float* seg = reinterpret_cast<float*>(xs + MAX_SLICE_STAGES * STAGE_ROWS * 16);     // per stage {sum x_q (as int bits), x scale}
    float* red = seg + MAX_SLICE_STAGES * 2;                                            // WK x 128
    float* slots = red + WK * TILE_N;                                                   // CS x 132 (leader: partials of the peers)
    float* att = slots + CS * 132;                                                      // attention scratch: q[128], p[256], red[8], o[16][128]
    __shared__ __align__(8) unsigned long long full_bar[NST], empty_bar[NST], red_bar, free_bar;
    __shared__ float s_rm;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster_ctarank();
    const int cluster = (int)cluster_id_x();
    const int n_clusters = a.n_clusters;
    const unsigned nctas = gridDim.x;

    if (tid == 0) {
        for (int i = 0; i < NST; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], WN); }
        mbar_init(&red_bar, 1); mbar_init(&free_bar, 1);
        if (rank == 0) mbar_expect_tx(&red_bar, XCH_BYTES);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");

    if (warp == CONSUMERS / 32) {
        // ================= producer warp: the whole layer's weight / scale / zero stream, in schedule order =================
        int j = 0;
        #pragma unroll 1
        for (int ph = 0; ph < 5; ph++) {
            if (ph == PH_ATT) continue;
            const PhaseGeom g = phase_geom(a, ph, rank);
            #pragma unroll 1
            for (int u = cluster; u < g.units; u += n_clusters) {
                #pragma unroll 1
                for (int sub = 0; sub < g.subs; sub++) {
                    int mi, ctile, N; resolve_unit(a, ph, g, u, sub, mi, ctile, N);
                    const int col0 = ctile * TILE_N;
                    const int tile_cols = min(TILE_N, N - col0);
                    const int nbox = (tile_cols + BOX_COLS - 1) / BOX_COLS;
                    #pragma unroll 1
                    for (int sg = g.sg0; sg < g.sg1; sg++, j++) {
                        const int slot = j % NST;
                        const uint32_t fb = smem_u32(&full_bar[slot]);
                        if (j >= NST) mbar_wait(smem_u32(&empty_bar[slot]), (uint32_t)((j / NST) + 1) & 1u);
                        const uint32_t base = smem_u32(ring) + slot * STAGE_STRIDE;
                        if (lane == 0) mbar_expect_tx(&full_bar[slot], (uint32_t)nbox * BOX_BYTES + (uint32_t)tile_cols * 2 + (uint32_t)tile_cols / 2);
                        __syncwarp();
                        if (lane < nbox) tma_load_2d(base + lane * BOX_BYTES, &a.tm[mi], col0 + lane * BOX_COLS, sg * STAGE_ROWS, fb);
                        if (lane == 4) bulk_g2s(base + META_SC, a.sc[mi] + (size_t)sg * N + col0, (uint32_t)tile_cols * 2, fb);
                        if (lane == 5) bulk_g2s(base + META_ZQ, a.qz[mi] + (size_t)sg * (N >> 3) + (col0 >> 3), (uint32_t)tile_cols / 2, fb);
                    }
                }
            }
        }
        return;
    }

    // ======================================================= consumer warps =======================================================
    const int wn = warp & (WN - 1), wk = warp >> 2;
    const int g8 = lane >> 2, t = lane & 3;
    const int pg = (g8 >> 1) | ((g8 & 1) << 2);            // column chunk of this lane (bank-conflict-free with the 128B swizzle)
    const int lane_col = wn * 32 + 4 * pg;
    unsigned gen = 0;                                       // grid-barrier generation (the host resets grid_bar[1] to 0 before the launch)
    int jcons = 0;                                          // ring position of the next stage this CTA consumes (all consumer warps agree)
    int exch = 0;                                           // exchanges done so far (red_bar / free_bar parities)
    bool cluster_up = false;

    // ---- exchange of one 128-float partial (+ extra floats) per CTA with the leader; returns in v the rank-ordered total ----
    auto exchange = [&](float& v, bool have_v, bool last) {
        // v: this thread's value (threads tid < 128 own column tid); leader gets the sum over ranks
        if (!cluster_up) { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); cluster_up = true; }
        if (rank != 0) {
            if (exch > 0) { for (unsigned i = 0; !mbar_try_cluster(smem_u32(&free_bar), (uint32_t)(exch - 1) & 1u); i++) if (i > SPIN_LIMIT) __trap(); }
            if (tid < 130) st_async_f32(mapa_shared(smem_u32(slots + rank * 132 + tid), 0), have_v ? v : 0.f, mapa_shared(smem_u32(&red_bar), 0));
        } else {
            mbar_wait(smem_u32(&red_bar), (uint32_t)exch & 1u);
            if (have_v) {
                #pragma unroll 1
                for (int r = 1; r < CS; r++) v += slots[r * 132 + tid];
            }
            consumer_sync();                                    // every thread has read its slot values
            if (!last) {
                if (tid == 0) mbar_expect_tx(&red_bar, XCH_BYTES);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                consumer_sync();
                if (tid >= 1 && tid < CS) mbar_arrive_remote(mapa_shared(smem_u32(&free_bar), (uint32_t)tid));
            }
        }
        exch++;
    };

    // ---- stage this CTA's K slice of the phase's input vector: (optional rms norm) -> 16-bit quantisation per stage (= group) ----
    auto stage_x = [&](const half* xin, const half* norm_w, const float* ssq, int ntile_ssq, const PhaseGeom& g) {
        if (norm_w) {
            // row factor from the per-tile sums of squares the producer of x left behind (or from x itself on the first layer)
            if (warp == 0) {
                float s = 0.f;
                if (ssq) { for (int i = lane; i < ntile_ssq; i += 32) s += __ldcg(ssq + i); }
                else {
                    for (int i = lane; i < a.hidden / 8; i += 32) {
                        const uint4 xv = __ldcg(reinterpret_cast<const uint4*>(xin) + i);
                        const half2* h = reinterpret_cast<const half2*>(&xv);
                        #pragma unroll
                        for (int q = 0; q < 4; q++) { const float2 f = __half22float2(h[q]); s = fmaf(f.x, f.x, s); s = fmaf(f.y, f.y, s); }
                    }
                }
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) s_rm = __half2float(__float2half_rn(rsqrtf(s / (float)a.hidden + a.eps)));
            }
            consumer_sync();
        }
        const int nrows = (g.sg1 - g.sg0) * STAGE_ROWS;
        for (int base = 0; base < nrows; base += CONSUMERS) {
            const int rr = base + tid;
            const bool act = rr < nrows;
            uint4 hv = make_uint4(0, 0, 0, 0);
            if (act) {
                const int k8 = g.sg0 * STAGE_ROWS + rr;
                hv = __ldcg(reinterpret_cast<const uint4*>(xin + (size_t)k8 * 8));      // produced by other SMs earlier in this launch: bypass L1
                if (norm_w) {
                    const half2 rm2 = __float2half2_rn(s_rm);
                    const uint4 wv = *reinterpret_cast<const uint4*>(norm_w + (size_t)k8 * 8);
                    half2* h = reinterpret_cast<half2*>(&hv);
                    const half2* w2 = reinterpret_cast<const half2*>(&wv);
                    #pragma unroll
                    for (int i = 0; i < 4; i++) h[i] = __hmul2(__hmul2(h[i], rm2), w2[i]);      // rms_norm.cu:118-131
                }
            }
            float mx = row_absmax(hv);
            #pragma unroll
            for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));      // 16 rows = one stage = one group
            uint4 q;
            int sum = quantise_row(hv, mx > 0.f ? 32767.0f / mx : 0.f, q);
            #pragma unroll
            for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (act) {
                *reinterpret_cast<uint4*>(xs + (size_t)rr * 16) = q;
                if ((rr & 15) == 0) { seg[(rr >> 4) * 2] = __int_as_float(sum); seg[(rr >> 4) * 2 + 1] = mx * (1.0f / 32767.0f); }
            }
        }
        consumer_sync();
    };

    // ---- one sub-tile: this CTA's K slice of a 128-column tile; returns the CTA partial of column tid (tid < 128) ----
    auto gemv_subtile = [&](const PhaseGeom& g, int tile_cols) -> float {
        Accum A;
        #pragma unroll
        for (int i = 0; i < 4; i++) A.acc[i] = 0.f;
        const bool col_ok = (wn * 32) < tile_cols;
        const uint32_t xa0 = smem_u32(xs) + (uint32_t)t * 16u;
        const uint32_t M4 = 0x0f0f0f0fu;
        const int nst = g.sg1 - g.sg0;
        // stages alternate between the two k-warp groups
        #pragma unroll 1
        for (int s = ((wk - jcons) & 1); s < nst; s += 2) {
            const int j = jcons + s;
            const int slot = j % NST;
            mbar_wait(smem_u32(&full_bar[slot]), (uint32_t)(j / NST) & 1u);
            if (col_ok) {
                const uint32_t sb = smem_u32(ring) + slot * STAGE_STRIDE;
                #pragma unroll
                for (int i = 0; i < 8; i++) { A.ia[i] = 0; A.ib[i] = 0; }
                int (&aA)[4] = *reinterpret_cast<int (*)[4]>(&A.ia[0]);
                int (&aB)[4] = *reinterpret_cast<int (*)[4]>(&A.ia[4]);
                int (&bA)[4] = *reinterpret_cast<int (*)[4]>(&A.ib[0]);
                int (&bB)[4] = *reinterpret_cast<int (*)[4]>(&A.ib[4]);
                #pragma unroll
                for (int u = 0; u < 4; u++) {
                    // rows of this lane inside the stage: r = u*4 + t; the 16-byte chunk pg is stored at pg ^ (r & 7)
                    const int r = u * 4 + t;
                    const uint4 w = lds128(sb + wn * BOX_BYTES + r * 128 + ((pg ^ (r & 7)) << 4));
                    const uint4 xb = lds128(xa0 + (uint32_t)(s * STAGE_ROWS + u * 4) * 16u);
                    const uint32_t lo0 = w.x & M4, hi0 = (w.x >> 4) & M4, lo1 = w.y & M4, hi1 = (w.y >> 4) & M4;
                    const uint32_t lo2 = w.z & M4, hi2 = (w.z >> 4) & M4, lo3 = w.w & M4, hi3 = (w.w >> 4) & M4;
                    imma_u8s8(aA, lo0, lo1, hi0, hi1, xb.x, xb.y);
                    imma_u8u8(bA, lo0, lo1, hi0, hi1, xb.z, xb.w);
                    imma_u8s8(aB, lo2, lo3, hi2, hi3, xb.x, xb.y);
                    imma_u8u8(bB, lo2, lo3, hi2, hi3, xb.z, xb.w);
                }
                // flush the stage (= quantisation group): acc += s * sx * (256 ia + ib - zp * sum x_q); only token 0 (D columns 0) matters
                uint2 sc2; uint32_t zw;
                asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(sc2.x), "=r"(sc2.y) : "r"(sb + META_SC + lane_col * 2));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(zw) : "r"(sb + META_ZQ + (lane_col >> 3) * 4));
                const int sxq = __float_as_int(seg[s * 2]);
                const float sxs = seg[s * 2 + 1];
                const half2 s01 = *reinterpret_cast<const half2*>(&sc2.x), s23 = *reinterpret_cast<const half2*>(&sc2.y);
                const float cs4[4] = {__low2float(s01), __high2float(s01), __low2float(s23), __high2float(s23)};
                const uint32_t z4 = zw >> ((lane_col & 4) * 4);
                #pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int zp = (int)((z4 >> (4 * c)) & 0xfu) + 1;
                    // D fragment: c0 = (row g8, col 2t), c2 = (row g8 + 8, col 2t): column c of this lane lives in tile (c >> 1), row half (c & 1)
                    const int jj = (c >> 1) * 4 + (c & 1) * 2;
                    const int val = A.ia[jj] * 256 + A.ib[jj] - zp * sxq;
                    A.acc[c] = fmaf(cs4[c] * sxs, (float)val, A.acc[c]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&empty_bar[slot]));
        }
        jcons += nst;
        // reduce the two k-warp groups through shared memory; only lanes with t == 0 hold token 0
        consumer_sync();
        if (t == 0) {
            #pragma unroll
            for (int c = 0; c < 4; c++) red[wk * TILE_N + lane_col + c] = col_ok ? A.acc[c] : 0.f;
        }
        consumer_sync();
        return tid < TILE_N ? red[tid] + red[TILE_N + tid] : 0.f;
    };


    // =========================================================== phase QKV ===========================================================
    {
        const PhaseGeom g = phase_geom(a, PH_QKV, rank);
        stage_x(a.x, a.ln1, a.ssq_in, a.hidden / TILE_N, g);
        half* hs = reinterpret_cast<half*>(att);
        for (int u = cluster; u < g.units; u += n_clusters) {
            int mi, ctile, N; resolve_unit(a, PH_QKV, g, u, 0, mi, ctile, N);
            float v = gemv_subtile(g, TILE_N);
            exchange(v, tid < TILE_N, false);
            if (rank == 0) {
                // tile == one head: rope on q / k (rope.cu:48-67), k / v rows into the cache (q4_attn.cu:32-51)
                if (tid < TILE_N) hs[tid] = __float2half_rn(v);
                consumer_sync();
                if (tid < TILE_N) {
                    half val = hs[tid];
                    if (mi < 2) {
                        const half* sr = a.sin + (size_t)a.past_len * TILE_N;
                        const half* cr = a.cos + (size_t)a.past_len * TILE_N;
                        const half other = hs[tid ^ 64];
                        val = tid < 64 ? __hfma(val, cr[tid], __hmul(other, __hneg(sr[tid]))) : __hfma(val, cr[tid], __hmul(other, sr[tid]));
                    }
                    if (mi == 0) a.q_buf[ctile * TILE_N + tid] = val;
                    else (mi == 1 ? a.kc : a.vc)[((size_t)ctile * a.max_seq + a.past_len) * TILE_N + tid] = val;
                }
                consumer_sync();
            }
        }
    }
    grid_barrier(a.grid_bar, nctas, gen, tid);

    // =========================================================== phase ATT ===========================================================
    {
        // the CS CTAs of a cluster split the sequence of one head (csrc/decode_attn.cu, without the register prefetch: the weight
        // ring of the next phases is already being filled by the producer warp while this runs)
        float* s_q = att; float* s_p = att + 128; float* s_r = att + 384; float* s_o = att + 392;     // [16][128]
        const int seq = a.past_len + 1;
        const int chunk = (seq + CS - 1) / CS;
        const int p0 = rank * chunk, p1 = min(seq, p0 + chunk);
        const float scale = rsqrtf((float)TILE_N);
        const int l16 = lane & 15, sub = lane >> 4, r16 = tid & 15, pgp = tid >> 4;
        for (int h = cluster; h < a.heads; h += n_clusters) {
            const half* kb = a.kc + (size_t)h * a.max_seq * TILE_N;
            const half* vb = a.vc + (size_t)h * a.max_seq * TILE_N;
            if (tid < TILE_N) s_q[tid] = __half2float(__ldcg(a.q_buf + h * TILE_N + tid)) * scale;
            consumer_sync();
            float qf[8];
            #pragma unroll
            for (int j = 0; j < 8; j++) qf[j] = s_q[l16 * 8 + j];
            float acc[8];
            #pragma unroll
            for (int j = 0; j < 8; j++) acc[j] = 0.f;
            float mx = -INFINITY, lsum = 0.f;
            for (int b0 = p0; b0 < p1; b0 += CONSUMERS) {
                const int b1 = min(p1, b0 + CONSUMERS);
                #pragma unroll 4
                for (int it = 0; it < 16; it++) {
                    const int pl = warp * 32 + it * 2 + sub, p = b0 + pl;
                    float sdot = 0.f;
                    if (p < b1) {
                        const uint4 kv = *(reinterpret_cast<const uint4*>(kb + (size_t)p * TILE_N) + l16);
                        const half2* hh = reinterpret_cast<const half2*>(&kv);
                        #pragma unroll
                        for (int j = 0; j < 4; j++) { const float2 f = __half22float2(hh[j]); sdot = fmaf(f.x, qf[2 * j], sdot); sdot = fmaf(f.y, qf[2 * j + 1], sdot); }
                    }
                    sdot += __shfl_xor_sync(0xffffffffu, sdot, 8); sdot += __shfl_xor_sync(0xffffffffu, sdot, 4);
                    sdot += __shfl_xor_sync(0xffffffffu, sdot, 2); sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
                    if (l16 == 0) s_p[pl] = (p < b1) ? sdot : -INFINITY;
                }
                consumer_sync();
                const float s = s_p[tid];
                float m = s;
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                if (lane == 0) s_r[warp] = m;
                consumer_sync();
                float bm = s_r[0];
                #pragma unroll
                for (int i = 1; i < CONSUMERS / 32; i++) bm = fmaxf(bm, s_r[i]);
                const float mnew = fmaxf(mx, bm);
                const float alpha = __expf(mx - mnew);
                const float e = __expf(s - mnew);
                float l = e;
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
                consumer_sync();
                s_p[tid] = e;
                if (lane == 0) s_r[warp] = l;
                consumer_sync();
                float bl = 0.f;
                #pragma unroll
                for (int i = 0; i < CONSUMERS / 32; i++) bl += s_r[i];
                #pragma unroll
                for (int j = 0; j < 8; j++) acc[j] *= alpha;
                #pragma unroll 4
                for (int it = 0; it < 16; it++) {
                    const int p = b0 + pgp + it * 16;
                    if (p < b1) {
                        const float w = s_p[pgp + it * 16];
                        const uint4 vv = *(reinterpret_cast<const uint4*>(vb + (size_t)p * TILE_N) + r16);
                        const half2* hh = reinterpret_cast<const half2*>(&vv);
                        #pragma unroll
                        for (int j = 0; j < 4; j++) { const float2 f = __half22float2(hh[j]); acc[2 * j] = fmaf(w, f.x, acc[2 * j]); acc[2 * j + 1] = fmaf(w, f.y, acc[2 * j + 1]); }
                    }
                }
                consumer_sync();
                lsum = lsum * alpha + bl;
                mx = mnew;
            }
            #pragma unroll
            for (int j = 0; j < 8; j++) s_o[pgp * TILE_N + r16 * 8 + j] = acc[j];
            consumer_sync();
            float osum = 0.f;
            if (tid < TILE_N) {
                #pragma unroll
                for (int i = 0; i < 16; i++) osum += s_o[i * TILE_N + tid];
            }
            // exchange (o[128], m, l): values 128 / 129 ride in threads 128 / 129
            float mine = tid < TILE_N ? osum : (tid == TILE_N ? mx : lsum);
            // the leader needs every rank's (o, m, l) separately -> no summation inside exchange(): deposit, then combine here
            if (!cluster_up) { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); cluster_up = true; }
            if (rank != 0) {
                if (exch > 0) { for (unsigned i = 0; !mbar_try_cluster(smem_u32(&free_bar), (uint32_t)(exch - 1) & 1u); i++) if (i > SPIN_LIMIT) __trap(); }
                if (tid < 130) st_async_f32(mapa_shared(smem_u32(slots + rank * 132 + tid), 0), mine, mapa_shared(smem_u32(&red_bar), 0));
            } else {
                mbar_wait(smem_u32(&red_bar), (uint32_t)exch & 1u);
                if (tid < 130) slots[tid] = mine;                       // the leader's own partial in slot 0
                consumer_sync();
                if (tid < TILE_N) {
                    float M = -INFINITY;
                    for (int r = 0; r < CS; r++) M = fmaxf(M, slots[r * 132 + 128]);
                    float L = 0.f, o = 0.f;
                    for (int r = 0; r < CS; r++) {
                        const float w = __expf(slots[r * 132 + 128] - M);
                        L = fmaf(slots[r * 132 + 129], w, L);
                        o = fmaf(slots[r * 132 + tid], w, o);
                    }
                    a.attn_out[h * TILE_N + tid] = __float2half_rn(o / L);
                }
                consumer_sync();
                if (tid == 0) mbar_expect_tx(&red_bar, XCH_BYTES);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                consumer_sync();
                if (tid >= 1 && tid < CS) mbar_arrive_remote(mapa_shared(smem_u32(&free_bar), (uint32_t)tid));
            }
            exch++;
        }
    }
    grid_barrier(a.grid_bar, nctas, gen, tid);

    // =========================================================== phase O ===========================================================
    {
        const PhaseGeom g = phase_geom(a, PH_O, rank);
        stage_x(a.attn_out, nullptr, nullptr, 0, g);
        for (int u = cluster; u < g.units; u += n_clusters) {
            float v = gemv_subtile(g, TILE_N);
            exchange(v, tid < TILE_N, false);
            if (rank == 0) {
                float sq = 0.f;
                if (tid < TILE_N) {
                    half* o = a.x + u * TILE_N + tid;
                    const half nv = __float2half_rn(v + __half2float(__ldcg(o)));
                    *o = nv;
                    sq = __half2float(nv) * __half2float(nv);
                }
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                if (lane == 0 && warp < 4) red[warp] = sq;
                consumer_sync();
                if (tid == 0) a.ssq_mid[u] = red[0] + red[1] + red[2] + red[3];
                consumer_sync();
            }
        }
    }
    grid_barrier(a.grid_bar, nctas, gen, tid);

    // =========================================================== phase GU ===========================================================
    {
        const PhaseGeom g = phase_geom(a, PH_GU, rank);
        stage_x(a.x, a.ln2, a.ssq_mid, a.hidden / TILE_N, g);
        for (int u = cluster; u < g.units; u += n_clusters) {
            const int tile_cols = min(TILE_N, a.inter - u * TILE_N);
            float vg = gemv_subtile(g, tile_cols);
            exchange(vg, tid < TILE_N, false);
            float vu = gemv_subtile(g, tile_cols);
            exchange(vu, tid < TILE_N, false);
            if (rank == 0 && tid < tile_cols)
                a.act[u * TILE_N + tid] = __hmul(silu_h(__float2half_rn(vg)), __float2half_rn(vu));      // q4_mlp.cu:27-36,46-88
        }
    }
    grid_barrier(a.grid_bar, nctas, gen, tid);

    // =========================================================== phase DOWN ===========================================================
    {
        const PhaseGeom g = phase_geom(a, PH_DOWN, rank);
        stage_x(a.act, nullptr, nullptr, 0, g);
        for (int u = cluster; u < g.units; u += n_clusters) {
            float v = gemv_subtile(g, TILE_N);
            const bool last = (u + n_clusters) >= g.units;
            exchange(v, tid < TILE_N, last);
            if (rank == 0) {
                float sq = 0.f;
                if (tid < TILE_N) {
                    half* o = a.x + u * TILE_N + tid;
                    const half nv = __float2half_rn(v + __half2float(__ldcg(o)));
                    *o = nv;
                    sq = __half2float(nv) * __half2float(nv);
                }
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                if (lane == 0 && warp < 4) red[warp] = sq;
                consumer_sync();
                if (tid == 0) a.ssq_out[u] = red[0] + red[1] + red[2] + red[3];
                consumer_sync();
            }
        }
    }
}

} // namespace

// Host side of the draft: geometry checks + launch.  `mats` = q, k, v, o, gate, up, down handles of one layer.
// scratch must hold: q_buf[hidden] + attn_out[hidden] + act[inter] halves, 3 x (hidden/128) floats, 2 unsigned (zeroed once).
int exl_decode_layer_draft(ExlDevice* ds, const exl_q4_matrix* const* mats, half* x, const half* ln1, const half* ln2, float eps,
                           const half* sin, const half* cos, int past_len, int max_seq, int heads, half* kc, half* vc,
                           unsigned char* scratch, bool first_layer, cudaStream_t stream)
{
    LayerArgs a;
    memset(&a, 0, sizeof(a));
    const int hidden = mats[0]->K, inter = mats[4]->N;
    for (int i = 0; i < 7; i++) {
        const exl_q4_matrix* w = mats[i];
        if (w->x_map || w->groupsize != 128) return exl_set_err(EXL_ERR_ARG, "decode_layer: groupsize 128 without act-order only");
        a.tm[i] = w->tmap_w; a.qz[i] = w->qzeros; a.sc[i] = w->scales;
    }
    if (hidden % (CS * STAGE_K) != 0 || heads * TILE_N != hidden || inter % 32 != 0 || inter % STAGE_K != 0)
        return exl_set_err(EXL_ERR_ARG, "decode_layer: unsupported geometry hidden %d inter %d heads %d", hidden, inter, heads);
    if ((inter / STAGE_K + CS - 1) / CS > MAX_SLICE_STAGES) return exl_set_err(EXL_ERR_ARG, "decode_layer: K slice too long");
    a.hidden = hidden; a.inter = inter; a.heads = heads;
    a.x = x; a.ln1 = ln1; a.ln2 = ln2; a.eps = eps; a.sin = sin; a.cos = cos; a.past_len = past_len; a.max_seq = max_seq;
    a.kc = kc; a.vc = vc;
    half* hp = reinterpret_cast<half*>(scratch);
    a.q_buf = hp; a.attn_out = hp + hidden; a.act = hp + 2 * hidden;
    float* fp = reinterpret_cast<float*>(hp + 2 * hidden + ((inter + 7) & ~7));
    const int nt = hidden / TILE_N;
    // ping-pong of the per-tile sums between consecutive layers is the caller's business; the draft keeps three arrays
    a.ssq_in = first_layer ? nullptr : fp; a.ssq_mid = fp + nt; a.ssq_out = fp;       // out overwrites in: read only in phase QKV, written in phase DOWN
    a.grid_bar = reinterpret_cast<unsigned*>(fp + 2 * nt);

    const size_t smem = 1024 + (size_t)NST * STAGE_STRIDE + MAX_SLICE_STAGES * STAGE_ROWS * 16 + MAX_SLICE_STAGES * 8 +
                        (size_t)WK * TILE_N * 4 + (size_t)CS * 132 * 4 + (392 + 16 * 128) * 4;
    static bool attr_done = false;
    if (!attr_done) {
        EXL_CUDA_TRY(cudaFuncSetAttribute(decode_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        EXL_CUDA_TRY(cudaFuncSetAttribute(decode_layer_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 0));
        attr_done = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;          // all CTAs co-resident: the grid barrier relies on it
    cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream; cfg.attrs = at; cfg.numAttrs = 2;
    int n_clusters = 0;
    cfg.gridDim = dim3(CS * 64);
    EXL_CUDA_TRY(cudaOccupancyMaxActiveClusters(&n_clusters, decode_layer_kernel, &cfg));
    if (n_clusters < 1) return exl_set_err(EXL_ERR_CUDA, "decode_layer: no co-resident cluster fits");
    a.n_clusters = n_clusters;
    cfg.gridDim = dim3((unsigned)(n_clusters * CS));
    EXL_CUDA_TRY(cudaMemsetAsync(a.grid_bar, 0, 2 * sizeof(unsigned), stream));
    cudaError_t e = cudaLaunchKernelEx(&cfg, decode_layer_kernel, a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return exl_set_err(EXL_ERR_CUDA, "launch of decode_layer_kernel failed: %s", cudaGetErrorString(e));
    return EXL_OK;
}
// syntax / resource check (no GPU needed):
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xptxas -v -c exllama_b200/csrc/experimental/decode_layer.cu -o /tmp/decode_layer.o
