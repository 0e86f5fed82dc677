// decode_step_sched.h -- the static work schedule of decode_step_kernel (decode_step.cu): which units of a phase a CTA owns,
// which CTA owns a unit, how many CTAs take part, and the host-side sizing that depends on it.
//
// Plain C++ with no CUDA types: the kernel includes it as device code, and tests/test_decode_schedule.py compiles the same
// header with g++ (tests/decode_sched_check.cpp) to check the partition invariants for every BASELINE shape, tensor-parallel
// shard and context length on the CPU -- a schedule mistake on the GPU is a trapped launch, so it is worth proving first.
#pragma once
#include <cstddef>

#if defined(__CUDACC__)
#define DS_HD __host__ __device__ __forceinline__
#else
#define DS_HD static inline
#endif

namespace ds_sched {

constexpr int TILE = 128;                        // columns per tile == K rows per stage
constexpr int W_BYTES = 8192;                    // packed weights of one unit: 16 k8-rows x 128 columns x 4 B
constexpr int STAGE_STRIDE = W_BYTES + 2048;     // ring stage: weights + scale rows + zero rows, 1 KB aligned (swizzle atom)
constexpr int PART_LD = 132;                     // attention partial: o[128], m, l, pad
constexpr int MAX_DEPTH = 6;                     // ring stages per pipeline
constexpr int ATT_CTAS_PER_HEAD = 7;             // at most this many CTAs share one head in the attention phase (when heads are few)

enum { PH_QKV = 0, PH_ATT = 1, PH_O = 2, PH_GU = 3, PH_DOWN = 4, PH_HEAD = 5 };

// contiguous share of n items for CTA c of G
DS_HD int share_lo(long long n, int c, int G) { return (int)(n * c / G); }
// share of CTA c in a phase that p_G <= G CTAs take part in: [lo, hi), empty for c >= p_G
DS_HD int range_lo(long long n, int c, int pG) { return c >= pG ? (int)n : (int)(n * c / pG); }
// the CTA whose share contains item u (inverse of share_lo)
DS_HD int cta_of(long long u, long long U, int G) { return (int)(((u + 1) * G - 1) / U); }

// U: units; spt: K stages per tile; tpm: tiles per matrix (ATT: units per head); G: CTAs that take part in the phase
struct Phase { int kind, U, spt, tpm, nmat, mat0, N, G; float* acc; };

// A: anything with the fields H, HQ, I, heads, past_len, vocab, lm_head, acc_qkv, acc_o, acc_gu, acc_d (the kernel's StepArgs)
template <class A>
DS_HD Phase phase_of(const A& a, int kind, int grid)
{
    Phase p; p.kind = kind; p.acc = nullptr; p.G = grid; p.spt = 0; p.tpm = 0; p.nmat = 0; p.mat0 = 0; p.N = 0; p.U = 0;
    if (kind == PH_QKV)       { p.spt = a.H / TILE;  p.N = a.HQ; p.nmat = 3; p.mat0 = 0; p.acc = a.acc_qkv; }
    else if (kind == PH_O)    { p.spt = a.HQ / TILE; p.N = a.H;  p.nmat = 1; p.mat0 = 3; p.acc = a.acc_o; }
    else if (kind == PH_GU)   { p.spt = a.H / TILE;  p.N = a.I;  p.nmat = 2; p.mat0 = 4; p.acc = a.acc_gu; }
    else if (kind == PH_DOWN) { p.spt = a.I / TILE;  p.N = a.H;  p.nmat = 1; p.mat0 = 6; p.acc = a.acc_d; }
    if (kind == PH_ATT) {
        const int nch = (a.past_len + 15) >> 4;
        p.tpm = nch > 0 ? nch : 1;                          // units per head (one empty unit when there is no history)
        p.U = a.heads * p.tpm;
        // at most ~7 CTAs per head: the O prologue combines one partial per CTA and head (two rounds of 4 loads); with few heads
        // per GPU (tensor parallel) the rest of the grid sits this phase out -- its bytes are small then
        if (a.heads * ATT_CTAS_PER_HEAD < grid) p.G = a.heads * ATT_CTAS_PER_HEAD;
    } else if (kind == PH_HEAD) {
        p.U = a.lm_head ? (int)(((long long)a.vocab * a.H * 2 + W_BYTES - 1) / W_BYTES) : 0;
    } else {
        p.tpm = p.N / TILE;
        p.U = p.nmat * p.tpm * p.spt;
    }
    return p;
}

// ---- host side (exl_decode_plan_create / exl_decode_step) ----
// Short contexts: with fewer attention units than participating CTAs (U < G) some CTAs between the first and the last CTA of a
// head own nothing, so the slot they would have written keeps whatever an EARLIER launch left there (a longer context fills
// every slot) -- and the O prologue combines slots [0, c_hi - c_lo] of the head.  exl_decode_step resets the table to the
// neutral partial (m = -inf, l = 0, o = 0) before such a launch.  With one unit per head (tpm == 1) a head has one slot: no gap.
// tests/decode_sched_check.cpp proves: a gap exists only where this predicate holds.
struct AttDims { int H, HQ, I, heads, past_len, vocab; const void* lm_head; float *acc_qkv, *acc_o, *acc_gu, *acc_d; };
static inline bool att_needs_reset(int heads, int past_len, int grid)
{
    AttDims a = {};
    a.heads = heads; a.past_len = past_len;
    const Phase p = phase_of(a, PH_ATT, grid);
    return p.U < p.G && p.tpm >= 2;
}

// slots of att_part per head: one per CTA that can hold a piece of the head (+ 2: a range may start and end inside the head)
static inline int att_slots_for(int grid, int heads)
{
    return (grid < ATT_CTAS_PER_HEAD * heads ? grid / heads : ATT_CTAS_PER_HEAD) + 2;
}
static inline int spt_max_for(int H, int HQ, int I)
{
    int m = (H > I ? H : I) / TILE;
    return HQ / TILE > m ? HQ / TILE : m;
}
constexpr size_t SMEM_STATIC_ALLOWANCE = 2048;   // static __shared__ of the kernel (mbarriers, schedule tables): ptxas reports 1664 B
// shared memory apart from the ring: alignment slack, x staging (+ a second slot of H / 128 stages with act-order), residual
// stream, attention scratch (partials, q / new k / new v in fp32, q in fp16), norm weights, static variables
static inline size_t smem_fixed_bytes(int H, int HQ, int I, bool act)
{
    return 1024 + (size_t)(spt_max_for(H, HQ, I) + (act ? H / TILE : 0)) * (256 + 32) + (size_t)H * 2
         + (size_t)(2 * 17 * PART_LD + 6 * TILE) * 4 + (size_t)2 * TILE * 2 /* fp16 q */ + (size_t)H * 2 /* wnorm */ + SMEM_STATIC_ALLOWANCE;
}
// dynamic shared memory to request for `depth` ring stages per pipeline
static inline size_t smem_dynamic_bytes(size_t fixed, int depth) { return fixed - SMEM_STATIC_ALLOWANCE + (size_t)4 * depth * STAGE_STRIDE; }
// ring stages per pipeline that fit (0: the model is too wide)
static inline int ring_depth_for(size_t dev_smem, size_t fixed)
{
    if (dev_smem <= fixed) return 0;
    int depth = (int)((dev_smem - fixed) / STAGE_STRIDE) / 4;
    return depth > MAX_DEPTH ? MAX_DEPTH : depth;
}

} // namespace ds_sched
