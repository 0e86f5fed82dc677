// decode_sched_check.cpp -- host-side proof obligations of the persistent decode kernel's static schedule.
//
// Compiled by tests/test_decode_schedule.py with g++ against exllama_b200/csrc/decode_step_sched.h -- the SAME header the kernel
// (decode_step.cu) includes, so the functions checked here are the ones that run on the GPU.  Test infrastructure, not product.
//
// For one (local) shape, grid size and a range of context lengths, checks:
//   P1  every phase's units are partitioned: shares are contiguous, disjoint, cover [0, U), CTAs >= Phase::G are empty,
//       and participating shares differ by at most one unit (every SM streams the same bytes)
//   P2  cta_of() is the inverse of range_lo() (the attention combine relies on it to find the CTAs of a head)
//   P3  ATT: a CTA touches at most 2 heads (the kernel keeps two head segments in shared memory)
//   P4  ATT: the partial slot a CTA writes for a head, cta - cta_of(first unit of the head), is inside [0, att_slots), no two
//       CTAs write the same (head, slot), and the O prologue's combine reads exactly the slots that were written -- except for
//       slots of idle CTAs at short contexts, which exist only where att_needs_reset() makes the host reset the table first
//   P5  act-order: a CTA's range touches at most 2 matrices of a multi-matrix phase (two staging slots)
//   P6  HEAD: the 8 KB stages cover every 1 KB chunk of lm_head once (hidden % 512 == 0)
//   P7  shared memory: >= 2 ring stages per pipeline fit, the head's normalised x fits the staging area
#include "../exllama_b200/csrc/decode_step_sched.h"
#include <cstdio>
#include <cstring>
#include <vector>

using namespace ds_sched;

namespace {

struct Dims
{
    int H, HQ, I, heads, past_len, vocab;
    const void* lm_head;
    float *acc_qkv, *acc_o, *acc_gu, *acc_d;
};

char* g_msg; int g_len;
int fail(const char* fmt, long long a = 0, long long b = 0, long long c = 0, long long d = 0, long long e = 0)
{
    if (g_msg && g_len > 0) snprintf(g_msg, (size_t)g_len, fmt, a, b, c, d, e);
    return 1;
}

int check_partition(const Phase& p, int grid, const char* name)
{
    static char buf[160];
    if (p.G < 1 || p.G > grid) { snprintf(buf, sizeof buf, "%s: taking-part CTAs %%lld outside [1, %%lld]", name); return fail(buf, p.G, grid); }
    int mn = 1 << 30, mx = 0, prev_hi = 0;
    for (int c = 0; c < grid; c++) {
        const int lo = range_lo(p.U, c, p.G), hi = range_lo(p.U, c + 1, p.G);
        if (lo != prev_hi) { snprintf(buf, sizeof buf, "%s: CTA %%lld starts at %%lld, previous ended at %%lld", name); return fail(buf, c, lo, prev_hi); }
        if (hi < lo) { snprintf(buf, sizeof buf, "%s: CTA %%lld has a negative share [%%lld, %%lld)", name); return fail(buf, c, lo, hi); }
        if (c >= p.G && hi != lo) { snprintf(buf, sizeof buf, "%s: CTA %%lld >= G = %%lld has work", name); return fail(buf, c, p.G); }
        if (c < p.G) { mn = hi - lo < mn ? hi - lo : mn; mx = hi - lo > mx ? hi - lo : mx; }
        for (int u = lo; u < hi; u++)
            if (cta_of(u, p.U, p.G) != c) { snprintf(buf, sizeof buf, "%s: cta_of(%%lld) = %%lld, owner is %%lld", name); return fail(buf, u, cta_of(u, p.U, p.G), c); }
        prev_hi = hi;
    }
    if (prev_hi != p.U) { snprintf(buf, sizeof buf, "%s: shares end at %%lld of %%lld units", name); return fail(buf, prev_hi, p.U); }
    if (p.U > 0 && mx - mn > 1) { snprintf(buf, sizeof buf, "%s: unbalanced shares %%lld .. %%lld", name); return fail(buf, mn, mx); }
    return 0;
}

} // namespace

// returns 0 when every property holds for every past_len in [past_lo, past_hi]; else 1 with a message
extern "C" int ds_check(int H, int HQ, int I, int heads, int vocab, int grid, int past_lo, int past_hi, int act,
                        long long dev_smem, char* msg, int msg_len, long long* n_gaps, int* n_resets)
{
    long long gaps = 0; int resets = 0;
    g_msg = msg; g_len = msg_len;
    if (msg && msg_len > 0) msg[0] = 0;
    Dims a; memset(&a, 0, sizeof a);
    a.H = H; a.HQ = HQ; a.I = I; a.heads = heads; a.vocab = vocab; a.lm_head = vocab ? (const void*)&a : nullptr;

    // ---- P7 ----
    const size_t fixed = smem_fixed_bytes(H, HQ, I, act != 0);
    const int depth = ring_depth_for((size_t)dev_smem, fixed);
    if (depth < 2) return fail("P7: only %lld ring stages per pipeline fit (fixed part %lld B of %lld B)", depth, (long long)fixed, dev_smem);
    if (smem_dynamic_bytes(fixed, depth) + SMEM_STATIC_ALLOWANCE > (size_t)dev_smem) return fail("P7: plan of %lld B exceeds %lld B", (long long)(smem_dynamic_bytes(fixed, depth) + SMEM_STATIC_ALLOWANCE), dev_smem);
    if ((size_t)H * 2 > (size_t)spt_max_for(H, HQ, I) * 256) return fail("P7: normalised x of the head (%lld B) does not fit the staging area", (long long)H * 2);
    if (H % 512 || HQ % TILE || I % TILE || HQ != heads * TILE) return fail("shape: widths must be multiples of 128 (hidden of 512), HQ == heads * 128");
    if (heads > grid) return fail("shape: more heads (%lld) than CTAs (%lld)", heads, grid);

    // ---- P1, P2, P5 for the phases that do not depend on the context length ----
    const int kinds[4] = {PH_QKV, PH_O, PH_GU, PH_DOWN};
    const char* names[4] = {"QKV", "O", "GU", "DOWN"};
    for (int k = 0; k < 4; k++) {
        const Phase p = phase_of(a, kinds[k], grid);
        if (p.U != p.nmat * (p.N / TILE) * p.spt || p.U <= 0) return fail("phase %lld: unit count %lld", kinds[k], p.U);
        if (check_partition(p, grid, names[k])) return 1;
        if (p.spt > spt_max_for(H, HQ, I)) return fail("phase %lld: %lld K stages exceed the staging area", kinds[k], p.spt);
        if (act) {
            const int per_mat = p.tpm * p.spt;
            for (int c = 0; c < grid; c++) {
                const int lo = range_lo(p.U, c, p.G), hi = range_lo(p.U, c + 1, p.G);
                if (hi > lo && (hi - 1) / per_mat - lo / per_mat > 1) return fail("P5: phase %lld CTA %lld touches %lld matrices", kinds[k], c, (hi - 1) / per_mat - lo / per_mat + 1);
                // the second staging slot holds H / 128 stages: only phases whose K is H may need it
                if (hi > lo && (hi - 1) / per_mat != lo / per_mat && p.spt != H / TILE) return fail("P5: phase %lld needs a second slot but K != hidden", kinds[k]);
            }
        }
    }
    // ---- P6 ----
    {
        const Phase p = phase_of(a, PH_HEAD, grid);
        if (vocab) {
            if (check_partition(p, grid, "HEAD")) return 1;
            const long long bytes = (long long)vocab * H * 2;
            if ((long long)p.U * W_BYTES < bytes || (long long)(p.U - 1) * W_BYTES >= bytes) return fail("P6: %lld stages for %lld bytes of lm_head", p.U, bytes);
            if (bytes % 1024) return fail("P6: lm_head is not a whole number of 1 KB chunks");
        } else if (p.U != 0) return fail("P6: no head but %lld units", p.U);
    }
    // ---- P1..P4 for ATT at every context length ----
    const int slots = att_slots_for(grid, heads);
    std::vector<int> written((size_t)heads * slots);
    for (int past = past_lo; past <= past_hi; past++) {
        a.past_len = past;
        const Phase p = phase_of(a, PH_ATT, grid);
        const int nph = p.tpm;
        if (nph != ((past + 15) / 16 > 0 ? (past + 15) / 16 : 1) || p.U != heads * nph) return fail("ATT past %lld: units %lld", past, p.U);
        if (check_partition(p, grid, "ATT")) { if (msg) { size_t n = strlen(msg); snprintf(msg + n, (size_t)msg_len - n, " (past %d)", past); } return 1; }
        std::fill(written.begin(), written.end(), -1);
        for (int c = 0; c < grid; c++) {
            const int u0 = range_lo(p.U, c, p.G), u1 = range_lo(p.U, c + 1, p.G);
            if (u1 <= u0) continue;
            const int h0 = u0 / nph, h1 = (u1 - 1) / nph;
            if (h1 - h0 + 1 > 2) return fail("P3: past %lld CTA %lld touches %lld heads", past, c, h1 - h0 + 1);
            for (int h = h0; h <= h1; h++) {
                const int slot = c - cta_of((long long)h * nph, p.U, p.G);
                if (slot < 0 || slot >= slots) return fail("P4: past %lld CTA %lld head %lld: slot %lld outside the table", past, c, h, slot);
                if (written[(size_t)h * slots + slot] >= 0) return fail("P4: past %lld head %lld slot %lld written twice (second writer CTA %lld)", past, h, slot, c);
                written[(size_t)h * slots + slot] = c;
            }
        }
        for (int h = 0; h < heads; h++) {
            // what combine8 (O prologue) reads
            const int c_lo = cta_of((long long)h * nph, p.U, p.G), c_hi = cta_of((long long)(h + 1) * nph - 1, p.U, p.G);
            const int ns = c_hi - c_lo + 1;
            if (ns < 1 || ns > slots) return fail("P4: past %lld head %lld: %lld partials, table has %lld slots", past, h, ns, slots);
            for (int s = 0; s < slots; s++) {
                const bool w = written[(size_t)h * slots + s] >= 0;
                if (w && s >= ns) return fail("P4: past %lld head %lld slot %lld is written but the combine reads only the first %lld", past, h, s, ns);
                if (!w && s < ns) {
                    // a CTA between the head's first and last CTA owns no unit: the combine reads a slot nobody wrote in THIS launch.
                    // Allowed only where the host resets the table before the launch (att_needs_reset).
                    gaps++;
                    if (!att_needs_reset(heads, past, grid)) return fail("P4: past %lld head %lld slot %lld is read but not written, and the host does not reset the table", past, h, s);
                }
            }
        }
        if (att_needs_reset(heads, past, grid)) resets++;
        if (att_needs_reset(heads, past, grid) && p.U >= p.G) return fail("P4: past %lld: reset requested although every CTA has work", past);
    }
    if (n_gaps) *n_gaps = gaps;
    if (n_resets) *n_resets = resets;
    return 0;
}

extern "C" int ds_plan(int H, int HQ, int I, int heads, int grid, int act, long long dev_smem, int* depth, long long* smem, int* att_slots)
{
    const size_t fixed = smem_fixed_bytes(H, HQ, I, act != 0);
    *depth = ring_depth_for((size_t)dev_smem, fixed);
    *smem = (long long)smem_dynamic_bytes(fixed, *depth);
    *att_slots = att_slots_for(grid, heads);
    return 0;
}
