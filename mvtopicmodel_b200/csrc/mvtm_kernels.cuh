// mvtm_kernels.cuh -- sm_100a kernels of the collapsed-Gibbs engine (no host code here).
//
// Reference tags as in include/mvtm.h (W = FastQMVWVWorkerRunnable, U = FastQMVWVUpdaterRunnable,
// M = FastQMVWVParallelTopicModel, FT = FTree).
//
// Design (DESIGN.md has the long form):
//  * one lane group per document-view (north_star a): G = 32 lanes (a whole warp) for large K, G = 16 or 8 lanes
//    for small K so that a warp hosts 2 or 4 document-views of near-equal length and the per-token control
//    instructions (shuffles, searches, TMA issue, bookkeeping) are shared between them; persistent CTAs, one per
//    SM, pull groups of work items from a longest-first list with an atomic counter;
//  * the document's topic counts n_d (u16) and its per-topic factor q[t] live in shared memory (b);
//    q[t] = (p_mm*n_d[t] + [t in S]*O_m[t] + gamma*alpha[t]) / (n_k[t] + betaSum), so that the weight of
//    topic t for a token of word w is simply (n_wk[w][t] + beta) * q[t]  -- the net distribution of
//    W:495-538 (doc bucket + tree bucket) evaluated densely;
//  * n_wk rows are staged global -> shared by the TMA engine (cp.async.bulk 1-D, mbarrier complete_tx)
//    through a per-warp ring that keeps `R` rows in flight, and consumed as 128-bit LDS (b);
//  * the conditional is sampled by a group-cooperative scan: per-lane partial sums, segmented __shfl_up
//    inclusive scan over the group's lanes, __ballot to find the lane, register-resident chunk sums to find the
//    chunk, a 4-element search to find the topic (c);
//  * RNG is Philox4x32-10 keyed on (seed; token position, global doc id, iteration, view|purpose) (d);
//  * count deltas: n_wk by global RED atomics issued by two lanes, n_k through a per-CTA shared-memory
//    delta vector flushed once per kernel (e).  No tensor cores: nothing here is a contraction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MVTM_MAXM 8
// registers per thread the DIRECT sweep kernels are compiled for (__maxnreg__).  The register file is split per SM sub-partition
// (16 K registers each): 168 registers = 3 warps per sub-partition (12 per SM), 128 = 4 (16 per SM); nothing in between launches.
// The row costs KS/G registers: 64 at (1024, 16) and (2048, 32) -> 168; 32 at (512, 16) -> 128.
__host__ __device__ constexpr int direct_regs(int KS, int G) { return (KS / G <= 32) ? 128 : 168; }

struct SweepParams {
    int M, K, Kp, m, V;
    int n_items;
    const int *order;                 // work list: local doc ids, longest first
    int *work_counter;
    const long long *doc_off[MVTM_MAXM];
    const int *word;                  // view m
    int *zv[MVTM_MAXM];               // assignments of every view (view m is written)
    int *nwk;                         // V x Kp
    const int *nk_frozen;             // Kp, snapshot taken before the launch
    int *nk_live;                     // Kp, receives the flushed deltas
    const float *ga_tree;             // gamma_m*alpha_m[t], 0 for inactive topics (M:2670-2678)
    const float *ga_full[MVTM_MAXM];  // gamma_i*alpha_i[t] unmasked (W:404)
    float beta, betaSum;
    float gas[MVTM_MAXM];             // gamma_i*alphaSum_i
    float ga_new[MVTM_MAXM];          // gamma_i*alpha_i[K]
    double pa[MVTM_MAXM], pb[MVTM_MAXM];   // p_a[m][i], p_b[m][i]
    unsigned char sparse[MVTM_MAXM];  // beta[i] == 0.0001 (W:335-336)
    int n_inactive, first_inactive;
    unsigned seed_lo, seed_hi, iteration;
    long long doc_id_base, doc_id_stride;
    int update_global;
    int beta_mallet;                  // MVTM_FLAG_BETA_MALLET: the view-coupling draw follows MALLET's Randoms.nextBeta (quirk Q5)
    int R;                            // ring depth
    unsigned long long *stats;        // [0] tokens, [1] changed, [2] new-topic draws
    float *oc_scratch;                // (multi-view) per resident document slot: Kp floats, see DocCtx::oc
    int oc_smem;                      // (multi-view) 1: keep DocCtx::oc in shared memory instead (views of short documents)
    unsigned *rbits;                  // Q1 mode: per document KS/32 words, bit t = "topic t is NOT in the document's dense index for the
                                      //   rest of this sweep" (it left it, or was gained while absent); carried across the view passes
    unsigned *rb_one;                 // probe: the flags of the probed document (overrides rbits)
    int *z_host;                      // mvtm_sweep_host: device alias of the caller's pinned array of view m (NULL: none); every
                                      //   token block's new assignments are stored there too, so no copy follows the pass
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mulhi_u32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
__host__ __device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t h0 = mulhi_u32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = mulhi_u32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
enum { PURPOSE_SAMPLE = 0, PURPOSE_INIT = 1, PURPOSE_PDRAW = 2 };

// ------------------------------------------------------------------------------------------------
// TMA 1-D bulk copy + mbarrier helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory"); }
__device__ __forceinline__ void fence_mbar_init()
{ asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async()
{ asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_row_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    } while (!ok);
}

// explicit .shared accesses through a 32-bit address held in a register (keeps ptxas from rematerialising the
// dynamic-smem base with S2UR SR_CgaCtaId on every access of the CTA-level vectors)
__device__ __forceinline__ float2 lds_f2(uint32_t a)
{ float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void reds_add(uint32_t a, int v)
{ asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// per-document state through ONE 32-bit base held in a register (DocCtx::sa) + compile-time offsets: the hot loop otherwise
// re-derives every address from %tid and the ring depth on each token (ptxas rematerialises them: ~45 instructions per step)
__device__ __forceinline__ int4 lds_i4(uint32_t a)
{ int4 v; asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ float4 lds_f4(uint32_t a)
{ float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ unsigned lds_u16(uint32_t a)
{ unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_u16(uint32_t a, unsigned v)
{ asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t a)
{ float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ unsigned lds_u32(uint32_t a)
{ unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_f32(uint32_t a, float v)
{ asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

#define FULL 0xffffffffu
__host__ __device__ constexpr int ilog2(int x) { return x <= 1 ? 0 : 1 + ilog2(x >> 1); }

// ------------------------------------------------------------------------------------------------
// per-document context: pointers into the lane group's shared-memory area + per-document scalars.
// Layout of the scan: topic t lives in chunk c = t >> 2 (4 topics, one 128-bit word); chunk c is owned by group
// lane gl = c % G and is that lane's j = c / G -th chunk (JG = KS / (4 G) chunks per lane).
// ------------------------------------------------------------------------------------------------
struct DocCtx {
    unsigned *rb;          // (Q1) the document's flag words in global memory, NULL if none are kept (single-view sweeps)
    uint32_t sa;           // .shared address of the document's area: q at +0, n_d at +4*KS (see carve_doc)
    float *q;              // KS
    unsigned short *nd;    // KS
    float *oc;             // Kp, GLOBAL scratch (MULTI) sum_i c_i * n_d[i][t]; only entries whose `om` bit is set are
                           //     ever written or read, so it is neither zeroed nor resident in shared memory
    unsigned *om;          // KS/32 (MULTI) bit t: some other view holds topic t
    float *cpar;           // 8    (MULTI) c_i = p[m][i] / (len_i + gas_i), 0 if len_i == 0 or i == m
    uint32_t ginv_sa;      // CTA-shared {ga_tree, 1/(n_k + betaSum)} float2[KS] as a .shared address
    const float *gaf;      // (MULTI) CTA-shared gamma_i*alpha_i[t] of every view, [M][KS] (W:404)
    int gaf_stride;        // = KS
    float pmm, coefm, C;   // p[m][m]; len_m + gas_m; new-topic mass per token (W:515)
};

template <bool MULTI>
__device__ __forceinline__ float q_value(float ndv, bool inS, float ocv, float pri, float coefm, float pmm, float2 gi)
{
    float v = ndv * pmm;
    if (MULTI) v += inS ? coefm * (ocv + pri) : 0.f;
    return (v + gi.x) * gi.y;
}

template <bool MULTI>
__device__ __forceinline__ float prior_other(const SweepParams &P, const DocCtx &c, int t)
{   // sum_i c_i * gamma_i*alpha_i[t] (W:404); c_i = 0 for the own view and for absent views, so no branch is needed
    float pri = 0.f;
    if (MULTI) {
        // two views: the sum has one term (c of the own view is 0).  Bit-identical to the loop, whose other term is
        // fmaf(0, x, acc) = acc.  (The generic loop compiles to ~45 instructions of unrolling scaffolding per call.)
        if (P.M == 2) return c.cpar[1 - P.m] * c.gaf[(size_t)(1 - P.m) * c.gaf_stride + t];
        if (P.M == 3) {     // the two other views in increasing order, as the loop visits them
            const int i1 = (P.m == 0) ? 1 : 0, i2 = (P.m == 2) ? 1 : 2;
            return fmaf(c.cpar[i2], c.gaf[(size_t)i2 * c.gaf_stride + t], c.cpar[i1] * c.gaf[(size_t)i1 * c.gaf_stride + t]);
        }
        for (int i = 0; i < P.M; i++) pri = fmaf(c.cpar[i], c.gaf[(size_t)i * c.gaf_stride + t], pri);
    }
    return pri;
}

// prior_other for the token loop: same sums, addressed from the two .shared bases held in registers (c.sa: cpar at +6 KS + 256;
// c.ginv_sa: gaf at +12 KS) -- through the generic pointers ptxas rebuilds both addresses from %tid on every call (~25 instructions)
template <int KS, bool MULTI>
__device__ __forceinline__ float prior_other_sa(const SweepParams &P, const DocCtx &c, int t)
{
    float pri = 0.f;
    if (MULTI) {
        const uint32_t cp = c.sa + (uint32_t)KS * 6u + 256u, ga = c.ginv_sa + (uint32_t)KS * 12u + 4u * (uint32_t)t;
        if (P.M == 2) return lds_f32(cp + 4u * (uint32_t)(1 - P.m)) * lds_f32(ga + (uint32_t)(1 - P.m) * (KS * 4u));
        if (P.M == 3) {
            const int i1 = (P.m == 0) ? 1 : 0, i2 = (P.m == 2) ? 1 : 2;
            return fmaf(lds_f32(cp + 4u * (uint32_t)i2), lds_f32(ga + (uint32_t)i2 * (KS * 4u)), lds_f32(cp + 4u * (uint32_t)i1) * lds_f32(ga + (uint32_t)i1 * (KS * 4u)));
        }
        for (int i = 0; i < P.M; i++) pri = fmaf(lds_f32(cp + 4u * (uint32_t)i), lds_f32(ga + (uint32_t)i * (KS * 4u)), pri);
    }
    return pri;
}

// n_d[t] += dl, then recompute q[t].  Executed by the lane that
// owns topic t (gl == (t >> 2) % G), W:434-471 / W:557-584.
// Q1 (reference-exact dense index, MVTM_FLAG_Q1_COMPAT): bit 15 of n_d[t] flags "t is not in S for the rest of the sweep".  The
// reference removes a topic from S when no view of the document holds it any more (W:441-468) and never inserts one (W:563-584 is
// dead code), so S = {held and not flagged} with the flag set when the topic leaves or is gained while absent
// (tests/test_reference_vectors.py::test_flag_rule_equals_reference_dense_index).  A flagged topic keeps only its tree mass.
template <int KS, int G, bool MULTI, bool Q1>
__device__ __forceinline__ void apply_count_delta(const SweepParams &P, DocCtx &c, int t, int dl)
{
    const uint32_t nd_a = c.sa + (uint32_t)KS * 4u + 2u * (uint32_t)t;
    const unsigned raw = lds_u16(nd_a);
    unsigned flag = Q1 ? (raw >> 15) : 0u;
    const unsigned before = Q1 ? (raw & 0x7fffu) : raw;
    const unsigned ndv_i = (unsigned)((int)before + dl);
    bool oth = false;
    if (MULTI) oth = (lds_u32(c.sa + (uint32_t)KS * 6u + 4u * (uint32_t)(t >> 5)) >> (t & 31)) & 1u;      // c.om[t >> 5]
    if (Q1) {
        const bool leaves = (dl < 0) && (ndv_i == 0u) && !oth;
        const bool gained_absent = (dl > 0) && (before == 0u) && !oth;
        if (!flag && (leaves || gained_absent)) {
            flag = 1u;
            if (c.rb) atomicOr(c.rb + (t >> 5), 1u << (t & 31));
        }
    }
    sts_u16(nd_a, Q1 ? (ndv_i | (flag << 15)) : ndv_i);
    const float ndv = (Q1 && flag) ? 0.f : (float)ndv_i;
    bool inS = false; float ocv = 0.f, pri = 0.f;
    if (MULTI) {
        // branch-light on purpose: the lane groups of a warp hold different documents and would serialise on branches
        inS = ((ndv_i > 0u) || oth) && !(Q1 && flag);
        pri = prior_other_sa<KS, MULTI>(P, c, t);
        if (oth) ocv = c.oc[t];
    }
    sts_f32(c.sa + 4u * (uint32_t)t, q_value<MULTI>(ndv, inS, ocv, pri, c.coefm, c.pmm, lds_f2(c.ginv_sa + 8u * (uint32_t)t)));
}

// one step: n_d[tinc]++ and n_d[tdec]-- (either may be -1 = none) by their owner lanes, in parallel when the owners
// differ (a second pass runs only when one lane owns both).  No warp-level synchronisation inside: the lane groups of a
// warp may diverge here.
template <int KS, int G, bool MULTI, bool Q1>
__device__ __forceinline__ void apply_pair(const SweepParams &P, DocCtx &c, int tinc, int tdec, int gl)
{
    if (tinc == tdec) return;                                   // same topic: the two changes cancel
    const int own_i = tinc >= 0 ? ((tinc >> 2) & (G - 1)) : -1, own_d = tdec >= 0 ? ((tdec >> 2) & (G - 1)) : -2;
    int t = -1, dl = 0, t2 = -1;
    if (gl == own_i) { t = tinc; dl = 1; if (own_i == own_d) t2 = tdec; } else if (gl == own_d) { t = tdec; dl = -1; }
#pragma unroll 1
    while (t >= 0) {
        apply_count_delta<KS, G, MULTI, Q1>(P, c, t, dl);
        t = t2; t2 = -1; dl = -1;
    }
}

// cc.mallet.util.Randoms.nextBeta(a, b) -- the law the reference draws p[m][i] from at W:333 -- restated literally over a
// counter-based stream (block k of the stream = Philox(k, c1, c2, c3); its words 1..3 are used, word 0 of block 0 is the default
// law's inversion uniform).  Quirk Q5: for a > 1, b == 1 the acceptance test compares against NaN (0 * log(inf)), so the first
// normal proposal inside [0, 1] is returned: a truncated N(1, 0.25/(a-1)) instead of Beta(a, 1).  a, b < 1 (the burn-in ramp
// starts at 0.31): Joehnk's method, which IS Beta(a, b).  Pinned against draws of the MALLET jar's own bytecode by
// tests/test_optim_host.py::test_engine_mallet_beta_law_matches_mallet_bytecode (mvtm_test_sampler which = 4).
struct BetaStream {
    uint32_t c1, c2, c3, k0, k1, blk; uint4 x; int used;
    __host__ __device__ double next()
    {   // 24-bit uniforms, three per Philox block
        if (used == 3) { x = philox4x32_10(blk++, c1, c2, c3, k0, k1); used = 0; }
        const uint32_t w = used == 0 ? x.y : (used == 1 ? x.z : x.w);
        used++;
        return (double)(w >> 8) * (1.0 / 16777216.0);
    }
    __host__ __device__ double gauss()
    {   double u1; do { u1 = next(); } while (u1 <= 0.0); const double u2 = next();
        return sqrt(-2.0 * log(u1)) * cos(2.0 * 3.14159265358979323846 * u2); }
};
__host__ __device__ inline double mallet_next_beta(BetaStream &r, double a, double b)
{
    if (a == 1.0 && b == 1.0) return r.next();
    if (a >= 1.0 && b >= 1.0) {
        const double A = a - 1.0, B = b - 1.0, C = A + B, L = C * log(C), mu = A / C, sigma = 0.5 / sqrt(C);
        double y = r.gauss(), x = sigma * y + mu;
        int guard = 0;
        while ((x < 0.0 || x > 1.0) && ++guard < 4096) { y = r.gauss(); x = sigma * y + mu; }
        double u = r.next();
        // b == 1: B*log((1-x)/B) = 0*log(inf) = NaN -> the comparison is false -> the first proposal is accepted (Q5)
        while (log(u) >= A * log(x / A) + B * log((1.0 - x) / B) + L + 0.5 * y * y && ++guard < 4096) {
            y = r.gauss(); x = sigma * y + mu;
            while ((x < 0.0 || x > 1.0) && ++guard < 4096) { y = r.gauss(); x = sigma * y + mu; }
            u = r.next();
        }
        return x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x);
    }
    double v1, v2; int guard = 0;
    do { v1 = pow(r.next(), 1.0 / a); v2 = pow(r.next(), 1.0 / b); } while (v1 + v2 > 1.0 && ++guard < 4096);
    return v1 / (v1 + v2);
}

// the view-coupling draw p[m][i] of W:327-337 for document gdoc (every lane of the group computes the same value).  Two functions,
// chosen by a uniform branch at the call site: the default law's has no nested call (putting MALLET's law behind a branch INSIDE it
// cost the short side-view passes 15 %, measured: profiles/r2_ab_compat_build_vs_round_start.log).
template <bool MALLET>
__device__ __noinline__ float draw_p_law(const SweepParams &P, int i, uint32_t gdoc, const double *p_override)
{
    int m = P.m;
    double r;
    if (p_override) r = p_override[i];
    else if (i == m) r = 1.0;
    else if (P.pa[i] == 0.0) r = 0.0;
    else {
        int lo = m < i ? m : i, hi = m < i ? i : m;
        const uint32_t tag = ((uint32_t)(lo * P.M + hi) << 8) | PURPOSE_PDRAW;
        uint4 x = philox4x32_10(0u, gdoc, P.iteration, tag, P.seed_lo, P.seed_hi);
        double b;
        if (MALLET) {
            BetaStream st{ gdoc, P.iteration, tag, P.seed_lo, P.seed_hi, 1u, x, 0 };
            b = mallet_next_beta(st, P.pa[i], P.pb[i]);
        } else {
            double u = (double)(x.x >> 8) * (1.0 / 16777216.0);
            b = pow(u, 1.0 / P.pa[i]);                          // Beta(a,1) by inversion (Q5: true law)
        }
        r = floor(1000.0 * b + 0.5) / 1000.0;                   // W:333 Math.round(1000*x)/1000
    }
    if (!p_override && i != 0 && P.sparse[i]) r = 0.0;          // W:335-336 (column i zeroed, incl. the diagonal)
    return (float)r;
}
__device__ __forceinline__ float draw_p(const SweepParams &P, int i, uint32_t gdoc, const double *p_override)
{ return P.beta_mallet ? draw_p_law<true>(P, i, gdoc, p_override) : draw_p_law<false>(P, i, gdoc, p_override); }

// Build n_d, (MULTI: oc/om/cpar) and q for document d of view P.m (len == 0: nothing to do, but every
// lane still walks the same __syncwarp sequence -- the groups of a warp hold different documents).
//
// Cost is O(KS/G) cheap vector work + O(tokens of the document, all views): q starts as the document-independent vector
// ga[t] / (n_k[t] + betaSum) (topics outside S, W:495-513) and only the topics some token of the document holds are
// recomputed; the other views' counts are histogrammed INTO the (still empty) n_d array and consumed by exchange, which
// leaves it empty again, so nothing is scanned densely.  (The first version zeroed, histogrammed and scanned all KS
// topics per other view and evaluated the full q expression for every topic: ~4000 instructions per document at
// K = 1000 -- more than sampling the 6-12 tokens of a side view.)
template <int KS, int G, bool MULTI, bool Q1>
__device__ __forceinline__ void doc_setup(const SweepParams &P, DocCtx &c, int d, int len, int gl, const double *p_override,
                                          bool skip_first)
{
    constexpr int JG = KS / (4 * G);
    const int m = P.m;
    c.rb = nullptr;
    if (Q1) c.rb = P.rb_one ? P.rb_one : ((MULTI && P.rbits) ? P.rbits + (size_t)d * (KS / 32) : nullptr);   // one pass: nothing to carry
    const long long b = P.doc_off[m][d];
    const uint32_t gdoc = (uint32_t)(P.doc_id_base + (long long)d * P.doc_id_stride);
    unsigned *nd32 = reinterpret_cast<unsigned *>(c.nd);
    {   // n_d = 0 (128-bit stores) and q = ga * ginv for every topic
        uint4 *nd128 = reinterpret_cast<uint4 *>(c.nd);
#pragma unroll
        for (int k = 0; k < (KS / 8 + G - 1) / G; k++) { const int i = gl + G * k; if (i < KS / 8) nd128[i] = make_uint4(0u, 0u, 0u, 0u); }
#pragma unroll(MULTI ? 2 : JG)
        for (int j = 0; j < JG; j++) {
            const int cidx = gl + G * j;
            const float4 g01 = lds_f4(c.ginv_sa + 32u * (uint32_t)cidx), g23 = lds_f4(c.ginv_sa + 32u * (uint32_t)cidx + 16u);
            reinterpret_cast<float4 *>(c.q)[cidx] = make_float4(g01.x * g01.y, g01.z * g01.w, g23.x * g23.y, g23.z * g23.w);
        }
    }
    c.pmm = 1.f; c.coefm = (float)len + P.gas[m]; c.C = 0.f;
    if (MULTI) {
        c.pmm = draw_p(P, m, gdoc, p_override);
        for (int k = gl; k < KS / 32; k += G) c.om[k] = 0u;
        float cdoc = 0.f;
        for (int i = 0; i < P.M; i++) {                           // uniform trip count; per-group work is predicated
            const long long bi = P.doc_off[i][d];
            const int leni = (int)(P.doc_off[i][d + 1] - bi);
            const float pmi = (i == m) ? c.pmm : draw_p(P, i, gdoc, p_override);
            const float denom = (float)leni + P.gas[i];
            cdoc += pmi * P.ga_new[i] / denom;                    // W:414-416 (every view, no length test)
            const bool other = (i != m && leni != 0 && len != 0);
            const float ci = other ? pmi / denom : 0.f;           // W:403-404
            __syncwarp();
            if (gl == 0) c.cpar[i] = ci;
            if (other) {                                          // histogram of view i into the empty n_d array
                const int *zi = P.zv[i] + bi;
                for (int k = gl; k < leni; k += G) { int t = zi[k]; if (t >= 0) atomicAdd(&nd32[t >> 1], 1u << ((t & 1) * 16)); }
            }
            __syncwarp();
            if (other) {                                          // consume it: the first lane to reach a topic takes its count
                const int *zi = P.zv[i] + bi;
                for (int k = gl; k < leni; k += G) {
                    const int t = zi[k];
                    if (t < 0) continue;
                    const unsigned sh = (unsigned)(t & 1) * 16u;
                    const unsigned cnt = (atomicAnd(&nd32[t >> 1], ~(0xffffu << sh)) >> sh) & 0xffffu;
                    if (cnt) {        // first view to hold the topic stores, later views accumulate (one lane per topic and view)
                        const bool seen = (c.om[t >> 5] >> (t & 31)) & 1u;
                        c.oc[t] = (seen ? c.oc[t] : 0.f) + ci * (float)cnt;
                        if (!seen) atomicOr(&c.om[t >> 5], 1u << (t & 31));
                    }
                }
            }
        }
        c.C = (P.n_inactive > 0) ? (cdoc * c.coefm) / (float)P.K : 0.f;   // W:418, W:515
    } else {
        // single view: p[0][0] = 1; C_doc = coefm * gamma*alpha[K] / (len + gas)  (W:413-418 with M = 1)
        c.C = (P.n_inactive > 0) ? (P.ga_new[m] / ((float)len + P.gas[m]) * c.coefm) / (float)P.K : 0.f;
    }
    __syncwarp();
    if (Q1) {
        // flags raised while an earlier view of this document was sampled in this sweep: the reference keeps ONE dense index per
        // document across its views (W:376-391 builds it once), so they still apply.  The barrier below is reached by every lane of
        // the warp (the lane groups hold different documents: the WORK is predicated, never the __syncwarp)
        if (c.rb && len != 0) {
            for (int k = gl; k < KS / 32; k += G) {
                unsigned wbits = c.rb[k];
                while (wbits) {
                    const int t = 32 * k + (__ffs(wbits) - 1);
                    wbits &= wbits - 1u;
                    atomicOr(&nd32[t >> 1], 0x8000u << ((t & 1) * 16));
                }
            }
        }
        __syncwarp();
    }
    int t_first = -1;
    {   // own-view histogram (W:352-359).  With skip_first the document's first token is left out: it is the first to
        // be resampled, so its removal (W:434-471) is folded into the setup
        const int *zm = P.zv[m] + b;
        for (int k = gl; k < len; k += G) {
            int t = zm[k];
            if (k == 0 && skip_first && (unsigned)__ldg(P.word + b) < (unsigned)P.V) { t_first = t; t = -1; }
            if (t >= 0) atomicAdd(&nd32[t >> 1], 1u << ((t & 1) * 16));
        }
    }
    __syncwarp();
    if (Q1 && t_first >= 0) {
        // the folded removal of the first token may take its topic out of the index (no view holds it any more, W:441-468)
        const unsigned raw = c.nd[t_first];
        bool oth = false;
        if (MULTI) oth = (c.om[t_first >> 5] >> (t_first & 31)) & 1u;
        if (raw == 0u && !oth) {
            c.nd[t_first] = (unsigned short)0x8000u;
            if (c.rb) atomicOr(c.rb + (t_first >> 5), 1u << (t_first & 31));
        }
    }
    if (Q1) __syncwarp();
    // q of the topics some token of the document holds (S of W:376-391 plus, harmlessly, the skipped first token's topic):
    // one token per lane, duplicates recompute the same value
    for (int i = 0; i < (MULTI ? P.M : 1); i++) {
        const int vi = MULTI ? i : m;
        const long long bi = P.doc_off[vi][d];
        const int leni = (len != 0) ? (int)(P.doc_off[vi][d + 1] - bi) : 0;
        const int *zi = P.zv[vi] + bi;
        for (int k = gl; k < leni; k += G) {
            const int t = zi[k];
            if (t < 0) continue;
            const unsigned raw = c.nd[t];
            if (Q1 && (raw >> 15)) continue;                      // held but not in the index: only its tree mass, already in q
            const float ndv = (float)(Q1 ? (raw & 0x7fffu) : raw);
            bool inS = false; float ocv = 0.f, pri = 0.f;
            if (MULTI) {
                const bool oth = (c.om[t >> 5] >> (t & 31)) & 1u;
                inS = (ndv > 0.f) || oth;
                if (inS) { ocv = oth ? c.oc[t] : 0.f; pri = prior_other<MULTI>(P, c, t); }
            }
            c.q[t] = q_value<MULTI>(ndv, inS, ocv, pri, c.coefm, c.pmm, lds_f2(c.ginv_sa + 8u * (uint32_t)t));
        }
    }
    __syncwarp();
}

// weight of one topic: (n + beta) * q evaluated as n*q + beta*q
__device__ __forceinline__ float topic_weight(int n, float q, float beta) { return fmaf(__int2float_rn(n), q, beta * q); }

// packed fp32 pairs (sm_100: add/mul/fma.f32x2 -- one issue slot for two lanes of arithmetic)
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b)
{ unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b))); return *reinterpret_cast<float2 *>(&d); }
__device__ __forceinline__ float2 f2_add(float2 a, float2 b)
{ unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b))); return *reinterpret_cast<float2 *>(&d); }
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c)
{ unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)), "l"(*reinterpret_cast<unsigned long long *>(&c))); return *reinterpret_cast<float2 *>(&d); }

// weight of one 4-topic chunk, sum_e (n_e + beta) * q_e, evaluated as beta * (q0 + q2, q1 + q3) + n * q on packed pairs: 4 packed
// instructions + 1 add for 4 topics, and nothing to maintain per chunk when a q changes (the first kernels kept beta * sum(q) of every
// chunk in registers: 2 JG instructions of select chain per n_d update, JG registers)
__device__ __forceinline__ float chunk_weight(const int4 r, const float4 qq, const float2 bb)
{
    float2 acc = f2_mul(f2_add(make_float2(qq.x, qq.y), make_float2(qq.z, qq.w)), bb);
    acc = f2_fma(make_float2(__int2float_rn(r.x), __int2float_rn(r.y)), make_float2(qq.x, qq.y), acc);
    acc = f2_fma(make_float2(__int2float_rn(r.z), __int2float_rn(r.w)), make_float2(qq.z, qq.w), acc);
    return acc.x + acc.y;
}

// per-lane weights of one row: cum[j] = running sum over the lane's chunks 0..j, each topic weighted (n + beta) * q.
// Returns the lane total (= cum[JG-1]).
template <int JG, int G>
__device__ __forceinline__ float lane_weights(uint32_t row_gl_sa, uint32_t q_gl_sa, float beta, float (&cum)[JG])
{   // row_gl_sa / q_gl_sa: .shared address of the lane's first chunk (base + 16*gl); chunk j sits 16*G*j bytes further
    float tot = 0.f;
    const float2 bb = make_float2(beta, beta);
#pragma unroll
    for (int j = 0; j < JG; j++) {
        tot += chunk_weight(lds_i4(row_gl_sa + 16u * G * j), lds_f4(q_gl_sa + 16u * G * j), bb);
        cum[j] = tot;
    }
    return tot;
}

// largest float below a positive x
__device__ __forceinline__ float next_below(float x) { return __int_as_float(__float_as_int(x) - 1); }

// group-cooperative selection: returns the topic whose cumulative weight (lane-major scan order inside the group) first
// exceeds target = u*(total + C) - C; -1 if the draw fell into the new-topic bucket (W:522).  Every lane of the warp
// must call it (full-mask shuffles); groups whose document is exhausted compute on stale data and ignore the result.
template <int JG, int G>
__device__ __forceinline__ int group_select(uint32_t row_gl_sa, uint32_t q_gl_sa, int lane, int gl, float beta, float u, float C)
{
    float cum[JG];
    const float lane_total = lane_weights<JG, G>(row_gl_sa, q_gl_sa, beta, cum);
    float incl = lane_total;
#pragma unroll
    for (int off = 1; off < G; off <<= 1) { float v = __shfl_up_sync(FULL, incl, off, G); if (gl >= off) incl += v; }
    const float total = __shfl_sync(FULL, incl, G - 1, G);
    float target = u * (total + C);
    const bool newbucket = (C > 0.f) && (target < C);
    target -= C;
    // the target is kept strictly below the total, so some lane always satisfies incl > target and the first one that does
    // has positive weight (lanes of weight zero repeat their predecessor's inclusive sum)
    target = fminf(target, next_below(total));
    const unsigned sh = (unsigned)(lane - gl);                                  // first lane of my group
    const unsigned gmask = (G == 32) ? FULL : ((1u << (G & 31)) - 1u);
    const unsigned hit = (__ballot_sync(FULL, incl > target) >> sh) & gmask;
    const int L = __ffs(hit) - 1;
    // residual inside the lane, clamped strictly below the lane total so that the chunk found has positive weight
    const float r = fminf(target - (incl - lane_total), next_below(lane_total));
    // chunk: jsel = number of running sums <= r, base = the last of them (0 if none).  Binary search over the register-resident
    // sums: one compare per level, the candidates of the next level chosen by selects (JG-1-log2 selects instead of 2 per sum)
    int jsel = 0; float base = 0.f;
    if (G == 32) {
        constexpr int P2 = (JG <= 1) ? 1 : (JG <= 2) ? 2 : (JG <= 4) ? 4 : (JG <= 8) ? 8 : 16;
        float win[P2];                                                          // thresholds cum[0..JG-2], +inf padded
#pragma unroll
        for (int j = 0; j < P2; j++) win[j] = (j < JG - 1) ? cum[j] : __int_as_float(0x7f800000);
#pragma unroll
        for (int half = P2 / 2; half >= 1; half >>= 1) {
            const float t = win[half - 1];
            const bool ge = (r >= t);
            base = ge ? t : base;
            jsel += ge ? half : 0;
#pragma unroll
            for (int j = 0; j + 1 < half; j++) win[j] = ge ? win[half + j] : win[j];
        }
    } else {                    // several documents per warp: the independent compares of the linear form schedule better (measured)
#pragma unroll
        for (int j = 0; j < JG - 1; j++) { const bool ge = (r >= cum[j]); jsel += ge ? 1 : 0; base = ge ? cum[j] : base; }
    }
    const int cidx = gl + G * jsel;
    const int4 rr = lds_i4(row_gl_sa + 16u * G * (uint32_t)jsel);
    const float4 qq = lds_f4(q_gl_sa + 16u * G * (uint32_t)jsel);
    const float w0 = topic_weight(rr.x, qq.x, beta), w1 = topic_weight(rr.y, qq.y, beta);
    const float w2 = topic_weight(rr.z, qq.z, beta), w3 = topic_weight(rr.w, qq.w, beta);
    const float c1 = w0 + w1, c2 = c1 + w2, c3 = c2 + w3;
    const float r2 = fminf(r - base, next_below(c3));                           // same clamp one level down
    const int e = (r2 >= w0 ? 1 : 0) + (r2 >= c1 ? 1 : 0) + (r2 >= c2 ? 1 : 0);
    const int mine = 4 * cidx + e;
    const int sel = __shfl_sync(FULL, mine, L, G);
    return newbucket ? -1 : sel;
}


// ------------------------------------------------------------------------------------------------
// DIRECT mode: the n_wk row of a token lives in REGISTERS (JG 128-bit words per lane, loaded straight from global memory / L2
// with ld.global.cg one token ahead) instead of a shared-memory ring filled by the TMA engine.  A document then needs only
// q + n_d in shared memory (6 KS bytes instead of 10 KS), so an SM holds ~1.6x the documents, which is what lets a warp host
// two (or four) document-views at K = 1024 (K = 512) without running out of warps.  Same arithmetic, same scan order.
// ------------------------------------------------------------------------------------------------
template <int JG, int G>
__device__ __forceinline__ void load_row_regs(int4 (&row)[JG], const int *rowp, int gl)
{   // .cg: L2 only -- the rows are modified by other SMs' RED atomics during the pass, an L1 copy could be arbitrarily stale.
    // Chunks beyond the row (KS > Kp) read the following row or the table's tail padding; their q is 0, so they weigh nothing
    const int4 *p = reinterpret_cast<const int4 *>(rowp) + gl;
#pragma unroll
    for (int j = 0; j < JG; j++) row[j] = __ldcg(p + G * j);
}

template <int JG, int G>
__device__ __forceinline__ float lane_weights_regs(const int4 (&row)[JG], uint32_t q_gl_sa, float beta, float (&cum)[JG])
{   // lane_weights with the row in registers (bit-identical arithmetic)
    float tot = 0.f;
    const float2 bb = make_float2(beta, beta);
#pragma unroll
    for (int j = 0; j < JG; j++) {
        tot += chunk_weight(row[j], lds_f4(q_gl_sa + 16u * G * j), bb);
        cum[j] = tot;
    }
    return tot;
}

// group_select over a register-resident row.  The chunk is found by bisection over the running sums and the row's 128-bit word of
// that chunk is carried along by selects (registers cannot be indexed); everything else as group_select.
template <int JG, int G>
__device__ __forceinline__ int group_select_regs(const int4 (&row)[JG], uint32_t q_gl_sa, int lane, int gl, float beta, float u, float C)
{
    float cum[JG];
    const float lane_total = lane_weights_regs<JG, G>(row, q_gl_sa, beta, cum);
    float incl = lane_total;
#pragma unroll
    for (int off = 1; off < G; off <<= 1) { float v = __shfl_up_sync(FULL, incl, off, G); if (gl >= off) incl += v; }
    const float total = __shfl_sync(FULL, incl, G - 1, G);
    float target = u * (total + C);
    const bool newbucket = (C > 0.f) && (target < C);
    target -= C;
    target = fminf(target, next_below(total));
    const unsigned sh = (unsigned)(lane - gl);
    const unsigned gmask = (G == 32) ? FULL : ((1u << (G & 31)) - 1u);
    const unsigned hit = (__ballot_sync(FULL, incl > target) >> sh) & gmask;
    const int L = __ffs(hit) - 1;
    const float r = fminf(target - (incl - lane_total), next_below(lane_total));
    int jsel = 0; float base = 0.f;
    constexpr int P2 = (JG <= 1) ? 1 : (JG <= 2) ? 2 : (JG <= 4) ? 4 : (JG <= 8) ? 8 : (JG <= 16) ? 16 : 32;
    float win[P2]; int4 rw[P2];
#pragma unroll
    for (int j = 0; j < P2; j++) { win[j] = (j < JG - 1) ? cum[j] : __int_as_float(0x7f800000); rw[j] = row[j < JG ? j : JG - 1]; }
#pragma unroll
    for (int half = P2 / 2; half >= 1; half >>= 1) {
        const float t = win[half - 1];
        const bool ge = (r >= t);
        base = ge ? t : base;
        jsel += ge ? half : 0;
#pragma unroll
        for (int j = 0; j + 1 < half; j++) win[j] = ge ? win[half + j] : win[j];
#pragma unroll
        for (int j = 0; j < half; j++) {
            rw[j].x = ge ? rw[half + j].x : rw[j].x; rw[j].y = ge ? rw[half + j].y : rw[j].y;
            rw[j].z = ge ? rw[half + j].z : rw[j].z; rw[j].w = ge ? rw[half + j].w : rw[j].w;
        }
    }
    const int cidx = gl + G * jsel;
    // (a branch on the selected lane's chunk index -- a switch over the JG registers, <= 32/G targets per warp -- instead of the
    // 4 (P2 - 1) selects was measured 5 % slower: the taken branches cost more than the selects they save)
    const int4 rr = rw[0];
    const float4 qq = lds_f4(q_gl_sa + 16u * G * (uint32_t)jsel);
    const float w0 = topic_weight(rr.x, qq.x, beta), w1 = topic_weight(rr.y, qq.y, beta);
    const float w2 = topic_weight(rr.z, qq.z, beta), w3 = topic_weight(rr.w, qq.w, beta);
    const float c1 = w0 + w1, c2 = c1 + w2, c3 = c2 + w3;
    const float r2 = fminf(r - base, next_below(c3));
    const int e = (r2 >= w0 ? 1 : 0) + (r2 >= c1 ? 1 : 0) + (r2 >= c2 ? 1 : 0);
    const int mine = 4 * cidx + e;
    const int sel = __shfl_sync(FULL, mine, L, G);
    return newbucket ? -1 : sel;
}

// shared-memory carve-up -------------------------------------------------------------------------
__host__ __device__ inline size_t smem_cta_bytes(int KS, int M_multi) { return (size_t)KS * 8 + (size_t)KS * 4 + (size_t)M_multi * KS * 4; }
// per-document area: [q KS*4][n_d KS*2][om 256 + cpar 64 (multi)][mbarriers 128][ring R*KS*4][oc KS*4 (multi, optional)]
// -- everything the hot loop addresses sits at a compile-time offset from the area's base, the ring (whose size depends on
// the run-time depth R) comes last
__host__ __device__ constexpr uint32_t doc_off_mbar(int KS, bool multi) { return (uint32_t)KS * 6u + (multi ? 320u : 0u); }
__host__ __device__ constexpr uint32_t doc_off_ring(int KS, bool multi) { return doc_off_mbar(KS, multi) + 128u; }
__host__ __device__ inline size_t smem_doc_bytes(int KS, int R, bool multi, bool oc_smem = false)
{
    size_t b = doc_off_ring(KS, multi) + (size_t)R * KS * 4;
    if (R == 0) b = doc_off_mbar(KS, multi);                      // DIRECT mode: rows in registers, no mbarriers and no ring
    if (multi && oc_smem) b += (size_t)KS * 4;                    // oc in shared memory
    return (b + 127) & ~(size_t)127;
}

__device__ __forceinline__ void carve_doc(unsigned char *base, int KS, int R, bool multi, DocCtx &c, int *&ring, unsigned long long *&mbar,
                                          bool oc_smem = false)
{
    c.sa = smem_u32(base);
    c.q = reinterpret_cast<float *>(base);
    c.nd = reinterpret_cast<unsigned short *>(base + (size_t)KS * 4);
    if (multi) { c.om = reinterpret_cast<unsigned *>(base + (size_t)KS * 6); c.cpar = reinterpret_cast<float *>(base + (size_t)KS * 6 + 256); }
    else { c.om = nullptr; c.cpar = nullptr; }
    mbar = reinterpret_cast<unsigned long long *>(base + doc_off_mbar(KS, multi));
    ring = reinterpret_cast<int *>(base + doc_off_ring(KS, multi));
    c.oc = nullptr;
    if (multi && oc_smem) c.oc = reinterpret_cast<float *>(base + doc_off_ring(KS, multi) + (size_t)R * KS * 4);
}

// ------------------------------------------------------------------------------------------------
// THE sweep kernel: one launch = one view pass over all documents of the shard  (W:186-233, W:301-597, U:197-218)
// ------------------------------------------------------------------------------------------------
// Threads per CTA the kernel is compiled for.  Shared memory, not the register file, limits how many documents an SM holds (one
// CTA per SM, smem_doc_bytes each), so the launch bound is the largest warp count the shared-memory budget admits for this
// (KS, G, >= 2 views if MULTI) at ring depth 1 -- and ptxas may use 65536 / bound registers per thread: 96 instead of 80 at
// K = 1024.  With the flat 768-thread bound of round 1 the K = 1024 kernel sat at the 80-register cap and its quality depended on
// the build (spilled loop-carried values in some, profiles/r2_ab_codegen_variants*.log).
__host__ __device__ constexpr int sweep_max_threads(int KS, int G, bool multi, bool direct = false)
{
    const size_t cta = (size_t)KS * 12 + (multi ? (size_t)2 * KS * 4 : 0);
    const size_t doc = direct ? (((size_t)KS * 6u + (multi ? 320u : 0u) + 127) & ~(size_t)127)                       // smem_doc_bytes(KS, 0, multi)
                              : (((size_t)KS * 6u + (multi ? 320u : 0u) + 128u + (size_t)KS * 4 + 127) & ~(size_t)127);  // smem_doc_bytes(KS, 1, multi)
    const int docs = (int)((227 * 1024 - 1024 - cta) / doc);
    int W = docs / (32 / G);
    // DIRECT: the row costs KS/G registers per thread on top of ~100, so the register file (64 K), not shared memory, is the bound
    const int wreg = direct ? (65536 / 32) / direct_regs(KS, G) : 24;
    W = W > wreg ? wreg : W;
    W = W > 24 ? 24 : (W < 1 ? 1 : W);
    return 32 * W;
}

template <int KS, int G, bool MULTI, bool Q1, bool DIRECT>
__device__ __forceinline__ void sweep_view_body(const SweepParams &P)
{
    constexpr int JG = KS / (4 * G), NSUB = 32 / G;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / G, gl = lane % G;
    const int R = DIRECT ? 0 : P.R;               // ring depth (DIRECT: no ring, the row is loaded into registers one token ahead)
    const int RA = DIRECT ? 1 : P.R;              // how many tokens ahead the row fetch runs

    {
        float2 *ginv = reinterpret_cast<float2 *>(smem);
        int *dnk = reinterpret_cast<int *>(smem + (size_t)KS * 8);
        for (int t = threadIdx.x; t < KS; t += blockDim.x) {
            float2 g = make_float2(0.f, 0.f);
            if (t < P.K) { g.x = __ldg(P.ga_tree + t); g.y = 1.0f / ((float)__ldg(P.nk_frozen + t) + P.betaSum); }
            ginv[t] = g;                                              // topics >= K get weight 0
            dnk[t] = 0;
        }
        if (MULTI) {
            float *gaf = reinterpret_cast<float *>(smem + (size_t)KS * 12);
            for (int i = 0; i < P.M; i++)
                for (int t = threadIdx.x; t < KS; t += blockDim.x) gaf[(size_t)i * KS + t] = (t < P.K) ? __ldg(P.ga_full[i] + t) : 0.f;
        }
    }
    DocCtx c; int *ring; unsigned long long *mbar;
    carve_doc(smem + smem_cta_bytes(KS, MULTI ? P.M : 0) + (size_t)(warp * NSUB + sub) * smem_doc_bytes(KS, R, MULTI, P.oc_smem != 0), KS, R, MULTI,
              c, ring, mbar, P.oc_smem != 0);
    c.gaf = reinterpret_cast<const float *>(smem + (size_t)KS * 12); c.gaf_stride = KS;
    if (MULTI && !P.oc_smem) c.oc = P.oc_scratch + ((size_t)blockIdx.x * (blockDim.x >> 5) * NSUB + (size_t)(warp * NSUB + sub)) * P.Kp;
    uint32_t cta_sa = smem_u32(smem);
    asm volatile("" : "+r"(cta_sa));                              // opaque: hold the base in a register
    c.ginv_sa = cta_sa;
    const uint32_t dnk_sa = cta_sa + (uint32_t)KS * 8u;
    if (!DIRECT) {
        if (gl == 0) { for (int s = 0; s < R; s++) mbar_init(smem_u32(mbar + s), 1); fence_mbar_init(); }
        for (int k = gl; k < R * KS; k += G) ring[k] = 0;
        fence_proxy_async();
    }
    __syncthreads();

    const uint32_t row_bytes = (uint32_t)P.Kp * 4u;
    asm volatile("" : "+r"(c.sa));                                // opaque: ONE register carries the document area's base
    const uint32_t mbar_u32 = c.sa + doc_off_mbar(KS, MULTI), ring_u32 = c.sa + doc_off_ring(KS, MULTI);
    const uint32_t q_gl_sa = c.sa + 16u * (uint32_t)gl;
    unsigned phasebits = 0u;
    unsigned n_tok = 0, n_changed = 0, n_new = 0;                 // per lane group: far below 2^32 per launch
    const int m = P.m;
    int *zmv = P.zv[m];
    int4 row[DIRECT ? JG : 1];                                    // DIRECT: the current token's n_wk row (this lane's chunks)

    // Work items are claimed two documents ahead and the next document's offsets and first tokens are loaded while the
    // current one is being sampled, so a new document starts with everything in registers (short documents -- side
    // views, SMS-sized corpora -- would otherwise pay three dependent global-memory latencies each).
    auto claim = [&]() { int it = 0; if (lane == 0) it = atomicAdd(P.work_counter, NSUB); return it; };
    // the row a token's TMA load fetches: its word's, or row 0 for an out-of-vocabulary id (the token is skipped, W:427-428, but
    // its slot of the ring is still filled and consumed) -- decided once when the word is loaded, not per token
    auto row_word = [&](int w) { return ((unsigned)w < (unsigned)P.V) ? w : 0; };
    int item0 = __shfl_sync(FULL, claim(), 0);
    int item_next_raw = claim();                                   // lane 0 holds it; broadcast when consumed
    int d = 0, len = 0, wcur = 0, zcur = -1, wnext = 0, wahead = 0;
    long long b = 0;
    {
        const bool have = item0 + sub < P.n_items;
        d = have ? __ldg(P.order + item0 + sub) : 0;
        b = P.doc_off[m][d];
        len = have ? (int)(P.doc_off[m][d + 1] - b) : 0;
        wcur = (gl < len) ? __ldg(P.word + b + gl) : 0;
        zcur = (gl < len) ? zmv[b + gl] : -1;
        wnext = (G + gl < len) ? __ldg(P.word + b + G + gl) : 0;
        wahead = (RA + gl < len) ? row_word(__ldg(P.word + b + RA + gl)) : 0;  // word of the token RA positions ahead
    }
    while (item0 < P.n_items) {
        const int maxlen = (NSUB == 1) ? len : __reduce_max_sync(FULL, len);
        const uint32_t gdoc = (uint32_t)(P.doc_id_base + (long long)d * P.doc_id_stride);
        // pipeline stage 1: claim the item after next, start loading the next document's id
        const int item_n = __shfl_sync(FULL, item_next_raw, 0);
        item_next_raw = claim();
        const bool have_n = item_n + sub < P.n_items;
        const int d_n = have_n ? __ldg(P.order + item_n + sub) : 0;

        // rows of the first R tokens go in flight before the per-document setup
        for (int i = 0; i < R; i++) {
            int w = __shfl_sync(FULL, wcur, i, G);
            if ((unsigned)w >= (unsigned)P.V) w = 0;
            if (gl == 0 && i < len) tma_row_load(ring_u32 + (uint32_t)i * KS * 4u, P.nwk + (size_t)w * P.Kp, row_bytes, mbar_u32 + 8u * i);
        }
        doc_setup<KS, G, MULTI, Q1>(P, c, d, len, gl, nullptr, true);
        if constexpr (DIRECT) {     // row of the document's first token (after the setup: JG 128-bit registers are not worth carrying through it)
            const int w0 = row_word(__shfl_sync(FULL, wcur, 0, G));
            if (len > 0) load_row_regs<JG, G>(row, P.nwk + (size_t)w0 * P.Kp, gl);
        }
        // pipeline stage 2 (d_n has arrived during the setup): next document's extent and first tokens
        const long long b_n = P.doc_off[m][d_n];
        const int len_n = have_n ? (int)(P.doc_off[m][d_n + 1] - b_n) : 0;
        int wcur_n = 0, zcur_n = -1, wnext_n = 0, wahead_n = 0;    // loaded after the first token block (stage 3)

        int slot = 0;
        for (int base = 0; base < maxlen; base += G) {
            const int nblk = min(max(len - base, 0), G);            // tokens of my document in this block
            const int nblk_max = min(maxlen - base, G);             // uniform trip count of the warp
            // one Philox call per lane covers the G tokens of the block
            const uint4 rnd = philox4x32_10((uint32_t)(base + gl), gdoc, P.iteration, ((uint32_t)m << 8) | PURPOSE_SAMPLE, P.seed_lo, P.seed_hi);
            const float umine = (float)(rnd.x >> 8) * (1.0f / 16777216.0f);
            // zeff: topic (>= 0), -1 = UNASSIGNED_TOPIC, -2 = out-of-vocabulary word: the token is skipped and neither
            // leaves nor joins n_d (W:427-428)
            const int zeff = ((unsigned)wcur < (unsigned)P.V) ? zcur : -2;
            int znew = zcur;
            // first token of the NEXT block: it leaves its topic (W:434-471) paired with this block's last increment
            // (the document's very first token was left out of n_d by doc_setup)
            const int znext = (base + G < len) ? zmv[b + base + G] : -1;
            const int wnext0 = __shfl_sync(FULL, wnext, 0, G);
            const int ot_nextblk = ((unsigned)wnext0 < (unsigned)P.V && base + G < len) ? znext : -1;
            int ot_carry = __shfl_sync(FULL, zeff, 0, G);
            {   // tokens this block resamples (counted once per block, not once per token)
                const unsigned live = __ballot_sync(FULL, gl < nblk && zeff != -2);
                n_tok += (unsigned)__popc((live >> (lane - gl)) & ((G == 32) ? FULL : ((1u << (G & 31)) - 1u)));
            }
            for (int i = 0; i < nblk_max; i++) {
                const bool act = (NSUB == 1) || (i < nblk);            // one document per warp: the trip count is its own
                const int w = __shfl_sync(FULL, wcur, i, G);
                const int ot = ot_carry;                                       // = zeff of token i (shuffled one step earlier)
                const float u = __shfl_sync(FULL, umine, i, G);
                int otn = __shfl_sync(FULL, zeff, (i + 1) & (G - 1), G);
                ot_carry = otn;
                if (i + 1 == G) otn = ot_nextblk;
                int wa = __shfl_sync(FULL, wahead, i, G);                      // word of token base+i+R (ring refill)
                const bool valid = act && (ot != -2);
                int nt;
                if constexpr (DIRECT) {
                    nt = group_select_regs<JG, G>(row, q_gl_sa, lane, gl, P.beta, u, c.C);
                    // row consumed: fetch the next token's (wahead holds in-vocabulary ids only: see row_word)
                    if (act && base + i + 1 < len) load_row_regs<JG, G>(row, P.nwk + (size_t)wa * P.Kp, gl);
                } else {
                    if (act) {
                        mbar_wait(mbar_u32 + 8u * slot, (phasebits >> slot) & 1u);
                        phasebits ^= 1u << slot;
                    }
                    __syncwarp();
                    nt = group_select<JG, G>(q_gl_sa + doc_off_ring(KS, MULTI) + (uint32_t)slot * (KS * 4u), q_gl_sa, lane, gl, P.beta, u, c.C);
                }
                if (valid) { if (nt < 0) { nt = P.first_inactive; n_new++; } }   // W:522-526
                else nt = ot;
                if (!DIRECT) __syncwarp();
                // slot consumed: refill it with the row of token base+i+R
                if (!DIRECT && act && gl == 0 && base + i + R < len) {       // (wahead holds in-vocabulary ids only: see row_word)
                    tma_row_load(ring_u32 + (uint32_t)slot * KS * 4u, P.nwk + (size_t)wa * P.Kp, row_bytes, mbar_u32 + 8u * slot);
                }
                // this token joins its new topic (W:557-560) while the next token of the block leaves its old one
                apply_pair<KS, G, MULTI, Q1>(P, c, valid ? nt : -1, (act && (i + 1 < nblk || i + 1 == G) && otn >= 0) ? otn : -1, gl);
                if (valid) {
                    if (nt != ot && P.update_global) {                       // U:197-218
                        const int tsel = (gl & 1) ? ot : nt, v = (gl & 1) ? -1 : 1;
                        if (tsel >= 0) {
                            if (gl < 2) atomicAdd(P.nwk + (size_t)w * P.Kp + tsel, v);
                            else if (gl < 4) reds_add(dnk_sa + 4u * (uint32_t)tsel, v);
                        }
                    }
                    n_changed += (nt != ot);
                    if (gl == i) znew = nt;
                }
                // (DIRECT: nothing in the token loop passes through shared memory between lanes -- q, n_d and the row are private to
                // their owner lane -- and the full-mask shuffles at the top of the next token reconverge the warp)
                if (!DIRECT) __syncwarp();
                if (!DIRECT && act) slot = (slot + 1 == R) ? 0 : slot + 1;
            }
            if (gl < nblk) {
                zmv[b + base + gl] = znew;
                if (P.z_host) P.z_host[b + base + gl] = znew;       // posted write over PCIe, one 4*G-byte run per block
            }
            wcur = wnext;
            wahead = (base + G + RA + gl < len) ? row_word(__ldg(P.word + b + base + G + RA + gl)) : 0;
            zcur = (base + G + gl < len) ? zmv[b + base + G + gl] : -1;
            wnext = (base + 2 * G + gl < len) ? __ldg(P.word + b + base + 2 * G + gl) : 0;
            if (base == 0) {   // pipeline stage 3: the next document's first tokens (its offsets arrived long ago)
                wcur_n = (gl < len_n) ? __ldg(P.word + b_n + gl) : 0;
                zcur_n = (gl < len_n) ? zmv[b_n + gl] : -1;
                wnext_n = (G + gl < len_n) ? __ldg(P.word + b_n + G + gl) : 0;
                wahead_n = (RA + gl < len_n) ? row_word(__ldg(P.word + b_n + RA + gl)) : 0;
            }
        }
        if (maxlen == 0) {     // (cannot happen: the work list holds non-empty documents only; keeps the pipeline total)
            wcur_n = (gl < len_n) ? __ldg(P.word + b_n + gl) : 0;
            zcur_n = (gl < len_n) ? zmv[b_n + gl] : -1;
            wnext_n = (G + gl < len_n) ? __ldg(P.word + b_n + G + gl) : 0;
            wahead_n = (RA + gl < len_n) ? row_word(__ldg(P.word + b_n + RA + gl)) : 0;
        }
        item0 = item_n; d = d_n; b = b_n; len = len_n;
        wcur = wcur_n; zcur = zcur_n; wnext = wnext_n; wahead = wahead_n;
    }
    if (gl == 0) {
        if (n_tok) atomicAdd(P.stats + 0, (unsigned long long)n_tok);
        if (n_changed) atomicAdd(P.stats + 1, (unsigned long long)n_changed);
        if (n_new) atomicAdd(P.stats + 2, (unsigned long long)n_new);
    }
    __syncthreads();
    if (P.update_global) {
        const int *dnk = reinterpret_cast<const int *>(smem + (size_t)KS * 8);
        for (int t = threadIdx.x; t < P.K; t += blockDim.x) { int v = dnk[t]; if (v) atomicAdd(P.nk_live + t, v); }
    }
}

// entry points: the TMA-ring kernel is bounded by its launch shape (the shared-memory budget decides the warps), the DIRECT kernel
// by an explicit register count (ptxas rounds a launch bound's thread count up to a multiple of 128 before it derives the
// register cap, which would leave 13..16-warp shapes with 128 registers and spills in the token loop)
template <int KS, int G, bool MULTI, bool Q1>
__global__ void __launch_bounds__(sweep_max_threads(KS, G, MULTI), 1) k_sweep_view(const SweepParams P)
{ sweep_view_body<KS, G, MULTI, Q1, false>(P); }

template <int KS, int G, bool MULTI>
__global__ void __maxnreg__(direct_regs(KS, G)) k_sweep_view_direct(const SweepParams P)
{ sweep_view_body<KS, G, MULTI, false, true>(P); }

// ------------------------------------------------------------------------------------------------
// parity probe: conditional of one token on frozen counts, same device functions as the sweep.  One warp; every lane
// group works on the same document in its own shared-memory area, group 0 writes the result.
// ------------------------------------------------------------------------------------------------
template <int KS, int G, bool MULTI, bool Q1>
__global__ void __launch_bounds__(32, 1) k_cond_probe(const SweepParams P, int d, int pos, const double *p_row, double *out)
{
    constexpr int JG = KS / (4 * G);
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x, sub = lane / G, gl = lane % G;
    float2 *ginv = reinterpret_cast<float2 *>(smem);
    for (int t = lane; t < KS; t += 32) {
        float2 g = make_float2(0.f, 0.f);
        if (t < P.K) { g.x = P.ga_tree[t]; g.y = 1.0f / ((float)P.nk_frozen[t] + P.betaSum); }
        ginv[t] = g;
    }
    if (MULTI) {
        float *gaf = reinterpret_cast<float *>(smem + (size_t)KS * 12);
        for (int i = 0; i < P.M; i++)
            for (int t = lane; t < KS; t += 32) gaf[(size_t)i * KS + t] = (t < P.K) ? P.ga_full[i][t] : 0.f;
    }
    DocCtx c; int *ring; unsigned long long *mbar;
    carve_doc(smem + smem_cta_bytes(KS, MULTI ? P.M : 0) + (size_t)sub * smem_doc_bytes(KS, 1, MULTI), KS, 1, MULTI, c, ring, mbar);
    c.gaf = reinterpret_cast<const float *>(smem + (size_t)KS * 12); c.gaf_stride = KS;
    c.ginv_sa = smem_u32(smem);
    if (MULTI) c.oc = P.oc_scratch + (size_t)sub * P.Kp;
    __syncwarp();
    const long long b = P.doc_off[P.m][d];
    const int len = (int)(P.doc_off[P.m][d + 1] - b);
    doc_setup<KS, G, MULTI, Q1>(P, c, d, len, gl, p_row, false);
    const int w = P.word[b + pos], ot = P.zv[P.m][b + pos];
    apply_pair<KS, G, MULTI, Q1>(P, c, -1, ot, gl);
    for (int t = gl; t < KS; t += G) ring[t] = (t < P.Kp) ? P.nwk[(size_t)w * P.Kp + t] : 0;
    __syncwarp();
    float cum[JG];
    float lt = lane_weights<JG, G>(smem_u32(ring) + 16u * (uint32_t)gl, c.sa + 16u * (uint32_t)gl, P.beta, cum);
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) lt += __shfl_xor_sync(FULL, lt, off, G);
    const float total = lt + c.C;
    if (sub == 0) {
        for (int t = gl; t < P.K; t += G) {
            float wgt = topic_weight(ring[t], c.q[t], P.beta);
            if (c.C > 0.f && t == P.first_inactive) wgt += c.C;
            out[t] = (double)(wgt / total);
        }
        if (gl == 0) out[P.K] = (double)(c.C / total);
    }
}

// ------------------------------------------------------------------------------------------------
// initialisation, counts, histogram, log-likelihood, invariants
// ------------------------------------------------------------------------------------------------
// random initialisation M:465-515 (previousModel == null): view 0 uniform over K, view m>0 uniform over the
// document's view-0 draws (uniform over K when the document has no view-0 tokens).  One thread per token;
// view 0 must be initialised first.
__global__ void k_init_assign(int m, int K, long long n_docs, const long long *off_m, const long long *off_0, const int *z0, int *z,
                              unsigned seed_lo, unsigned seed_hi, long long doc_id_base, long long doc_id_stride)
{
    // one warp per document keeps the doc id lookup trivial
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long d = wid; d < n_docs; d += nw) {
        const long long b = off_m[d]; const int len = (int)(off_m[d + 1] - b);
        const long long b0 = off_0[d]; const int len0 = (int)(off_0[d + 1] - b0);
        const uint32_t gdoc = (uint32_t)(doc_id_base + d * doc_id_stride);
        for (int i = lane; i < len; i += 32) {
            uint4 x = philox4x32_10((uint32_t)i, gdoc, 0u, ((uint32_t)m << 8) | PURPOSE_INIT, seed_lo, seed_hi);
            int t;
            if (m == 0 || len0 == 0) t = (int)__umulhi(x.x, (uint32_t)K);
            else t = z0[b0 + (long long)__umulhi(x.x, (uint32_t)len0)];
            z[b + i] = t;
        }
    }
}

// inferencer initialisation I:186-203: every in-vocabulary token draws its topic from the bare topic-word distribution
// phi_t = (n_wk[w][t] + beta) / (n_k[t] + betaSum) of the TRAINED counts (I:561-576 builds the trees without gamma*alpha);
// out-of-vocabulary tokens keep Java's default 0 (Q13).  FTree.sample (FT:111-136) on the heap-shaped tree walks the leaves
// in the order "topics rot..K-1, then 0..rot-1" with rot = 2^floor(log2(2K-1)) - K, which this linear scan reproduces.
__global__ void k_init_from_phi(int m, int K, int Kp, int V, long long n_docs, const long long *off, const int *word, int *z, const int *nwk,
                                const int *nk, double beta, double betaSum, int rot, unsigned seed_lo, unsigned seed_hi,
                                long long doc_id_base, long long doc_id_stride)
{
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long d = wid; d < n_docs; d += nw) {
        const long long b = off[d]; const int len = (int)(off[d + 1] - b);
        const uint32_t gdoc = (uint32_t)(doc_id_base + d * doc_id_stride);
        for (int i = lane; i < len; i += 32) {
            const int w = word[b + i];
            if ((unsigned)w >= (unsigned)V) { z[b + i] = 0; continue; }
            const int *row = nwk + (size_t)w * Kp;
            double total = 0.0;
            for (int t = 0; t < K; t++) total += ((double)row[t] + beta) / ((double)nk[t] + betaSum);
            uint4 x = philox4x32_10((uint32_t)i, gdoc, 0u, ((uint32_t)m << 8) | PURPOSE_INIT, seed_lo, seed_hi);
            const double target = (double)(x.x >> 8) * (1.0 / 16777216.0) * total;
            double cum = 0.0; int pick = -1, last = 0;
            for (int k = 0; k < K; k++) {
                int t = k + rot; if (t >= K) t -= K;
                cum += ((double)row[t] + beta) / ((double)nk[t] + betaSum);
                last = t;
                if (pick < 0 && target < cum) pick = t;
            }
            z[b + i] = pick < 0 ? last : pick;
        }
    }
}

// mvtm_sweep_host_dist: does the caller's array still hold what the device holds?  *ndiff receives the number of blocks that saw
// a difference (0 = identical)
__global__ void k_diff_assign(long long n, const int *a, const int *b, int *ndiff)
{
    int d = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) d |= (a[i] != b[i]);
    if (__syncthreads_or(d) && threadIdx.x == 0) atomicAdd(ndiff, 1);
}

// buildInitialTypeTopicCounts M:600-652: n_wk / n_k from (word, z); n_k through a shared-memory histogram
// fix != 0: an id >= K is rewritten to UNASSIGNED_TOPIC so that no sweep can index with it; fix == 0 (mvtm_check_invariants): the
// assignments are only read
__global__ void k_build_counts(long long n_tok, const int *word, int *z, int V, int K, int Kp, int *nwk, int *nk, int *bad, int fix)
{
    extern __shared__ int hk[];
    for (int t = threadIdx.x; t < K; t += blockDim.x) hk[t] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_tok; i += (long long)gridDim.x * blockDim.x) {
        int t = z[i], w = word[i];
        if (t < 0) continue;                                      // UNASSIGNED_TOPIC, M:634
        if (t >= K) { atomicAdd(bad, 1); if (fix) z[i] = -1; continue; }   // reported by the caller; never left where a sweep could index with it
        atomicAdd(hk + t, 1);
        if ((unsigned)w < (unsigned)V) atomicAdd(nwk + (size_t)w * Kp + t, 1);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < K; t += blockDim.x) { int v = hk[t]; if (v) atomicAdd(nk + t, v); }
}

// topicDocCounts (U:220-232 / M:647-649) recomputed: one warp per document, bins c >= 1
__global__ void k_doc_topic_hist(long long n_docs, const long long *off, const int *z, int K, int stride, int *hist)
{
    extern __shared__ int sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int *cnt = sm + (size_t)warp * K;
    for (int t = lane; t < K; t += 32) cnt[t] = 0;
    __syncwarp();
    for (long long d = (long long)blockIdx.x * nwarp + warp; d < n_docs; d += (long long)gridDim.x * nwarp) {
        const long long b = off[d]; const int len = (int)(off[d + 1] - b);
        if (len == 0) continue;
        for (int i = lane; i < len; i += 32) { int t = z[b + i]; if (t >= 0) atomicAdd(cnt + t, 1); }
        __syncwarp();
        for (int t = lane; t < K; t += 32) { int cv = cnt[t]; if (cv) { atomicAdd(hist + (size_t)t * stride + cv, 1); cnt[t] = 0; } }
        __syncwarp();
    }
}

__device__ __forceinline__ double log_gamma_stirling(double z)
{   // cc.mallet.types.Dirichlet.logGammaStirling
    int shift = 0;
    while (z < 2.0) { z += 1.0; shift++; }
    double r = 0.91893853320467274178 + (z - 0.5) * log(z) - z + 1.0 / (12.0 * z) - 1.0 / (360.0 * z * z * z) + 1.0 / (1260.0 * z * z * z * z * z);
    while (shift-- > 0) { z -= 1.0; r -= log(z); }
    return r;
}

// document part of modelLogLikelihood M:3341-3370: per-document term written to doc_ll[d] (0 for skipped docs),
// counted[d] = 1 when the document contributed (modalityCnt, M:3366)
__global__ void k_loglik_docs(long long n_docs, const long long *off, const int *z, const unsigned char *present, int K,
                              const double *ga, const double *tlg, double gas, int quirk_len2, double *doc_ll, unsigned char *counted)
{
    extern __shared__ int sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int *cnt = sm + (size_t)warp * K;
    for (int t = lane; t < K; t += 32) cnt[t] = 0;
    __syncwarp();
    for (long long d = (long long)blockIdx.x * nwarp + warp; d < n_docs; d += (long long)gridDim.x * nwarp) {
        const long long b = off[d]; const int len = (int)(off[d + 1] - b);
        const int arrlen = quirk_len2 ? (len < 2 ? 2 : len) : len;        // Q18
        if (!present[d] || arrlen == 0) { if (lane == 0) { doc_ll[d] = 0.0; counted[d] = 0; } continue; }
        // UNASSIGNED_TOPIC (-1: a fresh handle, an out-of-range id k_build_counts rejected, a restored state) counts for nothing
        for (int i = lane; i < len; i += 32) { const int t = z[b + i]; if (t >= 0 && t < K) atomicAdd(cnt + t, 1); }
        if (lane == 0 && arrlen > len) atomicAdd(cnt + 0, arrlen - len);
        __syncwarp();
        double acc = 0.0;
        for (int t = lane; t < K; t += 32) {
            int cv = cnt[t];
            if (cv > 0) { acc += log_gamma_stirling(ga[t] + (double)cv) - tlg[t]; cnt[t] = 0; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) { doc_ll[d] = acc - log_gamma_stirling(gas + (double)arrlen); counted[d] = 1; }
        __syncwarp();
    }
}

// held-out scoring by document completion (mvtm_heldout_loglik): one warp per document; n_d from the handle's own tokens of the
// view (the observed part, at their current assignments), every evaluation token w scored as
//   log sum_t (n_wk[w][t] + beta) / (n_k[t] + betaSum) * (n_d[t] + ga[t]) / (N_obs + sum_t ga[t])          (fp64)
__global__ void k_heldout_docs(long long n_docs, const long long *off, const int *z, const long long *eoff, const int *eword, int V, int K,
                               int Kp, const int *nwk, const int *nk, const double *ga, double ga_sum, double beta, double betaSum,
                               double *doc_ll, int *doc_n)
{
    extern __shared__ int sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int *cnt = sm + (size_t)warp * K;
    for (int t = lane; t < K; t += 32) cnt[t] = 0;
    __syncwarp();
    for (long long d = (long long)blockIdx.x * nwarp + warp; d < n_docs; d += (long long)gridDim.x * nwarp) {
        const long long b = off[d]; const int len = (int)(off[d + 1] - b);
        const long long eb = eoff[d]; const int elen = (int)(eoff[d + 1] - eb);
        if (elen == 0) { if (lane == 0) { doc_ll[d] = 0.0; doc_n[d] = 0; } continue; }
        int nobs = 0;
        for (int i = lane; i < len; i += 32) { int t = z[b + i]; if (t >= 0 && t < K) { atomicAdd(cnt + t, 1); nobs++; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nobs += __shfl_xor_sync(0xffffffffu, nobs, o);
        __syncwarp();
        const double denom = (double)nobs + ga_sum;
        double ll = 0.0; int n = 0;
        for (int i = 0; i < elen; i++) {
            const int w = eword[eb + i];
            if ((unsigned)w >= (unsigned)V) continue;                     // out-of-vocabulary: not scored (W:427-428 skips it too)
            const int *row = nwk + (size_t)w * Kp;
            double acc = 0.0;
            for (int t = lane; t < K; t += 32)
                acc += ((double)row[t] + beta) / ((double)nk[t] + betaSum) * ((double)cnt[t] + ga[t]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            ll += log(acc / denom); n++;
        }
        if (lane == 0) { doc_ll[d] = ll; doc_n[d] = n; }
        __syncwarp();
        for (int i = lane; i < len; i += 32) { int t = z[b + i]; if (t >= 0 && t < K) cnt[t] = 0; }
        __syncwarp();
    }
}

// topic-word part M:3389-3415: per-block partial sums of lgamma(beta + n) over cells with n > 0 and their count
__global__ void k_loglik_cells(int V, int K, int Kp, const int *nwk, double beta, double *part_sum, long long *part_nnz)
{
    __shared__ double ssum[32]; __shared__ long long snnz[32];
    double acc = 0.0; long long nnz = 0;
    const long long n = (long long)V * Kp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int t = (int)(i % Kp);
        int cv = nwk[i];
        if (t < K && cv > 0) { nnz++; acc += log_gamma_stirling(beta + (double)cv); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); nnz += __shfl_xor_sync(0xffffffffu, nnz, o); }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { ssum[warp] = acc; snnz[warp] = nnz; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0; long long c = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) { a += ssum[k]; c += snnz[k]; }
        part_sum[blockIdx.x] = a; part_nnz[blockIdx.x] = c;
    }
}

// invariants: scratch holds a recount of n_wk; compare cell by cell, flag negatives
__global__ void k_compare_counts(long long n, const int *a, const int *b, unsigned long long *bad)
{
    unsigned long long local = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        local += (a[i] != b[i]) || (a[i] < 0);
    if (local) atomicAdd(bad, local);
}

// multi-GPU delta plumbing: elementwise a -= b / a += b, and snapshot advance
__global__ void k_sub_inplace(long long n, int *a, const int *b)
{ for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] -= b[i]; }
__global__ void k_add_snapshot(long long n, int *a, int *snap)
{   // a holds the reduced delta: a = snap + a, snap = a
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) { int v = a[i] + snap[i]; a[i] = v; snap[i] = v; }
}
// Sum-form exchange: every rank entered the sweep with the same global table G (= snap) and now holds G + delta_r; after
// an in-place all-reduce a = N*G + sum_r delta_r, so the new global table is a - (N-1)*G.  One fused pass, no export pass.
__global__ void k_finish_sum_exchange4(long long n4, int4 *a, int4 *snap, int world_minus_1)
{   // 128-bit form for the overlapped exchange (row stride is a multiple of 32 ints, so the table is int4-divisible)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        int4 x = a[i]; const int4 g = snap[i];
        x.x -= world_minus_1 * g.x; x.y -= world_minus_1 * g.y; x.z -= world_minus_1 * g.z; x.w -= world_minus_1 * g.w;
        a[i] = x; snap[i] = x;
    }
}
__global__ void k_finish_sum_exchange(long long n, int *a, int *snap, int world_minus_1)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int v = a[i] - world_minus_1 * snap[i]; a[i] = v; snap[i] = v;
    }
}
