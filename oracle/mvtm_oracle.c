/*
 * mvtm_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the collapsed-Gibbs hot path of hmetaxa/MVTopicModel, used only as the
 * checker for the CUDA engine (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl
 * reference legs).  Nothing under mvtopicmodel_b200/ may link, import or execute this file.
 *
 * PARITY STATUS: pinned to outputs of the reference's own binaries.  The reference ships no tests or golden
 * vectors and its trainer cannot be launched here (no JVM in this image), but
 * tools/jvm_mini.py -- a JVM bytecode interpreter -- EXECUTES the reference's classes from the jars it ships:
 *   - FastQMVWVWorkerRunnable.sampleTopicsForOneDoc (the sampler, W:301-597) on five corpora; this file's
 *     reference-faithful mode reproduces its assignments token for token, sweep after sweep
 *     (tests/golden/reference_sampler_vectors.json, tests/test_reference_vectors.py);
 *   - FTree (build / sample / update), lower_bound, MALLET's logGammaStirling -- bit for bit
 *     (tests/golden/reference_vectors.json).
 * Not executed: the trainer's outer loop, modelLogLikelihood and the optimisers as a whole; those rest on
 * the citations below and the hand-derived known answers of SURVEY.md section 8(c) (tests/test_oracle.py).
 *
 * Reference citations use these tags (all under /root/reference/src/main/java/org/madgik/):
 *   W  = MVTopicModel/FastQMVWVWorkerRunnable.java
 *   U  = MVTopicModel/FastQMVWVUpdaterRunnable.java
 *   M  = MVTopicModel/FastQMVWVParallelTopicModel.java
 *   FT = utils/FTree.java            QD = utils/FastQDelta.java
 * MALLET 2.0.8 arithmetic (binary-only dependency, pom.xml:49-53) is restated from its published
 * algorithm as recovered in SURVEY.md section 8(c).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define ORC_MAXM 8
#define ORC_UNASSIGNED (-1)

/* sweep flags */
#define ORC_F_Q1_COMPAT     1u   /* W:560-584: never insert a newly gained topic into S during a sweep   */
#define ORC_F_STALE_TREES   2u   /* Q3: B bucket from the incrementally maintained trees (U:242-260)     */
#define ORC_F_DEFERRED      4u   /* apply all deltas at the end of the sweep instead of immediately      */
#define ORC_F_BETA_MALLET   8u   /* Q5: MALLET Randoms.nextBeta law instead of true Beta(a,1)=u^(1/a)    */
#define ORC_F_CHECK_RULE 256u     /* reference-faithful mode + Q1: also run the engine's flag rule for the dense index and count
                                     the tokens at which it disagrees with the reference's own list S (must stay 0)          */
#define ORC_F_ENGINE_MIRROR 16u  /* view-major order, dense single-scan sampler in the engine's order     */
#define ORC_F_DOC_ORDER     32u  /* (engine mirror) plain document order, no length sort                 */
#define ORC_F_FROZEN        64u  /* global counts frozen: the inferencer's nut = 0 mode, I:211-256 */
#define ORC_F_BARE_TREES    128u /* Q13: the inferencer's trees hold phi without gamma*alpha and ignore the inactive set (I:561-576) */

typedef struct {
    int M, K;
    int64_t D;
    int V[ORC_MAXM];
    int64_t *doc_off[ORC_MAXM];
    int32_t *word[ORC_MAXM];
    int32_t *z[ORC_MAXM];
    uint8_t *present[ORC_MAXM];      /* doc has an Assignments[m] object (may hold 0 tokens), MA:13-19 */
    int32_t *n_wk[ORC_MAXM];         /* typeTopicCounts[m][w][t], row-major by word, stride K          */
    int32_t *n_k[ORC_MAXM];          /* tokensPerTopic[m][t]                                            */
    int32_t *type_total[ORC_MAXM];   /* typeTotals[m][w], M:511                                         */
    double *tree[ORC_MAXM];          /* FTree per word: V*2K doubles (allocated lazily)                 */
    int32_t *hist[ORC_MAXM];         /* topicDocCounts[m][t][c], K*(maxlen+1)                           */
    int32_t *doclen_cnt[ORC_MAXM];   /* docLengthCounts[m][len], M:626                                  */
    int maxlen[ORC_MAXM];
    int64_t total_tokens[ORC_MAXM];
    int64_t docs_per_view[ORC_MAXM];
    double *alpha[ORC_MAXM];         /* K+1, slot K = new-topic prior */
    double alphaSum[ORC_MAXM], beta[ORC_MAXM], betaSum[ORC_MAXM], gamma[ORC_MAXM];
    double p_a[ORC_MAXM][ORC_MAXM], p_b[ORC_MAXM][ORC_MAXM];
    int n_inactive;
    int32_t *inactive;               /* ascending topic ids */
    uint64_t seed;
    int engine_G;                    /* lanes per document-view of the engine's scan (mvtm_scan_layout) */
    int64_t cnt_rule_bad;            /* ORC_F_CHECK_RULE: tokens at which the flag rule disagreed with the reference's S */
    uint8_t *rflag;                  /* engine mirror with Q1: D x K flags "topic is NOT in the dense index of this document for
                                        the rest of the sweep" (left it, or was gained while absent: W:441-468 removes, W:563-584
                                        never inserts); kept across the view passes of one sweep */
    int64_t doc_base, doc_stride;    /* global id of local document d = doc_base + d*doc_stride (keys the RNG) */
    int64_t cnt_new, cnt_doc, cnt_tree, cnt_changed;   /* W:33-35 bucket counters */
    char err[256];
} orc_t;

/* ------------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011, Random123).  The engine keys its draws the same way so that    */
/* oracle and engine consume the identical uniform for every (iteration, doc, view, token).          */
/* ------------------------------------------------------------------------------------------------ */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { ORC_PURPOSE_SAMPLE = 0, ORC_PURPOSE_INIT = 1, ORC_PURPOSE_PDRAW = 2 };

static void orc_draw(const orc_t *o, uint32_t pos, uint32_t doc, uint32_t iteration, uint32_t view_or_pair,
                     uint32_t purpose, uint32_t out[4])
{
    uint32_t ctr[4] = { pos, doc, iteration, (view_or_pair << 8) | purpose };
    uint32_t key[2] = { (uint32_t)o->seed, (uint32_t)(o->seed >> 32) };
    orc_philox4x32(ctr, key, out);
}
static inline double u24(uint32_t x) { return (double)(x >> 8) * (1.0 / 16777216.0); }

/* ------------------------------------------------------------------------------------------------ */
/* FTree (FT:96-147) on a double[2K] array                                                           */
/* ------------------------------------------------------------------------------------------------ */
void orc_ftree_build(double *tree, const double *w, int K)
{   /* FT:96-109 */
    tree[0] = 0.0;
    for (int i = 2 * K - 1; i > 0; --i) tree[i] = (i >= K) ? w[i - K] : tree[2 * i] + tree[2 * i + 1];
}
int orc_ftree_sample(const double *tree, int K, double u)
{   /* FT:111-136 */
    int i = 1;
    u = u * tree[i];
    while (i < K) {
        if (u < tree[2 * i]) i = 2 * i;
        else { u -= tree[2 * i]; i = 2 * i + 1; }
    }
    return i - K;
}
void orc_ftree_update(double *tree, int K, int t, double v)
{   /* FT:138-147 */
    int i = t + K;
    double d = v - tree[i];
    while (i > 0) { tree[i] += d; i /= 2; }
}
int orc_lower_bound(const double *arr, double key, int len)
{   /* W:257-277, literal control flow (Java int division truncates toward zero) */
    int lo = 0, hi = len - 1, mid = (lo + hi) / 2;
    for (;;) {
        if (arr[mid] >= key) { hi = mid - 1; if (hi < lo) return mid; }
        else { lo = mid + 1; if (hi < lo) return mid < len - 1 ? mid + 1 : -1; }
        mid = (lo + hi) / 2;
    }
}
double orc_log_gamma_stirling(double z)
{   /* cc.mallet.types.Dirichlet.logGammaStirling, SURVEY 8(c) */
    int shift = 0;
    while (z < 2) { z++; shift++; }
    double r = 0.5 * log(2.0 * M_PI) + (z - 0.5) * log(z) - z + 1.0 / (12.0 * z) - 1.0 / (360.0 * z * z * z)
             + 1.0 / (1260.0 * z * z * z * z * z);
    while (shift-- > 0) { z--; r -= log(z); }
    return r;
}

/* ------------------------------------------------------------------------------------------------ */
/* small deterministic generator for the MALLET-law Beta draw (Q5); seeded from Philox words          */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { uint64_t s; } orc_rng;
static inline double rng_u(orc_rng *r)
{   /* splitmix64 -> 53-bit uniform */
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
static double rng_gauss(orc_rng *r)
{   double u1, u2; do { u1 = rng_u(r); } while (u1 <= 0.0); u2 = rng_u(r);
    return sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2); }

double orc_next_beta_mallet(uint64_t seed, double a, double b)
{   /* cc.mallet.util.Randoms.nextBeta, SURVEY 8(c) Q5 */
    orc_rng r = { seed };
    if (a == 1 && b == 1) return rng_u(&r);
    if (a >= 1 && b >= 1) {
        double A = a - 1, B = b - 1, C = A + B, L = C * log(C), mu = A / C, sigma = 0.5 / sqrt(C);
        double y = rng_gauss(&r), x = sigma * y + mu;
        while (x < 0 || x > 1) { y = rng_gauss(&r); x = sigma * y + mu; }
        double u = rng_u(&r);
        /* with b == 1: B*log((1-x)/B) = 0*log(inf) = NaN -> comparison false -> first proposal accepted */
        while (log(u) >= A * log(x / A) + B * log((1 - x) / B) + L + 0.5 * y * y) {
            y = rng_gauss(&r); x = sigma * y + mu;
            while (x < 0 || x > 1) { y = rng_gauss(&r); x = sigma * y + mu; }
            u = rng_u(&r);
        }
        return x;
    }
    double v1, v2;
    do { v1 = pow(rng_u(&r), 1.0 / a); v2 = pow(rng_u(&r), 1.0 / b); } while (v1 + v2 > 1);
    return v1 / (v1 + v2);
}

/* ------------------------------------------------------------------------------------------------ */
/* lifecycle                                                                                         */
/* ------------------------------------------------------------------------------------------------ */
orc_t *orc_create(int M, int K, int64_t D, const int32_t *V, uint64_t seed)
{
    if (M < 1 || M > ORC_MAXM || K < 1 || D < 0) return NULL;
    orc_t *o = (orc_t *)calloc(1, sizeof(orc_t));
    o->M = M; o->K = K; o->D = D; o->seed = seed; o->doc_base = 0; o->doc_stride = 1; o->engine_G = 32;
    o->inactive = (int32_t *)calloc((size_t)K, sizeof(int32_t));
    for (int m = 0; m < M; m++) {
        o->V[m] = V[m];
        o->alpha[m] = (double *)calloc((size_t)K + 1, sizeof(double));
        /* ctor defaults, M:195-239 and S:149-159 */
        for (int t = 0; t <= K; t++) o->alpha[m][t] = 0.1;
        o->alphaSum[m] = 0.1 * K; o->beta[m] = 0.01; o->betaSum[m] = 0.01 * V[m]; o->gamma[m] = 1.0;
        for (int j = 0; j < M; j++) { o->p_a[m][j] = 0.2; o->p_b[m][j] = 1.0; }   /* M:1055-1058 */
    }
    return o;
}
void orc_destroy(orc_t *o)
{
    if (!o) return;
    for (int m = 0; m < o->M; m++) {
        free(o->doc_off[m]); free(o->word[m]); free(o->z[m]); free(o->present[m]); free(o->n_wk[m]);
        free(o->n_k[m]); free(o->type_total[m]); free(o->tree[m]); free(o->hist[m]); free(o->doclen_cnt[m]);
        free(o->alpha[m]);
    }
    free(o->inactive); free(o->rflag); free(o);
}
int orc_add_view(orc_t *o, int m, const int64_t *doc_off, const int32_t *word, const uint8_t *present)
{   /* MA:13-19 / M:410-463 restated as a doc-aligned CSR per view; absent view == empty range */
    if (m < 0 || m >= o->M) return 1;
    int64_t D = o->D, N = doc_off[D];
    o->doc_off[m] = (int64_t *)malloc((size_t)(D + 1) * 8); memcpy(o->doc_off[m], doc_off, (size_t)(D + 1) * 8);
    o->word[m] = (int32_t *)malloc((size_t)(N > 0 ? N : 1) * 4); memcpy(o->word[m], word, (size_t)N * 4);
    o->z[m] = (int32_t *)malloc((size_t)(N > 0 ? N : 1) * 4);
    for (int64_t i = 0; i < N; i++) o->z[m][i] = ORC_UNASSIGNED;
    o->present[m] = (uint8_t *)malloc((size_t)(D > 0 ? D : 1));
    int maxlen = 0; int64_t docs = 0;
    for (int64_t d = 0; d < D; d++) {
        int len = (int)(doc_off[d + 1] - doc_off[d]);
        if (len < 0) return 2;
        if (len > maxlen) maxlen = len;
        o->present[m][d] = present ? present[d] : (len > 0);
        docs += o->present[m][d];
    }
    o->maxlen[m] = maxlen; o->total_tokens[m] = N; o->docs_per_view[m] = docs;
    o->n_wk[m] = (int32_t *)calloc((size_t)o->V[m] * o->K, 4);
    o->n_k[m] = (int32_t *)calloc((size_t)o->K, 4);
    o->type_total[m] = (int32_t *)calloc((size_t)o->V[m], 4);
    o->hist[m] = (int32_t *)calloc((size_t)o->K * (maxlen + 1), 4);
    o->doclen_cnt[m] = (int32_t *)calloc((size_t)maxlen + 1, 4);
    return 0;
}
int orc_set_hyper(orc_t *o, const double *alpha, const double *alphaSum, const double *beta, const double *betaSum,
                  const double *gamma, const double *p_a, const double *p_b, const int32_t *inactive, int n_inactive)
{
    int M = o->M, K = o->K;
    for (int m = 0; m < M; m++) {
        if (alpha) memcpy(o->alpha[m], alpha + (size_t)m * (K + 1), (size_t)(K + 1) * 8);
        if (alphaSum) o->alphaSum[m] = alphaSum[m];
        if (beta) o->beta[m] = beta[m];
        if (betaSum) o->betaSum[m] = betaSum[m];
        if (gamma) o->gamma[m] = gamma[m];
        for (int j = 0; j < M; j++) {
            if (p_a) o->p_a[m][j] = p_a[m * M + j];
            if (p_b) o->p_b[m][j] = p_b[m * M + j];
        }
    }
    if (n_inactive >= 0) {
        o->n_inactive = n_inactive;
        for (int i = 0; i < n_inactive; i++) o->inactive[i] = inactive[i];
    }
    return 0;
}
static int is_inactive(const orc_t *o, int t)
{
    for (int i = 0; i < o->n_inactive; i++) if (o->inactive[i] == t) return 1;
    return 0;
}
static void remove_inactive(orc_t *o, int t)
{
    int j = 0;
    for (int i = 0; i < o->n_inactive; i++) if (o->inactive[i] != t) o->inactive[j++] = o->inactive[i];
    o->n_inactive = j;
}

/* tree leaves: M:2660-2691 */
static inline double leaf_value(const orc_t *o, int m, int w, int t)
{
    return o->gamma[m] * o->alpha[m][t] * ((o->n_wk[m][(size_t)w * o->K + t] + o->beta[m]) / (o->n_k[m][t] + o->betaSum[m]));
}
static inline double phi_value(const orc_t *o, int m, int w, int t)
{ return (o->n_wk[m][(size_t)w * o->K + t] + o->beta[m]) / (o->n_k[m][t] + o->betaSum[m]); }
static void build_tree_for_word_f(const orc_t *o, int m, int w, double *tree, double *tmp, unsigned flags)
{
    int K = o->K;
    if (flags & ORC_F_BARE_TREES) for (int t = 0; t < K; t++) tmp[t] = phi_value(o, m, w, t);          /* I:561-576 */
    else for (int t = 0; t < K; t++) tmp[t] = (o->n_inactive && is_inactive(o, t)) ? 0.0 : leaf_value(o, m, w, t);
    orc_ftree_build(tree, tmp, K);
}
static void build_tree_for_word(const orc_t *o, int m, int w, double *tree, double *tmp)
{ build_tree_for_word_f(o, m, w, tree, tmp, 0); }
int orc_rebuild_trees(orc_t *o)
{   /* buildFTrees, M:2660-2696 */
    int K = o->K;
    double *tmp = (double *)malloc((size_t)K * 8);
    for (int m = 0; m < o->M; m++) {
        if (!o->tree[m]) o->tree[m] = (double *)malloc((size_t)o->V[m] * 2 * K * 8);
        if (!o->tree[m]) { free(tmp); return 1; }
        for (int w = 0; w < o->V[m]; w++) build_tree_for_word(o, m, w, o->tree[m] + (size_t)w * 2 * K, tmp);
    }
    free(tmp);
    return 0;
}

int orc_rebuild_counts(orc_t *o)
{   /* buildInitialTypeTopicCounts M:600-652 (+ typeTotals M:511, initializeHistograms M:849-897) */
    int K = o->K;
    int32_t *local = (int32_t *)calloc((size_t)K, 4);
    for (int m = 0; m < o->M; m++) {
        memset(o->n_wk[m], 0, (size_t)o->V[m] * K * 4);
        memset(o->n_k[m], 0, (size_t)K * 4);
        memset(o->type_total[m], 0, (size_t)o->V[m] * 4);
        memset(o->hist[m], 0, (size_t)K * (o->maxlen[m] + 1) * 4);
        memset(o->doclen_cnt[m], 0, (size_t)(o->maxlen[m] + 1) * 4);
        for (int64_t d = 0; d < o->D; d++) {
            if (!o->present[m][d]) continue;
            int64_t b = o->doc_off[m][d], e = o->doc_off[m][d + 1];
            o->doclen_cnt[m][e - b]++;
            for (int64_t i = b; i < e; i++) {
                int w = o->word[m][i], t = o->z[m][i];
                if (w >= 0 && w < o->V[m]) o->type_total[m][w]++;
                if (t == ORC_UNASSIGNED) continue;
                local[t]++;
                o->n_k[m][t]++;
                if (w >= 0 && w < o->V[m]) o->n_wk[m][(size_t)w * K + t]++;
            }
            for (int t = 0; t < K; t++) { o->hist[m][(size_t)t * (o->maxlen[m] + 1) + local[t]]++; local[t] = 0; }
        }
    }
    free(local);
    return 0;
}

int orc_init_assignments(orc_t *o)
{   /* random initialisation, M:465-515 (previousModel == null path), Philox-keyed */
    int K = o->K;
    for (int64_t d = 0; d < o->D; d++) {
        int64_t b0 = o->doc_off[0][d]; int len0 = (int)(o->doc_off[0][d + 1] - b0);
        for (int m = 0; m < o->M; m++) {
            int64_t b = o->doc_off[m][d], e = o->doc_off[m][d + 1];
            for (int64_t i = b; i < e; i++) {
                uint32_t x[4];
                orc_draw(o, (uint32_t)(i - b), (uint32_t)(o->doc_base + d * o->doc_stride), 0, (uint32_t)m, ORC_PURPOSE_INIT, x);
                int t;
                if (m == 0 || len0 == 0) t = (int)(((uint64_t)x[0] * (uint64_t)K) >> 32);          /* M:500,506 */
                else t = o->z[0][b0 + (int64_t)(((uint64_t)x[0] * (uint64_t)len0) >> 32)];       /* M:503-504 */
                o->z[m][i] = t;
            }
        }
    }
    orc_rebuild_counts(o);
    return 0;
}
int orc_init_from_phi(orc_t *o)
{   /* inferencer initialisation I:186-203: topic = trees[m][type].sample(u) with bare-phi trees; OOV tokens keep 0 (Q13) */
    int K = o->K;
    double *tmp = (double *)malloc((size_t)K * 8), *tree = (double *)malloc((size_t)2 * K * 8);
    for (int64_t d = 0; d < o->D; d++)
        for (int m = 0; m < o->M; m++) {
            int64_t b = o->doc_off[m][d], e = o->doc_off[m][d + 1];
            for (int64_t i = b; i < e; i++) {
                int w = o->word[m][i];
                if (w < 0 || w >= o->V[m]) { o->z[m][i] = 0; continue; }
                build_tree_for_word_f(o, m, w, tree, tmp, ORC_F_BARE_TREES);
                uint32_t x[4];
                orc_draw(o, (uint32_t)(i - b), (uint32_t)(o->doc_base + d * o->doc_stride), 0, (uint32_t)m, ORC_PURPOSE_INIT, x);
                o->z[m][i] = orc_ftree_sample(tree, K, u24(x[0]));
            }
        }
    free(tmp); free(tree);
    return 0;
}
int orc_set_assignments(orc_t *o, int m, const int32_t *z)
{
    memcpy(o->z[m], z, (size_t)o->total_tokens[m] * 4);
    return 0;
}
int orc_get_assignments(const orc_t *o, int m, int32_t *z) { memcpy(z, o->z[m], (size_t)o->total_tokens[m] * 4); return 0; }
int orc_get_counts(const orc_t *o, int m, int32_t *n_wk, int32_t *n_k)
{
    if (n_wk) memcpy(n_wk, o->n_wk[m], (size_t)o->V[m] * o->K * 4);
    if (n_k) memcpy(n_k, o->n_k[m], (size_t)o->K * 4);
    return 0;
}
void orc_set_doc_ids(orc_t *o, int64_t base, int64_t stride) { o->doc_base = base; o->doc_stride = stride ? stride : 1; }
int orc_set_counts(orc_t *o, int m, const int32_t *n_wk, const int32_t *n_k)
{   /* multi-rank tests: install globally reduced counts */
    if (n_wk) memcpy(o->n_wk[m], n_wk, (size_t)o->V[m] * o->K * 4);
    if (n_k) memcpy(o->n_k[m], n_k, (size_t)o->K * 4);
    return 0;
}
int orc_maxlen(const orc_t *o, int m) { return o->maxlen[m]; }
int orc_get_hist(const orc_t *o, int m, int32_t *hist)
{ memcpy(hist, o->hist[m], (size_t)o->K * (o->maxlen[m] + 1) * 4); return 0; }
int orc_get_alpha(const orc_t *o, int m, double *alpha) { memcpy(alpha, o->alpha[m], (size_t)(o->K + 1) * 8); return 0; }
int orc_get_inactive(const orc_t *o, int32_t *out) { memcpy(out, o->inactive, (size_t)o->n_inactive * 4); return o->n_inactive; }
void orc_get_bucket_counters(const orc_t *o, int64_t *out4)
{ out4[0] = o->cnt_new; out4[1] = o->cnt_doc; out4[2] = o->cnt_tree; out4[3] = o->cnt_changed; }

/* ------------------------------------------------------------------------------------------------ */
/* delta application, U:197-270                                                                      */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { int32_t old_t, new_t, w, m, c_old, c_new; } orc_delta;   /* QD:14-34 */

static void apply_delta(orc_t *o, const orc_delta *d, int maintain_tree)
{
    int K = o->K, m = d->m;
    int32_t *row = o->n_wk[m] + (size_t)d->w * K;
    int32_t *hist = o->hist[m]; size_t hs = (size_t)o->maxlen[m] + 1;
    if (d->old_t != ORC_UNASSIGNED) row[d->old_t]--;                                  /* U:199-206 */
    row[d->new_t]++;                                                                  /* U:207     */
    if (d->old_t != ORC_UNASSIGNED) __atomic_fetch_sub(&o->n_k[m][d->old_t], 1, __ATOMIC_RELAXED);   /* U:209-216 */
    __atomic_fetch_add(&o->n_k[m][d->new_t], 1, __ATOMIC_RELAXED);                    /* U:218     */
    if (d->old_t != ORC_UNASSIGNED) {                                                 /* U:220-227 */
        __atomic_fetch_sub(&hist[d->old_t * hs + d->c_old + 1], 1, __ATOMIC_RELAXED);
        if (d->c_old > 0) __atomic_fetch_add(&hist[d->old_t * hs + d->c_old], 1, __ATOMIC_RELAXED);
    }
    if (d->c_new > 1) __atomic_fetch_sub(&hist[d->new_t * hs + d->c_new - 1], 1, __ATOMIC_RELAXED);   /* U:229-231 */
    __atomic_fetch_add(&hist[d->new_t * hs + d->c_new], 1, __ATOMIC_RELAXED);         /* U:232     */
    if (maintain_tree && o->tree[m]) {                                                /* U:242-260 */
        double *tree = o->tree[m] + (size_t)d->w * 2 * K;
        if (d->old_t != ORC_UNASSIGNED) orc_ftree_update(tree, K, d->old_t, leaf_value(o, m, d->w, d->old_t));
        orc_ftree_update(tree, K, d->new_t, leaf_value(o, m, d->w, d->new_t));
    }
    if (o->n_inactive && is_inactive(o, d->new_t)) {                                  /* U:263-270, Q16 */
        remove_inactive(o, d->new_t);
        o->alpha[m][d->new_t] = o->alpha[m][K];
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* per-document view-coupling draw, W:327-337                                                        */
/* ------------------------------------------------------------------------------------------------ */
static void draw_p(const orc_t *o, int64_t d, int iteration, unsigned flags, double p[ORC_MAXM][ORC_MAXM])
{
    int M = o->M;
    for (int m = 0; m < M; m++)
        for (int j = m; j < M; j++) {
            double r;
            if (m == j) r = 1.0;
            else if (o->p_a[m][j] == 0) r = 0.0;
            else {
                uint32_t x[4];
                orc_draw(o, 0, (uint32_t)(o->doc_base + d * o->doc_stride), (uint32_t)iteration, (uint32_t)(m * M + j), ORC_PURPOSE_PDRAW, x);
                double b;
                if (flags & ORC_F_BETA_MALLET) b = orc_next_beta_mallet(((uint64_t)x[1] << 32) | x[2], o->p_a[m][j], o->p_b[m][j]);
                else b = pow(u24(x[0]), 1.0 / o->p_a[m][j]);          /* true Beta(a,1), engine default (Q5) */
                r = floor(1000.0 * b + 0.5) / 1000.0;                   /* Java Math.round, W:333 (Q15) */
            }
            p[m][j] = (j != 0 && o->beta[j] == 0.0001) ? 0 : r;         /* W:335 (Q6) */
            p[j][m] = (m != 0 && o->beta[m] == 0.0001) ? 0 : r;         /* W:336      */
        }
}

/* ------------------------------------------------------------------------------------------------ */
/* reference-faithful sampler for one document, W:301-597                                            */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t *nd;       /* M*K  localTopicCounts            */
    int32_t *S;        /* K    localTopicIndex             */
    double *cum;       /* K    topicDocWordMasses          */
    double *O;         /* K    totalMassOtherModalities    */
    double *tmp, *ttree;   /* scratch for fresh trees      */
    orc_delta *dq; size_t ndq, capdq;   /* deferred deltas */
} orc_scratch;

static orc_scratch *scratch_new(const orc_t *o)
{
    orc_scratch *s = (orc_scratch *)calloc(1, sizeof(orc_scratch));
    int K = o->K;
    s->nd = (int32_t *)calloc((size_t)o->M * K, 4); s->S = (int32_t *)calloc((size_t)K + 1, 4);
    s->cum = (double *)calloc((size_t)K, 8); s->O = (double *)calloc((size_t)K, 8);
    s->tmp = (double *)calloc((size_t)K, 8); s->ttree = (double *)calloc((size_t)2 * K, 8);
    return s;
}
static void scratch_free(orc_scratch *s)
{ if (!s) return; free(s->nd); free(s->S); free(s->cum); free(s->O); free(s->tmp); free(s->ttree); free(s->dq); free(s); }

typedef void (*emit_fn)(void *ctx, const orc_delta *d);

static void sample_doc_reference(orc_t *o, int64_t d, int iteration, unsigned flags, orc_scratch *s,
                                 emit_fn emit, void *ctx, int64_t cnt[4])
{
    int M = o->M, K = o->K;
    double p[ORC_MAXM][ORC_MAXM];
    int len[ORC_MAXM] = { 0 };
    draw_p(o, d, iteration, flags, p);
    memset(s->nd, 0, (size_t)M * K * 4);
    for (int m = 0; m < M; m++) {                                       /* W:339-360 */
        int64_t b = o->doc_off[m][d], e = o->doc_off[m][d + 1];
        len[m] = (int)(e - b);
        for (int64_t i = b; i < e; i++) if (o->z[m][i] != ORC_UNASSIGNED) s->nd[m * K + o->z[m][i]]++;
    }
    int nz = 0;                                                          /* W:376-391 */
    for (int t = 0; t < K; t++)
        for (int i = 0; i < M; i++) if (s->nd[i * K + t] != 0) { s->S[nz++] = t; break; }
    uint8_t *rule = NULL;                                                /* the engine's "not in S" flags, checked against S */
    if ((flags & ORC_F_CHECK_RULE) && (flags & ORC_F_Q1_COMPAT)) rule = (uint8_t *)calloc((size_t)K, 1);

    for (int m = 0; m < M; m++) {                                        /* W:393 */
        memset(s->O, 0, (size_t)K * 8);
        double coefm = len[m] + o->gamma[m] * o->alphaSum[m];
        for (int di = 0; di < nz; di++) {                                /* W:399-410 */
            int t = s->S[di];
            double acc = 0;
            for (int i = 0; i < M; i++)
                if (i != m && len[i] != 0)
                    acc += p[m][i] * (s->nd[i * K + t] + o->gamma[i] * o->alpha[i][t]) / (len[i] + o->gamma[i] * o->alphaSum[i]);
            s->O[t] = acc * coefm;
        }
        double Cdoc = 0;                                                 /* W:413-418 */
        for (int i = 0; i < M; i++) Cdoc += p[m][i] * (o->gamma[i] * o->alpha[i][K]) / (len[i] + o->gamma[i] * o->alphaSum[i]);
        Cdoc *= coefm;

        int64_t b = o->doc_off[m][d];
        for (int pos = 0; pos < len[m]; pos++) {                         /* W:425 */
            int w = o->word[m][b + pos];
            if (w >= o->V[m] || w < 0) continue;                         /* W:427-428 */
            int old_t = o->z[m][b + pos];
            const int32_t *row = o->n_wk[m] + (size_t)w * K;
            if (old_t != ORC_UNASSIGNED) {                               /* W:434-471 */
                s->nd[m * K + old_t]--;
                int deleted = 1;
                for (int j = 0; j < M && deleted; j++) deleted = (s->nd[j * K + old_t] == 0);
                if (deleted && rule) rule[old_t] = 1;
                if (deleted) {
                    int di = 0;
                    while (di < nz && s->S[di] != old_t) di++;
                    if (di < nz) {   /* (under Q1 a topic gained this sweep may be missing from S) */
                        for (; di < nz - 1; di++) s->S[di] = s->S[di + 1];
                        nz--;
                    }
                }
            }
            if (rule) {                                                  /* S must equal {held and not flagged} at every token */
                int bad = 0, di = 0;
                for (int t = 0; t < K; t++) {
                    int held = 0;
                    for (int j = 0; j < M; j++) held |= (s->nd[j * K + t] != 0);
                    int in_rule = held && !rule[t];
                    int in_ref = (di < nz && s->S[di] == t);
                    if (in_ref) di++;
                    bad |= (in_rule != in_ref);
                }
                o->cnt_rule_bad += bad;
            }
            double acc = 0;                                              /* W:496-513 */
            for (int di = 0; di < nz; di++) {
                int t = s->S[di];
                double phi = (row[t] + o->beta[m]) / (__atomic_load_n(&o->n_k[m][t], __ATOMIC_RELAXED) + o->betaSum[m]);
                acc += (p[m][m] * s->nd[m * K + t] + s->O[t]) * phi;
                s->cum[di] = acc;
            }
            double C = o->n_inactive == 0 ? 0 : Cdoc / K;                /* W:515 */
            const double *tree;
            if (flags & ORC_F_STALE_TREES) tree = o->tree[m] + (size_t)w * 2 * K;
            else { build_tree_for_word_f(o, m, w, s->ttree, s->tmp, flags); tree = s->ttree; }
            double B = tree[1];
            uint32_t x[4];
            orc_draw(o, (uint32_t)pos, (uint32_t)(o->doc_base + d * o->doc_stride), (uint32_t)iteration, (uint32_t)m, ORC_PURPOSE_SAMPLE, x);
            double sample = u24(x[0]) * (C + acc + B);                   /* W:517-519 */
            int new_t;
            if (sample < C) { new_t = o->inactive[0]; cnt[0]++; }        /* W:522-526 */
            else {
                sample -= C;
                if (sample < acc) { int lb = orc_lower_bound(s->cum, sample, nz); new_t = lb < 0 ? -1 : s->S[lb]; cnt[1]++; }   /* W:529-531 */
                else { new_t = orc_ftree_sample(tree, K, u24(x[1])); cnt[2]++; }                                              /* W:533-535 */
            }
            if (new_t == -1) new_t = K - 1;                              /* W:549-553 */
            o->z[m][b + pos] = new_t;                                    /* W:557 */
            if (rule) {
                int held = 0;
                for (int j = 0; j < M; j++) held |= (s->nd[j * K + new_t] != 0);
                if (!held) rule[new_t] = 1;
            }
            s->nd[m * K + new_t]++;                                      /* W:560 */
            if (!(flags & ORC_F_Q1_COMPAT)) {
                /* intended semantics of W:563-584: insert when the topic was absent from every view */
                int isnew = (s->nd[m * K + new_t] == 1);
                for (int j = 0; j < M && isnew; j++) if (j != m) isnew = (s->nd[j * K + new_t] == 0);
                if (isnew) {
                    int di = nz;
                    while (di > 0 && s->S[di - 1] > new_t) { s->S[di] = s->S[di - 1]; di--; }
                    s->S[di] = new_t; nz++;
                    /* O for a freshly inserted topic: the prior part of W:404 (no counts elsewhere) */
                    double a2 = 0;
                    for (int i = 0; i < M; i++)
                        if (i != m && len[i] != 0) a2 += p[m][i] * (o->gamma[i] * o->alpha[i][new_t]) / (len[i] + o->gamma[i] * o->alphaSum[i]);
                    s->O[new_t] = a2 * coefm;
                }
            }
            if (new_t != old_t) {                                        /* W:587-589 */
                orc_delta dl = { old_t, new_t, w, m, old_t == ORC_UNASSIGNED ? 0 : s->nd[m * K + old_t], s->nd[m * K + new_t] };
                emit(ctx, &dl); cnt[3]++;
            }
        }
    }
    free(rule);
}

typedef struct { orc_t *o; orc_scratch *s; unsigned flags; } emit_ctx;
static void emit_immediate(void *c, const orc_delta *d)
{ emit_ctx *e = (emit_ctx *)c; apply_delta(e->o, d, 1); }
static void emit_nothing(void *c, const orc_delta *d) { (void)c; (void)d; }   /* nut = 0: W:587 never enqueues */
static void emit_deferred(void *c, const orc_delta *d)
{
    emit_ctx *e = (emit_ctx *)c; orc_scratch *s = e->s;
    if (s->ndq == s->capdq) { s->capdq = s->capdq ? s->capdq * 2 : 4096; s->dq = (orc_delta *)realloc(s->dq, s->capdq * sizeof(orc_delta)); }
    s->dq[s->ndq++] = *d;
}

/* ------------------------------------------------------------------------------------------------ */
/* engine-mirror sampler: the engine's target distribution (SURVEY Appendix A "net distribution")     */
/* evaluated densely in fp64, scanned in the engine's lane-major order with the same Philox uniform.   */
/* ------------------------------------------------------------------------------------------------ */
static inline int engine_order_topic(int idx, int JG, int G)
{   /* idx-th topic in scan order: lane-major over (lane, j, e); topic = 4*(lane + G*j) + e */
    int e = idx & 3, j = (idx >> 2) % JG, lane = (idx >> 2) / JG;
    return 4 * (lane + G * j) + e;
}
static int engine_slot_size(int K)
{   /* the engine's slot sizes: 128 * {1,2,3,4,6,8,12,16} */
    static const int opts[] = { 1, 2, 3, 4, 6, 8, 12, 16 };
    for (int i = 0; i < 8; i++) if (opts[i] * 128 >= K) return opts[i] * 128;
    return ((K + 127) / 128) * 128;
}

/* unnormalised engine weights for token (d, m, pos) given local counts nd (own token already removed),
 * other-view state and the frozen topic totals nk_frozen.  out has K entries. */
static unsigned g_engine_weight_flags = 0;   /* set by the sweep for the duration of a mirror pass (single-threaded) */
static const uint8_t *g_not_in_S = NULL;     /* K flags of the document being sampled (Q1), or NULL */
static void engine_weights(const orc_t *o, int m, int w, const int32_t *nd, const int len[ORC_MAXM],
                           double p[ORC_MAXM][ORC_MAXM], const int32_t *nk_frozen, double *out)
{
    int M = o->M, K = o->K;
    const int32_t *row = o->n_wk[m] + (size_t)w * K;
    double coefm = len[m] + o->gamma[m] * o->alphaSum[m];
    for (int t = 0; t < K; t++) {
        int inS = 0;
        for (int i = 0; i < M; i++) if (nd[i * K + t] != 0) { inS = 1; break; }
        if (g_not_in_S && g_not_in_S[t]) inS = 0;      /* Q1: held, but not in the dense index: neither n_d nor O enters (W:501-513 walks S) */
        double O = 0;
        if (inS && M > 1) {
            for (int i = 0; i < M; i++)
                if (i != m && len[i] != 0)
                    O += p[m][i] * (nd[i * K + t] + o->gamma[i] * o->alpha[i][t]) / (len[i] + o->gamma[i] * o->alphaSum[i]);
            O *= coefm;
        }
        double ga = (o->n_inactive && is_inactive(o, t)) ? 0.0 : o->gamma[m] * o->alpha[m][t];
        if (g_engine_weight_flags & ORC_F_BARE_TREES) ga = 1.0;
        double phi = (row[t] + o->beta[m]) / (nk_frozen[t] + o->betaSum[m]);
        out[t] = phi * ((inS ? p[m][m] * nd[m * K + t] : 0.0) + O + ga);
    }
}

int orc_engine_select_g(const double *wgt, int K, double u, double C, int first_inactive, int G)
{   /* C bucket first (W:522), then a single scan in engine order; returns the topic */
    int n = engine_slot_size(K), J = n / (4 * G);
    double total = 0;
    for (int idx = 0; idx < n; idx++) { int t = engine_order_topic(idx, J, G); if (t < K) total += wgt[t]; }
    double s = u * (total + C);
    if (s < C) return first_inactive;
    s -= C;
    double cum = 0; int last = -1;
    for (int idx = 0; idx < n; idx++) {
        int t = engine_order_topic(idx, J, G);
        if (t >= K) continue;
        cum += wgt[t];
        if (wgt[t] > 0) last = t;
        if (cum > s) return t;
    }
    return last;
}

int orc_engine_select(const double *wgt, int K, double u, double C, int first_inactive)
{ return orc_engine_select_g(wgt, K, u, C, first_inactive, 32); }
void orc_set_engine_group(orc_t *o, int G) { o->engine_G = G; }

static void sample_docview_engine(orc_t *o, int64_t d, int m, int iteration, unsigned flags, orc_scratch *s,
                                  const int32_t *nk_frozen, int32_t *dnk, int64_t cnt[4])
{
    int M = o->M, K = o->K;
    double p[ORC_MAXM][ORC_MAXM];
    int len[ORC_MAXM] = { 0 };
    draw_p(o, d, iteration, flags, p);
    memset(s->nd, 0, (size_t)M * K * 4);
    for (int i = 0; i < M; i++) {
        int64_t b = o->doc_off[i][d], e = o->doc_off[i][d + 1];
        len[i] = (int)(e - b);
        for (int64_t k = b; k < e; k++) if (o->z[i][k] != ORC_UNASSIGNED) s->nd[i * K + o->z[i][k]]++;
    }
    double coefm = len[m] + o->gamma[m] * o->alphaSum[m];
    double Cdoc = 0;
    for (int i = 0; i < M; i++) Cdoc += p[m][i] * (o->gamma[i] * o->alpha[i][K]) / (len[i] + o->gamma[i] * o->alphaSum[i]);
    Cdoc *= coefm;
    double C = o->n_inactive == 0 ? 0 : Cdoc / K;
    int64_t b = o->doc_off[m][d];
    uint8_t *rf = ((flags & ORC_F_Q1_COMPAT) && o->rflag) ? o->rflag + (size_t)d * K : NULL;
    g_not_in_S = rf;
    for (int pos = 0; pos < len[m]; pos++) {
        int w = o->word[m][b + pos];
        if (w >= o->V[m] || w < 0) continue;
        int old_t = o->z[m][b + pos];
        if (old_t != ORC_UNASSIGNED) {
            s->nd[m * K + old_t]--;
            if (rf) {                                   /* W:441-468: no view holds it any more -> it leaves the index for good */
                int held = 0;
                for (int i = 0; i < M; i++) held |= (s->nd[i * K + old_t] != 0);
                if (!held) rf[old_t] = 1;
            }
        }
        engine_weights(o, m, w, s->nd, len, p, nk_frozen, s->cum);
        uint32_t x[4];
        orc_draw(o, (uint32_t)pos, (uint32_t)(o->doc_base + d * o->doc_stride), (uint32_t)iteration, (uint32_t)m, ORC_PURPOSE_SAMPLE, x);
        int new_t = orc_engine_select_g(s->cum, K, u24(x[0]), C, o->n_inactive ? o->inactive[0] : -1, o->engine_G);
        o->z[m][b + pos] = new_t;
        if (rf) {                                       /* W:563-584 is dead code: a topic gained while absent is never inserted */
            int held = 0;
            for (int i = 0; i < M; i++) held |= (s->nd[i * K + new_t] != 0);
            if (!held) rf[new_t] = 1;
        }
        s->nd[m * K + new_t]++;
        if (new_t != old_t && !(flags & ORC_F_FROZEN)) {   /* n_wk live, n_k deferred to the end of the view pass (engine semantics) */
            int32_t *row = o->n_wk[m] + (size_t)w * K;
            if (old_t != ORC_UNASSIGNED) { row[old_t]--; dnk[old_t]--; }
            row[new_t]++; dnk[new_t]++;
            cnt[3]++;
        }
    }
}

static int cmp_len_desc(const void *a, const void *b, void *arg)
{
    const int64_t *off = (const int64_t *)arg;
    int64_t da = *(const int64_t *)a, db = *(const int64_t *)b;
    int64_t la = off[da + 1] - off[da], lb = off[db + 1] - off[db];
    if (la != lb) return la > lb ? -1 : 1;
    return da < db ? -1 : (da > db);
}

static void recount_nk_hist(orc_t *o, int m)
{
    int K = o->K;
    int32_t *local = (int32_t *)calloc((size_t)K, 4);
    memset(o->n_k[m], 0, (size_t)K * 4);
    memset(o->hist[m], 0, (size_t)K * (o->maxlen[m] + 1) * 4);
    for (int64_t d = 0; d < o->D; d++) {
        if (!o->present[m][d]) continue;
        for (int64_t i = o->doc_off[m][d]; i < o->doc_off[m][d + 1]; i++) { int t = o->z[m][i]; if (t >= 0) { local[t]++; o->n_k[m][t]++; } }
        for (int t = 0; t < K; t++) { o->hist[m][(size_t)t * (o->maxlen[m] + 1) + local[t]]++; local[t] = 0; }
    }
    free(local);
}

/* activation of inactive topics at a sweep boundary (engine semantics of U:263-270) */
static void activate_sampled_topics(orc_t *o)
{
    for (int i = 0; i < o->n_inactive; ) {
        int t = o->inactive[i], hit = 0;
        for (int m = 0; m < o->M; m++) if (o->n_k[m][t] > 0) { o->alpha[m][t] = o->alpha[m][o->K]; hit = 1; }
        if (hit) remove_inactive(o, t); else i++;
    }
}

long long orc_rule_violations(const orc_t *o) { return (long long)o->cnt_rule_bad; }

int orc_sweep(orc_t *o, int iteration, unsigned flags)
{
    orc_scratch *s = scratch_new(o);
    int64_t cnt[4] = { 0, 0, 0, 0 };
    if (flags & ORC_F_ENGINE_MIRROR) {
        int K = o->K;
        g_engine_weight_flags = flags;
        if (flags & ORC_F_Q1_COMPAT) {                   /* the dense index is rebuilt per document every sweep (W:376-391) */
            if (!o->rflag) o->rflag = (uint8_t *)malloc((size_t)(o->D > 0 ? o->D : 1) * K);
            memset(o->rflag, 0, (size_t)(o->D > 0 ? o->D : 1) * K);
        }
        int32_t *nk_frozen = (int32_t *)malloc((size_t)K * 4);
        int32_t *dnk = (int32_t *)malloc((size_t)K * 4);
        int64_t *order = (int64_t *)malloc((size_t)(o->D > 0 ? o->D : 1) * 8);
        for (int m = 0; m < o->M; m++) {
            memcpy(nk_frozen, o->n_k[m], (size_t)K * 4);
            for (int64_t d = 0; d < o->D; d++) order[d] = d;
            if (!(flags & ORC_F_DOC_ORDER)) qsort_r(order, (size_t)o->D, 8, cmp_len_desc, o->doc_off[m]);
            memset(dnk, 0, (size_t)K * 4);
            for (int64_t k = 0; k < o->D; k++) sample_docview_engine(o, order[k], m, iteration, flags, s, nk_frozen, dnk, cnt);
            g_not_in_S = NULL;
            if (!(flags & ORC_F_FROZEN)) {       /* the engine flushes its n_k deltas at the end of the view pass */
                memcpy(nk_frozen, o->n_k[m], (size_t)K * 4);
                recount_nk_hist(o, m);           /* local doc-topic histogram (and a local n_k, replaced below)  */
                for (int t = 0; t < K; t++) o->n_k[m][t] = nk_frozen[t] + dnk[t];
                if (m + 1 < o->M) activate_sampled_topics(o);    /* the engine activates after every view pass (mvtm.cu sweep_impl) */
            }
        }
        free(nk_frozen); free(dnk); free(order);
        g_engine_weight_flags = 0;
        if (!(flags & ORC_F_FROZEN)) activate_sampled_topics(o);
    } else {
        if ((flags & ORC_F_STALE_TREES) && !o->tree[0]) orc_rebuild_trees(o);
        emit_ctx ctx = { o, s, flags };
        emit_fn emit = (flags & ORC_F_FROZEN) ? emit_nothing : ((flags & ORC_F_DEFERRED) ? emit_deferred : emit_immediate);
        for (int64_t d = 0; d < o->D; d++) sample_doc_reference(o, d, iteration, flags, s, emit, &ctx, cnt);
        for (size_t i = 0; i < s->ndq; i++) apply_delta(o, &s->dq[i], 1);
    }
    o->cnt_new += cnt[0]; o->cnt_doc += cnt[1]; o->cnt_tree += cnt[2]; o->cnt_changed += cnt[3];
    scratch_free(s);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* conditional distribution of one token on frozen counts                                            */
/* ------------------------------------------------------------------------------------------------ */
int orc_cond_probs_ex(orc_t *o, int m, int64_t d, int pos, const double *p_in, int engine_form, const uint8_t *not_in_S, unsigned flags, double *out);
int orc_cond_probs_q1(orc_t *o, int m, int64_t d, int pos, const double *p_in, int engine_form, const uint8_t *not_in_S, double *out)
{ return orc_cond_probs_ex(o, m, d, pos, p_in, engine_form, not_in_S, 0, out); }
int orc_cond_probs(orc_t *o, int m, int64_t d, int pos, const double *p_in, int engine_form, double *out)
{ return orc_cond_probs_ex(o, m, d, pos, p_in, engine_form, NULL, 0, out); }

/* not_in_S (K flags, may be NULL): topics the document holds that the reference's dense index lacks at this token (quirk Q1:
 * gained earlier in the sweep); the token's own removal (W:434-471) is applied here on top of it. */
/* flags: ORC_F_BARE_TREES = the inferencer's trees (I:561-576: leaves hold phi without gamma*alpha and ignore the inactive set, Q13) */
int orc_cond_probs_ex(orc_t *o, int m, int64_t d, int pos, const double *p_in, int engine_form, const uint8_t *not_in_S, unsigned flags, double *out)
{   /* reference form: masses of the three buckets of W:495-538 with freshly built trees (M:2660-2691);
     * engine form: the dense net distribution.  Both from the sweep-start state of the document (Q1/Q3
     * do not matter there).  out[0..K) normalised probabilities, out[K] = share of the new-topic bucket. */
    int M = o->M, K = o->K;
    if (m < 0 || m >= M || d < 0 || d >= o->D) return 1;
    int64_t b = o->doc_off[m][d];
    if (pos < 0 || pos >= o->doc_off[m][d + 1] - b) return 2;
    double p[ORC_MAXM][ORC_MAXM];
    for (int i = 0; i < M; i++) for (int j = 0; j < M; j++) p[i][j] = p_in ? p_in[i * M + j] : (i == j ? 1.0 : 0.0);
    orc_scratch *s = scratch_new(o);
    int len[ORC_MAXM];
    for (int i = 0; i < M; i++) {
        int64_t bb = o->doc_off[i][d], e = o->doc_off[i][d + 1];
        len[i] = (int)(e - bb);
        for (int64_t k = bb; k < e; k++) if (o->z[i][k] != ORC_UNASSIGNED) s->nd[i * K + o->z[i][k]]++;
    }
    int w = o->word[m][b + pos], old_t = o->z[m][b + pos];
    if (w < 0 || w >= o->V[m]) { scratch_free(s); return 3; }
    if (old_t != ORC_UNASSIGNED) s->nd[m * K + old_t]--;
    double coefm = len[m] + o->gamma[m] * o->alphaSum[m];
    double Cdoc = 0;
    for (int i = 0; i < M; i++) Cdoc += p[m][i] * (o->gamma[i] * o->alpha[i][K]) / (len[i] + o->gamma[i] * o->alphaSum[i]);
    double C = o->n_inactive == 0 ? 0 : Cdoc * coefm / K;
    double total = C;
    if (engine_form) {
        g_not_in_S = not_in_S;
        g_engine_weight_flags = flags;
        engine_weights(o, m, w, s->nd, len, p, o->n_k[m], out);
        g_engine_weight_flags = 0;
        g_not_in_S = NULL;
        for (int t = 0; t < K; t++) total += out[t];
    } else {
        const int32_t *row = o->n_wk[m] + (size_t)w * K;
        for (int t = 0; t < K; t++) {
            int inS = 0;
            for (int i = 0; i < M; i++) if (s->nd[i * K + t] != 0) { inS = 1; break; }
            if (not_in_S && not_in_S[t]) inS = 0;
            double A = 0;
            if (inS) {
                double O = 0;
                for (int i = 0; i < M; i++)
                    if (i != m && len[i] != 0)
                        O += p[m][i] * (s->nd[i * K + t] + o->gamma[i] * o->alpha[i][t]) / (len[i] + o->gamma[i] * o->alphaSum[i]);
                O *= coefm;
                double phi = (row[t] + o->beta[m]) / (o->n_k[m][t] + o->betaSum[m]);
                A = (p[m][m] * s->nd[m * K + t] + O) * phi;
            }
            double leaf = (flags & ORC_F_BARE_TREES) ? phi_value(o, m, w, t)
                                                     : ((o->n_inactive && is_inactive(o, t)) ? 0.0 : leaf_value(o, m, w, t));
            out[t] = A + leaf;
            total += out[t];
        }
    }
    if (C > 0) out[o->inactive[0]] += C;
    for (int t = 0; t < K; t++) out[t] /= total;
    out[K] = C / total;
    scratch_free(s);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* modelLogLikelihood, M:3322-3452                                                                   */
/* ------------------------------------------------------------------------------------------------ */
int orc_loglik(const orc_t *o, double *ll_out, int quirk_len2)
{
    int K = o->K;
    int32_t *tc = (int32_t *)calloc((size_t)K, 4);
    double *tlg = (double *)malloc((size_t)K * 8);
    for (int m = 0; m < o->M; m++) {
        double ll = 0;
        for (int t = 0; t < K; t++) tlg[t] = orc_log_gamma_stirling(o->gamma[m] * o->alpha[m][t]);     /* M:3343 */
        int64_t modalityCnt = 0;
        for (int64_t d = 0; d < o->D; d++) {
            if (!o->present[m][d]) continue;                                                          /* M:3348 */
            int64_t b = o->doc_off[m][d], e = o->doc_off[m][d + 1];
            int len = (int)(e - b);
            int arrlen = quirk_len2 ? (len < 2 ? 2 : len) : len;     /* Q18: MALLET FeatureSequence capacity max(n,2) */
            if (arrlen > 0) {                                                                         /* M:3352 */
                for (int64_t i = b; i < e; i++) tc[o->z[m][i]]++;
                tc[0] += arrlen - len;                                /* phantom zero entries of the backing array */
                for (int t = 0; t < K; t++)
                    if (tc[t] > 0) ll += orc_log_gamma_stirling(o->gamma[m] * o->alpha[m][t] + tc[t]) - tlg[t];   /* M:3359 */
                ll -= orc_log_gamma_stirling(o->gamma[m] * o->alphaSum[m] + arrlen);                  /* M:3365 */
                modalityCnt++;
                for (int64_t i = b; i < e; i++) tc[o->z[m][i]] = 0;
                tc[0] = 0;
            }
        }
        ll += modalityCnt * orc_log_gamma_stirling(o->gamma[m] * o->alphaSum[m]);                     /* M:3373 */
        int64_t nnz = 0;
        size_t n = (size_t)o->V[m] * K;
        for (size_t i = 0; i < n; i++) {                                                              /* M:3389-3415 */
            int c = o->n_wk[m][i];
            if (c > 0) { nnz++; ll += orc_log_gamma_stirling(o->beta[m] + c); }
        }
        for (int t = 0; t < K; t++) ll -= orc_log_gamma_stirling(o->beta[m] * o->V[m] + o->n_k[m][t]);   /* M:3417-3419 */
        ll += orc_log_gamma_stirling(o->beta[m] * o->V[m]) * K;                                       /* M:3438 */
        ll -= orc_log_gamma_stirling(o->beta[m]) * nnz;                                               /* M:3441 */
        ll_out[m] = ll;
    }
    free(tc); free(tlg);
    return 0;
}

/* count invariants; returns number of violated cells */
int64_t orc_check_invariants(const orc_t *o)
{
    int K = o->K; int64_t bad = 0;
    for (int m = 0; m < o->M; m++) {
        int32_t *nwk = (int32_t *)calloc((size_t)o->V[m] * K, 4);
        int32_t *nk = (int32_t *)calloc((size_t)K, 4);
        for (int64_t i = 0; i < o->total_tokens[m]; i++) {
            int w = o->word[m][i], t = o->z[m][i];
            if (t < 0 || t >= K) { bad++; continue; }
            nk[t]++;
            if (w >= 0 && w < o->V[m]) nwk[(size_t)w * K + t]++;
        }
        for (size_t i = 0; i < (size_t)o->V[m] * K; i++) bad += (nwk[i] != o->n_wk[m][i]);
        for (int t = 0; t < K; t++) bad += (nk[t] != o->n_k[m][t]);
        free(nwk); free(nk);
    }
    return bad;
}

/* ------------------------------------------------------------------------------------------------ */
/* multithreaded restatement of the reference scheme (M:1033-1146,1213-1239): nst = 3T/4 sampler       */
/* threads over contiguous document blocks, nut = T/4 updater threads owning words w % nut, one queue  */
/* per (sampler, updater) pair, stale fp64 F+trees.  Used as the CPU baseline in bench.py.             */
/* ------------------------------------------------------------------------------------------------ */
#define Q_CAP (1u << 16)
typedef struct {
    orc_delta *buf;
    volatile uint64_t head __attribute__((aligned(64)));
    volatile uint64_t tail __attribute__((aligned(64)));
} spsc_q;

typedef struct {
    orc_t *o; int iteration; unsigned flags; int nst, nut, id;
    int64_t d0, d1; spsc_q *queues; int64_t cnt[4];
} mt_arg;

typedef struct { mt_arg *a; } mt_emit_ctx;
static void emit_queue(void *c, const orc_delta *d)
{   /* W:589: queue nst*(type % nut) + threadId */
    mt_arg *a = ((mt_emit_ctx *)c)->a;
    spsc_q *q = &a->queues[(size_t)a->nst * (d->w % a->nut) + a->id];
    uint64_t t = q->tail;
    while (t - __atomic_load_n(&q->head, __ATOMIC_ACQUIRE) >= Q_CAP) sched_yield();
    q->buf[t & (Q_CAP - 1)] = *d;
    __atomic_store_n(&q->tail, t + 1, __ATOMIC_RELEASE);
}
static void *mt_sampler(void *v)
{
    mt_arg *a = (mt_arg *)v;
    orc_scratch *s = scratch_new(a->o);
    mt_emit_ctx ctx = { a };
    for (int64_t d = a->d0; d < a->d1; d++)
        sample_doc_reference(a->o, d, a->iteration, a->flags | ORC_F_STALE_TREES, s, emit_queue, &ctx, a->cnt);
    orc_delta fin = { -1, -1, -1, -1, -1, -1 };                          /* W:216-218 sentinel */
    for (int ut = 0; ut < a->nut; ut++) {
        spsc_q *q = &a->queues[(size_t)a->nst * ut + a->id];
        uint64_t t = q->tail;
        while (t - __atomic_load_n(&q->head, __ATOMIC_ACQUIRE) >= Q_CAP) sched_yield();
        q->buf[t & (Q_CAP - 1)] = fin;
        __atomic_store_n(&q->tail, t + 1, __ATOMIC_RELEASE);
    }
    scratch_free(s);
    return NULL;
}
static void *mt_updater(void *v)
{   /* U:181-282; the reference sleeps 20 ms when drained, this restatement only yields */
    mt_arg *a = (mt_arg *)v;
    int finished = 0;
    while (finished < a->nst) {
        int progressed = 0;
        for (int st = 0; st < a->nst; st++) {
            spsc_q *q = &a->queues[(size_t)a->id * a->nst + st];
            uint64_t h = q->head;
            while (h < __atomic_load_n(&q->tail, __ATOMIC_ACQUIRE)) {
                orc_delta d = q->buf[h & (Q_CAP - 1)];
                __atomic_store_n(&q->head, ++h, __ATOMIC_RELEASE);
                progressed = 1;
                if (d.m == -1 && d.new_t == -1) { finished++; continue; }
                apply_delta(a->o, &d, 1);
            }
        }
        if (!progressed) sched_yield();
    }
    return NULL;
}
int orc_sweep_mt(orc_t *o, int iteration, int num_threads, unsigned flags)
{
    int nut = num_threads / 4, nst = 3 * num_threads / 4;            /* M:1036-1037 */
    if (nut < 1 || nst < 1) return 1;                                /* Q8 */
    if (o->n_inactive) return 2;                                     /* inactive-set mutation is not thread-safe here */
    if (!o->tree[0]) if (orc_rebuild_trees(o)) return 3;
    spsc_q *queues = (spsc_q *)calloc((size_t)nst * nut, sizeof(spsc_q));
    for (int i = 0; i < nst * nut; i++) queues[i].buf = (orc_delta *)malloc(sizeof(orc_delta) * Q_CAP);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)(nst + nut));
    mt_arg *args = (mt_arg *)calloc((size_t)(nst + nut), sizeof(mt_arg));
    int64_t per = o->D / nst;                                        /* M:1051 */
    for (int i = 0; i < nst; i++) {
        args[i] = (mt_arg){ o, iteration, flags, nst, nut, i, per * i, (i == nst - 1) ? o->D : per * (i + 1), queues, {0,0,0,0} };
        pthread_create(&th[i], NULL, mt_sampler, &args[i]);
    }
    for (int i = 0; i < nut; i++) {
        args[nst + i] = (mt_arg){ o, iteration, flags, nst, nut, i, 0, 0, queues, {0,0,0,0} };
        pthread_create(&th[nst + i], NULL, mt_updater, &args[nst + i]);
    }
    for (int i = 0; i < nst + nut; i++) pthread_join(th[i], NULL);   /* the CyclicBarrier of M:1231 */
    for (int i = 0; i < nst; i++) { o->cnt_new += args[i].cnt[0]; o->cnt_doc += args[i].cnt[1]; o->cnt_tree += args[i].cnt[2]; o->cnt_changed += args[i].cnt[3]; }
    for (int i = 0; i < nst * nut; i++) free(queues[i].buf);
    free(queues); free(th); free(args);
    return 0;
}
