// mvtm.cu -- host side of the C ABI declared in include/mvtm.h (libmvtm.so).
// Plain pointers and sizes only; no torch types.  There is no CPU fallback: every compute entry point
// needs a CUDA device and reports MVTM_ERR_CUDA otherwise.
#include "mvtm.h"
#include "mvtm_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

constexpr int HOST_CHUNKS = 8;                     // document ranges of the pipelined mvtm_sweep_host

struct ViewDev {
    bool added = false;
    int V = 0;
    long long n_tok = 0;
    int max_len = 0;
    long long docs_present = 0;
    int n_items = 0;
    long long *doc_off = nullptr;
    int *word = nullptr, *z = nullptr;
    unsigned char *present = nullptr;
    int *nwk = nullptr, *nk = nullptr, *nk_snap = nullptr;
    int *order = nullptr;
    int *z_mirror = nullptr;                        // device alias of a caller-owned pinned array that mirrors z (mvtm_set_host_mirror)
    int *z_host_once = nullptr;                     // the same for the passes of ONE call (mvtm_sweep_host_dist), wins over z_mirror
    int *z_stage = nullptr;                         // mvtm_sweep_host_dist: the uploaded assignments before they are compared with z
    std::vector<long long> chunk_tok_off;           // mvtm_sweep_host: HOST_CHUNKS + 1 token offsets of contiguous document ranges
    float *ga_tree = nullptr, *ga_full = nullptr, *ga_one = nullptr;   // ga_one: all ones, the inferencer's bare trees (Q13)
    int *snap_nwk = nullptr, *snap_nk = nullptr;    // multi-GPU delta snapshots
    std::vector<long long> h_doc_off;               // host copy (lengths, probe argument checks)
    std::vector<unsigned char> h_present;
    int tune_step = 0, ring_locked = 0;             // ring-depth autotune over the first sweeps (see ring_for_view)
    float tune_ms[8] = { 0.f };
};

}  // namespace

struct CommState;                                   // NCCL communicators and exchange state (mvtm_comm.inl)

struct mvtm_handle {
    CommState *comm = nullptr;                      // multi-GPU inside the library (mvtm_comm_init)
    int K = 0, M = 0, Kp = 0, J = 0, KS = 0, G = 32;
    unsigned long long mut_epoch = 1;               // bumped by everything that writes assignments or count tables on the device
    bool direct = false;                            // DIRECT sweep kernel: n_wk rows in registers instead of the TMA ring (mvtm_kernels.cuh)
    long long D = 0;
    int device = 0, num_sms = 0;
    unsigned flags = 0;
    unsigned long long seed = 0;
    long long doc_id_base = 0, doc_id_stride = 1;
    int cfg_warps = 0, cfg_ring = 0, cfg_ctas = 0;
    ViewDev v[MVTM_MAX_VIEWS];
    // hyper-parameters (host, fp64 as the reference keeps them)
    std::vector<double> alpha;                      // M x (K+1)
    double alphaSum[MVTM_MAX_VIEWS], beta[MVTM_MAX_VIEWS], betaSum[MVTM_MAX_VIEWS], gamma[MVTM_MAX_VIEWS];
    double p_a[MVTM_MAX_VIEWS][MVTM_MAX_VIEWS], p_b[MVTM_MAX_VIEWS][MVTM_MAX_VIEWS];
    std::vector<int> inactive;
    // hierarchical-DP state of the hyper-parameter step (M:136-139): root/view concentrations and table counts
    double gammaRoot = 10.0, rootTablesCnt = 0.0;
    double gammaView[MVTM_MAX_VIEWS] = { 0 }, tablesCnt[MVTM_MAX_VIEWS] = { 0 };
    double pMean[MVTM_MAX_VIEWS][MVTM_MAX_VIEWS] = { { 0 } };
    bool hyper_dirty = true;
    mvtm_stat_reducer reducer = nullptr;            // multi-rank hyper-parameter step (mvtm_set_stat_reducer)
    void *reducer_ctx = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;             // mvtm_sweep_host: H2D / D2H of z chunks beside the kernels
    std::vector<cudaEvent_t> host_ev;               // per (view, chunk): upload done
    cudaEvent_t ev[2 * MVTM_MAX_VIEWS + 2] = { nullptr };
    cudaEvent_t ev_done[MVTM_MAX_VIEWS] = { nullptr }, ev_ready[MVTM_MAX_VIEWS] = { nullptr };   // hand-over points with a caller-owned stream (async exchange)
    bool ready_pending[MVTM_MAX_VIEWS] = { false }, pass_queued[MVTM_MAX_VIEWS] = { false };
    bool sweep_open = false;
    int open_launches = 0, open_mode = 1;
    int *work_counter = nullptr;
    unsigned *rbits = nullptr;                      // MVTM_FLAG_Q1_COMPAT: per document KS/32 flag words (see SweepParams::rbits)
    float *oc_scratch = nullptr;                    // multi-view: Kp floats per resident document slot
    size_t oc_scratch_floats = 0;
    unsigned long long *d_stats = nullptr;
    int *d_bad = nullptr;
    mvtm_sweep_stats stats;
    std::string err;
};

static std::string g_create_err;

#define FAIL(h, code, ...)                                              \
    do {                                                                \
        char _b[512]; snprintf(_b, sizeof(_b), __VA_ARGS__);            \
        (h)->err = _b; return (code);                                   \
    } while (0)
#define CK(h, call)                                                                                          \
    do {                                                                                                     \
        cudaError_t _e = (call);                                                                             \
        if (_e != cudaSuccess) FAIL(h, MVTM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

static int wait_view_ready(mvtm_handle *h, int m);
static int wait_all_ready(mvtm_handle *h);
static void comm_teardown(mvtm_handle *h);
static int comm_reduce_stats(mvtm_handle *h, int op, long long *ints, long long n_ints, double *reals, long long n_reals);

static int pick_J(int K)
{
    static const int opts[] = { 1, 2, 3, 4, 6, 8, 12, 16 };
    for (int j : opts) if (j * 128 >= K) return j;
    return 0;
}
// lanes per document-view (the lane group G of mvtm_kernels.cuh) for a slot size KS = 128 * J.
// Compiled combinations: (128,8) (256,8) (384,16) (512,8) (512,16) (512,32) (768,16) (1024,16) (1024,32) (1536,32) (2048,32).
// DIRECT kernels compiled (mvtm_kernels.cuh): rows in registers cost KS/G registers per thread, so only combinations with
// KS/G <= 64; not in reference-compat (Q1) mode, whose kernels stay on the TMA ring.
// Measured on B200 (profiles/r2_ab_direct*.log), DIRECT vs TMA ring: K = 1000 two views (acm_2v) 2.14 vs 1.66 G tok/s, K = 1000 one
// view 2.62 vs 1.97, K = 1000 uniform words (HBM-bound) 1.67 vs 1.50, three views (pubmed_3v) 1.93 vs 1.59; K = 500 (four
// documents per warp at 168 registers) 3.19 vs 3.26, (two per warp at 128 registers, 16 warps per SM; compiled, MVTM_DIRECT=1 selects it)
// 3.24 vs 3.18: within 2 %, K <= 512 stays on the ring by default; K = 2000 four views (stress_4v, one document per warp, 12 instead
// of 8 warps per SM) 0.877 vs 0.671 at 200 K documents.
static bool direct_compiled(int KS, int G) { return (KS == 512 && G == 16) || (KS == 1024 && G == 16) || (KS == 2048 && G == 32); }
static bool pick_direct(int KS, bool multi, unsigned flags)
{
    (void)multi;
    if (flags & (MVTM_FLAG_Q1_COMPAT | MVTM_FLAG_TMA_RING)) return false;
    bool d = (KS == 1024 || KS == 2048);
    if (const char *e = getenv("MVTM_DIRECT")) d = atoi(e) != 0;
    return d;
}
static int pick_G(int KS, bool multi, bool direct = false)
{
    int g = KS <= 256 ? 8 : (KS <= 1024 ? 16 : 32);       // measured on B200: lda_100k 3.4 (G=16) vs 2.6 G tok/s (G=32)
    if (multi && KS >= 1024) g = 32;                      // measured: acm_2v 1.43 (G=32) vs 1.30 G tok/s (G=16)
    if (direct && KS == 1024) g = 16;                     // DIRECT: 6.5 KB of shared memory per document leave room for two per warp
    if (const char *e = getenv("MVTM_GROUP")) {
        int want = atoi(e);
        if (want == 32 && (KS == 512 || KS == 1024)) g = 32;
        if (want == 16 && (KS == 512 || KS == 1024)) g = 16;
        if (want == 8 && KS == 512) g = 8;
    }
    return g;
}

// ------------------------------------------------------------------------------------------------
extern "C" const char *mvtm_build_info(void)
{
    return "mvtm-b200 sm_100a; k_sweep_view<KS in {128..2048}, lanes per document G in {8,16,32}, single|multi view>: TMA 1-D bulk ring; "
           "k_sweep_view_direct<KS in {1024,2048}>: rows in registers; Philox4x32-10";
}

extern "C" const char *mvtm_last_error(const mvtm_handle *h) { return h ? h->err.c_str() : g_create_err.c_str(); }

extern "C" int mvtm_create(const mvtm_config *cfg, mvtm_handle **out)
{
    if (!cfg || !out) { g_create_err = "mvtm_create: NULL argument"; return MVTM_ERR_ARG; }
    *out = nullptr;
    if (cfg->num_topics < 1 || cfg->num_views < 1 || cfg->num_views > MVTM_MAX_VIEWS || cfg->num_docs < 0 || !cfg->vocab_sizes) {
        g_create_err = "mvtm_create: bad K / M / D / vocab_sizes"; return MVTM_ERR_ARG;
    }
    if (pick_J(cfg->num_topics) == 0) { g_create_err = "mvtm_create: K > 2048 is not supported by this build"; return MVTM_ERR_LIMIT; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("mvtm_create: no CUDA device (") + cudaGetErrorString(e) + "); this engine has no CPU fallback";
        return MVTM_ERR_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_err = "mvtm_create: bad device ordinal"; return MVTM_ERR_ARG; }
    mvtm_handle *h = new mvtm_handle();
    h->K = cfg->num_topics; h->M = cfg->num_views; h->D = cfg->num_docs; h->device = cfg->device;
    h->flags = cfg->flags; h->seed = cfg->seed;
    h->doc_id_base = cfg->doc_id_base; h->doc_id_stride = cfg->doc_id_stride ? cfg->doc_id_stride : 1;
    h->cfg_warps = cfg->warps_per_cta; h->cfg_ring = cfg->ring_depth; h->cfg_ctas = cfg->max_ctas;
    h->Kp = (h->K + 31) / 32 * 32;
    h->J = pick_J(h->K);
    h->KS = h->J * 128;
    h->direct = pick_direct(h->KS, h->M > 1, h->flags);
    h->G = pick_G(h->KS, h->M > 1, h->direct);
    if (h->direct && !direct_compiled(h->KS, h->G)) { h->direct = false; h->G = pick_G(h->KS, h->M > 1, false); }
    memset(&h->stats, 0, sizeof(h->stats));
    h->alpha.assign((size_t)h->M * (h->K + 1), 0.1);                    // S:149-159, M:195-239
    for (int m = 0; m < h->M; m++) {
        if (cfg->vocab_sizes[m] < 1) { g_create_err = "mvtm_create: vocab size < 1"; delete h; return MVTM_ERR_ARG;   /* nothing allocated yet */ }
        h->v[m].V = cfg->vocab_sizes[m];
        h->alphaSum[m] = 0.1 * h->K; h->beta[m] = 0.01; h->betaSum[m] = 0.01 * h->v[m].V; h->gamma[m] = 1.0;
        for (int j = 0; j < h->M; j++) { h->p_a[m][j] = 0.2; h->p_b[m][j] = 1.0; }   // M:1055-1058
    }
    // a failure below goes through the same teardown as mvtm_destroy (every member starts NULL, the destroy calls accept that)
#define CKC(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { g_create_err = std::string(#call) + ": " + cudaGetErrorString(_e); mvtm_destroy(h); return MVTM_ERR_CUDA; } } while (0)
    CKC(cudaSetDevice(h->device));
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, h->device));
    h->num_sms = prop.multiProcessorCount;
    CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto &ev : h->ev) CKC(cudaEventCreate(&ev));
    for (int m = 0; m < MVTM_MAX_VIEWS; m++) {
        CKC(cudaEventCreateWithFlags(&h->ev_done[m], cudaEventDisableTiming));
        CKC(cudaEventCreateWithFlags(&h->ev_ready[m], cudaEventDisableTiming));
    }
    CKC(cudaMalloc(&h->work_counter, sizeof(int)));
    CKC(cudaMalloc(&h->d_stats, 4 * sizeof(unsigned long long)));
    CKC(cudaMalloc(&h->d_bad, sizeof(int)));
#undef CKC
    *out = h;
    return MVTM_OK;
}

static void free_view(ViewDev &v)
{
    // nk / snap_nk are row V of the nwk / snap_nwk allocations (one all-reduce covers table and totals)
    cudaFree(v.doc_off); cudaFree(v.word); cudaFree(v.z); cudaFree(v.present); cudaFree(v.nwk);
    cudaFree(v.nk_snap); cudaFree(v.order); cudaFree(v.ga_tree); cudaFree(v.ga_full); cudaFree(v.ga_one); cudaFree(v.snap_nwk); cudaFree(v.z_stage);
    v = ViewDev();
}

extern "C" int mvtm_destroy(mvtm_handle *h)
{
    if (!h) return MVTM_OK;
    cudaSetDevice(h->device);
    // An overlapped exchange (the caller's all-reduce and mvtm_sum_exchange_finish_async on the caller's stream) and the copy
    // stream of mvtm_sweep_host may still be touching the tables: wait for the whole device, not just the handle's stream.
    cudaDeviceSynchronize();
    cudaGetLastError();
    comm_teardown(h);
    for (int m = 0; m < h->M; m++) free_view(h->v[m]);
    for (auto &ev : h->ev) if (ev) cudaEventDestroy(ev);
    for (int m = 0; m < MVTM_MAX_VIEWS; m++) { if (h->ev_done[m]) cudaEventDestroy(h->ev_done[m]); if (h->ev_ready[m]) cudaEventDestroy(h->ev_ready[m]); }
    cudaFree(h->work_counter); cudaFree(h->d_stats); cudaFree(h->d_bad); cudaFree(h->oc_scratch); cudaFree(h->rbits);
    for (auto &ev : h->host_ev) cudaEventDestroy(ev);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return MVTM_OK;
}

extern "C" int mvtm_row_stride(mvtm_handle *h, int32_t *stride_out)
{
    if (!h || !stride_out) return MVTM_ERR_ARG;
    *stride_out = h->Kp;
    return MVTM_OK;
}

extern "C" int mvtm_scan_layout(mvtm_handle *h, int32_t *lanes_per_doc, int32_t *chunks_per_lane)
{
    if (!h) return MVTM_ERR_ARG;
    if (lanes_per_doc) *lanes_per_doc = h->G;
    if (chunks_per_lane) *chunks_per_lane = h->KS / (4 * h->G);
    return MVTM_OK;
}

extern "C" int mvtm_add_view(mvtm_handle *h, int32_t m, const int64_t *doc_off, const int32_t *word_id, const uint8_t *present)
{
    if (!h) return MVTM_ERR_ARG;
    if (m < 0 || m >= h->M || !doc_off) FAIL(h, MVTM_ERR_ARG, "mvtm_add_view: bad view index or NULL doc_off");
    CK(h, cudaSetDevice(h->device));
    ViewDev &v = h->v[m];
    const long long D = h->D;
    if (doc_off[0] != 0) FAIL(h, MVTM_ERR_ARG, "mvtm_add_view: doc_off[0] must be 0");
    int max_len = 0;
    for (long long d = 0; d < D; d++) {
        long long len = doc_off[d + 1] - doc_off[d];
        if (len < 0) FAIL(h, MVTM_ERR_ARG, "mvtm_add_view: doc_off not monotone at doc %lld", d);
        if (len > 65535) FAIL(h, MVTM_ERR_LIMIT, "mvtm_add_view: doc %lld holds %lld tokens in view %d (limit 65535)", d, len, m);
        if ((h->flags & MVTM_FLAG_Q1_COMPAT) && len > 32767)
            FAIL(h, MVTM_ERR_LIMIT, "mvtm_add_view: doc %lld holds %lld tokens in view %d (limit 32767 with MVTM_FLAG_Q1_COMPAT)", d, len, m);
        max_len = std::max<int>(max_len, (int)len);
    }
    const long long N = doc_off[D];
    if (N > 0 && !word_id) FAIL(h, MVTM_ERR_ARG, "mvtm_add_view: NULL word_id");
    { int V = v.V; free_view(v); v.V = V; }
    v.n_tok = N; v.max_len = max_len;
    v.h_doc_off.assign(doc_off, doc_off + D + 1);
    // work list: documents that hold tokens, longest first (north_star a: bucketed by length and view)
    std::vector<int> order;
    order.reserve((size_t)D);
    std::vector<unsigned char> pres((size_t)std::max<long long>(D, 1));
    v.docs_present = 0;
    for (long long d = 0; d < D; d++) {
        long long len = doc_off[d + 1] - doc_off[d];
        if (len > 0) order.push_back((int)d);
        pres[(size_t)d] = present ? present[d] : (len > 0);
        v.docs_present += pres[(size_t)d];
    }
    if (!(h->flags & MVTM_FLAG_DOC_ORDER))
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return doc_off[a + 1] - doc_off[a] > doc_off[b + 1] - doc_off[b]; });
    v.n_items = (int)order.size();
    // HOST_CHUNKS contiguous document ranges of about equal token count: the upload units of mvtm_sweep_host
    v.chunk_tok_off.assign(1, 0);
    for (int cidx = 1; cidx <= HOST_CHUNKS; cidx++) {
        long long dhi = D;
        if (cidx < HOST_CHUNKS) dhi = std::lower_bound(doc_off, doc_off + D, N * cidx / HOST_CHUNKS) - doc_off;
        v.chunk_tok_off.push_back(std::max<long long>(v.chunk_tok_off.back(), (long long)doc_off[std::min(dhi, D)]));
    }
    const size_t Kp = (size_t)h->Kp;
    CK(h, cudaMalloc(&v.doc_off, (size_t)(D + 1) * 8));
    CK(h, cudaMalloc(&v.word, (size_t)std::max<long long>(N, 1) * 4));
    CK(h, cudaMalloc(&v.z, (size_t)std::max<long long>(N, 1) * 4));
    CK(h, cudaMalloc(&v.present, pres.size()));
    CK(h, cudaMalloc(&v.order, (size_t)std::max<int>(v.n_items, 1) * 4));
    // + KS ints of tail padding: the DIRECT kernel reads KS (>= Kp) ints per row, whatever follows the row weighs 0
    CK(h, cudaMalloc(&v.nwk, (((size_t)v.V + 1) * Kp + (size_t)h->KS) * 4));
    v.nk = v.nwk + (size_t)v.V * Kp;
    CK(h, cudaMalloc(&v.nk_snap, Kp * 4));
    CK(h, cudaMalloc(&v.ga_tree, Kp * 4));
    CK(h, cudaMalloc(&v.ga_full, Kp * 4));
    CK(h, cudaMalloc(&v.ga_one, Kp * 4));
    { std::vector<float> ones(Kp, 0.f); std::fill(ones.begin(), ones.begin() + h->K, 1.f); CK(h, cudaMemcpy(v.ga_one, ones.data(), Kp * 4, cudaMemcpyHostToDevice)); }
    CK(h, cudaMemcpy(v.doc_off, doc_off, (size_t)(D + 1) * 8, cudaMemcpyHostToDevice));
    if (N > 0) CK(h, cudaMemcpy(v.word, word_id, (size_t)N * 4, cudaMemcpyHostToDevice));
    CK(h, cudaMemcpy(v.present, pres.data(), pres.size(), cudaMemcpyHostToDevice));
    v.h_present = pres;
    if (v.n_items) CK(h, cudaMemcpy(v.order, order.data(), (size_t)v.n_items * 4, cudaMemcpyHostToDevice));
    CK(h, cudaMemset(v.z, 0xff, (size_t)std::max<long long>(N, 1) * 4));       // UNASSIGNED_TOPIC
    CK(h, cudaMemset(v.nwk, 0, (size_t)v.V * Kp * 4));
    CK(h, cudaMemset(v.nk, 0, Kp * 4));
    v.added = true;
    h->hyper_dirty = true;
    return MVTM_OK;
}

static int require_views(mvtm_handle *h, const char *who)
{
    for (int m = 0; m < h->M; m++) if (!h->v[m].added) FAIL(h, MVTM_ERR_STATE, "%s: view %d has not been added", who, m);
    return MVTM_OK;
}

static int rebuild_counts_view(mvtm_handle *h, int m)
{
    ViewDev &v = h->v[m];
    const size_t Kp = (size_t)h->Kp;
    h->mut_epoch++;
    if (int rc = wait_view_ready(h, m)) return rc;
    CK(h, cudaMemsetAsync(v.nwk, 0, (size_t)v.V * Kp * 4, h->stream));
    CK(h, cudaMemsetAsync(v.nk, 0, Kp * 4, h->stream));
    CK(h, cudaMemsetAsync(h->d_bad, 0, sizeof(int), h->stream));
    if (v.n_tok > 0) {
        int blocks = (int)std::min<long long>((v.n_tok + 255) / 256, (long long)h->num_sms * 8);
        k_build_counts<<<blocks, 256, (size_t)h->K * 4, h->stream>>>(v.n_tok, v.word, v.z, v.V, h->K, h->Kp, v.nwk, v.nk, h->d_bad, 1);
        CK(h, cudaGetLastError());
    }
    int bad = 0;
    CK(h, cudaMemcpyAsync(&bad, h->d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (bad) FAIL(h, MVTM_ERR_ARG, "assignments of view %d hold %d topic ids >= K", m, bad);
    return MVTM_OK;
}

extern "C" int mvtm_init_assignments(mvtm_handle *h)
{
    if (!h) return MVTM_ERR_ARG;
    if (int rc = require_views(h, "mvtm_init_assignments")) return rc;
    CK(h, cudaSetDevice(h->device));
    if (int rc = wait_all_ready(h)) return rc;
    for (int m = 0; m < h->M; m++) {
        ViewDev &v = h->v[m];
        if (v.n_tok > 0) {
            int blocks = (int)std::min<long long>((h->D * 32 + 255) / 256, (long long)h->num_sms * 16);
            k_init_assign<<<std::max(blocks, 1), 256, 0, h->stream>>>(m, h->K, h->D, v.doc_off, h->v[0].doc_off, h->v[0].z, v.z,
                                                                     (unsigned)h->seed, (unsigned)(h->seed >> 32), h->doc_id_base, h->doc_id_stride);
            CK(h, cudaGetLastError());
        }
        if (int rc = rebuild_counts_view(h, m)) return rc;
    }
    return MVTM_OK;
}

extern "C" int mvtm_set_assignments(mvtm_handle *h, int32_t m, const int32_t *z)
{
    if (!h) return MVTM_ERR_ARG;
    if (m < 0 || m >= h->M || !h->v[m].added) FAIL(h, MVTM_ERR_ARG, "mvtm_set_assignments: bad view %d", m);
    CK(h, cudaSetDevice(h->device));
    ViewDev &v = h->v[m];
    if (v.n_tok > 0) {
        if (!z) FAIL(h, MVTM_ERR_ARG, "mvtm_set_assignments: NULL z");
        CK(h, cudaMemcpyAsync(v.z, z, (size_t)v.n_tok * 4, cudaMemcpyHostToDevice, h->stream));
    }
    return rebuild_counts_view(h, m);
}

extern "C" int mvtm_set_counts(mvtm_handle *h, int32_t m, const int32_t *n_wk, const int32_t *n_k)
{
    if (!h) return MVTM_ERR_ARG;
    if (m < 0 || m >= h->M || !h->v[m].added || !n_wk || !n_k) FAIL(h, MVTM_ERR_ARG, "mvtm_set_counts: bad argument");
    CK(h, cudaSetDevice(h->device));
    if (int rc = wait_all_ready(h)) return rc;
    h->mut_epoch++;
    ViewDev &v = h->v[m];
    for (int t = 0; t < h->K; t++) if (n_k[t] < 0) FAIL(h, MVTM_ERR_CORRUPT, "mvtm_set_counts: negative n_k[%d]", t);
    CK(h, cudaMemsetAsync(v.nwk, 0, (size_t)v.V * h->Kp * 4, h->stream));
    CK(h, cudaMemsetAsync(v.nk, 0, (size_t)h->Kp * 4, h->stream));
    CK(h, cudaMemcpy2DAsync(v.nwk, (size_t)h->Kp * 4, n_wk, (size_t)h->K * 4, (size_t)h->K * 4, (size_t)v.V, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(v.nk, n_k, (size_t)h->K * 4, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return MVTM_OK;
}

extern "C" int mvtm_init_assignments_from_counts(mvtm_handle *h)
{
    if (!h) return MVTM_ERR_ARG;
    if (int rc = require_views(h, "mvtm_init_assignments_from_counts")) return rc;
    CK(h, cudaSetDevice(h->device));
    if (int rc = wait_all_ready(h)) return rc;
    h->mut_epoch++;
    const int K = h->K;
    int P2 = 1; while (P2 * 2 <= 2 * K - 1) P2 *= 2;                  // 2^floor(log2(2K-1)): first index of the deepest tree level
    const int rot = P2 - K;
    for (int m = 0; m < h->M; m++) {
        ViewDev &v = h->v[m];
        if (v.n_tok == 0) continue;
        int blocks = (int)std::min<long long>((h->D * 32 + 255) / 256, (long long)h->num_sms * 16);
        k_init_from_phi<<<std::max(blocks, 1), 256, 0, h->stream>>>(m, K, h->Kp, v.V, h->D, v.doc_off, v.word, v.z, v.nwk, v.nk, h->beta[m], h->betaSum[m],
                                                                   rot, (unsigned)h->seed, (unsigned)(h->seed >> 32), h->doc_id_base, h->doc_id_stride);
        CK(h, cudaGetLastError());
    }
    CK(h, cudaStreamSynchronize(h->stream));
    return MVTM_OK;
}

extern "C" int mvtm_get_assignments(mvtm_handle *h, int32_t m, int32_t *z_out)
{
    if (!h) return MVTM_ERR_ARG;
    if (m < 0 || m >= h->M || !h->v[m].added) FAIL(h, MVTM_ERR_ARG, "mvtm_get_assignments: bad view %d", m);
    CK(h, cudaSetDevice(h->device));
    if (h->v[m].n_tok > 0) {
        if (!z_out) FAIL(h, MVTM_ERR_ARG, "mvtm_get_assignments: NULL buffer");
        CK(h, cudaMemcpyAsync(z_out, h->v[m].z, (size_t)h->v[m].n_tok * 4, cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
    }
    return MVTM_OK;
}

extern "C" int mvtm_get_counts(mvtm_handle *h, int32_t m, int32_t *n_wk_out, int32_t *n_k_out)
{
    if (!h) return MVTM_ERR_ARG;
    if (m < 0 || m >= h->M || !h->v[m].added) FAIL(h, MVTM_ERR_ARG, "mvtm_get_counts: bad view %d", m);
    CK(h, cudaSetDevice(h->device));
    if (int rc = wait_view_ready(h, m)) return rc;
    ViewDev &v = h->v[m];
    if (n_wk_out)
        CK(h, cudaMemcpy2DAsync(n_wk_out, (size_t)h->K * 4, v.nwk, (size_t)h->Kp * 4, (size_t)h->K * 4, (size_t)v.V, cudaMemcpyDeviceToHost, h->stream));
    if (n_k_out) CK(h, cudaMemcpyAsync(n_k_out, v.nk, (size_t)h->K * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return MVTM_OK;
}

extern "C" int mvtm_set_hyper(mvtm_handle *h, const double *alpha, const double *alpha_sum, const double *beta, const double *beta_sum,
                              const double *gamma, const double *p_a, const double *p_b, const int32_t *inactive, int32_t n_inactive)
{
    if (!h) return MVTM_ERR_ARG;
    const int M = h->M, K = h->K;
    if (n_inactive > K || (n_inactive > 0 && !inactive)) FAIL(h, MVTM_ERR_ARG, "mvtm_set_hyper: bad inactive list");
    if (alpha) {
        for (size_t i = 0; i < (size_t)M * (K + 1); i++) if (!(alpha[i] >= 0.0) || !std::isfinite(alpha[i])) FAIL(h, MVTM_ERR_ARG, "mvtm_set_hyper: alpha[%zu] not finite/non-negative", i);
        h->alpha.assign(alpha, alpha + (size_t)M * (K + 1));
    }
    for (int m = 0; m < M; m++) {
        if (alpha_sum) h->alphaSum[m] = alpha_sum[m];
        if (beta) { if (!(beta[m] > 0.0)) FAIL(h, MVTM_ERR_ARG, "mvtm_set_hyper: beta[%d] must be > 0", m); h->beta[m] = beta[m]; }
        if (beta_sum) h->betaSum[m] = beta_sum[m];
        if (gamma) h->gamma[m] = gamma[m];
        for (int j = 0; j < M; j++) {
            if (p_a) h->p_a[m][j] = p_a[m * M + j];
            if (p_b) h->p_b[m][j] = p_b[m * M + j];
        }
    }
    if (n_inactive >= 0) {
        for (int i = 0; i < n_inactive; i++) if (inactive[i] < 0 || inactive[i] >= K) FAIL(h, MVTM_ERR_ARG, "mvtm_set_hyper: inactive topic %d out of range", inactive[i]);
        h->inactive.assign(inactive, inactive + n_inactive);
        std::sort(h->inactive.begin(), h->inactive.end());
        h->inactive.erase(std::unique(h->inactive.begin(), h->inactive.end()), h->inactive.end());
    }
    h->hyper_dirty = true;
    return MVTM_OK;
}

extern "C" int mvtm_get_hyper(mvtm_handle *h, double *alpha, double *alpha_sum, int32_t *inactive, int32_t *n_inactive)
{
    if (!h) return MVTM_ERR_ARG;
    if (alpha) memcpy(alpha, h->alpha.data(), h->alpha.size() * 8);
    if (alpha_sum) for (int m = 0; m < h->M; m++) alpha_sum[m] = h->alphaSum[m];
    if (inactive) for (size_t i = 0; i < h->inactive.size(); i++) inactive[i] = h->inactive[i];
    if (n_inactive) *n_inactive = (int)h->inactive.size();
    return MVTM_OK;
}

static int upload_hyper(mvtm_handle *h)
{
    if (!h->hyper_dirty) return MVTM_OK;
    const int K = h->K, Kp = h->Kp;
    std::vector<float> tree((size_t)Kp), full((size_t)Kp);
    for (int m = 0; m < h->M; m++) {
        if (!h->v[m].added) continue;
        std::fill(tree.begin(), tree.end(), 0.f); std::fill(full.begin(), full.end(), 0.f);
        for (int t = 0; t < K; t++) {
            double ga = h->gamma[m] * h->alpha[(size_t)m * (K + 1) + t];
            full[t] = (float)ga;                                         // W:404
            tree[t] = (float)ga;                                         // M:2678
        }
        for (int t : h->inactive) tree[t] = 0.f;                         // M:2670-2671
        CK(h, cudaMemcpyAsync(h->v[m].ga_tree, tree.data(), (size_t)Kp * 4, cudaMemcpyHostToDevice, h->stream));
        CK(h, cudaMemcpyAsync(h->v[m].ga_full, full.data(), (size_t)Kp * 4, cudaMemcpyHostToDevice, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
    }
    h->hyper_dirty = false;
    return MVTM_OK;
}

static void fill_params(mvtm_handle *h, int m, int iteration, int update_global, SweepParams &P)
{
    memset(&P, 0, sizeof(P));
    ViewDev &v = h->v[m];
    P.M = h->M; P.K = h->K; P.Kp = h->Kp; P.m = m; P.V = v.V;
    P.n_items = v.n_items; P.order = v.order; P.work_counter = h->work_counter;
    for (int i = 0; i < h->M; i++) {
        P.doc_off[i] = h->v[i].doc_off; P.zv[i] = h->v[i].z; P.ga_full[i] = h->v[i].ga_full;
        P.gas[i] = (float)(h->gamma[i] * h->alphaSum[i]);
        P.ga_new[i] = (float)(h->gamma[i] * h->alpha[(size_t)i * (h->K + 1) + h->K]);
        P.pa[i] = h->p_a[m][i]; P.pb[i] = h->p_b[m][i];
        P.sparse[i] = (h->beta[i] == 0.0001);                            // W:335-336
    }
    P.word = v.word; P.nwk = v.nwk; P.nk_frozen = v.nk_snap; P.nk_live = v.nk;
    P.ga_tree = (update_global == 2) ? v.ga_one : v.ga_tree;            // 2 = inferencer mode with bare-phi trees (I:561-576, Q13)
    P.beta = (float)h->beta[m]; P.betaSum = (float)h->betaSum[m];
    P.n_inactive = (int)h->inactive.size(); P.first_inactive = h->inactive.empty() ? 0 : h->inactive[0];
    if (update_global == 2) { update_global = 0; P.n_inactive = 0; P.first_inactive = 0; }   // I:243: the inferencer's inactive set is empty
    P.beta_mallet = (h->flags & MVTM_FLAG_BETA_MALLET) ? 1 : 0;
    P.seed_lo = (unsigned)h->seed; P.seed_hi = (unsigned)(h->seed >> 32); P.iteration = (unsigned)iteration;
    P.doc_id_base = h->doc_id_base; P.doc_id_stride = h->doc_id_stride;
    P.update_global = update_global;
    P.stats = h->d_stats;
    P.oc_scratch = h->oc_scratch;
    P.rbits = h->rbits; P.rb_one = nullptr;
}

static int ensure_oc_scratch(mvtm_handle *h, size_t doc_slots)
{
    if (h->M < 2) return MVTM_OK;
    const size_t need = doc_slots * (size_t)h->Kp;
    if (need <= h->oc_scratch_floats) return MVTM_OK;
    CK(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->oc_scratch); h->oc_scratch = nullptr; h->oc_scratch_floats = 0;
    CK(h, cudaMalloc(&h->oc_scratch, need * sizeof(float)));
    h->oc_scratch_floats = need;
    return MVTM_OK;
}

struct LaunchCfg { int R, W, grid, oc_smem; size_t smem; bool direct; };

// Ring depth of view m.  Unless fixed by the caller (mvtm_config.ring_depth / MVTM_RING), the first 2*RING_SAMPLES timed
// passes of the view alternate R = 1, 2, 1, 2, ... and the depth with the smaller MEDIAN pass time is kept (one sample per depth
// locked different depths on identical runs: BENCH_r01 vs SCALE_r01 N=1).  Cache-resident tables (Zipf corpora) favour R = 1
// (more resident documents), genuinely HBM-bound ones R = 2 (the row fetch latency is then worth a slot).  A depth that is
// within 2 % of the other is not worth a different launch shape: R = 1 wins ties.  The locked depth is reported in mvtm_stats.
constexpr int RING_SAMPLES = 3;
static int ring_for_view(mvtm_handle *h, int m)
{
    if (h->cfg_ring > 0) return h->cfg_ring;
    if (const char *e = getenv("MVTM_RING")) return atoi(e);
    ViewDev &v = h->v[m];
    if (v.ring_locked) return v.ring_locked;
    return (v.tune_step & 1) ? 2 : 1;
}
static void ring_record(mvtm_handle *h, int m, float ms)
{
    ViewDev &v = h->v[m];
    if (h->direct || h->cfg_ring > 0 || getenv("MVTM_RING") || v.ring_locked) return;
    v.tune_ms[v.tune_step] = ms;
    if (++v.tune_step == 2 * RING_SAMPLES) {
        float a[RING_SAMPLES], b[RING_SAMPLES];
        for (int i = 0; i < RING_SAMPLES; i++) { a[i] = v.tune_ms[2 * i]; b[i] = v.tune_ms[2 * i + 1]; }
        std::sort(a, a + RING_SAMPLES); std::sort(b, b + RING_SAMPLES);
        v.ring_locked = (b[RING_SAMPLES / 2] < 0.98f * a[RING_SAMPLES / 2]) ? 2 : 1;
    }
}

static int choose_launch(mvtm_handle *h, int m, int R, LaunchCfg &lc)
{
    const int KS = h->KS, G = h->G, NSUB = 32 / G;
    const bool multi = h->M > 1;
    // The other-view mass can live in shared memory instead of the global scratch (MVTM_OC_SMEM=1).  Measured on
    // pubmed_3v / acm_2v: the extra 4*KS bytes per document cost more resident documents than the faster reads give
    // back, even for the side views (5.7 vs 4.8 ms), so the default keeps it in global memory for every view.
    (void)m;
    const bool oc_smem = multi && !h->direct && getenv("MVTM_OC_SMEM") && atoi(getenv("MVTM_OC_SMEM")) != 0;
    R = h->direct ? 0 : std::max(1, std::min(R, 8));
    const size_t budget = 227 * 1024 - 1024;
    int docs;
    for (;;) {
        docs = (int)((budget - smem_cta_bytes(KS, multi ? h->M : 0)) / smem_doc_bytes(KS, R, multi, oc_smem));
        if (docs >= NSUB || R <= 1) break;
        R--;
    }
    if (docs < NSUB) FAIL(h, MVTM_ERR_LIMIT, "shared memory cannot hold one warp's state for K=%d", h->K);
    int W = docs / NSUB;
    W = std::min(W, sweep_max_threads(KS, G, multi, h->direct) / 32);
    if (h->cfg_warps > 0) W = std::min(W, h->cfg_warps);
    if (const char *e = getenv("MVTM_WARPS")) W = std::max(1, std::min(W, atoi(e)));
    int grid = h->num_sms;
    if (h->cfg_ctas > 0) grid = std::min(grid, h->cfg_ctas);
    if (const char *e = getenv("MVTM_CTAS")) grid = std::max(1, std::min(grid, atoi(e)));
    if (h->flags & MVTM_FLAG_SINGLE_WARP) { W = 1; grid = 1; }
    lc.R = R; lc.W = W; lc.grid = grid; lc.oc_smem = oc_smem ? 1 : 0; lc.direct = h->direct;
    lc.smem = smem_cta_bytes(KS, multi ? h->M : 0) + (size_t)W * NSUB * smem_doc_bytes(KS, R, multi, oc_smem);
    return MVTM_OK;
}

template <int KS, int G, bool MULTI>
static cudaError_t launch_sweep_t(const SweepParams &P, const LaunchCfg &lc, cudaStream_t s, bool q1)
{
    auto go = [&](auto kernel) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem);
        if (e != cudaSuccess) return e;
        kernel<<<lc.grid, lc.W * 32, lc.smem, s>>>(P);
        return cudaGetLastError();
    };
    if constexpr ((KS == 512 && G == 16) || (KS == 1024 && G == 16) || (KS == 2048 && G == 32)) {   // = direct_compiled
        if (lc.direct && !q1) return go(k_sweep_view_direct<KS, G, MULTI>);
    }
    if (lc.direct) return cudaErrorInvalidValue;
    return q1 ? go(k_sweep_view<KS, G, MULTI, true>) : go(k_sweep_view<KS, G, MULTI, false>);
}
template <int KS, int G, bool MULTI>
static cudaError_t launch_probe_t(const SweepParams &P, int d, int pos, const double *p_row, double *out, cudaStream_t s, bool q1)
{
    size_t smem = smem_cta_bytes(KS, MULTI ? P.M : 0) + (size_t)(32 / G) * smem_doc_bytes(KS, 1, MULTI);
    auto go = [&](auto kernel) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<1, 32, smem, s>>>(P, d, pos, p_row, out);
        return cudaGetLastError();
    };
    return q1 ? go(k_cond_probe<KS, G, MULTI, true>) : go(k_cond_probe<KS, G, MULTI, false>);
}

#define DISPATCH_KG(KS_, G_, MULTI_, CALL)                                                        \
    switch ((KS_) * 64 + (G_)) {                                                                  \
        case 128 * 64 + 8: CALL(128, 8, MULTI_); break;    case 256 * 64 + 8: CALL(256, 8, MULTI_); break;      \
        case 384 * 64 + 16: CALL(384, 16, MULTI_); break;  case 512 * 64 + 16: CALL(512, 16, MULTI_); break;    \
        case 512 * 64 + 32: CALL(512, 32, MULTI_); break;  case 768 * 64 + 16: CALL(768, 16, MULTI_); break;    \
        case 512 * 64 + 8: CALL(512, 8, MULTI_); break;                                                         \
        case 1024 * 64 + 16: CALL(1024, 16, MULTI_); break; case 1024 * 64 + 32: CALL(1024, 32, MULTI_); break; \
        case 1536 * 64 + 32: CALL(1536, 32, MULTI_); break; case 2048 * 64 + 32: CALL(2048, 32, MULTI_); break; \
        default: break;                                                                           \
    }

static cudaError_t launch_sweep(mvtm_handle *h, const SweepParams &P, const LaunchCfg &lc)
{
    cudaError_t e = cudaErrorInvalidValue;
    h->mut_epoch++;
#define CALL_SWEEP(KS_, G_, MU_) e = launch_sweep_t<KS_, G_, MU_>(P, lc, h->stream, (h->flags & MVTM_FLAG_Q1_COMPAT) != 0)
    if (h->M > 1) { DISPATCH_KG(h->KS, h->G, true, CALL_SWEEP) } else { DISPATCH_KG(h->KS, h->G, false, CALL_SWEEP) }
#undef CALL_SWEEP
    return e;
}

// activation of inactive topics that received tokens during the sweep (U:263-270 at sweep granularity)
static int activate_sampled_topics(mvtm_handle *h)
{
    if (h->inactive.empty()) return MVTM_OK;
    const int K = h->K;
    std::vector<std::vector<int>> nk((size_t)h->M, std::vector<int>((size_t)K));
    // on the handle's stream, behind anything a caller's stream still owes the views (the legacy default stream does not order
    // against these non-blocking streams)
    if (int rc = wait_all_ready(h)) return rc;
    for (int m = 0; m < h->M; m++) CK(h, cudaMemcpyAsync(nk[m].data(), h->v[m].nk, (size_t)K * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    std::vector<int> keep;
    for (int t : h->inactive) {
        bool hit = false;
        for (int m = 0; m < h->M; m++)
            if (nk[m][t] > 0) { h->alpha[(size_t)m * (K + 1) + t] = h->alpha[(size_t)m * (K + 1) + K]; hit = true; }   // Q16
        if (!hit) keep.push_back(t);
    }
    if (keep.size() != h->inactive.size()) { h->inactive.swap(keep); h->hyper_dirty = true; }
    return MVTM_OK;
}

// A view's tables may be owned for a while by work the CALLER queued on a stream of its own (the overlapped count exchange
// of a multi-GPU run, mvtm_view_wait_stream): everything here that touches them first orders the handle's stream behind it.
static int wait_view_ready(mvtm_handle *h, int m)
{
    if (h->ready_pending[m]) {
        CK(h, cudaStreamWaitEvent(h->stream, h->ev_ready[m], 0));
        h->ready_pending[m] = false;
    }
    return MVTM_OK;
}
static int wait_all_ready(mvtm_handle *h)
{
    for (int m = 0; m < h->M; m++) if (int rc = wait_view_ready(h, m)) return rc;
    return MVTM_OK;
}

// MVTM_FLAG_Q1_COMPAT: the reference rebuilds a document's dense index at the start of every sweep (W:376-391), so the "not in the
// index" flags that multi-view passes hand to each other start from zero each sweep.
static int clear_q1_flags(mvtm_handle *h)
{
    if (!(h->flags & MVTM_FLAG_Q1_COMPAT) || h->M < 2 || h->D == 0) return MVTM_OK;
    const size_t bytes = (size_t)h->D * (h->KS / 32) * 4;
    if (!h->rbits) CK(h, cudaMalloc(&h->rbits, bytes));
    CK(h, cudaMemsetAsync(h->rbits, 0, bytes, h->stream));
    return MVTM_OK;
}

// Queues view m's pass (n_k snapshot, work counter reset, k_sweep_view) on the handle's stream; no host synchronisation.
static int enqueue_view_pass(mvtm_handle *h, int iteration, int update_global, int m, int *launches)
{
    ViewDev &v = h->v[m];
    LaunchCfg lc;
    if (int rc = choose_launch(h, m, ring_for_view(h, m), lc)) return rc;
    if (int rc = ensure_oc_scratch(h, (size_t)lc.grid * lc.W * (32 / h->G))) return rc;
    if (int rc = wait_view_ready(h, m)) return rc;
    CK(h, cudaEventRecord(h->ev[2 + 2 * m], h->stream));
    if (v.n_items > 0) {
        CK(h, cudaMemcpyAsync(v.nk_snap, v.nk, (size_t)h->Kp * 4, cudaMemcpyDeviceToDevice, h->stream));
        CK(h, cudaMemsetAsync(h->work_counter, 0, sizeof(int), h->stream));
        SweepParams P;
        fill_params(h, m, iteration, update_global, P);
        P.R = lc.R; P.oc_smem = lc.oc_smem;
        h->stats.ring_depth[m] = lc.R;
        P.z_host = v.z_host_once ? v.z_host_once : v.z_mirror;
        CK(h, launch_sweep(h, P, lc));
        (*launches)++;
    }
    CK(h, cudaEventRecord(h->ev[3 + 2 * m], h->stream));
    h->pass_queued[m] = true;
    return MVTM_OK;
}

static int open_sweep(mvtm_handle *h)
{
    if (h->sweep_open) return MVTM_OK;
    if (int rc = require_views(h, "mvtm_sweep")) return rc;
    CK(h, cudaSetDevice(h->device));
    if (int rc = upload_hyper(h)) return rc;
    if (int rc = clear_q1_flags(h)) return rc;
    CK(h, cudaMemsetAsync(h->d_stats, 0, 4 * sizeof(unsigned long long), h->stream));
    CK(h, cudaEventRecord(h->ev[0], h->stream));
    for (int m = 0; m < h->M; m++) h->pass_queued[m] = false;
    h->open_launches = 0;
    h->sweep_open = true;
    return MVTM_OK;
}

// Host side of the barrier M:1231: waits for the queued passes, reads the counters and the per-view device times.
static int close_sweep(mvtm_handle *h, int update_global)
{
    if (!h->sweep_open) return MVTM_OK;
    h->sweep_open = false;
    CK(h, cudaEventRecord(h->ev[1], h->stream));
    unsigned long long st[4];
    CK(h, cudaMemcpyAsync(st, h->d_stats, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    h->stats.tokens = (long long)st[0]; h->stats.changed = (long long)st[1]; h->stats.new_topic = (long long)st[2];
    float ms = 0.f;
    CK(h, cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    h->stats.ms_total = ms;
    for (int m = 0; m < h->M; m++) {
        h->stats.ms_view[m] = 0.0;
        if (!h->pass_queued[m]) continue;
        CK(h, cudaEventElapsedTime(&ms, h->ev[2 + 2 * m], h->ev[3 + 2 * m]));
        h->stats.ms_view[m] = ms;
        if (h->v[m].n_items > 0) ring_record(h, m, ms);
    }
    h->stats.kernel_launches = h->open_launches;
    // multi-rank runs (a statistics reducer is installed) activate on the GLOBAL counts after the exchange: mvtm_activate_topics
    if (update_global == 1 && !h->reducer && !h->comm) if (int rc = activate_sampled_topics(h)) return rc;
    return MVTM_OK;
}

extern "C" int mvtm_activate_topics(mvtm_handle *h)
{
    if (!h) return MVTM_ERR_ARG;
    if (h->inactive.empty()) return MVTM_OK;
    CK(h, cudaSetDevice(h->device));
    if (int rc = wait_all_ready(h)) return rc;
    CK(h, cudaStreamSynchronize(h->stream));
    return activate_sampled_topics(h);
}

static int sweep_impl(mvtm_handle *h, int iteration, int update_global, bool sync_stats)
{
    if (h->sweep_open) FAIL(h, MVTM_ERR_STATE, "mvtm_sweep: passes queued by mvtm_sweep_view_async are still open (call mvtm_sweep_finish)");
    if (int rc = open_sweep(h)) return rc;
    for (int m = 0; m < h->M; m++) {
        if (int rc = enqueue_view_pass(h, iteration, update_global, m, &h->open_launches)) { h->sweep_open = false; return rc; }
        // The reference's updater activates a sampled inactive topic per delta (U:263-270).  The blocking single-handle sweep does
        // it after every VIEW PASS rather than once per sweep, so the following views of the same sweep already see the topic as
        // active (one host synchronisation per view, paid only while some topic is inactive, i.e. after an optimizeDP step).
        if (update_global == 1 && sync_stats && m + 1 < h->M && !h->inactive.empty() && !h->reducer && !h->comm) {
            if (int rc = activate_sampled_topics(h)) { h->sweep_open = false; return rc; }
            if (int rc = upload_hyper(h)) { h->sweep_open = false; return rc; }
        }
    }
    if (sync_stats) return close_sweep(h, update_global);
    h->sweep_open = false;
    CK(h, cudaEventRecord(h->ev[1], h->stream));
    return MVTM_OK;
}

extern "C" int mvtm_sweep(mvtm_handle *h, int32_t iteration, int32_t update_global)
{
    if (!h) return MVTM_ERR_ARG;
    return sweep_impl(h, iteration, update_global, true);
}

// ---- non-blocking form: one view pass at a time, so a multi-GPU caller can overlap the count exchange of view m with the
// sampling of the following views ------------------------------------------------------------------------------
extern "C" int mvtm_sweep_view_async(mvtm_handle *h, int32_t iteration, int32_t m, int32_t update_global)
{
    if (!h) return MVTM_ERR_ARG;
    if (m < 0 || m >= h->M) FAIL(h, MVTM_ERR_ARG, "mvtm_sweep_view_async: bad view %d", m);
    if (update_global < 0 || update_global > 2) FAIL(h, MVTM_ERR_ARG, "mvtm_sweep_view_async: bad update_global");
    if (int rc = open_sweep(h)) return rc;
    h->open_mode = update_global;
    return enqueue_view_pass(h, iteration, update_global, m, &h->open_launches);
}

extern "C" int mvtm_sweep_finish(mvtm_handle *h)
{
    if (!h) return MVTM_ERR_ARG;
    CK(h, cudaSetDevice(h->device));
    return close_sweep(h, h->open_mode);
}

extern "C" int mvtm_stream_wait_view(mvtm_handle *h, int32_t m, void *stream)
{
    if (!h) return MVTM_ERR_ARG;
    if (m < 0 || m >= h->M || !h->v[m].added) FAIL(h, MVTM_ERR_ARG, "mvtm_stream_wait_view: bad view %d", m);
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaEventRecord(h->ev_done[m], h->stream));
    CK(h, cudaStreamWaitEvent((cudaStream_t)stream, h->ev_done[m], 0));
    return MVTM_OK;
}

extern "C" int mvtm_view_wait_stream(mvtm_handle *h, int32_t m, void *stream)
{
    if (!h) return MVTM_ERR_ARG;
    if (m < 0 || m >= h->M || !h->v[m].added) FAIL(h, MVTM_ERR_ARG, "mvtm_view_wait_stream: bad view %d", m);
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaEventRecord(h->ev_ready[m], (cudaStream_t)stream));
    h->ready_pending[m] = true;
    return MVTM_OK;
}

// device-visible alias of a caller buffer when it is pinned + mapped (cudaHostAlloc / cudaHostRegister under UVA), else NULL
static int *mapped_alias(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
    return (int *)a.devicePointer;
}

extern "C" int mvtm_set_host_mirror(mvtm_handle *h, int32_t m, int32_t *z_host)
{
    if (!h) return MVTM_ERR_ARG;
    if (m < 0 || m >= h->M || !h->v[m].added) FAIL(h, MVTM_ERR_ARG, "mvtm_set_host_mirror: bad view %d", m);
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    if (!z_host) { h->v[m].z_mirror = nullptr; return MVTM_OK; }
    int *alias = mapped_alias(z_host);
    if (!alias) FAIL(h, MVTM_ERR_ARG, "mvtm_set_host_mirror: the array is not pinned + mapped host memory (cudaHostAlloc / cudaHostRegister)");
    h->v[m].z_mirror = alias;
    return MVTM_OK;
}

extern "C" int mvtm_sweep_host(mvtm_handle *h, int32_t iteration, int32_t *const *z_inout)
{
    // Upload: HOST_CHUNKS contiguous document ranges per view on copy_stream, each range added to the counts on `stream`
    // while the next one travels (the sampler needs ALL counts, so it starts after the last range).
    // Download: when the caller's arrays are pinned and mapped, the sweep kernel itself stores every token block's new
    // assignments to them next to its store to device memory (posted PCIe writes under the sampling, nothing left to copy
    // afterwards); pageable arrays get a plain copy after the pass.  One launch per view, exactly as mvtm_sweep.
    if (!h) return MVTM_ERR_ARG;
    if (!z_inout) FAIL(h, MVTM_ERR_ARG, "mvtm_sweep_host: NULL z_inout");
    if (h->sweep_open) FAIL(h, MVTM_ERR_STATE, "mvtm_sweep_host: passes queued by mvtm_sweep_view_async are still open");
    if (int rc = require_views(h, "mvtm_sweep_host")) return rc;
    CK(h, cudaSetDevice(h->device));
    if (int rc = wait_all_ready(h)) return rc;
    if (!h->copy_stream) CK(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    const size_t n_ev = (size_t)MVTM_MAX_VIEWS * HOST_CHUNKS;
    while (h->host_ev.size() < n_ev) { cudaEvent_t e; CK(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->host_ev.push_back(e); }
    auto ev_up = [&](int m, int cidx) { return h->host_ev[(size_t)(m * HOST_CHUNKS + cidx)]; };
    const size_t Kp = (size_t)h->Kp;
    int *alias[MVTM_MAX_VIEWS];
    const bool zero_copy = !(getenv("MVTM_HOST_ZEROCOPY") && atoi(getenv("MVTM_HOST_ZEROCOPY")) == 0);
    for (int m = 0; m < h->M; m++) {
        if (h->v[m].n_tok > 0 && !z_inout[m]) FAIL(h, MVTM_ERR_ARG, "mvtm_sweep_host: NULL z for view %d", m);
        alias[m] = (zero_copy && h->v[m].n_tok > 0) ? mapped_alias(z_inout[m]) : nullptr;
    }
    if (int rc = upload_hyper(h)) return rc;
    // the copy stream must not overtake earlier work on the handle's stream that still reads z
    CK(h, cudaEventRecord(h->ev_done[0], h->stream));
    CK(h, cudaStreamWaitEvent(h->copy_stream, h->ev_done[0], 0));
    CK(h, cudaMemsetAsync(h->d_bad, 0, sizeof(int), h->stream));
    for (int m = 0; m < h->M; m++) {
        ViewDev &v = h->v[m];
        CK(h, cudaMemsetAsync(v.nwk, 0, ((size_t)v.V + 1) * Kp * 4, h->stream));        // table and totals
        for (int cidx = 0; cidx < HOST_CHUNKS; cidx++) {
            const long long t0 = v.chunk_tok_off[cidx], n = v.chunk_tok_off[cidx + 1] - t0;
            if (n <= 0) continue;
            CK(h, cudaMemcpyAsync(v.z + t0, z_inout[m] + t0, (size_t)n * 4, cudaMemcpyHostToDevice, h->copy_stream));
            CK(h, cudaEventRecord(ev_up(m, cidx), h->copy_stream));
            CK(h, cudaStreamWaitEvent(h->stream, ev_up(m, cidx), 0));
            int blocks = (int)std::min<long long>((n + 255) / 256, (long long)h->num_sms * 8);
            k_build_counts<<<blocks, 256, (size_t)h->K * 4, h->stream>>>(n, v.word + t0, v.z + t0, v.V, h->K, h->Kp, v.nwk, v.nk, h->d_bad, 1);
            CK(h, cudaGetLastError());
        }
    }
    if (int rc = clear_q1_flags(h)) return rc;
    CK(h, cudaMemsetAsync(h->d_stats, 0, 4 * sizeof(unsigned long long), h->stream));
    CK(h, cudaEventRecord(h->ev[0], h->stream));
    int launches = 0;
    for (int m = 0; m < h->M; m++) {
        ViewDev &v = h->v[m];
        LaunchCfg lc;
        if (int rc = choose_launch(h, m, ring_for_view(h, m), lc)) return rc;
        if (int rc = ensure_oc_scratch(h, (size_t)lc.grid * lc.W * (32 / h->G))) return rc;
        CK(h, cudaEventRecord(h->ev[2 + 2 * m], h->stream));
        if (v.n_items > 0) {
            CK(h, cudaMemcpyAsync(v.nk_snap, v.nk, Kp * 4, cudaMemcpyDeviceToDevice, h->stream));
            CK(h, cudaMemsetAsync(h->work_counter, 0, sizeof(int), h->stream));
            SweepParams P;
            fill_params(h, m, iteration, 1, P);
            P.R = lc.R; P.oc_smem = lc.oc_smem;
            h->stats.ring_depth[m] = lc.R;
            P.z_host = alias[m];
            CK(h, launch_sweep(h, P, lc));
            launches++;
        }
        CK(h, cudaEventRecord(h->ev[3 + 2 * m], h->stream));
        if (v.n_tok > 0 && !alias[m])
            CK(h, cudaMemcpyAsync(z_inout[m], v.z, (size_t)v.n_tok * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(h, cudaEventRecord(h->ev[1], h->stream));
    unsigned long long st[4];
    int bad = 0;
    CK(h, cudaMemcpyAsync(st, h->d_stats, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(&bad, h->d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (bad) {      // those tokens were treated as UNASSIGNED_TOPIC on the device; the caller's arrays may hold their new topics
        FAIL(h, MVTM_ERR_ARG, "mvtm_sweep_host: the assignments hold %d topic ids >= K", bad);
    }
    h->stats.tokens = (long long)st[0]; h->stats.changed = (long long)st[1]; h->stats.new_topic = (long long)st[2];
    float ms = 0.f;
    CK(h, cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    h->stats.ms_total = ms;
    for (int m = 0; m < h->M; m++) {
        CK(h, cudaEventElapsedTime(&ms, h->ev[2 + 2 * m], h->ev[3 + 2 * m]));
        h->stats.ms_view[m] = ms;
        if (h->v[m].n_items > 0) ring_record(h, m, ms);
    }
    h->stats.kernel_launches = launches;
    return activate_sampled_topics(h);
}

extern "C" int mvtm_stats(mvtm_handle *h, mvtm_sweep_stats *out)
{
    if (!h || !out) return MVTM_ERR_ARG;
    for (int m = 0; m < h->M; m++) {
        const char *e = getenv("MVTM_RING");
        h->stats.ring_locked[m] = h->direct ? 0 : (h->cfg_ring > 0 ? h->cfg_ring : (e ? atoi(e) : h->v[m].ring_locked));   // DIRECT kernel: no ring
    }
    *out = h->stats;
    return MVTM_OK;
}

static int cond_probs_impl(mvtm_handle *h, int32_t m, int64_t doc, int32_t pos, const double *p_row, const int32_t *not_in_S, int32_t n_not,
                           double *probs_out, int tree_mode = 0);

extern "C" int mvtm_cond_probs(mvtm_handle *h, int32_t m, int64_t doc, int32_t pos, const double *p_row, double *probs_out)
{
    if (!h) return MVTM_ERR_ARG;
    return cond_probs_impl(h, m, doc, pos, p_row, nullptr, -1, probs_out);
}

extern "C" int mvtm_cond_probs_ex(mvtm_handle *h, int32_t m, int64_t doc, int32_t pos, const double *p_row, int32_t tree_mode,
                                  const int32_t *not_in_S, int32_t n_not_in_S, double *probs_out)
{
    if (!h) return MVTM_ERR_ARG;
    if (tree_mode != 0 && tree_mode != 2) FAIL(h, MVTM_ERR_ARG, "mvtm_cond_probs_ex: tree_mode must be 0 (trainer) or 2 (inferencer)");
    if (n_not_in_S < -1 || (n_not_in_S > 0 && !not_in_S)) FAIL(h, MVTM_ERR_ARG, "mvtm_cond_probs_ex: bad not_in_S list");
    for (int i = 0; i < n_not_in_S; i++)
        if (not_in_S[i] < 0 || not_in_S[i] >= h->K) FAIL(h, MVTM_ERR_ARG, "mvtm_cond_probs_ex: topic %d out of range", not_in_S[i]);
    return cond_probs_impl(h, m, doc, pos, p_row, not_in_S, n_not_in_S, probs_out, tree_mode);
}

static int cond_probs_impl(mvtm_handle *h, int32_t m, int64_t doc, int32_t pos, const double *p_row, const int32_t *not_in_S, int32_t n_not,
                           double *probs_out, int tree_mode)
{
    if (int rc = require_views(h, "mvtm_cond_probs")) return rc;
    if (m < 0 || m >= h->M || doc < 0 || doc >= h->D || !probs_out) FAIL(h, MVTM_ERR_ARG, "mvtm_cond_probs: bad argument");
    ViewDev &v = h->v[m];
    long long len = v.h_doc_off[(size_t)doc + 1] - v.h_doc_off[(size_t)doc];
    if (pos < 0 || pos >= len) FAIL(h, MVTM_ERR_ARG, "mvtm_cond_probs: position %d outside document of %lld tokens", pos, len);
    CK(h, cudaSetDevice(h->device));
    if (int rc = wait_all_ready(h)) return rc;
    if (int rc = upload_hyper(h)) return rc;
    if (int rc = ensure_oc_scratch(h, (size_t)(32 / h->G))) return rc;
    int w = 0;
    CK(h, cudaMemcpyAsync(&w, v.word + v.h_doc_off[(size_t)doc] + pos, 4, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (w < 0 || w >= v.V) FAIL(h, MVTM_ERR_ARG, "mvtm_cond_probs: token has out-of-vocabulary word id %d (W:427-428 skips it)", w);
    double *d_out = nullptr, *d_p = nullptr;
    CK(h, cudaMalloc(&d_out, (size_t)(h->K + 1) * 8));
    std::vector<double> prow((size_t)h->M, 0.0);
    if (p_row) prow.assign(p_row, p_row + h->M); else prow[(size_t)m] = 1.0;
    CK(h, cudaMalloc(&d_p, (size_t)h->M * 8));
    CK(h, cudaMemcpyAsync(d_p, prow.data(), (size_t)h->M * 8, cudaMemcpyHostToDevice, h->stream));   // ordered before the probe
    CK(h, cudaStreamSynchronize(h->stream));                                                           // prow is a local
    SweepParams P;
    fill_params(h, m, 0, tree_mode, P);
    P.nk_frozen = v.nk;                                                  // frozen counts = the current ones
    P.R = 1;
    P.rbits = nullptr;
    unsigned *d_rb = nullptr;
    const bool q1 = n_not >= 0;                                          // Q1 probe: the caller names the held topics the index lacks
    if (q1) {
        std::vector<unsigned> words((size_t)h->KS / 32, 0u);
        for (int i = 0; i < n_not; i++) words[(size_t)not_in_S[i] >> 5] |= 1u << (not_in_S[i] & 31);
        CK(h, cudaMalloc(&d_rb, words.size() * 4));
        CK(h, cudaMemcpy(d_rb, words.data(), words.size() * 4, cudaMemcpyHostToDevice));
        P.rb_one = d_rb;
    }
    cudaError_t e = cudaErrorInvalidValue;
#define CALL_PROBE(KS_, G_, MU_) e = launch_probe_t<KS_, G_, MU_>(P, (int)doc, pos, d_p, d_out, h->stream, q1)
    if (h->M > 1) { DISPATCH_KG(h->KS, h->G, true, CALL_PROBE) } else { DISPATCH_KG(h->KS, h->G, false, CALL_PROBE) }
#undef CALL_PROBE
    if (e == cudaSuccess) e = cudaMemcpyAsync(probs_out, d_out, (size_t)(h->K + 1) * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_out); cudaFree(d_p); cudaFree(d_rb);
    CK(h, e);
    return MVTM_OK;
}

extern "C" int mvtm_doc_topic_hist(mvtm_handle *h, int32_t m, int32_t *hist_out, int32_t *max_len_out)
{
    if (!h) return MVTM_ERR_ARG;
    if (m < 0 || m >= h->M || !h->v[m].added) FAIL(h, MVTM_ERR_ARG, "mvtm_doc_topic_hist: bad view %d", m);
    ViewDev &v = h->v[m];
    if (max_len_out) *max_len_out = v.max_len;
    if (!hist_out) return MVTM_OK;
    CK(h, cudaSetDevice(h->device));
    const int stride = v.max_len + 1;
    const size_t n = (size_t)h->K * stride;
    int *d_hist = nullptr;
    CK(h, cudaMalloc(&d_hist, n * 4));
    cudaError_t e = cudaMemsetAsync(d_hist, 0, n * 4, h->stream);
    if (e == cudaSuccess && v.n_tok > 0) {
        const int warps = 4;
        k_doc_topic_hist<<<h->num_sms * 4, warps * 32, (size_t)warps * h->K * 4, h->stream>>>(h->D, v.doc_off, v.z, h->K, stride, d_hist);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(hist_out, d_hist, n * 4, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_hist);
    CK(h, e);
    // bin 0 (M:647-649): documents of the view in which the topic is absent
    for (int t = 0; t < h->K; t++) {
        long long s = 0;
        for (int cidx = 1; cidx < stride; cidx++) s += hist_out[(size_t)t * stride + cidx];
        hist_out[(size_t)t * stride] = (int)(v.docs_present - s);
    }
    return MVTM_OK;
}

static double host_log_gamma_stirling(double z)
{
    int shift = 0;
    while (z < 2.0) { z += 1.0; shift++; }
    double r = 0.91893853320467274178 + (z - 0.5) * std::log(z) - z + 1.0 / (12.0 * z) - 1.0 / (360.0 * z * z * z) + 1.0 / (1260.0 * z * z * z * z * z);
    while (shift-- > 0) { z -= 1.0; r -= std::log(z); }
    return r;
}

static int loglik_impl(mvtm_handle *h, double *ll_out, double *doc_out, double *word_out, int32_t quirk_len2);

extern "C" int mvtm_loglik(mvtm_handle *h, double *ll_out, int32_t quirk_len2)
{
    if (!h) return MVTM_ERR_ARG;
    if (!ll_out) FAIL(h, MVTM_ERR_ARG, "mvtm_loglik: NULL output");
    return loglik_impl(h, ll_out, nullptr, nullptr, quirk_len2);
}

extern "C" int mvtm_loglik_parts(mvtm_handle *h, double *doc_part_out, double *word_part_out, int32_t quirk_len2)
{
    if (!h) return MVTM_ERR_ARG;
    if (!doc_part_out || !word_part_out) FAIL(h, MVTM_ERR_ARG, "mvtm_loglik_parts: NULL output");
    return loglik_impl(h, nullptr, doc_part_out, word_part_out, quirk_len2);
}

static int loglik_impl(mvtm_handle *h, double *ll_out, double *doc_out, double *word_out, int32_t quirk_len2)
{
    if (int rc = require_views(h, "mvtm_loglik")) return rc;
    CK(h, cudaSetDevice(h->device));
    if (int rc = wait_all_ready(h)) return rc;
    const int K = h->K;
    const long long D = h->D;
    double *d_ga = nullptr, *d_tlg = nullptr, *d_doc = nullptr, *d_part = nullptr;
    long long *d_nnz = nullptr; unsigned char *d_cnt = nullptr;
    const int cell_blocks = h->num_sms * 8;
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    step(cudaMalloc(&d_ga, (size_t)K * 8)); step(cudaMalloc(&d_tlg, (size_t)K * 8));
    step(cudaMalloc(&d_doc, (size_t)std::max<long long>(D, 1) * 8)); step(cudaMalloc(&d_cnt, (size_t)std::max<long long>(D, 1)));
    step(cudaMalloc(&d_part, (size_t)cell_blocks * 8)); step(cudaMalloc(&d_nnz, (size_t)cell_blocks * 8));
    std::vector<double> ga((size_t)K), tlg((size_t)K), doc_ll((size_t)D), part((size_t)cell_blocks);
    std::vector<long long> nnzp((size_t)cell_blocks);
    std::vector<unsigned char> counted((size_t)D);
    std::vector<int> nk((size_t)K);
    for (int m = 0; m < h->M && e == cudaSuccess; m++) {
        ViewDev &v = h->v[m];
        for (int t = 0; t < K; t++) { ga[t] = h->gamma[m] * h->alpha[(size_t)m * (K + 1) + t]; tlg[t] = host_log_gamma_stirling(ga[t]); }   // M:3343
        const double gas = h->gamma[m] * h->alphaSum[m];
        step(cudaMemcpyAsync(d_ga, ga.data(), (size_t)K * 8, cudaMemcpyHostToDevice, h->stream));
        step(cudaMemcpyAsync(d_tlg, tlg.data(), (size_t)K * 8, cudaMemcpyHostToDevice, h->stream));
        if (e == cudaSuccess && D > 0) {
            const int warps = 4;
            k_loglik_docs<<<h->num_sms * 4, warps * 32, (size_t)warps * K * 4, h->stream>>>(D, v.doc_off, v.z, v.present, K, d_ga, d_tlg, gas, quirk_len2, d_doc, d_cnt);
            step(cudaGetLastError());
        }
        if (e == cudaSuccess) { k_loglik_cells<<<cell_blocks, 256, 0, h->stream>>>(v.V, K, h->Kp, v.nwk, h->beta[m], d_part, d_nnz); step(cudaGetLastError()); }
        if (D > 0) { step(cudaMemcpyAsync(doc_ll.data(), d_doc, (size_t)D * 8, cudaMemcpyDeviceToHost, h->stream));
                     step(cudaMemcpyAsync(counted.data(), d_cnt, (size_t)D, cudaMemcpyDeviceToHost, h->stream)); }
        step(cudaMemcpyAsync(part.data(), d_part, (size_t)cell_blocks * 8, cudaMemcpyDeviceToHost, h->stream));
        step(cudaMemcpyAsync(nnzp.data(), d_nnz, (size_t)cell_blocks * 8, cudaMemcpyDeviceToHost, h->stream));
        step(cudaMemcpyAsync(nk.data(), v.nk, (size_t)K * 4, cudaMemcpyDeviceToHost, h->stream));
        step(cudaStreamSynchronize(h->stream));
        if (e != cudaSuccess) break;
        double ll = 0.0; long long modalityCnt = 0;
        for (long long d = 0; d < D; d++) { ll += doc_ll[(size_t)d]; modalityCnt += counted[(size_t)d]; }
        ll += (double)modalityCnt * host_log_gamma_stirling(gas);                                 // M:3373
        const double doc_part = ll;                      // everything above sums over THIS handle's documents
        // the topic-word terms are accumulated on their own: functions of the (global) count tables only, so that every rank of a
        // multi-GPU run computes bit-identical values whatever its local document part is
        double wp = 0.0;
        long long nnz = 0;
        for (int bidx = 0; bidx < cell_blocks; bidx++) { wp += part[(size_t)bidx]; nnz += nnzp[(size_t)bidx]; }
        const double bV = h->beta[m] * v.V;
        for (int t = 0; t < K; t++) wp -= host_log_gamma_stirling(bV + nk[(size_t)t]);           // M:3417-3419
        wp += host_log_gamma_stirling(bV) * K;                                                    // M:3438
        wp -= host_log_gamma_stirling(h->beta[m]) * (double)nnz;                                  // M:3441
        if (ll_out) ll_out[m] = doc_part + wp;
        if (doc_out) doc_out[m] = doc_part;
        if (word_out) word_out[m] = wp;
    }
    cudaFree(d_ga); cudaFree(d_tlg); cudaFree(d_doc); cudaFree(d_cnt); cudaFree(d_part); cudaFree(d_nnz);
    CK(h, e);
    return MVTM_OK;
}

extern "C" int mvtm_check_invariants(mvtm_handle *h, int64_t *violations_out)
{
    if (!h) return MVTM_ERR_ARG;
    if (!violations_out) FAIL(h, MVTM_ERR_ARG, "mvtm_check_invariants: NULL output");
    if (int rc = require_views(h, "mvtm_check_invariants")) return rc;
    CK(h, cudaSetDevice(h->device));
    if (int rc = wait_all_ready(h)) return rc;
    unsigned long long *d_bad = nullptr;
    CK(h, cudaMalloc(&d_bad, 8));
    cudaError_t e = cudaMemsetAsync(d_bad, 0, 8, h->stream);
    auto step = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    step(cudaMemsetAsync(h->d_bad, 0, 4, h->stream));                 // once: out-of-range ids accumulate over the views
    for (int m = 0; m < h->M && e == cudaSuccess; m++) {
        ViewDev &v = h->v[m];
        const size_t n = (size_t)v.V * h->Kp;
        int *s_nwk = nullptr, *s_nk = nullptr;
        step(cudaMalloc(&s_nwk, n * 4)); step(cudaMalloc(&s_nk, (size_t)h->Kp * 4));
        step(cudaMemsetAsync(s_nwk, 0, n * 4, h->stream)); step(cudaMemsetAsync(s_nk, 0, (size_t)h->Kp * 4, h->stream));
        if (e == cudaSuccess && v.n_tok > 0) {
            int blocks = (int)std::min<long long>((v.n_tok + 255) / 256, (long long)h->num_sms * 8);
            k_build_counts<<<blocks, 256, (size_t)h->K * 4, h->stream>>>(v.n_tok, v.word, v.z, v.V, h->K, h->Kp, s_nwk, s_nk, h->d_bad, 0);
            step(cudaGetLastError());
        }
        if (e == cudaSuccess) {
            k_compare_counts<<<h->num_sms * 8, 256, 0, h->stream>>>((long long)n, v.nwk, s_nwk, d_bad);
            k_compare_counts<<<1, 256, 0, h->stream>>>((long long)h->Kp, v.nk, s_nk, d_bad);
            step(cudaGetLastError());
        }
        step(cudaStreamSynchronize(h->stream));
        cudaFree(s_nwk); cudaFree(s_nk);
    }
    unsigned long long bad = 0; int badz = 0;
    step(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, h->stream));
    step(cudaMemcpyAsync(&badz, h->d_bad, 4, cudaMemcpyDeviceToHost, h->stream));
    step(cudaStreamSynchronize(h->stream));
    cudaFree(d_bad);
    CK(h, e);
    *violations_out = (int64_t)bad + badz;
    return MVTM_OK;
}

// held-out evaluation by document completion (see include/mvtm.h)
extern "C" int mvtm_heldout_loglik(mvtm_handle *h, int32_t m, const int64_t *eval_off, const int32_t *eval_word, double *ll_out, int64_t *n_out)
{
    if (!h) return MVTM_ERR_ARG;
    if (m < 0 || m >= h->M || !h->v[m].added || !eval_off || !ll_out || !n_out) FAIL(h, MVTM_ERR_ARG, "mvtm_heldout_loglik: bad argument");
    const long long D = h->D;
    if (eval_off[0] != 0) FAIL(h, MVTM_ERR_ARG, "mvtm_heldout_loglik: eval_off[0] must be 0");
    for (long long d = 0; d < D; d++) if (eval_off[d + 1] < eval_off[d]) FAIL(h, MVTM_ERR_ARG, "mvtm_heldout_loglik: eval_off not monotone at doc %lld", d);
    const long long NE = eval_off[D];
    if (NE > 0 && !eval_word) FAIL(h, MVTM_ERR_ARG, "mvtm_heldout_loglik: NULL eval_word");
    *ll_out = 0.0; *n_out = 0;
    if (D == 0 || NE == 0) return MVTM_OK;
    CK(h, cudaSetDevice(h->device));
    if (int rc = wait_view_ready(h, m)) return rc;
    ViewDev &v = h->v[m];
    const int K = h->K;
    std::vector<double> ga((size_t)K);
    double ga_sum = 0.0;
    for (int t = 0; t < K; t++) ga[(size_t)t] = h->gamma[m] * h->alpha[(size_t)m * (K + 1) + t];
    for (int t : h->inactive) ga[(size_t)t] = 0.0;                      // as the sampler's prior mass (M:2670-2671)
    for (int t = 0; t < K; t++) ga_sum += ga[(size_t)t];
    long long *d_eoff = nullptr; int *d_eword = nullptr, *d_n = nullptr; double *d_ga = nullptr, *d_ll = nullptr;
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    step(cudaMalloc(&d_eoff, (size_t)(D + 1) * 8)); step(cudaMalloc(&d_eword, (size_t)NE * 4)); step(cudaMalloc(&d_ga, (size_t)K * 8));
    step(cudaMalloc(&d_ll, (size_t)D * 8)); step(cudaMalloc(&d_n, (size_t)D * 4));
    step(cudaMemcpyAsync(d_eoff, eval_off, (size_t)(D + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    step(cudaMemcpyAsync(d_eword, eval_word, (size_t)NE * 4, cudaMemcpyHostToDevice, h->stream));
    step(cudaMemcpyAsync(d_ga, ga.data(), (size_t)K * 8, cudaMemcpyHostToDevice, h->stream));
    std::vector<double> ll((size_t)D); std::vector<int> n((size_t)D);
    if (e == cudaSuccess) {
        const int warps = 4;
        k_heldout_docs<<<h->num_sms * 4, warps * 32, (size_t)warps * K * 4, h->stream>>>(D, v.doc_off, v.z, d_eoff, d_eword, v.V, K, h->Kp, v.nwk, v.nk,
                                                                                        d_ga, ga_sum, h->beta[m], h->betaSum[m], d_ll, d_n);
        step(cudaGetLastError());
    }
    step(cudaMemcpyAsync(ll.data(), d_ll, (size_t)D * 8, cudaMemcpyDeviceToHost, h->stream));
    step(cudaMemcpyAsync(n.data(), d_n, (size_t)D * 4, cudaMemcpyDeviceToHost, h->stream));
    step(cudaStreamSynchronize(h->stream));
    cudaFree(d_eoff); cudaFree(d_eword); cudaFree(d_ga); cudaFree(d_ll); cudaFree(d_n);
    CK(h, e);
    double tot = 0.0; long long cnt = 0;
    for (long long d = 0; d < D; d++) { tot += ll[(size_t)d]; cnt += n[(size_t)d]; }      // fixed order: deterministic
    *ll_out = tot; *n_out = cnt;
    return MVTM_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU delta plumbing: snapshot, export local delta in place, import reduced delta
// ------------------------------------------------------------------------------------------------
extern "C" int mvtm_delta_begin(mvtm_handle *h)
{
    if (!h) return MVTM_ERR_ARG;
    h->mut_epoch++;
    if (int rc = require_views(h, "mvtm_delta_begin")) return rc;
    CK(h, cudaSetDevice(h->device));
    if (int rc = wait_all_ready(h)) return rc;
    for (int m = 0; m < h->M; m++) {
        ViewDev &v = h->v[m];
        const size_t n = (size_t)v.V * h->Kp;
        if (!v.snap_nwk) { CK(h, cudaMalloc(&v.snap_nwk, (n + h->Kp) * 4)); v.snap_nk = v.snap_nwk + n; }
        CK(h, cudaMemcpyAsync(v.snap_nwk, v.nwk, n * 4, cudaMemcpyDeviceToDevice, h->stream));
        CK(h, cudaMemcpyAsync(v.snap_nk, v.nk, (size_t)h->Kp * 4, cudaMemcpyDeviceToDevice, h->stream));
    }
    CK(h, cudaStreamSynchronize(h->stream));
    return MVTM_OK;
}

extern "C" int mvtm_delta_reset(mvtm_handle *h)
{
    if (!h) return MVTM_ERR_ARG;
    h->mut_epoch++;
    if (int rc = require_views(h, "mvtm_delta_reset")) return rc;
    CK(h, cudaSetDevice(h->device));
    if (int rc = wait_all_ready(h)) return rc;
    for (int m = 0; m < h->M; m++) {
        ViewDev &v = h->v[m];
        const size_t n = (size_t)v.V * h->Kp;
        if (!v.snap_nwk) { CK(h, cudaMalloc(&v.snap_nwk, (n + h->Kp) * 4)); v.snap_nk = v.snap_nwk + n; }
        CK(h, cudaMemsetAsync(v.snap_nwk, 0, n * 4, h->stream));
        CK(h, cudaMemsetAsync(v.snap_nk, 0, (size_t)h->Kp * 4, h->stream));
    }
    CK(h, cudaStreamSynchronize(h->stream));
    return MVTM_OK;
}

extern "C" int mvtm_delta_export(mvtm_handle *h, int32_t m, void **n_wk_dev, int64_t *n_wk_elems, void **n_k_dev, int64_t *n_k_elems)
{
    if (!h) return MVTM_ERR_ARG;
    h->mut_epoch++;
    if (m < 0 || m >= h->M || !h->v[m].added) FAIL(h, MVTM_ERR_ARG, "mvtm_delta_export: bad view %d", m);
    ViewDev &v = h->v[m];
    if (!v.snap_nwk) FAIL(h, MVTM_ERR_STATE, "mvtm_delta_export: call mvtm_delta_begin first");
    CK(h, cudaSetDevice(h->device));
    if (int rc = wait_view_ready(h, m)) return rc;
    const long long n = (long long)v.V * h->Kp;
    k_sub_inplace<<<h->num_sms * 8, 256, 0, h->stream>>>(n, v.nwk, v.snap_nwk);
    k_sub_inplace<<<1, 256, 0, h->stream>>>((long long)h->Kp, v.nk, v.snap_nk);
    CK(h, cudaGetLastError());
    CK(h, cudaStreamSynchronize(h->stream));
    if (n_wk_dev) *n_wk_dev = v.nwk;
    if (n_wk_elems) *n_wk_elems = n;
    if (n_k_dev) *n_k_dev = v.nk;
    if (n_k_elems) *n_k_elems = h->Kp;
    return MVTM_OK;
}

extern "C" int mvtm_sum_exchange_buffers(mvtm_handle *h, int32_t m, void **n_wk_dev, int64_t *n_wk_elems, void **n_k_dev, int64_t *n_k_elems)
{
    if (!h) return MVTM_ERR_ARG;
    h->mut_epoch++;
    if (m < 0 || m >= h->M || !h->v[m].added) FAIL(h, MVTM_ERR_ARG, "mvtm_sum_exchange_buffers: bad view %d", m);
    ViewDev &v = h->v[m];
    if (!v.snap_nwk) FAIL(h, MVTM_ERR_STATE, "mvtm_sum_exchange_buffers: call mvtm_delta_begin first");
    CK(h, cudaSetDevice(h->device));
    if (int rc = wait_view_ready(h, m)) return rc;
    CK(h, cudaStreamSynchronize(h->stream));
    if (n_wk_dev) *n_wk_dev = v.nwk;
    if (n_wk_elems) *n_wk_elems = (long long)v.V * h->Kp;
    if (n_k_dev) *n_k_dev = v.nk;
    if (n_k_elems) *n_k_elems = h->Kp;
    return MVTM_OK;
}

extern "C" int mvtm_sum_exchange_finish(mvtm_handle *h, int32_t m, int32_t world_size)
{
    if (!h) return MVTM_ERR_ARG;
    h->mut_epoch++;
    if (m < 0 || m >= h->M || !h->v[m].added || world_size < 1) FAIL(h, MVTM_ERR_ARG, "mvtm_sum_exchange_finish: bad argument");
    ViewDev &v = h->v[m];
    if (!v.snap_nwk) FAIL(h, MVTM_ERR_STATE, "mvtm_sum_exchange_finish: call mvtm_delta_begin first");
    CK(h, cudaSetDevice(h->device));
    if (int rc = wait_view_ready(h, m)) return rc;
    const long long n = (long long)v.V * h->Kp;
    k_finish_sum_exchange<<<h->num_sms * 8, 256, 0, h->stream>>>(n, v.nwk, v.snap_nwk, world_size - 1);
    k_finish_sum_exchange<<<1, 256, 0, h->stream>>>((long long)h->Kp, v.nk, v.snap_nk, world_size - 1);
    CK(h, cudaGetLastError());
    CK(h, cudaStreamSynchronize(h->stream));
    return MVTM_OK;
}

// The same finishing pass queued on the CALLER's stream (behind its all-reduce), no host synchronisation; table and totals
// are one allocation, so one launch covers both.  Pair it with mvtm_view_wait_stream so the view's next pass waits for it.
extern "C" int mvtm_sum_exchange_finish_async(mvtm_handle *h, int32_t m, int32_t world_size, void *stream, int32_t max_ctas)
{
    if (!h) return MVTM_ERR_ARG;
    h->mut_epoch++;
    if (m < 0 || m >= h->M || !h->v[m].added || world_size < 1) FAIL(h, MVTM_ERR_ARG, "mvtm_sum_exchange_finish_async: bad argument");
    ViewDev &v = h->v[m];
    if (!v.snap_nwk) FAIL(h, MVTM_ERR_STATE, "mvtm_sum_exchange_finish_async: call mvtm_delta_begin first");
    CK(h, cudaSetDevice(h->device));
    const long long n = ((long long)v.V + 1) * h->Kp;
    const int grid = max_ctas > 0 ? max_ctas : h->num_sms * 8;
    k_finish_sum_exchange4<<<grid, 512, 0, (cudaStream_t)stream>>>(n / 4, (int4 *)v.nwk, (int4 *)v.snap_nwk, world_size - 1);
    CK(h, cudaGetLastError());
    return MVTM_OK;
}

extern "C" int mvtm_delta_import(mvtm_handle *h, int32_t m)
{
    if (!h) return MVTM_ERR_ARG;
    h->mut_epoch++;
    if (m < 0 || m >= h->M || !h->v[m].added) FAIL(h, MVTM_ERR_ARG, "mvtm_delta_import: bad view %d", m);
    ViewDev &v = h->v[m];
    if (!v.snap_nwk) FAIL(h, MVTM_ERR_STATE, "mvtm_delta_import: call mvtm_delta_begin first");
    CK(h, cudaSetDevice(h->device));
    if (int rc = wait_view_ready(h, m)) return rc;
    const long long n = (long long)v.V * h->Kp;
    k_add_snapshot<<<h->num_sms * 8, 256, 0, h->stream>>>(n, v.nwk, v.snap_nwk);
    k_add_snapshot<<<1, 256, 0, h->stream>>>((long long)h->Kp, v.nk, v.snap_nk);
    CK(h, cudaGetLastError());
    CK(h, cudaStreamSynchronize(h->stream));
    return MVTM_OK;
}

#include "mvtm_optim.inl"
// test hook: the launch shape of a view pass for (K, M, flags), no device needed
extern "C" int mvtm_test_launch_shape(int32_t num_topics, int32_t num_views, uint32_t flags, int32_t *lanes_per_doc, int32_t *ring_depth,
                                      int32_t *warps_per_cta, int64_t *smem_bytes, int32_t *maxnreg)
{
    if (num_topics < 1 || pick_J(num_topics) == 0 || num_views < 1 || num_views > MVTM_MAX_VIEWS) return MVTM_ERR_ARG;
    if (!lanes_per_doc || !ring_depth || !warps_per_cta || !smem_bytes || !maxnreg) return MVTM_ERR_ARG;
    mvtm_handle h;
    h.K = num_topics; h.M = num_views; h.flags = flags; h.num_sms = 148;
    h.Kp = (h.K + 31) / 32 * 32; h.J = pick_J(h.K); h.KS = h.J * 128;
    h.direct = pick_direct(h.KS, h.M > 1, h.flags);
    h.G = pick_G(h.KS, h.M > 1, h.direct);
    if (h.direct && !direct_compiled(h.KS, h.G)) { h.direct = false; h.G = pick_G(h.KS, h.M > 1, false); }
    LaunchCfg lc;
    if (int rc = choose_launch(&h, 0, 1, lc)) return rc;
    *lanes_per_doc = h.G; *ring_depth = lc.R; *warps_per_cta = lc.W; *smem_bytes = (int64_t)lc.smem;
    *maxnreg = h.direct ? direct_regs(h.KS, h.G) : 0;
    return MVTM_OK;
}

#include "mvtm_comm.inl"
