// mvtm_comm.inl -- multi-GPU behind the C ABI (included by mvtm.cu): NCCL communicators owned by the handle, the per-sweep count
// exchange (sum form, overlapped with the other views' passes), the stateless multi-rank host sweep, the hyper-parameter
// statistics and the global log-likelihood reduced inside the library.  SURVEY 8(b)/(e): a JVM host drives N GPUs with N handles
// and nothing but these calls -- no torch, no Python.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already mapped into the process if there is one -- e.g. PyTorch's --
// else the system one), so libmvtm.so itself carries no NCCL dependency and single-GPU hosts never touch it.
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, ncclConfig_t *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    std::string err;
    bool load()
    {
        if (lib) return true;
        const char *names[] = { "libnccl.so.2", "libnccl.so" };
        for (const char *n : names) if ((lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_LOCAL))) break;   // already in the process?
        if (!lib) for (const char *n : names) if ((lib = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
        if (!lib) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
#define SYM(field, name) do { *(void **)(&field) = dlsym(lib, name); if (!field) { err = std::string("libnccl lacks ") + name; lib = nullptr; return false; } } while (0)
        SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommSplit, "ncclCommSplit");
        SYM(CommDestroy, "ncclCommDestroy"); SYM(AllReduce, "ncclAllReduce"); SYM(GetErrorString, "ncclGetErrorString");
        SYM(GetVersion, "ncclGetVersion");
#undef SYM
        return true;
    }
};
NcclApi g_nccl;

}  // namespace

struct CommState {
    ncclComm_t wide = nullptr, narrow = nullptr;    // narrow: CTA-limited, for exchanges hidden under another view's pass
    int rank = 0, world = 1, hidden_ctas = 0;
    cudaStream_t stream = nullptr;                  // the collectives and their finishing passes
    bool counts_global = false;                     // replicas hold global counts and snapshots of them (sum-form exchange valid)
    unsigned long long host_epoch = 0;              // mvtm_handle::mut_epoch at the end of the last mvtm_sweep_host_dist (0: none)
    int *d_flag = nullptr;                          // mvtm_sweep_host_dist: differences between the uploaded and the resident assignments
    cudaEvent_t ev_wide = nullptr;                  // end of the last exchange that used the WIDE communicator on the collective stream
    bool wide_pending = false;
    int host_fast_last = 0;                         // 1: the last mvtm_sweep_host_dist found its state intact and skipped the count rebuild
    long long bytes_last = 0;                       // bytes all-reduced by the last sweep
    int64_t *d_i = nullptr; double *d_r = nullptr; size_t cap_i = 0, cap_r = 0;   // staging of reduced statistics
};

#define NCK(h, call)                                                                                              \
    do {                                                                                                          \
        ncclResult_t _r = (call);                                                                                 \
        if (_r != ncclSuccess) FAIL(h, MVTM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(_r), __FILE__, __LINE__); \
    } while (0)

static void comm_teardown(mvtm_handle *h)
{
    CommState *c = h->comm;
    if (!c) return;
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->narrow) g_nccl.CommDestroy(c->narrow);
    if (c->wide) g_nccl.CommDestroy(c->wide);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->ev_wide) cudaEventDestroy(c->ev_wide);
    cudaFree(c->d_i); cudaFree(c->d_r); cudaFree(c->d_flag);
    delete c;
    h->comm = nullptr;
}

extern "C" int mvtm_comm_unique_id(void *id_out)
{
    if (!id_out) return MVTM_ERR_ARG;
    if (!g_nccl.load()) { g_create_err = g_nccl.err; return MVTM_ERR_CUDA; }
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) { g_create_err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return MVTM_ERR_CUDA; }
    static_assert(sizeof(ncclUniqueId) == MVTM_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    memcpy(id_out, &id, sizeof(id));
    return MVTM_OK;
}

extern "C" int mvtm_comm_init(mvtm_handle *h, const void *unique_id, int32_t rank, int32_t world, int32_t hidden_ctas)
{
    if (!h) return MVTM_ERR_ARG;
    if (!unique_id || world < 1 || rank < 0 || rank >= world) FAIL(h, MVTM_ERR_ARG, "mvtm_comm_init: bad rank / world / id");
    if (h->comm) FAIL(h, MVTM_ERR_STATE, "mvtm_comm_init: the handle already has a communicator");
    if (!g_nccl.load()) FAIL(h, MVTM_ERR_CUDA, "mvtm_comm_init: %s", g_nccl.err.c_str());
    CK(h, cudaSetDevice(h->device));
    CommState *c = new CommState();
    c->rank = rank; c->world = world; c->hidden_ctas = hidden_ctas > 0 ? hidden_ctas : 0;
    h->comm = c;
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    ncclResult_t r = g_nccl.CommInitRank(&c->wide, world, id, rank);
    if (r == ncclSuccess && c->hidden_ctas > 0 && h->M > 1 && world > 1) {
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.minCTAs = 1; cfg.maxCTAs = c->hidden_ctas;
        r = g_nccl.CommSplit(c->wide, 0, rank, &c->narrow, &cfg);
    }
    cudaError_t e = cudaSuccess;
    if (r == ncclSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (r == ncclSuccess && e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_wide, cudaEventDisableTiming);
    if (r != ncclSuccess || e != cudaSuccess) {
        std::string msg = r != ncclSuccess ? std::string("NCCL: ") + g_nccl.GetErrorString(r) : std::string("CUDA: ") + cudaGetErrorString(e);
        comm_teardown(h);
        FAIL(h, MVTM_ERR_CUDA, "mvtm_comm_init: %s", msg.c_str());
    }
    return MVTM_OK;
}

extern "C" int mvtm_comm_destroy(mvtm_handle *h)
{
    if (!h) return MVTM_ERR_ARG;
    cudaSetDevice(h->device);
    comm_teardown(h);
    return MVTM_OK;
}

extern "C" int mvtm_comm_info(mvtm_handle *h, int32_t *rank, int32_t *world, int32_t *nccl_version, int64_t *bytes_last_sweep)
{
    if (!h) return MVTM_ERR_ARG;
    if (!h->comm) FAIL(h, MVTM_ERR_STATE, "mvtm_comm_info: no communicator (mvtm_comm_init)");
    if (rank) *rank = h->comm->rank;
    if (world) *world = h->comm->world;
    if (nccl_version) { int v = 0; g_nccl.GetVersion(&v); *nccl_version = v; }
    if (bytes_last_sweep) *bytes_last_sweep = h->comm->bytes_last;
    return MVTM_OK;
}

extern "C" int mvtm_comm_last_host_step(mvtm_handle *h, int32_t *resident_counts_used)
{
    if (!h || !resident_counts_used) return MVTM_ERR_ARG;
    if (!h->comm) FAIL(h, MVTM_ERR_STATE, "mvtm_comm_last_host_step: no communicator (mvtm_comm_init)");
    *resident_counts_used = h->comm->host_fast_last;
    return MVTM_OK;
}

static int require_comm(mvtm_handle *h, const char *who)
{
    if (!h->comm) FAIL(h, MVTM_ERR_STATE, "%s: no communicator (call mvtm_comm_init first)", who);
    return MVTM_OK;
}

static int ensure_snapshots(mvtm_handle *h)
{
    for (int m = 0; m < h->M; m++) {
        ViewDev &v = h->v[m];
        const size_t n = (size_t)v.V * h->Kp;
        if (!v.snap_nwk) { CK(h, cudaMalloc(&v.snap_nwk, (n + h->Kp) * 4)); v.snap_nk = v.snap_nwk + n; }
    }
    return MVTM_OK;
}

// replicas hold this rank's LOCAL counts (right after mvtm_init_assignments / mvtm_set_assignments): one in-place all-reduce per
// view (table and totals are one allocation) makes them global, and the snapshot the sum-form exchange needs is taken.
extern "C" int mvtm_sync_counts(mvtm_handle *h, int32_t rebuild_from_assignments)
{
    if (!h) return MVTM_ERR_ARG;
    if (int rc = require_comm(h, "mvtm_sync_counts")) return rc;
    if (int rc = require_views(h, "mvtm_sync_counts")) return rc;
    if (h->sweep_open) FAIL(h, MVTM_ERR_STATE, "mvtm_sync_counts: view passes are still open (mvtm_sweep_finish)");
    CK(h, cudaSetDevice(h->device));
    CommState *c = h->comm;
    CK(h, cudaStreamSynchronize(c->stream));
    if (int rc = wait_all_ready(h)) return rc;
    if (int rc = ensure_snapshots(h)) return rc;
    if (rebuild_from_assignments)
        for (int m = 0; m < h->M; m++) if (int rc = rebuild_counts_view(h, m)) return rc;
    for (int m = 0; m < h->M; m++) {
        ViewDev &v = h->v[m];
        const size_t n = ((size_t)v.V + 1) * h->Kp;
        NCK(h, g_nccl.AllReduce(v.nwk, v.nwk, n, ncclInt32, ncclSum, c->wide, h->stream));
        CK(h, cudaMemcpyAsync(v.snap_nwk, v.nwk, n * 4, cudaMemcpyDeviceToDevice, h->stream));
    }
    CK(h, cudaStreamSynchronize(h->stream));
    c->counts_global = true;
    h->mut_epoch++;
    return MVTM_OK;
}

// queue the sum-form exchange of view m behind its pass: all-reduce of the replica (in place, table + totals) and the finishing
// pass  replica -= (N-1)*snapshot; snapshot = replica  on `s`; the view's next pass and every reader wait for it on the device
static int enqueue_exchange(mvtm_handle *h, int m, cudaStream_t s, ncclComm_t comm, int finish_ctas)
{
    CommState *c = h->comm;
    ViewDev &v = h->v[m];
    const size_t n = ((size_t)v.V + 1) * h->Kp;
    if (s != h->stream) {
        CK(h, cudaEventRecord(h->ev_done[m], h->stream));
        CK(h, cudaStreamWaitEvent(s, h->ev_done[m], 0));
    }
    NCK(h, g_nccl.AllReduce(v.nwk, v.nwk, n, ncclInt32, ncclSum, comm, s));
    const int grid = finish_ctas > 0 ? finish_ctas : h->num_sms * 8;
    k_finish_sum_exchange4<<<grid, 512, 0, s>>>((long long)(n / 4), (int4 *)v.nwk, (int4 *)v.snap_nwk, c->world - 1);
    CK(h, cudaGetLastError());
    if (s != h->stream) {
        CK(h, cudaEventRecord(h->ev_ready[m], s));
        h->ready_pending[m] = true;
        if (comm == c->wide) { CK(h, cudaEventRecord(c->ev_wide, s)); c->wide_pending = true; }
    }
    c->bytes_last += (long long)n * 4;
    return MVTM_OK;
}

static int critical_view(mvtm_handle *h)
{   // the view with the longest pass: only the other, shorter passes lie between two of its own, so its exchange cannot be hidden
    int best = 0;
    for (int m = 1; m < h->M; m++) if (h->v[m].n_tok > h->v[best].n_tok) best = m;
    return best;
}

// the passes of one sweep with their exchanges (replicas hold global counts + snapshots on entry and again when the exchanges end).
// z_pageable != NULL: host arrays that are not pinned + mapped get a copy of the view's new assignments behind its pass.
static int dist_passes(mvtm_handle *h, int iteration, int32_t *const *z_pageable)
{
    CommState *c = h->comm;
    if (int rc = open_sweep(h)) return rc;
    h->open_mode = 1;
    const bool overlap = h->M > 1 && c->world > 1;
    const int crit = critical_view(h);
    for (int m = 0; m < h->M; m++) {
        if (int rc = enqueue_view_pass(h, iteration, 1, m, &h->open_launches)) { h->sweep_open = false; return rc; }
        if (z_pageable && z_pageable[m] && h->v[m].n_tok > 0)
            CK(h, cudaMemcpyAsync(z_pageable[m], h->v[m].z, (size_t)h->v[m].n_tok * 4, cudaMemcpyDeviceToHost, h->stream));
        if (c->world == 1) continue;
        int rc;
        if (overlap) {
            const bool hidden = (m != crit) && c->narrow;
            rc = enqueue_exchange(h, m, c->stream, hidden ? c->narrow : c->wide, 0);
        } else {
            rc = enqueue_exchange(h, m, h->stream, c->wide, 0);       // a single view has nothing to hide its exchange under
        }
        if (rc) { h->sweep_open = false; return rc; }
        h->open_launches++;
    }
    return close_sweep(h, 1);                                           // host side of the barrier M:1231: the PASSES only
}

extern "C" int mvtm_sweep_dist(mvtm_handle *h, int32_t iteration)
{
    if (!h) return MVTM_ERR_ARG;
    if (int rc = require_comm(h, "mvtm_sweep_dist")) return rc;
    CommState *c = h->comm;
    if (!c->counts_global) FAIL(h, MVTM_ERR_STATE, "mvtm_sweep_dist: the replicas do not hold global counts (call mvtm_sync_counts)");
    if (h->sweep_open) FAIL(h, MVTM_ERR_STATE, "mvtm_sweep_dist: passes queued by mvtm_sweep_view_async are still open");
    CK(h, cudaSetDevice(h->device));
    // activation of inactive topics that the PREVIOUS sweep sampled, on the global counts its exchanges produced (U:263-270):
    // every rank takes the same decision.  Waits for those exchanges; skipped (no wait) while no topic is inactive.
    if (!h->inactive.empty()) if (int rc = mvtm_activate_topics(h)) return rc;
    c->bytes_last = 0;
    return dist_passes(h, iteration, nullptr);
}

extern "C" int mvtm_comm_drain(mvtm_handle *h)
{
    if (!h) return MVTM_ERR_ARG;
    if (int rc = require_comm(h, "mvtm_comm_drain")) return rc;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->comm->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return MVTM_OK;
}

// The multi-rank step of a host that keeps the assignments (`topicSequence` arrays) as its only state: assignments in, assignments
// out, counts consistent over all ranks in between.
//  * First call, or anything else touched the handle's assignments / tables since the last one, or the caller changed its arrays:
//    every rank uploads its shard, rebuilds its LOCAL counts from it, ONE all-reduce per view makes them global (view m's runs
//    while the next view's upload travels) and the snapshot of the sum-form exchange is taken.
//  * Otherwise (the usual case: the arrays come back exactly as the previous call returned them) the upload is COMPARED with the
//    resident assignments while the previous sweep's last exchange is still finishing; one 4-byte all-reduce tells every rank
//    whether all shards are intact, and if so the resident global counts are used as they are.  Any difference anywhere sends all
//    ranks down the rebuild path with the uploaded data, so the result never depends on which path ran.
// Then the passes run with their overlapped exchanges exactly as in mvtm_sweep_dist, and the new assignments land in the caller's
// arrays (written by the sweep kernel itself when they are pinned + mapped).  The replicas hold global counts afterwards.
extern "C" int mvtm_sweep_host_dist(mvtm_handle *h, int32_t iteration, int32_t *const *z_inout)
{
    if (!h) return MVTM_ERR_ARG;
    if (!z_inout) FAIL(h, MVTM_ERR_ARG, "mvtm_sweep_host_dist: NULL z_inout");
    if (int rc = require_comm(h, "mvtm_sweep_host_dist")) return rc;
    if (h->sweep_open) FAIL(h, MVTM_ERR_STATE, "mvtm_sweep_host_dist: passes queued by mvtm_sweep_view_async are still open");
    if (int rc = require_views(h, "mvtm_sweep_host_dist")) return rc;
    CK(h, cudaSetDevice(h->device));
    CommState *c = h->comm;
    if (!h->copy_stream) CK(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    const size_t n_ev = (size_t)MVTM_MAX_VIEWS * HOST_CHUNKS;
    while (h->host_ev.size() < n_ev) { cudaEvent_t e; CK(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->host_ev.push_back(e); }
    auto ev_up = [&](int m, int cidx) { return h->host_ev[(size_t)(m * HOST_CHUNKS + cidx)]; };
    const size_t Kp = (size_t)h->Kp;
    int *alias[MVTM_MAX_VIEWS];
    int32_t *pageable[MVTM_MAX_VIEWS];
    for (int m = 0; m < h->M; m++) {
        if (h->v[m].n_tok > 0 && !z_inout[m]) FAIL(h, MVTM_ERR_ARG, "mvtm_sweep_host_dist: NULL z for view %d", m);
        alias[m] = h->v[m].n_tok > 0 ? mapped_alias(z_inout[m]) : nullptr;
        pageable[m] = alias[m] ? nullptr : z_inout[m];
    }
    if (int rc = ensure_snapshots(h)) return rc;
    if (!c->d_flag) CK(h, cudaMalloc(&c->d_flag, sizeof(int)));
    c->bytes_last = 0;
    c->host_fast_last = 0;
    const bool compare = c->counts_global && c->host_epoch != 0 && c->host_epoch == h->mut_epoch &&
                         !(getenv("MVTM_HOST_COMPARE") && atoi(getenv("MVTM_HOST_COMPARE")) == 0);
    bool intact = false;
    // activation of inactive topics the previous sweep sampled, on the global counts its exchanges produced (as mvtm_sweep_dist does)
    if (compare && !h->inactive.empty()) if (int rc = mvtm_activate_topics(h)) return rc;
    // EVERY rank takes part in the verdict, whether or not it can compare (a rank whose state was touched reports "changed"): the
    // ranks must agree on the sequence of collectives that follows, and only the reduced verdict is the same everywhere
    static const int k_one = 1;
    if (compare) CK(h, cudaMemsetAsync(c->d_flag, 0, sizeof(int), h->stream));
    else CK(h, cudaMemcpyAsync(c->d_flag, &k_one, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (compare) {
        // the uploads go to a staging buffer and the comparisons only read z: neither waits for the exchanges the previous call
        // left running on the collective stream
        for (int m = 0; m < h->M; m++)
            if (h->v[m].n_tok > 0 && !h->v[m].z_stage) CK(h, cudaMalloc(&h->v[m].z_stage, (size_t)h->v[m].n_tok * 4));
        CK(h, cudaEventRecord(h->ev_done[0], h->stream));
        CK(h, cudaStreamWaitEvent(h->copy_stream, h->ev_done[0], 0));
        for (int m = 0; m < h->M; m++) {
            ViewDev &v = h->v[m];
            for (int cidx = 0; cidx < HOST_CHUNKS; cidx++) {
                const long long t0 = v.chunk_tok_off[cidx], n = v.chunk_tok_off[cidx + 1] - t0;
                if (n <= 0) continue;
                CK(h, cudaMemcpyAsync(v.z_stage + t0, z_inout[m] + t0, (size_t)n * 4, cudaMemcpyHostToDevice, h->copy_stream));
                CK(h, cudaEventRecord(ev_up(m, cidx), h->copy_stream));
                CK(h, cudaStreamWaitEvent(h->stream, ev_up(m, cidx), 0));
                int blocks = (int)std::min<long long>((n + 1023) / 1024, (long long)h->num_sms * 4);
                k_diff_assign<<<blocks, 256, 0, h->stream>>>(n, v.z_stage + t0, v.z + t0, c->d_flag);
                CK(h, cudaGetLastError());
            }
        }
    }
    {
        // the verdict travels on the WIDE communicator from the handle's stream, ordered (by an event) behind the last exchange that
        // used that communicator on the collective stream -- the longest view's, which has to end before the first pass anyway --
        // but NOT behind the CTA-limited exchanges of the other views still queued there: those keep running under the passes, and
        // one communicator is never used from two streams at once
        if (c->wide_pending) CK(h, cudaStreamWaitEvent(h->stream, c->ev_wide, 0));
        if (c->world > 1) NCK(h, g_nccl.AllReduce(c->d_flag, c->d_flag, 1, ncclInt32, ncclSum, c->wide, h->stream));
        int ndiff = 0;
        CK(h, cudaMemcpyAsync(&ndiff, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        intact = (ndiff == 0);                       // every rank compared, no rank found a difference
    }
    int bad = 0;
    if (!intact) {
        if (int rc = wait_all_ready(h)) return rc;
        c->counts_global = false; c->host_epoch = 0;
        h->mut_epoch++;
        // neither the copy stream nor the collective stream may overtake earlier work on the handle's stream
        CK(h, cudaEventRecord(h->ev_done[0], h->stream));
        CK(h, cudaStreamWaitEvent(h->copy_stream, h->ev_done[0], 0));
        CK(h, cudaMemsetAsync(h->d_bad, 0, sizeof(int), h->stream));
        for (int m = 0; m < h->M; m++) {
            ViewDev &v = h->v[m];
            const size_t n_tab = ((size_t)v.V + 1) * Kp;
            CK(h, cudaMemsetAsync(v.nwk, 0, n_tab * 4, h->stream));
            for (int cidx = 0; cidx < HOST_CHUNKS; cidx++) {
                const long long t0 = v.chunk_tok_off[cidx], n = v.chunk_tok_off[cidx + 1] - t0;
                if (n <= 0) continue;
                if (compare) {      // already on the device
                    CK(h, cudaMemcpyAsync(v.z + t0, v.z_stage + t0, (size_t)n * 4, cudaMemcpyDeviceToDevice, h->stream));
                } else {
                    CK(h, cudaMemcpyAsync(v.z + t0, z_inout[m] + t0, (size_t)n * 4, cudaMemcpyHostToDevice, h->copy_stream));
                    CK(h, cudaEventRecord(ev_up(m, cidx), h->copy_stream));
                    CK(h, cudaStreamWaitEvent(h->stream, ev_up(m, cidx), 0));
                }
                int blocks = (int)std::min<long long>((n + 255) / 256, (long long)h->num_sms * 8);
                k_build_counts<<<blocks, 256, (size_t)h->K * 4, h->stream>>>(n, v.word + t0, v.z + t0, v.V, h->K, h->Kp, v.nwk, v.nk, h->d_bad, 1);
                CK(h, cudaGetLastError());
            }
            if (c->world > 1) {      // local -> global + snapshot, on the collective stream while the next view's chunks travel and are counted
                CK(h, cudaEventRecord(h->ev_done[m], h->stream));
                CK(h, cudaStreamWaitEvent(c->stream, h->ev_done[m], 0));
                NCK(h, g_nccl.AllReduce(v.nwk, v.nwk, n_tab, ncclInt32, ncclSum, c->wide, c->stream));
                CK(h, cudaMemcpyAsync(v.snap_nwk, v.nwk, n_tab * 4, cudaMemcpyDeviceToDevice, c->stream));
                CK(h, cudaEventRecord(h->ev_ready[m], c->stream));
                h->ready_pending[m] = true;
                c->bytes_last += (long long)(n_tab * 4);
            }
        }
        CK(h, cudaMemcpyAsync(&bad, h->d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        // topic ids >= K were rewritten to UNASSIGNED_TOPIC on the device: the step still runs (a rank that left here would leave
        // its peers waiting in the exchanges) and reports the error at the end
        c->counts_global = true;
    }
    c->host_fast_last = intact ? 1 : 0;
    for (int m = 0; m < h->M; m++) h->v[m].z_host_once = alias[m];
    const int rc = dist_passes(h, iteration, pageable);
    for (int m = 0; m < h->M; m++) h->v[m].z_host_once = nullptr;
    if (rc) { c->counts_global = false; c->host_epoch = 0; return rc; }
    if (bad) {
        c->host_epoch = 0;
        FAIL(h, MVTM_ERR_ARG, "mvtm_sweep_host_dist: the assignments hold %d topic ids >= K (treated as unassigned)", bad);
    }
    c->host_epoch = h->mut_epoch;
    return MVTM_OK;
}

// all-reduce of host-side statistics over the handle's communicator (sum or max), staged through device memory
static int comm_reduce_stats(mvtm_handle *h, int op, long long *ints, long long n_ints, double *reals, long long n_reals)
{
    CommState *c = h->comm;
    if (c->world == 1) return MVTM_OK;
    const ncclRedOp_t rop = op == 0 ? ncclSum : ncclMax;
    if (c->wide_pending) CK(h, cudaStreamWaitEvent(h->stream, c->ev_wide, 0));   // the wide communicator is never used from two streams at once
    if ((size_t)n_ints > c->cap_i) { cudaFree(c->d_i); c->d_i = nullptr; c->cap_i = 0; CK(h, cudaMalloc(&c->d_i, (size_t)n_ints * 8)); c->cap_i = (size_t)n_ints; }
    if ((size_t)n_reals > c->cap_r) { cudaFree(c->d_r); c->d_r = nullptr; c->cap_r = 0; CK(h, cudaMalloc(&c->d_r, (size_t)n_reals * 8)); c->cap_r = (size_t)n_reals; }
    if (n_ints > 0 && ints) {
        CK(h, cudaMemcpyAsync(c->d_i, ints, (size_t)n_ints * 8, cudaMemcpyHostToDevice, h->stream));
        NCK(h, g_nccl.AllReduce(c->d_i, c->d_i, (size_t)n_ints, ncclInt64, rop, c->wide, h->stream));
        CK(h, cudaMemcpyAsync(ints, c->d_i, (size_t)n_ints * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    if (n_reals > 0 && reals) {
        CK(h, cudaMemcpyAsync(c->d_r, reals, (size_t)n_reals * 8, cudaMemcpyHostToDevice, h->stream));
        NCK(h, g_nccl.AllReduce(c->d_r, c->d_r, (size_t)n_reals, ncclFloat64, rop, c->wide, h->stream));
        CK(h, cudaMemcpyAsync(reals, c->d_r, (size_t)n_reals * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(h, cudaStreamSynchronize(h->stream));
    return MVTM_OK;
}

// modelLogLikelihood (M:3322-3452) of the WHOLE corpus: this rank's document parts summed over the ranks + the topic-word part
// (a function of the global tables, identical on every rank).  Every rank must call it.
extern "C" int mvtm_loglik_dist(mvtm_handle *h, double *ll_out, int32_t quirk_len2)
{
    if (!h) return MVTM_ERR_ARG;
    if (!ll_out) FAIL(h, MVTM_ERR_ARG, "mvtm_loglik_dist: NULL output");
    if (int rc = require_comm(h, "mvtm_loglik_dist")) return rc;
    if (!h->comm->counts_global) FAIL(h, MVTM_ERR_STATE, "mvtm_loglik_dist: the replicas do not hold global counts (call mvtm_sync_counts)");
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->comm->stream));
    double doc[MVTM_MAX_VIEWS], word[MVTM_MAX_VIEWS];
    if (int rc = mvtm_loglik_parts(h, doc, word, quirk_len2)) return rc;
    if (int rc = comm_reduce_stats(h, 0, nullptr, 0, doc, h->M)) return rc;
    for (int m = 0; m < h->M; m++) ll_out[m] = doc[m] + word[m];
    return MVTM_OK;
}
