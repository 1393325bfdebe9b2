// Multi-GPU driver for the C ABI WITHOUT torch or Python: N host threads, one handle per GPU, NCCL behind mvtm_comm_init.
// What a JVM host would do with one Java thread per device.  usage: dist_driver [world=2] [sweeps=6] [views=2]
// (views = 1: a single-view corpus, whose exchange cannot be hidden and runs on the handle's own stream)
//
// Corpus: D two-view documents generated from a fixed LCG; rank r holds documents r, r+N, ... (doc_id_base / doc_id_stride).
// Checks: (1) after mvtm_sync_counts every rank holds the same n_k, totalling the corpus; (2) after S mvtm_sweep_dist sweeps
// (overlapped exchange) the same again, and each rank's replica equals the histogram of ALL ranks' assignments (cell by cell, on
// the host); (3) the global log-likelihood is the same number on every rank and improved; (4) mvtm_sweep_host_dist
// steps through host arrays -- recounting on the first call, keeping the resident counts when no rank's arrays changed, recounting
// on EVERY rank when one rank edits one token -- leave consistent global tables that mvtm_sweep_dist continues from; (5) mvtm_optimize_hyper installs identical
// hyper-parameters on every rank (statistics reduced inside the library).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "mvtm.h"

namespace {
constexpr int K = 64, MMAX = 2;
int M = 2;                                   // views in use (argv[3])
const int32_t V[MMAX] = { 700, 90 };
struct Shard { std::vector<int64_t> off[MMAX]; std::vector<int32_t> word[MMAX]; int64_t D = 0; };
struct Result { std::vector<int32_t> nk[MMAX], z[MMAX], nwk[MMAX]; double ll0[MMAX], ll1[MMAX]; std::vector<double> alpha; int rc = 0; std::string err; };

uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

Shard make_shard(int D_total, int rank, int world)
{
    Shard sh;
    for (int m = 0; m < MMAX; m++) sh.off[m].push_back(0);
    for (int d = 0; d < D_total; d++) {
        uint32_t s = 12345u + 7919u * (uint32_t)d;            // per-document stream: the shard of a document does not change its words
        const int topic = (int)(lcg(s) % 16);
        const int len0 = 8 + (int)(lcg(s) % 40), len1 = (d % 5 == 4) ? 0 : 2 + (int)(lcg(s) % 4);
        std::vector<int32_t> w0, w1;
        for (int i = 0; i < len0; i++) w0.push_back((int32_t)((topic * 40 + lcg(s) % 60 + (lcg(s) % 6 == 0 ? lcg(s) % 700 : 0)) % 700));
        for (int i = 0; i < len1; i++) w1.push_back((int32_t)((topic * 5 + lcg(s) % 8) % 90));
        if (d % world != rank) continue;
        sh.word[0].insert(sh.word[0].end(), w0.begin(), w0.end()); sh.off[0].push_back((int64_t)sh.word[0].size());
        sh.word[1].insert(sh.word[1].end(), w1.begin(), w1.end()); sh.off[1].push_back((int64_t)sh.word[1].size());
        sh.D++;
    }
    return sh;
}

#define CHECK(call) do { int _rc = (call); if (_rc) { res.rc = _rc; res.err = std::string(#call) + ": " + mvtm_last_error(h); if (h) mvtm_destroy(h); return; } } while (0)

void rank_main(int rank, int world, int D_total, int sweeps, const unsigned char *id, Result &res)
{
    mvtm_handle *h = nullptr;
    Shard sh = make_shard(D_total, rank, world);
    mvtm_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.num_topics = K; cfg.num_views = M; cfg.num_docs = sh.D; cfg.vocab_sizes = V; cfg.seed = 99; cfg.device = rank;
    cfg.doc_id_base = rank; cfg.doc_id_stride = world; cfg.max_ctas = 120; cfg.ring_depth = 1;
    if (int rc = mvtm_create(&cfg, &h)) { res.rc = rc; res.err = std::string("mvtm_create: ") + mvtm_last_error(nullptr); return; }
    for (int m = 0; m < M; m++) CHECK(mvtm_add_view(h, m, sh.off[m].data(), sh.word[m].data(), nullptr));
    CHECK(mvtm_comm_init(h, id, rank, world, 8));
    CHECK(mvtm_init_assignments(h));
    CHECK(mvtm_sync_counts(h, 0));
    CHECK(mvtm_loglik_dist(h, res.ll0, 0));
    for (int it = 1; it <= sweeps; it++) CHECK(mvtm_sweep_dist(h, it));
    CHECK(mvtm_comm_drain(h));
    // a stateless host step on every rank, then back to resident sweeps
    std::vector<int32_t> zh[MMAX]; int32_t *zp[MMAX];
    for (int m = 0; m < M; m++) { zh[m].resize(sh.word[m].size() + 1); CHECK(mvtm_get_assignments(h, m, zh[m].data())); zp[m] = zh[m].data(); }
    int32_t kept = -1;
    CHECK(mvtm_sweep_host_dist(h, sweeps + 1, zp));                      // first host step: recount + all-reduce
    CHECK(mvtm_comm_last_host_step(h, &kept));
    if (kept != 0) { res.rc = 100; res.err = "first mvtm_sweep_host_dist did not recount"; mvtm_destroy(h); return; }
    CHECK(mvtm_sweep_host_dist(h, sweeps + 1, zp));                      // arrays unchanged on every rank: resident counts kept
    CHECK(mvtm_comm_last_host_step(h, &kept));
    if (kept != 1) { res.rc = 101; res.err = "second mvtm_sweep_host_dist recounted although nothing changed"; mvtm_destroy(h); return; }
    if (rank == world - 1 && !zh[0].empty()) zh[0][0] = (zh[0][0] + 1) % K;   // ONE rank edits ONE token: every rank must recount
    CHECK(mvtm_sweep_host_dist(h, sweeps + 1, zp));
    CHECK(mvtm_comm_last_host_step(h, &kept));
    if (kept != 0) { res.rc = 102; res.err = "mvtm_sweep_host_dist missed an edit made on another rank"; mvtm_destroy(h); return; }
    CHECK(mvtm_sweep_dist(h, sweeps + 2));                               // global counts were left behind
    CHECK(mvtm_comm_drain(h));
    CHECK(mvtm_optimize_hyper(h, 50, MVTM_OPT_ALL));
    CHECK(mvtm_sweep_dist(h, sweeps + 3));
    CHECK(mvtm_comm_drain(h));
    CHECK(mvtm_loglik_dist(h, res.ll1, 0));
    for (int m = 0; m < M; m++) {
        res.nk[m].resize(K); res.nwk[m].resize((size_t)V[m] * K); res.z[m].resize(sh.word[m].size() + 1);
        CHECK(mvtm_get_counts(h, m, res.nwk[m].data(), res.nk[m].data()));
        CHECK(mvtm_get_assignments(h, m, res.z[m].data()));
        res.z[m].resize(sh.word[m].size());
    }
    res.alpha.resize((size_t)M * (K + 1));
    double asum[MMAX]; int32_t ina[K], nin = 0;
    CHECK(mvtm_get_hyper(h, res.alpha.data(), asum, ina, &nin));
    int32_t r = -1, w = -1, ver = 0; int64_t bytes = 0;
    CHECK(mvtm_comm_info(h, &r, &w, &ver, &bytes));
    if (rank == 0) std::printf("NCCL %d, world %d, %lld bytes all-reduced by the last sweep\n", ver, w, (long long)bytes);
    mvtm_destroy(h);
}
}  // namespace

int main(int argc, char **argv)
{
    const int world = argc > 1 ? atoi(argv[1]) : 2, sweeps = argc > 2 ? atoi(argv[2]) : 6, D_total = 6000;
    M = (argc > 3 && atoi(argv[3]) == 1) ? 1 : 2;
    unsigned char id[MVTM_COMM_ID_BYTES];
    if (mvtm_comm_unique_id(id)) { std::printf("mvtm_comm_unique_id: %s\nFAIL\n", mvtm_last_error(nullptr)); return 2; }
    std::vector<Result> res((size_t)world);
    std::vector<std::thread> th;
    for (int r = 0; r < world; r++) th.emplace_back(rank_main, r, world, D_total, sweeps, id, std::ref(res[(size_t)r]));
    for (auto &t : th) t.join();
    for (int r = 0; r < world; r++) if (res[(size_t)r].rc) { std::printf("rank %d: status %d: %s\nFAIL\n", r, res[(size_t)r].rc, res[(size_t)r].err.c_str()); return 1; }
    int bad = 0;
    for (int m = 0; m < M; m++) {
        // global histogram of all ranks' assignments, on the host
        std::vector<int32_t> nwk((size_t)V[m] * K, 0), nk((size_t)K, 0);
        long long ntok = 0;
        for (int r = 0; r < world; r++) {
            Shard sh = make_shard(D_total, r, world);
            for (size_t i = 0; i < sh.word[m].size(); i++) { int t = res[(size_t)r].z[m][i]; nwk[(size_t)sh.word[m][i] * K + t]++; nk[(size_t)t]++; ntok++; }
        }
        for (int r = 0; r < world; r++) {
            const Result &R = res[(size_t)r];
            auto report = [&](bool wrong, const char *what) { if (wrong) { bad++; std::printf("MISMATCH rank %d view %d: %s\n", r, m, what); } };
            report(R.nk[m] != nk, "n_k != histogram of all ranks' assignments");
            report(R.nwk[m] != nwk, "n_wk != histogram of all ranks' assignments");
            report(R.ll0[m] != res[0].ll0[m], "initial global LL differs between ranks");
            report(R.ll1[m] != res[0].ll1[m], "final global LL differs between ranks");
            report(!(R.ll1[m] > R.ll0[m]), "LL did not improve");
            report(R.alpha != res[0].alpha, "alpha differs between ranks after mvtm_optimize_hyper");
            if (R.ll1[m] != res[0].ll1[m]) std::printf("   %.17g vs %.17g\n", R.ll1[m], res[0].ll1[m]);
        }
        std::printf("view %d: %lld tokens, LL/token %.4f -> %.4f\n", m, ntok, res[0].ll0[m] / (double)ntok, res[0].ll1[m] / (double)ntok);
    }
    std::printf(bad ? "FAIL (%d mismatches)\n" : "OK\n", bad);
    return bad ? 1 : 0;
}
