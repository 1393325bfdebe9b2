// mvtm_optim.inl -- hyper-parameter step of estimate() (M:1173-1210): optimizeP, optimizeDP, optimizeGamma,
// optimizeBeta.  Included by mvtm.cu (needs mvtm_handle).  SURVEY.md section 8(f) rank 1.
//
// The reference runs these on the host from statistics the sampling path maintains; so does this file: the device
// only produces the statistics (k_p_stats, k_doc_topic_hist, k_value_hist), the samplers run on the host in fp64.
// Random numbers: the reference draws from MALLET Randoms / knowceans samplers, none of them reproducible across
// runs (Q9); here every draw comes from one counter-based Philox stream keyed (seed, iteration), and the samplers are
// textbook algorithms with the same laws (Gamma: Marsaglia-Tsang; Beta: ratio of Gammas, KR:267-271; Antoniak: sum of
// Bernoulli(alpha/(alpha+i)), the law KS:1089-1110 samples by Stirling-number inversion; Dirichlet: normalised Gammas
// with the reference's 1e-4 floor, M:2593-2632).

namespace {

struct OptRng {
    uint64_t seed; uint32_t iteration; uint64_t n;
    // test hook (mvtm_test_hyper_core): when `script` is set, every sampler call returns the next scripted value instead of drawing,
    // and records (kind, a, b) -- kind 1 Gamma(a,1), 2 Beta(a,b), 3 Bernoulli(a), 4 Antoniak(a, n=b) -- so that the argument of
    // every draw of optimizeDP / optimizeGamma can be compared with what the reference's bytecode passes to its samplers
    const double *script = nullptr; long long script_len = 0, script_pos = 0; double *arg_log = nullptr; bool overrun = false;
    double take(double kind, double a, double b)
    {
        if (script_pos >= script_len) { overrun = true; script_pos++; return 1.0; }
        if (arg_log) { arg_log[3 * script_pos] = kind; arg_log[3 * script_pos + 1] = a; arg_log[3 * script_pos + 2] = b; }
        return script[script_pos++];
    }
    static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
    {
        for (int r = 0; r < 10; r++) {
            uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
            uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
    double next()
    {   // 53-bit uniform in [0,1): purpose 3 of the engine's counter layout
        uint32_t x[4];
        philox((uint32_t)n, (uint32_t)(n >> 32), iteration, 3u, (uint32_t)seed, (uint32_t)(seed >> 32), x);
        n++;
        return (double)(((uint64_t)(x[0] >> 5) << 26) | (uint64_t)(x[1] >> 6)) * (1.0 / 9007199254740992.0);
    }
    double next_open() { double u; do { u = next(); } while (u <= 0.0); return u; }
    double normal()
    {   // Box-Muller, one value per call
        const double u1 = next_open(), u2 = next();
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925 * u2);
    }
};

static double rand_gamma(OptRng &g, double shape)
{   // Gamma(shape, 1).  shape <= 0 -> 0 like KR:298-300
    if (!(shape > 0.0)) return 0.0;
    if (g.script) return g.take(1, shape, 0);
    if (shape < 1.0) return rand_gamma(g, shape + 1.0) * std::pow(g.next_open(), 1.0 / shape);
    const double d = shape - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
    for (;;) {
        double x, v;
        do { x = g.normal(); v = 1.0 + c * x; } while (v <= 0.0);
        v = v * v * v;
        const double u = g.next_open();
        if (u < 1.0 - 0.0331 * x * x * x * x) return d * v;
        if (std::log(u) < 0.5 * x * x + d * (1.0 - v + std::log(v))) return d * v;
    }
}
static double rand_beta(OptRng &g, double a, double b)
{   // KR:267-271: first component of a 2-dimensional Dirichlet drawn as normalised Gammas
    if (g.script) return g.take(2, a, b);
    const double x = rand_gamma(g, a), y = rand_gamma(g, b);
    return x / (x + y);
}
static int rand_bernoulli(OptRng &g, double p) { return g.script ? (int)g.take(3, p, 0) : (g.next() < p ? 1 : 0); }   // KR:789-795
static int rand_antoniak(OptRng &g, double alpha, int n)
{   // number of tables of a CRP(alpha) after n customers (Antoniak 1974), KS:1089-1110
    if (n > 20000) throw std::range_error("MAXSTIRLING");           // KS:1023: the reference's table ends here
    if (g.script) return (int)g.take(4, alpha, n);
    int m = 0;
    for (int i = 0; i < n; i++) m += rand_bernoulli(g, alpha / (alpha + i));
    return m < 1 ? 1 : m;
}
static void sample_dirichlet(OptRng &g, const std::vector<double> &p, std::vector<double> &out)
{   // M:2593-2632
    out.resize(p.size());
    double sum = 0.0;
    for (size_t i = 0; i < p.size(); i++) {
        double v = 1e-4;
        if (p[i] > 0.0) { v = rand_gamma(g, p[i]); if (v <= 0.0) v = 1e-4; }
        out[i] = v; sum += v;
    }
    for (double &v : out) v /= sum;
}

static double mallet_digamma(double z)
{   // cc.mallet.types.Dirichlet.digamma as compiled in MALLET 2.0.8: every Bernoulli term is constant-folded to 0 (Q19)
    if (z < 1e-6) return -0.5772156649015329 - 1.0 / z;
    double acc = 0.0;
    while (z < 9.5) { acc -= 1.0 / z; z += 1.0; }
    return acc + std::log(z) - 1.0 / (2.0 * z);
}

// cc.mallet.types.Dirichlet.learnSymmetricConcentration (call site M:2327), SURVEY 8(c): 200 fixed-point iterations; the
// denominator's "iterate up" branch never advances its lower index (previousLength stays 0), reproduced literally.
static double learn_symmetric_concentration(const std::vector<long long> &countHist, const std::vector<long long> &lengthHist,
                                            int numDimensions, double currentValue)
{
    int largestNonZeroCount = 0;
    for (size_t i = 0; i < countHist.size(); i++) if (countHist[i] > 0) largestNonZeroCount = (int)i;
    std::vector<int> nonZeroLength;
    for (size_t i = 0; i < lengthHist.size(); i++) if (lengthHist[i] > 0) nonZeroLength.push_back((int)i);
    for (int iteration = 1; iteration <= 200; iteration++) {
        const double currentParameter = currentValue / numDimensions;
        double currentDigamma = 0.0, numerator = 0.0;
        for (int index = 1; index <= largestNonZeroCount; index++) {
            currentDigamma += 1.0 / (currentParameter + index - 1);
            numerator += (double)countHist[(size_t)index] * currentDigamma;
        }
        currentDigamma = 0.0;
        double denominator = 0.0;
        const int previousLength = 0;
        const double cachedDigamma = mallet_digamma(currentValue);
        for (int length : nonZeroLength) {
            if (length - previousLength > 20) currentDigamma = mallet_digamma(currentValue + length) - cachedDigamma;
            else for (int index = previousLength; index < length; index++) currentDigamma += 1.0 / (currentValue + index);
            denominator += currentDigamma * (double)lengthHist[(size_t)length];
        }
        currentValue = currentParameter * numerator / denominator;
    }
    return currentValue;
}

}  // namespace

// ---- device statistics -----------------------------------------------------------------------------------------
// optimizeP's per-document overlap statistic (M:2706-2782), one warp per document.  For every document the views are
// ordered by length, descending, through the reference's TreeMap<Integer,Byte>: views of equal length collide and only
// the highest view index survives (Q11).  For each later view m and each earlier view i:
//     pDistr_Mean[m][i][doc] = #{tokens of view m whose topic also occurs in view i} / len_m.
// Output: per-block partial sums over documents, part[block][M*M] (symmetric entries both filled).
__global__ void k_p_stats(int M, int K, long long n_docs, const SweepParams P, double *part)
{
    extern __shared__ unsigned sm_bits[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int words = (K + 31) / 32;
    unsigned *bits = sm_bits + (size_t)warp * M * words;
    double acc[MVTM_MAXM * MVTM_MAXM];
    for (int k = 0; k < M * M; k++) acc[k] = 0.0;
    for (long long d = (long long)blockIdx.x * nwarp + warp; d < n_docs; d += (long long)gridDim.x * nwarp) {
        int len[MVTM_MAXM], kept[MVTM_MAXM], nkept = 0;
        for (int i = 0; i < M; i++) len[i] = (int)(P.doc_off[i][d + 1] - P.doc_off[i][d]);
        // descending distinct lengths; for a given length the last view put() wins
        int prev = 0x7fffffff;
        for (;;) {
            int best = -1;
            for (int i = 0; i < M; i++) if (len[i] < prev && len[i] > best) best = len[i];
            if (best < 0) break;
            int who = 0;
            for (int i = 0; i < M; i++) if (len[i] == best) who = i;
            kept[nkept++] = who; prev = best;
        }
        for (int k = lane; k < M * words; k += 32) bits[k] = 0u;
        __syncwarp();
        for (int i = 0; i < M; i++) {
            const int *zi = P.zv[i] + P.doc_off[i][d];
            for (int k = lane; k < len[i]; k += 32) { int t = zi[k]; if (t >= 0) atomicOr(&bits[(size_t)i * words + (t >> 5)], 1u << (t & 31)); }
        }
        __syncwarp();
        for (int r = 1; r < nkept; r++) {
            const int m = kept[r];
            if (len[m] == 0) continue;
            const int *zm = P.zv[m] + P.doc_off[m][d];
            for (int q = 0; q < r; q++) {
                const int i = kept[q];
                int cnt = 0;
                for (int k = lane; k < len[m]; k += 32) { int t = zm[k]; if (t >= 0) cnt += (bits[(size_t)i * words + (t >> 5)] >> (t & 31)) & 1u; }
                for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                const double v = (double)cnt / (double)len[m];
                acc[m * M + i] += v; acc[i * M + m] += v;
            }
        }
        __syncwarp();
    }
    // block reduction in a fixed order: warps write to shared, thread 0 adds them up
    __shared__ double red[32 * MVTM_MAXM * MVTM_MAXM / 4];   // up to 8 warps x 64 pairs
    if (lane == 0) for (int k = 0; k < M * M; k++) red[warp * M * M + k] = acc[k];
    __syncthreads();
    if (threadIdx.x < M * M) {
        double s = 0.0;
        for (int w = 0; w < nwarp; w++) s += red[w * M * M + threadIdx.x];
        part[(size_t)blockIdx.x * M * M + threadIdx.x] = s;
    }
}

// histogram of the strictly positive n_wk cell values (optimizeBeta's countHistogram, M:2295-2309)
__global__ void k_value_hist(int V, int K, int Kp, const int *nwk, int max_value, unsigned long long *hist)
{
    const long long n = (long long)V * Kp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % Kp), c = nwk[i];
        if (t < K && c > 0 && c <= max_value) atomicAdd(hist + c, 1ull);
    }
}

// ---- host -------------------------------------------------------------------------------------------------------
// Multi-rank runs: the statistics gathered from this handle's documents pass through the caller's reducer (sum / max over
// the ranks) before they are used, so every rank derives the same hyper-parameters (mvtm_set_stat_reducer).
static int reduce_stats(mvtm_handle *h, int op, long long *ints, long long n_ints, double *reals, long long n_reals)
{
    if (!h->reducer) return h->comm ? comm_reduce_stats(h, op, ints, n_ints, reals, n_reals) : MVTM_OK;   // NCCL inside the library
    if (h->reducer(h->reducer_ctx, op, (int64_t *)ints, n_ints, reals, n_reals) != 0)
        FAIL(h, MVTM_ERR_STATE, "mvtm_optimize_hyper: the statistics reducer reported a failure");
    return MVTM_OK;
}
static int global_max_len(mvtm_handle *h, int m, int &ml)
{
    long long v = h->v[m].max_len;
    if (int rc = reduce_stats(h, 1, &v, 1, nullptr, 0)) return rc;
    ml = (int)v;
    return MVTM_OK;
}

// topicDocCounts of view m over ALL ranks' documents: K x (global max length + 1)
static int doc_topic_hist_host(mvtm_handle *h, int m, std::vector<long long> &hist, int &stride)
{
    int ml_local = 0, ml = 0;
    if (int rc = mvtm_doc_topic_hist(h, m, nullptr, &ml_local)) return rc;
    if (int rc = global_max_len(h, m, ml)) return rc;
    const int ls = ml_local + 1;
    std::vector<int> local((size_t)h->K * ls, 0);
    if (int rc = mvtm_doc_topic_hist(h, m, local.data(), &ml_local)) return rc;
    stride = ml + 1;
    hist.assign((size_t)h->K * stride, 0);
    for (int t = 0; t < h->K; t++)
        for (int i = 1; i < ls; i++) hist[(size_t)t * stride + i] = local[(size_t)t * ls + i];      // bins c >= 1 (bin 0 is never read, M:2461)
    return reduce_stats(h, 0, hist.data(), (long long)hist.size(), nullptr, 0);
}

extern "C" int mvtm_p_statistics(mvtm_handle *h, double *psum_out, int64_t *docs_per_view_out)
{
    if (!h) return MVTM_ERR_ARG;
    if (int rc = require_views(h, "mvtm_p_statistics")) return rc;
    CK(h, cudaSetDevice(h->device));
    const int M = h->M;
    if (docs_per_view_out) for (int m = 0; m < M; m++) docs_per_view_out[m] = h->v[m].docs_present;
    if (!psum_out) return MVTM_OK;
    const int warps = 8, blocks = h->num_sms * 2;
    SweepParams P;
    fill_params(h, 0, 0, 0, P);
    double *d_part = nullptr;
    CK(h, cudaMalloc(&d_part, (size_t)blocks * M * M * 8));
    const size_t smem = (size_t)warps * M * ((h->K + 31) / 32) * 4;
    k_p_stats<<<blocks, warps * 32, smem, h->stream>>>(M, h->K, h->D, P, d_part);
    std::vector<double> part((size_t)blocks * M * M);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(part.data(), d_part, part.size() * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_part);
    CK(h, e);
    for (int k = 0; k < M * M; k++) { double s = 0.0; for (int b = 0; b < blocks; b++) s += part[(size_t)b * M * M + k]; psum_out[k] = s; }
    return MVTM_OK;
}

static int optimize_p(mvtm_handle *h)
{   // M:2698-2819
    const int M = h->M;
    std::vector<double> psum((size_t)M * M);
    std::vector<long long> docs((size_t)M);
    if (int rc = mvtm_p_statistics(h, psum.data(), nullptr)) return rc;
    for (int m = 0; m < M; m++) docs[(size_t)m] = h->v[m].docs_present;
    if (int rc = reduce_stats(h, 0, docs.data(), M, psum.data(), (long long)M * M)) return rc;
    for (int m = 0; m < M; m++) {
        h->pMean[m][m] = 1.0;
        for (int i = m + 1; i < M; i++) {
            const double denom = (double)std::min(docs[(size_t)m], docs[(size_t)i]);
            const double mean = psum[(size_t)m * M + i] / denom;                            // M:2793
            h->pMean[m][i] = h->pMean[i][m] = mean;
            const double a = (mean == 1.0) ? 5000.0 : -1.0 / std::log(mean);                // M:2797
            h->p_a[m][i] = h->p_a[i][m] = std::min(a, 100.0);                                // M:2805-2806
            h->p_b[m][i] = h->p_b[i][m] = 1.0;
        }
    }
    return MVTM_OK;
}

// The values optimizeDP / optimizeGamma read and write (M:2369-2591), detached from the handle: the engine points this at its own
// fields, the CPU test hook at plain arrays -- one body of code for both.
struct HyperState {
    int M, K;
    double *alpha;                       // M x (K+1)
    double *alphaSum, *gamma, *gammaView, *tablesCnt;   // [M]
    double *gammaRoot, *rootTablesCnt;
    std::vector<int> *inactive;
};

// hist_of(m, hist, stride): topicDocCounts of view m as K rows of `stride` bins (bin c = documents holding the topic c times)
template <class HistFn>
static int optimize_dp_core(HyperState &S, HistFn &&hist_of, OptRng &g)
{   // M:2440-2591
    const int M = S.M, K = S.K;
    std::vector<std::vector<double>> mk((size_t)M, std::vector<double>((size_t)K + 1, 0.0));
    std::vector<double> mk_root((size_t)K + 1, 0.0);
    std::vector<char> active((size_t)K, 0);
    std::vector<long long> hist; int stride = 0;
    for (int m = 0; m < M; m++) {
        if (int rc = hist_of(m, hist, stride)) return rc;
        for (int t = 0; t < K; t++) {
            const double ga = S.gamma[m] * S.alpha[(size_t)m * (K + 1) + t];
            for (int i = 1; i < stride; i++) {
                const long long cnt = hist[(size_t)t * stride + i];
                if (cnt <= 0) continue;
                active[(size_t)t] = 1;
                if (i > 1) {
                    int tbl;
                    try { tbl = rand_antoniak(g, ga, i); } catch (...) { tbl = 1; }              // M:2468-2475
                    mk[m][t] += (double)cnt * tbl;
                } else mk[m][t] += cnt;                                                          // M:2481-2485
            }
        }
    }
    for (int t = 0; t < K; t++)
        for (int m = 0; m < M; m++) {
            if (mk[m][t] > 1) {
                int tbl;
                try { tbl = rand_antoniak(g, *S.gammaRoot, (int)std::ceil(mk[m][t])); } catch (...) { tbl = 1; }
                mk_root[t] += tbl;
            } else if (mk[m][t] == 1) mk_root[t] += 1;
        }
    mk_root[(size_t)K] = *S.gammaRoot;                                                          // M:2520
    *S.rootTablesCnt = std::accumulate(mk_root.begin(), mk_root.end(), 0.0);
    std::vector<double> v((size_t)K + 1, 0.0), tt;
    for (int s = 0; s < 10; s++) { sample_dirichlet(g, mk_root, tt); for (int k = 0; k <= K; k++) v[k] += tt[k] / 10.0; }
    S.inactive->clear();
    for (int t = 0; t < K; t++) if (!active[(size_t)t]) S.inactive->push_back(t);
    for (int m = 0; m < M; m++) {
        for (int t = 0; t < K; t++) mk[m][t] += v[t] * *S.gammaRoot;                            // M:2553
        mk[m][(size_t)K] = S.gammaView[m] + v[(size_t)K] * *S.gammaRoot;                        // M:2557
        S.tablesCnt[m] = std::accumulate(mk[m].begin(), mk[m].end(), 0.0);
        double *al = &S.alpha[(size_t)m * (K + 1)];
        std::fill(al, al + K + 1, 0.0);
        double asum = 0.0;
        for (int s = 0; s < 10; s++) { sample_dirichlet(g, mk[m], tt); for (int k = 0; k <= K; k++) { al[k] += tt[k] / 10.0; asum += tt[k] / 10.0; } }
        S.alphaSum[m] = asum;
    }
    return MVTM_OK;
}

// lencnt_of(m, lencnt): docLengthCounts of view m (M:626), bin j = documents of length j (all ranks)
template <class LenFn>
static int optimize_gamma_core(HyperState &S, LenFn &&lencnt_of, OptRng &g)
{   // M:2369-2438 (Escobar & West 1995 / Teh et al. 2006 auxiliary-variable updates)
    const int K = S.K;
    const double aalpha = 5, balpha = 0.1, agamma = 5, bgamma = 0.1;
    const int R = 10;
    for (int r = 0; r < R; r++) {
        const double eta = rand_beta(g, *S.gammaRoot + 1, *S.rootTablesCnt);
        const double bloge = bgamma - std::log(eta);
        const double pie = 1.0 / (1.0 + (*S.rootTablesCnt * bloge / (agamma + K - 1)));
        const int u = rand_bernoulli(g, pie);
        *S.gammaRoot = rand_gamma(g, agamma + K - 1 + u) * (1.0 / bloge);
    }
    std::vector<long long> lencnt;
    for (int m = 0; m < S.M; m++) {
        if (int rc = lencnt_of(m, lencnt)) return rc;
        for (int r = 0; r < R; r++) {
            const double prev = S.gamma[m];
            const double eta = rand_beta(g, S.gammaView[m] + 1, S.tablesCnt[m]);
            const double bloge = bgamma - std::log(eta);
            const double pie = 1.0 / (1.0 + (S.tablesCnt[m] * bloge / (agamma + K - 1)));
            const int u = rand_bernoulli(g, pie);
            S.gammaView[m] = rand_gamma(g, agamma + K - 1 + u) * (1.0 / bloge);
            double qs = 0.0, qw = 0.0;
            for (size_t j = 1; j < lencnt.size(); j++)            // j = 0: Bernoulli(0) = 0 and log Beta(.,0) = log 1 = 0
                for (long long i = 0; i < lencnt[j]; i++) {
                    qs += rand_bernoulli(g, (double)j / ((double)j + S.gamma[m]));
                    qw += std::log(rand_beta(g, S.gamma[m] + 1, (double)j));
                }
            S.gamma[m] = rand_gamma(g, aalpha + S.tablesCnt[m] - qs) * (1.0 / (balpha - qw));
            if (S.gamma[m] == 0 || !std::isfinite(S.gamma[m])) S.gamma[m] = prev;              // M:2425-2428
        }
    }
    return MVTM_OK;
}

static HyperState hyper_state_of(mvtm_handle *h)
{
    return HyperState{ h->M, h->K, h->alpha.data(), h->alphaSum, h->gamma, h->gammaView, h->tablesCnt, &h->gammaRoot, &h->rootTablesCnt, &h->inactive };
}

static int optimize_dp(mvtm_handle *h, OptRng &g)
{
    HyperState S = hyper_state_of(h);
    return optimize_dp_core(S, [h](int m, std::vector<long long> &hist, int &stride) { return doc_topic_hist_host(h, m, hist, stride); }, g);
}

static int optimize_gamma(mvtm_handle *h, OptRng &g)
{
    HyperState S = hyper_state_of(h);
    return optimize_gamma_core(S, [h](int m, std::vector<long long> &lencnt) -> int {
        ViewDev &v = h->v[m];
        int gml = 0;
        if (int rc = global_max_len(h, m, gml)) return rc;
        lencnt.assign((size_t)gml + 1, 0);                                                // docLengthCounts, M:626 (all ranks)
        for (long long d = 0; d < h->D; d++) {
            const long long len = v.h_doc_off[(size_t)d + 1] - v.h_doc_off[(size_t)d];
            if (len > 0 || v.h_present[(size_t)d]) lencnt[(size_t)len]++;
        }
        return reduce_stats(h, 0, lencnt.data(), (long long)lencnt.size(), nullptr, 0);
    }, g);
}

static int optimize_beta(mvtm_handle *h)
{   // M:2288-2367
    const int K = h->K;
    for (int m = 0; m < h->M; m++) {
        ViewDev &v = h->v[m];
        const double prevBetaSum = h->betaSum[m];
        std::vector<int> nk((size_t)K);
        if (int rc = wait_view_ready(h, m)) return rc;              // an overlapped exchange may still own the view's tables
        CK(h, cudaMemcpyAsync(nk.data(), v.nk, (size_t)K * 4, cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        const int maxTopicSize = *std::max_element(nk.begin(), nk.end());
        std::vector<long long> sizeHist((size_t)maxTopicSize + 1, 0);
        for (int t = 0; t < K; t++) sizeHist[(size_t)nk[(size_t)t]]++;
        // a cell cannot exceed its topic's size, so maxTopicSize bounds the value histogram (the reference sizes it
        // by maxTypeCount, M:2295; bins above the largest cell are zero either way)
        std::vector<long long> countHist((size_t)maxTopicSize + 1, 0);
        {
            unsigned long long *d_hist = nullptr;
            CK(h, cudaMalloc(&d_hist, countHist.size() * 8));
            cudaError_t e = cudaMemsetAsync(d_hist, 0, countHist.size() * 8, h->stream);
            if (e == cudaSuccess) { k_value_hist<<<h->num_sms * 8, 256, 0, h->stream>>>(v.V, K, h->Kp, v.nwk, maxTopicSize, d_hist); e = cudaGetLastError(); }
            if (e == cudaSuccess) e = cudaMemcpyAsync(countHist.data(), d_hist, countHist.size() * 8, cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            cudaFree(d_hist);
            CK(h, e);
        }
        double bs = learn_symmetric_concentration(countHist, sizeHist, v.V, h->betaSum[m]);     // M:2327
        if (bs < v.V * 0.0001) { h->beta[m] = 0.0001; h->betaSum[m] = h->beta[m] * v.V; }        // M:2332-2336 (Q6 sentinel)
        else if (std::isnan(bs)) {
            if (h->beta[m] == 0.01) { h->beta[m] = 0.0001; h->betaSum[m] = h->beta[m] * v.V; }   // M:2340-2344
            else { h->betaSum[m] = prevBetaSum; h->beta[m] = prevBetaSum / v.V; }
        } else { h->betaSum[m] = bs; h->beta[m] = bs / v.V; }
    }
    return MVTM_OK;
}

extern "C" int mvtm_set_stat_reducer(mvtm_handle *h, mvtm_stat_reducer fn, void *ctx)
{
    if (!h) return MVTM_ERR_ARG;
    h->reducer = fn; h->reducer_ctx = ctx;
    return MVTM_OK;
}

extern "C" int mvtm_optimize_hyper(mvtm_handle *h, int32_t iteration, uint32_t which)
{
    if (!h) return MVTM_ERR_ARG;
    if (int rc = require_views(h, "mvtm_optimize_hyper")) return rc;
    CK(h, cudaSetDevice(h->device));
    if (int rc = wait_all_ready(h)) return rc;
    OptRng g{ h->seed, (uint32_t)iteration, 0 };
    if (which & MVTM_OPT_P) if (int rc = optimize_p(h)) return rc;
    if (which & MVTM_OPT_DP) if (int rc = optimize_dp(h, g)) return rc;
    if (which & MVTM_OPT_GAMMA) if (int rc = optimize_gamma(h, g)) return rc;
    if (which & MVTM_OPT_BETA) if (int rc = optimize_beta(h)) return rc;
    h->hyper_dirty = true;
    return MVTM_OK;
}

extern "C" int mvtm_get_hyper_full(mvtm_handle *h, double *alpha, double *alpha_sum, double *beta, double *beta_sum, double *gamma,
                                   double *p_a, double *p_b, double *p_mean, double *gamma_root, double *gamma_view, double *tables_cnt)
{
    if (!h) return MVTM_ERR_ARG;
    const int M = h->M;
    if (alpha) memcpy(alpha, h->alpha.data(), h->alpha.size() * 8);
    for (int m = 0; m < M; m++) {
        if (alpha_sum) alpha_sum[m] = h->alphaSum[m];
        if (beta) beta[m] = h->beta[m];
        if (beta_sum) beta_sum[m] = h->betaSum[m];
        if (gamma) gamma[m] = h->gamma[m];
        if (gamma_view) gamma_view[m] = h->gammaView[m];
        if (tables_cnt) tables_cnt[m] = h->tablesCnt[m];
        for (int j = 0; j < M; j++) {
            if (p_a) p_a[m * M + j] = h->p_a[m][j];
            if (p_b) p_b[m * M + j] = h->p_b[m][j];
            if (p_mean) p_mean[m * M + j] = h->pMean[m][j];
        }
    }
    if (gamma_root) *gamma_root = h->gammaRoot;
    return MVTM_OK;
}

// test hooks for the samplers (distribution tests against exact laws)
extern "C" int mvtm_test_sampler(uint64_t seed, int32_t which, double a, double b, int32_t n, double *out)
{
    OptRng g{ seed, 0u, 0 };
    for (int i = 0; i < n; i++) {
        switch (which) {
            case 0: out[i] = g.next(); break;
            case 1: out[i] = rand_gamma(g, a); break;
            case 2: out[i] = rand_beta(g, a, b); break;
            case 3: try { out[i] = rand_antoniak(g, a, (int)b); } catch (...) { out[i] = 1; } break;
            case 4: {   // the sweep kernel's MVTM_FLAG_BETA_MALLET draw (mallet_next_beta, mvtm_kernels.cuh) evaluated on the host
                BetaStream st{ (uint32_t)i, 0u, PURPOSE_PDRAW, (uint32_t)seed, (uint32_t)(seed >> 32), 0u, make_uint4(0u, 0u, 0u, 0u), 3 };
                out[i] = mallet_next_beta(st, a, b);
                break;
            }
            default: return MVTM_ERR_ARG;
        }
    }
    return MVTM_OK;
}
// optimizeDP (which & MVTM_OPT_DP) then optimizeGamma (which & MVTM_OPT_GAMMA) on plain host arrays with SCRIPTED draws: the same
// core the engine runs, no device.  hist[m] is K x stride[m] (topicDocCounts), lencnt[m] has n_len[m] bins (docLengthCounts).
// scal = { gammaRoot, rootTablesCnt } in/out.  arg_log receives 3 doubles per consumed draw; *n_used the number consumed.
// Returns MVTM_ERR_ARG when the script is too short.
extern "C" int mvtm_test_hyper_core(int32_t M, int32_t K, uint32_t which, const int64_t *const *hist, const int32_t *stride,
                                    const int64_t *const *lencnt, const int32_t *n_len, double *alpha, double *alpha_sum, double *gamma,
                                    double *gamma_view, double *tables_cnt, double *scal, int32_t *inactive, int32_t *n_inactive,
                                    const double *script, int64_t script_len, double *arg_log, int64_t *n_used)
{
    if (M < 1 || K < 1 || !alpha || !alpha_sum || !gamma || !gamma_view || !tables_cnt || !scal || !script) return MVTM_ERR_ARG;
    OptRng g{ 0, 0u, 0 };
    g.script = script; g.script_len = script_len; g.arg_log = arg_log;
    std::vector<int> inact;
    HyperState S{ M, K, alpha, alpha_sum, gamma, gamma_view, tables_cnt, &scal[0], &scal[1], &inact };
    if (which & MVTM_OPT_DP) {
        if (!hist || !stride) return MVTM_ERR_ARG;
        optimize_dp_core(S, [&](int m, std::vector<long long> &h, int &st) { st = stride[m]; h.assign(hist[m], hist[m] + (size_t)K * st); return 0; }, g);
        if (inactive) for (size_t i = 0; i < inact.size(); i++) inactive[i] = inact[i];
        if (n_inactive) *n_inactive = (int32_t)inact.size();
    }
    if (which & MVTM_OPT_GAMMA) {
        if (!lencnt || !n_len) return MVTM_ERR_ARG;
        optimize_gamma_core(S, [&](int m, std::vector<long long> &l) { l.assign(lencnt[m], lencnt[m] + n_len[m]); return 0; }, g);
    }
    if (n_used) *n_used = g.script_pos;
    return g.overrun ? MVTM_ERR_ARG : MVTM_OK;
}
extern "C" double mvtm_test_learn_symmetric_concentration(const int64_t *count_hist, int32_t n_count, const int64_t *length_hist, int32_t n_length,
                                                          int32_t num_dimensions, double current)
{
    std::vector<long long> a(count_hist, count_hist + n_count), b(length_hist, length_hist + n_length);
    return learn_symmetric_concentration(a, b, num_dimensions, current);
}
