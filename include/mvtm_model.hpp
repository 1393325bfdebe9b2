// mvtm_model.hpp -- C++ host-side mirror of org.madgik.MVTopicModel.FastQMVWVParallelTopicModel for the sampling
// path, written over the C ABI of mvtm.h (the reference is JVM code and this image has no JVM, so the host side
// above the ABI is C++; INTEGRATION.md shows the JNI / Panama binding a Java maintainer would add instead).
// Same method names, argument meaning and iteration schedule as the reference:
//   constructor M:183-247, setters M:273-335, addInstances M:396-533, estimate M:1033-1356,
//   modelLogLikelihood M:3322-3452  (M = FastQMVWVParallelTopicModel.java).
// Errors: the reference swallows exceptions and logs (W:230-232); this mirror throws std::runtime_error carrying
// mvtm_last_error().
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "mvtm.h"

namespace mvtm {

struct Instance {                 // MALLET Instance with a FeatureSequence payload
    std::string name;             // instance.getName(): the entity id documents are joined on (M:437)
    std::vector<int32_t> features;
};
struct InstanceList {             // one view
    std::vector<Instance> instances;
    int32_t alphabetSize = 0;     // training[m].getDataAlphabet().size(), M:413
};

class FastQMVWVParallelTopicModel {
public:
    int numTopics, numModalities;
    std::vector<double> alpha, alphaSum, beta, betaSum, gamma, p_a, p_b;   // alpha: M x (K+1); p_a/p_b: M x M
    int numIterations = 1000, burninPeriod = 200, optimizeInterval = 50;   // M:111-126
    long long randomSeed = -1;
    int numThreads = 1;
    std::vector<std::vector<double>> perplexities;                          // [M][200], M:234
    std::vector<std::string> data;                                          // entity ids in document order
    std::vector<int64_t> totalTokens;
    int device = 0;

    FastQMVWVParallelTopicModel(int numTopics_, int numModalities_, double alpha_, double beta_, bool useCycleProposals = false,
                                const std::string & /*SQLConnectionString*/ = "", bool useTypeVectors = false,
                                double vectorsLambda = 0.0, bool trainTypeVectors = false)
        : numTopics(numTopics_), numModalities(numModalities_)
    {
        if (useCycleProposals || useTypeVectors || trainTypeVectors || vectorsLambda != 0.0)
            throw std::invalid_argument("cycle proposals / type vectors are outside the accelerated path (SURVEY 8)");
        const int K = numTopics, M = numModalities;
        alpha.assign((size_t)M * (K + 1), alpha_); alphaSum.assign(M, alpha_ * K); beta.assign(M, beta_); betaSum.assign(M, 0.0);
        gamma.assign(M, 1.0); p_a.assign((size_t)M * M, 0.2); p_b.assign((size_t)M * M, 1.0);
        perplexities.assign(M, std::vector<double>(200, 0.0)); totalTokens.assign(M, 0);
    }
    ~FastQMVWVParallelTopicModel() { if (h_) mvtm_destroy(h_); }
    FastQMVWVParallelTopicModel(const FastQMVWVParallelTopicModel &) = delete;
    FastQMVWVParallelTopicModel &operator=(const FastQMVWVParallelTopicModel &) = delete;

    void setNumIterations(int n) { numIterations = n; }
    void setBurninPeriod(int n) { burninPeriod = n; }
    void setRandomSeed(long long s) { randomSeed = s; }
    void setOptimizeInterval(int n) { optimizeInterval = n; }
    void setNumThreads(int n) { numThreads = n; }

    // addInstances M:396-533: join the views by entity id (view 0 always appends; later views attach to an existing id
    // or append), hand the doc-aligned CSR to the engine, random initialisation + initial counts on the device.
    void addInstances(const std::vector<InstanceList> &training, const std::string & /*batchId*/ = "", int /*vectorSize*/ = 0)
    {
        const int M = numModalities;
        if ((int)training.size() != M) throw std::invalid_argument("one InstanceList per modality is required");
        std::map<std::string, size_t> entityPosition;
        std::vector<std::vector<const Instance *>> docs;
        std::vector<int32_t> V(M);
        for (int m = 0; m < M; m++) {
            V[m] = std::max<int32_t>(1, training[m].alphabetSize);
            betaSum[m] = beta[m] * training[m].alphabetSize;                       // M:420
            for (const Instance &inst : training[m].instances) {
                auto it = entityPosition.find(inst.name);
                if (m != 0 && it != entityPosition.end()) docs[it->second][m] = &inst;
                else {
                    docs.emplace_back(M, nullptr); docs.back()[m] = &inst;
                    entityPosition[inst.name] = docs.size() - 1; data.push_back(inst.name);
                }
            }
        }
        const int64_t D = (int64_t)docs.size();
        mvtm_config cfg{};
        cfg.num_topics = numTopics; cfg.num_views = M; cfg.num_docs = D; cfg.vocab_sizes = V.data();
        cfg.seed = randomSeed == -1 ? 0x9E3779B97F4A7C15ull : (uint64_t)randomSeed; cfg.device = device; cfg.doc_id_stride = 1;
        if (mvtm_create(&cfg, &h_)) throw std::runtime_error(mvtm_last_error(nullptr));
        for (int m = 0; m < M; m++) {
            std::vector<int64_t> off(D + 1, 0); std::vector<int32_t> words; std::vector<uint8_t> present(std::max<int64_t>(D, 1), 0);
            for (int64_t d = 0; d < D; d++) {
                const Instance *in = docs[d][m];
                if (in) { present[d] = 1; words.insert(words.end(), in->features.begin(), in->features.end()); }
                off[d + 1] = (int64_t)words.size();
            }
            totalTokens[m] = (int64_t)words.size();
            ck(mvtm_add_view(h_, m, off.data(), words.data(), present.data()));
        }
        pushHyper();
        ck(mvtm_init_assignments(h_));
    }

    // estimate M:1033-1356: the iteration loop with the burn-in ramp of p_a and the LL series; the body of each
    // iteration (M:1213-1239: worker + updater threads up to the barrier) is one mvtm_sweep.
    void estimate()
    {
        if (!h_) throw std::logic_error("addInstances must be called before estimate");
        std::fill(p_a.begin(), p_a.end(), 0.2); std::fill(p_b.begin(), p_b.end(), 1.0);     // M:1055-1058
        pushHyper();
        for (int iteration = 1; iteration <= numIterations; iteration++) {
            if (iteration < burninPeriod && numModalities > 1)
                std::fill(p_a.begin(), p_a.end(), std::min(iteration / 100.0 + 0.3, 1.1));  // M:1166-1169
            else if (iteration > burninPeriod && optimizeInterval != 0 && iteration % optimizeInterval == 0) {
                // optimizeP, optimizeDP, optimizeGamma, optimizeBeta + buildFTrees(false), M:1173-1210
                ck(mvtm_optimize_hyper(h_, iteration, MVTM_OPT_ALL));
                pullHyper();
            }
            if (iteration < burninPeriod && numModalities > 1) pushHyper();
            ck(mvtm_sweep(h_, iteration, 1));
            if (iteration % 10 == 0 && iteration / 10 < 200) {                               // M:1296-1304
                std::vector<double> ll = modelLogLikelihood();
                for (int m = 0; m < numModalities; m++) perplexities[m][iteration / 10] = ll[m] / std::max<int64_t>(1, totalTokens[m]);
            }
            int32_t nin = 0;
            ck(mvtm_get_hyper(h_, alpha.data(), alphaSum.data(), nullptr, &nin));           // topics activated by the sweep
        }
    }

    std::vector<double> modelLogLikelihood(bool quirkLen2 = false)
    {
        std::vector<double> ll(numModalities);
        ck(mvtm_loglik(h_, ll.data(), quirkLen2 ? 1 : 0));
        return ll;
    }
    std::vector<int32_t> tokensPerTopic(int m) { std::vector<int32_t> nk(numTopics); ck(mvtm_get_counts(h_, m, nullptr, nk.data())); return nk; }
    std::vector<int32_t> typeTopicCounts(int m, int32_t V) { std::vector<int32_t> c((size_t)V * numTopics); ck(mvtm_get_counts(h_, m, c.data(), nullptr)); return c; }
    std::vector<int32_t> topicAssignments(int m) { std::vector<int32_t> z((size_t)totalTokens[m]); ck(mvtm_get_assignments(h_, m, z.data())); return z; }
    int64_t checkInvariants() { int64_t v = -1; ck(mvtm_check_invariants(h_, &v)); return v; }
    mvtm_handle *handle() { return h_; }

private:
    mvtm_handle *h_ = nullptr;
    void ck(int rc) { if (rc) throw std::runtime_error(std::string("mvtm status ") + std::to_string(rc) + ": " + mvtm_last_error(h_)); }
    void pullHyper()
    {
        ck(mvtm_get_hyper_full(h_, alpha.data(), alphaSum.data(), beta.data(), betaSum.data(), gamma.data(), p_a.data(), p_b.data(),
                               nullptr, nullptr, nullptr, nullptr));
    }
    void pushHyper()
    {
        ck(mvtm_set_hyper(h_, alpha.data(), alphaSum.data(), beta.data(), betaSum.data(), gamma.data(), p_a.data(), p_b.data(), nullptr, -1));
    }
};

}  // namespace mvtm
