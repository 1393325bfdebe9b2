// C++ test driver for the C-ABI library through the FastQMVWVParallelTopicModel mirror (include/mvtm_model.hpp).
// Builds a small two-view corpus, runs estimate(), checks the count invariants and that LL/token improved.
#include <cstdio>
#include <random>

#include "mvtm_model.hpp"

int main()
{
    const int K = 40, M = 2, D = 2000;
    std::mt19937 rng(7);
    std::vector<mvtm::InstanceList> training(M);
    training[0].alphabetSize = 500; training[1].alphabetSize = 60;
    for (int d = 0; d < D; d++) {
        int topic = d % 10;
        mvtm::Instance text; text.name = "doc" + std::to_string(d);
        int len = 10 + rng() % 30;
        for (int i = 0; i < len; i++) text.features.push_back((topic * 50 + rng() % 50 + (rng() % 5 == 0 ? rng() % 500 : 0)) % 500);
        training[0].instances.push_back(text);
        if (d % 4 != 3) {                                    // a quarter of the documents lack the side view
            mvtm::Instance side; side.name = text.name;
            for (int i = 0; i < 3; i++) side.features.push_back((topic * 6 + rng() % 6) % 60);
            training[1].instances.push_back(side);
        }
    }
    try {
        mvtm::FastQMVWVParallelTopicModel model(K, M, 0.1, 0.01);
        model.setRandomSeed(11);
        model.setNumIterations(30);
        model.setBurninPeriod(50);
        model.addInstances(training, "batch0", 0);
        std::vector<double> ll0 = model.modelLogLikelihood();
        model.estimate();
        std::vector<double> ll1 = model.modelLogLikelihood();
        long long bad = model.checkInvariants();
        std::printf("LL/token view0 %.4f -> %.4f, view1 %.4f -> %.4f, invariant violations %lld\n", ll0[0] / model.totalTokens[0],
                    ll1[0] / model.totalTokens[0], ll0[1] / model.totalTokens[1], ll1[1] / model.totalTokens[1], bad);
        if (bad != 0 || !(ll1[0] > ll0[0]) || !(ll1[1] > ll0[1])) { std::printf("FAIL\n"); return 1; }
        std::printf("OK\n");
    } catch (const std::exception &e) {
        std::printf("exception: %s\n", e.what());
        return 2;
    }
    return 0;
}
