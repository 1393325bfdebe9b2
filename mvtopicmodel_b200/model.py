"""Host-side mirror of the reference's trainer class for the sampling path.

`FastQMVWVParallelTopicModel` keeps the method names, argument meaning and the iteration schedule of
org.madgik.MVTopicModel.FastQMVWVParallelTopicModel (M) for the calls on the hot path -- constructor M:183-247,
setters M:273-335, addInstances M:396-533, estimate M:1033-1356, modelLogLikelihood M:3322-3452 -- and routes
them to the C ABI (include/mvtm.h).  The readers that turn the sampled state into the reference's text files
(printState M:3269-3320, printTypeTopicCounts M:2076-2102, printTopicWordWeights M:2104-2129, top words M:1792-1890)
are host code in state_io.py.  Everything outside the path (DB output, diagnostics, embeddings) is out of scope
 and raises NotImplementedError instead of silently doing something else.
"""
import numpy as np

from . import state_io
from .engine import Engine


class Instance:
    """Stand-in for a MALLET Instance whose data is a FeatureSequence: a name (entity id) + word ids."""
    __slots__ = ("name", "features")

    def __init__(self, name, features):
        self.name = str(name)
        self.features = np.asarray(features, dtype=np.int32)


class InstanceList(list):
    """Stand-in for a MALLET InstanceList of one view: a list of Instance + the size of its data alphabet."""

    def __init__(self, instances=(), alphabet_size=None):
        super().__init__(instances)
        self._alphabet_size = alphabet_size

    def alphabet_size(self):
        if self._alphabet_size is not None:
            return int(self._alphabet_size)
        return int(max((int(i.features.max()) for i in self if len(i.features)), default=-1) + 1)


class FastQMVWVParallelTopicModel:
    UNASSIGNED_TOPIC = -1

    def __init__(self, numTopics, numModalities, alpha=0.1, beta=0.01, useCycleProposals=False, SQLConnectionString="",
                 useTypeVectors=False, vectorsLambda=0.0, trainTypeVectors=False):
        if useTypeVectors or trainTypeVectors or vectorsLambda:
            raise NotImplementedError("word-embedding mixing (W:504-507) is outside the accelerated path; keep vectorsLambda = 0")
        if useCycleProposals:
            raise NotImplementedError("cycle proposals are commented out in the reference (W:614-858)")
        self.numTopics, self.numModalities = int(numTopics), int(numModalities)
        K, M = self.numTopics, self.numModalities
        # M:195-239
        self.alpha = np.full((M, K + 1), float(alpha))
        self.alphaSum = np.full(M, float(alpha) * K)
        self.beta = np.full(M, float(beta))
        self.betaSum = np.zeros(M)
        self.gamma = np.ones(M)
        self.p_a = np.full((M, M), 0.2)
        self.p_b = np.ones((M, M))
        self.numIterations, self.burninPeriod, self.optimizeInterval = 1000, 200, 50      # M:111-126
        self.showTopicsInterval, self.wordsPerTopic = 50, 7
        self.randomSeed, self.numThreads = -1, 1
        self.perplexities = np.zeros((M, 200))                                                  # M:234 (Q10)
        self.data = []            # entity ids in document order (M:67)
        self.numTypes = [0] * M
        self.totalTokens = [0] * M
        self.engine = None
        self.device = 0
        self.iterationsSoFar = 0
        self.gammaRoot, self.gammaView, self.pMean, self.inActiveTopicIndex = 10.0, np.zeros(M), np.eye(M), []
        self.sweep_ms = []
        self.saveStateInterval, self.stateFilename = 0, None
        self.engineFlags = 0      # mvtm_config.flags; setReferenceCompat(True) = the reference's behaviour incl. quirks Q1 and Q5
        self.alphabet = [None] * M      # per view: object with lookup_object(i) (ingest.Alphabet) or None -> the id as text
        self.views, self.present = None, None

    # ---- setters, M:273-335 ------------------------------------------------------------------------
    def setNumIterations(self, n): self.numIterations = int(n)
    def setBurninPeriod(self, n): self.burninPeriod = int(n)
    def setTopicDisplay(self, interval, n): self.showTopicsInterval, self.wordsPerTopic = int(interval), int(n)
    def setRandomSeed(self, seed): self.randomSeed = int(seed)

    def setReferenceCompat(self, on=True):
        """Not in the reference (it IS the reference): sample with its dense-index quirk Q1 (W:441-468, W:560-584) and MALLET's
        Beta law for the view-coupling draw (W:333, Q5) instead of the intended semantics.  Call before addInstances."""
        from ._lib import FLAG_REFERENCE_COMPAT
        self.engineFlags = (self.engineFlags | FLAG_REFERENCE_COMPAT) if on else (self.engineFlags & ~FLAG_REFERENCE_COMPAT)

    def setOptimizeInterval(self, interval): self.optimizeInterval = int(interval)
    def setNumThreads(self, threads): self.numThreads = int(threads)   # kept for API parity; the GPU bounds asynchrony itself
    def setSymmetricAlpha(self, b): pass
    def setSaveState(self, interval, filename): self.saveStateInterval, self.stateFilename = int(interval), filename    # M:320-323
    def setSaveSerializedModel(self, interval, filename): raise NotImplementedError("Java serialisation is out of scope")

    # ---- addInstances, M:396-533 ---------------------------------------------------------------------
    def addInstances(self, training, batchId="", vectorSize=0, previousModel=None):
        if previousModel is not None:
            raise NotImplementedError("warm start from a previous model's trees (M:488-496) is not on the accelerated path")
        M, K = self.numModalities, self.numTopics
        if len(training) != M:
            raise ValueError("one InstanceList per modality is required")
        entityPosition, docs = {}, []           # M:398, M:437-455: view 0 always appends, views m > 0 join by name
        for m in range(M):
            self.numTypes[m] = training[m].alphabet_size()
            self.betaSum[m] = self.beta[m] * self.numTypes[m]                                   # M:420
            for inst in training[m]:
                if m != 0 and inst.name in entityPosition:
                    docs[entityPosition[inst.name]][m] = inst.features
                else:
                    row = [None] * M
                    row[m] = inst.features
                    docs.append(row)
                    entityPosition[inst.name] = len(docs) - 1
                    self.data.append(inst.name)
        D = len(docs)
        views, present = [], []
        for m in range(M):
            lens = np.array([0 if r[m] is None else len(r[m]) for r in docs], dtype=np.int64)
            off = np.zeros(D + 1, dtype=np.int64)
            np.cumsum(lens, out=off[1:])
            words = np.concatenate([r[m] for r in docs if r[m] is not None and len(r[m])]) if off[-1] else np.zeros(0, np.int32)
            views.append((off, words.astype(np.int32)))
            present.append(np.array([r[m] is not None for r in docs], dtype=np.uint8))
            self.totalTokens[m] = int(off[-1])
        seed = self.randomSeed if self.randomSeed != -1 else int(np.random.SeedSequence().entropy % (1 << 63))
        for m in range(M):
            self.alphabet[m] = getattr(training[m], "alphabet", None)
        self.views, self.present = views, present
        self.engine = Engine(K, [max(1, v) for v in self.numTypes], views, seed=seed, device=self.device, present=present,
                             flags=self.engineFlags)
        self._push_hyper()
        self.engine.init_assignments()          # M:465-515 + buildInitialTypeTopicCounts M:600-652

    def _push_hyper(self):
        self.engine.set_hyper(alpha=self.alpha, alphaSum=self.alphaSum, beta=self.beta, betaSum=self.betaSum, gamma=self.gamma,
                              p_a=self.p_a, p_b=self.p_b)

    # ---- estimate, M:1033-1356 -------------------------------------------------------------------------
    def estimate(self):
        if self.engine is None:
            raise RuntimeError("addInstances must be called before estimate")
        self.p_a[:] = 0.2; self.p_b[:] = 1.0                                                        # M:1055-1058
        self._push_hyper()
        for iteration in range(1, self.numIterations + 1):                                          # M:1146
            if self.saveStateInterval != 0 and iteration % self.saveStateInterval == 0:             # M:1154-1155 (before the sweep)
                self.printState(f"{self.stateFilename}.{iteration}")
            if iteration < self.burninPeriod and self.numModalities > 1:
                self.p_a[:] = min(iteration / 100.0 + 0.3, 1.1)                                     # M:1166-1169
                self._push_hyper()
            elif iteration > self.burninPeriod and self.optimizeInterval != 0 and iteration % self.optimizeInterval == 0:
                self.optimizeHyperParameters(iteration)                                            # M:1173-1210
            self.engine.sweep(iteration)                                                           # M:1213-1239
            self.sweep_ms.append(self.engine.stats()["ms_total"])
            if iteration % 10 == 0:                                                                # M:1296-1304
                ll = self.modelLogLikelihood()
                if iteration // 10 < self.perplexities.shape[1]:
                    for m in range(self.numModalities):
                        self.perplexities[m, iteration // 10] = ll[m] / max(1, self.totalTokens[m])
            self.iterationsSoFar = iteration
            alpha, alphaSum, ina = self.engine.get_hyper()    # topics activated by the sweep (U:263-270)
            self.alpha, self.alphaSum, self.inActiveTopicIndex = alpha, alphaSum, list(ina)

    def optimizeHyperParameters(self, iteration):
        """optimizeP / optimizeDP / optimizeGamma / optimizeBeta in the reference's order (M:1173-1210), run by the engine's
        host code from device statistics (mvtm_optimize_hyper); the new values are mirrored into this object."""
        self.engine.optimize_hyper(iteration)
        self._pull_hyper()

    def optimizeP(self, appendMetadata=False):
        self.engine.optimize_hyper(self.iterationsSoFar, 1); self._pull_hyper()

    def optimizeBeta(self):
        self.engine.optimize_hyper(self.iterationsSoFar, 8); self._pull_hyper()

    def _pull_hyper(self):
        hf = self.engine.get_hyper_full()
        self.alpha, self.alphaSum, self.beta, self.betaSum, self.gamma = hf["alpha"], hf["alphaSum"], hf["beta"], hf["betaSum"], hf["gamma"]
        self.p_a, self.p_b, self.pMean = hf["p_a"], hf["p_b"], hf["pMean"]
        self.gammaRoot, self.gammaView, self.inActiveTopicIndex = hf["gammaRoot"], hf["gammaView"], list(hf["inactive"])

    # ---- readers ---------------------------------------------------------------------------------------
    def modelLogLikelihood(self, quirk_len2=False):
        return self.engine.loglik(quirk_len2)

    @property
    def typeTopicCounts(self):
        return [self.engine.get_counts(m)[0] for m in range(self.numModalities)]

    @property
    def tokensPerTopic(self):
        return [self.engine.get_counts(m, want_nwk=False)[1] for m in range(self.numModalities)]

    def topicDocCounts(self, m):
        return self.engine.doc_topic_hist(m)

    def getTopicAssignments(self, m):
        return self.engine.get_assignments(m)

    def docLengthCounts(self, m):
        """Histogram of document lengths of view m (M:107, filled at M:626): bin j = documents of length j that have the view."""
        off = self.views[m][0]
        lens = (off[1:] - off[:-1])[np.asarray(self.present[m], dtype=bool) | ((off[1:] - off[:-1]) > 0)]
        return np.bincount(lens)

    def getTopicProbabilities(self, instance, modality=None):
        """M:2134-2171.  getTopicProbabilities(instanceID) -> one smoothed topic distribution per view of that training document
        (a view the document lacks raises, as the reference's null Assignments entry does);
        getTopicProbabilities(topics, modality) -> the distribution of one topic sequence (e.g. from the inferencer)."""
        if modality is not None:
            return state_io.topic_probabilities(instance, self.numTopics, self.gamma[modality], self.alpha[modality])
        d, out = int(instance), []
        for m in range(self.numModalities):
            off = self.views[m][0]
            if not (self.present[m][d] or off[d + 1] > off[d]):
                raise ValueError(f"document {d} has no view {m}")
            z = self.engine.get_assignments(m)[off[d]:off[d + 1]]
            out.append(state_io.topic_probabilities(z, self.numTopics, self.gamma[m], self.alpha[m]))
        return out

    def getInferencer(self):
        """M:3457-3463"""
        return FastQMVWVTopicInferencer(self)

    # ---- text readers (state_io.py) -------------------------------------------------------------------
    def _lookup(self, m):
        a = self.alphabet[m]
        return (lambda i: str(a.lookup_object(int(i)))) if a is not None else (lambda i: str(int(i)))

    def _lookups(self):
        return [self._lookup(m) for m in range(self.numModalities)]

    def getSortedWords(self, modality):
        """M:1792-1809: per topic the (type, count) pairs with positive count in MALLET IDSorter order."""
        nwk = self.engine.get_counts(modality)[0]
        return [state_io.sorted_words(nwk, t) for t in range(self.numTopics)]

    def getTopWords(self, numWords, modality):
        """M:1819-1845 (argument order as in the reference: numWords, modality)."""
        return state_io.top_words(self.engine.get_counts(modality)[0], numWords, self._lookup(modality))

    def displayTopWords(self, numWords, numLabels=0, usingNewLines=False):
        """M:1851-1888"""
        return state_io.display_top_words(self.typeTopicCounts, self.alpha, self._lookups(), numWords, usingNewLines)

    def printState(self, f):
        """M:3269-3320: path -> gzip file (printState(File)); file object -> plain text (printState(PrintStream))."""
        zs = [self.engine.get_assignments(m) for m in range(self.numModalities)]
        args = (self.views, zs, self.present, self._lookups(), self.gamma, self.alpha, self.beta)
        if isinstance(f, (str, bytes)):
            state_io.write_state_gz(f, *args)
        else:
            state_io.write_state(f, *args)

    def readState(self, f):
        """Restores the assignments from a printState file of the same corpus (a Java run's or our own) and rebuilds the
        counts (the initialisation path M:534-573 takes for saved states)."""
        zs, header = state_io.read_state(f, self.views, self.present)
        for m in range(self.numModalities):
            self.engine.set_assignments(m, zs[m])
        return header

    def printTypeTopicCounts(self, path):
        """M:2076-2102"""
        with open(path, "w", encoding="utf-8") as out:
            state_io.write_type_topic_counts(out, self.typeTopicCounts, self._lookups())

    def printTopicWordWeights(self, f):
        """M:2104-2129"""
        if isinstance(f, (str, bytes)):
            with open(f, "w", encoding="utf-8") as out:
                state_io.write_topic_word_weights(out, self.typeTopicCounts, self.beta, self._lookups())
        else:
            state_io.write_topic_word_weights(f, self.typeTopicCounts, self.beta, self._lookups())


def split_for_completion(views):
    """Document completion split: within every document-view the tokens at even positions are OBSERVED (folded in), those at
    odd positions are EVALUATED.  Returns (observed views, evaluation views), both doc-aligned CSR like the input."""
    obs, ev = [], []
    for off, w in views:
        lens = (off[1:] - off[:-1]).astype(np.int64)
        pos = np.arange(len(w), dtype=np.int64) - np.repeat(off[:-1], lens)
        is_obs = (pos % 2) == 0
        for keep, dst in ((is_obs, obs), (~is_obs, ev)):
            cnt = np.zeros(len(off) - 1, dtype=np.int64)
            if len(w):
                np.add.at(cnt, np.repeat(np.arange(len(off) - 1), lens), keep.astype(np.int64))
            noff = np.zeros(len(off), dtype=np.int64)
            np.cumsum(cnt, out=noff[1:])
            dst.append((noff, np.ascontiguousarray(w[keep], dtype=np.int32)))
    return obs, ev


class FastQMVWVTopicInferencer:
    """Mirror of org.madgik.MVTopicModel.FastQMVWVTopicInferencer (I) for the sampling path: folds NEW documents into a
    trained model with the global counts frozen (the worker with nut = 0, I:211-256).

    Built from a trained FastQMVWVParallelTopicModel like M:3457-3463 (`model.getInferencer()`).  `inferTopicDistributions`
    follows I:114-330: join the new documents' views, draw every in-vocabulary token's topic from the trained topic-word
    distribution (I:186-203), run `numIterations` = 10 sweeps (I:74,561), return the document-topic proportions of I:385-412.
    quirk_bare_trees=True reproduces Q13 (the inferencer's trees omit gamma*alpha)."""

    def __init__(self, model):
        self.K, self.M = model.numTopics, model.numModalities
        self.numTypes = list(model.numTypes)
        self.counts = [model.engine.get_counts(m) for m in range(self.M)]
        hf = model.engine.get_hyper_full()
        self.hyper = {k: hf[k] for k in ("alpha", "alphaSum", "beta", "betaSum", "gamma", "p_a", "p_b")}
        self.inactive = list(hf["inactive"])
        self.pMean = hf["pMean"] if model.iterationsSoFar > model.burninPeriod else np.ones((self.M, self.M))
        self.numIterations = 10
        self.device = model.device
        self.seed = model.randomSeed if model.randomSeed != -1 else 1
        self.engine = None

    def _join(self, instances):
        """Views of the new documents joined by name exactly as addInstances does (M:437-455) -> (names, CSR views)."""
        M = self.M
        entityPosition, docs, names = {}, [], []
        for m in range(M):
            for inst in instances[m]:
                if m != 0 and inst.name in entityPosition:
                    docs[entityPosition[inst.name]][m] = inst.features
                else:
                    row = [None] * M; row[m] = inst.features
                    docs.append(row); entityPosition[inst.name] = len(docs) - 1; names.append(inst.name)
        D = len(docs)
        views = []
        for m in range(M):
            lens = np.array([0 if r[m] is None else len(r[m]) for r in docs], dtype=np.int64)
            off = np.zeros(D + 1, dtype=np.int64); np.cumsum(lens, out=off[1:])
            words = np.concatenate([r[m] for r in docs if r[m] is not None and len(r[m])]) if off[-1] else np.zeros(0, np.int32)
            views.append((off, words.astype(np.int32)))
        return names, views

    def _fold_in(self, views, quirk_bare_trees=False):
        """I:114-330: trained counts frozen, tree-draw initialisation, numIterations sweeps over `views`."""
        self.engine = e = Engine(self.K, [max(1, v) for v in self.numTypes], views, seed=self.seed, device=self.device)
        e.set_hyper(inactive=self.inactive, **self.hyper)
        for m in range(self.M):
            e.set_counts(m, *self.counts[m])
        e.init_assignments_from_counts()
        for it in range(1, self.numIterations + 1):
            e.sweep(it, update_global=2 if quirk_bare_trees else 0)
        return e

    def heldOutPerplexity(self, instances, numSamples=1, lastSweeps=1):
        """Per-view held-out perplexity by document completion (this build's estimator: the reference never evaluates one,
        S:191): even-position tokens of every held-out document-view are folded in with the inferencer's frozen sweeps,
        odd-position tokens are scored by mvtm_heldout_loglik.  One fold-in is a single Gibbs sample; `numSamples` fold-ins
        with different seeds x the last `lastSweeps` sweeps of each are averaged (in the log domain) to cut its variance.
        Returns (perplexity[M], tokens scored[M])."""
        _, views = self._join(instances)
        obs, ev = split_for_completion(views)
        acc, cnt, seed0 = np.zeros(self.M), np.zeros(self.M, dtype=np.int64), self.seed
        for s in range(int(numSamples)):
            self.seed = seed0 + s
            self.engine = e = Engine(self.K, [max(1, v) for v in self.numTypes], obs, seed=self.seed, device=self.device)
            e.set_hyper(inactive=self.inactive, **self.hyper)
            for m in range(self.M):
                e.set_counts(m, *self.counts[m])
            e.init_assignments_from_counts()
            for it in range(1, self.numIterations + 1):
                e.sweep(it, update_global=0)
                if it > self.numIterations - int(lastSweeps):
                    for m in range(self.M):
                        ll, n = e.heldout_loglik(m, ev[m][0], ev[m][1])
                        cnt[m] = n
                        acc[m] += ll / max(n, 1)
        self.seed = seed0
        ppl = np.exp(-acc / (int(numSamples) * int(lastSweeps)))
        ppl[cnt == 0] = np.nan
        return ppl, cnt

    def inferTopicDistributions(self, instances, discrWeightPerModality=None, quirk_bare_trees=False):
        M, K = self.M, self.K
        entityPosition, docs, names = {}, [], []
        for m in range(M):
            for inst in instances[m]:
                if m != 0 and inst.name in entityPosition:
                    docs[entityPosition[inst.name]][m] = inst.features
                else:
                    row = [None] * M; row[m] = inst.features
                    docs.append(row); entityPosition[inst.name] = len(docs) - 1; names.append(inst.name)
        D = len(docs)
        views = []
        for m in range(M):
            lens = np.array([0 if r[m] is None else len(r[m]) for r in docs], dtype=np.int64)
            off = np.zeros(D + 1, dtype=np.int64); np.cumsum(lens, out=off[1:])
            words = np.concatenate([r[m] for r in docs if r[m] is not None and len(r[m])]) if off[-1] else np.zeros(0, np.int32)
            views.append((off, words.astype(np.int32)))
        self.engine = e = Engine(K, [max(1, v) for v in self.numTypes], views, seed=self.seed, device=self.device)
        e.set_hyper(inactive=self.inactive, **self.hyper)
        for m in range(M):
            e.set_counts(m, *self.counts[m])
        e.init_assignments_from_counts()
        for it in range(1, self.numIterations + 1):
            e.sweep(it, update_global=2 if quirk_bare_trees else 0)
        w = np.ones(M) if discrWeightPerModality is None else np.asarray(discrWeightPerModality, dtype=np.float64)
        theta = np.zeros((D, K))
        norm = np.zeros(D)
        for m in range(M):
            off = views[m][0]
            z = e.get_assignments(m)
            lens = (off[1:] - off[:-1]).astype(np.float64)
            has = lens > 0
            cnt = np.zeros((D, K))
            np.add.at(cnt, (np.repeat(np.arange(D), (off[1:] - off[:-1])), z), 1.0)
            wm = (1.0 if m == 0 else w[m]) * self.pMean[0, m]
            ga = self.hyper["gamma"][m] * self.hyper["alpha"][m][:K]
            contrib = wm * (cnt + ga) / (lens + self.hyper["gamma"][m] * self.hyper["alphaSum"][m])[:, None]
            theta[has] += contrib[has]
            norm[has] += wm
        theta[norm > 0] /= norm[norm > 0][:, None]
        return names, theta
