"""Golden vectors of the SAMPLER produced by the reference's own binary: FastQMVWVWorkerRunnable.sampleTopicsForOneDoc from
output/MVTopicModel-1.0-SNAPSHOT.jar (a prebuilt, slightly older build of W:301-597 -- same statements, one shared queue, int[][]
tokensPerTopic) is EXECUTED by tools/jvm_mini.py document after document, sweep after sweep, on small corpora.

What is real and what is shimmed
  executed from the jar : the whole per-document sampler (view-coupling matrix, local counts, dense topic index and its in-sweep
                          maintenance incl. the dead insertion code Q1, other-view mass, new-topic mass, per-token masses, bucket
                          choice, lower_bound, FTree.sample), FastQDelta, FTree (construction, sample, update)
  shimmed (host Python) : the containers the sampler reads (ArrayList, MALLET Instance / FeatureSequence / LabelSequence) as thin
                          views over the same arrays; ThreadLocalRandom.nextDouble and Randoms.nextBeta, which return the uniforms
                          the ORACLE draws for the same (token position, document, iteration, view) -- Philox4x32-10, 24-bit, see
                          oracle/mvtm_oracle.c orc_draw -- so that both consume identical randomness (the reference's own RNG is
                          unseedable, Q9); the queue between the two runnables: Queue.add hands the delta (plus the end-of-sweep
                          sentinel) straight to FastQMVWVUpdaterRunnable.run -- ALSO executed from the jar (U:164-297: counts,
                          totals, topicDocCounts histogram, the two F+tree leaves, activation of inactive topics) -- so every
                          delta is applied at once; the initial tables and histograms come from the jar's
                          initializeHistograms + buildInitialTypeTopicCounts (M:849-897, M:600-652), the F+trees from its
                          recalcTrees (this build's name for buildFTrees, M:2660-2696)

Output: tests/golden/reference_sampler_vectors.json -- corpus, hyper-parameters, initial assignments, and the assignments after
every sweep.  tests/test_reference_vectors.py replays them through the C oracle (reference-faithful mode) and demands equality
token for token.  Needs /root/reference (build container only).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import jvm_mini  # noqa: E402
from jvm_mini import JObject  # noqa: E402
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/output"
W = "org/madgik/MVTopicModel/FastQMVWVWorkerRunnable"
FT = "org/madgik/utils/FTree"
UP = "org/madgik/MVTopicModel/FastQMVWVUpdaterRunnable"
MC = "org/madgik/MVTopicModel/FastQMVWVParallelTopicModel"
PURPOSE_SAMPLE, PURPOSE_PDRAW = 0, 2


def u24(x):
    return float(int(x) >> 8) * (1.0 / 16777216.0)


class RefSampler:
    """One model state driven through the jar's sampler."""

    def __init__(self, K, Vs, views, z0, seed, alpha, alphaSum, beta, gamma, p_a, p_b, inactive=()):
        self.vm = vm = jvm_mini.MiniJVM([os.path.join(REF, "lib", "mallet-2.0.8.jar"), os.path.join(REF, "MVTopicModel-1.0-SNAPSHOT.jar")])
        self.K, self.M, self.Vs, self.views, self.seed = K, len(Vs), Vs, views, seed
        M = self.M
        self.D = len(views[0][0]) - 1
        self.z = [[int(t) for t in z0[m]] for m in range(M)]
        self.nwk = [[[0] * K for _ in range(Vs[m])] for m in range(M)]       # filled by the jar's buildInitialTypeTopicCounts below
        self.nk = [[0] * K for _ in range(M)]
        self.alpha = [list(map(float, a)) for a in alpha]
        self.alphaSum, self.beta, self.gamma = list(map(float, alphaSum)), list(map(float, beta)), list(map(float, gamma))
        self.betaSum = [self.beta[m] * Vs[m] for m in range(M)]
        self.inactive = list(inactive)
        # per-document containers: data.get(d).Assignments[m] = {instance -> FeatureSequence shim, topicSequence -> LabelSequence shim}
        self.docs = []
        for d in range(self.D):
            ent = JObject("org/madgik/utils/MixTopicModelTopicAssignment")
            arr = []
            for m in range(M):
                b, e = int(views[m][0][d]), int(views[m][0][d + 1])
                if e == b and m > 0:
                    arr.append(None)                       # the document lacks this view (MA:13-19)
                    continue
                ta = JObject("cc/mallet/topics/TopicAssignment")
                zslice = self.z[m][b:e]                    # LabelSequence.getFeatures(): a LIVE array the sampler writes into
                ta.fields["topicSequence"] = ("labels", m, b, zslice)
                ta.fields["instance"] = ("instance", [int(x) for x in views[m][1][b:e]])
                arr.append(ta)
            ent.fields["Assignments"] = arr
            self.docs.append(ent)
        self.iteration, self.doc = 0, 0
        self.counters = {"new": 0, "doc": 0, "tree": 0, "tree_bucket": 0, "deltas": 0, "beta_draws": 0}
        sh = vm.shims
        sh["java/util/ArrayList.get:(I)Ljava/lang/Object;"] = lambda loc, r, a, pc: r[1][a[0]]
        sh["cc/mallet/types/LabelSequence.getFeatures:()[I"] = lambda loc, r, a, pc: r[3]
        sh["cc/mallet/types/Instance.getData:()Ljava/lang/Object;"] = lambda loc, r, a, pc: ("fs", r[1])
        sh["cc/mallet/types/FeatureSequence.getLength:()I"] = lambda loc, r, a, pc: len(r[1])
        sh["cc/mallet/types/FeatureSequence.getIndexAtPosition:(I)I"] = lambda loc, r, a, pc: r[1][a[0]]
        sh["java/util/List.isEmpty:()Z"] = lambda loc, r, a, pc: int(len(self.inactive) == 0)
        sh["java/util/List.get:(I)Ljava/lang/Object;"] = lambda loc, r, a, pc: (("queue",) if r[0] == "queues" else self.inactive[a[0]])
        sh["java/util/List.size:()I"] = lambda loc, r, a, pc: (1 if r[0] == "queues" else len(self.inactive))
        unbox = lambda v: v.fields["value"] if isinstance(v, JObject) else v
        sh["java/util/List.contains:(Ljava/lang/Object;)Z"] = lambda loc, r, a, pc: int(unbox(a[0]) in self.inactive)

        def list_remove(loc, r, a, pc):
            v = unbox(a[0])
            if v in self.inactive:
                self.inactive.remove(v); return 1
            return 0
        sh["java/util/List.remove:(Ljava/lang/Object;)Z"] = list_remove
        sh["java/lang/Integer.valueOf:(I)Ljava/lang/Integer;"] = lambda loc, r, a, pc: a[0]

        def integer_init(loc, r, a, pc):
            r.fields["value"] = a[0]
        sh["java/lang/Integer.<init>:(I)V"] = integer_init
        sh["java/util/Queue.poll:()Ljava/lang/Object;"] = lambda loc, r, a, pc: (self.pending.pop(0) if self.pending else None)

        def set_add(loc, r, a, pc):
            r.fields.setdefault("items", set()).add(a[0]); return 1
        sh["java/util/Set.add:(Ljava/lang/Object;)Z"] = set_add
        sh["java/util/Set.size:()I"] = lambda loc, r, a, pc: len(r.fields.get("items", ()))
        sh["java/lang/Thread.currentThread:()Ljava/lang/Thread;"] = lambda loc, r, a, pc: ("thread",)

        sh["java/lang/Thread.sleep:(J)V"] = lambda loc, r, a, pc: None      # U:276-280: the 20 ms nap after every pass over the queues
        sh["java/util/concurrent/CyclicBarrier.await:()I"] = lambda loc, r, a, pc: 0
        sh["org/apache/log4j/Logger.info:(Ljava/lang/Object;)V"] = lambda loc, r, a, pc: None
        vm.statics[(UP, "logger")] = JObject("logger")
        sh["java/lang/Integer.intValue:()I"] = lambda loc, r, a, pc: r
        sh["java/util/concurrent/ThreadLocalRandom.current:()Ljava/util/concurrent/ThreadLocalRandom;"] = lambda loc, r, a, pc: ("tlr",)
        sh["java/util/concurrent/ThreadLocalRandom.nextDouble:()D"] = self.next_double
        sh["cc/mallet/util/Randoms.nextBeta:(DD)D"] = self.next_beta
        def count_bucket(loc, r, a, pc):       # newMassCnt / topicDocMassCnt / wordFTreeMassCnt by call site (W:523, W:530, W:533)
            self.counters[{1257: "new", 1326: "doc", 1350: "tree_bucket"}[pc]] += 1
            return 0
        sh["java/util/concurrent/atomic/AtomicInteger.getAndIncrement:()I"] = count_bucket
        sh["java/util/Queue.add:(Ljava/lang/Object;)Z"] = self.apply_delta
        for d in ("(Ljava/lang/String;)Ljava/lang/StringBuilder;", "(D)Ljava/lang/StringBuilder;", "(I)Ljava/lang/StringBuilder;"):
            sh["java/lang/StringBuilder.append:" + d] = lambda loc, r, a, pc: r
        sh["java/lang/StringBuilder.toString:()Ljava/lang/String;"] = lambda loc, r, a, pc: ""
        sh["java/io/PrintStream.println:(Ljava/lang/String;)V"] = lambda loc, r, a, pc: None

        def boom(loc, r, a, pc):
            raise RuntimeError("the reference's sampler threw inside sampleTopicsForOneDoc")
        sh["java/lang/Exception.printStackTrace:()V"] = boom

        # counts, totals, topicDocCounts and docLengthCounts by the jar's own initialisation code (M:849-897, M:600-652)
        model = _model_object(self)
        vm.call(MC, "initializeHistograms", "()V", [model])
        vm.call(MC, "buildInitialTypeTopicCounts", "()V", [model])
        self.hist, self.doc_len_counts = model.fields["topicDocCounts"], model.fields["docLengthCounts"]
        # F+trees by the jar's recalcTrees(true) (this build's name for buildFTrees, M:2660-2696): leaves
        # gamma*alpha*(n_wk+beta)/(n_k+betaSum), 0 for inactive topics, one `new FTree(temp)` per word type
        self.trees = [[None] * Vs[m] for m in range(M)]
        model.fields["trees"], model.fields["inActiveTopicIndex"] = self.trees, ("inactive",)
        model.fields["docSmoothingOnlyMass"] = [0.0] * M
        self.model = model
        vm.call(MC, "recalcTrees", "(Z)V", [model, 1])
        # the updater (U:78-147 fields set directly): one queue, drained by running its run() after every enqueue
        self.pending = []
        up = JObject(UP)
        up.fields.update(dict(typeTopicCounts=self.nwk, tokensPerTopic=self.nk, trees=self.trees, queues=("queues",), alpha=self.alpha,
                              alphaSum=self.alphaSum, beta=self.beta, betaSum=self.betaSum, gamma=self.gamma, numTopics=K,
                              numModalities=M, numTypes=list(Vs), topicDocCounts=self.hist, docLengthCounts=self.doc_len_counts,
                              inActiveTopicIndex=("inactive",), optimizeParams=0, isFinished=1, useCycleProposals=0,
                              cyclicBarrier=("barrier",)))
        self.updater = up
        wk = JObject(W)
        wk.fields.update(dict(data=("arraylist", self.docs), numModalities=M, numTopics=K, alpha=self.alpha, alphaSum=self.alphaSum,
                              beta=self.beta, betaSum=self.betaSum, gamma=self.gamma, p_a=[list(map(float, r)) for r in p_a],
                              p_b=[list(map(float, r)) for r in p_b], typeTopicCounts=self.nwk, tokensPerTopic=self.nk, trees=self.trees,
                              random=("randoms",), queue=("queue",), inActiveTopicIndex=("inactive",), useTypeVectors=0,
                              useTypeVectorsProb=0.0, typeTopicSimilarity=None, threadId=0))
        self.worker = wk

    def rebuild_trees(self):
        """recalcTrees(false) executed from the jar: what estimate() does after every optimise step (M:1209)."""
        self.vm.call(MC, "recalcTrees", "(Z)V", [self.model, 0])

    # --- randomness: exactly the oracle's draws (oracle/mvtm_oracle.c: orc_draw, draw_p) ---------------------------------
    def philox(self, pos, view_or_pair, purpose):
        ctr = [pos, self.doc, self.iteration, (view_or_pair << 8) | purpose]
        return O.philox(ctr, [self.seed & 0xFFFFFFFF, (self.seed >> 32) & 0xFFFFFFFF])

    def next_double(self, loc, recv, args, pc):
        m, pos = loc[20], loc[22]          # the sampler's view and position loop variables (bytecode locals 20 and 22)
        x = self.philox(pos, m, PURPOSE_SAMPLE)
        if pc == 1223:                     # u = ThreadLocalRandom.nextDouble() of W:517
            return u24(x[0])
        if pc == 1357:                     # u2 of W:534 (the F+tree bucket)
            self.counters["tree"] += 1
            return u24(x[1])
        raise RuntimeError(f"unexpected nextDouble call site {pc}")

    def next_beta(self, loc, recv, args, pc):
        m, j = loc[18], loc[19]            # W:327-337 loop variables
        self.counters["beta_draws"] += 1
        x = self.philox(0, m * self.M + j, PURPOSE_PDRAW)
        return u24(x[0]) ** (1.0 / args[0])            # Beta(a, 1) by inversion, the oracle's default law (Q5)

    # --- the queue: every delta goes straight through the jar's FastQMVWVUpdaterRunnable.run (U:164-297) ---------------------
    def apply_delta(self, loc, recv, args, pc):
        sentinel = JObject("org/madgik/utils/FastQDelta")                 # W:215-222: FastQDelta(-1, -1, -1, -1, ...) ends a worker's stream
        sentinel.fields.update(dict(NewTopic=-1, OldTopic=-1, Type=-1, Modality=-1, DocOldTopicCnt=-1, DocNewTopicCnt=-1))
        self.pending.extend([args[0], sentinel])
        self.updater.fields["isFinished"] = 1
        self.vm.call(UP, "run", "()V", [self.updater])
        assert not self.pending
        self.counters["deltas"] += 1
        return 1

    def sweep(self, iteration):
        self.iteration = iteration
        for d in range(self.D):
            self.doc = d
            self.vm.call(W, "sampleTopicsForOneDoc", "(I)V", [self.worker, d])
        # gather the live per-document arrays back into CSR order
        for d, ent in enumerate(self.docs):
            for m, ta in enumerate(ent.fields["Assignments"]):
                if ta is not None:
                    _, _, b, zs = ta.fields["topicSequence"]
                    self.z[m][b:b + len(zs)] = zs
        return [list(z) for z in self.z]


def reference_loglik(ref):
    """FastQMVWVParallelTopicModel.modelLogLikelihood (M:3322-3452) EXECUTED from the jar on the sampler's current state.  The
    documents are presented the way MALLET holds them: a LabelSequence's backing array has capacity max(length, 2)
    (FeatureSequence(Alphabet, int) allocates max(capacity, 2) ints), which is what quirk Q18 is about."""
    vm, M, K = ref.vm, ref.M, ref.K
    docs = []
    for d in range(ref.D):
        ent = JObject("org/madgik/utils/MixTopicModelTopicAssignment")
        arr = []
        for m in range(M):
            b, e = int(ref.views[m][0][d]), int(ref.views[m][0][d + 1])
            if e == b and m > 0:
                arr.append(None)
                continue
            ta = JObject("cc/mallet/topics/TopicAssignment")
            feats = list(ref.z[m][b:e])
            ta.fields["topicSequence"] = ("labels", m, b, feats + [0] * (max(2, len(feats)) - len(feats)))
            arr.append(ta)
        ent.fields["Assignments"] = arr
        docs.append(ent)
    model = JObject(MC)
    model.fields.update(dict(numModalities=M, numTopics=K, data=("arraylist", docs), typeTopicCounts=ref.nwk, tokensPerTopic=ref.nk,
                             alpha=ref.alpha, alphaSum=ref.alphaSum, beta=ref.beta, betaSum=ref.betaSum, gamma=ref.gamma,
                             numTypes=list(ref.Vs)))
    sh = vm.shims
    sh["java/util/ArrayList.size:()I"] = lambda loc, r, a, pc: len(r[1])
    sh["java/lang/Byte.valueOf:(B)Ljava/lang/Byte;"] = lambda loc, r, a, pc: a[0]
    sh["java/lang/Byte.byteValue:()B"] = lambda loc, r, a, pc: r
    sh["java/lang/StringBuilder.append:(Ljava/lang/Object;)Ljava/lang/StringBuilder;"] = lambda loc, r, a, pc: r
    sh["org/apache/log4j/Logger.info:(Ljava/lang/Object;)V"] = lambda loc, r, a, pc: None

    def warn(loc, r, a, pc):
        raise RuntimeError("modelLogLikelihood logged a warning (NaN / infinite term)")
    sh["org/apache/log4j/Logger.warn:(Ljava/lang/Object;)V"] = warn
    vm.statics[(MC, "logger")] = JObject("logger")
    old = vm.strict_fields
    vm.strict_fields = True
    try:
        return list(vm.call(MC, "modelLogLikelihood", "()[D", [model]))
    finally:
        vm.strict_fields = old


def _model_object(ref, docs=None):
    """A FastQMVWVParallelTopicModel object over the sampler's state (fields set directly, no constructor)."""
    vm, M, K = ref.vm, ref.M, ref.K
    model = JObject(MC)
    model.fields.update(dict(numModalities=M, numTopics=K, data=("arraylist", docs if docs is not None else ref.docs),
                             typeTopicCounts=ref.nwk, tokensPerTopic=ref.nk, alpha=ref.alpha, alphaSum=ref.alphaSum, beta=ref.beta,
                             betaSum=ref.betaSum, gamma=ref.gamma, numTypes=list(ref.Vs), totalTokens=[len(z) for z in ref.z],
                             maxTypeCount=[int(max(sum(row) for row in ref.nwk[m])) for m in range(M)], formatter=("nf",),
                             docLengthCounts=None, topicDocCounts=None, histogramSize=[0] * M,
                             totalDocsPerModality=[sum(1 for e in ref.docs if e.fields["Assignments"][m] is not None) for m in range(M)]))
    sh = vm.shims
    sh["java/util/ArrayList.size:()I"] = lambda loc, r, a, pc: len(r[1])
    sh["java/lang/Byte.valueOf:(B)Ljava/lang/Byte;"] = lambda loc, r, a, pc: a[0]
    sh["java/lang/Byte.byteValue:()B"] = lambda loc, r, a, pc: r
    sh["java/lang/StringBuilder.append:(Ljava/lang/Object;)Ljava/lang/StringBuilder;"] = lambda loc, r, a, pc: r
    sh["org/apache/log4j/Logger.info:(Ljava/lang/Object;)V"] = lambda loc, r, a, pc: None
    sh["org/apache/log4j/Logger.warn:(Ljava/lang/Object;)V"] = lambda loc, r, a, pc: None
    sh["java/text/NumberFormat.format:(D)Ljava/lang/String;"] = lambda loc, r, a, pc: ""
    sh["java/text/NumberFormat.format:(Ljava/lang/Object;)Ljava/lang/String;"] = lambda loc, r, a, pc: ""
    sh[MC + ".appendMetadata:(Ljava/lang/String;)V"] = lambda loc, r, a, pc: None
    sh["cc/mallet/types/FeatureSequence.getFeatures:()[I"] = lambda loc, r, a, pc: (r[3] if r[0] == "labels" else r[1])
    sh["cc/mallet/types/FeatureSequence.size:()I"] = lambda loc, r, a, pc: len(r[1])

    class _It:
        def __init__(self, lst):
            self.l, self.i = lst, 0
    sh["java/util/ArrayList.iterator:()Ljava/util/Iterator;"] = lambda loc, r, a, pc: _It(r[1])
    sh["java/util/Iterator.hasNext:()Z"] = lambda loc, r, a, pc: int(r.i < len(r.l))

    def _next(loc, r, a, pc):
        v = r.l[r.i]; r.i += 1
        return v
    sh["java/util/Iterator.next:()Ljava/lang/Object;"] = _next
    vm.statics[(MC, "logger")] = JObject("logger")
    return model


def reference_counts_and_histograms(ref):
    """initializeHistograms (M:849-897) + buildInitialTypeTopicCounts (M:600-652) EXECUTED from the jar over the current
    assignments, on COPIES of the tables: type-topic counts, topic totals, topicDocCounts[m][t][c] and docLengthCounts."""
    import copy
    model = _model_object(ref)
    model.fields["typeTopicCounts"] = copy.deepcopy(ref.nwk)
    model.fields["tokensPerTopic"] = copy.deepcopy(ref.nk)
    # the per-document arrays must show the CURRENT assignments (the sampler's live slices already do)
    ref.vm.strict_fields = True
    try:
        ref.vm.call(MC, "initializeHistograms", "()V", [model])
        ref.vm.call(MC, "buildInitialTypeTopicCounts", "()V", [model])
    finally:
        ref.vm.strict_fields = False
    return {"typeTopicCounts": model.fields["typeTopicCounts"], "tokensPerTopic": model.fields["tokensPerTopic"],
            "topicDocCounts": model.fields["topicDocCounts"], "docLengthCounts": model.fields["docLengthCounts"]}


def reference_optimize_beta(ref):
    """optimizeBeta (M:2288-2367) EXECUTED from the jar (it calls MALLET's learnSymmetricConcentration from the MALLET jar) on
    copies of beta / betaSum: returns the values it would install."""
    model = _model_object(ref)
    model.fields["beta"], model.fields["betaSum"] = list(ref.beta), list(ref.betaSum)
    ref.vm.strict_fields = True
    try:
        ref.vm.call(MC, "optimizeBeta", "()V", [model])
    finally:
        ref.vm.strict_fields = False
    return {"beta": model.fields["beta"], "betaSum": model.fields["betaSum"]}


def reference_conditionals(ref, iteration, max_tokens=600, rebuild=True):
    """north_star check (b) against the reference itself: one sweep of the jar's sampler with the GLOBAL counts frozen (deltas
    dropped, the inferencer's nut = 0 mode, W:587) while the per-token masses it computes are read out of its frame at the moment
    it draws u (W:517): dense index S, cumulative document masses (W:496-513), new-topic mass C (W:515) and the leaves of the
    word's F+tree (the B bucket).  Net conditional P(t) = (A_t + leaf_t [+ C on the first inactive topic]) / total.  A token is
    recorded only when S equals the set of topics the document currently holds (so the dead insertion code Q1 has had no
    effect on it) together with the document's assignments at that moment, from which any implementation can rebuild n_d.  Tokens
    on which Q1 HAS had an effect are recorded too, with the list of held topics the index lacks (`not_in_S`)."""
    vm, K, M = ref.vm, ref.K, ref.M
    recs = []
    if rebuild:                  # (the inferencer's own trees -- make_reference_inference_vectors.py -- must stay as they are)
        ref.rebuild_trees()      # fresh trees: during a sweep only two leaves per delta are refreshed (Q3), the check is on frozen, consistent state
    saved_apply = vm.shims["java/util/Queue.add:(Ljava/lang/Object;)Z"]
    vm.shims["java/util/Queue.add:(Ljava/lang/Object;)Z"] = lambda loc, r, a, pc: 1          # counts stay frozen
    saved_nd = vm.shims["java/util/concurrent/ThreadLocalRandom.nextDouble:()D"]

    def nd(loc, recv, args, pc):
        r = saved_nd(loc, recv, args, pc)
        if pc == 1223 and len(recs) < max_tokens:
            m, pos, nz = loc[20], loc[22], loc[19]
            S, cum, C = loc[6][:nz], loc[7][:nz], loc[25]
            counts = loc[13]
            held = sorted(t for t in range(K) if any(counts[i][t] != 0 for i in range(M)))
            q1 = list(S) != held
            if q1 and not set(S) <= set(held):
                raise RuntimeError("the reference's dense index holds a topic no view holds")
            if (not q1) or len([r for r in recs if r.get("not_in_S")]) < max_tokens // 2:
                leaves = loc[11].fields["tree"][K:2 * K]
                mass = [float(x) for x in leaves]
                prev = 0.0
                for t, c in zip(S, cum):
                    mass[t] += c - prev; prev = c
                total = sum(mass) + C
                probs = [x / total for x in mass]
                if C > 0:
                    probs[ref.inactive[0]] += C / total
                ent = ref.docs[ref.doc].fields["Assignments"]
                zdoc = [None if ta is None else list(ta.fields["topicSequence"][3]) for ta in ent]
                rec = {"doc": ref.doc, "view": m, "pos": pos, "p_row": [float(x) for x in loc[17][m]], "z_doc": zdoc,
                       "probs": probs, "new_share": C / total}
                if q1:      # quirk Q1 at work: topics the document holds that the reference's dense index lacks (gained this sweep)
                    rec["not_in_S"] = [t for t in held if t not in set(S)]
                recs.append(rec)
        return r
    vm.shims["java/util/concurrent/ThreadLocalRandom.nextDouble:()D"] = nd
    try:
        ref.sweep(iteration)
    finally:
        vm.shims["java/util/Queue.add:(Ljava/lang/Object;)Z"] = saved_apply
        vm.shims["java/util/concurrent/ThreadLocalRandom.nextDouble:()D"] = saved_nd
    return recs


def make_case(name, K, Vs, means, D, seed, sweeps, p_a=0.0, inactive=(), alpha_new=0.1, rng_seed=0, sparse_view=None, unassigned=0):
    from helpers import random_corpus
    views = random_corpus(rng_seed, D, K, Vs, means, empty_frac=0.1)
    M = len(Vs)
    o = O.Oracle(K, Vs, views, seed=seed)
    o.init_assignments()
    z0 = [o.get_assignments(m).tolist() for m in range(M)]
    alpha = np.full((M, K + 1), 0.1); alpha[:, K] = alpha_new
    alphaSum = np.full(M, 0.1 * K)
    beta, gamma = np.full(M, 0.01), np.ones(M)
    pa, pb = np.full((M, M), p_a), np.ones((M, M))
    if inactive:
        # topics without tokens only: move their tokens to topic 0 first
        for m in range(M):
            z0[m] = [0 if t in inactive else t for t in z0[m]]
    if sparse_view is not None:
        beta[sparse_view] = 0.0001                      # the "too sparse" sentinel of optimizeBeta (M:2332-2336): W:335-336 zero p
    if unassigned:
        for m in range(M):                              # UNASSIGNED_TOPIC tokens (W:434: no decrement, delta with OldTopic = -1)
            for i in range(0, len(z0[m]), unassigned):
                z0[m][i] = -1
    ref = RefSampler(K, Vs, views, z0, seed, alpha, alphaSum, beta, gamma, pa, pb, inactive)
    out = {"name": name, "K": K, "V": Vs, "seed": seed, "views": [{"off": v[0].tolist(), "word": v[1].tolist()} for v in views],
           "z0": z0, "alpha": alpha.tolist(), "alphaSum": alphaSum.tolist(), "beta": beta.tolist(), "betaSum": ref.betaSum, "gamma": gamma.tolist(),
           "p_a": pa.tolist(), "p_b": pb.tolist(), "inactive": list(inactive), "z_after": []}
    out["loglik_after"] = []
    for it in range(1, sweeps + 1):
        out["z_after"].append(ref.sweep(it))
        out["loglik_after"].append(reference_loglik(ref))          # modelLogLikelihood from the jar on the state just reached
    # frozen-count conditionals of the state just reached (the assignments keep moving, the tables do not)
    out["frozen_counts_z"] = [list(z) for z in ref.z]
    out["frozen_alpha"] = [list(a) for a in ref.alpha]
    out["frozen_inactive"] = list(ref.inactive)
    out["counts_and_histograms"] = reference_counts_and_histograms(ref)
    out["optimize_beta"] = reference_optimize_beta(ref)
    out["conditionals"] = reference_conditionals(ref, sweeps + 1)
    out["counters"] = dict(ref.counters)
    out["hist_maintained"] = [[list(r) for r in h] for h in ref.hist]      # topicDocCounts as the updater left it (U:220-232)
    out["nk_final"] = [list(r) for r in ref.nk]
    print(name, "tokens", [len(z) for z in z0], "sweeps", sweeps, "conditionals", len(out["conditionals"]), ref.counters, "bytecode steps", ref.vm.steps, flush=True)
    return out


def main():
    cases = [
        make_case("single_view", K=8, Vs=[30], means=[7], D=40, seed=11, sweeps=3, rng_seed=1),
        make_case("two_views_coupled", K=10, Vs=[40, 12], means=[8, 3], D=30, seed=12, sweeps=3, p_a=0.7, rng_seed=2),
        make_case("three_views_inactive_topics", K=12, Vs=[30, 10, 8], means=[7, 3, 2], D=25, seed=13, sweeps=2, p_a=1.3,
                  inactive=(4, 9), alpha_new=6.0, rng_seed=3),
        make_case("sparse_sentinel_and_unassigned", K=9, Vs=[25, 10, 6], means=[6, 3, 2], D=25, seed=14, sweeps=2, p_a=0.9,
                  sparse_view=1, unassigned=7, rng_seed=4),
        make_case("k37_longer_docs", K=37, Vs=[120, 30], means=[25, 5], D=20, seed=15, sweeps=2, p_a=0.4, rng_seed=5),
    ]
    json.dump({"source": "output/MVTopicModel-1.0-SNAPSHOT.jar FastQMVWVWorkerRunnable.sampleTopicsForOneDoc executed by tools/jvm_mini.py",
               "cases": cases}, open(os.path.join(HERE, "reference_sampler_vectors.json"), "w"))


if __name__ == "__main__":
    main()
