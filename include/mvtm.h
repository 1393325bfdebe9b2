/*
 * mvtm.h -- C ABI of the B200-native collapsed-Gibbs engine for the multi-view HDP topic model (MViHDP).
 *
 * This is the drop-in boundary for ONE path of hmetaxa/MVTopicModel: the per-token sampling done by
 * FastQMVWVWorkerRunnable and the count updates done by FastQMVWVUpdaterRunnable, as driven by
 * FastQMVWVParallelTopicModel.addInstances()/estimate().  The reference has no FFI of its own (it is pure
 * Java); each entry point below names the reference code whose job it takes over.
 * Reference tags (under /root/reference/src/main/java/org/madgik/):
 *   W = MVTopicModel/FastQMVWVWorkerRunnable.java     U = MVTopicModel/FastQMVWVUpdaterRunnable.java
 *   M = MVTopicModel/FastQMVWVParallelTopicModel.java  I = MVTopicModel/FastQMVWVTopicInferencer.java
 *
 * Conventions: every function returns 0 on success and a non-zero mvtm_status otherwise, never throws or
 * aborts; mvtm_last_error() returns a UTF-8 message for the last failure on that handle (or for a failed
 * mvtm_create when h == NULL).  The caller owns every buffer it passes; inputs are copied.  One caller
 * thread per handle; one handle drives one GPU.  There is NO CPU fallback: without a CUDA device every
 * compute entry point fails with MVTM_ERR_CUDA.
 */
#ifndef MVTM_H
#define MVTM_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVTM_MAX_VIEWS 8
#define MVTM_UNASSIGNED (-1)          /* FastQMVWVParallelTopicModel.UNASSIGNED_TOPIC */

typedef enum {
    MVTM_OK = 0,
    MVTM_ERR_ARG = 1,        /* bad argument (K/M mismatch, NULL, out of range)             */
    MVTM_ERR_STATE = 2,      /* call order (e.g. sweep before every view was added)         */
    MVTM_ERR_CUDA = 3,       /* CUDA runtime failure, including "no device"                 */
    MVTM_ERR_CORRUPT = 4,    /* negative count / invariant violation detected on the device */
    MVTM_ERR_LIMIT = 5       /* a size this build does not support (doc-view > 65535 tokens, K > 2048) */
} mvtm_status;

/* mvtm_config.flags */
#define MVTM_FLAG_DOC_ORDER   1u   /* work list in plain document order instead of longest-first (tests)     */
#define MVTM_FLAG_SINGLE_WARP 2u   /* run each view pass with one warp: fully sequential, deterministic (tests) */
#define MVTM_FLAG_Q1_COMPAT   4u   /* reproduce the reference's dense topic index exactly (quirk Q1): a topic is removed from the
                                      document's index when no view holds it any more (W:441-468) and is NEVER re-inserted (the
                                      insertion code W:563-584 is dead), so a topic gained later in the same sweep keeps only its
                                      F+tree mass until the next sweep.  Default off: the index is "topics the document holds".
                                      Limits a document-view to 32767 tokens. */
#define MVTM_FLAG_BETA_MALLET 8u   /* draw the view-coupling matrix p[m][i] (W:327-337) from the law of MALLET's Randoms.nextBeta, which the
                                      reference calls at W:333, instead of the true Beta(p_a, 1): for p_a > 1 that is a truncated normal
                                      N(1, 0.25/(p_a - 1)) on [0, 1] (quirk Q5: its rejection test compares against NaN).  Same law for
                                      p_a <= 1. */
#define MVTM_FLAG_TMA_RING   16u   /* stage n_wk rows through the shared-memory ring filled by the TMA engine for every K.  Default (flag clear):
                                      for K in (768, 1024] and (1536, 2048] the DIRECT kernel keeps the row of a token in registers
                                      (ld.global.cg one token ahead), which leaves room for more documents per SM -- measured faster
                                      there; compat mode (Q1) always uses the ring.  Same arithmetic and scan order for a given
                                      lane-group size, so frozen sweeps of the two kernels agree token for token. */
#define MVTM_FLAG_REFERENCE_COMPAT (MVTM_FLAG_Q1_COMPAT | MVTM_FLAG_BETA_MALLET)   /* both: the reference's behaviour, quirks included */

typedef struct mvtm_config {
    int32_t num_topics;              /* K, M:183 numTopics                                                   */
    int32_t num_views;               /* M, M:183 numModalities                                               */
    int64_t num_docs;                /* documents held by THIS handle (its shard)                            */
    const int32_t *vocab_sizes;      /* V_m = alphabet[m].size(), M:413                                      */
    uint64_t seed;                   /* Philox key (M:403-408 randomSeed)                                    */
    int32_t device;                  /* CUDA device ordinal                                                  */
    uint32_t flags;
    int64_t doc_id_base;             /* global id of local document d is doc_id_base + d * doc_id_stride;     */
    int64_t doc_id_stride;           /*   it keys the RNG, so a sharded run draws what the unsharded one does */
    int32_t warps_per_cta;           /* 0 = auto                                                             */
    int32_t ring_depth;              /* n_wk rows in flight per warp (TMA ring kernel only), 0 = auto        */
    int32_t max_ctas;                /* 0 = one persistent CTA per SM; smaller values bound the number of documents
                                        sampled concurrently (the asynchrony the reference bounds by numThreads, M:1036) */
} mvtm_config;

typedef struct mvtm_sweep_stats {
    int64_t tokens;                  /* token updates performed by the last sweep (all views)                */
    int64_t changed;                 /* tokens whose topic changed (= deltas of W:587-589)                   */
    int64_t new_topic;               /* draws from the new-topic bucket (W:523 newMassCnt)                   */
    double ms_total;                 /* device time of the last sweep, CUDA events                           */
    double ms_view[MVTM_MAX_VIEWS];  /* device time of each view's sampling kernel                           */
    int32_t kernel_launches;         /* kernels launched by the last sweep                                   */
    int32_t ring_depth[MVTM_MAX_VIEWS];  /* TMA ring depth each view's last pass ran with; 0 = DIRECT kernel     */
    int32_t ring_locked[MVTM_MAX_VIEWS]; /* the depth the autotune settled on (0 = still sampling, or DIRECT kernel; = the configured depth when fixed) */
} mvtm_sweep_stats;

typedef struct mvtm_handle mvtm_handle;

/* Lifecycle.  Replaces the constructor M:183-247 + initSpace M:575-598 (device allocation). */
int mvtm_create(const mvtm_config *cfg, mvtm_handle **out);
int mvtm_destroy(mvtm_handle *h);
const char *mvtm_last_error(const mvtm_handle *h);

/* Corpus.  Replaces the per-view document collection of addInstances M:410-463: view m as a doc-aligned
 * CSR (doc_off has num_docs+1 entries; a document that lacks view m has an empty range).  `present`
 * (nullable, num_docs bytes) marks documents that own an Assignments[m] object even if it is empty
 * (only the log-likelihood quirk Q18 looks at it). */
int mvtm_add_view(mvtm_handle *h, int32_t m, const int64_t *doc_off, const int32_t *word_id, const uint8_t *present);

/* Initial assignments.  mvtm_init_assignments = the random initialisation M:465-515 followed by
 * buildInitialTypeTopicCounts M:600-652; mvtm_set_assignments imports z (state restore, M:534-573) and
 * rebuilds the view's counts. */
int mvtm_init_assignments(mvtm_handle *h);
int mvtm_set_assignments(mvtm_handle *h, int32_t m, const int32_t *z);

/* Inference on new documents (FastQMVWVTopicInferencer, I:114-330): a handle built over the NEW documents receives the trained
 * counts (mvtm_set_counts: typeTopicCounts[m] as V_m x K row-major and tokensPerTopic[m]), draws the initial topic of every
 * in-vocabulary token from the bare topic-word distribution of those counts (mvtm_init_assignments_from_counts, I:186-203;
 * out-of-vocabulary tokens get topic 0, Q13) and is then swept with update_global = 0 (or 2 for the reference's trees without
 * gamma*alpha, Q13). */
int mvtm_set_counts(mvtm_handle *h, int32_t m, const int32_t *n_wk, const int32_t *n_k);
int mvtm_init_assignments_from_counts(mvtm_handle *h);

/* Hyper-parameters as the worker/updater constructors receive them (W:80-150, U:78-147).  Any pointer may
 * be NULL (= keep).  alpha is M x (K+1) row-major (slot K = new-topic prior), p_a/p_b are M x M;
 * inactive = inActiveTopicIndex (M:95) as a list, n_inactive < 0 = keep.  Defaults after create are the
 * driver's: alpha 0.1, alphaSum 0.1*K, beta 0.01, betaSum beta*V, gamma 1, p_a 0.2, p_b 1, no inactive topics. */
int mvtm_set_hyper(mvtm_handle *h, const double *alpha, const double *alpha_sum, const double *beta,
                   const double *beta_sum, const double *gamma, const double *p_a, const double *p_b,
                   const int32_t *inactive, int32_t n_inactive);
int mvtm_get_hyper(mvtm_handle *h, double *alpha, double *alpha_sum, int32_t *inactive, int32_t *n_inactive);

/* One Gibbs sweep over every view = one iteration of estimate()'s loop body M:1213-1239: the work of all
 * FastQMVWVWorkerRunnable.run (W:186-233, W:301-597) and FastQMVWVUpdaterRunnable.run (U:164-297) threads up to
 * the barrier.  update_global = 0 freezes n_wk / n_k (the inferencer's nut = 0 mode, I:211-256); update_global = 2 does the
 * same with the inferencer's own trees, which hold phi without gamma*alpha (I:561-576, quirk Q13).  Blocking. */
int mvtm_sweep(mvtm_handle *h, int32_t iteration, int32_t update_global);

/* The same sweep through HOST buffers: uploads z (one array per view, caller memory, pinned or pageable),
 * rebuilds the counts from it, sweeps, and downloads the new z into the same arrays.  This is the call a
 * stateless host (a JVM holding topicSequence arrays) makes; bench.py's e2e number times it. */
int mvtm_sweep_host(mvtm_handle *h, int32_t iteration, int32_t *const *z_inout);

/* Host mirror of the assignments: z_host (caller-owned, tokens-of-view-m ints, PINNED and MAPPED host memory -- cudaHostAlloc or
 * cudaHostRegister) receives every new assignment that subsequent sweeps of view m store, written by the sweep kernel itself
 * next to its store to device memory (posted PCIe writes under the sampling).  A JVM-side topicSequence buffer thus stays
 * current without a copy after the sweep; the array is complete when the sweep call returns (mvtm_sweep) or mvtm_sweep_finish
 * has returned.  Tokens the sweep skips (out-of-vocabulary words) are not rewritten.  NULL stops mirroring. */
int mvtm_set_host_mirror(mvtm_handle *h, int32_t m, int32_t *z_host);

/* Readers.  z: LabelSequence.getFeatures() of every doc concatenated in CSR order; n_wk: typeTopicCounts[m]
 * as V_m x K row-major by word (M:584); n_k: tokensPerTopic[m]. */
int mvtm_get_assignments(mvtm_handle *h, int32_t m, int32_t *z_out);
int mvtm_get_counts(mvtm_handle *h, int32_t m, int32_t *n_wk_out, int32_t *n_k_out);

/* topicDocCounts[m][t][c] (U:220-232, M:647-649), recomputed from the assignments: K x (max_len+1) ints,
 * bin c = number of documents in which topic t holds exactly c tokens of view m.  *max_len_out receives the
 * longest document of the view; call with hist_out == NULL to query it. */
int mvtm_doc_topic_hist(mvtm_handle *h, int32_t m, int32_t *hist_out, int32_t *max_len_out);

/* modelLogLikelihood M:3322-3452 with MALLET's logGammaStirling; ll_out has M entries.  quirk_len2 != 0
 * reproduces Q18 (phantom topic-0 tokens of documents shorter than two tokens). */
int mvtm_loglik(mvtm_handle *h, double *ll_out, int32_t quirk_len2);
/* The same value split for multi-rank hosts: doc_part sums over THIS handle's documents (M:3341-3373), word_part is the
 * topic-word part (M:3389-3441), a function of the count tables only -- identical on every rank once the counts are global.
 * Global log-likelihood = sum over ranks of doc_part + word_part of any one rank. */
int mvtm_loglik_parts(mvtm_handle *h, double *doc_part_out, double *word_part_out, int32_t quirk_len2);

/* Held-out evaluation by document completion.  The reference constructs MALLET's MarginalProbEstimator (M:3470-3478) but never
 * calls it (S:191 hard-codes perplexity = 0), so the estimator is this build's own and is applied identically to the CPU oracle:
 * the handle holds the OBSERVED part of every held-out document (folded in with mvtm_set_counts + mvtm_init_assignments_from_counts
 * + frozen sweeps, I:114-330), eval_off / eval_word is a doc-aligned CSR of the evaluation tokens of view m, and every
 * in-vocabulary evaluation token w of document d scores
 *     log sum_t (n_wk[w][t] + beta) / (n_k[t] + betaSum) * (n_d[t] + gamma*alpha[t]) / (N_obs + sum_t gamma*alpha[t])      (fp64)
 * with n_d from the document's observed tokens of view m at their current assignments and gamma*alpha = 0 on inactive topics.
 * *ll_out = sum of the logs, *n_out = tokens scored; perplexity = exp(-ll / n). */
int mvtm_heldout_loglik(mvtm_handle *h, int32_t m, const int64_t *eval_off, const int32_t *eval_word, double *ll_out, int64_t *n_out);

/* Parity probe: the conditional distribution of token (m, doc, pos) on the current (frozen) counts, computed
 * by the same device code the sweep uses.  p_row = row m of the view-coupling matrix p (W:327-337), M entries
 * (NULL = identity).  probs_out[0..K) normalised, probs_out[K] = share of the new-topic bucket (W:515). */
int mvtm_cond_probs(mvtm_handle *h, int32_t m, int64_t doc, int32_t pos, const double *p_row, double *probs_out);

/* The same probe, extended.  tree_mode: 0 = the trainer's trees (gamma*alpha*phi leaves, inactive topics masked, M:2660-2691);
 * 2 = the inferencer's trees (bare phi leaves, empty inactive set, I:561-576 / I:243, quirk Q13) -- what mvtm_sweep's
 * update_global = 2 samples from.  not_in_S (n_not_in_S >= 0; -1 = none): the reference's dense index under quirk Q1 -- topics the
 * document holds that the index lacks at this token (gained earlier in the sweep); their document and other-view terms are dropped
 * exactly as W:501-513 does when it walks S.  The token's own removal (W:434-471) is applied on top.  Works on any handle. */
int mvtm_cond_probs_ex(mvtm_handle *h, int32_t m, int64_t doc, int32_t pos, const double *p_row, int32_t tree_mode,
                       const int32_t *not_in_S, int32_t n_not_in_S, double *probs_out);

/* n_wk == histogram of (word, z), n_k == histogram of z, no negative cell.
 * *violations_out = number of offending cells (0 = consistent). */
int mvtm_check_invariants(mvtm_handle *h, int64_t *violations_out);

int mvtm_stats(mvtm_handle *h, mvtm_sweep_stats *out);

/* Multi-GPU plumbing (one handle per rank, documents sharded, n_wk / n_k replicated -- SURVEY 8e).
 * mvtm_delta_begin snapshots the replicas; after local sweeps mvtm_delta_export turns view m's replica into
 * its local delta IN PLACE and returns the device pointers (n_wk: V_m x row_stride int32, n_k: row_stride int32)
 * for the caller's all-reduce (NCCL through torch.distributed); mvtm_delta_import adds the reduced delta back
 * onto the snapshot, giving every rank the same bit-exact global counts. */
int mvtm_delta_begin(mvtm_handle *h);
int mvtm_delta_reset(mvtm_handle *h);      /* snapshot := 0 (the replicas hold purely local counts, e.g. right after a rebuild) */
int mvtm_delta_export(mvtm_handle *h, int32_t m, void **n_wk_dev, int64_t *n_wk_elems, void **n_k_dev, int64_t *n_k_elems);
int mvtm_delta_import(mvtm_handle *h, int32_t m);
/* Sum-form of the same exchange, one pass cheaper: when every rank entered the sweep with the same global counts G (true
 * after any completed exchange), all-reduce the replicas themselves (buffers from mvtm_sum_exchange_buffers) and call
 * mvtm_sum_exchange_finish, which turns N*G + sum(delta) into G + sum(delta) using the snapshot of G.  Needs
 * world_size * (largest cell) < 2^31. */
int mvtm_sum_exchange_buffers(mvtm_handle *h, int32_t m, void **n_wk_dev, int64_t *n_wk_elems, void **n_k_dev, int64_t *n_k_elems);
int mvtm_sum_exchange_finish(mvtm_handle *h, int32_t m, int32_t world_size);
int mvtm_row_stride(mvtm_handle *h, int32_t *stride_out);

/* Overlapped exchange.  The reference's barrier M:1231 only requires that view m's counts are complete before view m is
 * sampled AGAIN, so the exchange of view m may run while the following views (and the next sweep's earlier views) are being
 * sampled.  The sweep is queued one view pass at a time without host synchronisation:
 *     for m: mvtm_sweep_view_async(h, it, m, 1)            -- waits (on the device) for anything handed over for view m
 *            mvtm_stream_wait_view(h, m, comm)              -- the caller's stream `comm` (a cudaStream_t) waits for that pass
 *            <all-reduce of mvtm_sum_exchange_buffers(m) on comm; n_k follows n_wk in the same allocation, so ONE
 *             all-reduce of n_wk_elems + n_k_elems ints starting at n_wk_dev covers both>
 *            mvtm_sum_exchange_finish_async(h, m, N, comm, ctas)   -- the finishing pass, queued on comm
 *            mvtm_view_wait_stream(h, m, comm)              -- hand view m back: its next pass / any reader waits for comm
 *     mvtm_sweep_finish(h)                                  -- host side of the barrier: waits for the queued PASSES only
 *                                                              (not for comm), fills mvtm_stats
 * The caller leaves SMs free for the collective by creating the handle with mvtm_config.max_ctas < the SM count
 * (the sweep kernel is persistent and otherwise owns every SM's shared memory until it ends). */
int mvtm_sweep_view_async(mvtm_handle *h, int32_t iteration, int32_t m, int32_t update_global);
int mvtm_sweep_finish(mvtm_handle *h);
int mvtm_stream_wait_view(mvtm_handle *h, int32_t m, void *stream);
int mvtm_view_wait_stream(mvtm_handle *h, int32_t m, void *stream);
int mvtm_sum_exchange_finish_async(mvtm_handle *h, int32_t m, int32_t world_size, void *stream, int32_t max_ctas);

/* ---- Multi-GPU inside the library ("NCCL communicators created" by the boundary; 8e) --------------------------------
 * One handle per GPU / rank, documents sharded over the ranks (mvtm_config.doc_id_base / doc_id_stride name the global ids),
 * n_wk / n_k replicated.  The host only moves 128 bytes: rank 0 calls mvtm_comm_unique_id and hands the id to every rank (over
 * whatever it has: a socket, MPI, a Java RMI call), then every rank calls mvtm_comm_init.  NCCL itself is loaded at run time
 * (libnccl.so.2), so single-GPU hosts do not need it.  Replaces, across GPUs, what the reference's barrier M:1231 and its
 * updater threads (U:164-297) do across CPU threads.
 *   mvtm_comm_init(h, id, rank, world, hidden_ctas)  communicator over `world` ranks; hidden_ctas > 0 additionally creates a
 *       CTA-limited communicator for exchanges that run UNDER another view's pass -- create the handle with
 *       mvtm_config.max_ctas = SMs - hidden_ctas so that those SMs stay free (the sweep kernel is persistent).
 *   mvtm_sync_counts(h, rebuild)   local counts -> global counts (one in-place all-reduce per view) + snapshot; call after
 *       mvtm_init_assignments / mvtm_set_assignments on every rank (rebuild != 0 first recounts from the resident assignments).
 *   mvtm_sweep_dist(h, it)         one Gibbs sweep over this rank's shard + the count exchange: per view, the pass, then an all-reduce
 *       of the replica and a fused finishing pass on the library's own stream -- view m's exchange runs while the following
 *       views (and the next sweep's earlier views) are sampled; only the view with the longest pass is exchanged on the wide
 *       communicator.  Returns when the PASSES are done (mvtm_stats valid); the exchanges are ordered on the device before
 *       anything that touches the view again.  Inactive topics sampled by a sweep are activated (U:263-270) at the start of
 *       the next one, on the global counts.  mvtm_comm_drain waits for everything.
 *   mvtm_sweep_host_dist(h, it, z) the step of a host whose only state are its assignment arrays (the reference's topicSequence),
 *       on every rank: assignments in, assignments out.  The first call -- and any call after something else wrote the handle's
 *       assignments or tables, or after the caller changed its arrays -- uploads, recounts locally, makes the counts global with
 *       ONE all-reduce per view (overlapped with the next view's upload) and takes the exchange snapshot.  Otherwise the upload
 *       is only COMPARED with the resident assignments (one 4-byte all-reduce tells every rank whether all shards are intact)
 *       and the resident global counts are used as they are; a difference on any rank sends all ranks down the recount path, so
 *       the result never depends on which path ran.  Every rank takes part in that 4-byte verdict on EVERY call (a rank that cannot
 *       compare reports "changed"), so the ranks always agree on the collectives that follow; topic ids >= K in the arrays are treated
 *       as unassigned, the step still completes on every rank and the call then returns MVTM_ERR_ARG.  Then the passes with their overlapped exchanges as in mvtm_sweep_dist; new z
 *       is written to the caller's arrays (by the sweep kernel itself when they are pinned + mapped).  The replicas hold global
 *       counts afterwards (mvtm_sweep_dist / mvtm_loglik_dist may follow).  mvtm_comm_last_host_step tells which path ran;
 *       MVTM_HOST_COMPARE=0 in the environment forces the recount path.
 *   mvtm_loglik_dist(h, ll, q)     modelLogLikelihood of the whole corpus (document parts summed over ranks).
 * With a communicator and no mvtm_set_stat_reducer callback, mvtm_optimize_hyper reduces its statistics over the communicator
 * itself, so every rank installs identical hyper-parameters. */
#define MVTM_COMM_ID_BYTES 128
int mvtm_comm_unique_id(void *id_out);
int mvtm_comm_init(mvtm_handle *h, const void *unique_id, int32_t rank, int32_t world, int32_t hidden_ctas);
int mvtm_comm_destroy(mvtm_handle *h);
int mvtm_comm_info(mvtm_handle *h, int32_t *rank, int32_t *world, int32_t *nccl_version, int64_t *bytes_last_sweep);
int mvtm_sync_counts(mvtm_handle *h, int32_t rebuild_from_assignments);
int mvtm_sweep_dist(mvtm_handle *h, int32_t iteration);
int mvtm_comm_drain(mvtm_handle *h);
int mvtm_sweep_host_dist(mvtm_handle *h, int32_t iteration, int32_t *const *z_inout);
int mvtm_comm_last_host_step(mvtm_handle *h, int32_t *resident_counts_used);
int mvtm_loglik_dist(mvtm_handle *h, double *ll_out, int32_t quirk_len2);

/* Scan layout of the sampler (for order-exact checkers): a document-view is sampled by `lanes_per_doc` lanes (8, 16 or
 * 32); topic t sits in 4-topic chunk c = t/4 owned by lane c % lanes_per_doc as its (c / lanes_per_doc)-th chunk, and the
 * cumulative scan runs lane-major (all chunks of lane 0, then lane 1, ...). */
int mvtm_scan_layout(mvtm_handle *h, int32_t *lanes_per_doc, int32_t *chunks_per_lane);

/* ---- hyper-parameter step of estimate() (M:1173-1210), SURVEY 8(f) rank 1 -------------------------------------------
 * mvtm_optimize_hyper runs, in the reference's order, the selected parts of
 *   optimizeP (M:2698-2819), optimizeDP (M:2440-2591), optimizeGamma (M:2369-2438), optimizeBeta (M:2288-2367)
 * on the host from statistics computed on the device, and installs the new alpha / alphaSum / beta / betaSum / gamma /
 * p_a / p_b / inactive-topic set in the handle (the F+trees the reference rebuilds at M:1209 are implicit here).
 * Draws come from the handle's Philox stream keyed on `iteration`.  mvtm_p_statistics exposes optimizeP's sufficient
 * statistic: psum[m*M+i] = sum over documents of pDistr_Mean[m][i][doc] (M:2706-2782), docs_per_view = totalDocsPerModality. */
#define MVTM_OPT_P     1u
#define MVTM_OPT_DP    2u
#define MVTM_OPT_GAMMA 4u
#define MVTM_OPT_BETA  8u
#define MVTM_OPT_ALL   15u
int mvtm_optimize_hyper(mvtm_handle *h, int32_t iteration, uint32_t which);
/* Multi-rank hyper-parameter step ("at optimise steps, the doc-topic histogram").  The statistics that
 * mvtm_optimize_hyper gathers from THIS handle's documents -- optimizeP's overlap sums and per-view document counts, the longest
 * document, the doc-topic histograms of optimizeDP, docLengthCounts of optimizeGamma -- are passed to `fn` before use:
 * op 0 = replace each element by its sum over all ranks, op 1 = by its maximum; `ints` / `reals` may be NULL with a zero count;
 * return 0 on success.  With an all-reduce behind `fn`, the same seed and the same iteration every rank installs identical
 * hyper-parameters.  optimizeBeta needs no reduction: it reads the (already global) count tables.  fn = NULL (default): single handle. */
/* Activation of inactive topics that received tokens (U:263-270: alpha[m][t] = alpha[m][K], topic leaves inActiveTopicIndex).
 * A single handle does this after every view pass of mvtm_sweep (the reference: per delta).  With a statistics reducer installed (multi-rank) the sweep leaves it to
 * the caller, who calls mvtm_activate_topics AFTER the count exchange so that every rank decides on the same global counts. */
int mvtm_activate_topics(mvtm_handle *h);
typedef int (*mvtm_stat_reducer)(void *ctx, int32_t op, int64_t *ints, int64_t n_ints, double *reals, int64_t n_reals);
int mvtm_set_stat_reducer(mvtm_handle *h, mvtm_stat_reducer fn, void *ctx);
int mvtm_p_statistics(mvtm_handle *h, double *psum_out, int64_t *docs_per_view_out);
int mvtm_get_hyper_full(mvtm_handle *h, double *alpha, double *alpha_sum, double *beta, double *beta_sum, double *gamma,
                        double *p_a, double *p_b, double *p_mean, double *gamma_root, double *gamma_view, double *tables_cnt);

/* Test hooks for the host-side samplers of the hyper-parameter step (no device work): `which` 0 = uniform, 1 = Gamma(a,1),
 * 2 = Beta(a,b), 3 = Antoniak(alpha = a, n = b), 4 = the sweep's MVTM_FLAG_BETA_MALLET draw (MALLET's Randoms.nextBeta law); and MALLET's learnSymmetricConcentration as restated in this build. */
int mvtm_test_sampler(uint64_t seed, int32_t which, double a, double b, int32_t n, double *out);
double mvtm_test_learn_symmetric_concentration(const int64_t *count_hist, int32_t n_count, const int64_t *length_hist,
                                               int32_t n_length, int32_t num_dimensions, double current);
/* optimizeDP (which & MVTM_OPT_DP, M:2440-2591) then optimizeGamma (which & MVTM_OPT_GAMMA, M:2369-2438) on plain host arrays with
 * SCRIPTED draws -- the same code mvtm_optimize_hyper runs, without a device, so that the arguments of every sampler call and the
 * resulting alpha / gamma can be compared with the reference's bytecode fed the same draws (tests/test_optim_host.py).
 * hist[m]: K x stride[m] bins of topicDocCounts; lencnt[m]: n_len[m] bins of docLengthCounts; alpha: M x (K+1);
 * scal = { gammaRoot, rootTablesCnt } (in/out); script: one value per draw in call order (Gamma(a,1), Beta, Bernoulli, Antoniak);
 * arg_log: 3 doubles (kind 1..4, a, b) per consumed draw; *n_used: draws consumed.  MVTM_ERR_ARG when the script is too short. */
/* The launch shape a handle with these parameters would give a view pass, computed without a device (148 SMs assumed): lanes per
 * document, ring depth (0 = DIRECT kernel), warps per CTA, dynamic shared memory per CTA, and the register count the kernel is
 * compiled for (0 = bounded by its launch shape).  A CPU test checks every K against the SM's limits (227 KB of shared memory,
 * 16 K registers per sub-partition) together with the register counts of the compiled kernels. */
int mvtm_test_launch_shape(int32_t num_topics, int32_t num_views, uint32_t flags, int32_t *lanes_per_doc, int32_t *ring_depth,
                           int32_t *warps_per_cta, int64_t *smem_bytes, int32_t *maxnreg);
int mvtm_test_hyper_core(int32_t M, int32_t K, uint32_t which, const int64_t *const *hist, const int32_t *stride,
                         const int64_t *const *lencnt, const int32_t *n_len, double *alpha, double *alpha_sum, double *gamma,
                         double *gamma_view, double *tables_cnt, double *scal, int32_t *inactive, int32_t *n_inactive,
                         const double *script, int64_t script_len, double *arg_log, int64_t *n_used);

/* Build information: "sm_100a", kernel variants compiled in. */
const char *mvtm_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* MVTM_H */
