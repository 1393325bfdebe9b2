"""ctypes binding of the CPU oracle (oracle/mvtm_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (mvtopicmodel_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmvtm_oracle.so")

F_Q1_COMPAT, F_STALE_TREES, F_DEFERRED, F_BETA_MALLET, F_ENGINE_MIRROR, F_DOC_ORDER, F_FROZEN, F_BARE_TREES = 1, 2, 4, 8, 16, 32, 64, 128
F_CHECK_RULE = 256


_STAMP = _SO + ".cpuflags"


_ISA_PREFIXES = ("sse", "ssse", "avx", "fma", "bmi", "adx", "aes", "pclmul", "popcnt", "lzcnt", "abm", "movbe", "f16c", "sha", "vaes",
                 "vpclmul", "gfni", "amx", "rdrnd", "rdseed", "clflushopt", "clwb", "movdir", "serialize", "waitpkg", "xsave", "fsgsbase")


def _cpu_flags():
    """the instruction-set flags of /proc/cpuinfo (only those -march=native can turn into instructions)"""
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("flags"):
                    return {w for w in ln.split(":", 1)[1].split() if w.startswith(_ISA_PREFIXES)}
    except OSError:
        pass
    return set()


def _built_for_this_cpu():
    """The Makefile compiles with -march=native (BASELINE.md's CPU arm), and the built file travels with the repo snapshot: if the
    host it lands on lacks an ISA extension of the host it was built on, rebuild there instead of dying on an illegal instruction."""
    if not os.path.exists(_STAMP):
        return True
    with open(_STAMP) as f:
        return set(f.read().split()) <= _cpu_flags()


def build(force=False):
    src = os.path.join(_HERE, "mvtm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src) or not _built_for_this_cpu():
        try:
            subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
        except (OSError, subprocess.CalledProcessError):
            if force or not os.path.exists(_SO):
                raise
            return _SO                       # no compiler on this host: keep the file that travelled here
        with open(_STAMP, "w") as f:
            f.write(" ".join(sorted(_cpu_flags())))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or not _built_for_this_cpu():
            build()
        L = C.CDLL(_SO)
        p, i32, i64, u64, dbl, u32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double, C.c_uint
        sig = {
            "orc_philox4x32": (None, [p, p, p]),
            "orc_ftree_build": (None, [p, p, i32]),
            "orc_ftree_sample": (i32, [p, i32, dbl]),
            "orc_ftree_update": (None, [p, i32, i32, dbl]),
            "orc_lower_bound": (i32, [p, dbl, i32]),
            "orc_log_gamma_stirling": (dbl, [dbl]),
            "orc_next_beta_mallet": (dbl, [u64, dbl, dbl]),
            "orc_create": (p, [i32, i32, i64, p, u64]),
            "orc_destroy": (None, [p]),
            "orc_add_view": (i32, [p, i32, p, p, p]),
            "orc_set_hyper": (i32, [p, p, p, p, p, p, p, p, p, i32]),
            "orc_rebuild_trees": (i32, [p]),
            "orc_rebuild_counts": (i32, [p]),
            "orc_init_assignments": (i32, [p]),
            "orc_init_from_phi": (i32, [p]),
            "orc_set_assignments": (i32, [p, i32, p]),
            "orc_get_assignments": (i32, [p, i32, p]),
            "orc_get_counts": (i32, [p, i32, p, p]),
            "orc_maxlen": (i32, [p, i32]),
            "orc_set_doc_ids": (None, [p, i64, i64]),
            "orc_set_engine_group": (None, [p, i32]),
            "orc_set_counts": (i32, [p, i32, p, p]),
            "orc_get_hist": (i32, [p, i32, p]),
            "orc_get_alpha": (i32, [p, i32, p]),
            "orc_get_inactive": (i32, [p, p]),
            "orc_get_bucket_counters": (None, [p, p]),
            "orc_sweep": (i32, [p, i32, u32]),
            "orc_sweep_mt": (i32, [p, i32, i32, u32]),
            "orc_cond_probs": (i32, [p, i32, i64, i32, p, i32, p]),
            "orc_rule_violations": (C.c_longlong, [p]),
            "orc_cond_probs_q1": (i32, [p, i32, i64, i32, p, i32, p, p]),
            "orc_cond_probs_ex": (i32, [p, i32, i64, i32, p, i32, p, u32, p]),
            "orc_engine_select": (i32, [p, i32, dbl, dbl, i32]),
            "orc_loglik": (i32, [p, p, i32]),
            "orc_check_invariants": (i64, [p]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32(_ptr(c), _ptr(k), _ptr(out))
    return out


def ftree_build(w):
    w = np.ascontiguousarray(w, dtype=np.float64)
    tree = np.zeros(2 * len(w), dtype=np.float64)
    lib().orc_ftree_build(_ptr(tree), _ptr(w), len(w))
    return tree


def ftree_sample(tree, u):
    return lib().orc_ftree_sample(_ptr(tree), len(tree) // 2, float(u))


def ftree_update(tree, t, v):
    lib().orc_ftree_update(_ptr(tree), len(tree) // 2, int(t), float(v))


def lower_bound(arr, key, n=None):
    a = np.ascontiguousarray(arr, dtype=np.float64)
    return lib().orc_lower_bound(_ptr(a), float(key), len(a) if n is None else n)


def log_gamma_stirling(z):
    return lib().orc_log_gamma_stirling(float(z))


def next_beta_mallet(seed, a, b):
    return lib().orc_next_beta_mallet(int(seed), float(a), float(b))


def engine_select(w, u, C_mass=0.0, first_inactive=-1):
    w = np.ascontiguousarray(w, dtype=np.float64)
    return lib().orc_engine_select(_ptr(w), len(w), float(u), float(C_mass), int(first_inactive))


class Oracle:
    """One corpus + model state held by the C oracle.  `views` is a list of (doc_off int64[D+1], word int32[N])."""

    def __init__(self, K, V, views, seed=1, present=None):
        self.M, self.K = len(views), int(K)
        self.V = np.asarray(V, dtype=np.int32)
        self.D = len(views[0][0]) - 1
        self.h = lib().orc_create(self.M, self.K, self.D, _ptr(self.V), int(seed))
        if not self.h:
            raise ValueError("orc_create failed")
        self.ntok = []
        for m, (off, word) in enumerate(views):
            off = np.ascontiguousarray(off, dtype=np.int64)
            word = np.ascontiguousarray(word, dtype=np.int32)
            assert len(off) == self.D + 1
            pr = None if present is None or present[m] is None else np.ascontiguousarray(present[m], dtype=np.uint8)
            rc = lib().orc_add_view(self.h, m, _ptr(off), _ptr(word), _ptr(pr))
            if rc:
                raise ValueError(f"orc_add_view rc={rc}")
            self.ntok.append(int(off[-1]))

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_hyper(self, alpha=None, alphaSum=None, beta=None, betaSum=None, gamma=None, p_a=None, p_b=None,
                  inactive=None):
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        alpha, alphaSum, beta, betaSum, gamma, p_a, p_b = map(f, (alpha, alphaSum, beta, betaSum, gamma, p_a, p_b))
        if inactive is None:
            ina, n = None, -1
        else:
            ina = np.ascontiguousarray(sorted(inactive), dtype=np.int32)
            n = len(ina)
        lib().orc_set_hyper(self.h, _ptr(alpha), _ptr(alphaSum), _ptr(beta), _ptr(betaSum), _ptr(gamma), _ptr(p_a),
                            _ptr(p_b), _ptr(ina), n)

    def init_assignments(self):
        lib().orc_init_assignments(self.h)

    def init_from_phi(self):
        """inferencer initialisation (I:186-203) from the counts currently installed (set_counts)"""
        lib().orc_init_from_phi(self.h)

    def set_assignments(self, zs):
        for m, z in enumerate(zs):
            z = np.ascontiguousarray(z, dtype=np.int32)
            assert len(z) == self.ntok[m]
            lib().orc_set_assignments(self.h, m, _ptr(z))
        lib().orc_rebuild_counts(self.h)

    def rebuild_trees(self):
        if lib().orc_rebuild_trees(self.h):
            raise MemoryError("oracle trees")

    def get_assignments(self, m):
        z = np.zeros(self.ntok[m], dtype=np.int32)
        lib().orc_get_assignments(self.h, m, _ptr(z))
        return z

    def get_counts(self, m):
        nwk = np.zeros((int(self.V[m]), self.K), dtype=np.int32)
        nk = np.zeros(self.K, dtype=np.int32)
        lib().orc_get_counts(self.h, m, _ptr(nwk), _ptr(nk))
        return nwk, nk

    def set_engine_group(self, lanes_per_doc):
        """scan layout of the engine under test (Engine.scan_layout()[0]) for the F_ENGINE_MIRROR mode"""
        lib().orc_set_engine_group(self.h, int(lanes_per_doc))

    def set_doc_ids(self, base, stride):
        lib().orc_set_doc_ids(self.h, int(base), int(stride))

    def set_counts(self, m, nwk=None, nk=None):
        nwk = None if nwk is None else np.ascontiguousarray(nwk, dtype=np.int32)
        nk = None if nk is None else np.ascontiguousarray(nk, dtype=np.int32)
        lib().orc_set_counts(self.h, m, _ptr(nwk), _ptr(nk))

    def get_hist(self, m):
        ml = lib().orc_maxlen(self.h, m)
        h = np.zeros((self.K, ml + 1), dtype=np.int32)
        lib().orc_get_hist(self.h, m, _ptr(h))
        return h

    def get_alpha(self, m):
        a = np.zeros(self.K + 1, dtype=np.float64)
        lib().orc_get_alpha(self.h, m, _ptr(a))
        return a

    def get_inactive(self):
        out = np.zeros(self.K, dtype=np.int32)
        n = lib().orc_get_inactive(self.h, _ptr(out))
        return out[:n].copy()

    def bucket_counters(self):
        out = np.zeros(4, dtype=np.int64)
        lib().orc_get_bucket_counters(self.h, _ptr(out))
        return out

    def sweep(self, iteration, flags=0):
        rc = lib().orc_sweep(self.h, int(iteration), int(flags))
        if rc:
            raise RuntimeError(f"orc_sweep rc={rc}")

    def sweep_mt(self, iteration, threads, flags=0):
        rc = lib().orc_sweep_mt(self.h, int(iteration), int(threads), int(flags))
        if rc:
            raise RuntimeError(f"orc_sweep_mt rc={rc}")

    def cond_probs(self, m, doc, pos, p=None, engine_form=False, not_in_S=None, flags=0):
        """not_in_S: topics the document holds that the reference's dense index lacks at this token (quirk Q1).
        flags: F_BARE_TREES for the inferencer's trees (phi leaves without gamma*alpha, Q13)."""
        out = np.zeros(self.K + 1, dtype=np.float64)
        pm = None if p is None else np.ascontiguousarray(p, dtype=np.float64)
        ex = None
        if not_in_S is not None:
            ex = np.zeros(self.K, dtype=np.uint8); ex[list(not_in_S)] = 1
        rc = lib().orc_cond_probs_ex(self.h, int(m), int(doc), int(pos), _ptr(pm), int(engine_form), _ptr(ex), int(flags), _ptr(out))
        if rc:
            raise ValueError(f"orc_cond_probs rc={rc}")
        return out

    def rule_violations(self):
        return int(lib().orc_rule_violations(self.h))

    def loglik(self, quirk_len2=False):
        out = np.zeros(self.M, dtype=np.float64)
        lib().orc_loglik(self.h, _ptr(out), int(quirk_len2))
        return out

    def check_invariants(self):
        return int(lib().orc_check_invariants(self.h))


def doc_topic_proportions(views, zs, K, gamma, alpha, alphaSum, p_mean0, weights):
    """theta_d[t] of the inferencer's output (I:385-412): sum_m w_m*pMean[0][m]*(n_d[m][t]+gamma_m*alpha_m[t])/(len_m+gamma_m*alphaSum_m)
    over the views the document has, normalised by sum_m w_m*pMean[0][m] (w_0 = 1)."""
    M, D = len(views), len(views[0][0]) - 1
    out = np.zeros((D, K))
    for d in range(D):
        norm = 0.0
        for m in range(M):
            b, e = views[m][0][d], views[m][0][d + 1]
            if e == b:
                continue            # Assignments[m] == null (absent view)
            cnt = np.bincount(zs[m][b:e], minlength=K).astype(np.float64)
            wm = (1.0 if m == 0 else weights[m]) * p_mean0[m]
            out[d] += wm * (cnt + gamma[m] * alpha[m][:K]) / ((e - b) + gamma[m] * alphaSum[m])
            norm += wm
        if norm > 0:
            out[d] /= norm
    return out


def heldout_loglik(obs_view, z_obs, nwk, nk, eval_view, ga, beta, betaSum):
    """Document-completion score of ONE view, the restatement of mvtm_heldout_loglik (include/mvtm.h) in numpy fp64:
    sum over in-vocabulary evaluation tokens of log( phi[w] . (n_d + ga) / (N_obs + sum ga) ), n_d from the observed tokens.
    ga = gamma*alpha[0..K) with zeros on inactive topics.  Returns (sum of logs, tokens scored)."""
    off, _ = obs_view
    eoff, ew = eval_view
    V, K = nwk.shape
    phi = (nwk.astype(np.float64) + beta) / (nk.astype(np.float64) + betaSum)[None, :]
    ga = np.asarray(ga, dtype=np.float64)
    ll, n = 0.0, 0
    for d in range(len(off) - 1):
        words = ew[eoff[d]:eoff[d + 1]]
        words = words[(words >= 0) & (words < V)]
        if len(words) == 0:
            continue
        zd = z_obs[off[d]:off[d + 1]]
        zd = zd[(zd >= 0) & (zd < K)]
        theta = (np.bincount(zd, minlength=K).astype(np.float64) + ga) / (len(zd) + ga.sum())
        ll += float(np.log(phi[words] @ theta).sum())
        n += len(words)
    return ll, n
