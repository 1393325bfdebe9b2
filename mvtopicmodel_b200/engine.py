"""Thin numpy-facing wrapper over the C ABI (one Engine = one mvtm_handle = one GPU)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import MvtmConfig, MvtmSweepStats


class MvtmError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"mvtm status {status}: {message}")
        self.status = status


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    def __init__(self, K, V, views, seed=1, device=0, flags=0, doc_id_base=0, doc_id_stride=1, present=None,
                 warps_per_cta=0, ring_depth=0, max_ctas=0):
        """views: list of (doc_off int64[D+1], word_id int32[N]) -- a doc-aligned CSR per view (MA:13-19)."""
        self.L = _lib.lib()
        self.K, self.M = int(K), len(views)
        self.V = np.ascontiguousarray(V, dtype=np.int32)
        assert len(self.V) == self.M
        self.D = len(views[0][0]) - 1
        cfg = MvtmConfig(self.K, self.M, self.D, self.V.ctypes.data_as(C.POINTER(C.c_int32)), int(seed), int(device),
                         int(flags), int(doc_id_base), int(doc_id_stride), int(warps_per_cta), int(ring_depth), int(max_ctas))
        h = C.c_void_p()
        rc = self.L.mvtm_create(C.byref(cfg), C.byref(h))
        if rc:
            raise MvtmError(rc, self.L.mvtm_last_error(None).decode())
        self.h = h
        self.ntok = []
        for m, (off, word) in enumerate(views):
            off = np.ascontiguousarray(off, dtype=np.int64)
            word = np.ascontiguousarray(word, dtype=np.int32)
            if len(off) != self.D + 1:
                raise ValueError("every view needs num_docs+1 offsets")
            pr = None if present is None or present[m] is None else np.ascontiguousarray(present[m], dtype=np.uint8)
            self._ck(self.L.mvtm_add_view(self.h, m, _ptr(off), _ptr(word), _ptr(pr)))
            self.ntok.append(int(off[-1]))

    def _ck(self, rc):
        if rc:
            raise MvtmError(rc, self.L.mvtm_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.mvtm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- state ---------------------------------------------------------------------------------
    def init_assignments(self):
        self._ck(self.L.mvtm_init_assignments(self.h))

    def set_assignments(self, m, z):
        z = np.ascontiguousarray(z, dtype=np.int32)
        if len(z) != self.ntok[m]:
            raise ValueError("assignment array length != tokens of the view")
        self._ck(self.L.mvtm_set_assignments(self.h, m, _ptr(z)))

    def set_counts(self, m, nwk, nk):
        nwk = np.ascontiguousarray(nwk, dtype=np.int32); nk = np.ascontiguousarray(nk, dtype=np.int32)
        if nwk.shape != (int(self.V[m]), self.K) or nk.shape != (self.K,):
            raise ValueError("n_wk must be V_m x K and n_k K")
        self._ck(self.L.mvtm_set_counts(self.h, m, _ptr(nwk), _ptr(nk)))

    def init_assignments_from_counts(self):
        self._ck(self.L.mvtm_init_assignments_from_counts(self.h))

    def get_assignments(self, m, out=None):
        z = np.empty(self.ntok[m], dtype=np.int32) if out is None else out
        self._ck(self.L.mvtm_get_assignments(self.h, m, _ptr(z)))
        return z

    def get_counts(self, m, want_nwk=True):
        nwk = np.empty((int(self.V[m]), self.K), dtype=np.int32) if want_nwk else None
        nk = np.empty(self.K, dtype=np.int32)
        self._ck(self.L.mvtm_get_counts(self.h, m, _ptr(nwk), _ptr(nk)))
        return nwk, nk

    def set_hyper(self, alpha=None, alphaSum=None, beta=None, betaSum=None, gamma=None, p_a=None, p_b=None,
                  inactive=None):
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        alpha, alphaSum, beta, betaSum, gamma, p_a, p_b = map(f, (alpha, alphaSum, beta, betaSum, gamma, p_a, p_b))
        if alpha is not None and alpha.size != self.M * (self.K + 1):
            raise ValueError("alpha must be M x (K+1)")
        if inactive is None:
            ina, n = None, -1
        else:
            ina = np.ascontiguousarray(sorted(inactive), dtype=np.int32)
            n = len(ina)
        self._ck(self.L.mvtm_set_hyper(self.h, _ptr(alpha), _ptr(alphaSum), _ptr(beta), _ptr(betaSum), _ptr(gamma),
                                       _ptr(p_a), _ptr(p_b), _ptr(ina), n))

    def get_hyper(self):
        alpha = np.empty((self.M, self.K + 1), dtype=np.float64)
        asum = np.empty(self.M, dtype=np.float64)
        ina = np.empty(self.K, dtype=np.int32)
        n = C.c_int32()
        self._ck(self.L.mvtm_get_hyper(self.h, _ptr(alpha), _ptr(asum), _ptr(ina), C.byref(n)))
        return alpha, asum, ina[:n.value].copy()

    # --- hyper-parameter step (M:1173-1210) --------------------------------------------------------
    def optimize_hyper(self, iteration, which=15):
        self._ck(self.L.mvtm_optimize_hyper(self.h, int(iteration), int(which)))

    def activate_topics(self):
        self._ck(self.L.mvtm_activate_topics(self.h))

    def set_stat_reducer(self, fn):
        """fn(op, ints, reals): op 0 = sum, 1 = max over the ranks, IN PLACE on the two numpy views (either may be None);
        None removes the reducer.  Used by the multi-rank hyper-parameter step (mvtm_set_stat_reducer)."""
        if fn is None:
            self._reducer = None
            self._ck(self.L.mvtm_set_stat_reducer(self.h, C.cast(None, _lib.STAT_REDUCER), None))
            return

        def trampoline(ctx, op, ints, n_ints, reals, n_reals):
            try:
                a = np.ctypeslib.as_array(ints, shape=(n_ints,)) if n_ints else None
                b = np.ctypeslib.as_array(reals, shape=(n_reals,)) if n_reals else None
                fn(int(op), a, b)
                return 0
            except Exception:          # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return 1
        self._reducer = _lib.STAT_REDUCER(trampoline)      # keep the callback object alive
        self._ck(self.L.mvtm_set_stat_reducer(self.h, self._reducer, None))

    def p_statistics(self):
        psum = np.empty((self.M, self.M), dtype=np.float64)
        docs = np.empty(self.M, dtype=np.int64)
        self._ck(self.L.mvtm_p_statistics(self.h, _ptr(psum), _ptr(docs)))
        return psum, docs

    def get_hyper_full(self):
        M, K = self.M, self.K
        d = dict(alpha=np.empty((M, K + 1)), alphaSum=np.empty(M), beta=np.empty(M), betaSum=np.empty(M), gamma=np.empty(M),
                 p_a=np.empty((M, M)), p_b=np.empty((M, M)), pMean=np.empty((M, M)), gammaRoot=np.empty(1), gammaView=np.empty(M),
                 tablesCnt=np.empty(M))
        self._ck(self.L.mvtm_get_hyper_full(self.h, *[_ptr(d[k]) for k in ("alpha", "alphaSum", "beta", "betaSum", "gamma", "p_a", "p_b",
                                                                              "pMean", "gammaRoot", "gammaView", "tablesCnt")]))
        d["gammaRoot"] = float(d["gammaRoot"][0])
        d["inactive"] = self.get_hyper()[2]
        return d

    # --- the hot path ----------------------------------------------------------------------------
    def sweep(self, iteration, update_global=True):
        mode = update_global if update_global in (0, 1, 2) else int(bool(update_global))
        self._ck(self.L.mvtm_sweep(self.h, int(iteration), int(mode)))

    def sweep_host(self, iteration, z_arrays):
        """z_arrays: one int32 numpy array per view, updated in place (host buffers in, host buffers out)."""
        ptrs = (C.c_void_p * self.M)()
        for m, z in enumerate(z_arrays):
            assert z.dtype == np.int32 and z.flags["C_CONTIGUOUS"] and len(z) == self.ntok[m]
            ptrs[m] = z.ctypes.data
        self._ck(self.L.mvtm_sweep_host(self.h, int(iteration), ptrs))

    # --- multi-GPU inside the library (NCCL behind the C ABI, include/mvtm.h "Multi-GPU inside the library") ------------------
    @staticmethod
    def comm_unique_id():
        """128 bytes from ncclGetUniqueId; rank 0 creates it, the host hands it to every rank."""
        buf = C.create_string_buffer(128)
        L = _lib.lib()
        rc = L.mvtm_comm_unique_id(buf)
        if rc:
            raise MvtmError(rc, L.mvtm_last_error(None).decode())
        return buf.raw

    def comm_init(self, unique_id, rank, world, hidden_ctas=0):
        assert len(unique_id) == 128
        self._ck(self.L.mvtm_comm_init(self.h, C.c_char_p(unique_id), int(rank), int(world), int(hidden_ctas)))

    def comm_info(self):
        r, w, v, b = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        self._ck(self.L.mvtm_comm_info(self.h, C.byref(r), C.byref(w), C.byref(v), C.byref(b)))
        return {"rank": r.value, "world": w.value, "nccl_version": v.value, "bytes_last_sweep": b.value}

    def sync_counts(self, rebuild=False):
        self._ck(self.L.mvtm_sync_counts(self.h, int(bool(rebuild))))

    def sweep_dist(self, iteration):
        self._ck(self.L.mvtm_sweep_dist(self.h, int(iteration)))

    def comm_drain(self):
        self._ck(self.L.mvtm_comm_drain(self.h))

    def sweep_host_dist(self, iteration, z_arrays):
        ptrs = (C.c_void_p * self.M)()
        for m, z in enumerate(z_arrays):
            assert z.dtype == np.int32 and z.flags["C_CONTIGUOUS"] and len(z) == self.ntok[m]
            ptrs[m] = z.ctypes.data
        self._ck(self.L.mvtm_sweep_host_dist(self.h, int(iteration), ptrs))

    def last_host_step_used_resident_counts(self):
        """True if the last sweep_host_dist found the caller's arrays equal to the resident assignments on every rank and skipped the recount."""
        f = C.c_int32()
        self._ck(self.L.mvtm_comm_last_host_step(self.h, C.byref(f)))
        return bool(f.value)

    def loglik_dist(self, quirk_len2=False):
        out = np.empty(self.M, dtype=np.float64)
        self._ck(self.L.mvtm_loglik_dist(self.h, _ptr(out), int(bool(quirk_len2))))
        return out

    def set_host_mirror(self, m, z_host):
        """z_host: pinned int32 numpy array (e.g. torch.empty(n, dtype=torch.int32).pin_memory().numpy()) or None."""
        if z_host is not None:
            assert z_host.dtype == np.int32 and z_host.flags["C_CONTIGUOUS"] and len(z_host) == self.ntok[m]
        self._ck(self.L.mvtm_set_host_mirror(self.h, int(m), None if z_host is None else C.c_void_p(z_host.ctypes.data)))

    def stats(self):
        s = MvtmSweepStats()
        self._ck(self.L.mvtm_stats(self.h, C.byref(s)))
        return {"tokens": s.tokens, "changed": s.changed, "new_topic": s.new_topic, "ms_total": s.ms_total,
                "ms_view": [s.ms_view[m] for m in range(self.M)], "kernel_launches": s.kernel_launches,
                "ring_depth": [s.ring_depth[m] for m in range(self.M)], "ring_locked": [s.ring_locked[m] for m in range(self.M)]}

    # --- readers -----------------------------------------------------------------------------------
    def cond_probs(self, m, doc, pos, p_row=None, not_in_S=None, tree_mode=0):
        """not_in_S (optional): held topics the reference's dense index lacks at this token (quirk Q1); tree_mode 2 = the
        inferencer's bare-phi trees (mvtm_cond_probs_ex)."""
        out = np.empty(self.K + 1, dtype=np.float64)
        pr = None if p_row is None else np.ascontiguousarray(p_row, dtype=np.float64)
        if not_in_S is None and tree_mode == 0:
            self._ck(self.L.mvtm_cond_probs(self.h, int(m), int(doc), int(pos), _ptr(pr), _ptr(out)))
        else:
            ex = None if not_in_S is None else np.ascontiguousarray(list(not_in_S), dtype=np.int32)
            self._ck(self.L.mvtm_cond_probs_ex(self.h, int(m), int(doc), int(pos), _ptr(pr), int(tree_mode), _ptr(ex),
                                               -1 if ex is None else len(ex), _ptr(out)))
        return out

    def loglik(self, quirk_len2=False):
        out = np.empty(self.M, dtype=np.float64)
        self._ck(self.L.mvtm_loglik(self.h, _ptr(out), int(quirk_len2)))
        return out

    def heldout_loglik(self, m, eval_off, eval_word):
        """(sum of log p(w | d) over the evaluation tokens of view m, tokens scored) -- document completion, see mvtm.h."""
        eo = np.ascontiguousarray(eval_off, dtype=np.int64); ew = np.ascontiguousarray(eval_word, dtype=np.int32)
        if len(eo) != self.D + 1 or eo[-1] != len(ew):
            raise ValueError("eval CSR must be aligned with the handle's documents")
        ll, n = C.c_double(), C.c_int64()
        self._ck(self.L.mvtm_heldout_loglik(self.h, int(m), _ptr(eo), _ptr(ew), C.byref(ll), C.byref(n)))
        return ll.value, n.value

    def loglik_parts(self, quirk_len2=False):
        """(document part over this handle's documents, topic-word part of the count tables), each M doubles."""
        doc, word = np.empty(self.M, dtype=np.float64), np.empty(self.M, dtype=np.float64)
        self._ck(self.L.mvtm_loglik_parts(self.h, _ptr(doc), _ptr(word), int(quirk_len2)))
        return doc, word

    def doc_topic_hist(self, m):
        ml = C.c_int32()
        self._ck(self.L.mvtm_doc_topic_hist(self.h, m, None, C.byref(ml)))
        hist = np.empty((self.K, ml.value + 1), dtype=np.int32)
        self._ck(self.L.mvtm_doc_topic_hist(self.h, m, _ptr(hist), C.byref(ml)))
        return hist

    def check_invariants(self):
        v = C.c_int64()
        self._ck(self.L.mvtm_check_invariants(self.h, C.byref(v)))
        return v.value

    # --- multi-GPU plumbing ------------------------------------------------------------------------
    def row_stride(self):
        s = C.c_int32()
        self._ck(self.L.mvtm_row_stride(self.h, C.byref(s)))
        return s.value

    def scan_layout(self):
        g, j = C.c_int32(), C.c_int32()
        self._ck(self.L.mvtm_scan_layout(self.h, C.byref(g), C.byref(j)))
        return g.value, j.value

    def delta_begin(self):
        self._ck(self.L.mvtm_delta_begin(self.h))

    def delta_reset(self):
        self._ck(self.L.mvtm_delta_reset(self.h))

    def delta_export(self, m):
        p1, n1, p2, n2 = C.c_void_p(), C.c_int64(), C.c_void_p(), C.c_int64()
        self._ck(self.L.mvtm_delta_export(self.h, m, C.byref(p1), C.byref(n1), C.byref(p2), C.byref(n2)))
        return (p1.value, n1.value), (p2.value, n2.value)

    def sum_exchange_buffers(self, m):
        p1, n1, p2, n2 = C.c_void_p(), C.c_int64(), C.c_void_p(), C.c_int64()
        self._ck(self.L.mvtm_sum_exchange_buffers(self.h, m, C.byref(p1), C.byref(n1), C.byref(p2), C.byref(n2)))
        return (p1.value, n1.value), (p2.value, n2.value)

    def sum_exchange_finish(self, m, world):
        self._ck(self.L.mvtm_sum_exchange_finish(self.h, m, int(world)))

    # overlapped exchange (mvtm.h "Overlapped exchange"): streams are raw cudaStream_t handles (ints)
    def sweep_view_async(self, iteration, m, update_global=1):
        self._ck(self.L.mvtm_sweep_view_async(self.h, int(iteration), int(m), int(update_global)))

    def sweep_finish(self):
        self._ck(self.L.mvtm_sweep_finish(self.h))

    def stream_wait_view(self, m, stream):
        self._ck(self.L.mvtm_stream_wait_view(self.h, int(m), C.c_void_p(int(stream))))

    def view_wait_stream(self, m, stream):
        self._ck(self.L.mvtm_view_wait_stream(self.h, int(m), C.c_void_p(int(stream))))

    def sum_exchange_finish_async(self, m, world, stream, max_ctas=0):
        self._ck(self.L.mvtm_sum_exchange_finish_async(self.h, int(m), int(world), C.c_void_p(int(stream)), int(max_ctas)))

    def delta_import(self, m):
        self._ck(self.L.mvtm_delta_import(self.h, m))
