"""Python mirror of the drop-in boundary: one object per (data set, tree) whose ``eval`` is one call of the
reference's ``com.plfun`` (lfun / lfundG, src/treesub.c:7764 / 7608) on the GPU.

It is a thin ctypes layer over the C ABI (include/paml_b200.h): the same calls the C host shim
(paml_b200/host/plfun_b200.inc) makes from inside codeml/baseml.  Used by the tests, bench.py and smoke().
"""
import ctypes as C
import numpy as np

from ._lib import lib, check, PB200Error  # noqa: F401
from .fixture import Problem, SEM_LFUN, SEM_LFUNDG  # noqa: F401


def _ptr(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def plan_tree(nnode, ns, root, nson, sons, scale_flags=None, ps=False):
    """Host-side planner only (no GPU): returns (program int32 array, nslots).  ps: the P-stationary kernel's program."""
    L = lib()
    fn = L.pb200_plan_tree_ps if ps else L.pb200_plan_tree
    nson = np.ascontiguousarray(nson, np.int32)
    sons = np.ascontiguousarray(sons, np.int32)
    maxsons = sons.shape[1]
    flags = None if scale_flags is None else np.ascontiguousarray(scale_flags, np.uint8)
    nslots = C.c_int(0)
    need = fn(C.c_int(nnode), C.c_int(ns), C.c_int(root), _ptr(nson, C.c_int), _ptr(sons, C.c_int),
                             C.c_int(maxsons), _ptr(flags, C.c_ubyte), None, C.c_int(0), C.byref(nslots))
    if need < 0:
        raise PB200Error(L.pb200_last_error().decode())
    prog = np.zeros(need, np.int32)
    fn(C.c_int(nnode), C.c_int(ns), C.c_int(root), _ptr(nson, C.c_int), _ptr(sons, C.c_int),
                      C.c_int(maxsons), _ptr(flags, C.c_ubyte), _ptr(prog, C.c_int), C.c_int(need), C.byref(nslots))
    return prog, nslots.value


def plan_tree4(nnode, ns, root, nson, sons, scale_flags=None):
    """Planner of the 4-state kernel's macro-op program (no GPU): returns (ops int32 [nops, 2], tip_order, nslots)."""
    L = lib()
    nson = np.ascontiguousarray(nson, np.int32)
    sons = np.ascontiguousarray(sons, np.int32)
    maxsons = sons.shape[1]
    flags = None if scale_flags is None else np.ascontiguousarray(scale_flags, np.uint8)
    nslots, ntip = C.c_int(0), C.c_int(0)
    args = (C.c_int(nnode), C.c_int(ns), C.c_int(root), _ptr(nson, C.c_int), _ptr(sons, C.c_int), C.c_int(maxsons), _ptr(flags, C.c_ubyte))
    need = L.pb200_plan_tree4(*args, None, C.c_int(0), None, C.c_int(0), C.byref(ntip), C.byref(nslots))
    if need < 0:
        raise PB200Error(L.pb200_last_error().decode())
    ops = np.zeros(need, np.int32)
    tips = np.zeros(ntip.value, np.int32)
    L.pb200_plan_tree4(*args, _ptr(ops, C.c_int), C.c_int(need), _ptr(tips, C.c_int), C.c_int(ntip.value), C.byref(ntip), C.byref(nslots))
    return ops.reshape(-1, 2), tips, nslots.value


def default_scale_flags(nnode, root, nson, sons, every):
    nson = np.ascontiguousarray(nson, np.int32)
    sons = np.ascontiguousarray(sons, np.int32)
    flags = np.zeros(nnode, np.uint8)
    k = lib().pb200_default_scale_flags(C.c_int(nnode), C.c_int(root), _ptr(nson, C.c_int), _ptr(sons, C.c_int),
                                        C.c_int(sons.shape[1]), C.c_int(every), _ptr(flags, C.c_ubyte))
    return k, flags


class TreeLikelihood:
    """GPU tree-likelihood context for one shard of site patterns on one device."""

    def __init__(self, device=0):
        self._L = lib()
        self._ctx = C.c_void_p()
        check(self._L.pb200_create(C.byref(self._ctx), C.c_int(device)))
        self.npatt = 0
        self.nclass = 1
        self.nnode = 0

    def close(self):
        if self._ctx:
            self._L.pb200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- static part ------------------------------------------------------------------------------
    def set_data(self, n, z, fpatt, cleandata=1, nchara=None, charamap=None):
        z = np.ascontiguousarray(z, np.uint8)
        fpatt = np.ascontiguousarray(fpatt, np.float64)
        ns, npatt = z.shape
        assert fpatt.shape == (npatt,)
        nchara = None if nchara is None else np.ascontiguousarray(nchara, np.uint8)
        charamap = None if charamap is None else np.ascontiguousarray(charamap, np.uint8)
        check(self._L.pb200_set_data(self._ctx, C.c_int(n), C.c_int(ns), C.c_longlong(npatt), _ptr(z, C.c_ubyte),
                                     _ptr(fpatt, C.c_double), C.c_int(cleandata), _ptr(nchara, C.c_ubyte),
                                     _ptr(charamap, C.c_ubyte)))
        self.n, self.ns, self.npatt = n, ns, npatt

    def set_tree(self, root, nson, sons, scale_flags=None):
        nson = np.ascontiguousarray(nson, np.int32)
        sons = np.ascontiguousarray(sons, np.int32)
        flags = None if scale_flags is None else np.ascontiguousarray(scale_flags, np.uint8)
        nnode = nson.shape[0]
        check(self._L.pb200_set_tree(self._ctx, C.c_int(nnode), C.c_int(root), _ptr(nson, C.c_int), _ptr(sons, C.c_int),
                                     C.c_int(sons.shape[1]), _ptr(flags, C.c_ubyte)))
        self.nnode = nnode

    def set_problem(self, p, h0=0, h1=None):
        """Upload the static part of a Problem, optionally only the pattern shard [h0, h1)."""
        h1 = p.npatt if h1 is None else h1
        self.set_data(p.n, p.z[:, h0:h1], p.fpatt[h0:h1], p.cleandata, p.nchara, p.charamap)
        self.set_tree(p.root, p.nson, p.sons, p.scale_flags)

    # -- per evaluation ---------------------------------------------------------------------------
    def eval(self, semantics, freqK, set_kind, nr, set_data, branch_set, branch_t, pi):
        freqK = np.ascontiguousarray(freqK, np.float64)
        set_data = np.ascontiguousarray(set_data, np.float64)
        if set_data.ndim == 1:
            set_data = set_data.reshape(1, -1)
        branch_set = np.ascontiguousarray(branch_set, np.int32)
        branch_t = np.ascontiguousarray(branch_t, np.float64)
        pi = np.ascontiguousarray(pi, np.float64)
        nclass = freqK.shape[0]
        assert branch_set.shape == (nclass, self.nnode) and branch_t.shape == (nclass, self.nnode)
        out = C.c_double(0)
        check(self._L.pb200_eval(self._ctx, C.c_int(semantics), C.c_int(nclass), _ptr(freqK, C.c_double),
                                 C.c_int(set_kind), C.c_int(nr), C.c_int(set_data.shape[0]),
                                 C.c_longlong(set_data.shape[1]), _ptr(set_data, C.c_double),
                                 _ptr(branch_set, C.c_int), _ptr(branch_t, C.c_double), _ptr(pi, C.c_double),
                                 C.byref(out)))
        self.nclass = nclass
        return out.value

    def eval_problem(self, p):
        return self.eval(p.semantics, p.freqK, p.set_kind, p.nr, p.set_data, p.branch_set, p.branch_t, p.pi)

    def eval_begin_problem(self, p):
        """pb200_eval_begin: enqueue the evaluation and return without waiting (collect with eval_end)."""
        freqK = np.ascontiguousarray(p.freqK, np.float64)
        set_data = np.ascontiguousarray(p.set_data, np.float64)
        if set_data.ndim == 1:
            set_data = set_data.reshape(1, -1)
        branch_set = np.ascontiguousarray(p.branch_set, np.int32)
        branch_t = np.ascontiguousarray(p.branch_t, np.float64)
        pi = np.ascontiguousarray(p.pi, np.float64)
        check(self._L.pb200_eval_begin(self._ctx, C.c_int(p.semantics), C.c_int(freqK.shape[0]), _ptr(freqK, C.c_double),
                                       C.c_int(p.set_kind), C.c_int(p.nr), C.c_int(set_data.shape[0]),
                                       C.c_longlong(set_data.shape[1]), _ptr(set_data, C.c_double),
                                       _ptr(branch_set, C.c_int), _ptr(branch_t, C.c_double), _ptr(pi, C.c_double)))
        self.nclass = freqK.shape[0]

    def eval_end(self):
        out = C.c_double(0)
        check(self._L.pb200_eval_end(self._ctx, C.byref(out)))
        return out.value

    def lnf(self):
        out = np.zeros(self.npatt)
        check(self._L.pb200_get_lnf(self._ctx, _ptr(out, C.c_double)))
        return out

    def fhK(self):
        out = np.zeros((self.nclass, self.npatt))
        check(self._L.pb200_get_fhK(self._ctx, _ptr(out, C.c_double)))
        return out

    def pmat(self, n, set_kind, nr, set_data, mat_set, mat_t):
        set_data = np.ascontiguousarray(set_data, np.float64)
        if set_data.ndim == 1:
            set_data = set_data.reshape(1, -1)
        mat_set = np.ascontiguousarray(mat_set, np.int32)
        mat_t = np.ascontiguousarray(mat_t, np.float64)
        nmat = mat_set.shape[0]
        out = np.zeros((nmat, n, n))
        check(self._L.pb200_pmat(self._ctx, C.c_int(n), C.c_int(set_kind), C.c_int(nr), C.c_int(set_data.shape[0]),
                                 C.c_longlong(set_data.shape[1]), _ptr(set_data, C.c_double), C.c_int(nmat),
                                 _ptr(mat_set, C.c_int), _ptr(mat_t, C.c_double), _ptr(out, C.c_double)))
        return out

    # -- conditional vectors resident in HBM (minbranches family) ---------------------------------------
    def cond_update(self, set_kind, nr, set_data, branch_set, branch_t, updates, scale_flags=None):
        """pb200_cond_update.  updates: list of (node, [sons]) in the order to compute (sons before fathers)."""
        set_data = np.ascontiguousarray(set_data, np.float64)
        if set_data.ndim == 1:
            set_data = set_data.reshape(1, -1)
        branch_set = np.ascontiguousarray(branch_set, np.int32)
        branch_t = np.ascontiguousarray(branch_t, np.float64)
        nclass = branch_set.shape[0]
        maxsons = max([len(sn) for _, sn in updates] + [1])
        un = np.array([u for u, _ in updates], np.int32)
        uc = np.array([len(sn) for _, sn in updates], np.int32)
        us = np.full((max(1, len(updates)), maxsons), -1, np.int32)
        for i, (_, sn) in enumerate(updates):
            us[i, :len(sn)] = sn
        flags = None if scale_flags is None else np.ascontiguousarray(scale_flags, np.uint8)
        check(self._L.pb200_cond_update(self._ctx, C.c_int(nclass), C.c_int(set_kind), C.c_int(nr), C.c_int(set_data.shape[0]),
                                        C.c_longlong(set_data.shape[1]), _ptr(set_data, C.c_double), _ptr(branch_set, C.c_int),
                                        _ptr(branch_t, C.c_double), C.c_int(len(updates)), _ptr(un, C.c_int), _ptr(uc, C.c_int),
                                        _ptr(us, C.c_int), C.c_int(maxsons), _ptr(flags, C.c_ubyte)))
        self.cond_nclass = nclass
        self.cond_nscale = 0 if flags is None else int(flags.astype(bool).sum())

    def branch_lnl(self, freqK, a, b, t, mult, set_of_class, pi, semantics, want_dd=True):
        """pb200_branch_lnl -> (-lnL, d/dt, d2/dt2) for the branch b -> a at length t."""
        mult = np.ascontiguousarray(mult, np.float64)
        soc = np.ascontiguousarray(set_of_class, np.int32)
        pi = np.ascontiguousarray(pi, np.float64)
        fk = None if freqK is None else np.ascontiguousarray(freqK, np.float64)
        out = np.zeros(3)
        check(self._L.pb200_branch_lnl(self._ctx, C.c_int(mult.shape[0]), _ptr(fk, C.c_double), C.c_int(a), C.c_int(b),
                                       C.c_double(float(t)), _ptr(mult, C.c_double), _ptr(soc, C.c_int), _ptr(pi, C.c_double),
                                       C.c_int(semantics), C.c_int(1 if want_dd else 0), _ptr(out, C.c_double)))
        return out

    def branch_newton(self, freqK, a, b, t, l0, mult, set_of_class, pi, semantics, first_cycle=False, e=1e-6, tb=(1e-8, 50.0)):
        """pb200_branch_newton -> (t, -lnL, dl, ddl, number of evaluations): the whole iteration of one branch on the device."""
        mult = np.ascontiguousarray(mult, np.float64)
        soc = np.ascontiguousarray(set_of_class, np.int32)
        pi = np.ascontiguousarray(pi, np.float64)
        fk = None if freqK is None else np.ascontiguousarray(freqK, np.float64)
        tbv = np.ascontiguousarray(tb, np.float64)
        out = np.zeros(4)
        nev = C.c_int(0)
        check(self._L.pb200_branch_newton(self._ctx, C.c_int(mult.shape[0]), _ptr(fk, C.c_double), C.c_int(a), C.c_int(b),
                                          C.c_double(float(t)), C.c_double(float(l0)), _ptr(mult, C.c_double), _ptr(soc, C.c_int),
                                          _ptr(pi, C.c_double), C.c_int(semantics), C.c_int(1 if first_cycle else 0), C.c_double(float(e)),
                                          _ptr(tbv, C.c_double), _ptr(out, C.c_double), C.byref(nev)))
        return out[0], out[1], out[2], out[3], nev.value

    def cond_get(self, node):
        """Resident vectors of one internal node [nclass, npatt, n] and the scale factors [nclass, nscale, npatt]."""
        cond = np.zeros((self.cond_nclass, self.npatt, self.n))
        sc = np.zeros((self.cond_nclass, max(1, self.cond_nscale), self.npatt))
        check(self._L.pb200_cond_get(self._ctx, C.c_int(node), _ptr(cond, C.c_double), _ptr(sc, C.c_double)))
        return cond, sc[:, :self.cond_nscale]

    def eigen_qrev(self, Q, pi):
        """pb200_eigen_qrev: Q [njob, n, n] (or [n, n]), pi [njob, n] -> Root [njob, n], U, V [njob, n, n], sweeps [njob]."""
        Q = np.ascontiguousarray(Q, np.float64)
        pi = np.ascontiguousarray(pi, np.float64)
        if Q.ndim == 2:
            Q = Q[None]; pi = pi[None]
        njob, n = Q.shape[0], Q.shape[1]
        Root = np.zeros((njob, n)); U = np.zeros((njob, n, n)); V = np.zeros((njob, n, n)); sw = np.zeros(njob, np.int32)
        check(self._L.pb200_eigen_qrev(self._ctx, C.c_int(n), C.c_int(njob), _ptr(Q, C.c_double), _ptr(pi, C.c_double),
                                       _ptr(Root, C.c_double), _ptr(U, C.c_double), _ptr(V, C.c_double), _ptr(sw, C.c_int)))
        return Root, U, V, sw

    # -- multi-GPU --------------------------------------------------------------------------------
    @staticmethod
    def nccl_unique_id():
        buf = C.create_string_buffer(128)
        check(lib().pb200_nccl_unique_id(buf))
        return buf.raw

    def comm_init(self, nranks, rank, uid):
        check(self._L.pb200_comm_init(self._ctx, C.c_int(nranks), C.c_int(rank), C.c_char_p(uid)))

    def p2p_local_handle(self):
        buf = C.create_string_buffer(64)
        check(self._L.pb200_p2p_local_handle(self._ctx, buf))
        return buf.raw

    def p2p_init(self, nranks, rank, handles):
        """handles: list of the nranks 64-byte handles in rank order (all-gathered by the launcher)."""
        blob = b"".join(handles)
        assert len(blob) == 64 * nranks
        check(self._L.pb200_p2p_init(self._ctx, C.c_int(nranks), C.c_int(rank), C.c_char_p(blob)))

    def p2p_enable(self, on=True):
        check(self._L.pb200_p2p_enable(self._ctx, C.c_int(1 if on else 0)))

    # -- timing / introspection -------------------------------------------------------------------
    def set_timing(self, enable=True):
        check(self._L.pb200_set_timing(self._ctx, C.c_int(1 if enable else 0)))

    def last_timing(self):
        ms = (C.c_double * 5)()
        check(self._L.pb200_last_timing(self._ctx, ms))
        return dict(total=ms[0], pt=ms[1], prune=ms[2], mix=ms[3], allreduce=ms[4])

    def launch_count(self):
        return int(self._L.pb200_launch_count(self._ctx))

    def last_kernel(self):
        return self._L.pb200_last_kernel(self._ctx).decode()

    def measure_fp64_peak(self):
        a, b = C.c_double(0), C.c_double(0)
        check(self._L.pb200_measure_fp64_peak(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def bench_resident(self, steps, flush_bytes=0):
        """Per-step device times (ms) of `steps` back-to-back evaluations with everything resident in HBM."""
        arrs = [np.zeros(steps) for _ in range(4)]
        check(self._L.pb200_bench_resident(self._ctx, C.c_int(steps), C.c_longlong(flush_bytes),
                                           *[_ptr(a, C.c_double) for a in arrs]))
        return dict(step=arrs[0], pt=arrs[1], prune=arrs[2], mix=arrs[3])

    def bench_e2e(self, p, steps, flush_bytes=0):
        """Seconds per pb200_eval call (host clock inside the library, host buffers in, -lnL out), `steps` calls."""
        freqK = np.ascontiguousarray(p.freqK, np.float64)
        sets = np.ascontiguousarray(p.set_data, np.float64)
        bset = np.ascontiguousarray(p.branch_set, np.int32)
        bt = np.ascontiguousarray(p.branch_t, np.float64)
        pi = np.ascontiguousarray(p.pi, np.float64)
        sec = np.zeros(steps)
        out = C.c_double(0)
        check(self._L.pb200_bench_e2e(self._ctx, C.c_int(steps), C.c_longlong(flush_bytes), C.c_int(p.semantics),
                                      C.c_int(freqK.shape[0]), _ptr(freqK, C.c_double), C.c_int(p.set_kind), C.c_int(p.nr),
                                      C.c_int(sets.shape[0]), C.c_longlong(sets.shape[1]), _ptr(sets, C.c_double),
                                      _ptr(bset, C.c_int), _ptr(bt, C.c_double), _ptr(pi, C.c_double),
                                      _ptr(sec, C.c_double), C.byref(out)))
        self.nclass = freqK.shape[0]
        return sec, out.value

    def flush_l2(self, nbytes=256 << 20):
        check(self._L.pb200_flush_l2(self._ctx, C.c_longlong(nbytes)))
