"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
        _LIB = C.CDLL(so)
        _LIB.orc_lnl.restype = C.c_int
    return _LIB


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def pmat(t, n, kind, nr, set_row):
    P = np.zeros((n, n))
    s = np.ascontiguousarray(set_row, dtype=np.float64)
    _lib().orc_pmat(_p(P, C.c_double), C.c_double(float(t)), C.c_int(n), C.c_int(kind), C.c_int(nr), _p(s, C.c_double))
    return P


def set_node_scale(p, every):
    flags = np.zeros(p.nnode, np.uint8)
    nson = np.ascontiguousarray(p.nson, np.int32)
    sons = np.ascontiguousarray(p.sons, np.int32)
    k = _lib().orc_set_node_scale(C.c_int(p.nnode), C.c_int(p.root), _p(nson, C.c_int), _p(sons, C.c_int),
                                  C.c_int(p.maxsons), C.c_int(every), _p(flags, C.c_ubyte))
    return k, flags


def lnl(p, want_conP=False):
    """Evaluate problem p (paml_b200.fixture.Problem) with the CPU oracle. Returns dict."""
    N, K = int(p.npatt), int(p.nclass)
    z = np.ascontiguousarray(p.z, np.uint8)
    fpatt = np.ascontiguousarray(p.fpatt, np.float64)
    nson = np.ascontiguousarray(p.nson, np.int32)
    sons = np.ascontiguousarray(p.sons, np.int32)
    flags = np.ascontiguousarray(p.scale_flags if p.scale_flags is not None else np.zeros(p.nnode, np.uint8), np.uint8)
    nchara = np.ascontiguousarray(p.nchara, np.uint8)
    charamap = np.ascontiguousarray(p.charamap, np.uint8)
    freqK = np.ascontiguousarray(p.freqK, np.float64)
    sets = np.ascontiguousarray(p.set_data, np.float64)
    bset = np.ascontiguousarray(p.branch_set, np.int32)
    bt = np.ascontiguousarray(p.branch_t, np.float64)
    pi = np.ascontiguousarray(p.pi, np.float64)
    out = C.c_double(0)
    lnf = np.zeros(N)
    fhK = np.zeros((K, N))
    conP = np.zeros((K, p.nnode - p.ns, N, p.n)) if want_conP else None
    rc = _lib().orc_lnl(C.c_int(p.n), C.c_int(p.ns), C.c_int(p.nnode), C.c_int(p.root), C.c_int(p.maxsons),
                        C.c_longlong(N), C.c_int(p.cleandata), _p(nson, C.c_int), _p(sons, C.c_int),
                        _p(flags, C.c_ubyte), _p(z, C.c_ubyte), _p(fpatt, C.c_double),
                        _p(nchara, C.c_ubyte), _p(charamap, C.c_ubyte),
                        C.c_int(p.semantics), C.c_int(K), _p(freqK, C.c_double),
                        C.c_int(p.set_kind), C.c_int(p.nr), C.c_int(p.nset), C.c_longlong(p.set_stride),
                        _p(sets, C.c_double), _p(bset, C.c_int), _p(bt, C.c_double), _p(pi, C.c_double),
                        C.byref(out), _p(lnf, C.c_double), _p(fhK, C.c_double), _p(conP, C.c_double))
    if rc != 0:
        raise MemoryError("oracle out of memory")
    return dict(neg_lnL=out.value, lnf=lnf, fhK=fhK, conP=conP)
