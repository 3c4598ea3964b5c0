"""Reader for the .pbfx golden fixtures (written by oracle/ref_hook.inc from the unmodified reference).

A fixture holds one complete likelihood evaluation: every hot-path input in the flat layout of the
C-ABI (include/paml_b200.h) plus the reference's outputs.  This module is plain file parsing and is
used by tests/, bench.py and __graft_entry__.smoke(); it has no dependency on oracle/.
"""
from dataclasses import dataclass, field
import numpy as np

KIND_UVROOT, KIND_CIJK, KIND_K80, KIND_JC69LIKE = 0, 1, 2, 3
SEM_LFUN, SEM_LFUNDG = 0, 1


@dataclass
class Problem:
    """One likelihood problem: static part (data + tree) and per-eval part (spectral sets, t's)."""
    program: int = 1
    n: int = 61
    ns: int = 0
    nnode: int = 0
    root: int = 0
    maxsons: int = 3
    cleandata: int = 1
    npatt: int = 0
    father: np.ndarray = None      # int32 [nnode]
    nson: np.ndarray = None        # int32 [nnode]
    sons: np.ndarray = None        # int32 [nnode, maxsons]
    scale_flags: np.ndarray = None  # uint8 [nnode]
    z: np.ndarray = None           # uint8 [ns, npatt]
    fpatt: np.ndarray = None       # float64 [npatt]
    nchara: np.ndarray = None      # uint8 [256]
    charamap: np.ndarray = None    # uint8 [256, 64]
    pi: np.ndarray = None          # float64 [n]
    semantics: int = SEM_LFUN
    nclass: int = 1
    freqK: np.ndarray = None       # float64 [nclass]
    set_kind: int = KIND_UVROOT
    nr: int = 0
    nset: int = 1
    set_stride: int = 0
    set_data: np.ndarray = None    # float64 [nset, set_stride]
    branch_set: np.ndarray = None  # int32 [nclass, nnode]
    branch_t: np.ndarray = None    # float64 [nclass, nnode]
    # reference outputs (fixtures only)
    ref_neg_lnL: float = None
    ref_lnf: np.ndarray = None     # [npatt]
    ref_fhK: np.ndarray = None     # [nclass, npatt]
    ref_P: list = field(default_factory=list)   # [(class, node, P[n,n])]
    x: np.ndarray = None
    seqtype: int = -1
    model: int = -1
    nscale: int = 0
    branch_len: np.ndarray = None  # float64 [nnode]: the tree's own branch lengths (synthetic problems; branch_t = these x class rate)
    model_name: str = None
    branch_rec: dict = None        # branch fixtures: a, b, siteclass, t and the reference's lfuntdd outputs l, dl, ddl

    @property
    def ninternal(self):
        return self.nnode - self.ns

    def updates_per_eval(self):
        """site-pattern x node x class updates per evaluation."""
        return int(self.npatt) * self.ninternal * self.nclass


def default_ambiguity_maps(n):
    nchara = np.zeros(256, np.uint8)
    charamap = np.zeros((256, 64), np.uint8)
    nchara[:n] = 1
    charamap[np.arange(n), 0] = np.arange(n)
    return nchara, charamap


def load_fixture(path):
    buf = open(path, "rb").read()
    assert buf[:8] == b"PBFX0001", "bad fixture magic"
    off = 8

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(buf, dtype=dtype, count=count, offset=off)
        off += a.nbytes
        return a.copy()

    hdr = take(np.int32, 16)
    p = Problem()
    (p.program, p.n, p.ns, p.nnode, p.root, p.maxsons, p.cleandata, p.nclass, p.nset, p.set_kind, p.nr,
     p.semantics, p.nscale, ndump, p.seqtype, p.model) = [int(v) for v in hdr]
    p.npatt, p.set_stride, np_ = [int(v) for v in take(np.int64, 3)]
    N, n = p.npatt, p.n
    p.father = take(np.int32, p.nnode)
    p.nson = take(np.int32, p.nnode)
    p.sons = take(np.int32, p.nnode * p.maxsons).reshape(p.nnode, p.maxsons)
    p.scale_flags = take(np.uint8, p.nnode)
    p.z = take(np.uint8, p.ns * N).reshape(p.ns, N)
    p.fpatt = take(np.float64, N)
    p.nchara = take(np.uint8, 256)
    p.charamap = take(np.uint8, 256 * 64).reshape(256, 64)
    p.pi = take(np.float64, 64)[:n].copy()
    p.freqK = take(np.float64, 128)[:p.nclass].copy()
    p.set_data = take(np.float64, p.nset * p.set_stride).reshape(p.nset, p.set_stride)
    p.branch_set = take(np.int32, p.nclass * p.nnode).reshape(p.nclass, p.nnode)
    p.branch_t = take(np.float64, p.nclass * p.nnode).reshape(p.nclass, p.nnode)
    p.ref_neg_lnL = float(take(np.float64, 1)[0])
    p.ref_lnf = take(np.float64, N)
    if p.semantics == SEM_LFUNDG:
        p.ref_fhK = take(np.float64, p.nclass * N).reshape(p.nclass, N)
    cls = take(np.int32, ndump)
    node = take(np.int32, ndump)
    Ps = take(np.float64, ndump * n * n).reshape(ndump, n, n)
    p.ref_P = [(int(cls[i]), int(node[i]), Ps[i]) for i in range(ndump)]
    p.x = take(np.float64, max(np_, 0))
    if buf[off:off + 8] == b"BRLN0001":        # branch fixture (oracle/ref_hook.inc, PBX_BRANCH_DUMP)
        off += 8
        a, b, sc, _ = [int(v) for v in take(np.int32, 4)]
        t, l, dl, ddl, blen = [float(v) for v in take(np.float64, 5)]
        p.branch_rec = dict(a=a, b=b, siteclass=sc, t=t, l=l, dl=dl, ddl=ddl, branch_len=blen)
    assert off == len(buf), "fixture has trailing bytes"
    return p
