"""Site-pattern compression in the reference's order (PatternWeight, src/treesub.c:1386-1516).

The reference collapses identical alignment columns and keeps the distinct ones SORTED in ascending
strcmp order of the column's raw characters (src/treesub.c:1421-1456); fpatt[] are the multiplicities and
pose[] maps each site to its pattern.  For coded states whose code order equals the character order of
the alphabet used here, sorting the code columns lexicographically gives the same order; callers with raw
characters pass them as bytes.
"""
import ctypes as C

import numpy as np

from . import _lib


def compress_patterns(cols):
    """cols: uint8 [ns, nsites] (codes or raw characters).  Returns (z [ns, npatt], counts [npatt]) with
    patterns in ascending lexicographic (= strcmp) order of the column bytes."""
    cols = np.ascontiguousarray(cols, np.uint8)
    if cols.shape[1] > 50_000 and int(cols.max()) < 255:
        # large inputs: the library's hash + sort (csrc/patterns.cpp, host code) gives the same order and counts far faster
        r = pattern_weight(cols)
        return r["z"], r["fpatt"].astype(np.int64)
    z, counts = np.unique(np.ascontiguousarray(cols.T), axis=0, return_counts=True)
    return np.ascontiguousarray(z.T), counts


def compress_with_positions(cols):
    """As compress_patterns, plus pose[nsites] (site -> pattern index), the reference's com.pose."""
    z, inv, counts = np.unique(np.ascontiguousarray(cols.T), axis=0, return_inverse=True, return_counts=True)
    return np.ascontiguousarray(z.T), counts, inv.reshape(-1)


def pattern_weight(z, width=1, gene=None, ngene=1):
    """The library's PatternWeight (pb200_compress_patterns, csrc/patterns.cpp) through the C ABI.
    z: uint8 [ns, ls*width] (raw characters or codes); gene: int32 [ls] labels or None.
    Returns dict(npatt, posG [ngene+1], lgene [ngene], pose [ls], fpatt [npatt], rep [npatt], z [ns, npatt*width])."""
    z = np.ascontiguousarray(z, np.uint8)
    ns, ls = z.shape[0], z.shape[1] // width
    rows = (C.POINTER(C.c_ubyte) * ns)(*[z[j].ctypes.data_as(C.POINTER(C.c_ubyte)) for j in range(ns)])
    g = None if gene is None else np.ascontiguousarray(gene, np.int32)
    npatt = C.c_longlong(0)
    posG = np.zeros(ngene + 1, np.int64); lgene = np.zeros(ngene, np.int64)
    pose = np.zeros(ls, np.int32); fpatt = np.zeros(ls); rep = np.zeros(ls, np.int64)
    rc = _lib.lib().pb200_compress_patterns(
        C.c_int(ns), C.c_int(width), C.c_longlong(ls), rows, None if g is None else g.ctypes.data_as(C.POINTER(C.c_int)),
        C.c_int(ngene), C.byref(npatt), posG.ctypes.data_as(C.POINTER(C.c_longlong)), lgene.ctypes.data_as(C.POINTER(C.c_longlong)),
        pose.ctypes.data_as(C.POINTER(C.c_int)), fpatt.ctypes.data_as(C.POINTER(C.c_double)), rep.ctypes.data_as(C.POINTER(C.c_longlong)))
    if rc:
        raise _lib.PB200Error("pb200_compress_patterns failed (%d)" % rc)
    n = npatt.value
    rep = rep[:n]
    cols = (rep[:, None] * width + np.arange(width)[None, :]).reshape(-1)
    return dict(npatt=n, posG=posG, lgene=lgene, pose=pose, fpatt=fpatt[:n].copy(), rep=rep.copy(), z=z[:, cols].copy())
