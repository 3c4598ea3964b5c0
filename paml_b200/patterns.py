"""Site-pattern compression in the reference's order (PatternWeight, src/treesub.c:1386-1516).

The reference collapses identical alignment columns and keeps the distinct ones SORTED in ascending
strcmp order of the column's raw characters (src/treesub.c:1421-1456); fpatt[] are the multiplicities and
pose[] maps each site to its pattern.  For coded states whose code order equals the character order of
the alphabet used here, sorting the code columns lexicographically gives the same order; callers with raw
characters pass them as bytes.
"""
import numpy as np


def compress_patterns(cols):
    """cols: uint8 [ns, nsites] (codes or raw characters).  Returns (z [ns, npatt], counts [npatt]) with
    patterns in ascending lexicographic (= strcmp) order of the column bytes."""
    z, counts = np.unique(np.ascontiguousarray(cols.T), axis=0, return_counts=True)
    return np.ascontiguousarray(z.T), counts


def compress_with_positions(cols):
    """As compress_patterns, plus pose[nsites] (site -> pattern index), the reference's com.pose."""
    z, inv, counts = np.unique(np.ascontiguousarray(cols.T), axis=0, return_inverse=True, return_counts=True)
    return np.ascontiguousarray(z.T), counts, inv.reshape(-1)
