"""Writers for the reference's input files (pattern-format sequence file, tree file, control file), so a
synthetic Problem can be handed to the UNMODIFIED reference binaries (oracle/_ref) for the CPU baseline.

Pattern format (src/treesub.c:549, 951-983): first line ``ns L P``; each sequence as a name followed by two
spaces and the characters; then the npatt pattern counts.  The reference keeps the patterns in file order."""
import numpy as np

from . import models


def _tree_newick(p, branch):
    def rec(v):
        if p.nson[v] == 0:
            return "%d:%.8f" % (v + 1, branch[v])
        inner = ",".join(rec(int(c)) for c in p.sons[v, :p.nson[v]])
        return "(%s)" % inner + ("" if v == p.root else ":%.8f" % branch[v])
    return rec(p.root) + ";"


def write_pattern_files(p, prefix, h1=None, genes=None):
    """Write <prefix>.phy (pattern format) and <prefix>.tre for Problem p (clean data only).
    Branch lengths are class 0's t (for a one-class problem: the branch lengths themselves).
    genes: pattern counts per gene -> option ``PG`` with a ``G ngene n1 n2 ..`` line (src/treesub.c:624-657)."""
    assert p.cleandata, "pattern writer handles resolved states only"
    h1 = p.npatt if h1 is None else min(h1, p.npatt)
    if p.n == 61:
        from61, _ = models.sense_codons()
        chars = np.array([[models.BASES[ic // 16], models.BASES[(ic // 4) % 4], models.BASES[ic % 4]] for ic in from61])
        width = 3
    elif p.n == 20:
        chars = np.array([[c] for c in models.AAS]); width = 1
    else:
        chars = np.array([[c] for c in models.BASES]); width = 1
    with open(prefix + ".phy", "w") as f:
        if genes:
            assert sum(genes) == h1
            f.write("%d %d PG\nG %d %s\n" % (p.ns, h1 * width, len(genes), " ".join(str(g) for g in genes)))
        else:
            f.write("%d %d P\n" % (p.ns, h1 * width))
        for i in range(p.ns):
            f.write("s%d  " % (i + 1))
            f.write("".join("".join(row) for row in chars[p.z[i, :h1]]))
            f.write("\n")
        f.write(" ".join("%d" % w for w in p.fpatt[:h1]))
        f.write("\n")
    with open(prefix + ".tre", "w") as f:
        f.write("%d 1\n%s\n" % (p.ns, _tree_newick(p, p.branch_len if getattr(p, "branch_len", None) is not None else p.branch_t[0])))
    return h1


def write_sequence_files(p, prefix, seed=1):
    """Write <prefix>.phy as an ordinary sequential alignment (every pattern repeated fpatt times, the sites shuffled)
    and <prefix>.tre: the reference then compresses the sites itself (PatternWeight, src/treesub.c:1386-1516)."""
    assert p.cleandata
    if p.n == 61:
        from61, _ = models.sense_codons()
        chars = np.array(["".join(models.BASES[k] for k in (ic // 16, (ic // 4) % 4, ic % 4)) for ic in from61]); width = 3
    elif p.n == 20:
        chars = np.array(list(models.AAS)); width = 1
    else:
        chars = np.array(list(models.BASES)); width = 1
    sites = np.repeat(np.arange(p.npatt), np.asarray(p.fpatt, dtype=np.int64))
    np.random.default_rng(seed).shuffle(sites)
    with open(prefix + ".phy", "w") as f:
        f.write("%d %d\n" % (p.ns, len(sites) * width))
        for i in range(p.ns):
            f.write("s%d  %s\n" % (i + 1, "".join(chars[p.z[i, sites]])))
    with open(prefix + ".tre", "w") as f:
        f.write("%d 1\n%s\n" % (p.ns, _tree_newick(p, p.branch_t[0])))
    return len(sites)


def write_codeml_ctl(path, seqfile, treefile, seqtype=1, model=0, NSsites=0, kappa=5.0, omega=0.3, extra=None):
    """A codeml control file with every parameter fixed (fix_blength = 2: branch lengths from the tree file),
    so the run is one likelihood evaluation at a known point."""
    keys = dict(seqfile=seqfile, treefile=treefile, outfile="mlc", noisy=0, verbose=0, runmode=0, seqtype=seqtype,
                CodonFreq=2, clock=0, model=model, NSsites=NSsites, icode=0, fix_kappa=1, kappa=kappa, fix_omega=1,
                omega=omega, fix_alpha=1, alpha=0, ncatG=1, getSE=0, RateAncestor=0, Small_Diff=".5e-6",
                cleandata=1, fix_blength=2, method=0)
    keys.update(extra or {})
    with open(path, "w") as f:
        for k, v in keys.items():
            f.write("%s = %s\n" % (k, v))


def write_baseml_ctl(path, seqfile, treefile, model=7, alpha=0.5, ncatG=8, extra=None):
    """A baseml control file with branch lengths taken from the tree file (fix_blength = 2) and alpha fixed; the
    exchangeabilities of REV stay parameters -- fix them for a single evaluation with an `in.baseml` line `-1 a b c d e`."""
    keys = dict(seqfile=seqfile, treefile=treefile, outfile="mlb", noisy=0, verbose=0, runmode=0, model=model, Mgene=0,
                clock=0, fix_kappa=0, kappa=2, fix_alpha=1, alpha=alpha, Malpha=0, ncatG=ncatG, nhomo=0, getSE=0,
                RateAncestor=0, Small_Diff="7e-6", cleandata=1, fix_blength=2, method=0)
    keys.update(extra or {})
    with open(path, "w") as f:
        for k, v in keys.items():
            f.write("%s = %s\n" % (k, v))
