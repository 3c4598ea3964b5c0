"""Site-pattern compression (pb200_compress_patterns, the library's PatternWeight -- SURVEY.md 8f rank 3): bit-exact
pattern order, counts and site map against (a) an independent restatement of the reference's definition
(src/treesub.c:1386-1516: per gene, distinct columns in ascending strcmp order of character+1, first site represents)
and (b) the unmodified reference binary's own "site pattern counts" listing.  Host code: runs without a GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

from paml_b200 import paml_io, patterns, simulate

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, "oracle", "_ref")


def restatement(z, width, gene, ngene):
    """Plain-Python PatternWeight: what the reference's sorted insertion ends with."""
    ns, ls = z.shape[0], z.shape[1] // width
    cols = [bytes(((z[:, h * width:(h + 1) * width].astype(np.int32) + 1) & 255).astype(np.uint8).reshape(-1)) for h in range(ls)]
    cols = [c.split(b"\0")[0] for c in cols]                 # strcmp stops at a zero byte
    posG, pats, first = [0], [], []
    pose = np.zeros(ls, np.int64)
    for g in range(ngene):
        sites = [h for h in range(ls) if (gene[h] if gene is not None else 0) == g]
        seen = {}
        for h in sites:
            seen.setdefault(cols[h], h)
        keys = sorted(seen)                                   # bytes compare as unsigned char, like strcmp
        idx = {k: len(pats) + i for i, k in enumerate(keys)}
        for h in sites:
            pose[h] = idx[cols[h]]
        pats += keys
        first += [seen[k] for k in keys]
        posG.append(len(pats))
    fpatt = np.bincount(pose, minlength=len(pats)).astype(float)
    return dict(npatt=len(pats), posG=np.array(posG), pose=pose, fpatt=fpatt, rep=np.array(first))


def _check(z, width=1, gene=None, ngene=1):
    got = patterns.pattern_weight(z, width, gene, ngene)
    exp = restatement(np.asarray(z, np.uint8), width, gene, ngene)
    assert got["npatt"] == exp["npatt"]
    np.testing.assert_array_equal(got["posG"], exp["posG"])
    np.testing.assert_array_equal(got["pose"], exp["pose"])
    np.testing.assert_array_equal(got["fpatt"], exp["fpatt"])
    np.testing.assert_array_equal(got["rep"], exp["rep"])
    assert got["fpatt"].sum() == z.shape[1] // width
    return got


@pytest.mark.parametrize("ns,ls,alphabet,width", [(5, 400, b"TCAG", 1), (12, 3000, b"TCAG", 1), (7, 900, b"ARNDCQEGHILKMFPSTWYV", 1),
                                                   (4, 500, b"TCAG", 3), (1, 50, b"TCAG", 1), (9, 1, b"TCAG", 1)])
def test_random_alignments(ns, ls, alphabet, width):
    rng = np.random.default_rng(ns * 1000 + ls)
    anc = rng.integers(0, len(alphabet), ls * width)
    z = np.array([[alphabet[a if rng.random() > 0.15 else rng.integers(0, len(alphabet))] for a in anc] for _ in range(ns)], np.uint8)
    _check(z, width)


def test_genes_and_ambiguity_characters():
    rng = np.random.default_rng(7)
    alphabet = b"TCAGNRY-?"
    ls = 1200
    z = np.array([[alphabet[i] for i in rng.choice(len(alphabet), ls, p=[.3, .3, .15, .15, .03, .02, .02, .02, .01])] for _ in range(6)], np.uint8)
    z[1:] = np.where(rng.random((5, ls)) < 0.7, z[0], z[1:])
    gene = rng.integers(0, 3, ls).astype(np.int32)
    got = _check(z, 1, gene, 3)
    assert (np.diff(got["posG"]) > 0).all()
    # site marks on a later line, genes interleaved: every pattern index lies inside its gene's range
    for g in range(3):
        assert ((got["pose"][gene == g] >= got["posG"][g]) & (got["pose"][gene == g] < got["posG"][g + 1])).all()


def test_coded_bytes_including_255_and_zero():
    """Coded data (0..n-1) works like raw characters; a byte of 255 becomes the string terminator after the +1."""
    rng = np.random.default_rng(11)
    z = rng.integers(0, 4, (5, 300)).astype(np.uint8)
    z[2, ::17] = 255
    _check(z)


def test_all_identical_and_all_distinct():
    z = np.full((4, 64), ord("A"), np.uint8)
    got = _check(z)
    assert got["npatt"] == 1 and got["fpatt"][0] == 64
    z = np.array([[b"TCAG"[(h >> (2 * j)) & 3] for h in range(256)] for j in range(4)], np.uint8)
    got = _check(z)
    assert got["npatt"] == 256


def test_against_reference_listing(tmp_path):
    """The unmodified reference prints the compressed alignment (patterns in ITS order, then the counts) when
    verbose = 1: ours must be identical, character for character."""
    exe = os.path.join(REF, "baseml")
    if not os.path.exists(exe):
        pytest.skip("reference binary not built")
    p = simulate.nuc_gtr_gamma_problem(ns=7, npatt=400, seed=61, K=1)
    ls = paml_io.write_sequence_files(p, str(tmp_path / "s"), seed=5)
    with open(tmp_path / "run.ctl", "w") as f:
        f.write("seqfile = s.phy\ntreefile = s.tre\noutfile = mlb\nnoisy = 0\nverbose = 1\nrunmode = 0\nmodel = 4\nMgene = 0\n"
                "clock = 0\nfix_kappa = 1\nkappa = 2\nfix_alpha = 1\nalpha = 0\nMalpha = 0\nncatG = 1\nnhomo = 0\ngetSE = 0\n"
                "RateAncestor = 0\ncleandata = 1\nfix_blength = 2\nmethod = 0\n")
    r = subprocess.run([exe, "run.ctl"], cwd=str(tmp_path), stdin=subprocess.DEVNULL, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:]
    txt = open(tmp_path / "mlb").read()
    blk = txt[txt.index("Printing out site pattern counts"):]
    m = re.search(r"\n\s*(\d+)\s+(\d+)\s+P\s*\n", blk)
    ns, npatt = int(m.group(1)), int(m.group(2))
    rest = blk[m.end():]
    seqs = []
    lines = [ln for ln in rest.splitlines()]
    i = 0
    while len(seqs) < ns:
        ln = lines[i]; i += 1
        if ln.startswith("s"):
            seqs.append("".join(ln.split()[1:]))
    counts = []
    while len(counts) < npatt:
        counts += [float(t) for t in lines[i].split()]; i += 1
    # our compression of the same file
    raw = [ln.split(None, 1)[1].strip() for ln in open(tmp_path / "s.phy").read().splitlines()[1:] if ln.strip()]
    z = np.array([list(s.encode()) for s in raw], np.uint8)
    assert z.shape == (ns, ls)
    got = patterns.pattern_weight(z)
    assert got["npatt"] == npatt
    assert ["".join(chr(c) for c in row) for row in got["z"]] == seqs
    np.testing.assert_array_equal(got["fpatt"], np.array(counts))
