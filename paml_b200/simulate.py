"""Synthetic alignments and ready-made likelihood problems of the BASELINE.json shapes.

There is no network and no reference tree on the GPU box, so bench.py and the -m gpu tests manufacture
their inputs here: sequences are evolved down a seeded random tree under the named model (the job the
reference's `evolver` does, src/evolver.c:753), collapsed into site patterns, and wrapped in a Problem
(paml_b200.fixture.Problem) in exactly the flat layout the C ABI takes.
"""
import numpy as np

from . import models
from .fixture import Problem, KIND_UVROOT, KIND_CIJK, SEM_LFUN, SEM_LFUNDG, default_ambiguity_maps
from .likelihood import default_scale_flags
from .patterns import compress_patterns
from .tree import parse_newick, random_tree


def evolve(tree, pmat_of_branch, pi, nsites, rng):
    """Evolve nsites independent sites down `tree`.  pmat_of_branch(node) -> P[n,n].  Returns uint8 [ns, nsites]."""
    n = len(pi)
    states = np.zeros((tree.nnode, nsites), np.uint8)
    states[tree.root] = rng.choice(n, size=nsites, p=pi / pi.sum())
    order = [tree.root]
    for v in order:
        for k in range(tree.nson[v]):
            c = int(tree.sons[v, k])
            P = pmat_of_branch(c)
            cdf = np.cumsum(P, axis=1)
            cdf /= cdf[:, -1:]
            # one sorted table for all parent states: entry s*n + j = s + cdf[s, j]; the child of a site with parent s and
            # uniform u is the number of entries of row s that are <= u, i.e. searchsorted(table, s + u) - s*n
            table = (np.arange(n)[:, None] + cdf).ravel()
            par = states[v].astype(np.int64)
            key = par + rng.random(nsites)
            child = np.minimum(np.searchsorted(table, key, side="right") - par * n, n - 1).astype(np.uint8)
            states[c] = child
            order.append(c)
    return states[:tree.ns]


def _finish(p, tree, every):
    p.ns, p.nnode, p.root, p.maxsons = tree.ns, tree.nnode, tree.root, tree.maxsons
    p.father, p.nson, p.sons = tree.father, tree.nson, tree.sons
    p.nscale, p.scale_flags = default_scale_flags(tree.nnode, tree.root, tree.nson, tree.sons, every)
    p.nchara, p.charamap = default_ambiguity_maps(p.n)
    p.cleandata = 1
    p.branch_len = tree.branch.copy()        # the tree's own branch lengths (branch_t carries them times the class rates)
    return p


def _patterns(z_sites, npatt, rng, n):
    """Collapse sites to patterns; top up with extra random distinct columns if the simulation gave fewer
    than npatt distinct patterns, or keep the first npatt (weights = counts)."""
    z, w = compress_patterns(z_sites)
    if z.shape[1] >= npatt:
        return z[:, :npatt].copy(), w[:npatt].astype(float)
    return z, w.astype(float)


def _grow_patterns(npatt, first_sites, make_sites, rng, n, rounds=10):
    """Simulate until `npatt` distinct site patterns exist (the first npatt are kept, weights = counts): make_sites(k)
    returns k more simulated sites [ns, k].  Discrete-gamma and low-omega classes yield many constant sites, so the
    number of sites needed is not known in advance."""
    z = None; w = None
    k = first_sites
    for _ in range(rounds):
        zs = make_sites(k)
        if z is not None:
            zs = np.concatenate([np.repeat(z, w.astype(np.int64), axis=1), zs], axis=1)
        z, w = _patterns(zs, npatt, rng, n)
        if z.shape[1] >= npatt:
            break
        got = max(1, z.shape[1])
        k = int((npatt - got) * max(1.3, 1.15 * zs.shape[1] / got)) + 64       # the shortfall at the yield seen so far
    return z, w


def codon_m0_problem(ns=8, npatt=100_000, seed=13147, kappa=5.0, omega=0.3, newick=None, bl=(0.05, 0.3), sim_factor=2.2):
    """BASELINE config 2: codeml M0 (seqtype=1, model=0, NSsites=0, CodonFreq=2), 61 states."""
    rng = np.random.default_rng(seed)
    if newick is None and ns == 8:
        b = lambda: rng.uniform(*bl)
        newick = "(((1:%f,2:%f):%f,(3:%f,4:%f):%f):%f,(5:%f,6:%f):%f,(7:%f,8:%f):%f);" % tuple(b() for _ in range(13))
    tree = parse_newick(newick) if newick else random_tree(ns, seed, bl[0], bl[1])
    fb = np.array([[0.25, 0.30, 0.20, 0.25], [0.30, 0.20, 0.30, 0.20], [0.20, 0.30, 0.25, 0.25]])
    pi = models.f3x4(fb)
    Root, U, V = models.codon_m0(kappa, omega, pi)
    cache = {}
    def P(c):
        if c not in cache:
            cache[c] = models.pmat_uvroot(tree.branch[c], Root, U, V)
        return cache[c]
    z = np.zeros((tree.ns, 0), np.uint8); w = np.zeros(0)
    nsites = int(npatt * sim_factor)
    for _ in range(8):
        zs = evolve(tree, P, pi, nsites, rng)
        z, w = _patterns(np.concatenate([np.repeat(z, w.astype(int), axis=1), zs], axis=1) if z.shape[1] else zs, npatt, rng, 61)
        if z.shape[1] >= npatt:
            break
    p = Problem(program=1, n=61, npatt=z.shape[1], z=z, fpatt=w, pi=pi, semantics=SEM_LFUN, nclass=1,
                freqK=np.ones(1), set_kind=KIND_UVROOT, nr=61, nset=1)
    p.set_data = models.uvroot_set(Root, U, V).reshape(1, -1)
    p.set_stride = p.set_data.shape[1]
    _finish(p, tree, 15)
    p.branch_set = np.zeros((1, tree.nnode), np.int32); p.branch_set[0, tree.root] = -1
    p.branch_t = tree.branch.reshape(1, -1).copy()
    p.x = np.array([kappa, omega])
    p.seqtype = 1
    return p


def codon_sites_problem(ns=32, npatt=50_000, seed=13149, kappa=2.0, freqs=(0.6, 0.3, 0.1), omegas=(0.1, 1.0, 2.5), bl=(0.02, 0.2)):
    """BASELINE config 3: codeml NSsites models (M1a/M2a/M3...): one spectral set per site class,
    all scaled by the shared Qfactor (src/codeml.c:2585-2606), lfundG semantics."""
    rng = np.random.default_rng(seed)
    tree = random_tree(ns, seed, bl[0], bl[1])
    fb = np.array([[0.25, 0.30, 0.20, 0.25], [0.30, 0.20, 0.30, 0.20], [0.20, 0.30, 0.25, 0.25]])
    pi = models.f3x4(fb)
    K = len(freqs)
    freqs = np.asarray(freqs, float); freqs = freqs / freqs.sum()
    sets, unscaled = [], []
    for w_ in omegas:
        Root, U, V = models.codon_m0(kappa, w_, pi)     # Root scaled to mean rate 1 for this class alone
        unscaled.append((Root, U, V))
    # shared scale: mean rate over classes = 1.  Recover each class's own mean rate from its scaling.
    mrs = []
    for w_ in omegas:
        R1, _, _ = models.codon_m0(kappa, w_, pi)
        mrs.append(1.0)                                  # placeholder, refined below
    # recompute unscaled roots to get true per-class mean rates
    from .models import sense_codons
    Qf = 0.0
    raw = []
    for w_ in omegas:
        Root, U, V, mr = _codon_raw(kappa, w_, pi)
        raw.append((Root, U, V, mr))
    Qf = sum(f * r[3] for f, r in zip(freqs, raw))
    for Root, U, V, mr in raw:
        sets.append(models.uvroot_set(Root / Qf, U, V))
    def make_sites(nsites):
        site_cls = rng.choice(K, size=nsites, p=freqs)
        zs = np.zeros((tree.ns, site_cls.size), np.uint8)
        for k in range(K):
            idx = np.nonzero(site_cls == k)[0]
            Root, U, V, mr = raw[k]
            zs[:, idx] = evolve(tree, lambda c: models.pmat_uvroot(tree.branch[c], Root / Qf, U, V), pi, idx.size, rng)
        return zs
    z, w = _grow_patterns(npatt, int(npatt * 1.6), make_sites, rng, 61)
    p = Problem(program=1, n=61, npatt=z.shape[1], z=z, fpatt=w, pi=pi, semantics=SEM_LFUNDG, nclass=K,
                freqK=freqs, set_kind=KIND_UVROOT, nr=61, nset=K)
    p.set_data = np.stack(sets); p.set_stride = p.set_data.shape[1]
    _finish(p, tree, 15)
    p.branch_set = np.repeat(np.arange(K, dtype=np.int32)[:, None], tree.nnode, axis=1); p.branch_set[:, tree.root] = -1
    p.branch_t = np.repeat(tree.branch.reshape(1, -1), K, axis=0)
    p.seqtype = 1
    return p


def _codon_raw(kappa, omega, pi):
    """Unscaled spectral form and mean rate of one codon class."""
    Root, U, V = models.codon_m0(kappa, omega, pi)       # Root already divided by mr
    # mean rate: rebuild Q's diagonal cheaply from the scaled spectral form is not possible; recompute
    from .models import sense_codons, _STD_CODE
    from61, _ = sense_codons()
    n = 61
    mr = 0.0
    for i in range(n):
        a = from61[i]; fa = (a // 16, (a // 4) % 4, a % 4)
        for j in range(n):
            if i == j: continue
            b = from61[j]; fb = (b // 16, (b // 4) % 4, b % 4)
            d = [k for k in range(3) if fa[k] != fb[k]]
            if len(d) != 1: continue
            q = kappa if fa[d[0]] + fb[d[0]] in (1, 5) else 1.0
            if _STD_CODE[a] != _STD_CODE[b]: q *= omega
            mr += pi[i] * q * pi[j]
    return Root * mr, U, V, mr


def wag_model():
    """The WAG model (dat/wag.dat) in spectral form, exactly as the reference's eigenQaa left it: read from the golden fixture
    tests/golden/stewart_wag_g4.pbfx (dumped from the unmodified reference running aaml with aaRatefile = wag.dat, model = 2,
    i.e. WAG's own frequencies).  Returns None when the fixture is not there."""
    import os
    from .fixture import load_fixture
    f = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "stewart_wag_g4.pbfx")
    if not os.path.exists(f):
        return None
    q = load_fixture(f)
    n = 20
    sd = np.asarray(q.set_data).reshape(-1)
    return sd[:n].copy(), sd[n:n + n * n].reshape(n, n).copy(), sd[n + n * n:n + 2 * n * n].reshape(n, n).copy(), np.asarray(q.pi[:n], float).copy()


def aa_gamma_problem(ns=64, npatt=200_000, seed=13151, alpha=0.5, K=4, bl=(0.02, 0.2), wag=False):
    """BASELINE config 4 shape: aaml empirical-matrix model + discrete gamma (K classes share one set, t*rK).
    wag=True: the WAG matrix of dat/wag.dat (see wag_model); otherwise a seeded random reversible 20-state model."""
    rng = np.random.default_rng(seed)
    tree = random_tree(ns, seed, bl[0], bl[1])
    m = wag_model() if wag else None
    Root, U, V, pi = m if m is not None else models.synthetic_aa_model()
    freqK, rK = models.discrete_gamma(alpha, K)
    def make_sites(nsites):
        site_cls = rng.integers(0, K, size=nsites)
        zs = np.zeros((tree.ns, site_cls.size), np.uint8)
        for k in range(K):
            idx = np.nonzero(site_cls == k)[0]
            zs[:, idx] = evolve(tree, lambda c: models.pmat_uvroot(tree.branch[c] * rK[k], Root, U, V), pi, idx.size, rng)
        return zs
    z, w = _grow_patterns(npatt, int(npatt * 1.02) + 64, make_sites, rng, 20)
    p = Problem(program=1, n=20, npatt=z.shape[1], z=z, fpatt=w, pi=pi, semantics=SEM_LFUNDG, nclass=K,
                freqK=freqK, set_kind=KIND_UVROOT, nr=20, nset=1)
    p.set_data = models.uvroot_set(Root, U, V).reshape(1, -1); p.set_stride = p.set_data.shape[1]
    _finish(p, tree, 50)
    p.branch_set = np.zeros((K, tree.nnode), np.int32); p.branch_set[:, tree.root] = -1
    p.branch_t = rK[:, None] * tree.branch[None, :]
    p.seqtype = 2
    p.model_name = "WAG (dat/wag.dat)" if m is not None else "seeded random reversible 20-state model"
    return p


def nuc_gtr_gamma_problem(ns=128, npatt=1_000_000, seed=1234567, alpha=0.5, K=8, bl=(0.01, 0.15),
                          rates=(0.9, 0.3, 0.2, 0.4, 0.25), pi=(0.25318, 0.32894, 0.31196, 0.10592)):
    """BASELINE config 5 shape: baseml REV(GTR) + discrete gamma, Cijk spectral form."""
    rng = np.random.default_rng(seed)
    tree = random_tree(ns, seed, bl[0], bl[1])
    pi = np.asarray(pi, float); pi = pi / pi.sum()
    cset, nr = models.gtr_cijk(rates, pi)
    Root = cset[:4]; Cijk = cset[4:].reshape(4, 4, 4)
    freqK, rK = models.discrete_gamma(alpha, K)
    def P(t):
        M = (Cijk * np.expm1(t * Root)[None, None, :]).sum(-1) + np.eye(4)
        return np.maximum(M, 0)
    def make_sites(nsites):
        site_cls = rng.integers(0, K, size=nsites)
        zs = np.zeros((tree.ns, site_cls.size), np.uint8)
        for k in range(K):
            idx = np.nonzero(site_cls == k)[0]
            zs[:, idx] = evolve(tree, lambda c: P(tree.branch[c] * rK[k]), pi, idx.size, rng)
        return zs
    z, w = _grow_patterns(npatt, int(npatt * 1.36) + 64, make_sites, rng, 4)
    p = Problem(program=0, n=4, npatt=z.shape[1], z=z, fpatt=w, pi=pi, semantics=SEM_LFUNDG, nclass=K,
                freqK=freqK, set_kind=KIND_CIJK, nr=nr, nset=1)
    p.set_data = cset.reshape(1, -1); p.set_stride = p.set_data.shape[1]
    _finish(p, tree, 100)
    p.branch_set = np.zeros((K, tree.nnode), np.int32); p.branch_set[:, tree.root] = -1
    p.branch_t = rK[:, None] * tree.branch[None, :]
    p.seqtype = 0
    return p


def add_ambiguity(p, frac, seed=3):
    """Turn a clean problem into an 'unclean' one: a fraction of tip states becomes ambiguity codes
    (code n = fully missing, code n+1 = a two-state ambiguity), exercising nChara/CharaMap."""
    rng = np.random.default_rng(seed)
    n = p.n
    q = Problem(**{k: getattr(p, k) for k in p.__dataclass_fields__})
    q.z = p.z.copy()
    q.nchara = p.nchara.copy(); q.charamap = p.charamap.copy()
    q.nchara[n] = n; q.charamap[n, :n] = np.arange(n)
    q.nchara[n + 1] = 2; q.charamap[n + 1, :2] = [0, n - 1]
    mask = rng.random(q.z.shape) < frac
    codes = np.where(rng.random(q.z.shape) < 0.5, n, n + 1).astype(np.uint8)
    q.z[mask] = codes[mask]
    q.cleandata = 0
    return q
