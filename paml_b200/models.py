"""Host-side model numerics that stay on the CPU: building rate matrices, their
spectral decomposition and the discrete-gamma rates.  In the drop-in these come from the reference's own C
(eigenQcodon / eigenQREV / DiscreteGamma called by SetParameters); this module is the Python mirror the
tests and bench.py use to manufacture inputs of the same form (Root, U, V / Root, Cijk).
"""
import numpy as np

BASES = "TCAG"                       # src/tools.c:15
AAS = "ARNDCQEGHILKMFPSTWYV"         # src/tools.c:19 (first 20)
# standard genetic code (icode 0) in TCAG codon order, src/tools.c:23-25
_STD_CODE = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"


def sense_codons():
    """FROM61: indices (0..63, TCAG order) of the 61 sense codons; FROM64: inverse map (-1 for stops)."""
    from61 = [i for i in range(64) if _STD_CODE[i] != "*"]
    from64 = [-1] * 64
    for k, ic in enumerate(from61):
        from64[ic] = k
    return np.array(from61), np.array(from64)


def f3x4(fb3x4):
    """Codon frequencies from 3x4 position-specific base frequencies (CodonFreq = 2)."""
    from61, _ = sense_codons()
    fb = np.asarray(fb3x4, float).reshape(3, 4)
    pi = np.array([fb[0, ic // 16] * fb[1, (ic // 4) % 4] * fb[2, ic % 4] for ic in from61])
    return pi / pi.sum()


def eigen_reversible(Q, pi):
    """Spectral form of a reversible Q (what eigenQREV returns, src/tools.c:5023-5110):
    Q = U diag(Root) V with U = pi^-1/2 R, V = R' pi^1/2, R the eigenvectors of pi^1/2 Q pi^-1/2."""
    sq = np.sqrt(pi)
    S = (sq[:, None] * Q) / sq[None, :]
    S = 0.5 * (S + S.T)
    w, R = np.linalg.eigh(S)
    order = np.argsort(-w)           # descending, zero eigenvalue first (EigenSort, src/tools.c:5136)
    w, R = w[order], R[:, order]
    U = R / sq[:, None]
    V = R.T * sq[None, :]
    return w, U, V


def codon_m0(kappa, omega, pi):
    """Goldman-Yang / M0 codon model (eigenQcodon mode 1, HKY-style kappa; src/codeml.c:3272-3321).
    Returns Root (scaled so the mean rate is 1), U, V."""
    from61, _ = sense_codons()
    n = len(from61)
    Q = np.zeros((n, n))
    for i in range(n):
        a = from61[i]
        fa = (a // 16, (a // 4) % 4, a % 4)
        for j in range(i):
            b = from61[j]
            fb = (b // 16, (b // 4) % 4, b % 4)
            diff = [k for k in range(3) if fa[k] != fb[k]]
            if len(diff) != 1:
                continue
            k = diff[0]
            q = 1.0
            if fa[k] + fb[k] in (1, 5):      # T<->C or A<->G
                q = kappa
            if _STD_CODE[a] != _STD_CODE[b]:
                q *= omega
            Q[i, j] = Q[j, i] = q
    Q = Q * pi[None, :]
    Q[np.diag_indices(n)] = -Q.sum(1)
    mr = -(pi * np.diag(Q)).sum()
    Root, U, V = eigen_reversible(Q, pi)
    return Root / mr, U, V


def uvroot_set(Root, U, V):
    """Pack one PB200_KIND_UVROOT spectral set: Root[n], U[n*n], V[n*n]."""
    return np.concatenate([Root, U.ravel(), V.ravel()])


def gtr_cijk(rates, pi):
    """REV/GTR nucleotide model in PAML's Cijk form (eigenQREVbase, src/treesub.c:2488-2540):
    rates = (a..e) for TC, TA, TG, CA, CG relative to AG = 1; state order TCAG.
    Returns one PB200_KIND_CIJK set (Root[4], Cijk[4*4*4]) and nr = 4."""
    a, b, c, d, e = rates
    R = np.array([[0, a, b, c], [a, 0, d, e], [b, d, 0, 1.0], [c, e, 1.0, 0]])
    pi = np.asarray(pi, float)
    Q = R * pi[None, :]
    Q[np.diag_indices(4)] = -Q.sum(1)
    mr = -(pi * np.diag(Q)).sum()
    Root, U, V = eigen_reversible(Q / mr, pi)
    Root[0] = 0.0
    Cijk = np.einsum("ik,kj->ijk", U, V)
    return np.concatenate([Root, Cijk.ravel()]), 4


def synthetic_aa_model(seed=7):
    """A reversible 20-state model with seeded random exchangeabilities and frequencies (stands in for an
    empirical matrix such as dat/wag.dat, which is reference data and not available on the GPU box)."""
    rng = np.random.default_rng(seed)
    S = rng.gamma(0.6, 1.0, (20, 20)) + 0.02
    S = np.triu(S, 1); S = S + S.T
    pi = rng.dirichlet(np.full(20, 8.0))
    Q = S * pi[None, :]
    Q[np.diag_indices(20)] = -Q.sum(1)
    mr = -(pi * np.diag(Q)).sum()
    Root, U, V = eigen_reversible(Q / mr, pi)
    return Root, U, V, pi


def discrete_gamma(alpha, K):
    """Mean-of-category discrete gamma rates (DiscreteGamma with UseMedian = 0, src/tools.c:2601-2627)."""
    from scipy.special import gammainc
    from scipy.stats import gamma as gdist
    cuts = gdist.ppf(np.arange(1, K) / K, a=alpha, scale=1.0 / alpha)
    lo = np.concatenate([[0.0], cuts]); hi = np.concatenate([cuts, [np.inf]])
    cdf1 = lambda x: np.where(np.isinf(x), 1.0, gammainc(alpha + 1, np.where(np.isinf(x), 0.0, x) * alpha))
    rK = (cdf1(hi) - cdf1(lo)) * K
    return np.full(K, 1.0 / K), rK


def pmat_uvroot(t, Root, U, V):
    """Reference semantics of PMatUVRoot in numpy (host-side helper for the simulator)."""
    P = (U * np.expm1(t * Root)[None, :]) @ V + np.eye(len(Root))
    P[P < 0] = 0
    return P
