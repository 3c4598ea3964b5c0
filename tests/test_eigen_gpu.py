"""Device eigen-solver (pb200_eigen_qrev, csrc/eigen.cuh -- SURVEY.md 8f rank 1) against the reference's own eigenQREV.

The checker is the reference itself: every UVROOT fixture under tests/golden/ carries (Root, U, V) exactly as the
reference's eigenQcodon / eigenQaa left them at the example's MLE.  Q = U diag(Root) V rebuilds the rate matrix they
decompose; the device solution of that Q (with the fixture's pi) must give the same spectrum and the same P(t) -- the
factors themselves are unique only up to the sign of each eigenvector, so they are compared through P(t), U V = I and
Q = U diag(Root) V.  Dropped states (pi = 0) and batches are checked against a numpy restatement (test code)."""
import time

import numpy as np
import pytest

from conftest import GOLDEN, golden_ids
from paml_b200 import models
from paml_b200.fixture import load_fixture
from paml_b200.likelihood import TreeLikelihood

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tl():
    t = TreeLikelihood(0)
    yield t
    t.close()


def _pt(Root, U, V, t):
    return (U * np.exp(Root * t)[None, :]) @ V


def _check(Q, pi, Root, U, V, tol=5e-13):
    n = len(pi)
    scale = max(1.0, np.abs(Q).max())
    keep = pi > 1e-100
    assert np.abs(U @ V - np.eye(n)).max() < tol
    assert np.abs((U * Root[None, :]) @ V - np.where(np.outer(keep, keep), Q, 0)).max() < tol * scale
    r = Root[keep]
    assert (np.diff(r) <= 0).all(), "eigenvalues of the kept states must come in decreasing order"
    assert (Root[~keep] == 0).all()


@pytest.mark.parametrize("path", GOLDEN, ids=golden_ids())
def test_fixture_sets(tl, path):
    p = load_fixture(path)
    if p.set_kind != 0:
        pytest.skip("not a UVROOT fixture")
    n = p.n
    pi = np.asarray(p.pi[:n], np.float64)
    Qs, refs = [], []
    for iset in range(p.nset):
        s = p.set_data[iset]
        Root, U, V = s[:n], s[n:n + n * n].reshape(n, n), s[n + n * n:n + 2 * n * n].reshape(n, n)
        if not np.any(U):
            continue                                   # unused _UU[] slot
        Qs.append((U * Root[None, :]) @ V)
        refs.append((Root, U, V))
    R2, U2, V2, sw = tl.eigen_qrev(np.array(Qs), np.tile(pi, (len(Qs), 1)))
    assert (sw > 0).all() and (sw <= 12).all(), sw
    for k, (Root, U, V) in enumerate(refs):
        _check(Qs[k], pi, R2[k], U2[k], V2[k], tol=2e-12)
        np.testing.assert_allclose(np.sort(R2[k]), np.sort(Root), rtol=0, atol=1e-11 * max(1.0, np.abs(Root).max()))
        for t in (1e-3, 0.05, 0.4, 3.0):
            np.testing.assert_allclose(_pt(R2[k], U2[k], V2[k], t), _pt(Root, U, V, t), rtol=0, atol=2e-13)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 20, 61, 63, 64])
def test_sizes_random_reversible(tl, n):
    rng = np.random.default_rng(n)
    njob = 5
    Qs, pis = [], []
    for _ in range(njob):
        S = rng.random((n, n)); S = S + S.T
        pi = rng.random(n) + 0.05; pi /= pi.sum()
        Q = S * pi[None, :]
        np.fill_diagonal(Q, 0); np.fill_diagonal(Q, -Q.sum(1))
        Qs.append(Q); pis.append(pi)
    R, U, V, sw = tl.eigen_qrev(np.array(Qs), np.array(pis))
    for k in range(njob):
        _check(Qs[k], pis[k], R[k], U[k], V[k])
        Rn, Un, Vn = models.eigen_reversible(Qs[k], pis[k])
        np.testing.assert_allclose(np.sort(R[k]), np.sort(Rn), rtol=0, atol=1e-12 * np.abs(Rn).max())
        np.testing.assert_allclose(_pt(R[k], U[k], V[k], 0.3), _pt(Rn, Un, Vn, 0.3), rtol=0, atol=1e-13)
        assert abs(R[k].max()) < 1e-12 * np.abs(R[k]).max()        # the zero root of a rate matrix


def test_dropped_states(tl):
    """pi = 0 for some states (codon frequencies with empty codons): the reference solves the reduced problem and puts
    identity rows/columns back (src/tools.c:5055-5102)."""
    rng = np.random.default_rng(3)
    n = 61
    S = rng.random((n, n)); S = S + S.T
    pi = rng.random(n); pi[[0, 7, 33, 60]] = 0; pi /= pi.sum()
    Q = S * pi[None, :]
    np.fill_diagonal(Q, 0); np.fill_diagonal(Q, -Q.sum(1))
    R, U, V, sw = tl.eigen_qrev(Q, pi)
    _check(Q, pi, R[0], U[0], V[0])
    keep = pi > 0
    for j in np.where(~keep)[0]:
        e = np.zeros(n); e[j] = 1
        assert (U[0][j] == e).all() and (U[0][:, j] == e).all() and (V[0][j] == e).all() and (V[0][:, j] == e).all()
    Qk = Q[np.ix_(keep, keep)]
    Rn, Un, Vn = models.eigen_reversible(Qk, pi[keep])
    np.testing.assert_allclose(_pt(R[0], U[0], V[0], 0.2)[np.ix_(keep, keep)], _pt(Rn, Un, Vn, 0.2), rtol=0, atol=1e-13)


def test_timing_report(tl):
    """Not a gate: prints what one codon-sized solve costs through the C ABI (host buffers in and out)."""
    p = load_fixture([g for g in GOLDEN if g.endswith("mhc_m0.pbfx")][0])
    n = p.n
    s = p.set_data[0]
    Root, U, V = s[:n], s[n:n + n * n].reshape(n, n), s[n + n * n:n + 2 * n * n].reshape(n, n)
    Q = (U * Root[None, :]) @ V
    pi = np.asarray(p.pi[:n], np.float64)
    for njob in (1, 8, 64):
        Qb, pb = np.tile(Q, (njob, 1, 1)), np.tile(pi, (njob, 1))
        tl.eigen_qrev(Qb, pb)
        t0 = time.perf_counter()
        for _ in range(20):
            sw = tl.eigen_qrev(Qb, pb)[3]
        dt = (time.perf_counter() - t0) / 20
        t0 = time.perf_counter()
        for _ in range(20 if njob == 1 else 2):
            for k in range(njob):
                models.eigen_reversible(Q, pi)
        dn = (time.perf_counter() - t0) / (20 if njob == 1 else 2)
        print("\neigen n=%d njob=%d: device %.3f ms per call (%.3f ms per matrix, %d sweeps); numpy eigh on the host %.3f ms"
              % (n, njob, dt * 1e3, dt * 1e3 / njob, sw[0], dn * 1e3))


def test_warm_start_follows_the_optimiser(tl):
    """The Jacobi iteration starts from the eigenvectors of the previous call's solution (job for job): ming2 perturbs one
    parameter at a time, so the next matrix is a small perturbation of the last one and two sweeps (one, for a finite-difference
    step) replace seven or eight.  The result must not depend on where the iteration started: every solution of a 200-step chain of
    perturbed matrices is checked like a cold one (U V = I, Q = U diag(Root) V, spectrum against numpy), i.e. the basis handed down
    from call to call does not drift (every 32nd call starts cold)."""
    rng = np.random.default_rng(11)
    n = 61
    S = rng.random((n, n)); S = S + S.T
    pi = rng.random(n) + 0.05; pi /= pi.sum()

    def make(S):
        Q = S * pi[None, :]
        np.fill_diagonal(Q, 0); np.fill_diagonal(Q, -Q.sum(1))
        return Q

    R, U, V, sw = tl.eigen_qrev(make(S), pi)
    warm_sweeps, t_warm = [], 0.0
    for step in range(200):
        rel = 1e-7 if step % 3 else 3e-2                # finite-difference steps and line-search steps
        S = S * (1 + rel * rng.standard_normal((n, n)))
        S = (S + S.T) / 2
        Q = make(S)
        t0 = time.perf_counter()
        R, U, V, sw = tl.eigen_qrev(Q, pi)
        t_warm += time.perf_counter() - t0
        assert sw[0] > 0
        warm_sweeps.append(int(sw[0]))
        _check(Q, pi, R[0], U[0], V[0])
        if step % 20 == 0:
            Rn, Un, Vn = models.eigen_reversible(Q, pi)
            np.testing.assert_allclose(np.sort(R[0]), np.sort(Rn), rtol=0, atol=1e-12 * np.abs(Rn).max())
            np.testing.assert_allclose(_pt(R[0], U[0], V[0], 0.3), _pt(Rn, Un, Vn, 0.3), rtol=0, atol=1e-13)
    ws = np.array(warm_sweeps)
    small = np.array([s for i, s in enumerate(warm_sweeps) if i % 3 and (i + 1) % 32])
    print("\nwarm-started eigen_qrev, n = 61: %.3f ms per call on average; sweeps: median %d, after a 1e-7 perturbation %.2f on average, max %d"
          % (t_warm / 200 * 1e3, np.median(ws), small.mean(), ws.max()))
    assert small.mean() <= 2.5, small
