"""The host model's linear solvers against a plain restatement of util.rs::linsolve (libaom mathutils.h linsolve, the
routine behind every AR and noise-strength solve of the reference): bit for bit, including the pivot bubbling, the
failure returns, and the tridiagonal fast path the strength systems take."""
import ctypes as C

import numpy as np
import pytest

from grav1synth_b200 import diff as D

TINY = 1.0e-16


def linsolve_ref(n, A, b, x):
    """util.rs::linsolve, statement by statement (numpy float64 scalars round like the reference's f64)."""
    A, b, x = A.copy(), b.copy(), x.copy()
    for k in range(n - 1):
        for i in range(n - 1, k, -1):
            if abs(A[i - 1, k]) < abs(A[i, k]):
                A[[i - 1, i]] = A[[i, i - 1]]
                b[[i - 1, i]] = b[[i, i - 1]]
        for i in range(k, n - 1):
            if abs(A[k, k]) < TINY:
                return False, x
            c = A[i + 1, k] / A[k, k]
            for j in range(n):
                A[i + 1, j] = A[i + 1, j] - c * A[k, j]
            b[i + 1] = b[i + 1] - c * b[k]
    for i in range(n - 1, -1, -1):
        if abs(A[i, i]) < TINY:
            return False, x
        c = np.float64(0)
        for j in range(i + 1, n):
            c = c + A[i, j] * x[j]
        x[i] = (b[i] - c) / A[i, i]
    return True, x


def probe(which, A, b, x0):
    L = D.lib()
    L.g1s_linsolve_probe.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.g1s_linsolve_probe.restype = C.c_int
    A = np.ascontiguousarray(A, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    x = np.array(x0, np.float64)
    rc = L.g1s_linsolve_probe(which, len(b), A.ctypes.data, b.ctypes.data, x.ctypes.data)
    return rc, x


def same_bits(a, b):
    return np.array_equal(np.asarray(a, np.float64).view(np.uint64), np.asarray(b, np.float64).view(np.uint64))


@pytest.mark.parametrize("n", [1, 2, 5, 20, 24, 25, 31, 40])
def test_dense_elimination_matches_the_reference_statement_by_statement(n):
    rng = np.random.default_rng(n)
    for trial in range(12):
        A = rng.normal(size=(n, n))
        if trial % 3 == 0:   # symmetric positive definite, as the AR systems are
            A = A @ A.T + n * np.eye(n)
        elif trial % 3 == 1:  # rows that make the bubble pass work
            A[rng.permutation(n)[: max(1, n // 2)]] *= 1e3
        b = rng.normal(size=n)
        x0 = rng.normal(size=n) if trial % 2 else np.zeros(n)
        ok, want = linsolve_ref(n, A, b, x0)
        rc, got = probe(0, A, b, x0)
        assert (rc == 1) == ok
        assert same_bits(got, want)


def test_dense_elimination_failures_leave_x_as_the_reference_does():
    n = 6
    A = np.eye(n)
    A[3, 3] = 0.0          # fails in the back substitution at row 3: rows 5 and 4 are written, the rest is not
    b = np.arange(1.0, n + 1)
    x0 = np.full(n, 7.0)
    ok, want = linsolve_ref(n, A, b, x0)
    rc, got = probe(0, A, b, x0)
    assert not ok and rc == -1 and same_bits(got, want)
    A = np.zeros((n, n))   # fails at the first pivot: x untouched
    ok, want = linsolve_ref(n, A, b, x0)
    rc, got = probe(0, A, b, x0)
    assert not ok and rc == -1 and same_bits(got, x0) and same_bits(want, x0)


def strength_like(rng, n=20, empty_bins=()):
    """A noise strength system as NoiseStrengthSolver::solve builds it: tridiagonal sums + the regulariser."""
    A = np.zeros((n, n))
    b = np.zeros(n)
    m = 0
    for _ in range(int(rng.integers(50, 4000))):
        bin_ = rng.uniform(0, n - 1)
        i0 = int(bin_)
        if i0 in empty_bins:
            continue
        i1 = min(n - 1, i0 + 1)
        a = bin_ - i0
        s = rng.uniform(0.5, 6.0)
        A[i0, i0] += (1 - a) * (1 - a)
        A[i1, i0] += a * (1 - a)
        A[i1, i1] += a * a
        A[i0, i1] += a * (1 - a)
        b[i0] += (1 - a) * s
        b[i1] += a * s
        m += 1
    alpha = 2.0 * m / n
    for i in range(n):
        lo, hi = max(0, i - 1), min(n - 1, i + 1)
        A[i, lo] -= alpha
        A[i, i] += 2 * alpha
        A[i, hi] -= alpha
    for i in range(n):
        A[i, i] += 1.0 / 8192.0
    return A, b + 3.0 / 8192.0


def test_tridiagonal_fast_path_is_the_dense_elimination_bit_for_bit():
    rng = np.random.default_rng(7)
    for trial in range(40):
        A, b = strength_like(rng, empty_bins=(3, 4, 5) if trial % 4 == 0 else ())
        x0 = np.zeros(20)
        ok, want = linsolve_ref(20, A, b, x0)
        rc, got = probe(1, A, b, x0)
        assert ok and rc == 1
        assert same_bits(got, want)
        rc0, got0 = probe(0, A, b, x0)
        assert rc0 == 1 and same_bits(got0, want)


def test_tridiagonal_fast_path_declines_when_a_row_swap_is_needed_and_reports_failures():
    rng = np.random.default_rng(11)
    n = 12
    A = np.diag(rng.uniform(1, 2, n)) + np.diag(rng.uniform(0.1, 0.5, n - 1), 1) + np.diag(rng.uniform(0.1, 0.5, n - 1), -1)
    b = rng.normal(size=n)
    A[5, 4] = 9.0          # larger than the pivot above it: the bubble pass swaps rows 4 and 5
    rc, _ = probe(1, A, b, np.zeros(n))
    assert rc == 0
    ok, want = linsolve_ref(n, A, b, np.zeros(n))
    rc, got = probe(0, A, b, np.zeros(n))
    assert ok and rc == 1 and same_bits(got, want)
    A = np.diag(np.ones(n))
    A[n - 3, n - 3] = 0.0  # back substitution fails at row n-3: the rows below it are written, as the reference does
    x0 = np.full(n, 5.0)
    ok, want = linsolve_ref(n, A, b, x0)
    rc, got = probe(1, A, b, x0)
    assert not ok and rc == -1 and same_bits(got, want)
