"""Shared input generators for the parity tests (CPU hostcheck tests and GPU tests use the same cases)."""
import numpy as np


def g1_cases(n, seed=5):
    """Arbitrary G1 byte triples: random coordinates (mostly off-curve), ~10% infinite flags with garbage
    coordinates, x-collisions with equal / negated / unrelated y, and y = 0 points."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 101, (n, 3), dtype=np.uint8)
    b = rng.integers(0, 101, (n, 3), dtype=np.uint8)
    a[:, 2] = rng.random(n) < 0.1
    b[:, 2] = rng.random(n) < 0.1
    q = n // 4
    b[:q, :2] = a[:q, :2]
    b[q:n // 3, 0] = a[q:n // 3, 0]
    b[q:n // 3, 1] = (101 - a[q:n // 3, 1]) % 101
    b[n // 3:n // 2, 0] = a[n // 3:n // 2, 0]
    a[n // 2:n // 2 + n // 50, 1] = 0
    return a, b


def scalars_u64(n, seed=6):
    rng = np.random.default_rng(seed)
    s = rng.integers(0, 1 << 63, n, dtype=np.uint64)
    s[: n // 2] %= 40
    s[:17] = np.arange(17)
    return s


def poly_cases(n, la, lb, seed=7):
    rng = np.random.default_rng(seed + 31 * la + lb)
    a = rng.integers(0, 17, (n, la), dtype=np.uint8)
    b = rng.integers(0, 17, (n, lb), dtype=np.uint8)
    a[rng.random((n, la)) < 0.3] = 0
    b[rng.random((n, lb)) < 0.3] = 0
    al = rng.integers(1, la + 1, n).astype(np.uint8)
    bl = rng.integers(1, lb + 1, n).astype(np.uint8)
    a[: n // 20] = 0        # zero polynomials
    b[n // 20: n // 10] = 0
    return a, al, b, bl


def subgroup_points(W, oracle, n, seed=3):
    """P = a*G, Q = b*H with a, b in [1,17): the config-4 inputs."""
    ai, bi, s = W.make_group_items(seed, 0, n)
    P = W.g1_subgroup_table()[ai]
    Q = oracle.g2_mul(np.tile(np.array([[36, 31]], np.uint8), (n, 1)), bi.astype(np.uint64))
    return np.ascontiguousarray(P), np.ascontiguousarray(Q), s


SRS_MODES = {
    "identity6": lambda W: W.identity_srs(6),      # exactly what the reference's test builds (plonk-test.c:125-129)
    "generator6": lambda W: W.generator_srs(6),
    "identity9": lambda W: W.identity_srs(9),      # the benchmark size (no item dies on the SRS guard)
    "generator9": lambda W: W.generator_srs(9),
    "generator4": lambda W: W.generator_srs(4),    # every item dies on the first SRS guard
}


def garbage_srs(n=10, seed=12):
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 101, (n, 3), dtype=np.uint8)
    g[:, 2] = rng.random(n) < 0.3
    return g, np.array([36, 31, 90, 82], np.uint8)
