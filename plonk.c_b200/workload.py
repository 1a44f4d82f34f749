"""Synthetic workloads for the plonk-test circuit (host side, numpy only -- no device work here).

Everything is a pure function of (seed, item index), through a counter-based generator, so the
CPU reference arm, the oracle and the CUDA path are fed byte-identical inputs without sharing state
(SURVEY.md section 8(d)).

Byte layouts (shared by the C-ABI in include/plonk_b200.h and by both CPU checkers):
  circuit  44 B   q_l[4] q_r[4] q_o[4] q_m[4] q_c[4] | c_a.type[4] c_a.index[4] | c_b.. | c_c..
  witness  12 B   a[4] b[4] c[4]                      (ASSIGNMENTS, constraints.h:57-62)
  rand      9 B   b1..b9                              (plonk_prove's HF rand[9], plonk.h:228)
  chal      5 B   alpha beta gamma z v                (CHALLENGE, plonk.h:16-22)
  proof    34 B   9 x G1{x,y,infinite} + 7 x HF       (PROOF, plonk.h:24-41)
  srs      g1s[len][3], g2[4] = g2_1.x g2_1.y g2_s.x g2_s.y   (SRS, srs.h:11-16)
"""
import numpy as np

P17, P101 = 17, 101

# The plonk-test circuit, verbatim from the reference's only end-to-end test
# (plonk-test.c:157-213): three x*x gates and one x^2+y^2=z^2 gate.
COPYOF_A, COPYOF_B, COPYOF_C = 0, 1, 2
PLONK_TEST_CIRCUIT = np.array(
    [0, 0, 0, 1,            # q_l
     0, 0, 0, 1,            # q_r
     16, 16, 16, 16,        # q_o = -1
     1, 1, 1, 0,            # q_m
     0, 0, 0, 0,            # q_c
     COPYOF_B, COPYOF_B, COPYOF_B, COPYOF_C, 1, 2, 3, 1,    # c_a = b1 b2 b3 c1
     COPYOF_A, COPYOF_A, COPYOF_A, COPYOF_C, 1, 2, 3, 2,    # c_b = a1 a2 a3 c2
     COPYOF_A, COPYOF_B, COPYOF_C, COPYOF_C, 4, 4, 4, 3],   # c_c = a4 b4 c4 c3
    dtype=np.uint8)

# The shipped test vector (plonk-test.c:229-267).
GOLDEN_WITNESS = np.array([[3, 4, 5, 9, 3, 4, 5, 16, 9, 16, 8, 8]], dtype=np.uint8)
GOLDEN_RAND = np.array([[7, 4, 11, 12, 16, 2, 14, 11, 7]], dtype=np.uint8)
GOLDEN_CHALLENGE = np.array([[15, 12, 13, 5, 12]], dtype=np.uint8)
GOLDEN_U = np.array([4], dtype=np.uint8)     # the verifier's extra challenge

G2_GENERATOR = (36, 31)   # g2.h:19-21
G2_TIMES_2 = (90, 82)     # g2_mul(H, 2), pinned by g2-test.c:17


def _g1_add(p, q):
    """Affine addition on y^2 = x^3 + 3 over F_101; None is the identity.  Host-side helper used only
    to lay out the generator SRS (a handful of points), never on a hot path."""
    if p is None:
        return q
    if q is None:
        return p
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        if (y1 + y2) % P101 == 0:
            return None
        m = 3 * x1 * x1 * pow(2 * y1, P101 - 2, P101) % P101
    else:
        m = (y2 - y1) * pow(x2 - x1, P101 - 2, P101) % P101
    x3 = (m * m - x1 - x2) % P101
    return x3, (m * (x1 - x3) - y1) % P101


def g1_multiple(k, base=(1, 2)):
    acc = None
    for _ in range(k):
        acc = _g1_add(acc, base)
    return acc


def _enc(p):
    return [0, 0, 1] if p is None else [p[0], p[1], 0]


def identity_srs(n, secret=2):
    """What srs_create(secret, n) really produces: every g1s[i] is {0,0,infinite} because the reference
    multiplies the identity, not the generator (srs.h:27-35, pinned by srs-test.c:15-17)."""
    assert secret == 2, "g2_s is tabulated for the test secret only"
    g1s = np.tile(np.array([0, 0, 1], np.uint8), (n + 1, 1))
    return g1s, np.array([*G2_GENERATOR, *G2_TIMES_2], np.uint8)


def generator_srs(n, secret=2):
    """SURVEY.md section 8(d) mode (ii): g1s[i] = g1_mul(G, secret^i mod 101), g2_s = g2_mul(H, secret),
    handed over through the public SRS struct.  (For i >= 7 the exponent 2^i mod 101 is no longer
    2^i mod 17, so entries 7.. are not powers of one secret in the order-17 group; kept as specified.)"""
    assert secret == 2
    g1s = np.array([_enc(g1_multiple(pow(secret, i, P101) % 17)) for i in range(n + 1)], np.uint8)
    return g1s, np.array([*G2_GENERATOR, *G2_TIMES_2], np.uint8)


def satisfying_witnesses():
    """The 289 (x, y, z) in F_17^3 with x^2 + y^2 = z^2, lexicographic; row 71 is (3, 4, 5).
    Returns [289][12] witness rows a=[x,y,z,x^2] b=[x,y,z,y^2] c=[x^2,y^2,z^2,z^2]."""
    rows = []
    for x in range(17):
        for y in range(17):
            for z in range(17):
                if (x * x + y * y - z * z) % 17 == 0:
                    xx, yy, zz = x * x % 17, y * y % 17, z * z % 17
                    rows.append([x, y, z, xx, x, y, z, yy, xx, yy, zz, zz])
    return np.array(rows, dtype=np.uint8)


_WITNESS_TABLE = None
DRAWS_PER_ITEM = 16


def splitmix64(x):
    """Vectorised splitmix64 finaliser on uint64 counters."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def make_batch(seed, start, count, variant="U17"):
    """Items [start, start+count) of the synthetic stream `seed`.

    Draw j of item i is splitmix64(seed + i*16 + j): j=0 picks the witness row (mod 289), j=1..9 the
    blinding scalars, j=10..14 the challenges, j=15 the verifier's u.  Variant "U17" is uniform on
    [0,17); "NZ" is uniform on [1,17) for blinding and challenges.
    Returns (witness[count][12], rand[count][9], chal[count][5], u[count]) as uint8.
    """
    global _WITNESS_TABLE
    if _WITNESS_TABLE is None:
        _WITNESS_TABLE = satisfying_witnesses()
    idx = np.arange(start, start + count, dtype=np.uint64)
    with np.errstate(over="ignore"):
        base = np.uint64(seed) + idx * np.uint64(DRAWS_PER_ITEM)
        draws = splitmix64(base[:, None] + np.arange(DRAWS_PER_ITEM, dtype=np.uint64)[None, :])
    wit = _WITNESS_TABLE[(draws[:, 0] % np.uint64(289)).astype(np.int64)]
    if variant == "U17":
        sc = (draws[:, 1:] % np.uint64(17)).astype(np.uint8)
    elif variant == "NZ":
        sc = (draws[:, 1:] % np.uint64(16)).astype(np.uint8) + np.uint8(1)
    else:
        raise ValueError(variant)
    return (np.ascontiguousarray(wit), np.ascontiguousarray(sc[:, 0:9]),
            np.ascontiguousarray(sc[:, 9:14]), np.ascontiguousarray(sc[:, 14]))


def make_poly_items(seed, start, count, top_nonzero=False):
    """Config-2 items: A[6], B[6] coefficients, x, vals[4], all uniform on [0,17)
    (optionally with the top coefficients forced non-zero to isolate the trimming path)."""
    idx = np.arange(start, start + count, dtype=np.uint64)
    with np.errstate(over="ignore"):
        base = np.uint64(seed) + idx * np.uint64(32)
        d = splitmix64(base[:, None] + np.arange(17, dtype=np.uint64)[None, :])
    v = (d % np.uint64(17)).astype(np.uint8)
    a, b, x, vals = v[:, 0:6].copy(), v[:, 6:12].copy(), v[:, 12].copy(), v[:, 13:17].copy()
    if top_nonzero:
        a[:, 5] = (d[:, 5] % np.uint64(16)).astype(np.uint8) + 1
        b[:, 5] = (d[:, 11] % np.uint64(16)).astype(np.uint8) + 1
    return a, b, x, vals


def make_group_items(seed, start, count):
    """Config-3/4 items: P = a*G, Q = b*H with a, b uniform on [1,17) as *indices*; the caller maps
    them to points with the implementation under test.  Returns (a[count], b[count], s[count]) with s
    a scalar uniform on [0,17)."""
    idx = np.arange(start, start + count, dtype=np.uint64)
    with np.errstate(over="ignore"):
        base = np.uint64(seed) + idx * np.uint64(4)
        d = splitmix64(base[:, None] + np.arange(3, dtype=np.uint64)[None, :])
    a = (d[:, 0] % np.uint64(16)).astype(np.uint8) + 1
    b = (d[:, 1] % np.uint64(16)).astype(np.uint8) + 1
    s = (d[:, 2] % np.uint64(17)).astype(np.uint8)
    return a, b, s


def g1_subgroup_table():
    """[17][3]: k*G for k = 0..16 (row 0 is the identity {0,0,1})."""
    return np.array([_enc(g1_multiple(k)) for k in range(17)], np.uint8)
