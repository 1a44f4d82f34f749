"""GPU tests at BASELINE.json's full sizes, through size-independent properties (the CPU oracle cannot
finish 2^22..2^24 items in a test), plus oracle spot checks on random windows of the big batch."""
import numpy as np
import pytest

import parity_suite as ps

pytestmark = pytest.mark.gpu

N22 = 1 << 22


def _t(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_config2_poly_properties_2p22(host, oracle, W):
    """2^22 items {A[6], B[6], x, vals[4]}: eval is a ring homomorphism, division undoes multiplication,
    interpolation inverts evaluation on H."""
    import torch
    a, b, x, vals = W.make_poly_items(42, 0, N22)
    b[:, 0] |= (b.sum(1) == 0).astype(np.uint8)                       # keep B non-zero so that A*B / B is defined
    six = np.full(N22, 6, np.uint8)
    A, B, X, V, L6 = _t(a), _t(b), _t(x), _t(vals), _t(six)
    prod, plen = host.poly_mul(A, L6, B, L6)
    ea, eb, ep = host.poly_eval(A, L6, X), host.poly_eval(B, L6, X), host.poly_eval(prod, plen, X)
    assert torch.equal(host.hf_mul(ea, eb), ep), "eval(A*B) != eval(A) eval(B)"
    q, ql, r, rl, st = host.poly_divide(prod, plen, B, L6, sq=11, sr=5)
    assert int(st.sum()) == 0 and int(r.sum()) == 0, "A*B / B left a remainder"
    at, al = host.poly_add(A, L6, torch.zeros_like(A), torch.ones_like(L6))    # poly_new(A): the trimmed form
    assert torch.equal(q[:, :6], at) and torch.equal(ql, al), "(A*B)/B != A"
    ctx = host.Plonk(W.PLONK_TEST_CIRCUIT, *W.identity_srs(6))
    f, fl = ctx.interpolate_at_h(V)
    for k, h in enumerate((1, 4, 16, 13)):
        assert torch.equal(host.poly_eval(f, fl, torch.full_like(X, h)), V[:, k]), "interpolate_at_h(vals)(h_k) != vals[k]"
    # oracle spot check on a window
    lo = 3_000_000
    sl = slice(lo, lo + 4096)
    ps.eq("poly_mul window", (prod[sl].cpu().numpy(), plen[sl].cpu().numpy()), oracle.poly_binop(2, a[sl], six[sl], b[sl], six[sl], 11))
    zh = np.tile(np.array([16, 0, 0, 0, 1], np.uint8), (4096, 1))
    five = np.full(4096, 5, np.uint8)
    got = host.poly_divide(prod[sl], plen[sl], _t(zh), _t(five), sq=7, sr=4)
    ps.eq("poly_divide by Z_H window", tuple(t.cpu().numpy() for t in got),
          oracle.poly_divide(prod[sl].cpu().numpy(), plen[sl].cpu().numpy(), zh, five, 7, 4))


def test_config3_g1_mul_2p24_group_law(host, oracle, W):
    """2^24 scalar multiplications against the test SRS: s*P + t*P = ((s+t) mod 17)*P on the order-17 subgroup,
    and the results agree with the oracle on a window."""
    import torch
    n = 1 << 24
    g1s, _ = W.generator_srs(9)
    ai, bi, s = W.make_group_items(8, 0, n)
    P = g1s[ai % 10]
    t = ((s.astype(np.int64) * 7 + 3) % 17).astype(np.uint8)
    Pd, S, T = _t(P), _t(s), _t(t)
    sP, tP = host.g1_mul(Pd, S), host.g1_mul(Pd, T)
    st = _t(((s.astype(np.int64) + t) % 17).astype(np.uint8))
    assert torch.equal(host.g1_add(sP, tP), host.g1_mul(Pd, st))
    assert bool(host.g1_is_on_curve(sP).all())
    sl = slice(9_000_000, 9_000_000 + 8192)
    ps.eq("g1_mul window", sP[sl].cpu().numpy(), oracle.g1_mul(P[sl], s[sl].astype(np.uint64)))


def test_config4_pairing_2p22_bilinearity(host, oracle, W):
    """2^22 pairings: e(aG, bH) = e(G, H)^(ab) -- checked with gtp_pow on the device -- and
    e(P, Q) e(R, Q) = e(P + R, Q) when no point is the identity (e(identity, .) = (0,0) breaks it, hazard C-10)."""
    import torch
    ai, bi, s = W.make_group_items(15, 0, N22)
    T17 = W.g1_subgroup_table()
    P = T17[ai]
    H = np.tile(np.array([[36, 31]], np.uint8), (N22, 1))
    Q = host.g2_mul(_t(H), _t(bi.astype(np.int64)))
    e = host.pairing(_t(P), Q)
    base = np.tile(np.array([[7, 28]], np.uint8), (N22, 1))           # e(G, H), SURVEY.md A.3
    want = host.gtp_pow(_t(base), _t(ai.astype(np.int64) * bi.astype(np.int64)))
    assert torch.equal(e, want), "e(aG, bH) != e(G,H)^(ab)"
    ri = (ai.astype(np.int64) * 3) % 17
    keep = (ri != 0) & ((ai + ri) % 17 != 0)
    R = T17[ri]
    lhs = host.gtp_mul(e, host.pairing(_t(R), Q))
    rhs = host.pairing(host.g1_add(_t(P), _t(R)), Q)
    k = _t(keep)
    assert torch.equal(lhs[k], rhs[k])
    sl = slice(1_234_567, 1_234_567 + 8192)
    ps.eq("pairing window", e[sl].cpu().numpy(), oracle.pairing(P[sl], Q[sl].cpu().numpy(), 8))


@pytest.mark.parametrize("mode", ["identity", "generator"])
def test_config5_prove_verify_2p22(host, oracle, W, mode):
    """2^22 proofs (the per-GPU share of config 5 is 2^21): status histogram matches SURVEY.md Appendix B, the run
    is deterministic, every identity-SRS proof verifies, and two windows agree with the oracle byte for byte."""
    import torch
    g1s, g2 = (W.identity_srs if mode == "identity" else W.generator_srs)(9)
    pk = host.Plonk(W.PLONK_TEST_CIRCUIT, g1s, g2)
    wit, rnd, chal, u = W.make_batch(2025, 0, N22, "U17")
    d = [_t(x) for x in (wit, rnd, chal, u)]
    proofs, status, verdict = pk.prove_verify(*d)
    p2, s2, v2 = pk.prove_verify(*d)
    assert torch.equal(proofs, p2) and torch.equal(status, s2) and torch.equal(verdict, v2)
    counts = torch.zeros(18, dtype=torch.int64, device="cuda")
    host.tally(proofs, status, verdict, counts)
    torch.cuda.synchronize()
    c = counts.cpu().numpy()
    hist = np.bincount(status.cpu().numpy(), minlength=15)
    assert c[:15].tolist() == hist.tolist() and c[15] == 0
    assert c[17] == int(proofs.to(torch.int64).sum()), "checksum of proof bytes"
    frac = hist / N22
    assert abs(frac[0] - 0.607) < 0.01 and abs(frac[8] - 0.335) < 0.01 and abs(frac[9] - 0.058) < 0.005 and hist[[0, 8, 9]].sum() == N22
    done = status == 0
    assert bool((verdict[~done] == 0xFF).all())
    if mode == "identity":
        assert bool((verdict[done] == 1).all()) and c[16] == hist[0]
    else:
        acc = int((verdict[done] == 1).sum()) / int(done.sum())
        assert 0.20 < acc < 0.26, acc          # the reference's non-standard r(x): SURVEY.md Appendix C-9 (22.8%)
    for lo in (0, 3_777_000):
        sl = slice(lo, lo + 8192)
        rp, rs = oracle.plonk_prove_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, wit[sl], rnd[sl], chal[sl], 8)
        rv, _ = oracle.plonk_verify_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, rp, chal[sl], u[sl], 8)
        rv = np.where(rs == 0, rv, 0xFF).astype(np.uint8)
        ps.eq(f"window {lo}", (proofs[sl].cpu().numpy(), status[sl].cpu().numpy(), verdict[sl].cpu().numpy()), (rp, rs, rv))


def test_tally_matches_numpy(host):
    """pb_tally_dev against plonk_c_b200.shard.tally_host: arbitrary bytes (status values beyond 14 fold into bin 15),
    ragged sizes, pointers that are not 16-byte aligned (byte path), absent arrays, accumulation over calls."""
    import torch
    from plonk_c_b200 import shard
    rng = np.random.default_rng(5)
    for n, off in ((1, 0), (15, 0), (16, 0), (4097, 0), (100003, 0), (100003, 1), (65536, 3), (1 << 20, 0)):
        proofs = rng.integers(0, 256, (n + 4, 34), dtype=np.uint8)
        status = rng.integers(0, 256, n + 4, dtype=np.uint8)
        status[rng.random(n + 4) < 0.7] = 0
        status[rng.random(n + 4) < 0.2] = 8
        verdict = rng.integers(0, 3, n + 4, dtype=np.uint8)
        dp, ds, dv = (_t(x)[off:off + n] for x in (proofs, status, verdict))
        hp, hs, hv = proofs[off:off + n], status[off:off + n], verdict[off:off + n]
        counts = torch.zeros(18, dtype=torch.int64, device="cuda")
        host.tally(dp, ds, dv, counts)
        host.tally(dp, ds, dv, counts)            # accumulates
        torch.cuda.synchronize()
        assert counts.cpu().numpy().tolist() == (2 * shard.tally_host(hp, hs, hv)).tolist(), (n, off)
        counts.zero_()
        host.tally(None, ds, None, counts)
        torch.cuda.synchronize()
        assert counts.cpu().numpy().tolist() == shard.tally_host(None, hs, None).tolist(), (n, off, "status only")
