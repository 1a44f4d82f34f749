"""The .pbatch on-disk format (plonk.c_b200/wire.py): round trip, and the oracle's results stored next to the inputs."""
import numpy as np
import pytest

from plonk_c_b200 import wire


def test_round_trip_with_oracle_results(tmp_path, oracle, W):
    n = 3001
    g1s, g2 = W.generator_srs(9)
    wit, rnd, chal, u = W.make_batch(8, 0, n)
    proofs, status = oracle.plonk_prove_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal)
    verdict, _ = oracle.plonk_verify_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, proofs, chal, u)
    path = tmp_path / "batch.pbatch"
    wire.write_batch(path, W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal, u, proofs, status, verdict)
    assert path.stat().st_size == 28 + 44 + 30 + 4 + n * (12 + 9 + 5 + 1 + 34 + 1 + 1)
    b = wire.read_batch(path)
    for k, v in dict(circuit=W.PLONK_TEST_CIRCUIT, srs_g1s=g1s, srs_g2=g2, witness=wit, rand=rnd, chal=chal, u=u,
                     proofs=proofs, status=status, verdict=verdict).items():
        assert np.array_equal(b[k], v), k
    # inputs only
    wire.write_batch(path, W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal, u)
    b = wire.read_batch(path)
    assert b["flags"] == 0 and "proofs" not in b and b["n"] == n


def test_rejects_garbage(tmp_path):
    p = tmp_path / "x.pbatch"
    p.write_bytes(b"not a batch file at all, definitely not" * 3)
    with pytest.raises(ValueError):
        wire.read_batch(p)


# ---------------------------------------------------------------- packed wire v2 (csrc/wire.cuh, pb_wire_*, wire.py)
def _hc():
    import ctypes as C
    import os
    import __graft_entry__ as g
    return C.CDLL(g.build_hostcheck())


def test_packed_inputs_three_implementations_agree(host, W):
    """numpy (wire.py) = C helpers (pb_wire_*) = the device function the kernels call (wire.cuh, compiled for the host)."""
    import ctypes as C
    n = 50021
    wit, rnd, chal, u = W.make_batch(77, 5, n)
    wit[0], rnd[0], chal[0], u[0] = 16, 16, 16, 16       # every digit at its maximum
    wit[1], rnd[1], chal[1], u[1] = 0, 0, 0, 0
    packed = wire.pack_inputs(wit, rnd, chal, u)
    assert packed.shape == (n, 16) and np.array_equal(packed, host.wire_pack_inputs(wit, rnd, chal, u))
    assert packed[0].view("<u4").max() == 17 ** 7 - 1 and (packed[0].view("<u4")[3] == 17 ** 6 - 1)
    for unpack in (wire.unpack_inputs, host.wire_unpack_inputs):
        w2, r2, c2, u2, valid = unpack(packed)
        assert valid.all() and np.array_equal(w2, wit) and np.array_equal(r2, rnd) and np.array_equal(c2, chal) and np.array_equal(u2, u)
    hc = _hc()
    v27, ok = np.zeros((n, 27), np.uint8), np.zeros(n, np.uint8)
    hc.hc_unpack_input16(packed.ctypes.data_as(C.c_void_p), v27.ctypes.data_as(C.c_void_p), ok.ctypes.data_as(C.c_void_p), C.c_size_t(n))
    assert ok.all() and np.array_equal(v27, np.concatenate([wit, rnd, chal, u[:, None]], axis=1))


def test_packed_inputs_reject_non_encodings(host):
    import ctypes as C
    rng = np.random.default_rng(3)
    n = 4096
    packed = rng.integers(0, 256, (n, 16), dtype=np.uint8)          # random bytes: almost never an encoding
    words = packed.view("<u4")
    words[:8] = 17 ** 7 - 1
    words[:8, 3] = 17 ** 6 - 1                                           # valid: spare digit 0
    words[8:16] = 17 ** 7 - 1                                            # invalid only because the spare digit is 16
    words[16:24] = 5
    words[16:24, 1] = 17 ** 7                                            # first non-encoding word value
    _, _, _, _, valid = wire.unpack_inputs(packed)
    assert valid[:8].all() and not valid[8:24].any() and valid.sum() < 40
    w2, r2, c2, u2, v2 = host.wire_unpack_inputs(packed)
    assert np.array_equal(v2, valid)
    a = wire.unpack_inputs(packed)
    for x, y in zip(a[:4], (w2, r2, c2, u2)):
        assert np.array_equal(x, y)
    hc = _hc()
    v27, ok = np.zeros((n, 27), np.uint8), np.zeros(n, np.uint8)
    hc.hc_unpack_input16(packed.ctypes.data_as(C.c_void_p), v27.ctypes.data_as(C.c_void_p), ok.ctypes.data_as(C.c_void_p), C.c_size_t(n))
    assert np.array_equal(ok.astype(bool), valid)
    # an input byte > 16 has no encoding: the packers emit the all-ones record
    wit = np.zeros((3, 12), np.uint8); rnd = np.zeros((3, 9), np.uint8); chal = np.zeros((3, 5), np.uint8); u = np.zeros(3, np.uint8)
    wit[1, 4] = 17; u[2] = 200
    for pack in (wire.pack_inputs, host.wire_pack_inputs):
        p = pack(wit, rnd, chal, u)
        assert (p[0] == 0).all() and (p[1] == 0xFF).all() and (p[2] == 0xFF).all()


def test_unpack7_against_division():
    """The digit extraction of the kernels (multiply-high divisions) on 2^32 / 61 evenly spread words plus both edges."""
    import ctypes as C
    hc = _hc()
    hc.hc_check_unpack7.restype = C.c_uint64
    assert hc.hc_check_unpack7(C.c_uint32(0), C.c_uint32(0xFFFFFFFF), C.c_uint32(61)) == 0
    assert hc.hc_check_unpack7(C.c_uint32(17 ** 7 - 70000), C.c_uint32(17 ** 7 + 70000), C.c_uint32(1)) == 0
    assert hc.hc_check_unpack7(C.c_uint32(0xFFFF0000), C.c_uint32(0xFFFFFFFF), C.c_uint32(1)) == 0
    assert hc.hc_check_unpack7(C.c_uint32(0), C.c_uint32(200000), C.c_uint32(1)) == 0


def test_packed_proofs_round_trip(host, oracle, W):
    n = 20000
    g1s, g2 = W.generator_srs(9)
    wit, rnd, chal, u = W.make_batch(31, 0, n)
    proofs, status = oracle.plonk_prove_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal)
    done = proofs[status == 0]
    pp = wire.pack_proofs(done)
    assert pp.shape == (done.shape[0], 22) and np.array_equal(pp, host.wire_pack_proofs(done))
    assert np.array_equal(wire.unpack_proofs(pp), done) and np.array_equal(host.wire_unpack_proofs(pp), done)
    assert np.array_equal(wire.scatter_proofs(done, status), proofs) and np.array_equal(host.wire_scatter_proofs(done, status), proofs)
    # arbitrary in-range bytes (garbage-SRS proofs hold off-curve points and identities with coordinates)
    rng = np.random.default_rng(4)
    g = np.concatenate([rng.integers(0, 101, (5000, 27), dtype=np.uint8), rng.integers(0, 17, (5000, 7), dtype=np.uint8)], axis=1)
    g[:, 2:27:3] = rng.integers(0, 2, (5000, 9), dtype=np.uint8)
    assert np.array_equal(wire.unpack_proofs(wire.pack_proofs(g)), g) and np.array_equal(host.wire_pack_proofs(g), wire.pack_proofs(g))
    st = rng.integers(0, 13, 1000).astype(np.uint8); st[::7] = 254
    vd = rng.integers(0, 4, 1000).astype(np.uint8); vd[st != 0] = 0xFF
    sv = wire.make_sv(st, vd)
    for split in (wire.split_sv, host.wire_split_sv):
        s2, v2 = split(sv)
        assert np.array_equal(s2, st) and np.array_equal(v2, vd)


# ---------------------------------------------------------------- packed wire v3 (14 B in, 12 B per proof; points as curve indices)
def test_v3_curve_point_list():
    px, py, base = wire.curve_points()
    assert len(px) == 102 and px[0] == 0 and py[0] == 0
    x, y = px[1:].astype(np.int64), py[1:].astype(np.int64)
    assert ((y * y - x * x * x - 3) % 101 == 0).all()                                    # all on the curve
    keys = x * 101 + y
    assert (np.diff(keys) > 0).all()                                                     # strictly ordered by (x, y): no duplicates
    assert all(base[v] == 1 + (x < v).sum() for v in range(101)) and base[101] == 102
    # the encoder's rule "second point of an abscissa = the one with 2y > 101" matches the list
    idx = base[x] + (2 * y > 101)
    assert np.array_equal(idx, np.arange(1, 102))


def test_v3_inputs_three_way_and_rejections(host):
    rng = np.random.default_rng(12)
    n = 40000
    wit, rnd, chal, u = (rng.integers(0, 17, s, dtype=np.uint8) for s in ((n, 12), (n, 9), (n, 5), (n,)))
    wit[0], rnd[0], chal[0], u[0] = 16, 16, 16, 16
    wit[1], rnd[1], chal[1], u[1] = 0, 0, 0, 0
    p = wire.pack_inputs3(wit, rnd, chal, u)
    assert p.shape == (n, 14) and np.array_equal(p, host.wire3_pack_inputs(wit, rnd, chal, u))
    w0 = p[0, :12].view("<u4")
    assert ((w0 & 0x1FFFFFFF) == 17 ** 7 - 1).all()
    G = int(p[0, 12:].view("<u2")[0]) | sum(int(w0[k] >> 29) << (16 + 3 * k) for k in range(3))
    assert G == 17 ** 6 - 1
    for unpack in (wire.unpack_inputs3, host.wire3_unpack_inputs):
        w2, r2, c2, u2, valid = unpack(p)
        assert valid.all() and np.array_equal(w2, wit) and np.array_equal(r2, rnd) and np.array_equal(c2, chal) and np.array_equal(u2, u)
    # non-encodings: a 29-bit field >= 17^7, G >= 17^6, random bytes
    bad = p[:64].copy()
    words = bad[:, :12].view("<u4")
    words[:16, 1] = (words[:16, 1] & 0xE0000000) | 17 ** 7
    bad[16:32, 12:] = 0xFF; words[16:32] |= 0xE0000000                                    # G = 2^25 - 1 > 17^6
    bad[32:64] = rng.integers(0, 256, (32, 14), dtype=np.uint8)
    bad[32:64, 3] |= 0x1F; bad[32:64, 2] = 0xFF                                             # word 0's field = 2^29 - 1: never an encoding
    a, b = wire.unpack_inputs3(bad), host.wire3_unpack_inputs(bad)
    assert not a[4].any() and all(np.array_equal(x, y) for x, y in zip(a, b)) and (a[0] == 0xFF).all()
    # an input byte > 16 has no encoding
    wit[5, 3] = 17; u[6] = 99
    for pack in (wire.pack_inputs3, host.wire3_pack_inputs):
        q = pack(wit, rnd, chal, u)
        assert (q[5] == 0xFF).all() and (q[6] == 0xFF).all() and np.array_equal(q[7], p[7])
    assert not wire.unpack_inputs3(q)[4][5:7].any()


def test_v3_proofs_three_way(host, oracle, W):
    n = 20000
    for g1s, g2 in (W.generator_srs(9), W.identity_srs(9)):
        wit, rnd, chal, u = W.make_batch(32, 0, n)
        proofs, status = oracle.plonk_prove_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal)
        done = proofs[status == 0]
        pp = wire.pack_proofs3(done)
        assert pp.shape == (done.shape[0], 12) and np.array_equal(pp, host.wire3_pack_proofs(done))
        assert np.array_equal(wire.unpack_proofs3(pp), done) and np.array_equal(host.wire3_unpack_proofs(pp), done)
    # every point of the curve in every slot, every opening value
    px, py, _ = wire.curve_points()
    rng = np.random.default_rng(9)
    m = 102 * 9
    rec = np.zeros((m, 34), np.uint8)
    idx = rng.integers(0, 102, (m, 9))
    for j in range(9):
        idx[j * 102:(j + 1) * 102, j] = np.arange(102)
    rec[:, 0:27:3], rec[:, 1:27:3], rec[:, 2:27:3] = px[idx], py[idx], idx == 0
    rec[:, 27:] = rng.integers(0, 17, (m, 7)); rec[:17, 33] = np.arange(17); rec[17:34, 27:] = 16
    pp = wire.pack_proofs3(rec)
    assert np.array_equal(pp, host.wire3_pack_proofs(rec)) and np.array_equal(wire.unpack_proofs3(pp), rec)
    assert np.array_equal(host.wire3_unpack_proofs(pp), rec)
    # records without an encoding are refused by both packers, non-encodings by both unpackers
    off = rec[:1].copy(); off[0, 1] = (off[0, 1] + 1) % 101
    if off[0, 2]: off[0, 0] = 5
    with pytest.raises(ValueError):
        wire.pack_proofs3(off)
    with pytest.raises(host.PlonkB200Error):
        host.wire3_pack_proofs(off)
    for word, value in ((0, 102), (1, 127 << 7), (2, 289 << 21), (2, (3 << 30) | (3 << 21))):      # index 102; index 127; pair 289; 7th opening >= 17
        q = pp[40:41].copy()
        w = q.view("<u4")
        if value == 102: w[0, word] = (w[0, word] & ~np.uint32(0x7F)) | np.uint32(102)
        elif value == 127 << 7: w[0, word] |= np.uint32(127 << 7)
        elif value == 289 << 21: w[0, word] = (w[0, word] & ~np.uint32(0x1FF << 21)) | np.uint32(289 << 21)
        else: w[0, 0] |= np.uint32(3 << 30); w[0, 1] |= np.uint32(3 << 30); w[0, 2] |= np.uint32(1 << 30)   # 31
        with pytest.raises(ValueError):
            wire.unpack_proofs3(q)
        with pytest.raises(host.PlonkB200Error):
            host.wire3_unpack_proofs(q)


def test_device_generator_equals_workload(W):
    """synth_item (wire.cuh), the generator of the seeded mode, is workload.make_batch bit for bit -- both variants,
    far-apart start offsets, and its packed records are the packing of its struct arrays."""
    import ctypes as C
    hc = _hc()
    tab = np.ascontiguousarray(W.satisfying_witnesses())
    for seed, start, variant in ((2025, 0, "U17"), (7, 123456789012, "NZ"), (2**63 + 5, 2**40 + 3, "U17")):
        n = 3000
        wit, rnd, chal, u = (np.zeros((n, 12), np.uint8), np.zeros((n, 9), np.uint8), np.zeros((n, 5), np.uint8), np.zeros(n, np.uint8))
        packed = np.zeros((n, 16), np.uint8)
        hc.hc_synth(C.c_uint64(seed), C.c_uint64(start), C.c_int(0 if variant == "U17" else 1), tab.ctypes.data_as(C.c_void_p),
                    *(x.ctypes.data_as(C.c_void_p) for x in (wit, rnd, chal, u, packed)), C.c_size_t(n))
        ref = W.make_batch(seed, start, n, variant)
        for x, y in zip((wit, rnd, chal, u), ref):
            assert np.array_equal(x, y)
        assert np.array_equal(packed, wire.pack_inputs(*ref))


def test_pbatch_v2_round_trip(tmp_path, oracle, W):
    n = 3001
    g1s, g2 = W.generator_srs(9)
    wit, rnd, chal, u = W.make_batch(8, 0, n)
    proofs, status = oracle.plonk_prove_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal)
    verdict, _ = oracle.plonk_verify_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, proofs, chal, u)
    verdict = np.where(status == 0, verdict, 0xFF).astype(np.uint8)
    path = tmp_path / "batch2.pbatch"
    wire.write_batch(path, W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal, u, proofs, status, verdict, version=2)
    k = int((status == 0).sum())
    assert path.stat().st_size == 28 + 44 + 30 + 4 + n * 16 + 8 + n + k * 22
    b = wire.read_batch(path)
    assert b["version"] == 2 and b["n_done"] == k and b["valid"].all()
    for key, v in dict(circuit=W.PLONK_TEST_CIRCUIT, srs_g1s=g1s, srs_g2=g2, witness=wit, rand=rnd, chal=chal, u=u,
                       proofs=proofs, status=status, verdict=verdict).items():
        assert np.array_equal(b[key], v), key
    raw = wire.read_batch(path, raw=True)
    assert raw["packed_inputs"].shape == (n, 16) and raw["packed_proofs"].shape == (k, 22)


def test_pbatch_v3_round_trip(tmp_path, oracle, W):
    import util
    n = 3001
    g1s, g2 = W.generator_srs(9)
    wit, rnd, chal, u = W.make_batch(9, 0, n)
    proofs, status = oracle.plonk_prove_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal)
    verdict, _ = oracle.plonk_verify_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, proofs, chal, u)
    verdict = np.where(status == 0, verdict, 0xFF).astype(np.uint8)
    path = tmp_path / "batch3.pbatch"
    wire.write_batch(path, W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal, u, proofs, status, verdict, version=3)
    k = int((status == 0).sum())
    assert path.stat().st_size == 28 + 44 + 30 + 4 + n * 14 + 8 + n + k * 12
    b = wire.read_batch(path)
    assert b["version"] == 3 and b["n_done"] == k and b["valid"].all()
    for key, v in dict(circuit=W.PLONK_TEST_CIRCUIT, srs_g1s=g1s, srs_g2=g2, witness=wit, rand=rnd, chal=chal, u=u,
                       proofs=proofs, status=status, verdict=verdict).items():
        assert np.array_equal(b[key], v), key
    raw = wire.read_batch(path, raw=True)
    assert raw["packed_inputs"].shape == (n, 14) and raw["packed_proofs"].shape == (k, 12)
    # inputs only; and proofs over an SRS off the curve have no v3 encoding
    wire.write_batch(path, W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal, u, version=3)
    b = wire.read_batch(path)
    assert b["flags"] == 0 and "proofs" not in b and np.array_equal(b["witness"], wit)
    gs, g2g = util.garbage_srs()
    gp, gst = oracle.plonk_prove_batch(W.PLONK_TEST_CIRCUIT, gs, g2g, wit[:500], rnd[:500], chal[:500])
    if (gst == 0).any():
        with pytest.raises(ValueError):
            wire.write_batch(path, W.PLONK_TEST_CIRCUIT, gs, g2g, wit[:500], rnd[:500], chal[:500], u[:500], gp, gst,
                             np.where(gst == 0, 0, 0xFF).astype(np.uint8), version=3)


@pytest.mark.gpu
def test_file_to_gpu_to_file(tmp_path, host, oracle, W):
    """A batch file is proved and verified on the GPU straight from its memory-mapped arrays; results equal the oracle's."""
    n = 20011
    g1s, g2 = W.generator_srs(9)
    wit, rnd, chal, u = W.make_batch(9, 0, n)
    path = tmp_path / "in.pbatch"
    wire.write_batch(path, W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal, u)
    b = wire.read_batch(path)
    pk = host.Plonk(b["circuit"], b["srs_g1s"], b["srs_g2"])
    proofs, status, verdict = pk.prove_verify(np.array(b["witness"]), np.array(b["rand"]), np.array(b["chal"]), np.array(b["u"]))
    rp, rs = oracle.plonk_prove_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, wit, rnd, chal)
    rv, _ = oracle.plonk_verify_batch(W.PLONK_TEST_CIRCUIT, g1s, g2, rp, chal, u)
    rv = np.where(rs == 0, rv, 0xFF).astype(np.uint8)
    assert np.array_equal(proofs, rp) and np.array_equal(status, rs) and np.array_equal(verdict, rv)
