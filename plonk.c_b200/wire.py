"""On-disk / on-wire batch formats shared by the CPU baseline, the oracle fixtures and the GPU path (SURVEY.md section
8(f) rank 3).

Version 1 -- raw arrays of the reference's own structs.  A `.pbatch` file is a little-endian header followed by
item-major arrays, exactly the buffers the C ABI takes (include/plonk_b200.h), so a file can be read straight into
pinned memory and handed to pb_plonk_prove_verify without any conversion:

    magic   8 bytes  b"PLONKB2\\0"
    version u32      1
    n       u64      number of items
    srs_len u32      number of G1 points in the SRS
    flags   u32      bit 0: proofs/status present, bit 1: verdicts present
    circuit 44 bytes q_l q_r q_o q_m q_c | c_a.type c_a.index | c_b.. | c_c..      (constraints.h:35-47)
    srs_g1s srs_len x 3 bytes  G1{x, y, infinite}                                  (g1.h:8-11)
    srs_g2  4 bytes  g2_1.x g2_1.y g2_s.x g2_s.y                                   (srs.h:11-16)
    witness n x 12   a[4] b[4] c[4]                                                (constraints.h:57-62)
    rand    n x 9    HF rand[9]                                                    (plonk.h:228)
    chal    n x 5    CHALLENGE                                                     (plonk.h:16-22)
    u       n x 1    verifier challenge
    [proofs n x 34   PROOF (plonk.h:24-41)]  [status n x 1  SURVEY.md Appendix B row]   if flags & 1
    [verdict n x 1]                                                                      if flags & 2

Version 2 -- the packed wire format (csrc/wire.cuh; what pb_plonk_prove_verify_packed moves over PCIe).  Same header
(version = 2, flags bit 2 = packed outputs present), same circuit / SRS block, then

    inputs  n x 16   four little-endian u32 per item; word k = sum_j v[7k + j] * 17^j over the 27 values
                     a[4] b[4] c[4] | rand[9] | alpha beta gamma z v | u   (+ one spare digit, 0)
    [n_done u64, sv n x 1 (low nibble status, 15 = bad input; high nibble verdict, 15 = not verified),
     proofs n_done x 22: nine u16 points x | y << 7 | infinite << 14, one u32 with the seven openings as base-17 digits;
     only completed proofs (status 0), in item order]                                   if flags & 4

Version 3 -- packed wire v3 (csrc/wire.cuh; what pb_plonk_prove_verify_packed3 moves): as version 2 with 14-byte input
records and 12-byte proof records (commitments as 7-bit indices into the 102 points of the curve: only for SRSs on the
curve; write_batch raises ValueError for a proof with a point off it).

The numpy functions below are the formats' reference implementation; the C helpers pb_wire_* and the device functions
of csrc/wire.cuh are checked against them (tests/test_wire.py).
"""
import struct

import numpy as np

MAGIC = b"PLONKB2\0"
_HDR = struct.Struct("<8sIQII")
PACKED_IN_BYTES, PACKED_PROOF_BYTES = 16, 22
_POW17 = 17 ** np.arange(7, dtype=np.uint64)


# ---------------------------------------------------------------- packed wire v2 (numpy reference implementation)
def pack_inputs(witness, rand, chal, u):
    """[n][12], [n][9], [n][5], [n] -> [n][16] packed records.  An item with a byte > 16 gets the all-ones record,
    which is not an encoding (the prover reports it as PB_PROVE_BAD_INPUT, as it does for the struct bytes)."""
    n = int(witness.shape[0])
    v = np.zeros((n, 28), np.uint64)
    v[:, 0:12] = witness
    v[:, 12:21] = rand
    v[:, 21:26] = chal
    v[:, 26] = u
    words = (v.reshape(n, 4, 7) * _POW17[None, None, :]).sum(axis=2).astype(np.uint32)
    words[(v > 16).any(axis=1)] = 0xFFFFFFFF
    return np.ascontiguousarray(words).view(np.uint8).reshape(n, 16)


def unpack_inputs(packed):
    """[n][16] -> (witness, rand, chal, u, valid).  Invalid records (word >= 17^7 or spare digit != 0) decode to 0xFF."""
    n = int(packed.shape[0])
    words = np.ascontiguousarray(packed, np.uint8).reshape(n, 16).view("<u4").astype(np.uint64)
    digits = (words[:, :, None] // _POW17[None, None, :]) % np.uint64(17)
    valid = (words < np.uint64(17 ** 7)).all(axis=1) & (digits[:, 3, 6] == 0)
    v = digits.reshape(n, 28).astype(np.uint8)
    v[~valid] = 0xFF
    return (np.ascontiguousarray(v[:, 0:12]), np.ascontiguousarray(v[:, 12:21]), np.ascontiguousarray(v[:, 21:26]),
            np.ascontiguousarray(v[:, 26]), valid)


def pack_proofs(proofs):
    """[m][34] PROOF structs -> [m][22] packed records."""
    p = np.ascontiguousarray(proofs, np.uint8).reshape(-1, 34)
    m = p.shape[0]
    pts = p[:, :27].reshape(m, 9, 3).astype(np.uint16)
    w16 = pts[:, :, 0] | (pts[:, :, 1] << 7) | ((pts[:, :, 2] != 0).astype(np.uint16) << 14)
    sc = (p[:, 27:34].astype(np.uint64) * _POW17[None, :]).sum(axis=1).astype("<u4")
    out = np.empty((m, 22), np.uint8)
    out[:, :18] = w16.astype("<u2").view(np.uint8).reshape(m, 18)
    out[:, 18:] = sc.view(np.uint8).reshape(m, 4)
    return out


def unpack_proofs(packed):
    """[m][22] -> [m][34]"""
    q = np.ascontiguousarray(packed, np.uint8).reshape(-1, 22)
    m = q.shape[0]
    w16 = np.ascontiguousarray(q[:, :18]).view("<u2").reshape(m, 9)
    sc = np.ascontiguousarray(q[:, 18:]).view("<u4").reshape(m).astype(np.uint64)
    if (sc >= 17 ** 7).any():
        raise ValueError("packed proof record is not a canonical encoding")
    out = np.empty((m, 34), np.uint8)
    pts = np.stack([w16 & 0x7F, (w16 >> 7) & 0x7F, w16 >> 14], axis=2)
    out[:, :27] = pts.reshape(m, 27)
    out[:, 27:] = ((sc[:, None] // _POW17[None, :]) % np.uint64(17)).astype(np.uint8)
    return out


# ---------------------------------------------------------------- packed wire v3 (numpy reference implementation)
# 14 B in, 12 B per completed proof out; layout in csrc/wire.cuh and include/plonk_b200.h.
PACKED3_IN_BYTES, PACKED3_PROOF_BYTES = 14, 12


def curve_points():
    """The 102 points of E(F_101): y^2 = x^3 + 3 in wire order: index 0 = infinity (stored as x = y = 0), then the affine
    points by (x, y).  -> (px[102], py[102], base[102]) with base[x] = index of the first point with abscissa x."""
    px, py, base = [0], [0], []
    for x in range(101):
        base.append(len(px))
        rhs = (x * x * x + 3) % 101
        for y in range(101):
            if y * y % 101 == rhs:
                px.append(x)
                py.append(y)
    base.append(len(px))
    assert len(px) == 102
    return np.array(px, np.uint8), np.array(py, np.uint8), np.array(base, np.uint8)


_PX, _PY, _BASE = curve_points()


def pack_inputs3(witness, rand, chal, u):
    """[n][12], [n][9], [n][5], [n] -> [n][14] v3 records (all-ones record for an item with a byte > 16)."""
    n = int(witness.shape[0])
    v = np.zeros((n, 27), np.uint64)
    v[:, 0:12] = witness
    v[:, 12:21] = rand
    v[:, 21:26] = chal
    v[:, 26] = u
    words = (v[:, :21].reshape(n, 3, 7) * _POW17[None, None, :]).sum(axis=2)
    G = (v[:, 21:27] * _POW17[None, :6]).sum(axis=1)
    for k in range(3):
        words[:, k] |= ((G >> np.uint64(16 + 3 * k)) & np.uint64(7)) << np.uint64(29)
    out = np.empty((n, 14), np.uint8)
    out[:, :12] = np.ascontiguousarray(words.astype("<u4")).view(np.uint8).reshape(n, 12)
    out[:, 12:] = np.ascontiguousarray((G & np.uint64(0xFFFF)).astype("<u2")).view(np.uint8).reshape(n, 2)
    out[(v > 16).any(axis=1)] = 0xFF
    return out


def unpack_inputs3(packed):
    """[n][14] -> (witness, rand, chal, u, valid).  Invalid records decode to 0xFF."""
    q = np.ascontiguousarray(packed, np.uint8).reshape(-1, 14)
    n = q.shape[0]
    words = np.ascontiguousarray(q[:, :12]).view("<u4").reshape(n, 3).astype(np.uint64)
    G = np.ascontiguousarray(q[:, 12:]).view("<u2").reshape(n).astype(np.uint64)
    for k in range(3):
        G |= (words[:, k] >> np.uint64(29)) << np.uint64(16 + 3 * k)
    low = words & np.uint64(0x1FFFFFFF)
    valid = (low < np.uint64(17 ** 7)).all(axis=1) & (G < np.uint64(17 ** 6))
    v = np.empty((n, 27), np.uint8)
    v[:, :21] = ((low[:, :, None] // _POW17[None, None, :]) % np.uint64(17)).reshape(n, 21)
    v[:, 21:] = (G[:, None] // _POW17[None, :6]) % np.uint64(17)
    v[~valid] = 0xFF
    return (np.ascontiguousarray(v[:, 0:12]), np.ascontiguousarray(v[:, 12:21]), np.ascontiguousarray(v[:, 21:26]),
            np.ascontiguousarray(v[:, 26]), valid)


def pack_proofs3(proofs):
    """[m][34] PROOF structs whose points are on the curve -> [m][12] v3 records."""
    p = np.ascontiguousarray(proofs, np.uint8).reshape(-1, 34)
    m = p.shape[0]
    pts = p[:, :27].reshape(m, 9, 3).astype(np.int64)
    x, y, inf = pts[:, :, 0], pts[:, :, 1], pts[:, :, 2] != 0
    on = np.where(inf, (x == 0) & (y == 0), (x < 101) & (y < 101) & ((y * y - x * x * x - 3) % 101 == 0))
    if not on.all() or (p[:, 27:] > 16).any():
        raise ValueError("proof record has no packed v3 encoding (a point off the curve or an opening > 16)")
    idx = np.where(inf, 0, _BASE[np.minimum(x, 100)].astype(np.int64) + (2 * y > 101)).astype(np.uint64)
    sc = p[:, 27:].astype(np.uint64)
    words = np.zeros((m, 3), np.uint64)
    for k in range(3):
        words[:, k] = (idx[:, 3 * k] | idx[:, 3 * k + 1] << np.uint64(7) | idx[:, 3 * k + 2] << np.uint64(14)
                       | (sc[:, 2 * k] + np.uint64(17) * sc[:, 2 * k + 1]) << np.uint64(21)
                       | ((sc[:, 6] >> np.uint64(2 * k)) & np.uint64(3)) << np.uint64(30))
    return np.ascontiguousarray(words.astype("<u4")).view(np.uint8).reshape(m, 12)


def unpack_proofs3(packed):
    """[m][12] -> [m][34]"""
    q = np.ascontiguousarray(packed, np.uint8).reshape(-1, 12)
    m = q.shape[0]
    w = q.view("<u4").reshape(m, 3).astype(np.uint64)
    out = np.empty((m, 34), np.uint8)
    last = np.zeros(m, np.uint64)
    for k in range(3):
        for j in range(3):
            i = (w[:, k] >> np.uint64(7 * j)) & np.uint64(0x7F)
            if (i >= 102).any():
                raise ValueError("packed v3 proof record is not a canonical encoding")
            i = i.astype(np.int64)
            out[:, 3 * (3 * k + j)] = _PX[i]
            out[:, 3 * (3 * k + j) + 1] = _PY[i]
            out[:, 3 * (3 * k + j) + 2] = i == 0
        pair = (w[:, k] >> np.uint64(21)) & np.uint64(0x1FF)
        if (pair >= 289).any():
            raise ValueError("packed v3 proof record is not a canonical encoding")
        out[:, 27 + 2 * k] = pair % np.uint64(17)
        out[:, 28 + 2 * k] = pair // np.uint64(17)
        last |= (w[:, k] >> np.uint64(30)) << np.uint64(2 * k)
    if (last >= 17).any():
        raise ValueError("packed v3 proof record is not a canonical encoding")
    out[:, 33] = last
    return out


def make_sv(status, verdict):
    s = np.where(np.asarray(status) > 14, 15, status).astype(np.uint8)
    v = np.where(np.asarray(verdict) > 14, 15, verdict).astype(np.uint8)
    return s | (v << 4)


def split_sv(sv):
    s, v = sv & 15, sv >> 4
    return np.where(s == 15, 254, s).astype(np.uint8), np.where(v == 15, 0xFF, v).astype(np.uint8)


def scatter_proofs(proofs_dense, status):
    """dense completed proofs + status bytes -> the full [n][34] array of the struct API (zero where status != 0)"""
    status = np.asarray(status)
    out = np.zeros((status.shape[0], 34), np.uint8)
    out[status == 0] = np.asarray(proofs_dense).reshape(-1, 34)[: int((status == 0).sum())]
    return out


# ---------------------------------------------------------------- .pbatch files
def write_batch(path, circuit, srs_g1s, srs_g2, witness, rand, chal, u, proofs=None, status=None, verdict=None, version=1):
    n = int(witness.shape[0])
    g1s = np.ascontiguousarray(srs_g1s, np.uint8)
    assert (proofs is None) == (status is None)
    assert version in (1, 2, 3)
    if version == 1:
        flags = (1 if proofs is not None else 0) | (2 if verdict is not None else 0)
    else:
        assert proofs is None or verdict is not None, "packed outputs carry status and verdict together (sv byte)"
        flags = 4 if proofs is not None else 0
    with open(path, "wb") as f:
        f.write(_HDR.pack(MAGIC, version, n, int(g1s.shape[0]), flags))
        for arr, shape in ((circuit, (44,)), (g1s, (g1s.shape[0], 3)), (srs_g2, (4,))):
            a = np.ascontiguousarray(arr, np.uint8)
            assert a.shape == shape, (a.shape, shape)
            f.write(a.tobytes())
        if version == 1:
            for arr, shape in ((witness, (n, 12)), (rand, (n, 9)), (chal, (n, 5)), (u, (n,))):
                a = np.ascontiguousarray(arr, np.uint8)
                assert a.shape == shape, (a.shape, shape)
                f.write(a.tobytes())
            if proofs is not None:
                f.write(np.ascontiguousarray(proofs, np.uint8).reshape(n, 34).tobytes())
                f.write(np.ascontiguousarray(status, np.uint8).reshape(n).tobytes())
            if verdict is not None:
                f.write(np.ascontiguousarray(verdict, np.uint8).reshape(n).tobytes())
        else:
            f.write((pack_inputs3 if version == 3 else pack_inputs)(witness, rand, chal, u).tobytes())
            if proofs is not None:
                status = np.ascontiguousarray(status, np.uint8).reshape(n)
                done = np.ascontiguousarray(proofs, np.uint8).reshape(n, 34)[status == 0]
                f.write(struct.pack("<Q", int(done.shape[0])))
                f.write(make_sv(status, np.ascontiguousarray(verdict, np.uint8).reshape(n)).tobytes())
                f.write((pack_proofs3 if version == 3 else pack_proofs)(done).tobytes())     # v3: ValueError for a point off the curve


def read_batch(path, raw=False):
    """Returns a dict of numpy arrays keyed like write_batch's arguments (v1: memory-mapped, read-only; v2: decoded to the
    struct arrays, plus "packed_inputs" / "packed_proofs" / "sv" exactly as stored when raw=True)."""
    with open(path, "rb") as f:
        magic, version, n, srs_len, flags = _HDR.unpack(f.read(_HDR.size))
    if magic != MAGIC or version not in (1, 2, 3):
        raise ValueError(f"{path}: not a .pbatch v1/v2/v3 file")
    mm = np.memmap(path, dtype=np.uint8, mode="r", offset=_HDR.size)
    out, pos = {"n": n, "flags": flags, "version": version}, 0

    def take(name, shape):
        nonlocal pos
        size = int(np.prod(shape))
        if pos + size > mm.size:
            raise ValueError(f"{path}: truncated at {name}")
        out[name] = mm[pos:pos + size].reshape(shape)
        pos += size
    take("circuit", (44,)); take("srs_g1s", (srs_len, 3)); take("srs_g2", (4,))
    if version == 1:
        take("witness", (n, 12)); take("rand", (n, 9)); take("chal", (n, 5)); take("u", (n,))
        if flags & 1:
            take("proofs", (n, 34)); take("status", (n,))
        if flags & 2:
            take("verdict", (n,))
    else:
        v3 = version == 3
        take("packed_inputs", (n, 14 if v3 else 16))
        out["witness"], out["rand"], out["chal"], out["u"], out["valid"] = (unpack_inputs3 if v3 else unpack_inputs)(np.array(out["packed_inputs"]))
        if flags & 4:
            take("n_done", (8,))
            out["n_done"] = int(np.array(out["n_done"]).view("<u8")[0])
            take("sv", (n,)); take("packed_proofs", (out["n_done"], 12 if v3 else 22))
            out["status"], out["verdict"] = split_sv(np.array(out["sv"]))
            out["proofs"] = scatter_proofs((unpack_proofs3 if v3 else unpack_proofs)(np.array(out["packed_proofs"])), out["status"])
        if not raw:
            for k in ("packed_inputs", "packed_proofs", "sv"):
                out.pop(k, None)
    if pos != mm.size:
        raise ValueError(f"{path}: {mm.size - pos} trailing bytes")
    return out
