"""On-disk batch format shared by the CPU baseline, the oracle fixtures and the GPU path.

A `.pbatch` file is a little-endian header followed by raw, item-major arrays of the reference's own structs -- exactly the
buffers the C ABI takes (include/plonk_b200.h), so a file can be read straight into pinned memory and handed to
pb_plonk_prove_verify without any conversion:

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
"""
import struct

import numpy as np

MAGIC = b"PLONKB2\0"
_HDR = struct.Struct("<8sIQII")


def write_batch(path, circuit, srs_g1s, srs_g2, witness, rand, chal, u, proofs=None, status=None, verdict=None):
    n = int(witness.shape[0])
    g1s = np.ascontiguousarray(srs_g1s, np.uint8)
    flags = (1 if proofs is not None else 0) | (2 if verdict is not None else 0)
    assert (proofs is None) == (status is None)
    with open(path, "wb") as f:
        f.write(_HDR.pack(MAGIC, 1, n, int(g1s.shape[0]), flags))
        for arr, shape in ((circuit, (44,)), (g1s, (g1s.shape[0], 3)), (srs_g2, (4,)), (witness, (n, 12)), (rand, (n, 9)),
                           (chal, (n, 5)), (u, (n,))):
            a = np.ascontiguousarray(arr, np.uint8)
            assert a.shape == shape, (a.shape, shape)
            f.write(a.tobytes())
        if proofs is not None:
            f.write(np.ascontiguousarray(proofs, np.uint8).reshape(n, 34).tobytes())
            f.write(np.ascontiguousarray(status, np.uint8).reshape(n).tobytes())
        if verdict is not None:
            f.write(np.ascontiguousarray(verdict, np.uint8).reshape(n).tobytes())


def read_batch(path):
    """Returns a dict of numpy arrays (memory-mapped, read-only) keyed like write_batch's arguments."""
    with open(path, "rb") as f:
        magic, version, n, srs_len, flags = _HDR.unpack(f.read(_HDR.size))
    if magic != MAGIC or version != 1:
        raise ValueError(f"{path}: not a .pbatch v1 file")
    mm = np.memmap(path, dtype=np.uint8, mode="r", offset=_HDR.size)
    out, pos = {"n": n, "flags": flags}, 0

    def take(name, shape):
        nonlocal pos
        size = int(np.prod(shape))
        if pos + size > mm.size:
            raise ValueError(f"{path}: truncated at {name}")
        out[name] = mm[pos:pos + size].reshape(shape)
        pos += size
    take("circuit", (44,)); take("srs_g1s", (srs_len, 3)); take("srs_g2", (4,))
    take("witness", (n, 12)); take("rand", (n, 9)); take("chal", (n, 5)); take("u", (n,))
    if flags & 1:
        take("proofs", (n, 34)); take("status", (n,))
    if flags & 2:
        take("verdict", (n,))
    if pos != mm.size:
        raise ValueError(f"{path}: {mm.size - pos} trailing bytes")
    return out
