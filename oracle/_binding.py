"""TEST INFRASTRUCTURE -- one ctypes binding for both CPU checkers (same exported functions,
prefix `ref_` for the compiled reference, `port_` for the C restatement)."""
import ctypes as C
import os

import numpy as np

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)


def _p(a):
    return None if a is None else a.ctypes.data_as(u8p)


def _u8(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if shape is not None:
        a = a.reshape(shape)
    return a


class OracleLib:
    def __init__(self, path, prefix):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built (run `make -C oracle` or __graft_entry__.build())")
        self.path = path
        self.prefix = prefix
        self.lib = C.CDLL(path)
        self.kind = "reference" if prefix == "ref_" else "port"

    def _f(self, name, argtypes, restype=None):
        fn = getattr(self.lib, self.prefix + name)
        fn.argtypes = argtypes
        fn.restype = restype
        return fn

    # ---------------------------------------------------------------- family 1
    def field_op(self, field, op, a, b=None):
        a = _u8(a)
        b = None if b is None else _u8(b)
        out = np.empty_like(a)
        self._f("field_op", [C.c_int, C.c_int, u8p, u8p, u8p, C.c_size_t])(field, op, _p(a), _p(b), _p(out), a.size)
        return out

    def hf_new(self, v):
        return self._f("hf_new", [C.c_int64], C.c_uint8)(v)

    def gf_new(self, v):
        return self._f("gf_new", [C.c_int64], C.c_uint8)(v)

    # ---------------------------------------------------------------- family 2
    def poly_binop(self, op, a, alen, b, blen, so):
        a, b = _u8(a), _u8(b)
        alen, blen = _u8(alen), _u8(blen)
        n = a.shape[0]
        out = np.zeros((n, so), np.uint8)
        olen = np.zeros(n, np.uint8)
        self._f("poly_binop", [C.c_int, u8p, u8p, C.c_size_t, u8p, u8p, C.c_size_t, u8p, u8p, C.c_size_t, C.c_size_t])(
            op, _p(a), _p(alen), a.shape[1], _p(b), _p(blen), b.shape[1], _p(out), _p(olen), so, n)
        return out, olen

    def poly_divide(self, num, nlen, den, dlen, sq, sr):
        num, den, nlen, dlen = _u8(num), _u8(den), _u8(nlen), _u8(dlen)
        n = num.shape[0]
        quot = np.zeros((n, sq), np.uint8)
        rem = np.zeros((n, sr), np.uint8)
        qlen = np.zeros(n, np.uint8)
        rlen = np.zeros(n, np.uint8)
        status = np.zeros(n, np.uint8)
        self._f("poly_divide", [u8p, u8p, C.c_size_t] * 4 + [u8p, C.c_size_t])(
            _p(num), _p(nlen), num.shape[1], _p(den), _p(dlen), den.shape[1],
            _p(quot), _p(qlen), sq, _p(rem), _p(rlen), sr, _p(status), n)
        return quot, qlen, rem, rlen, status

    def poly_eval(self, p, plen, x):
        p, plen, x = _u8(p), _u8(plen), _u8(x)
        n = p.shape[0]
        out = np.zeros(n, np.uint8)
        self._f("poly_eval", [u8p, u8p, C.c_size_t, u8p, u8p, C.c_size_t])(_p(p), _p(plen), p.shape[1], _p(x), _p(out), n)
        return out

    def poly_unop(self, op, p, plen, k, so):
        p, plen = _u8(p), _u8(plen)
        k = None if k is None else _u8(k)
        n = p.shape[0]
        out = np.zeros((n, so), np.uint8)
        olen = np.zeros(n, np.uint8)
        self._f("poly_unop", [C.c_int, u8p, u8p, C.c_size_t, u8p, u8p, u8p, C.c_size_t, C.c_size_t])(
            op, _p(p), _p(plen), p.shape[1], _p(k), _p(out), _p(olen), so, n)
        return out, olen

    def poly_slice(self, p, plen, start, end, so):
        p, plen, start, end = _u8(p), _u8(plen), _u8(start), _u8(end)
        n = p.shape[0]
        out = np.zeros((n, so), np.uint8)
        olen = np.zeros(n, np.uint8)
        status = np.zeros(n, np.uint8)
        self._f("poly_slice", [u8p, u8p, C.c_size_t, u8p, u8p, u8p, u8p, C.c_size_t, u8p, C.c_size_t])(
            _p(p), _p(plen), p.shape[1], _p(start), _p(end), _p(out), _p(olen), so, _p(status), n)
        return out, olen, status

    def poly_z(self, points, so=32):
        points = _u8(points)
        out = np.zeros(so, np.uint8)
        olen = np.zeros(1, np.uint8)
        self._f("poly_z", [u8p, C.c_size_t, u8p, u8p, C.c_size_t])(_p(points), points.size, _p(out), _p(olen), so)
        return out, int(olen[0])

    def poly_lagrange(self, xs, ys, so):
        xs, ys = _u8(xs), _u8(ys)
        n, ln = xs.shape
        out = np.zeros((n, so), np.uint8)
        olen = np.zeros(n, np.uint8)
        status = np.zeros(n, np.uint8)
        self._f("poly_lagrange", [u8p, u8p, C.c_size_t, C.c_size_t, u8p, u8p, C.c_size_t, u8p])(
            _p(xs), _p(ys), ln, n, _p(out), _p(olen), so, _p(status))
        return out, olen, status

    def matrix_mul(self, a, b):
        a, b = _u8(a), _u8(b)
        out = np.zeros((a.shape[0], b.shape[1]), np.uint8)
        self._f("matrix_mul", [u8p, C.c_size_t, C.c_size_t, u8p, C.c_size_t, C.c_size_t, u8p])(
            _p(a), a.shape[0], a.shape[1], _p(b), b.shape[0], b.shape[1], _p(out))
        return out

    def matrix_inv(self, a):
        a = _u8(a)
        out = np.zeros_like(a)
        self._f("matrix_inv", [u8p, C.c_size_t, u8p])(_p(a), a.shape[0], _p(out))
        return out

    def matrix_gauss_jordan(self, a):
        a = _u8(a).copy()
        self._f("matrix_gauss_jordan", [u8p, C.c_size_t, C.c_size_t])(_p(a), a.shape[0], a.shape[1])
        return a

    # ---------------------------------------------------------------- family 3
    def g1_op(self, op, a, b=None):
        a = _u8(a)
        b = None if b is None else _u8(b)
        out = np.zeros_like(a)
        self._f("g1_op", [C.c_int, u8p, u8p, u8p, C.c_size_t])(op, _p(a), _p(b), _p(out), a.shape[0])
        return out

    def g1_mul(self, p, scalars, nthreads=1):
        p = _u8(p)
        s = np.ascontiguousarray(scalars, dtype=np.uint64)
        out = np.zeros_like(p)
        self._f("g1_mul", [u8p, u64p, u8p, C.c_size_t, C.c_int])(_p(p), s.ctypes.data_as(u64p), _p(out), p.shape[0], nthreads)
        return out

    def g1_is_on_curve(self, p):
        p = _u8(p)
        out = np.zeros(p.shape[0], np.uint8)
        self._f("g1_is_on_curve", [u8p, u8p, C.c_size_t])(_p(p), _p(out), p.shape[0])
        return out

    def g2_op(self, op, a, b=None):
        a = _u8(a)
        b = None if b is None else _u8(b)
        out = np.zeros_like(a)
        self._f("g2_op", [C.c_int, u8p, u8p, u8p, C.c_size_t])(op, _p(a), _p(b), _p(out), a.shape[0])
        return out

    def g2_mul(self, p, scalars):
        p = _u8(p)
        s = np.ascontiguousarray(scalars, dtype=np.uint64)
        out = np.zeros_like(p)
        self._f("g2_mul", [u8p, u64p, u8p, C.c_size_t])(_p(p), s.ctypes.data_as(u64p), _p(out), p.shape[0])
        return out

    def gtp_mul(self, a, b):
        a, b = _u8(a), _u8(b)
        out = np.zeros_like(a)
        self._f("gtp_mul", [u8p, u8p, u8p, C.c_size_t])(_p(a), _p(b), _p(out), a.shape[0])
        return out

    def gtp_pow(self, a, e):
        a = _u8(a)
        e = np.ascontiguousarray(e, dtype=np.uint64)
        out = np.zeros_like(a)
        self._f("gtp_pow", [u8p, u64p, u8p, C.c_size_t])(_p(a), e.ctypes.data_as(u64p), _p(out), a.shape[0])
        return out

    def srs_create(self, secret, n):
        g1s = np.zeros((n + 1, 3), np.uint8)
        g2 = np.zeros(4, np.uint8)
        self._f("srs_create", [C.c_uint8, C.c_uint32, u8p, u8p])(secret, n, _p(g1s), _p(g2))
        return g1s, g2

    def srs_eval_at_s(self, g1s, g2, polys, plen, nthreads=1):
        g1s, g2, polys, plen = _u8(g1s), _u8(g2), _u8(polys), _u8(plen)
        n = polys.shape[0]
        out = np.zeros((n, 3), np.uint8)
        status = np.zeros(n, np.uint8)
        self._f("srs_eval_at_s", [u8p, C.c_uint32, u8p, u8p, u8p, C.c_size_t, u8p, u8p, C.c_size_t, C.c_int])(
            _p(g1s), g1s.shape[0], _p(g2), _p(polys), _p(plen), polys.shape[1], _p(out), _p(status), n, nthreads)
        return out, status

    # ---------------------------------------------------------------- family 4
    def line(self, a, b):
        a, b = _u8(a), _u8(b)
        out = np.zeros((a.shape[0], 3), np.uint8)
        self._f("line", [u8p, u8p, u8p, C.c_size_t])(_p(a), _p(b), _p(out), a.shape[0])
        return out

    def pairing(self, p, q, nthreads=1):
        p, q = _u8(p), _u8(q)
        out = np.zeros((p.shape[0], 2), np.uint8)
        self._f("pairing", [u8p, u8p, u8p, C.c_size_t, C.c_int])(_p(p), _p(q), _p(out), p.shape[0], nthreads)
        return out

    def pairing_f(self, r, p, q):
        p, q = _u8(p), _u8(q)
        out = np.zeros((p.shape[0], 2), np.uint8)
        self._f("pairing_f", [C.c_uint64, u8p, u8p, u8p, C.c_size_t])(r, _p(p), _p(q), _p(out), p.shape[0])
        return out

    # ---------------------------------------------------------------- protocol
    def plonk_setup_dump(self):
        out = np.zeros(37, np.uint8)
        self._f("plonk_setup_dump", [u8p])(_p(out))
        return dict(h=out[0:4].copy(), k1_h=out[4:8].copy(), k2_h=out[8:12].copy(),
                    h_pows_inv=out[12:28].reshape(4, 4).copy(), z_h=out[28:28 + out[36]].copy())

    def copy_constraints_to_roots(self, types, idx):
        types, idx = _u8(types), _u8(idx)
        out = np.zeros(types.size, np.uint8)
        self._f("copy_constraints_to_roots", [u8p, u8p, C.c_size_t, u8p])(_p(types), _p(idx), types.size, _p(out))
        return out

    def interpolate_at_h(self, vals):
        vals = _u8(vals)
        n = vals.shape[0]
        out = np.zeros((n, 4), np.uint8)
        olen = np.zeros(n, np.uint8)
        self._f("interpolate_at_h", [u8p, u8p, u8p, C.c_size_t])(_p(vals), _p(out), _p(olen), n)
        return out, olen

    def plonk_prove_batch(self, circuit, g1s, g2, wit, rnd, chal, nthreads=1):
        circuit, g1s, g2 = _u8(circuit), _u8(g1s), _u8(g2)
        wit, rnd, chal = _u8(wit), _u8(rnd), _u8(chal)
        n = wit.shape[0]
        assert wit.shape == (n, 12) and rnd.shape == (n, 9) and chal.shape == (n, 5) and circuit.size == 44
        proofs = np.zeros((n, 34), np.uint8)
        status = np.zeros(n, np.uint8)
        self._f("plonk_prove_batch", [u8p, u8p, C.c_uint32, u8p, u8p, u8p, u8p, C.c_size_t, u8p, u8p, C.c_int])(
            _p(circuit), _p(g1s), g1s.shape[0], _p(g2), _p(wit), _p(rnd), _p(chal), n, _p(proofs), _p(status), nthreads)
        return proofs, status

    def plonk_verify_batch(self, circuit, g1s, g2, proofs, chal, u, nthreads=1, want_gt=True):
        circuit, g1s, g2 = _u8(circuit), _u8(g1s), _u8(g2)
        proofs, chal, u = _u8(proofs), _u8(chal), _u8(u)
        n = proofs.shape[0]
        verdict = np.zeros(n, np.uint8)
        gt = np.zeros((n, 4), np.uint8) if want_gt else None
        self._f("plonk_verify_batch", [u8p, u8p, C.c_uint32, u8p, u8p, u8p, u8p, C.c_size_t, u8p, u8p, C.c_int])(
            _p(circuit), _p(g1s), g1s.shape[0], _p(g2), _p(proofs), _p(chal), _p(u), n, _p(verdict), _p(gt), nthreads)
        return verdict, gt

    # ---- Fiat-Shamir mode (spec: oracle/fs_spec.inc)
    def plonk_prove_fs_batch(self, circuit, g1s, g2, wit, rnd, nthreads=1):
        """-> proofs [n][34], status [n], chal [n][6] = alpha beta gamma z v u (0xFF where not drawn)"""
        circuit, g1s, g2 = _u8(circuit), _u8(g1s), _u8(g2)
        wit, rnd = _u8(wit), _u8(rnd)
        n = wit.shape[0]
        assert wit.shape == (n, 12) and rnd.shape == (n, 9) and circuit.size == 44
        proofs = np.zeros((n, 34), np.uint8)
        status = np.zeros(n, np.uint8)
        chal = np.zeros((n, 6), np.uint8)
        self._f("plonk_prove_fs_batch", [u8p, u8p, C.c_uint32, u8p, u8p, u8p, C.c_size_t, u8p, u8p, u8p, C.c_int])(
            _p(circuit), _p(g1s), g1s.shape[0], _p(g2), _p(wit), _p(rnd), n, _p(proofs), _p(status), _p(chal), nthreads)
        return proofs, status, chal

    def fs_seed(self, circuit, g1s, g2):
        """the transcript state after absorbing circuit and SRS: four 32-bit words as one Python int (v0 | v1 << 32 | v2 << 64 | v3 << 96)"""
        circuit, g1s, g2 = _u8(circuit), _u8(g1s), _u8(g2)
        out = np.zeros(4, np.uint32)
        self._f("fs_seed", [u8p, u8p, C.c_uint32, u8p, C.c_void_p], None)(_p(circuit), _p(g1s), g1s.shape[0], _p(g2), out.ctypes.data_as(C.c_void_p))
        return sum(int(w) << (32 * k) for k, w in enumerate(out))

    def fs_challenges(self, circuit, g1s, g2, proofs):
        return self.fs_derive(self.fs_seed(circuit, g1s, g2), proofs)

    def plonk_verify_fs_batch(self, circuit, g1s, g2, proofs, nthreads=1, want_gt=True):
        """the verifier's side of Fiat-Shamir mode: derive the six challenges, then the explicit-challenge verifier"""
        d = self.fs_challenges(circuit, g1s, g2, proofs)
        return self.plonk_verify_batch(circuit, g1s, g2, proofs, np.ascontiguousarray(d[:, :5]), np.ascontiguousarray(d[:, 5]),
                                       nthreads, want_gt)

    def fs_derive(self, seed, proofs):
        proofs = _u8(proofs)
        n = proofs.shape[0]
        chal = np.zeros((n, 6), np.uint8)
        words = np.array([(int(seed) >> (32 * k)) & 0xFFFFFFFF for k in range(4)], np.uint32)
        self._f("fs_derive", [C.c_void_p, u8p, C.c_size_t, u8p])(words.ctypes.data_as(C.c_void_p), _p(proofs), n, _p(chal))
        return chal

    def verifier_key(self, circuit, g1s, g2):
        circuit, g1s, g2 = _u8(circuit), _u8(g1s), _u8(g2)
        out = np.zeros((9, 3), np.uint8)
        self._f("verifier_key", [u8p, u8p, C.c_uint32, u8p, u8p])(_p(circuit), _p(g1s), g1s.shape[0], _p(g2), _p(out))
        return out


_HERE = os.path.dirname(os.path.abspath(__file__))


def load_ref(libc_malloc=False):
    name = "libref_oracle_libc.so" if libc_malloc else "libref_oracle.so"
    return OracleLib(os.path.join(_HERE, "_ref", name), "ref_")


def load_port():
    return OracleLib(os.path.join(_HERE, "libplonk_port.so"), "port_")


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_oracle.so"))
