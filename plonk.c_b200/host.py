"""ctypes binding of the C ABI (include/plonk_b200.h) -- the host-side mirror of the reference's header API.

Every function takes either numpy arrays (HOST path: the library copies in, launches, copies out) or
torch CUDA tensors (DEVICE path: `*_dev` entry points, enqueued on torch's current stream, outputs are
torch tensors).  Names follow the reference: poly_mul, poly_divide, poly_eval, interpolate_at_h, g1_mul,
srs_eval_at_s, pairing, plonk_prove ... (src/poly.h, src/g1.h, src/srs.h, src/pairing.h, src/plonk.h).

There is no fallback: if the shared library is missing, or no CUDA device is present, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PB_LIB", os.path.join(_HERE, "libplonk_b200.so"))   # PB_LIB: build variants for tuning experiments

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)

PB_OK, PB_ERR_NO_DEVICE, PB_ERR_CUDA, PB_ERR_ARG = 0, -1, -2, -3
OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_INV, OP_POW = range(7)
POLY_ADD, POLY_SUB, POLY_MUL = range(3)
POLY_SCALE, POLY_NEGATE, POLY_SHIFT, POLY_ADD_HF = range(4)
G_ADD, G_DOUBLE, G_NEG = range(3)


class PlonkB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


_lib = None


def lib():
    """Load libplonk_b200.so (built by __graft_entry__.build()).  Fails loudly when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _lib.pb_last_error.restype = C.c_char_p
    return _lib


def _check(rc):
    if rc != 0:
        raise PlonkB200Error(rc, lib().pb_last_error().decode())


def device_count():
    n = lib().pb_device_count()
    if n < 0:
        _check(n)
    return n


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class _Call:
    """Marshals one C-ABI call: numpy arrays -> host entry point, torch CUDA tensors -> `_dev` entry point."""

    def __init__(self, name, *inputs):
        self.name = name
        self.dev = any(_is_torch(x) for x in inputs if x is not None)
        self.keep = []
        if self.dev:
            import torch
            self.torch = torch
            self.device = next(x.device for x in inputs if x is not None and _is_torch(x))

    def inp(self, x, dtype=np.uint8):
        if x is None:
            return None
        if self.dev:
            t = x
            if not _is_torch(t):
                t = self.torch.as_tensor(np.ascontiguousarray(x, dtype=dtype), device=self.device)
            want = self.torch.uint8 if dtype == np.uint8 else self.torch.int64
            if dtype == np.uint64 and t.dtype in (self.torch.uint64,):
                pass
            elif t.dtype != want:
                t = t.to(want)
            t = t.contiguous()
            self.keep.append(t)
            return C.c_void_p(t.data_ptr())
        a = np.ascontiguousarray(x, dtype=dtype)
        self.keep.append(a)
        return a.ctypes.data_as(C.c_void_p)

    def outbuf(self, x, shape):
        """A caller-owned OUTPUT buffer: it must already be C-contiguous uint8 of the right size -- coercing it would make
        the library write into a temporary copy and leave the caller's buffer untouched."""
        if self.dev:
            if not (_is_torch(x) and x.dtype == self.torch.uint8 and x.is_contiguous() and tuple(x.shape) == tuple(shape)):
                raise ValueError(f"output buffer must be a contiguous torch.uint8 CUDA tensor of shape {tuple(shape)}")
            self.keep.append(x)
            return C.c_void_p(x.data_ptr())
        if not (isinstance(x, np.ndarray) and x.dtype == np.uint8 and x.flags.c_contiguous and x.flags.writeable
                and tuple(x.shape) == tuple(shape)):
            raise ValueError(f"output buffer must be a writeable C-contiguous uint8 numpy array of shape {tuple(shape)}")
        self.keep.append(x)
        return x.ctypes.data_as(C.c_void_p)

    def out(self, shape, dtype=np.uint8):
        if self.dev:
            t = self.torch.empty(shape, dtype=self.torch.uint8 if dtype == np.uint8 else self.torch.int64, device=self.device)
            self.keep.append(t)
            return t, C.c_void_p(t.data_ptr())
        a = np.empty(shape, dtype=dtype)
        self.keep.append(a)
        return a, a.ctypes.data_as(C.c_void_p)

    def run(self, *args, dev_tail=()):
        """dev_tail: arguments the `_dev` entry point takes after the stream"""
        fn = getattr(lib(), self.name + ("_dev" if self.dev else ""))
        fn.restype = C.c_int
        if self.dev:
            with self.torch.cuda.device(self.device):
                stream = C.c_void_p(self.torch.cuda.current_stream().cuda_stream)
                _check(fn(*args, stream, *dev_tail))
        else:
            _check(fn(*args))


def _n(x):
    return int(x.shape[0])


# ---------------------------------------------------------------- family (1): hf.h / gf.h
def field_op(field, op, a, b=None):
    c = _Call("pb_field_op", a, b)
    n = int(np.prod(a.shape))
    out, po = c.out(tuple(a.shape))
    c.run(C.c_int(field), C.c_int(op), c.inp(a), c.inp(b), po, C.c_size_t(n))
    return out


def field_new(field, v):
    """hf_new / gf_new over an int64 array (hf.h:25-35, gf.h:24-34)."""
    c = _Call("pb_field_new", v)
    n = int(np.prod(v.shape))
    out, po = c.out(tuple(v.shape))
    c.run(C.c_int(field), c.inp(v, np.int64), po, C.c_size_t(n))
    return out


def hf_new(v): return field_new(17, v)
def gf_new(v): return field_new(101, v)
def hf_add(a, b): return field_op(17, OP_ADD, a, b)
def hf_sub(a, b): return field_op(17, OP_SUB, a, b)
def hf_mul(a, b): return field_op(17, OP_MUL, a, b)
def hf_div(a, b): return field_op(17, OP_DIV, a, b)
def hf_neg(a): return field_op(17, OP_NEG, a)
def hf_inv(a): return field_op(17, OP_INV, a)
def hf_pow(a, e): return field_op(17, OP_POW, a, e)
def gf_add(a, b): return field_op(101, OP_ADD, a, b)
def gf_sub(a, b): return field_op(101, OP_SUB, a, b)
def gf_mul(a, b): return field_op(101, OP_MUL, a, b)
def gf_div(a, b): return field_op(101, OP_DIV, a, b)
def gf_neg(a): return field_op(101, OP_NEG, a)
def gf_inv(a): return field_op(101, OP_INV, a)
def gf_pow(a, e): return field_op(101, OP_POW, a, e)


# ---------------------------------------------------------------- family (2): poly.h / matrix.h
def poly_binop(op, a, alen, b, blen, so=None):
    c = _Call("pb_poly_binop", a, b)
    n, sa, sb = _n(a), int(a.shape[1]), int(b.shape[1])
    if so is None:
        so = sa + sb - 1 if op == POLY_MUL else max(sa, sb)
    out, po = c.out((n, so))
    olen, pl = c.out((n,))
    c.run(C.c_int(op), c.inp(a), c.inp(alen), C.c_size_t(sa), c.inp(b), c.inp(blen), C.c_size_t(sb), po, pl,
          C.c_size_t(so), C.c_size_t(n))
    return out, olen


def poly_add(a, alen, b, blen, so=None): return poly_binop(POLY_ADD, a, alen, b, blen, so)
def poly_sub(a, alen, b, blen, so=None): return poly_binop(POLY_SUB, a, alen, b, blen, so)
def poly_mul(a, alen, b, blen, so=None): return poly_binop(POLY_MUL, a, alen, b, blen, so)


def poly_divide(num, nlen, den, dlen, sq=None, sr=None):
    c = _Call("pb_poly_divide", num, den)
    n, sn, sd = _n(num), int(num.shape[1]), int(den.shape[1])
    sq = sn if sq is None else sq
    sr = max(sd - 1, 1) if sr is None else sr
    quot, pq = c.out((n, sq))
    qlen, pql = c.out((n,))
    rem, pr = c.out((n, sr))
    rlen, prl = c.out((n,))
    status, ps = c.out((n,))
    c.run(c.inp(num), c.inp(nlen), C.c_size_t(sn), c.inp(den), c.inp(dlen), C.c_size_t(sd), pq, pql, C.c_size_t(sq),
          pr, prl, C.c_size_t(sr), ps, C.c_size_t(n))
    return quot, qlen, rem, rlen, status


def poly_eval(p, plen, x):
    c = _Call("pb_poly_eval", p, x)
    n = _n(p)
    out, po = c.out((n,))
    c.run(c.inp(p), c.inp(plen), C.c_size_t(int(p.shape[1])), c.inp(x), po, C.c_size_t(n))
    return out


def poly_unop(op, p, plen, k, so):
    c = _Call("pb_poly_unop", p)
    n = _n(p)
    out, po = c.out((n, so))
    olen, pl = c.out((n,))
    c.run(C.c_int(op), c.inp(p), c.inp(plen), C.c_size_t(int(p.shape[1])), c.inp(k), po, pl, C.c_size_t(so), C.c_size_t(n))
    return out, olen


def poly_scale(p, plen, k, so=None): return poly_unop(POLY_SCALE, p, plen, k, so or int(p.shape[1]))
def poly_negate(p, plen, so=None): return poly_unop(POLY_NEGATE, p, plen, None, so or int(p.shape[1]))
def poly_shift(p, plen, k, so): return poly_unop(POLY_SHIFT, p, plen, k, so)
def poly_add_hf(p, plen, k, so=None): return poly_unop(POLY_ADD_HF, p, plen, k, so or int(p.shape[1]))


def poly_slice(p, plen, start, end, so=None):
    c = _Call("pb_poly_slice", p)
    n = _n(p)
    so = so or int(p.shape[1])
    out, po = c.out((n, so))
    olen, pl = c.out((n,))
    status, ps = c.out((n,))
    c.run(c.inp(p), c.inp(plen), C.c_size_t(int(p.shape[1])), c.inp(start), c.inp(end), po, pl, C.c_size_t(so), ps, C.c_size_t(n))
    return out, olen, status


def poly_lagrange(xs, ys, so=None):
    c = _Call("pb_poly_lagrange", xs, ys)
    n, ln = _n(xs), int(xs.shape[1])
    so = so or ln
    out, po = c.out((n, so))
    olen, pl = c.out((n,))
    status, ps = c.out((n,))
    c.run(c.inp(xs), c.inp(ys), C.c_size_t(ln), po, pl, C.c_size_t(so), ps, C.c_size_t(n))
    return out, olen, status


def matrix_mul(a, b):
    """a: [n][m][k], b: [n][k][c] -> [n][m][c]   (matrix_mul, matrix.h:81-98)"""
    c = _Call("pb_matrix_mul", a, b)
    n, m, k, cc = _n(a), int(a.shape[1]), int(a.shape[2]), int(b.shape[2])
    out, po = c.out((n, m, cc))
    c.run(c.inp(a), c.inp(b), po, C.c_uint32(m), C.c_uint32(k), C.c_uint32(cc), C.c_size_t(n))
    return out


def matrix_inv(a):
    """a: [n][d][d] -> [n][d][d]   (matrix_inv, matrix.h:151-176)"""
    c = _Call("pb_matrix_inv", a)
    n, d = _n(a), int(a.shape[1])
    out, po = c.out((n, d, d))
    c.run(c.inp(a), po, C.c_uint32(d), C.c_size_t(n))
    return out


# ---------------------------------------------------------------- family (3): g1.h / g2.h
def g1_op(op, a, b=None):
    c = _Call("pb_g1_op", a, b)
    n = _n(a)
    out, po = c.out((n, 3))
    c.run(C.c_int(op), c.inp(a), c.inp(b), po, C.c_size_t(n))
    return out


def g1_add(a, b): return g1_op(G_ADD, a, b)
def g1_double(a): return g1_op(G_DOUBLE, a)
def g1_neg(a): return g1_op(G_NEG, a)


def g1_mul(points, scalars):
    """g1_mul (g1.h:91-103).  scalars: uint64 (raw, never reduced) or uint8."""
    is_u8 = (scalars.dtype == np.uint8) if not _is_torch(scalars) else (str(scalars.dtype) == "torch.uint8")
    c = _Call("pb_g1_mul_u8" if is_u8 else "pb_g1_mul", points, scalars)
    n = _n(points)
    out, po = c.out((n, 3))
    c.run(c.inp(points), c.inp(scalars, np.uint8 if is_u8 else np.uint64), po, C.c_size_t(n))
    return out


def g1_is_on_curve(points):
    c = _Call("pb_g1_is_on_curve", points)
    n = _n(points)
    out, po = c.out((n,))
    c.run(c.inp(points), po, C.c_size_t(n))
    return out


def g2_op(op, a, b=None):
    c = _Call("pb_g2_op", a, b)
    n = _n(a)
    out, po = c.out((n, 2))
    c.run(C.c_int(op), c.inp(a), c.inp(b), po, C.c_size_t(n))
    return out


def g2_add(a, b): return g2_op(G_ADD, a, b)
def g2_neg(a): return g2_op(G_NEG, a)


def g2_mul(points, scalars):
    c = _Call("pb_g2_mul", points, scalars)
    n = _n(points)
    out, po = c.out((n, 2))
    c.run(c.inp(points), c.inp(scalars, np.uint64), po, C.c_size_t(n))
    return out


# ---------------------------------------------------------------- family (4): gt.h / pairing.h
def gtp_mul(a, b):
    c = _Call("pb_gtp_mul", a, b)
    n = _n(a)
    out, po = c.out((n, 2))
    c.run(c.inp(a), c.inp(b), po, C.c_size_t(n))
    return out


def gtp_pow(a, e):
    c = _Call("pb_gtp_pow", a, e)
    n = _n(a)
    out, po = c.out((n, 2))
    c.run(c.inp(a), c.inp(e, np.uint64), po, C.c_size_t(n))
    return out


def line(a, b):
    c = _Call("pb_line", a, b)
    n = _n(a)
    out, po = c.out((n, 3))
    c.run(c.inp(a), c.inp(b), po, C.c_size_t(n))
    return out


def pairing(p, q):
    c = _Call("pb_pairing", p, q)
    n = _n(p)
    out, po = c.out((n, 2))
    c.run(c.inp(p), c.inp(q), po, C.c_size_t(n))
    return out


def pairing_f(r, p, q):
    c = _Call("pb_pairing_f", p, q)
    n = _n(p)
    out, po = c.out((n, 2))
    c.run(C.c_uint64(r), c.inp(p), c.inp(q), po, C.c_size_t(n))
    return out


# ---------------------------------------------------------------- context: plonk_new + circuit constants
class Plonk:
    """One circuit + one SRS on one device: the batch counterpart of the reference's PLONK struct
    (plonk.h:43-51) together with CONSTRAINTS (constraints.h:35-47)."""

    def __init__(self, circuit, srs_g1s, srs_g2, device=0):
        circuit = np.ascontiguousarray(circuit, dtype=np.uint8)
        g1s = np.ascontiguousarray(srs_g1s, dtype=np.uint8)
        g2 = np.ascontiguousarray(srs_g2, dtype=np.uint8)
        assert circuit.size == 44 and g1s.ndim == 2 and g1s.shape[1] == 3 and g2.size == 4
        self.srs_len = int(g1s.shape[0])
        self.device = device
        self._h = C.c_void_p()
        fn = lib().pb_ctx_create
        fn.restype = C.c_int
        _check(fn(C.byref(self._h), C.c_int(device), circuit.ctypes.data_as(u8p), g1s.ctypes.data_as(u8p),
                  C.c_uint32(self.srs_len), g2.ctypes.data_as(u8p)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().pb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setup_dump(self):
        out = np.zeros(37, np.uint8)
        _check(lib().pb_ctx_setup_dump(self._h, out.ctypes.data_as(u8p)))
        return dict(h=out[0:4].copy(), k1_h=out[4:8].copy(), k2_h=out[8:12].copy(),
                    h_pows_inv=out[12:28].reshape(4, 4).copy(), z_h=out[28:28 + out[36]].copy())

    def circuit_dump(self):
        out = np.zeros(48, np.uint8)
        _check(lib().pb_ctx_circuit_dump(self._h, out.ctypes.data_as(u8p)))
        return dict(sigma=out[0:12].reshape(3, 4).copy(), s_sigma=out[12:24].reshape(3, 4).copy(),
                    q_polys=out[24:44].reshape(5, 4).copy(), l1=out[44:48].copy())

    def verifier_key(self):
        out = np.zeros(27, np.uint8)
        _check(lib().pb_ctx_verifier_key(self._h, out.ctypes.data_as(u8p)))
        return out.reshape(9, 3)

    def srs_table(self):
        out = np.zeros((self.srs_len, 17, 3), np.uint8)
        _check(lib().pb_ctx_srs_table(self._h, out.ctypes.data_as(u8p)))
        return out

    def interpolate_at_h(self, vals):
        c = _Call("pb_interpolate_at_h", vals)
        n = _n(vals)
        out, po = c.out((n, 4))
        olen, pl = c.out((n,))
        c.run(self._h, c.inp(vals), po, pl, C.c_size_t(n))
        return out, olen

    def poly_divide_zh(self, num, nlen):
        """poly_divide(num, Z_H) with the context's Z_H = x^4 - 1 (plonk.h:505) -> (quot, qlen, rem, rlen, status)."""
        c = _Call("pb_poly_divide_zh", num)
        n, sn = _n(num), int(num.shape[1])
        quot, pq = c.out((n, sn - 4))
        qlen, pql = c.out((n,))
        rem, pr = c.out((n, 4))
        rlen, prl = c.out((n,))
        status, ps = c.out((n,))
        c.run(self._h, c.inp(num), c.inp(nlen), C.c_size_t(sn), pq, pql, pr, prl, ps, C.c_size_t(n))
        return quot, qlen, rem, rlen, status

    def config2_items_into(self, a, b, x, vals, outs):
        """Host path with caller-owned (pinned) numpy buffers: outs = (prod[n][11], prod_len[n], quot[n][7], quot_len[n],
        rem[n][4], rem_len[n], evals[n], interp[n][4], interp_len[n])."""
        n = _n(a)
        shapes = ((n, 11), (n,), (n, 7), (n,), (n, 4), (n,), (n,), (n, 4), (n,))
        c = _Call("pb_config2_items", a, b, x, vals)
        assert not c.dev
        c.run(self._h, c.inp(a), c.inp(b), c.inp(x), c.inp(vals), *(c.outbuf(o, sh) for o, sh in zip(outs, shapes)), C.c_size_t(n))

    def config2_items(self, a, b, x, vals):
        """BASELINE config 2 in one launch: returns (prod, prod_len, quot, quot_len, rem, rem_len, evals, interp, interp_len),
        byte-identical to poly_mul / poly_divide(., Z_H) / poly_eval / interpolate_at_h.  torch CUDA tensors -> device path;
        numpy arrays -> the pipelined host path."""
        if not _is_torch(a):
            n = _n(a)
            outs = tuple(np.empty(sh, np.uint8) for sh in ((n, 11), (n,), (n, 7), (n,), (n, 4), (n,), (n,), (n, 4), (n,)))
            self.config2_items_into(a, b, x, vals, outs)
            return outs
        import torch
        n = _n(a)
        dev = a.device
        mk = lambda *shape: torch.empty(shape, dtype=torch.uint8, device=dev)
        outs = (mk(n, 11), mk(n), mk(n, 7), mk(n), mk(n, 4), mk(n), mk(n), mk(n, 4), mk(n))
        with torch.cuda.device(dev):
            fn = lib().pb_config2_items_dev
            fn.restype = C.c_int
            _check(fn(self._h, *(C.c_void_p(t.data_ptr()) for t in (a, b, x, vals)), *(C.c_void_p(t.data_ptr()) for t in outs),
                      C.c_size_t(n), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return outs

    def srs_eval_at_s(self, polys, plen):
        c = _Call("pb_srs_eval_at_s", polys)
        n = _n(polys)
        out, po = c.out((n, 3))
        status, ps = c.out((n,))
        c.run(self._h, c.inp(polys), c.inp(plen), C.c_size_t(int(polys.shape[1])), po, ps, C.c_size_t(n))
        return out, status

    def constraints_satisfy(self, witness):
        c = _Call("pb_constraints_satisfy", witness)
        n = _n(witness)
        out, po = c.out((n,))
        c.run(self._h, c.inp(witness), po, C.c_size_t(n))
        return out

    def prove(self, witness, rnd, chal):
        """plonk_prove over a batch -> (proofs[n][34], status[n])."""
        c = _Call("pb_plonk_prove", witness, rnd, chal)
        n = _n(witness)
        proofs, pp = c.out((n, 34))
        status, ps = c.out((n,))
        c.run(self._h, c.inp(witness), c.inp(rnd), c.inp(chal), pp, ps, C.c_size_t(n))
        return proofs, status

    def verify(self, proofs, chal, u, want_gt=False):
        c = _Call("pb_plonk_verify", proofs, chal, u)
        n = _n(proofs)
        verdict, pv = c.out((n,))
        if want_gt:
            gt, pg = c.out((n, 4))
        else:
            gt, pg = None, None
        c.run(self._h, c.inp(proofs), c.inp(chal), c.inp(u), pv, pg, C.c_size_t(n))
        return (verdict, gt) if want_gt else verdict

    def prove_verify(self, witness, rnd, chal, u):
        """prove, then verify every completed proof -> (proofs, status, verdict); verdict 0xFF where status != 0."""
        c = _Call("pb_plonk_prove_verify", witness, rnd, chal, u)
        n = _n(witness)
        proofs, pp = c.out((n, 34))
        status, ps = c.out((n,))
        verdict, pv = c.out((n,))
        c.run(self._h, c.inp(witness), c.inp(rnd), c.inp(chal), c.inp(u), pp, ps, pv, C.c_size_t(n))
        return proofs, status, verdict

    def prove_verify_tally_dev(self, witness, rnd, chal, u, counts):
        """Device path (torch CUDA tensors): prove, verify, and counts (torch int64[18], CUDA) += the batch's counters, in
        one call -> (proofs, status, verdict)."""
        import torch
        n = _n(witness)
        dev = witness.device
        proofs = torch.empty((n, 34), dtype=torch.uint8, device=dev)
        status = torch.empty(n, dtype=torch.uint8, device=dev)
        verdict = torch.empty(n, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _check(lib().pb_plonk_prove_verify_tally_dev(
                self._h, *(C.c_void_p(t.data_ptr()) for t in (witness, rnd, chal, u, proofs, status, verdict, counts)),
                C.c_size_t(n), C.c_void_p(torch.cuda.current_stream().cuda_stream), None))
        return proofs, status, verdict

    # ---- Fiat-Shamir mode (include/plonk_b200.h; specification oracle/fs_spec.inc)
    def fs_seed(self):
        """the transcript state after absorbing circuit and SRS, as one int: v0 | v1 << 32 | v2 << 64 | v3 << 96"""
        out = np.zeros(4, np.uint32)
        _check(lib().pb_ctx_fs_seed(self._h, out.ctypes.data_as(C.c_void_p)))
        return sum(int(w) << (32 * k) for k, w in enumerate(out))

    def prove_fs(self, witness, rnd, want_challenges=False):
        """-> (proofs[n][34], status[n]) and, on request, chal[n][6] = alpha beta gamma z v u (0xFF where not drawn)."""
        c = _Call("pb_plonk_prove_fs", witness, rnd)
        n = _n(witness)
        proofs, pp = c.out((n, 34))
        status, ps = c.out((n,))
        chal, pc = c.out((n, 6)) if want_challenges else (None, None)
        c.run(self._h, c.inp(witness), c.inp(rnd), pp, ps, pc, C.c_size_t(n))
        return (proofs, status, chal) if want_challenges else (proofs, status)

    def verify_fs(self, proofs, want_gt=False):
        c = _Call("pb_plonk_verify_fs", proofs)
        n = _n(proofs)
        verdict, pv = c.out((n,))
        gt, pg = c.out((n, 4)) if want_gt else (None, None)
        c.run(self._h, c.inp(proofs), pv, pg, C.c_size_t(n))
        return (verdict, gt) if want_gt else verdict

    def prove_verify_fs(self, witness, rnd):
        c = _Call("pb_plonk_prove_verify_fs", witness, rnd)
        n = _n(witness)
        proofs, pp = c.out((n, 34))
        status, ps = c.out((n,))
        verdict, pv = c.out((n,))
        c.run(self._h, c.inp(witness), c.inp(rnd), pp, ps, pv, C.c_size_t(n), dev_tail=(None,))
        return proofs, status, verdict

    def prove_verify_fs_into(self, witness, rnd, proofs, status, verdict, mid_event=None):
        c = _Call("pb_plonk_prove_verify_fs", witness, rnd, proofs, status, verdict)
        n = _n(witness)
        c.run(self._h, c.inp(witness), c.inp(rnd), c.outbuf(proofs, (n, 34)), c.outbuf(status, (n,)), c.outbuf(verdict, (n,)),
              C.c_size_t(n), dev_tail=(mid_event,))

    def fs_challenges(self, proofs):
        """the six challenges of each PROOF record as a verifier derives them -> chal[n][6]"""
        c = _Call("pb_fs_challenges", proofs)
        n = _n(proofs)
        chal, pc = c.out((n, 6))
        c.run(self._h, c.inp(proofs), pc, C.c_size_t(n))
        return chal

    def prove_verify_into(self, witness, rnd, chal, u, proofs, status, verdict):
        """Same with caller-owned buffers (numpy arrays over pinned memory, or torch CUDA tensors): no allocation
        on the call path.  This is what bench.py times."""
        c = _Call("pb_plonk_prove_verify", witness, rnd, chal, u, proofs, status, verdict)
        n = _n(witness)
        c.run(self._h, c.inp(witness), c.inp(rnd), c.inp(chal), c.inp(u), c.outbuf(proofs, (n, 34)), c.outbuf(status, (n,)),
              c.outbuf(verdict, (n,)), C.c_size_t(n))

    # ---- outputs that cost less PCIe (include/plonk_b200.h: compact output, packed wire v2, seeded mode)
    def prove_verify_compact_into(self, witness, rnd, chal, u, proofs_dense, status, verdict):
        """Host path (numpy).  proofs_dense: caller-owned [n][34] capacity; returns the number of completed proofs, whose
        PROOF structs are proofs_dense[:n_done] in item order."""
        c = _Call("pb_plonk_prove_verify_compact", witness, rnd, chal, u)
        assert not c.dev
        n = _n(witness)
        n_done = C.c_size_t(0)
        c.run(self._h, c.inp(witness), c.inp(rnd), c.inp(chal), c.inp(u), c.outbuf(proofs_dense, (n, 34)), C.byref(n_done),
              c.outbuf(status, (n,)), c.outbuf(verdict, (n,)), C.c_size_t(n))
        return int(n_done.value)

    def prove_verify_compact(self, witness, rnd, chal, u):
        n = _n(witness)
        dense, status, verdict = np.empty((n, 34), np.uint8), np.empty(n, np.uint8), np.empty(n, np.uint8)
        k = self.prove_verify_compact_into(witness, rnd, chal, u, dense, status, verdict)
        return dense[:k], status, verdict

    def prove_verify_packed_into(self, packed_in, packed_proofs, sv):
        """Host path (numpy), packed wire v2: packed_in [n][16] -> packed_proofs[:n_done] ([n][22] capacity), sv[n]."""
        c = _Call("pb_plonk_prove_verify_packed", packed_in)
        assert not c.dev
        n = _n(packed_in)
        n_done = C.c_size_t(0)
        c.run(self._h, c.inp(packed_in), c.outbuf(packed_proofs, (n, 22)), C.byref(n_done), c.outbuf(sv, (n,)), C.c_size_t(n))
        return int(n_done.value)

    def prove_verify_packed(self, packed_in):
        n = _n(packed_in)
        pp, sv = np.empty((n, 22), np.uint8), np.empty(n, np.uint8)
        k = self.prove_verify_packed_into(packed_in, pp, sv)
        return pp[:k], sv

    def prove_verify_packed_dev(self, packed_in, v3=False):
        """Device path (torch CUDA tensors): -> (packed_proofs [n][22] (v3: [n][12]) capacity, n_done (device int32[1]), sv[n])."""
        import torch
        n = _n(packed_in)
        dev = packed_in.device
        lib().pb_packed_workspace_bytes.restype = C.c_size_t
        ws = torch.empty(lib().pb_packed_workspace_bytes(C.c_size_t(n)), dtype=torch.uint8, device=dev)
        pp = torch.empty((n, 12 if v3 else 22), dtype=torch.uint8, device=dev)
        sv = torch.empty(n, dtype=torch.uint8, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        fn = lib().pb_plonk_prove_verify_packed3_dev if v3 else lib().pb_plonk_prove_verify_packed_dev
        with torch.cuda.device(dev):
            _check(fn(self._h, C.c_void_p(packed_in.data_ptr()), C.c_void_p(pp.data_ptr()),
                      C.c_void_p(cnt.data_ptr()), C.c_void_p(sv.data_ptr()), C.c_void_p(ws.data_ptr()),
                      C.c_size_t(n), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return pp, cnt, sv

    def prove_verify_packed3_into(self, packed_in, packed_proofs, sv):
        """Host path (numpy), packed wire v3: packed_in [n][14] -> packed_proofs[:n_done] ([n][12] capacity), sv[n].
        Only for an SRS whose points are canonical points of the curve (PlonkError otherwise)."""
        c = _Call("pb_plonk_prove_verify_packed3", packed_in)
        assert not c.dev
        n = _n(packed_in)
        n_done = C.c_size_t(0)
        c.run(self._h, c.inp(packed_in), c.outbuf(packed_proofs, (n, 12)), C.byref(n_done), c.outbuf(sv, (n,)), C.c_size_t(n))
        return int(n_done.value)

    def prove_verify_packed3(self, packed_in):
        n = _n(packed_in)
        pp, sv = np.empty((n, 12), np.uint8), np.empty(n, np.uint8)
        k = self.prove_verify_packed3_into(packed_in, pp, sv)
        return pp[:k], sv

    def prove_verify_seeded(self, seed, start, count, variant="U17"):
        """Inputs generated on the device from (seed, start, count); only the 18 counters come back."""
        counts = np.zeros(18, np.int64)
        _check(lib().pb_plonk_prove_verify_seeded(self._h, C.c_uint64(seed), C.c_uint64(start), C.c_size_t(count),
                                                  C.c_int(VARIANTS[variant]), counts.ctypes.data_as(C.c_void_p)))
        return counts

    def seeded_workspace(self, n, device):
        import torch
        lib().pb_seeded_workspace_bytes.restype = C.c_size_t
        return torch.empty(lib().pb_seeded_workspace_bytes(C.c_size_t(n)), dtype=torch.uint8, device=device)

    def prove_verify_seeded_dev(self, seed, start, n, variant, workspace, counts):
        """Enqueue on torch's current stream: counts (torch int64[18], CUDA) += the tally of items [start, start + n)."""
        import torch
        with torch.cuda.device(counts.device):
            _check(lib().pb_plonk_prove_verify_seeded_dev(self._h, C.c_uint64(seed), C.c_uint64(start), C.c_size_t(n),
                                                          C.c_int(VARIANTS[variant]), C.c_void_p(workspace.data_ptr()),
                                                          C.c_void_p(counts.data_ptr()),
                                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def synth_batch(self, seed, start, n, variant="U17", device=None):
        """The synthetic stream generated on the device: -> (witness, rnd, chal, u, packed) as torch CUDA tensors."""
        import torch
        dev = torch.device("cuda", self.device) if device is None else device
        mk = lambda *shape: torch.empty(shape, dtype=torch.uint8, device=dev)
        wit, rnd, chal, u, packed = mk(n, 12), mk(n, 9), mk(n, 5), mk(n), mk(n, 16)
        with torch.cuda.device(dev):
            _check(lib().pb_synth_batch_dev(self._h, C.c_uint64(seed), C.c_uint64(start), C.c_size_t(n), C.c_int(VARIANTS[variant]),
                                            *(C.c_void_p(t.data_ptr()) for t in (wit, rnd, chal, u, packed)),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return wit, rnd, chal, u, packed


VARIANTS = {"U17": 0, "NZ": 1}


def gather_completed(proofs, status):
    """torch CUDA tensors proofs[n][34], status[n] -> (dense [n][34] capacity, n_done device int32[1]): the PROOF structs of
    the completed items (status 0) in item order."""
    import torch
    n = _n(proofs)
    dev = proofs.device
    dense = torch.empty((n, 34), dtype=torch.uint8, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    offs = torch.empty(n // 128 + 4, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _check(lib().pb_gather_completed_dev(C.c_void_p(proofs.data_ptr()), C.c_void_p(status.data_ptr()), C.c_void_p(dense.data_ptr()),
                                             C.c_void_p(cnt.data_ptr()), C.c_void_p(offs.data_ptr()), C.c_size_t(n),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return dense, cnt


# ---- packed wire v2: the C helpers (host-side format conversion; plonk.c_b200/wire.py is the numpy twin)
def wire_pack_inputs(witness, rnd, chal, u):
    n = _n(witness)
    out = np.empty((n, 16), np.uint8)
    a = [np.ascontiguousarray(x, np.uint8) for x in (witness, rnd, chal, u)]
    _check(lib().pb_wire_pack_inputs(*(x.ctypes.data_as(C.c_void_p) for x in a), out.ctypes.data_as(C.c_void_p), C.c_size_t(n)))
    return out


def wire_unpack_inputs(packed):
    packed = np.ascontiguousarray(packed, np.uint8)
    n = _n(packed)
    wit, rnd, chal, u, valid = np.empty((n, 12), np.uint8), np.empty((n, 9), np.uint8), np.empty((n, 5), np.uint8), np.empty(n, np.uint8), np.empty(n, np.uint8)
    _check(lib().pb_wire_unpack_inputs(packed.ctypes.data_as(C.c_void_p), *(x.ctypes.data_as(C.c_void_p) for x in (wit, rnd, chal, u, valid)),
                                       C.c_size_t(n)))
    return wit, rnd, chal, u, valid.astype(bool)


def wire_pack_proofs(proofs):
    proofs = np.ascontiguousarray(proofs, np.uint8).reshape(-1, 34)
    out = np.empty((proofs.shape[0], 22), np.uint8)
    _check(lib().pb_wire_pack_proofs(proofs.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_size_t(proofs.shape[0])))
    return out


def wire_unpack_proofs(packed):
    packed = np.ascontiguousarray(packed, np.uint8).reshape(-1, 22)
    out = np.empty((packed.shape[0], 34), np.uint8)
    _check(lib().pb_wire_unpack_proofs(packed.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_size_t(packed.shape[0])))
    return out


def wire3_pack_inputs(witness, rnd, chal, u):
    n = _n(witness)
    out = np.empty((n, 14), np.uint8)
    a = [np.ascontiguousarray(x, np.uint8) for x in (witness, rnd, chal, u)]
    _check(lib().pb_wire3_pack_inputs(*(x.ctypes.data_as(C.c_void_p) for x in a), out.ctypes.data_as(C.c_void_p), C.c_size_t(n)))
    return out


def wire3_unpack_inputs(packed):
    packed = np.ascontiguousarray(packed, np.uint8).reshape(-1, 14)
    n = packed.shape[0]
    wit, rnd, chal, u, valid = np.empty((n, 12), np.uint8), np.empty((n, 9), np.uint8), np.empty((n, 5), np.uint8), np.empty(n, np.uint8), np.empty(n, np.uint8)
    _check(lib().pb_wire3_unpack_inputs(packed.ctypes.data_as(C.c_void_p), *(x.ctypes.data_as(C.c_void_p) for x in (wit, rnd, chal, u, valid)),
                                        C.c_size_t(n)))
    return wit, rnd, chal, u, valid.astype(bool)


def wire3_pack_proofs(proofs):
    proofs = np.ascontiguousarray(proofs, np.uint8).reshape(-1, 34)
    out = np.empty((proofs.shape[0], 12), np.uint8)
    _check(lib().pb_wire3_pack_proofs(proofs.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_size_t(proofs.shape[0])))
    return out


def wire3_unpack_proofs(packed):
    packed = np.ascontiguousarray(packed, np.uint8).reshape(-1, 12)
    out = np.empty((packed.shape[0], 34), np.uint8)
    _check(lib().pb_wire3_unpack_proofs(packed.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_size_t(packed.shape[0])))
    return out


def wire_scatter_proofs(proofs_dense, status):
    status = np.ascontiguousarray(status, np.uint8)
    dense = np.ascontiguousarray(proofs_dense, np.uint8)
    out = np.empty((status.shape[0], 34), np.uint8)
    _check(lib().pb_wire_scatter_proofs(dense.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                        C.c_size_t(status.shape[0])))
    return out


def wire_split_sv(sv):
    sv = np.ascontiguousarray(sv, np.uint8)
    st, vd = np.empty_like(sv), np.empty_like(sv)
    _check(lib().pb_wire_split_sv(sv.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p), vd.ctypes.data_as(C.c_void_p), C.c_size_t(sv.shape[0])))
    return st, vd


def srs_eval_at_s_raw(srs_g1s, polys, plen, trim=True):
    """srs_eval_at_s against an SRS passed per call (no context): -> (points[n][3], status[n])"""
    g = np.ascontiguousarray(srs_g1s, np.uint8)
    c = _Call("pb_srs_eval_at_s_raw", polys)
    n = _n(polys)
    out, po = c.out((n, 3))
    status, ps = c.out((n,))
    if c.dev:
        g = c.torch.as_tensor(g, device=c.device)
        c.keep.append(g)
        pg = C.c_void_p(g.data_ptr())
    else:
        pg = g.ctypes.data_as(C.c_void_p)
    c.run(pg, C.c_uint32(g.shape[0]), c.inp(polys), c.inp(plen), C.c_size_t(int(polys.shape[1])), C.c_int(1 if trim else 0), po, ps, C.c_size_t(n))
    return out, status


def constraints_satisfy_rows(selectors, a, b, c_):
    """selectors[5][rows], a / b / c [n][rows] (numpy) -> first failing row per item (int32, -1 = satisfied)"""
    q = np.ascontiguousarray(selectors, np.uint8)
    a, b, c_ = (np.ascontiguousarray(v, np.uint8) for v in (a, b, c_))
    n, rows = a.shape
    out = np.empty(n, np.int32)
    _check(lib().pb_constraints_satisfy_rows(q.ctypes.data_as(C.c_void_p), C.c_uint32(rows), a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                             c_.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_size_t(n)))
    return out


def tally(proofs, status, verdict, counts):
    """counts (torch int64[18], CUDA) += per-status histogram, accept count, proof-byte checksum."""
    import torch
    with torch.cuda.device(counts.device):
        fn = lib().pb_tally_dev
        fn.restype = C.c_int
        n = int(status.shape[0])
        _check(fn(C.c_void_p(proofs.data_ptr()) if proofs is not None else None, C.c_void_p(status.data_ptr()),
                  C.c_void_p(verdict.data_ptr()) if verdict is not None else None, C.c_size_t(n),
                  C.c_void_p(counts.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
