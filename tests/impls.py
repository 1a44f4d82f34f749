"""Adapters that give every implementation under test the call signatures of oracle._binding.OracleLib, so one
parity suite (parity_suite.py) runs against: the C restatement, the host-compiled kernel source (hostcheck),
and the CUDA library through its C ABI (host-pointer path and device-pointer path)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)


def _p(a):
    return a.ctypes.data_as(u8p)


class GpuImpl:
    """libplonk_b200.so through plonk_c_b200.host.  path='host': numpy in/out (host-pointer entry points);
    path='device': torch CUDA tensors (the `_dev` entry points on torch's current stream)."""
    kind = "cuda"

    def __init__(self, host, path="device"):
        self.h = host
        self.path = path
        self._ctx = {}
        if path == "device":
            import torch
            self.torch = torch

    def _in(self, x, dtype=np.uint8):
        if x is None:
            return None
        x = np.ascontiguousarray(x, dtype=dtype)
        if self.path == "device":
            if dtype == np.uint64:
                return self.torch.from_numpy(x.view(np.int64)).cuda()
            return self.torch.from_numpy(x).cuda()
        return x

    def _out(self, x):
        if isinstance(x, tuple):
            return tuple(self._out(y) for y in x)
        if x is None:
            return None
        if self.path == "device":
            return x.cpu().numpy()
        return x

    def ctx(self, circuit, g1s, g2):
        key = (bytes(np.asarray(circuit, np.uint8)), bytes(np.asarray(g1s, np.uint8)), bytes(np.asarray(g2, np.uint8)))
        if key not in self._ctx:
            self._ctx[key] = self.h.Plonk(circuit, g1s, g2, device=0)
        return self._ctx[key]

    def field_op(self, field, op, a, b=None):
        return self._out(self.h.field_op(field, op, self._in(a), self._in(b)))

    def poly_binop(self, op, a, alen, b, blen, so):
        return self._out(self.h.poly_binop(op, self._in(a), self._in(alen), self._in(b), self._in(blen), so))

    def poly_divide(self, num, nlen, den, dlen, sq, sr):
        return self._out(self.h.poly_divide(self._in(num), self._in(nlen), self._in(den), self._in(dlen), sq, sr))

    def poly_eval(self, p, plen, x):
        return self._out(self.h.poly_eval(self._in(p), self._in(plen), self._in(x)))

    def poly_unop(self, op, p, plen, k, so):
        return self._out(self.h.poly_unop(op, self._in(p), self._in(plen), self._in(k), so))

    def poly_slice(self, p, plen, start, end, so):
        return self._out(self.h.poly_slice(self._in(p), self._in(plen), self._in(start), self._in(end), so))

    def poly_lagrange(self, xs, ys, so):
        return self._out(self.h.poly_lagrange(self._in(xs), self._in(ys), so))

    def matrix_mul(self, a, b):
        return self._out(self.h.matrix_mul(self._in(a[None]), self._in(b[None])))[0]

    def matrix_inv(self, a):
        return self._out(self.h.matrix_inv(self._in(a[None])))[0]

    def g1_op(self, op, a, b=None):
        return self._out(self.h.g1_op(op, self._in(a), self._in(b)))

    def g1_mul(self, p, s, nthreads=1):
        s = np.asarray(s)
        return self._out(self.h.g1_mul(self._in(p), self._in(s, s.dtype if s.dtype == np.uint8 else np.uint64)))

    def g1_is_on_curve(self, p):
        return self._out(self.h.g1_is_on_curve(self._in(p)))

    def g2_op(self, op, a, b=None):
        return self._out(self.h.g2_op(op, self._in(a), self._in(b)))

    def g2_mul(self, p, s):
        return self._out(self.h.g2_mul(self._in(p), self._in(s, np.uint64)))

    def gtp_mul(self, a, b):
        return self._out(self.h.gtp_mul(self._in(a), self._in(b)))

    def gtp_pow(self, a, e):
        return self._out(self.h.gtp_pow(self._in(a), self._in(e, np.uint64)))

    def line(self, a, b):
        return self._out(self.h.line(self._in(a), self._in(b)))

    def pairing(self, p, q, nthreads=1):
        return self._out(self.h.pairing(self._in(p), self._in(q)))

    def pairing_f(self, r, p, q):
        return self._out(self.h.pairing_f(r, self._in(p), self._in(q)))

    def srs_eval_at_s(self, g1s, g2, polys, plen, nthreads=1):
        return self._out(self.ctx(DEFAULT_CIRCUIT(), g1s, g2).srs_eval_at_s(self._in(polys), self._in(plen)))

    def plonk_setup_dump(self):
        from plonk_c_b200 import workload as W
        return self.ctx(DEFAULT_CIRCUIT(), *W.identity_srs(6)).setup_dump()

    def interpolate_at_h(self, vals):
        from plonk_c_b200 import workload as W
        return self._out(self.ctx(DEFAULT_CIRCUIT(), *W.identity_srs(6)).interpolate_at_h(self._in(vals)))

    def verifier_key(self, circuit, g1s, g2):
        return self.ctx(circuit, g1s, g2).verifier_key()

    def plonk_prove_batch(self, circuit, g1s, g2, wit, rnd, chal, nthreads=1):
        return self._out(self.ctx(circuit, g1s, g2).prove(self._in(wit), self._in(rnd), self._in(chal)))

    def plonk_verify_batch(self, circuit, g1s, g2, proofs, chal, u, nthreads=1, want_gt=True):
        return self._out(self.ctx(circuit, g1s, g2).verify(self._in(proofs), self._in(chal), self._in(u), want_gt=True))

    def plonk_verdict_only(self, circuit, g1s, g2, proofs, chal, u):
        """verdicts without the two GT values (the fast verifier then shares one final exponentiation)"""
        return self._out(self.ctx(circuit, g1s, g2).verify(self._in(proofs), self._in(chal), self._in(u), want_gt=False))

    # Fiat-Shamir mode
    def plonk_prove_fs_batch(self, circuit, g1s, g2, wit, rnd, nthreads=1):
        return self._out(self.ctx(circuit, g1s, g2).prove_fs(self._in(wit), self._in(rnd), want_challenges=True))

    def plonk_verify_fs_batch(self, circuit, g1s, g2, proofs, want_gt=True):
        return self._out(self.ctx(circuit, g1s, g2).verify_fs(self._in(proofs), want_gt=True))

    def plonk_prove_verify_fs_batch(self, circuit, g1s, g2, wit, rnd):
        return self._out(self.ctx(circuit, g1s, g2).prove_verify_fs(self._in(wit), self._in(rnd)))

    def fs_seed(self, circuit, g1s, g2):
        return self.ctx(circuit, g1s, g2).fs_seed()

    def fs_challenges(self, circuit, g1s, g2, proofs):
        return self._out(self.ctx(circuit, g1s, g2).fs_challenges(self._in(proofs)))


def DEFAULT_CIRCUIT():
    from plonk_c_b200 import workload as W
    return W.PLONK_TEST_CIRCUIT


class HostcheckImpl:
    """tests/hostcheck/libpb_hostcheck.so: field.cuh / curve.cuh / prover.cuh / verifier.cuh compiled for the host.
    Per-circuit constants and tables are taken from `setup` (an oracle), as cabi.cu takes them from its setup kernels."""
    kind = "hostcheck"

    def __init__(self, setup_oracle, fast=False):
        self.fast = fast          # fast=True: pair-table prover + joint double-and-add verifier (canonical on-curve SRS only)
        path = os.path.join(HERE, "hostcheck", "libpb_hostcheck.so")
        if not os.path.exists(path):
            import __graft_entry__ as g
            g.build_hostcheck()
        self.lib = C.CDLL(path)
        self.o = setup_oracle

    def g1_op(self, op, a, b=None):
        a = np.ascontiguousarray(a, np.uint8)
        b = a if b is None else np.ascontiguousarray(b, np.uint8)
        out = np.zeros_like(a)
        self.lib.hc_g1_op(op, _p(a), _p(b), _p(out), C.c_size_t(a.shape[0]))
        return out

    def g1_mul(self, p, s, nthreads=1):
        p = np.ascontiguousarray(p, np.uint8)
        s = np.ascontiguousarray(s, np.uint64)
        out = np.zeros_like(p)
        self.lib.hc_g1_mul(_p(p), s.ctypes.data_as(u64p), _p(out), C.c_size_t(p.shape[0]))
        return out

    def g1_is_on_curve(self, p):
        p = np.ascontiguousarray(p, np.uint8)
        out = np.zeros(p.shape[0], np.uint8)
        self.lib.hc_g1_on_curve(_p(p), _p(out), C.c_size_t(p.shape[0]))
        return out

    def g2_op(self, op, a, b=None):
        a = np.ascontiguousarray(a, np.uint8)
        b = a if b is None else np.ascontiguousarray(b, np.uint8)
        out = np.zeros_like(a)
        self.lib.hc_g2_op(op, _p(a), _p(b), _p(out), C.c_size_t(a.shape[0]))
        return out

    def g2_mul(self, p, s):
        p = np.ascontiguousarray(p, np.uint8)
        s = np.ascontiguousarray(s, np.uint64)
        out = np.zeros_like(p)
        self.lib.hc_g2_mul(_p(p), s.ctypes.data_as(u64p), _p(out), C.c_size_t(p.shape[0]))
        return out

    def gtp_mul(self, a, b):
        a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
        out = np.zeros_like(a)
        self.lib.hc_gtp_mul(_p(a), _p(b), _p(out), C.c_size_t(a.shape[0]))
        return out

    def gtp_pow(self, a, e):
        a = np.ascontiguousarray(a, np.uint8)
        e = np.ascontiguousarray(e, np.uint64)
        out = np.zeros_like(a)
        self.lib.hc_gtp_pow(_p(a), e.ctypes.data_as(u64p), _p(out), C.c_size_t(a.shape[0]))
        return out

    def line(self, a, b):
        a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
        out = np.zeros((a.shape[0], 3), np.uint8)
        self.lib.hc_line(_p(a), _p(b), _p(out), C.c_size_t(a.shape[0]))
        return out

    def pairing(self, p, q, nthreads=1):
        p, q = np.ascontiguousarray(p, np.uint8), np.ascontiguousarray(q, np.uint8)
        out = np.zeros((p.shape[0], 2), np.uint8)
        self.lib.hc_pairing(_p(p), _p(q), _p(out), C.c_size_t(p.shape[0]))
        return out

    def pairing_f(self, r, p, q):
        p, q = np.ascontiguousarray(p, np.uint8), np.ascontiguousarray(q, np.uint8)
        out = np.zeros((p.shape[0], 2), np.uint8)
        self.lib.hc_pairing_f(C.c_uint64(r), _p(p), _p(q), _p(out), C.c_size_t(p.shape[0]))
        return out

    def fs_seed(self, circuit, g1s, g2):
        circuit, g1s, g2 = (np.ascontiguousarray(x, np.uint8) for x in (circuit, g1s, g2))
        out = np.zeros(4, np.uint32)
        self.lib.hc_fs_seed(_p(circuit), _p(g1s), C.c_uint32(g1s.shape[0]), _p(g2), out.ctypes.data_as(C.c_void_p))
        return sum(int(w) << (32 * k) for k, w in enumerate(out))

    @staticmethod
    def _seed_words(seed):
        return np.array([(int(seed) >> (32 * k)) & 0xFFFFFFFF for k in range(4)], np.uint32)

    def fs_challenges(self, circuit, g1s, g2, proofs):
        proofs = np.ascontiguousarray(proofs, np.uint8)
        out = np.zeros((proofs.shape[0], 6), np.uint8)
        sw = self._seed_words(self.fs_seed(circuit, g1s, g2))
        self.lib.hc_fs_challenges(sw.ctypes.data_as(C.c_void_p), _p(proofs), _p(out), C.c_size_t(proofs.shape[0]))
        return out

    def _prove_fn(self):
        # fast: False = sequential tables (any SRS); True = pair tables; "wide" = one-look-up T6 tables (prover only)
        return {False: self.lib.hc_prove, True: self.lib.hc_prove_pairs, "wide": self.lib.hc_prove_wide, "log": self.lib.hc_prove_pairs}[self.fast]

    def wide_tables(self, g1s):
        """(T6 [17^6][3], T3 [17^3][3]) as the wide-table builder produces them for this SRS"""
        tb = self._table(np.ascontiguousarray(g1s, np.uint8))
        t6 = np.zeros((17 ** 6, 3), np.uint8)
        t3 = np.zeros((17 ** 3, 3), np.uint8)
        self.lib.hc_wide_tables(tb.ctypes.data_as(C.c_void_p), _p(t6), _p(t3))
        return t6, t3

    def _cc_words(self, circuit, srs_len, fs_seed=0):
        o = self.o
        circuit = np.asarray(circuit, np.uint8)
        qv = circuit[:20].reshape(5, 4)
        QP, _ = o.interpolate_at_h(qv)
        sig = np.stack([o.copy_constraints_to_roots(circuit[20 + 8 * s:24 + 8 * s], circuit[24 + 8 * s:28 + 8 * s]) for s in range(3)])
        SP, _ = o.interpolate_at_h(sig)
        vinv = o.plonk_setup_dump()["h_pows_inv"]
        l1, _ = o.interpolate_at_h(np.array([[1, 0, 0, 0]], np.uint8))
        w = np.concatenate([qv.ravel(), QP.ravel(), sig.ravel(), SP.ravel(), vinv.ravel(), l1.ravel(), [srs_len, 0], self._seed_words(fs_seed)]).astype(np.uint32)
        assert w.size * 4 == self.lib.hc_sizeof_cc()
        return np.ascontiguousarray(w)

    def _table(self, g1s):
        rows = []
        for i in range(9):
            if i < g1s.shape[0]:
                pts = self.o.g1_mul(np.tile(g1s[i:i + 1], (17, 1)), np.arange(17, dtype=np.uint64))
            else:
                pts = np.tile(np.array([[0, 0, 1]], np.uint8), (17, 1))
            rows.append(pts[:, 0].astype(np.uint32) | (pts[:, 1].astype(np.uint32) << 8) | (pts[:, 2].astype(np.uint32) << 16))
        return np.ascontiguousarray(np.stack(rows).astype(np.uint32))

    def plonk_prove_batch(self, circuit, g1s, g2, wit, rnd, chal, nthreads=1):
        g1s = np.ascontiguousarray(g1s, np.uint8)
        ccw, tb = self._cc_words(circuit, g1s.shape[0]), self._table(g1s)
        wit, rnd, chal = (np.ascontiguousarray(x, np.uint8) for x in (wit, rnd, chal))
        n = wit.shape[0]
        proofs, status = np.zeros((n, 34), np.uint8), np.zeros(n, np.uint8)
        self._prove_fn()(ccw.ctypes.data_as(C.c_void_p), tb.ctypes.data_as(C.c_void_p), _p(wit), _p(rnd), _p(chal),
                          _p(proofs), _p(status), None, C.c_size_t(n))
        return proofs, status

    def plonk_prove_fs_batch(self, circuit, g1s, g2, wit, rnd, nthreads=1):
        g1s = np.ascontiguousarray(g1s, np.uint8)
        ccw, tb = self._cc_words(circuit, g1s.shape[0], self.fs_seed(circuit, g1s, g2)), self._table(g1s)
        wit, rnd = (np.ascontiguousarray(x, np.uint8) for x in (wit, rnd))
        n = wit.shape[0]
        proofs, status, chal = np.zeros((n, 34), np.uint8), np.zeros(n, np.uint8), np.zeros((n, 6), np.uint8)
        self._prove_fn()(ccw.ctypes.data_as(C.c_void_p), tb.ctypes.data_as(C.c_void_p), _p(wit), _p(rnd), None,
                          _p(proofs), _p(status), _p(chal), C.c_size_t(n))
        return proofs, status, chal

    def _verify_fn(self):
        # fast=True: Straus + Miller-loop verifier; fast="log": the table path (discrete logarithms, pairing tables); else exact
        return self.lib.hc_verify_log if self.fast == "log" else self.lib.hc_verify_fast if self.fast is True else self.lib.hc_verify

    def plonk_verify_batch(self, circuit, g1s, g2, proofs, chal, u, nthreads=1, want_gt=True):
        key = np.concatenate([self.o.verifier_key(circuit, g1s, g2).ravel(), np.asarray(g2, np.uint8)]).astype(np.uint8)
        proofs, chal, u = (np.ascontiguousarray(x, np.uint8) for x in (proofs, chal, u))
        n = proofs.shape[0]
        verdict, gt = np.zeros(n, np.uint8), np.zeros((n, 4), np.uint8)
        self._verify_fn()(_p(key), self._seed_words(0).ctypes.data_as(C.c_void_p), _p(proofs), _p(chal), _p(u), _p(verdict), _p(gt), C.c_size_t(n))
        return verdict, gt

    def plonk_verdict_only(self, circuit, g1s, g2, proofs, chal, u):
        key = np.concatenate([self.o.verifier_key(circuit, g1s, g2).ravel(), np.asarray(g2, np.uint8)]).astype(np.uint8)
        proofs, chal, u = (np.ascontiguousarray(x, np.uint8) for x in (proofs, chal, u))
        n = proofs.shape[0]
        verdict = np.zeros(n, np.uint8)
        self._verify_fn()(_p(key), self._seed_words(0).ctypes.data_as(C.c_void_p), _p(proofs), _p(chal), _p(u), _p(verdict), None, C.c_size_t(n))
        return verdict

    def plonk_verify_fs_batch(self, circuit, g1s, g2, proofs, want_gt=True):
        key = np.concatenate([self.o.verifier_key(circuit, g1s, g2).ravel(), np.asarray(g2, np.uint8)]).astype(np.uint8)
        proofs = np.ascontiguousarray(proofs, np.uint8)
        n = proofs.shape[0]
        verdict, gt = np.zeros(n, np.uint8), np.zeros((n, 4), np.uint8)
        self._verify_fn()(_p(key), self._seed_words(self.fs_seed(circuit, g1s, g2)).ctypes.data_as(C.c_void_p), _p(proofs), None, None,
                                                                      _p(verdict), _p(gt), C.c_size_t(n))
        return verdict, gt
