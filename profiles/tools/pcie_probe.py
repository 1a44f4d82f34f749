import torch, time
torch.cuda.init()
n = 1 << 21
def bw(label, fn, nbytes, reps=10):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / reps
    print(f"{label:50s} {dt*1e3:8.3f} ms  {nbytes/dt/1e9:7.1f} GB/s")
hin = torch.empty(n*27, dtype=torch.uint8).pin_memory(); din = torch.empty(n*27, dtype=torch.uint8, device='cuda')
hout = torch.empty(n*36, dtype=torch.uint8).pin_memory(); dout = torch.empty(n*36, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
bw("H2D one copy 56.6MB", lambda: din.copy_(hin, non_blocking=True), n*27)
bw("D2H one copy 75.5MB", lambda: hout.copy_(dout, non_blocking=True), n*36)
def both():
    with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
    with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
bw("H2D+D2H concurrent (bytes = both)", both, n*63)
for chunks in (4, 16, 64):
    def chunked():
        c = n*27//chunks
        for k in range(chunks): din[k*c:(k+1)*c].copy_(hin[k*c:(k+1)*c], non_blocking=True)
    bw(f"H2D in {chunks} chunks", chunked, n*27)
    def chunked4():
        c = n*27//chunks//4
        for k in range(chunks*4): din[k*c:(k+1)*c].copy_(hin[k*c:(k+1)*c], non_blocking=True)
    bw(f"H2D in {chunks*4} pieces", chunked4, n*27)
# zero-copy read kernel: a simple device copy from host-mapped memory
x = torch.empty(n*27, dtype=torch.uint8, device='cuda')
import ctypes
# torch pinned memory is mapped (UVA): a device kernel can read it via the same pointer. Use torch's copy kernel through a cuda tensor alias.
try:
    from torch.utils.cpp_extension import load_inline
except Exception as e:
    print(e)
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from plonk_c_b200 import host, workload as W
lib = host.lib()
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
m = n * 27
ha = torch.randint(0, 101, (m,), dtype=torch.uint8).pin_memory(); hb = torch.randint(0, 101, (m,), dtype=torch.uint8).pin_memory()
dc = torch.empty(m, dtype=torch.uint8, device='cuda'); hc = torch.empty(m, dtype=torch.uint8).pin_memory()
da, db = ha.cuda(), hb.cuda()
P = lambda t: C.c_void_p(t.data_ptr())
bw("zero-copy READ 2 streams -> device (gf_add)", lambda: host._check(lib.pb_field_op_dev(101, 0, P(ha), P(hb), P(dc), C.c_size_t(m), sp)), 2*m)
bw("device -> zero-copy WRITE (gf_add)", lambda: host._check(lib.pb_field_op_dev(101, 0, P(da), P(db), P(hc), C.c_size_t(m), sp)), m)
bw("zero-copy read + write", lambda: host._check(lib.pb_field_op_dev(101, 0, P(ha), P(hb), P(hc), C.c_size_t(m), sp)), 3*m)
assert torch.equal(hc, ((ha.to(torch.int32) + hb.to(torch.int32)) % 101).to(torch.uint8))
# prove+verify with inputs read straight from pinned host memory
pk = host.Plonk(W.PLONK_TEST_CIRCUIT, *W.generator_srs(9))
wit, rnd, chal, u = [torch.from_numpy(x).pin_memory() for x in W.make_batch(1, 0, n)]
proofs = torch.empty((n, 34), dtype=torch.uint8, device='cuda'); status = torch.empty(n, dtype=torch.uint8, device='cuda'); verdict = torch.empty(n, dtype=torch.uint8, device='cuda')
hp = torch.empty((n, 34), dtype=torch.uint8).pin_memory(); hs = torch.empty(n, dtype=torch.uint8).pin_memory(); hv = torch.empty(n, dtype=torch.uint8).pin_memory()
def zc_in():
    host._check(lib.pb_plonk_prove_verify_dev(pk._h, P(wit), P(rnd), P(chal), P(u), P(proofs), P(status), P(verdict), C.c_size_t(n), sp))
bw("prove_verify: inputs zero-copy, outputs device", zc_in, n*27)
def zc_in_copy_out():
    zc_in(); hp.copy_(proofs, non_blocking=True); hs.copy_(status, non_blocking=True); hv.copy_(verdict, non_blocking=True)
bw("prove_verify: zero-copy in, 3 D2H copies out", zc_in_copy_out, n*63)
def zc_all():
    host._check(lib.pb_plonk_prove_verify_dev(pk._h, P(wit), P(rnd), P(chal), P(u), P(hp), P(hs), P(hv), C.c_size_t(n), sp))
bw("prove_verify: everything zero-copy (in and out)", zc_all, n*63)
dw = [t.cuda() for t in (wit, rnd, chal, u)]
host._check(lib.pb_plonk_prove_verify_dev(pk._h, P(dw[0]), P(dw[1]), P(dw[2]), P(dw[3]), P(proofs), P(status), P(verdict), C.c_size_t(n), sp)); torch.cuda.synchronize()
print("zero-copy results match device results:", torch.equal(hp, proofs.cpu()), torch.equal(hs, status.cpu()), torch.equal(hv, verdict.cpu()))
