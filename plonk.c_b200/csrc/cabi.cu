// cabi.cu -- the C ABI declared in include/plonk_b200.h: context creation, kernel launches, and the
// host-pointer wrappers (copy in, launch, copy out).  No arithmetic of the path happens on the host:
// even the per-circuit constants (inverse Vandermonde matrix, interpolations, SRS table, verifier
// key) are produced by the batch kernels at context creation.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <chrono>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/plonk_b200.h"
#include "kernels.cuh"
#include "poly_fast.cuh"

using namespace pb;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInitializationError)
    return fail(PB_ERR_NO_DEVICE, std::string("plonk_b200: no usable CUDA device (") + cudaGetErrorString(e) +
                                      "); this library has no CPU fallback [" + what + "]");
  return fail(PB_ERR_CUDA, std::string("plonk_b200: CUDA error in ") + what + ": " + cudaGetErrorString(e));
}
#define CU(call)                                        \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)
#define ARG(cond)                                                         \
  do {                                                                    \
    if (!(cond)) return fail(PB_ERR_ARG, "plonk_b200: bad argument: " #cond); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline unsigned blocks_for(size_t n, int block) { return (unsigned)((n + block - 1) / block); }
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
  if (n <= 0) return fail(PB_ERR_NO_DEVICE, "plonk_b200: no CUDA device; this library has no CPU fallback");
  return PB_OK;
}

// RAII device scratch for the host-pointer wrappers
struct Dev {
  void* p = nullptr;
  cudaError_t err = cudaSuccess;
  explicit Dev(size_t bytes) { err = cudaMalloc(&p, bytes ? bytes : 16); }
  ~Dev() { if (p) cudaFree(p); }
  template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};
#define LAUNCH_CHECK(what) do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return cuda_fail(e__, what); } while (0)

constexpr size_t PIPE_CHUNK_MAX = 1u << 20;   // largest chunk of the host-pointer pipelines
#ifndef PB_PIPE_TWO_STREAMS
#define PB_PIPE_TWO_STREAMS 1
#endif
#ifndef PB_PIPE_SLOTS
#define PB_PIPE_SLOTS 8
#endif
#ifndef PB_PIPE_LOOKAHEAD
#define PB_PIPE_LOOKAHEAD 6
#endif
constexpr size_t PIPE_TAIL_MIN = 1u << 17;       // smallest chunk of the schedule's tail (PB_PIPE_SMALL overrides, for tuning)
constexpr int PIPE_SLOTS = PB_PIPE_SLOTS;
constexpr int PIPE_LOOKAHEAD = PB_PIPE_LOOKAHEAD;   // compact / packed modes: chunks enqueued ahead of the one whose count the host waits for
static_assert(PIPE_LOOKAHEAD >= 1 && PIPE_LOOKAHEAD < PIPE_SLOTS, "the D2H copy of a slot's previous chunk must be enqueued before the slot is reused");
// items per pipeline chunk of the host-pointer prove/verify (PB_PIPE_CHUNK overrides, for tuning; multiple of 128)
size_t pipe_chunk() {
  static size_t v = [] {
    size_t c = 1u << 19;   // measured (e2e, 2^21 items): 2^17 0.88, 2^18 0.95, 3*2^17..2^20 1.00-1.01 G proofs/s
    if (const char* e = getenv("PB_PIPE_CHUNK")) { size_t x = strtoull(e, nullptr, 10); if (x >= 128 && x <= PIPE_CHUNK_MAX) c = x & ~(size_t)127; }
    return c;
  }();
  return v;
}

// one buffer set of the host-pointer pipelines, sized for `cap` items (a multiple of 128, grown on demand)
struct PipeSlot {
  cudaEvent_t ev_in = nullptr, ev_k = nullptr, ev_out = nullptr;   // inputs landed / kernels done / outputs copied out
  uint8_t* in = nullptr;       // [cap * 26]: witness | rnd | chal (struct inputs), or the packed records [cap * 16]
  uint8_t* proofs = nullptr;   // [cap * 34]: PROOF structs at item positions
  uint8_t* dense = nullptr;    // [cap * 34 + 16]: completed proofs only (compact / packed outputs)
  uint32_t* offs = nullptr;    // [cap / 128 + 4]: group offsets of the dense list; the last word is the chunk's count
};
enum PipeMode { PIPE_PROVE = 0, PIPE_STRUCT = 1, PIPE_COMPACT = 2, PIPE_PACKED = 3, PIPE_PACKED3 = 4 };

}  // namespace

struct pb_ctx {
  int device = 0;
  CircuitConst cc{};
  VerifyKey vk{};
  bool vk_valid = true;               // false when a selector polynomial is longer than the SRS
  ProverTables* d_tables = nullptr;   // device: exact sequential path (any SRS)
  ProverPairTables* d_pair_tables = nullptr;   // device: fast path, only when srs_canonical
  ProverWideTables* d_wide_tables = nullptr;   // device: widest fast path (48 MB look-up table), only when srs_canonical and not PB_WIDE_TABLES=0
  uint16_t* d_wide_store = nullptr;            // 3 * 17^3 + 17^6 packed points behind d_wide_tables
  VerifyTables* d_verify_tables = nullptr;     // device: fast path, only when key_canonical
  VerifyLogTables* d_verify_log = nullptr;     // device: table path (discrete logarithms + pairing tables), only when key_canonical and not PB_VERIFY_TABLES=0
  bool srs_canonical = false;         // every SRS point is a canonically encoded point of E(F_101)
  bool key_canonical = false;         // same for the nine verifier-key points
  bool force_exact = false;           // PB_FORCE_EXACT=1: always take the sequential path (tests)
  uint32_t* d_srs_table = nullptr;    // device, [srs_len][17]
  uint32_t srs_len = 0;
  std::vector<uint8_t> srs_g1s;       // host copy, [srs_len][3]
  uint8_t srs_g2[4] = {0, 0, 0, 0};
  uint8_t setup[37] = {0};            // h k1_h k2_h h_pows_inv z_h len
  uint8_t circuit_dump[48] = {0};
  uint8_t vkey_bytes[27] = {0};
  std::vector<uint8_t> table_bytes;   // [srs_len][17][3]
  // scratch for the dense list of completed proofs: one buffer per stream (work on one stream is ordered, so a
  // buffer can be reused by the next call on that stream), grown on demand, freed with the context
  std::mutex scratch_mu;
  std::map<cudaStream_t, std::pair<uint32_t*, size_t>> scratch;
  std::mutex pipe_mu;
  bool pipe_ready = false;            // streams and events exist
  cudaStream_t s_in = nullptr, s_k = nullptr, s_k2 = nullptr, s_out = nullptr;   // H2D engine, SMs (even / odd chunks), D2H engine
  uint8_t *d_u = nullptr, *d_status = nullptr, *d_verdict = nullptr, *d_sv = nullptr;   // whole-batch one-byte arrays of the host-pointer pipelines
  size_t small_cap = 0;
  size_t slot_cap = 0;                // items each slot's buffers hold
  PipeSlot slots[PIPE_SLOTS];
  uint32_t* h_count = nullptr;        // pinned + mapped, [PIPE_SLOTS]: completed proofs of the chunk in each slot (written by the kernel)
  uint32_t* h_count_dev = nullptr;    // the device-side address of h_count
  uint8_t* d_wtab = nullptr;          // seeded mode: the 289-row witness table of the synthetic stream
  uint8_t* d_seed_ws = nullptr;       // seeded mode: workspace (packed inputs, proofs, status, verdict, counters)
  size_t seed_ws_items = 0;
};

namespace {

struct DeviceGuard {
  int prev = 0;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
  }
  ~DeviceGuard() { if (ok) cudaSetDevice(prev); }
};

// `_dev` entry points that take a context launch on the CALLER's current device: it must be the context's (tables, keys and
// scratch live there); anything else would be a peer access or a fault
int on_ctx_device(const pb_ctx* ctx) {
  int cur = -1;
  CU(cudaGetDevice(&cur));
  if (cur != ctx->device)
    return fail(PB_ERR_ARG, "plonk_b200: the context lives on device " + std::to_string(ctx->device) + " but the current device is " + std::to_string(cur) +
                                " (call cudaSetDevice before a *_dev entry point)");
  return PB_OK;
}

// device scratch of (n + 4) words for `stream`: [0] = counter, [4..] = list
int scratch_for(const pb_ctx* cctx, cudaStream_t st, size_t n, uint32_t** out) {
  pb_ctx* c = const_cast<pb_ctx*>(cctx);
  std::lock_guard<std::mutex> lock(c->scratch_mu);
  auto& slot = c->scratch[st];
  if (slot.second < n + 4) {
    if (slot.first) { CU(cudaStreamSynchronize(st)); CU(cudaFree(slot.first)); slot.first = nullptr; slot.second = 0; }
    size_t cap = n + 4;
    if (cap < (1u << 16)) cap = 1u << 16;
    CU(cudaMalloc(reinterpret_cast<void**>(&slot.first), cap * sizeof(uint32_t)));
    slot.second = cap;
  }
  *out = slot.first;
  return PB_OK;
}

// streams, events and the pinned count mirror (once); slot buffers for chunks of up to `cap` items (grown on demand, so a
// batch of one -- the drop-in plonk_prove -- does not allocate buffers for a million).  A failure part-way leaves
// everything allocated so far owned by the context: the next call resumes, pb_ctx_destroy frees.
int pipe_init(pb_ctx* c, size_t cap) {
  if (!c->pipe_ready) {
    if (!c->s_in) CU(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    if (!c->s_k) CU(cudaStreamCreateWithFlags(&c->s_k, cudaStreamNonBlocking));
    if (!c->s_k2) CU(cudaStreamCreateWithFlags(&c->s_k2, cudaStreamNonBlocking));
    if (!c->s_out) CU(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    for (auto& s : c->slots) {
      if (!s.ev_in) CU(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
      if (!s.ev_k) CU(cudaEventCreateWithFlags(&s.ev_k, cudaEventDisableTiming));
      if (!s.ev_out) CU(cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming));
    }
    if (!c->h_count) {
      CU(cudaHostAlloc(reinterpret_cast<void**>(&c->h_count), PIPE_SLOTS * sizeof(uint32_t), cudaHostAllocMapped));
      CU(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->h_count_dev), c->h_count, 0));
    }
    c->pipe_ready = true;
  }
  cap = (cap + 4095) & ~(size_t)4095;
  if (c->slot_cap >= cap) return PB_OK;
  CU(cudaStreamSynchronize(c->s_in)); CU(cudaStreamSynchronize(c->s_k)); CU(cudaStreamSynchronize(c->s_k2)); CU(cudaStreamSynchronize(c->s_out));
  c->slot_cap = 0;
  for (auto& s : c->slots) {
    for (void** p : {reinterpret_cast<void**>(&s.in), reinterpret_cast<void**>(&s.proofs), reinterpret_cast<void**>(&s.dense),
                     reinterpret_cast<void**>(&s.offs)}) {
      if (*p) CU(cudaFree(*p));
      *p = nullptr;
    }
    CU(cudaMalloc(&s.in, cap * 26));
    CU(cudaMalloc(&s.proofs, cap * 34));
    CU(cudaMalloc(&s.dense, cap * 34 + 16));
    CU(cudaMalloc(reinterpret_cast<void**>(&s.offs), (cap / 128 + 4) * sizeof(uint32_t)));
  }
  c->slot_cap = cap;
  return PB_OK;
}


// ------------------------------------------------------------------ host-pointer wrappers: one chunked pipeline for all
// Every host-pointer entry point of families (1)-(4) -- and the drop-in headers behind them -- runs through this: per
// device, three streams (H2D engine, SMs, D2H engine) over a ring of four buffer sets that are owned by the library and
// grown on demand, so a call allocates nothing once warm and a large batch overlaps its copies with its kernels.  Round 1
// paid cudaMalloc + cudaFree + blocking copies + a device-wide synchronise per call.
struct HArr { const void* in; void* out; size_t item; };           // one host array: `item` bytes per batch item; in XOR out
struct GenericPipe {
  std::mutex mu;
  int device = 0;
  bool ready = false;
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  static constexpr int SLOTS = 4;
  struct Slot { uint8_t* buf = nullptr; size_t cap = 0; cudaEvent_t ev_in = nullptr, ev_k = nullptr, ev_out = nullptr; } slots[SLOTS];
};
GenericPipe* pipe_for_device(int dev) {
  static std::mutex mu;
  static std::map<int, GenericPipe*> pipes;
  std::lock_guard<std::mutex> lock(mu);
  auto& p = pipes[dev];
  if (!p) { p = new GenericPipe(); p->device = dev; }
  return p;
}
constexpr size_t GP_CHUNK_BYTES = 8u << 20;    // per chunk, the larger of the two directions: well above the ~1 MB PCIe knee
inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

// launch(din, dout, m, stream): enqueue the kernels for m items whose inputs are at din[k] and outputs go to dout[k]
template <typename F>
int piped(int device, std::initializer_list<HArr> arrs, size_t n, F&& launch) {
  if (n == 0) return PB_OK;
  int rc = require_device();
  if (rc) return rc;
  if (device < 0) CU(cudaGetDevice(&device));
  DeviceGuard g(device);
  if (!g.ok) return fail(PB_ERR_CUDA, "plonk_b200: cudaSetDevice failed");
  GenericPipe* gp = pipe_for_device(device);
  std::lock_guard<std::mutex> lock(gp->mu);
  if (!gp->ready) {
    if (!gp->s_in) CU(cudaStreamCreateWithFlags(&gp->s_in, cudaStreamNonBlocking));
    if (!gp->s_k) CU(cudaStreamCreateWithFlags(&gp->s_k, cudaStreamNonBlocking));
    if (!gp->s_out) CU(cudaStreamCreateWithFlags(&gp->s_out, cudaStreamNonBlocking));
    for (auto& s : gp->slots) {
      if (!s.ev_in) CU(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
      if (!s.ev_k) CU(cudaEventCreateWithFlags(&s.ev_k, cudaEventDisableTiming));
      if (!s.ev_out) CU(cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming));
    }
    gp->ready = true;
  }
  std::vector<HArr> a(arrs);
  size_t in_item = 0, out_item = 0;
  for (auto& x : a) { (x.in ? in_item : out_item) += x.item; }
  const size_t per = in_item > out_item ? in_item : out_item;
  size_t chunk = GP_CHUNK_BYTES / (per ? per : 1);
  chunk = chunk < 4096 ? 4096 : (chunk & ~(size_t)4095);           // a multiple of 4096 items: every array's chunk offset stays 16-byte aligned
  if (chunk > n) chunk = n;
  size_t need = 0;
  std::vector<size_t> off(a.size());
  for (size_t k = 0; k < a.size(); k++) { off[k] = need; need += up256(chunk * a[k].item); }
  const size_t nchunks = (n + chunk - 1) / chunk;
  const int used = nchunks < (size_t)GenericPipe::SLOTS ? (int)nchunks : GenericPipe::SLOTS;
  for (int k = 0; k < used; k++) {
    auto& s = gp->slots[k];
    if (s.cap < need) {
      CU(cudaStreamSynchronize(gp->s_in)); CU(cudaStreamSynchronize(gp->s_k)); CU(cudaStreamSynchronize(gp->s_out));
      if (s.buf) CU(cudaFree(s.buf));
      s.buf = nullptr; s.cap = 0;
      const size_t cap = need < (1u << 16) ? (1u << 16) : need;
      CU(cudaMalloc(&s.buf, cap));
      s.cap = cap;
    }
  }
  std::vector<uint8_t*> din, dout;
  size_t done = 0;
  for (size_t c = 0; c < nchunks; c++) {
    auto& s = gp->slots[c % GenericPipe::SLOTS];
    const size_t m = n - done < chunk ? n - done : chunk;
    const bool reuse = c >= (size_t)GenericPipe::SLOTS;
    if (reuse) { CU(cudaStreamWaitEvent(gp->s_in, s.ev_k, 0)); CU(cudaStreamWaitEvent(gp->s_in, s.ev_out, 0)); }
    din.clear(); dout.clear();
    for (size_t k = 0; k < a.size(); k++) {
      if (a[k].in) {
        din.push_back(s.buf + off[k]);
        CU(cudaMemcpyAsync(s.buf + off[k], static_cast<const uint8_t*>(a[k].in) + done * a[k].item, m * a[k].item, cudaMemcpyHostToDevice, gp->s_in));
      } else {
        dout.push_back(s.buf + off[k]);
      }
    }
    CU(cudaEventRecord(s.ev_in, gp->s_in));
    CU(cudaStreamWaitEvent(gp->s_k, s.ev_in, 0));
    if ((rc = launch(din.data(), dout.data(), m, gp->s_k))) return rc;
    CU(cudaEventRecord(s.ev_k, gp->s_k));
    CU(cudaStreamWaitEvent(gp->s_out, s.ev_k, 0));
    for (size_t k = 0; k < a.size(); k++)
      if (!a[k].in) CU(cudaMemcpyAsync(static_cast<uint8_t*>(a[k].out) + done * a[k].item, s.buf + off[k], m * a[k].item, cudaMemcpyDeviceToHost, gp->s_out));
    CU(cudaEventRecord(s.ev_out, gp->s_out));
    done += m;
  }
  CU(cudaStreamSynchronize(gp->s_out));
  CU(cudaStreamSynchronize(gp->s_k));
  CU(cudaStreamSynchronize(gp->s_in));
  return PB_OK;
}
#define HIN(p, item) HArr{(p), nullptr, (size_t)(item)}
#define HOUT(p, item) HArr{nullptr, (p), (size_t)(item)}

}  // namespace

extern "C" {

const char* pb_last_error(void) { return g_err.c_str(); }
int pb_abi_version(void) { return 2; }
int pb_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
  return n;
}
int pb_host_alloc(void** out, size_t bytes) {
  ARG(out != nullptr);
  int rc = require_device();
  if (rc) return rc;
  CU(cudaHostAlloc(out, bytes ? bytes : 16, cudaHostAllocDefault));
  return PB_OK;
}
int pb_host_free(void* p) {
  if (p) CU(cudaFreeHost(p));
  return PB_OK;
}

// ------------------------------------------------------------------ family (1)
int pb_field_op_dev(int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(field == 17 || field == 101);
  ARG(op >= 0 && op <= 6);
  ARG(a && out);
  ARG(b || op == PB_OP_NEG || op == PB_OP_INV);
  ARG(aligned16(a) && aligned16(out) && (!b || aligned16(b)));
  if (n == 0) return PB_OK;
  unsigned grid = blocks_for((n + 15) / 16, BLOCK_LIGHT);
  if (grid > 148u * 16u) grid = 148u * 16u;
  if (op == PB_OP_POW) {   // table-driven (kernels.cuh: field_pow_kernel); fewer, longer-lived blocks amortise the table
    if (grid > 148u * 8u) grid = 148u * 8u;
    if (field == 17) field_pow_kernel<17><<<grid, BLOCK_LIGHT, 0, S(stream)>>>(a, b, out, n);
    else field_pow_kernel<101><<<grid, BLOCK_LIGHT, 0, S(stream)>>>(a, b, out, n);
  } else {
#define PB_FIELD_OP(F_, O_) if (field == F_ && op == O_) field_op_kernel<F_, O_><<<grid, BLOCK_LIGHT, 0, S(stream)>>>(a, b, out, n);
    PB_FIELD_OP(17, 0) PB_FIELD_OP(17, 1) PB_FIELD_OP(17, 2) PB_FIELD_OP(17, 3) PB_FIELD_OP(17, 4) PB_FIELD_OP(17, 5)
    PB_FIELD_OP(101, 0) PB_FIELD_OP(101, 1) PB_FIELD_OP(101, 2) PB_FIELD_OP(101, 3) PB_FIELD_OP(101, 4) PB_FIELD_OP(101, 5)
#undef PB_FIELD_OP
  }
  LAUNCH_CHECK("field_op_kernel");
  return PB_OK;
}
int pb_field_op(int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(a && out);
  ARG(b || op == PB_OP_NEG || op == PB_OP_INV);
  if (b) return piped(-1, {HIN(a, 1), HIN(b, 1), HOUT(out, 1)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_field_op_dev(field, op, di[0], di[1], dq[0], m, st); });
  return piped(-1, {HIN(a, 1), HOUT(out, 1)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_field_op_dev(field, op, di[0], nullptr, dq[0], m, st); });
}

int pb_field_new_dev(int field, const int64_t* v, uint8_t* out, size_t n, void* stream) {
  if (n == 0) return PB_OK;
  ARG((field == 17 || field == 101) && v && out);
  unsigned grid = blocks_for(n, BLOCK_LIGHT);
  if (grid > 148u * 16u) grid = 148u * 16u;
  if (field == 17) field_new_kernel<17><<<grid, BLOCK_LIGHT, 0, S(stream)>>>(reinterpret_cast<const long long*>(v), out, n);
  else field_new_kernel<101><<<grid, BLOCK_LIGHT, 0, S(stream)>>>(reinterpret_cast<const long long*>(v), out, n);
  LAUNCH_CHECK("field_new_kernel");
  return PB_OK;
}
int pb_field_new(int field, const int64_t* v, uint8_t* out, size_t n) {
  if (n == 0) return PB_OK;
  ARG(v && out);
  return piped(-1, {HIN(v, 8), HOUT(out, 1)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_field_new_dev(field, reinterpret_cast<const int64_t*>(di[0]), dq[0], m, st); });
}

// ------------------------------------------------------------------ family (2)
int pb_poly_binop_dev(int op, const uint8_t* a, const uint8_t* alen, size_t sa, const uint8_t* b, const uint8_t* blen, size_t sb,
                      uint8_t* out, uint8_t* olen, size_t so, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(op >= 0 && op <= 2);
  ARG(a && alen && b && blen && out && olen);
  ARG(sa >= 1 && sb >= 1 && sa <= PB_POLY_MAX && sb <= PB_POLY_MAX);
  ARG(so >= (op == 2 ? sa + sb - 1 : (sa > sb ? sa : sb)) && so <= 2 * PB_POLY_MAX);
  if (n == 0) return PB_OK;
  // register-resident fast path for the shapes of BASELINE config 2 and of the prover (natural output stride, aligned)
  const bool al = aligned16(a) && aligned16(b) && aligned16(out);
  if (op == PB_POLY_MUL && al && so == sa + sb - 1) {
    // full groups of PF4_ITEMS items: four items per thread (poly_fast.cuh); the ragged tail: one item per thread
    const size_t m4 = (aligned16(alen) && aligned16(blen) && aligned16(olen)) ? n / PF4_ITEMS * PF4_ITEMS : 0;
#define PB_MUL_FAST(A_, B_)                                                                                         \
    if (sa == A_ && sb == B_) {                                                                                     \
      if (m4) poly_mul_fast4_kernel<A_, B_><<<(unsigned)(m4 / PF4_ITEMS), PF4_BLOCK, 0, S(stream)>>>(a, alen, b, blen, out, olen); \
      if (n > m4) poly_mul_fast_kernel<A_, B_><<<blocks_for(n - m4, PF_BLOCK), PF_BLOCK, 0, S(stream)>>>(           \
          a + m4 * A_, alen + m4, b + m4 * B_, blen + m4, out + m4 * (A_ + B_ - 1), olen + m4, n - m4);              \
      LAUNCH_CHECK("poly_mul_fast_kernel");                                                                         \
      return PB_OK;                                                                                                 \
    }
    PB_MUL_FAST(6, 6) PB_MUL_FAST(11, 6) PB_MUL_FAST(16, 7) PB_MUL_FAST(11, 4) PB_MUL_FAST(6, 4) PB_MUL_FAST(7, 4)
#undef PB_MUL_FAST
  }
  poly_binop_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(op, a, alen, (int)sa, b, blen, (int)sb, out, olen, (int)so, n);
  LAUNCH_CHECK("poly_binop_kernel");
  return PB_OK;
}
int pb_poly_binop(int op, const uint8_t* a, const uint8_t* alen, size_t sa, const uint8_t* b, const uint8_t* blen, size_t sb,
                  uint8_t* out, uint8_t* olen, size_t so, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(a && alen && b && blen && out && olen);
  return piped(-1, {HIN(a, sa), HIN(alen, 1), HIN(b, sb), HIN(blen, 1), HOUT(out, so), HOUT(olen, 1)}, n,
               [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
                 return pb_poly_binop_dev(op, di[0], di[1], sa, di[2], di[3], sb, dq[0], dq[1], so, m, st); });
}

int pb_poly_divide_dev(const uint8_t* num, const uint8_t* nlen, size_t sn, const uint8_t* den, const uint8_t* dlen, size_t sd,
                       uint8_t* quot, uint8_t* qlen, size_t sq, uint8_t* rem, uint8_t* rlen, size_t sr,
                       uint8_t* status, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(num && nlen && den && dlen && quot && qlen && rem && rlen && status);
  ARG(sn >= 1 && sd >= 1 && sn <= PB_POLY_MAX && sd <= PB_POLY_MAX && sq >= 1 && sr >= 1);
  if (n == 0) return PB_OK;
  const bool al = aligned16(num) && aligned16(den) && aligned16(quot) && aligned16(rem);
  if (al && sn >= sd && sq == sn - sd + 1 && sr == sd - 1) {
#define PB_DIV_FAST(N_, D_)                                                                                                       \
    if (sn == N_ && sd == D_) {                                                                                                   \
      poly_divide_fast_kernel<N_, D_><<<blocks_for(n, PF_BLOCK), PF_BLOCK, 0, S(stream)>>>(num, nlen, den, dlen, quot, qlen, rem, rlen, status, n); \
      LAUNCH_CHECK("poly_divide_fast_kernel");                                                                                    \
      return PB_OK;                                                                                                               \
    }
    PB_DIV_FAST(11, 5) PB_DIV_FAST(22, 5) PB_DIV_FAST(10, 2) PB_DIV_FAST(7, 2)
#undef PB_DIV_FAST
  }
  poly_divide_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(num, nlen, (int)sn, den, dlen, (int)sd, quot, qlen, (int)sq,
                                                                              rem, rlen, (int)sr, status, n);
  LAUNCH_CHECK("poly_divide_kernel");
  return PB_OK;
}
int pb_poly_divide(const uint8_t* num, const uint8_t* nlen, size_t sn, const uint8_t* den, const uint8_t* dlen, size_t sd,
                   uint8_t* quot, uint8_t* qlen, size_t sq, uint8_t* rem, uint8_t* rlen, size_t sr, uint8_t* status, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(num && nlen && den && dlen && quot && qlen && rem && rlen && status);
  return piped(-1, {HIN(num, sn), HIN(nlen, 1), HIN(den, sd), HIN(dlen, 1), HOUT(quot, sq), HOUT(qlen, 1), HOUT(rem, sr), HOUT(rlen, 1), HOUT(status, 1)}, n,
               [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
                 return pb_poly_divide_dev(di[0], di[1], sn, di[2], di[3], sd, dq[0], dq[1], sq, dq[2], dq[3], sr, dq[4], m, st); });
}

// poly_divide(p, Z_H) against the context's Z_H = x^4 - 1 (plonk.h:505): numerators of stride 11 (config 2: a product of two
// 6-coefficient polynomials) or 22 (the prover's t_numer); quotient sn - 4 columns, remainder 4 columns
int pb_poly_divide_zh_dev(const pb_ctx* ctx, const uint8_t* num, const uint8_t* nlen, size_t sn, uint8_t* quot, uint8_t* qlen, uint8_t* rem,
                          uint8_t* rlen, uint8_t* status, size_t n, void* stream) {
  if (n == 0) return PB_OK;
  ARG(ctx && num && nlen && quot && qlen && rem && rlen && status);
  ARG(sn == 11 || sn == 22);
  ARG(aligned16(num) && aligned16(quot) && aligned16(rem));
  const size_t m4 = (aligned16(nlen) && aligned16(qlen) && aligned16(rlen) && aligned16(status)) ? n / PF4_ITEMS * PF4_ITEMS : 0;
#define PB_DIVZH(N_)                                                                                                                          \
  if (sn == N_) {                                                                                                                             \
    if (m4) poly_divide_zh4_kernel<N_><<<(unsigned)(m4 / PF4_ITEMS), PF4_BLOCK, 0, S(stream)>>>(num, nlen, quot, qlen, rem, rlen, status);   \
    if (n > m4) poly_divide_zh_kernel<N_><<<blocks_for(n - m4, PF_BLOCK), PF_BLOCK, 0, S(stream)>>>(                                         \
        num + m4 * N_, nlen + m4, quot + m4 * (N_ - 4), qlen + m4, rem + m4 * 4, rlen + m4, status + m4, n - m4);                            \
  }
  PB_DIVZH(11) PB_DIVZH(22)
#undef PB_DIVZH
  LAUNCH_CHECK("poly_divide_zh_kernel");
  return PB_OK;
}
int pb_poly_divide_zh(const pb_ctx* ctx, const uint8_t* num, const uint8_t* nlen, size_t sn, uint8_t* quot, uint8_t* qlen, uint8_t* rem,
                      uint8_t* rlen, uint8_t* status, size_t n) {
  if (n == 0) return PB_OK;
  ARG(ctx && num && nlen && quot && qlen && rem && rlen && status);
  ARG(sn == 11 || sn == 22);
  return piped(ctx->device, {HIN(num, sn), HIN(nlen, 1), HOUT(quot, sn - 4), HOUT(qlen, 1), HOUT(rem, 4), HOUT(rlen, 1), HOUT(status, 1)}, n,
               [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
                 return pb_poly_divide_zh_dev(ctx, di[0], di[1], sn, dq[0], dq[1], dq[2], dq[3], dq[4], m, st); });
}

int pb_poly_eval_dev(const uint8_t* p, const uint8_t* plen, size_t sp, const uint8_t* x, uint8_t* out, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(p && plen && x && out && sp >= 1);
  if (n == 0) return PB_OK;
  if (aligned16(p)) {
    const size_t m4 = (aligned16(plen) && aligned16(x) && aligned16(out)) ? n / PF4_ITEMS * PF4_ITEMS : 0;
#define PB_EVAL_FAST(P_)                                                                                   \
    if (sp == P_) {                                                                                        \
      if (m4) poly_eval_fast4_kernel<P_><<<(unsigned)(m4 / PF4_ITEMS), PF4_BLOCK, 0, S(stream)>>>(p, plen, x, out); \
      if (n > m4) poly_eval_fast_kernel<P_><<<blocks_for(n - m4, PF_BLOCK), PF_BLOCK, 0, S(stream)>>>(p + m4 * P_, plen + m4, x + m4, out + m4, n - m4); \
      LAUNCH_CHECK("poly_eval_fast_kernel");                                                               \
      return PB_OK;                                                                                        \
    }
    PB_EVAL_FAST(4) PB_EVAL_FAST(6) PB_EVAL_FAST(7) PB_EVAL_FAST(11) PB_EVAL_FAST(18) PB_EVAL_FAST(22)
#undef PB_EVAL_FAST
  }
  poly_eval_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(p, plen, (int)sp, x, out, n);
  LAUNCH_CHECK("poly_eval_kernel");
  return PB_OK;
}
int pb_poly_eval(const uint8_t* p, const uint8_t* plen, size_t sp, const uint8_t* x, uint8_t* out, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(p && plen && x && out);
  return piped(-1, {HIN(p, sp), HIN(plen, 1), HIN(x, 1), HOUT(out, 1)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_poly_eval_dev(di[0], di[1], sp, di[2], dq[0], m, st); });
}

int pb_poly_unop_dev(int op, const uint8_t* p, const uint8_t* plen, size_t sp, const uint8_t* k, uint8_t* out, uint8_t* olen,
                     size_t so, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(op >= 0 && op <= 3);
  ARG(p && plen && out && olen && sp >= 1 && sp <= PB_POLY_MAX && so >= sp && so <= 2 * PB_POLY_MAX);
  ARG(k || op == PB_POLY_NEGATE);
  if (n == 0) return PB_OK;
  poly_unop_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(op, p, plen, (int)sp, k, out, olen, (int)so, n);
  LAUNCH_CHECK("poly_unop_kernel");
  return PB_OK;
}
int pb_poly_unop(int op, const uint8_t* p, const uint8_t* plen, size_t sp, const uint8_t* k, uint8_t* out, uint8_t* olen, size_t so, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(p && plen && out && olen);
  if (k) return piped(-1, {HIN(p, sp), HIN(plen, 1), HIN(k, 1), HOUT(out, so), HOUT(olen, 1)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_poly_unop_dev(op, di[0], di[1], sp, di[2], dq[0], dq[1], so, m, st); });
  return piped(-1, {HIN(p, sp), HIN(plen, 1), HOUT(out, so), HOUT(olen, 1)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_poly_unop_dev(op, di[0], di[1], sp, nullptr, dq[0], dq[1], so, m, st); });
}

int pb_poly_slice_dev(const uint8_t* p, const uint8_t* plen, size_t sp, const uint8_t* start, const uint8_t* end, uint8_t* out,
                      uint8_t* olen, size_t so, uint8_t* status, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(p && plen && start && end && out && olen && status && sp >= 1 && sp <= PB_POLY_MAX && so >= sp);
  if (n == 0) return PB_OK;
  poly_slice_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(p, plen, (int)sp, start, end, out, olen, (int)so, status, n);
  LAUNCH_CHECK("poly_slice_kernel");
  return PB_OK;
}
int pb_poly_slice(const uint8_t* p, const uint8_t* plen, size_t sp, const uint8_t* start, const uint8_t* end, uint8_t* out,
                  uint8_t* olen, size_t so, uint8_t* status, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(p && plen && start && end && out && olen && status);
  return piped(-1, {HIN(p, sp), HIN(plen, 1), HIN(start, 1), HIN(end, 1), HOUT(out, so), HOUT(olen, 1), HOUT(status, 1)}, n,
               [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
                 return pb_poly_slice_dev(di[0], di[1], sp, di[2], di[3], dq[0], dq[1], so, dq[2], m, st); });
}

int pb_poly_lagrange_dev(const uint8_t* xs, const uint8_t* ys, size_t len, uint8_t* out, uint8_t* olen, size_t so, uint8_t* status,
                         size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(xs && ys && out && olen && status && len >= 1 && len <= 16 && so >= len);
  if (n == 0) return PB_OK;
  poly_lagrange_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(xs, ys, (int)len, out, olen, (int)so, status, n);
  LAUNCH_CHECK("poly_lagrange_kernel");
  return PB_OK;
}
int pb_poly_lagrange(const uint8_t* xs, const uint8_t* ys, size_t len, uint8_t* out, uint8_t* olen, size_t so, uint8_t* status, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(xs && ys && out && olen && status);
  return piped(-1, {HIN(xs, len), HIN(ys, len), HOUT(out, so), HOUT(olen, 1), HOUT(status, 1)}, n,
               [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
                 return pb_poly_lagrange_dev(di[0], di[1], len, dq[0], dq[1], so, dq[2], m, st); });
}

int pb_interpolate_at_h_dev(const pb_ctx* ctx, const uint8_t* vals, uint8_t* out, uint8_t* olen, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && vals && out && olen);
  { int rc_dev = on_ctx_device(ctx); if (rc_dev) return rc_dev; }
  ARG(aligned16(vals) && aligned16(out));
  if (n == 0) return PB_OK;
  const size_t m4 = aligned16(olen) ? n / PF4_ITEMS * PF4_ITEMS : 0;
  if (m4) interpolate4_kernel<<<(unsigned)(m4 / PF4_ITEMS), PF4_BLOCK, 0, S(stream)>>>(ctx->cc, vals, out, olen);
  if (n > m4) interpolate_kernel<<<blocks_for(n - m4, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(ctx->cc, vals + 4 * m4, out + 4 * m4, olen + m4, n - m4);
  LAUNCH_CHECK("interpolate_kernel");
  return PB_OK;
}
int pb_interpolate_at_h(const pb_ctx* ctx, const uint8_t* vals, uint8_t* out, uint8_t* olen, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && vals && out && olen);
  return piped(ctx->device, {HIN(vals, 4), HOUT(out, 4), HOUT(olen, 1)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_interpolate_at_h_dev(ctx, di[0], dq[0], dq[1], m, st); });
}

int pb_config2_items_dev(const pb_ctx* ctx, const uint8_t* a, const uint8_t* b, const uint8_t* x, const uint8_t* vals, uint8_t* prod,
                         uint8_t* prod_len, uint8_t* quot, uint8_t* quot_len, uint8_t* rem, uint8_t* rem_len, uint8_t* evals, uint8_t* interp,
                         uint8_t* interp_len, size_t n, void* stream) {
  if (n == 0) return PB_OK;
  ARG(ctx && a && b && x && vals && prod && prod_len && quot && quot_len && rem && rem_len && evals && interp && interp_len);
  { int rc_dev = on_ctx_device(ctx); if (rc_dev) return rc_dev; }
  ARG(aligned16(a) && aligned16(b) && aligned16(vals) && aligned16(prod) && aligned16(quot) && aligned16(rem) && aligned16(interp));
  // full groups of PF4_ITEMS items: four items per thread, whole-word accesses (poly_fast.cuh); the ragged tail: one item per thread
  const size_t blocks4 = (aligned16(x) && aligned16(prod_len) && aligned16(quot_len) && aligned16(rem_len) && aligned16(evals) && aligned16(interp_len))
                             ? n / PF4_ITEMS : 0;
  const size_t m = blocks4 * PF4_ITEMS;
  if (blocks4) {
    config2_kernel4<<<(unsigned)blocks4, PF4_BLOCK, 0, S(stream)>>>(ctx->cc, a, b, x, vals, prod, prod_len, quot, quot_len, rem, rem_len, evals, interp,
                                                                   interp_len);
    LAUNCH_CHECK("config2_kernel4");
  }
  if (n > m) {
    config2_kernel<<<blocks_for(n - m, PF_BLOCK), PF_BLOCK, 0, S(stream)>>>(ctx->cc, a + 6 * m, b + 6 * m, x + m, vals + 4 * m, prod + 11 * m, prod_len + m,
                                                                           quot + 7 * m, quot_len + m, rem + 4 * m, rem_len + m, evals + m,
                                                                           interp + 4 * m, interp_len + m, n - m);
    LAUNCH_CHECK("config2_kernel");
  }
  return PB_OK;
}

int pb_config2_items(const pb_ctx* ctx, const uint8_t* a, const uint8_t* b, const uint8_t* x, const uint8_t* vals, uint8_t* prod, uint8_t* prod_len,
                     uint8_t* quot, uint8_t* quot_len, uint8_t* rem, uint8_t* rem_len, uint8_t* evals, uint8_t* interp, uint8_t* interp_len, size_t n) {
  if (n == 0) return PB_OK;
  ARG(ctx && a && b && x && vals && prod && prod_len && quot && quot_len && rem && rem_len && evals && interp && interp_len);
  return piped(ctx->device, {HIN(a, 6), HIN(b, 6), HIN(x, 1), HIN(vals, 4), HOUT(prod, 11), HOUT(prod_len, 1), HOUT(quot, 7), HOUT(quot_len, 1), HOUT(rem, 4),
                             HOUT(rem_len, 1), HOUT(evals, 1), HOUT(interp, 4), HOUT(interp_len, 1)}, n,
               [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
                 return pb_config2_items_dev(ctx, di[0], di[1], di[2], di[3], dq[0], dq[1], dq[2], dq[3], dq[4], dq[5], dq[6], dq[7], dq[8], m, st); });
}

int pb_matrix_mul_dev(const uint8_t* a, const uint8_t* b, uint8_t* out, uint32_t m, uint32_t k, uint32_t c, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(a && b && out && m >= 1 && k >= 1 && c >= 1 && m <= 8 && k <= 8 && c <= 8);
  if (n == 0) return PB_OK;
  matrix_mul_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(a, b, out, (int)m, (int)k, (int)c, n);
  LAUNCH_CHECK("matrix_mul_kernel");
  return PB_OK;
}
int pb_matrix_mul(const uint8_t* a, const uint8_t* b, uint8_t* out, uint32_t m, uint32_t k, uint32_t c, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(a && b && out && m >= 1 && k >= 1 && c >= 1 && m <= 8 && k <= 8 && c <= 8);
  return piped(-1, {HIN(a, m * k), HIN(b, k * c), HOUT(out, m * c)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t cnt, cudaStream_t st) {
    return pb_matrix_mul_dev(di[0], di[1], dq[0], m, k, c, cnt, st); });
}
int pb_matrix_inv_dev(const uint8_t* a, uint8_t* out, uint32_t dim, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(a && out && dim >= 1 && dim <= 8);
  if (n == 0) return PB_OK;
  matrix_inv_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(a, out, (int)dim, n);
  LAUNCH_CHECK("matrix_inv_kernel");
  return PB_OK;
}
int pb_matrix_inv(const uint8_t* a, uint8_t* out, uint32_t dim, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(a && out && dim >= 1 && dim <= 8);
  return piped(-1, {HIN(a, dim * dim), HOUT(out, dim * dim)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_matrix_inv_dev(di[0], dq[0], dim, m, st); });
}

int pb_matrix_gauss_jordan_dev(uint8_t* a, uint32_t rows, uint32_t cols, size_t n, void* stream) {
  if (n == 0) return PB_OK;
  ARG(a && rows >= 1 && cols >= 1 && rows <= 8 && cols <= 16);
  matrix_rref_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(a, (int)rows, (int)cols, n);
  LAUNCH_CHECK("matrix_rref_kernel");
  return PB_OK;
}
int pb_matrix_gauss_jordan(uint8_t* a, uint32_t rows, uint32_t cols, size_t n) {
  if (n == 0) return PB_OK;
  ARG(a && rows >= 1 && cols >= 1 && rows <= 8 && cols <= 16);
  return piped(-1, {HIN(a, rows * cols), HOUT(a, rows * cols)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    CU(cudaMemcpyAsync(dq[0], di[0], m * rows * cols, cudaMemcpyDeviceToDevice, st));     // the kernel reduces in place
    return pb_matrix_gauss_jordan_dev(dq[0], rows, cols, m, st); });
}

// ------------------------------------------------------------------ family (3)
int pb_g1_op_dev(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(op >= 0 && op <= 2 && a && out && (b || op != PB_G_ADD));
  if (n == 0) return PB_OK;
  g1_op_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(op, a, b, out, n);
  LAUNCH_CHECK("g1_op_kernel");
  return PB_OK;
}
int pb_g1_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(a && out && (b || op != PB_G_ADD));
  if (b) return piped(-1, {HIN(a, 3), HIN(b, 3), HOUT(out, 3)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_g1_op_dev(op, di[0], di[1], dq[0], m, st); });
  return piped(-1, {HIN(a, 3), HOUT(out, 3)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_g1_op_dev(op, di[0], nullptr, dq[0], m, st); });
}
int pb_g1_mul_dev(const uint8_t* points, const uint64_t* scalars, uint8_t* out, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(points && scalars && out);
  if (n == 0) return PB_OK;
  g1_mul_kernel<uint64_t><<<blocks_for(n, BLOCK), BLOCK, 0, S(stream)>>>(points, scalars, out, n);
  LAUNCH_CHECK("g1_mul_kernel");
  return PB_OK;
}
int pb_g1_mul(const uint8_t* points, const uint64_t* scalars, uint8_t* out, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(points && scalars && out);
  return piped(-1, {HIN(points, 3), HIN(scalars, 8), HOUT(out, 3)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_g1_mul_dev(di[0], reinterpret_cast<const uint64_t*>(di[1]), dq[0], m, st); });
}
int pb_g1_mul_u8_dev(const uint8_t* points, const uint8_t* scalars, uint8_t* out, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(points && scalars && out);
  if (n == 0) return PB_OK;
  g1_mul_kernel<uint8_t><<<blocks_for(n, BLOCK), BLOCK, 0, S(stream)>>>(points, scalars, out, n);
  LAUNCH_CHECK("g1_mul_kernel");
  return PB_OK;
}
int pb_g1_mul_u8(const uint8_t* points, const uint8_t* scalars, uint8_t* out, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(points && scalars && out);
  return piped(-1, {HIN(points, 3), HIN(scalars, 1), HOUT(out, 3)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_g1_mul_u8_dev(di[0], di[1], dq[0], m, st); });
}
int pb_g1_is_on_curve_dev(const uint8_t* points, uint8_t* out, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(points && out);
  if (n == 0) return PB_OK;
  g1_on_curve_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(points, out, n);
  LAUNCH_CHECK("g1_on_curve_kernel");
  return PB_OK;
}
int pb_g1_is_on_curve(const uint8_t* points, uint8_t* out, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(points && out);
  return piped(-1, {HIN(points, 3), HOUT(out, 1)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_g1_is_on_curve_dev(di[0], dq[0], m, st); });
}
int pb_g2_op_dev(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG((op == PB_G_ADD || op == PB_G_NEG) && a && out && (b || op != PB_G_ADD));
  if (n == 0) return PB_OK;
  g2_op_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(op, a, b, out, n);
  LAUNCH_CHECK("g2_op_kernel");
  return PB_OK;
}
int pb_g2_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(a && out && (b || op != PB_G_ADD));
  if (b) return piped(-1, {HIN(a, 2), HIN(b, 2), HOUT(out, 2)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_g2_op_dev(op, di[0], di[1], dq[0], m, st); });
  return piped(-1, {HIN(a, 2), HOUT(out, 2)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_g2_op_dev(op, di[0], nullptr, dq[0], m, st); });
}
int pb_g2_mul_dev(const uint8_t* points, const uint64_t* scalars, uint8_t* out, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(points && scalars && out);
  if (n == 0) return PB_OK;
  g2_mul_kernel<<<blocks_for(n, BLOCK), BLOCK, 0, S(stream)>>>(points, scalars, out, n);
  LAUNCH_CHECK("g2_mul_kernel");
  return PB_OK;
}
int pb_g2_mul(const uint8_t* points, const uint64_t* scalars, uint8_t* out, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(points && scalars && out);
  return piped(-1, {HIN(points, 2), HIN(scalars, 8), HOUT(out, 2)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_g2_mul_dev(di[0], reinterpret_cast<const uint64_t*>(di[1]), dq[0], m, st); });
}

int pb_srs_eval_at_s_dev(const pb_ctx* ctx, const uint8_t* polys, const uint8_t* plen, size_t sp, uint8_t* out, uint8_t* status,
                         size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && polys && plen && out && status && sp >= 1 && sp <= PB_POLY_MAX);
  { int rc_dev = on_ctx_device(ctx); if (rc_dev) return rc_dev; }
  if (n == 0) return PB_OK;
  commit_kernel<<<blocks_for(n, BLOCK), BLOCK, ctx->srs_len * 17 * sizeof(uint32_t), S(stream)>>>(ctx->d_srs_table, ctx->srs_len, polys, plen,
                                                                                                 (int)sp, out, status, n);
  LAUNCH_CHECK("commit_kernel");
  return PB_OK;
}
int pb_srs_eval_at_s(const pb_ctx* ctx, const uint8_t* polys, const uint8_t* plen, size_t sp, uint8_t* out, uint8_t* status, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && polys && plen && out && status);
  return piped(ctx->device, {HIN(polys, sp), HIN(plen, 1), HOUT(out, 3), HOUT(status, 1)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_srs_eval_at_s_dev(ctx, di[0], di[1], sp, dq[0], dq[1], m, st); });
}

int pb_srs_eval_at_s_raw_dev(const uint8_t* srs_g1s, uint32_t srs_len, const uint8_t* polys, const uint8_t* plen, size_t sp, int trim,
                             uint8_t* out, uint8_t* status, size_t n, void* stream) {
  if (n == 0) return PB_OK;
  ARG(srs_g1s && polys && plen && out && status && sp >= 1 && sp <= PB_POLY_MAX && srs_len >= 1);
  commit_raw_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(srs_g1s, srs_len, polys, plen, (int)sp, trim, out, status, n);
  LAUNCH_CHECK("commit_raw_kernel");
  return PB_OK;
}
int pb_srs_eval_at_s_raw(const uint8_t* srs_g1s, uint32_t srs_len, const uint8_t* polys, const uint8_t* plen, size_t sp, int trim, uint8_t* out,
                         uint8_t* status, size_t n) {
  if (n == 0) return PB_OK;
  ARG(srs_g1s && polys && plen && out && status && srs_len >= 1);
  // the SRS is one shared array, not a per-item one: it rides along as a single "item" of srs_len * 3 bytes per chunk
  int rc = require_device();
  if (rc) return rc;
  void* d_srs = nullptr;
  CU(cudaMalloc(&d_srs, 3 * (size_t)srs_len));
  cudaError_t e = cudaMemcpy(d_srs, srs_g1s, 3 * (size_t)srs_len, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    rc = piped(-1, {HIN(polys, sp), HIN(plen, 1), HOUT(out, 3), HOUT(status, 1)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
      return pb_srs_eval_at_s_raw_dev(static_cast<const uint8_t*>(d_srs), srs_len, di[0], di[1], sp, trim, dq[0], dq[1], m, st); });
  cudaFree(d_srs);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  return rc;
}

// ------------------------------------------------------------------ family (4)
int pb_gtp_mul_dev(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(a && b && out);
  if (n == 0) return PB_OK;
  gtp_mul_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(a, b, out, n);
  LAUNCH_CHECK("gtp_mul_kernel");
  return PB_OK;
}
int pb_gtp_mul(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(a && b && out);
  return piped(-1, {HIN(a, 2), HIN(b, 2), HOUT(out, 2)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_gtp_mul_dev(di[0], di[1], dq[0], m, st); });
}
int pb_gtp_pow_dev(const uint8_t* a, const uint64_t* e, uint8_t* out, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(a && e && out);
  if (n == 0) return PB_OK;
  gtp_pow_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(a, e, out, n);
  LAUNCH_CHECK("gtp_pow_kernel");
  return PB_OK;
}
int pb_gtp_pow(const uint8_t* a, const uint64_t* e, uint8_t* out, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(a && e && out);
  return piped(-1, {HIN(a, 2), HIN(e, 8), HOUT(out, 2)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_gtp_pow_dev(di[0], reinterpret_cast<const uint64_t*>(di[1]), dq[0], m, st); });
}
int pb_line_dev(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(a && b && out);
  if (n == 0) return PB_OK;
  line_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(a, b, out, n);
  LAUNCH_CHECK("line_kernel");
  return PB_OK;
}
int pb_line(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(a && b && out);
  return piped(-1, {HIN(a, 3), HIN(b, 3), HOUT(out, 3)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_line_dev(di[0], di[1], dq[0], m, st); });
}
int pb_pairing_dev(const uint8_t* p, const uint8_t* q, uint8_t* out, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(p && q && out);
  if (n == 0) return PB_OK;
  pairing_kernel<<<blocks_for(n, BLOCK), BLOCK, 0, S(stream)>>>(p, q, out, n);
  LAUNCH_CHECK("pairing_kernel");
  return PB_OK;
}
int pb_pairing(const uint8_t* p, const uint8_t* q, uint8_t* out, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(p && q && out);
  return piped(-1, {HIN(p, 3), HIN(q, 2), HOUT(out, 2)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_pairing_dev(di[0], di[1], dq[0], m, st); });
}
int pb_pairing_f_dev(uint64_t r, const uint8_t* p, const uint8_t* q, uint8_t* out, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(p && q && out && r >= 1);
  if (n == 0) return PB_OK;
  pairing_f_kernel<<<blocks_for(n, BLOCK), BLOCK, 0, S(stream)>>>(r, p, q, out, n);
  LAUNCH_CHECK("pairing_f_kernel");
  return PB_OK;
}
int pb_pairing_f(uint64_t r, const uint8_t* p, const uint8_t* q, uint8_t* out, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(p && q && out);
  return piped(-1, {HIN(p, 3), HIN(q, 2), HOUT(out, 2)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_pairing_f_dev(r, di[0], di[1], dq[0], m, st); });
}

// ------------------------------------------------------------------ context
int pb_ctx_create(pb_ctx** out, int device, const uint8_t circuit[PB_CIRCUIT_BYTES], const uint8_t* srs_g1s, uint32_t srs_len,
                  const uint8_t srs_g2[4]) {
  ARG(out && circuit && srs_g1s && srs_g2);
  ARG(srs_len >= 1 && srs_len <= PB_SRS_MAX);
  int rc = require_device();
  if (rc) return rc;
  for (int i = 0; i < 20; i++) ARG(circuit[i] < 17);
  for (int s = 0; s < 3; s++)
    for (int i = 0; i < 4; i++) ARG(circuit[24 + 8 * s + i] >= 1 && circuit[24 + 8 * s + i] <= 4);   // 1-based index into H (plonk.h:144)
  for (uint32_t i = 0; i < srs_len; i++) ARG(srs_g1s[3 * i] < 101 && srs_g1s[3 * i + 1] < 101);
  for (int i = 0; i < 4; i++) ARG(srs_g2[i] < 101);
  DeviceGuard g(device);
  if (!g.ok) return fail(PB_ERR_CUDA, "plonk_b200: cudaSetDevice failed");

  pb_ctx* c = new pb_ctx();
  c->device = device;
  c->srs_len = srs_len;
  c->srs_g1s.assign(srs_g1s, srs_g1s + 3 * srs_len);
  memcpy(c->srs_g2, srs_g2, 4);
  auto bail = [&](int code) { pb_ctx_destroy(c); return code; };

  // plonk_new (plonk.h:53-119): H, k1 H, k2 H, inverse Vandermonde matrix, Z_H
  uint8_t h[4], k1h[4], k2h[4], V[16], Vinv[16];
  for (int i = 0; i < 4; i++) h[i] = (uint8_t)pow17(4u, (uint64_t)i);
  for (int i = 0; i < 4; i++) { k1h[i] = (uint8_t)mul17(h[i], 2u); k2h[i] = (uint8_t)mul17(h[i], 3u); }
  for (int r = 0; r < 4; r++) for (int col = 0; col < 4; col++) V[r * 4 + col] = (uint8_t)pow17(h[r], (uint64_t)col);
  if ((rc = pb_matrix_inv(V, Vinv, 4, 1))) return bail(rc);
  for (int r = 0; r < 4; r++)
    for (int col = 0; col < 4; col++) {
      uint32_t s = 0;
      for (int k = 0; k < 4; k++) s += (uint32_t)V[r * 4 + k] * Vinv[k * 4 + col];
      if (red17(s) != (r == col ? 1u : 0u)) return bail(fail(PB_ERR_CUDA, "plonk_b200: h_pows_inv is not the inverse of the Vandermonde matrix"));
    }
  // Z_H = prod (x - h_i) must be x^4 - 1: prover.cuh divides by that shape
  uint8_t zh[8] = {1, 0, 0, 0, 0, 0, 0, 0};
  int zl = 1;
  for (int i = 0; i < 4; i++) {
    uint8_t t[8] = {0};
    for (int k = 0; k < zl; k++) { t[k] = (uint8_t)add17(t[k], mul17(zh[k], neg17(h[i]))); t[k + 1] = (uint8_t)add17(t[k + 1], zh[k]); }
    zl++;
    memcpy(zh, t, 8);
  }
  const uint8_t zh_expect[5] = {16, 0, 0, 0, 1};
  if (zl != 5 || memcmp(zh, zh_expect, 5) != 0) return bail(fail(PB_ERR_CUDA, "plonk_b200: Z_H is not x^4 - 1"));
  memcpy(c->setup, h, 4); memcpy(c->setup + 4, k1h, 4); memcpy(c->setup + 8, k2h, 4); memcpy(c->setup + 12, Vinv, 16);
  memcpy(c->setup + 28, zh, 8); c->setup[36] = 5;

  CircuitConst& cc = c->cc;
  for (int s = 0; s < 5; s++) for (int i = 0; i < 4; i++) cc.qv[s][i] = circuit[4 * s + i];
  for (int r = 0; r < 4; r++) for (int col = 0; col < 4; col++) cc.vinv[r][col] = Vinv[r * 4 + col];
  cc.srs_len = srs_len;
  cc.bad_copy = 0;
  cc.fs_seed = fs_seed_host(circuit, srs_g1s, srs_len, srs_g2);
  // copy_constraints_to_roots (plonk.h:142-160)
  uint8_t rows[9][4];
  for (int s = 0; s < 5; s++) memcpy(rows[s], circuit + 4 * s, 4);
  for (int s = 0; s < 3; s++)
    for (int i = 0; i < 4; i++) {
      uint8_t type = circuit[20 + 8 * s + i], idx = (uint8_t)(circuit[24 + 8 * s + i] - 1);
      uint8_t v = 0;
      if (type == 0) v = h[idx]; else if (type == 1) v = k1h[idx]; else if (type == 2) v = k2h[idx]; else cc.bad_copy = 1;
      rows[5 + s][i] = v;
      cc.sig[s][i] = v;
    }
  rows[8][0] = 1; rows[8][1] = rows[8][2] = rows[8][3] = 0;      // L1 = interpolate([1,0,0,0]), plonk.h:385-391
  // the nine witness-independent interpolations of plonk.h:268-275,391 -- on the device
  uint8_t polys[9][4], plens[9];
  if ((rc = pb_interpolate_at_h(c, &rows[0][0], &polys[0][0], plens, 9))) return bail(rc);
  for (int s = 0; s < 5; s++) for (int i = 0; i < 4; i++) cc.QP[s][i] = polys[s][i];
  for (int s = 0; s < 3; s++) for (int i = 0; i < 4; i++) cc.SP[s][i] = polys[5 + s][i];
  for (int i = 0; i < 4; i++) cc.l1[i] = polys[8][i];
  for (int s = 0; s < 3; s++) memcpy(c->circuit_dump + 4 * s, rows[5 + s], 4);
  for (int s = 0; s < 3; s++) memcpy(c->circuit_dump + 12 + 4 * s, polys[5 + s], 4);
  for (int s = 0; s < 5; s++) memcpy(c->circuit_dump + 24 + 4 * s, polys[s], 4);
  memcpy(c->circuit_dump + 44, polys[8], 4);

  // SRS fixed-base table T[i][c] = g1_mul(g1s[i], c), and the prover's nine-row copy with the field tables.
  // Every temporary is an RAII `Dev` (freed on every return path); every copy and launch is checked.
#define CUB(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return bail(cuda_fail(e__, #call)); } while (0)
#define DEVB(name, bytes) Dev name(bytes); if (name.err != cudaSuccess) return bail(cuda_fail(name.err, "cudaMalloc"))
#define LAUNCHB(what) do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return bail(cuda_fail(e__, what)); } while (0)
  const uint32_t rows_full = srs_len;
  ProverTables pt;
  memset(&pt, 0, sizeof pt);
  {
    DEVB(d_g1s, 3 * srs_len + 16);
    DEVB(d_prow, PROVER_SRS_ROWS * 17 * sizeof(uint32_t));
    CUB(cudaMemcpy(d_g1s.p, srs_g1s, 3 * srs_len, cudaMemcpyHostToDevice));
    CUB(cudaMalloc(&c->d_srs_table, rows_full * 17 * sizeof(uint32_t)));
    srs_table_kernel<<<1, 256>>>(d_g1s.as<uint8_t>(), srs_len, rows_full, c->d_srs_table);
    LAUNCHB("srs_table_kernel");
    srs_table_kernel<<<1, 256>>>(d_g1s.as<uint8_t>(), srs_len, PROVER_SRS_ROWS, d_prow.as<uint32_t>());
    LAUNCHB("srs_table_kernel");
    CUB(cudaMemcpy(pt.T, d_prow.p, sizeof pt.T, cudaMemcpyDeviceToHost));
  }
  for (uint32_t i = 0; i < 256; i++) pt.ft.inv101[i] = (uint8_t)pow101(i % 101u, 99);
  for (uint32_t i = 0; i < 17; i++) pt.ft.inv17[i] = (uint8_t)pow17(i, 15);
  for (uint32_t zz = 0; zz < 17; zz++) for (uint32_t k = 0; k < 20; k++) pt.pow17[zz][k] = (uint8_t)pow17(zz, k);
  for (uint32_t i = 0; i < MOD17_RANGE; i++) pt.mod17[i] = (uint8_t)(i % 17u);
  CUB(cudaMalloc(&c->d_tables, sizeof(ProverTables)));
  CUB(cudaMemcpy(c->d_tables, &pt, sizeof pt, cudaMemcpyHostToDevice));
  {
    // fast-path eligibility: canonical encodings (checked here) and curve membership (checked on the device)
    std::vector<uint8_t> on(srs_len);
    if ((rc = pb_g1_is_on_curve(srs_g1s, on.data(), srs_len))) return bail(rc);
    bool canon = true;
    for (uint32_t i = 0; i < srs_len; i++) {
      const uint8_t* g = srs_g1s + 3 * i;
      canon = canon && on[i] && g[2] <= 1 && !(g[2] == 1 && (g[0] | g[1]));
    }
    c->srs_canonical = canon;
    if (const char* e2 = getenv("PB_FORCE_EXACT")) c->force_exact = e2[0] == '1';
    if (canon) {
      std::vector<ProverPairTables> ppt_store(1);     // 7 KB: off the stack
      ProverPairTables& ppt = ppt_store[0];
      memset(&ppt, 0, sizeof ppt);
      ppt.ft = pt.ft;
      memcpy(ppt.pow17, pt.pow17, sizeof ppt.pow17);
      memcpy(ppt.mod17, pt.mod17, sizeof ppt.mod17);
      DEVB(d_single, sizeof pt.T);
      {
        DEVB(d_pairs, sizeof ppt.T2);
        CUB(cudaMemcpy(d_single.p, pt.T, sizeof pt.T, cudaMemcpyHostToDevice));
        pair_table_kernel<<<8, 256>>>(d_single.as<uint32_t>(), PROVER_SRS_ROWS, PROVER_PAIR_ROWS, d_pairs.as<uint32_t>());
        LAUNCHB("pair_table_kernel");
        CUB(cudaMemcpy(ppt.T2, d_pairs.p, sizeof ppt.T2, cudaMemcpyDeviceToHost));
      }
      CUB(cudaMalloc(&c->d_pair_tables, sizeof ppt));
      CUB(cudaMemcpy(c->d_pair_tables, &ppt, sizeof ppt, cudaMemcpyHostToDevice));
      const char* ew = getenv("PB_WIDE_TABLES");
      if (!(ew && ew[0] == '0')) {
        // one-look-up commitments: T3 for SRS rows 0-2 / 3-5 / 6-8, then T6 = T3[0] (+) T3[1]  (prover.cuh)
        const size_t n16 = 3u * (size_t)WIDE_T3_ENTRIES + WIDE_T6_ENTRIES;
        if (cudaMalloc(&c->d_wide_store, n16 * sizeof(uint16_t)) != cudaSuccess) {
          // no room for the 48 MB table: not an error, the context stays on the pair tables (same results)
          cudaGetLastError();
          c->d_wide_store = nullptr;
        } else {
          CUB(cudaMalloc(&c->d_wide_tables, sizeof(ProverWideTables)));
          uint16_t* t3 = c->d_wide_store;
          uint16_t* t6 = c->d_wide_store + 3u * (size_t)WIDE_T3_ENTRIES;
          wide_t3_kernel<<<58, 256>>>(d_single.as<uint32_t>(), PROVER_SRS_ROWS, t3);
          LAUNCHB("wide_t3_kernel");
          wide_t6_kernel<<<148 * 8, 256>>>(t3, t6);
          LAUNCHB("wide_t6_kernel");
          CUB(cudaDeviceSynchronize());
          ProverWideTables wt;
          memset(&wt, 0, sizeof wt);
          wt.ft = pt.ft;
          memcpy(wt.pow17, pt.pow17, sizeof wt.pow17);
          memcpy(wt.mod17, pt.mod17, sizeof wt.mod17);
          wt.T6 = t6;
          wt.T3 = t3 + 2u * (size_t)WIDE_T3_ENTRIES;
          CUB(cudaMemcpy(c->d_wide_tables, &wt, sizeof wt, cudaMemcpyHostToDevice));
        }
      }
    }
  }
  std::vector<uint32_t> tab(rows_full * 17);
  CUB(cudaMemcpy(tab.data(), c->d_srs_table, tab.size() * 4, cudaMemcpyDeviceToHost));
  c->table_bytes.resize(tab.size() * 3);
  for (size_t k = 0; k < tab.size(); k++) {
    c->table_bytes[3 * k] = tab[k] & 0xFF; c->table_bytes[3 * k + 1] = (tab[k] >> 8) & 0xFF; c->table_bytes[3 * k + 2] = (tab[k] >> 16) & 1;
  }

  // verifier key: srs_eval_at_s of q_M q_L q_R q_O q_C S1 S2 S3 (on the device)
  {
    const int order[8] = {3, 0, 1, 2, 4, 5, 6, 7};
    uint8_t kp[8][4];
    for (int j = 0; j < 8; j++) memcpy(kp[j], polys[order[j]], 4);
    uint32_t packed[8];
    {
      DEVB(d_kp, 32);
      DEVB(d_out, 32);
      CUB(cudaMemcpy(d_kp.p, kp, 32, cudaMemcpyHostToDevice));
      verifier_key_kernel<<<1, 32>>>(c->d_srs_table, srs_len, d_kp.as<uint8_t>(), d_out.as<uint32_t>());
      LAUNCHB("verifier_key_kernel");
      CUB(cudaMemcpy(packed, d_out.p, 32, cudaMemcpyDeviceToHost));
    }
    G1* dst[8] = {&c->vk.qm, &c->vk.ql, &c->vk.qr, &c->vk.qo, &c->vk.qc, &c->vk.s1, &c->vk.s2, &c->vk.s3};
    for (int j = 0; j < 8; j++) {
      if (packed[j] == 0xFFFFFFFFu) { c->vk_valid = false; packed[j] = pack_g1(0, 0, 1); }
      *dst[j] = G1{packed[j] & 0xFFu, (packed[j] >> 8) & 0xFFu, (packed[j] >> 16) & 1u};
      c->vkey_bytes[3 * j] = (uint8_t)dst[j]->x; c->vkey_bytes[3 * j + 1] = (uint8_t)dst[j]->y; c->vkey_bytes[3 * j + 2] = (uint8_t)dst[j]->inf;
    }
    c->vk.g1_one = G1{srs_g1s[0], srs_g1s[1], srs_g1s[2] ? 1u : 0u};
    c->vkey_bytes[24] = srs_g1s[0]; c->vkey_bytes[25] = srs_g1s[1]; c->vkey_bytes[26] = srs_g1s[2] ? 1 : 0;
    c->vk.g2_one = G2{srs_g2[0], srs_g2[1]};
    c->vk.g2_s = G2{srs_g2[2], srs_g2[3]};
    c->vk.fs_seed = c->cc.fs_seed;
    // fast-path verifier tables: only when all nine key points are canonically encoded curve points
    uint8_t on[9];
    if ((rc = pb_g1_is_on_curve(c->vkey_bytes, on, 9))) return bail(rc);
    bool canon = c->vk_valid;
    for (int j = 0; j < 9; j++) {
      const uint8_t* g = c->vkey_bytes + 3 * j;
      canon = canon && on[j] && g[2] <= 1 && !(g[2] == 1 && (g[0] | g[1]));
    }
    c->key_canonical = canon;
    if (canon) {
      DEVB(d_key, 32);
      DEVB(d_kt, 9 * 17 * 4);
      CUB(cudaMalloc(&c->d_verify_tables, sizeof(VerifyTables)));
      CUB(cudaMemcpy(d_key.p, c->vkey_bytes, 27, cudaMemcpyHostToDevice));
      srs_table_kernel<<<1, 256>>>(d_key.as<uint8_t>(), 9, 9, d_kt.as<uint32_t>());                  // rows c * K_j by the reference's g1_mul
      LAUNCHB("srs_table_kernel");
      verify_tables_kernel<<<1, 256>>>(d_kt.as<uint32_t>(), c->d_verify_tables);
      LAUNCHB("verify_tables_kernel");
      CUB(cudaDeviceSynchronize());
      const char* evt = getenv("PB_VERIFY_TABLES");
      if (!(evt && evt[0] == '0')) {
        DEVB(d_ok, 16);
        CUB(cudaMalloc(&c->d_verify_log, sizeof(VerifyLogTables)));
        verify_log_tables_kernel<<<1, 256>>>(c->vk, c->d_verify_tables, c->d_verify_log, d_ok.as<uint32_t>());
        LAUNCHB("verify_log_tables_kernel");
        uint32_t ok = 0;
        CUB(cudaMemcpy(&ok, d_ok.p, 4, cudaMemcpyDeviceToHost));
        if (!ok) { cudaFree(c->d_verify_log); c->d_verify_log = nullptr; }
      }
    }
  }
#undef CUB
#undef DEVB
#undef LAUNCHB
  *out = c;
  return PB_OK;
}

int pb_ctx_destroy(pb_ctx* c) {
  if (!c) return PB_OK;
  DeviceGuard g(c->device);
  if (c->d_tables) cudaFree(c->d_tables);
  if (c->d_pair_tables) cudaFree(c->d_pair_tables);
  if (c->d_wide_tables) cudaFree(c->d_wide_tables);
  if (c->d_wide_store) cudaFree(c->d_wide_store);
  for (auto& kv : c->scratch) if (kv.second.first) cudaFree(kv.second.first);
  if (c->d_verify_tables) cudaFree(c->d_verify_tables);
  if (c->d_verify_log) cudaFree(c->d_verify_log);
  if (c->d_srs_table) cudaFree(c->d_srs_table);
  for (auto& s : c->slots) {
    for (cudaEvent_t e : {s.ev_in, s.ev_k, s.ev_out}) if (e) cudaEventDestroy(e);
    void* bufs[4] = {s.in, s.proofs, s.dense, s.offs};
    for (auto b : bufs) if (b) cudaFree(b);
  }
  for (cudaStream_t st : {c->s_in, c->s_k, c->s_k2, c->s_out}) if (st) cudaStreamDestroy(st);
  for (uint8_t* p : {c->d_u, c->d_status, c->d_verdict, c->d_sv, c->d_wtab, c->d_seed_ws}) if (p) cudaFree(p);
  if (c->h_count) cudaFreeHost(c->h_count);
  delete c;
  return PB_OK;
}
int pb_ctx_setup_dump(const pb_ctx* ctx, uint8_t out[37]) { ARG(ctx && out); memcpy(out, ctx->setup, 37); return PB_OK; }
int pb_ctx_circuit_dump(const pb_ctx* ctx, uint8_t out[48]) { ARG(ctx && out); memcpy(out, ctx->circuit_dump, 48); return PB_OK; }
int pb_ctx_verifier_key(const pb_ctx* ctx, uint8_t out[27]) { ARG(ctx && out); memcpy(out, ctx->vkey_bytes, 27); return PB_OK; }
int pb_ctx_srs_table(const pb_ctx* ctx, uint8_t* out) { ARG(ctx && out); memcpy(out, ctx->table_bytes.data(), ctx->table_bytes.size()); return PB_OK; }

// ------------------------------------------------------------------ protocol
int pb_constraints_satisfy_dev(const pb_ctx* ctx, const uint8_t* witness, uint8_t* out, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && witness && out);
  { int rc_dev = on_ctx_device(ctx); if (rc_dev) return rc_dev; }
  if (n == 0) return PB_OK;
  satisfy_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(ctx->cc, witness, out, n);
  LAUNCH_CHECK("satisfy_kernel");
  return PB_OK;
}
int pb_constraints_satisfy(const pb_ctx* ctx, const uint8_t* witness, uint8_t* out, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && witness && out);
  return piped(ctx->device, {HIN(witness, 12), HOUT(out, 1)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_constraints_satisfy_dev(ctx, di[0], dq[0], m, st); });
}

int pb_constraints_satisfy_rows_dev(const uint8_t* selectors, uint32_t rows, const uint8_t* a, const uint8_t* b, const uint8_t* c, int32_t* first_bad,
                                    size_t n, void* stream) {
  if (n == 0) return PB_OK;
  ARG(selectors && a && b && c && first_bad && rows >= 1);
  satisfy_rows_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(selectors, rows, a, b, c, first_bad, n);
  LAUNCH_CHECK("satisfy_rows_kernel");
  return PB_OK;
}
int pb_constraints_satisfy_rows(const uint8_t* selectors, uint32_t rows, const uint8_t* a, const uint8_t* b, const uint8_t* c, int32_t* first_bad, size_t n) {
  if (n == 0) return PB_OK;
  ARG(selectors && a && b && c && first_bad && rows >= 1);
  int rc = require_device();
  if (rc) return rc;
  void* d_q = nullptr;
  CU(cudaMalloc(&d_q, 5 * (size_t)rows));
  cudaError_t e = cudaMemcpy(d_q, selectors, 5 * (size_t)rows, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    rc = piped(-1, {HIN(a, rows), HIN(b, rows), HIN(c, rows), HOUT(first_bad, 4)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
      return pb_constraints_satisfy_rows_dev(static_cast<const uint8_t*>(d_q), rows, di[0], di[1], di[2], reinterpret_cast<int32_t*>(dq[0]), m, st); });
  cudaFree(d_q);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  return rc;
}

// prove launch: pair tables when the SRS is canonical, the sequential tables otherwise
// chal == nullptr selects the Fiat-Shamir instantiation (challenges drawn in the kernel; chal_out optional);
// packed != nullptr selects the packed-input instantiation (wire.cuh): witness / rnd / chal are not read
static int launch_prove(const pb_ctx* ctx, const uint8_t* witness, const uint8_t* rnd, const uint8_t* chal, uint8_t* proofs, uint8_t* status,
                        size_t n, cudaStream_t st, uint32_t* done_list, uint32_t* done_count, uint8_t* verdict, uint8_t* chal_out = nullptr,
                        const uint8_t* packed = nullptr, int wire3 = 0) {
  const bool pair = ctx->srs_canonical && !ctx->force_exact;
  const bool wide = pair && ctx->d_wide_tables != nullptr;
  const unsigned grid = blocks_for(n, PBLOCK);
#define PB_LAUNCH_PROVE(TABLES, FSMODE, PK, TB, W, CH, CO) \
  prove_kernel<TABLES, FSMODE, PK><<<grid, PBLOCK, 0, st>>>(ctx->cc, TB, W, rnd, CH, proofs, status, n, done_list, done_count, verdict, CO, wire3)
  if (packed) {
    if (wide) PB_LAUNCH_PROVE(ProverWideTables, false, true, ctx->d_wide_tables, packed, nullptr, nullptr);
    else if (pair) PB_LAUNCH_PROVE(ProverPairTables, false, true, ctx->d_pair_tables, packed, nullptr, nullptr);
    else PB_LAUNCH_PROVE(ProverTables, false, true, ctx->d_tables, packed, nullptr, nullptr);
  } else if (chal) {
    if (wide) PB_LAUNCH_PROVE(ProverWideTables, false, false, ctx->d_wide_tables, witness, chal, nullptr);
    else if (pair) PB_LAUNCH_PROVE(ProverPairTables, false, false, ctx->d_pair_tables, witness, chal, nullptr);
    else PB_LAUNCH_PROVE(ProverTables, false, false, ctx->d_tables, witness, chal, nullptr);
  } else {
    if (wide) PB_LAUNCH_PROVE(ProverWideTables, true, false, ctx->d_wide_tables, witness, nullptr, chal_out);
    else if (pair) PB_LAUNCH_PROVE(ProverPairTables, true, false, ctx->d_pair_tables, witness, nullptr, chal_out);
    else PB_LAUNCH_PROVE(ProverTables, true, false, ctx->d_tables, witness, nullptr, chal_out);
  }
#undef PB_LAUNCH_PROVE
  LAUNCH_CHECK("prove_kernel");
  return PB_OK;
}
static bool table_verifier(const pb_ctx* ctx) { return ctx->key_canonical && !ctx->force_exact && ctx->d_verify_log != nullptr; }
static int launch_verify(const pb_ctx* ctx, const uint8_t* proofs, const uint8_t* chal, const uint8_t* u, const uint8_t* status,
                         const uint32_t* done_list, const uint32_t* done_count, uint8_t* verdict, uint8_t* gt, size_t n, cudaStream_t st,
                         const uint8_t* packed = nullptr, int wire3 = 0, int64_t* counts = nullptr) {
  const uint32_t* pk = reinterpret_cast<const uint32_t*>(packed);
  const bool fast = ctx->key_canonical && !ctx->force_exact && !(status && !done_list);
  if (table_verifier(ctx)) {     // the dense list, every item, or (status without a list) each block compacting its own items
    const uint8_t* by_status = done_list ? nullptr : status;
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(by_status ? counts : nullptr);   // fused tally: status mode only
    if (gt) verify_log_kernel<true><<<blocks_for(n, VLBLOCK), VLBLOCK, 0, st>>>(ctx->vk, ctx->d_verify_log, proofs, chal, u, done_list, done_count, verdict, gt, n, pk, wire3, by_status, cnt);
    else verify_log_kernel<false><<<blocks_for(n, VLBLOCK), VLBLOCK, 0, st>>>(ctx->vk, ctx->d_verify_log, proofs, chal, u, done_list, done_count, verdict, nullptr, n, pk, wire3, by_status, cnt);
  } else if (fast)
    if (gt) verify_fast_kernel<true><<<blocks_for(n, BLOCK), BLOCK, 0, st>>>(ctx->vk, ctx->d_verify_tables, proofs, chal, u, done_list, done_count, verdict, gt, n, pk, wire3);
    else verify_fast_kernel<false><<<blocks_for(n, BLOCK), BLOCK, 0, st>>>(ctx->vk, ctx->d_verify_tables, proofs, chal, u, done_list, done_count, verdict, nullptr, n, pk, wire3);
  else
    verify_kernel<<<blocks_for(n, BLOCK), BLOCK, 0, st>>>(ctx->vk, proofs, chal, u, status, verdict, gt, n, pk, wire3);
  LAUNCH_CHECK("verify_kernel");
  return PB_OK;
}

int pb_plonk_prove_dev(const pb_ctx* ctx, const uint8_t* witness, const uint8_t* rnd, const uint8_t* chal, uint8_t* proofs,
                       uint8_t* status, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && witness && rnd && chal && proofs && status);
  { int rc_dev = on_ctx_device(ctx); if (rc_dev) return rc_dev; }
  ARG(aligned16(witness) && aligned16(rnd) && aligned16(chal) && aligned16(proofs) && aligned16(status));
  if (n == 0) return PB_OK;
  return launch_prove(ctx, witness, rnd, chal, proofs, status, n, S(stream), nullptr, nullptr, nullptr);
}
int pb_plonk_verify_dev(const pb_ctx* ctx, const uint8_t* proofs, const uint8_t* chal, const uint8_t* u, uint8_t* verdict, uint8_t* gt,
                        size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && proofs && chal && u && verdict);
  { int rc_dev = on_ctx_device(ctx); if (rc_dev) return rc_dev; }
  ARG(ctx->vk_valid);
  ARG(aligned16(proofs) && aligned16(chal) && (!gt || aligned16(gt)));
  if (n == 0) return PB_OK;
  return launch_verify(ctx, proofs, chal, u, nullptr, nullptr, nullptr, verdict, gt, n, S(stream));
}
int pb_plonk_verify_completed_dev(const pb_ctx* ctx, const uint8_t* proofs, const uint8_t* chal, const uint8_t* u, const uint8_t* status,
                                  uint8_t* verdict, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && proofs && chal && u && status && verdict);
  { int rc_dev = on_ctx_device(ctx); if (rc_dev) return rc_dev; }
  ARG(ctx->vk_valid);
  ARG(aligned16(proofs) && aligned16(chal));
  ARG(n < 0xFFFFFFFFull);
  cudaStream_t st = S(stream);
  if (!(ctx->key_canonical && !ctx->force_exact) || table_verifier(ctx)) return launch_verify(ctx, proofs, chal, u, status, nullptr, nullptr, verdict, nullptr, n, st);
  uint32_t* scratch = nullptr;
  int rc = scratch_for(ctx, st, n, &scratch);
  if (rc) return rc;
  CU(cudaMemsetAsync(scratch, 0, 4 * sizeof(uint32_t), st));
  compact_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, st>>>(status, n, scratch + 4, scratch, verdict);
  return launch_verify(ctx, proofs, chal, u, status, scratch + 4, scratch, verdict, nullptr, n, st);
}
int pb_plonk_prove_verify_dev(const pb_ctx* ctx, const uint8_t* witness, const uint8_t* rnd, const uint8_t* chal, const uint8_t* u,
                              uint8_t* proofs, uint8_t* status, uint8_t* verdict, size_t n, void* stream) {
  return pb_plonk_prove_verify_ex_dev(ctx, witness, rnd, chal, u, proofs, status, verdict, n, stream, nullptr);
}
// chal == u == nullptr: Fiat-Shamir mode; packed != nullptr: packed input records instead of witness / rnd / chal / u
static int prove_verify_dev(const pb_ctx* ctx, const uint8_t* witness, const uint8_t* rnd, const uint8_t* chal, const uint8_t* u,
                            uint8_t* proofs, uint8_t* status, uint8_t* verdict, size_t n, void* stream, void* mid_event,
                            const uint8_t* packed = nullptr, int wire3 = 0, int64_t* counts = nullptr) {
  ARG(verdict);
  ARG(ctx && ctx->vk_valid);
  { int rc_dev = on_ctx_device(ctx); if (rc_dev) return rc_dev; }
  ARG((packed || (witness && rnd)) && proofs && status);
  ARG((chal == nullptr) == (u == nullptr));
  ARG(aligned16(witness) && aligned16(rnd) && aligned16(chal) && aligned16(proofs) && aligned16(status) && aligned16(packed));
  ARG(n < 0xFFFFFFFFull);
  // The prover appends the indices of the completed proofs to a dense list (stream-ordered scratch), the verifier walks
  // that list: no lane idles on the ~40% of random inputs on which the reference exits.
  cudaStream_t st = S(stream);
  if (table_verifier(ctx)) {
    // The table-path verifier compacts each block's items by their status bytes itself: no dense list in global memory,
    // no atomics in the prover (178.6 us against 184-186), no memset launch.
    int rcs = launch_prove(ctx, witness, rnd, chal, proofs, status, n, st, nullptr, nullptr, nullptr, nullptr, packed, wire3);
    if (mid_event) cudaEventRecord(reinterpret_cast<cudaEvent_t>(mid_event), st);
    if (!rcs) rcs = launch_verify(ctx, proofs, chal, u, status, nullptr, nullptr, verdict, nullptr, n, st, packed, wire3, counts);   // counters in the verifier's epilogue
    return rcs;
  }
  uint32_t* scratch = nullptr;
  int rc = scratch_for(ctx, st, n, &scratch);
  if (rc) return rc;
  CU(cudaMemsetAsync(scratch, 0, 4 * sizeof(uint32_t), st));
  rc = launch_prove(ctx, witness, rnd, chal, proofs, status, n, st, scratch + 4, scratch, verdict, nullptr, packed, wire3);
  if (mid_event) cudaEventRecord(reinterpret_cast<cudaEvent_t>(mid_event), st);
  if (!rc) rc = launch_verify(ctx, proofs, chal, u, status, scratch + 4, scratch, verdict, nullptr, n, st, packed, wire3);
  if (!rc && counts) rc = pb_tally_dev(proofs, status, verdict, n, counts, stream);      // no table path: the separate pass
  return rc;
}
int pb_plonk_prove_verify_tally_dev(const pb_ctx* ctx, const uint8_t* witness, const uint8_t* rnd, const uint8_t* chal, const uint8_t* u,
                                    uint8_t* proofs, uint8_t* status, uint8_t* verdict, int64_t* counts, size_t n, void* stream, void* mid_event) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(counts && chal && u);
  ARG((reinterpret_cast<uintptr_t>(counts) & 7u) == 0);      // 64-bit atomics
  return prove_verify_dev(ctx, witness, rnd, chal, u, proofs, status, verdict, n, stream, mid_event, nullptr, 0, counts);
}
int pb_plonk_prove_verify_ex_dev(const pb_ctx* ctx, const uint8_t* witness, const uint8_t* rnd, const uint8_t* chal, const uint8_t* u,
                                 uint8_t* proofs, uint8_t* status, uint8_t* verdict, size_t n, void* stream, void* mid_event) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(chal && u);
  return prove_verify_dev(ctx, witness, rnd, chal, u, proofs, status, verdict, n, stream, mid_event);
}

// ---- Fiat-Shamir mode (transcript.cuh; specification oracle/fs_spec.inc)
int pb_ctx_fs_seed(const pb_ctx* ctx, uint32_t out[4]) { ARG(ctx && out); memcpy(out, ctx->cc.fs_seed.v, 16); return PB_OK; }
int pb_plonk_prove_fs_dev(const pb_ctx* ctx, const uint8_t* witness, const uint8_t* rnd, uint8_t* proofs, uint8_t* status,
                          uint8_t* chal_out, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && witness && rnd && proofs && status);
  { int rc_dev = on_ctx_device(ctx); if (rc_dev) return rc_dev; }
  ARG(aligned16(witness) && aligned16(rnd) && aligned16(proofs) && aligned16(status));
  return launch_prove(ctx, witness, rnd, nullptr, proofs, status, n, S(stream), nullptr, nullptr, nullptr, chal_out);
}
int pb_plonk_verify_fs_dev(const pb_ctx* ctx, const uint8_t* proofs, uint8_t* verdict, uint8_t* gt, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && proofs && verdict);
  { int rc_dev = on_ctx_device(ctx); if (rc_dev) return rc_dev; }
  ARG(ctx->vk_valid);
  ARG(aligned16(proofs) && (!gt || aligned16(gt)));
  return launch_verify(ctx, proofs, nullptr, nullptr, nullptr, nullptr, nullptr, verdict, gt, n, S(stream));
}
int pb_plonk_prove_verify_fs_dev(const pb_ctx* ctx, const uint8_t* witness, const uint8_t* rnd, uint8_t* proofs, uint8_t* status,
                                 uint8_t* verdict, size_t n, void* stream, void* mid_event) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  return prove_verify_dev(ctx, witness, rnd, nullptr, nullptr, proofs, status, verdict, n, stream, mid_event);
}
int pb_fs_challenges_dev(const pb_ctx* ctx, const uint8_t* proofs, uint8_t* chal6, size_t n, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && proofs && chal6);
  { int rc_dev = on_ctx_device(ctx); if (rc_dev) return rc_dev; }
  fs_challenges_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(ctx->cc.fs_seed, proofs, chal6, n);
  LAUNCH_CHECK("fs_challenges_kernel");
  return PB_OK;
}

// dense list of the completed proofs of a chunk (stream-ordered): offs[0 .. groups) and the count in offs[cap / 128 + 3]
static int launch_gather(const uint8_t* proofs, const uint8_t* status, const uint8_t* verdict, size_t m, uint32_t* offs, uint32_t* count,
                         uint8_t* dense, uint8_t* sv, int pack /*0 structs, 1 packed v2, 2 packed v3*/, cudaStream_t st, uint32_t* count_host = nullptr) {
  done_counts_kernel<<<blocks_for((m + GBLOCK - 1) / GBLOCK, 256), 256, 0, st>>>(status, m, offs);
  done_offsets_kernel<<<1, 1024, 0, st>>>(m, offs, count, count_host);
  LAUNCH_CHECK("done_offsets_kernel");
  if (pack == 2) gather_done_kernel<2><<<blocks_for(m, GBLOCK), GBLOCK, 0, st>>>(proofs, status, verdict, m, offs, dense, sv);
  else if (pack == 1) gather_done_kernel<1><<<blocks_for(m, GBLOCK), GBLOCK, 0, st>>>(proofs, status, verdict, m, offs, dense, sv);
  else gather_done_kernel<0><<<blocks_for(m, GBLOCK), GBLOCK, 0, st>>>(proofs, status, verdict, m, offs, dense, sv);
  LAUNCH_CHECK("gather_done_kernel");
  return PB_OK;
}

// host-pointer versions: a three-stage pipeline over chunks, one stream per copy engine and two for the SMs -- s_in (H2D copy
// engine), s_k / s_k2 (SMs, chunks alternating), s_out (D2H copy engine) -- tied together by events, over a ring of PIPE_SLOTS
// buffer sets:
//   H2D(c) waits for kernels(c - SLOTS) (input buffers free);  kernels(c) wait for H2D(c) and D2H(c - SLOTS) (output buffers free);
//   D2H(c) waits for kernels(c).
// The input stage therefore never waits for the (slower) output stage, so once the inputs are up the D2H engine has the
// PCIe link to itself.  PCIe throughput on these hosts drops sharply for pieces below ~1 MB
// (profiles/r1/pcie_probe.txt), so the one-byte-per-item arrays (u in; status, verdict or sv out) are not chunked: u goes
// up with the first chunk, status and verdict come back after the last one.
//
// Modes (PipeMode): PROVE / STRUCT move the reference's structs both ways (27 B in, 34 + 2 B out per item).
// COMPACT: struct inputs; only the proofs that exist come back -- the completed ones, dense, in item order (the zero
// records of items on which the reference exits do not travel).  PACKED: packed wire v2 both ways (wire.cuh: 16 B in,
// 22 B per completed proof + 1 B per item out); PACKED3: packed wire v3 (14 B in, 12 B per completed proof + 1 B per item
// out; SRS on the curve only).  In the dense modes the size of a chunk's D2H copy is known only when its
// kernels have run: the host waits for chunk c - PIPE_LOOKAHEAD's kernels (the GPU has the chunks in between queued),
// reads the count from pinned memory and issues that chunk's copy.
static int pipeline(const pb_ctx* cctx, const uint8_t* witness, const uint8_t* rnd, const uint8_t* chal, const uint8_t* u,
                    uint8_t* proofs, uint8_t* status, uint8_t* verdict, size_t n, int mode, size_t* n_done = nullptr) {
  pb_ctx* ctx = const_cast<pb_ctx*>(cctx);
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(PB_ERR_CUDA, "plonk_b200: cudaSetDevice failed");
  std::lock_guard<std::mutex> lock(ctx->pipe_mu);
  const size_t chunk = pipe_chunk();
  int rc = pipe_init(ctx, n < chunk ? n : chunk);
  if (rc) return rc;
  if (ctx->small_cap < n) {
    for (uint8_t** p : {&ctx->d_u, &ctx->d_status, &ctx->d_verdict, &ctx->d_sv}) { if (*p) CU(cudaFree(*p)); *p = nullptr; }
    ctx->small_cap = 0;
    size_t cap = (n + 255) & ~(size_t)255;
    CU(cudaMalloc(&ctx->d_u, cap)); CU(cudaMalloc(&ctx->d_status, cap)); CU(cudaMalloc(&ctx->d_verdict, cap)); CU(cudaMalloc(&ctx->d_sv, cap));
    ctx->small_cap = cap;
  }
  const bool wire3 = mode == PIPE_PACKED3, packed = mode == PIPE_PACKED || wire3, dense = mode == PIPE_COMPACT || packed;
  const size_t rec = wire3 ? PACKED3_PROOF_BYTES : packed ? PACKED_PROOF_BYTES : 34;
  const size_t rec_in = wire3 ? PACKED3_IN_BYTES : PACKED_IN_BYTES;
  // Chunk schedule.  The H2D engine is the stage that is busy from the first byte to the last, so what is exposed is the
  // time AFTER the last input byte has landed: the last chunk's kernels and its D2H copy.  Full-size chunks first (large
  // copies run at the link's rate, profiles/r1/pcie_probe.txt), then a tail that halves down to PIPE_TAIL_MIN items.
  // PB_PIPE_RAMP_UP=1 also ramps the head up from PIPE_TAIL_MIN (round 1's symmetric schedule).
  std::vector<size_t> sched;
  {
    static const size_t small_env = getenv("PB_PIPE_SMALL") ? strtoull(getenv("PB_PIPE_SMALL"), nullptr, 10) & ~(size_t)127 : 0;
    static const bool ramp_up = getenv("PB_PIPE_RAMP_UP") && getenv("PB_PIPE_RAMP_UP")[0] == '1';
    const size_t small = small_env >= 128 ? small_env : PIPE_TAIL_MIN;
    const size_t ragged = n % 128;          // every chunk but the very last starts at a multiple of 128 items (16-byte aligned slices)
    size_t left = n - ragged;
    std::vector<size_t> tail;               // ascending
    for (size_t t = small; t < chunk && left > 2 * t; t *= 2) { tail.push_back(t); left -= t; }
    if (ramp_up)
      for (size_t t = small; t < chunk && left > 2 * t; t *= 2) { sched.push_back(t); left -= t; }
    while (left > 0) { size_t m = left < chunk ? left : chunk; sched.push_back(m); left -= m; }
    for (size_t k = tail.size(); k-- > 0;) sched.push_back(tail[k]);
    if (ragged) sched.push_back(ragged);
  }
  // PB_PIPE_TRACE=1: print, per chunk, when its inputs landed / its kernels finished / its proofs were copied out
  // PB_PIPE_TRACE=2: only the host-side line (no events between the stages, which cost API calls of their own)
  static const int trace_level = getenv("PB_PIPE_TRACE") ? atoi(getenv("PB_PIPE_TRACE")) : 0;
  const bool trace = trace_level == 1;
  std::vector<cudaEvent_t> tev;
  cudaEvent_t t0 = nullptr;
  if (trace) {
    tev.resize(3 * sched.size());
    for (auto& e : tev) CU(cudaEventCreate(&e));
    CU(cudaEventCreate(&t0));
    CU(cudaEventRecord(t0, ctx->s_in));
  }
  size_t total_done = 0;
  // compact / packed modes: the chunk's kernels have run -> its count is in pinned memory -> issue its D2H copy
  auto finish = [&](size_t c) -> int {
    PipeSlot& s = ctx->slots[c % PIPE_SLOTS];
    CU(cudaEventSynchronize(s.ev_k));
    const size_t cnt = ctx->h_count[c % PIPE_SLOTS];
    if (cnt > sched[c]) return fail(PB_ERR_CUDA, "plonk_b200: dense-list count exceeds the chunk");
    if (cnt) CU(cudaMemcpyAsync(proofs + total_done * rec, s.dense, cnt * rec, cudaMemcpyDeviceToHost, ctx->s_out));
    CU(cudaEventRecord(s.ev_out, ctx->s_out));
    if (trace) CU(cudaEventRecord(tev[3 * c + 2], ctx->s_out));
    total_done += cnt;
    return PB_OK;
  };
  size_t done = 0, c = 0;
  const auto host_t0 = std::chrono::steady_clock::now();
  auto host_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
  double host_enq = 0;
  for (size_t m : sched) {
    if (dense && c >= (size_t)PIPE_LOOKAHEAD && (rc = finish(c - PIPE_LOOKAHEAD))) return rc;
    PipeSlot& s = ctx->slots[c % PIPE_SLOTS];
    const bool reuse = c >= (size_t)PIPE_SLOTS;
    uint8_t *d_wit = s.in, *d_rnd = s.in + ctx->slot_cap * 12, *d_chal = s.in + ctx->slot_cap * 21;
    if (reuse) CU(cudaStreamWaitEvent(ctx->s_in, s.ev_k, 0));
    if (packed) {
      CU(cudaMemcpyAsync(s.in, witness + done * rec_in, m * rec_in, cudaMemcpyHostToDevice, ctx->s_in));
    } else {
      CU(cudaMemcpyAsync(d_wit, witness + done * 12, m * 12, cudaMemcpyHostToDevice, ctx->s_in));
      CU(cudaMemcpyAsync(d_rnd, rnd + done * 9, m * 9, cudaMemcpyHostToDevice, ctx->s_in));
      if (chal) CU(cudaMemcpyAsync(d_chal, chal + done * 5, m * 5, cudaMemcpyHostToDevice, ctx->s_in));   // absent in Fiat-Shamir mode
      if (mode != PIPE_PROVE && c == 0 && u) CU(cudaMemcpyAsync(ctx->d_u, u, n, cudaMemcpyHostToDevice, ctx->s_in));
    }
    CU(cudaEventRecord(s.ev_in, ctx->s_in));
    if (trace) CU(cudaEventRecord(tev[3 * c], ctx->s_in));
    // Two compute streams, chunks alternating: the latency-bound tail of one chunk (the last wave of its kernels, the small
    // count / offset / gather launches and the gaps between dependent launches) overlaps the next chunk's prover.
    cudaStream_t sk = PB_PIPE_TWO_STREAMS && (c & 1) ? ctx->s_k2 : ctx->s_k;
    CU(cudaStreamWaitEvent(sk, s.ev_in, 0));
    if (reuse) CU(cudaStreamWaitEvent(sk, s.ev_out, 0));
    if (packed)
      rc = prove_verify_dev(ctx, nullptr, nullptr, nullptr, nullptr, s.proofs, ctx->d_status + done, ctx->d_verdict + done, m, sk, nullptr, s.in, wire3);
    else if (mode != PIPE_PROVE)
      rc = prove_verify_dev(ctx, d_wit, d_rnd, chal ? d_chal : nullptr, u ? ctx->d_u + done : nullptr, s.proofs, ctx->d_status + done,
                            ctx->d_verdict + done, m, sk, nullptr);
    else if (chal)
      rc = pb_plonk_prove_dev(ctx, d_wit, d_rnd, d_chal, s.proofs, ctx->d_status + done, m, sk);
    else
      rc = pb_plonk_prove_fs_dev(ctx, d_wit, d_rnd, s.proofs, ctx->d_status + done, nullptr, m, sk);
    if (rc) return rc;
    if (dense) {
      uint32_t* cnt = s.offs + ctx->slot_cap / 128 + 3;
      // the count also lands in mapped pinned memory (a store by the kernel, not a copy queued behind the D2H engine's proofs)
      if ((rc = launch_gather(s.proofs, ctx->d_status + done, ctx->d_verdict + done, m, s.offs, cnt, s.dense, ctx->d_sv + done, wire3 ? 2 : packed ? 1 : 0, sk,
                              ctx->h_count_dev + c % PIPE_SLOTS))) return rc;
    }
    CU(cudaEventRecord(s.ev_k, sk));
    if (trace) CU(cudaEventRecord(tev[3 * c + 1], sk));
    if (!dense) {
      CU(cudaStreamWaitEvent(ctx->s_out, s.ev_k, 0));
      CU(cudaMemcpyAsync(proofs + done * 34, s.proofs, m * 34, cudaMemcpyDeviceToHost, ctx->s_out));
      CU(cudaEventRecord(s.ev_out, ctx->s_out));
      if (trace) CU(cudaEventRecord(tev[3 * c + 2], ctx->s_out));
    }
    done += m;
    c++;
  }
  host_enq = host_ms();
  if (dense)
    for (size_t k = c > (size_t)PIPE_LOOKAHEAD ? c - PIPE_LOOKAHEAD : 0; k < c; k++)
      if ((rc = finish(k))) return rc;
  // s_out is ordered after the last kernels: ev_k of the last chunk of each compute stream
  if (!dense && c) CU(cudaStreamWaitEvent(ctx->s_out, ctx->slots[(c - 1) % PIPE_SLOTS].ev_k, 0));
  if (!dense && c > 1) CU(cudaStreamWaitEvent(ctx->s_out, ctx->slots[(c - 2) % PIPE_SLOTS].ev_k, 0));
  if (packed) {
    CU(cudaMemcpyAsync(status /* = sv */, ctx->d_sv, n, cudaMemcpyDeviceToHost, ctx->s_out));
  } else {
    CU(cudaMemcpyAsync(status, ctx->d_status, n, cudaMemcpyDeviceToHost, ctx->s_out));
    if (mode != PIPE_PROVE) CU(cudaMemcpyAsync(verdict, ctx->d_verdict, n, cudaMemcpyDeviceToHost, ctx->s_out));
  }
  CU(cudaStreamSynchronize(ctx->s_out));
  CU(cudaStreamSynchronize(ctx->s_in));
  CU(cudaStreamSynchronize(ctx->s_k));
  CU(cudaStreamSynchronize(ctx->s_k2));
  if (n_done) *n_done = total_done;
  if (trace_level) fprintf(stderr, "pipe host: %zu chunks enqueued at %.3f ms, call done at %.3f ms\n", sched.size(), host_enq, host_ms());
  if (trace) {
    for (size_t k = 0; k < sched.size(); k++) {
      float a = 0, b = 0, d = 0;
      cudaEventElapsedTime(&a, t0, tev[3 * k]); cudaEventElapsedTime(&b, t0, tev[3 * k + 1]); cudaEventElapsedTime(&d, t0, tev[3 * k + 2]);
      fprintf(stderr, "pipe chunk %2zu items %7zu  in %.3f ms  kernels %.3f ms  out %.3f ms\n", k, sched[k], a, b, d);
    }
    for (auto& e : tev) cudaEventDestroy(e);
    cudaEventDestroy(t0);
  }
  return PB_OK;
}
int pb_plonk_prove(const pb_ctx* ctx, const uint8_t* witness, const uint8_t* rnd, const uint8_t* chal, uint8_t* proofs, uint8_t* status, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && witness && rnd && chal && proofs && status);
  return pipeline(ctx, witness, rnd, chal, nullptr, proofs, status, nullptr, n, PIPE_PROVE);
}
int pb_plonk_prove_verify(const pb_ctx* ctx, const uint8_t* witness, const uint8_t* rnd, const uint8_t* chal, const uint8_t* u,
                          uint8_t* proofs, uint8_t* status, uint8_t* verdict, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && witness && rnd && chal && u && proofs && status && verdict);
  ARG(ctx->vk_valid);
  return pipeline(ctx, witness, rnd, chal, u, proofs, status, verdict, n, PIPE_STRUCT);
}
// ---- compact output, packed wire v2, seeded mode (wire.cuh) ------------------------------------------------------------
int pb_plonk_prove_verify_compact(const pb_ctx* ctx, const uint8_t* witness, const uint8_t* rnd, const uint8_t* chal, const uint8_t* u,
                                  uint8_t* proofs_dense, size_t* n_done, uint8_t* status, uint8_t* verdict, size_t n) {
  if (n_done) *n_done = 0;
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && witness && rnd && chal && u && proofs_dense && n_done && status && verdict);
  ARG(ctx->vk_valid);
  return pipeline(ctx, witness, rnd, chal, u, proofs_dense, status, verdict, n, PIPE_COMPACT, n_done);
}
int pb_plonk_prove_verify_packed(const pb_ctx* ctx, const uint8_t* packed_in, uint8_t* packed_proofs, size_t* n_done, uint8_t* sv, size_t n) {
  if (n_done) *n_done = 0;
  if (n == 0) return PB_OK;
  ARG(ctx && packed_in && packed_proofs && n_done && sv);
  ARG(ctx->vk_valid);
  return pipeline(ctx, packed_in, nullptr, nullptr, nullptr, packed_proofs, sv, nullptr, n, PIPE_PACKED, n_done);
}
int pb_plonk_prove_verify_packed3(const pb_ctx* ctx, const uint8_t* packed_in, uint8_t* packed_proofs, size_t* n_done, uint8_t* sv, size_t n) {
  if (n_done) *n_done = 0;
  if (n == 0) return PB_OK;
  ARG(ctx && packed_in && packed_proofs && n_done && sv);
  ARG(ctx->vk_valid);
  if (!ctx->srs_canonical) return fail(PB_ERR_ARG, "plonk_b200: packed wire v3 needs an SRS whose points are canonical points of the curve (use v2)");
  return pipeline(ctx, packed_in, nullptr, nullptr, nullptr, packed_proofs, sv, nullptr, n, PIPE_PACKED3, n_done);
}
size_t pb_packed_workspace_bytes(size_t n) {
  const size_t cap = (n + 127) & ~(size_t)127;
  return cap * 34 + cap + cap + (cap / 128 + 4) * sizeof(uint32_t) + 64;   // proofs | status | verdict | offs
}
static int packed_dev(const pb_ctx* ctx, const uint8_t* packed_in, uint8_t* packed_proofs, uint32_t* n_done_dev, uint8_t* sv,
                      void* workspace, size_t n, void* stream, int wire3) {
  if (n == 0) return PB_OK;
  ARG(ctx && packed_in && packed_proofs && n_done_dev && sv && workspace);
  ARG(aligned16(packed_in) && aligned16(packed_proofs) && aligned16(workspace));
  if (wire3 && !ctx->srs_canonical) return fail(PB_ERR_ARG, "plonk_b200: packed wire v3 needs an SRS whose points are canonical points of the curve (use v2)");
  const size_t cap = (n + 127) & ~(size_t)127;
  uint8_t* proofs = static_cast<uint8_t*>(workspace);
  uint8_t* status = proofs + cap * 34;
  uint8_t* verdict = status + cap;
  uint32_t* offs = reinterpret_cast<uint32_t*>(verdict + cap);
  int rc = prove_verify_dev(ctx, nullptr, nullptr, nullptr, nullptr, proofs, status, verdict, n, stream, nullptr, packed_in, wire3);
  if (rc) return rc;
  return launch_gather(proofs, status, verdict, n, offs, n_done_dev, packed_proofs, sv, wire3 ? 2 : 1, S(stream));
}
int pb_plonk_prove_verify_packed_dev(const pb_ctx* ctx, const uint8_t* packed_in, uint8_t* packed_proofs, uint32_t* n_done_dev, uint8_t* sv,
                                     void* workspace, size_t n, void* stream) {
  return packed_dev(ctx, packed_in, packed_proofs, n_done_dev, sv, workspace, n, stream, 0);
}
int pb_plonk_prove_verify_packed3_dev(const pb_ctx* ctx, const uint8_t* packed_in, uint8_t* packed_proofs, uint32_t* n_done_dev, uint8_t* sv,
                                      void* workspace, size_t n, void* stream) {
  return packed_dev(ctx, packed_in, packed_proofs, n_done_dev, sv, workspace, n, stream, 1);
}
int pb_gather_completed_dev(const uint8_t* proofs, const uint8_t* status, uint8_t* proofs_dense, uint32_t* n_done_dev, uint32_t* offs_scratch,
                            size_t n, void* stream) {
  if (n == 0) return PB_OK;
  ARG(proofs && status && proofs_dense && n_done_dev && offs_scratch);
  ARG(aligned16(proofs) && aligned16(proofs_dense));
  return launch_gather(proofs, status, nullptr, n, offs_scratch, n_done_dev, proofs_dense, nullptr, 0, S(stream));
}

// format conversion on the host (no arithmetic of the path): the reference's structs <-> packed wire v2
int pb_wire_pack_inputs(const uint8_t* witness, const uint8_t* rnd, const uint8_t* chal, const uint8_t* u, uint8_t* packed, size_t n) {
  if (n == 0) return PB_OK;
  ARG(witness && rnd && chal && u && packed);
  for (size_t i = 0; i < n; i++) {
    uint32_t v[PACKED_VALUES], w[4];
    for (int k = 0; k < 12; k++) v[k] = witness[i * 12 + k];
    for (int k = 0; k < 9; k++) v[12 + k] = rnd[i * 9 + k];
    for (int k = 0; k < 5; k++) v[21 + k] = chal[i * 5 + k];
    v[26] = u[i];
    bool ok = true;
    for (int k = 0; k < PACKED_VALUES; k++) ok = ok && v[k] < 17u;
    if (ok) pack_input16(v, w);
    else w[0] = w[1] = w[2] = w[3] = 0xFFFFFFFFu;   // not an encoding: the prover reports PB_PROVE_BAD_INPUT, as for the struct bytes
    for (int k = 0; k < 4; k++) for (int b = 0; b < 4; b++) packed[i * 16 + 4 * k + b] = (uint8_t)(w[k] >> (8 * b));
  }
  return PB_OK;
}
int pb_wire_unpack_inputs(const uint8_t* packed, uint8_t* witness, uint8_t* rnd, uint8_t* chal, uint8_t* u, uint8_t* valid, size_t n) {
  if (n == 0) return PB_OK;
  ARG(packed && witness && rnd && chal && u);
  for (size_t i = 0; i < n; i++) {
    uint32_t w[4], v[PACKED_VALUES];
    for (int k = 0; k < 4; k++) { w[k] = 0; for (int b = 0; b < 4; b++) w[k] |= (uint32_t)packed[i * 16 + 4 * k + b] << (8 * b); }
    const bool ok = unpack_input16(w[0], w[1], w[2], w[3], v);
    for (int k = 0; k < 12; k++) witness[i * 12 + k] = ok ? (uint8_t)v[k] : 0xFF;
    for (int k = 0; k < 9; k++) rnd[i * 9 + k] = ok ? (uint8_t)v[12 + k] : 0xFF;
    for (int k = 0; k < 5; k++) chal[i * 5 + k] = ok ? (uint8_t)v[21 + k] : 0xFF;
    u[i] = ok ? (uint8_t)v[26] : 0xFF;
    if (valid) valid[i] = ok ? 1 : 0;
  }
  return PB_OK;
}
int pb_wire_pack_proofs(const uint8_t* proofs, uint8_t* packed, size_t n) {
  if (n == 0) return PB_OK;
  ARG(proofs && packed);
  for (size_t i = 0; i < n; i++) {
    uint16_t r[11];
    pack_proof22(proofs + i * 34, r);
    for (int k = 0; k < 11; k++) { packed[i * 22 + 2 * k] = (uint8_t)(r[k] & 0xFF); packed[i * 22 + 2 * k + 1] = (uint8_t)(r[k] >> 8); }
  }
  return PB_OK;
}
int pb_wire_unpack_proofs(const uint8_t* packed, uint8_t* proofs, size_t n) {
  if (n == 0) return PB_OK;
  ARG(packed && proofs);
  for (size_t i = 0; i < n; i++) {
    uint16_t r[11];
    for (int k = 0; k < 11; k++) r[k] = (uint16_t)(packed[i * 22 + 2 * k] | packed[i * 22 + 2 * k + 1] << 8);
    if (!unpack_proof22(r, proofs + i * 34)) return fail(PB_ERR_ARG, "plonk_b200: packed proof record is not a canonical encoding");
  }
  return PB_OK;
}
// the same for packed wire v3 (14-byte input records, 12-byte proof records of points ON THE CURVE)
static const CurveIndexImage k_curve_index = CurveIndexImage();
int pb_wire3_pack_inputs(const uint8_t* witness, const uint8_t* rnd, const uint8_t* chal, const uint8_t* u, uint8_t* packed, size_t n) {
  if (n == 0) return PB_OK;
  ARG(witness && rnd && chal && u && packed);
  for (size_t i = 0; i < n; i++) {
    uint32_t v[PACKED_VALUES];
    uint16_t h[7];
    for (int k = 0; k < 12; k++) v[k] = witness[i * 12 + k];
    for (int k = 0; k < 9; k++) v[12 + k] = rnd[i * 9 + k];
    for (int k = 0; k < 5; k++) v[21 + k] = chal[i * 5 + k];
    v[26] = u[i];
    bool ok = true;
    for (int k = 0; k < PACKED_VALUES; k++) ok = ok && v[k] < 17u;
    if (ok) pack_input14(v, h);
    else for (int k = 0; k < 7; k++) h[k] = 0xFFFFu;   // not an encoding: the prover reports PB_PROVE_BAD_INPUT
    for (int k = 0; k < 7; k++) { packed[i * 14 + 2 * k] = (uint8_t)(h[k] & 0xFF); packed[i * 14 + 2 * k + 1] = (uint8_t)(h[k] >> 8); }
  }
  return PB_OK;
}
int pb_wire3_unpack_inputs(const uint8_t* packed, uint8_t* witness, uint8_t* rnd, uint8_t* chal, uint8_t* u, uint8_t* valid, size_t n) {
  if (n == 0) return PB_OK;
  ARG(packed && witness && rnd && chal && u);
  for (size_t i = 0; i < n; i++) {
    uint16_t h[7];
    uint32_t w[4], v[PACKED_VALUES];
    for (int k = 0; k < 7; k++) h[k] = (uint16_t)(packed[i * 14 + 2 * k] | packed[i * 14 + 2 * k + 1] << 8);
    input14_words(h, w);
    const bool ok = unpack_input16(w[0], w[1], w[2], w[3], v);
    for (int k = 0; k < 12; k++) witness[i * 12 + k] = ok ? (uint8_t)v[k] : 0xFF;
    for (int k = 0; k < 9; k++) rnd[i * 9 + k] = ok ? (uint8_t)v[12 + k] : 0xFF;
    for (int k = 0; k < 5; k++) chal[i * 5 + k] = ok ? (uint8_t)v[21 + k] : 0xFF;
    u[i] = ok ? (uint8_t)v[26] : 0xFF;
    if (valid) valid[i] = ok ? 1 : 0;
  }
  return PB_OK;
}
int pb_wire3_pack_proofs(const uint8_t* proofs, uint8_t* packed, size_t n) {
  if (n == 0) return PB_OK;
  ARG(proofs && packed);
  for (size_t i = 0; i < n; i++) {
    const uint8_t* r = proofs + i * 34;
    for (int j = 0; j < 9; j++) {        // every point must be one of the 102: otherwise the record has no v3 encoding
      const uint32_t x = r[3 * j], y = r[3 * j + 1];
      const bool on = r[3 * j + 2] ? (x == 0 && y == 0) : (x < 101u && y < 101u && y * y % 101u == (x * x % 101u * x + 3u) % 101u);
      if (!on) return fail(PB_ERR_ARG, "plonk_b200: a proof point is not a canonical point of the curve: no packed v3 encoding");
    }
    for (int j = 0; j < 7; j++) if (r[27 + j] >= 17u) return fail(PB_ERR_ARG, "plonk_b200: an opening is not a canonical F17 element");
    uint32_t w[3];
    pack_proof12(r, k_curve_index.base, w);
    for (int k = 0; k < 3; k++) for (int b = 0; b < 4; b++) packed[i * 12 + 4 * k + b] = (uint8_t)(w[k] >> (8 * b));
  }
  return PB_OK;
}
int pb_wire3_unpack_proofs(const uint8_t* packed, uint8_t* proofs, size_t n) {
  if (n == 0) return PB_OK;
  ARG(packed && proofs);
  for (size_t i = 0; i < n; i++) {
    uint32_t w[3];
    for (int k = 0; k < 3; k++) { w[k] = 0; for (int b = 0; b < 4; b++) w[k] |= (uint32_t)packed[i * 12 + 4 * k + b] << (8 * b); }
    if (!unpack_proof12(w, k_curve_index.px, k_curve_index.py, proofs + i * 34))
      return fail(PB_ERR_ARG, "plonk_b200: packed v3 proof record is not a canonical encoding");
  }
  return PB_OK;
}
// dense completed proofs + status bytes -> the full [n][34] array of the struct API (zero records where status != 0)
int pb_wire_scatter_proofs(const uint8_t* proofs_dense, const uint8_t* status, uint8_t* proofs, size_t n) {
  if (n == 0) return PB_OK;
  ARG(proofs_dense && status && proofs);
  size_t k = 0;
  for (size_t i = 0; i < n; i++) {
    if (status[i] == 0) memcpy(proofs + i * 34, proofs_dense + (k++) * 34, 34);
    else memset(proofs + i * 34, 0, 34);
  }
  return PB_OK;
}
int pb_wire_split_sv(const uint8_t* sv, uint8_t* status, uint8_t* verdict, size_t n) {
  if (n == 0) return PB_OK;
  ARG(sv);
  for (size_t i = 0; i < n; i++) {
    const uint8_t s = sv[i] & 15u, v = sv[i] >> 4;
    if (status) status[i] = s == 15u ? 254 : s;
    if (verdict) verdict[i] = v == 15u ? 0xFF : v;
  }
  return PB_OK;
}

// seeded mode: inputs are generated on the device from (seed, start, count) -- the counter-based stream of
// plonk.c_b200/workload.py -- and only the 18 counters come back.  No batch data crosses PCIe.
static int seeded_tables(pb_ctx* ctx) {
  if (ctx->d_wtab) return PB_OK;
  uint8_t tab[SYNTH_WITNESS_ROWS * 12];
  int rows = 0;
  for (uint32_t x = 0; x < 17; x++)
    for (uint32_t y = 0; y < 17; y++)
      for (uint32_t z = 0; z < 17; z++)
        if ((x * x + y * y) % 17u == (z * z) % 17u) {
          const uint8_t xx = (uint8_t)(x * x % 17u), yy = (uint8_t)(y * y % 17u), zz = (uint8_t)(z * z % 17u);
          const uint8_t row[12] = {(uint8_t)x, (uint8_t)y, (uint8_t)z, xx, (uint8_t)x, (uint8_t)y, (uint8_t)z, yy, xx, yy, zz, zz};
          if (rows < SYNTH_WITNESS_ROWS) memcpy(tab + 12 * rows, row, 12);
          rows++;
        }
  if (rows != SYNTH_WITNESS_ROWS) return fail(PB_ERR_CUDA, "plonk_b200: witness table does not have 289 rows");
  CU(cudaMalloc(&ctx->d_wtab, sizeof tab));
  CU(cudaMemcpy(ctx->d_wtab, tab, sizeof tab, cudaMemcpyHostToDevice));
  return PB_OK;
}
size_t pb_seeded_workspace_bytes(size_t n) { return ((n + 127) & ~(size_t)127) * 16 + pb_packed_workspace_bytes(n); }
int pb_synth_batch_dev(const pb_ctx* cctx, uint64_t seed, uint64_t start, size_t n, int variant, uint8_t* witness, uint8_t* rnd, uint8_t* chal,
                       uint8_t* u, uint8_t* packed, void* stream) {
  if (n == 0) return PB_OK;
  pb_ctx* ctx = const_cast<pb_ctx*>(cctx);
  ARG(ctx && (variant == 0 || variant == 1));
  ARG(packed || (witness && rnd && chal && u));
  ARG(aligned16(packed));
  if (!ctx->d_wtab) {   // first use on this context (pb_plonk_prove_verify_seeded has done this already, under the same lock)
    std::lock_guard<std::mutex> lock(ctx->pipe_mu);
    int rc = seeded_tables(ctx);
    if (rc) return rc;
  }
  if (packed) synth_packed_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(seed, start, variant, ctx->d_wtab, packed, n);
  if (witness) synth_struct_kernel<<<blocks_for(n, BLOCK_LIGHT), BLOCK_LIGHT, 0, S(stream)>>>(seed, start, variant, ctx->d_wtab, witness, rnd, chal, u, n);
  LAUNCH_CHECK("synth_kernel");
  return PB_OK;
}
int pb_plonk_prove_verify_seeded_dev(const pb_ctx* ctx, uint64_t seed, uint64_t start, size_t n, int variant, void* workspace,
                                     int64_t* counts_dev, void* stream) {
  if (n == 0) return PB_OK;
  ARG(ctx && workspace && counts_dev && aligned16(workspace));
  ARG(ctx->vk_valid);
  const size_t cap = (n + 127) & ~(size_t)127;
  uint8_t* packed = static_cast<uint8_t*>(workspace);
  uint8_t* proofs = packed + cap * 16;
  uint8_t* status = proofs + cap * 34;
  uint8_t* verdict = status + cap;
  int rc = pb_synth_batch_dev(ctx, seed, start, n, variant, nullptr, nullptr, nullptr, nullptr, packed, stream);
  if (!rc) rc = prove_verify_dev(ctx, nullptr, nullptr, nullptr, nullptr, proofs, status, verdict, n, stream, nullptr, packed, 0, counts_dev);
  return rc;
}
int pb_plonk_prove_verify_seeded(const pb_ctx* cctx, uint64_t seed, uint64_t start, size_t count, int variant, int64_t counts[18]) {
  ARG(cctx && counts);
  ARG(cctx->vk_valid);
  for (int k = 0; k < 18; k++) counts[k] = 0;
  if (count == 0) return PB_OK;
  pb_ctx* ctx = const_cast<pb_ctx*>(cctx);
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(PB_ERR_CUDA, "plonk_b200: cudaSetDevice failed");
  const size_t chunk = 1u << 21;
  const size_t items = count < chunk ? count : chunk;
  int64_t* d_counts = nullptr;
  {
    std::lock_guard<std::mutex> lock(ctx->pipe_mu);
    if (ctx->seed_ws_items < items) {
      if (ctx->d_seed_ws) CU(cudaFree(ctx->d_seed_ws));
      ctx->d_seed_ws = nullptr; ctx->seed_ws_items = 0;
      CU(cudaMalloc(&ctx->d_seed_ws, pb_seeded_workspace_bytes(items) + 256));
      ctx->seed_ws_items = items;
    }
    int rc = pipe_init(ctx, 128);
    if (!rc) rc = seeded_tables(ctx);
    if (rc) return rc;
    d_counts = reinterpret_cast<int64_t*>(ctx->d_seed_ws + ((pb_seeded_workspace_bytes(items) + 15) & ~(size_t)15));
    CU(cudaMemsetAsync(d_counts, 0, 18 * sizeof(int64_t), ctx->s_k));
    for (size_t done = 0; done < count; done += chunk) {
      const size_t m = count - done < chunk ? count - done : chunk;
      rc = pb_plonk_prove_verify_seeded_dev(ctx, seed, start + done, m, variant, ctx->d_seed_ws, d_counts, ctx->s_k);
      if (rc) return rc;
    }
    CU(cudaMemcpyAsync(counts, d_counts, 18 * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->s_k));
    CU(cudaStreamSynchronize(ctx->s_k));
  }
  return PB_OK;
}

int pb_plonk_prove_fs(const pb_ctx* ctx, const uint8_t* witness, const uint8_t* rnd, uint8_t* proofs, uint8_t* status, uint8_t* chal_out, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && witness && rnd && proofs && status);
  if (!chal_out) return pipeline(ctx, witness, rnd, nullptr, nullptr, proofs, status, nullptr, n, PIPE_PROVE);
  // with the challenge read-back: the generic chunked pipeline
  return piped(ctx->device, {HIN(witness, 12), HIN(rnd, 9), HOUT(proofs, 34), HOUT(status, 1), HOUT(chal_out, 6)}, n,
               [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
                 return pb_plonk_prove_fs_dev(ctx, di[0], di[1], dq[0], dq[1], dq[2], m, st); });
}
int pb_plonk_prove_verify_fs(const pb_ctx* ctx, const uint8_t* witness, const uint8_t* rnd, uint8_t* proofs, uint8_t* status,
                             uint8_t* verdict, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && witness && rnd && proofs && status && verdict);
  ARG(ctx->vk_valid);
  return pipeline(ctx, witness, rnd, nullptr, nullptr, proofs, status, verdict, n, PIPE_STRUCT);
}
int pb_plonk_verify_fs(const pb_ctx* ctx, const uint8_t* proofs, uint8_t* verdict, uint8_t* gt, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && proofs && verdict);
  if (gt) return piped(ctx->device, {HIN(proofs, 34), HOUT(verdict, 1), HOUT(gt, 4)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_plonk_verify_fs_dev(ctx, di[0], dq[0], dq[1], m, st); });
  return piped(ctx->device, {HIN(proofs, 34), HOUT(verdict, 1)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_plonk_verify_fs_dev(ctx, di[0], dq[0], nullptr, m, st); });
}
int pb_fs_challenges(const pb_ctx* ctx, const uint8_t* proofs, uint8_t* chal6, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && proofs && chal6);
  return piped(ctx->device, {HIN(proofs, 34), HOUT(chal6, 6)}, n, [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) {
    return pb_fs_challenges_dev(ctx, di[0], dq[0], m, st); });
}
int pb_plonk_verify(const pb_ctx* ctx, const uint8_t* proofs, const uint8_t* chal, const uint8_t* u, uint8_t* verdict, uint8_t* gt, size_t n) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(ctx && proofs && chal && u && verdict);
  if (gt) return piped(ctx->device, {HIN(proofs, 34), HIN(chal, 5), HIN(u, 1), HOUT(verdict, 1), HOUT(gt, 4)}, n,
                       [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) { return pb_plonk_verify_dev(ctx, di[0], di[1], di[2], dq[0], dq[1], m, st); });
  return piped(ctx->device, {HIN(proofs, 34), HIN(chal, 5), HIN(u, 1), HOUT(verdict, 1)}, n,
               [&](uint8_t* const* di, uint8_t* const* dq, size_t m, cudaStream_t st) { return pb_plonk_verify_dev(ctx, di[0], di[1], di[2], dq[0], nullptr, m, st); });
}

// ---- circuit front-end: what eval_expr / gate_list_append produce (constraints.h:227-309) -> the
// 44-byte circuit a context takes.  Host-side authoring, no arithmetic of the path.  The reference never lowers a GATE_LIST
// itself: plonk-test.c:139-213 fills CONSTRAINTS by hand; this is that hand work, mechanised.
int pb_circuit_from_gates(const uint8_t* gates, const size_t* a_idx, const size_t* b_idx, const size_t* c_idx, size_t num_gates,
                          const size_t* equal_pairs, size_t n_equal, uint8_t circuit[PB_CIRCUIT_BYTES]) {
  ARG(circuit && (num_gates == 0 || (gates && a_idx && b_idx && c_idx)) && (n_equal == 0 || equal_pairs));
  if (num_gates > 4)
    return fail(PB_ERR_ARG, "plonk_b200: a circuit has at most 4 gates: the domain H is generated by omega = 4, of order 4 in F17 "
                            "(plonk.h:12), and interpolate_at_h exits unless num_constraints == h_len (plonk.h:164-167)");
  for (size_t i = 0; i < 5 * num_gates; i++) ARG(gates[i] < 17);
  memset(circuit, 0, PB_CIRCUIT_BYTES);
  for (size_t i = 0; i < num_gates; i++)
    for (int s = 0; s < 5; s++) circuit[4 * s + i] = gates[5 * i + s];        // GATE{q_l q_r q_o q_m q_c} -> q_l[4] q_r[4] ...
  // the variable on each of the twelve wire positions, type-major: A1..A4 B1..B4 C1..C4; an unused row gets a variable of its own
  size_t var[12], next_free = 0;
  for (size_t i = 0; i < num_gates; i++) {
    var[i] = a_idx[i]; var[4 + i] = b_idx[i]; var[8 + i] = c_idx[i];
    for (size_t v : {a_idx[i], b_idx[i], c_idx[i]}) if (v + 1 > next_free) next_free = v + 1;
  }
  for (size_t k = 0; k < n_equal; k++) for (int j = 0; j < 2; j++) if (equal_pairs[2 * k + j] + 1 > next_free) next_free = equal_pairs[2 * k + j] + 1;
  for (size_t i = num_gates; i < 4; i++) { var[i] = next_free++; var[4 + i] = next_free++; var[8 + i] = next_free++; }
  // variables asserted equal (an expression's output that must equal another's) share one copy cycle: union-find over the pairs
  std::map<size_t, size_t> parent;
  auto find = [&](size_t v) { while (parent.count(v) && parent[v] != v) v = parent[v]; return v; };
  for (size_t k = 0; k < n_equal; k++) {
    const size_t a = find(equal_pairs[2 * k]), b = find(equal_pairs[2 * k + 1]);
    if (a != b) parent[a > b ? a : b] = a > b ? b : a;
  }
  // copy constraints: every position points at the next position (in type-major order, cyclically) that holds the same variable
  for (int p = 0; p < 12; p++) {
    int q = p;
    for (int step = 1; step <= 12; step++) {
      const int cand = (p + step) % 12;
      if (find(var[cand]) == find(var[p])) { q = cand; break; }
    }
    circuit[20 + 8 * (p / 4) + (p % 4)] = (uint8_t)(q / 4);              // COPY_OF.type: COPYOF_A / _B / _C
    circuit[24 + 8 * (p / 4) + (p % 4)] = (uint8_t)(q % 4 + 1);          // COPY_OF.index, 1-based (plonk.h:144)
  }
  return PB_OK;
}
// witness rows a[4] b[4] c[4] of each item from its variable values: var_values[n][n_vars] (rows beyond num_gates are zero)
int pb_witness_from_values(const size_t* a_idx, const size_t* b_idx, const size_t* c_idx, size_t num_gates, const uint8_t* var_values,
                           size_t n_vars, uint8_t* witness, size_t n) {
  if (n == 0) return PB_OK;
  ARG(num_gates <= 4 && a_idx && b_idx && c_idx && var_values && witness);
  for (size_t i = 0; i < num_gates; i++) ARG(a_idx[i] < n_vars && b_idx[i] < n_vars && c_idx[i] < n_vars);
  for (size_t k = 0; k < n; k++) {
    uint8_t* w = witness + 12 * k;
    memset(w, 0, 12);
    for (size_t i = 0; i < num_gates; i++) {
      w[i] = var_values[k * n_vars + a_idx[i]]; w[4 + i] = var_values[k * n_vars + b_idx[i]]; w[8 + i] = var_values[k * n_vars + c_idx[i]];
    }
  }
  return PB_OK;
}

int pb_tally_dev(const uint8_t* proofs, const uint8_t* status, const uint8_t* verdict, size_t n, int64_t* counts, void* stream) {
  if (n == 0) return PB_OK;   // an empty batch is valid and touches no pointer
  ARG(counts);
  if (n == 0) return PB_OK;
  unsigned grid = blocks_for(n, BLOCK_LIGHT);
  if (grid > 148u * 4u) grid = 148u * 4u;   // measured: 148 blocks 47 us, 296: 34, 592: 26.5, 1184: 28.4, 2368: 31 (2^21 items)
  const int vec = aligned16(proofs) && aligned16(status) && aligned16(verdict);   // null pointers count as aligned
  tally_kernel<<<grid, BLOCK_LIGHT, 0, S(stream)>>>(proofs, status, verdict, n, reinterpret_cast<unsigned long long*>(counts), vec);
  LAUNCH_CHECK("tally_kernel");
  return PB_OK;
}

int pb_peak_probe_dev(int kind, uint32_t iters, uint64_t* ops_out_host, uint32_t* sink_dev, void* stream) {
  ARG(kind >= 0 && kind <= 7 && ops_out_host && sink_dev);
  const unsigned grid = 148u * 8u, block = 256u;
  peak_probe_kernel<<<grid, block, 0, S(stream)>>>(kind, iters, sink_dev);
  LAUNCH_CHECK("peak_probe_kernel");
  *ops_out_host = (uint64_t)grid * block * iters * (uint64_t)PROBE_OPS_PER_ITER;
  return PB_OK;
}

}  // extern "C"
