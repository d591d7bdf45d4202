// Shared infrastructure of libb200lbfgs.so: status/error plumbing, the context object
// (device, stream, NCCL communicator), launch accounting and warp/block reductions.
#pragma once

#include "../../include/b200_lbfgs.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h> // header-only (NVTX 3): ranges cost nothing unless a tool is attached

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

namespace b200 {

// ---- error plumbing -----------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<long> g_launches;

#define B200_CUDA(call)                                                                         \
  do {                                                                                          \
    cudaError_t _e = (call);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::b200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));  \
      return B200_ERR_CUDA;                                                                     \
    }                                                                                           \
  } while (0)

#define B200_TRY(call)            \
  do {                            \
    int _s = (call);              \
    if (_s != B200_OK) return _s; \
  } while (0)

#define B200_REQUIRE(cond, msg)                                        \
  do {                                                                 \
    if (!(cond)) {                                                     \
      ::b200::set_error("%s:%d: %s", __FILE__, __LINE__, msg);        \
      return B200_ERR_INVALID;                                         \
    }                                                                  \
  } while (0)

// every kernel launch in the library goes through this so b200_launch_count() is exact
#define B200_LAUNCH(kernel, grid, block, smem, stream, ...)                                                   \
  do {                                                                                                        \
    B200_CUDA(::b200::launch_ex(kernel, dim3(grid), dim3(block), (size_t)(smem), (stream), 1, __VA_ARGS__)); \
    ::b200::g_launches.fetch_add(1, std::memory_order_relaxed);                                               \
  } while (0)

inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

// ---- debugging switches (environment), read ONCE at library load and again only by b200_debug_reload_env(): nothing on the
// per-iteration path calls getenv -------------------------------------------------------------------
struct EnvFlags {
  bool no_fused_direction = false; // B200_NO_FUSED_DIRECTION: three-kernel direction (dots, solve, apply)
  bool no_graph = false;           // B200_NO_GRAPH: plain launches instead of the captured iteration graph
  bool tc_timing = false;          // B200_TC_TIMING: per-CTA clock64 instrumentation of the tcgen05 kernels
  bool no_zero_copy = false;       // B200_NO_ZERO_COPY: D2H copy nodes instead of kernels writing the pinned mailbox
  bool no_speculation = false;     // B200_NO_SPECULATION: never launch the next iteration's graph ahead of the host
  bool no_p2p = false;             // B200_NO_P2P: NCCL all-reduce instead of the peer-memory one
  bool nvtx = false;               // B200_NVTX: NVTX ranges around direction / evaluation / collective
  int fwd16 = -1, dw16 = -1, tail = -1, tail_fwd = -1; // B200_FWD16 / _DW16 / _TAIL / _TAIL_FWD: -1 unset, else the integer
  int mid16 = -1;                  // B200_MID16: 0 = hidden layers on the generic TF32 kernels
  int wide16 = 1;                  // B200_WIDE16=0: wide hidden layers on the generic TF32 kernels instead of the fp16 pair kernels
  long wide16_min = 1L << 30;      // B200_WIDE16_MIN: smallest GEMM (samples x in x out multiply-adds) that takes the wide kernels
  int wide_chunk = 0;              // B200_WIDE_CHUNK: K blocks of 64 per TMEM accumulation chunk of the wide kernels (0 = default 4)
  int pair = 1;                    // B200_PAIR: bit0 = layer-0 forward as CTA pairs (cta_group::2, weights split between the two SMs)
  bool pdl = true;                 // B200_PDL=0: plain stream order instead of programmatic dependent launches
  int dw_tail = 0;                 // B200_DW_TAIL=1: the one-tile last feature group of the fp16 dW kernel gets fewer, longer slices
  int prep_pub = 1;                // B200_PREP_PUB=0: every CTA of the layer-1 forward job recomputes the 128 feature scales from W_0
  int side = 1;                    // B200_SIDE=0: no side stream (every kernel of an evaluation in one stream)
  int ring = 0;                    // B200_RING: bit0 layer-0 forward, bit1 layer-0 dW: deeper X ring than weight / delta ring
  int diag = 0;                    // B200_DIAG: timing experiments of the fp16 kernels (parts switched off; results are wrong)
  int tc_mask = 7;                 // B200_TC_MASK: bit0 FWD, bit1 DX, bit2 DW on the tensor cores; bit3 / bit4 see gemm_tc.cu
  int dw_bn = 0;                   // B200_DW_BN: 128 / 256, 0 = by precision mode
  int dots_bulk = 3;               // B200_DOTS_BULK: history pass of long vectors. 0 register-staged kernel; bulk-copy ring with 1: 8 consumer warps, 2: 16, 3 (default): 16 + the tile's vectors converted once per CTA
  long dots_bulk_min = 1L << 20;   // B200_DOTS_BULK_MIN: shortest vector that takes the bulk-copy dots kernel
  long p2p_spin_limit = 0;         // B200_P2P_SPIN_LIMIT: polls before p2p_reduce_kernel gives up on a peer (0 = default)
};
const EnvFlags &env();
void env_reload();

// ---- launches -------------------------------------------------------------------------------------
// Programmatic dependent launch (B200_PDL, default on): every kernel of the library starts with pdl_enter() — it lets the
// NEXT kernel of the stream be launched at once (griddepcontrol.launch_dependents) and then waits until the PREVIOUS one has
// completed and its memory is visible (griddepcontrol.wait). Launch latency, CTA rasterisation and whatever a kernel does
// before it touches global memory overlap the tail of its predecessor; a CUDA graph captured from these launches carries
// programmatic edges. Every kernel waits before it returns on every path, so completion stays transitive along the stream.
// Measured (B200): launch-bound paths gain most — S-LBFGS 70 -> 92 epochs/s, GD 4 610 -> 5 850 it/s; the L-BFGS iteration 0.5 %.
// Tried and rejected: running the tcgen05 kernels' set-up (barriers, TMEM allocation) BEFORE the wait. A successor's
// tcgen05.alloc then blocks on an SM whose columns the predecessor's CTA still holds, and S-LBFGS epochs became 3x slower.
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, Args &&...args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  unsigned na = 0;
  if (env().pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster_x > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = (unsigned)cluster_x; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
#endif

// ---- NCCL, loaded lazily with dlopen so the single-GPU path has no NCCL dependency -------------
struct NcclApi;
NcclApi *nccl_api(); // nullptr (and error set) if libnccl cannot be loaded

} // namespace b200

// ---- per-launch profiling (CUDA events on the launching stream; no sync until the report) ---------
namespace b200 {
struct Profiler {
  bool on = false;
  std::vector<std::string> names;
  struct Entry { int id; cudaEvent_t a, b; };
  std::vector<Entry> entries;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  cudaEvent_t get() {
    if (used == pool.size()) {
      if (pool.size() >= (1u << 17)) return nullptr;
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
      pool.push_back(e);
    }
    return pool[used++];
  }
  int id_of(const char *name) {
    for (size_t i = 0; i < names.size(); ++i) if (names[i] == name) return (int)i;
    names.emplace_back(name);
    return (int)names.size() - 1;
  }
};
} // namespace b200

// ---- context ---------------------------------------------------------------------------------------
namespace b200 {
constexpr int kHostScalars = 128;
constexpr int kHostErrSlot = 64; // h_scalars[kHostErrSlot] != 0: a device-side wait gave up (p2p_reduce_kernel: dead or stalled peer)
} // namespace b200
struct b200_ctx {
  b200::Profiler prof;
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  // side stream for work off the critical path of an evaluation (the dW of a middle layer runs beside the dX / dW chain that
  // follows it): forked from and joined to `stream` with the two events, also inside a stream capture
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // multi-GPU (one process per GPU)
  void *comm = nullptr; // ncclComm_t
  int rank = 0, world = 1;
  // small pinned host mailbox for scalar read-backs
  double *h_scalars = nullptr; // kHostScalars doubles, pinned (zero-copy: kernels write results and error flags into it)
  double *d_scalars = nullptr; // 64 doubles, device
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  // solver objects parked by b200_lbfgs_destroy for reuse (work space + instantiated CUDA graphs): a solve call from host
  // buffers otherwise pays ~15 ms of cudaMalloc / cudaFree / graph instantiation per call (reference: 6 + 2m DeviceBuffers per solve)
  // one-shot all-reduce over NVLink peer memory (context.cu: ctx_p2p_setup): every rank owns a symmetric buffer
  // {2 x [slot_floats floats | loss double], flags[world], epoch}; peers[r] is rank r's buffer mapped into this process
  struct P2P {
    bool tried = false, ready = false;
    long gen = 0; // bumped whenever the buffers are (re)made: captured graphs bake their addresses in
    size_t slot_floats = 0, slot_bytes = 0, req_floats = 0;
    char *local = nullptr;        // this rank's buffer
    char **peers_dev = nullptr;   // device array of world pointers (own entry = local)
    std::vector<void *> opened;   // cudaIpcOpenMemHandle results (to close)
  } p2p;
  std::vector<void *> lbfgs_pool;
  void (*lbfgs_pool_free)(void *) = nullptr;
};

namespace b200 {

// brackets the launches issued during its lifetime with two events when profiling is enabled, and with an NVTX range
// (B200_NVTX=1) so that a timeline tool shows direction / evaluation / collective phases by name
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char *name) : on(env().nvtx) { if (on) nvtxRangePushA(name); }
  ~NvtxRange() { if (on) nvtxRangePop(); }
};
struct ProfScope {
  b200_ctx *ctx;
  cudaEvent_t b = nullptr;
  NvtxRange nvtx;
  ProfScope(b200_ctx *c, const char *name) : ctx(c), nvtx(name) {
    if (!c->prof.on) return;
    cudaEvent_t a = c->prof.get();
    b = c->prof.get();
    if (!a || !b) { b = nullptr; return; }
    c->prof.entries.push_back({c->prof.id_of(name), a, b});
    cudaEventRecord(a, c->stream);
  }
  ~ProfScope() {
    if (b) cudaEventRecord(b, ctx->stream);
  }
};

// Handles may be destroyed in any order (a garbage-collected host language does): objects that hold a pointer to another handle
// ask whether it is still alive before touching it.
bool ctx_is_live(const b200_ctx *ctx);
void net_register(unsigned long long uid, void *net); // nullptr: unregister
void *net_lookup(unsigned long long uid);             // nullptr once the network has been destroyed

int ctx_allreduce(b200_ctx *ctx, float *grad, size_t n, double *loss_dev); // grad (float) + 1 double
// collective (every rank calls it with the same n): maps the peers' symmetric buffers; ctx->p2p.ready tells whether the
// peer-memory all-reduce can be used (all ranks agree), otherwise the callers stay on NCCL
int ctx_p2p_setup(b200_ctx *ctx, size_t n_floats);
// after a stream synchronisation: B200_ERR_COMM if a device-side wait on a peer timed out since the last check
int ctx_check_device_error(b200_ctx *ctx);
int ctx_allreduce_f64(b200_ctx *ctx, double *v, size_t n);
int ctx_reduce_shards(b200_ctx *ctx, const float *full, float *shard_out, size_t n, size_t chunk);
int ctx_allgather_shards(b200_ctx *ctx, float *full, size_t n, size_t chunk);

// ---- speculative iterations (solvers.cu) -----------------------------------------------------------
// The L-BFGS loop launches the graph of iteration i+1 BEFORE the host has seen the result of iteration i, assuming the first
// trial point is accepted (the common case). The kernel that finishes an evaluation decides on the device, with the host's own
// arithmetic, whether that assumption holds (gate_next); every kernel of a speculatively launched graph returns at once when it
// does not, so a wrong guess costs a few empty launches and the host falls back to its line search.
struct SpecState {
  int gate_next;    // 1: the iteration whose evaluation just finished was accepted at its first trial point and has not converged
  int pad;
  double loss_prev; // loss at the start of the iteration being evaluated
  double c1, tol;
  const double *alpha0, *gdotp; // first step length and g.p of the direction being evaluated (device header fields)
};
__device__ __forceinline__ bool spec_skip(const SpecState *st, int spec) {
  return spec != 0 && st != nullptr && *reinterpret_cast<const volatile int *>(&st->gate_next) == 0;
}
// run by the one thread that has just produced (loss, ||g||^2) of an evaluation
__device__ __forceinline__ void spec_decide(SpecState *st, double loss, double gnorm2) {
  if (!st) return;
  const double alpha = (double)(float)*st->alpha0;
  const bool ok = loss <= st->loss_prev + st->c1 * alpha * *st->gdotp;
  const bool conv = sqrt(gnorm2) < st->tol;
  st->gate_next = (ok && !conv) ? 1 : 0;
  if (ok) st->loss_prev = loss;
}

// ---- device helpers -------------------------------------------------------------------------------
// Compiler-only barrier for memory operations: a batch of independent loads written before it is ISSUED before anything after it.
// Without it nvcc sinks each load of an unrolled batch next to its first use (fewer live registers), and an in-order SM then
// waits one full L2 round trip per load instead of one per batch.
__device__ __forceinline__ void loads_in_flight() { asm volatile("" ::: "memory"); }
// Loads that stay where they are written (volatile asm statements keep their order): a batch of them is issued back to back.
__device__ __forceinline__ float ldg_pinned(const float *p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// acc + (t0 + ... + t7) in fp64, a fixed tree; volatile like the loads, so that it stays BEHIND a batch of ldg_pinned (the
// compiler otherwise converts and adds each value right after its load, which stalls an in-order SM once per load)
__device__ __forceinline__ double sum8_pinned(double acc, float t0, float t1, float t2, float t3, float t4, float t5, float t6, float t7) {
  double r;
  asm volatile(
      "{\n\t.reg .f64 d<8>;\n\t"
      "cvt.f64.f32 d0, %2;\n\tcvt.f64.f32 d1, %3;\n\tcvt.f64.f32 d2, %4;\n\tcvt.f64.f32 d3, %5;\n\t"
      "cvt.f64.f32 d4, %6;\n\tcvt.f64.f32 d5, %7;\n\tcvt.f64.f32 d6, %8;\n\tcvt.f64.f32 d7, %9;\n\t"
      "add.f64 d0, d0, d1;\n\tadd.f64 d2, d2, d3;\n\tadd.f64 d4, d4, d5;\n\tadd.f64 d6, d6, d7;\n\t"
      "add.f64 d0, d0, d2;\n\tadd.f64 d4, d4, d6;\n\tadd.f64 d0, d0, d4;\n\tadd.f64 %0, %1, d0;\n\t}"
      : "=d"(r)
      : "d"(acc), "f"(t0), "f"(t1), "f"(t2), "f"(t3), "f"(t4), "f"(t5), "f"(t6), "f"(t7));
  return r;
}
__device__ __forceinline__ float4 ldg4_pinned(const float4 *p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// Every use of t[0..15] below this point depends on it, and it depends on all sixteen values: the loads that produce them are
// issued first, whatever arithmetic follows (no instruction is emitted).
#define B200_PIN16_C(t, c)                                                                                                       \
  asm volatile("" : "+f"(t[0].c), "+f"(t[1].c), "+f"(t[2].c), "+f"(t[3].c), "+f"(t[4].c), "+f"(t[5].c), "+f"(t[6].c), "+f"(t[7].c), \
                    "+f"(t[8].c), "+f"(t[9].c), "+f"(t[10].c), "+f"(t[11].c), "+f"(t[12].c), "+f"(t[13].c), "+f"(t[14].c), "+f"(t[15].c))
#define B200_PIN16_F4(t) do { B200_PIN16_C(t, x); B200_PIN16_C(t, y); B200_PIN16_C(t, z); B200_PIN16_C(t, w); } while (0)
__device__ __forceinline__ float4 ldcg4_pinned(const float4 *p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ double ldcg_pinned(const double *p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum, result valid in thread 0. `red` must hold >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double *red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  if (w == 0) {
    v = (lane < nw) ? red[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}

__device__ __forceinline__ float act_apply(int act, float v) {
  switch (act) {
  case B200_ACT_TANH: return tanhf(v);
  case B200_ACT_RELU: return v > 0.0f ? v : 0.0f;
  case B200_ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
  default: return v;
  }
}
// derivative from the POST-activation value, as the reference CUDA backend does
// (src/cuda/kernels.cuh:109-133); equal to the CPU backend's prime(z) for all four activations.
__device__ __forceinline__ float act_deriv_from_output(int act, float a) {
  switch (act) {
  case B200_ACT_TANH: return 1.0f - a * a;
  case B200_ACT_RELU: return a > 0.0f ? 1.0f : 0.0f;
  case B200_ACT_SIGMOID: return a * (1.0f - a);
  default: return 1.0f;
  }
}

} // namespace b200
