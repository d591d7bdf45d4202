// Context, error plumbing, raw device memory and the NCCL bridge of libb200lbfgs.so.
//
// b200_ctx replaces cuda_mlp::CublasHandle (src/cuda/cublas_handle.cuh:22-39): where the reference
// carries a cuBLAS handle on the legacy default stream, this carries a non-blocking stream, a pinned
// scalar mailbox and (multi-GPU) an NCCL communicator. No cuBLAS is linked.
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h> // types only; the functions are resolved with dlopen (no link-time NCCL dependency)

#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <set>

namespace b200 {

static thread_local std::string t_error;
std::atomic<long> g_launches{0};

void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_error = buf;
}

static EnvFlags g_env;
static std::once_flag g_env_once;
void env_reload() {
  auto flag = [](const char *n) { return std::getenv(n) != nullptr; };
  auto num = [](const char *n, int dflt) { const char *e = std::getenv(n); return e ? std::atoi(e) : dflt; };
  EnvFlags e;
  e.no_fused_direction = flag("B200_NO_FUSED_DIRECTION");
  e.no_graph = flag("B200_NO_GRAPH");
  e.tc_timing = flag("B200_TC_TIMING");
  e.no_zero_copy = flag("B200_NO_ZERO_COPY");
  e.no_speculation = flag("B200_NO_SPECULATION");
  e.no_p2p = flag("B200_NO_P2P");
  e.nvtx = flag("B200_NVTX");
  e.fwd16 = num("B200_FWD16", -1);
  e.dw16 = num("B200_DW16", -1);
  e.tail = num("B200_TAIL", -1);
  e.tail_fwd = num("B200_TAIL_FWD", -1);
  e.mid16 = num("B200_MID16", -1);
  e.wide16 = num("B200_WIDE16", 1);
  e.wide_chunk = num("B200_WIDE_CHUNK", 0);
  e.pair = num("B200_PAIR", 1);
  e.diag = num("B200_DIAG", 0);
  e.ring = num("B200_RING", 0);
  e.side = num("B200_SIDE", 1);
  e.prep_pub = num("B200_PREP_PUB", 1);
  e.dw_tail = num("B200_DW_TAIL", 0);
  e.pdl = num("B200_PDL", 1) != 0;
  e.tc_mask = num("B200_TC_MASK", 7);
  e.dw_bn = num("B200_DW_BN", 0);
  if (const char *s = std::getenv("B200_P2P_SPIN_LIMIT")) e.p2p_spin_limit = std::atol(s);
  if (const char *s = std::getenv("B200_WIDE16_MIN")) e.wide16_min = std::atol(s);
  if (const char *s = std::getenv("B200_DOTS_BULK")) e.dots_bulk = std::atoi(s);
  if (const char *s = std::getenv("B200_DOTS_BULK_MIN")) e.dots_bulk_min = std::atol(s);
  g_env = e;
}
const EnvFlags &env() {
  std::call_once(g_env_once, env_reload);
  return g_env;
}

static std::mutex g_live_mu;
static std::set<const b200_ctx *> g_live_ctx;
static std::map<unsigned long long, void *> g_live_nets;
bool ctx_is_live(const b200_ctx *ctx) {
  std::lock_guard<std::mutex> lk(g_live_mu);
  return g_live_ctx.count(ctx) != 0;
}
void net_register(unsigned long long uid, void *net) {
  std::lock_guard<std::mutex> lk(g_live_mu);
  if (net) g_live_nets[uid] = net; else g_live_nets.erase(uid);
}
void *net_lookup(unsigned long long uid) {
  std::lock_guard<std::mutex> lk(g_live_mu);
  auto it = g_live_nets.find(uid);
  return it == g_live_nets.end() ? nullptr : it->second;
}

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                                cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  static bool ok = false;
  std::call_once(once, [] {
    // if the host process (torch.distributed) already mapped an NCCL, the SONAME lookup returns that one
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
#define B200_SYM(field, sym)                                              \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym)); \
  if (!api.field) return;
    B200_SYM(GetUniqueId, "ncclGetUniqueId")
    B200_SYM(CommInitRank, "ncclCommInitRank")
    B200_SYM(CommDestroy, "ncclCommDestroy")
    B200_SYM(AllReduce, "ncclAllReduce")
    B200_SYM(ReduceScatter, "ncclReduceScatter")
    B200_SYM(AllGather, "ncclAllGather")
    B200_SYM(Reduce, "ncclReduce")
    B200_SYM(Broadcast, "ncclBroadcast")
    B200_SYM(GroupStart, "ncclGroupStart")
    B200_SYM(GroupEnd, "ncclGroupEnd")
    B200_SYM(GetErrorString, "ncclGetErrorString")
#undef B200_SYM
    ok = true;
  });
  if (!ok) {
    set_error("NCCL could not be loaded (dlopen libnccl.so.2): %s", dlerror() ? dlerror() : "missing symbol");
    return nullptr;
  }
  return &api;
}

#define B200_NCCL(api, call)                                                                   \
  do {                                                                                         \
    ncclResult_t _r = (call);                                                                  \
    if (_r != ncclSuccess) {                                                                   \
      ::b200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, (api)->GetErrorString(_r)); \
      return B200_ERR_COMM;                                                                    \
    }                                                                                          \
  } while (0)

// sum-allreduce of the flat gradient and (optionally) the double loss scalar as ONE grouped NCCL launch
int ctx_allreduce(b200_ctx *ctx, float *grad, size_t n, double *loss_dev) {
  if (ctx->world <= 1) return B200_OK;
  NcclApi *api = nccl_api();
  if (!api) return B200_ERR_COMM;
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  B200_NCCL(api, api->GroupStart());
  if (grad && n) B200_NCCL(api, api->AllReduce(grad, grad, n, ncclFloat, ncclSum, comm, ctx->stream));
  if (loss_dev) B200_NCCL(api, api->AllReduce(loss_dev, loss_dev, 1, ncclDouble, ncclSum, comm, ctx->stream));
  B200_NCCL(api, api->GroupEnd());
  return B200_OK;
}

// Symmetric buffers for the peer-memory all-reduce. Layout per rank: slot 0 | slot 1 | flags[world] (u32) | epoch (u32), each
// slot = slot_floats floats (gradient) + one double (loss partial) padded to 256 B. IPC handles travel through an NCCL
// all-gather, so no extra rendezvous is needed. Any failure (IPC not permitted, no peer access) leaves ready = false on EVERY
// rank (agreed with an all-reduce) and the callers keep using NCCL.
int ctx_p2p_setup(b200_ctx *ctx, size_t n_floats) {
  b200_ctx::P2P &pp = ctx->p2p;
  if (ctx->world <= 1) return B200_OK;
  if (pp.tried && n_floats <= pp.req_floats) return B200_OK;
  NcclApi *api = nccl_api();
  if (!api) return B200_ERR_COMM;
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  const int W = ctx->world;
  if (pp.tried) {
    // a larger network than the buffers were sized for (every rank gets here at the same evaluation): nobody may still be reading
    // the old slots when they are unmapped, so drain and meet the peers first
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    B200_TRY(ctx_allreduce_f64(ctx, ctx->d_scalars + 63, 1));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    for (void *p : pp.opened) cudaIpcCloseMemHandle(p);
    pp.opened.clear();
    if (pp.peers_dev) cudaFree(pp.peers_dev);
    if (pp.local) cudaFree(pp.local);
    pp.peers_dev = nullptr;
    pp.local = nullptr;
    pp.ready = false;
    cudaGetLastError();
  }
  pp.tried = true;
  pp.req_floats = n_floats;
  ++pp.gen;
  int ok = env().no_p2p ? 0 : 1;
  pp.slot_floats = (n_floats + 63) & ~size_t(63);
  pp.slot_bytes = ((pp.slot_floats * 4 + 8) + 255) & ~size_t(255);
  const size_t total = 2 * pp.slot_bytes + ((sizeof(unsigned) * (W + 1) + 255) & ~size_t(255));
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok && cudaMalloc(&pp.local, total) != cudaSuccess) ok = 0;
  if (ok && cudaMemset(pp.local, 0, total) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&mine, pp.local) != cudaSuccess) ok = 0;
  cudaGetLastError();
  // all-gather {ok flag, handle}
  struct Item { int ok; int pad; cudaIpcMemHandle_t h; };
  Item item{ok, 0, mine}, *d_item = nullptr, *d_all = nullptr;
  std::vector<Item> all(W);
  B200_CUDA(cudaMalloc(&d_item, sizeof(Item)));
  B200_CUDA(cudaMalloc(&d_all, sizeof(Item) * W));
  B200_CUDA(cudaMemcpy(d_item, &item, sizeof(Item), cudaMemcpyHostToDevice));
  B200_NCCL(api, api->AllGather(d_item, d_all, sizeof(Item), ncclChar, comm, ctx->stream));
  B200_CUDA(cudaStreamSynchronize(ctx->stream));
  B200_CUDA(cudaMemcpy(all.data(), d_all, sizeof(Item) * W, cudaMemcpyDeviceToHost));
  for (int r = 0; r < W; ++r) ok &= all[r].ok;
  std::vector<char *> peers(W, nullptr);
  if (ok) {
    for (int r = 0; r < W && ok; ++r) {
      if (r == ctx->rank) { peers[r] = pp.local; continue; }
      void *ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
      pp.opened.push_back(ptr);
      peers[r] = (char *)ptr;
    }
  }
  // agree: everybody opened everybody
  item.ok = ok;
  B200_CUDA(cudaMemcpy(d_item, &item, sizeof(Item), cudaMemcpyHostToDevice));
  B200_NCCL(api, api->AllGather(d_item, d_all, sizeof(Item), ncclChar, comm, ctx->stream));
  B200_CUDA(cudaStreamSynchronize(ctx->stream));
  B200_CUDA(cudaMemcpy(all.data(), d_all, sizeof(Item) * W, cudaMemcpyDeviceToHost));
  for (int r = 0; r < W; ++r) ok &= all[r].ok;
  cudaFree(d_item);
  cudaFree(d_all);
  if (ok) {
    B200_CUDA(cudaMalloc(&pp.peers_dev, sizeof(char *) * W));
    B200_CUDA(cudaMemcpy(pp.peers_dev, peers.data(), sizeof(char *) * W, cudaMemcpyHostToDevice));
    pp.ready = true;
  } else {
    for (void *p : pp.opened) cudaIpcCloseMemHandle(p);
    pp.opened.clear();
    cudaGetLastError();
  }
  return B200_OK;
}

// reduce-scatter with exact (possibly uneven) shards: rank i receives sum over ranks of full[i*chunk .. min(n, (i+1)*chunk))
// in shard_out; one grouped launch of W ncclReduce calls
int ctx_reduce_shards(b200_ctx *ctx, const float *full, float *shard_out, size_t n, size_t chunk) {
  if (ctx->world <= 1) return B200_OK;
  NcclApi *api = nccl_api();
  if (!api) return B200_ERR_COMM;
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  B200_NCCL(api, api->GroupStart());
  for (int i = 0; i < ctx->world; ++i) {
    const size_t lo = std::min(n, (size_t)i * chunk), len = std::min(chunk, n - lo);
    if (len) B200_NCCL(api, api->Reduce(full + lo, shard_out, len, ncclFloat, ncclSum, i, comm, ctx->stream));
  }
  B200_NCCL(api, api->GroupEnd());
  return B200_OK;
}

// all-gather with the same shard layout, in place on `full` (rank i is the root of its own shard)
int ctx_allgather_shards(b200_ctx *ctx, float *full, size_t n, size_t chunk) {
  if (ctx->world <= 1) return B200_OK;
  NcclApi *api = nccl_api();
  if (!api) return B200_ERR_COMM;
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  B200_NCCL(api, api->GroupStart());
  for (int i = 0; i < ctx->world; ++i) {
    const size_t lo = std::min(n, (size_t)i * chunk), len = std::min(chunk, n - lo);
    if (len) B200_NCCL(api, api->Broadcast(full + lo, full + lo, len, ncclFloat, i, comm, ctx->stream));
  }
  B200_NCCL(api, api->GroupEnd());
  return B200_OK;
}

int ctx_check_device_error(b200_ctx *ctx) {
  volatile double *flag = ctx->h_scalars + kHostErrSlot;
  if (*flag == 0.0) return B200_OK;
  const double code = *flag;
  *flag = 0.0;
  if (code == 2.0) {
    set_error("prep_w16_kernel: the layer-0 CTAs did not publish the feature scales within the wait limit; results of this evaluation are invalid");
    return B200_ERR_CUDA;
  }
  set_error("rank %d: a peer did not publish its gradient slot within the wait limit of the peer-memory all-reduce "
            "(dead or stalled rank); results of this evaluation are invalid", ctx->rank);
  return B200_ERR_COMM;
}

int ctx_allreduce_f64(b200_ctx *ctx, double *v, size_t n) {
  if (ctx->world <= 1) return B200_OK;
  NcclApi *api = nccl_api();
  if (!api) return B200_ERR_COMM;
  B200_NCCL(api, api->AllReduce(v, v, n, ncclDouble, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
  return B200_OK;
}

} // namespace b200

using namespace b200;

extern "C" {

const char *b200_last_error(void) { return t_error.c_str(); }
int b200_abi_version(void) { return B200_ABI_VERSION; }
long b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int b200_debug_reload_env(void) {
  env(); // (first use reads the environment itself)
  env_reload();
  return B200_OK;
}

int b200_ctx_create(int device, b200_ctx **out) {
  B200_REQUIRE(out, "null out pointer");
  int count = 0;
  B200_CUDA(cudaGetDeviceCount(&count));
  B200_REQUIRE(device >= 0 && device < count, "no such CUDA device (this library has no CPU fallback)");
  B200_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  B200_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libb200lbfgs.so is built for sm_100a only", device, prop.major, prop.minor);
    return B200_ERR_UNSUPPORTED;
  }
  b200_ctx *ctx = new b200_ctx;
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  B200_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  B200_CUDA(cudaMallocHost(&ctx->h_scalars, sizeof(double) * kHostScalars));
  memset(ctx->h_scalars, 0, sizeof(double) * kHostScalars);
  B200_CUDA(cudaMalloc(&ctx->d_scalars, sizeof(double) * 64));
  B200_CUDA(cudaMemset(ctx->d_scalars, 0, sizeof(double) * 64));
  B200_CUDA(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
  B200_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  B200_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  B200_CUDA(cudaEventCreate(&ctx->ev_a));
  B200_CUDA(cudaEventCreate(&ctx->ev_b));
  { std::lock_guard<std::mutex> lk(g_live_mu); g_live_ctx.insert(ctx); }
  *out = ctx;
  return B200_OK;
}

int b200_ctx_destroy(b200_ctx *ctx) {
  if (!ctx || !ctx_is_live(ctx)) return B200_OK;
  { std::lock_guard<std::mutex> lk(g_live_mu); g_live_ctx.erase(ctx); }
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->comm) {
    if (NcclApi *api = nccl_api()) api->CommDestroy((ncclComm_t)ctx->comm);
  }
  for (void *p : ctx->p2p.opened) cudaIpcCloseMemHandle(p);
  if (ctx->p2p.peers_dev) cudaFree(ctx->p2p.peers_dev);
  if (ctx->p2p.local) cudaFree(ctx->p2p.local);
  if (ctx->lbfgs_pool_free)
    for (void *p : ctx->lbfgs_pool) ctx->lbfgs_pool_free(p);
  ctx->lbfgs_pool.clear();
  for (cudaEvent_t e : ctx->prof.pool) cudaEventDestroy(e);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
  cudaEventDestroy(ctx->ev_a);
  cudaEventDestroy(ctx->ev_b);
  cudaFreeHost(ctx->h_scalars);
  cudaFree(ctx->d_scalars);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return B200_OK;
}

int b200_ctx_profile(b200_ctx *ctx, int enable) {
  B200_REQUIRE(ctx, "null ctx");
  B200_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->prof.on = enable != 0;
  ctx->prof.entries.clear();
  ctx->prof.used = 0;
  return B200_OK;
}

int b200_ctx_profile_report(b200_ctx *ctx, char *json_out, size_t capacity) {
  B200_REQUIRE(ctx && json_out && capacity > 2, "bad argument");
  B200_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<double> ms(ctx->prof.names.size(), 0.0);
  std::vector<long> calls(ctx->prof.names.size(), 0);
  for (const auto &e : ctx->prof.entries) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, e.a, e.b) == cudaSuccess) {
      ms[e.id] += t;
      calls[e.id] += 1;
    }
  }
  std::string out = "{";
  for (size_t i = 0; i < ms.size(); ++i) {
    if (!calls[i]) continue;
    char buf[256];
    snprintf(buf, sizeof(buf), "%s\"%s\": [%ld, %.6f]", out.size() > 1 ? ", " : "", ctx->prof.names[i].c_str(), calls[i], ms[i]);
    out += buf;
  }
  out += "}";
  B200_REQUIRE(out.size() + 1 <= capacity, "report buffer too small");
  memcpy(json_out, out.c_str(), out.size() + 1);
  return B200_OK;
}

int b200_ctx_synchronize(b200_ctx *ctx) {
  B200_REQUIRE(ctx, "null ctx");
  B200_CUDA(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}
int b200_ctx_set_stream(b200_ctx *ctx, void *cuda_stream) {
  B200_REQUIRE(ctx, "null ctx");
  B200_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return B200_OK;
}
void *b200_ctx_stream(b200_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int b200_ctx_device(b200_ctx *ctx) { return ctx ? ctx->device : -1; }
int b200_ctx_rank(b200_ctx *ctx) { return ctx ? ctx->rank : 0; }
int b200_ctx_world(b200_ctx *ctx) { return ctx ? ctx->world : 1; }

int b200_comm_unique_id(void *out_128_bytes) {
  B200_REQUIRE(out_128_bytes, "null out pointer");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  NcclApi *api = nccl_api();
  if (!api) return B200_ERR_COMM;
  ncclUniqueId id;
  B200_NCCL(api, api->GetUniqueId(&id));
  memcpy(out_128_bytes, &id, sizeof(id));
  return B200_OK;
}

int b200_ctx_init_comm(b200_ctx *ctx, const void *unique_id_128_bytes, int rank, int world) {
  B200_REQUIRE(ctx && unique_id_128_bytes, "null argument");
  B200_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank/world");
  B200_REQUIRE(!ctx->comm, "communicator already initialised");
  if (world == 1) return B200_OK;
  NcclApi *api = nccl_api();
  if (!api) return B200_ERR_COMM;
  B200_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, unique_id_128_bytes, sizeof(id));
  ncclComm_t comm;
  B200_NCCL(api, api->CommInitRank(&comm, world, id, rank));
  ctx->comm = comm;
  ctx->rank = rank;
  ctx->world = world;
  return B200_OK;
}

int b200_ctx_allreduce_f32(b200_ctx *ctx, float *dev, size_t n) {
  B200_REQUIRE(ctx && dev, "null argument");
  return ctx_allreduce(ctx, dev, n, nullptr);
}

// ---- raw device memory (DeviceBuffer<T>, src/cuda/device_buffer.cuh:7-96) ------------------------
int b200_malloc(void **dev, size_t bytes) {
  B200_REQUIRE(dev, "null out pointer");
  *dev = nullptr;
  if (bytes == 0) return B200_OK;
  B200_CUDA(cudaMalloc(dev, bytes));
  return B200_OK;
}
int b200_free(void *dev) {
  if (dev) B200_CUDA(cudaFree(dev));
  return B200_OK;
}
int b200_memcpy_h2d(void *dev, const void *host, size_t bytes) {
  if (bytes) {
    B200_CUDA(cudaDeviceSynchronize()); // library work runs on a non-blocking stream the legacy stream does not order with
    B200_CUDA(cudaMemcpy(dev, host, bytes, cudaMemcpyHostToDevice));
  }
  return B200_OK;
}
int b200_memcpy_d2h(void *host, const void *dev, size_t bytes) {
  if (bytes) {
    B200_CUDA(cudaDeviceSynchronize());
    B200_CUDA(cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost));
  }
  return B200_OK;
}
int b200_memcpy_d2d(void *dst, const void *src, size_t bytes) {
  if (bytes) {
    B200_CUDA(cudaDeviceSynchronize());
    B200_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToDevice));
    B200_CUDA(cudaDeviceSynchronize());
  }
  return B200_OK;
}
int b200_memset(void *dev, int value, size_t bytes) {
  if (bytes) {
    B200_CUDA(cudaDeviceSynchronize());
    B200_CUDA(cudaMemset(dev, value, bytes));
    B200_CUDA(cudaDeviceSynchronize());
  }
  return B200_OK;
}
int b200_host_alloc_pinned(void **host, size_t bytes) {
  B200_REQUIRE(host, "null out pointer");
  B200_CUDA(cudaMallocHost(host, bytes ? bytes : 1));
  return B200_OK;
}
int b200_host_free_pinned(void *host) {
  if (host) B200_CUDA(cudaFreeHost(host));
  return B200_OK;
}

} // extern "C"
