// The MLP objective: forward, loss, backward — replaces cuda_mlp::CudaNetwork / CudaDenseLayer
// (src/cuda/network.cuh:16-156, src/cuda/layer.cuh:13-117). Per evaluation the reference issues
// 5 SGEMMs + 11 small kernels/memsets/memcpys + one blocking dot for a 2-layer net; here it is
// L forward GEMMs (bias+activation, and loss/delta on the last one, in the epilogue), L-1 dX GEMMs
// (activation derivative in the epilogue), L split-K dW GEMMs (bias gradient as an extra ones-row)
// and one finalize pass that reduces the split-K partials straight into the caller's gradient buffer.
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "network.cuh"

#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>

namespace b200 {

namespace {

constexpr int kMaxLayers = 16;

// a contiguous run of gradient elements [off, off + size) whose split-K partials are `splits` slices `stride` floats apart
// starting at partials + part_off (a layer's [dW; db], or — when they come from different kernels — its dW and its db apart)
struct FinLayer {
  unsigned long long off, size, stride;
  const float *part; // slice 0
  int splits;
};
struct FinParams {
  FinLayer L[2 * kMaxLayers];
  int nl;
  unsigned long long n;
  const float *partials;
  const float *w;
  float lam;
  float *grad;
  double *fin_part; // [gridDim.x][2] = {sum g^2, sum w^2}
  const float *zero; // a float 0 in global memory: where the loads of slices past the last one point
  // the LAST CTA to finish also reduces the scalars (what eval_scalars_kernel did in a launch of its own)
  unsigned *done_count; // zero on entry, reset by the last CTA
  const double *loss_part;
  int n_loss;
  double inv_batch, lam_d;
  int want_gnorm;
  EvalOut *out;
  // multi-GPU over peer memory: the gradient goes to this rank's symmetric slot (epoch & 1) instead of `grad`, and the last CTA
  // publishes the loss partial there and raises this rank's flag in every peer's buffer (p2p_reduce_kernel consumes them)
  double *host_out;   // optional pinned-host {loss, gnorm2}
  SpecState *spec_st; // decide the speculation gate after the scalars (single GPU) ...
  int spec;           // ... and skip the whole kernel when launched speculatively on a wrong guess
  char *sym_local;
  char *const *peers;
  unsigned long long slot_bytes, slot_floats;
  int rank, world;
};

// grad[j] = sum_s partial_l[s][j - off_l] (+ lam * w[j]); per-CTA partials of ||g||^2, ||w||^2.
// A CTA takes 32 consecutive gradient elements at a time; its 8 warps take the splits round-robin (every
// load is one coalesced 128-byte row segment, up to 8 in flight per thread), accumulate in fp64 and are
// combined in a fixed warp order, so the result is deterministic and rounded once.
__device__ __forceinline__ unsigned *p2p_flags(char *buf, unsigned long long slot_bytes) {
  return reinterpret_cast<unsigned *>(buf + 2 * slot_bytes);
}

__global__ void __launch_bounds__(256) finalize_grad_kernel(const FinParams p) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  if (spec_skip(p.spec_st, p.spec)) return;
  __shared__ double sh[8][32];
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double g2 = 0.0, w2 = 0.0;
  unsigned epoch = 0;
  float *gdst = p.grad;
  if (p.sym_local) { // epoch: completed peer all-reduces on this context (advanced by p2p_reduce_kernel)
    epoch = *reinterpret_cast<volatile unsigned *>(p2p_flags(p.sym_local, p.slot_bytes) + p.world);
    gdst = reinterpret_cast<float *>(p.sym_local + (epoch & 1u) * p.slot_bytes);
  }
  const unsigned long long ngroups = (p.n + 31) / 32;
  // A group = 32 consecutive gradient elements (one coalesced 128-byte segment per slice). Two passes over the groups:
  //  (1) "thin" groups (at most kThinSplits slices, e.g. the 37-way split-K of layer 0): ONE WARP per group with every slice load
  //      in flight at once, no shared memory and no barrier — eight groups per CTA at a time;
  //  (2) "fat" groups (hundreds of slices: one per CTA of the last-layer kernels / per SM of the mid-layer dW): the CTA's 8 warps
  //      take the slices round-robin (up to 8 loads in flight per thread) and are combined in a fixed warp order.
  // Either way an element is summed in fp64 in a fixed order: deterministic, rounded once.
  constexpr int kThinSplits = 40;
  auto layer_of = [&](unsigned long long j) {
    int l = 0;
#pragma unroll 1
    while (l + 1 < p.nl && j >= p.L[l + 1].off) ++l;
    return l;
  };
  auto emit = [&](unsigned long long j, double tot) {
    if (p.lam != 0.0f) {
      const float wv = __ldg(p.w + j);
      tot = fma((double)p.lam, (double)wv, tot);
      w2 += (double)wv * (double)wv;
    }
    const float s = (float)tot;
    gdst[j] = s;
    g2 += (double)s * (double)s;
  };
  // (2) first: the long poles start at once. Uniform per CTA: the class of a group is that of its first element's layer. Only the
  // groups of the layers with more than kThinSplits slices are visited: a scan of ALL groups with a layer look-up each (eight warps
  // per CTA, 626 000 groups at 2·10⁷ elements) was two thirds of this kernel's 0.5 ms at BASELINE configs[4] (ncu source view).
  // The fat groups of all layers are dealt round-robin over the CTAs as ONE list (a CTA rarely gets two of these long poles).
  unsigned long long dealt = 0;
  for (int lf = p.nl - 1; lf >= 0; --lf) { // (from the END: the last layer's elements have the most slices)
    if (p.L[lf].splits <= kThinSplits) continue;
    const unsigned long long g_lo = (p.L[lf].off + 31) / 32; // first group whose FIRST element lies in this layer
    unsigned long long g_hi = (p.L[lf].off + p.L[lf].size + 31) / 32;
    if (g_hi > ngroups) g_hi = ngroups;
    if (g_hi <= g_lo) continue;
    const unsigned long long first = (blockIdx.x + gridDim.x - dealt % gridDim.x) % gridDim.x; // this CTA's first group of the layer
    dealt += g_hi - g_lo;
  for (unsigned long long it = first; g_lo + it < g_hi; it += gridDim.x) {
    const unsigned long long grp = g_hi - 1 - it;
    const unsigned long long j0 = grp * 32 + lane;
    const unsigned long long j = j0 < p.n ? j0 : p.n - 1; // (lanes past the end repeat the last element and are not emitted)
    double acc = 0.0;
    {
      const FinLayer &L = p.L[layer_of(j)];
      const float *src = L.part + (j - L.off);
      for (int sp = warp; sp < L.splits; sp += 128) { // 16 slices of this warp in flight per step
        float t[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) t[u] = ldg_pinned(sp + 8 * u < L.splits ? src + (unsigned long long)(sp + 8 * u) * L.stride : p.zero);
#pragma unroll
        for (int u = 0; u < 16; u += 8) acc = sum8_pinned(acc, t[u], t[u + 1], t[u + 2], t[u + 3], t[u + 4], t[u + 5], t[u + 6], t[u + 7]);
      }
    }
    __syncthreads();
    sh[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && j0 < p.n)
      emit(j, ((sh[0][lane] + sh[1][lane]) + (sh[2][lane] + sh[3][lane])) + ((sh[4][lane] + sh[5][lane]) + (sh[6][lane] + sh[7][lane])));
  }
  }
  // (1)
  auto do_group = [&](unsigned long long grp) { // (grp is warp-uniform)
    if (p.L[layer_of(grp * 32)].splits > kThinSplits) return;
    const unsigned long long j0 = grp * 32 + lane;
    const bool live = j0 < p.n;
    const unsigned long long j = live ? j0 : p.n - 1; // (every lane stays in the loop: the warp-level barrier below needs them all)
    const FinLayer &L = p.L[layer_of(j)]; // (a group that straddles two layers: each lane follows its own)
    const float *src = L.part + (j - L.off);
    double acc = 0.0;
    if (__all_sync(0xffffffffu, L.splits <= 4)) { // wide layers (one slice per layer: gemm "wide16") and tiny batches: four loads, not 40
      float t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = ldg_pinned(u < L.splits ? src + (unsigned long long)u * L.stride : p.zero);
      acc = ((double)t[0] + (double)t[1]) + ((double)t[2] + (double)t[3]); // (the tree of the general form: same bits)
      if (live) emit(j, acc);
      return;
    }
    for (int sp0 = 0; sp0 < L.splits; sp0 += kThinSplits) {
      float t[kThinSplits];
#pragma unroll
      for (int u = 0; u < kThinSplits; ++u) t[u] = ldg_pinned(sp0 + u < L.splits ? src + (unsigned long long)(sp0 + u) * L.stride : p.zero);
      // (pinned: the 40 loads are issued back to back, ONE L2 round trip per batch)
#pragma unroll
      for (int u = 0; u < kThinSplits; u += 8) acc = sum8_pinned(acc, t[u], t[u + 1], t[u + 2], t[u + 3], t[u + 4], t[u + 5], t[u + 6], t[u + 7]);
    }
    if (live) emit(j, acc);
  };
  if (p.n < (1ull << 18)) {
    for (unsigned long long grp = (unsigned long long)blockIdx.x * 8 + warp; grp < ngroups; grp += (unsigned long long)gridDim.x * 8) do_group(grp);
  } else {
    // long gradients (wide layers: 2·10⁷ elements at BASELINE configs[4], one slice per layer): a warp takes FOUR consecutive groups;
    // when they lie in one layer with at most four slices every load of the four is in flight at once (one group per warp and round
    // trip is latency-bound at ~1 TB/s). Per-element sums as in do_group: same bits.
    const unsigned long long nquads = (p.n + 127) / 128;
    for (unsigned long long quad = (unsigned long long)blockIdx.x * 8 + warp; quad < nquads; quad += (unsigned long long)gridDim.x * 8) {
      const unsigned long long first = quad * 128, lastj = first + 127 < p.n ? first + 127 : p.n - 1;
      const int l = layer_of(first);
      if (l != layer_of(lastj) || p.L[l].splits > 4) { // (warp-uniform)
#pragma unroll 1
        for (int g = 0; g < 4; ++g)
          if (quad * 4 + g < ngroups) do_group(quad * 4 + g);
        continue;
      }
      const FinLayer &L = p.L[l];
      float t[4][4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const unsigned long long j0 = first + g * 32 + lane, j = j0 < p.n ? j0 : p.n - 1;
        const float *src = L.part + (j - L.off);
        t[g][0] = ldg_pinned(src);
#pragma unroll
        for (int u = 1; u < 4; ++u) t[g][u] = L.splits > 1 ? ldg_pinned(u < L.splits ? src + (unsigned long long)u * L.stride : p.zero) : 0.0f;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const unsigned long long j0 = first + g * 32 + lane;
        if (j0 < p.n) emit(j0, ((double)t[g][0] + (double)t[g][1]) + ((double)t[g][2] + (double)t[g][3]));
      }
    }
  }
  const double a = block_sum(g2, red);
  const double b = block_sum(w2, red);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    p.fin_part[2 * blockIdx.x + 0] = a;
    p.fin_part[2 * blockIdx.x + 1] = b;
    __threadfence(); // (peer-memory form: the LAST CTA's system-scope fence below publishes what it has observed of the others)
    last = atomicAdd(p.done_count, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  // loss = 0.5 * inv_batch * sum(loss_part) + 0.5 * lam * sum(w^2 parts); gnorm2 = sum(g^2 parts): fixed order (deterministic)
  __threadfence();
  double l = 0.0, sg = 0.0, sw = 0.0;
  { // every load of this thread in flight before the first add (fixed order per thread)
    double tl[4], tg[4], tw[4];
    for (int i0 = threadIdx.x; i0 < p.n_loss; i0 += 4 * blockDim.x) {
#pragma unroll
      for (int u = 0; u < 4; ++u) tl[u] = (i0 + u * (int)blockDim.x < p.n_loss) ? __ldcg(p.loss_part + i0 + u * blockDim.x) : 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) l += tl[u];
    }
    for (int i0 = threadIdx.x; i0 < (int)gridDim.x; i0 += 4 * blockDim.x) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * (int)blockDim.x;
        const double2 v = (i < (int)gridDim.x) ? __ldcg(reinterpret_cast<const double2 *>(p.fin_part) + i) : make_double2(0.0, 0.0);
        tg[u] = v.x; tw[u] = v.y;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { sg += tg[u]; sw += tw[u]; }
    }
  }
  l = block_sum(l, red);
  sg = block_sum(sg, red);
  sw = block_sum(sw, red);
  if (threadIdx.x == 0) {
    const double loss = 0.5 * p.inv_batch * l + 0.5 * p.lam_d * sw;
    p.out->loss = loss;
    if (p.want_gnorm) p.out->gnorm2 = sg;
    *p.done_count = 0u;
    if (!p.sym_local && p.want_gnorm) {
      spec_decide(p.spec_st, loss, sg);
      if (p.host_out) { p.host_out[0] = loss; p.host_out[1] = sg; __threadfence_system(); }
    }
    if (p.sym_local) { // publish: loss partial into the slot, then this rank's flag (= epoch + 1) into every peer's buffer
      *reinterpret_cast<double *>(p.sym_local + (epoch & 1u) * p.slot_bytes + p.slot_floats * 4) = loss;
      __threadfence_system();
      for (int r = 0; r < p.world; ++r)
        if (r != p.rank) *reinterpret_cast<volatile unsigned *>(p2p_flags(p.peers[r], p.slot_bytes) + p.rank) = epoch + 1u;
    }
  }
}

// One-shot all-reduce over NVLink peer memory, fused with ||g||^2: every rank waits until all peers have published their slot of
// this epoch, then sums the W slots in rank order (identical order on every rank => bit-identical replicas of g) straight out of
// the peers' memory into grad_out. Replaces ncclAllReduce(grad) + ncclAllReduce(loss) + two norm kernels; at n ~ 1e5 the NCCL
// ring / tree costs ~35 us of latency, this ~10. Two slots alternate by epoch: a rank can only overwrite a slot after every
// peer has signalled the NEXT epoch, i.e. after it finished reading this one.
struct P2PReduceParams {
  char *const *peers;
  char *local;
  unsigned long long slot_bytes, slot_floats, n;
  int rank, world;
  float *grad_out;
  double *fin_part;
  unsigned *done_count;
  EvalOut *out;
  SpecState *spec_st;
  int spec;
  double *host_out;
  double *host_err;          // pinned-host flag: set when a peer's slot did not arrive within spin_limit polls
  unsigned long long spin_limit;
};
__global__ void __launch_bounds__(256) p2p_reduce_kernel(const P2PReduceParams p) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  if (spec_skip(p.spec_st, p.spec)) return;
  __shared__ double red[32];
  __shared__ bool last;
  unsigned *flags = p2p_flags(p.local, p.slot_bytes);
  const unsigned epoch = *reinterpret_cast<volatile unsigned *>(flags + p.world);
  if (threadIdx.x < p.world && (int)threadIdx.x != p.rank) {
    // bounded wait: a dead or stalled peer must not hang every other rank inside a kernel. On a timeout the flag goes up in the
    // host mailbox (the solvers turn it into B200_ERR_COMM at their next synchronisation) and the kernel runs to completion on
    // whatever the slot holds.
    const volatile unsigned *f = flags + threadIdx.x;
    unsigned ns = 32;
    unsigned long long polls = 0;
    while (*f < epoch + 1u) {
      __nanosleep(ns);
      if (ns < 256) ns <<= 1; // (the wake-up granularity is part of the all-reduce latency)
      if (++polls > p.spin_limit) {
        if (p.host_err) { *p.host_err = 1.0; __threadfence_system(); }
        break;
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  const unsigned long long off = (epoch & 1u) * p.slot_bytes;
  const char *peer[8]; // (the mapped peer buffers: loaded once, not once per element)
#pragma unroll
  for (int u = 0; u < 8; ++u) peer[u] = p.peers[u < p.world ? u : 0];
  const unsigned long long nv = (p.n + 3) / 4; // slots are padded to 64 floats: whole float4s are always readable
  double g2 = 0.0;
  for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (unsigned long long)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.world <= 8) { // the peer loads of a step go out together: ONE NVLink round trip, not one per rank (pinned: the compiler
      float4 x[8];      // otherwise pairs every load with its add)
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (u < p.world) x[u] = ldcg4_pinned(reinterpret_cast<const float4 *>(peer[u] + off) + v);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (u < p.world) { acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w; } // rank order
    } else {
      for (int r = 0; r < p.world; ++r) {
        const float4 x = __ldcg(reinterpret_cast<const float4 *>(p.peers[r] + off) + v);
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
      }
    }
    const unsigned long long j = 4 * v;
    if (j + 3 < p.n) {
      *reinterpret_cast<float4 *>(p.grad_out + j) = acc;
      g2 += (double)acc.x * acc.x + (double)acc.y * acc.y + (double)acc.z * acc.z + (double)acc.w * acc.w;
    } else {
      const float a[4] = {acc.x, acc.y, acc.z, acc.w};
      for (int e = 0; e < 4; ++e)
        if (j + e < p.n) { p.grad_out[j + e] = a[e]; g2 += (double)a[e] * a[e]; }
    }
  }
  g2 = block_sum(g2, red);
  if (threadIdx.x == 0) {
    p.fin_part[2 * blockIdx.x] = g2;
    __threadfence();
    last = atomicAdd(p.done_count, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double s = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += __ldcg(p.fin_part + 2 * i);
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    double loss = 0.0;
    for (int r = 0; r < p.world; ++r) loss += __ldcg(reinterpret_cast<const double *>(p.peers[r] + off + p.slot_floats * 4));
    p.out->loss = loss;
    p.out->gnorm2 = s;
    *p.done_count = 0u;
    spec_decide(p.spec_st, loss, s);
    if (p.host_out) { p.host_out[0] = loss; p.host_out[1] = s; __threadfence_system(); }
    flags[p.world] = epoch + 1u; // this context's epoch: the slot parity of the next evaluation
  }
}

// loss = 0.5 * inv_batch * sum(loss_part) + 0.5 * lam * sum(w^2 parts); gnorm2 = sum(g^2 parts)
__global__ void __launch_bounds__(1024) eval_scalars_kernel(const double *loss_part, int n_loss, const double *fin_part,
                                                           int n_fin, double inv_batch, double lam, int want_gnorm,
                                                           EvalOut *out) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  __shared__ double red[32];
  double l = 0.0, g2 = 0.0, w2 = 0.0;
  for (int i = threadIdx.x; i < n_loss; i += blockDim.x) l += loss_part[i];
  for (int i = threadIdx.x; i < n_fin; i += blockDim.x) {
    g2 += fin_part[2 * i];
    w2 += fin_part[2 * i + 1];
  }
  l = block_sum(l, red);
  g2 = block_sum(g2, red);
  w2 = block_sum(w2, red);
  if (threadIdx.x == 0) {
    out->loss = 0.5 * inv_batch * l + 0.5 * lam * w2;
    if (want_gnorm) out->gnorm2 = g2;
  }
}

__global__ void __launch_bounds__(256) sumsq_part_kernel(const float *x, unsigned long long n, double *part) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  __shared__ double red[32];
  double s = 0.0;
  for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
       j += (unsigned long long)gridDim.x * blockDim.x) {
    const float v = x[j];
    s += (double)v * (double)v;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) part[2 * blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) gnorm_from_parts_kernel(const double *part, int nparts, EvalOut *out) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += part[2 * i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out->gnorm2 = s;
}

// UnifiedLauncher<CudaBackend>::evaluate (src/unified_launcher.hpp:154-199) on the device:
// per sample arg-max of prediction vs target (first maximum wins, strict >), squared error sum.
__global__ void __launch_bounds__(256) evaluate_kernel(const float *out, const float *tgt, long batch, int od,
                                                       double *part) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  __shared__ double red[32];
  double se = 0.0, correct = 0.0;
  for (long b = (long)blockIdx.x * blockDim.x + threadIdx.x; b < batch; b += (long)gridDim.x * blockDim.x) {
    int pi = 0, ti = 0;
    double pm = -1e20, tm = -1e20;
    for (int r = 0; r < od; ++r) {
      const double v = out[b * od + r], tv = tgt[b * od + r];
      se += (v - tv) * (v - tv);
      if (v > pm) { pm = v; pi = r; }
      if (tv > tm) { tm = tv; ti = r; }
    }
    if (pi == ti) correct += 1.0;
  }
  se = block_sum(se, red);
  correct = block_sum(correct, red);
  if (threadIdx.x == 0) {
    part[2 * blockIdx.x] = se;
    part[2 * blockIdx.x + 1] = correct;
  }
}

// [dW; db] split-K partials of a skinny layer (out <= 16). Each CTA takes a slice of the batch; its 8 warps
// take samples round-robin, a lane owns the input features lane, lane+32, ... (feature `in` is the bias row,
// reading as 1), so every A_prev row is one coalesced 4*in-byte read and the 16 delta values are smem broadcasts.
// The per-warp accumulators are combined in a fixed warp order (deterministic). Replaces a 128x16-tile GEMM
// that would idle 15/16 of its lanes (and the reference's serial sum_rows_kernel, src/cuda/kernels.cuh:144-153).
constexpr int kSkinnyTile = 64;
template <int FPL>
__global__ void __launch_bounds__(256, 2) skinny_dw_kernel(const float *__restrict__ A, const float *__restrict__ D, int ldd, int in,
                                                        int out, long batch, int chunk, float *__restrict__ partial,
                                                        unsigned long long pstride, const SpecState *spec_st, int spec) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  if (spec_skip(spec_st, spec)) return;
  __shared__ __align__(16) float sd[kSkinnyTile][16];
  __shared__ float red[32 * FPL * 16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long b0 = (long)blockIdx.x * chunk, b1 = min(batch, b0 + (long)chunk);
  float acc[FPL][16];
#pragma unroll
  for (int c = 0; c < FPL; ++c)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[c][j] = 0.0f;
  for (long bt = b0; bt < b1; bt += kSkinnyTile) {
    const int nb = (int)min((long)kSkinnyTile, b1 - bt);
    __syncthreads();
    for (int e = threadIdx.x; e < kSkinnyTile * 16; e += blockDim.x) {
      const int r = e >> 4, j = e & 15;
      sd[r][j] = (r < nb && j < out) ? __ldg(D + (bt + r) * ldd + j) : 0.0f;
    }
    __syncthreads();
#pragma unroll 2
    for (int r = warp; r < nb; r += 8) {
      float a[FPL];
#pragma unroll
      for (int c = 0; c < FPL; ++c) {
        const int i = lane + 32 * c;
        a[c] = (i < in) ? __ldg(A + (bt + r) * in + i) : (i == in ? 1.0f : 0.0f);
      }
      const float4 d0 = *reinterpret_cast<const float4 *>(&sd[r][0]), d1 = *reinterpret_cast<const float4 *>(&sd[r][4]);
      const float4 d2 = *reinterpret_cast<const float4 *>(&sd[r][8]), d3 = *reinterpret_cast<const float4 *>(&sd[r][12]);
      const float d[16] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x, d2.y, d2.z, d2.w, d3.x, d3.y, d3.z, d3.w};
#pragma unroll
      for (int c = 0; c < FPL; ++c)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[c][j] = fmaf(a[c], d[j], acc[c][j]);
    }
  }
  for (int w = 0; w < 8; ++w) { // fixed-order cross-warp combine
    __syncthreads();
    if (warp == w) {
#pragma unroll
      for (int c = 0; c < FPL; ++c)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float *slot = &red[((lane + 32 * c) * 16) + j];
          *slot = (w == 0) ? acc[c][j] : *slot + acc[c][j];
        }
    }
  }
  __syncthreads();
  float *dst = partial + (unsigned long long)blockIdx.x * pstride;
  for (int e = threadIdx.x; e < (in + 1) * out; e += blockDim.x) {
    const int i = e / out, j = e - i * out;
    dst[e] = red[i * 16 + j];
  }
}

// x -> u = round(255 x) as uint8 AND as fp16 (exact: u <= 255 has 8 significant bits); *flag &= (every x is exactly
// float(u)/255.0f with 0 <= u <= 255). The fp16 copy is feature-block-major, [nblocks16][rows][64]: features [in | 1 | zero padding];
// column `in` is the ones feature whose "weight gradient" is the bias gradient (gemm_dw16.cu), and every operand tile of the layer-0
// GEMMs is one contiguous chunk that TMA feeds to the tensor cores as it is.
__global__ void __launch_bounds__(256) quantize_u8_kernel(const float *__restrict__ x, unsigned long long n, int in, int nblocks16,
                                                          uint8_t *__restrict__ q, __half *__restrict__ q16, int *flag) {
  pdl_enter(); // programmatic dependent launch: this grid may start while its predecessor drains (common.cuh)
  int ok = 1;
  const unsigned long long nv = n / 4;
  for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < nv;
       v += (unsigned long long)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(x) + v);
    const float f[4] = {a.x, a.y, a.z, a.w};
    unsigned packed = 0;
    __half hv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int u = __float2int_rn(f[e] * 255.0f);
      ok &= (u >= 0 && u <= 255 && __fdiv_rn((float)u, 255.0f) == f[e]) ? 1 : 0;
      packed |= (unsigned)(u & 255) << (8 * e);
      hv[e] = __int2half_rn(u & 255);
    }
    reinterpret_cast<unsigned *>(q)[v] = packed;
    const unsigned long long e0 = v * 4, row = e0 / (unsigned)in, col = e0 - row * (unsigned)in; // in % 4 == 0: one row per vector
    const unsigned long long rows_all = n / (unsigned)in;
    *reinterpret_cast<uint2 *>(q16 + ((col >> 6) * rows_all + row) * 64 + (col & 63)) = *reinterpret_cast<const uint2 *>(hv);
  }
  const unsigned long long rows = n / (unsigned)in;
  for (unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows;
       r += (unsigned long long)gridDim.x * blockDim.x) {
    for (int c = in; c < 64 * nblocks16; ++c) q16[((unsigned long long)(c >> 6) * rows + r) * 64 + (c & 63)] = __float2half_rn(c == in ? 1.0f : 0.0f);
  }
  if (!__all_sync(0xffffffffu, ok)) {
    if ((threadIdx.x & 31) == 0) atomicAnd(flag, 0);
  }
}

int free_batch_buffers(b200_net *net) {
  for (float *p : net->act) if (p) cudaFree(p);
  for (float *p : net->delta) if (p) cudaFree(p);
  net->act.assign(net->nlayers(), nullptr);
  net->delta.assign(net->nlayers(), nullptr);
  if (net->loss_part) { cudaFree(net->loss_part); net->loss_part = nullptr; }
  net->cap = 0;
  return B200_OK;
}

} // namespace

// Per-batch buffers. Unlike the reference (re-cudaMalloc on every batch-size change,
// src/cuda/network.cuh:133-147) capacity only grows, so SGD mini-batches alternating with
// full-batch recorder evaluations do not churn the allocator.
int net_ensure(b200_net *net, long batch) {
  B200_REQUIRE(batch > 0, "batch must be positive");
  const int L = net->nlayers();
  if (batch > net->cap) {
    free_batch_buffers(net);
    for (int l = 0; l < L; ++l) {
      B200_CUDA(cudaMalloc(&net->act[l], sizeof(float) * (size_t)net->dims[l + 1] * batch));
      B200_CUDA(cudaMalloc(&net->delta[l], sizeof(float) * (size_t)net->ldd[l] * batch));
      B200_CUDA(cudaMemsetAsync(net->delta[l], 0, sizeof(float) * (size_t)net->ldd[l] * batch, net->ctx->stream)); // padding columns stay 0
    }
    net->loss_part_cap = std::max(4 * ceil_div(batch, kBM) * ceil_div(net->dims[L], 16), 2 * net->ctx->num_sms);
    B200_CUDA(cudaMalloc(&net->loss_part, sizeof(double) * net->loss_part_cap));
    ++net->config_gen;
    net->cap = batch;
  }
  if (batch != net->partials_batch) {
    // split-K plan: enough CTAs to fill the machine ~2x, each split a multiple of 16 samples
    const int target = 2 * net->ctx->num_sms;
    size_t total = 0;
    for (int l = 0; l < L; ++l) {
      const int M = net->dims[l] + 1, N = net->dims[l + 1];
      const int tn = (N > 64) ? 8 : (N > 32 ? 4 : (N > 16 ? 2 : 1));
      const int tiles = ceil_div(M, kBM) * ceil_div(N, 16 * tn);
      int s = std::max(1, std::min(ceil_div(target, tiles), ceil_div(batch, 64)));
      int kc = ceil_div(ceil_div(batch, s), kBK) * kBK;
      s = ceil_div(batch, kc);
      net->splits[l] = s;
      net->k_chunk[l] = kc;
      net->part_off[l] = total;
      int s_tc = 1;
      tc_dw_plan(net, l, batch, &s_tc); // room for whichever path runs
      net->skinny_splits[l] = 2 * net->ctx->num_sms;
      int s_16 = 1;
      if (l == 0) dw16_plan(net, batch, &s_16);
      if (l == 1) mid16_dw_plan(net, batch, &s_16);
      // (the per-CTA partials of the skinny / one-pass last-layer kernels exist only for narrow layers: reserving them for a
      // 4096-wide layer would be 2 * SMs copies of a 67 MB matrix)
      const int s_skinny = (N <= 16) ? net->skinny_splits[l] : 1;
      total += (size_t)std::max(std::max(std::max(s, s_tc), s_16), s_skinny) * M * N;
    }
    if (total > net->partials_cap) {
      if (net->partials) cudaFree(net->partials);
      B200_CUDA(cudaMalloc(&net->partials, sizeof(float) * total));
      ++net->config_gen;
      net->partials_cap = total;
    }
    net->partials_batch = batch;
  }
  return B200_OK;
}

void net_xq_clear(b200_net *net) {
  if (net->xq.valid) ++net->config_gen;
  net->xq.valid = false;
  net->xg.src = nullptr;
  net->xq.user = false;
  net->xq.src = nullptr;
  net->xq.rows = 0;
  net->w16x.x_src = nullptr;
  net->w16x.x_rows = 0;
  net->w16x.x_done = false;
}

// x is not (or cannot be checked to be) 8-bit pixels: remember that the caller holds it constant (b200_net::Wide16::x_src)
static void net_hold_float_input(b200_net *net, const float *x, long batch) {
  net->w16x.x_src = batch > 0 ? x : nullptr;
  net->w16x.x_rows = batch;
  net->w16x.x_done = false; // (a refresh re-splits: the buffer may hold new data)
}

int net_quantize_input(b200_net *net, const float *x, long batch, bool refresh) {
  const int in = net->dims[0];
  const bool cached = net->xq.valid && net->xq.src == x && net->xq.rows == batch;
  if (cached && !refresh) return B200_OK;
  // refresh: same buffers (no captured graph is invalidated), contents re-derived from what x holds NOW — a caller may have
  // uploaded new data to the same device buffer since the copy was made
  if (!cached) net_xq_clear(net);
  // TMA needs 16-byte row strides on the uint8 copy; float4 reads need an aligned source
  if (in % 16 != 0 || (reinterpret_cast<uintptr_t>(x) & 15u) != 0 || batch <= 0) {
    net_hold_float_input(net, x, batch);
    return B200_OK;
  }
  const size_t bytes = (size_t)batch * in;
  cudaStream_t st = net->ctx->stream;
  const int nb16 = (in + 1 + 63) / 64;
  if (bytes > net->xq.cap) {
    if (net->xq.data) cudaFree(net->xq.data);
    if (net->xq.data16) cudaFree(net->xq.data16);
    net->xq.data = nullptr;
    net->xq.data16 = nullptr;
    net->xq.cap = 0;
    B200_CUDA(cudaMalloc(&net->xq.data, bytes));
    B200_CUDA(cudaMalloc(&net->xq.data16, sizeof(__half) * (size_t)batch * 64 * nb16));
    net->xq.nblocks16 = nb16;
    ++net->config_gen;
    net->xq.cap = bytes;
  }
  if (!net->xq.flag) B200_CUDA(cudaMalloc(&net->xq.flag, sizeof(int)));
  const int one = 1;
  B200_CUDA(cudaMemcpyAsync(net->xq.flag, &one, sizeof(int), cudaMemcpyHostToDevice, st));
  B200_LAUNCH(quantize_u8_kernel, 8 * net->ctx->num_sms, 256, 0, st, x, (unsigned long long)bytes, in, nb16, net->xq.data,
              (__half *)net->xq.data16, net->xq.flag);
  int ok = 0;
  B200_CUDA(cudaMemcpyAsync(&ok, net->xq.flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  if (ok && !cached) {
    ++net->config_gen;
    net->xq.valid = true;
    net->xq.src = x;
    net->xq.rows = batch;
  } else if (!ok && cached) {
    net_xq_clear(net); // the buffer no longer holds 8-bit pixel data
  }
  if (!ok) net_hold_float_input(net, x, batch);
  return B200_OK;
}

void *net_gather16_buffer(b200_net *net, long rows) {
  if (!net->xq.valid || !net->xq.data16) return nullptr;
  if (rows > net->xg.cap) {
    if (net->xg.data16) cudaFree(net->xg.data16);
    net->xg.data16 = nullptr;
    net->xg.cap = 0;
    if (cudaMalloc(&net->xg.data16, sizeof(__half) * (size_t)rows * 64 * net->xq.nblocks16) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    net->xg.cap = rows;
    ++net->config_gen;
  }
  net->xg.src = nullptr; // (contents about to change)
  return net->xg.data16;
}
void net_gather16_register(b200_net *net, const float *src, long rows) {
  net->xg.src = src;
  net->xg.rows = rows;
}
const void *net_x16_source(const b200_net *net, long *rows_total, int *nblocks) {
  if (!net->xq.valid || !net->xq.data16) return nullptr;
  *rows_total = net->xq.rows;
  *nblocks = net->xq.nblocks16;
  return net->xq.data16;
}

bool net_x16_view(b200_net *net, const float *x, long batch, X16View *v) {
  if (net->xg.src == x && net->xg.rows == batch && net->xg.data16 && net->xq.valid) { // a gathered mini-batch
    v->base = net->xg.data16;
    v->rows_total = batch;
    v->row0 = 0;
    v->nblocks = net->xq.nblocks16;
    return true;
  }
  if (!net->xq.valid || !net->xq.data16 || x < net->xq.src) return false;
  const size_t off = (size_t)(x - net->xq.src);
  const int in = net->dims[0];
  if (off % in != 0) return false;
  const long row0 = (long)(off / in);
  if (row0 + batch > net->xq.rows) return false;
  v->base = net->xq.data16;
  v->rows_total = net->xq.rows;
  v->row0 = row0;
  v->nblocks = net->xq.nblocks16;
  return true;
}

// Every kernel of an evaluation honours the speculation gate (common.cuh), so any network can be evaluated speculatively; what
// remains excluded is what enqueues something other than the library's own kernels behind the gate: the NCCL fall-back of the
// gradient all-reduce (the solver checks that itself) and the debugging variants.
bool net_spec_capable(b200_net *net, const float *x, long batch) {
  (void)x; (void)batch;
  return env().fwd16 < 0 && !env().tc_timing;
}

const uint8_t *net_xq_lookup(b200_net *net, const float *x, long batch) {
  if (!net->xq.valid || x < net->xq.src) return nullptr;
  const size_t off = (size_t)(x - net->xq.src);
  const int in = net->dims[0];
  if (off % in != 0) return nullptr;
  const long row0 = (long)(off / in);
  if (row0 + batch > net->xq.rows) return nullptr;
  return net->xq.data + (size_t)row0 * in;
}

static int launch_fwd_layer(b200_net *net, int l, const float *params, const float *in, long batch, bool last,
                            const float *t, float inv_batch) {
  const int K = net->dims[l], N = net->dims[l + 1];
  const float *W = params + net->offs[l];
  GemmParams p{};
  p.A = in; p.lda = K;
  p.B = W; p.ldb = N;
  p.M = (int)batch; p.N = N; p.K = K;
  p.vecA = aligned16(in) && (K % 4 == 0);
  p.vecB = aligned16(W) && (N % 4 == 0);
  p.a_ones_row = -1;
  p.k_chunk = K;
  p.spec_st = net->spec_st; p.spec = net->spec_flag;
  p.bias = W + (size_t)K * N;
  p.act = net->acts[l];
  p.out = net->act[l]; p.ldo = N;
  if (last) {
    p.aux = t; p.ld_aux = N;
    p.delta = net->delta[l];
    p.ldd = net->ldd[l];
    p.inv_batch = inv_batch;
    p.loss_part = net->loss_part;
    const int tn = (N > 64) ? 8 : (N > 32 ? 4 : (N > 16 ? 2 : 1));
    net->loss_part_n = ceil_div(batch, kBM) * ceil_div(N, 16 * tn);
    return launch_gemm_simt<true, false, EPI_FWD_LAST>(p, 1, net->ctx->stream);
  }
  return launch_gemm_simt<true, false, EPI_FWD>(p, 1, net->ctx->stream);
}

int net_forward(b200_net *net, const float *params, const float *x, long batch) {
  B200_TRY(net_ensure(net, batch));
  net->w16_params = nullptr;
  net->chain_ready = false;
  X16View xv0;
  const bool mid = mid16_applicable(net) && net_x16_view(net, x, batch, &xv0);
  net->m16.on = mid;
  net->m16.act0_stale = false;
  net->split_src = nullptr;
  wide16_begin(net);
  if (mid) B200_TRY(mid16_ensure(net, batch));
  else B200_TRY(tc_split_params(net, params));
  if (net->prec != B200_PREC_FP32 && net_x16_view(net, x, batch, &xv0)) B200_TRY(fwd16_prepare(net, params));
  const float *cur = x;
  for (int l = 0; l < net->nlayers(); ++l) {
    bool done = false;
    char nm[16];
    snprintf(nm, sizeof(nm), "fwdonly%d", l);
    ProfScope ps(net->ctx, nm);
    if (mid && l == 1) {
      B200_TRY(mid16_forward_layer1(net, params, batch));
      cur = net->act[1];
      continue;
    }
    if (net->prec != B200_PREC_FP32 && l + 1 < net->nlayers())
      B200_TRY(tc_forward_layer(net, l, params, cur, batch, nullptr, &done, nullptr));
    if (!done) B200_TRY(launch_fwd_layer(net, l, params, cur, batch, false, nullptr, 0.f));
    cur = net->act[l];
  }
  net->last_batch = batch;
  return B200_OK;
}

int net_eval(b200_net *net, const float *params, const float *x, const float *t, long batch, long batch_global,
             float *grad_out, EvalOut *out) {
  B200_REQUIRE(net && params && x && t && grad_out && out, "null argument");
  NvtxRange nvtx_eval("evaluation");
  B200_TRY(net_ensure(net, batch));
  b200_ctx *ctx = net->ctx;
  cudaStream_t st = ctx->stream;
  const int L = net->nlayers();
  if (batch_global <= 0) batch_global = net->batch_global > 0 ? net->batch_global : batch * ctx->world;
  const float inv_batch = 1.0f / (float)batch_global;

  net->w16_params = nullptr;
  net->chain_ready = false;
  // hidden layer 1 on the fp16 kernels (b200_net::Mid16): needs the fp16 copy of an 8-bit-pixel input for layer 0
  X16View xv0;
  const bool mid = mid16_applicable(net) && net_x16_view(net, x, batch, &xv0);
  net->m16.on = mid;
  net->m16.act0_stale = false;
  net->split_src = nullptr;
  wide16_begin(net);
  if (mid) B200_TRY(mid16_ensure(net, batch));
  else B200_TRY(tc_split_params(net, params)); // (the TF32 hi / lo split of the parameters feeds the generic kernels only)
  const bool have16 = net->prec != B200_PREC_FP32 && net_x16_view(net, x, batch, &xv0); // fp16 copy of this batch (input slice or gathered)
  if (have16) B200_TRY(fwd16_prepare(net, params));
  // forward sweep
  const float *cur = x;
  bool fused_last = false; // last layer, loss, delta_L and delta_{L-1} produced by the penultimate layer's epilogue
  bool tail_done = false;  // ... or by the one-pass last-layer kernel, which also produces the [dW_L; db_L] partials
  const bool use_tail = tail_applicable(net);
  const bool use_dw16 = use_tail && have16 && dw16_applicable(net) && (L == 2 || net->chain_ready); // layer-0 dW on the fp16 tensor cores: the tail writes delta_0 as scaled fp16 {hi | lo}
  for (int l = 0; l < L; ++l) {
    const bool last = (l == L - 1);
    if (last && fused_last) break;
    if (mid && l == 1) {
      ProfScope ps(ctx, "fwd1");
      B200_TRY(mid16_forward_layer1(net, params, batch));
      cur = net->act[1];
      continue;
    }
    if (last && use_tail) {
      // mid: delta_1 leaves only as its fp16 pair (feeds both the dX and the dW kernel of layer 1)
      B200_TRY(tail_layer(net, params, t, batch, inv_batch, /*want32=*/!mid && (!use_dw16 || L > 2),
                          /*want16=*/mid || (use_dw16 && L == 2), /*chain16=*/use_dw16 && L > 2));
      tail_done = true;
      break;
    }
    bool done = false;
    {
      char nm[16];
      snprintf(nm, sizeof(nm), "fwd%d", l);
      ProfScope ps(ctx, nm);
      if (!last && net->prec != B200_PREC_FP32) {
        const TcFuseLast fuse{t, inv_batch};
        B200_TRY(tc_forward_layer(net, l, params, cur, batch, use_tail ? nullptr : &fuse, &done, &fused_last));
      }
      B200_REQUIRE(done || !mid, "the fp16 layer-0 kernel did not take the layer the mid16 path was planned on");
      if (!done) B200_TRY(launch_fwd_layer(net, l, params, cur, batch, last, t, inv_batch));
    }
    cur = net->act[l];
  }
  net->last_batch = batch;

  // backward sweep
  bool d16_ready = use_dw16 && L == 2; // fp16 {hi | lo} delta_0 in net->delta16
  bool side_pending = false;
  for (int l = L - 1; l >= 0; --l) {
    const int K = net->dims[l], N = net->dims[l + 1];
    const float *W = params + net->offs[l];
    const float *in = (l == 0) ? x : net->act[l - 1];
    const bool side_dw1 = mid && l == 1 && env().side && ctx->side_stream && !ctx->prof.on;
    if (side_dw1) B200_CUDA(cudaEventRecord(ctx->ev_fork, st)); // (before dX_1 is launched: dW_1 does not wait for it)
    if (l > 0 && !(l == L - 1 && (fused_last || tail_done))) { // delta_{l-1} = (delta_l W_l^T) .* act'_{l-1}(A_{l-1})
      char nm[16];
      snprintf(nm, sizeof(nm), "dx%d", l);
      ProfScope ps(ctx, nm);
      bool done = false;
      if (mid && l == 1) {
        B200_TRY(mid16_dx_layer1(net, batch));
        d16_ready = true;
        done = true;
      }
      if (!done && l == L - 1 && net->prec == B200_PREC_TF32X3 && wide16_last_dx_applicable(net, batch)) {
        B200_TRY(wide16_last_dx(net, params, batch)); // delta_{L-1} only as the fp16 pair the wide kernels of layer L-1 read
        done = true;
      }
      if (!done && net->prec != B200_PREC_FP32) {
        bool emit16 = (l == 1 && use_dw16 && L > 2);
        B200_TRY(tc_dx_layer(net, l, params, batch, &done, &emit16));
        if (emit16) d16_ready = true;
      }
      if (!done) {
        GemmParams p{};
        p.A = net->delta[l]; p.lda = net->ldd[l];
        p.B = W; p.ldb = N;
        p.M = (int)batch; p.N = K; p.K = N;
        p.vecA = aligned16(p.A) && (N % 4 == 0) && (net->ldd[l] % 4 == 0);
        p.vecB = aligned16(W) && (N % 4 == 0);
        p.a_ones_row = -1;
        p.k_chunk = N;
        p.spec_st = net->spec_st; p.spec = net->spec_flag;
        p.act = net->acts[l - 1];
        p.out = net->delta[l - 1]; p.ldo = net->ldd[l - 1];
        p.aux = net->act[l - 1]; p.ld_aux = K;
        B200_TRY((launch_gemm_simt<true, true, EPI_DX>(p, 1, st)));
      }
    }
    if (!(l == L - 1 && tail_done)) { // [dW; db] partials = [A_{l-1} | 1]^T delta_l over batch slices
      char nm[16];
      snprintf(nm, sizeof(nm), "dw%d", l);
      ProfScope ps(ctx, nm);
      bool done = false;
      if (mid && l == 1) {
        if (side_dw1) {
          // dW_1 needs A_1 and delta_1 only: it runs on the side stream, beside the dX_1 (launched first) -> dW_0 chain; its
          // CTAs fill the SMs that chain leaves idle in its last rounds, and it is joined before the split-K combine
          B200_CUDA(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
          ctx->stream = ctx->side_stream;
          const int rc = mid16_dw_layer1(net, batch);
          ctx->stream = st;
          B200_TRY(rc);
          B200_CUDA(cudaEventRecord(ctx->ev_join, ctx->side_stream));
          side_pending = true;
        } else {
          B200_TRY(mid16_dw_layer1(net, batch));
        }
        done = true;
      }
      if (l == 0) net->dw0_tail_row0 = -1; // (set by dw16_layer when its last feature group has its own number of slices)
      if (l == 0 && d16_ready) {
        X16View xv;
        if (net_x16_view(net, x, batch, &xv)) B200_TRY(dw16_layer(net, xv, batch, &done));
      }
      if (!done && net->prec != B200_PREC_FP32) B200_TRY(tc_dw_layer(net, l, in, batch, &done));
      if (!done && N <= 16 && K + 1 <= 160 && net->prec != B200_PREC_FP32) {
        const int splits = std::max(1, std::min(net->skinny_splits[l], ceil_div(batch, kSkinnyTile)));
        const int chunk = ceil_div(ceil_div(batch, splits), kSkinnyTile) * kSkinnyTile;
        const int s_used = ceil_div(batch, chunk);
        float *part = net->partials + net->part_off[l];
        const unsigned long long ps = (unsigned long long)(K + 1) * N;
        if (K + 1 <= 96) B200_LAUNCH(skinny_dw_kernel<3>, s_used, 256, 0, st, in, net->delta[l], net->ldd[l], K, N, batch, chunk, part, ps,
                                     net->spec_st, net->spec_flag);
        else B200_LAUNCH(skinny_dw_kernel<5>, s_used, 256, 0, st, in, net->delta[l], net->ldd[l], K, N, batch, chunk, part, ps,
                         net->spec_st, net->spec_flag);
        net->splits_used[l] = s_used;
        done = true;
      }
      if (!done) {
        GemmParams p{};
        p.A = in; p.lda = K;
        p.B = net->delta[l]; p.ldb = net->ldd[l];
        p.M = K + 1; p.N = N; p.K = (int)batch;
        p.vecA = aligned16(in) && (K % 4 == 0);
        p.vecB = aligned16(p.B) && (N % 4 == 0);
        p.a_ones_row = K;
        p.k_chunk = net->k_chunk[l];
        p.spec_st = net->spec_st; p.spec = net->spec_flag;
        p.out = net->partials + net->part_off[l];
        B200_TRY((launch_gemm_simt<false, false, EPI_DW>(p, net->splits[l], st)));
        net->splits_used[l] = net->splits[l];
      }
    }
  }

  if (side_pending) B200_CUDA(cudaStreamWaitEvent(st, ctx->ev_join, 0));
  // split-K reduction straight into the caller's gradient buffer (+ L2 term, + ||g||^2 partials)
  FinParams fp{};
  fp.nl = 0;
  for (int l = 0; l < L; ++l) {
    FinLayer &f = fp.L[fp.nl++];
    f.off = net->offs[l];
    f.size = f.stride = (unsigned long long)(net->dims[l] + 1) * net->dims[l + 1];
    f.part = net->partials + net->part_off[l];
    f.splits = net->splits_used[l];
    if (l == 0 && net->dw0_tail_row0 >= 0) { // the last feature group of the fp16 dW kernel has its own number of slices
      const unsigned long long head = (unsigned long long)net->dw0_tail_row0 * net->dims[1];
      FinLayer &t = fp.L[fp.nl++];
      t = f;
      t.off = f.off + head; t.size = f.size - head; t.part = f.part + head; t.splits = net->dw0_tail_splits;
      f.size = head; // (f stays valid: fp.L is an array)
    }
    if (l == 1 && mid) { // dW_1 from the split-K kernel, db_1 from the last-layer backward kernel's column sums
      f.size = (unsigned long long)net->dims[1] * net->dims[2];
      FinLayer &b = fp.L[fp.nl++];
      b.off = f.off + f.size;
      b.size = b.stride = (unsigned long long)net->dims[2];
      b.part = net->m16.db_part;
      b.splits = net->m16.db_splits;
    }
  }
  fp.n = net->n;
  fp.partials = net->partials;
  fp.w = params;
  fp.lam = net->l2 / (float)ctx->world; // each rank adds its share; the all-reduce sums them
  fp.grad = grad_out;
  fp.fin_part = net->fin_part;
  fp.zero = reinterpret_cast<const float *>(net->fin_part + 2 * net->fin_blocks);
  const bool multi = ctx->world > 1 && !net->defer_reduce;
  // collective (every rank evaluates the same networks in the same order); the first evaluation of a network is never inside a
  // graph capture. A network larger than the buffers were made for re-makes them.
  if (multi && (!ctx->p2p.tried || net->n > ctx->p2p.req_floats)) B200_TRY(ctx_p2p_setup(ctx, net->n));
  const bool p2p = multi && ctx->p2p.ready && net->n <= ctx->p2p.slot_floats && (reinterpret_cast<uintptr_t>(grad_out) & 15u) == 0;
  {
    ProfScope ps(ctx, "finalize");
    fp.done_count = net->fin_done;
    fp.loss_part = net->loss_part; fp.n_loss = net->loss_part_n;
    fp.inv_batch = (double)inv_batch; fp.lam_d = (double)fp.lam;
    fp.want_gnorm = multi ? 0 : 1;
    fp.out = out;
    fp.spec_st = net->spec_st; fp.spec = net->spec_flag; fp.host_out = net->host_out;
    if (p2p) {
      fp.sym_local = ctx->p2p.local; fp.peers = ctx->p2p.peers_dev;
      fp.slot_bytes = ctx->p2p.slot_bytes; fp.slot_floats = ctx->p2p.slot_floats;
      fp.rank = ctx->rank; fp.world = ctx->world;
    }
    B200_LAUNCH(finalize_grad_kernel, net->fin_blocks, 256, 0, st, fp);
  }
  if (p2p) {
    ProfScope ps(ctx, "allreduce_p2p");
    P2PReduceParams rp{ctx->p2p.peers_dev, ctx->p2p.local, ctx->p2p.slot_bytes, ctx->p2p.slot_floats, net->n, ctx->rank, ctx->world,
                       grad_out, net->fin_part, net->fin_done, out, net->spec_st, net->spec_flag, net->host_out,
                       ctx->h_scalars + kHostErrSlot,
                       (unsigned long long)(env().p2p_spin_limit > 0 ? env().p2p_spin_limit : 20000000L)}; // ~20 s at 1 us per poll
    const int blocks = std::max(1, std::min(net->fin_blocks, ceil_div((long)((net->n + 3) / 4), 256)));
    B200_LAUNCH(p2p_reduce_kernel, blocks, 256, 0, st, rp);
  } else if (multi) {
    ProfScope ps(ctx, "allreduce");
    B200_TRY(ctx_allreduce(ctx, grad_out, net->n, &out->loss));
    B200_LAUNCH(sumsq_part_kernel, net->fin_blocks, 256, 0, st, grad_out, (unsigned long long)net->n, net->fin_part);
    B200_LAUNCH(gnorm_from_parts_kernel, 1, 256, 0, st, net->fin_part, net->fin_blocks, out);
  }
  return B200_OK;
}

} // namespace b200

using namespace b200;

// =====================================================================================================
// C ABI: network
// =====================================================================================================
extern "C" {

int b200_net_create(b200_ctx *ctx, int nlayers, const int *dims, const int *acts, b200_net **out) {
  B200_REQUIRE(ctx && dims && acts && out, "null argument");
  B200_REQUIRE(nlayers >= 1 && nlayers <= kMaxLayers, "1..16 layers supported");
  for (int l = 0; l <= nlayers; ++l) B200_REQUIRE(dims[l] > 0, "layer dimensions must be positive");
  for (int l = 0; l < nlayers; ++l) B200_REQUIRE(acts[l] >= 0 && acts[l] <= 3, "unknown activation");
  b200_net *net = new b200_net;
  static std::atomic<unsigned long long> next_uid{1};
  net->uid = next_uid.fetch_add(1);
  net_register(net->uid, net);
  net->ctx = ctx;
  net->dims.assign(dims, dims + nlayers + 1);
  net->acts.assign(acts, acts + nlayers);
  net->offs.resize(nlayers);
  size_t off = 0;
  for (int l = 0; l < nlayers; ++l) {
    net->offs[l] = off;
    off += (size_t)dims[l + 1] * dims[l] + dims[l + 1];
  }
  net->n = off;
  net->act.assign(nlayers, nullptr);
  net->delta.assign(nlayers, nullptr);
  net->ldd.resize(nlayers);
  for (int l = 0; l < nlayers; ++l) net->ldd[l] = (dims[l + 1] + 3) & ~3;
  net->splits.assign(nlayers, 1);
  net->splits_used.assign(nlayers, 1);
  net->skinny_splits.assign(nlayers, 1);
  net->k_chunk.assign(nlayers, 16);
  net->part_off.assign(nlayers, 0);
  net->fin_blocks = std::max(1, std::min(4 * ctx->num_sms, ceil_div((long)net->n, 256))); // (a warp per group of 32 elements)
  cudaSetDevice(ctx->device);
  B200_CUDA(cudaMalloc(&net->fin_part, sizeof(double) * (2 * net->fin_blocks + 2))); // (+ a zero: FinParams::zero)
  B200_CUDA(cudaMemset(net->fin_part, 0, sizeof(double) * (2 * net->fin_blocks + 2)));
  B200_CUDA(cudaMalloc(&net->eval_out, sizeof(EvalOut)));
  B200_CUDA(cudaMalloc(&net->fin_done, sizeof(unsigned)));
  B200_CUDA(cudaMemset(net->fin_done, 0, sizeof(unsigned)));
  *out = net;
  return B200_OK;
}

int b200_net_destroy(b200_net *net) {
  if (!net) return B200_OK;
  net_register(net->uid, nullptr);
  if (ctx_is_live(net->ctx)) { // (the context may have been destroyed first: its stream and pooled solvers went with it)
    cudaSetDevice(net->ctx->device);
    cudaStreamSynchronize(net->ctx->stream);
    lbfgs_pool_forget_net(net->ctx, net->uid);
  } else {
    cudaDeviceSynchronize();
  }
  free_batch_buffers(net);
  tc_release(net);
  tail_release(net);
  if (net->xq.data) cudaFree(net->xq.data);
  if (net->xq.data16) cudaFree(net->xq.data16);
  if (net->xq.flag) cudaFree(net->xq.flag);
  if (net->xg.data16) cudaFree(net->xg.data16);
  if (net->partials) cudaFree(net->partials);
  if (net->fin_part) cudaFree(net->fin_part);
  if (net->eval_out) cudaFree(net->eval_out);
  if (net->fin_done) cudaFree(net->fin_done);
  if (net->params) cudaFree(net->params);
  if (net->grads) cudaFree(net->grads);
  delete net;
  return B200_OK;
}

size_t b200_net_params_size(b200_net *net) { return net ? net->n : 0; }
int b200_net_output_size(b200_net *net) { return net ? net->dims.back() : 0; }

static float activation_scale(int act) { // src/cuda/kernels.cuh:61-71
  return act == B200_ACT_RELU ? 1.41421356f : 1.0f;
}

int b200_net_bind_params(b200_net *net, unsigned seed) {
  B200_REQUIRE(net, "null net");
  cudaSetDevice(net->ctx->device);
  if (!net->params) B200_CUDA(cudaMalloc(&net->params, sizeof(float) * net->n));
  if (!net->grads) B200_CUDA(cudaMalloc(&net->grads, sizeof(float) * net->n));
  // host-side init, same RNG consumption as CudaNetwork::bindParams (src/cuda/network.cuh:37-59)
  std::vector<float> host(net->n);
  std::mt19937 gen(seed);
  size_t off = 0;
  for (int l = 0; l < net->nlayers(); ++l) {
    const size_t wc = (size_t)net->dims[l + 1] * net->dims[l], bc = net->dims[l + 1];
    const float sd = activation_scale(net->acts[l]) * std::sqrt(1.0f / (float)net->dims[l]);
    std::normal_distribution<float> dist(0.0f, sd);
    for (size_t i = 0; i < wc; ++i) host[off + i] = dist(gen);
    for (size_t i = 0; i < bc; ++i) host[off + wc + i] = 0.0f;
    off += wc + bc;
  }
  B200_CUDA(cudaMemcpyAsync(net->params, host.data(), sizeof(float) * net->n, cudaMemcpyHostToDevice, net->ctx->stream));
  B200_CUDA(cudaMemsetAsync(net->grads, 0, sizeof(float) * net->n, net->ctx->stream));
  B200_CUDA(cudaStreamSynchronize(net->ctx->stream));
  return B200_OK;
}

float *b200_net_params_data(b200_net *net) { return net ? net->params : nullptr; }
float *b200_net_grads_data(b200_net *net) { return net ? net->grads : nullptr; }

int b200_net_zero_grads(b200_net *net) {
  B200_REQUIRE(net && net->grads, "params not bound");
  B200_CUDA(cudaMemsetAsync(net->grads, 0, sizeof(float) * net->n, net->ctx->stream));
  return B200_OK;
}

int b200_net_set_precision(b200_net *net, int prec) {
  B200_REQUIRE(net, "null net");
  B200_REQUIRE(prec >= B200_PREC_FP32 && prec <= B200_PREC_TF32, "unknown precision mode");
  net->prec = prec;
  ++net->config_gen;
  return B200_OK;
}
int b200_net_get_precision(b200_net *net) { return net ? net->prec : -1; }

int b200_net_set_l2(b200_net *net, float lambda) {
  B200_REQUIRE(net, "null net");
  net->l2 = lambda;
  ++net->config_gen;
  return B200_OK;
}

int b200_net_set_global_batch(b200_net *net, long batch_global) {
  B200_REQUIRE(net && batch_global >= 0, "bad argument");
  net->batch_global = batch_global;
  ++net->config_gen;
  return B200_OK;
}

int b200_net_quantize_input(b200_net *net, const float *x_dev, long batch, int *quantized) {
  B200_REQUIRE(net && x_dev, "null argument");
  cudaSetDevice(net->ctx->device);
  B200_TRY(net_quantize_input(net, x_dev, batch, true));
  if (net->xq.valid) net->xq.user = true;
  if (quantized) *quantized = net->xq.valid ? 1 : 0;
  return B200_OK;
}
int b200_net_clear_input_cache(b200_net *net) {
  B200_REQUIRE(net, "null net");
  net_xq_clear(net);
  return B200_OK;
}

int b200_net_forward(b200_net *net, const float *x_dev, long batch) {
  B200_REQUIRE(net && net->params && x_dev, "null argument / params not bound");
  cudaSetDevice(net->ctx->device);
  return net_forward(net, net->params, x_dev, batch);
}

int b200_net_loss_grad_async(b200_net *net, const float *params_dev, const float *x_dev, const float *t_dev, long batch,
                             float *grad_dev, double *loss_dev) {
  B200_REQUIRE(net && x_dev && t_dev, "null argument");
  const float *pp = params_dev ? params_dev : net->params;
  float *gg = grad_dev ? grad_dev : net->grads;
  B200_REQUIRE(pp && gg, "params not bound");
  cudaSetDevice(net->ctx->device);
  B200_TRY(net_eval(net, pp, x_dev, t_dev, batch, 0, gg, (EvalOut *)net->eval_out));
  if (loss_dev)
    B200_CUDA(cudaMemcpyAsync(loss_dev, net->eval_out, sizeof(double), cudaMemcpyDeviceToDevice, net->ctx->stream));
  return B200_OK;
}

int b200_net_loss_grad(b200_net *net, const float *x_dev, const float *t_dev, long batch, float *loss_host) {
  B200_REQUIRE(net && net->params, "params not bound");
  B200_TRY(b200_net_loss_grad_async(net, nullptr, x_dev, t_dev, batch, nullptr, nullptr));
  b200_ctx *ctx = net->ctx;
  B200_CUDA(cudaMemcpyAsync(ctx->h_scalars, net->eval_out, sizeof(EvalOut), cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(cudaStreamSynchronize(ctx->stream));
  if (loss_host) *loss_host = (float)ctx->h_scalars[0];
  return ctx_check_device_error(ctx); // (a device-side wait that gave up during this evaluation)
}

int b200_net_copy_output_to_host(b200_net *net, float *host, size_t n) {
  B200_REQUIRE(net && host, "null argument");
  if (net->act.empty() || !net->act.back() || net->last_batch == 0) return B200_OK; // network.cuh:122-124
  B200_REQUIRE(n <= (size_t)net->dims.back() * net->last_batch, "more elements requested than the last batch produced");
  B200_CUDA(cudaMemcpyAsync(host, net->act.back(), sizeof(float) * n, cudaMemcpyDeviceToHost, net->ctx->stream));
  B200_CUDA(cudaStreamSynchronize(net->ctx->stream));
  return B200_OK;
}

int b200_net_last_batch(b200_net *net) { return net ? (int)net->last_batch : 0; }

int b200_net_copy_activation_to_host(b200_net *net, int layer, float *host, size_t n) {
  B200_REQUIRE(net && host, "null argument");
  B200_REQUIRE(layer >= 0 && layer < net->nlayers(), "no such layer");
  B200_REQUIRE(net->act[layer] && net->last_batch > 0, "no forward pass has run");
  B200_REQUIRE(n <= (size_t)net->dims[layer + 1] * net->last_batch, "more elements requested than the last batch produced");
  if (layer == 0) B200_TRY(mid16_reconstruct_act0(net)); // (A_1 may exist only as its fp16 pair)
  B200_CUDA(cudaMemcpyAsync(host, net->act[layer], sizeof(float) * n, cudaMemcpyDeviceToHost, net->ctx->stream));
  B200_CUDA(cudaStreamSynchronize(net->ctx->stream));
  return B200_OK;
}

int b200_net_evaluate(b200_net *net, const float *x_dev, const float *t_dev, long batch, double *mse, double *accuracy) {
  B200_REQUIRE(net && net->params && x_dev && t_dev, "null argument / params not bound");
  cudaSetDevice(net->ctx->device);
  B200_TRY(net_forward(net, net->params, x_dev, batch));
  const int od = net->dims.back();
  const int blocks = std::max(1, std::min(net->fin_blocks, ceil_div(batch, 256)));
  B200_LAUNCH(evaluate_kernel, blocks, 256, 0, net->ctx->stream, net->act.back(), t_dev, batch, od, net->fin_part);
  std::vector<double> part(2 * blocks);
  B200_CUDA(cudaMemcpyAsync(part.data(), net->fin_part, sizeof(double) * 2 * blocks, cudaMemcpyDeviceToHost,
                            net->ctx->stream));
  B200_CUDA(cudaStreamSynchronize(net->ctx->stream));
  double se = 0.0, correct = 0.0;
  for (int i = 0; i < blocks; ++i) { se += part[2 * i]; correct += part[2 * i + 1]; }
  if (mse) *mse = se / ((double)batch * od);               // unified_launcher.hpp:196
  if (accuracy) *accuracy = correct / (double)batch * 100.0; // :197
  return B200_OK;
}

} // extern "C"
