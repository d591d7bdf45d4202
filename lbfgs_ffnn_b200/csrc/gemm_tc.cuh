// tcgen05 / TMEM / TMA tensor-core GEMMs of the objective (B200_PREC_TF32X3 and B200_PREC_TF32 modes).
// Each entry sets *done = true when it handled the layer; otherwise the caller uses the FFMA kernel.
#pragma once

#include "network.cuh"

#include <cuda.h>

namespace b200 {

struct TcFuseLast { // request to fuse the last layer + loss + deltas into the penultimate layer's forward epilogue
  const float *targets;
  float inv_batch;
};
int tc_forward_layer(b200_net *net, int l, const float *params, const float *in, long batch, const TcFuseLast *fuse,
                     bool *done, bool *fused);
// emit16 (in/out): write delta_{l-1} as scaled fp16 {hi | lo} into net->delta16 instead of fp32 (see gemm_dw16.cu); false on
// return when the fp32 rows were written after all
int tc_dx_layer(b200_net *net, int l, const float *params, long batch, bool *done, bool *emit16 = nullptr);
int tc_dw_layer(b200_net *net, int l, const float *in, long batch, bool *done);
// split-K plan of the tensor-core dW kernel: returns K blocks per split, *splits = number of splits
int tc_dw_plan(b200_net *net, int l, long batch, int *splits);
int tc_split_params(b200_net *net, const float *params);
// persistent fp16 forward of layer 0 on the uint8 copy of the input (gemm_fwd16.cu)
int fwd16_forward_layer(b200_net *net, int l, const float *params, const X16View &x16, long batch, bool *done);
int fwd16_prepare(b200_net *net, const float *params);
void fwd16_release(b200_net *net);
// fp16 dW of layer 0 from the uint8 input copy and the fp16 {hi | lo} delta of tail_layer (gemm_dw16.cu)
bool dw16_applicable(const b200_net *net);
int dw16_plan(const b200_net *net, long batch, int *splits); // K blocks per split; *splits = slices of the batch
int dw16_layer(b200_net *net, const X16View &x16, long batch, bool *done);
// hidden layer 1 of a three-layer net on the fp16 kernels (b200_net::Mid16): gemm_fwd16.cu / gemm_dw16.cu
int mid16_forward_layer1(b200_net *net, const float *params, long batch);
int mid16_dx_layer1(b200_net *net, long batch);
int mid16_dw_layer1(b200_net *net, long batch);
int mid16_dw_plan(const b200_net *net, long batch, int *splits);
void tc_release(b200_net *net);

// 2-D fp32 tensor map {dim0 contiguous, dim1 rows} with SWIZZLE_128B (gemm_fwd16.cu)
int tc_make_map_2d_f32(CUtensorMap *tm, const float *ptr, unsigned long long dim0, unsigned long long dim1, unsigned box0, unsigned box1);

} // namespace b200
