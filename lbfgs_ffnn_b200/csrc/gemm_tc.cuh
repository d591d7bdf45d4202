// tcgen05 / TMEM / TMA tensor-core GEMMs of the objective (B200_PREC_TF32X3 and B200_PREC_TF32 modes).
// Each entry sets *done = true when it handled the layer; otherwise the caller uses the FFMA kernel.
#pragma once

#include "network.cuh"

namespace b200 {

int tc_forward_layer(b200_net *net, int l, const float *params, const float *in, long batch, bool *done);
int tc_dw_layer(b200_net *net, int l, const float *in, long batch, bool *done);
void tc_release(b200_net *net);

} // namespace b200
