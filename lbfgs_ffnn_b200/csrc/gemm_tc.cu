#include "gemm_tc.cuh"

namespace b200 {

int tc_forward_layer(b200_net *, int, const float *, const float *, long, bool *done) {
  *done = false;
  return B200_OK;
}
int tc_dw_layer(b200_net *, int, const float *, long, bool *done) {
  *done = false;
  return B200_OK;
}
void tc_release(b200_net *) {}

} // namespace b200
