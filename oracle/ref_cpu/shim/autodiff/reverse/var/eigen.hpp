// The reference includes this autodiff header (src/layer.hpp:10-11, src/minimizer/lbfgs.hpp:6) but uses nothing from it on the MLP path;
// the vendored library needs the real Eigen, so an empty stand-in is put in front of it (test infrastructure).
#pragma once
