// forwards to the Eigen-API stand-in (test infrastructure; see eigen_shim.hpp)
#pragma once
#include "../../../eigen_shim.hpp"
