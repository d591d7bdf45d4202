// drop-in for the reference header src/cuda/lbfgs.cuh
#pragma once
#include "../../../unified/unified.hpp"
