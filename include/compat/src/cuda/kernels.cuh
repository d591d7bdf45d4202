// drop-in for the reference header src/cuda/kernels.cuh
#pragma once
#include "../../../unified/unified.hpp"
