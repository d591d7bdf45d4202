// drop-in for the reference header src/cuda/cublas_handle.cuh
#pragma once
#include "../../../unified/unified.hpp"
