// drop-in for the reference header src/cuda/minimizer_base.cuh
#pragma once
#include "../../../unified/unified.hpp"
