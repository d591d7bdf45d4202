// drop-in for the reference header src/cuda/common.cuh
#pragma once
#include "../../../unified/unified.hpp"
