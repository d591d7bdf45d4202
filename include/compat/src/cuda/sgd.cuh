// drop-in for the reference header src/cuda/sgd.cuh
#pragma once
#include "../../../unified/unified.hpp"
