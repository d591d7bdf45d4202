// drop-in for the reference header src/cuda/layer.cuh
#pragma once
#include "../../../unified/unified.hpp"
