// drop-in for the reference header src/cuda/network.cuh
#pragma once
#include "../../../unified/unified.hpp"
