// drop-in for the reference header src/cuda/gd.cuh
#pragma once
#include "../../../unified/unified.hpp"
