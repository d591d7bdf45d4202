// drop-in for the reference header src/cuda/device_buffer.cuh
#pragma once
#include "../../../unified/unified.hpp"
