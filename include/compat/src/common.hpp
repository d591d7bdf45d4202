// drop-in for the reference header src/common.hpp (CUDA-backend builds only need checkParallelism())
#pragma once
#include "../../unified/unified.hpp"
#include <iostream>
inline void checkParallelism() { std::cout << "[B200] CUDA backend: libb200lbfgs (sm_100a), no CPU path" << std::endl; }
