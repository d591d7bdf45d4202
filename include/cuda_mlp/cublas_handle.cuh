// forwarding header: the reference includes "cuda/cublas_handle.cuh"; everything lives in cuda_mlp.hpp + unified.hpp
#pragma once
#include "../unified/unified.hpp"
