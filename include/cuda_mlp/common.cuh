// forwarding header: the reference includes "cuda/common.cuh"; everything lives in cuda_mlp.hpp + unified.hpp
#pragma once
#include "../unified/unified.hpp"
