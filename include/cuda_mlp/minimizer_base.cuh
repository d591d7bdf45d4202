// forwarding header: the reference includes "cuda/minimizer_base.cuh"; everything lives in cuda_mlp.hpp + unified.hpp
#pragma once
#include "../unified/unified.hpp"
