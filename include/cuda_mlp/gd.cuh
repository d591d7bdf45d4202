// forwarding header: the reference includes "cuda/gd.cuh"; everything lives in cuda_mlp.hpp + unified.hpp
#pragma once
#include "../unified/unified.hpp"
