// forwarding header: the reference includes "cuda/device_buffer.cuh"; everything lives in cuda_mlp.hpp + unified.hpp
#pragma once
#include "../unified/unified.hpp"
