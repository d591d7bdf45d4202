// drop-in for the reference header src/network_wrapper.hpp
#pragma once
#include "../../unified/unified.hpp"
