// drop-in for the reference header src/unified_optimization.hpp
#pragma once
#include "../../unified/unified.hpp"
