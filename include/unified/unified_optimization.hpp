// forwarding header for the reference include "src/unified_optimization.hpp"
#pragma once
#include "unified.hpp"
