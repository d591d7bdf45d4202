// drop-in for the reference header src/unified_launcher.hpp
#pragma once
#include "../../unified/unified.hpp"
