// drop-in for the reference header src/seed.hpp
#pragma once
#include "../../unified/unified.hpp"
