// drop-in for the reference header src/iteration_recorder.hpp
#pragma once
#include "../../unified/unified.hpp"
