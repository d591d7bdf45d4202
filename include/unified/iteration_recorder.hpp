// forwarding header for the reference include "src/iteration_recorder.hpp"
#pragma once
#include "unified.hpp"
