// forwarding header for the reference include "src/seed.hpp"
#pragma once
#include "unified.hpp"
