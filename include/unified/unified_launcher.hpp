// forwarding header for the reference include "src/unified_launcher.hpp"
#pragma once
#include "unified.hpp"
