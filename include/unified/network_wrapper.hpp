// forwarding header for the reference include "src/network_wrapper.hpp"
#pragma once
#include "unified.hpp"
