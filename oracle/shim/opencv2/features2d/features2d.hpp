// oracle/shim/opencv2/features2d/features2d.hpp -- TEST INFRASTRUCTURE ONLY: ORBmatcher.hpp includes it and uses nothing of it.
#pragma once
#include "../core/core.hpp"
