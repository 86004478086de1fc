/* oracle/shim/dbow2/TemplatedVocabulary.h -- TEST INFRASTRUCTURE ONLY.
 * The reference's vendored ScoringObject.cpp reaches its own header through "TemplatedVocabulary.h" (an
 * OpenCV-backed template it uses nothing else of); the build recipe compiles ScoringObject.cpp through a link placed
 * beside this file, which passes the include on to the reference's ScoringObject.h and the standard headers the real
 * TemplatedVocabulary.h brings along. */
#pragma once
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>
#include "ScoringObject.h"
