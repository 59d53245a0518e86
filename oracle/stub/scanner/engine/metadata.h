// Stand-in for the reference's scanner/engine/metadata.h while oracle/Makefile compiles the reference's
// scanner/engine/sampler.cpp unmodified (test infrastructure): sampler.cpp needs none of the table metadata
// classes, only the standard headers that arrive through this file upstream.
#pragma once
#include <cassert>
#include <functional>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "scanner/util/common.h"
