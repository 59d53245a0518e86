// glog is not in this image; LOG / VLOG come from this repo's scanner/util/common.h (glog-shaped).
#pragma once
#include "scanner/util/common.h"
