// stand-in for <opencv2/core/core.hpp>: see ../../mini_cv.h
#pragma once
#include "../../mini_cv.h"
