// stand-in for <opencv2/core/eigen.hpp>: see ../../mini_cv.h
#pragma once
#include "../../mini_eigen.h"
#include "../../mini_cv.h"
