// stand-in for <opencv2/imgproc.hpp>: see ../mini_cv.h
#pragma once
#include "../mini_cv.h"
