// stand-in for <opencv2/opencv.hpp>: see ../mini_cv.h
#pragma once
#include "../mini_cv.h"
