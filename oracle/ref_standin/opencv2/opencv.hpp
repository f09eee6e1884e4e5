// opencv2/opencv.hpp STAND-IN (test infrastructure)
#pragma once
#include <opencv2/core.hpp>
