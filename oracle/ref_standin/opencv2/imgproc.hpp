// opencv2/imgproc.hpp STAND-IN: nothing from imgproc is used by the compiled sources
#pragma once
#include <opencv2/core.hpp>
