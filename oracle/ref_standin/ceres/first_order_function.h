// ceres/first_order_function.h STAND-IN (test infrastructure): see ceres/ceres.h
#pragma once
#include <ceres/ceres.h>
