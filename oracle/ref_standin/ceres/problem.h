// ceres/problem.h STAND-IN (test infrastructure): see ceres/ceres.h
#pragma once
#include <ceres/ceres.h>
