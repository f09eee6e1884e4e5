// ceres/ceres.h STAND-IN (test infrastructure): types that let estimate_pose_bfgs() in
// src/vlcal/calib/visual_camera_calibration.cpp COMPILE.  There is no solver behind them: ceres::Solve throws.  The
// pin tests only run the Nelder-Mead branch of that file.
#pragma once

#include <ceres/jet.h>

#include <stdexcept>
#include <vector>

namespace ceres {

enum CallbackReturnType { SOLVER_CONTINUE, SOLVER_ABORT, SOLVER_TERMINATE_SUCCESSFULLY };
enum LineSearchDirectionType { STEEPEST_DESCENT, NONLINEAR_CONJUGATE_GRADIENT, LBFGS, BFGS };

struct IterationSummary {};

class IterationCallback {
public:
  virtual ~IterationCallback() {}
  virtual CallbackReturnType operator()(const IterationSummary& summary) = 0;
};

class FirstOrderFunction {
public:
  virtual ~FirstOrderFunction() {}
  virtual bool Evaluate(const double* const parameters, double* cost, double* gradient) const = 0;
  virtual int NumParameters() const = 0;
};

class Manifold {
public:
  virtual ~Manifold() {}
};

template <class Functor, int N>
class AutoDiffFirstOrderFunction : public FirstOrderFunction {
public:
  explicit AutoDiffFirstOrderFunction(Functor* f) : functor(f) {}
  ~AutoDiffFirstOrderFunction() override { delete functor; }
  bool Evaluate(const double* const, double*, double*) const override { return false; }  // no autodiff behind the stand-in
  int NumParameters() const override { return N; }

private:
  Functor* functor;
};

class GradientProblem {
public:
  GradientProblem(FirstOrderFunction* f, Manifold* m) : function(f), manifold(m) {}
  ~GradientProblem() {
    delete function;
    delete manifold;
  }

private:
  FirstOrderFunction* function;
  Manifold* manifold;
};

struct GradientProblemSolver {
  struct Options {
    bool minimizer_progress_to_stdout = false;
    bool update_state_every_iteration = false;
    LineSearchDirectionType line_search_direction_type = LBFGS;
    std::vector<IterationCallback*> callbacks;
  };
  struct Summary {
    std::vector<IterationSummary> iterations;
    double final_cost = 0.0;
  };
};

inline void Solve(const GradientProblemSolver::Options&, const GradientProblem&, double*, GradientProblemSolver::Summary*) {
  throw std::runtime_error("ceres stand-in: no solver (only the Nelder-Mead branch of the reference is runnable here)");
}

}  // namespace ceres
