// Stand-in with the shape of NIDCostParams in the reference's include/vlcal/calib/cost_calculator_nid.hpp:9-14 (test scaffolding only).
#pragma once
namespace vlcal {
struct NIDCostParams {
  NIDCostParams() : bins(16) {}  // src/vlcal/calib/cost_calculator_nid.cpp:7-9
  int bins;
};
}  // namespace vlcal
