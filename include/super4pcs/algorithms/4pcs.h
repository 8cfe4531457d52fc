// super4pcs-b200: GlobalRegistration::Match4PCS.
//
// The reference's Match4PCS (src/super4pcs/algorithms/4pcs.{h,cc}) is the legacy O(n^2) 4PCS kept
// for comparison (demo flag -x); it is outside the hot path this project rebuilds (SURVEY.md 2, row
// 11).  So that callers that name the type keep compiling and running (demos/Super4PCS/
// super4pcs_test.cc:141, the Meshlab plugin), Match4PCS is provided with the reference's
// constructor signature and runs the same GPU pipeline as MatchSuper4PCS.
#ifndef SUPER4PCS_B200_ALGO_4PCS_H_
#define SUPER4PCS_B200_ALGO_4PCS_H_

#include "super4pcs/algorithms/super4pcs.h"

namespace GlobalRegistration {

class Match4PCS : public MatchSuper4PCS {
 public:
  using Base = Match4PCSBase;
  using Scalar = typename Base::Scalar;
  using PairsVector = typename Base::PairsVector;
  using VectorType = typename Base::VectorType;

  explicit Match4PCS(const Match4PCSOptions& options, const Utils::Logger& logger)
      : MatchSuper4PCS(options, logger) {}

  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};

}  // namespace GlobalRegistration

#endif  // SUPER4PCS_B200_ALGO_4PCS_H_
