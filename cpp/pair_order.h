// super4pcs-b200 (internal to the C++ layer): the ORDER in which the reference emits the pairs of one ExtractPairs call.
//
// The reference does not sort its pair lists (src/super4pcs/algorithms/super4pcs.cc:183-224): they reach
// FindCongruentQuadrilaterals in the emission order of IntersectionFunctor::process
// (accelerators/pairExtraction/intersectionFunctor.h:104-236) -- for every primitive (= point id i, ascending) the leaves
// of an octree over the unit cube in a fixed order, inside a leaf the points in the order of an id array that is
// partitioned IN PLACE by every split (intersectionNode.h:165-249) and never reset between calls
// (pairCreationFunctor.h:118-121).  The candidate order that follows decides which of several candidates with EQUAL
// inlier counts is kept (match4pcsBase.hpp:468).  The device produces the same pair SET, sorted; this class replays the
// reference's sequential traversal on the host -- cheap at the sample sizes where ties occur -- and yields, per call, the
// "leaf position" of every point id.  An ordered pair (a, b) then sorts by (max(a,b), leaf position of min(a,b), a > b).
#ifndef SUPER4PCS_B200_CPP_PAIR_ORDER_H_
#define SUPER4PCS_B200_CPP_PAIR_ORDER_H_

#include <cstddef>
#include <cstdint>
#include <vector>

namespace GlobalRegistration {
namespace detail {

class PairOrder {
 public:
  static constexpr uint32_t kAbsent = 0xFFFFFFFFu;  ///< the point is in no leaf of this call (it has no pair)

  /// unit-cube coordinates (3 floats per point) of sampled Q; the id array becomes the identity
  void Reset(const std::vector<float>& unit_xyz, float ratio);
  bool empty() const { return ids_.empty(); }
  size_t size() const { return ids_.size(); }

  /// Replays one ExtractPairs(pair_distance, pair_distance_epsilon) call: advances the id array exactly like the
  /// reference and fills leaf_position[id] for every point id.
  void Replay(float pair_distance, float pair_distance_epsilon, std::vector<uint32_t>* leaf_position);

  /// the history-dependent state (for rolling back calls of bases that were selected ahead but never tried)
  const std::vector<uint32_t>& state() const { return ids_; }
  void set_state(const std::vector<uint32_t>& ids) { ids_ = ids; }

  /// sort key of the ordered pair (a, b) of a call whose leaf positions are `pos`
  static inline uint64_t Key(const std::vector<uint32_t>& pos, int a, int b) {
    const uint32_t hi = uint32_t(a > b ? a : b), lo = uint32_t(a > b ? b : a);
    return (uint64_t(hi) << 33) | (uint64_t(pos[lo]) << 1) | uint64_t(a > b ? 1 : 0);
  }

 private:
  struct Node {
    float c[3];
    uint32_t begin, end;
  };
  bool ShellHitsBox(const float* centre, float radius, const float* box, float half) const;
  uint32_t Partition(int first, int last, int dim, float value);
  void Split(const Node& n, float half, std::vector<Node>* out);

  std::vector<float> unit_;  ///< 3 floats per point
  std::vector<uint32_t> ids_;
  float ratio_ = 1.f;
};

}  // namespace detail
}  // namespace GlobalRegistration

#endif  // SUPER4PCS_B200_CPP_PAIR_ORDER_H_
