// super4pcs-b200: host replay of the reference's pair emission order (see cpp/pair_order.h).
#include "pair_order.h"

#include <algorithm>
#include <cmath>
#include <utility>

namespace GlobalRegistration {
namespace detail {

constexpr uint32_t PairOrder::kAbsent;

void PairOrder::Reset(const std::vector<float>& unit_xyz, float ratio) {
  unit_ = unit_xyz;
  ratio_ = ratio;
  ids_.resize(unit_xyz.size() / 3);
  for (size_t i = 0; i < ids_.size(); ++i) ids_[i] = uint32_t(i);
}

// Box-sphere test on the SHELL of radius r (Arvo; the reference's HyperSphere::intersect,
// intersectionPrimitive.h:108-130): the box reaches inside the sphere and is not swallowed by it.
bool PairOrder::ShellHitsBox(const float* centre, float radius, const float* box, float half) const {
  float inside[3], farthest[3];
  for (int d = 0; d < 3; ++d) {
    const float lo = box[d] - half, hi = box[d] + half;
    const float to_lo = (centre[d] - lo) * (centre[d] - lo), to_hi = (centre[d] - hi) * (centre[d] - hi);
    inside[d] = centre[d] < lo ? to_lo : (centre[d] > hi ? to_hi : 0.f);
    farthest[d] = std::max(to_lo, to_hi);
  }
  const float nearest2 = inside[0] + (inside[1] + inside[2]);  // summation order of Eigen's 3-element reduction
  const float farthest2 = farthest[0] + (farthest[1] + farthest[2]);
  const float r2 = radius * radius;
  return nearest2 < r2 && r2 < farthest2;
}

// In-place two-sided partition of ids_[first, last) by coordinate `dim` < value (NdNode::_split); returns the cut.
uint32_t PairOrder::Partition(int first, int last, int dim, float value) {
  auto coord = [this, dim](int k) { return unit_[3 * size_t(ids_[size_t(k)]) + size_t(dim)]; };
  int lo = first, hi = last - 1;
  while (lo < hi) {
    while (lo < last && coord(lo) < value) ++lo;
    while (hi >= first && coord(hi) >= value) --hi;
    if (lo > hi) break;
    std::swap(ids_[size_t(lo)], ids_[size_t(hi)]);
    ++lo;
    --hi;
  }
  if (lo >= last) return uint32_t(last);
  return coord(lo) < value ? uint32_t(lo + 1) : uint32_t(lo);
}

// Eight children, dimension by dimension (NdNode::split); empty children are not kept.
void PairOrder::Split(const Node& n, float half, std::vector<Node>* out) {
  Node kid[8];
  for (Node& k : kid) k = n;
  const float quarter = half / 2.f;
  for (int d = 0; d < 3; ++d) {
    const int groups = 1 << d, span = 8 >> d, mid = span / 2;
    for (int g = 0; g < groups; ++g) {
      Node* first = kid + g * span;
      const float centre = first->c[d];
      const uint32_t cut = Partition(int(first->begin), int(first[span - 1].end), d, centre);
      for (int i = 0; i < mid; ++i) {
        first[i].c[d] = centre - quarter;
        first[i].end = cut;
      }
      for (int i = mid; i < span; ++i) {
        first[i].c[d] = centre + quarter;
        first[i].begin = cut;
      }
    }
  }
  for (const Node& k : kid)
    if (k.end != k.begin) out->push_back(k);
}

void PairOrder::Replay(float pair_distance, float pair_distance_epsilon, std::vector<uint32_t>* leaf_position) {
  const int n = int(ids_.size());
  leaf_position->assign(size_t(n), kAbsent);
  const float radius = pair_distance / ratio_;                      // PairCreationFunctor::setRadius
  const float eps_norm = pair_distance_epsilon / ratio_;            // getNormalizedEpsilon
  const int lvl_max = int(-std::log2(eps_norm));                    // GetRoundedEpsilonValue
  const float eps = float(1.f / std::pow(2, lvl_max));
  const unsigned kMinNodeSize = 50;                                 // super4pcs.cc:219
  std::vector<Node> level, next;
  std::vector<std::pair<Node, float>> early;
  next.push_back(Node{{0.5f, 0.5f, 0.5f}, 0u, uint32_t(n)});
  for (int lvl = 0; lvl != lvl_max - 1 && !next.empty(); ++lvl) {
    const float edge = float(1.0 / std::pow(2, lvl));
    const float half = edge / 2.f;
    level.swap(next);
    next.clear();
    for (const Node& node : level) {
      for (int prim = 0; prim < n; ++prim) {
        if (!ShellHitsBox(&unit_[3 * size_t(prim)], radius, node.c, half + eps)) continue;
        if (node.end - node.begin > kMinNodeSize) Split(node, half, &next);
        else early.emplace_back(node, half + eps);
        break;
      }
    }
  }
  // per primitive the reference sweeps the last level's nodes, then the early leaves: a point's position in that sequence
  uint32_t position = 0;
  for (const Node& node : next)
    for (uint32_t k = node.begin; k < node.end; ++k) (*leaf_position)[ids_[k]] = position++;
  for (const auto& leaf : early)
    for (uint32_t k = leaf.first.begin; k < leaf.first.end; ++k) (*leaf_position)[ids_[k]] = position++;
}

}  // namespace detail
}  // namespace GlobalRegistration
