// super4pcs-b200: Sampling::UniformDistSampler, behaviour of the reference's
// src/super4pcs/sampling.h:59-121: one point per voxel of edge options.delta -- the FIRST point
// of every voxel in input order -- emitted in input order.  The voxel of a point is
// (int(floor(x*s)), int(floor(y*s)), int(floor(z*s))) with s = 1.0f / delta (float), exactly
// as the reference computes it; which hash container finds duplicates does not change the output.
#ifndef SUPER4PCS_B200_SAMPLING_H_
#define SUPER4PCS_B200_SAMPLING_H_

#include <cmath>
#include <cstdint>
#include <unordered_set>
#include <vector>

#include "super4pcs/shared4pcs.h"

namespace GlobalRegistration {
namespace Sampling {

struct UniformDistSampler {
 private:
  struct Voxel {
    int x, y, z;
    bool operator==(const Voxel& o) const { return x == o.x && y == o.y && z == o.z; }
  };
  struct VoxelHash {
    std::size_t operator()(const Voxel& v) const {
      std::uint64_t h = 0x9E3779B97F4A7C15ull;
      for (std::uint64_t c : {std::uint64_t(std::uint32_t(v.x)), std::uint64_t(std::uint32_t(v.y)),
                              std::uint64_t(std::uint32_t(v.z))}) {
        h ^= c + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h *= 0xBF58476D1CE4E5B9ull;
      }
      return std::size_t(h ^ (h >> 31));
    }
  };

 public:
  template <typename Point>
  inline void operator()(const std::vector<Point>& in, const Match4PCSOptions& options,
                         std::vector<Point>& out) const {
    using Scalar = typename Point::Scalar;
    out.clear();
    const Scalar scale = 1.0f / options.delta;
    std::unordered_set<Voxel, VoxelHash> seen;
    seen.reserve(in.size());
    for (const Point& p : in) {
      const Voxel v{int(std::floor(p.x() * scale)), int(std::floor(p.y() * scale)),
                    int(std::floor(p.z() * scale))};
      if (seen.insert(v).second) out.push_back(p);
    }
  }
};

}  // namespace Sampling
}  // namespace GlobalRegistration

#endif  // SUPER4PCS_B200_SAMPLING_H_
