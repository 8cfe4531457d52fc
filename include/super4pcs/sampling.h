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

namespace detail {
/// Device version of the voxel sampler (libsuper4pcs_b200 -> s4g_voxel_sample): indices of the first
/// point of every voxel, ascending.  Throws std::runtime_error when the GPU path fails.
void GpuVoxelSample(const float* xyz, std::size_t n, float voxel, std::vector<int>& keep);
/// inputs at least this large go to the device (S4PCS_GPU_SAMPLER_MIN overrides; 0 disables)
std::size_t GpuSamplerThreshold();
}  // namespace detail

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
    const std::size_t gpu_min = detail::GpuSamplerThreshold();
    if (gpu_min != 0 && in.size() >= gpu_min) {
      // same voxel arithmetic on the device; identical output (tests/test_sampler_gpu.py)
      std::vector<float> xyz(3 * in.size());
      for (std::size_t i = 0; i < in.size(); ++i) {
        xyz[3 * i] = in[i].x(); xyz[3 * i + 1] = in[i].y(); xyz[3 * i + 2] = in[i].z();
      }
      std::vector<int> keep;
      detail::GpuVoxelSample(xyz.data(), in.size(), options.delta, keep);
      out.reserve(keep.size());
      for (int k : keep) out.push_back(in[std::size_t(k)]);
      return;
    }
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
