// super4pcs-b200: Utils::CleanInvalidNormals, interface of the reference's
// src/super4pcs/utils/geometry.h:56-82: vertices whose normal is (near) zero lose it, the others
// get unit normals; the parallel normal array is kept in sync.
#ifndef SUPER4PCS_B200_UTILS_GEOMETRY_H_
#define SUPER4PCS_B200_UTILS_GEOMETRY_H_

#include <cstddef>
#include <iostream>

namespace GlobalRegistration {
namespace Utils {

template <typename PointContainer, typename VecContainer>
static inline void CleanInvalidNormals(PointContainer& v, VecContainer& normals) {
  if (v.size() != normals.size()) return;
  using Vector = typename VecContainer::value_type;
  std::size_t dropped = 0;
  for (std::size_t i = 0; i < v.size(); ++i) {
    if (v[i].normal().squaredNorm() < 0.01) {
      normals[i] = Vector::Zero();
      v[i].set_normal(Vector::Zero());
      ++dropped;
    } else {
      normals[i].normalize();
      v[i].normalize();
    }
  }
  if (dropped) std::cout << "Found " << dropped << " vertices with invalid normals" << std::endl;
}

}  // namespace Utils
}  // namespace GlobalRegistration

#endif  // SUPER4PCS_B200_UTILS_GEOMETRY_H_
