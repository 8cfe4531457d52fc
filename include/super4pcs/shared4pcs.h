// super4pcs-b200: public value types of the GlobalRegistration API.
//
// Header-compatible replacement of the reference's src/super4pcs/shared4pcs.h: the same names,
// members, defaults and semantics (Point3D shared4pcs.h:61-111, Quadrilateral :116-138,
// Match4PCSOptions :148-190) so that callers written against the reference compile unchanged.
// Written from the interface description in SURVEY.md 8(b); no reference code is reused.
#ifndef SUPER4PCS_B200_SHARED4PCS_H_
#define SUPER4PCS_B200_SHARED4PCS_H_

#include <Eigen/Core>

#include <array>
#include <cstddef>
#include <fstream>
#include <iostream>
#include <random>
#include <vector>

namespace GlobalRegistration {

/// Four point indices; ordered lexicographically.
struct Quadrilateral {
  std::array<int, 4> vertices;

  Quadrilateral(int v0, int v1, int v2, int v3) : vertices{{v0, v1, v2, v3}} {}

  bool operator<(const Quadrilateral& o) const { return vertices < o.vertices; }
  bool operator==(const Quadrilateral& o) const { return vertices == o.vertices; }
  int operator[](int i) const { return vertices[i]; }
  int& operator[](int i) { return vertices[i]; }
};

inline std::ofstream& operator<<(std::ofstream& os, const Quadrilateral& q) {
  os << "[" << q[0] << " " << q[1] << " " << q[2] << " " << q[3] << "]";
  return os;
}

/// A 3D sample: position, (optional) unit normal, (optional) colour.
/// Defaults: position 0, normal 0 ("no normal"), rgb -1 ("no colour").
class Point3D {
 public:
  using Scalar = float;
  using VectorType = Eigen::Matrix<Scalar, 3, 1>;

  Point3D() = default;
  Point3D(const Point3D&) = default;
  Point3D& operator=(const Point3D&) = default;
  Point3D(Scalar px, Scalar py, Scalar pz) : pos_(px, py, pz) {}
  template <typename S>
  explicit Point3D(const Eigen::Matrix<S, 3, 1>& p) : pos_(Scalar(p(0)), Scalar(p(1)), Scalar(p(2))) {}

  VectorType& pos() { return pos_; }
  const VectorType& pos() const { return pos_; }
  const VectorType& normal() const { return normal_; }
  const VectorType& rgb() const { return rgb_; }

  void set_rgb(const VectorType& c) { rgb_ = c; }
  /// stores the normalised vector
  void set_normal(const VectorType& n) { normal_ = n.normalized(); }
  void normalize() { normal_.normalize(); }
  bool hasColor() const { return rgb_.squaredNorm() > Scalar(0.001); }

  Scalar& x() { return pos_.coeffRef(0); }
  Scalar& y() { return pos_.coeffRef(1); }
  Scalar& z() { return pos_.coeffRef(2); }
  Scalar x() const { return pos_.coeff(0); }
  Scalar y() const { return pos_.coeff(1); }
  Scalar z() const { return pos_.coeff(2); }

 private:
  VectorType pos_{Scalar(0), Scalar(0), Scalar(0)};
  VectorType normal_{Scalar(0), Scalar(0), Scalar(0)};
  VectorType rgb_{Scalar(-1), Scalar(-1), Scalar(-1)};
};

/// Algorithm parameters.  delta and the overlap estimate are the application knobs.
struct Match4PCSOptions {
  using Scalar = typename Point3D::Scalar;
  Match4PCSOptions() {}

  // --- geometry
  Scalar delta = Scalar(5.0);  ///< two points closer than this count as matched (the LCP radius)
  // --- optional pair filters; a negative value switches the filter off
  Scalar max_normal_difference = Scalar(-1);     ///< [deg] between the normals of corresponding points
  Scalar max_translation_distance = Scalar(-1);  ///< between corresponding points
  Scalar max_angle = Scalar(-1);                 ///< [deg] of the sought rotation
  Scalar max_color_distance = Scalar(-1);        ///< RGB distance between corresponding points
  // --- search budget
  size_t sample_size = 200;    ///< upper bound on the number of samples drawn from Q
  int max_time_seconds = 60;   ///< the search may be stopped after this long, keeping the best so far
  unsigned int randomSeed = std::mt19937::default_seed;  ///< seed of the matcher's std::mt19937

  /// false (and nothing changes) when the termination threshold is below the overlap
  bool configureOverlap(Scalar overlap, Scalar terminate_thr = Scalar(1)) {
    if (terminate_thr < overlap) return false;
    overlap_estimation = overlap;
    terminate_threshold = terminate_thr;
    return true;
  }
  Scalar getTerminateThreshold() const { return terminate_threshold; }
  Scalar getOverlapEstimation() const { return overlap_estimation; }

 private:
  Scalar terminate_threshold = Scalar(1.0);  ///< stop once the LCP exceeds this
  Scalar overlap_estimation = Scalar(0.2);   ///< expected overlap fraction
};

}  // namespace GlobalRegistration

#endif  // SUPER4PCS_B200_SHARED4PCS_H_
