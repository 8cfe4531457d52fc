// Test driver (ours) for the REFERENCE's PCL wrapper class pcl::Super4PCS<>, which is included unchanged
// from <reference>/demos/PCLWrapper and compiled against the product's headers with the PCL stub of
// tests/stubs/.  Usage: pcl_wrapper_test a.xyz b.xyz overlap delta n_points  ->  prints "Score-matrix:" rows.
#include <pcl/registration/super4pcs.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>

static bool load(const char* path, pcl::PointCloud<pcl::PointXYZ>& c) {
  std::ifstream f(path);
  pcl::PointXYZ p;
  while (f >> p.x >> p.y >> p.z) c.push_back(p);
  return c.size() > 0;
}

int main(int argc, char** argv) {
  if (argc < 6) return 2;
  typedef pcl::PointCloud<pcl::PointXYZ> Cloud;
  Cloud::Ptr scene(new Cloud), object(new Cloud), aligned(new Cloud);
  if (!load(argv[1], *scene) || !load(argv[2], *object)) return 3;
  pcl::Super4PCS<pcl::PointXYZ, pcl::PointXYZ> align;
  if (!align.options_.configureOverlap(float(std::atof(argv[3])))) return 4;
  align.options_.delta = float(std::atof(argv[4]));
  align.options_.sample_size = std::size_t(std::atoi(argv[5]));
  align.options_.max_time_seconds = 1000;
  align.setInputSource(object);
  align.setInputTarget(scene);
  align.align(*aligned);
  if (!align.hasConverged() || aligned->size() != object->size()) return 5;
  const Eigen::Matrix4f T = align.getFinalTransformation();
  std::printf("\n");
  for (int r = 0; r < 4; ++r) std::printf("Score-matrix: %.9g %.9g %.9g %.9g\n", T(r, 0), T(r, 1), T(r, 2), T(r, 3));
  std::printf("first-aligned: %.9g %.9g %.9g\n", (*aligned)[0].x, (*aligned)[0].y, (*aligned)[0].z);
  return 0;
}
