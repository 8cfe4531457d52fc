// TEST STUB -- a few dozen lines standing in for the parts of the Point Cloud Library that the
// reference's PCL wrapper (demos/PCLWrapper/pcl/registration/super4pcs.h + impl/super4pcs.hpp) touches.
// PCL is not installed in this image; the stub only exists so that tests/cpp/pcl_wrapper_main.cc can
// compile the reference's wrapper UNCHANGED against the product's headers (compile-compat of the second
// caller named in SURVEY.md 2, row 15).  It makes no attempt to be PCL.
#ifndef S4_TEST_STUB_PCL_REGISTRATION_H_
#define S4_TEST_STUB_PCL_REGISTRATION_H_

#include <Eigen/Core>
#include <cstdarg>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

namespace pcl {

struct PointXYZ { float x, y, z; };

struct PointIndices {
  typedef std::shared_ptr<PointIndices> Ptr;
  typedef std::shared_ptr<const PointIndices> ConstPtr;
  std::vector<int> indices;
};

template <typename PointT>
class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  std::vector<PointT> points;
  std::size_t size() const { return points.size(); }
  const PointT& operator[](std::size_t i) const { return points[i]; }
  PointT& operator[](std::size_t i) { return points[i]; }
  void push_back(const PointT& p) { points.push_back(p); }
};

namespace console {
inline void print_highlight(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  std::vprintf(fmt, ap);
  va_end(ap);
}
}  // namespace console

template <typename PointT>
void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Matrix4f& T) {
  out.points.resize(in.size());
  for (std::size_t i = 0; i < in.size(); ++i) {
    const Eigen::Vector4f p = T * Eigen::Vector4f(in[i].x, in[i].y, in[i].z, 1.f);
    out[i] = in[i];
    out[i].x = p[0]; out[i].y = p[1]; out[i].z = p[2];
  }
}

namespace registration {
template <typename S, typename T>
struct TransformationEstimation { virtual ~TransformationEstimation() {} };
}  // namespace registration

template <typename PointSource, typename PointTarget>
class Registration {
 public:
  typedef Eigen::Matrix4f Matrix4;
  typedef PointCloud<PointSource> PointCloudSource;
  typedef typename PointCloudSource::Ptr PointCloudSourcePtr;
  typedef typename PointCloudSource::ConstPtr PointCloudSourceConstPtr;
  typedef PointCloud<PointTarget> PointCloudTarget;
  typedef typename PointCloudTarget::ConstPtr PointCloudTargetConstPtr;

  Registration() : final_transformation_(Matrix4::Identity()), converged_(false) {}
  virtual ~Registration() {}
  void setInputSource(const PointCloudSourceConstPtr& c) { input_ = c; }
  void setInputTarget(const PointCloudTargetConstPtr& c) { target_ = c; }
  void align(PointCloudSource& output) { computeTransformation(output, Matrix4::Identity()); }
  bool hasConverged() const { return converged_; }
  Matrix4 getFinalTransformation() const { return final_transformation_; }
  const std::string& getClassName() const { return reg_name_; }

 protected:
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;
  std::string reg_name_;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  std::shared_ptr<registration::TransformationEstimation<PointSource, PointTarget>> transformation_estimation_;
  Matrix4 final_transformation_;
  bool converged_;
};

}  // namespace pcl
#endif
