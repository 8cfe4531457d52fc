// super4pcs-b200: IOManager -- OBJ / PLY reading, PLY / OBJ / matrix writing, with the interface
// of the reference's src/super4pcs/io/io.h (ReadObject, WriteObject, WriteMatrix; `tripple`) so
// that the reference's demo (demos/Super4PCS/super4pcs_test.cc) compiles unchanged.  Host-side
// file I/O, not part of the accelerated path.  PTX scans and texture
// look-ups (OpenCV) of the reference are not supported.
#ifndef SUPER4PCS_B200_IO_IO_H_
#define SUPER4PCS_B200_IO_IO_H_

#include <string>
#include <vector>

#include <Eigen/Core>

#include "super4pcs/shared4pcs.h"

/// triangle: vertex ids a,b,c (1-based, OBJ convention) + normal ids n* + texture ids t*
struct tripple {
  int a, b, c;
  int n1, n2, n3;
  int t1, t2, t3;
  tripple() : a(0), b(0), c(0), n1(0), n2(0), n3(0), t1(0), t2(0), t3(0) {}
  tripple(int a_, int b_, int c_) : a(a_), b(b_), c(c_), n1(0), n2(0), n3(0), t1(0), t2(0), t3(0) {}
};

class IOManager {
 public:
  enum MATRIX_MODE { POLYWORKS };

  /// dispatches on the extension (.obj, .ply); false on failure / unsupported format
  bool ReadObject(const char* name, std::vector<GlobalRegistration::Point3D>& v,
                  std::vector<Eigen::Matrix2f>& tex_coords,
                  std::vector<typename GlobalRegistration::Point3D::VectorType>& normals,
                  std::vector<tripple>& tris, std::vector<std::string>& mtls);
  /// writes a .ply when there are no triangles, a .obj otherwise (extension replaced / appended)
  bool WriteObject(const char* name, const std::vector<GlobalRegistration::Point3D>& v,
                   const std::vector<Eigen::Matrix2f>& tex_coords,
                   const std::vector<typename GlobalRegistration::Point3D::VectorType>& normals,
                   const std::vector<tripple>& tris, const std::vector<std::string>& mtls);
  bool WriteMatrix(const std::string& name, const Eigen::Ref<const Eigen::Matrix<double, 4, 4> >& mat,
                   MATRIX_MODE mode);

 private:
  bool ReadPly(const char* name, std::vector<GlobalRegistration::Point3D>& v,
               std::vector<typename GlobalRegistration::Point3D::VectorType>& normals);
  bool ReadObj(const char* name, std::vector<GlobalRegistration::Point3D>& v,
               std::vector<Eigen::Matrix2f>& tex_coords,
               std::vector<typename GlobalRegistration::Point3D::VectorType>& normals,
               std::vector<tripple>& tris, std::vector<std::string>& mtls);
  bool WritePly(const std::string& name, const std::vector<GlobalRegistration::Point3D>& v,
                const std::vector<typename GlobalRegistration::Point3D::VectorType>& normals);
  bool WriteObj(const std::string& name, const std::vector<GlobalRegistration::Point3D>& v,
                const std::vector<Eigen::Matrix2f>& tex_coords,
                const std::vector<typename GlobalRegistration::Point3D::VectorType>& normals,
                const std::vector<tripple>& tris, const std::vector<std::string>& mtls);
};

#endif  // SUPER4PCS_B200_IO_IO_H_
