// super4pcs-b200 -- file I/O shim.
//
// The reference's demo (demos/Super4PCS/super4pcs_test.cc) talks to a class called IOManager and a
// triangle record called `tripple` (reference src/super4pcs/io/io.h).  This header offers both with
// call-compatible signatures so that the demo builds untouched; the implementation (cpp/io.cc) is
// new: OBJ, PLY (ascii / binary little or big endian) and PTX in; PLY / OBJ / Polyworks matrix out, in
// the reference's exact file formats (tests/test_io_cpu.py compares both byte for byte).
// Not supported: texture look-ups through OpenCV (optional in the reference, off by default).
#ifndef SUPER4PCS_B200_IO_IO_H_
#define SUPER4PCS_B200_IO_IO_H_

#include <string>
#include <vector>

#include <Eigen/Core>

#include "super4pcs/shared4pcs.h"

/// One triangle in OBJ conventions (indices are 1-based): corners a b c, their normal ids n1 n2 n3
/// and texture-coordinate ids t1 t2 t3.
struct tripple {
  int a = 0, b = 0, c = 0;
  int n1 = 0, n2 = 0, n3 = 0;
  int t1 = 0, t2 = 0, t3 = 0;
  tripple() = default;
  tripple(int corner_a, int corner_b, int corner_c) : a(corner_a), b(corner_b), c(corner_c) {}
};

class IOManager {
  using Cloud = std::vector<GlobalRegistration::Point3D>;
  using Normals = std::vector<typename GlobalRegistration::Point3D::VectorType>;
  using TexCoords = std::vector<Eigen::Matrix2f>;
  using Faces = std::vector<tripple>;
  using Names = std::vector<std::string>;
  using Mat4d = Eigen::Matrix<double, 4, 4>;

 public:
  enum MATRIX_MODE { POLYWORKS };

  /// Loads `path` (.obj, .ply or .ptx, chosen by the extension).  False when the file cannot be read,
  /// holds no vertex or has an unsupported extension.
  bool ReadObject(const char* path, Cloud& vertices, TexCoords& tex, Normals& normals, Faces& faces, Names& mtllibs);

  /// Saves a cloud: PLY when `faces` is empty, OBJ otherwise; the extension of `path` is replaced
  /// (or appended when there is none).
  bool WriteObject(const char* path, const Cloud& vertices, const TexCoords& tex, const Normals& normals,
                   const Faces& faces, const Names& mtllibs);

  /// Saves a 4x4 matrix in the Polyworks text layout.
  bool WriteMatrix(const std::string& path, const Eigen::Ref<const Mat4d>& matrix, MATRIX_MODE mode);

 private:
  bool ReadObj(const char* path, Cloud& vertices, TexCoords& tex, Normals& normals, Faces& faces, Names& mtllibs);
  bool ReadPly(const char* path, Cloud& vertices, Normals& normals);
  bool ReadPtx(const char* path, Cloud& vertices);
  bool WriteObj(const std::string& path, const Cloud& vertices, const TexCoords& tex, const Normals& normals,
                const Faces& faces, const Names& mtllibs);
  bool WritePly(const std::string& path, const Cloud& vertices, const Normals& normals);
};

#endif  // SUPER4PCS_B200_IO_IO_H_
