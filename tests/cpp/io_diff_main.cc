// TEST INFRASTRUCTURE: one program, compiled twice -- against the reference's IOManager
// (/root/reference/src/super4pcs/io) and against the product's (include/super4pcs/io/io.h +
// cpp/io.cc) -- so that tests/test_io_cpu.py can compare what the two read from the same files
// and what they write, byte for byte.  Uses only the interface both share.
//   io_diff <input> <dump.txt> <out-path-for-WriteObject> <out.matrix>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <Eigen/Core>

#include "super4pcs/io/io.h"
#include "super4pcs/utils/geometry.h"

using GlobalRegistration::Point3D;

static void hexf(FILE* o, float x) {
  unsigned u;
  std::memcpy(&u, &x, 4);
  std::fprintf(o, " %08x", u);
}

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  std::vector<Point3D> v;
  std::vector<Eigen::Matrix2f> tex;
  std::vector<Point3D::VectorType> normals;
  std::vector<tripple> tris;
  std::vector<std::string> mtls;
  IOManager io;
  const bool ok = io.ReadObject(argv[1], v, tex, normals, tris, mtls);
  FILE* o = std::fopen(argv[2], "w");
  std::fprintf(o, "ok %d v %zu tex %zu normals %zu tris %zu mtls %zu\n", ok ? 1 : 0, v.size(), tex.size(), normals.size(),
               tris.size(), mtls.size());
  for (const Point3D& p : v) {
    std::fprintf(o, "p");
    for (int k = 0; k < 3; ++k) hexf(o, p.pos()[k]);
    for (int k = 0; k < 3; ++k) hexf(o, p.normal()[k]);
    for (int k = 0; k < 3; ++k) hexf(o, p.rgb()[k]);
    std::fprintf(o, " c%d\n", p.hasColor() ? 1 : 0);
  }
  for (const auto& n : normals) {
    std::fprintf(o, "n");
    for (int k = 0; k < 3; ++k) hexf(o, n[k]);
    std::fprintf(o, "\n");
  }
  for (const auto& t : tex) {
    std::fprintf(o, "t");
    hexf(o, t.coeff(0));
    hexf(o, t.coeff(1));
    std::fprintf(o, "\n");
  }
  for (const tripple& t : tris) std::fprintf(o, "f %d %d %d | %d %d %d | %d %d %d\n", t.a, t.b, t.c, t.n1, t.n2, t.n3, t.t1, t.t2, t.t3);
  for (const std::string& m : mtls) std::fprintf(o, "m [%s]\n", m.c_str());
  if (ok) {
    // the demo's sequence: drop invalid normals when there is no mesh, then write
    if (tris.empty()) GlobalRegistration::Utils::CleanInvalidNormals(v, normals);
    std::fprintf(o, "after-clean v %zu normals %zu\n", v.size(), normals.size());
    const bool w = io.WriteObject(argv[3], v, tex, normals, tris, mtls);
    std::fprintf(o, "write %d\n", w ? 1 : 0);
  }
  Eigen::Matrix<double, 4, 4> M;
  M << 0.7399, 0.062655, -0.669793, -0.097583, -0.104949, 0.994213, -0.022932, -0.005567, 0.66448, 0.087262, 0.742194,
      -0.0321, 0, 0, 0, 1;
  std::fprintf(o, "matrix %d\n", io.WriteMatrix(argv[4], M, IOManager::POLYWORKS) ? 1 : 0);
  std::fclose(o);
  return 0;
}
