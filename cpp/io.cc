// super4pcs-b200: IOManager (host file I/O).
#include "super4pcs/io/io.h"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <locale>
#include <sstream>

using GlobalRegistration::Point3D;
using Vec3 = Point3D::VectorType;

namespace {

std::string Extension(const std::string& f) { return f.size() >= 4 ? f.substr(f.size() - 3) : std::string(); }

std::string WithExtension(const std::string& f, const char* ext) {
  if (f.size() >= 4 && f[f.size() - 4] == '.') return f.substr(0, f.size() - 3) + ext;
  return f + "." + ext;
}

// one "f" record: v, v/t, v//n or v/t/n per corner
bool ParseFace(const char* line, tripple* t) {
  int idx[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};  // [corner][v, t, n]
  const char* p = line + 1;
  for (int c = 0; c < 3; ++c) {
    while (*p == ' ' || *p == '\t') ++p;
    if (!*p) return false;
    int field = 0;
    while (*p && *p != ' ' && *p != '\t' && *p != '\r' && *p != '\n') {
      if (*p == '/') { ++field; ++p; continue; }
      char* end = nullptr;
      long val = std::strtol(p, &end, 10);
      if (end == p) return false;
      if (field < 3) idx[c][field] = int(val);
      p = end;
    }
  }
  t->a = idx[0][0]; t->b = idx[1][0]; t->c = idx[2][0];
  t->t1 = idx[0][1]; t->t2 = idx[1][1]; t->t3 = idx[2][1];
  t->n1 = idx[0][2]; t->n2 = idx[1][2]; t->n3 = idx[2][2];
  return true;
}

}  // namespace

bool IOManager::ReadObject(const char* name, std::vector<Point3D>& v, std::vector<Eigen::Matrix2f>& tex_coords,
                           std::vector<Vec3>& normals, std::vector<tripple>& tris, std::vector<std::string>& mtls) {
  const std::string ext = Extension(name);
  if (ext == "ply") return ReadPly(name, v, normals);
  if (ext == "obj") return ReadObj(name, v, tex_coords, normals, tris, mtls);
  if (ext == "ptx") return ReadPtx(name, v);
  std::cerr << "Unsupported file format" << std::endl;
  return false;
}

bool IOManager::ReadObj(const char* name, std::vector<Point3D>& v, std::vector<Eigen::Matrix2f>& tex_coords,
                        std::vector<Vec3>& normals, std::vector<tripple>& tris, std::vector<std::string>& mtls) {
  std::ifstream f(name);
  if (!f) return false;
  v.clear();
  tris.clear();
  // Record type = first blank-delimited token of the line.  Two behaviours of the reference's reader
  // (io.cc:148-190) are kept on purpose, because they change what a caller gets from common files:
  //  * a line WITHOUT any token (a blank line, or the empty read after a final newline) is treated as
  //    another record of the previous line's type whose numbers failed to parse -- and numbers that fail
  //    to parse keep their previous values.  A "v"-only file that ends with a newline therefore yields
  //    its last vertex twice (same for a trailing "vn");
  //  * x, y, z are shared between "v" and "vn" records for that purpose.
  // For "vt", "f" and "mtllib" the reference reads uninitialised memory in that situation; such repeats
  // are dropped here.
  enum Kind { kOther, kV, kVt, kVn, kF, kMtllib } kind = kOther;
  float x = 0, y = 0, z = 0;
  std::string line;
  for (bool last = false; !last;) {
    // the reference loops "while (!eof) getline": a final newline leads to one more, empty, read
    const bool got = bool(std::getline(f, line));
    if (!got) line.clear();
    last = !got || f.eof();
    char tok[128];
    const bool has_token = std::sscanf(line.c_str(), "%127s", tok) == 1;
    if (has_token) {
      kind = !std::strcmp(tok, "v") ? kV : !std::strcmp(tok, "vt") ? kVt : !std::strcmp(tok, "vn") ? kVn
           : !std::strcmp(tok, "f") ? kF : !std::strcmp(tok, "mtllib") ? kMtllib : kOther;
    }
    const char* s = line.c_str();
    switch (kind) {
      case kV:
        std::sscanf(s, "%*s %f %f %f", &x, &y, &z);
        v.emplace_back(x, y, z);
        v.back().set_rgb(Vec3::Zero());  // OBJ vertices start with a (black) colour, like the reference
        break;
      case kVn:
        std::sscanf(s, "%*s %f %f %f", &x, &y, &z);
        normals.push_back(Vec3(x, y, z));
        break;
      case kVt:
        if (has_token) {
          Eigen::Matrix2f tc = Eigen::Matrix2f::Zero();
          std::sscanf(s, "%*s %f %f", &tc.coeffRef(0), &tc.coeffRef(1));
          tex_coords.push_back(tc);
        }
        break;
      case kF: {
        tripple t;
        const char* rec = std::strchr(s, 'f');
        if (!has_token || rec == nullptr || !ParseFace(rec, &t)) break;
        tris.push_back(t);
        if (!normals.empty()) {
          const int vi[3] = {t.a, t.b, t.c}, ni[3] = {t.n1, t.n2, t.n3};
          for (int c = 0; c < 3; ++c)
            if (vi[c] >= 1 && size_t(vi[c]) <= v.size() && ni[c] >= 1 && size_t(ni[c]) <= normals.size())
              v[vi[c] - 1].set_normal(normals[ni[c] - 1]);
        }
        break;
      }
      case kMtllib:
        if (has_token) mtls.push_back(line.size() > 7 ? line.substr(7) : std::string());
        break;
      case kOther:
        break;
    }
  }
  if (tris.empty()) {
    if (v.size() == normals.size())
      for (size_t i = 0; i < v.size(); ++i) v[i].set_normal(normals[i]);
  } else if (!normals.empty()) {
    // one normal per vertex, in vertex order
    normals.clear();
    normals.reserve(v.size());
    for (const Point3D& p : v) normals.push_back(p.normal());
  }
  return !v.empty();
}

bool IOManager::ReadPly(const char* name, std::vector<Point3D>& v, std::vector<Vec3>& normals) {
  std::ifstream f(name, std::ios::binary);
  if (!f) return false;
  std::string line, format;
  size_t n_vertices = 0;
  struct Prop { std::string type, name; };
  std::vector<Prop> props;
  bool in_vertex = false, header_ok = false;
  std::getline(f, line);
  if (line.compare(0, 3, "ply") != 0) return false;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    std::istringstream ls(line);
    std::string tok;
    ls >> tok;
    if (tok == "format") ls >> format;
    else if (tok == "element") {
      std::string el;
      size_t cnt;
      ls >> el >> cnt;
      in_vertex = (el == "vertex");
      if (in_vertex) n_vertices = cnt;
    } else if (tok == "property" && in_vertex) {
      Prop p;
      ls >> p.type >> p.name;
      if (p.type == "list") return false;
      props.push_back(p);
    } else if (tok == "end_header") { header_ok = true; break; }
  }
  if (!header_ok || n_vertices == 0) return false;
  auto size_of = [](const std::string& t) -> int {
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
    if (t == "double" || t == "float64") return 8;
    return 0;
  };
  v.clear();
  normals.clear();
  v.reserve(n_vertices);
  const bool ascii = format == "ascii";
  const bool big_endian = format == "binary_big_endian";
  if (!ascii && !big_endian && format != "binary_little_endian") return false;
  for (size_t i = 0; i < n_vertices; ++i) {
    float px = 0, py = 0, pz = 0, nx = 0, ny = 0, nz = 0, r = -1, g = -1, b = -1;
    bool has_n = false, has_c = false;
    for (size_t k = 0; k < props.size(); ++k) {
      double d = 0;
      if (ascii) {
        if (!(f >> d)) return false;
      } else {
        char buf[8];
        const int sz = size_of(props[k].type);
        if (sz == 0 || !f.read(buf, sz)) return false;
        if (big_endian)
          for (int lo = 0, hi = sz - 1; lo < hi; ++lo, --hi) std::swap(buf[lo], buf[hi]);
        const std::string& t = props[k].type;
        if (t == "float" || t == "float32") { float x; std::memcpy(&x, buf, 4); d = x; }
        else if (t == "double" || t == "float64") { std::memcpy(&d, buf, 8); }
        else if (sz == 1) d = (t[0] == 'u') ? double(uint8_t(buf[0])) : double(int8_t(buf[0]));
        else if (sz == 2) { int16_t x; std::memcpy(&x, buf, 2); d = (t[0] == 'u') ? double(uint16_t(x)) : double(x); }
        else { int32_t x; std::memcpy(&x, buf, 4); d = (t[0] == 'u') ? double(uint32_t(x)) : double(x); }
      }
      const std::string& nme = props[k].name;
      if (nme == "x") px = float(d); else if (nme == "y") py = float(d); else if (nme == "z") pz = float(d);
      else if (nme == "nx") { nx = float(d); has_n = true; } else if (nme == "ny") ny = float(d);
      else if (nme == "nz") nz = float(d);
      else if (nme == "red") { r = float(d); has_c = true; } else if (nme == "green") g = float(d);
      else if (nme == "blue") b = float(d);
    }
    v.emplace_back(px, py, pz);
    if (has_n) {
      v.back().set_normal(Vec3(nx, ny, nz));
      normals.push_back(Vec3(nx, ny, nz));
    }
    if (has_c) v.back().set_rgb(Vec3(r, g, b));
  }
  return !v.empty();
}

bool IOManager::WriteObject(const char* name, const std::vector<Point3D>& v,
                            const std::vector<Eigen::Matrix2f>& tex_coords, const std::vector<Vec3>& normals,
                            const std::vector<tripple>& tris, const std::vector<std::string>& mtls) {
  const std::string f(name);
  if (tris.empty()) return WritePly(WithExtension(f, "ply"), v, normals);
  return WriteObj(WithExtension(f, "obj"), v, tex_coords, normals, tris, mtls);
}

// Leica PTX scan: "columns", "rows", eight lines of scanner / cloud matrices, then one
// "x y z intensity r g b" record per line.  Reads exactly columns*rows records or fails.
bool IOManager::ReadPtx(const char* name, std::vector<Point3D>& v) {
  std::ifstream f(name);
  if (!f) {
    std::cerr << "(PTX) error opening file" << std::endl;
    return false;
  }
  std::string line;
  long dims[2] = {0, 0};
  for (long& d : dims) {
    std::getline(f, line);
    std::istringstream(line) >> d;
  }
  for (int skip = 0; skip < 8; ++skip) std::getline(f, line);
  const long expected = dims[0] * dims[1];
  v.clear();
  if (expected > 0) v.reserve(size_t(std::min(expected, 1L << 24)));  // (a hint only: the header is not trusted)
  // a field that fails to parse keeps the value of the previous record (stream semantics the
  // reference relies on for short lines), hence the state outside the loop
  Point3D rec;
  float intensity = 0.f;
  Vec3 rgb;
  for (long i = 0; i < expected && !f.eof(); ++i) {
    std::getline(f, line);
    std::istringstream ls(line);
    ls >> rec.x();
    ls >> rec.y();
    ls >> rec.z();
    ls >> intensity;
    ls >> rgb[0];
    ls >> rgb[1];
    ls >> rgb[2];
    rec.set_rgb(rgb);
    v.push_back(rec);
  }
  return long(v.size()) == expected;
}

// Output formats are those of the reference byte for byte (io.cc:328-458), so that files written
// after a switch-over are interchangeable with files written before it:
//  * PLY: binary little endian, floats x y z [nx ny nz] then uchar r g b when any point "has a
//    colour" (Point3D::hasColor(): squared norm of rgb > 0.001 -- true for the default rgb of -1,
//    which is written as 255);
//  * OBJ: stream default precision (6 significant digits), "v x y z " with a trailing blank, the
//    colour appended when its first channel is non-zero, faces as "a/t" when there are texture
//    coordinates and as "a/n" (single slash -- a quirk kept for compatibility) when there are
//    only normals.
bool IOManager::WritePly(const std::string& name, const std::vector<Point3D>& v, const std::vector<Vec3>& normals) {
  std::ofstream f(name, std::ios::out | std::ios::trunc | std::ios::binary);
  if (!f.is_open()) {
    std::cerr << "Cannot open file to write!" << std::endl;
    return false;
  }
  const bool with_normals = normals.size() == v.size();
  const bool with_color = std::any_of(v.begin(), v.end(), [](const Point3D& p) { return p.hasColor(); });
  f.imbue(std::locale::classic());
  f << "ply\nformat binary_little_endian 1.0\ncomment Super4PCS output file\nelement vertex " << v.size() << "\n"
    << "property float x\nproperty float y\nproperty float z\n";
  if (with_normals) f << "property float nx\nproperty float ny\nproperty float nz\n";
  if (with_color) f << "property uchar red\nproperty uchar green\nproperty uchar blue\n";
  f << "end_header\n";
  std::vector<char> rec;
  rec.reserve(27);
  auto put_float = [&rec](float x) {
    char b[4];
    std::memcpy(b, &x, 4);
    rec.insert(rec.end(), b, b + 4);
  };
  for (size_t i = 0; i < v.size(); ++i) {
    rec.clear();
    put_float(v[i].x());
    put_float(v[i].y());
    put_float(v[i].z());
    if (with_normals)
      for (int k = 0; k < 3; ++k) put_float(normals[i][k]);
    if (with_color)
      for (int k = 0; k < 3; ++k) rec.push_back(static_cast<char>(static_cast<int>(v[i].rgb()[k])));
    f.write(rec.data(), std::streamsize(rec.size()));
  }
  f.close();
  return true;
}

bool IOManager::WriteObj(const std::string& name, const std::vector<Point3D>& v,
                         const std::vector<Eigen::Matrix2f>& tex_coords, const std::vector<Vec3>& normals,
                         const std::vector<tripple>& tris, const std::vector<std::string>& mtls) {
  std::ofstream f(name);
  if (!f) return false;
  for (const std::string& m : mtls) f << "mtllib " << m << "\n";
  for (const Point3D& p : v) {
    f << "v " << p.x() << " " << p.y() << " " << p.z() << " ";
    if (p.rgb()[0] != 0) f << p.rgb()[0] << " " << p.rgb()[1] << " " << p.rgb()[2];
    f << "\n";
  }
  for (const Vec3& n : normals) f << "vn " << n[0] << " " << n[1] << " " << n[2] << "\n";
  for (const Eigen::Matrix2f& t : tex_coords) f << "vt " << t.coeff(0) << " " << t.coeff(1) << "\n";
  const bool with_tex = !tex_coords.empty(), plain = normals.empty() && tex_coords.empty();
  for (const tripple& t : tris) {
    if (plain) f << "f " << t.a << " " << t.b << " " << t.c << "\n";
    else if (with_tex) f << "f " << t.a << "/" << t.t1 << " " << t.b << "/" << t.t2 << " " << t.c << "/" << t.t3 << "\n";
    else f << "f " << t.a << "/" << t.n1 << " " << t.b << "/" << t.n2 << " " << t.c << "/" << t.n3 << "\n";
  }
  f.close();
  return true;
}

// Polyworks text matrix: VERSION / MATRIX header, 4 rows, values padded with a blank when >= 0
bool IOManager::WriteMatrix(const std::string& name, const Eigen::Ref<const Eigen::Matrix<double, 4, 4> >& mat,
                            MATRIX_MODE mode) {
  if (mode != POLYWORKS) return false;
  std::ofstream f(name, std::ofstream::out | std::ofstream::trunc);
  if (!f) return false;
  f << "VERSION\t=\t1\nMATRIX\t=\n";
  for (int r = 0; r < 4; ++r) {
    for (int c = 0; c < 4; ++c) {
      const double x = mat(r, c);
      f << (x >= 0. ? " " : "") << std::to_string(x) << (c < 3 ? "  " : "\n");
    }
  }
  return bool(f);
}
