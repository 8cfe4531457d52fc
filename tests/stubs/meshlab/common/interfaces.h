// TEST INFRASTRUCTURE: the handful of MeshLab / Qt / vcglib names the reference's MeshLab plugin
// (demos/MeshlabPlugin/filter_globalregistration/globalregistration.{h,cpp}) touches, so that the plugin can be
// compiled UNCHANGED against the product's headers in an image without MeshLab, Qt or vcglib
// (tests/test_meshlab_plugin.py).  Nothing here is shipped.
#ifndef S4_TEST_STUB_MESHLAB_INTERFACES_H_
#define S4_TEST_STUB_MESHLAB_INTERFACES_H_

#include <cassert>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include <Eigen/Core>

#define Q_OBJECT
#define Q_INTERFACES(x)
#define MESH_FILTER_INTERFACE_IID "stub"
#define MESHLAB_PLUGIN_IID_EXPORTER(x)
#define MESHLAB_PLUGIN_NAME_EXPORTER(x)
#define foreach(decl, container) for (decl : container)

class QObject {
 public:
  virtual ~QObject() {}
};

class QString {
 public:
  QString() {}
  QString(const char* s) : s_(s) {}
  const std::string& str() const { return s_; }

 private:
  std::string s_;
};

class QAction {
 public:
  QAction(const QString& text, QObject*) : text_(text) {}
  const QString& text() const { return text_; }

 private:
  QString text_;
};

template <typename T>
class QList : public std::vector<T> {
 public:
  QList& operator<<(const T& v) {
    this->push_back(v);
    return *this;
  }
};

namespace vcg {
typedef bool CallBackPos(const int, const char*);

struct Point3f {
  float v[3];
  template <typename EigenVector>
  void ToEigenVector(EigenVector& out) const {
    for (int k = 0; k < 3; ++k) out[k] = v[k];
  }
};

struct Matrix44f {
  float m[4][4];
  template <typename EigenMatrix>
  void FromEigenMatrix(const EigenMatrix& e) {
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) m[r][c] = float(e(r, c));
  }
};
}  // namespace vcg

struct CVertexO {
  vcg::Point3f p;
  const vcg::Point3f& P() const { return p; }
};

struct CMeshO {
  std::vector<CVertexO> vert;
  vcg::Matrix44f Tr;
};

class MeshModel {
 public:
  enum { MM_VERTCOORD = 1 };
  CMeshO cm;
};

class MeshDocument {
 public:
  MeshModel* mm() { return current; }
  MeshModel* current = nullptr;
};

// parameters: the plugin declares them with Rich* objects and reads them back by name
struct RichParameter {
  std::string name;
  double number = 0;
  MeshModel* mesh = nullptr;
  virtual ~RichParameter() {}
};
struct RichMesh : RichParameter {
  RichMesh(const char* n, MeshModel* m, MeshDocument*, const char*, const char*) { name = n; mesh = m; }
};
struct RichAbsPerc : RichParameter {
  RichAbsPerc(const char* n, float def, float, float, const char*, const char*) { name = n; number = def; }
};
struct RichFloat : RichParameter {
  RichFloat(const char* n, float def, const char*, const char*) { name = n; number = def; }
};
struct RichInt : RichParameter {
  RichInt(const char* n, int def, const char*, const char*) { name = n; number = def; }
};
struct RichBool : RichParameter {
  RichBool(const char* n, bool def, const char*, const char*) { name = n; number = def ? 1 : 0; }
};

class RichParameterSet {
 public:
  ~RichParameterSet() {
    for (auto& kv : params_) delete kv.second;
  }
  void addParam(RichParameter* p) {
    delete params_[p->name];
    params_[p->name] = p;
  }
  RichParameter& at(const std::string& n) { return *params_.at(n); }
  MeshModel* getMesh(const char* n) { return at(n).mesh; }
  float getAbsPerc(const char* n) { return float(at(n).number); }
  float getFloat(const char* n) { return float(at(n).number); }
  int getInt(const char* n) { return int(at(n).number); }
  bool getBool(const char* n) { return at(n).number != 0; }

 private:
  std::map<std::string, RichParameter*> params_;
};

class MeshFilterInterface {
 public:
  typedef int FilterIDType;
  enum FilterClass { Generic, PointSet };
  enum FILTER_ARITY { SINGLE_MESH };
  virtual ~MeshFilterInterface() {
    for (QAction* a : actionList) delete a;
  }
  QList<FilterIDType> types() const { return typeList; }
  FilterIDType ID(QAction* a) const {
    for (size_t i = 0; i < actionList.size(); ++i)
      if (actionList[i] == a) return typeList[i];
    assert(0);
    return -1;
  }
  void Log(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    std::printf("[meshlab-log] ");
    std::vprintf(fmt, ap);
    std::printf("\n");
    va_end(ap);
  }
  QList<FilterIDType> typeList;
  QList<QAction*> actionList;
};

#endif  // S4_TEST_STUB_MESHLAB_INTERFACES_H_
