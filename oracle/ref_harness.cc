// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the
// product (libs4g.so / super4pcs_b200). Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the library this builds.
//
// oracle/_ref/liboracle_ref.so = the UNMODIFIED reference (compiled from the sources
// where they lie under /root/reference, see oracle/Makefile) + this thin C-ABI
// harness, which reaches the protected stages of GlobalRegistration::MatchSuper4PCS
// through a subclass, the sanctioned pattern of the reference's own tests
// (reference tests/testing.h:71-154, Testing::TestMatcher).
//
// No reference source is copied here: this file only #includes the reference's
// public headers at build time.
#include "super4pcs/algorithms/super4pcs.h"
#include "super4pcs/algorithms/4pcs.h"
#include "super4pcs/utils/logger.h"

#include <cstdint>
#include <cstring>
#include <vector>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace GlobalRegistration;

extern "C" {
struct RefOptions {
  float delta;
  float max_normal_difference;
  float max_translation_distance;
  float max_angle;
  float max_color_distance;
  uint64_t sample_size;
  int32_t max_time_seconds;
  uint32_t random_seed;
  float overlap;
  float terminate_threshold;
};
}

namespace {

// keeps every input point, in input order (stage tests want sampled == input)
struct IdentitySampler {
  template <typename Point>
  void operator()(const std::vector<Point>& in, const Match4PCSOptions&,
                  std::vector<Point>& out) const { out = in; }
};

struct CountingVisitor {
  mutable long n_candidate_calls = 0;
  inline void operator()(float fraction, float, Eigen::Ref<Match4PCSBase::MatrixType>) const {
    if (fraction < 0) ++n_candidate_calls;
  }
  constexpr bool needsGlobalTransformation() const { return false; }
};

class Probe : public MatchSuper4PCS {
 public:
  Probe(const Match4PCSOptions& o, const Utils::Logger& l) : MatchSuper4PCS(o, l) {}
  using MatchSuper4PCS::ExtractPairs;
  using MatchSuper4PCS::FindCongruentQuadrilaterals;
  using Match4PCSBase::init;
  using Match4PCSBase::SelectQuadrilateral;
  using Match4PCSBase::ComputeRigidTransformation;
  using Match4PCSBase::Verify;
  using Match4PCSBase::TryCongruentSet;
  using Match4PCSBase::TryOneBase;
  using Match4PCSBase::Perform_N_steps;
  using Match4PCSBase::best_LCP_;
  using Match4PCSBase::base_3D_;
  using Match4PCSBase::base_;
  using Match4PCSBase::current_congruent_;
  using Match4PCSBase::transform_;
  using Match4PCSBase::sampled_P_3D_;
  using Match4PCSBase::sampled_Q_3D_;
  using Match4PCSBase::number_of_trials_;
  using Match4PCSBase::P_diameter_;
  using Match4PCSBase::max_base_diameter_;
  using Match4PCSBase::centroid_P_;
  using Match4PCSBase::centroid_Q_;
  using Match4PCSBase::qcentroid1_;
  using Match4PCSBase::qcentroid2_;
};

struct Handle {
  Utils::Logger logger{Utils::NoLog};
  Match4PCSOptions opt;
  Probe* m = nullptr;
  std::vector<Point3D> P, Q;
  std::vector<std::pair<int, int>> pairs;
  std::vector<Quadrilateral> quads;
  ~Handle() { delete m; }
};

Match4PCSOptions to_options(const RefOptions* o) {
  Match4PCSOptions r;
  r.delta = o->delta;
  r.max_normal_difference = o->max_normal_difference;
  r.max_translation_distance = o->max_translation_distance;
  r.max_angle = o->max_angle;
  r.max_color_distance = o->max_color_distance;
  r.sample_size = o->sample_size;
  r.max_time_seconds = o->max_time_seconds;
  r.randomSeed = o->random_seed;
  r.configureOverlap(o->overlap, o->terminate_threshold);
  return r;
}

void fill_cloud(std::vector<Point3D>& c, const float* xyz, const float* nrm,
                const float* rgb, int n) {
  c.clear();
  c.reserve(n);
  for (int i = 0; i < n; ++i) {
    Point3D p(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    if (nrm) {
      // stored verbatim (callers pass already-normalised or zero normals);
      // set_normal() would normalise and turn zero normals into NaN
      Point3D::VectorType v(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
      if (v.squaredNorm() > 0) p.set_normal(v);
    }
    if (rgb) p.set_rgb(Point3D::VectorType(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]));
    c.push_back(p);
  }
}

}  // namespace

extern "C" {

int ref_abi_version() { return 1; }

int ref_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// Creates a matcher and runs the reference's init() (sampling, centring, kd-tree,
// diameter, initial LCP). use_identity_sampler!=0 keeps every point.
void* ref_create(const float* Pxyz, const float* Pnrm, const float* Prgb, int nP,
                 const float* Qxyz, const float* Qnrm, const float* Qrgb, int nQ,
                 const RefOptions* o, int use_identity_sampler) {
  Handle* h = new Handle;
  h->opt = to_options(o);
  fill_cloud(h->P, Pxyz, Pnrm, Prgb, nP);
  fill_cloud(h->Q, Qxyz, Qnrm, Qrgb, nQ);
  h->m = new Probe(h->opt, h->logger);
  if (use_identity_sampler)
    h->m->init(h->P, h->Q, IdentitySampler());
  else
    h->m->init(h->P, h->Q, Sampling::UniformDistSampler());
  return h;
}

void ref_destroy(void* hv) { delete static_cast<Handle*>(hv); }

int ref_n_sampled_p(void* hv) { return (int)static_cast<Handle*>(hv)->m->sampled_P_3D_.size(); }
int ref_n_sampled_q(void* hv) { return (int)static_cast<Handle*>(hv)->m->sampled_Q_3D_.size(); }

static void dump_cloud(const std::vector<Point3D>& c, float* xyz, float* nrm, float* rgb) {
  for (size_t i = 0; i < c.size(); ++i)
    for (int k = 0; k < 3; ++k) {
      if (xyz) xyz[3 * i + k] = c[i].pos()[k];
      if (nrm) nrm[3 * i + k] = c[i].normal()[k];
      if (rgb) rgb[3 * i + k] = c[i].rgb()[k];
    }
}
// centred sampled clouds (what every later stage works on)
void ref_get_sampled_p(void* hv, float* xyz, float* nrm, float* rgb) {
  dump_cloud(static_cast<Handle*>(hv)->m->sampled_P_3D_, xyz, nrm, rgb);
}
void ref_get_sampled_q(void* hv, float* xyz, float* nrm, float* rgb) {
  dump_cloud(static_cast<Handle*>(hv)->m->sampled_Q_3D_, xyz, nrm, rgb);
}

// scalars of init(): [0]=best_LCP_ (initial), [1]=P_diameter_, [2]=max_base_diameter_,
// [3]=number_of_trials_, [4..6]=centroid_P, [7..9]=centroid_Q
void ref_get_init_state(void* hv, float* out10) {
  Probe* m = static_cast<Handle*>(hv)->m;
  out10[0] = m->best_LCP_;
  out10[1] = m->P_diameter_;
  out10[2] = m->max_base_diameter_;
  out10[3] = (float)m->number_of_trials_;
  for (int k = 0; k < 3; ++k) { out10[4 + k] = m->centroid_P_[k]; out10[7 + k] = m->centroid_Q_[k]; }
}

void ref_set_best_lcp(void* hv, float v) { static_cast<Handle*>(hv)->m->best_LCP_ = v; }
float ref_get_best_lcp(void* hv) { return static_cast<Handle*>(hv)->m->best_LCP_; }

// base_3D_ <- four explicit points (pos, normal, rgb each 3 floats; nrm/rgb may be null)
void ref_set_base3d(void* hv, const float* xyz, const float* nrm, const float* rgb) {
  Probe* m = static_cast<Handle*>(hv)->m;
  std::vector<Point3D> b;
  fill_cloud(b, xyz, nrm, rgb, 4);
  m->base_3D_ = b;
}
void ref_get_base3d(void* hv, float* xyz, float* nrm, float* rgb) {
  dump_cloud(static_cast<Handle*>(hv)->m->base_3D_, xyz, nrm, rgb);
}

// Reference base selection (consumes the member RNG). Returns 1 on success.
int ref_select_quadrilateral(void* hv, float* inv1, float* inv2, int* ids4) {
  Probe* m = static_cast<Handle*>(hv)->m;
  float a = 0, b = 0;
  bool ok = m->SelectQuadrilateral(a, b, ids4[0], ids4[1], ids4[2], ids4[3]);
  *inv1 = a; *inv2 = b;
  return ok ? 1 : 0;
}

// MatchSuper4PCS::ExtractPairs; result kept in the handle, returns the count.
// Order is the reference's emission order (not canonical).
long ref_extract_pairs(void* hv, float d, float normal_angle, float eps, int b1, int b2) {
  Handle* h = static_cast<Handle*>(hv);
  h->m->ExtractPairs(d, normal_angle, eps, b1, b2, &h->pairs);
  return (long)h->pairs.size();
}
void ref_get_pairs(void* hv, int* out2n) {
  Handle* h = static_cast<Handle*>(hv);
  for (size_t i = 0; i < h->pairs.size(); ++i) {
    out2n[2 * i] = h->pairs[i].first;
    out2n[2 * i + 1] = h->pairs[i].second;
  }
}

// MatchSuper4PCS::FindCongruentQuadrilaterals on explicit pair lists.
long ref_find_quads(void* hv, float inv1, float inv2, float thr1, float thr2,
                    const int* pairs1, long n1, const int* pairs2, long n2) {
  Handle* h = static_cast<Handle*>(hv);
  std::vector<std::pair<int, int>> p1(n1), p2(n2);
  for (long i = 0; i < n1; ++i) p1[i] = {pairs1[2 * i], pairs1[2 * i + 1]};
  for (long i = 0; i < n2; ++i) p2[i] = {pairs2[2 * i], pairs2[2 * i + 1]};
  h->m->FindCongruentQuadrilaterals(inv1, inv2, thr1, thr2, p1, p2, &h->quads);
  return (long)h->quads.size();
}
void ref_get_quads(void* hv, int* out4n) {
  Handle* h = static_cast<Handle*>(hv);
  for (size_t i = 0; i < h->quads.size(); ++i)
    for (int k = 0; k < 4; ++k) out4n[4 * i + k] = h->quads[i].vertices[k];
}

// Match4PCSBase::ComputeRigidTransformation for K candidate quads of sampled_Q
// against base ids (into sampled_P), exactly as TryCongruentSet prepares its
// arguments (match4pcsBase.hpp:373-434). out_T: K x 16 floats, COLUMN-major
// (Eigen default); out_rms: K; out_ok: K (the bool the function returned).
void ref_rigid_batch(void* hv, const int* base_ids4, const int* quads4k, long K,
                     float* out_T, float* out_rms, int* out_ok) {
  Handle* h = static_cast<Handle*>(hv);
  Probe* m = h->m;
  static const double pi = std::acos(-1);
  const Point3D& b1 = m->sampled_P_3D_[base_ids4[0]];
  const Point3D& b2 = m->sampled_P_3D_[base_ids4[1]];
  const Point3D& b3 = m->sampled_P_3D_[base_ids4[2]];
  const Point3D& b4 = m->sampled_P_3D_[base_ids4[3]];
  const std::array<Point3D, 4> congruent_base{{b1, b2, b3, b4}};
  Eigen::Matrix<float, 3, 1> centroid1 = (b1.pos() + b2.pos() + b3.pos()) / float(3);
  for (long i = 0; i < K; ++i) {
    std::array<Point3D, 4> cand;
    for (int k = 0; k < 4; ++k) cand[k] = m->sampled_Q_3D_[quads4k[4 * i + k]];
    Eigen::Matrix<float, 3, 1> centroid2 =
        (cand[0].pos() + cand[1].pos() + cand[2].pos()) / float(3.);
    Eigen::Matrix<float, 4, 4> T = Eigen::Matrix<float, 4, 4>::Zero();
    float rms = -1;
    bool ok = m->ComputeRigidTransformation(congruent_base, cand, centroid1, centroid2,
                                            h->opt.max_angle * pi / 180.0, T, rms, false);
    std::memcpy(out_T + 16 * i, T.data(), 16 * sizeof(float));
    out_rms[i] = rms;
    out_ok[i] = ok ? 1 : 0;
  }
}

// Match4PCSBase::Verify for K transforms (column-major 4x4 each). best_lcp is
// written to best_LCP_ before every call (0 disables the early exit). With
// nthreads>1 candidates are spread over OpenMP threads -- the reference's own
// parallelisation of this loop (match4pcsBase.hpp:390-393); Verify is const.
// Returns elapsed seconds.
double ref_verify_batch(void* hv, const float* T16k, long K, float best_lcp,
                        int nthreads, float* out_lcp) {
  Probe* m = static_cast<Handle*>(hv)->m;
  m->best_LCP_ = best_lcp;
  auto t0 = std::chrono::steady_clock::now();
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(dynamic, 1)
#endif
  for (long i = 0; i < K; ++i) {
    Eigen::Matrix<float, 4, 4> T;
    std::memcpy(T.data(), T16k + 16 * i, 16 * sizeof(float));
    out_lcp[i] = m->Verify(T);
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// Match4PCSBase::TryCongruentSet as shipped (rigid fit + gate + Verify with early
// exit + first-max rule). Outputs the matcher state afterwards.
// out_state: [0]=best_LCP_, [1]=nbCongruent (gate-passing quads), [2]=visitor calls
// out_T: transform_ (column-major 16), out_ids: base_[4], current_congruent_[4]
int ref_try_congruent_set(void* hv, const int* base_ids4, const int* quads4k, long K,
                          float* out_state3, float* out_T, int* out_ids8) {
  Handle* h = static_cast<Handle*>(hv);
  Probe* m = h->m;
  std::vector<Quadrilateral> quads;
  quads.reserve(K);
  for (long i = 0; i < K; ++i)
    quads.emplace_back(quads4k[4 * i], quads4k[4 * i + 1], quads4k[4 * i + 2], quads4k[4 * i + 3]);
  size_t nb = 0;
  CountingVisitor v;
  bool r = m->TryCongruentSet(base_ids4[0], base_ids4[1], base_ids4[2], base_ids4[3], quads, v, nb);
  out_state3[0] = m->best_LCP_;
  out_state3[1] = (float)nb;
  out_state3[2] = (float)v.n_candidate_calls;
  std::memcpy(out_T, m->transform_.data(), 16 * sizeof(float));
  for (int k = 0; k < 4; ++k) { out_ids8[k] = m->base_[k]; out_ids8[4 + k] = m->current_congruent_[k]; }
  return r ? 1 : 0;
}

// n RANSAC steps on an initialised matcher (Match4PCSBase::Perform_N_steps, hpp:208-274; the
// Meshlab plugin's stepwise usage).  out_state3 = (best_LCP, #progress reports, #candidate reports);
// the matcher stays alive, so its RNG / base state can be probed afterwards.
int ref_perform_n_steps(void* hv, int n, float* out_state3, float* out_T16_colmajor) {
  Handle* h = static_cast<Handle*>(hv);
  struct V {
    mutable long progress = 0, candidates = 0;
    inline void operator()(float fraction, float, Eigen::Ref<Match4PCSBase::MatrixType>) const {
      if (fraction < 0) ++candidates; else ++progress;
    }
    constexpr bool needsGlobalTransformation() const { return false; }
  } v;
  std::vector<Point3D> Q = h->Q;
  Match4PCSBase::MatrixType mat(Match4PCSBase::MatrixType::Identity());
  const bool r = h->m->Perform_N_steps(n, mat, &Q, v);
  out_state3[0] = h->m->best_LCP_;
  out_state3[1] = (float)v.progress;
  out_state3[2] = (float)v.candidates;
  std::memcpy(out_T16_colmajor, mat.data(), 16 * sizeof(float));
  return r ? 1 : 0;
}

// Whole pipeline, Match4PCSBase::ComputeTransformation with the default sampler.
// Q is transformed in place like the reference does. Returns the score.
float ref_compute_transformation(const float* Pxyz, const float* Pnrm, const float* Prgb, int nP,
                                 float* Qxyz, const float* Qnrm, const float* Qrgb, int nQ,
                                 const RefOptions* o, float* out_T16_colmajor) {
  Utils::Logger logger(Utils::NoLog);
  Match4PCSOptions opt = to_options(o);
  std::vector<Point3D> P, Q;
  fill_cloud(P, Pxyz, Pnrm, Prgb, nP);
  fill_cloud(Q, Qxyz, Qnrm, Qrgb, nQ);
  MatchSuper4PCS matcher(opt, logger);
  Match4PCSBase::MatrixType mat(Match4PCSBase::MatrixType::Identity());
  float score = matcher.ComputeTransformation(P, &Q, mat);
  std::memcpy(out_T16_colmajor, mat.data(), 16 * sizeof(float));
  for (int i = 0; i < nQ; ++i)
    for (int k = 0; k < 3; ++k) Qxyz[3 * i + k] = Q[i].pos()[k];
  return score;
}


// Whole pipeline through a BASE-CLASS pointer (the Meshlab plugin's usage pattern,
// demos/MeshlabPlugin/.../globalregistration.cpp:161-198: polymorphic new / delete) with a visitor that
// asks for GLOBAL transforms and records every per-iteration report (fraction >= 0):
// out_trace: max_trace x 18 floats = (fraction, best_LCP, 4x4 column-major). Returns the number of reports.
int ref_compute_transformation_traced(const float* Pxyz, int nP, const float* Qxyz, int nQ, const RefOptions* o,
                                      float* out_score, float* out_T16_colmajor, float* out_trace, int max_trace) {
  struct Trace {
    mutable std::vector<float> rows;
    mutable int n = 0;
    inline void operator()(float fraction, float best, Eigen::Ref<Match4PCSBase::MatrixType> T) const {
      if (fraction < 0) return;
      rows.push_back(fraction);
      rows.push_back(best);
      Match4PCSBase::MatrixType M = T;
      for (int i = 0; i < 16; ++i) rows.push_back(M.data()[i]);
      ++n;
    }
    constexpr bool needsGlobalTransformation() const { return true; }
  };
  Utils::Logger logger(Utils::NoLog);
  Match4PCSOptions opt = to_options(o);
  std::vector<Point3D> P, Q;
  fill_cloud(P, Pxyz, nullptr, nullptr, nP);
  fill_cloud(Q, Qxyz, nullptr, nullptr, nQ);
  Match4PCSBase* matcher = new MatchSuper4PCS(opt, logger);
  Match4PCSBase::MatrixType mat(Match4PCSBase::MatrixType::Identity());
  Trace tr;
  *out_score = matcher->ComputeTransformation(P, &Q, mat, Sampling::UniformDistSampler(), tr);
  delete matcher;
  std::memcpy(out_T16_colmajor, mat.data(), 16 * sizeof(float));
  const int n = tr.n < max_trace ? tr.n : max_trace;
  std::memcpy(out_trace, tr.rows.data(), sizeof(float) * 18 * n);
  return tr.n;
}

}  // extern "C"
