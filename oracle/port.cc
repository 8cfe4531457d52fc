// TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the Super4PCS hot path.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load the library this file builds (oracle/liboracle_port.so). The product
// (libs4g.so, super4pcs_b200/) never includes, links or calls anything from oracle/.
//
// Plain C++ (no Eigen, no reference headers): every Eigen expression of the reference
// is spelled out operation by operation, in the association order the reference's
// binary actually evaluates (g++ -O3, SSE2, no FMA contraction; this file is built
// with -ffp-contract=off). The orders were pinned empirically bit-for-bit against
// oracle/_ref (the unmodified reference) -- see tests/test_oracle_port_vs_ref.py and
// the golden vectors in tests/golden/ that were generated from oracle/_ref.
//
// Citations are file:line into /root/reference/src/super4pcs/.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <set>
#include <utility>
#include <vector>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct V3 { float x, y, z; };

inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 mul(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 smul(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 divs(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
// Eigen redux of a 3-vector: a0 + (a1 + a2)   (re-verified)
inline float sum3(float a, float b, float c) { return a + (b + c); }
inline float dot(V3 a, V3 b) { return sum3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float sqnorm(V3 a) { return dot(a, a); }
inline float norm(V3 a) { return std::sqrt(sqnorm(a)); }
// Eigen MatrixBase::cross for 3-vectors (Geometry/OrthoMethods.h)
inline V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// MatrixBase::normalized()/normalize() (Core/Dot.h): divide by sqrt(squaredNorm) if >0
inline V3 normalized(V3 a) {
  float z = sqnorm(a);
  if (z > 0.f) return divs(a, std::sqrt(z));
  return a;
}

struct M3 { float m[3][3]; };  // m[row][col]

// coefficient of a lazy 3x3 product: sum_k a(i,k) b(k,j), Eigen redux order
inline float prod_coeff(const M3& a, const M3& b, int i, int j) {
  return sum3(a.m[i][0] * b.m[0][j], a.m[i][1] * b.m[1][j], a.m[i][2] * b.m[2][j]);
}
inline V3 mulMV(const M3& a, V3 v) {
  return {sum3(a.m[0][0] * v.x, a.m[0][1] * v.y, a.m[0][2] * v.z),
          sum3(a.m[1][0] * v.x, a.m[1][1] * v.y, a.m[1][2] * v.z),
          sum3(a.m[2][0] * v.x, a.m[2][1] * v.y, a.m[2][2] * v.z)};
}

const float kLargeNumber = 1e9f;

// ---------------------------------------------------------------------------------
// a6: Match4PCSBase::ComputeRigidTransformation (algorithms/match4pcsBase.cc:365-500)
// with computeScale == false. T is COLUMN-major 4x4 (Eigen default).
// Returns the bool the reference returns ("return kLargeNumber" from a bool function
// => true with rms = 1e9, match4pcsBase.cc:417-434).
// ---------------------------------------------------------------------------------
bool rigid(const V3 ref[4], const V3 cand[4], V3 centroid1, V3 centroid2,
           float max_angle, float* T, float* rms_out) {
  *rms_out = kLargeNumber;
  for (int i = 0; i < 16; ++i) T[i] = 0.f;
  const float kSmallNumber = 1e-6f;
  V3 p0 = ref[0], p1 = ref[1], p2 = ref[2];
  V3 q0 = cand[0], q1 = cand[1], q2 = cand[2];

  V3 vp1 = sub(p1, p0);                                   // cc:415
  if (sqnorm(vp1) == 0) return true;
  vp1 = normalized(vp1);
  V3 d = sub(p2, p0);
  V3 vp2 = sub(d, smul(dot(d, vp1), vp1));                // cc:418
  if (sqnorm(vp2) == 0) return true;
  vp2 = normalized(vp2);
  V3 vp3 = cross(vp1, vp2);
  if (sqnorm(vp3) == 0) return true;
  vp3 = normalized(vp3);

  V3 vq1 = sub(q1, q0);                                   // cc:425
  if (sqnorm(vq1) == 0) return true;
  vq1 = normalized(vq1);
  V3 e = sub(q2, q0);
  V3 vq2 = sub(e, smul(dot(e, vq1), vq1));
  if (sqnorm(vq2) == 0) return true;
  vq2 = normalized(vq2);
  V3 vq3 = cross(vq1, vq2);
  if (sqnorm(vq3) == 0) return true;
  vq3 = normalized(vq3);

  // rotate_p / rotate_q hold the frames as ROWS (cc:439-447); R = rotate_p^T * rotate_q
  M3 fp = {{{vp1.x, vp1.y, vp1.z}, {vp2.x, vp2.y, vp2.z}, {vp3.x, vp3.y, vp3.z}}};
  M3 fq = {{{vq1.x, vq1.y, vq1.z}, {vq2.x, vq2.y, vq2.z}, {vq3.x, vq3.y, vq3.z}}};
  M3 fpt;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) fpt.m[i][j] = fp.m[j][i];
  M3 R;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R.m[i][j] = prod_coeff(fpt, fq, i, j);

  // cc:453: ((R*R).diagonal().array() - 1 > 1e-6).any()  (R*R, not R*R^T)
  for (int i = 0; i < 3; ++i)
    if (prod_coeff(R, R, i, i) - 1.f > kSmallNumber) return false;

  if (max_angle >= 0) {                                   // cc:457-472
    bool ok =
        std::abs(std::atan2(R.m[2][1], R.m[2][2])) <= max_angle &&
        std::abs(std::atan2(-R.m[2][0], std::sqrt(std::pow(R.m[2][1], 2) + std::pow(R.m[2][2], 2)))) <= max_angle &&
        std::abs(std::atan2(R.m[1][0], R.m[0][0])) <= max_angle;
    if (!ok) return false;
  }

  float rms = 0.f;                                        // cc:477-489
  for (int i = 0; i < 3; ++i) {
    V3 first = sub(smul(1.f, cand[i]), centroid2);
    V3 tr = mulMV(R, first);
    rms += norm(add(sub(tr, ref[i]), centroid1));
  }
  rms /= 4.f;
  *rms_out = rms;

  // cc:491-497: Identity.scale(1).translate(c1).rotate(R).translate(-c2)
  //   => linear = R, translation = c1 + R*(-c2)
  V3 t = add(centroid1, mulMV(R, neg(centroid2)));
  for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) T[4 * c + r] = R.m[r][c];
  T[12] = t.x; T[13] = t.y; T[14] = t.z; T[15] = 1.f;
  return true;
}

// (M * q.homogeneous()).head<3>() of a column-major 4x4: ((m0 x + m1 y) + m2 z) + m3
// (match4pcsBase.cc:532; order measured in SURVEY Appendix B.3)
inline V3 xform(const float* T, V3 q) {
  return {((T[0] * q.x + T[4] * q.y) + T[8] * q.z) + T[12],
          ((T[1] * q.x + T[5] * q.y) + T[9] * q.z) + T[13],
          ((T[2] * q.x + T[6] * q.y) + T[10] * q.z) + T[14]};
}

// ---------------------------------------------------------------------------------
// a8 accelerator: uniform hash grid over P (stand-in for accelerators/kdtree.h; the
// semantics restated are "exists p: ||Tq - p||^2 <= delta^2", kdtree.h:388-453 with
// the leaf test of line 418).
// ---------------------------------------------------------------------------------
struct Grid {
  float h = 0, inv_h = 0; V3 lo{0, 0, 0}; int nx = 0, ny = 0, nz = 0;
  std::vector<uint32_t> start; std::vector<V3> pts;
  inline int cx(float x) const { return (int)std::floor((x - lo.x) * inv_h); }
  inline int cy(float y) const { return (int)std::floor((y - lo.y) * inv_h); }
  inline int cz(float z) const { return (int)std::floor((z - lo.z) * inv_h); }
  void build(const V3* P, int n, float delta) {
    h = delta * 1.01f; if (!(h > 0)) h = 1.f;
    V3 mn = P[0], mx = P[0];
    for (int i = 1; i < n; ++i) {
      mn.x = std::min(mn.x, P[i].x); mn.y = std::min(mn.y, P[i].y); mn.z = std::min(mn.z, P[i].z);
      mx.x = std::max(mx.x, P[i].x); mx.y = std::max(mx.y, P[i].y); mx.z = std::max(mx.z, P[i].z);
    }
    // cap the table size
    for (;;) {
      inv_h = 1.f / h; lo = {mn.x - h, mn.y - h, mn.z - h};
      double ex = (mx.x - lo.x) / h + 2, ey = (mx.y - lo.y) / h + 2, ez = (mx.z - lo.z) / h + 2;
      if (ex * ey * ez < 6.0e7) { nx = (int)ex + 1; ny = (int)ey + 1; nz = (int)ez + 1; break; }
      h *= 1.5f;
    }
    size_t nc = (size_t)nx * ny * nz;
    start.assign(nc + 1, 0);
    std::vector<uint32_t> key(n);
    for (int i = 0; i < n; ++i) {
      key[i] = (uint32_t)(((size_t)cz(P[i].z) * ny + cy(P[i].y)) * nx + cx(P[i].x));
      start[key[i] + 1]++;
    }
    for (size_t c = 0; c < nc; ++c) start[c + 1] += start[c];
    pts.resize(n);
    std::vector<uint32_t> fill(start.begin(), start.end() - 1);
    for (int i = 0; i < n; ++i) pts[fill[key[i]]++] = P[i];
  }
  inline bool any_within(V3 t, float sq_eps) const {
    // neighbourhood wide enough for radius delta whatever h >= 1.01 delta is
    int x0 = cx(t.x), y0 = cy(t.y), z0 = cz(t.z);
    for (int z = z0 - 1; z <= z0 + 1; ++z) {
      if (z < 0 || z >= nz) continue;
      for (int y = y0 - 1; y <= y0 + 1; ++y) {
        if (y < 0 || y >= ny) continue;
        int xa = std::max(x0 - 1, 0), xb = std::min(x0 + 1, nx - 1);
        if (xa > xb) continue;
        size_t row = ((size_t)z * ny + y) * nx;
        for (uint32_t k = start[row + xa]; k < start[row + xb + 1]; ++k) {
          V3 d = sub(t, pts[k]);
          if (sqnorm(d) <= sq_eps) return true;           // kdtree.h:417-418
        }
      }
    }
    return false;
  }
};

struct Port {
  std::vector<V3> P, Q, Qn, Qrgb;   // centred sampled clouds (Q with normals / rgb)
  float delta = 0;
  Grid grid;
  // a1 state (pairCreationFunctor.h:90-122)
  V3 gcenter{0, 0, 0}; float ratio = 1.f; std::vector<V3> qunit;
  std::vector<int32_t> pairs, quads;
};

// a8: Match4PCSBase::Verify (match4pcsBase.cc:508-567). best_lcp drives the early exit
// exactly like best_LCP_ (terminate_value = size_t(best_LCP_*N), cc:520,558-560).
float verify(const Port& s, const float* T, float best_lcp, uint32_t* good_out) {
  const float epsilon = s.delta;
  const float sq_eps = epsilon * epsilon;                  // cc:522
  const size_t n = s.Q.size();
  const size_t terminate_value = (size_t)(best_lcp * n);
  uint32_t good = 0;
  for (size_t i = 0; i < n; ++i) {
    V3 t = xform(T, s.Q[i]);
    if (s.grid.any_within(t, sq_eps)) good++;
    if (n - i + good < terminate_value) break;            // cc:558
  }
  if (good_out) *good_out = good;
  return float(good) / float(n);                          // cc:566
}

}  // namespace

extern "C" {

int port_abi_version() { return 1; }
int port_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// P, Q: CENTRED sampled clouds (what Match4PCSBase::init leaves in sampled_P_3D_ /
// sampled_Q_3D_, match4pcsBase.hpp:142-149). Qnrm / Qrgb may be null (=> zero normals,
// rgb = -1, the Point3D defaults, shared4pcs.h:104-109).
void* port_create(const float* Pxyz, int nP, const float* Qxyz, const float* Qnrm,
                  const float* Qrgb, int nQ, float delta) {
  Port* s = new Port;
  s->delta = delta;
  s->P.resize(nP); s->Q.resize(nQ); s->Qn.resize(nQ); s->Qrgb.resize(nQ);
  for (int i = 0; i < nP; ++i) s->P[i] = {Pxyz[3 * i], Pxyz[3 * i + 1], Pxyz[3 * i + 2]};
  for (int i = 0; i < nQ; ++i) {
    s->Q[i] = {Qxyz[3 * i], Qxyz[3 * i + 1], Qxyz[3 * i + 2]};
    s->Qn[i] = Qnrm ? V3{Qnrm[3 * i], Qnrm[3 * i + 1], Qnrm[3 * i + 2]} : V3{0, 0, 0};
    s->Qrgb[i] = Qrgb ? V3{Qrgb[3 * i], Qrgb[3 * i + 1], Qrgb[3 * i + 2]} : V3{-1, -1, -1};
  }
  if (nP > 0) s->grid.build(s->P.data(), nP, delta);

  // a1: PairCreationFunctor::synch3DContent (pairCreationFunctor.h:90-122)
  if (nQ > 0) {
    V3 mn = s->Q[0], mx = s->Q[0];
    for (int i = 1; i < nQ; ++i) {
      V3 q = s->Q[i];
      mn.x = std::min(mn.x, q.x); mn.y = std::min(mn.y, q.y); mn.z = std::min(mn.z, q.z);
      mx.x = std::max(mx.x, q.x); mx.y = std::max(mx.y, q.y); mx.z = std::max(mx.z, q.z);
    }
    // AlignedBox::center() = (min+max)/2 ; diagonal() = max-min
    s->gcenter = {(mn.x + mx.x) / 2.f, (mn.y + mx.y) / 2.f, (mn.z + mx.z) / 2.f};
    V3 dg = sub(mx, mn);
    float mc = std::max(dg.x, std::max(dg.y, dg.z));
    s->ratio = (float)((double)mc + 0.001);                // h:111 (float + double literal)
    s->qunit.resize(nQ);
    for (int i = 0; i < nQ; ++i) {                         // worldToUnit, h:66-70
      V3 q = s->Q[i];
      s->qunit[i] = {(q.x - s->gcenter.x) / s->ratio + 0.5f,
                     (q.y - s->gcenter.y) / s->ratio + 0.5f,
                     (q.z - s->gcenter.z) / s->ratio + 0.5f};
    }
  }
  return s;
}
void port_destroy(void* h) { delete static_cast<Port*>(h); }

void port_get_normalization(void* h, float* out5) {
  Port* s = static_cast<Port*>(h);
  out5[0] = s->gcenter.x; out5[1] = s->gcenter.y; out5[2] = s->gcenter.z; out5[3] = s->ratio;
  out5[4] = 0;
}

// a6 for K quads, arguments prepared like TryCongruentSet does (match4pcsBase.hpp:373-434)
void port_rigid_batch(void* h, const int* base_ids4, const int* quads4k, long K,
                      float max_angle_deg, float* out_T, float* out_rms, int* out_ok) {
  Port* s = static_cast<Port*>(h);
  static const double pi = std::acos(-1);
  V3 ref[4];
  for (int k = 0; k < 4; ++k) ref[k] = s->P[base_ids4[k]];
  V3 c1 = divs(add(add(ref[0], ref[1]), ref[2]), 3.f);     // hpp:385
  float max_angle = (float)(max_angle_deg * pi / 180.0);   // hpp:426 (double -> Scalar param)
  for (long i = 0; i < K; ++i) {
    V3 cand[4];
    for (int k = 0; k < 4; ++k) cand[k] = s->Q[quads4k[4 * i + k]];
    V3 c2 = divs(add(add(cand[0], cand[1]), cand[2]), 3.f);  // hpp:415-417
    out_ok[i] = rigid(ref, cand, c1, c2, max_angle, out_T + 16 * i, out_rms + i) ? 1 : 0;
  }
}

// a8 for K transforms; returns elapsed seconds. out_good (nullable) = integer counts.
double port_verify_batch(void* h, const float* T16k, long K, float best_lcp, int nthreads,
                         float* out_lcp, uint32_t* out_good) {
  Port* s = static_cast<Port*>(h);
  auto t0 = std::chrono::steady_clock::now();
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(dynamic, 1)
#endif
  for (long i = 0; i < K; ++i) {
    uint32_t g = 0;
    out_lcp[i] = verify(*s, T16k + 16 * i, best_lcp, &g);
    if (out_good) out_good[i] = g;
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// brute-force a8 (no accelerator), for validating the grid above on small inputs
void port_verify_bruteforce(void* h, const float* T16k, long K, uint32_t* out_good) {
  Port* s = static_cast<Port*>(h);
  const float sq_eps = s->delta * s->delta;
  for (long c = 0; c < K; ++c) {
    uint32_t good = 0;
    for (size_t i = 0; i < s->Q.size(); ++i) {
      V3 t = xform(T16k + 16 * c, s->Q[i]);
      bool hit = false;
      for (size_t j = 0; j < s->P.size() && !hit; ++j) hit = sqnorm(sub(t, s->P[j])) <= sq_eps;
      good += hit;
    }
    out_good[c] = good;
  }
}

// a7: Match4PCSBase::TryCongruentSet (match4pcsBase.hpp:363-497): loop over quads in the
// given order, rigid fit, gate ok && 0 <= rms < 2*delta, Verify with the running
// best_LCP_ (early exit), strict first-max rule.
// out_state: [0]=best_lcp after, [1]=#gate-passing, out_best_index = index of the winning
// quad in the input list (-1 if none beat best_lcp_in), out_T = its transform.
void port_try_congruent_set(void* h, const int* base_ids4, const int* quads4k, long K,
                            float max_angle_deg, float best_lcp_in, float* out_state2,
                            long* out_best_index, float* out_T) {
  Port* s = static_cast<Port*>(h);
  static const double pi = std::acos(-1);
  V3 ref[4];
  for (int k = 0; k < 4; ++k) ref[k] = s->P[base_ids4[k]];
  V3 c1 = divs(add(add(ref[0], ref[1]), ref[2]), 3.f);
  float max_angle = (float)(max_angle_deg * pi / 180.0);
  float best = best_lcp_in; long best_i = -1; long nb = 0;
  for (int i = 0; i < 16; ++i) out_T[i] = (i % 5 == 0) ? 1.f : 0.f;
  for (long i = 0; i < K; ++i) {
    V3 cand[4];
    for (int k = 0; k < 4; ++k) cand[k] = s->Q[quads4k[4 * i + k]];
    V3 c2 = divs(add(add(cand[0], cand[1]), cand[2]), 3.f);
    float T[16], rms = -1;
    bool ok = rigid(ref, cand, c1, c2, max_angle, T, &rms);
    if (ok && rms >= 0.f && rms < 2.0f * s->delta) {       // hpp:436-439
      nb++;
      float lcp = verify(*s, T, best, nullptr);
      if (lcp > best) { best = lcp; best_i = i; std::memcpy(out_T, T, sizeof(T)); }  // hpp:468
    }
  }
  out_state2[0] = best; out_state2[1] = (float)nb; *out_best_index = best_i;
}

}  // extern "C"
