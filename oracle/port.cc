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
  // PairCreationFunctor::ids (pairCreationFunctor.h:118-121): identity after synch3DContent, then partially sorted IN PLACE
  // by every octree split of every ExtractPairs call -- the emission order of the pairs depends on this history
  std::vector<uint32_t> oct_ids;
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


// ---------------------------------------------------------------------------------
// a2 + a3: MatchSuper4PCS::ExtractPairs (algorithms/super4pcs.cc:183-224) restated as a
// brute-force sweep over i > j applying
//   (1) the accelerator's point test in UNIT-cube coordinates,
//       HyperSphere::intersectPoint (accelerators/pairExtraction/intersectionPrimitive.h:154-157)
//       with the rounded epsilon of GetRoundedEpsilonValue (intersectionFunctor.h:59-67) and
//       radius d/_ratio (pairCreationFunctor.h:124-129);
//   (2) the exact predicate PairCreationFunctor::process (pairCreationFunctor.h:151-218).
// The octree descent itself (intersectionFunctor.h:154-191) is a conservative pre-filter
// (box grown by epsilon vs shell) and is not restated; the reference's own test asserts the
// result equals brute force (tests/pair_extraction.cc:307-311).
// ---------------------------------------------------------------------------------
struct PairParams {
  float pair_distance, pair_normals_angle, pair_distance_epsilon;
  V3 b1_pos, b1_nrm, b1_rgb, b2_pos, b2_nrm, b2_rgb;
  float max_normal_difference, max_translation_distance, max_angle, max_color_distance;
};

inline float rounded_epsilon(float eps_norm, int* lvl) {
  const int lvlMax = -std::log2(eps_norm);             // intersectionFunctor.h:60 (float -> int)
  if (lvl) *lvl = lvlMax;
  return 1.f / std::pow(2, lvlMax);                    // double expression narrowed to Scalar
}

// returns bit 0: emit (j,i), bit 1: emit (i,j)
inline int pair_predicate(const Port& s, const PairParams& pp, int i, int j, float nRadius,
                          float eps_round, V3 segment1) {
  // (1) intersectPoint in unit coordinates: SQR(|pos - center| - radius) < SQR(epsilon)
  {
    float dn = norm(sub(s.qunit[j], s.qunit[i])) - nRadius;
    if (!(dn * dn < eps_round * eps_round)) return 0;
  }
  // (2) process(i, j), i > j: p = Q_[j], q = Q_[i]
  V3 p = s.Q[j], q = s.Q[i];
  const float distance = norm(sub(q, p));                               // h:160
  const double pair_distance = pp.pair_distance, pair_eps = pp.pair_distance_epsilon;
  if (std::abs((double)distance - pair_distance) > pair_eps) return 0;  // h:162 (double compare)
  V3 pn = s.Qn[j], qn = s.Qn[i];
  if (pp.max_normal_difference > 0 && sqnorm(qn) > 0 && sqnorm(pn) > 0) {   // h:165-180
    const float norm_threshold = (float)(0.5 * pp.max_normal_difference * M_PI / 180.0);
    const double first_normal_angle = norm(sub(qn, pn));
    const double second_normal_angle = norm(add(qn, pn));
    const double pna = pp.pair_normals_angle;
    const float first_norm_distance =
        (float)std::min(std::abs(first_normal_angle - pna), std::abs(second_normal_angle - pna));
    if (first_norm_distance > norm_threshold) return 0;
  }
  if (pp.max_color_distance > 0) {                                          // h:182-192
    V3 pc = s.Qrgb[j], qc = s.Qrgb[i];
    const bool use_rgb = (pc.x >= 0 && qc.x >= 0 && pp.b1_rgb.x >= 0 && pp.b2_rgb.x >= 0);
    bool color_good = norm(sub(pc, pp.b1_rgb)) < pp.max_color_distance &&
                      norm(sub(qc, pp.b2_rgb)) < pp.max_color_distance;
    if (use_rgb && !color_good) return 0;
  }
  if (pp.max_translation_distance > 0) {                                    // h:194-200
    const bool dist_good = norm(sub(p, pp.b1_pos)) < pp.max_translation_distance &&
                           norm(sub(q, pp.b2_pos)) < pp.max_translation_distance;
    if (!dist_good) return 0;
  }
  if (pp.max_angle > 0) {                                                   // h:203-212
    V3 segment2 = normalized(sub(q, p));
    int r = 0;
    if (std::acos(dot(segment1, segment2)) <= pp.max_angle * M_PI / 180.0) r |= 1;
    if (std::acos(dot(segment1, neg(segment2))) <= pp.max_angle * M_PI / 180.0) r |= 2;
    return r;
  }
  return 3;
}


// ---------------------------------------------------------------------------------
// The EMISSION ORDER of a2/a3 (the pair set is restated above as a brute-force sweep): the
// reference feeds FindCongruentQuadrilaterals the pairs in the order IntersectionFunctor::process
// produces them (accelerators/pairExtraction/intersectionFunctor.h:104-236) -- for every primitive
// (= point id i, ascending) the leaves of an octree over the unit cube in a fixed order, and inside
// a leaf the points in the order of the shared, persistent id array.  The candidate order that
// follows from it decides which of several candidates with EQUAL inlier counts is kept
// (match4pcsBase.hpp:468), so it is part of the observable behaviour.  Restated here:
//   level loop (h:141-179): a node that intersects some primitive is split when it holds more than
//   minNodeSize ids (NdNode::split, intersectionNode.h:187-249: one in-place partition per
//   dimension, NdNode::_split :165-185), else parked as an "early" leaf; others are dropped;
//   emission loop (h:186-233): per primitive, the last level's nodes, then the early leaves.
// ---------------------------------------------------------------------------------
struct OctNode { V3 c; uint32_t b, e; };

inline float v3get(const V3& v, int d) { return d == 0 ? v.x : d == 1 ? v.y : v.z; }
inline void v3set(V3& v, int d, float x) { (d == 0 ? v.x : d == 1 ? v.y : v.z) = x; }

// HyperSphere::intersect (intersectionPrimitive.h:108-130): Arvo's box-sphere test on the SHELL of
// radius r -- the box must reach inside the sphere and not lie entirely inside it
inline bool shell_hits_box(const V3& centre, float radius, const V3& box, float half) {
  float near2[3], far2[3];
  for (int d = 0; d < 3; ++d) {
    const float c = v3get(centre, d), lo = v3get(box, d) - half, hi = v3get(box, d) + half;
    const float qlo = (c - lo) * (c - lo), qhi = (c - hi) * (c - hi);
    near2[d] = c < lo ? qlo : (c > hi ? qhi : 0.f);
    far2[d] = std::max(qlo, qhi);
  }
  const float dmin = near2[0] + (near2[1] + near2[2]);   // Eigen redux of a 3-array
  const float dmax = far2[0] + (far2[1] + far2[2]);
  const float r2 = radius * radius;
  return dmin < r2 && r2 < dmax;
}

// NdNode::_split: two-sided sweep that moves ids with coordinate < value to the front
inline uint32_t oct_partition(const std::vector<V3>& pts, std::vector<uint32_t>& ids, int first, int last, int dim, float value) {
  int lo = first, hi = last - 1;
  while (lo < hi) {
    while (lo < last && v3get(pts[ids[lo]], dim) < value) ++lo;
    while (hi >= first && v3get(pts[ids[hi]], dim) >= value) --hi;
    if (lo > hi) break;
    std::swap(ids[lo], ids[hi]);
    ++lo; --hi;
  }
  if (lo >= last) return (uint32_t)last;
  return v3get(pts[ids[lo]], dim) < value ? (uint32_t)(lo + 1) : (uint32_t)lo;
}

// NdNode::split: eight children in the order of the dimension-by-dimension halving, empty ones removed
inline void oct_split(const OctNode& n, float half, const std::vector<V3>& pts, std::vector<uint32_t>& ids,
                      std::vector<OctNode>& out) {
  OctNode kid[8];
  for (OctNode& k : kid) k = n;
  const float quarter = half / 2.f;
  for (int d = 0; d < 3; ++d) {
    const int groups = 1 << d, span = 8 >> d, mid = span / 2;
    for (int g = 0; g < groups; ++g) {
      OctNode* first = kid + g * span;
      const float centre_d = v3get(first->c, d);
      const uint32_t cut = oct_partition(pts, ids, (int)first->b, (int)first[span - 1].e, d, centre_d);
      for (int i = 0; i < mid; ++i) { v3set(first[i].c, d, centre_d - quarter); first[i].e = cut; }
      for (int i = mid; i < span; ++i) { v3set(first[i].c, d, centre_d + quarter); first[i].b = cut; }
    }
  }
  for (const OctNode& k : kid)
    if (k.e != k.b) out.push_back(k);
}

// Ordered pairs of one ExtractPairs call in the reference's emission order.  Updates s.oct_ids.
inline void extract_pairs_in_reference_order(Port& s, const PairParams& pp, float nRadius, float eps_norm, V3 segment1,
                                             std::vector<int32_t>& out) {
  const int n = (int)s.qunit.size();
  if ((int)s.oct_ids.size() != n) {
    s.oct_ids.resize(n);
    for (int i = 0; i < n; ++i) s.oct_ids[i] = (uint32_t)i;
  }
  int lvl_max = 0;
  const float eps = rounded_epsilon(eps_norm, &lvl_max);
  std::vector<OctNode> level, next;
  std::vector<std::pair<OctNode, float>> early;
  next.push_back(OctNode{V3{0.5f, 0.5f, 0.5f}, 0u, (uint32_t)n});
  for (int lvl = 0; lvl != lvl_max - 1 && !next.empty(); ++lvl) {
    const float edge = (float)(1.0 / std::pow(2, lvl));
    const float half = edge / 2.f;
    level.swap(next);
    next.clear();
    for (const OctNode& node : level) {
      for (int prim = 0; prim < n; ++prim) {
        if (!shell_hits_box(s.qunit[prim], nRadius, node.c, half + eps)) continue;
        if ((int)(node.e - node.b) > 50) oct_split(node, half, s.qunit, s.oct_ids, next);
        else early.emplace_back(node, half + eps);
        break;
      }
    }
  }
  out.clear();
  auto sweep = [&](int prim, const OctNode& node) {
    for (uint32_t k = node.b; k < node.e; ++k) {
      const int id = (int)s.oct_ids[k];
      if (prim <= id) continue;
      const int r = pair_predicate(s, pp, prim, id, nRadius, eps, segment1);
      if (r & 1) { out.push_back(id); out.push_back(prim); }
      if (r & 2) { out.push_back(prim); out.push_back(id); }
    }
  };
  for (int prim = 0; prim < n; ++prim) {
    for (const OctNode& node : next)
      if (shell_hits_box(s.qunit[prim], nRadius, node.c, eps * 2.f)) sweep(prim, node);
    for (const auto& leaf : early)
      if (shell_hits_box(s.qunit[prim], nRadius, leaf.first.c, leaf.second)) sweep(prim, leaf.first);
  }
}

// ---------------------------------------------------------------------------------
// a4: IndexedNormalSet<Point,3,7,float> (accelerators/normalset.h:71-153, normalset.hpp)
// a5: MatchSuper4PCS::FindCongruentQuadrilaterals (algorithms/super4pcs.cc:80-177)
// ---------------------------------------------------------------------------------
struct Quat { float x, y, z, w; };

// third column of householderQ() of ColPivHouseholderQR(A), A (3x2) = [v0 | v1]: the rotation
// axis Eigen's setFromTwoVectors takes from JacobiSVD<Matrix<float,2,3>>(m, ComputeFullV)
// when the two vectors are nearly opposite (Geometry/Quaternion.h:593-604; the Jacobi sweeps
// and the final sort only touch columns 0 and 1 of V).  Restates QR/ColPivHouseholderQR.h
// computeInPlace, Householder/Householder.h makeHouseholder / applyHouseholderOnTheLeft and
// HouseholderSequence::evalTo for this fixed shape.
inline void make_householder(const float* v, int n, float* essential, float* tau, float* beta) {
  float tailSq = 0.f;
  for (int i = 1; i < n; ++i) tailSq = (i == 1) ? v[i] * v[i] : tailSq + v[i] * v[i];
  if (n == 1) tailSq = 0.f;
  float c0 = v[0];
  const float tol = 1.17549435e-38f;  // numeric_limits<float>::min()
  if (tailSq <= tol) {
    *tau = 0.f; *beta = c0;
    for (int i = 0; i < n - 1; ++i) essential[i] = 0.f;
  } else {
    float b = std::sqrt(c0 * c0 + tailSq);
    if (c0 >= 0.f) b = -b;
    for (int i = 0; i < n - 1; ++i) essential[i] = v[i + 1] / (c0 - b);
    *tau = (b - c0) / b;
    *beta = b;
  }
}
// M (rows x cols, row-major, leading dim ld) <- H M, H = I - tau [1;ess][1;ess]^T
inline void apply_householder_left(float* M, int ld, int rows, int cols, const float* ess, float tau) {
  if (rows == 1) { for (int j = 0; j < cols; ++j) M[j] *= (1.f - tau); return; }
  if (tau == 0.f) return;
  float tmp[3];
  for (int j = 0; j < cols; ++j) {
    float acc = 0.f;
    for (int i = 0; i < rows - 1; ++i) {
      float t = ess[i] * M[(i + 1) * ld + j];
      acc = (i == 0) ? t : acc + t;
    }
    tmp[j] = acc + M[j];                                  // tmp += row(0)
  }
  for (int j = 0; j < cols; ++j) M[j] -= tau * tmp[j];    // row(0) -= tau * tmp
  for (int i = 0; i < rows - 1; ++i)
    for (int j = 0; j < cols; ++j) M[(i + 1) * ld + j] -= (tau * ess[i]) * tmp[j];
}
inline V3 svd_null_axis(V3 v0, V3 v1) {
  // scale = max |coeff|; m / scale
  float sc = 0.f;
  const float all[6] = {v0.x, v0.y, v0.z, v1.x, v1.y, v1.z};
  for (float a : all) sc = std::max(sc, std::fabs(a));
  if (sc == 0.f) sc = 1.f;
  float A[3][2] = {{v0.x / sc, v1.x / sc}, {v0.y / sc, v1.y / sc}, {v0.z / sc, v1.z / sc}};
  float nrm[2];
  for (int k = 0; k < 2; ++k) nrm[k] = std::sqrt(sum3(A[0][k] * A[0][k], A[1][k] * A[1][k], A[2][k] * A[2][k]));
  float hc[2] = {0.f, 0.f};
  float ess[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k = 0; k < 2; ++k) {
    int big = k;
    if (k == 0 && nrm[1] > nrm[0]) big = 1;               // maxCoeff keeps the first maximum
    if (big != k) {
      for (int r = 0; r < 3; ++r) std::swap(A[r][k], A[r][big]);
      std::swap(nrm[k], nrm[big]);
    }
    float col[3];
    for (int r = k; r < 3; ++r) col[r - k] = A[r][k];
    float beta;
    make_householder(col, 3 - k, ess[k], &hc[k], &beta);
    A[k][k] = beta;
    if (k == 0) {
      // apply to the remaining column (rows k.., col 1)
      float M[3] = {A[0][1], A[1][1], A[2][1]};
      apply_householder_left(M, 1, 3, 1, ess[0], hc[0]);
      A[0][1] = M[0]; A[1][1] = M[1]; A[2][1] = M[2];
      // (the column-norm down-date only matters for choosing later pivots; with 2 columns the
      //  second pivot is forced)
    }
  }
  // Q = H0 H1 applied to the identity, HouseholderSequence::evalTo: k = 1 then k = 0
  float Qm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  apply_householder_left(&Qm[4], 3, 2, 2, ess[1], hc[1]);
  apply_householder_left(&Qm[0], 3, 3, 3, ess[0], hc[0]);
  return {Qm[2], Qm[5], Qm[8]};
}

// QuaternionBase::setFromTwoVectors(a, b) (Geometry/Quaternion.h:578-612)
inline Quat quat_from_two_vectors(V3 a, V3 b) {
  V3 v0 = normalized(a), v1 = normalized(b);
  float c = dot(v1, v0);
  Quat q;
  if (c < -1.f + 1e-5f) {
    c = std::max(c, -1.f);
    V3 axis = svd_null_axis(v0, v1);
    float w2 = (1.f + c) * 0.5f;
    q.w = std::sqrt(w2);
    float k = std::sqrt(1.f - w2);
    q.x = axis.x * k; q.y = axis.y * k; q.z = axis.z * k;
    return q;
  }
  V3 axis = cross(v0, v1);
  float sq = std::sqrt((1.f + c) * 2.f);
  float invs = 1.f / sq;
  q.x = axis.x * invs; q.y = axis.y * invs; q.z = axis.z * invs;
  q.w = sq * 0.5f;
  return q;
}
// QuaternionBase::_transformVector (Geometry/Quaternion.h:472-481)
inline V3 quat_rotate(Quat q, V3 v) {
  V3 qv = {q.x, q.y, q.z};
  V3 uv = cross(qv, v);
  uv = add(uv, uv);
  return add(add(v, smul(q.w, uv)), cross(qv, uv));
}

struct NormalSetParams {
  float nepsilon;   // 1/7 + 1e-5 (normalset.h:115)
  float epsilon;    // 1 / egSize (h:121)
  int egSize;
};
inline NormalSetParams nset_params(float eps) {
  NormalSetParams p;
  p.nepsilon = (float)(1.f / 7.f + 0.00001);             // float + double literal -> Scalar
  const int gridDepth = -std::log2(eps);                 // h:119 (float -> int truncation)
  p.egSize = (int)std::pow(2, gridDepth);                // h:120
  p.epsilon = 1.f / p.egSize;                            // h:121
  return p;
}
// UnrollIndexLoop (accelerators/utils.h:139-148) on p/_epsilon, truncating casts
inline long index_pos(const NormalSetParams& g, V3 p) {
  V3 c = divs(p, g.epsilon);
  return ((long)(int)c.z * g.egSize + (int)c.y) * (long)g.egSize + (int)c.x;
}
inline int index_normal(const NormalSetParams& g, V3 n) {
  V3 c = {(n.x / 2.f + 0.5f) / g.nepsilon, (n.y / 2.f + 0.5f) / g.nepsilon, (n.z / 2.f + 0.5f) / g.nepsilon};
  return ((int)c.z * 7 + (int)c.y) * 7 + (int)c.x;
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


// a2+a3. base_p1 / base_p2: 9 floats each (pos, normal, rgb) of base_3D_[base_point1/2];
// filters4 = {max_normal_difference, max_translation_distance, max_angle, max_color_distance}.
// Result (sorted lexicographically) is kept in the handle; returns the number of ordered pairs.
long port_extract_pairs(void* h, float pair_distance, float pair_normals_angle, float eps,
                        const float* base_p1, const float* base_p2, const float* filters4) {
  Port* s = static_cast<Port*>(h);
  PairParams pp;
  pp.pair_distance = pair_distance; pp.pair_normals_angle = pair_normals_angle; pp.pair_distance_epsilon = eps;
  pp.b1_pos = {base_p1[0], base_p1[1], base_p1[2]}; pp.b1_nrm = {base_p1[3], base_p1[4], base_p1[5]};
  pp.b1_rgb = {base_p1[6], base_p1[7], base_p1[8]};
  pp.b2_pos = {base_p2[0], base_p2[1], base_p2[2]}; pp.b2_nrm = {base_p2[3], base_p2[4], base_p2[5]};
  pp.b2_rgb = {base_p2[6], base_p2[7], base_p2[8]};
  pp.max_normal_difference = filters4[0]; pp.max_translation_distance = filters4[1];
  pp.max_angle = filters4[2]; pp.max_color_distance = filters4[3];
  const float nRadius = pair_distance / s->ratio;                 // setRadius, h:124-129
  const float eps_norm = eps / s->ratio;                          // getNormalizedEpsilon, h:131-133
  const float eps_round = rounded_epsilon(eps_norm, nullptr);
  const V3 segment1 = normalized(sub(pp.b2_pos, pp.b1_pos));      // setBase, h:135-143
  const int n = (int)s->Q.size();
  std::vector<std::vector<int32_t>> rows(n);                      // rows[a] = partners b of (a,b)
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64)
#endif
  for (int a = 0; a < n; ++a) {
    std::vector<int32_t>& row = rows[a];
    for (int b = 0; b < n; ++b) {
      if (a == b) continue;
      int i = std::max(a, b), j = std::min(a, b);
      int r = pair_predicate(*s, pp, i, j, nRadius, eps_round, segment1);
      // bit0 -> (j,i), bit1 -> (i,j)
      if ((a == j && (r & 1)) || (a == i && (r & 2))) row.push_back(b);
    }
  }
  s->pairs.clear();
  for (int a = 0; a < n; ++a)
    for (int32_t b : rows[a]) { s->pairs.push_back(a); s->pairs.push_back(b); }
  return (long)(s->pairs.size() / 2);
}
// a2+a3 in the reference's EMISSION order (see extract_pairs_in_reference_order).  Stateful like the reference: the
// order depends on every earlier call on this handle.  The pair SET equals port_extract_pairs'.
long port_extract_pairs_ordered(void* h, float pair_distance, float pair_normals_angle, float eps,
                                const float* base_p1, const float* base_p2, const float* filters4) {
  Port* s = static_cast<Port*>(h);
  PairParams pp;
  pp.pair_distance = pair_distance; pp.pair_normals_angle = pair_normals_angle; pp.pair_distance_epsilon = eps;
  pp.b1_pos = {base_p1[0], base_p1[1], base_p1[2]}; pp.b1_nrm = {base_p1[3], base_p1[4], base_p1[5]};
  pp.b1_rgb = {base_p1[6], base_p1[7], base_p1[8]};
  pp.b2_pos = {base_p2[0], base_p2[1], base_p2[2]}; pp.b2_nrm = {base_p2[3], base_p2[4], base_p2[5]};
  pp.b2_rgb = {base_p2[6], base_p2[7], base_p2[8]};
  pp.max_normal_difference = filters4[0]; pp.max_translation_distance = filters4[1];
  pp.max_angle = filters4[2]; pp.max_color_distance = filters4[3];
  const float nRadius = pair_distance / s->ratio;
  const float eps_norm = eps / s->ratio;
  const V3 segment1 = normalized(sub(pp.b2_pos, pp.b1_pos));
  extract_pairs_in_reference_order(*s, pp, nRadius, eps_norm, segment1, s->pairs);
  return (long)(s->pairs.size() / 2);
}

void port_get_pairs(void* h, int32_t* out) {
  Port* s = static_cast<Port*>(h);
  std::memcpy(out, s->pairs.data(), s->pairs.size() * sizeof(int32_t));
}


// a4+a5 on explicit pair lists (P_pairs = pairs1, Q_pairs = pairs2, K x 2 int32 each).
// base_xyz: base_3D_[0..3] positions (12 floats).  Quads (sorted like the std::set<(id,i)>
// of super4pcs.cc:127) are kept in the handle; returns their number.
long port_find_quads(void* h, float invariant1, float invariant2, float distance_threshold2,
                     const float* base_xyz, const int32_t* pairs1, long n1, const int32_t* pairs2, long n2) {
  Port* s = static_cast<Port*>(h);
  V3 b[4];
  for (int k = 0; k < 4; ++k) b[k] = {base_xyz[3 * k], base_xyz[3 * k + 1], base_xyz[3 * k + 2]};
  const float alpha_cos = dot(normalized(sub(b[1], b[0])), normalized(sub(b[3], b[2])));   // cc:109-111
  const float eps = distance_threshold2 / s->ratio;                                      // cc:114
  const NormalSetParams g = nset_params(eps);

  // build: key (cell, bin) -> ids in insertion order (cc:116-124, normalset.hpp:110-127)
  std::vector<std::pair<std::pair<long, int>, uint32_t>> entries(n1);
  for (long i = 0; i < n1; ++i) {
    V3 p1 = s->qunit[pairs1[2 * i]], p2 = s->qunit[pairs1[2 * i + 1]];
    V3 d = sub(p2, p1);
    V3 n = normalized(d);
    V3 pos = add(p1, smul(invariant1, d));
    entries[i] = {{index_pos(g, pos), index_normal(g, n)}, (uint32_t)i};
  }
  std::stable_sort(entries.begin(), entries.end(),
                   [](const std::pair<std::pair<long, int>, uint32_t>& a,
                      const std::pair<std::pair<long, int>, uint32_t>& c) { return a.first < c.first; });

  // getNeighbors constants (normalset.hpp:174-181)
  const float alpha = std::acos(alpha_cos);
  const float perimeter = (float)(2.f * M_PI * std::atan(alpha));
  const unsigned int nbSample = (unsigned int)(2 * std::ceil(perimeter * 7.f / 2.f));
  const float angleStep = (float)(2.f * M_PI / float(nbSample));
  const float sinAlpha = std::sin(alpha);
  std::vector<V3> ring(nbSample);
  for (unsigned int a = 0; a < nbSample; ++a) {
    float theta = float(a) * angleStep;
    ring[a] = {sinAlpha * std::cos(theta), sinAlpha * std::sin(theta), alpha_cos};
  }

  std::vector<std::vector<std::pair<uint32_t, uint32_t>>> found(
#ifdef _OPENMP
      omp_get_max_threads()
#else
      1
#endif
  );
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 256)
#endif
  for (long i = 0; i < n2; ++i) {
#ifdef _OPENMP
    std::vector<std::pair<uint32_t, uint32_t>>& out = found[omp_get_thread_num()];
#else
    std::vector<std::pair<uint32_t, uint32_t>>& out = found[0];
#endif
    const int a0 = pairs2[2 * i], a1 = pairs2[2 * i + 1];
    V3 p1 = s->qunit[a0], p2 = s->qunit[a1];
    V3 pq1 = s->Q[a0], pq2 = s->Q[a1];
    V3 query = add(p1, smul(invariant2, sub(p2, p1)));              // cc:141
    V3 queryQ = add(pq1, smul(invariant2, sub(pq2, pq1)));          // cc:142
    V3 queryn = normalized(sub(p2, p1));
    const long cell = index_pos(g, query);
    auto lo = std::lower_bound(entries.begin(), entries.end(), std::make_pair(std::make_pair(cell, 0), 0u));
    if (lo == entries.end() || lo->first.first != cell) continue;   // angularGrid(p) == NULL
    Quat q = quat_from_two_vectors({0.f, 0.f, 1.f}, queryn);
    bool colored[343] = {false};
    for (unsigned int a = 0; a < nbSample; ++a) {
      V3 dir = normalized(quat_rotate(q, ring[a]));
      int id = index_normal(g, dir);
      if (id >= 0 && id < 343) colored[id] = true;   // emptiness is checked by the range scan below
    }
    for (auto it = lo; it != entries.end() && it->first.first == cell; ++it) {
      if (!colored[it->first.second]) continue;
      const uint32_t id = it->second;
      V3 pp1 = s->Q[pairs1[2 * id]], pp2 = s->Q[pairs1[2 * id + 1]];
      V3 invPoint = add(pp1, mul(sub(pp2, pp1), invariant1));       // cc:157
      if (sqnorm(sub(queryQ, invPoint)) <= distance_threshold2)    // cc:160 (squared vs un-squared, sic)
        out.emplace_back(id, (uint32_t)i);
    }
  }
  std::vector<std::pair<uint32_t, uint32_t>> comb;
  for (auto& f : found) comb.insert(comb.end(), f.begin(), f.end());
  std::sort(comb.begin(), comb.end());                               // std::set order, cc:127
  s->quads.clear();
  for (auto& c : comb) {
    s->quads.push_back(pairs1[2 * c.first]); s->quads.push_back(pairs1[2 * c.first + 1]);
    s->quads.push_back(pairs2[2 * c.second]); s->quads.push_back(pairs2[2 * c.second + 1]);
  }
  return (long)(s->quads.size() / 4);
}
void port_get_quads(void* h, int32_t* out) {
  Port* s = static_cast<Port*>(h);
  std::memcpy(out, s->quads.data(), s->quads.size() * sizeof(int32_t));
}

}  // extern "C"
