// a8 -- Match4PCSBase::Verify (reference algorithms/match4pcsBase.cc:508-567) for a batch of
// candidate transforms, on the bricked uniform grid + delta-field built by s4g_set_cloud_p.
//
//   counts[k] = #{ q in sampled_Q : exists p in sampled_P, ||T_k q - p||^2 <= delta^2 }
//
// Arithmetic of the DECISION follows the reference's binary operation by operation (SURVEY.md
// A.3 / B.3):
//   T q  = ((m0 x + m1 y) + m2 z) + m3   per row, fp32, no FMA      (match4pcsBase.cc:532)
//   d^2  = dx^2 + (dy^2 + dz^2)                                     (kdtree.h:417)
//   hit  = d^2 <= delta*delta                                       (kdtree.h:418, cc:522)
//
// Schedule (B200).  The working set (Q 16 MB, delta-field 27 + 35 MB, sorted P 16 MB, cellStart, brick
// tables, a 4 MB summed-area table: ~110 MB at 1M points) is mostly L2-resident; the kernel is bound by
// instruction issue and L2 latency, so the design minimises instructions per (query, candidate) pair -- by
// deciding as many pairs as possible from pre-computed bits instead of point tests:
//  * sampled_Q is streamed in Morton order, one coalesced float4 per thread per tile; a CTA stages a
//    chunk of 16 candidate transforms and owns 8 tiles of 128 queries.
//  * phase 0 (two levels, one thread per (tile, candidate), then one per (surviving pair, warp)): the bounding sphere of
//    the 128 queries of a tile -- then of the 32 Morton-consecutive queries one WARP owns -- transformed and grown by
//    delta, is tested against the summed-area table of the coarse occupancy (8 look-ups); a warp only visits the
//    candidates that survive for its own queries.
//  * phase 1 (every surviving pair, ~40 instructions): the VOXEL-space image V q (V = the transform
//    pre-multiplied by the world->voxel map, voxel edge = h/4 ~ delta/2; 9 FMAs) addresses the
//    delta-field (GridDev::vox): v-brick table entry, then 2 bits:
//        neither  -> no P point within delta of any location of that voxel: not an inlier, done;
//        CERTAIN  -> some P point is within delta of every location of the voxel: inlier, done;
//        MAYBE    -> refined by the same two bits of the 2x2x2 sub-voxel (edge h/8) that holds the position
//                    (GridDev::vfine); only a pair that is MAYBE again is queued for the exact test.
//    The field is built with a margin (GridDev::vslack) that covers the rounding difference between
//    V q (FMA chain) and the reference-order T q for rigid motions of clouds of this size; every
//    candidate's rounding bound is checked against it when the CTA stages it, and a candidate that
//    exceeds it (huge coefficients, non-rigid 4x4, NaN) takes the robust path: the voxel is derived
//    from the reference-order T q itself, whose voxel coordinate is accurate to 1e-3 voxel whatever T.
//    Either way the bits only replace point tests whose outcome they imply, so counts stay exact.
//  * phase 2 (dense, one queued pair per lane; every WARP has its own queue across its tiles and flushes it on
//    its own -- no CTA barrier inside the tile loop): exact T q in the reference's operation order, the 2x2x2
//    cell block that contains every P point within delta, its occupancy nibble (GridDev::vocc), the non-empty
//    rows' contiguous point runs, d^2 <= delta^2, first hit wins.
//
// Probe (phase 2): cell edge h >= 2.02*delta, so the delta-ball around t = T q touches at most 2 cells per
// axis: x0 = floor(u - 0.5), u = (t - o)/h, cells {x0, x0+1}.  For a point p with |t - p|_x <= delta(1+1e-6):
// |u - v| <= 0.4951, u in [x0+0.5, x0+1.5) => v in (x0+0.0049, x0+1.9951); the 0.0049-cell margin
// dominates the rounding of u and v (< 4e-4 cell for grids <= 2048 cells per axis), so the block is
// conservative and the count exact.
#include "s4g_internal.cuh"

namespace {

constexpr int kThreads = kVerifyTile;   // 128: one query per thread per tile
constexpr int kCandPerBlock = 16;   // transforms staged per CTA
constexpr int kTilesPerBlock = kThreads / 16;  // (tile, candidate) pairs of a CTA = one per thread in the cull phase
// Compile-time knobs for A/B runs on the GPU (scripts/verify_ab.sh rebuilds libs4g with S4G_NVCC_DEFINES and re-runs
// parity + bench).
#ifndef S4G_QUEUE_CAP
#define S4G_QUEUE_CAP 2048
#endif
#ifndef S4G_VERIFY_MIN_BLOCKS
#define S4G_VERIFY_MIN_BLOCKS 12
#endif
#ifndef S4G_RASTER_Y
#define S4G_RASTER_Y 0             // 1: candidate chunks vary fastest across consecutive CTAs (A/B of the L2 behaviour)
#endif
#ifndef S4G_SUBVOXEL
#define S4G_SUBVOXEL 1             // 0: ignore the second level of the delta-field (A/B)
#endif
#ifndef S4G_FLUSH_MIN
#define S4G_FLUSH_MIN (S4G_QUEUE_CAP / 2)
#endif
constexpr int kQueueCap = S4G_QUEUE_CAP;   // queued pairs of a CTA (uint16 entries: candidate << 10 | tile << 7 | thread)
constexpr int kWarpQueue = kQueueCap / (kThreads / 32);   // ... of one warp
constexpr int kWarpFlush = S4G_FLUSH_MIN / (kThreads / 32);   // a warp runs phase 2 once this many of its pairs wait (or at its last tile)
static_assert(kQueueCap >= kThreads && kQueueCap <= 16384, "queue capacity (uint16 entries in shared memory)");
static_assert(kTilesPerBlock == 8 && kCandPerBlock == 16, "entry layout: 4 + 3 + 7 bits; 4-bit certain counters hold <= 8");

struct ProbeStats {
  unsigned long long tested = 0, ranges = 0, bricks = 0, bitmap = 0, culled = 0;
};

// Scan one contiguous run of P points; single exit, flag based.  Two points are in flight per
// iteration (the loads of a run are independent; the kernel is latency-bound here).
template <bool kStats>
__device__ __forceinline__ bool probe_run(const GridDev& g, uint32_t s, uint32_t e, float tx, float ty,
                                          float tz, float sq_eps, ProbeStats& st) {
  bool found = false;
  for (uint32_t k = s; k < e && !found; k += 2) {
    const bool two = k + 1 < e;
    const float4 p = __ldg(&g.pts[k]);
    const float4 r = __ldg(&g.pts[two ? k + 1 : k]);
    const float dx = __fsub_rn(tx, p.x), dy = __fsub_rn(ty, p.y), dz = __fsub_rn(tz, p.z);
    const float ex = __fsub_rn(tx, r.x), ey = __fsub_rn(ty, r.y), ez = __fsub_rn(tz, r.z);
    const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dz, dz)));
    const float e2 = __fadd_rn(__fmul_rn(ex, ex), __fadd_rn(__fmul_rn(ey, ey), __fmul_rn(ez, ez)));
    if (kStats) st.tested += two ? 2 : 1;
    found = d2 <= sq_eps || e2 <= sq_eps;
  }
  return found;
}

// Does any P point of row r (cells (x0..x0+1, y0 + (r&1), z0 + (r>>1))) of the block with origin
// (x0,y0,z0) lie within delta of t?  The caller guarantees -1 <= x0 <= nx-1 (same for y, z).
template <bool kStats>
__device__ __forceinline__ bool walk_row(const GridDev& g, int x0, int y0, int z0, int r, float tx, float ty,
                                         float tz, float sq_eps, ProbeStats& st) {
  const int bs = g.bshift, m = (1 << bs) - 1;
  const int xa = max(x0, 0), xb = min(x0 + 1, g.nx - 1);
  const bool same_brick = (xa >> bs) == (xb >> bs);
  const int cz = z0 + (r >> 1), cy = y0 + (r & 1);
  bool found = false;
  if (cz >= 0 && cz < g.nz && cy >= 0 && cy < g.ny) {
    const int rowb = ((cz >> bs) * g.tby + (cy >> bs)) * g.tbx;
    const uint32_t rowl = (uint32_t)((((cz & m) << bs) | (cy & m)) << bs);
    const int ra = __ldg(&g.top[rowb + (xa >> bs)]);
    if (kStats) st.bricks++;
    if (ra >= 0) {
      const uint32_t idx = ((uint32_t)ra << (3 * bs)) | rowl | (uint32_t)(xa & m);
      const uint32_t s = __ldg(&g.cellStart[idx]);
      const uint32_t e = __ldg(&g.cellStart[idx + (same_brick ? (uint32_t)(xb - xa) : 0u) + 1u]);
      if (kStats) st.ranges++;
      found = probe_run<kStats>(g, s, e, tx, ty, tz, sq_eps, st);
    }
    if (!same_brick && !found) {
      const int rb = __ldg(&g.top[rowb + (xb >> bs)]);
      if (kStats) st.bricks++;
      if (rb >= 0) {
        const uint32_t idx = ((uint32_t)rb << (3 * bs)) | rowl | (uint32_t)(xb & m);
        const uint32_t s = __ldg(&g.cellStart[idx]);
        const uint32_t e = __ldg(&g.cellStart[idx + 1u]);
        if (kStats) st.ranges++;
        found = probe_run<kStats>(g, s, e, tx, ty, tz, sq_eps, st);
      }
    }
  }
  return found;
}

// Occupancy nibble of the 2x2x2 block with origin cell (x0,y0,z0) (GridDev::vocc): bit r = row r holds points.  Origins
// outside the lattice or outside the v-bricks have no points in their block.
__device__ __forceinline__ uint32_t block_rows(const GridDev& g, int x0, int y0, int z0) {
  uint32_t nib = 0xFu;
  if (g.vocc != nullptr) {
    nib = 0u;
    if (x0 >= 0 && y0 >= 0 && z0 >= 0) {
      const int bs = g.bshift, m = (1 << bs) - 1;
      const int rank = __ldg(&g.vtop[((z0 >> bs) * g.tby + (y0 >> bs)) * g.tbx + (x0 >> bs)]);
      if (rank >= 0) {
        const uint32_t cell = ((uint32_t)rank << (3 * bs)) | (uint32_t)((((z0 & m) << bs) | (y0 & m)) << bs) | (uint32_t)(x0 & m);
        nib = (__ldg(&g.vocc[cell >> 3]) >> ((cell & 7u) * 4u)) & 0xFu;
      }
    }
  }
  return nib;
}

// T q in the reference's operation order: ((m0 x + m1 y) + m2 z) + m3 per row
__device__ __forceinline__ void exact_tq(const float* __restrict__ m, float4 q, float& tx, float& ty, float& tz) {
  tx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], q.x), __fmul_rn(m[1], q.y)), __fmul_rn(m[2], q.z)), m[3]);
  ty = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[4], q.x), __fmul_rn(m[5], q.y)), __fmul_rn(m[6], q.z)), m[7]);
  tz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[8], q.x), __fmul_rn(m[9], q.y)), __fmul_rn(m[10], q.z)), m[11]);
}

// voxel coordinates of the query under one candidate.  Fast path: FMA chain of the voxel-space matrix (selection only,
// its rounding is covered by the field's margin for candidates that passed the bound check).  Robust path: from the
// reference-order T q (accurate to 1e-3 voxel for any T).
template <bool kRobust>
__device__ __forceinline__ void voxel_of(const GridDev& g, const float* __restrict__ v, const float* __restrict__ m,
                                         float4 q, float& ux, float& uy, float& uz) {
  if (!kRobust) {
    ux = __fmaf_rn(v[0], q.x, __fmaf_rn(v[1], q.y, __fmaf_rn(v[2], q.z, v[3])));
    uy = __fmaf_rn(v[4], q.x, __fmaf_rn(v[5], q.y, __fmaf_rn(v[6], q.z, v[7])));
    uz = __fmaf_rn(v[8], q.x, __fmaf_rn(v[9], q.y, __fmaf_rn(v[10], q.z, v[11])));
  } else {
    float tx, ty, tz;
    exact_tq(m, q, tx, ty, tz);
    ux = __fmul_rn(__fsub_rn(tx, g.ox), g.inv_v);
    uy = __fmul_rn(__fsub_rn(ty, g.oy), g.inv_v);
    uz = __fmul_rn(__fsub_rn(tz, g.oz), g.inv_v);
  }
}

// Tile-level cull: can ANY point of a query tile (bounding sphere `sph`, world units) come within
// delta of a P point under the candidate whose voxel-space matrix is `v`?  Conservative test on the
// summed-area table of the coarse occupancy ((2^cshift)^3-cell blocks): the transformed sphere (radius scaled by
// `scale` >= the operator norm of the candidate's 3x3 part), grown by delta, is boxed and the number of occupied blocks
// the box touches follows from 8 table look-ups.  Run by ONE thread per (tile, candidate).
__device__ bool tile_live(const GridDev& g, const float* __restrict__ v, float4 sph, float scale) {
  // single exit, flag based (see DESIGN.md section 7 on early returns + warp votes)
  bool live = true;
  if (g.csat != nullptr) {
    // centre in cell coordinates, radius in cells: r/h + delta/h (< 0.4951) + slack
    const float cx = 0.25f * __fmaf_rn(v[0], sph.x, __fmaf_rn(v[1], sph.y, __fmaf_rn(v[2], sph.z, v[3])));
    const float cy = 0.25f * __fmaf_rn(v[4], sph.x, __fmaf_rn(v[5], sph.y, __fmaf_rn(v[6], sph.z, v[7])));
    const float cz = 0.25f * __fmaf_rn(v[8], sph.x, __fmaf_rn(v[9], sph.y, __fmaf_rn(v[10], sph.z, v[11])));
    const float R = sph.w * g.inv_h * scale * 1.0001f + 0.52f;
    const float fx0 = cx - R, fx1 = cx + R, fy0 = cy - R, fy1 = cy + R, fz0 = cz - R, fz1 = cz + R;
    // boxes entirely outside the grid cannot match anything
    const bool inside = fx1 >= 0.f && fy1 >= 0.f && fz1 >= 0.f && fx0 < (float)g.nx && fy0 < (float)g.ny &&
                        fz0 < (float)g.nz;
    live = false;
    if (inside) {
      // number of occupied coarse blocks the box touches, from the summed-area table: 8 look-ups
      const int cs = g.cshift;
      const int x0 = max(0, __float2int_rd(fx0)) >> cs, x1 = (min(g.nx - 1, __float2int_rd(fx1)) >> cs) + 1;
      const int y0 = max(0, __float2int_rd(fy0)) >> cs, y1 = (min(g.ny - 1, __float2int_rd(fy1)) >> cs) + 1;
      const int z0 = max(0, __float2int_rd(fz0)) >> cs, z1 = (min(g.nz - 1, __float2int_rd(fz1)) >> cs) + 1;
      const uint32_t sx = (uint32_t)(g.cnx + 1), sxy = sx * (uint32_t)(g.cny + 1);
      const uint32_t* __restrict__ S = g.csat;
      const uint32_t a = (uint32_t)z1 * sxy, b = (uint32_t)z0 * sxy, c = (uint32_t)y1 * sx, d = (uint32_t)y0 * sx;
      const uint32_t cnt = (__ldg(&S[a + c + x1]) - __ldg(&S[a + c + x0]) - __ldg(&S[a + d + x1]) + __ldg(&S[a + d + x0])) -
                           (__ldg(&S[b + c + x1]) - __ldg(&S[b + c + x0]) - __ldg(&S[b + d + x1]) + __ldg(&S[b + d + x0]));
      live = cnt != 0u;
    }
  }
  return live;
}

// delta-field addressing of voxel (X,Y,Z) (bit 0 MAYBE, bit 1 CERTAIN; no v-brick = neither).  kBS > 0: brick shift
// known at compile time (2 = the common 4x4x4-cell bricks), 0: read from the grid.
template <int kBS>
__device__ __forceinline__ int vtop_index(const GridDev& g, int X, int Y, int Z) {
  const int bs = kBS > 0 ? kBS : g.bshift;
  return ((Z >> (bs + 2)) * g.tby + (Y >> (bs + 2))) * g.tbx + (X >> (bs + 2));
}
// delta-field cell of voxel (X,Y,Z) inside v-brick `rank`, and the bit position of the voxel's 2 bits in its slab word
template <int kBS>
__device__ __forceinline__ uint32_t vox_cell(const GridDev& g, int rank, int X, int Y, int Z) {
  const int bs = kBS > 0 ? kBS : g.bshift, m = (1 << bs) - 1;
  return ((uint32_t)rank << (3 * bs)) | (uint32_t)(((((Z >> 2) & m) << bs) | ((Y >> 2) & m)) << bs | ((X >> 2) & m));
}
__device__ __forceinline__ uint32_t vox_shift(int X, int Y) { return 2u * (uint32_t)(((Y & 3) << 2) | (X & 3)); }

// second level: state of the sub-voxel (edge h/8) of boundary voxel (X,Y,Z) that holds the position (ux,uy,uz); w = the
// voxel's slab word, sh = its bit position (GridDev::vfine: slot = the cell's base + the boundary voxels before it)
__device__ __forceinline__ uint32_t vox_refine(const GridDev& g, uint32_t cell, uint32_t w, uint32_t sh, int X, int Y, int Z, float ux,
                                               float uy, float uz) {
  const uint4 cw = __ldg(reinterpret_cast<const uint4*>(g.vox) + cell);
  const int vz = Z & 3;
  uint32_t slot = __ldg(&g.vbase[cell]) + (uint32_t)__popc((w & ~(w >> 1) & 0x55555555u) & ((1u << sh) - 1u));
  slot += vz > 0 ? (uint32_t)__popc(cw.x & ~(cw.x >> 1) & 0x55555555u) : 0u;
  slot += vz > 1 ? (uint32_t)__popc(cw.y & ~(cw.y >> 1) & 0x55555555u) : 0u;
  slot += vz > 2 ? (uint32_t)__popc(cw.z & ~(cw.z >> 1) & 0x55555555u) : 0u;
  const uint32_t f = __ldg(&g.vfine[slot]);
  const uint32_t ch = ((ux - (float)X) >= 0.5f ? 1u : 0u) | ((uy - (float)Y) >= 0.5f ? 2u : 0u) | ((uz - (float)Z) >= 0.5f ? 4u : 0u);
  return ((f >> ch) & 1u) | (((f >> (8u + ch)) & 1u) << 1);
}

// state of the voxel (X,Y,Z) = floor of the voxel-space position (ux,uy,uz): bit 0 MAYBE, bit 1 CERTAIN; a BOUNDARY voxel
// (MAYBE, not CERTAIN) is refined to the state of its sub-voxel.
template <int kBS>
__device__ __forceinline__ uint32_t vox_state(const GridDev& g, int rank, int X, int Y, int Z, float ux, float uy, float uz) {
  const uint32_t cell = vox_cell<kBS>(g, rank, X, Y, Z), sh = vox_shift(X, Y);
  const uint32_t w = __ldg(&g.vox[(cell << 2) | (uint32_t)(Z & 3)]);
  uint32_t s = (w >> sh) & 3u;
  if (S4G_SUBVOXEL && s == 1u && g.vfine != nullptr) s = vox_refine(g, cell, w, sh, X, Y, Z, ux, uy, uz);
  return s;
}

// T12: K x 12 floats, row-major 3x4 (r00 r01 r02 t0 | r10 ... ), the top three rows of T.
// grid.x = query super-tiles (kThreads*kTilesPerBlock queries), grid.y = candidate chunks.
template <bool kStats, int kBS>
__global__ void __launch_bounds__(kThreads, kStats ? 1 : S4G_VERIFY_MIN_BLOCKS)
k_verify(GridDev g, const float4* __restrict__ Q, const float4* __restrict__ tiles, const float4* __restrict__ subs, int nQ,
         const float* __restrict__ T12, int K, float sq_eps, float qax, float qay, float qaz,
         uint32_t* __restrict__ counts, unsigned long long* __restrict__ stats, const uint32_t* __restrict__ dK) {
  if (dK != nullptr) {                            // candidate count known only on the device (s4g_try_bases): K is its upper bound
    K = min(K, (int)__ldg(dK));
    if ((int)((S4G_RASTER_Y ? blockIdx.x : blockIdx.y) * kCandPerBlock) >= K) return;   // CTA-uniform
  }
  __shared__ __align__(16) float sT[kCandPerBlock * 12];     // exact transforms (decision arithmetic)
  __shared__ __align__(16) float sV[kCandPerBlock * 12];     // voxel-space transforms (selection only)
  __shared__ float sScale[kCandPerBlock];                    // tile-cull radius scale; < 0: robust path, no cull
  __shared__ uint32_t sLive[kTilesPerBlock * (kThreads / 32)];   // [tile * 4 + warp] bit c: candidate c may hit that warp's 32 queries
  __shared__ uint32_t sCnt[kCandPerBlock];
  __shared__ uint16_t sQueue[kQueueCap];                     // (candidate, tile, query) pairs waiting for the exact test, one quarter per warp
  __shared__ uint8_t sPair[kThreads];                        // phase 0: surviving (tile << 4 | candidate) pairs
  __shared__ uint32_t sPairN[kThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31;
  const int c0 = (S4G_RASTER_Y ? blockIdx.x : blockIdx.y) * kCandPerBlock;
  const int nc = min(kCandPerBlock, K - c0);
  for (int i = tid; i < nc * 12; i += kThreads) {
    const float t = T12[(size_t)c0 * 12 + i];
    sT[i] = t;
    const int col = i & 3, row = (i % 12) >> 2;
    const float o = row == 0 ? g.ox : row == 1 ? g.oy : g.oz;
    sV[i] = col < 3 ? t * g.inv_v : (t - o) * g.inv_v;
  }
  if (tid < kCandPerBlock) {
    sCnt[tid] = 0;
    float sc = -1.f;
    if (tid < nc) {
      // rounding bound of this candidate's voxel position (fast path) against the margin the field was built with:
      // |x~ - fl(T q)| <= 2^-20 max_r (sum_j |T_rj| |q_j| + |T_r3| + |o_r|)   (16 roundings of relative size 2^-24)
      const float* __restrict__ m = &T12[(size_t)(c0 + tid) * 12];
      const float w0 = fabsf(m[0]) * qax + fabsf(m[1]) * qay + fabsf(m[2]) * qaz + fabsf(m[3]) + fabsf(g.ox);
      const float w1 = fabsf(m[4]) * qax + fabsf(m[5]) * qay + fabsf(m[6]) * qaz + fabsf(m[7]) + fabsf(g.oy);
      const float w2 = fabsf(m[8]) * qax + fabsf(m[9]) * qay + fabsf(m[10]) * qaz + fabsf(m[11]) + fabsf(g.oz);
      const float E = fmaxf(w0, fmaxf(w1, w2)) * 9.5367431640625e-7f;   // 2^-20
      // operator norm of the 3x3 part: ||A||_2^2 = lambda_max(A^T A) <= max row sum of |A^T A| (= 1 for a rotation)
      const float g00 = m[0] * m[0] + m[4] * m[4] + m[8] * m[8], g11 = m[1] * m[1] + m[5] * m[5] + m[9] * m[9],
                  g22 = m[2] * m[2] + m[6] * m[6] + m[10] * m[10];
      const float g01 = fabsf(m[0] * m[1] + m[4] * m[5] + m[8] * m[9]), g02 = fabsf(m[0] * m[2] + m[4] * m[6] + m[8] * m[10]),
                  g12 = fabsf(m[1] * m[2] + m[5] * m[6] + m[9] * m[10]);
      const float n2 = fmaxf(g00 + g01 + g02, fmaxf(g01 + g11 + g12, g02 + g12 + g22));
      const float s = sqrtf(n2) * 1.00001f;
      const bool precise = (E <= g.vslack) && (s <= 1.0e6f);     // false for NaN / Inf
      sc = precise ? s : -1.f;
    }
    sScale[tid] = sc;
  }
  __syncthreads();
  uint32_t imprec = 0;                            // bit c: candidate c takes the robust path (CTA-uniform)
#pragma unroll
  for (int c = 0; c < kCandPerBlock; ++c) imprec |= (sScale[c] < 0.f ? 1u : 0u) << c;

  ProbeStats st;
  uint16_t* const wq = &sQueue[(tid >> 5) * kWarpQueue];   // this warp's queue
  uint32_t qn = 0u;                               // entries waiting in it (warp-uniform)
  const long long qbase = (long long)(S4G_RASTER_Y ? blockIdx.y : blockIdx.x) * (kThreads * kTilesPerBlock);
  // ---- phase 0, two levels.  (a) one thread per (128-query tile, candidate): the tile's bounding sphere against the
  // summed-area table; survivors (~1 in 4) are compacted into a list.  (b) one thread per (surviving pair, warp of the
  // tile): the bounding sphere of that warp's 32 queries -- half the radius, an eighth of the box -- against the table
  // again.  sLive[tile * 4 + warp] = candidates that warp has to visit for that tile.
  static_assert(kCandPerBlock == 16 && kThreads == 128 && kVerifySub == 32 && kTilesPerBlock * kCandPerBlock == kThreads, "cull mapping");
  {
    const int t = tid >> 4, c = tid & 15;
    const long long tile0 = qbase + (long long)t * kThreads;
    bool live = false;
    if (tile0 < nQ && c < nc)
      live = ((imprec >> c) & 1u) ? true : tile_live(g, &sV[c * 12], __ldg(&tiles[tile0 / kThreads]), sScale[c]);
    const unsigned b = __ballot_sync(0xffffffffu, live);
    if (lane == 0) sPairN[tid >> 5] = (uint32_t)__popc(b);
    if (tid < kTilesPerBlock * (kThreads / 32)) sLive[tid] = 0u;
    __syncthreads();
    uint32_t before = 0, nLive = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      const uint32_t k = sPairN[w];
      before += w < (tid >> 5) ? k : 0u;
      nLive += k;
    }
    if (live) sPair[before + __popc(b & ((1u << lane) - 1u))] = (uint8_t)tid;     // tid == tile << 4 | candidate
    __syncthreads();
#pragma unroll 1
    for (uint32_t j = (uint32_t)tid; j < 4u * nLive; j += kThreads) {
      const uint32_t code = sPair[j >> 2], sub = (code >> 4) * 4u + (j & 3u), cc = code & 15u;
      const long long q0 = qbase + (long long)sub * kVerifySub;
      if (q0 < nQ) {
        const bool l2 = ((imprec >> cc) & 1u) ? true : tile_live(g, &sV[cc * 12], __ldg(&subs[q0 / kVerifySub]), sScale[cc]);
        if (l2) atomicOr(&sLive[sub], 1u << cc);
      }
    }
  }
  __syncthreads();
  int last_tile = 0;                              // last tile of this CTA that exists (CTA-uniform)
  for (int t = kTilesPerBlock - 1; t > 0; --t)
    if (qbase + (long long)t * kThreads < nQ) { last_tile = t; break; }
  unsigned long long cert = 0ull;                 // 4-bit counters: CERTAIN inliers of this thread's queries per candidate
#pragma unroll 1
  for (int t = 0; t <= last_tile; ++t) {
    const long long tile0 = qbase + (long long)t * kThreads;
    uint32_t live_mask = sLive[t * 4 + (tid >> 5)];                 // this warp's candidates (warp-uniform)
    if (kStats) st.culled += (lane == 0 && tile0 + (tid & ~31) < nQ) ? (unsigned long long)(nc - __popc(live_mask)) : 0ull;
    const bool last = t == last_tile;
    if (live_mask == 0u && !last) continue;       // warp-uniform: nothing to do for this warp's 32 queries (no CTA barrier below)
    const long long qi = tile0 + tid;
    const bool valid = qi < nQ;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid && live_mask) q = __ldg(&Q[qi]);

    // ---- phase 1: delta-field state per (query, candidate), live candidates only; two candidates per
    // iteration so that two table loads are in flight.  Fast loop: every live candidate passed the rounding bound
    // (finite, bounded coefficients), so the voxel comes from the FMA chain and the range test is an unsigned compare.
    uint32_t pend = 0u;                           // bit c: pair (this query, candidate c) needs the exact test
    const uint32_t limX = (uint32_t)g.nx << 2, limY = (uint32_t)g.ny << 2, limZ = (uint32_t)g.nz << 2;
    if ((live_mask & imprec) == 0u) {
      const uint32_t lx = valid ? limX : 0u;        // a lane without a query fails the range test of every candidate
#pragma unroll 1
      while (live_mask) {
        const int ca = __ffs(live_mask) - 1;
        live_mask &= live_mask - 1;
        const bool two = live_mask != 0u;
        const int cb = two ? __ffs(live_mask) - 1 : ca;
        live_mask &= live_mask - 1;                  // (0 stays 0)
        float ax, ay, az, bx, by, bz;
        voxel_of<false>(g, &sV[ca * 12], nullptr, q, ax, ay, az);
        voxel_of<false>(g, &sV[cb * 12], nullptr, q, bx, by, bz);
        const int aX = __float2int_rd(ax), aY = __float2int_rd(ay), aZ = __float2int_rd(az);
        const int bX = __float2int_rd(bx), bY = __float2int_rd(by), bZ = __float2int_rd(bz);
        const bool ina = (uint32_t)aX < lx && (uint32_t)aY < limY && (uint32_t)aZ < limZ;
        const bool inb = two && (uint32_t)bX < lx && (uint32_t)bY < limY && (uint32_t)bZ < limZ;
        const int ra = ina ? __ldg(&g.vtop[vtop_index<kBS>(g, aX, aY, aZ)]) : -1;
        const int rb = inb ? __ldg(&g.vtop[vtop_index<kBS>(g, bX, bY, bZ)]) : -1;
        if (kStats) st.bitmap += (ina ? 1 : 0) + (inb ? 1 : 0);
        if ((ra & rb) >= 0) {                        // some lane of the warp landed in a v-brick (for a or for b): both words in flight
          const uint32_t cella = vox_cell<kBS>(g, ra, aX, aY, aZ), cellb = vox_cell<kBS>(g, rb, bX, bY, bZ);
          const uint32_t sha = vox_shift(aX, aY), shb = vox_shift(bX, bY);
          const uint32_t wa = ra >= 0 ? __ldg(&g.vox[(cella << 2) | (uint32_t)(aZ & 3)]) : 0u;
          const uint32_t wb = rb >= 0 ? __ldg(&g.vox[(cellb << 2) | (uint32_t)(bZ & 3)]) : 0u;
          uint32_t sa = (wa >> sha) & 3u, sb = (wb >> shb) & 3u;
          if (kStats) st.bitmap += (ra >= 0 ? 1 : 0) + (rb >= 0 ? 1 : 0);
          if (S4G_SUBVOXEL && g.vfine != nullptr) {
            if (sa == 1u) sa = vox_refine(g, cella, wa, sha, aX, aY, aZ, ax, ay, az);
            if (sb == 1u) sb = vox_refine(g, cellb, wb, shb, bX, bY, bZ, bx, by, bz);
          }
          if (sa & 2u) cert += 1ull << (4 * ca);
          else if (sa & 1u) pend |= 1u << ca;
          if (sb & 2u) cert += 1ull << (4 * cb);
          else if (sb & 1u) pend |= 1u << cb;
        }
      }
    } else {
      // robust loop (a live candidate failed the bound: huge / non-rigid / NaN coefficients): voxel from the reference-order
      // T q, float range test (NaN and out-of-range coordinates fail it before any int conversion is used)
#pragma unroll 1
      while (live_mask) {
        const int c = __ffs(live_mask) - 1;
        live_mask &= live_mask - 1;
        float ux, uy, uz;
        voxel_of<true>(g, &sV[c * 12], &sT[c * 12], q, ux, uy, uz);
        const bool in = valid && ux >= 0.f && uy >= 0.f && uz >= 0.f && ux < (float)limX && uy < (float)limY && uz < (float)limZ;
        const int X = in ? __float2int_rd(ux) : 0, Y = in ? __float2int_rd(uy) : 0, Z = in ? __float2int_rd(uz) : 0;
        const int r = in ? __ldg(&g.vtop[vtop_index<kBS>(g, X, Y, Z)]) : -1;
        if (kStats) st.bitmap += in ? 1 : 0;
        if (r >= 0) {
          const uint32_t sa = vox_state<kBS>(g, r, X, Y, Z, ux, uy, uz);
          if (kStats) st.bitmap++;
          if (sa & 2u) cert += 1ull << (4 * c);
          else if (sa & 1u) pend |= 1u << c;
        }
      }
    }
    // ---- per-WARP queue: rounds of { compaction of pending pairs -> this warp's queue ; flush = phase 2 when enough
    // pairs wait, the queue is full, or this is the last tile }.  Warp-synchronous: no CTA barrier in the tile loop, the
    // four warps of the CTA drift apart freely (their live candidates differ).
    bool again;
    do {
      if (__any_sync(0xffffffffu, pend != 0u)) {
        const uint32_t cnt = (uint32_t)__popc(pend);
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t base = qn + incl - cnt;
        while (pend && base < (uint32_t)kWarpQueue) {
          const int c = __ffs(pend) - 1;
          pend &= pend - 1u;
          wq[base++] = (uint16_t)((c << 10) | (t << 7) | tid);
        }
        qn = min(qn + total, (uint32_t)kWarpQueue);
      }
      again = __any_sync(0xffffffffu, pend != 0u);                  // leftovers: the queue is full
      if (again || last || qn >= (uint32_t)kWarpFlush) {
        __syncwarp();                                               // the queue entries of all lanes are visible
        // ---- phase 2: one queued pair per lane, densely packed
#pragma unroll 1
        for (uint32_t i = (uint32_t)lane; i < qn; i += 32u) {
          const uint32_t e = wq[i];
          const int c = (int)(e >> 10);
          const float4 qq = __ldg(&Q[qbase + (long long)(e & 1023u)]);
          float tx, ty, tz;
          exact_tq(&sT[c * 12], qq, tx, ty, tz);
          // origin of the 2x2x2 cell block around t (conservative, see the header)
          const float ux = __fsub_rn(__fmul_rn(__fsub_rn(tx, g.ox), g.inv_h), 0.5f);
          const float uy = __fsub_rn(__fmul_rn(__fsub_rn(ty, g.oy), g.inv_h), 0.5f);
          const float uz = __fsub_rn(__fmul_rn(__fsub_rn(tz, g.oz), g.inv_h), 0.5f);
          const bool in = ux >= -1.f && uy >= -1.f && uz >= -1.f && ux < (float)g.nx && uy < (float)g.ny && uz < (float)g.nz;
          bool found = false;
          if (in) {
            const int x0 = __float2int_rd(ux), y0 = __float2int_rd(uy), z0 = __float2int_rd(uz);
            uint32_t nib = block_rows(g, x0, y0, z0);
            if (kStats) st.bitmap += 2;
#pragma unroll 1
            while (nib && !found) {
              const int r = __ffs(nib) - 1;
              nib &= nib - 1u;
              found = walk_row<kStats>(g, x0, y0, z0, r, tx, ty, tz, sq_eps, st);
            }
          }
          if (found) atomicAdd(&sCnt[c], 1u);
        }
        __syncwarp();                                               // every lane is done with the queue
        qn = 0u;
      }
    } while (again);
  }
  // CERTAIN inliers: 16 packed 4-bit counters per thread -> one warp reduction per candidate
#pragma unroll
  for (int c = 0; c < kCandPerBlock; ++c) {
    const uint32_t v = __reduce_add_sync(0xffffffffu, (uint32_t)(cert >> (4 * c)) & 15u);
    if (lane == 0 && v) atomicAdd(&sCnt[c], v);
  }
  __syncthreads();
  if (tid < nc) {
    const uint32_t v = sCnt[tid];
    if (v) atomicAdd(&counts[c0 + tid], v);
  }
  if (kStats) {
    atomicAdd(&stats[0], st.tested);
    atomicAdd(&stats[1], st.ranges);
    atomicAdd(&stats[2], st.bricks);
    atomicAdd(&stats[3], st.bitmap);
    atomicAdd(&stats[4], st.culled);   // (32-query sub-tile, candidate) pairs culled (lane 0 of every warp counted them)
  }
}

// column-major 4x4 (16 floats) -> row-major 3x4 (12 floats)
__global__ void k_pack_T12(const float* __restrict__ T16, int K, float* __restrict__ T12) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * 12) return;
  int k = i / 12, e = i % 12, r = e / 4, c = e % 4;
  T12[i] = T16[(size_t)k * 16 + c * 4 + r];
}

}  // namespace

// Enqueue Verify for K transforms given as row-major 3x4 (device). counts must be zeroed by us.
int s4g_launch_verify(s4g_ctx* ctx, const float* d_T12, int K, uint32_t* d_counts, bool timed, const uint32_t* d_K) {
  if (K <= 0) return S4G_OK;
  cudaStream_t st = ctx->stream;
  S4G_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)K * sizeof(uint32_t), st));
  float sq_eps = ctx->delta * ctx->delta;  // epsilon*epsilon, match4pcsBase.cc:522
  const int per_block = kThreads * kTilesPerBlock;
  dim3 grid((unsigned)((ctx->nQ + per_block - 1) / per_block), 1, 1);
  // gridDim.y is limited to 65535: loop over slabs of candidate chunks
  const int max_chunks = 65535;
  if (timed) S4G_EV_START(ctx, S4G_EV_VERIFY);
  for (int c0 = 0; c0 < K; c0 += max_chunks * kCandPerBlock) {
    int kk = (K - c0 < max_chunks * kCandPerBlock) ? (K - c0) : max_chunks * kCandPerBlock;
    grid.y = (unsigned)((kk + kCandPerBlock - 1) / kCandPerBlock);
    if (S4G_RASTER_Y) { const unsigned t = grid.x; grid.x = grid.y; grid.y = t; }
    if (ctx->grid.bshift == 2)
      k_verify<false, 2><<<grid, kThreads, 0, st>>>(ctx->grid, ctx->dQmorton.as<float4>(), ctx->dQtiles.as<float4>(), ctx->dQtiles.as<float4>() + (ctx->nQ + kVerifyTile - 1) / kVerifyTile,
                                                    ctx->nQ, d_T12 + (size_t)c0 * 12, kk, sq_eps, ctx->qabs[0],
                                                    ctx->qabs[1], ctx->qabs[2], d_counts + c0, nullptr, d_K);
    else
      k_verify<false, 0><<<grid, kThreads, 0, st>>>(ctx->grid, ctx->dQmorton.as<float4>(), ctx->dQtiles.as<float4>(), ctx->dQtiles.as<float4>() + (ctx->nQ + kVerifyTile - 1) / kVerifyTile,
                                                    ctx->nQ, d_T12 + (size_t)c0 * 12, kk, sq_eps, ctx->qabs[0],
                                                    ctx->qabs[1], ctx->qabs[2], d_counts + c0, nullptr, d_K);
    if (S4G_RASTER_Y) { const unsigned t = grid.x; grid.x = grid.y; grid.y = t; }
    ctx->launches++;
  }
  if (timed) S4G_EV_STOP(ctx, S4G_EV_VERIFY);
  S4G_CUDA(cudaGetLastError());
  return S4G_OK;
}

static int check_ready(s4g_ctx* ctx, const char* who) {
  if (ctx->nP <= 0 || ctx->nQ <= 0) {
    ctx->err = std::string(who) + ": call s4g_set_cloud_p and s4g_set_cloud_q first";
    return S4G_ERR_STATE;
  }
  return S4G_OK;
}

extern "C" int s4g_verify_dev(s4g_ctx* ctx, const float* d_T, int K, uint32_t* d_counts) {
  if (!ctx) return S4G_ERR_ARG;
  if (K < 0 || (K > 0 && (!d_T || !d_counts))) { ctx->err = "s4g_verify_dev: bad arguments"; return S4G_ERR_ARG; }
  S4G_TRY(check_ready(ctx, "s4g_verify_dev"));
  if (K == 0) return S4G_OK;
  S4G_CUDA(cudaSetDevice(ctx->device));
  S4G_TRY(s4g_reserve(ctx, ctx->dT12, (size_t)K * 12 * sizeof(float)));
  k_pack_T12<<<(K * 12 + 255) / 256, 256, 0, ctx->stream>>>(d_T, K, ctx->dT12.as<float>());
  ctx->launches++;
  return s4g_launch_verify(ctx, ctx->dT12.as<float>(), K, d_counts, true, nullptr);
}

extern "C" int s4g_verify(s4g_ctx* ctx, const float* T, int K, uint32_t* counts) {
  if (!ctx) return S4G_ERR_ARG;
  if (K < 0 || (K > 0 && (!T || !counts))) { ctx->err = "s4g_verify: bad arguments"; return S4G_ERR_ARG; }
  S4G_TRY(check_ready(ctx, "s4g_verify"));
  if (K == 0) return S4G_OK;
  S4G_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  S4G_TRY(s4g_reserve(ctx, ctx->dScratchA, (size_t)K * 16 * sizeof(float)));
  S4G_TRY(s4g_reserve(ctx, ctx->dCounts, (size_t)K * sizeof(uint32_t)));
  S4G_CUDA(cudaMemcpyAsync(ctx->dScratchA.p, T, (size_t)K * 16 * sizeof(float), cudaMemcpyHostToDevice, st));
  S4G_TRY(s4g_verify_dev(ctx, ctx->dScratchA.as<float>(), K, ctx->dCounts.as<uint32_t>()));
  S4G_CUDA(cudaMemcpyAsync(counts, ctx->dCounts.p, (size_t)K * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  S4G_CUDA(cudaStreamSynchronize(st));
  return S4G_OK;
}

extern "C" int s4g_verify_probe_stats(s4g_ctx* ctx, const float* T, int K, uint64_t* out5) {
  if (!ctx) return S4G_ERR_ARG;
  if (K <= 0 || !T || !out5) { ctx->err = "s4g_verify_probe_stats: bad arguments"; return S4G_ERR_ARG; }
  S4G_TRY(check_ready(ctx, "s4g_verify_probe_stats"));
  S4G_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  S4G_TRY(s4g_reserve(ctx, ctx->dScratchA, (size_t)K * 16 * sizeof(float)));
  S4G_TRY(s4g_reserve(ctx, ctx->dCounts, (size_t)K * sizeof(uint32_t)));
  S4G_TRY(s4g_reserve(ctx, ctx->dT12, (size_t)K * 12 * sizeof(float)));
  S4G_TRY(s4g_reserve(ctx, ctx->dMisc, 256));
  S4G_CUDA(cudaMemcpyAsync(ctx->dScratchA.p, T, (size_t)K * 16 * sizeof(float), cudaMemcpyHostToDevice, st));
  S4G_CUDA(cudaMemsetAsync(ctx->dMisc.p, 0, 40, st));
  S4G_CUDA(cudaMemsetAsync(ctx->dCounts.p, 0, (size_t)K * sizeof(uint32_t), st));
  k_pack_T12<<<(K * 12 + 255) / 256, 256, 0, st>>>(ctx->dScratchA.as<float>(), K, ctx->dT12.as<float>());
  const int per_block = kThreads * kTilesPerBlock;
  dim3 grid((unsigned)((ctx->nQ + per_block - 1) / per_block), (unsigned)((K + kCandPerBlock - 1) / kCandPerBlock), 1);
  if (grid.y > 65535) { ctx->err = "s4g_verify_probe_stats: K too large"; return S4G_ERR_ARG; }
  if (S4G_RASTER_Y) { const unsigned t = grid.x; grid.x = grid.y; grid.y = t; }
  k_verify<true, 0><<<grid, kThreads, 0, st>>>(ctx->grid, ctx->dQmorton.as<float4>(), ctx->dQtiles.as<float4>(), ctx->dQtiles.as<float4>() + (ctx->nQ + kVerifyTile - 1) / kVerifyTile, ctx->nQ,
                                            ctx->dT12.as<float>(), K, ctx->delta * ctx->delta, ctx->qabs[0], ctx->qabs[1],
                                            ctx->qabs[2], ctx->dCounts.as<uint32_t>(), ctx->dMisc.as<unsigned long long>(), nullptr);
  ctx->launches += 2;
  S4G_CUDA(cudaGetLastError());
  unsigned long long h[5] = {0, 0, 0, 0, 0};
  S4G_CUDA(cudaMemcpyAsync(h, ctx->dMisc.p, 40, cudaMemcpyDeviceToHost, st));
  S4G_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < 5; ++i) out5[i] = h[i];
  return S4G_OK;
}
