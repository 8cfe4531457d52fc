// a8 -- Match4PCSBase::Verify (reference algorithms/match4pcsBase.cc:508-567) for a batch of
// candidate transforms, on the bricked uniform grid built by s4g_set_cloud_p.
//
//   counts[k] = #{ q in sampled_Q : exists p in sampled_P, ||T_k q - p||^2 <= delta^2 }
//
// Arithmetic of the DECISION follows the reference's binary operation by operation (SURVEY.md
// A.3 / B.3):
//   T q  = ((m0 x + m1 y) + m2 z) + m3   per row, fp32, no FMA      (match4pcsBase.cc:532)
//   d^2  = dx^2 + (dy^2 + dz^2)                                     (kdtree.h:417)
//   hit  = d^2 <= delta*delta                                       (kdtree.h:418, cc:522)
//
// Schedule (B200).  The working set (Q 16 MB, sorted P 16 MB, cellStart, brick table, occupancy
// bitmap: ~45 MB at 1M points) is L2-resident; the kernel is instruction-issue bound, so the
// design minimises instructions per (query, candidate) pair:
//  * sampled_Q is streamed in Morton order, one coalesced float4 per thread per tile, kept in
//    registers and in a shared-memory tile; a CTA stages a chunk of 16 candidate transforms.
//  * phase 0 (one thread per (tile, candidate)): the tile's bounding sphere, transformed and grown
//    by delta, is tested against the summed-area table of the coarse occupancy (2x2x2-cell blocks at
//    1M points, 8 look-ups, no loop); most wrong candidates
//    miss P entirely over most tiles and cost nothing further.
//  * phase 1 (cheap, every surviving pair): the CELL-space image u = U q (U = the transform
//    pre-multiplied by the world->cell map, 9 FMAs -- used only to pick cells, never for the
//    decision) gives the origin of the 2x2x2 cell block that must contain every P point within
//    delta; a 4-bit entry of the occupancy map tells which of the block's four x-rows (two
//    x-adjacent cells each) hold points.  Each thread collects its live (candidate, row) bits in
//    a 64-bit register; after the candidate loop they are compacted into a shared-memory queue
//    (one warp scan + one shared atomic per warp).
//  * phase 2 (dense, one queue entry = one non-empty row per thread): exact T q, the row's one or
//    two contiguous point runs, d^2 <= delta^2; hits set a bit per (candidate, query) in shared
//    memory (a query can hit in several rows), the bits are counted per candidate at the end of
//    the tile, one global atomicAdd per CTA per candidate at the end of the CTA.
//
// Probe: cell edge h >= 2.02*delta, so the delta-ball around T q touches at most 2 cells per
// axis: x0 = floor(u - 0.5), cells {x0, x0+1}.  For a point p with |T q - p|_x <= delta(1+1e-6):
// |u - v| <= 0.4951, u in [x0+0.5, x0+1.5) => v in (x0+0.0049, x0+1.9951); the 0.0049-cell margin
// dominates the rounding of u (FMA chain vs exact: < 1e-3 cell for grids <= 2048 cells per axis)
// and of v, so the block is conservative and the count exact.
#include "s4g_internal.cuh"

namespace {

constexpr int kThreads = kVerifyTile;   // 128: one query per thread per tile
constexpr int kCandPerBlock = 16;   // transforms staged per CTA
constexpr int kTilesPerBlock = kThreads / 16;  // (tile, candidate) pairs of a CTA = one per thread in the cull phase
// Compile-time knobs for A/B runs on the GPU (scripts/verify_ab.sh rebuilds libs4g with S4G_NVCC_DEFINES and re-runs
// parity + bench).  The defaults are the measured round-1 configuration; a default build is unchanged.
#ifndef S4G_QUEUE_CAP
#define S4G_QUEUE_CAP 3072
#endif
#ifndef S4G_VERIFY_MIN_BLOCKS
#define S4G_VERIFY_MIN_BLOCKS 12
#endif
constexpr int kQueueCap = S4G_QUEUE_CAP;   // queue entries per round (a tile with more live rows takes extra rounds)
static_assert(kQueueCap >= kThreads && kQueueCap <= 16384, "queue capacity (uint16 entries in shared memory)");

struct ProbeStats {
  unsigned long long tested = 0, ranges = 0, bricks = 0, bitmap = 0, culled = 0;
};

// Scan one contiguous run of P points; single exit, flag based.  Two points are in flight per
// iteration (the loads of a run are independent; the kernel is latency-bound here).
template <bool kStats>
__device__ __forceinline__ bool probe_run(const GridDev& g, uint32_t s, uint32_t e, float tx, float ty,
                                          float tz, float sq_eps, ProbeStats& st) {
  bool found = false;
#ifdef S4G_PROBE4
  // variant: four points in flight per iteration (the index is clamped to the run's last point: re-testing it cannot
  // change the answer).  Same decision arithmetic, fewer dependent iterations for the lanes with long runs.
  for (uint32_t k = s; k < e && !found; k += 4) {
    const uint32_t last = e - 1u;
    const float4 p0 = __ldg(&g.pts[k]);
    const float4 p1 = __ldg(&g.pts[min(k + 1u, last)]);
    const float4 p2 = __ldg(&g.pts[min(k + 2u, last)]);
    const float4 p3 = __ldg(&g.pts[min(k + 3u, last)]);
    const float ax = __fsub_rn(tx, p0.x), ay = __fsub_rn(ty, p0.y), az = __fsub_rn(tz, p0.z);
    const float bx = __fsub_rn(tx, p1.x), by = __fsub_rn(ty, p1.y), bz = __fsub_rn(tz, p1.z);
    const float cx = __fsub_rn(tx, p2.x), cy = __fsub_rn(ty, p2.y), cz = __fsub_rn(tz, p2.z);
    const float dx = __fsub_rn(tx, p3.x), dy = __fsub_rn(ty, p3.y), dz = __fsub_rn(tz, p3.z);
    const float a2 = __fadd_rn(__fmul_rn(ax, ax), __fadd_rn(__fmul_rn(ay, ay), __fmul_rn(az, az)));
    const float b2 = __fadd_rn(__fmul_rn(bx, bx), __fadd_rn(__fmul_rn(by, by), __fmul_rn(bz, bz)));
    const float c2 = __fadd_rn(__fmul_rn(cx, cx), __fadd_rn(__fmul_rn(cy, cy), __fmul_rn(cz, cz)));
    const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dz, dz)));
    if (kStats) st.tested += min(4u, e - k);
    found = a2 <= sq_eps || b2 <= sq_eps || c2 <= sq_eps || d2 <= sq_eps;
  }
  return found;
#endif
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

// Address of the occupancy nibble of block origin (ox,oy,oz) (already shifted by +1): the map is
// tiled in 4x4x4-origin bricks = 64 nibbles = one 32-byte sector, so the handful of neighbouring
// origins a warp touches share a sector instead of spanning one cache line per (y,z) row.
__device__ __forceinline__ uint32_t occ_index(const GridDev& g, uint32_t ox, uint32_t oy, uint32_t oz) {
  const uint32_t t = ((oz >> 2) * (uint32_t)g.oty + (oy >> 2)) * (uint32_t)g.otx + (ox >> 2);
  return (t << 6) | ((oz & 3u) << 4) | ((oy & 3u) << 2) | (ox & 3u);
}

// origin of the 2x2x2 block in cell space: floor(U q) with U carrying the -0.5 shift
__device__ __forceinline__ void block_origin(const float* __restrict__ u, float4 q, int& x0, int& y0, int& z0) {
  x0 = __float2int_rd(__fmaf_rn(u[0], q.x, __fmaf_rn(u[1], q.y, __fmaf_rn(u[2], q.z, u[3]))));
  y0 = __float2int_rd(__fmaf_rn(u[4], q.x, __fmaf_rn(u[5], q.y, __fmaf_rn(u[6], q.z, u[7]))));
  z0 = __float2int_rd(__fmaf_rn(u[8], q.x, __fmaf_rn(u[9], q.y, __fmaf_rn(u[10], q.z, u[11]))));
}

// Tile-level cull: can ANY point of a query tile (bounding sphere `sph`, world units) come within
// delta of a P point under the candidate whose cell-space matrix is `u`?  Conservative test on the
// summed-area table of the coarse occupancy ((2^cshift)^3-cell blocks): the transformed sphere, grown by delta, is
// boxed and the number of occupied blocks the box touches follows from 8 table look-ups.  Run by ONE thread per (tile, candidate).
__device__ bool tile_live(const GridDev& g, const float* __restrict__ u, float4 sph) {
  // single exit, flag based (see DESIGN.md section 7 on early returns + warp votes)
  bool live = true;
  if (g.csat != nullptr) {
    // centre in cell coordinates (u carries a -0.5 shift), radius in cells: r/h + delta/h (< 0.4951) + slack
    const float cx = __fmaf_rn(u[0], sph.x, __fmaf_rn(u[1], sph.y, __fmaf_rn(u[2], sph.z, u[3]))) + 0.5f;
    const float cy = __fmaf_rn(u[4], sph.x, __fmaf_rn(u[5], sph.y, __fmaf_rn(u[6], sph.z, u[7]))) + 0.5f;
    const float cz = __fmaf_rn(u[8], sph.x, __fmaf_rn(u[9], sph.y, __fmaf_rn(u[10], sph.z, u[11]))) + 0.5f;
    const float R = sph.w * g.inv_h * 1.0001f + 0.52f;
    const float fx0 = cx - R, fx1 = cx + R, fy0 = cy - R, fy1 = cy + R, fz0 = cz - R, fz1 = cz + R;
    // NaN transforms and boxes entirely outside the grid cannot match anything
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

// T12: K x 12 floats, row-major 3x4 (r00 r01 r02 t0 | r10 ... ), the top three rows of T.
// grid.x = query super-tiles (kThreads*kTilesPerBlock queries), grid.y = candidate chunks.
template <bool kStats>
__global__ void __launch_bounds__(kThreads, kStats ? 1 : S4G_VERIFY_MIN_BLOCKS)
k_verify(GridDev g, const float4* __restrict__ Q, const float4* __restrict__ tiles, int nQ,
         const float* __restrict__ T12, int K, float sq_eps, uint32_t* __restrict__ counts,
         unsigned long long* __restrict__ stats) {
  __shared__ __align__(16) float sT[kCandPerBlock * 12];     // exact transforms (decision arithmetic)
  __shared__ __align__(16) float sU[kCandPerBlock * 12];     // cell-space transforms (cell selection only)
  __shared__ uint32_t sLive[kTilesPerBlock];                 // bit c: candidate c may hit tile t
  __shared__ uint32_t sCnt[kCandPerBlock];
  __shared__ float4 sQ[kThreads];
  __shared__ uint16_t sQueue[kQueueCap];                        // (candidate, row, query) entries of one round
  __shared__ uint32_t sHit[kCandPerBlock * (kThreads / 32)];    // one bit per (candidate, query)
  __shared__ uint32_t sQn[2];                                   // entry counters, alternating per round
  const int tid = threadIdx.x, lane = tid & 31;
  const int c0 = blockIdx.y * kCandPerBlock;
  const int nc = min(kCandPerBlock, K - c0);
  for (int i = tid; i < nc * 12; i += kThreads) {
    const float t = T12[(size_t)c0 * 12 + i];
    sT[i] = t;
    const int col = i & 3, row = (i % 12) >> 2;
    const float o = row == 0 ? g.ox : row == 1 ? g.oy : g.oz;
    sU[i] = col < 3 ? t * g.inv_h : (t - o) * g.inv_h - 0.5f;
  }
  if (tid < kCandPerBlock) sCnt[tid] = 0;
  if (tid < kCandPerBlock * (kThreads / 32)) sHit[tid] = 0;
  if (tid < 2) sQn[tid] = 0;
  __syncthreads();

  ProbeStats st;
  uint32_t rnd = 0;                               // CTA-uniform round counter (selects sQn[rnd & 1])
  const long long qbase = (long long)blockIdx.x * (kThreads * kTilesPerBlock);
  // ---- phase 0: cull whole (tile, candidate) pairs on the coarse occupancy; one thread per pair
  static_assert(kCandPerBlock == 16 && kTilesPerBlock * kCandPerBlock == kThreads, "cull mapping: one thread per pair");
  {
    const int t = tid >> 4, c = tid & 15;
    const long long tile0 = qbase + (long long)t * kThreads;
    bool live = false;
    if (tile0 < nQ && c < nc) live = tile_live(g, &sU[c * 12], __ldg(&tiles[tile0 / kThreads]));
    const unsigned b = __ballot_sync(0xffffffffu, live);
    if ((tid & 31) == 0) {
      sLive[2 * (tid >> 5)] = b & 0xFFFFu;
      sLive[2 * (tid >> 5) + 1] = b >> 16;
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int t = 0; t < kTilesPerBlock; ++t) {
    const long long tile0 = qbase + (long long)t * kThreads;
    if (tile0 >= nQ) break;                       // CTA-uniform
    uint32_t live_mask = sLive[t];
    if (kStats) st.culled += (unsigned long long)(nc - __popc(live_mask));
    if (live_mask == 0u) continue;                // CTA-uniform: the whole tile is culled
    const long long qi = tile0 + tid;
    const bool valid = qi < nQ;
    const float4 q = valid ? __ldg(&Q[qi]) : make_float4(0.f, 0.f, 0.f, 0.f);
    sQ[tid] = q;                                  // (the previous tile's readers passed its last barrier)

    // ---- phase 1: one occupancy nibble per (query, candidate), live candidates only; two
    // candidates per iteration so that two map loads are in flight
    unsigned long long rows = 0ull;               // bit 4c + r: row r of candidate c's block holds points
#pragma unroll 1
    while (live_mask) {
      const int ca = __ffs(live_mask) - 1;
      live_mask &= live_mask - 1;
      const bool two = live_mask != 0u;
      const int cb = two ? __ffs(live_mask) - 1 : ca;
      live_mask &= live_mask - 1;                  // (0 stays 0)
      int xa, ya, za, xb, yb, zb;
      block_origin(&sU[ca * 12], q, xa, ya, za);
      block_origin(&sU[cb * 12], q, xb, yb, zb);
      const bool ina = valid && (unsigned)(xa + 1) <= (unsigned)g.nx && (unsigned)(ya + 1) <= (unsigned)g.ny &&
                       (unsigned)(za + 1) <= (unsigned)g.nz;
      const bool inb = two && valid && (unsigned)(xb + 1) <= (unsigned)g.nx && (unsigned)(yb + 1) <= (unsigned)g.ny &&
                       (unsigned)(zb + 1) <= (unsigned)g.nz;
      uint32_t na = ina ? 0xFu : 0u, nb = inb ? 0xFu : 0u;
      if (g.occ != nullptr) {
        const uint32_t oa = ina ? occ_index(g, (uint32_t)(xa + 1), (uint32_t)(ya + 1), (uint32_t)(za + 1)) : 0u;
        const uint32_t ob = inb ? occ_index(g, (uint32_t)(xb + 1), (uint32_t)(yb + 1), (uint32_t)(zb + 1)) : 0u;
        const uint32_t wa = ina ? __ldg(&g.occ[oa >> 3]) : 0u;
        const uint32_t wb = inb ? __ldg(&g.occ[ob >> 3]) : 0u;
        na = (wa >> ((oa & 7u) * 4u)) & 0xFu;
        nb = (wb >> ((ob & 7u) * 4u)) & 0xFu;
        if (kStats) st.bitmap += (ina ? 1 : 0) + (inb ? 1 : 0);
      }
      rows |= (unsigned long long)na << (4 * ca);
      rows |= (unsigned long long)nb << (4 * cb);
    }
    // ---- rounds of { compaction of live (candidate, row) bits -> queue ; phase 2 }.  One round
    // unless the tile has more than kQueueCap live rows.
    bool more;
    do {
      uint32_t* const qn = &sQn[rnd & 1u];
      {
        const uint32_t cnt = (uint32_t)__popcll(rows);
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t base = 0;
        if (lane == 31 && total) base = atomicAdd(qn, total);
        base = __shfl_sync(0xffffffffu, base, 31) + incl - cnt;
        while (rows && base < (uint32_t)kQueueCap) {
          const int bit = __ffsll((long long)rows) - 1;
          rows &= rows - 1ull;
          sQueue[base++] = (uint16_t)((bit << 8) | tid);   // tid < 256
        }
      }
      __syncthreads();                            // queue (and, in the first round, sQ) complete
      const uint32_t n = min(*qn, (uint32_t)kQueueCap);
      if (tid == 0) sQn[(rnd + 1u) & 1u] = 0;      // the other counter is idle until the next round

      // ---- phase 2: one non-empty row per thread, densely packed
#pragma unroll 1
      for (uint32_t i = (uint32_t)tid; i < n; i += kThreads) {
        const uint32_t e = sQueue[i];
        const int c = (int)(e >> 10), r = (int)((e >> 8) & 3u), qid = (int)(e & 0xFFu);
        const float4 qq = sQ[qid];
        const float* m = &sT[c * 12];
        // ((m0 x + m1 y) + m2 z) + m3
        const float tx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], qq.x), __fmul_rn(m[1], qq.y)), __fmul_rn(m[2], qq.z)), m[3]);
        const float ty = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[4], qq.x), __fmul_rn(m[5], qq.y)), __fmul_rn(m[6], qq.z)), m[7]);
        const float tz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[8], qq.x), __fmul_rn(m[9], qq.y)), __fmul_rn(m[10], qq.z)), m[11]);
        int x0, y0, z0;
        block_origin(&sU[c * 12], qq, x0, y0, z0);
        if (walk_row<kStats>(g, x0, y0, z0, r, tx, ty, tz, sq_eps, st))
          atomicOr(&sHit[c * (kThreads / 32) + (qid >> 5)], 1u << (qid & 31));
      }
      more = __syncthreads_or(rows != 0ull) != 0;  // barrier: phase 2 done with the queue / sQ / sHit
      ++rnd;
    } while (more);
    // count and clear the hit bits of this tile (the next writers of sHit are two barriers away)
    if (tid < kCandPerBlock * (kThreads / 32)) {
      const uint32_t h = sHit[tid];
      if (h) {
        atomicAdd(&sCnt[tid / (kThreads / 32)], (uint32_t)__popc(h));
        sHit[tid] = 0;
      }
    }
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
    if (tid == 0) atomicAdd(&stats[4], st.culled);   // (tile, candidate) pairs culled, counted once per CTA
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
int s4g_launch_verify(s4g_ctx* ctx, const float* d_T12, int K, uint32_t* d_counts, bool timed) {
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
    k_verify<false><<<grid, kThreads, 0, st>>>(ctx->grid, ctx->dQmorton.as<float4>(), ctx->dQtiles.as<float4>(),
                                               ctx->nQ, d_T12 + (size_t)c0 * 12, kk, sq_eps, d_counts + c0, nullptr);
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
  return s4g_launch_verify(ctx, ctx->dT12.as<float>(), K, d_counts, true);
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
  k_verify<true><<<grid, kThreads, 0, st>>>(ctx->grid, ctx->dQmorton.as<float4>(), ctx->dQtiles.as<float4>(), ctx->nQ,
                                            ctx->dT12.as<float>(), K, ctx->delta * ctx->delta,
                                            ctx->dCounts.as<uint32_t>(), ctx->dMisc.as<unsigned long long>());
  ctx->launches += 2;
  S4G_CUDA(cudaGetLastError());
  unsigned long long h[5] = {0, 0, 0, 0, 0};
  S4G_CUDA(cudaMemcpyAsync(h, ctx->dMisc.p, 40, cudaMemcpyDeviceToHost, st));
  S4G_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < 5; ++i) out5[i] = h[i];
  return S4G_OK;
}
