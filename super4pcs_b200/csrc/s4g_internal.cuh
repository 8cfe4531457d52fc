// Internal declarations shared by the .cu files of libs4g.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "s4g.h"

#define S4G_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      char b__[512];                                                                     \
      snprintf(b__, sizeof b__, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,            \
               cudaGetErrorString(e__));                                                 \
      ctx->err = b__;                                                                    \
      return S4G_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define S4G_TRY(expr)                 \
  do {                                \
    int rc__ = (expr);                \
    if (rc__ != S4G_OK) return rc__;  \
  } while (0)

// grow-only device buffer
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  template <class T> T* as() const { return static_cast<T*>(p); }
};

// Uniform grid over the centred sampled P, bricked: a dense top-level table of bricks
// (edge 2^bshift cells) holds the rank of each occupied brick; occupied bricks own a dense
// block of cellStart entries; P is sorted by (brick rank, local cell) so every cell -- and
// every x-adjacent cell pair inside a brick -- is one contiguous run of float4 points.
struct GridDev {
  float ox, oy, oz;     // world coordinate of cell (0,0,0)'s low corner
  float inv_h;          // 1 / cell edge
  int nx, ny, nz;       // extent in cells
  int bshift;           // log2(brick edge in cells)
  int tbx, tby, tbz;    // extent in bricks
  const int* top;       // [tbx*tby*tbz] brick rank or -1
  const uint32_t* cellStart;  // [(nBricks << 3*bshift) + 1]
  const float4* pts;    // sorted points, w = original index (bit pattern)
  const uint32_t* csat; // summed-area table of the coarse occupancy ((2^cshift)^3-cell blocks):
                        // csat[(Z*(cny+1)+Y)*(cnx+1)+X] = #occupied blocks with x<X, y<Y, z<Z; or nullptr
  int cnx, cny, cnz, cshift;
  // ---- delta-field (verify.cu phase 1): 2 bits per voxel of edge h/4 (4x4x4 voxels per cell), stored per "v-brick"
  // (the bricks of `top`'s lattice that hold at least one voxel within delta of a P point): bit 0 = MAYBE (some P
  // point may lie within delta of some location of the voxel), bit 1 = CERTAIN (one P point lies within delta of
  // EVERY location of the voxel).  Neither bit: no P point within delta of any location of the voxel.
  const int* vtop;      // [tbx*tby*tbz] v-brick rank or -1
  const uint32_t* vox;  // [nVBricks << (3*bshift + 2)] words: (rank << (3*bshift) | local cell) * 4 + (vz & 3); bits 2*((vy&3)*4 + (vx&3))
  // second level, for the BOUNDARY voxels (MAYBE but not CERTAIN) only: the same two bits for each of the voxel's 2x2x2
  // sub-voxels (edge h/8 ~ delta/4).  vbase[cell] = slot of the cell's first boundary voxel (cells in v-brick order, voxels
  // in bit order of the cell's 4 words); vfine[slot] = MAYBE bits of the 8 children (bits 0-7, child = sx | sy << 1 | sz << 2)
  // | CERTAIN bits (bits 8-15).
  // occupancy of the 2x2x2-cell blocks the exact test probes: 4 bits per block ORIGIN cell (which of the block's four x-rows
  // hold points), stored for the cells of the v-bricks only (every origin of a non-empty block lies in one): nibble
  // (rank << 3*bshift | local cell) of vocc
  const uint32_t* vocc;
  const uint32_t* vbase;
  const uint16_t* vfine;
  float inv_v;          // 4 * inv_h (voxels per world unit)
  float vslack;         // world-unit uncertainty of a query's voxel position the field was built to tolerate
};

struct s4g_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  std::string err;
  int sm_count = 148;

  // ---- P side
  int nP = 0;
  float delta = 0.f;
  float cell_h = 0.f;
  GridDev grid{};
  long long nBricks = 0, nCells = 0;
  DevBuf dP, dPsorted, dTop, dCellStart, dCsat, dVtop, dVox, dVocc, dVbase, dVfine;
  long long nVBricks = 0, nVBoundary = 0;

  // ---- Q side
  int nQ = 0;
  DevBuf dQ;        // float4 original order (w = index bits)
  DevBuf dQmorton;  // float4 Morton order (w = original index bits)
  DevBuf dQn;       // float4 normals (w = 0)
  DevBuf dQrgb;     // float4 rgb (w = 0)
  DevBuf dQunit;    // float4 unit-cube coordinates (pairCreationFunctor.h:66-70)
  DevBuf dQmside;   // Morton-ordered copies of unit coordinates | normals | rgb (3 x n float4) for the pair predicate
  DevBuf dQtiles;   // bounding spheres (centre, radius): [nTiles] of every kVerifyTile consecutive Morton points, then [nSubs] of every kVerifySub
  DevBuf dQgroups;  // AABBs of the 64-point groups / 64-group supergroups of the Morton order
  bool pair_index_ready = false;
  bool q_has_normals = false, q_has_rgb = false;
  float qabs[3] = {0, 0, 0};   // largest |coordinate| of sampled Q per axis (rounding bound of the cell-space transform)
  float gcenter[3] = {0, 0, 0};
  float ratio = 1.f;

  // ---- pairs / quads
  DevBuf dPairs[2];
  long long nPairs[2] = {0, 0};
  bool pairs_sorted[2] = {false, false};
  DevBuf dQuads;
  long long nQuads = 0;

  // ---- several bases per launch chain (s4g_try_bases): shared lists keyed by the base index
  DevBuf bArgs, bCounts, bPairKeys[2], bQKeys[2], bQVals[2], bQCnt, bQuadKeys[2], bQuads, bMisc, bResults;

  // ---- scratch
  DevBuf dScratchA, dScratchB, dScratchC, dScratchD, dCub;
  DevBuf dT12, dRms, dOk, dCandIdx, dCounts, dResult, dMisc;
  void* hPinned = nullptr;  // small pinned staging block
  size_t hPinnedBytes = 0;

  // ---- timing
  // event pairs: 0 = Verify, 1 = rigid fit, 2 = pair extraction, 3 = quad extraction
  cudaEvent_t ev[4][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  bool ev_pending[4] = {false, false, false, false};
  double ms[4] = {0, 0, 0, 0};
  unsigned long long launches = 0;

  // ---- communicator of the sharded candidate set (comm.cu; null = the caller merges the shards)
  void* comm = nullptr;  // ncclComm_t
  int comm_ranks = 1, comm_rank = 0;
  int comm_timeout_s = 60;
  bool stuck = false;    // a timed-out collective whose stream never drained: s4g_destroy does not wait for it
  unsigned long long collectives = 0;
};

// queries per Verify tile (= threads per Verify CTA)
constexpr int kVerifyTile = 128;
// queries per cull unit (= one warp of a Verify CTA); s4g_set_cloud_q pre-computes one bounding sphere per unit
constexpr int kVerifySub = 32;

int s4g_reserve(s4g_ctx* ctx, DevBuf& b, size_t bytes);

// comm.cu: the reduction of the shards' winners on the device (active when a communicator is attached and shard_world > 1)
bool s4g_comm_active(const s4g_ctx* ctx, int shard_world);
int s4g_comm_check_shard(s4g_ctx* ctx, int shard_rank, int shard_world);
int s4g_comm_max_u64(s4g_ctx* ctx, const unsigned long long* d_in, unsigned long long* d_out, cudaStream_t st);
int s4g_comm_reduce_result(s4g_ctx* ctx, const unsigned long long* d_local, unsigned long long* d_global,
                           s4g_tcs_result* rec, cudaStream_t st);
int s4g_comm_wait(s4g_ctx* ctx, cudaStream_t st);  // stream synchronize, with the communicator's deadline when one is attached

// host-side state of one s4g_try_bases call (the three stages live next to the kernels they share code with)
constexpr int kBatchMaxBases = 64;
constexpr int kBatchIdBits = 26;                 // point / list indices inside the packed 64-bit keys
constexpr int kBatchSegShift = 2 * kBatchIdBits; // (segment or base) << 52 | first << 26 | second
struct BatchHost {
  int B = 0;
  unsigned long long nPairs = 0;                 // ordered pairs of all 2B extractions
  uint32_t segCount[2 * kBatchMaxBases] = {};    // ... per extraction (segment 2b + slot)
  uint32_t segOff[2 * kBatchMaxBases + 1] = {};  // exclusive prefix
  unsigned long long nPPairs = 0;                // pairs of the even (P-pair) segments
  unsigned long long nQuads = 0;
  uint32_t quadOff[kBatchMaxBases + 1] = {};     // first quad of every base in the shared quad list
};
int s4g_batch_pairs(s4g_ctx* ctx, const s4g_base_desc* bases, int B, float eps, const s4g_pair_filters* f, BatchHost& bh);
int s4g_batch_quads(s4g_ctx* ctx, const s4g_base_desc* bases, float thr2, BatchHost& bh);
int s4g_batch_tcs(s4g_ctx* ctx, const s4g_base_desc* bases, float max_angle_deg, float rms_threshold, BatchHost& bh,
                  s4g_base_result* out);
enum { S4G_EV_VERIFY = 0, S4G_EV_RIGID = 1, S4G_EV_PAIRS = 2, S4G_EV_QUADS = 3 };
// record the start / stop event of a timed kernel group on the context's stream
#define S4G_EV_START(ctx, which) S4G_CUDA(cudaEventRecord((ctx)->ev[which][0], (ctx)->stream))
#define S4G_EV_STOP(ctx, which)                                          \
  do {                                                                   \
    S4G_CUDA(cudaEventRecord((ctx)->ev[which][1], (ctx)->stream));       \
    (ctx)->ev_pending[which] = true;                                     \
  } while (0)

// ---- device helpers: the reference's float arithmetic, operation by operation -------------
// The whole library is compiled with -fmad=false; the _rn intrinsics below additionally
// make the association order explicit where parity with the reference depends on it.
__device__ __forceinline__ float s4_sum3(float a, float b, float c) {
  // Eigen's redux of a fixed 3-vector: a0 + (a1 + a2)
  return __fadd_rn(a, __fadd_rn(b, c));
}
__device__ __forceinline__ float s4_dot(float3 a, float3 b) {
  return s4_sum3(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ float s4_sqnorm(float3 a) { return s4_dot(a, a); }
__device__ __forceinline__ float3 s4_sub(float3 a, float3 b) {
  return make_float3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z));
}
__device__ __forceinline__ float3 s4_add(float3 a, float3 b) {
  return make_float3(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z));
}
__device__ __forceinline__ float3 s4_scale(float s, float3 a) {
  return make_float3(__fmul_rn(s, a.x), __fmul_rn(s, a.y), __fmul_rn(s, a.z));
}
__device__ __forceinline__ float3 s4_div(float3 a, float s) {
  return make_float3(__fdiv_rn(a.x, s), __fdiv_rn(a.y, s), __fdiv_rn(a.z, s));
}
__device__ __forceinline__ float3 s4_cross(float3 a, float3 b) {
  return make_float3(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)),
                     __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
                     __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
// MatrixBase::normalized(): n / sqrt(squaredNorm) when squaredNorm > 0
__device__ __forceinline__ float3 s4_normalized(float3 a) {
  float z = s4_sqnorm(a);
  if (z > 0.f) return s4_div(a, __fsqrt_rn(z));
  return a;
}
__device__ __forceinline__ float3 s4_xyz(float4 v) { return make_float3(v.x, v.y, v.z); }
