// f2 -- Sampling::UniformDistSampler on the GPU (reference src/super4pcs/sampling.h:59-121):
// keep the FIRST point (smallest input index) of every voxel of edge `voxel`, output in input
// order.  The voxel of a point is (int(floor(x*s)), int(floor(y*s)), int(floor(z*s))) with
// s = 1.0f / voxel, exactly the reference's float arithmetic (sampling.h:75,90-92).
//
// One pass inserts every point into an open-addressing hash table keyed by the packed voxel
// (3 x 21 bits, relative to the voxel of the cloud's bounding-box minimum: any absolute coordinate range, e.g.
// georeferenced scans, as long as the cloud spans fewer than 2^21 voxels per axis) and keeps the minimum index per voxel with atomicMin; a second pass flags the points
// that ARE their voxel's minimum; cub::DeviceSelect compacts the flagged indices (ascending, so the
// output order is the input order).  HBM-streaming work: 12 B read + ~16 B of table traffic per point.
#include "s4g_internal.cuh"
#include <cub/cub.cuh>
#include <cmath>

namespace {

constexpr unsigned long long kEmpty = ~0ull;

struct VoxelBase { int x, y, z; };   // voxel of the bounding-box minimum

__device__ __forceinline__ bool voxel_key(const float* __restrict__ xyz, long long i, float scale, VoxelBase b,
                                          unsigned long long& key) {
  const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
  const long long cx = (long long)(int)floorf(__fmul_rn(x, scale)) - b.x, cy = (long long)(int)floorf(__fmul_rn(y, scale)) - b.y,
                  cz = (long long)(int)floorf(__fmul_rn(z, scale)) - b.z;
  const long long lim = 1ll << 21;
  if (cx < 0 || cx >= lim || cy < 0 || cy >= lim || cz < 0 || cz >= lim) return false;
  key = ((unsigned long long)cx << 42) | ((unsigned long long)cy << 21) | (unsigned long long)cz;
  return true;
}
__device__ __forceinline__ unsigned long long mix(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}

__global__ void k_voxel_insert(const float* __restrict__ xyz, long long n, float scale, VoxelBase vb,
                               unsigned long long* __restrict__ keys, unsigned int* __restrict__ minidx,
                               unsigned long long mask, int* __restrict__ err) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long key;
  if (!voxel_key(xyz, i, scale, vb, key)) { *err = 1; return; }
  unsigned long long slot = mix(key) & mask;
  for (;;) {
    unsigned long long prev = keys[slot];
    if (prev == kEmpty) prev = atomicCAS(&keys[slot], kEmpty, key);
    if (prev == kEmpty || prev == key) { atomicMin(&minidx[slot], (unsigned int)i); return; }
    slot = (slot + 1) & mask;
  }
}

__global__ void k_voxel_flag(const float* __restrict__ xyz, long long n, float scale, VoxelBase vb,
                             const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ minidx,
                             unsigned long long mask, unsigned char* __restrict__ flags) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long key;
  if (!voxel_key(xyz, i, scale, vb, key)) { flags[i] = 0; return; }
  unsigned long long slot = mix(key) & mask;
  while (keys[slot] != key) slot = (slot + 1) & mask;
  flags[i] = minidx[slot] == (unsigned int)i;
}

}  // namespace

extern "C" int s4g_voxel_sample(s4g_ctx* ctx, const float* xyz, int64_t n, float voxel, int32_t* out_indices,
                                int64_t* n_out) {
  if (!ctx) return S4G_ERR_ARG;
  if (!xyz || n <= 0 || !(voxel > 0.f) || !out_indices || !n_out || n >= (1ll << 31)) {
    ctx->err = "s4g_voxel_sample: bad arguments";
    return S4G_ERR_ARG;
  }
  S4G_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  unsigned long long cap = 1;
  while (cap < 2ull * (unsigned long long)n) cap <<= 1;
  S4G_TRY(s4g_reserve(ctx, ctx->dScratchA, (size_t)n * 3 * sizeof(float)));
  S4G_TRY(s4g_reserve(ctx, ctx->dScratchB, (size_t)cap * sizeof(unsigned long long)));
  S4G_TRY(s4g_reserve(ctx, ctx->dScratchC, (size_t)cap * sizeof(unsigned int)));
  S4G_TRY(s4g_reserve(ctx, ctx->dScratchD, (size_t)n * (sizeof(unsigned char) + 2 * sizeof(int32_t)) + 64));
  S4G_TRY(s4g_reserve(ctx, ctx->dMisc, 256));
  float* d_xyz = ctx->dScratchA.as<float>();
  unsigned long long* keys = ctx->dScratchB.as<unsigned long long>();
  unsigned int* minidx = ctx->dScratchC.as<unsigned int>();
  int32_t* iota = ctx->dScratchD.as<int32_t>();
  int32_t* sel = iota + n;
  unsigned char* flags = reinterpret_cast<unsigned char*>(sel + n);
  int* d_err = ctx->dMisc.as<int>();
  int* d_num = d_err + 1;
  S4G_CUDA(cudaMemcpyAsync(d_xyz, xyz, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
  S4G_CUDA(cudaMemsetAsync(keys, 0xFF, (size_t)cap * sizeof(unsigned long long), st));
  S4G_CUDA(cudaMemsetAsync(minidx, 0xFF, (size_t)cap * sizeof(unsigned int), st));
  S4G_CUDA(cudaMemsetAsync(d_err, 0, 8, st));
  const float scale = 1.0f / voxel;                                   // sampling.h:75
  VoxelBase vb;
  {
    // voxel of the bounding-box minimum (floor and the float product are monotone: it is the smallest voxel coordinate);
    // non-finite coordinates are rejected here
    float mn[3] = {xyz[0], xyz[1], xyz[2]};
    for (int64_t i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) {
        const float v = xyz[3 * i + k];
        if (!std::isfinite(v)) { ctx->err = "s4g_voxel_sample: NaN or infinite coordinate"; return S4G_ERR_ARG; }
        mn[k] = v < mn[k] ? v : mn[k];
      }
    volatile float px = mn[0] * scale, py = mn[1] * scale, pz = mn[2] * scale;   // one rounding each, like __fmul_rn
    vb.x = (int)std::floor((float)px);
    vb.y = (int)std::floor((float)py);
    vb.z = (int)std::floor((float)pz);
  }
  const unsigned nb = (unsigned)((n + 255) / 256);
  k_voxel_insert<<<nb, 256, 0, st>>>(d_xyz, n, scale, vb, keys, minidx, cap - 1, d_err);
  k_voxel_flag<<<nb, 256, 0, st>>>(d_xyz, n, scale, vb, keys, minidx, cap - 1, flags);
  cub::CountingInputIterator<int32_t> counting(0);
  size_t cub_bytes = 0;
  cub::DeviceSelect::Flagged(nullptr, cub_bytes, counting, flags, sel, d_num, (int)n, st);
  S4G_TRY(s4g_reserve(ctx, ctx->dCub, cub_bytes));
  cub::DeviceSelect::Flagged(ctx->dCub.p, cub_bytes, counting, flags, sel, d_num, (int)n, st);
  ctx->launches += 4;
  S4G_CUDA(cudaGetLastError());
  int h[2] = {0, 0};
  S4G_CUDA(cudaMemcpyAsync(h, d_err, 8, cudaMemcpyDeviceToHost, st));
  S4G_CUDA(cudaStreamSynchronize(st));
  if (h[0]) { ctx->err = "s4g_voxel_sample: the cloud spans more than 2^21 voxels along an axis"; return S4G_ERR_ARG; }
  S4G_CUDA(cudaMemcpyAsync(out_indices, sel, (size_t)h[1] * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  S4G_CUDA(cudaStreamSynchronize(st));
  *n_out = h[1];
  return S4G_OK;
}
