// a8 -- Match4PCSBase::Verify (reference algorithms/match4pcsBase.cc:508-567) for a batch of
// candidate transforms, on the bricked uniform grid built by s4g_set_cloud_p.
//
//   counts[k] = #{ q in sampled_Q : exists p in sampled_P, ||T_k q - p||^2 <= delta^2 }
//
// Arithmetic follows the reference's binary operation by operation:
//   T q  = ((m0 x + m1 y) + m2 z) + m3   per row, fp32, no FMA      (match4pcsBase.cc:532)
//   d^2  = dx^2 + (dy^2 + dz^2)                                     (kdtree.h:417)
//   hit  = d^2 <= delta*delta                                       (kdtree.h:418, cc:522)
//
// Layout / schedule (B200): sampled_Q is streamed in Morton order (float4, one coalesced
// 16-byte load per thread per tile) so the 32 queries of a warp land in a handful of
// neighbouring grid cells after the rigid motion; the P grid (points + cellStart + brick table,
// ~30 MB at 1M points) is L2-resident and read through the read-only path.  Each thread keeps its
// query in registers and loops over a chunk of candidate transforms staged in shared memory;
// inlier votes are reduced with __ballot_sync/__popc and one shared-memory add per warp per
// candidate, then one global atomicAdd per block per candidate.
//
// Probe: cell edge h >= 2.02*delta, so the delta-ball around T q touches at most 2 cells per
// axis: x0 = floor(u - 0.5), cells {x0, x0+1} (u = cell coordinate of T q).  For a point p with
// |T q - p|_x <= delta(1+1e-6): |u - v| <= 0.4951, u in [x0+0.5, x0+1.5) => v in
// (x0+0.0049, x0+1.9951): the 0.0049-cell margin dominates the rounding of u and v (<= 2 ulp of
// a coordinate < 8192 cells = 0.001), so the probe is conservative and the count exact.
// Inside a brick the two x-neighbours are adjacent cellStart entries => one contiguous run.
#include "s4g_internal.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kCandPerBlock = 16;   // transforms staged per block
constexpr int kTilesPerBlock = 4;   // query tiles (of kThreads) per block

// Scan one contiguous run of P points; single exit, flag based (no multi-level early returns).
template <bool kStats>
__device__ __forceinline__ bool probe_run(const GridDev& g, uint32_t s, uint32_t e, float tx, float ty,
                                          float tz, float sq_eps, unsigned long long& n_tested) {
  bool found = false;
  for (uint32_t k = s; k < e && !found; ++k) {
    float4 p = __ldg(&g.pts[k]);
    float dx = __fsub_rn(tx, p.x), dy = __fsub_rn(ty, p.y), dz = __fsub_rn(tz, p.z);
    float d2 = __fadd_rn(__fmul_rn(dx, dx), __fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dz, dz)));
    if (kStats) n_tested++;
    found = d2 <= sq_eps;
  }
  return found;
}

// Does any P point lie within delta of (tx,ty,tz)?
template <bool kStats>
__device__ __forceinline__ bool any_within(const GridDev& g, float tx, float ty, float tz, float sq_eps,
                                           unsigned long long& n_tested, unsigned long long& n_ranges) {
  float ux = (tx - g.ox) * g.inv_h, uy = (ty - g.oy) * g.inv_h, uz = (tz - g.oz) * g.inv_h;
  // queries outside the padded grid (or NaN) cannot have a neighbour
  bool inside = ux > -1.f && uy > -1.f && uz > -1.f && ux < (float)(g.nx + 1) && uy < (float)(g.ny + 1) &&
                uz < (float)(g.nz + 1);
  if (!inside) { ux = uy = uz = -8.f; }
  const int x0 = (int)floorf(ux - 0.5f), y0 = (int)floorf(uy - 0.5f), z0 = (int)floorf(uz - 0.5f);
  const int bs = g.bshift, m = (1 << bs) - 1;
  const int xa = max(x0, 0), xb = min(x0 + 1, g.nx - 1);
  const bool same_brick = (xa >> bs) == (xb >> bs);
  bool found = false;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int cz = z0 + (r >> 1), cy = y0 + (r & 1);
    const bool row_ok = inside && !found && xa <= xb && cz >= 0 && cz < g.nz && cy >= 0 && cy < g.ny;
    if (row_ok) {
      const int rowb = ((cz >> bs) * g.tby + (cy >> bs)) * g.tbx;
      const uint32_t rowl = (uint32_t)((((cz & m) << bs) | (cy & m)) << bs);
      // first (or only) x cell -- when both x cells share a brick they are ONE contiguous run
      const int ra = __ldg(&g.top[rowb + (xa >> bs)]);
      if (ra >= 0) {
        const uint32_t idx = ((uint32_t)ra << (3 * bs)) | rowl | (uint32_t)(xa & m);
        const uint32_t s = __ldg(&g.cellStart[idx]);
        const uint32_t e = __ldg(&g.cellStart[idx + (same_brick ? (uint32_t)(xb - xa) : 0u) + 1u]);
        if (kStats) n_ranges++;
        found = probe_run<kStats>(g, s, e, tx, ty, tz, sq_eps, n_tested);
      }
      if (!same_brick && !found) {
        const int rb = __ldg(&g.top[rowb + (xb >> bs)]);
        if (rb >= 0) {
          const uint32_t idx = ((uint32_t)rb << (3 * bs)) | rowl | (uint32_t)(xb & m);
          const uint32_t s = __ldg(&g.cellStart[idx]);
          const uint32_t e = __ldg(&g.cellStart[idx + 1u]);
          if (kStats) n_ranges++;
          found = probe_run<kStats>(g, s, e, tx, ty, tz, sq_eps, n_tested);
        }
      }
    }
  }
  return found;
}

// T12: K x 12 floats, row-major 3x4 (r00 r01 r02 t0 | r10 ... ), the top three rows of T.
// grid.x = query super-tiles (kThreads*kTilesPerBlock queries), grid.y = candidate chunks.
template <bool kStats>
__global__ void __launch_bounds__(kThreads)
k_verify(GridDev g, const float4* __restrict__ Q, int nQ, const float* __restrict__ T12, int K,
         float sq_eps, uint32_t* __restrict__ counts, unsigned long long* __restrict__ stats) {
  __shared__ float sT[kCandPerBlock * 12];
  __shared__ uint32_t sCnt[kCandPerBlock];
  const int c0 = blockIdx.y * kCandPerBlock;
  const int nc = min(kCandPerBlock, K - c0);
  for (int i = threadIdx.x; i < nc * 12; i += kThreads) sT[i] = T12[(size_t)c0 * 12 + i];
  if (threadIdx.x < kCandPerBlock) sCnt[threadIdx.x] = 0;
  __syncthreads();

  unsigned long long n_tested = 0, n_ranges = 0;
  const int lane = threadIdx.x & 31;
  const long long qbase = (long long)blockIdx.x * (kThreads * kTilesPerBlock);
#pragma unroll 1
  for (int t = 0; t < kTilesPerBlock; ++t) {
    long long qi = qbase + (long long)t * kThreads + threadIdx.x;
    const bool valid = qi < nQ;
    float4 q = valid ? __ldg(&Q[qi]) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (__ballot_sync(0xffffffffu, valid) == 0u) break;
#pragma unroll 1
    for (int c = 0; c < nc; ++c) {
      const float* m = &sT[c * 12];
      // ((m0 x + m1 y) + m2 z) + m3
      float tx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], q.x), __fmul_rn(m[1], q.y)), __fmul_rn(m[2], q.z)), m[3]);
      float ty = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[4], q.x), __fmul_rn(m[5], q.y)), __fmul_rn(m[6], q.z)), m[7]);
      float tz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[8], q.x), __fmul_rn(m[9], q.y)), __fmul_rn(m[10], q.z)), m[11]);
      bool hit = valid && any_within<kStats>(g, tx, ty, tz, sq_eps, n_tested, n_ranges);
      unsigned b = __ballot_sync(0xffffffffu, hit);
      if (lane == 0 && b) atomicAdd(&sCnt[c], (uint32_t)__popc(b));
      __syncwarp();
    }
  }
  __syncthreads();
  if (threadIdx.x < nc) {
    uint32_t v = sCnt[threadIdx.x];
    if (v) atomicAdd(&counts[c0 + threadIdx.x], v);
  }
  if (kStats) {
    atomicAdd(&stats[0], n_tested);
    atomicAdd(&stats[1], n_ranges);
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
    k_verify<false><<<grid, kThreads, 0, st>>>(ctx->grid, ctx->dQmorton.as<float4>(), ctx->nQ,
                                               d_T12 + (size_t)c0 * 12, kk, sq_eps, d_counts + c0, nullptr);
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

extern "C" int s4g_verify_probe_stats(s4g_ctx* ctx, const float* T, int K, uint64_t* out2) {
  if (!ctx) return S4G_ERR_ARG;
  if (K <= 0 || !T || !out2) { ctx->err = "s4g_verify_probe_stats: bad arguments"; return S4G_ERR_ARG; }
  S4G_TRY(check_ready(ctx, "s4g_verify_probe_stats"));
  S4G_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  S4G_TRY(s4g_reserve(ctx, ctx->dScratchA, (size_t)K * 16 * sizeof(float)));
  S4G_TRY(s4g_reserve(ctx, ctx->dCounts, (size_t)K * sizeof(uint32_t)));
  S4G_TRY(s4g_reserve(ctx, ctx->dT12, (size_t)K * 12 * sizeof(float)));
  S4G_TRY(s4g_reserve(ctx, ctx->dMisc, 256));
  S4G_CUDA(cudaMemcpyAsync(ctx->dScratchA.p, T, (size_t)K * 16 * sizeof(float), cudaMemcpyHostToDevice, st));
  S4G_CUDA(cudaMemsetAsync(ctx->dMisc.p, 0, 16, st));
  S4G_CUDA(cudaMemsetAsync(ctx->dCounts.p, 0, (size_t)K * sizeof(uint32_t), st));
  k_pack_T12<<<(K * 12 + 255) / 256, 256, 0, st>>>(ctx->dScratchA.as<float>(), K, ctx->dT12.as<float>());
  const int per_block = kThreads * kTilesPerBlock;
  dim3 grid((unsigned)((ctx->nQ + per_block - 1) / per_block), (unsigned)((K + kCandPerBlock - 1) / kCandPerBlock), 1);
  if (grid.y > 65535) { ctx->err = "s4g_verify_probe_stats: K too large"; return S4G_ERR_ARG; }
  k_verify<true><<<grid, kThreads, 0, st>>>(ctx->grid, ctx->dQmorton.as<float4>(), ctx->nQ, ctx->dT12.as<float>(), K,
                                            ctx->delta * ctx->delta, ctx->dCounts.as<uint32_t>(),
                                            ctx->dMisc.as<unsigned long long>());
  ctx->launches += 2;
  S4G_CUDA(cudaGetLastError());
  unsigned long long h[2] = {0, 0};
  S4G_CUDA(cudaMemcpyAsync(h, ctx->dMisc.p, 16, cudaMemcpyDeviceToHost, st));
  S4G_CUDA(cudaStreamSynchronize(st));
  out2[0] = h[0];
  out2[1] = h[1];
  return S4G_OK;
}
