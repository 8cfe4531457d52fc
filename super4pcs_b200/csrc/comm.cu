// super4pcs-b200: the one collective of the path, inside the library (SURVEY.md section 8, row e).
//
// The candidates of a base are sharded over contexts by index (shard_rank / shard_world of include/s4g.h); what has to be
// exchanged is the winner: the maximum of the packed key (count << 32) | (0xFFFFFFFF - index) -- the reference's
// first-maximum rule, match4pcsBase.hpp:467-484 -- and the owner's 136-byte result record.  With a communicator attached to
// the contexts both happen on the device, on the stream that ran k_verify: one ncclAllReduce(ncclMax, uint64), then the
// non-owners zero their record and one ncclAllReduce(ncclSum, uint32 x 34) leaves the owner's record (and the sum of the
// shards' gate counts) on every rank.  No host hop between the argmax and the result read-back.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2 on the first s4g_comm_* call): a process that already carries an
// NCCL (torch's bundled one under torchrun) shares it, a C++ caller gets the system library, and callers that never attach a
// communicator never load it.  There is no substitute path: without NCCL the s4g_comm_* calls fail with S4G_ERR_COMM.
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <future>
#include <memory>
#include <cstddef>
#include <cstring>
#include <mutex>
#include <thread>

#include "s4g_internal.cuh"

namespace {

struct NcclApi {
  void* lib = nullptr;
  std::string why;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  bool ok() const { return lib != nullptr && why.empty(); }
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names)
      if ((api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
    if (!api.lib) {
      const char* e = dlerror();
      api.why = std::string("NCCL is not loadable (") + (e ? e : "dlopen failed") + ")";
      return;
    }
    auto need = [](const char* sym, void** out) {
      *out = dlsym(api.lib, sym);
      if (!*out && api.why.empty()) api.why = std::string("NCCL lacks ") + sym;
    };
    need("ncclGetVersion", reinterpret_cast<void**>(&api.GetVersion));
    need("ncclGetUniqueId", reinterpret_cast<void**>(&api.GetUniqueId));
    need("ncclCommInitRank", reinterpret_cast<void**>(&api.CommInitRank));
    need("ncclCommInitAll", reinterpret_cast<void**>(&api.CommInitAll));
    need("ncclCommDestroy", reinterpret_cast<void**>(&api.CommDestroy));
    need("ncclCommAbort", reinterpret_cast<void**>(&api.CommAbort));
    need("ncclAllReduce", reinterpret_cast<void**>(&api.AllReduce));
    need("ncclGetErrorString", reinterpret_cast<void**>(&api.GetErrorString));
    need("ncclGroupStart", reinterpret_cast<void**>(&api.GroupStart));
    need("ncclGroupEnd", reinterpret_cast<void**>(&api.GroupEnd));
  });
  return api;
}

int comm_fail(s4g_ctx* ctx, const char* what, ncclResult_t r) {
  if (ctx) ctx->err = std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error");
  return S4G_ERR_COMM;
}

#define S4G_CUDA_CTX(c, call)                                                    \
  do {                                                                           \
    cudaError_t e_ = (call);                                                     \
    if (e_ != cudaSuccess) {                                                     \
      (c)->err = std::string(#call) + ": " + cudaGetErrorString(e_);             \
      return S4G_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)

#define S4G_NCCL(ctx, call)                                     \
  do {                                                          \
    ncclResult_t r_ = (call);                                   \
    if (r_ != ncclSuccess) return comm_fail((ctx), #call, r_);  \
  } while (0)

// non-owners clear their record (all of it but the gate count, which adds up); the owner is the rank whose local key is
// the global one -- indices are disjoint between shards, so exactly one -- or rank 0 when no shard verified anything
__global__ void k_mask_result(const unsigned long long* __restrict__ local, const unsigned long long* __restrict__ global,
                              int rank, uint32_t* __restrict__ rec, int words, int gate_word) {
  const unsigned long long g = *global;
  const bool owner = (*local == g) && (g != 0ull || rank == 0);
  if (owner) return;
  for (int w = threadIdx.x; w < words; w += blockDim.x)
    if (w != gate_word) rec[w] = 0u;
}

}  // namespace

// a sharded call on a context with a communicator is reduced over it (a communicator of one rank included: the same chain,
// NCCL copying in place); shard_world == 1 on a context of a larger communicator stays the plain local call
bool s4g_comm_active(const s4g_ctx* ctx, int shard_world) { return ctx->comm != nullptr && shard_world == ctx->comm_ranks; }

int s4g_comm_check_shard(s4g_ctx* ctx, int shard_rank, int shard_world) {
  if (!ctx->comm || (shard_world == 1 && ctx->comm_ranks != 1)) return S4G_OK;
  if (shard_world != ctx->comm_ranks || shard_rank != ctx->comm_rank) {
    ctx->err = "shard_rank / shard_world differ from the rank / size of the communicator attached to this context";
    return S4G_ERR_ARG;
  }
  return S4G_OK;
}

int s4g_comm_max_u64(s4g_ctx* ctx, const unsigned long long* d_in, unsigned long long* d_out, cudaStream_t st) {
  S4G_NCCL(ctx, nccl().AllReduce(d_in, d_out, 1, ncclUint64, ncclMax, static_cast<ncclComm_t>(ctx->comm), st));
  ctx->collectives++;
  return S4G_OK;
}

// d_local: this shard's key, d_global: scratch for the reduced key, rec: this shard's s4g_tcs_result (device)
int s4g_comm_reduce_result(s4g_ctx* ctx, const unsigned long long* d_local, unsigned long long* d_global,
                           s4g_tcs_result* rec, cudaStream_t st) {
  static_assert(sizeof(s4g_tcs_result) % 4 == 0, "the record is reduced as 32-bit words");
  const int words = int(sizeof(s4g_tcs_result) / 4), gate_word = int(offsetof(s4g_tcs_result, n_gate_pass) / 4);
  S4G_TRY(s4g_comm_max_u64(ctx, d_local, d_global, st));
  k_mask_result<<<1, 64, 0, st>>>(d_local, d_global, ctx->comm_rank, reinterpret_cast<uint32_t*>(rec), words, gate_word);
  ctx->launches++;
  S4G_CUDA(cudaGetLastError());
  S4G_NCCL(ctx, nccl().AllReduce(rec, rec, size_t(words), ncclUint32, ncclSum, static_cast<ncclComm_t>(ctx->comm), st));
  ctx->collectives++;
  return S4G_OK;
}

// cudaStreamSynchronize with a deadline: a peer that failed before its collective would otherwise block this rank for ever
int s4g_comm_wait(s4g_ctx* ctx, cudaStream_t st) {
  if (!ctx->comm) {
    S4G_CUDA(cudaStreamSynchronize(st));
    return S4G_OK;
  }
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(ctx->comm_timeout_s);
  for (unsigned spins = 0;; ++spins) {
    const cudaError_t e = cudaStreamQuery(st);
    if (e == cudaSuccess) return S4G_OK;
    if (e != cudaErrorNotReady) {
      ctx->err = std::string("collective wait: ") + cudaGetErrorString(e);
      return S4G_ERR_CUDA;
    }
    if (std::chrono::steady_clock::now() > deadline) {
      // ncclCommAbort makes the collective's kernel exit; it runs on a thread of its own so that an NCCL that is itself
      // stuck cannot hold this rank.  If the stream does not drain afterwards the context is marked unusable: s4g_destroy
      // then returns without synchronising (device memory is left to the process teardown).
      ncclComm_t comm = static_cast<ncclComm_t>(ctx->comm);
      ctx->comm = nullptr;
      ctx->comm_ranks = 1;
      ctx->comm_rank = 0;
      auto done = std::make_shared<std::promise<void>>();
      std::future<void> aborted = done->get_future();
      std::thread([comm, done] {
        nccl().CommAbort(comm);
        done->set_value();
      }).detach();
      bool drained = false;
      if (aborted.wait_for(std::chrono::seconds(5)) == std::future_status::ready) {
        const auto until = std::chrono::steady_clock::now() + std::chrono::seconds(5);
        while (!drained && std::chrono::steady_clock::now() < until) {
          drained = cudaStreamQuery(st) != cudaErrorNotReady;
          if (!drained) std::this_thread::yield();
        }
      }
      ctx->stuck = !drained;
      ctx->err = "collective wait: a peer did not reach the reduction within the time limit (communicator aborted)";
      return S4G_ERR_COMM;
    }
    if (spins > 2000) std::this_thread::yield();
  }
}

extern "C" int s4g_comm_unique_id(unsigned char* out_id) {
  if (!out_id) return S4G_ERR_ARG;
  if (!nccl().ok()) return S4G_ERR_COMM;
  ncclUniqueId id;
  if (nccl().GetUniqueId(&id) != ncclSuccess) return S4G_ERR_COMM;
  static_assert(sizeof id == S4G_COMM_ID_BYTES, "s4g.h mirrors NCCL_UNIQUE_ID_BYTES");
  std::memcpy(out_id, &id, sizeof id);
  return S4G_OK;
}

extern "C" int s4g_comm_init_rank(s4g_ctx* ctx, const unsigned char* id_bytes, int n_ranks, int rank) {
  if (!ctx) return S4G_ERR_ARG;
  if (!id_bytes || n_ranks < 1 || rank < 0 || rank >= n_ranks) { ctx->err = "s4g_comm_init_rank: bad arguments"; return S4G_ERR_ARG; }
  if (ctx->comm) { ctx->err = "s4g_comm_init_rank: a communicator is already attached (s4g_comm_destroy first)"; return S4G_ERR_STATE; }
  if (!nccl().ok()) { ctx->err = "s4g_comm_init_rank: " + nccl().why; return S4G_ERR_COMM; }
  S4G_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  std::memcpy(&id, id_bytes, sizeof id);
  ncclComm_t comm = nullptr;
  S4G_NCCL(ctx, nccl().CommInitRank(&comm, n_ranks, id, rank));
  ctx->comm = comm;
  ctx->comm_ranks = n_ranks;
  ctx->comm_rank = rank;
  // NCCL connects its transports on the first collective, and that is a blocking exchange between the ranks: do it here,
  // where every rank is known to be present, so that later collectives are plain stream-ordered launches
  S4G_TRY(s4g_reserve(ctx, ctx->dMisc, 256));
  S4G_NCCL(ctx, nccl().AllReduce(ctx->dMisc.p, ctx->dMisc.p, 1, ncclUint64, ncclMax, comm, ctx->stream));
  S4G_CUDA(cudaStreamSynchronize(ctx->stream));
  return S4G_OK;
}

extern "C" int s4g_comm_init_all(s4g_ctx** ctxs, int n) {
  if (!ctxs || n < 1 || n > 64) return S4G_ERR_ARG;
  for (int r = 0; r < n; ++r)
    if (!ctxs[r]) return S4G_ERR_ARG;
  s4g_ctx* first = ctxs[0];
  for (int r = 0; r < n; ++r) {
    if (ctxs[r]->comm) { first->err = "s4g_comm_init_all: a communicator is already attached (s4g_comm_destroy first)"; return S4G_ERR_STATE; }
    for (int q = 0; q < r; ++q)
      if (ctxs[q]->device == ctxs[r]->device) { first->err = "s4g_comm_init_all: two contexts on one device (NCCL wants one rank per GPU)"; return S4G_ERR_ARG; }
  }
  if (!nccl().ok()) { first->err = "s4g_comm_init_all: " + nccl().why; return S4G_ERR_COMM; }
  int devs[64];
  ncclComm_t comms[64];
  for (int r = 0; r < n; ++r) devs[r] = ctxs[r]->device;
  S4G_NCCL(first, nccl().CommInitAll(comms, n, devs));
  for (int r = 0; r < n; ++r) {
    ctxs[r]->comm = comms[r];
    ctxs[r]->comm_ranks = n;
    ctxs[r]->comm_rank = r;
  }
  // first collective = transport set-up, a blocking exchange between the ranks (see s4g_comm_init_rank): one group call
  // from this thread for all of them, so that the per-context threads only ever enqueue
  for (int r = 0; r < n; ++r) {
    S4G_CUDA_CTX(first, cudaSetDevice(ctxs[r]->device));
    if (s4g_reserve(ctxs[r], ctxs[r]->dMisc, 256) != S4G_OK) { first->err = ctxs[r]->err; return S4G_ERR_NOMEM; }
  }
  S4G_NCCL(first, nccl().GroupStart());
  for (int r = 0; r < n; ++r) {
    const ncclResult_t e = nccl().AllReduce(ctxs[r]->dMisc.p, ctxs[r]->dMisc.p, 1, ncclUint64, ncclMax, comms[r], ctxs[r]->stream);
    if (e != ncclSuccess) { nccl().GroupEnd(); return comm_fail(first, "ncclAllReduce (transport set-up)", e); }
  }
  S4G_NCCL(first, nccl().GroupEnd());
  for (int r = 0; r < n; ++r) {
    S4G_CUDA_CTX(first, cudaSetDevice(ctxs[r]->device));
    S4G_CUDA_CTX(first, cudaStreamSynchronize(ctxs[r]->stream));
  }
  return S4G_OK;
}

extern "C" int s4g_comm_destroy(s4g_ctx* ctx) {
  if (!ctx) return S4G_ERR_ARG;
  if (!ctx->comm) return S4G_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  nccl().CommDestroy(static_cast<ncclComm_t>(ctx->comm));
  ctx->comm = nullptr;
  ctx->comm_ranks = 1;
  ctx->comm_rank = 0;
  return S4G_OK;
}

extern "C" int s4g_comm_info(s4g_ctx* ctx, int* out4) {
  if (!ctx || !out4) return S4G_ERR_ARG;
  int version = 0;
  if (ctx->comm && nccl().ok()) nccl().GetVersion(&version);
  out4[0] = ctx->comm ? ctx->comm_ranks : 0;
  out4[1] = ctx->comm ? ctx->comm_rank : 0;
  out4[2] = version;
  out4[3] = int(ctx->collectives & 0x7fffffff);
  return S4G_OK;
}

extern "C" int s4g_comm_set_timeout(s4g_ctx* ctx, int seconds) {
  if (!ctx || seconds < 1) return S4G_ERR_ARG;
  ctx->comm_timeout_s = seconds;
  return S4G_OK;
}
