// super4pcs-b200: candidate-set sharding across several device contexts owned by ONE process (SURVEY.md section 8,
// row e, inside the C++ layer; S4PCS_DEVICES).  The candidates (congruent quads) of a base are independent given the
// replicated clouds, so context r of W runs TryCongruentSet on the quads with index % W == r (the shard_rank /
// shard_world arguments of include/s4g.h) and the W shard results are combined by the maximum of the packed key
// (count << 32) | (0xFFFFFFFF - quad index): highest inlier count, ties -> smallest index = the reference's
// first-maximum rule (match4pcsBase.hpp:468).  Between processes that maximum is one NCCL allreduce (bench.py,
// super4pcs_b200/sharding.py); here all W contexts belong to the calling process and it is a W-element loop.
#ifndef SUPER4PCS_B200_CPP_SHARDS_H_
#define SUPER4PCS_B200_CPP_SHARDS_H_

#include <cstddef>
#include <exception>
#include <thread>
#include <vector>

#include "s4g.h"

namespace GlobalRegistration {
namespace detail {

/// Calls fn(context, rank, world) once per shard: rank 0 = `primary` on the calling thread, rank r > 0 = peers[r - 1]
/// on a thread of its own (a shard is a chain of stream launches with blocking size read-backs, so host threads are
/// what lets the devices run at the same time).  All shards are joined; the first failure in rank order is rethrown.
template <typename Fn>
void ForEachShard(s4g_ctx* primary, const std::vector<s4g_ctx*>* peers, Fn fn) {
  const int world = 1 + (peers ? int(peers->size()) : 0);
  if (world == 1) {
    fn(primary, 0, 1);
    return;
  }
  std::vector<std::exception_ptr> errors(static_cast<size_t>(world));
  auto guarded = [&fn, &errors, world](s4g_ctx* ctx, int rank) {
    try {
      fn(ctx, rank, world);
    } catch (...) {
      errors[size_t(rank)] = std::current_exception();
    }
  };
  std::vector<std::thread> workers;
  workers.reserve(size_t(world - 1));
  for (int r = 1; r < world; ++r) workers.emplace_back(guarded, (*peers)[size_t(r - 1)], r);
  guarded(primary, 0);
  for (std::thread& w : workers) w.join();
  for (const std::exception_ptr& e : errors)
    if (e) std::rethrow_exception(e);
}

/// The single reduction of row e: the shard with the largest key holds the winner; gate passes (= Verify calls) add up.
inline s4g_tcs_result MergeShards(const std::vector<s4g_tcs_result>& shards) {
  size_t win = 0;
  unsigned long long gate = 0;
  for (size_t r = 0; r < shards.size(); ++r) {
    gate += shards[r].n_gate_pass;
    const bool has = shards[r].best_index >= 0, cur = shards[win].best_index >= 0;
    if (has && (!cur || shards[r].key > shards[win].key)) win = r;
  }
  s4g_tcs_result out = shards[win];
  out.n_gate_pass = uint32_t(gate);
  return out;
}

}  // namespace detail
}  // namespace GlobalRegistration

#endif  // SUPER4PCS_B200_CPP_SHARDS_H_
