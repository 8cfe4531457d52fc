// super4pcs-b200: candidate-set sharding across several device contexts owned by ONE process (SURVEY.md section 8,
// row e, inside the C++ layer; S4PCS_DEVICES).  The candidates (congruent quads) of a base are independent given the
// replicated clouds, so context r of W runs TryCongruentSet on the quads with index % W == r (the shard_rank /
// shard_world arguments of include/s4g.h) and the W shard results are combined by the maximum of the packed key
// (count << 32) | (0xFFFFFFFF - quad index): highest inlier count, ties -> smallest index = the reference's
// first-maximum rule (match4pcsBase.hpp:468).  Here all W contexts belong to the calling process: by default the
// maximum is a W-element loop over the records the shard threads read back; with S4PCS_NCCL=1 the contexts share a
// communicator and libs4g reduces on the devices (ncclAllReduce, csrc/comm.cu) -- the shard calls then all return the
// global record.  Between processes (one per GPU) the same library path is used with s4g_comm_init_rank (bench.py).
#ifndef SUPER4PCS_B200_CPP_SHARDS_H_
#define SUPER4PCS_B200_CPP_SHARDS_H_

#include <condition_variable>
#include <cstddef>
#include <cstring>
#include <exception>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

#include "s4g.h"

namespace GlobalRegistration {
namespace detail {

/// S4PCS_NCCL: the library-side reduction waits for every rank (up to its deadline), so a shard thread should only enter it
/// when every other shard thread will.  Each thread calls Pass() right before its TryCongruentSet call; a thread that leaves its
/// pass without having arrived (an exception or an early return in an earlier stage) is recorded by ForEachShard, and Pass()
/// then returns false on all the others, which skip the call.  One gate per ForEachShard call.
class ShardGate {
 public:
  explicit ShardGate(int world) : arrived_(size_t(world), 0), left_(size_t(world), 0) {}
  bool Pass(int rank) {
    std::unique_lock<std::mutex> lock(m_);
    arrived_[size_t(rank)] = 1;
    cv_.notify_all();
    cv_.wait(lock, [this] {
      for (size_t r = 0; r < arrived_.size(); ++r)
        if (!arrived_[r] && !left_[r]) return false;
      return true;
    });
    for (char gone : left_)
      if (gone) return false;
    return true;
  }
  void Leave(int rank) {
    std::lock_guard<std::mutex> lock(m_);
    if (!arrived_[size_t(rank)]) left_[size_t(rank)] = 1;
    cv_.notify_all();
  }

 private:
  std::mutex m_;
  std::condition_variable cv_;
  std::vector<char> arrived_, left_;
};

/// Calls fn(context, rank, world) once per shard: rank 0 = `primary` on the calling thread, rank r > 0 = peers[r - 1]
/// on a thread of its own (a shard is a chain of stream launches with blocking size read-backs, so host threads are
/// what lets the devices run at the same time).  All shards are joined; the first failure in rank order is rethrown.
template <typename Fn>
void ForEachShard(s4g_ctx* primary, const std::vector<s4g_ctx*>* peers, Fn fn, ShardGate* gate = nullptr) {
  const int world = 1 + (peers ? int(peers->size()) : 0);
  if (world == 1) {
    fn(primary, 0, 1);
    return;
  }
  std::vector<std::exception_ptr> errors(static_cast<size_t>(world));
  auto guarded = [&fn, &errors, world, gate](s4g_ctx* ctx, int rank) {
    try {
      fn(ctx, rank, world);
    } catch (...) {
      errors[size_t(rank)] = std::current_exception();
    }
    if (gate) gate->Leave(rank);  // (no-op for a shard that went through the gate)
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

/// With a communicator attached (S4PCS_NCCL) libs4g has reduced key and record on the devices: every shard call returned
/// the global result.  Anything else is a broken collective, not a tie to resolve.
inline s4g_tcs_result SameOnAllShards(const std::vector<s4g_tcs_result>& shards) {
  for (size_t r = 1; r < shards.size(); ++r)
    if (shards[r].key != shards[0].key || shards[r].best_index != shards[0].best_index ||
        shards[r].best_count != shards[0].best_count || shards[r].n_gate_pass != shards[0].n_gate_pass ||
        std::memcmp(shards[r].best_T, shards[0].best_T, sizeof shards[0].best_T) != 0)
      throw std::runtime_error("super4pcs-b200: S4PCS_NCCL: the devices returned different records after the reduction");
  return shards[0];
}

/// the result of a base whose shards were reduced on the devices (S4PCS_NCCL) or, by default, are merged here
inline s4g_tcs_result CombineShards(const std::vector<s4g_tcs_result>& shards, bool reduced_on_device) {
  return reduced_on_device && shards.size() > 1 ? SameOnAllShards(shards) : MergeShards(shards);
}

}  // namespace detail
}  // namespace GlobalRegistration

#endif  // SUPER4PCS_B200_CPP_SHARDS_H_
