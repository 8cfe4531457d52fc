// CPU unit test of cpp/shards.h (candidate-set sharding inside the C++ layer): the shard reduction keeps the reference's
// first-maximum rule, and ForEachShard runs every shard, joins all of them and rethrows the first failure in rank order.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "shards.h"

using GlobalRegistration::detail::ForEachShard;
using GlobalRegistration::detail::MergeShards;

static s4g_tcs_result shard(int index, unsigned count, unsigned gate) {
  s4g_tcs_result r;
  std::memset(&r, 0, sizeof r);
  r.best_index = index;
  r.best_count = count;
  r.n_gate_pass = gate;
  r.key = index < 0 ? 0 : ((uint64_t(count) << 32) | uint64_t(0xFFFFFFFFu - uint32_t(index)));
  return r;
}

#define CHECK(c) do { if (!(c)) { std::printf("FAIL line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main() {
  // highest count wins, wherever it sits
  s4g_tcs_result m = MergeShards({shard(4, 10, 3), shard(1, 12, 2), shard(2, 11, 5)});
  CHECK(m.best_index == 1 && m.best_count == 12 && m.n_gate_pass == 10);
  // equal counts: the smallest quad index (= first in the reference's candidate order) wins
  m = MergeShards({shard(8, 7, 1), shard(5, 7, 1), shard(6, 7, 1)});
  CHECK(m.best_index == 5 && m.n_gate_pass == 3);
  // a winner with zero inliers still beats "nothing verified"; index 0 has the largest key of its count
  m = MergeShards({shard(-1, 0, 0), shard(3, 0, 1), shard(0, 0, 1)});
  CHECK(m.best_index == 0 && m.best_count == 0 && m.n_gate_pass == 2);
  // nothing verified anywhere
  m = MergeShards({shard(-1, 0, 0), shard(-1, 0, 0)});
  CHECK(m.best_index == -1 && m.key == 0 && m.n_gate_pass == 0);
  // one shard
  m = MergeShards({shard(9, 2, 4)});
  CHECK(m.best_index == 9 && m.n_gate_pass == 4);

  // ForEachShard: rank 0 on the primary, rank r on peers[r - 1], world = 1 + #peers
  s4g_ctx* const primary = reinterpret_cast<s4g_ctx*>(0x10);
  std::vector<s4g_ctx*> peers = {reinterpret_cast<s4g_ctx*>(0x20), reinterpret_cast<s4g_ctx*>(0x30),
                                 reinterpret_cast<s4g_ctx*>(0x40)};
  std::atomic<int> calls{0}, ok{0};
  ForEachShard(primary, &peers, [&](s4g_ctx* c, int rank, int world) {
    ++calls;
    if (world == 4 && c == (rank == 0 ? primary : peers[size_t(rank - 1)])) ++ok;
  });
  CHECK(calls == 4 && ok == 4);
  calls = 0;
  ForEachShard(primary, nullptr, [&](s4g_ctx* c, int rank, int world) { calls += (c == primary && rank == 0 && world == 1); });
  CHECK(calls == 1);
  // failures: every shard still runs and is joined, the first failure in rank order surfaces
  calls = 0;
  std::string what;
  try {
    ForEachShard(primary, &peers, [&](s4g_ctx*, int rank, int) {
      ++calls;
      if (rank == 1 || rank == 3) throw std::runtime_error("shard " + std::to_string(rank));
    });
  } catch (const std::runtime_error& e) {
    what = e.what();
  }
  CHECK(calls == 4 && what == "shard 1");
  std::printf("OK\n");
  return 0;
}
