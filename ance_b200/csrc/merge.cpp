// merge.cpp — host k-way merge of per-shard top-k lists.
//
// Replaces the reference's gather-through-the-filesystem + single-node search:
//   utils/util.py:87-146 (barrier_array_merge) and drivers/run_ann_data_gen.py:265-303; the
//   reference's own sharded precedent is utils/eval_mrr.py:175-183 (argsort over gathered top-k).
// Each shard list is sorted by (score desc, label asc); so is the output.  Labels -1 (padding of a
// shard with fewer than k rows) are skipped; the output is padded with -1 / -FLT_MAX like faiss.
#include <float.h>
#include <stdint.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "common.h"

namespace {

void merge_range(const float* const* D, const int64_t* const* I, int n_shards, int64_t q0, int64_t q1, int k,
                 float* D_out, int64_t* I_out) {
  // heads of the shard lists are kept in small local arrays; only the winner's head is refilled per output
  std::vector<int> pos(n_shards);
  std::vector<float> hs(n_shards);
  std::vector<int64_t> hi(n_shards);
  auto refill = [&](int s, size_t base) {
    int p = pos[s];
    while (p < k && I[s][base + p] < 0) ++p;  // skip padding
    pos[s] = p;
    if (p < k) {
      hs[s] = D[s][base + p];
      hi[s] = I[s][base + p];
    } else {
      hi[s] = -1;  // exhausted
    }
  };
  for (int64_t q = q0; q < q1; ++q) {
    const size_t base = static_cast<size_t>(q) * k;
    for (int s = 0; s < n_shards; ++s) {
      pos[s] = 0;
      refill(s, base);
    }
    for (int o = 0; o < k; ++o) {
      int best = -1;
      for (int s = 0; s < n_shards; ++s) {
        if (hi[s] < 0) continue;
        if (best < 0 || hs[s] > hs[best] || (hs[s] == hs[best] && hi[s] < hi[best])) best = s;
      }
      if (best < 0) {
        D_out[base + o] = -FLT_MAX;
        I_out[base + o] = -1;
      } else {
        D_out[base + o] = hs[best];
        I_out[base + o] = hi[best];
        ++pos[best];
        refill(best, base);
      }
    }
  }
}

}  // namespace

extern "C" int ance_merge_topk_host(const float* const* D, const int64_t* const* I, int n_shards, int64_t nq, int k,
                                    float* D_out, int64_t* I_out, int n_threads) {
  ANCE_REQUIRE(D && I && D_out && I_out, "ance_merge_topk_host: null buffer");
  ANCE_REQUIRE(n_shards > 0 && nq >= 0 && k > 0, "ance_merge_topk_host: bad shape");
  for (int s = 0; s < n_shards; ++s) ANCE_REQUIRE(D[s] && I[s], "ance_merge_topk_host: shard %d is null", s);
  if (nq == 0) return ANCE_OK;
  if (n_threads <= 0) n_threads = static_cast<int>(std::thread::hardware_concurrency());
  n_threads = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(n_threads, (nq + 1023) / 1024)));
  if (n_threads == 1) {
    merge_range(D, I, n_shards, 0, nq, k, D_out, I_out);
    return ANCE_OK;
  }
  std::vector<std::thread> th;
  const int64_t per = (nq + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    const int64_t q0 = t * per, q1 = std::min<int64_t>(nq, q0 + per);
    if (q0 >= q1) break;
    th.emplace_back(merge_range, D, I, n_shards, q0, q1, k, D_out, I_out);
  }
  for (auto& t : th) t.join();
  return ANCE_OK;
}
