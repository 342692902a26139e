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

// ------------------------------------------------------------------------------------------------
// ann_training_data_N writer — replaces the per-query Python formatting of drivers/run_ann_data_gen.py:318-329:
//   f.write("{}\t{}\t{}\n".format(query_id, pos_pid, ','.join(str(neg_pid) for neg_pid in ...)))
// Lines are emitted in `order` (the shuffled query order); a query without negatives (count 0) still gets its line with an
// empty list, as the reference's join of an empty list does.
// ------------------------------------------------------------------------------------------------
#include <stdio.h>

#include <charconv>
#include <string>

extern "C" int ance_write_training_data_host(const char* path, const int64_t* qids, const int64_t* pos, const int64_t* neg,
                                             const int64_t* counts, const int64_t* order, int64_t n, int neg_stride,
                                             int64_t* lines_written) {
  ANCE_REQUIRE(path && (n == 0 || (qids && pos && neg && counts && order)) && n >= 0 && neg_stride >= 0,
               "ance_write_training_data_host: bad arguments");
  FILE* f = fopen(path, "wb");
  ANCE_REQUIRE(f != nullptr, "ance_write_training_data_host: cannot open %s", path);
  std::string buf;
  buf.reserve(1 << 20);
  char tmp[24];
  auto put = [&](int64_t v) {
    auto r = std::to_chars(tmp, tmp + sizeof(tmp), v);
    buf.append(tmp, r.ptr);
  };
  int64_t written = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t q = order[i];
    if (q < 0 || q >= n || counts[q] < 0 || counts[q] > neg_stride) {
      fclose(f);
      ance::set_error("ance_write_training_data_host: order / counts out of range at line %lld", (long long)i);
      return ANCE_ERR_INVALID;
    }
    put(qids[q]);
    buf.push_back('\t');
    put(pos[q]);
    buf.push_back('\t');
    for (int64_t c = 0; c < counts[q]; ++c) {
      if (c) buf.push_back(',');
      put(neg[q * neg_stride + c]);
    }
    buf.push_back('\n');
    ++written;
    if (buf.size() > (1 << 20) - 8192) {
      if (fwrite(buf.data(), 1, buf.size(), f) != buf.size()) { fclose(f); ance::set_error("ance_write_training_data_host: write failed"); return ANCE_ERR_INVALID; }
      buf.clear();
    }
  }
  const bool ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
  if (fclose(f) != 0 || !ok) { ance::set_error("ance_write_training_data_host: write failed"); return ANCE_ERR_INVALID; }
  if (lines_written) *lines_written = written;
  return ANCE_OK;
}
