/* flat_ip_oracle.c -- C restatement of the exact flat inner-product top-k.  TEST INFRASTRUCTURE ONLY.
 *
 * Same contract as oracle/flat_ip_oracle.py (see its header for the provenance: faiss.IndexFlatIP as called at
 * drivers/run_ann_data_gen.py:269-276,303 and drivers/run_ann_data_gen_dpr.py:238-252; faiss itself is an unpinned,
 * un-vendored dependency, so the published contract is restated and the two degrees of freedom it leaves open are
 * fixed): canonical score = fp32 inputs, products and sum in fp64, ONE rounding to fp32; order (score descending,
 * row ascending); fewer than k rows -> label -1 with the lowest float.
 *
 * An independent second restatement: the numpy oracle goes through BLAS, this file is plain loops; the CPU tests
 * require the two to agree bit for bit (tests/test_oracle_golden.py), and it is fast enough (OpenMP over queries)
 * to serve as a definition-level check at sizes the numpy brute force cannot hold in memory.
 * Built by ance_b200.build.build_oracle() into oracle/_build/liboracle.so; never loaded by the product. */
#include <float.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct {
  float s;
  int64_t r;
} cand_t;

/* a "worse" than b in the result order (score desc, row asc) */
static inline int worse(const cand_t a, const cand_t b) { return a.s < b.s || (a.s == b.s && a.r > b.r); }

static void sift_down(cand_t* h, int n, int i) { /* heap with the WORST kept candidate at the root */
  for (;;) {
    int l = 2 * i + 1, r = l + 1, m = i;
    if (l < n && worse(h[l], h[m])) m = l;
    if (r < n && worse(h[r], h[m])) m = r;
    if (m == i) return;
    cand_t t = h[i];
    h[i] = h[m];
    h[m] = t;
    i = m;
  }
}

static int cmp_result_order(const void* pa, const void* pb) {
  const cand_t a = *(const cand_t*)pa, b = *(const cand_t*)pb;
  if (a.s != b.s) return a.s > b.s ? -1 : 1;
  return a.r < b.r ? -1 : (a.r > b.r ? 1 : 0);
}

/* P [n, dim], Q [nq, dim] row-major fp32; D [nq, k] fp32, I [nq, k] int64.  Returns 0, or -1 on allocation failure. */
int flat_ip_search_exact(const float* P, int64_t n, const float* Q, int64_t nq, int dim, int k, float* D, int64_t* I) {
  int failed = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t q = 0; q < nq; ++q) {
    cand_t* heap = (cand_t*)malloc(sizeof(cand_t) * (size_t)(k > 0 ? k : 1));
    if (!heap) {
      failed = 1;
      continue;
    }
    int m = 0;
    const float* qv = Q + (size_t)q * dim;
    for (int64_t r = 0; r < n; ++r) {
      const float* pv = P + (size_t)r * dim;
      double acc = 0.0;
      for (int d = 0; d < dim; ++d) acc += (double)qv[d] * (double)pv[d];
      const cand_t c = {(float)acc, r};
      if (m < k) {
        heap[m++] = c;
        if (m == k)
          for (int i = k / 2 - 1; i >= 0; --i) sift_down(heap, k, i);
      } else if (k > 0 && worse(heap[0], c)) {
        heap[0] = c;
        sift_down(heap, k, 0);
      }
    }
    qsort(heap, (size_t)m, sizeof(cand_t), cmp_result_order);
    for (int j = 0; j < k; ++j) {
      D[(size_t)q * k + j] = j < m ? heap[j].s : -FLT_MAX;
      I[(size_t)q * k + j] = j < m ? heap[j].r : -1;
    }
    free(heap);
  }
  return failed ? -1 : 0;
}
