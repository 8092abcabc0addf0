/* CPU oracle: chunk-free exact nearest-neighbour search for the Chamfer loss.
 * TEST INFRASTRUCTURE ONLY - never linked into, or called by, the product library.
 *
 * Restates modules/loss/chamfer_distance.py:14-23 of the reference per pair:
 *     diff = p1 - p2                      (:14)
 *     dist = (dx*dx + dy*dy) + dz*dz      (:15, torch.sum over the last dim of 3, left to right)
 *     v    = sqrt(dist)                   (:19-20, sqrt BEFORE the min)
 *     min over the other cloud, first index on ties (:22-23, torch.min semantics)
 * Every operation is a separately rounded IEEE fp32 op: build with -ffp-contract=off so the
 * compiler cannot fuse mul+add (an FMA chain differs from the reference in ~21% of pairs by 1 ulp).
 * Memory is O(P + M) instead of the reference's O(P*M); results are bit-identical (checked against
 * tests/golden/hotpath_golden.npz by tests/test_oracle_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline float pair_dist(const float* a, const float* b) {
  float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  float xx = dx * dx, yy = dy * dy, zz = dz * dz;
  float s = xx + yy;
  s = s + zz;
  return sqrtf(s);
}

/* p1 (B,P,3), p2 (B,M,3) -> min1 (B,P), idx1 (B,P), min2 (B,M), idx2 (B,M).  Returns 0. */
int vpn_oracle_chamfer_nn(const float* p1, const float* p2, int B, int P, int M,
                          float* min1, int64_t* idx1, float* min2, int64_t* idx2, int nthreads) {
  if (B < 0 || P <= 0 || M <= 0) return -1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
  int nt = omp_get_max_threads();
#else
  int nt = 1;
#endif
  float* cmin = (float*)malloc((size_t)nt * M * sizeof(float));
  int64_t* cidx = (int64_t*)malloc((size_t)nt * M * sizeof(int64_t));
  if (!cmin || !cidx) { free(cmin); free(cidx); return -2; }
  for (int b = 0; b < B; ++b) {
    const float* a = p1 + (size_t)b * P * 3;
    const float* t = p2 + (size_t)b * M * 3;
#pragma omp parallel
    {
#ifdef _OPENMP
      int tid = omp_get_thread_num(), nth = omp_get_num_threads();
#else
      int tid = 0, nth = 1;
#endif
      float* lm = cmin + (size_t)tid * M;
      int64_t* li = cidx + (size_t)tid * M;
      for (int j = 0; j < M; ++j) { lm[j] = INFINITY; li[j] = 0; }
      /* contiguous, ordered row ranges per thread: merging in thread order keeps the first index */
      int lo = (int)((int64_t)P * tid / nth), hi = (int)((int64_t)P * (tid + 1) / nth);
      for (int i = lo; i < hi; ++i) {
        float best = INFINITY; int64_t bj = 0;
        for (int j = 0; j < M; ++j) {
          float v = pair_dist(a + 3 * i, t + 3 * j);
          if (v < best) { best = v; bj = j; }
          if (v < lm[j]) { lm[j] = v; li[j] = i; }
        }
        /* torch.min propagates NaN; a NaN row yields NaN with the index of the first NaN */
        min1[(size_t)b * P + i] = best; idx1[(size_t)b * P + i] = bj;
      }
#pragma omp barrier
#pragma omp for schedule(static)
      for (int j = 0; j < M; ++j) {
        float best = INFINITY; int64_t bi = 0;
        for (int k = 0; k < nth; ++k) {
          float v = cmin[(size_t)k * M + j];
          if (v < best) { best = v; bi = cidx[(size_t)k * M + j]; }
        }
        min2[(size_t)b * M + j] = best; idx2[(size_t)b * M + j] = bi;
      }
    }
  }
  free(cmin); free(cidx);
  return 0;
}

int vpn_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
