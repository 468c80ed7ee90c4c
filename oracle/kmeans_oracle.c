/*
 * C restatement of the KMeans.fit() Lloyd loop — TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may link or execute this file.  The product library (libb2kmeans.so) never does.
 *
 * Same semantics as oracle/kmeans_oracle.py (which documents the reference call sites:
 * spark_rapids_ml/clustering.py:381-415 fit, :584-601 predict; arithmetic lives in the
 * absent third-party cuML 25.12 KMeansMG).  fp64 evaluation from f32 inputs, lowest index
 * on ties, empty cluster keeps its previous center, stop on sum_j||dc_j||^2 < tol.
 * tests/test_oracle.py checks this file against the NumPy oracle and the reference's
 * known-answer vectors (tests/golden/kmeans_known_answers.json).
 *
 * Built by oracle/Makefile into oracle/_build/libkmeans_oracle.so (OpenMP: uses every host
 * thread it is given; OMP_NUM_THREADS bounds it).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* bench.py's CPU arm: explicit thread count (torchrun exports OMP_NUM_THREADS=1, which must not decide the baseline). */
void oracle_set_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
}

/* Parallel first touch: copy src into dst with the same static row partitioning the Lloyd loops use, so that on a
 * multi-socket host every thread's rows live on its own NUMA node. */
void oracle_parallel_copy(float* dst, const float* src, long long n, int d) {
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < n; ++i) {
    for (int t = 0; t < d; ++t) dst[i * d + t] = src[i * d + t];
  }
}


/* labels[i] = argmin_j ||x_i - c_j||^2 (fp64, expanded form), mind[i] = that distance.
 * mind may be NULL. */
void oracle_assign(const float* X, int64_t n, int d, const float* C, int k,
                   int32_t* labels, double* mind) {
  /* Centres are processed JB at a time with the centre index as the vector dimension: every (row, centre) dot product is
   * still the sequential sum over t = 0..d-1 in fp64 (same order, same rounding as a scalar loop), but JB of them run
   * as independent SIMD lanes, which is what lets the compiler vectorise a reduction it must not re-associate. */
  enum { JB = 32 };
  const int kp = (k + JB - 1) / JB * JB;
  double* cn = (double*)malloc(sizeof(double) * (size_t)k);
  double* Ct = (double*)calloc((size_t)d * kp, sizeof(double)); /* [d][kp], zero padded */
  for (int j = 0; j < k; ++j) {
    double s = 0.0;
    for (int t = 0; t < d; ++t) {
      s += (double)C[(size_t)j * d + t] * (double)C[(size_t)j * d + t];
      Ct[(size_t)t * kp + j] = (double)C[(size_t)j * d + t];
    }
    cn[j] = s;
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const float* x = X + (size_t)i * d;
    double xn = 0.0;
    for (int t = 0; t < d; ++t) xn += (double)x[t] * (double)x[t];
    double best = DBL_MAX;
    int bj = 0;
    for (int j0 = 0; j0 < k; j0 += JB) {
      double dot[JB];
      for (int jj = 0; jj < JB; ++jj) dot[jj] = 0.0;
      for (int t = 0; t < d; ++t) {
        const double xt = (double)x[t];
        const double* ct = Ct + (size_t)t * kp + j0;
#pragma omp simd
        for (int jj = 0; jj < JB; ++jj) dot[jj] += xt * ct[jj];
      }
      const int jn = (k - j0 < JB) ? (k - j0) : JB;
      for (int jj = 0; jj < jn; ++jj) { /* index order, strict '<': lowest index wins ties */
        double dist = xn + cn[j0 + jj] - 2.0 * dot[jj];
        if (dist < 0.0) dist = 0.0;
        if (dist < best) { best = dist; bj = j0 + jj; }
      }
    }
    labels[i] = bj;
    if (mind) mind[i] = best;
  }
  free(Ct);
  free(cn);
}

/* One rank's partial sums: S[k*d] += sum of rows per label, w[k] += counts. */
void oracle_partial_sums(const float* X, int64_t n, int d, const int32_t* labels, int k,
                         double* S, double* w) {
#ifdef _OPENMP
  int nt = omp_get_max_threads();
#else
  int nt = 1;
#endif
  size_t kd = (size_t)k * d;
  double* Sl = (double*)calloc((size_t)nt * kd, sizeof(double));
  double* wl = (double*)calloc((size_t)nt * k, sizeof(double));
#pragma omp parallel
  {
#ifdef _OPENMP
    int tid = omp_get_thread_num();
#else
    int tid = 0;
#endif
    double* Sm = Sl + (size_t)tid * kd;
    double* wm = wl + (size_t)tid * k;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      int j = labels[i];
      const float* x = X + (size_t)i * d;
      double* s = Sm + (size_t)j * d;
      for (int t = 0; t < d; ++t) s[t] += (double)x[t];
      wm[j] += 1.0;
    }
  }
  for (int t = 0; t < nt; ++t) {
    for (size_t e = 0; e < kd; ++e) S[e] += Sl[(size_t)t * kd + e];
    for (int j = 0; j < k; ++j) w[j] += wl[(size_t)t * k + j];
  }
  free(Sl);
  free(wl);
}

/* Full Lloyd loop on one partition.  C (k*d f32) holds the initial centers on entry and the
 * final ones on exit.  Returns n_iter.  tol == 0 is mapped to FLT_MIN by the caller
 * (reference: clustering.py:113-123).  inertia_out/labels_out (final assign) may be NULL. */
int oracle_lloyd(const float* X, int64_t n, int d, float* C, int k, int max_iter, double tol,
                 double* inertia_out, int32_t* labels_out) {
  size_t kd = (size_t)k * d;
  int32_t* lab = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  double* S = (double*)malloc(sizeof(double) * kd);
  double* w = (double*)malloc(sizeof(double) * (size_t)k);
  int n_iter = 0;
  for (int it = 1; it <= max_iter; ++it) {
    oracle_assign(X, n, d, C, k, lab, NULL);
    memset(S, 0, sizeof(double) * kd);
    memset(w, 0, sizeof(double) * (size_t)k);
    oracle_partial_sums(X, n, d, lab, k, S, w);
    double shift = 0.0;
    for (int j = 0; j < k; ++j) {
      for (int t = 0; t < d; ++t) {
        float old = C[(size_t)j * d + t];
        float nw = old;
        if (w[j] > 0.0) nw = (float)(S[(size_t)j * d + t] / w[j]);
        double df = (double)nw - (double)old;
        shift += df * df;
        C[(size_t)j * d + t] = nw;
      }
    }
    n_iter = it;
    if (shift < tol) break;
  }
  if (inertia_out || labels_out) {
    double* md = (double*)malloc(sizeof(double) * (size_t)n);
    oracle_assign(X, n, d, C, k, lab, md);
    if (inertia_out) {
      double s = 0.0;
      for (int64_t i = 0; i < n; ++i) s += md[i];
      *inertia_out = s;
    }
    if (labels_out) memcpy(labels_out, lab, sizeof(int32_t) * (size_t)n);
    free(md);
  }
  free(lab);
  free(S);
  free(w);
  return n_iter;
}
