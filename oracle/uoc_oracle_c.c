/*
 * uoc_oracle_c.c -- plain C restatement of the DISCRETE decisions of the clustering path, in the
 * canonical fp32 arithmetic the CUDA kernels use.  TEST INFRASTRUCTURE ONLY (the checker for the
 * bit-exact parity tests; never linked into or called by the product library).
 *
 * Why a second oracle next to oracle/uoc_oracle.py: the reference computes its dot products with
 * torch.mm, whose summation order is library defined; an arg-max / arg-min / threshold on top of
 * it is only reproducible bit for bit against a restatement with a DEFINED order.  Canonical
 * order = one sequential fmaf chain over channels k = 0..d-1 from +0.0f; cosine distance =
 * 0.5f * (1.0f - dot).  tests/test_oracle_c.py pins this file against oracle/uoc_oracle.py (and
 * through it against the reference) on inputs with margin.
 *
 * Follows lib/utils/mean_shift.py of NVlabs/UnseenObjectClustering @ f5a00c7:
 *   uoc_oracle_select_seeds   :128-189   farthest point sampling (uoc_oracle_select_seeds_init: continued from given seeds)
 *   uoc_oracle_label_seeds    :41-76     greedy epsilon-neighbourhood labelling (+ :30-38 mode)
 *   uoc_oracle_assign         :206-227   nearest seed, histogram over range(len(unique)), label-0 swap
 *   uoc_oracle_hill_climb     :79-109    mean-shift updates (double accumulation: a tolerance oracle)
 * X is planar: element (point p, channel k) at X[k * stride_d + p].
 *
 * Build: gcc -O2 -fPIC -shared -std=c11 -ffp-contract=off -fopenmp (see unseenobjectclustering_b200/build.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static float dot_chain(const float* a, int64_t sa, const float* b, int64_t sb, int d) {
  float acc = 0.0f;
  for (int k = 0; k < d; ++k) acc = fmaf(a[k * sa], b[k * sb], acc);
  return acc;
}

/* metric 0 = cosine: 0.5f * (1 - dot chain);  metric 1 = euclidean (the 'euclidean' branches of mean_shift.py:21-24,
 * 58-60,159-160,207-209): sqrtf of the chain acc = fmaf(t, t, acc), t = a_k - b_k. */
static float distance(const float* a, int64_t sa, const float* b, int64_t sb, int d, int metric) {
  if (metric == 0) return 0.5f * (1.0f - dot_chain(a, sa, b, sb, d));
  float acc = 0.0f;
  for (int k = 0; k < d; ++k) {
    const float t = a[k * sa] - b[k * sb];
    acc = fmaf(t, t, acc);
  }
  return sqrtf(acc);
}

/* mean_shift.py:128-189.  selected[m], seeds[m*d]. */
int uoc_oracle_select_seeds_metric(const float* X, int64_t n, int d, int64_t stride_d, int m, int64_t first,
                                   int64_t* selected, float* seeds, int metric) {
  if (first < 0 || first >= n) return 1;
  float* r = (float*)malloc(sizeof(float) * (size_t)n);
  float* s = (float*)malloc(sizeof(float) * (size_t)d);
  if (!r || !s) return 2;
  int64_t idx = first;
  for (int i = 0; i < m; ++i) {
    selected[i] = idx;
    for (int k = 0; k < d; ++k) { s[k] = X[k * stride_d + idx]; seeds[(size_t)i * d + k] = s[k]; }
    if (i + 1 == m) break;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n; ++p) {
      const float dist = distance(X + p, stride_d, s, 1, d, metric);
      r[p] = (i == 0) ? dist : (dist < r[p] ? dist : r[p]);       /* running min == min over columns [:i+1] (:174) */
    }
    float best = r[0];
    int64_t bi = 0;
    for (int64_t p = 1; p < n; ++p)
      if (r[p] > best) { best = r[p]; bi = p; }                  /* first maximum, like torch.argmax */
    idx = bi;
  }
  free(r);
  free(s);
  return 0;
}

/* mean_shift.py:128-189 continued from seeds chosen before (:144-149, :164-169): init [num_init][d].  selected = -1 for them. */
int uoc_oracle_select_seeds_init(const float* X, int64_t n, int d, int64_t stride_d, int m, const float* init, int num_init,
                                 int64_t* selected, float* seeds, int metric) {
  if (num_init < 1 || num_init > m) return 1;
  float* r = (float*)malloc(sizeof(float) * (size_t)n);
  float* s = (float*)malloc(sizeof(float) * (size_t)d);
  if (!r || !s) return 2;
  int64_t idx = -1;
  for (int i = 0; i < m; ++i) {
    selected[i] = (i < num_init) ? -1 : idx;
    for (int k = 0; k < d; ++k) {
      s[k] = (i < num_init) ? init[(size_t)i * d + k] : X[k * stride_d + idx];
      seeds[(size_t)i * d + k] = s[k];
    }
    if (i + 1 == m) break;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n; ++p) {
      const float dist = distance(X + p, stride_d, s, 1, d, metric);
      r[p] = (i == 0) ? dist : (dist < r[p] ? dist : r[p]);
    }
    float best = r[0];
    int64_t bi = 0;
    for (int64_t p = 1; p < n; ++p)
      if (r[p] > best) { best = r[p]; bi = p; }
    idx = bi;
  }
  free(r);
  free(s);
  return 0;
}

int uoc_oracle_select_seeds(const float* X, int64_t n, int d, int64_t stride_d, int m, int64_t first,
                            int64_t* selected, float* seeds) {
  return uoc_oracle_select_seeds_metric(X, n, d, stride_d, m, first, selected, seeds, 0);
}

/* mean_shift.py:41-76 + :30-38.  Z [m][d] row-major.  labels[m]; returns len(unique(labels)). */
int uoc_oracle_label_seeds_metric(const float* Z, int m, int d, float eps, int32_t* labels, int metric) {
  unsigned char* comp = (unsigned char*)malloc((size_t)m);
  int* cnt = (int*)malloc(sizeof(int) * (size_t)(m + 1));
  for (int j = 0; j < m; ++j) labels[j] = -1;
  int K = 0;
  for (int i = 0; i < m; ++i) {
    if (labels[i] != -1) continue;
    int any = 0;
    for (int l = 0; l < K; ++l) cnt[l] = 0;
    for (int j = 0; j < m; ++j) {
      const float dist = distance(Z + (size_t)j * d, 1, Z + (size_t)i * d, 1, d, metric);
      comp[j] = dist <= eps;
      if (comp[j] && labels[j] != -1) { cnt[labels[j]] += 1; any = 1; }
    }
    int lab;
    if (any) {                                   /* mode of the labelled members, ties -> smallest label */
      lab = 0;
      for (int l = 1; l < K; ++l)
        if (cnt[l] > cnt[lab]) lab = l;
    } else {
      lab = K++;
    }
    for (int j = 0; j < m; ++j)
      if (comp[j]) labels[j] = lab;              /* overwrite, labelled or not (:74) */
  }
  int uniq = 0;
  for (int l = 0; l < K; ++l) cnt[l] = 0;
  for (int j = 0; j < m; ++j) cnt[labels[j]] = 1;
  for (int l = 0; l < K; ++l) uniq += cnt[l];
  free(comp);
  free(cnt);
  return uniq;
}

int uoc_oracle_label_seeds(const float* Z, int m, int d, float eps, int32_t* labels) {
  return uoc_oracle_label_seeds_metric(Z, m, d, eps, labels, 0);
}

/* mean_shift.py:206-227.  labels_out[n]. */
int uoc_oracle_assign_metric(const float* X, int64_t n, int d, int64_t stride_d, const float* Z, int m,
                             const int32_t* seed_labels, int num_unique, int32_t* labels_out, int metric) {
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < n; ++p) {
    float best = 0.f;
    int bj = 0;
    for (int j = 0; j < m; ++j) {
      const float dist = distance(X + p, stride_d, Z + (size_t)j * d, 1, d, metric);
      if (j == 0 || dist < best) { best = dist; bj = j; }
    }
    labels_out[p] = seed_labels[bj];
  }
  int64_t* count = (int64_t*)calloc((size_t)(num_unique > 0 ? num_unique : 1), sizeof(int64_t));
  for (int64_t p = 0; p < n; ++p)
    if (labels_out[p] >= 0 && labels_out[p] < num_unique) count[labels_out[p]] += 1;
  int lm = 0;
  for (int i = 1; i < num_unique; ++i)
    if (count[i] > count[lm]) lm = i;
  if (lm != 0) {
    for (int64_t p = 0; p < n; ++p) {
      if (labels_out[p] == 0) labels_out[p] = lm;
      else if (labels_out[p] == lm) labels_out[p] = 0;
    }
  }
  free(count);
  return 0;
}

int uoc_oracle_assign(const float* X, int64_t n, int d, int64_t stride_d, const float* Z, int m,
                      const int32_t* seed_labels, int num_unique, int32_t* labels_out) {
  return uoc_oracle_assign_metric(X, n, d, stride_d, Z, m, seed_labels, num_unique, labels_out, 0);
}

/* mean_shift.py:79-109 in double precision (tolerance oracle for the tensor-core loop). Z [m][d] in/out.
 * metric 1: weights exp(-kappa ||x - z||^2), rows divided by max(sum of weights, 1) (:21-24, :101-105). */
int uoc_oracle_hill_climb_metric(const float* X, int64_t n, int d, int64_t stride_d, float* Z, int m, float kappa, int iters,
                                 int metric) {
  double* acc = (double*)malloc(sizeof(double) * (size_t)m * d);
  double* wsum = (double*)malloc(sizeof(double) * (size_t)m);
  if (!acc || !wsum) return 2;
  for (int it = 0; it < iters; ++it) {
#pragma omp parallel for schedule(static)
    for (int j = 0; j < m; ++j) {
      double* a = acc + (size_t)j * d;
      for (int k = 0; k < d; ++k) a[k] = 0.0;
      double ws = 0.0;
      for (int64_t p = 0; p < n; ++p) {
        double s = 0.0;
        if (metric == 0) {
          for (int k = 0; k < d; ++k) s += (double)X[k * stride_d + p] * (double)Z[(size_t)j * d + k];
        } else {
          for (int k = 0; k < d; ++k) {
            const double t = (double)X[k * stride_d + p] - (double)Z[(size_t)j * d + k];
            s -= t * t;
          }
        }
        const double w = exp((double)kappa * s);
        ws += w;
        for (int k = 0; k < d; ++k) a[k] += w * (double)X[k * stride_d + p];
      }
      wsum[j] = ws;
    }
    for (int j = 0; j < m; ++j) {
      double nrm;
      if (metric == 0) {
        double ss = 0.0;
        for (int k = 0; k < d; ++k) ss += acc[(size_t)j * d + k] * acc[(size_t)j * d + k];
        nrm = sqrt(ss);
        if (nrm < 1e-12) nrm = 1e-12;
      } else {
        nrm = wsum[j] < 1.0 ? 1.0 : wsum[j];
      }
      for (int k = 0; k < d; ++k) Z[(size_t)j * d + k] = (float)(acc[(size_t)j * d + k] / nrm);
    }
  }
  free(acc);
  free(wsum);
  return 0;
}

int uoc_oracle_hill_climb(const float* X, int64_t n, int d, int64_t stride_d, float* Z, int m, float kappa, int iters) {
  return uoc_oracle_hill_climb_metric(X, n, d, stride_d, Z, m, kappa, iters, 0);
}
