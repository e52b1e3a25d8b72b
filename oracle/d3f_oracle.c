/* TEST INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * CPU restatement ("port") of the native half of the D3Feat hot path, written from the
 * algorithm of the reference (paths relative to /root/reference), NOT copied from it:
 *
 *   orc_grid_subsample        <- cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-105
 *                                tf_custom_ops/tf_subsampling/grid_subsampling/grid_subsampling.cpp:5-97
 *                                (SampledData: grid_subsampling.h:10-80 / :10-69)
 *   orc_batch_grid_subsample  <- tf_custom_ops/tf_subsampling/grid_subsampling/grid_subsampling.cpp:101-149
 *   orc_batch_neighbors       <- tf_custom_ops/tf_neighbors/neighbors/neighbors.cpp:211-332 (result-set
 *                                semantics of nanoflann.hpp:249-253, 432-440, 1280-1289; the KD-tree is NOT
 *                                restated -- this is an exhaustive search with the same fp32 arithmetic)
 *   orc_ordered_neighbors     <- tf_custom_ops/tf_neighbors/neighbors/neighbors.cpp:58-123
 *
 * Pinning: checked bit-for-bit against the compiled reference cores (oracle/_ref, built by
 * oracle/Makefile from /root/reference) in tests/test_oracle_vs_ref.py and against the committed
 * fixtures tests/golden/*.npz generated from those cores by scripts/make_golden.py.
 *
 * Canonical orders (the reference's own order is an artefact of libstdc++ internals):
 *   - subsampled cells are emitted in ascending 64-bit cell key (per batch element);
 *   - neighbours are emitted in ascending (d2, support index).
 *
 * Compile: gcc -std=c11 -O2 -ffp-contract=off (no FMA contraction: the reference is built with
 * plain g++ -O2 for baseline x86-64, tf_custom_ops/compile_op.sh:8, so every mul/add rounds separately).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <unistd.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  uint64_t key;
  int32_t idx;
} KeyIdx;

static int cmp_keyidx(const void* a, const void* b) {
  const KeyIdx* x = (const KeyIdx*)a;
  const KeyIdx* y = (const KeyIdx*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return (x->idx > y->idx) - (x->idx < y->idx);
}

/* (size_t)floor(v) as gcc/x86-64 evaluates it for |v| < 2^63: convert through int64, reinterpret. */
static uint64_t f2u64(float v) { return (uint64_t)(int64_t)v; }

/* One cloud. Outputs must hold N rows (upper bound). feats/classes may be NULL.
 * out_keys (optional, may be NULL) receives each cell's 64-bit key. Returns M. */
int orc_grid_subsample(const float* pts, int N, const float* feats, int fdim, const int* classes,
                       int ldim, float dl, float* out_pts, float* out_feats, int* out_classes,
                       uint64_t* out_keys) {
  if (N <= 0) return 0;
  /* limits of the cloud (cloud.cpp:27-67) */
  float mn[3] = {pts[0], pts[1], pts[2]}, mx[3] = {pts[0], pts[1], pts[2]};
  for (int i = 0; i < N; ++i)
    for (int a = 0; a < 3; ++a) {
      float v = pts[3 * i + a];
      if (v < mn[a]) mn[a] = v;
      if (v > mx[a]) mx[a] = v;
    }
  /* originCorner = floor(minCorner * (1/sampleDl)) * sampleDl  (grid_subsampling.cpp:27) */
  float inv = 1 / dl;
  float org[3];
  for (int a = 0; a < 3; ++a) org[a] = floorf(mn[a] * inv) * dl;
  uint64_t NX = f2u64(floorf((mx[0] - org[0]) / dl)) + 1;
  uint64_t NY = f2u64(floorf((mx[1] - org[1]) / dl)) + 1;

  KeyIdx* ki = (KeyIdx*)malloc(sizeof(KeyIdx) * (size_t)N);
  for (int i = 0; i < N; ++i) {
    uint64_t iX = f2u64(floorf((pts[3 * i + 0] - org[0]) / dl));
    uint64_t iY = f2u64(floorf((pts[3 * i + 1] - org[1]) / dl));
    uint64_t iZ = f2u64(floorf((pts[3 * i + 2] - org[2]) / dl));
    ki[i].key = iX + NX * iY + NX * NY * iZ; /* mod 2^64, like size_t arithmetic */
    ki[i].idx = i;
  }
  qsort(ki, (size_t)N, sizeof(KeyIdx), cmp_keyidx);

  int M = 0;
  int i = 0;
  while (i < N) {
    int j = i;
    float sx = 0.f, sy = 0.f, sz = 0.f; /* point = PointXYZ() then += p in input order */
    int count = 0;
    if (feats && fdim > 0)
      for (int c = 0; c < fdim; ++c) out_feats[(size_t)M * fdim + c] = 0.f;
    if (classes && ldim > 0)
      for (int c = 0; c < ldim; ++c) out_classes[(size_t)M * ldim + c] = INT32_MIN;
    while (j < N && ki[j].key == ki[i].key) {
      int p = ki[j].idx;
      sx += pts[3 * p];
      sy += pts[3 * p + 1];
      sz += pts[3 * p + 2];
      count += 1;
      if (feats && fdim > 0)
        for (int c = 0; c < fdim; ++c) out_feats[(size_t)M * fdim + c] += feats[(size_t)p * fdim + c];
      if (classes && ldim > 0)
        for (int c = 0; c < ldim; ++c) {
          /* max_element over pair<const int,int> compares .first (the label) first and labels are
           * unique keys, so the reference returns the LARGEST LABEL PRESENT in the cell
           * (grid_subsampling.cpp:97-101), not the most frequent one. */
          int l = classes[(size_t)p * ldim + c];
          if (l > out_classes[(size_t)M * ldim + c]) out_classes[(size_t)M * ldim + c] = l;
        }
      ++j;
    }
    /* point * (1.0 / count): double reciprocal narrowed to float by operator*(PointXYZ, float) */
    float r = (float)(1.0 / (double)count);
    out_pts[3 * M] = sx * r;
    out_pts[3 * M + 1] = sy * r;
    out_pts[3 * M + 2] = sz * r;
    if (feats && fdim > 0) {
      float fc = (float)count;
      for (int c = 0; c < fdim; ++c) out_feats[(size_t)M * fdim + c] = out_feats[(size_t)M * fdim + c] / fc;
    }
    if (out_keys) out_keys[M] = ki[i].key;
    ++M;
    i = j;
  }
  free(ki);
  return M;
}

/* Stacked clouds. out_pts holds N rows (upper bound), out_batches[B]. Returns total M. */
int orc_batch_grid_subsample(const float* pts, int N, const int* batches, int B, float dl,
                             float* out_pts, int* out_batches) {
  (void)N;
  int start = 0, M = 0;
  for (int b = 0; b < B; ++b) {
    int m = orc_grid_subsample(pts + 3 * (size_t)start, batches[b], NULL, 0, NULL, 0, dl,
                               out_pts + 3 * (size_t)M, NULL, NULL, NULL);
    out_batches[b] = m;
    M += m;
    start += batches[b];
  }
  return M;
}

typedef struct {
  float d2;
  int32_t idx;
} DistIdx;

static int cmp_distidx(const void* a, const void* b) {
  const DistIdx* x = (const DistIdx*)a;
  const DistIdx* y = (const DistIdx*)b;
  if (x->d2 != y->d2) return x->d2 < y->d2 ? -1 : 1;
  return (x->idx > y->idx) - (x->idx < y->idx);
}

/* d2 = ((dx*dx) + dy*dy) + dz*dz with diff = query - support, each op rounded to fp32
 * (nanoflann.hpp:432-440 L2_Simple_Adaptor::evalMetric; identical value to PointXYZ::sq_norm). */
static inline float sqdist(const float* q, const float* s) {
  float dx = q[0] - s[0], dy = q[1] - s[1], dz = q[2] - s[2];
  float r = dx * dx;
  r = r + dy * dy;
  r = r + dz * dz;
  return r;
}

/* Phase 1: counts[Nq] and the return value = max count.
 * Phase 2 (out != NULL): fills out[Nq*cols]; rows longer than cols are truncated (nearest first),
 * shorter rows padded with pad_value. Per-batch isolation; indices are global (offset by the
 * batch element's first support, neighbors.cpp:319-321). Queries are split over pthreads. */
typedef struct {
  const float *q, *s;
  int q0, q1, s0, s1;
  float r2;
  int *counts, *out;
  int cols, pad_value, maxc;
} NbJob;

static void* nb_worker(void* arg) {
  NbJob* jb = (NbJob*)arg;
  DistIdx* buf = (DistIdx*)malloc(sizeof(DistIdx) * (size_t)(jb->s1 - jb->s0 + 1));
  int lmax = 0;
  for (int i = jb->q0; i < jb->q1; ++i) {
    int n = 0;
    for (int j = jb->s0; j < jb->s1; ++j) {
      float d2 = sqdist(jb->q + 3 * (size_t)i, jb->s + 3 * (size_t)j);
      if (d2 < jb->r2) {
        buf[n].d2 = d2;
        buf[n].idx = j;
        ++n;
      }
    }
    if (jb->counts) jb->counts[i] = n;
    if (n > lmax) lmax = n;
    if (jb->out) {
      qsort(buf, (size_t)n, sizeof(DistIdx), cmp_distidx);
      for (int c = 0; c < jb->cols; ++c)
        jb->out[(size_t)i * jb->cols + c] = c < n ? buf[c].idx : jb->pad_value;
    }
  }
  jb->maxc = lmax;
  free(buf);
  return NULL;
}

static int n_threads(void) {
  const char* e = getenv("ORC_THREADS");
  int t = e ? atoi(e) : 0;
  if (t <= 0) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    t = n > 0 ? (int)n : 1;
  }
  return t > 64 ? 64 : t;
}

static int batch_neighbors_impl(const float* q, int Nq, const float* s, int Ns, const int* qb,
                                const int* sb, int B, float radius, int* counts, int* out, int cols,
                                int pad_value) {
  float r2 = radius * radius; /* neighbors.cpp:226 */
  (void)Nq;
  (void)Ns;
  int maxc = 0, q0 = 0, s0 = 0;
  int T = n_threads();
  for (int b = 0; b < B; ++b) {
    int nq = qb[b], ns = sb[b];
    int t_use = nq < 256 ? 1 : T;
    pthread_t th[64];
    NbJob jobs[64];
    for (int t = 0; t < t_use; ++t) {
      NbJob* jb = &jobs[t];
      jb->q = q; jb->s = s;
      jb->q0 = q0 + (int)((long long)nq * t / t_use);
      jb->q1 = q0 + (int)((long long)nq * (t + 1) / t_use);
      jb->s0 = s0; jb->s1 = s0 + ns;
      jb->r2 = r2; jb->counts = counts; jb->out = out;
      jb->cols = cols; jb->pad_value = pad_value; jb->maxc = 0;
      if (t_use > 1) pthread_create(&th[t], NULL, nb_worker, jb);
      else nb_worker(jb);
    }
    for (int t = 0; t < t_use; ++t) {
      if (t_use > 1) pthread_join(th[t], NULL);
      if (jobs[t].maxc > maxc) maxc = jobs[t].maxc;
    }
    q0 += nq;
    s0 += ns;
  }
  return maxc;
}

int orc_batch_neighbors_count(const float* q, int Nq, const float* s, int Ns, const int* qb,
                              const int* sb, int B, float radius, int* counts) {
  return batch_neighbors_impl(q, Nq, s, Ns, qb, sb, B, radius, counts, NULL, 0, 0);
}

/* pad_value: Ns for the batch op (neighbors.cpp:324), -1 for the non-batch op (neighbors.cpp:117) */
void orc_batch_neighbors_fill(const float* q, int Nq, const float* s, int Ns, const int* qb,
                              const int* sb, int B, float radius, int cols, int pad_value, int* out) {
  batch_neighbors_impl(q, Nq, s, Ns, qb, sb, B, radius, NULL, out, cols, pad_value);
}
