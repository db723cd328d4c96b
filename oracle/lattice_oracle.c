/* TEST INFRASTRUCTURE - CPU oracle for the permutohedral-lattice build.  Never shipped, never linked
 * into libefgh_b200.so; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.
 *
 * Plain-C restatement, for dim d=3 (d1=4), of
 *   reference nets/generate_data.py:9-54    constants (elevation matrix, std, canonical simplex)
 *   reference nets/generate_data.py:56-112  get_keys_and_barycentric
 *   reference nets/generate_data.py:128-179 per-level driver (scale, key box, build, next coords)
 *   reference nets/transforms.py:62-92      key2int / int2key mixed-radix packing
 *   reference nets/transforms.py:125-184    build_it (insertion-ordered dedup + neighbour lookup)
 * Float semantics pinned against the live reference in this container (tests/golden/, see
 * oracle/make_golden.py): MKL sgemm == k-ascending fmaf chain, then one multiply by float(std);
 * torch.round == rintf (half to even); torch.sort(descending) is stable on ties.
 * Compile with -ffp-contract=off so that only the explicit fmaf() calls fuse.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "i2i_map.h"

#define D1 4

/* generate_data.py:15-20: E = (triu(ones(4,3)) + [0; diag(-1,-2,-3)]) * diag(1/sqrt(i(i+1))), in f32. */
static void elevate_matrix(float E[4][3]) {
  float v[3];
  for (int i = 0; i < 3; ++i) v[i] = 1.0f / sqrtf((float)((i + 1) * (i + 2)));
  float L[4][3] = {{1, 1, 1}, {-1, 1, 1}, {0, -2, 1}, {0, 0, -3}};
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 3; ++c) E[r][c] = L[r][c] * v[c];
}

void efgh_oracle_elevate_matrix(float *out12) {
  float E[4][3];
  elevate_matrix(E);
  memcpy(out12, E, sizeof(E));
}

/* generate_data.py:19 */
double efgh_oracle_expected_std(void) { return 4.0 * sqrt(2.0 / 3.0); }

/* generate_data.py:26-30: canonical[i][j] = j if j <= 3-i else j-4 */
static inline int canonical(int i, int j) { return (j <= 3 - i) ? j : j - 4; }

/* One point of get_keys_and_barycentric (generate_data.py:56-112).
 * p: already-scaled xyz.  key[c][r], bary[4], diff[4] (= el_minus_gr after the greedy fix-up). */
static void point_keys(const float E[4][3], float stdf, const float p[3], int64_t key[4][4],
                       float bary[4], float diff[4]) {
  float el[4], gr[4], d0[4];
  int rank[4];
  for (int c = 0; c < 4; ++c) {
    float acc = E[c][0] * p[0];            /* :67  sgemm == k-ascending fma chain ... */
    acc = fmaf(E[c][1], p[1], acc);
    acc = fmaf(E[c][2], p[2], acc);
    el[c] = acc * stdf;                    /* ... then * float32(expected_std) */
    gr[c] = rintf(el[c] / 4.0f) * 4.0f;    /* :70  round half to even */
    d0[c] = el[c] - gr[c];                 /* :72 */
  }
  for (int c = 0; c < 4; ++c) {            /* :73-78 inverse permutation of a stable descending sort */
    int r = 0;
    for (int j = 0; j < 4; ++j) r += (d0[j] > d0[c]) || (d0[j] == d0[c] && j < c);
    rank[c] = r;
  }
  float rsf = (((gr[0] + gr[1]) + gr[2]) + gr[3]) / 4.0f; /* :81 */
  int rs = (int)rsf;
  for (int c = 0; c < 4; ++c) {            /* :83-93 */
    if (rs > 0 && rank[c] >= D1 - rs) { gr[c] -= 4.0f; rank[c] -= D1; }
    else if (rs < 0 && rank[c] < -rs) { gr[c] += 4.0f; rank[c] += D1; }
    rank[c] += rs;
  }
  float b[5] = {0, 0, 0, 0, 0};
  for (int c = 0; c < 4; ++c) diff[c] = el[c] - gr[c];       /* :96 */
  for (int c = 0; c < 4; ++c) b[3 - rank[c]] += diff[c];     /* :100 (distinct slots per point) */
  for (int c = 0; c < 4; ++c) b[4 - rank[c]] -= diff[c];     /* :101 */
  for (int s = 0; s < 5; ++s) b[s] /= 4.0f;                  /* :102 */
  b[0] += 1.0f + b[4];                                       /* :103 */
  for (int s = 0; s < 4; ++s) bary[s] = b[s];                /* :104 */
  for (int c = 0; c < 4; ++c)                                /* :106 */
    for (int r = 0; r < 4; ++r) key[c][r] = (int64_t)gr[c] + canonical(rank[c], r);
}

/* get_keys_and_barycentric over a cloud: pts (3,n) rows with stride ld; keys (4,n,4) int64,
 * bary (4,n), elmgr (4,n) - the layouts the reference returns. */
void efgh_oracle_keys(const float *pts, int64_t n, int64_t ld, int64_t *keys, float *bary, float *elmgr) {
  float E[4][3];
  elevate_matrix(E);
  float stdf = (float)efgh_oracle_expected_std();
  for (int64_t i = 0; i < n; ++i) {
    float p[3] = {pts[i], pts[ld + i], pts[2 * ld + i]}, b[4], d[4];
    int64_t k[4][4];
    point_keys(E, stdf, p, k, b, d);
    for (int c = 0; c < 4; ++c) {
      bary[c * n + i] = b[c];
      elmgr[c * n + i] = d[c];
      for (int r = 0; r < 4; ++r) keys[(c * n + i) * 4 + r] = k[c][r];
    }
  }
}

/* transforms.py:62-78 */
static inline int64_t key2int(const int64_t key[4], const int64_t maxs[4], const int64_t mins[4]) {
  int64_t res = 0;
  for (int i = 0; i < 3; ++i) {
    res += key[i] - mins[i];
    res *= maxs[i + 1] - mins[i + 1] + 1;
  }
  return res + (key[3] - mins[3]);
}

static inline int64_t floor_mod(int64_t a, int64_t b) { int64_t m = a % b; return (m != 0 && ((m < 0) != (b < 0))) ? m + b : m; }
static inline int64_t floor_div(int64_t a, int64_t b) { int64_t q = a / b; return ((a % b != 0) && ((a < 0) != (b < 0))) ? q - 1 : q; }

/* transforms.py:81-92 (Python floor semantics for % and //) */
static inline void int2key(int64_t v, const int64_t maxs[4], const int64_t mins[4], int64_t key[4]) {
  for (int i = 3; i > 0; --i) {
    int64_t s = maxs[i] - mins[i] + 1;
    key[i] = floor_mod(v, s);
    v -= key[i];
    v = floor_div(v, s);
  }
  key[0] = v;
  for (int i = 0; i < 4; ++i) key[i] += mins[i];
}

typedef struct {
  int64_t n, hash_cnt, F;
  int has_next;
  float *scaled;   /* (3,n) the level's points after `*= scale` */
  float *bary;     /* (4,n) */
  float *elmgr;    /* (4,n) */
  int64_t *loff;   /* (4,n) */
  int64_t *nbr;    /* (F,hash_cnt) */
  float *next;     /* (3,hash_cnt) next level's points (before its own `*= scale`) */
  int64_t mins[4], maxs[4];
} oracle_level;

/* One iteration of the loop at generate_data.py:128-179.
 * pts: (3,n) rows with stride ld, the previous level's output (or the input cloud).
 * scale: this level's scale; offs: (F,4) blur offsets or F=-1 for "no blur" (radius -1);
 * has_next: 0 for the last level (assign_last False). */
oracle_level *efgh_oracle_level_build(const float *pts, int64_t n, int64_t ld, double scale,
                                      const int64_t *offs, int64_t F, int has_next) {
  oracle_level *L = (oracle_level *)calloc(1, sizeof(oracle_level));
  float E[4][3];
  elevate_matrix(E);
  const double std = efgh_oracle_expected_std();
  const float stdf = (float)std, scalef = (float)scale;
  L->n = n; L->F = F; L->has_next = has_next;
  L->scaled = (float *)malloc(sizeof(float) * 3 * (n ? n : 1));
  L->bary = (float *)malloc(sizeof(float) * 4 * (n ? n : 1));
  L->elmgr = (float *)malloc(sizeof(float) * 4 * (n ? n : 1));
  L->loff = (int64_t *)malloc(sizeof(int64_t) * 4 * (n ? n : 1));
  int64_t *keys = (int64_t *)malloc(sizeof(int64_t) * 16 * (n ? n : 1)); /* [point][coord][rem] */

  for (int c = 0; c < 4; ++c) { L->mins[c] = INT64_MAX; L->maxs[c] = INT64_MIN; }
  for (int64_t i = 0; i < n; ++i) {
    float p[3], b[4], d[4];
    int64_t k[4][4];
    for (int a = 0; a < 3; ++a) { p[a] = pts[a * ld + i] * scalef; L->scaled[a * n + i] = p[a]; } /* :130 */
    point_keys(E, stdf, p, k, b, d);
    for (int c = 0; c < 4; ++c) {
      L->bary[c * n + i] = b[c];
      L->elmgr[c * n + i] = d[c];
      for (int r = 0; r < 4; ++r) {
        keys[i * 16 + c * 4 + r] = k[c][r];
        if (k[c][r] < L->mins[c]) L->mins[c] = k[c][r];     /* :135-136 */
        if (k[c][r] > L->maxs[c]) L->maxs[c] = k[c][r];
      }
    }
  }

  /* build_it pass 1 (transforms.py:152-166): insertion-ordered dedup, point-major / remainder-minor */
  i2i_map *k2i = i2i_new(), *i2k = i2i_new();
  int64_t cnt = 0;
  for (int64_t i = 0; i < n; ++i)
    for (int r = 0; r < 4; ++r) {
      int64_t key[4] = {keys[i * 16 + r], keys[i * 16 + 4 + r], keys[i * 16 + 8 + r], keys[i * 16 + 12 + r]};
      int64_t packed = key2int(key, L->maxs, L->mins);
      int64_t idx = i2i_get(k2i, packed, -1);
      if (idx == -1) {
        i2i_set(k2i, packed, cnt);
        i2i_set(i2k, cnt, packed);
        idx = cnt++;
      }
      L->loff[r * n + i] = idx;
    }
  L->hash_cnt = cnt;
  free(keys);

  /* pass 2 (transforms.py:168-180) + next-level coordinates (generate_data.py:162-163,175-178) */
  int64_t Fa = F > 0 ? F : 0;
  L->nbr = (int64_t *)malloc(sizeof(int64_t) * ((Fa * cnt) > 0 ? Fa * cnt : 1));
  L->next = (float *)malloc(sizeof(float) * 3 * (cnt ? cnt : 1));
  const float divisor = (float)(std * scale);                 /* :177 python double product -> f32 scalar */
  for (int64_t h = 0; h < cnt; ++h) {
    int64_t key[4];
    int2key(i2i_get(i2k, h, -1), L->maxs, L->mins, key);
    for (int64_t f = 0; f < Fa; ++f) {
      int64_t nk[4] = {key[0] + offs[f * 4], key[1] + offs[f * 4 + 1], key[2] + offs[f * 4 + 2], key[3] + offs[f * 4 + 3]};
      L->nbr[f * cnt + h] = i2i_get(k2i, key2int(nk, L->maxs, L->mins), -1);
    }
    if (has_next) {
      float q[4];
      for (int c = 0; c < 4; ++c) q[c] = (float)key[c] / divisor;
      for (int a = 0; a < 3; ++a) {                            /* :178  E^T (3x4) @ last (4xH) */
        float acc = E[0][a] * q[0];
        acc = fmaf(E[1][a], q[1], acc);
        acc = fmaf(E[2][a], q[2], acc);
        acc = fmaf(E[3][a], q[3], acc);
        L->next[a * cnt + h] = acc;
      }
    }
  }
  i2i_free(k2i);
  i2i_free(i2k);
  return L;
}

int64_t efgh_oracle_level_hash_cnt(const oracle_level *L) { return L->hash_cnt; }

void efgh_oracle_level_key_box(const oracle_level *L, int64_t *mins, int64_t *maxs) {
  memcpy(mins, L->mins, sizeof(L->mins));
  memcpy(maxs, L->maxs, sizeof(L->maxs));
}

/* Any output pointer may be NULL. */
void efgh_oracle_level_export(const oracle_level *L, float *scaled, float *bary, float *elmgr,
                              int64_t *loff, int64_t *nbr, float *next) {
  size_t n = (size_t)L->n, H = (size_t)L->hash_cnt;
  if (scaled) memcpy(scaled, L->scaled, sizeof(float) * 3 * n);
  if (bary) memcpy(bary, L->bary, sizeof(float) * 4 * n);
  if (elmgr) memcpy(elmgr, L->elmgr, sizeof(float) * 4 * n);
  if (loff) memcpy(loff, L->loff, sizeof(int64_t) * 4 * n);
  if (nbr && L->F > 0) memcpy(nbr, L->nbr, sizeof(int64_t) * (size_t)L->F * H);
  if (next && L->has_next) memcpy(next, L->next, sizeof(float) * 3 * H);
}

const float *efgh_oracle_level_next_ptr(const oracle_level *L) { return L->next; }

void efgh_oracle_level_free(oracle_level *L) {
  if (!L) return;
  free(L->scaled); free(L->bary); free(L->elmgr); free(L->loff); free(L->nbr); free(L->next); free(L);
}
