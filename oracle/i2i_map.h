/* TEST INFRASTRUCTURE - CPU oracle only, never linked into the product library.
 *
 * int64 -> int64 insert/lookup map with the behaviour the reference gets from klib khash 0.2.8
 * through reference lib/khash_int2int.h:8-33 (init / get-with-default / set / destroy):
 *   - open addressing over a power-of-two bucket array, triangular probe sequence
 *     i, i+1, i+3, i+6 ... (reference lib/khash.h:232-247 `i = (i + (++step)) & mask`),
 *   - 64 -> 32 bit fold of the key `(k>>33) ^ k ^ (k<<11)` (reference lib/khash.h:387),
 *   - grow when occupancy would pass 0.77 of the buckets (reference lib/khash.h:194,299-311).
 * The lattice build never deletes, so tombstones are not modelled.  Which bucket a key lands in is not
 * observable through get/set, so results are identical to the reference's for any input.
 *
 * When the oracle is compiled with -DEFGH_ORACLE_REF_KHASH (and -I<reference>/lib) the real reference
 * header is used instead; see oracle/Makefile (target _ref).
 */
#ifndef EFGH_ORACLE_I2I_MAP_H
#define EFGH_ORACLE_I2I_MAP_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef EFGH_ORACLE_REF_KHASH

#include "khash_int2int.h" /* the reference's own header, found through -I */
typedef void i2i_map;
static inline i2i_map *i2i_new(void) { return khash_int2int_init(); }
static inline void i2i_free(i2i_map *m) { khash_int2int_destroy(m); }
static inline int64_t i2i_get(i2i_map *m, int64_t k, int64_t dflt) {
  return (int64_t)khash_int2int_get(m, (khint64_t)k, (khint64_t)dflt);
}
static inline void i2i_set(i2i_map *m, int64_t k, int64_t v) {
  khash_int2int_set(m, (khint64_t)k, (khint64_t)v);
}

#else

typedef struct {
  uint32_t n_buckets, n_used, grow_at;
  uint8_t *used;
  int64_t *keys, *vals;
} i2i_map;

static inline uint32_t i2i_fold(int64_t key) {
  uint64_t k = (uint64_t)key;
  return (uint32_t)((k >> 33) ^ k ^ (k << 11));
}

static inline i2i_map *i2i_new(void) { return (i2i_map *)calloc(1, sizeof(i2i_map)); }

static inline void i2i_free(i2i_map *m) {
  if (!m) return;
  free(m->used); free(m->keys); free(m->vals); free(m);
}

static inline uint32_t i2i_find_slot(const i2i_map *m, int64_t key) {
  uint32_t mask = m->n_buckets - 1, i = i2i_fold(key) & mask, step = 0;
  while (m->used[i] && m->keys[i] != key) i = (i + (++step)) & mask;
  return i;
}

static inline void i2i_rehash(i2i_map *m, uint32_t new_n) {
  i2i_map old = *m;
  m->n_buckets = new_n;
  m->grow_at = (uint32_t)(new_n * 0.77 + 0.5);
  m->used = (uint8_t *)calloc(new_n, 1);
  m->keys = (int64_t *)malloc(sizeof(int64_t) * new_n);
  m->vals = (int64_t *)malloc(sizeof(int64_t) * new_n);
  for (uint32_t b = 0; b < old.n_buckets; ++b)
    if (old.used[b]) {
      uint32_t i = i2i_find_slot(m, old.keys[b]);
      m->used[i] = 1; m->keys[i] = old.keys[b]; m->vals[i] = old.vals[b];
    }
  free(old.used); free(old.keys); free(old.vals);
}

static inline int64_t i2i_get(const i2i_map *m, int64_t key, int64_t dflt) {
  if (!m->n_buckets) return dflt;
  uint32_t i = i2i_find_slot(m, key);
  return m->used[i] ? m->vals[i] : dflt;
}

static inline void i2i_set(i2i_map *m, int64_t key, int64_t val) {
  if (m->n_used >= m->grow_at) i2i_rehash(m, m->n_buckets ? m->n_buckets * 2 : 4);
  uint32_t i = i2i_find_slot(m, key);
  if (!m->used[i]) { m->used[i] = 1; m->keys[i] = key; ++m->n_used; }
  m->vals[i] = val;
}

#endif /* EFGH_ORACLE_REF_KHASH */
#endif
