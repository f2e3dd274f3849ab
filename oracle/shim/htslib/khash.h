/* oracle/shim/htslib/khash.h -- TEST INFRASTRUCTURE ONLY.
 * A small open-addressing (linear probing) string-keyed hash map offering the macro surface
 * that the reference names (reader.cpp:33 KHASH_MAP_INIT_STR(read_name, bam1_t*), kh_init,
 * kh_get, kh_put, kh_value, kh_key, kh_del, kh_begin, kh_end, kh_exist, kh_size, kh_clear,
 * kh_destroy).  Written for this repo; it is not klib's implementation.
 */
#ifndef ORACLE_SHIM_HTSLIB_KHASH_H
#define ORACLE_SHIM_HTSLIB_KHASH_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint32_t khint_t;
typedef khint_t khiter_t;

static inline uint32_t shim_kh_strhash(const char *s)
{
   uint32_t h = 2166136261u;
   for (; *s; ++s) { h ^= (uint8_t)*s; h *= 16777619u; }
   return h;
}

#define KHASH_MAP_INIT_STR(name, val_t)                                                     \
   typedef struct kh_##name##_s {                                                           \
      khint_t cap, live, used;                                                              \
      uint8_t *state; /* 0 empty, 1 live, 2 tombstone */                                    \
      const char **keys;                                                                    \
      val_t *vals;                                                                          \
   } kh_##name##_t;                                                                         \
   static inline kh_##name##_t *kh_init_##name(void)                                        \
   { return (kh_##name##_t *)calloc(1, sizeof(kh_##name##_t)); }                            \
   static inline void kh_clear_##name(kh_##name##_t *h)                                     \
   { if (h && h->state) { memset(h->state, 0, h->cap); h->live = h->used = 0; } }           \
   static inline void kh_destroy_##name(kh_##name##_t *h)                                   \
   { if (h) { free(h->state); free((void *)h->keys); free(h->vals); free(h); } }            \
   static inline khint_t kh_get_##name(const kh_##name##_t *h, const char *key)             \
   {                                                                                        \
      if (!h->cap) return 0;                                                                \
      khint_t mask = h->cap - 1, i = shim_kh_strhash(key) & mask, n = 0;                    \
      while (h->state[i] != 0 && n++ < h->cap) {                                            \
         if (h->state[i] == 1 && strcmp(h->keys[i], key) == 0) return i;                    \
         i = (i + 1) & mask;                                                                \
      }                                                                                     \
      return h->cap;                                                                        \
   }                                                                                        \
   static inline void kh_rehash_##name(kh_##name##_t *h, khint_t ncap)                      \
   {                                                                                        \
      uint8_t *os = h->state; const char **ok = h->keys; val_t *ov = h->vals;               \
      khint_t ocap = h->cap;                                                                \
      h->state = (uint8_t *)calloc(ncap, 1);                                                \
      h->keys = (const char **)calloc(ncap, sizeof(char *));                                \
      h->vals = (val_t *)calloc(ncap, sizeof(val_t));                                       \
      h->cap = ncap; h->live = h->used = 0;                                                 \
      for (khint_t j = 0; j < ocap; ++j) if (os[j] == 1) {                                  \
         khint_t mask = ncap - 1, i = shim_kh_strhash(ok[j]) & mask;                        \
         while (h->state[i]) i = (i + 1) & mask;                                            \
         h->state[i] = 1; h->keys[i] = ok[j]; h->vals[i] = ov[j]; h->live++; h->used++;     \
      }                                                                                     \
      free(os); free((void *)ok); free(ov);                                                 \
   }                                                                                        \
   static inline khint_t kh_put_##name(kh_##name##_t *h, const char *key, int *ret)         \
   {                                                                                        \
      if ((h->used + 1) * 2 > h->cap) kh_rehash_##name(h, h->cap ? h->cap * 2 : 64);        \
      khint_t mask = h->cap - 1, i = shim_kh_strhash(key) & mask, tomb = h->cap;            \
      while (h->state[i] != 0) {                                                            \
         if (h->state[i] == 1 && strcmp(h->keys[i], key) == 0) { *ret = 0; return i; }      \
         if (h->state[i] == 2 && tomb == h->cap) tomb = i;                                  \
         i = (i + 1) & mask;                                                                \
      }                                                                                     \
      if (tomb != h->cap) { i = tomb; *ret = 2; } else { h->used++; *ret = 1; }             \
      h->state[i] = 1; h->keys[i] = key; h->live++;                                         \
      return i;                                                                             \
   }                                                                                        \
   static inline void kh_del_##name(kh_##name##_t *h, khint_t i)                            \
   { if (i < h->cap && h->state[i] == 1) { h->state[i] = 2; h->live--; } }

#define khash_t(name) kh_##name##_t
#define kh_init(name) kh_init_##name()
#define kh_destroy(name, h) kh_destroy_##name(h)
#define kh_clear(name, h) kh_clear_##name(h)
#define kh_get(name, h, k) kh_get_##name(h, k)
#define kh_put(name, h, k, r) kh_put_##name(h, k, r)
#define kh_del(name, h, k) kh_del_##name(h, k)
#define kh_exist(h, x) ((h)->state[(x)] == 1)
#define kh_key(h, x) ((h)->keys[x])
#define kh_val(h, x) ((h)->vals[x])
#define kh_value(h, x) ((h)->vals[x])
#define kh_begin(h) ((khint_t)0)
#define kh_end(h) ((h)->cap)
#define kh_size(h) ((h)->live)
#endif
