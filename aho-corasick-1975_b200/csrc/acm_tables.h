/* Table images produced by the finalise step (acm_finalise.c) and consumed by the CUDA side (acm_device.cu),
 * plus the hash functions both sides must agree on bit for bit.  Plain C / CUDA C++ compatible. */
#ifndef ACM_TABLES_H
#define ACM_TABLES_H

#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define ACM_HD __host__ __device__ __forceinline__
#else
#define ACM_HD static inline
#endif

/* 64-bit finaliser (splitmix64) for the open-addressing tables living in global memory. */
ACM_HD uint64_t
acm_mix64 (uint64_t x) {
  x ^= x >> 30;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27;
  x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

/* q-gram key -> 32-bit value fed to the shared-memory filter hashes (identity for keys that fit 32 bits). */
ACM_HD uint32_t
acm_fold_key (uint64_t key) {
  return (uint32_t)key ^ ((uint32_t)(key >> 32) * 0x85EBCA6Bu);
}

#define ACM_BLOOM_C1 0x9E3779B1u
#define ACM_BLOOM_C2 0x85EBCA77u
/* word index of a folded key in a filter of nwords 32-bit words (nwords need not be a power of two) */
ACM_HD uint32_t
acm_bloom_word (uint32_t folded, uint32_t nwords) {
  uint32_t h = folded * ACM_BLOOM_C1;
#if defined(__CUDA_ARCH__)
  return __umulhi (h, nwords);
#else
  return (uint32_t)(((uint64_t)h * nwords) >> 32);
#endif
}
/* k (1..3) bit positions inside the word, from an independent multiplicative hash */
ACM_HD uint32_t
acm_bloom_mask (uint32_t folded, uint32_t k) {
  uint32_t g = folded * ACM_BLOOM_C2;
  uint32_t m = 1u << (g >> 27);
  if (k > 1)
    m |= 1u << ((g >> 22) & 31u);
  if (k > 2)
    m |= 1u << ((g >> 17) & 31u);
  return m;
}

/* One slot of the global-memory hash tables (16 bytes, read with one vector load).
 * edge table : key = (node << 32) | symbol           -> child node, keyword ending at the child (or ACM_TAB_NONE)
 * q-gram table: key = packed last q symbols of a keyword -> reverse-trie node at depth q, keyword ending there (or NONE)
 * An empty slot has node == ACM_TAB_NONE. */
typedef struct {
  uint64_t key;
  uint32_t node;
  uint32_t keyword;
} acm_slot;
#define ACM_TAB_NONE 0xFFFFFFFFu

typedef struct {
  uint32_t keyword;
  uint32_t length;
} acm_output; /* one entry of the CSR output sets of the DFA engines */

struct acm_tables {
  int engine;        /* ACM_B200_ENGINE_* */
  int width;         /* bytes per device symbol: 1, 2 or 4 */
  uint32_t nb_states, nb_keywords, lmax, lmin;
  /* --- DFA engines (width 1) --- */
  uint32_t nb_classes;        /* columns of delta; class 0 = bytes that occur in no keyword */
  uint8_t class_of_byte[256];
  uint32_t out_threshold;     /* dfa states >= out_threshold have a non-empty output set */
  uint32_t nb_dfa_states;
  int delta_entry_bytes;      /* 2 (uint16, shared memory) or 4 (uint32, global memory) */
  void *delta;                /* [nb_dfa_states][nb_classes] next dfa state */
  size_t delta_bytes;
  uint32_t *out_offsets;      /* [nb_dfa_states - out_threshold + 1] */
  acm_output *out_entries;    /* longest first */
  uint64_t nb_out_entries;
  uint32_t *dfa_of_state;     /* host state id -> dfa state (to start from a carried cursor) */
  /* --- filter engine --- */
  uint32_t q;                 /* symbols per filter window = min(lmin, 4 for bytes / 2 otherwise) */
  uint32_t *bloom;
  uint32_t bloom_words, bloom_k;
  acm_slot *qgrams;
  uint64_t qgram_slots;       /* power of two */
  acm_slot *edges;
  uint64_t edge_slots;        /* power of two */
  uint32_t nb_rev_nodes;
};

#ifdef __cplusplus
extern "C" {
#endif
struct _ac_machine;
/* Builds the images for the machine's current dictionary; returns 0 or an ACM_B200_ERR_* code. */
int acm_build_tables (struct _ac_machine *m, struct acm_tables *t, uint64_t smem_budget);
void acm_free_tables (struct acm_tables *t);
/* Device symbol of the edge entering host state s (raw value or class id). */
uint32_t acm_symbol_of_state (const struct _ac_machine *m, const struct _ac_state *s);
#ifdef __cplusplus
}
#endif
#endif
