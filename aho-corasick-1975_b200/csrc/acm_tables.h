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

/* Hash of the open-addressing tables living in global memory (q-gram -> node, reverse-trie edges; callers mask it with
 * slots - 1): the top half of key * 2^64/phi, byte-swapped so that the best-mixed bits of the product -- its top ones -- are the
 * low ones.  Four instructions on the GPU (IMAD.WIDE + 2 IMAD + PRMT); the splitmix64 finaliser used before cost 21, a tenth of
 * all instructions of the verification kernel (profiles/r4_verify_c5.txt). */
ACM_HD uint64_t
acm_mix64 (uint64_t x) {
  const uint32_t top = (uint32_t)((x * 0x9E3779B97F4A7C15ull) >> 32);
#if defined(__CUDA_ARCH__)
  return __byte_perm (top, 0u, 0x0123);
#else
  return __builtin_bswap32 (top);
#endif
}

/* q-gram key -> 32-bit value fed to the shared-memory filter hashes (identity for keys that fit 32 bits). */
ACM_HD uint32_t
acm_fold_key (uint64_t key) {
  return (uint32_t)key ^ ((uint32_t)(key >> 32) * 0x85EBCA6Bu);
}

#define ACM_BLOOM_C1 0x9E3779B1u
/* Blocked Bloom filter in shared memory: one 32-bit word per key, two bits inside it.
 *   p1 = folded * C1 (64-bit product): word index = mulhi (lo32 (p1), nwords)   (nwords need not be a power of two)
 *                                      bit a      = hi32 (p1) & 31              (bits 32..36 of the product: depend on every input bit)
 *                                      bit b      = lo32 (p1) & 31              (a bijection of the key's low 5 bits)
 * The bit positions sit in the low 5 bits of a register on purpose: the GPU's funnel shift takes its amount modulo 32, so the
 * kernel needs no extraction instruction, and the multiplies run on the FMA pipe next to the ALU pipe that does the shifts.
 * (Measured and dropped: a third bit per key -- fewer false positives do not pay for two more instructions per test; and a
 * second bit at a FIXED distance from the first, which lets the finalise step AND the pair together and the kernel test with one
 * shift -- 7 instead of 9 instructions per test, but the second bit then carries no information and the pass rate of the
 * config-3 filter rose from 3.5 % to 8.9 %.) */
ACM_HD uint32_t
acm_mulhi32 (uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi (a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
ACM_HD uint32_t
acm_bloom_word (uint32_t folded, uint32_t nwords) {
  return acm_mulhi32 (folded * ACM_BLOOM_C1, nwords);
}
ACM_HD uint32_t
acm_bloom_mask (uint32_t folded) {
  const uint32_t lo = folded * ACM_BLOOM_C1, hi = acm_mulhi32 (folded, ACM_BLOOM_C1);
  return (1u << (hi & 31u)) | (1u << (lo & 31u));
}
ACM_HD int
acm_bloom_test (const uint32_t *filter, uint32_t nwords, uint32_t folded) { /* what a kernel's test computes */
  const uint32_t m = acm_bloom_mask (folded);
  return (filter[acm_bloom_word (folded, nwords)] & m) == m;
}

/* Second-level filter for dictionaries too large for shared memory: same blocked layout, in global memory (L2 resident),
 * independent hash constants, 2 bits per key, consulted only by the survivors of the shared-memory level. */
#define ACM_BLOOM_C3 0xC2B2AE35u
#define ACM_BLOOM_C4 0x27D4EB2Fu
ACM_HD uint32_t
acm_bloom2_word (uint32_t folded, uint32_t nwords) {
  return acm_mulhi32 (folded * ACM_BLOOM_C3, nwords);
}
ACM_HD uint32_t
acm_bloom2_mask (uint32_t folded) {
  const uint32_t h = acm_mulhi32 (folded, ACM_BLOOM_C4);
  return (1u << (h & 31u)) | (1u << ((h >> 5) & 31u));
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
/* node field of a slot: this flag + keyword id = "only that keyword lies below: compare its remaining symbols directly" */
#define ACM_TAIL_FLAG 0x80000000u

/* Compact q-gram key set for keys that fit 32 bits (byte and 16-bit alphabets), used by the confirmation step of the filter kernel:
 * buckets of 4 keys (16 bytes = one vector load); a key lives in its home bucket or, if that is full, in the next ones.
 * Empty cells hold ACM_QSET_EMPTY; a dictionary whose q-gram equals that value sets qset_has_empty_key instead of storing it.
 * A lookup reads the home bucket and stops at the first bucket that has an empty cell. */
#define ACM_QSET_EMPTY 0xFFFFFFFFu
ACM_HD uint32_t
acm_qset_bucket (uint32_t key, uint32_t shift) {
  return (key * 0x9E3779B1u) >> shift; /* top bits of a multiplicative hash; buckets = 1 << (32 - shift) */
}

/* Stride-2 filter (byte alphabets, shortest keyword >= 4): only every second text position is tested, on a 3-byte window.
 * Every keyword CHOOSES two of its own 3-byte windows, one whose last byte lies an even number of bytes before the keyword's last
 * byte (distance dA = 0, 2, 4, ...) and one at an odd distance (dB = 1, 3, ...): whatever the parity of the position an occurrence
 * ends on, exactly one of the two windows ends on a sampled (even) position s, and the occurrence ends at s + d.  The finalise
 * step picks the distances greedily so that the windows' filter bits collide as much as possible (the fill of the shared-memory
 * filter, hence its false-positive rate, falls by more than half against "always the last window"); a keyword whose proper suffix
 * is a keyword takes that suffix's distances, so that all keywords ending on one position share them.
 * The kernel builds the 32-bit filter key with one byte permute: the three window bytes in text order, the third one repeated. */
ACM_HD uint32_t
acm_s2_key (uint32_t b0, uint32_t b1, uint32_t b2) {
  return b0 | (b1 << 8) | (b2 << 16) | (b2 << 24);
}
/* Second level, global memory (L2 resident): the distance table.  One 32-bit word per hashed 3-byte window holding two 16-bit
 * entries {valid, right, distance d, extension byte}: "a keyword whose chosen window is this one, extended by this byte on the
 * left (text[s-3]) or on the right (text[s+1]), ends d bytes after the window".  One word load turns a filter hit at s into the
 * candidate end positions s + d (exact on four bytes); a word with more than two entries continues in the next word (CONT). */
#define ACM_PAIR_C0 0x9E3779B1u
ACM_HD uint32_t
acm_pair_word (uint32_t gram3 /* 24 bits, first byte lowest */, uint32_t log2_words) {
  return (gram3 * ACM_PAIR_C0) >> (32 - log2_words);
}
#define ACM_S2D_VALID 0x8000u
#define ACM_S2D_CONT 0x4000u  /* low entry only: more entries of this word's windows follow in the next word */
#define ACM_S2D_RIGHT 0x2000u /* the extension byte is text[s+1]; otherwise text[s-3] */
#define ACM_S2D_DIST(e) (((e) >> 8) & 31u)
#define ACM_S2_DMAX 29u       /* largest distance a keyword may choose: (d >> 1) indexes the 15 bits of a candidate's distance mask */
#define ACM_S2_DMASK_ALL 0x7FFFu
/* A candidate of the verification kernels: end position in the low 48 bits, distance mask in the high 16 (0 = every distance).
 * A keyword found ending there is reported only if its own chosen distance for that parity is in the mask: every occurrence is
 * reported by exactly one candidate however many sampled windows point at its end. */
#define ACM_CAND_POS_MASK 0xFFFFFFFFFFFFull
#define ACM_S2_CAND_CAP 64u /* confirmed candidates a warp can hold per span */
/* shared memory of one warp of the stride-2 kernel: header, staged hits (16-bit each), the span's candidates */
#define ACM_S2_WARP_BYTES(hit_cap) (16u + (hit_cap) * 2u + ACM_S2_CAND_CAP * 4u)

typedef struct {
  uint32_t keyword;
  uint32_t length;
} acm_output; /* one entry of the CSR output sets of the DFA engines */

/* Host-only structures of the filter engine that survive a finalise, so that keywords appended later (Meyer-style insertions
 * between two scans) are added to the resident tables in place instead of rebuilding everything (acm_patch_filter_tables): the
 * uncompressed reverse trie with its per-node keyword counts.  Kept only for dictionaries of moderate size. */
struct acm_filter_builder {
  acm_slot *full;      /* uncompressed reverse trie: (node << 32 | symbol) -> child */
  uint64_t full_slots, full_used;
  uint32_t *parent, *depth, *count, *only_kw, *term_kw;
  size_t max_nodes;
  uint32_t nodes;
  uint64_t nq;          /* distinct q-grams (keys of the q-gram table / set / filter) */
  uint64_t pool_syms;   /* symbols used in kw_pool */
  uint64_t pool_cap_syms, kw_cap; /* capacities (host arrays are allocated with room to grow) */
  uint64_t rpool_words_used, rpool_cap_words;
  uint64_t edges_used, qset_cells;
  uint32_t keywords_done; /* keywords 0 .. keywords_done-1 are in the tables */
};

/* One in-place change of a device table, applied by a small kernel after an append-only update (acm_device.cu). */
enum { ACM_PATCH_BLOOM = 0, ACM_PATCH_BLOOM2 = 1, ACM_PATCH_QGRAMS = 2, ACM_PATCH_QSET = 3, ACM_PATCH_EDGES = 4 };
struct acm_patch {
  uint32_t array, pad;
  uint64_t index;    /* element index: 32-bit words for the filters and the q-gram set cells, 16-byte slots for the hash tables */
  uint32_t value[4]; /* one word, or a whole slot */
};
struct acm_patch_list {
  struct acm_patch *items;
  uint64_t nb, cap;
  /* the appended tails of the per-keyword arrays, uploaded as contiguous copies: [first, first + nb) */
  uint64_t kw_first, kw_nb, pool_first_sym, pool_nb_syms, rpool_first_word, rpool_nb_words;
};

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
  uint32_t max_out_records;   /* largest output set */
  uint32_t *dfa_of_state;     /* host state id -> dfa state (to start from a carried cursor) */
  /* --- filter engine --- */
  uint32_t q;                 /* symbols per filter window = min(lmin, 4 for bytes / 2 otherwise) */
  uint32_t *bloom;
  uint32_t bloom_words;
  uint32_t *bloom2;           /* optional second level in global memory (0 when the first level is selective enough) */
  uint32_t bloom2_words;
  double bloom_fp;            /* expected false-positive rate of one probe, from the actual fill of every word */
  uint32_t *bloom_s2;         /* stride-2 filter (width 1, shortest keyword >= 4): 3-byte keys, two per keyword; 0 when not applicable */
  uint32_t bloom_s2_words;
  uint32_t s2_hit_cap;        /* raw filter hits a warp can stage per 2 KiB tile: 1.5 x the expected number + 32 */
  double bloom_s2_hit_rate;   /* expected fraction of sampled positions that pass (false positives + true 3-byte windows) */
  uint32_t *s2_dist;          /* second level of the stride-2 filter: the distance table, 1 << s2_dist_log2 words */
  uint32_t s2_dist_log2;
  uint16_t *kw_dist;          /* keyword id -> chosen distances, dA | dB << 8 */
  acm_slot *qgrams;
  uint64_t qgram_slots;       /* power of two */
  uint32_t *qset;             /* same keys as a compact set (4 keys per 16-byte bucket), widths 1 and 2 only */
  uint32_t qset_shift;        /* buckets = 1 << (32 - shift) */
  uint32_t qset_has_empty_key;
  acm_slot *edges;
  uint64_t edge_slots;        /* power of two */
  uint32_t nb_rev_nodes;
  uint32_t *kw_len;           /* keyword id -> length in symbols */
  uint64_t *kw_off;           /* keyword id -> first symbol in kw_pool */
  void *kw_pool;              /* every keyword's symbols, forward, `width` bytes each */
  uint64_t kw_pool_bytes;
  uint32_t *kw_meta;          /* width 1: keyword id -> {length, first word in kw_rpool} (one 8-byte load) */
  uint32_t *kw_rpool;         /* width 1: every keyword's bytes REVERSED, each keyword padded to whole 32-bit words */
  uint64_t kw_rpool_words;
  struct acm_filter_builder *builder; /* host only (never stored in a blob); non-null: the host images above stay allocated */
};

#ifdef __cplusplus
extern "C" {
#endif
struct _ac_machine;
/* Builds the images for the machine's current dictionary; returns 0 or an ACM_B200_ERR_* code. */
int acm_build_tables (struct _ac_machine *m, struct acm_tables *t, uint64_t smem_budget, uint64_t smem_optin);
/* Adds the keywords appended since the tables were built (ids t->builder->keywords_done .. nb_keywords-1) to the host images in
 * place and lists what changed.  Returns 0, or ACM_B200_ERR_INVALID when the update cannot be done in place (then rebuild). */
int acm_patch_filter_tables (struct _ac_machine *m, struct acm_tables *t, struct acm_patch_list *patches);
void acm_free_tables (struct acm_tables *t);
/* Device symbol of the edge entering host state s (raw value or class id). */
uint32_t acm_symbol_of_state (const struct _ac_machine *m, const struct _ac_state *s);
#ifdef __cplusplus
}
#endif
#endif
