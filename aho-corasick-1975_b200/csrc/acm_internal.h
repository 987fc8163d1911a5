/* Internal layout shared by the host C library (acm_host.c, acm_finalise.c) and the CUDA side (acm_device.cu).
 * Nothing here is part of the ABI. */
#ifndef ACM_INTERNAL_H
#define ACM_INTERNAL_H

#include "aho_corasick.h"
#include "acm_b200.h"
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACM_NONE 0xFFFFFFFFu

/* One node of the keyword trie == one state of the machine (role of reference aho_corasick.c:44-65; layout is ours). */
struct _ac_state {
  struct _ac_machine *machine;
  struct _ac_state *parent; /* previous state; 0 for state 0 */
  void *letter;             /* first-inserted letter object of the incoming edge */
  struct _ac_state *fail;   /* f(s); 0 for state 0 */
  struct _ac_state **children; /* sorted by cmp on child->letter; points to inline_child while there is one child */
  struct _ac_state *inline_child;
  uint32_t nb_children, cap_children;
  struct _ac_state **ifs; /* Meyer'85 inverse fail set: the states x with f(x) == this */
  uint32_t nb_ifs, cap_ifs;
  uint32_t if_index; /* position of this state in fail->ifs */
  uint32_t depth;    /* number of letters from state 0 */
  uint32_t id;       /* creation order, state 0 is 0 */
  uint32_t rank;     /* keyword id (first-termination order) or ACM_NONE */
  size_t nb_outputs; /* keywords that are suffixes of this state's string, itself included */
  void *value;
  void (*value_dtor) (void *);
};

struct acm_state_block {
  struct acm_state_block *next;
  uint32_t used;
  struct _ac_state states[];
};
#define ACM_STATES_PER_BLOCK 4096u

enum acm_symbol_kind {
  ACM_SYM_CUSTOM = 0, /* user comparator: the GPU path sees class ids produced by acm_b200_remap_text */
  ACM_SYM_RAW1 = 1,   /* ACM_CMP_DEFAULT, 1-byte letters */
  ACM_SYM_RAW2 = 2,
  ACM_SYM_RAW4 = 4,
  ACM_SYM_RAWN = 8    /* ACM_CMP_DEFAULT with another letter size: treated like CUSTOM on the GPU path */
};

struct acm_device_image; /* owned by acm_device.cu */
struct acm_arena {
  struct acm_arena *next;
  size_t used, cap;
  unsigned char bytes[];
};

struct _ac_machine {
  struct _ac_state *root;
  size_t nb_sequences; /* distinct keywords */
  size_t nb_states;
  CMP_TYPE cmp;
  void *cmp_arg;
  DESTROY_TYPE dtor;
  int symbol_kind;
  size_t symbol_size;
  void *token; /* mtx_t *, serialises insertions and finalise (role of reference aho_corasick.c:81) */
  struct acm_state_block *blocks, *last_block;
  struct _ac_state **keywords; /* rank -> terminal state */
  size_t cap_keywords;
  uint32_t lmax, lmin; /* longest / shortest keyword, in letters */
  uint32_t max_depth;  /* deepest state (>= lmax while a keyword is being inserted) */
  uint64_t generation; /* bumped by every change of the trie or of the keyword set (role of `reconstruct`, :70) */
  struct _ac_state **scratch; /* work stack for the IF traversals */
  size_t nb_scratch, cap_scratch;
  /* user-comparator alphabets: stable class ids (1..nb_class, first-seen order); 0 = letter of no keyword */
  const void **class_letter; /* sorted by cmp */
  uint32_t *class_sorted_id; /* id of class_letter[k] */
  uint32_t nb_class, cap_class;
  uint32_t *class_of_state;  /* state id -> class id of its incoming letter */
  size_t class_states_done, cap_class_states;
  int bulk; /* set while acm_b200_insert_keywords loads many keywords: links are rebuilt by one BFS at the end */
  struct acm_arena *arena; /* letters copied by acm_b200_insert_keywords */
  /* a machine loaded from a blob (acm_blob.c): dictionary still packed (rank order), tables ready for upload */
  void *lazy_symbols;
  uint64_t *lazy_offsets;
  uint64_t lazy_nb;
  int lazy_tables_use_state_ids; /* the preloaded tables belong to a DFA engine (their cursor map is keyed by state id) */
  size_t owned_symbol_size;      /* cmp_arg of such a machine */
  struct acm_tables *preloaded;  /* tables of the blob, valid while generation == preloaded_generation */
  uint64_t preloaded_generation, preloaded_budget, preloaded_s2_smem;
  uint64_t last_smem_budget, last_s2_smem; /* shared-memory sizes of the last finalise (what acm_b200_save builds for) */
  /* GPU side */
  struct acm_device_image *device;
  uint64_t device_generation;
  char engine_override[16];
  uint64_t option_bloom_words, option_threads, option_stream_bytes;
  int option_no_stride2, option_no_events, option_no_tma, option_lean;
  uint64_t option_s2_smem_kb, option_s2_batches, option_s2_dist_log2;
  int force_rebuild;   /* an option that shapes the tables changed: the next finalise builds them from scratch */
  int option_no_patch; /* 0: tables are patched in place after append-only insertions where possible */
};

/* acm_host.c */
void acm_fatal (const char *function, const char *message);
void acm_lock (struct _ac_machine *m);
void acm_unlock (struct _ac_machine *m);
const struct _ac_state *acm_host_goto (const struct _ac_state *s, const void *letter);
struct _ac_state *acm_find_child (const struct _ac_state *s, const void *letter);
void acm_ensure_trie (struct _ac_machine *m);
void acm_ensure_trie_locked (struct _ac_machine *m); /* machine lock held */
struct acm_tables;
void acm_free_tables (struct acm_tables *t);

/* acm_device.cu */
void acm_device_release (struct acm_device_image *image);

#ifdef __cplusplus
}
#endif
#endif
