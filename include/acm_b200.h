/* acm_b200.h -- batch (GPU) entry points of libac75.so: the data-parallel replacement of the reference's scan loop.
 *
 * The reference scans a text with one call per symbol (reference examples/test.c:17-23):
 *     for each symbol:  nb = acm_match (&cursor, &text[i]);              -- aho_corasick.c:434-448 -> state_goto :167-192
 *                       for j < nb:  acm_get_match (cursor, j, &holder); -- aho_corasick.c:451-482
 * The functions below replace that loop for a whole text: the machine's goto/fail/output trie is compiled ("finalise")
 * into device tables, the text is scanned on the GPU (hand-written sm_100a kernels, no CPU fallback) and every
 * occurrence comes back as one 16-byte record.  Records are returned in the reference's own emission order:
 * end position ascending, and for one end position longest keyword first (aho_corasick.c:459-466).
 *
 * Plain C ABI: pointers, sizes, integers.  Errors are returned as codes (a CUDA failure is not a contract violation,
 * unlike the reference's ACM_ASSERT convention kept for the per-symbol API).
 */
#ifndef ACM_B200_H
#define ACM_B200_H

#include "aho_corasick.h"
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One occurrence.  The reference reports (cursor state, index) -> MatchHolder{letters,length,value}
 * (aho_corasick.h:23-28,81); a record carries the same information in position-independent form. */
typedef struct {
  uint64_t end;     /* base + index, in symbols, of the LAST symbol of the occurrence */
  uint32_t keyword; /* keyword id: 0-based rank among distinct keywords in first-termination order, i.e. the value of
                       acm_nb_keywords() just before the acm_insert_end_of_keyword call that first ended it (aho_corasick.c:346-355) */
  uint32_t length;  /* number of symbols of the keyword (MatchHolder.length, aho_corasick.c:472-474) */
} ACMB200Match;

enum {
  ACM_B200_OK = 0,
  ACM_B200_ERR_INVALID = 1,   /* null machine / bad argument */
  ACM_B200_ERR_NO_DEVICE = 2, /* no CUDA device: there is no CPU fallback */
  ACM_B200_ERR_CUDA = 3,      /* a CUDA call failed, see acm_b200_last_error */
  ACM_B200_ERR_NOMEM = 4,
  ACM_B200_ERR_ALPHABET = 5,  /* raw text given to a machine whose comparator needs acm_b200_remap_text */
  ACM_B200_ERR_CAPACITY = 6   /* not an error of the scan: more records were found than `capacity`; *nb_matches holds the total */
};

enum { /* scan engines, chosen at finalise; see DESIGN.md */
  ACM_B200_ENGINE_AUTO = 0,
  ACM_B200_ENGINE_DFA_SMEM = 1,   /* dense delta table over byte classes, resident in shared memory */
  ACM_B200_ENGINE_DFA_GLOBAL = 2, /* dense delta table in (L2-resident) global memory */
  ACM_B200_ENGINE_FILTER = 3      /* exact suffix q-gram filter in shared memory + reverse-trie verification */
};

typedef struct {
  const void *text;        /* nb_symbols symbols of acm_b200_symbol_width() bytes each; host or device memory.  Host text larger than
                              "stream_bytes" is streamed through two bounded device buffers, copy of chunk i+1 overlapping the scan of chunk i
                              (pinned host memory makes the copies asynchronous) */
  uint64_t nb_symbols;
  uint64_t lead;           /* the first `lead` symbols are left context only: occurrences ENDING inside them are not reported.
                              A shard [a,b) of a larger text is scanned as text+a-lead with lead = min(a, acm_b200_max_keyword_length()-1) */
  uint64_t base;           /* added to every reported end position (e.g. the shard's offset a-lead) */
  int text_on_device;      /* non-zero: `text` is a device pointer on `device`, 16-byte aligned */
  int matches_on_device;   /* non-zero: `matches` is a device pointer */
  ACMB200Match *matches;   /* room for `capacity` records (may be 0 with capacity 0 to only count) */
  uint64_t capacity;
  const ACState **cursor;  /* optional, in/out: scan continues from *cursor (as acm_match would, aho_corasick.c:447) and *cursor is
                              advanced to the state after the last symbol; 0 = start from state 0, nothing returned */
  void *stream;            /* cudaStream_t to run on (0 = a stream of the library's own) */
} ACMB200Scan;

typedef struct {
  int engine;               /* ACM_B200_ENGINE_* actually used */
  int symbol_width;         /* bytes per device symbol */
  uint32_t nb_states, nb_keywords, max_keyword_length, min_keyword_length, nb_classes;
  uint64_t table_bytes;     /* device bytes of the automaton tables */
  uint64_t smem_bytes;      /* dynamic shared memory of the main kernel */
  uint64_t finalise_count;  /* number of rebuild+upload cycles so far (Meyer insertions trigger one at the next scan) */
  double finalise_ms;       /* host build + upload of the last finalise */
  /* last scan, measured with CUDA events on the scan's stream */
  double scan_kernel_ms;    /* first kernel start -> last kernel end (device resident input) */
  double main_kernel_ms;    /* the text-streaming kernel(s) alone */
  double h2d_ms, d2h_ms;
  uint64_t last_nb_symbols, last_nb_matches, last_nb_candidates;
  uint64_t main_kernel_launches, total_kernel_launches; /* since machine creation */
  uint64_t fallback_count;  /* scans re-run in the filter engine's dense mode because a candidate buffer overflowed */
  double filter_fp;         /* expected pass rate of one test of the shared-memory filter on unrelated text (0 for the DFA engines) */
  uint64_t hot_spans;       /* 32 KiB spans the stride-2 kernel handed to the exact follow-up kernel because a stage overflowed (since creation) */
  uint64_t dfa_event_scans; /* DFA engines: scans whose second pass expanded the events recorded by the first instead of walking the text again */
  uint64_t filter_stride;   /* last scan, filter engine: text positions per filter test (2 = the stride-2 kernel, 1 = every position; 0 = DFA) */
  uint64_t dense_scans;     /* filter engine: scans that went through the dense mode (text dense in candidates; since creation) */
  uint64_t patch_count;     /* finalises that patched the resident tables in place instead of rebuilding them (append-only insertions) */
  uint64_t blob_loads;      /* finalises that uploaded the tables of a blob (acm_b200_load) as they were */
  uint64_t dfa_tma_scans;   /* DFA scans whose count pass staged the text through shared memory with TMA (since creation) */
  uint64_t dfa_lean_scans;  /* ... of which: the count pass recorded events only and the records were counted from the event lists */
} ACMB200Stats;

/* Number of CUDA devices visible (0 if none / no driver). */
int acm_b200_device_count (void);

/* Compiles the machine for `device` (-1 = current device) if it changed since the last call: cmp-ordered symbol remap,
 * BFS state order, dense delta table / q-gram filter + reverse trie, CSR output sets; then uploads.  Idempotent; the role of the
 * reference's lazy state_fail_state_construct guarded by `reconstruct` (aho_corasick.c:386-417).  Called implicitly by the scans. */
int acm_b200_finalise (ACMachine *machine, int device);

/* General scan. *nb_matches receives the number of occurrences found (even when larger than capacity). */
int acm_b200_scan_ex (ACMachine *machine, const ACMB200Scan *scan, uint64_t *nb_matches);

/* Convenience: host text, host records, reference order, from *cursor (may be 0). */
int acm_b200_scan (ACMachine *machine, const ACState **cursor, const void *text, uint64_t nb_symbols, ACMB200Match *matches, uint64_t capacity,
                   uint64_t *nb_matches);

/* Bulk form of acm_insert_letter_of_keyword / acm_insert_end_of_keyword (aho_corasick.h:53,65) for packed dictionaries:
 * keyword k = symbols[offsets[k] .. offsets[k+1]), letters of acm_b200_symbol_width() bytes, copied into machine-owned storage.
 * Machines created with ACM_CMP_DEFAULT over 1/2/4-byte letters and no letter destructor only.  ids[k] (optional) = keyword id. */
int acm_b200_insert_keywords (ACMachine *machine, const void *symbols, const uint64_t *offsets, uint64_t nb_keywords, uint32_t *ids);

/* keyword id -> what acm_get_match would have put in the holder (letters, length, value; reference aho_corasick.c:468-481).  The
 * holder follows the reference's rules (acm_matcher_init before, acm_matcher_release after). */
int acm_b200_keyword (const ACMachine *machine, uint32_t keyword, MatchHolder *holder);

/* The keyword ids in the order acm_foreach_keyword enumerates the keywords (reference aho_corasick.c:490-531): ids[k] = id of the
 * k-th keyword its callback receives.  capacity 0 only counts (*nb). */
int acm_b200_keyword_order (const ACMachine *machine, uint32_t *ids, uint64_t capacity, uint64_t *nb);

/* A finalised dictionary on disk (the reference has no serialisation, aho_corasick.h:45-98).  acm_b200_save writes the packed
 * dictionary and the table images of the machine's current dictionary (no GPU needed); acm_b200_load returns a machine -- created as
 * by acm_create (ACM_CMP_DEFAULT, &letter_size, 0) -- whose first batch scan uploads those images without inserting or building
 * anything; the keyword trie behind the per-symbol API is rebuilt from the packed dictionary the first time it is needed.  Keyword
 * ids are preserved; user values are not stored.  ACM_CMP_DEFAULT machines over 1/2/4-byte letters only. */
int acm_b200_save (ACMachine *machine, const char *path);
ACMachine *acm_b200_load (const char *path, int *error);

/* Bytes per symbol the batch scan expects: the letter size for ACM_CMP_DEFAULT machines with 1/2/4-byte letters, else 4
 * (class ids produced by acm_b200_remap_text). */
size_t acm_b200_symbol_width (const ACMachine *machine);
uint32_t acm_b200_max_keyword_length (const ACMachine *machine);

/* For machines with a user comparator (or a default comparator over letters that are not 1/2/4 bytes): maps `nb` letters of
 * `letter_size` bytes to uint32 class ids with the machine's own comparator (equal under cmp <=> same id; letters that occur in no
 * keyword get id 0).  The ids are what acm_b200_scan consumes for such machines. */
int acm_b200_remap_text (ACMachine *machine, const void *letters, size_t letter_size, uint64_t nb, uint32_t *class_ids);

/* Tuning / test knobs: "engine" = auto|dfa_smem|dfa_global|filter ; "bloom_words", "threads";
 * "stream_bytes" = size of the chunks host text is streamed in (default 256 MiB; host texts up to that size are copied whole);
 * "dfa_events" = 0 makes the second pass of the DFA engines walk the text again instead of expanding recorded events;
 * "stride2" = 0 keeps the one-test-per-position filter kernel where the stride-2 kernel would apply (byte alphabet, shortest keyword
 * >= 4 bytes); "s2_smem_kb" = shared-memory carve-out the stride-2 kernel is sized for (default 196: the rest of the SM stays L1). */
int acm_b200_set_option (ACMachine *machine, const char *key, const char *value);
int acm_b200_get_stats (ACMachine *machine, ACMB200Stats *stats);
const char *acm_b200_last_error (void);
const char *acm_b200_version (void);

#ifdef __cplusplus
}
#endif
#endif
