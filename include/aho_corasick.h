/* aho_corasick.h -- drop-in public interface of the B200-native Aho-Corasick / Meyer matching engine.
 *
 * This header keeps, declaration for declaration, the public C interface of farhiongit/aho-corasick-1975
 * (reference aho_corasick.h:23-98) so that programs written against the reference (its examples/test.c and
 * examples/aho_corasick_generic_test.c) compile and run unchanged against libac75.so built from this repository.
 * It was written from the interface contract in SURVEY.md section 8(b); the implementation behind it
 * (aho-corasick-1975_b200/csrc) shares no code with the reference.
 *
 * The per-symbol entry points below run on the host, exactly like the reference's.  The data-parallel scan
 * path (whole texts, on the GPU) is the batch interface declared in acm_b200.h.
 */
#ifndef __ACM__
#define __ACM__

#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* A matched (or enumerated) keyword.  Replaces reference aho_corasick.h:23-28.
 * `letters` points to `length` pointers to the dictionary's own letter objects (the first-inserted letter of every
 * edge, reference aho_corasick.c:248,304-307); they stay valid until acm_release. */
typedef struct
{
  const void **letters;
  size_t length;
  void *value;
} MatchHolder;

typedef struct _ac_state ACState;     /* reference aho_corasick.h:30 */
typedef struct _ac_machine ACMachine; /* reference aho_corasick.h:31 */

/* Three-way comparison of two letters (reference aho_corasick.h:33).  Two letters are the same symbol iff it returns 0. */
typedef int (*CMP_TYPE) (const void *letter_a, const void *letter_b, const void *eq_arg);
/* Destructor of a letter, compatible with free (reference aho_corasick.h:34). */
typedef void (*DESTROY_TYPE) (void *letter);
/* memcmp over *(size_t *) cmp_arg bytes (reference aho_corasick.h:35, aho_corasick.c:134-138).  With it and a letter size
 * of 1, 2 or 4 bytes the GPU path consumes raw text; any other comparator goes through acm_b200_remap_text. */
extern const CMP_TYPE ACM_CMP_DEFAULT;

/* reference aho_corasick.h:45 -- `cmp` is mandatory, `cmp_arg` is borrowed for the life of the machine, `dtor` (may be 0)
 * is applied to every letter the machine stops needing. */
ACMachine *acm_create (CMP_TYPE cmp, void *cmp_arg, DESTROY_TYPE dtor);

/* reference aho_corasick.h:48 -- state 0, to start inserting a keyword or scanning a text. */
ACState *acm_initiate (ACMachine *machine);

/* reference aho_corasick.h:53 -- appends one letter to the keyword being inserted.  The letter pointer is kept when a new
 * edge is created, and handed to `dtor` at once when the edge already exists. */
void acm_insert_letter_of_keyword (ACState **state, void *letter);

/* reference aho_corasick.h:65 -- ends the keyword.  Returns 0 if the keyword had no value yet (the machine then owns
 * `value` and will apply `dtor` to it at release), otherwise the value it already had (the caller keeps `value`).
 * `*state` is reset to state 0. */
void *acm_insert_end_of_keyword (ACState **state, void *value, void (*dtor) (void *));

/* reference aho_corasick.h:70 -- feeds one symbol of the text; returns how many keywords end on it. */
size_t acm_match (const ACState **state, const void *letter);

/* reference aho_corasick.h:74,81,84 -- reusable holder; index 0 is the longest keyword ending at the current symbol. */
void acm_matcher_init (MatchHolder *matcher);
void acm_get_match (const ACState *state, size_t index, MatchHolder *matcher);
void acm_matcher_release (MatchHolder *matcher);

/* reference aho_corasick.h:87,90,93 */
size_t acm_nb_keywords (const ACMachine *machine);
void acm_foreach_keyword (const ACMachine *machine, void (*op) (MatchHolder));
void acm_release (ACMachine *machine);

/* reference aho_corasick.h:96-98 */
typedef int (*PRINT_TYPE) (FILE *, const void *letter);
void acm_print (ACMachine *machine, FILE *stream, PRINT_TYPE printer);
extern const int ACM_INCREMENTAL_STRING_MATCHING; /* 1: fail links are maintained incrementally (Meyer, 1985) */

#ifdef __cplusplus
}
#endif
#endif
