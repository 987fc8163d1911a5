/* TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference library through its public API
 * (aho_corasick.h:45-97) and records what its own scan loop reports.
 *
 * Compiled by oracle/Makefile together with /root/reference/aho_corasick.c (where it lies) and the
 * minimaps stand-in into oracle/_ref/libacref_{meyer,classic}.so.  Only tests/, bench.py's
 * cpu_baseline/--impl reference legs and __graft_entry__.smoke() load it.
 *
 * The loop is the reference's own usage pattern (examples/test.c:17-23): one acm_match per symbol,
 * then acm_get_match for every index 0..nb-1 (index 0 = longest, aho_corasick.c:459-466).
 * Keyword identity: the reference has no keyword id; the harness recovers the
 * "rank among distinct keywords in first-termination order" by passing value = (void *)(rank + 1)
 * to acm_insert_end_of_keyword (aho_corasick.c:357-359 keeps the first non-zero value and returns it
 * for duplicates).
 */
#include "aho_corasick.h"
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
  uint64_t end; /* index of the last symbol of the occurrence, + base */
  uint32_t id;  /* rank of the keyword (first-termination order) */
  uint32_t len; /* number of symbols of the keyword */
} refh_match;

struct arena_block {
  struct arena_block *next;
  size_t used, cap;
  unsigned char bytes[];
};

typedef struct {
  ACMachine *machine;
  const ACState *cursor;
  size_t width;     /* bytes per symbol; persistent: its address is cmp_arg */
  size_t nb_ranks;  /* distinct keywords so far */
  struct arena_block *arena;
  MatchHolder holder;
} refh;

static void *
arena_alloc (refh *h, size_t n) {
  n = (n + 7) & ~(size_t)7;
  if (!h->arena || h->arena->used + n > h->arena->cap) {
    size_t cap = n > (1u << 20) ? n : (1u << 20);
    struct arena_block *b = malloc (sizeof (*b) + cap);
    if (!b)
      abort ();
    b->next = h->arena;
    b->used = 0;
    b->cap = cap;
    h->arena = b;
  }
  void *p = h->arena->bytes + h->arena->used;
  h->arena->used += n;
  return p;
}

refh *
refh_create (size_t width) {
  refh *h = calloc (1, sizeof (*h));
  if (!h)
    return 0;
  h->width = width;
  h->machine = acm_create (ACM_CMP_DEFAULT, &h->width, 0);
  h->cursor = acm_initiate (h->machine);
  acm_matcher_init (&h->holder);
  return h;
}

void
refh_release (refh *h) {
  if (!h)
    return;
  acm_matcher_release (&h->holder);
  acm_release (h->machine);
  while (h->arena) {
    struct arena_block *b = h->arena;
    h->arena = b->next;
    free (b);
  }
  free (h);
}

int
refh_incremental (void) {
  return ACM_INCREMENTAL_STRING_MATCHING;
}

/* Inserts one keyword of len symbols; returns its rank (existing rank for a duplicate). */
uint32_t
refh_insert (refh *h, const void *symbols, size_t len) {
  unsigned char *copy = arena_alloc (h, len * h->width); /* letters are kept by pointer, aho_corasick.c:248 */
  memcpy (copy, symbols, len * h->width);
  ACState *s = acm_initiate (h->machine);
  for (size_t i = 0; i < len; i++)
    acm_insert_letter_of_keyword (&s, copy + i * h->width);
  void *prev = acm_insert_end_of_keyword (&s, (void *)(uintptr_t)(h->nb_ranks + 1), 0);
  if (prev)
    return (uint32_t)((uintptr_t)prev - 1);
  return (uint32_t)h->nb_ranks++;
}

/* Packed bulk insert: keywords i = symbols[offsets[i] .. offsets[i+1]) ; ranks (may be 0) receives the rank of each. */
void
refh_insert_many (refh *h, const void *symbols, const uint64_t *offsets, size_t nb, uint32_t *ranks) {
  for (size_t k = 0; k < nb; k++) {
    uint32_t r = refh_insert (h, (const unsigned char *)symbols + offsets[k] * h->width, (size_t)(offsets[k + 1] - offsets[k]));
    if (ranks)
      ranks[k] = r;
  }
}

/* Ranks of the keywords in the order the reference's acm_foreach_keyword (aho_corasick.c:490-531) enumerates them.  Its callback
 * has no user argument: the collector is thread-local. */
static _Thread_local uint32_t *foreach_out;
static _Thread_local size_t foreach_cap, foreach_nb;
static void
foreach_collect (MatchHolder holder) {
  if (foreach_nb < foreach_cap)
    foreach_out[foreach_nb] = (uint32_t)((uintptr_t)holder.value - 1);
  foreach_nb++;
}

size_t
refh_foreach_ranks (refh *h, uint32_t *out, size_t cap) {
  foreach_out = out;
  foreach_cap = cap;
  foreach_nb = 0;
  acm_foreach_keyword (h->machine, foreach_collect);
  return foreach_nb;
}

size_t
refh_nb_keywords (const refh *h) {
  return acm_nb_keywords (h->machine);
}

void
refh_reset_cursor (refh *h) {
  h->cursor = acm_initiate (h->machine);
}

/* Scans n symbols from the handle's carried cursor.  Writes at most cap records, returns the number found.
 * mode bit 0: call acm_get_match for every match (the reference's usage pattern); if clear only acm_match is called
 * and nothing can be recorded. */
uint64_t
refh_scan (refh *h, const void *text, uint64_t n, uint64_t base, refh_match *out, uint64_t cap, int mode) {
  const unsigned char *t = text;
  const size_t w = h->width;
  const ACState *cur = h->cursor;
  uint64_t found = 0;
  for (uint64_t i = 0; i < n; i++) {
    size_t nb = acm_match (&cur, t + i * w);
    if (!(mode & 1)) {
      found += nb;
      continue;
    }
    for (size_t j = 0; j < nb; j++) {
      acm_get_match (cur, j, &h->holder);
      if (found < cap && out)
        out[found] = (refh_match){ .end = base + i, .id = (uint32_t)((uintptr_t)h->holder.value - 1), .len = (uint32_t)h->holder.length };
      found++;
    }
  }
  h->cursor = cur;
  return found;
}

/* Multi-threaded timing helper for the reference arm: the reference allows many threads to scan one machine, each with
 * its own cursor (README.md:266,364; no lock in acm_match in the Meyer build).  Thread k scans slice k from state 0. */
struct mt_arg {
  refh *h;
  const unsigned char *text;
  uint64_t n;
  int mode;
  uint64_t found;
};

static void *
mt_worker (void *p) {
  struct mt_arg *a = p;
  const size_t w = a->h->width;
  const ACState *cur = acm_initiate (a->h->machine);
  MatchHolder holder;
  acm_matcher_init (&holder);
  uint64_t found = 0;
  for (uint64_t i = 0; i < a->n; i++) {
    size_t nb = acm_match (&cur, a->text + i * w);
    if (a->mode & 1)
      for (size_t j = 0; j < nb; j++)
        acm_get_match (cur, j, &holder);
    found += nb;
  }
  acm_matcher_release (&holder);
  a->found = found;
  return 0;
}

uint64_t
refh_scan_mt (refh *h, const void *text, uint64_t n, int nthreads, int mode, double *seconds) {
  if (nthreads < 1)
    nthreads = 1;
  pthread_t *th = calloc ((size_t)nthreads, sizeof (*th));
  struct mt_arg *args = calloc ((size_t)nthreads, sizeof (*args));
  uint64_t per = n / (uint64_t)nthreads, found = 0;
  struct timespec t0, t1;
  clock_gettime (CLOCK_MONOTONIC, &t0);
  for (int k = 0; k < nthreads; k++) {
    uint64_t lo = per * (uint64_t)k, hi = k == nthreads - 1 ? n : lo + per;
    args[k] = (struct mt_arg){ h, (const unsigned char *)text + lo * h->width, hi - lo, mode, 0 };
    pthread_create (&th[k], 0, mt_worker, &args[k]);
  }
  for (int k = 0; k < nthreads; k++) {
    pthread_join (th[k], 0);
    found += args[k].found;
  }
  clock_gettime (CLOCK_MONOTONIC, &t1);
  if (seconds)
    *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  free (th);
  free (args);
  return found;
}
