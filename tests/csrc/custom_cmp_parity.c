/* GPU parity for machines with a user comparator (reference aho_corasick_generic_test.c:48-54: wchar_t letters, case-insensitive).
 * The batch scan over class ids (acm_b200_remap_text + acm_b200_scan) must report exactly what the per-symbol loop
 * acm_match / acm_get_match reports on the same machine, including insertions between scans on a carried cursor.
 * Every round also prints the FNV-1a-64 of the GPU records (end, id, len as three little-endian u64, emission order -- the hash of
 * SURVEY.md Appendix B): the Python side of the test recomputes it from the ORACLE on a case-folded copy of the same workload, so
 * the GPU result is held against the reference itself, not only against this library's own host loop.
 * Exit status 0 = identical.  Usage: custom_cmp_parity <text file> */
#include "acm_b200.h"
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <wchar.h>
#include <wctype.h>

static int
alphacmp (const void *k, const void *t, const void *arg) {
  (void)arg;
  wint_t a = towlower (*(const wint_t *)k), b = towlower (*(const wint_t *)t);
  return a > b ? 1 : (a < b ? -1 : 0);
}

static uint64_t
fnv1a64_records (const ACMB200Match *r, uint64_t n) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (uint64_t i = 0; i < n; i++) {
    const uint64_t w[3] = { r[i].end, r[i].keyword, r[i].length };
    for (int k = 0; k < 3; k++)
      for (int b = 0; b < 8; b++) {
        h ^= (w[k] >> (8 * b)) & 0xFFu;
        h *= 0x100000001b3ull;
      }
  }
  return h;
}

static size_t
loop_scan (const ACState **cursor, const wchar_t *text, size_t n, uint64_t base, ACMB200Match *out, size_t cap, ACMachine *m, MatchHolder *h, size_t *rank_of_value) {
  size_t found = 0;
  (void)m;
  for (size_t i = 0; i < n; i++) {
    size_t nb = acm_match (cursor, &text[i]);
    for (size_t j = 0; j < nb; j++) {
      acm_get_match (*cursor, j, h);
      if (found < cap)
        out[found] = (ACMB200Match){ base + i, (uint32_t)(*(size_t *)h->value), (uint32_t)h->length };
      found++;
    }
  }
  (void)rank_of_value;
  return found;
}

int
main (int argc, char **argv) {
  if (argc < 2)
    return 2;
  FILE *f = fopen (argv[1], "rb");
  if (!f)
    return 2;
  static unsigned char raw[1 << 20];
  size_t nraw = fread (raw, 1, sizeof raw, f);
  fclose (f);
  /* bytes -> wchar_t letters, mixed case kept (the comparator folds it) */
  wchar_t *text = malloc (nraw * sizeof (wchar_t));
  for (size_t i = 0; i < nraw; i++)
    text[i] = (wchar_t)raw[i];

  ACMachine *m = acm_create (alphacmp, 0, 0);
  static size_t ids[4096];
  size_t nk = 0;
  const ACState *loop_cursor = acm_initiate (m), *gpu_cursor = acm_initiate (m);
  MatchHolder h;
  acm_matcher_init (&h);
  ACMB200Match *a = malloc (sizeof (*a) * 2000000), *b = malloc (sizeof (*b) * 2000000);
  uint32_t *classes = malloc (nraw * sizeof (uint32_t));
  int rc = 0;
  const size_t rounds = 4, per = nraw / rounds;
  for (size_t r = 0; r < rounds && !rc; r++) {
    /* insert the words that start in this slice (keywords grow between scans: Meyer-style updates + rebuild) */
    ACState *s = acm_initiate (m);
    for (size_t i = r * per; i < (r + 1) * per && nk < 4000;) {
      if (iswalpha ((wint_t)text[i])) {
        size_t j = i;
        while (j < nraw && iswalpha ((wint_t)text[j]))
          acm_insert_letter_of_keyword (&s, &text[j++]);
        ids[nk] = acm_nb_keywords (m);
        if (acm_insert_end_of_keyword (&s, &ids[nk], 0) == 0)
          nk++;
        i = j + 37; /* not every word */
      } else
        i++;
    }
    const wchar_t *slice = text + r * per;
    size_t na = loop_scan (&loop_cursor, slice, per, r * per, a, 2000000, m, &h, ids);
    if (acm_b200_remap_text (m, slice, sizeof (wchar_t), per, classes) != ACM_B200_OK) {
      fprintf (stderr, "remap failed\n");
      rc = 3;
      break;
    }
    uint64_t nb = 0;
    int e = acm_b200_scan (m, &gpu_cursor, classes, per, b, 2000000, &nb);
    if (e) {
      fprintf (stderr, "scan failed: %d %s\n", e, acm_b200_last_error ());
      rc = 4;
      break;
    }
    for (uint64_t k = 0; k < nb; k++)
      b[k].end += r * per; /* the loop reports absolute positions */
    if (na != nb || memcmp (a, b, na * sizeof (*a)) || loop_cursor != gpu_cursor) {
      fprintf (stderr, "round %zu: loop %zu records, gpu %llu, cursors %s\n", r, na, (unsigned long long)nb, loop_cursor == gpu_cursor ? "equal" : "DIFFER");
      rc = 1;
    } else
      printf ("round %zu: %zu keywords, %zu records identical, gpu fnv %016llx\n", r, acm_nb_keywords (m), na, (unsigned long long)fnv1a64_records (b, nb));
  }
  ACMB200Stats st;
  acm_b200_get_stats (m, &st);
  printf ("engine %d width %d finalise_count %llu\n", st.engine, st.symbol_width, (unsigned long long)st.finalise_count);
  acm_matcher_release (&h);
  acm_release (m);
  return rc;
}
