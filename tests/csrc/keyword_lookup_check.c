/* acm_b200_keyword (id) against the reference-shaped per-symbol API of the same library: for every keyword, walking its letters
 * with acm_match reaches its terminal state, where acm_get_match (state, 0, &holder) (reference aho_corasick.c:451-482, index 0 =
 * the longest keyword ending there = the keyword itself) must give the same letters (the SAME pointers: the dictionary's own
 * first-inserted letter objects, :248,477-479), length and value as acm_b200_keyword.  Also checks acm_b200_keyword_order against
 * an acm_foreach_keyword pass.  Host only (links acm_host.c, no GPU). */
#include "acm_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t rng_state = 12345;
static uint32_t
rnd (void) {
  rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
  return (uint32_t)(rng_state >> 33);
}

static uint64_t foreach_seen, foreach_errors;
static const uint32_t *expected_order;
static void
on_keyword (MatchHolder h) {
  /* value = id + 1 (set below) */
  const uint32_t id = (uint32_t)((uintptr_t)h.value - 1);
  if (expected_order[foreach_seen] != id)
    foreach_errors++;
  foreach_seen++;
}

int
main (int argc, char **argv) {
  const uint32_t nk = argc > 1 ? (uint32_t)atoi (argv[1]) : 2000;
  const uint32_t alphabet = argc > 2 ? (uint32_t)atoi (argv[2]) : 6;
  size_t sz = 1;
  ACMachine *m = acm_create (ACM_CMP_DEFAULT, &sz, 0);
  /* per-letter insertion with caller-owned letters (never copied) and a value per keyword; duplicates and nested keywords on purpose */
  unsigned char *letters = malloc ((size_t)nk * 16);
  size_t at = 0;
  uint64_t errors = 0;
  for (uint32_t k = 0; k < nk; k++) {
    const uint32_t len = 1 + rnd () % 9;
    ACState *s = acm_initiate (m);
    for (uint32_t i = 0; i < len; i++) {
      letters[at] = (unsigned char)('a' + rnd () % alphabet);
      acm_insert_letter_of_keyword (&s, &letters[at++]);
    }
    const size_t before = acm_nb_keywords (m);
    acm_insert_end_of_keyword (&s, (void *)(uintptr_t)(before + 1), 0); /* kept only if the keyword is new: value = id + 1 */
  }
  const uint32_t n = (uint32_t)acm_nb_keywords (m);
  MatchHolder by_id, by_match;
  acm_matcher_init (&by_id);
  acm_matcher_init (&by_match);
  for (uint32_t id = 0; id < n; id++) {
    if (acm_b200_keyword (m, id, &by_id) != ACM_B200_OK) {
      errors++;
      continue;
    }
    const ACState *cur = acm_initiate (m);
    size_t nb = 0;
    for (size_t i = 0; i < by_id.length; i++)
      nb = acm_match (&cur, by_id.letters[i]);
    if (!nb) {
      errors++;
      continue;
    }
    acm_get_match (cur, 0, &by_match);
    if (by_match.length != by_id.length || by_match.value != by_id.value || (uintptr_t)by_id.value != (uintptr_t)id + 1
        || memcmp (by_match.letters, by_id.letters, by_id.length * sizeof (void *)))
      errors++;
  }
  if (acm_b200_keyword (m, n, &by_id) != ACM_B200_ERR_INVALID)
    errors++;
  acm_matcher_release (&by_id);
  acm_matcher_release (&by_match);
  /* foreach order */
  uint32_t *order = malloc ((size_t)n * 4 + 4);
  uint64_t got = 0;
  if (acm_b200_keyword_order (m, order, n, &got) != ACM_B200_OK || got != n)
    errors++;
  expected_order = order;
  acm_foreach_keyword (m, on_keyword);
  if (foreach_seen != n || foreach_errors)
    errors++;
  printf ("keywords %u checked %u foreach %llu errors %llu\n", n, n, (unsigned long long)foreach_seen, (unsigned long long)errors);
  free (order);
  free (letters);
  acm_release (m);
  return errors ? 1 : 0;
}
