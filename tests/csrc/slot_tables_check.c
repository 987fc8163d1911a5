/* CPU check of the two open-addressing tables the verification kernel walks (q-gram -> node, reverse-trie edges; acm_tables.h):
 * they must be at most a third full right after a build, every stored key must be found by the probe sequence the GPU uses
 * (acm_mix64 (key) & (slots - 1), then linear), and the probe sequences must be short -- a miss, the usual end of a walk, runs up
 * to the next empty slot, and on the GPU a warp waits for its slowest lane.  Links the library's object files directly.
 * Usage: slot_tables_check <keywords> <alphabet size> <symbol bytes: 1|4> */
#include "acm_internal.h"
#include "acm_tables.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t rng_state = 4242;
static uint32_t
rnd (void) {
  rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
  return (uint32_t)(rng_state >> 33);
}

static uint64_t
check_table (const char *name, const acm_slot *tab, uint64_t slots) {
  uint64_t used = 0, errors = 0, hit_probes = 0, longest_hit = 0, miss_probes = 0, longest_miss = 0;
  for (uint64_t i = 0; i < slots; i++) {
    /* a miss that starts at slot i ends at the next empty slot */
    uint64_t run = 1;
    for (uint64_t j = i; tab[j].node != ACM_TAB_NONE; j = (j + 1) & (slots - 1))
      run++;
    miss_probes += run;
    if (run > longest_miss)
      longest_miss = run;
    if (tab[i].node == ACM_TAB_NONE)
      continue;
    used++;
    uint64_t j = acm_mix64 (tab[i].key) & (slots - 1), probes = 1;
    while (j != i && tab[j].node != ACM_TAB_NONE)
      j = (j + 1) & (slots - 1), probes++;
    if (j != i)
      errors++; /* an empty slot before the key: the GPU's lookup would miss it */
    hit_probes += probes;
    if (probes > longest_hit)
      longest_hit = probes;
  }
  const double fill = (double)used / (double)slots, mean_hit = used ? (double)hit_probes / (double)used : 0, mean_miss = (double)miss_probes / (double)slots;
  printf ("%s: %llu of %llu slots used (%.3f), probes per hit %.3f (longest %llu), per miss %.3f (longest %llu), unreachable keys %llu\n", name, (unsigned long long)used,
          (unsigned long long)slots, fill, mean_hit, (unsigned long long)longest_hit, mean_miss, (unsigned long long)longest_miss, (unsigned long long)errors);
  if (slots > 64 && fill > 0.34)
    errors++;
  if (mean_hit > 1.35 || mean_miss > 1.8 || longest_miss > 40)
    errors++;
  return errors;
}

int
main (int argc, char **argv) {
  const uint32_t nk = argc > 1 ? (uint32_t)atoi (argv[1]) : 20000, alphabet = argc > 2 ? (uint32_t)atoi (argv[2]) : 256;
  size_t sz = argc > 3 ? (size_t)atoi (argv[3]) : 1;
  ACMachine *m = acm_create (ACM_CMP_DEFAULT, &sz, 0);
  unsigned char *sym = malloc ((size_t)nk * 40 * sz);
  uint64_t *off = malloc (((size_t)nk + 1) * 8), at = 0;
  for (uint32_t k = 0; k < nk; k++) { /* random keywords of 4..32 symbols; every 4th one extends an earlier one to the left (shared suffixes: real trie nodes) */
    off[k] = at;
    uint32_t len = 4 + rnd () % 29, copy = 0;
    const uint32_t src = k >= 4 && k % 4 == 0 ? rnd () % k : 0;
    if (k >= 4 && k % 4 == 0 && off[src + 1] - off[src] < len)
      copy = (uint32_t)(off[src + 1] - off[src]);
    for (uint32_t i = 0; i < len - copy; i++) {
      const uint32_t s = rnd () % alphabet;
      memcpy (sym + (at + i) * sz, &s, sz);
    }
    if (copy)
      memcpy (sym + (at + len - copy) * sz, sym + off[src] * sz, (size_t)copy * sz);
    at += len;
  }
  off[nk] = at;
  if (acm_b200_insert_keywords (m, sym, off, nk, 0)) {
    fprintf (stderr, "insert failed\n");
    return 2;
  }
  snprintf (m->engine_override, sizeof m->engine_override, "filter");
  struct acm_tables t;
  if (acm_build_tables (m, &t, 200 * 1024, 195 * 1024) || t.engine != ACM_B200_ENGINE_FILTER) {
    fprintf (stderr, "no filter tables\n");
    return 3;
  }
  uint64_t errors = check_table ("qgrams", t.qgrams, t.qgram_slots);
  errors += check_table ("edges", t.edges, t.edge_slots);
  printf ("errors %llu\n", (unsigned long long)errors);
  return errors ? 1 : 0;
}
