/* CPU check of the append-only table update (acm_finalise.c: acm_patch_filter_tables): a dictionary is built, its
 * filter tables are compiled, keywords are appended in rounds (nested, suffix and extending keywords on purpose) and patched into the
 * host images in place; after every round the patched images must hold exactly what a from-scratch build of the same dictionary
 * holds -- the reverse-trie edge table and the q-gram table as maps (node numbering is identical: both number nodes in keyword
 * order), keyword lengths / offsets / pool, filter words.  Links the library's object files directly (the table compiler is
 * internal).  Usage: patch_tables_check <initial keywords> <keywords per round> <alphabet size> */
#include "acm_internal.h"
#include "acm_tables.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static uint64_t rng_state = 777;
static uint32_t rnd (void) { rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng_state >> 33); }
static const acm_slot *find (const acm_slot *tab, uint64_t n, uint64_t key) {
  uint64_t j = acm_mix64 (key) & (n - 1);
  while (tab[j].node != ACM_TAB_NONE) { if (tab[j].key == key) return &tab[j]; j = (j + 1) & (n - 1); }
  return 0;
}
static uint64_t cmp_maps (const char *name, const acm_slot *a, uint64_t na, const acm_slot *b, uint64_t nb) {
  uint64_t errs = 0, cnt = 0;
  for (uint64_t i = 0; i < na; i++) if (a[i].node != ACM_TAB_NONE) {
    cnt++;
    const acm_slot *o = find (b, nb, a[i].key);
    if (!o || o->node != a[i].node || o->keyword != a[i].keyword) { if (errs < 5) printf ("%s: key %llx patched (%x,%x) rebuilt %s (%x,%x)\n", name, (unsigned long long)a[i].key, a[i].node, a[i].keyword, o ? "" : "MISSING", o ? o->node : 0, o ? o->keyword : 0); errs++; }
  }
  printf ("%s: %llu entries, %llu differences\n", name, (unsigned long long)cnt, (unsigned long long)errs);
  return errs;
}
int main (int argc, char **argv) {
  const uint32_t nk0 = argc > 1 ? atoi (argv[1]) : 3000, add = argc > 2 ? atoi (argv[2]) : 300, alphabet = argc > 3 ? atoi (argv[3]) : 40, rounds = 5;
  size_t sz = 4;
  ACMachine *m = acm_create (ACM_CMP_DEFAULT, &sz, 0);
  uint32_t *sym = malloc (64u << 20); uint64_t *off = malloc (8u << 20);
  uint64_t at = 0; uint32_t k = 0;
  uint32_t hist[4096][12]; uint32_t hlen[4096]; uint32_t nh = 0;
  #define GEN(count) do { uint32_t k0 = k; uint64_t at0 = at; for (uint32_t i = 0; i < (count); i++) { off[k] = at; uint32_t len = 2 + rnd () % 8; uint32_t r = rnd () % 10; \
      if (r < 2 && nh) { uint32_t h = rnd () % nh; sym[at++] = rnd () % alphabet; sym[at++] = rnd () % alphabet; memcpy (sym + at, hist[h], hlen[h] * 4); at += hlen[h]; } \
      else if (r < 3 && nh && hlen[nh - 1] > 3) { memcpy (sym + at, hist[nh - 1] + 1, (hlen[nh - 1] - 1) * 4); at += hlen[nh - 1] - 1; } \
      else for (uint32_t j = 0; j < len; j++) sym[at++] = rnd () % alphabet; \
      if (at - off[k] <= 10) { hlen[nh % 4096] = at - off[k]; memcpy (hist[nh % 4096], sym + off[k], (at - off[k]) * 4); if (nh < 4096) nh++; } k++; } off[k] = at; \
      uint64_t *o2 = malloc ((k - k0 + 1) * 8); for (uint32_t i = k0; i <= k; i++) o2[i - k0] = off[i] - at0; acm_b200_insert_keywords (m, sym + at0, o2, k - k0, 0); free (o2); } while (0)
  GEN (nk0);
  struct acm_tables t1, t2;
  if (acm_build_tables (m, &t1, 200 * 1024, 195 * 1024) || !t1.builder) { printf ("no builder (engine %d)\n", t1.engine); return 2; }
  uint64_t errors = 0;
  for (uint32_t r = 0; r < rounds; r++) {
    GEN (add);
    struct acm_patch_list pl;
    int rc = acm_patch_filter_tables (m, &t1, &pl);
    if (rc) { printf ("round %u: patch refused (%d)\n", r, rc); return 3; }
    printf ("round %u: keywords %zu patches %llu\n", r, acm_nb_keywords (m), (unsigned long long)pl.nb);
    free (pl.items);
    m->option_no_patch = 1;
    if (acm_build_tables (m, &t2, 200 * 1024, 195 * 1024)) return 4;
    m->option_no_patch = 0;
    if (t1.q != t2.q || t1.nb_keywords != t2.nb_keywords) { printf ("q/nk differ\n"); errors++; }
    errors += cmp_maps ("edges patched->rebuilt", t1.edges, t1.edge_slots, t2.edges, t2.edge_slots);
    errors += cmp_maps ("edges rebuilt->patched", t2.edges, t2.edge_slots, t1.edges, t1.edge_slots);
    errors += cmp_maps ("qgrams patched->rebuilt", t1.qgrams, t1.qgram_slots, t2.qgrams, t2.qgram_slots);
    errors += cmp_maps ("qgrams rebuilt->patched", t2.qgrams, t2.qgram_slots, t1.qgrams, t1.qgram_slots);
    for (uint32_t i = 0; i < t2.nb_keywords; i++) if (t1.kw_len[i] != t2.kw_len[i] || t1.kw_off[i] != t2.kw_off[i]) { errors++; break; }
    {
      uint64_t used = 0;
      for (uint32_t i = 0; i < t2.nb_keywords; i++)
        used += t2.kw_len[i];
      if (memcmp (t1.kw_pool, t2.kw_pool, used * 4)) { printf ("pool differs\n"); errors++; }
    }
    if (t1.bloom_words == t2.bloom_words && memcmp (t1.bloom, t2.bloom, t1.bloom_words * 4)) { printf ("bloom differs\n"); errors++; }
    acm_free_tables (&t2);
  }
  printf ("errors %llu\n", (unsigned long long)errors);
  return errors != 0;
}
