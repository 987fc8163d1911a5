/* CPU check of the stride-2 tables (acm_finalise.c: window choice, shared-memory filter, distance table, kw_dist).
 * Emulates on the host what filter_scan_s2_kernel + filter_verify_kernel do with them -- test every even position's 3-byte window
 * in the filter, probe the distance table, turn entries into candidate ends, report a keyword ending there only if its own chosen
 * distance produced the candidate -- and compares with the per-symbol host API (acm_match / acm_get_match, the reference's loop):
 * every occurrence must be reported exactly once.  Links the library's object files directly (acm_build_tables is internal). */
#include "acm_internal.h"
#include "acm_tables.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t
rnd (void) {
  rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
  return (uint32_t)(rng_state >> 33);
}

int
main (int argc, char **argv) {
  const uint32_t nk = argc > 1 ? (uint32_t)atoi (argv[1]) : 3000;
  const uint32_t alphabet = argc > 2 ? (uint32_t)atoi (argv[2]) : 256;
  const uint64_t n = argc > 3 ? (uint64_t)atoll (argv[3]) : 300000;
  const uint32_t lmaxk = argc > 4 ? (uint32_t)atoi (argv[4]) : 32;
  rng_state += nk * 7919u + alphabet;
  size_t sz = 1;
  ACMachine *m = acm_create (ACM_CMP_DEFAULT, &sz, 0);
  /* packed dictionary: random keywords of 4..lmaxk bytes; every 5th one is a suffix or an extension of an earlier one (nested) */
  uint8_t *sym = malloc ((size_t)nk * (lmaxk + 8));
  uint64_t *off = malloc (((size_t)nk + 1) * 8);
  uint64_t at = 0;
  for (uint32_t k = 0; k < nk; k++) {
    off[k] = at;
    uint32_t len = 4 + rnd () % (lmaxk - 3);
    if (k >= 5 && k % 5 == 0) {
      const uint32_t src = rnd () % k, slen = (uint32_t)(off[src + 1] - off[src]);
      if (k % 10 == 0 && slen > 4) { /* proper suffix of an earlier keyword */
        len = 4 + rnd () % (slen - 4);
        memcpy (sym + at, sym + off[src + 1] - len, len);
      } else { /* earlier keyword with bytes prepended */
        const uint32_t extra = 1 + rnd () % 6;
        len = slen + extra <= lmaxk ? slen + extra : slen;
        for (uint32_t i = 0; i < len - slen; i++)
          sym[at + i] = (uint8_t)(rnd () % alphabet);
        memcpy (sym + at + (len - slen), sym + off[src], slen);
      }
    } else
      for (uint32_t i = 0; i < len; i++)
        sym[at + i] = (uint8_t)(rnd () % alphabet);
    at += len;
  }
  off[nk] = at;
  uint32_t *ids = malloc ((size_t)nk * 4);
  if (acm_b200_insert_keywords (m, sym, off, nk, ids)) {
    fprintf (stderr, "insert failed\n");
    return 2;
  }
  /* text: random bytes with keywords planted every ~97 bytes (overlapping plants allowed) */
  uint8_t *text = malloc (n + 64);
  for (uint64_t i = 0; i < n; i++)
    text[i] = (uint8_t)(rnd () % alphabet);
  for (uint64_t p = 0; p + 64 < n; p += 60 + rnd () % 75) {
    const uint32_t k = rnd () % nk;
    memcpy (text + p, sym + off[k], off[k + 1] - off[k]);
  }
  if (argc > 6) /* engine override, e.g. "filter" for dictionaries the automatic choice gives to a DFA engine */
    snprintf (m->engine_override, sizeof m->engine_override, "%s", argv[6]);
  struct acm_tables t;
  const int rc = acm_build_tables (m, &t, 200 * 1024, (argc > 5 ? (uint64_t)atoi (argv[5]) : 195) * 1024);
  if (rc || t.engine != ACM_B200_ENGINE_FILTER || !t.bloom_s2) {
    fprintf (stderr, "no stride-2 tables (rc %d engine %d)\n", rc, t.engine);
    return 3;
  }
  /* expected: the reference's loop */
  uint64_t expected = 0, cap = 1 << 20;
  uint64_t *exp_end = malloc (cap * 8);
  uint32_t *exp_kw = malloc (cap * 4);
  {
    const ACState *cur = acm_initiate (m);
    for (uint64_t i = 0; i < n; i++) {
      const size_t nb = acm_match (&cur, &text[i]);
      const struct _ac_state *s = (const struct _ac_state *)cur;
      size_t left = nb;
      for (; left; s = s->fail)
        if (s->rank != ACM_NONE) {
          if (expected == cap) {
            cap *= 2;
            exp_end = realloc (exp_end, cap * 8);
            exp_kw = realloc (exp_kw, cap * 4);
          }
          exp_end[expected] = i;
          exp_kw[expected++] = s->rank;
          left--;
        }
    }
  }
  /* emulation */
  uint8_t *seen = calloc (expected ? expected : 1, 1);
  uint64_t hits = 0, tests = 0, cands = 0, reported = 0, errors = 0;
  const uint32_t mask = (1u << t.s2_dist_log2) - 1u;
  for (uint64_t s = 2; s < n; s += 2) {
    const uint32_t key = acm_s2_key (text[s - 2], text[s - 1], text[s]);
    tests++;
    if (!acm_bloom_test (t.bloom_s2, t.bloom_s2_words, key))
      continue;
    hits++;
    const uint32_t gram3 = text[s - 2] | ((uint32_t)text[s - 1] << 8) | ((uint32_t)text[s] << 16);
    const int left = s >= 3 ? text[s - 3] : -1, right = s + 1 < n ? text[s + 1] : -1;
    uint32_t d_seen = 0; /* two entries (left / right extension) may name the same distance: the kernel merges equal ends of a span */
    for (uint32_t idx = acm_pair_word (gram3, t.s2_dist_log2);; idx = (idx + 1) & mask) {
      const uint32_t word = t.s2_dist[idx];
      for (int h = 0; h < 2; h++) {
        const uint32_t ent = (word >> (16 * h)) & 0xFFFFu;
        if (!(ent & ACM_S2D_VALID) || ((ent & ACM_S2D_RIGHT) ? right : left) != (int)(ent & 0xFFu))
          continue;
        const uint32_t d = ACM_S2D_DIST (ent);
        const uint64_t e = s + d;
        if (e >= n || ((d_seen >> d) & 1u))
          continue;
        d_seen |= 1u << d;
        cands++;
        /* every expected occurrence ending at e whose keyword chose distance d for this parity is reported by this candidate */
        uint64_t lo = 0, hi = expected;
        while (lo < hi) {
          const uint64_t mid = (lo + hi) / 2;
          if (exp_end[mid] < e)
            lo = mid + 1;
          else
            hi = mid;
        }
        for (uint64_t i = lo; i < expected && exp_end[i] == e; i++) {
          const uint32_t dd = t.kw_dist[exp_kw[i]], dk = (e & 1) ? dd >> 8 : dd & 0xFFu;
          if (dk == d) {
            if (seen[i]++) {
              errors++;
              fprintf (stderr, "occurrence (end %llu, keyword %u) reported twice\n", (unsigned long long)e, exp_kw[i]);
            }
            reported++;
          }
        }
      }
      if (!(word & ACM_S2D_CONT))
        break;
    }
  }
  for (uint64_t i = 0; i < expected; i++)
    if (!seen[i] && exp_end[i] >= 0) {
      if (errors < 10)
        fprintf (stderr, "occurrence (end %llu, keyword %u, len %u, dist %04x) missed\n", (unsigned long long)exp_end[i], exp_kw[i], t.kw_len ? t.kw_len[exp_kw[i]] : 0, t.kw_dist[exp_kw[i]]);
      errors++;
    }
  printf ("keywords %zu occurrences %llu reported %llu candidates %llu filter_hit_rate %.4f (expected %.4f) words %u hit_cap %u errors %llu\n", acm_nb_keywords (m),
          (unsigned long long)expected, (unsigned long long)reported, (unsigned long long)cands, (double)hits / (double)tests, t.bloom_s2_hit_rate, t.bloom_s2_words, t.s2_hit_cap,
          (unsigned long long)errors);
  acm_free_tables (&t);
  acm_release (m);
  return errors ? 1 : 0;
}
