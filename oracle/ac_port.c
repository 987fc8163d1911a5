/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference's scan path in plain C.
 *
 * This file is the oracle's own statement of WHAT farhiongit/aho-corasick-1975 computes on the scan path.
 * It shares no code with the product (aho-corasick-1975_b200/csrc) and none with the reference; every function
 * cites the reference lines whose behaviour it restates.  Only tests/, bench.py's cpu_baseline / --impl reference
 * legs and __graft_entry__.smoke() may load it.  It is pinned (tests/test_oracle.py) against
 *   - README.md:92-93 (the exact match list of examples/test.c),
 *   - the unmodified reference compiled in oracle/_ref/ in both builds (Meyer'85 and -DNMEYER_85), on the config-1
 *     dictionary/text (11,676 and 298,855 matches; FNV-1a-64 4cb7510699888d13 / e6ce99c887bfa45c) and on random
 *     dictionaries with heavy overlaps, including insertions interleaved with scanning on a carried cursor.
 *
 * Restated rules:
 *   goto/fail walk          aho_corasick.c:167-192   (state_goto; root self-loop simulated at :185-186)
 *   fail + nb_outputs       aho_corasick.c:194-208 and the BFS of :386-417 (classic build), lazy on a dirty counter (:353,:389)
 *   keyword identity/dedup  aho_corasick.c:340-363   (first termination gives the rank; re-insertion keeps it)
 *   enumeration order       aho_corasick.c:451-482   (terminal states along s, f(s), f(f(s)).. = longest first; length = depth)
 *   carried cursor          aho_corasick.c:360       (insertions never move a scan cursor)
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
  uint64_t end;
  uint32_t id;
  uint32_t len;
} acport_match;

#define NONE 0xFFFFFFFFu

typedef struct {
  size_t width;
  /* trie: states are numbered in creation order (aho_corasick.c:101), state 0 is the root */
  uint32_t nb_states, cap_states;
  uint32_t *parent, *depth, *rank; /* rank == NONE: not the end of a keyword (is_end_of_keyword, :54) */
  uint32_t *first_edge;            /* per state: head of its edge list */
  /* edges */
  uint32_t nb_edges, cap_edges;
  uint32_t *edge_sym, *edge_to, *edge_next;
  /* (state,symbol) -> edge hash, open addressing */
  uint64_t *hkey;
  uint32_t *hval;
  uint64_t hcap, hused;
  /* derived by the BFS (Algorithm 3) */
  uint32_t *fail, *nb_outputs;
  uint32_t *delta; /* optional dense table for 1-byte symbols: delta[s*256+a] */
  int dense_ok;
  uint64_t dirty; /* "reconstruct" counter, :70 */
  uint32_t nb_ranks;
  uint32_t cursor; /* carried scan cursor */
  uint32_t lmax;
} acport;

static uint64_t
mix64 (uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

static uint32_t
read_sym (const acport *h, const unsigned char *p) {
  /* symbol identity is memcmp over width bytes (cmp_default, :134-138): equal bytes <=> equal integers */
  switch (h->width) {
    case 1:
      return *p;
    case 2: {
      uint16_t v;
      memcpy (&v, p, 2);
      return v;
    }
    default: {
      uint32_t v;
      memcpy (&v, p, 4);
      return v;
    }
  }
}

static void
hash_grow (acport *h) {
  uint64_t ncap = h->hcap ? h->hcap * 2 : 1024;
  uint64_t *nk = malloc (ncap * sizeof (*nk));
  uint32_t *nv = malloc (ncap * sizeof (*nv));
  if (!nk || !nv)
    abort ();
  memset (nv, 0xFF, ncap * sizeof (*nv));
  for (uint64_t i = 0; i < h->hcap; i++)
    if (h->hval[i] != NONE) {
      uint64_t j = mix64 (h->hkey[i]) & (ncap - 1);
      while (nv[j] != NONE)
        j = (j + 1) & (ncap - 1);
      nk[j] = h->hkey[i];
      nv[j] = h->hval[i];
    }
  free (h->hkey);
  free (h->hval);
  h->hkey = nk;
  h->hval = nv;
  h->hcap = ncap;
}

/* g(s, a) on the trie only: the child or NONE ("fail"), :175 */
static inline uint32_t
child_of (const acport *h, uint32_t s, uint32_t a) {
  if (!h->hcap)
    return NONE;
  uint64_t key = ((uint64_t)s << 32) | a, j = mix64 (key) & (h->hcap - 1);
  while (h->hval[j] != NONE) {
    if (h->hkey[j] == key)
      return h->edge_to[h->hval[j]];
    j = (j + 1) & (h->hcap - 1);
  }
  return NONE;
}

static uint32_t
new_state (acport *h, uint32_t parent) {
  if (h->nb_states == h->cap_states) {
    h->cap_states = h->cap_states ? h->cap_states * 2 : 1024;
    h->parent = realloc (h->parent, h->cap_states * sizeof (uint32_t));
    h->depth = realloc (h->depth, h->cap_states * sizeof (uint32_t));
    h->rank = realloc (h->rank, h->cap_states * sizeof (uint32_t));
    h->first_edge = realloc (h->first_edge, h->cap_states * sizeof (uint32_t));
    if (!h->parent || !h->depth || !h->rank || !h->first_edge)
      abort ();
  }
  uint32_t s = h->nb_states++;
  h->parent[s] = parent;
  h->depth[s] = parent == NONE ? 0 : h->depth[parent] + 1;
  h->rank[s] = NONE;
  h->first_edge[s] = NONE;
  return s;
}

acport *
acport_create (size_t width) {
  if (width != 1 && width != 2 && width != 4)
    return 0;
  acport *h = calloc (1, sizeof (*h));
  if (!h)
    return 0;
  h->width = width;
  new_state (h, NONE);
  return h;
}

void
acport_release (acport *h) {
  if (!h)
    return;
  free (h->parent);
  free (h->depth);
  free (h->rank);
  free (h->first_edge);
  free (h->edge_sym);
  free (h->edge_to);
  free (h->edge_next);
  free (h->hkey);
  free (h->hval);
  free (h->fail);
  free (h->nb_outputs);
  free (h->delta);
  free (h);
}

int
acport_incremental (void) {
  return 0; /* restates the classic build: lazy full BFS */
}

/* Algorithm 2 (:291-316, :242-267) + end of keyword (:340-363). Returns the rank. */
uint32_t
acport_insert (acport *h, const void *symbols, size_t len) {
  const unsigned char *p = symbols;
  uint32_t s = 0;
  for (size_t i = 0; i < len; i++) {
    uint32_t a = read_sym (h, p + i * h->width), c = child_of (h, s, a);
    if (c == NONE) {
      if ((h->hused + 1) * 2 > h->hcap)
        hash_grow (h);
      if (h->nb_edges == h->cap_edges) {
        h->cap_edges = h->cap_edges ? h->cap_edges * 2 : 1024;
        h->edge_sym = realloc (h->edge_sym, h->cap_edges * sizeof (uint32_t));
        h->edge_to = realloc (h->edge_to, h->cap_edges * sizeof (uint32_t));
        h->edge_next = realloc (h->edge_next, h->cap_edges * sizeof (uint32_t));
        if (!h->edge_sym || !h->edge_to || !h->edge_next)
          abort ();
      }
      c = new_state (h, s);
      uint32_t e = h->nb_edges++;
      h->edge_sym[e] = a;
      h->edge_to[e] = c;
      h->edge_next[e] = h->first_edge[s];
      h->first_edge[s] = e;
      uint64_t key = ((uint64_t)s << 32) | a, j = mix64 (key) & (h->hcap - 1);
      while (h->hval[j] != NONE)
        j = (j + 1) & (h->hcap - 1);
      h->hkey[j] = key;
      h->hval[j] = e;
      h->hused++;
      h->dirty++; /* new states need f() before the next scan */
    }
    s = c;
  }
  if (s == 0)
    return NONE; /* empty keyword: the reference asserts (:345) */
  if (h->rank[s] == NONE) {
    h->rank[s] = h->nb_ranks++;
    h->dirty++;
    if (h->depth[s] > h->lmax)
      h->lmax = h->depth[s];
  }
  return h->rank[s];
}

void
acport_insert_many (acport *h, const void *symbols, const uint64_t *offsets, size_t nb, uint32_t *ranks) {
  for (size_t k = 0; k < nb; k++) {
    uint32_t r = acport_insert (h, (const unsigned char *)symbols + offsets[k] * h->width, (size_t)(offsets[k + 1] - offsets[k]));
    if (ranks)
      ranks[k] = r;
  }
}

size_t
acport_nb_keywords (const acport *h) {
  return h->nb_ranks;
}

void
acport_reset_cursor (acport *h) {
  h->cursor = 0;
}

uint32_t
acport_lmax (const acport *h) {
  return h->lmax;
}

uint32_t
acport_nb_states (const acport *h) {
  return h->nb_states;
}

/* delta(s,a) by Algorithm 1's loop (:172-191) over trie + fail */
static inline uint32_t
step_sparse (const acport *h, uint32_t s, uint32_t a) {
  for (;;) {
    uint32_t c = child_of (h, s, a);
    if (c != NONE)
      return c;
    if (s == 0)
      return 0; /* LOOP_0 simulated, :185-186 */
    s = h->fail[s];
  }
}

/* Algorithm 3 (:386-417): BFS from the root; f(child of root) = root (:205); f(T[r,a]) = delta(f(r), a) (:202);
 * nb_outputs(s) = [s terminal] + nb_outputs(f(s)) (:381,:207). */
static void
rebuild (acport *h) {
  if (!h->dirty)
    return;
  uint32_t n = h->nb_states;
  h->fail = realloc (h->fail, n * sizeof (uint32_t));
  h->nb_outputs = realloc (h->nb_outputs, n * sizeof (uint32_t));
  uint32_t *queue = malloc (n * sizeof (uint32_t));
  if (!h->fail || !h->nb_outputs || !queue)
    abort ();
  free (h->delta);
  h->delta = 0;
  h->dense_ok = h->width == 1 && (uint64_t)n * 256 * 4 <= (3ull << 30);
  if (h->dense_ok) {
    h->delta = malloc ((size_t)n * 256 * sizeof (uint32_t));
    if (!h->delta)
      h->dense_ok = 0;
  }
  size_t qh = 0, qt = 0;
  queue[qt++] = 0;
  h->fail[0] = 0;
  h->nb_outputs[0] = 0;
  while (qh < qt) {
    uint32_t r = queue[qh++];
    if (h->dense_ok) { /* row of r: inherit the row of f(r) (already final: f(r) is shallower), then overwrite trie edges */
      uint32_t *row = h->delta + (size_t)r * 256;
      if (r == 0)
        memset (row, 0, 256 * sizeof (uint32_t));
      else
        memcpy (row, h->delta + (size_t)h->fail[r] * 256, 256 * sizeof (uint32_t));
    }
    for (uint32_t e = h->first_edge[r]; e != NONE; e = h->edge_next[e]) {
      uint32_t s = h->edge_to[e], a = h->edge_sym[e];
      h->fail[s] = r == 0 ? 0 : step_sparse (h, h->fail[r], a);
      h->nb_outputs[s] = (h->rank[s] != NONE ? 1u : 0u) + h->nb_outputs[h->fail[s]];
      queue[qt++] = s;
    }
    if (h->dense_ok)
      for (uint32_t e = h->first_edge[r]; e != NONE; e = h->edge_next[e])
        h->delta[(size_t)r * 256 + h->edge_sym[e]] = h->edge_to[e];
  }
  free (queue);
  h->dirty = 0;
}

void
acport_prepare (acport *h) {
  rebuild (h);
}

/* The reference's usage loop (examples/test.c:17-23): acm_match per symbol (:434-448), then acm_get_match for
 * index 0..nb-1 (:451-482).  Symbols with index < lead are context only: matches ending there are not reported
 * (used by checkers that split a text into overlapped slices). */
uint64_t
acport_scan_lead (acport *h, const void *text, uint64_t n, uint64_t lead, uint64_t base, acport_match *out, uint64_t cap, int mode,
                  uint32_t *cursor) {
  rebuild (h);
  const unsigned char *t = text;
  uint32_t s = *cursor;
  uint64_t found = 0;
  for (uint64_t i = 0; i < n; i++) {
    if (h->dense_ok)
      s = h->delta[(size_t)s * 256 + t[i]];
    else
      s = step_sparse (h, s, read_sym (h, t + i * h->width));
    uint32_t nb = h->nb_outputs[s];
    if (!nb || i < lead)
      continue;
    if (!(mode & 1)) {
      found += nb;
      continue;
    }
    uint32_t q = s;
    for (uint32_t j = 0; j < nb; j++) { /* :459-466 */
      while (h->rank[q] == NONE)
        q = h->fail[q];
      if (found < cap && out)
        out[found] = (acport_match){ .end = base + i, .id = h->rank[q], .len = h->depth[q] /* :472-474 */ };
      found++;
      q = h->fail[q];
    }
  }
  *cursor = s;
  return found;
}

uint64_t
acport_scan (acport *h, const void *text, uint64_t n, uint64_t base, acport_match *out, uint64_t cap, int mode) {
  return acport_scan_lead (h, text, n, 0, base, out, cap, mode, &h->cursor);
}

struct mt_arg {
  acport *h;
  const unsigned char *text;
  uint64_t n;
  int mode;
  uint64_t found;
};

static void *
mt_worker (void *p) {
  struct mt_arg *a = p;
  uint32_t cur = 0;
  a->found = acport_scan_lead (a->h, a->text, a->n, 0, 0, 0, 0, a->mode, &cur);
  return 0;
}

/* Timing helper with the same meaning as refh_scan_mt: thread k scans slice k from state 0. */
uint64_t
acport_scan_mt (acport *h, const void *text, uint64_t n, int nthreads, int mode, double *seconds) {
  rebuild (h);
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
