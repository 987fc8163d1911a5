/* acm_finalise.c -- compiles the host trie (goto / fail / outputs) into the flat images the GPU scans.
 *
 * Role of the reference's Algorithm 3 pass (aho_corasick.c:386-417) plus everything the reference leaves in pointer form:
 *   symbol remap     device symbols are raw letters for ACM_CMP_DEFAULT with 1/2/4-byte letters, else stable class ids
 *                    assigned with the user's comparator (symbol identity is cmp()==0, aho_corasick.c:102,299);
 *   DFA engines      states renumbered by (has outputs, depth); delta[s][class] = child, else delta[f(s)][class]
 *                    (Appendix C rules 1-2), filled in depth order so f(s)'s row is final; CSR output sets list the
 *                    keyword states along s, f(s), ... longest first (rule 4) with length = depth (rule 5);
 *   filter engine    the last q symbols of every keyword feed a blocked Bloom filter (shared memory) and an exact q-gram
 *                    hash table; a reverse trie (keywords read right to left) in a hash table verifies candidates.
 * Engine choice and sizes are in DESIGN.md.
 */
#include "acm_internal.h"
#include "acm_tables.h"
#include <stdlib.h>
#include <string.h>

uint32_t
acm_symbol_of_state (const struct _ac_machine *m, const struct _ac_state *s) {
  switch (m->symbol_kind) {
    case ACM_SYM_RAW1:
      return *(const uint8_t *)s->letter;
    case ACM_SYM_RAW2: {
      uint16_t v;
      memcpy (&v, s->letter, 2);
      return v;
    }
    case ACM_SYM_RAW4: {
      uint32_t v;
      memcpy (&v, s->letter, 4);
      return v;
    }
    default:
      return m->class_of_state[s->id];
  }
}

/* ---- class ids for user comparators -------------------------------------------------------------------------------- */
static int64_t
class_slot (const struct _ac_machine *m, const void *letter) {
  uint32_t lo = 0, hi = m->nb_class;
  while (lo < hi) {
    uint32_t mid = lo + (hi - lo) / 2;
    int c = m->cmp (letter, m->class_letter[mid], m->cmp_arg);
    if (c == 0)
      return mid;
    if (c < 0)
      hi = mid;
    else
      lo = mid + 1;
  }
  return ~(int64_t)lo;
}

static int
extend_classes (struct _ac_machine *m) {
  if (m->cap_class_states < m->nb_states) {
    size_t cap = m->nb_states + m->nb_states / 2 + 16;
    uint32_t *grown = realloc (m->class_of_state, cap * sizeof (*grown));
    if (!grown)
      return ACM_B200_ERR_NOMEM;
    m->class_of_state = grown;
    m->cap_class_states = cap;
  }
  size_t id = 0;
  for (struct acm_state_block *b = m->blocks; b; b = b->next)
    for (uint32_t i = 0; i < b->used; i++, id++) {
      if (id < m->class_states_done)
        continue;
      struct _ac_state *s = &b->states[i];
      if (!s->parent) {
        m->class_of_state[id] = 0;
        continue;
      }
      int64_t k = class_slot (m, s->letter);
      if (k < 0) {
        if (m->nb_class == m->cap_class) {
          uint32_t cap = m->cap_class ? m->cap_class * 2 : 64;
          const void **gl = realloc (m->class_letter, cap * sizeof (*gl));
          uint32_t *gi = realloc (m->class_sorted_id, cap * sizeof (*gi));
          if (gl)
            m->class_letter = gl;
          if (gi)
            m->class_sorted_id = gi;
          if (!gl || !gi)
            return ACM_B200_ERR_NOMEM;
          m->cap_class = cap;
        }
        uint32_t at = (uint32_t)~k;
        memmove (m->class_letter + at + 1, m->class_letter + at, (m->nb_class - at) * sizeof (*m->class_letter));
        memmove (m->class_sorted_id + at + 1, m->class_sorted_id + at, (m->nb_class - at) * sizeof (*m->class_sorted_id));
        m->class_letter[at] = s->letter;
        m->class_sorted_id[at] = ++m->nb_class; /* ids start at 1 */
        k = at;
      }
      m->class_of_state[id] = m->class_sorted_id[k];
    }
  m->class_states_done = m->nb_states;
  return ACM_B200_OK;
}

int
acm_b200_remap_text (ACMachine *m, const void *letters, size_t letter_size, uint64_t nb, uint32_t *class_ids) {
  if (!m || (!letters && nb) || !class_ids || !letter_size)
    return ACM_B200_ERR_INVALID;
  if (m->symbol_kind == ACM_SYM_RAW1 || m->symbol_kind == ACM_SYM_RAW2 || m->symbol_kind == ACM_SYM_RAW4) {
    /* raw machines need no remap; give the identity so callers can be generic */
    for (uint64_t i = 0; i < nb; i++) {
      uint32_t v = 0;
      memcpy (&v, (const char *)letters + i * letter_size, letter_size < 4 ? letter_size : 4);
      class_ids[i] = v;
    }
    return ACM_B200_OK;
  }
  acm_lock (m);
  int rc = extend_classes (m);
  if (rc == ACM_B200_OK)
    for (uint64_t i = 0; i < nb; i++) {
      int64_t k = class_slot (m, (const char *)letters + i * letter_size);
      class_ids[i] = k >= 0 ? m->class_sorted_id[k] : 0;
    }
  acm_unlock (m);
  return rc;
}

/* ---- helpers -------------------------------------------------------------------------------------------------------- */
static uint64_t
pow2_at_least (uint64_t x) {
  uint64_t p = 1;
  while (p < x)
    p <<= 1;
  return p;
}

static void
slot_insert (acm_slot *tab, uint64_t nslots, uint64_t key, uint32_t node, uint32_t keyword) {
  uint64_t j = acm_mix64 (key) & (nslots - 1);
  while (tab[j].node != ACM_TAB_NONE)
    j = (j + 1) & (nslots - 1);
  tab[j] = (acm_slot){ key, node, keyword };
}

static acm_slot *
slot_find (acm_slot *tab, uint64_t nslots, uint64_t key) {
  uint64_t j = acm_mix64 (key) & (nslots - 1);
  while (tab[j].node != ACM_TAB_NONE) {
    if (tab[j].key == key)
      return &tab[j];
    j = (j + 1) & (nslots - 1);
  }
  return 0;
}

void
acm_free_tables (struct acm_tables *t) {
  free (t->delta);
  free (t->out_offsets);
  free (t->out_entries);
  free (t->dfa_of_state);
  free (t->bloom);
  free (t->bloom2);
  free (t->bloom_s2);
  free (t->s2_dist);
  free (t->kw_dist);
  free (t->qgrams);
  free (t->qset);
  free (t->kw_len);
  free (t->kw_off);
  free (t->kw_pool);
  free (t->kw_meta);
  free (t->kw_rpool);
  free (t->edges);
  if (t->builder) {
    free (t->builder->full);
    free (t->builder->parent);
    free (t->builder->depth);
    free (t->builder->count);
    free (t->builder->only_kw);
    free (t->builder->term_kw);
    free (t->builder);
  }
  memset (t, 0, sizeof (*t));
}

/* ---- DFA engines ---------------------------------------------------------------------------------------------------- */
static int
build_dfa (struct _ac_machine *m, struct acm_tables *t, struct _ac_state **by_depth) {
  const uint32_t n = (uint32_t)m->nb_states, K = t->nb_classes;
  /* renumber: states without outputs first (depth order, state 0 -> 0), then states with outputs */
  t->dfa_of_state = malloc ((size_t)n * sizeof (uint32_t));
  if (!t->dfa_of_state)
    return ACM_B200_ERR_NOMEM;
  uint32_t next = 0;
  for (uint32_t i = 0; i < n; i++)
    if (!by_depth[i]->nb_outputs)
      t->dfa_of_state[by_depth[i]->id] = next++;
  t->out_threshold = next;
  for (uint32_t i = 0; i < n; i++)
    if (by_depth[i]->nb_outputs)
      t->dfa_of_state[by_depth[i]->id] = next++;
  t->nb_dfa_states = n;

  t->delta_bytes = (size_t)n * K * (size_t)t->delta_entry_bytes;
  t->delta = malloc (t->delta_bytes ? t->delta_bytes : 1);
  if (!t->delta)
    return ACM_B200_ERR_NOMEM;
  uint16_t *d16 = t->delta_entry_bytes == 2 ? t->delta : 0;
  uint32_t *d32 = t->delta_entry_bytes == 4 ? t->delta : 0;
  for (uint32_t i = 0; i < n; i++) { /* depth order: the row of f(s) is final before the row of s */
    const struct _ac_state *s = by_depth[i];
    size_t row = (size_t)t->dfa_of_state[s->id] * K;
    if (!s->parent) {
      if (d16)
        memset (d16 + row, 0, K * 2);
      else
        memset (d32 + row, 0, (size_t)K * 4);
    } else {
      size_t frow = (size_t)t->dfa_of_state[s->fail->id] * K;
      if (d16)
        memcpy (d16 + row, d16 + frow, K * 2);
      else
        memcpy (d32 + row, d32 + frow, (size_t)K * 4);
    }
    for (uint32_t k = 0; k < s->nb_children; k++) {
      const struct _ac_state *c = s->children[k];
      uint32_t cls = t->class_of_byte[*(const uint8_t *)c->letter], to = t->dfa_of_state[c->id];
      if (d16)
        d16[row + cls] = (uint16_t)to;
      else
        d32[row + cls] = to;
    }
  }
  /* CSR output sets */
  uint32_t nout = n - t->out_threshold;
  t->out_offsets = malloc (((size_t)nout + 1) * sizeof (uint32_t));
  if (!t->out_offsets)
    return ACM_B200_ERR_NOMEM;
  struct _ac_state **of_dfa = malloc (((size_t)nout + 1) * sizeof (*of_dfa));
  if (!of_dfa)
    return ACM_B200_ERR_NOMEM;
  for (uint32_t i = 0; i < n; i++)
    if (by_depth[i]->nb_outputs)
      of_dfa[t->dfa_of_state[by_depth[i]->id] - t->out_threshold] = by_depth[i];
  uint64_t total = 0;
  for (uint32_t i = 0; i < nout; i++) {
    t->out_offsets[i] = (uint32_t)total;
    total += of_dfa[i]->nb_outputs;
    if (of_dfa[i]->nb_outputs > t->max_out_records)
      t->max_out_records = of_dfa[i]->nb_outputs > 0xFFFFFFFFu ? 0xFFFFFFFFu : (uint32_t)of_dfa[i]->nb_outputs;
  }
  if (total > 0xFFFFFFFFull) {
    free (of_dfa);
    return ACM_B200_ERR_NOMEM;
  }
  t->out_offsets[nout] = (uint32_t)total;
  t->nb_out_entries = total;
  t->out_entries = malloc ((total ? total : 1) * sizeof (acm_output));
  if (!t->out_entries) {
    free (of_dfa);
    return ACM_B200_ERR_NOMEM;
  }
  for (uint32_t i = 0; i < nout; i++) {
    acm_output *o = t->out_entries + t->out_offsets[i];
    size_t left = of_dfa[i]->nb_outputs;
    for (const struct _ac_state *s = of_dfa[i]; left; s = s->fail)
      if (s->rank != ACM_NONE) {
        *o++ = (acm_output){ s->rank, s->depth };
        left--;
      }
  }
  free (of_dfa);
  return ACM_B200_OK;
}

/* ---- stride-2 filter: window choice ------------------------------------------------------------------------------ */
/* Every keyword without a proper suffix that is a keyword ("root") chooses, for each parity, the distance d of the 3-byte window
 * it puts into the shared-memory filter (acm_tables.h); the others inherit the distances of their longest suffix keyword (the
 * same bytes, hence the same filter bits and the same distance-table entries).  Choice = the window that adds the fewest new
 * bits, weighted towards emptier filter words (the false-positive rate is the mean of the squared word fills), refined by
 * re-choosing every window a few times against a counting filter. */
struct s2_build {
  uint32_t words;
  uint32_t *cnt; /* [words][32] keys per bit */
  uint8_t *pop;  /* [words] bits set */
};

static inline void
s2_bits_of (uint32_t key, uint32_t words, uint32_t *word, uint32_t *b0, uint32_t *b1) {
  const uint32_t mask = acm_bloom_mask (key);
  *word = acm_bloom_word (key, words);
  *b0 = (uint32_t)__builtin_ctz (mask);
  *b1 = 31u - (uint32_t)__builtin_clz (mask); /* == b0 when both hash bits coincide */
}

static void
s2_apply (struct s2_build *b, uint32_t key, int add) {
  uint32_t w, b0, b1;
  s2_bits_of (key, b->words, &w, &b0, &b1);
  for (int i = 0; i < (b0 == b1 ? 1 : 2); i++) {
    uint32_t *c = &b->cnt[(size_t)w * 32 + (i ? b1 : b0)];
    if (add) {
      if ((*c)++ == 0)
        b->pop[w]++;
    } else if (--(*c) == 0)
      b->pop[w]--;
  }
}

static uint32_t
s2_cost (const struct s2_build *b, uint32_t key) {
  uint32_t w, b0, b1;
  s2_bits_of (key, b->words, &w, &b0, &b1);
  uint32_t fresh = b->cnt[(size_t)w * 32 + b0] == 0;
  if (b1 != b0)
    fresh += b->cnt[(size_t)w * 32 + b1] == 0;
  const uint32_t c = b->pop[w];
  return (c + fresh) * (c + fresh) - c * c;
}

/* window of keyword bytes k[0..len) whose last byte lies d bytes before the keyword's last byte */
static inline uint32_t
s2_window_key (const uint8_t *k, uint32_t len, uint32_t d) {
  const uint8_t *w = k + len - 3 - d;
  return acm_s2_key (w[0], w[1], w[2]);
}

static void
s2_dist_insert (uint32_t *tab, uint32_t lg, uint32_t gram3, uint32_t entry /* VALID | RIGHT? | d << 8 | ext */) {
  const uint32_t mask = (1u << lg) - 1u;
  for (uint32_t idx = acm_pair_word (gram3, lg);; idx = (idx + 1) & mask) {
    const uint32_t lo = tab[idx] & 0xFFFFu, hi = tab[idx] >> 16;
    if ((lo & ~ACM_S2D_CONT) == entry || hi == entry)
      return; /* shared with another keyword */
    if (!(lo & ACM_S2D_VALID)) {
      tab[idx] |= entry;
      return;
    }
    if (!(hi & ACM_S2D_VALID)) {
      tab[idx] |= entry << 16;
      return;
    }
    tab[idx] |= ACM_S2D_CONT;
  }
}

static int
build_stride2 (struct _ac_machine *m, struct acm_tables *t, uint64_t smem_optin) {
  const uint32_t nk = t->nb_keywords;
  const uint8_t *pool = t->kw_pool;
  int rc = ACM_B200_ERR_NOMEM;
  uint32_t *parent = malloc ((size_t)nk * sizeof (uint32_t)), *order = malloc ((size_t)nk * sizeof (uint32_t));
  uint32_t *start = calloc ((size_t)t->lmax + 2, sizeof (uint32_t));
  uint8_t *dist = malloc ((size_t)nk * 2);
  struct s2_build b = { 0, 0, 0 };
  uint32_t *bl = 0;
  if (!parent || !order || !start || !dist)
    goto done;
  /* longest proper suffix that is a keyword: the first terminal state on the fail chain (Appendix C rule 4) */
  for (uint32_t r = 0; r < nk; r++) {
    parent[r] = ACM_NONE;
    const struct _ac_state *s = m->keywords[r];
    if (s->nb_outputs > 1)
      for (const struct _ac_state *f = s->fail; f && f->parent; f = f->fail)
        if (f->rank != ACM_NONE) {
          parent[r] = f->rank;
          break;
        }
  }
  /* keywords by ascending length (a suffix keyword comes before the keywords that inherit from it) */
  for (uint32_t r = 0; r < nk; r++)
    start[t->kw_len[r] + 1]++;
  for (uint32_t l = 0; l <= t->lmax; l++)
    start[l + 1] += start[l];
  for (uint32_t r = 0; r < nk; r++)
    order[start[t->kw_len[r]]++] = r;
  uint64_t roots = 0;
  for (uint32_t r = 0; r < nk; r++)
    roots += parent[r] == ACM_NONE;

  /* The stage capacity depends on the hit rate, which depends on the filter size, which is what the stages leave of the shared
   * memory the kernel may use (smem_optin here): start from a small stage and grow it until the expected hits fit with a margin. */
  uint32_t hit_cap = 64;
  for (int attempt = 0; attempt < 22; attempt++, hit_cap += 32) {
    const uint64_t room = smem_optin - 2048;
    if (room < 32ull * ACM_S2_WARP_BYTES (hit_cap) + 4096)
      break;
    const uint64_t s2_max = (room - 32ull * ACM_S2_WARP_BYTES (hit_cap)) / 4;
    uint64_t s2_want = m->option_bloom_words ? m->option_bloom_words : pow2_at_least ((2ull * nk * 24 + 31) / 32);
    if (s2_want > s2_max)
      s2_want = s2_max;
    if (s2_want < 64)
      s2_want = 64;
    if (b.words != (uint32_t)s2_want) {
      free (b.cnt);
      free (b.pop);
      b.words = (uint32_t)s2_want;
      b.cnt = calloc ((size_t)b.words * 32, sizeof (uint32_t));
      b.pop = calloc (b.words, 1);
      if (!b.cnt || !b.pop)
        goto done;
      memset (dist, 0xFF, (size_t)nk * 2);
      for (int pass = 0; pass < (nk > 300000 ? 2 : 4); pass++) /* (huge dictionaries: two rounds of re-choosing are enough) */
        for (uint32_t i = 0; i < nk; i++) {
          const uint32_t r = order[i], len = t->kw_len[r];
          if (parent[r] != ACM_NONE)
            continue;
          const uint8_t *k = pool + t->kw_off[r];
          const uint32_t dmax = len - 3 < ACM_S2_DMAX ? len - 3 : ACM_S2_DMAX;
          for (uint32_t role = 0; role < 2; role++) {
            uint8_t *chosen = &dist[2 * (size_t)r + role];
            if (*chosen != 0xFF)
              s2_apply (&b, s2_window_key (k, len, *chosen), 0);
            /* a window at the very start of the keyword has no keyword byte before it: the distance table extends it to the
             * RIGHT instead -- the rare case for the kernel; such a window is chosen only when the keyword has no other of this
             * parity (4-byte keywords, odd distance) */
            uint32_t best_d = role, best = 0xFFFFFFFFu;
            const uint32_t dlast = role + 4 > len ? role : (len - 4 < dmax ? len - 4 : dmax);
            for (uint32_t d = role; d <= dlast; d += 2) {
              const uint32_t c = s2_cost (&b, s2_window_key (k, len, d)) * 64 + d; /* ties: the window nearest to the end */
              if (c < best)
                best = c, best_d = d;
            }
            *chosen = (uint8_t)best_d;
            s2_apply (&b, s2_window_key (k, len, best_d), 1);
          }
        }
    }
    double fp2 = 0;
    for (uint32_t i = 0; i < b.words; i++) {
      const double f = b.pop[i] / 32.0;
      fp2 += f * f;
    }
    const double hit_rate = fp2 / b.words + 2.0 * (double)roots / 16777216.0;
    if (hit_rate * 1024 * 1.5 + 32 > hit_cap) { /* expected hits per tile (1024 tests) must leave a 1.5x margin in the stage */
      /* jump to the stage this hit rate asks for (+ one step: the smaller filter will let a little more through) instead of
       * growing by 32 and redoing the window choice each time -- with 10^6 keywords one choice takes seconds */
      const uint32_t want_cap = ((uint32_t)(hit_rate * 1024 * 1.5 + 32) + 31) / 32 * 32 + 32;
      if (want_cap > hit_cap + 32)
        hit_cap = want_cap - 32; /* the loop adds 32 */
      continue;
    }
    for (uint32_t i = 0; i < nk; i++) { /* ascending length: the suffix keyword's distances are final */
      const uint32_t r = order[i];
      if (parent[r] != ACM_NONE) {
        dist[2 * (size_t)r] = dist[2 * (size_t)parent[r]];
        dist[2 * (size_t)r + 1] = dist[2 * (size_t)parent[r] + 1];
      }
    }
    bl = calloc (b.words, sizeof (uint32_t));
    /* distance table: 32 words per keyword while that stays at 16 MB (config 3), never more than that unless the table would be more
     * than half full: it must stay L2-RESIDENT next to the streaming text -- at 64 MB (10^6 keywords, first version) every lookup
     * was a DRAM access, 11.7 GB read for a 2 GiB text (profiles/r2_filter_scan_s2_c4s_2gib.txt) */
    uint32_t lg = 16;
    const uint32_t lg_cap = m->option_s2_dist_log2 ? (uint32_t)m->option_s2_dist_log2 : 23;
    while (lg < 26 && ((lg < lg_cap && (1ull << lg) < 32ull * nk) || (1ull << lg) < 2ull * nk))
      lg++;
    t->s2_dist = calloc ((size_t)1 << lg, sizeof (uint32_t));
    t->kw_dist = malloc (((size_t)nk + 1) * sizeof (uint16_t));
    if (!bl || !t->s2_dist || !t->kw_dist)
      goto done;
    for (uint32_t w = 0; w < b.words; w++) {
      for (uint32_t bit = 0; bit < 32; bit++)
        if (b.cnt[(size_t)w * 32 + bit])
          bl[w] |= 1u << bit;
    }
    t->s2_dist_log2 = lg;
    for (uint32_t r = 0; r < nk; r++) {
      t->kw_dist[r] = (uint16_t)(dist[2 * (size_t)r] | (dist[2 * (size_t)r + 1] << 8));
      if (parent[r] != ACM_NONE)
        continue;
      const uint32_t len = t->kw_len[r];
      const uint8_t *k = pool + t->kw_off[r];
      for (uint32_t role = 0; role < 2; role++) {
        const uint32_t d = dist[2 * (size_t)r + role];
        const uint8_t *w = k + len - 3 - d; /* the window; extended to the left when a keyword byte precedes it, else to the right */
        const uint32_t gram3 = w[0] | ((uint32_t)w[1] << 8) | ((uint32_t)w[2] << 16);
        const uint32_t entry = d + 4 <= len ? (ACM_S2D_VALID | (d << 8) | w[-1]) : (ACM_S2D_VALID | ACM_S2D_RIGHT | (d << 8) | w[3]);
        s2_dist_insert (t->s2_dist, lg, gram3, entry);
      }
    }
    t->bloom_s2 = bl;
    bl = 0;
    t->bloom_s2_words = b.words;
    t->bloom_s2_hit_rate = hit_rate;
    t->s2_hit_cap = hit_cap;
    break;
  }
  rc = ACM_B200_OK;
done:
  free (parent);
  free (order);
  free (start);
  free (dist);
  free (b.cnt);
  free (b.pop);
  free (bl);
  return rc;
}

/* ---- filter engine -------------------------------------------------------------------------------------------------- */
/* Reverse trie (keywords read right to left), path-compressed: an edge whose subtree holds exactly ONE keyword does not lead to a
 * node but to a "tail" (ACM_TAIL_FLAG | keyword id): the rest of that keyword is compared directly against the text from the
 * keyword pool.  Only edges leaving nodes with two or more keywords below them are stored. */
static int
build_filter (struct _ac_machine *m, struct acm_tables *t, uint64_t smem_budget, uint64_t smem_optin) {
  const uint32_t nk = (uint32_t)m->nb_sequences;
  uint64_t total_syms = 0;
  for (uint32_t r = 0; r < nk; r++)
    total_syms += m->keywords[r]->depth;
  t->q = m->lmin == ACM_NONE ? 1 : m->lmin;
  uint32_t qmax = t->width == 1 ? 4 : 2;
  if (t->q > qmax)
    t->q = qmax;
  const uint32_t q = t->q;
  const int shift = t->width == 1 ? 8 : (t->width == 2 ? 16 : 32);
  int rc = ACM_B200_ERR_NOMEM;

  /* keyword pool (forward symbols) for the tail compares */
  /* dictionaries of moderate size keep their build structures and get room to grow: keywords appended later are then added in
   * place (acm_patch_filter_tables).  The byte-alphabet path with its reversed pool is rebuilt instead (its stride-2 tables are
   * optimised as a whole). */
  const int keep = !m->option_no_patch && total_syms <= (8u << 20) && t->width != 1;
  const uint64_t kw_cap = keep ? (uint64_t)nk + nk / 4 + 8192 : (uint64_t)nk + 1;
  const uint64_t pool_cap_syms = keep ? total_syms + total_syms / 4 + 65536 : total_syms + 4;
  t->kw_len = malloc ((size_t)kw_cap * sizeof (uint32_t));
  t->kw_off = malloc ((size_t)kw_cap * sizeof (uint64_t));
  t->kw_pool_bytes = pool_cap_syms * (uint64_t)t->width;
  t->kw_pool = calloc (1, t->kw_pool_bytes);
  /* full (uncompressed) reverse trie, host only */
  const uint64_t full_slots = pow2_at_least (2 * pool_cap_syms + 16);
  acm_slot *full = malloc (full_slots * sizeof (acm_slot));
  const size_t max_nodes = (size_t)pool_cap_syms + 2;
  uint32_t *parent = malloc (max_nodes * sizeof (uint32_t)), *depth = malloc (max_nodes * sizeof (uint32_t));
  uint32_t *count = calloc (max_nodes, sizeof (uint32_t)), *only_kw = malloc (max_nodes * sizeof (uint32_t));
  uint32_t *term_kw = malloc (max_nodes * sizeof (uint32_t));
  uint64_t cap_q = 1024, nq = 0;
  uint64_t *qkeys = malloc (cap_q * sizeof (*qkeys));
  uint32_t *qnodes = malloc (cap_q * sizeof (*qnodes));
  if (!t->kw_len || !t->kw_off || !t->kw_pool || !full || !parent || !depth || !count || !only_kw || !term_kw || !qkeys || !qnodes)
    goto done;
  memset (full, 0xFF, full_slots * sizeof (acm_slot));
  memset (term_kw, 0xFF, max_nodes * sizeof (uint32_t));
  uint32_t nodes = 1;
  parent[0] = 0;
  depth[0] = 0;
  uint64_t pool_at = 0;
  for (uint32_t r = 0; r < nk; r++) {
    const uint32_t len = m->keywords[r]->depth;
    t->kw_len[r] = len;
    t->kw_off[r] = pool_at;
    uint32_t node = 0, d = 0;
    uint64_t key = 0;
    for (const struct _ac_state *s = m->keywords[r]; s->parent; s = s->parent) { /* last letter first */
      const uint32_t sym = acm_symbol_of_state (m, s);
      const uint64_t at = pool_at + (len - 1 - d);
      if (t->width == 1)
        ((uint8_t *)t->kw_pool)[at] = (uint8_t)sym;
      else if (t->width == 2)
        ((uint16_t *)t->kw_pool)[at] = (uint16_t)sym;
      else
        ((uint32_t *)t->kw_pool)[at] = sym;
      const uint64_t ekey = ((uint64_t)node << 32) | sym;
      acm_slot *e = slot_find (full, full_slots, ekey);
      int created = 0;
      if (!e) {
        parent[nodes] = node;
        depth[nodes] = d + 1;
        slot_insert (full, full_slots, ekey, nodes++, ACM_TAB_NONE);
        e = slot_find (full, full_slots, ekey);
        created = 1;
      }
      node = e->node;
      d++;
      if (d <= q) {
        key = shift == 32 ? ((d == 1 ? 0 : key << 32) | sym) : ((key << shift) | sym);
        if (d == q && created) {
          if (nq == cap_q) {
            cap_q *= 2;
            qkeys = realloc (qkeys, cap_q * sizeof (*qkeys));
            qnodes = realloc (qnodes, cap_q * sizeof (*qnodes));
            if (!qkeys || !qnodes)
              goto done;
          }
          qkeys[nq] = key;
          qnodes[nq++] = node;
        }
      }
    }
    term_kw[node] = r; /* distinct keywords end at distinct nodes */
    pool_at += len;
  }
  t->nb_rev_nodes = nodes;
  if (t->width == 1 && total_syms + 3ull * nk < 0xFFFFFFF0ull * 4) {
    /* reversed, word-aligned copy of the pool: the tail compare of the verification kernels walks the text leftwards four bytes
     * at a time against aligned 32-bit words */
    uint64_t words = 0;
    for (uint32_t r = 0; r < nk; r++)
      words += (t->kw_len[r] + 3) / 4;
    t->kw_meta = malloc (((size_t)nk + 1) * 2 * sizeof (uint32_t));
    t->kw_rpool = calloc (words + 1, sizeof (uint32_t));
    if (!t->kw_meta || !t->kw_rpool)
      goto done;
    t->kw_rpool_words = words + 1;
    uint64_t at = 0;
    for (uint32_t r = 0; r < nk; r++) {
      const uint32_t len = t->kw_len[r];
      const uint8_t *fwd = (const uint8_t *)t->kw_pool + t->kw_off[r];
      t->kw_meta[2 * (size_t)r] = len;
      t->kw_meta[2 * (size_t)r + 1] = (uint32_t)at;
      for (uint32_t j = 0; j < len; j++)
        t->kw_rpool[at + (j >> 2)] |= (uint32_t)fwd[len - 1 - j] << (8 * (j & 3));
      at += (len + 3) / 4;
    }
  }
  /* keywords per subtree; children have larger ids than their parent */
  for (uint32_t v = nodes - 1; v >= 1; v--) {
    if (term_kw[v] != ACM_TAB_NONE) {
      count[v]++;
      only_kw[v] = term_kw[v];
    }
    count[parent[v]] += count[v];
    only_kw[parent[v]] = only_kw[v]; /* meaningful only where the final count is 1 */
  }
  /* stored edges: those leaving a node at depth >= q that has >= 2 keywords below it */
  uint64_t kept = 0;
  for (uint64_t i = 0; i < full_slots; i++)
    if (full[i].node != ACM_TAB_NONE) {
      const uint32_t from = (uint32_t)(full[i].key >> 32);
      if (depth[from] >= q && count[from] >= 2)
        kept++;
    }
  t->edge_slots = pow2_at_least (3 * kept + 16);
  t->edges = malloc (t->edge_slots * sizeof (acm_slot));
  if (!t->edges)
    goto done;
  memset (t->edges, 0xFF, t->edge_slots * sizeof (acm_slot));
  for (uint64_t i = 0; i < full_slots; i++)
    if (full[i].node != ACM_TAB_NONE) {
      const uint32_t from = (uint32_t)(full[i].key >> 32), to = full[i].node;
      if (depth[from] >= q && count[from] >= 2) {
        if (count[to] == 1)
          slot_insert (t->edges, t->edge_slots, full[i].key, ACM_TAIL_FLAG | only_kw[to], ACM_TAB_NONE);
        else
          slot_insert (t->edges, t->edge_slots, full[i].key, to, term_kw[to]);
      }
    }
  /* exact q-gram table: depth-q node (or tail), and the keyword of exactly q symbols ending there */
  t->qgram_slots = pow2_at_least (3 * nq + 16);
  t->qgrams = malloc (t->qgram_slots * sizeof (acm_slot));
  if (!t->qgrams)
    goto done;
  memset (t->qgrams, 0xFF, t->qgram_slots * sizeof (acm_slot));
  for (uint64_t i = 0; i < nq; i++) {
    const uint32_t v = qnodes[i];
    if (count[v] == 1)
      slot_insert (t->qgrams, t->qgram_slots, qkeys[i], ACM_TAIL_FLAG | only_kw[v], ACM_TAB_NONE);
    else
      slot_insert (t->qgrams, t->qgram_slots, qkeys[i], v, term_kw[v]);
  }
  if (t->width != 4) { /* compact confirmation set, load factor <= 1/4 */
    uint32_t bits = 8;
    while ((4ull << bits) < 4 * nq + 16 && bits < 29)
      bits++;
    t->qset_shift = 32 - bits;
    t->qset = malloc ((size_t)16 << bits);
    if (!t->qset)
      goto done;
    memset (t->qset, 0xFF, (size_t)16 << bits);
    const uint32_t mask = (1u << bits) - 1u;
    for (uint64_t i = 0; i < nq; i++) {
      const uint32_t key = (uint32_t)qkeys[i];
      if (key == ACM_QSET_EMPTY) {
        t->qset_has_empty_key = 1;
        continue;
      }
      for (uint32_t b = acm_qset_bucket (key, t->qset_shift), placed = 0; !placed; b = (b + 1) & mask)
        for (int c = 0; c < 4 && !placed; c++)
          if (t->qset[4 * (size_t)b + c] == ACM_QSET_EMPTY) {
            t->qset[4 * (size_t)b + c] = key;
            placed = 1;
          }
    }
  }
  /* blocked Bloom filter sized to the shared-memory budget: ~24 bits per q-gram, at most the budget */
  uint64_t want_words = m->option_bloom_words ? m->option_bloom_words : pow2_at_least ((nq * 24 + 31) / 32);
  uint64_t max_words = smem_budget / 4;
  if (want_words > max_words)
    want_words = max_words;
  if (want_words < 64)
    want_words = 64;
  t->bloom_words = (uint32_t)want_words;
  t->bloom = calloc (t->bloom_words, sizeof (uint32_t));
  if (!t->bloom)
    goto done;
  for (uint64_t i = 0; i < nq; i++) {
    uint32_t f = acm_fold_key (qkeys[i]);
    t->bloom[acm_bloom_word (f, t->bloom_words)] |= acm_bloom_mask (f);
  }
  double fp = 0;
  for (uint32_t i = 0; i < t->bloom_words; i++) {
    const double f = __builtin_popcount (t->bloom[i]) / 32.0;
    fp += f * f;
  }
  t->bloom_fp = fp / t->bloom_words;
  if (t->bloom_fp > 0.08) { /* the shared-memory level alone would flood the confirmation step: add the global level */
    t->bloom2_words = (uint32_t)pow2_at_least (nq < 1024 ? 1024 : nq); /* >= 32 bits per key */
    t->bloom2 = calloc (t->bloom2_words, sizeof (uint32_t));
    if (!t->bloom2)
      goto done;
    for (uint64_t i = 0; i < nq; i++) {
      const uint32_t f = acm_fold_key (qkeys[i]);
      t->bloom2[acm_bloom2_word (f, t->bloom2_words)] |= acm_bloom2_mask (f);
    }
    double fp2 = 0;
    for (uint32_t i = 0; i < t->bloom2_words; i++) {
      const double f = __builtin_popcount (t->bloom2[i]) / 32.0;
      fp2 += f * f;
    }
    t->bloom_fp *= fp2 / t->bloom2_words;
  }
  /* stride-2 filter (acm_tables.h): byte alphabet, every keyword at least 4 bytes long, dictionary small enough for the
   * shared-memory level to stay selective with two keys per keyword */
  if (t->width == 1 && q == 4 && nk && smem_optin >= 65536) {
    const int s2rc = build_stride2 (m, t, smem_optin);
    if (s2rc != ACM_B200_OK) {
      rc = s2rc;
      goto done;
    }
  }
  rc = ACM_B200_OK;
  if (keep && (t->builder = calloc (1, sizeof (*t->builder)))) { /* the build structures stay: appended keywords are added in place */
    struct acm_filter_builder *b = t->builder;
    b->full = full, b->full_slots = full_slots, b->full_used = nodes - 1;
    b->parent = parent, b->depth = depth, b->count = count, b->only_kw = only_kw, b->term_kw = term_kw;
    b->max_nodes = max_nodes;
    b->nodes = nodes;
    b->nq = nq;
    b->pool_syms = pool_at;
    b->pool_cap_syms = pool_cap_syms;
    b->kw_cap = kw_cap;
    b->edges_used = kept;
    b->keywords_done = nk;
    full = 0, parent = depth = count = only_kw = term_kw = 0;
  }
done:
  free (full);
  free (parent);
  free (depth);
  free (count);
  free (only_kw);
  free (term_kw);
  free (qkeys);
  free (qnodes);
  return rc;
}

/* ---- append-only update of the filter tables --------------------------------------------------------- */
/* The reference's Meyer bookkeeping (aho_corasick.c:210-240, 318-338) says which fail links and output counts a new keyword
 * touches; the filter engine has neither -- its tables are keyed by the REVERSE trie -- so the set that changes is smaller still:
 * the nodes on the new keyword's own right-to-left path.  For each of them the stored form is recomputed from the per-node
 * keyword counts kept by the builder: an edge is stored iff it leaves a node at depth >= q with two or more keywords below it,
 * and leads to a tail (one keyword below) or to a node; the q-gram table holds the same for the depth-q node.  A node whose
 * count goes from 1 to 2 additionally gets the edge of the keyword that used to be alone below it. */
static int
patch_push (struct acm_patch_list *pl, uint32_t array, uint64_t index, const void *value, size_t bytes) {
  if (pl->nb == pl->cap) {
    const uint64_t cap = pl->cap ? pl->cap * 2 : 4096;
    struct acm_patch *grown = realloc (pl->items, cap * sizeof (*grown));
    if (!grown)
      return ACM_B200_ERR_NOMEM;
    pl->items = grown;
    pl->cap = cap;
  }
  struct acm_patch *p = &pl->items[pl->nb++];
  memset (p, 0, sizeof (*p));
  p->array = array;
  p->index = index;
  memcpy (p->value, value, bytes);
  return ACM_B200_OK;
}

/* sets key -> (node, keyword) in an open-addressing table, recording the slot if it changed */
static int
slot_upsert (acm_slot *tab, uint64_t nslots, uint64_t *used, uint64_t key, uint32_t node, uint32_t keyword, uint32_t array, struct acm_patch_list *pl) {
  uint64_t j = acm_mix64 (key) & (nslots - 1);
  while (tab[j].node != ACM_TAB_NONE && tab[j].key != key)
    j = (j + 1) & (nslots - 1);
  if (tab[j].node == ACM_TAB_NONE)
    (*used)++;
  else if (tab[j].node == node && tab[j].keyword == keyword)
    return ACM_B200_OK;
  tab[j] = (acm_slot){ key, node, keyword };
  return patch_push (pl, array, j, &tab[j], sizeof (acm_slot));
}

static uint32_t
pool_symbol (const struct acm_tables *t, uint64_t at) {
  return t->width == 1 ? ((const uint8_t *)t->kw_pool)[at] : (t->width == 2 ? ((const uint16_t *)t->kw_pool)[at] : ((const uint32_t *)t->kw_pool)[at]);
}

int
acm_patch_filter_tables (struct _ac_machine *m, struct acm_tables *t, struct acm_patch_list *pl) {
  struct acm_filter_builder *b = t->builder;
  const uint32_t nk = (uint32_t)m->nb_sequences, q = t->q;
  memset (pl, 0, sizeof (*pl));
  if (!b || t->engine != ACM_B200_ENGINE_FILTER || t->bloom_s2 || t->kw_meta || nk < b->keywords_done || m->engine_override[0])
    return ACM_B200_ERR_INVALID;
  if (m->symbol_kind != ACM_SYM_RAW1 && m->symbol_kind != ACM_SYM_RAW2 && m->symbol_kind != ACM_SYM_RAW4)
    return ACM_B200_ERR_INVALID; /* class-id alphabets: new letters may add classes */
  const int shift = t->width == 1 ? 8 : (t->width == 2 ? 16 : 32);
  /* everything must fit the room the build left */
  uint64_t new_syms = 0;
  for (uint32_t r = b->keywords_done; r < nk; r++) {
    if (m->keywords[r]->depth < q)
      return ACM_B200_ERR_INVALID; /* a shorter filter window: every table changes */
    new_syms += m->keywords[r]->depth;
  }
  if (nk + 1 > b->kw_cap || b->pool_syms + new_syms + 4 > b->pool_cap_syms || b->nodes + new_syms + 2 > b->max_nodes || (b->full_used + new_syms) * 10 > b->full_slots * 7
      || (b->edges_used + 3ull * (nk - b->keywords_done)) * 10 > t->edge_slots * 7 || (b->nq + (nk - b->keywords_done)) * 10 > t->qgram_slots * 7
      || (t->qset && (b->nq + (nk - b->keywords_done)) * 2 > ((uint64_t)4 << (32 - t->qset_shift))))
    return ACM_B200_ERR_INVALID;
  pl->kw_first = b->keywords_done;
  pl->pool_first_sym = b->pool_syms;
  uint32_t *path = malloc (((size_t)m->lmax + 2) * sizeof (uint32_t)), *path_sym = malloc (((size_t)m->lmax + 2) * sizeof (uint32_t));
  uint32_t *old_count = malloc (((size_t)m->lmax + 2) * sizeof (uint32_t)), *old_only = malloc (((size_t)m->lmax + 2) * sizeof (uint32_t));
  int rc = ACM_B200_ERR_NOMEM;
  if (!path || !path_sym || !old_count || !old_only)
    goto done;
  for (uint32_t r = b->keywords_done; r < nk; r++) {
    const uint32_t len = m->keywords[r]->depth;
    t->kw_len[r] = len;
    t->kw_off[r] = b->pool_syms;
    /* 1. the keyword's right-to-left path in the uncompressed trie (new nodes appended) and its symbols into the pool */
    uint32_t node = 0, d = 0;
    uint64_t qkey = 0;
    int q_created = 0;
    path[0] = 0;
    for (const struct _ac_state *s = m->keywords[r]; s->parent; s = s->parent) {
      const uint32_t sym = acm_symbol_of_state (m, s);
      const uint64_t at = b->pool_syms + (len - 1 - d);
      if (t->width == 1)
        ((uint8_t *)t->kw_pool)[at] = (uint8_t)sym;
      else if (t->width == 2)
        ((uint16_t *)t->kw_pool)[at] = (uint16_t)sym;
      else
        ((uint32_t *)t->kw_pool)[at] = sym;
      const uint64_t ekey = ((uint64_t)node << 32) | sym;
      acm_slot *e = slot_find (b->full, b->full_slots, ekey);
      int created = 0;
      if (!e) {
        b->parent[b->nodes] = node;
        b->depth[b->nodes] = d + 1;
        b->count[b->nodes] = 0;
        b->term_kw[b->nodes] = ACM_TAB_NONE;
        slot_insert (b->full, b->full_slots, ekey, b->nodes++, ACM_TAB_NONE);
        b->full_used++;
        e = slot_find (b->full, b->full_slots, ekey);
        created = 1;
      }
      node = e->node;
      d++;
      path[d] = node;
      path_sym[d] = sym;
      if (d <= q) {
        qkey = shift == 32 ? ((d == 1 ? 0 : qkey << 32) | sym) : ((qkey << shift) | sym);
        if (d == q)
          q_created = created;
      }
    }
    b->pool_syms += len;
    /* 2. counts along the path (the old ones are needed below) */
    for (uint32_t i = 1; i <= len; i++) {
      old_count[i] = b->count[path[i]];
      old_only[i] = b->only_kw[path[i]];
      b->count[path[i]]++;
      if (b->count[path[i]] == 1)
        b->only_kw[path[i]] = r;
    }
    b->term_kw[path[len]] = r; /* distinct keywords end at distinct nodes */
    /* 3. the stored form of every node of the path from depth q on */
    for (uint32_t i = q; i <= len; i++) {
      const uint32_t v = path[i];
      const uint32_t to_node = b->count[v] == 1 ? (ACM_TAIL_FLAG | b->only_kw[v]) : v, to_kw = b->count[v] == 1 ? ACM_TAB_NONE : b->term_kw[v];
      if (i == q) {
        uint64_t used = b->nq;
        if ((rc = slot_upsert (t->qgrams, t->qgram_slots, &used, qkey, to_node, to_kw, ACM_PATCH_QGRAMS, pl)))
          goto done;
        if (q_created) { /* a q-gram no keyword ended with so far: filter bits, confirmation set */
          b->nq++;
          const uint32_t f = acm_fold_key (qkey);
          const uint32_t w1 = acm_bloom_word (f, t->bloom_words), before1 = t->bloom[w1];
          t->bloom[w1] |= acm_bloom_mask (f);
          if (t->bloom[w1] != before1 && (rc = patch_push (pl, ACM_PATCH_BLOOM, w1, &t->bloom[w1], 4)))
            goto done;
          if (t->bloom2) {
            const uint32_t w2 = acm_bloom2_word (f, t->bloom2_words), before2 = t->bloom2[w2];
            t->bloom2[w2] |= acm_bloom2_mask (f);
            if (t->bloom2[w2] != before2 && (rc = patch_push (pl, ACM_PATCH_BLOOM2, w2, &t->bloom2[w2], 4)))
              goto done;
          }
          if (t->qset) {
            const uint32_t key = (uint32_t)qkey, mask = (1u << (32 - t->qset_shift)) - 1u;
            if (key == ACM_QSET_EMPTY)
              t->qset_has_empty_key = 1;
            else
              for (uint32_t bk = acm_qset_bucket (key, t->qset_shift), placed = 0; !placed; bk = (bk + 1) & mask)
                for (int c = 0; c < 4 && !placed; c++)
                  if (t->qset[4 * (size_t)bk + c] == ACM_QSET_EMPTY) {
                    t->qset[4 * (size_t)bk + c] = key;
                    placed = 1;
                    if ((rc = patch_push (pl, ACM_PATCH_QSET, 4 * (uint64_t)bk + c, &key, 4)))
                      goto done;
                  }
          }
        }
      } else if (b->count[path[i - 1]] >= 2) {
        if ((rc = slot_upsert (t->edges, t->edge_slots, &b->edges_used, ((uint64_t)path[i - 1] << 32) | path_sym[i], to_node, to_kw, ACM_PATCH_EDGES, pl)))
          goto done;
      }
      /* a node that had ONE keyword below it and now has two: that keyword's own edge out of it was not stored so far */
      if (old_count[i] == 1 && i < len + 1) {
        const uint32_t old = old_only[i], old_len = t->kw_len[old];
        if (old_len > i) {
          const uint32_t sym = pool_symbol (t, t->kw_off[old] + (old_len - 1 - i));
          const acm_slot *e = slot_find (b->full, b->full_slots, ((uint64_t)v << 32) | sym);
          const uint32_t c = e->node;
          const uint32_t c_node = b->count[c] == 1 ? (ACM_TAIL_FLAG | b->only_kw[c]) : c, c_kw = b->count[c] == 1 ? ACM_TAB_NONE : b->term_kw[c];
          if ((rc = slot_upsert (t->edges, t->edge_slots, &b->edges_used, ((uint64_t)v << 32) | sym, c_node, c_kw, ACM_PATCH_EDGES, pl)))
            goto done;
        }
      }
    }
  }
  /* a word or slot may have changed more than once (two new keywords sharing an edge): the kernel applies the list in parallel,
   * so every entry carries the FINAL content of its target */
  for (uint64_t i = 0; i < pl->nb; i++) {
    struct acm_patch *p = &pl->items[i];
    switch (p->array) {
      case ACM_PATCH_BLOOM:
        p->value[0] = t->bloom[p->index];
        break;
      case ACM_PATCH_BLOOM2:
        p->value[0] = t->bloom2[p->index];
        break;
      case ACM_PATCH_QSET:
        p->value[0] = t->qset[p->index];
        break;
      case ACM_PATCH_QGRAMS:
        memcpy (p->value, &t->qgrams[p->index], sizeof (acm_slot));
        break;
      case ACM_PATCH_EDGES:
        memcpy (p->value, &t->edges[p->index], sizeof (acm_slot));
        break;
    }
  }
  pl->kw_nb = nk - pl->kw_first;
  pl->pool_nb_syms = b->pool_syms - pl->pool_first_sym;
  b->keywords_done = nk;
  t->nb_keywords = nk;
  t->nb_states = (uint32_t)m->nb_states;
  t->lmax = m->lmax;
  t->nb_rev_nodes = b->nodes;
  rc = ACM_B200_OK;
done:
  free (path);
  free (path_sym);
  free (old_count);
  free (old_only);
  if (rc) {
    free (pl->items);
    memset (pl, 0, sizeof (*pl));
  }
  return rc;
}

/* ---- entry ---------------------------------------------------------------------------------------------------------- */
int
acm_build_tables (struct _ac_machine *m, struct acm_tables *t, uint64_t smem_budget, uint64_t smem_optin) {
  memset (t, 0, sizeof (*t));
  const uint32_t n = (uint32_t)m->nb_states;
  t->nb_states = n;
  t->nb_keywords = (uint32_t)m->nb_sequences;
  t->lmax = m->lmax;
  t->lmin = m->lmin == ACM_NONE ? 0 : m->lmin;
  const int raw = m->symbol_kind == ACM_SYM_RAW1 || m->symbol_kind == ACM_SYM_RAW2 || m->symbol_kind == ACM_SYM_RAW4;
  t->width = raw ? (int)m->symbol_size : 4;
  if (!raw) {
    int rc = extend_classes (m);
    if (rc)
      return rc;
  }
  /* byte classes (DFA engines) */
  uint32_t K = 1;
  if (t->width == 1) {
    int used[256] = { 0 };
    for (struct acm_state_block *b = m->blocks; b; b = b->next)
      for (uint32_t i = 0; i < b->used; i++)
        if (b->states[i].parent)
          used[*(const uint8_t *)b->states[i].letter] = 1;
    for (int c = 0; c < 256; c++)
      t->class_of_byte[c] = used[c] ? (uint8_t)K++ : 0;
    if (K > 255) { /* all 256 byte values used: no room for an "other" class, none needed */
      K = 256;
      for (int c = 0; c < 256; c++)
        t->class_of_byte[c] = (uint8_t)c;
    }
  }
  t->nb_classes = K;

  /* engine choice */
  int engine = ACM_B200_ENGINE_AUTO;
  if (!strcmp (m->engine_override, "dfa_smem"))
    engine = ACM_B200_ENGINE_DFA_SMEM;
  else if (!strcmp (m->engine_override, "dfa_global"))
    engine = ACM_B200_ENGINE_DFA_GLOBAL;
  else if (!strcmp (m->engine_override, "filter"))
    engine = ACM_B200_ENGINE_FILTER;
  const uint64_t smem_table = (uint64_t)n * K * 2;
  const int smem_ok = t->width == 1 && n <= 0xFFFF && smem_table + 1024 + 32 * 896 <= smem_budget; /* + the emit pass's per-warp queues */
  const int global_ok = t->width == 1 && (uint64_t)n * K * 4 <= (8ull << 30);
  if (engine == ACM_B200_ENGINE_DFA_SMEM && !smem_ok)
    engine = ACM_B200_ENGINE_AUTO;
  if (engine == ACM_B200_ENGINE_DFA_GLOBAL && !global_ok)
    engine = ACM_B200_ENGINE_AUTO;
  if (engine == ACM_B200_ENGINE_AUTO) {
    if (smem_ok)
      engine = ACM_B200_ENGINE_DFA_SMEM;
    else if (t->width == 1 && t->lmin < 3 && global_ok)
      engine = ACM_B200_ENGINE_DFA_GLOBAL; /* short keywords: a suffix filter would let most positions through */
    else
      engine = ACM_B200_ENGINE_FILTER;
  }
  if (t->nb_keywords == 0 && t->width != 1)
    engine = ACM_B200_ENGINE_FILTER;
  t->engine = engine;

  int rc;
  if (engine == ACM_B200_ENGINE_FILTER)
    rc = build_filter (m, t, smem_budget, smem_optin);
  else {
    t->delta_entry_bytes = engine == ACM_B200_ENGINE_DFA_SMEM ? 2 : 4;
    /* states sorted by depth (counting sort) */
    uint32_t maxd = 0;
    for (struct acm_state_block *b = m->blocks; b; b = b->next)
      for (uint32_t i = 0; i < b->used; i++)
        if (b->states[i].depth > maxd)
          maxd = b->states[i].depth;
    uint32_t *start = calloc ((size_t)maxd + 2, sizeof (uint32_t));
    struct _ac_state **by_depth = malloc ((size_t)n * sizeof (*by_depth));
    if (!start || !by_depth)
      return ACM_B200_ERR_NOMEM;
    for (struct acm_state_block *b = m->blocks; b; b = b->next)
      for (uint32_t i = 0; i < b->used; i++)
        start[b->states[i].depth + 1]++;
    for (uint32_t d = 0; d <= maxd; d++)
      start[d + 1] += start[d];
    for (struct acm_state_block *b = m->blocks; b; b = b->next)
      for (uint32_t i = 0; i < b->used; i++)
        by_depth[start[b->states[i].depth]++] = &b->states[i];
    rc = build_dfa (m, t, by_depth);
    free (start);
    free (by_depth);
  }
  if (rc)
    acm_free_tables (t);
  return rc;
}
