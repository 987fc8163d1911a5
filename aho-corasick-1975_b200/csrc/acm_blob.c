/* acm_blob.c -- a finalised dictionary on disk (the reference has no serialisation at all, aho_corasick.h:45-98).
 *
 * acm_b200_save writes the packed dictionary (keywords in id order) and the table images the finalise step produces -- class map,
 * delta table + CSR output sets, or the filter tables (shared-memory filters, distance table, q-gram set and table, reverse-trie
 * edges, keyword pools); acm_b200_load gives back a machine whose first batch scan uploads those images as they are: no insertion,
 * no table build.  The keyword trie behind the per-symbol API (acm_match, acm_get_match, carried cursors, further insertions,
 * acm_foreach_keyword) is rebuilt from the packed dictionary the first time somebody asks for it (acm_ensure_trie, acm_host.c).
 * Only machines over ACM_CMP_DEFAULT with 1/2/4-byte letters can be saved: a user comparator is code, not data.
 * User values (the void * of acm_insert_end_of_keyword) are pointers into the saving process and are not stored: keyword ids are.
 *
 * Layout: header, keyword offsets, keyword symbols, struct acm_tables (pointers meaningless), then the arrays in the order of
 * blob_arrays () each preceded by its byte count.  Little-endian, same-ABI files only (the header says so).
 */
#include "acm_internal.h"
#include "acm_tables.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ACM_BLOB_MAGIC 0x424F4C4235374341ull /* "AC75BLOB" */
#define ACM_BLOB_VERSION 3u

struct blob_header {
  uint64_t magic;
  uint32_t version, sizeof_tables;
  uint32_t width, nb_keywords, lmax, lmin, max_depth, nb_states;
  uint64_t total_symbols;
  uint64_t smem_budget, s2_smem; /* shared-memory sizes the tables were built for */
};

struct blob_array {
  void **ptr;
  size_t bytes;
};

/* every array of the table images with its size in bytes (sizes follow from the scalar fields of the struct) */
static int
blob_arrays (struct acm_tables *t, struct blob_array *a) {
  int n = 0;
  const size_t nk = t->nb_keywords;
#define ARR(field, nbytes) a[n++] = (struct blob_array){ (void **)&t->field, (size_t)(nbytes) }
  ARR (delta, t->delta_bytes);
  ARR (out_offsets, t->engine == ACM_B200_ENGINE_FILTER ? 0 : ((size_t)t->nb_dfa_states - t->out_threshold + 1) * 4);
  ARR (out_entries, t->nb_out_entries * sizeof (acm_output));
  ARR (dfa_of_state, t->engine == ACM_B200_ENGINE_FILTER ? 0 : (size_t)t->nb_states * 4);
  ARR (bloom, (size_t)t->bloom_words * 4);
  ARR (bloom2, (size_t)t->bloom2_words * 4);
  ARR (bloom_s2, (size_t)t->bloom_s2_words * 4);
  ARR (s2_dist, t->bloom_s2_words ? (size_t)4 << t->s2_dist_log2 : 0);
  ARR (kw_dist, t->bloom_s2_words ? (nk + 1) * 2 : 0);
  ARR (qgrams, (size_t)t->qgram_slots * sizeof (acm_slot));
  ARR (qset, t->engine == ACM_B200_ENGINE_FILTER && t->width != 4 ? (size_t)16 << (32 - t->qset_shift) : 0);
  ARR (edges, (size_t)t->edge_slots * sizeof (acm_slot));
  ARR (kw_len, t->engine == ACM_B200_ENGINE_FILTER ? (nk + 1) * 4 : 0);
  ARR (kw_off, t->engine == ACM_B200_ENGINE_FILTER ? (nk + 1) * 8 : 0);
  ARR (kw_pool, t->kw_pool_bytes);
  ARR (kw_meta, t->kw_rpool_words ? (nk + 1) * 8 : 0);
  ARR (kw_rpool, (size_t)t->kw_rpool_words * 4);
#undef ARR
  return n;
}

static int
put (FILE *f, const void *p, size_t n) {
  return n == 0 || fwrite (p, 1, n, f) == n;
}

static int
get (FILE *f, void *p, size_t n) {
  return n == 0 || fread (p, 1, n, f) == n;
}

int
acm_b200_save (ACMachine *m, const char *path) {
  if (!m || !path)
    return ACM_B200_ERR_INVALID;
  if (m->symbol_kind != ACM_SYM_RAW1 && m->symbol_kind != ACM_SYM_RAW2 && m->symbol_kind != ACM_SYM_RAW4)
    return ACM_B200_ERR_ALPHABET;
  acm_ensure_trie (m);
  acm_lock (m);
  int rc = ACM_B200_ERR_NOMEM;
  struct acm_tables t;
  memset (&t, 0, sizeof t);
  FILE *f = 0;
  uint64_t *offsets = 0;
  unsigned char *symbols = 0;
  /* the shared-memory sizes of the device the machine was last finalised for; without one, those of a B200 (227 KB opt-in) */
  const uint64_t optin = 232448;
  uint64_t budget = m->last_smem_budget ? m->last_smem_budget : optin - 32 * (192 * 2 + 16) - 2048;
  uint64_t s2_smem = m->last_s2_smem ? m->last_s2_smem : (m->option_no_stride2 ? 0 : (m->option_s2_smem_kb ? m->option_s2_smem_kb * 1024 - 1024 : 195 * 1024));
  if ((rc = acm_build_tables (m, &t, budget, s2_smem)))
    goto done;
  rc = ACM_B200_ERR_NOMEM;
  const size_t w = m->symbol_size, nk = m->nb_sequences;
  uint64_t total = 0;
  for (size_t r = 0; r < nk; r++)
    total += m->keywords[r]->depth;
  offsets = malloc ((nk + 1) * sizeof (*offsets));
  symbols = malloc (total * w + 1);
  if (!offsets || !symbols)
    goto done;
  uint64_t at = 0;
  for (size_t r = 0; r < nk; r++) {
    const uint32_t len = m->keywords[r]->depth;
    offsets[r] = at;
    uint32_t k = len;
    for (const struct _ac_state *s = m->keywords[r]; s->parent; s = s->parent)
      memcpy (symbols + (at + --k) * w, s->letter, w);
    at += len;
  }
  offsets[nk] = at;
  struct blob_header h = { ACM_BLOB_MAGIC, ACM_BLOB_VERSION, (uint32_t)sizeof (struct acm_tables), (uint32_t)w, (uint32_t)nk, m->lmax, m->lmin, m->max_depth,
                           (uint32_t)m->nb_states, total, budget, s2_smem };
  rc = ACM_B200_ERR_INVALID;
  if (!(f = fopen (path, "wb")))
    goto done;
  struct blob_array arr[32];
  const int na = blob_arrays (&t, arr);
  struct acm_tables image = t; /* the scalar fields; this process's pointers mean nothing to a reader and would make equal dictionaries give different files */
  {
    struct blob_array image_arr[32];
    const int ni = blob_arrays (&image, image_arr);
    for (int i = 0; i < ni; i++)
      *image_arr[i].ptr = 0;
    image.builder = 0;
  }
  int ok = put (f, &h, sizeof h) && put (f, offsets, (nk + 1) * sizeof (*offsets)) && put (f, symbols, total * w) && put (f, &image, sizeof image);
  for (int i = 0; i < na && ok; i++) {
    const uint64_t bytes = *arr[i].ptr ? arr[i].bytes : 0;
    ok = put (f, &bytes, sizeof bytes) && put (f, *arr[i].ptr, bytes);
  }
  ok = fclose (f) == 0 && ok;
  f = 0;
  rc = ok ? ACM_B200_OK : ACM_B200_ERR_INVALID;
done:
  if (f)
    fclose (f);
  free (offsets);
  free (symbols);
  acm_free_tables (&t);
  acm_unlock (m);
  return rc;
}

ACMachine *
acm_b200_load (const char *path, int *error) {
  int rc = ACM_B200_ERR_INVALID;
  ACMachine *m = 0;
  struct acm_tables *t = 0;
  uint64_t *offsets = 0;
  void *symbols = 0;
  FILE *f = path ? fopen (path, "rb") : 0;
  struct blob_header h;
  if (!f || !get (f, &h, sizeof h) || h.magic != ACM_BLOB_MAGIC || h.version != ACM_BLOB_VERSION || h.sizeof_tables != sizeof (struct acm_tables)
      || (h.width != 1 && h.width != 2 && h.width != 4))
    goto done;
  rc = ACM_B200_ERR_NOMEM;
  offsets = malloc (((size_t)h.nb_keywords + 1) * sizeof (*offsets));
  symbols = malloc (h.total_symbols * h.width + 1);
  t = calloc (1, sizeof (*t));
  if (!offsets || !symbols || !t)
    goto done;
  rc = ACM_B200_ERR_INVALID;
  if (!get (f, offsets, ((size_t)h.nb_keywords + 1) * sizeof (*offsets)) || offsets[h.nb_keywords] != h.total_symbols || !get (f, symbols, h.total_symbols * h.width)
      || !get (f, t, sizeof (*t)))
    goto done;
  {
    struct blob_array arr[32];
    const int na = blob_arrays (t, arr);
    for (int i = 0; i < na; i++)
      *arr[i].ptr = 0; /* the pointers of the saving process mean nothing here */
    t->builder = 0;    /* (nor its build structures: tables from a blob are rebuilt, not patched, after an insertion) */
    for (int i = 0; i < na; i++) {
      uint64_t bytes;
      if (!get (f, &bytes, sizeof bytes) || (bytes && bytes != arr[i].bytes))
        goto done;
      if (bytes) {
        if (!(*arr[i].ptr = malloc (bytes))) {
          rc = ACM_B200_ERR_NOMEM;
          goto done;
        }
        if (!get (f, *arr[i].ptr, bytes))
          goto done;
      }
    }
  }
  /* the machine: ACM_CMP_DEFAULT over letters of h.width bytes; cmp_arg lives inside the machine */
  {
    size_t probe = h.width;
    m = acm_create (ACM_CMP_DEFAULT, &probe, 0);
    m->owned_symbol_size = h.width;
    m->cmp_arg = &m->owned_symbol_size;
  }
  m->lazy_symbols = symbols;
  m->lazy_offsets = offsets;
  m->lazy_nb = h.nb_keywords;
  m->lazy_tables_use_state_ids = t->engine != ACM_B200_ENGINE_FILTER;
  m->lmax = h.lmax;
  m->lmin = h.lmin;
  m->max_depth = h.max_depth;
  m->preloaded = t;
  m->preloaded_generation = m->generation;
  m->preloaded_budget = h.smem_budget;
  m->preloaded_s2_smem = h.s2_smem;
  symbols = offsets = 0;
  t = 0;
  rc = ACM_B200_OK;
done:
  if (f)
    fclose (f);
  free (offsets);
  free (symbols);
  if (t) {
    acm_free_tables (t);
    free (t);
  }
  if (error)
    *error = rc;
  return rc == ACM_B200_OK ? m : 0;
}
