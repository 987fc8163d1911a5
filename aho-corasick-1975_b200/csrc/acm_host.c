/* acm_host.c -- host side of libac75.so: the keyword trie and the per-symbol (legacy) API.
 *
 * Same observable behaviour as farhiongit/aho-corasick-1975 (reference aho_corasick.c), own implementation:
 *  - states live in arena blocks and keep their children in a cmp-sorted pointer array (the reference keeps one ordered
 *    map per state in an external container library, aho_corasick.c:47,102);
 *  - fail links, inverse fail sets and output counts are maintained incrementally on every insertion (Meyer, 1985),
 *    so insertions can be interleaved with scanning on a carried cursor (reference aho_corasick.c:194-267,318-338);
 *  - every change bumps machine->generation; the GPU tables are rebuilt lazily from it (acm_finalise.c / acm_device.cu).
 * Contract violations follow the reference's convention (aho_corasick.c:24-36): message on stderr, thrd_exit(EXIT_FAILURE).
 */
#include "acm_internal.h"
#include <stdlib.h>
#include <string.h>
#include <threads.h>

#define REQUIRE(cond, msg)           \
  do {                               \
    if (!(cond))                     \
      acm_fatal (__func__, (msg));   \
  } while (0)

void
acm_fatal (const char *function, const char *message) {
  fflush (stdout);
  fprintf (stderr, "FATAL ERROR: A prerequisite is not fulfilled in function %s.\n", function);
  fprintf (stderr, "             %s\n", message);
  thrd_exit (EXIT_FAILURE);
}

void
acm_lock (struct _ac_machine *m) {
  REQUIRE (mtx_lock ((mtx_t *)m->token) == thrd_success, "The machine lock could not be taken.");
}

void
acm_unlock (struct _ac_machine *m) {
  REQUIRE (mtx_unlock ((mtx_t *)m->token) == thrd_success, "The machine lock could not be released.");
}

/* ---- letters ------------------------------------------------------------------------------------------------------ */
static int
default_compare (const void *a, const void *b, const void *cmp_arg) {
  return memcmp (a, b, *(const size_t *)cmp_arg); /* reference aho_corasick.c:134-138 */
}
const CMP_TYPE ACM_CMP_DEFAULT = default_compare;
const int ACM_INCREMENTAL_STRING_MATCHING = 1;

static inline int
compare_letters (const struct _ac_machine *m, const void *a, const void *b) {
  switch (m->symbol_kind) {
    case ACM_SYM_RAW1:
      return (int)*(const unsigned char *)a - (int)*(const unsigned char *)b;
    case ACM_SYM_RAW2:
      return memcmp (a, b, 2);
    case ACM_SYM_RAW4:
      return memcmp (a, b, 4);
    default:
      return m->cmp (a, b, m->cmp_arg);
  }
}

/* ACM_CMP_DEFAULT over 1/2/4-byte letters: the letter as an integer whose order is the comparator's (memcmp) order.  A state with
 * more than one child keeps these keys right behind its child pointers (one allocation), so that looking a letter up reads one
 * contiguous array instead of two dependent cache lines per probe (child state, then its letter): acm_find_child was 58 % of the
 * time it takes to load 10^6 keywords. */
static inline int
raw_letters (const struct _ac_machine *m) {
  return m->symbol_kind == ACM_SYM_RAW1 || m->symbol_kind == ACM_SYM_RAW2 || m->symbol_kind == ACM_SYM_RAW4;
}

static inline uint32_t
letter_key (const struct _ac_machine *m, const void *letter) {
  const unsigned char *b = letter;
  switch (m->symbol_kind) {
    case ACM_SYM_RAW1:
      return b[0];
    case ACM_SYM_RAW2:
      return ((uint32_t)b[0] << 8) | b[1];
    default:
      return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
  }
}

static inline uint32_t *
child_keys (const struct _ac_state *s) { /* only for heap-allocated child arrays of raw-letter machines */
  return (uint32_t *)(s->children + s->cap_children);
}

/* Position of `letter` among the children of s: index of the match, or ~(insertion point). */
static inline int64_t
child_slot (const struct _ac_state *s, const void *letter) {
  const struct _ac_machine *m = s->machine;
  uint32_t lo = 0, hi = s->nb_children;
  if (s->children != &s->inline_child && raw_letters (m)) {
    const uint32_t key = letter_key (m, letter), *keys = child_keys (s);
    while (lo < hi) {
      const uint32_t mid = lo + (hi - lo) / 2;
      if (keys[mid] == key)
        return mid;
      if (key < keys[mid])
        hi = mid;
      else
        lo = mid + 1;
    }
    return ~(int64_t)lo;
  }
  while (lo < hi) {
    uint32_t mid = lo + (hi - lo) / 2;
    int c = compare_letters (m, letter, s->children[mid]->letter);
    if (c == 0)
      return mid;
    if (c < 0)
      hi = mid;
    else
      lo = mid + 1;
  }
  return ~(int64_t)lo;
}

struct _ac_state *
acm_find_child (const struct _ac_state *s, const void *letter) {
  int64_t k = child_slot (s, letter);
  return k >= 0 ? s->children[k] : 0;
}

/* delta(s, a): the child if there is one, else retry from f(s); state 0 loops on itself (reference aho_corasick.c:167-192). */
const struct _ac_state *
acm_host_goto (const struct _ac_state *s, const void *letter) {
  for (;;) {
    struct _ac_state *c = acm_find_child (s, letter);
    if (c)
      return c;
    if (!s->parent)
      return s;
    s = s->fail;
  }
}

/* ---- states ------------------------------------------------------------------------------------------------------- */
static struct _ac_state *
new_state (struct _ac_machine *m) {
  if (!m->last_block || m->last_block->used == ACM_STATES_PER_BLOCK) {
    struct acm_state_block *b = calloc (1, sizeof (*b) + ACM_STATES_PER_BLOCK * sizeof (struct _ac_state));
    REQUIRE (b, "Out of memory.");
    if (m->last_block)
      m->last_block->next = b;
    else
      m->blocks = b;
    m->last_block = b;
  }
  struct _ac_state *s = &m->last_block->states[m->last_block->used++];
  s->machine = m;
  s->id = (uint32_t)m->nb_states++;
  s->rank = ACM_NONE;
  s->children = &s->inline_child;
  s->cap_children = 1;
  return s;
}

static void
add_child (struct _ac_state *s, struct _ac_state *child, uint32_t at) {
  const int keyed = raw_letters (s->machine);
  if (s->nb_children == s->cap_children) {
    uint32_t cap = s->cap_children * 2;
    struct _ac_state **grown;
    if (keyed) { /* pointers, then keys: a new block, both arrays copied over */
      grown = malloc (cap * (sizeof (*grown) + sizeof (uint32_t)));
      REQUIRE (grown, "Out of memory.");
      uint32_t *keys = (uint32_t *)(grown + cap);
      if (s->children == &s->inline_child) {
        grown[0] = s->inline_child;
        keys[0] = letter_key (s->machine, s->inline_child->letter);
      } else {
        memcpy (grown, s->children, s->nb_children * sizeof (*grown));
        memcpy (keys, child_keys (s), s->nb_children * sizeof (uint32_t));
        free (s->children);
      }
    } else if (s->children == &s->inline_child) {
      grown = malloc (cap * sizeof (*grown));
      if (grown)
        grown[0] = s->inline_child;
    } else
      grown = realloc (s->children, cap * sizeof (*grown));
    REQUIRE (grown, "Out of memory.");
    s->children = grown;
    s->cap_children = cap;
  }
  memmove (s->children + at + 1, s->children + at, (s->nb_children - at) * sizeof (*s->children));
  s->children[at] = child;
  if (keyed && s->children != &s->inline_child) {
    uint32_t *keys = child_keys (s);
    memmove (keys + at + 1, keys + at, (s->nb_children - at) * sizeof (uint32_t));
    keys[at] = letter_key (s->machine, child->letter);
  }
  s->nb_children++;
}

static void
ifs_add (struct _ac_state *target, struct _ac_state *x) {
  if (target->nb_ifs == target->cap_ifs) {
    uint32_t cap = target->cap_ifs ? target->cap_ifs * 2 : 2;
    struct _ac_state **grown = realloc (target->ifs, cap * sizeof (*grown));
    REQUIRE (grown, "Out of memory.");
    target->ifs = grown;
    target->cap_ifs = cap;
  }
  x->if_index = target->nb_ifs;
  target->ifs[target->nb_ifs++] = x;
  x->fail = target;
}

static void
ifs_remove (struct _ac_state *x) {
  struct _ac_state *t = x->fail;
  struct _ac_state *last = t->ifs[--t->nb_ifs];
  t->ifs[x->if_index] = last;
  last->if_index = x->if_index;
}

static void
scratch_push (struct _ac_machine *m, struct _ac_state *s) {
  if (m->nb_scratch == m->cap_scratch) {
    size_t cap = m->cap_scratch ? m->cap_scratch * 2 : 256;
    struct _ac_state **grown = realloc (m->scratch, cap * sizeof (*grown));
    REQUIRE (grown, "Out of memory.");
    m->scratch = grown;
    m->cap_scratch = cap;
  }
  m->scratch[m->nb_scratch++] = s;
}

/* Meyer'85: a new node n' = T[n,c] becomes the longest proper suffix of every existing x' = T[x,c] where x ranges over the
 * states reachable from IF[n] through IF without meeting a c-edge on the way (reference aho_corasick.c:224-239,264).
 * Iterative over the machine's work stack; the stack is seeded with a copy of IF[n] because re-pointing x' may edit IF[n]. */
static void
repoint_longer_suffixes (struct _ac_machine *m, struct _ac_state *n, struct _ac_state *nprime) {
  size_t floor = m->nb_scratch;
  for (uint32_t i = 0; i < n->nb_ifs; i++)
    scratch_push (m, n->ifs[i]);
  while (m->nb_scratch > floor) {
    struct _ac_state *x = m->scratch[--m->nb_scratch];
    struct _ac_state *xprime = acm_find_child (x, nprime->letter);
    if (xprime) {
      if (xprime != nprime && xprime->fail != nprime) {
        ifs_remove (xprime);
        ifs_add (nprime, xprime);
      }
    } else
      for (uint32_t i = 0; i < x->nb_ifs; i++)
        scratch_push (m, x->ifs[i]);
  }
}

/* reference aho_corasick.c:242-267 (enter_child) + :194-208 (complete_fail_state) */
static struct _ac_state *
enter_child (struct _ac_state *n, void *letter, uint32_t at) {
  struct _ac_machine *m = n->machine;
  struct _ac_state *nprime = new_state (m);
  nprime->parent = n;
  nprime->letter = letter; /* kept by pointer, never copied */
  nprime->depth = n->depth + 1;
  if (nprime->depth > m->max_depth)
    m->max_depth = nprime->depth;
  /* f(n') before n' is linked: delta(f(n), c), or state 0 for the children of state 0 */
  if (m->bulk) { /* links are rebuilt in one pass by rebuild_links() */
    add_child (n, nprime, at);
    m->generation++;
    return nprime;
  }
  struct _ac_state *f = n->parent ? (struct _ac_state *)acm_host_goto (n->fail, letter) : n;
  add_child (n, nprime, at);
  ifs_add (f, nprime);
  nprime->nb_outputs = f->nb_outputs;
  repoint_longer_suffixes (m, n, nprime);
  m->generation++;
  return nprime;
}

/* Whole-machine recomputation of f, IF and the output counts by one breadth-first pass (Aho-Corasick Algorithm 3,
 * reference aho_corasick.c:386-417); used after a bulk load instead of per-node incremental maintenance. Same result. */
static void
rebuild_links (struct _ac_machine *m) {
  const size_t n = m->nb_states;
  struct _ac_state **order = malloc (n * sizeof (*order));
  REQUIRE (order, "Out of memory.");
  for (struct acm_state_block *b = m->blocks; b; b = b->next)
    for (uint32_t i = 0; i < b->used; i++)
      b->states[i].nb_ifs = 0;
  size_t head = 0, tail = 0;
  order[tail++] = m->root;
  m->root->fail = 0;
  m->root->nb_outputs = 0;
  while (head < tail) {
    struct _ac_state *r = order[head++];
    for (uint32_t k = 0; k < r->nb_children; k++) {
      struct _ac_state *s = r->children[k];
      struct _ac_state *f = r->parent ? (struct _ac_state *)acm_host_goto (r->fail, s->letter) : r;
      ifs_add (f, s);
      s->nb_outputs = (s->rank != ACM_NONE ? 1u : 0u) + f->nb_outputs;
      order[tail++] = s;
    }
  }
  free (order);
}

/* ---- public: lifecycle -------------------------------------------------------------------------------------------- */
ACMachine *
acm_create (CMP_TYPE cmp, void *cmp_arg, DESTROY_TYPE dtor) {
  REQUIRE (cmp, "A comparison function should be provided.");
  struct _ac_machine *m = calloc (1, sizeof (*m));
  REQUIRE (m, "Out of memory.");
  m->cmp = cmp;
  m->cmp_arg = cmp_arg;
  m->dtor = dtor;
  m->symbol_kind = ACM_SYM_CUSTOM;
  if (cmp == ACM_CMP_DEFAULT) {
    REQUIRE (cmp_arg, "ACM_CMP_DEFAULT needs a pointer to the size of a letter.");
    m->symbol_size = *(const size_t *)cmp_arg;
    m->symbol_kind = m->symbol_size == 1 ? ACM_SYM_RAW1 : m->symbol_size == 2 ? ACM_SYM_RAW2 : m->symbol_size == 4 ? ACM_SYM_RAW4 : ACM_SYM_RAWN;
  }
  m->token = malloc (sizeof (mtx_t));
  REQUIRE (m->token && mtx_init ((mtx_t *)m->token, mtx_plain) == thrd_success, "Out of memory.");
  m->lmin = ACM_NONE;
  m->root = new_state (m);
  m->generation = 1;
  return m;
}

void
acm_release (ACMachine *m) {
  REQUIRE (m, "Invalid null machine.");
  if (m->device)
    acm_device_release (m->device);
  for (struct acm_state_block *b = m->blocks; b;) {
    for (uint32_t i = 0; i < b->used; i++) {
      struct _ac_state *s = &b->states[i];
      if (s->parent && m->dtor)
        m->dtor (s->letter);
      if (s->value_dtor)
        s->value_dtor (s->value);
      if (s->children != &s->inline_child)
        free (s->children);
      free (s->ifs);
    }
    struct acm_state_block *next = b->next;
    free (b);
    b = next;
  }
  while (m->arena) {
    struct acm_arena *a = m->arena;
    m->arena = a->next;
    free (a);
  }
  free (m->lazy_symbols);
  free (m->lazy_offsets);
  if (m->preloaded) {
    acm_free_tables (m->preloaded);
    free (m->preloaded);
  }
  free (m->keywords);
  free (m->scratch);
  free (m->class_letter);
  free (m->class_sorted_id);
  free (m->class_of_state);
  mtx_destroy ((mtx_t *)m->token);
  free (m->token);
  free (m);
}

ACState *
acm_initiate (ACMachine *m) {
  REQUIRE (m, "Invalid null machine.");
  acm_ensure_trie (m);
  return m->root;
}

/* ---- public: dictionary ------------------------------------------------------------------------------------------- */
/* insertion of one letter with the machine lock held */
static void
insert_letter_locked (struct _ac_machine *m, ACState **state, void *letter) {
  int64_t k = child_slot (*state, letter);
  if (k >= 0) {
    *state = (*state)->children[k];
    if (m->dtor)
      m->dtor (letter); /* the edge keeps its first letter; the duplicate is handed back at once */
  } else
    *state = enter_child (*state, letter, (uint32_t)~k);
}

void
acm_insert_letter_of_keyword (ACState **state, void *letter) {
  REQUIRE (state && *state && letter, "Invalid null state or letter.");
  struct _ac_machine *m = (*state)->machine;
  acm_lock (m);
  insert_letter_locked (m, state, letter);
  acm_unlock (m);
}

/* +1 on the node and on every node that has it as a suffix, i.e. its IF closure (reference aho_corasick.c:330-338) */
static void
count_new_output (struct _ac_machine *m, struct _ac_state *n) {
  size_t floor = m->nb_scratch;
  scratch_push (m, n);
  while (m->nb_scratch > floor) {
    struct _ac_state *x = m->scratch[--m->nb_scratch];
    x->nb_outputs++;
    for (uint32_t i = 0; i < x->nb_ifs; i++)
      scratch_push (m, x->ifs[i]);
  }
}

/* end of a keyword with the machine lock held; *state must not be state 0 */
static void *
insert_end_locked (struct _ac_machine *m, ACState **state, void *value, void (*dtor) (void *)) {
  struct _ac_state *s = *state;
  if (s->rank == ACM_NONE) {
    if (m->nb_sequences == m->cap_keywords) {
      size_t cap = m->cap_keywords ? m->cap_keywords * 2 : 64;
      struct _ac_state **grown = realloc (m->keywords, cap * sizeof (*grown));
      REQUIRE (grown, "Out of memory.");
      m->keywords = grown;
      m->cap_keywords = cap;
    }
    s->rank = (uint32_t)m->nb_sequences;
    m->keywords[m->nb_sequences++] = s;
    if (!m->bulk)
      count_new_output (m, s);
    if (s->depth > m->lmax)
      m->lmax = s->depth;
    if (s->depth < m->lmin)
      m->lmin = s->depth;
    m->generation++;
  }
  void *previous = s->value;
  if (!s->value) {
    s->value = value;
    s->value_dtor = dtor;
  }
  *state = m->root;
  return previous;
}

void *
acm_insert_end_of_keyword (ACState **state, void *value, void (*dtor) (void *)) {
  REQUIRE (state && *state, "Invalid null state.");
  struct _ac_machine *m = (*state)->machine;
  acm_lock (m);
  if (!(*state)->parent) {
    acm_unlock (m); /* do not die holding the lock */
    acm_fatal (__func__, "acm_insert_letter_of_keyword should be called first.");
  }
  void *previous = insert_end_locked (m, state, value, dtor);
  acm_unlock (m);
  return previous;
}

size_t
acm_nb_keywords (const ACMachine *m) {
  REQUIRE (m, "Invalid null machine.");
  return m->lazy_symbols ? (size_t)m->lazy_nb : m->nb_sequences;
}

/* ---- public: per-symbol scan -------------------------------------------------------------------------------------- */
size_t
acm_match (const ACState **state, const void *letter) {
  REQUIRE (state && *state && letter, "Invalid null state or letter.");
  return (*state = acm_host_goto (*state, letter))->nb_outputs;
}

void
acm_matcher_init (MatchHolder *matcher) {
  REQUIRE (matcher, "Invalid null matcher.");
  memset (matcher, 0, sizeof (*matcher));
}

void
acm_matcher_release (MatchHolder *matcher) {
  REQUIRE (matcher, "Invalid null matcher.");
  free (matcher->letters);
  memset (matcher, 0, sizeof (*matcher));
}

static void
fill_holder (const struct _ac_state *terminal, MatchHolder *matcher) {
  matcher->length = terminal->depth;
  const void **grown = realloc (matcher->letters, (terminal->depth ? terminal->depth : 1) * sizeof (*grown));
  REQUIRE (grown, "Out of memory.");
  matcher->letters = grown;
  size_t k = terminal->depth;
  for (const struct _ac_state *s = terminal; s->parent; s = s->parent)
    matcher->letters[--k] = s->letter;
  matcher->value = terminal->value;
}

void
acm_get_match (const ACState *state, size_t index, MatchHolder *matcher) {
  REQUIRE (state, "Invalid null state.");
  REQUIRE (state->parent, "acm_match should be called first and acm_matcher_init called on the MatchHolder.");
  REQUIRE (index < state->nb_outputs, "Index out of bounds.");
  /* the index-th keyword state met along s, f(s), f(f(s)), ... : longest first (reference aho_corasick.c:459-466) */
  for (size_t seen = 0;; state = state->fail) {
    if (state->rank == ACM_NONE)
      continue;
    if (seen++ == index)
      break;
  }
  if (matcher)
    fill_holder (state, matcher);
}

int
acm_b200_keyword (const ACMachine *m, uint32_t keyword, MatchHolder *holder) {
  if (!m || !holder)
    return ACM_B200_ERR_INVALID;
  acm_ensure_trie ((ACMachine *)m);
  if (keyword >= m->nb_sequences)
    return ACM_B200_ERR_INVALID;
  fill_holder (m->keywords[keyword], holder);
  return ACM_B200_OK;
}

/* ---- public: introspection ---------------------------------------------------------------------------------------- */
static int acm_walk_keywords (const ACMachine *m, void (*op) (MatchHolder), uint32_t *ids, uint64_t capacity, uint64_t *nb);

struct walk_frame {
  const struct _ac_state *state;
  uint32_t next_child;
};

void
acm_foreach_keyword (const ACMachine *m, void (*op) (MatchHolder)) {
  REQUIRE (m, "Invalid null machine.");
  if (!op)
    return;
  acm_ensure_trie ((ACMachine *)m);
  acm_walk_keywords (m, op, 0, 0, 0);
}

/* keyword ids in the order acm_foreach_keyword visits the keywords (reference aho_corasick.c:490-531: pre-order, children in
 * comparator order): ids[k] = id of the k-th keyword the callback would receive.  *nb receives the number of keywords. */
int
acm_b200_keyword_order (const ACMachine *m, uint32_t *ids, uint64_t capacity, uint64_t *nb) {
  if (!m || (capacity && !ids))
    return ACM_B200_ERR_INVALID;
  acm_ensure_trie ((ACMachine *)m);
  return acm_walk_keywords (m, 0, ids, capacity, nb);
}

static int
acm_walk_keywords (const ACMachine *m, void (*op) (MatchHolder), uint32_t *ids, uint64_t capacity, uint64_t *nb) {
  /* pre-order, children in comparator order (reference aho_corasick.c:512-531) */
  size_t cap = (size_t)m->lmax + 2, top = 0;
  uint64_t seen = 0;
  struct walk_frame *stack = malloc (cap * sizeof (*stack));
  const void **letters = malloc (cap * sizeof (*letters));
  REQUIRE (stack && letters, "Out of memory.");
  stack[top++] = (struct walk_frame){ m->root, 0 };
  while (top) {
    struct walk_frame *f = &stack[top - 1];
    if (f->next_child == 0 && f->state->rank != ACM_NONE) {
      if (op)
        op ((MatchHolder){ .letters = letters, .length = f->state->depth, .value = f->state->value });
      if (seen < capacity)
        ids[seen] = f->state->rank;
      seen++;
    }
    if (f->next_child < f->state->nb_children) {
      const struct _ac_state *c = f->state->children[f->next_child++];
      if (c->depth + 1 >= cap) { /* keywords longer than lmax may be under construction */
        cap *= 2;
        stack = realloc (stack, cap * sizeof (*stack));
        letters = realloc (letters, cap * sizeof (*letters));
        REQUIRE (stack && letters, "Out of memory.");
      }
      letters[c->depth - 1] = c->letter;
      stack[top++] = (struct walk_frame){ c, 0 };
    } else
      top--;
  }
  free (stack);
  free (letters);
  if (nb)
    *nb = seen;
  return seen > capacity && capacity ? ACM_B200_ERR_CAPACITY : ACM_B200_OK;
}

/* Debug dump, same layout as the reference's (aho_corasick.c:533-594): one line per branch,
 * (id)---letter-->(id)[+outputs](v fail id), continuation branches aligned under their parent with an 'L'. */
static void
print_branches (const struct _ac_state *s, FILE *stream, int *column, int indent, PRINT_TYPE printer) {
  REQUIRE (s->rank == ACM_NONE || s->nb_outputs, "Keyword without defined output.");
  REQUIRE (s->parent ? s->fail != 0 : s->fail == 0, "Incorrect fail state.");
  for (uint32_t k = 0; k < s->nb_children; k++) {
    const struct _ac_state *c = s->children[k];
    if (indent < *column) {
      fputc ('\n', stream);
      *column = 0;
      if (indent) {
        for (int t = 0; t < indent - 1; t++)
          *column += fprintf (stream, " ");
        *column += fprintf (stream, "L");
      }
    } else
      for (int pad = indent - *column; pad > 0; pad--)
        *column += fprintf (stream, " ");
    if (!s->parent)
      *column += fprintf (stream, "(%03zu)", (size_t)s->id);
    *column += fprintf (stream, "---");
    REQUIRE (c->parent == s, "Incorrect previous state.");
    if (printer)
      *column += printer (stream, c->letter);
    *column += fprintf (stream, "-->(%03zu)", (size_t)c->id);
    if (c->rank != ACM_NONE)
      *column += fprintf (stream, "[+%zu]", c->nb_outputs);
    if (c->fail != s->machine->root)
      *column += fprintf (stream, "(v %03zu)", (size_t)c->fail->id);
    print_branches (c, stream, column, *column, printer);
  }
}

void
acm_print (ACMachine *m, FILE *stream, PRINT_TYPE printer) {
  REQUIRE (m, "Invalid null machine.");
  if (!stream)
    return;
  acm_ensure_trie (m);
  int column = 0;
  fprintf (stream, "\n");
  print_branches (m->root, stream, &column, 0, printer);
  fprintf (stream, "\n");
}

/* ---- batch helpers that need no GPU ------------------------------------------------------------------------------- */
size_t
acm_b200_symbol_width (const ACMachine *m) {
  if (!m)
    return 0;
  return (m->symbol_kind == ACM_SYM_RAW1 || m->symbol_kind == ACM_SYM_RAW2 || m->symbol_kind == ACM_SYM_RAW4) ? m->symbol_size : 4;
}

uint32_t
acm_b200_max_keyword_length (const ACMachine *m) {
  return m ? m->lmax : 0;
}

/* Bulk insertion for callers that hold the dictionary as packed arrays: keyword k = symbols[offsets[k] .. offsets[k+1]) in letters
 * of acm_b200_symbol_width() bytes.  The letters are copied into storage owned by the machine (the per-letter API keeps the
 * caller's pointers, reference aho_corasick.c:248).  Only for machines created without a letter destructor.  ids[k] (optional)
 * receives the keyword id, an existing id for a duplicate. */
static int
insert_packed_locked (ACMachine *m, const void *symbols, const uint64_t *offsets, uint64_t nb, uint32_t *ids) {
  const size_t w = m->symbol_size;
  const int bulk = nb >= 256; /* below that the incremental maintenance is cheaper than a full pass */
  m->bulk = bulk;
  int rc = ACM_B200_OK;
  for (uint64_t k = 0; k < nb && rc == ACM_B200_OK; k++) {
    const size_t len = (size_t)(offsets[k + 1] - offsets[k]), bytes = (len * w + 7) & ~(size_t)7;
    if (!len) {
      rc = ACM_B200_ERR_INVALID; /* the empty keyword is forbidden (reference aho_corasick.c:345) */
      break;
    }
    if (!m->arena || m->arena->used + bytes > m->arena->cap) {
      size_t cap = bytes > (1u << 22) ? bytes : (1u << 22);
      struct acm_arena *a = malloc (sizeof (*a) + cap);
      if (!a) {
        rc = ACM_B200_ERR_NOMEM;
        break;
      }
      a->next = m->arena;
      a->used = 0;
      a->cap = cap;
      m->arena = a;
    }
    unsigned char *copy = m->arena->bytes + m->arena->used;
    memcpy (copy, (const unsigned char *)symbols + offsets[k] * w, len * w);
    ACState *s = m->root;
    const size_t states_before = m->nb_states;
    for (size_t i = 0; i < len; i++)
      insert_letter_locked (m, &s, copy + i * w);
    if (m->nb_states != states_before)
      m->arena->used += bytes; /* at least one letter pointer was kept */
    const uint32_t id = s->rank != ACM_NONE ? s->rank : (uint32_t)m->nb_sequences;
    insert_end_locked (m, &s, 0, 0);
    if (ids)
      ids[k] = id;
  }
  if (bulk) { /* the states inserted above have no links yet: nobody else sees them, the lock is still held */
    m->bulk = 0;
    rebuild_links (m);
  }
  return rc;
}

int
acm_b200_insert_keywords (ACMachine *m, const void *symbols, const uint64_t *offsets, uint64_t nb, uint32_t *ids) {
  if (!m || !offsets || (!symbols && nb && offsets[nb]) || m->dtor)
    return ACM_B200_ERR_INVALID;
  if (m->symbol_kind != ACM_SYM_RAW1 && m->symbol_kind != ACM_SYM_RAW2 && m->symbol_kind != ACM_SYM_RAW4)
    return ACM_B200_ERR_ALPHABET;
  acm_ensure_trie (m);
  acm_lock (m); /* held for the whole load: a scan or an insertion on another thread never sees a state without its links */
  const int rc = insert_packed_locked (m, symbols, offsets, nb, ids);
  acm_unlock (m);
  return rc;
}

/* A machine loaded from a blob (acm_blob.c) keeps its dictionary packed until somebody needs the trie: the per-symbol API, a
 * carried cursor, an insertion, the keyword enumeration.  The batch scan itself only needs the tables of the blob. */
void
acm_ensure_trie_locked (ACMachine *m) {
  if (!m->lazy_symbols)
    return;
  void *symbols = m->lazy_symbols;
  uint64_t *offsets = m->lazy_offsets;
  const uint64_t generation = m->generation;
  const int rc = insert_packed_locked (m, symbols, offsets, m->lazy_nb, 0); /* rank order: every keyword gets its id back */
  REQUIRE (rc == ACM_B200_OK && m->nb_sequences == m->lazy_nb, "The packed dictionary of the blob is inconsistent.");
  m->lazy_symbols = 0;
  m->lazy_offsets = 0;
  free (symbols);
  free (offsets);
  /* same dictionary, same tables -- except the state-id map of the DFA engines (ids are creation order), which a carried
   * cursor needs: those tables are rebuilt at the next scan */
  m->generation = m->lazy_tables_use_state_ids ? generation + 1 : generation;
  if (!m->lazy_tables_use_state_ids) {
    if (m->preloaded)
      m->preloaded_generation = m->generation;
  }
}

void
acm_ensure_trie (ACMachine *m) {
  if (!m->lazy_symbols)
    return;
  acm_lock (m);
  acm_ensure_trie_locked (m);
  acm_unlock (m);
}
