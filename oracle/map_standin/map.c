/* TEST INFRASTRUCTURE ONLY -- stand-in for minimaps, see map.h.
 *
 * Ordered unique-key container as a treap with parent links.  Design points forced by the
 * reference's usage:
 *  - an operator may free the element before asking for its removal (aho_corasick.c:114-115), so
 *    unlinking never re-reads the element or its key: nodes are unlinked by node pointer;
 *  - operators mutate maps re-entrantly, including removing a not-yet-visited element from the map
 *    being traversed (aho_corasick.c:264 -> :236 -> :217, e.g. inserting "aa" while "a","ba","baa","baaa"
 *    exist moves "baa" out of IF["a"] while IF["a"] is being walked).  Meyer's loop "for x in IF[n]" is over
 *    the set as it was when the loop started, so traversal runs over a SNAPSHOT and still visits elements
 *    unlinked meanwhile (their data, &transition->state, stays valid in the reference's usage).  Skipping
 *    them instead makes the result depend on pointer order (the IF sets are keyed by pointer value): observed
 *    298,853 instead of 298,855 matches on the config-1 dictionary.  Unlinked nodes are only marked dead
 *    while a traversal is active and reclaimed afterwards, so the snapshot never dangles;
 *  - map_find_key returns the number of elements the operator was applied to, counting an element
 *    for which the operator returned 0 (aho_corasick.c:232 relies on this);
 *  - map_find_key with MAP_GET_ONE writes nothing to the map: the reference lets several threads scan one machine without
 *    a lock (README.md:266,364), and the --impl reference arm of bench.py does exactly that;
 *  - no call to rand(): the reference's third test depends on the unseeded rand() sequence.
 */
#include "map.h"
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

struct node {
  struct node *left, *right, *parent;
  void *data;
  uint32_t prio;
  int dead;
  struct node *next_dead;
};

struct map {
  struct node *root;
  size_t size;
  map_key_extractor get_key;
  map_key_comparator cmp;
  const void *cmp_arg;
  int unique;
  int traversals;          /* depth of active traversals over this map */
  struct node *dead_list;  /* unlinked while a traversal was active */
};

static int
generic_cmp (const void *a, const void *b, const void *arg) {
  return memcmp (a, b, *(const size_t *)arg);
}
static int
get_one (void *data, void *op_arg, int *remove) {
  (void)remove;
  *(void **)op_arg = data;
  return 0;
}
static int
remove_all (void *data, void *op_arg, int *remove) {
  (void)data;
  (void)op_arg;
  *remove = 1;
  return 1;
}
const map_key_comparator MAP_GENERIC_CMP = generic_cmp;
const map_operator MAP_GET_ONE = get_one;
const map_operator MAP_REMOVE_ALL = remove_all;

static uint32_t
next_prio (void) {
  static _Thread_local uint64_t s = 0x9E3779B97F4A7C15ull;
  s ^= s << 13;
  s ^= s >> 7;
  s ^= s << 17;
  return (uint32_t)(s >> 32);
}

static const void *
key_of (const map *m, void *data) {
  return m->get_key ? m->get_key (data) : data;
}

map *
map_create (map_key_extractor get_key, map_key_comparator cmp, const void *cmp_arg, int unique) {
  if (!cmp)
    return 0;
  map *m = calloc (1, sizeof (*m));
  if (!m)
    return 0;
  m->get_key = get_key;
  m->cmp = cmp;
  m->cmp_arg = cmp_arg;
  m->unique = unique;
  return m;
}

int
map_destroy (map *m) {
  if (!m || m->size || m->traversals)
    return 0;
  free (m);
  return 1;
}

size_t
map_size (const map *m) {
  return m ? m->size : 0;
}

static void
replace_child (map *m, struct node *parent, struct node *old, struct node *new) {
  if (!parent)
    m->root = new;
  else if (parent->left == old)
    parent->left = new;
  else
    parent->right = new;
  if (new)
    new->parent = parent;
}

static void
rotate_up (map *m, struct node *n) { /* n replaces its parent */
  struct node *p = n->parent, *g = p->parent;
  if (p->left == n) {
    p->left = n->right;
    if (n->right)
      n->right->parent = p;
    n->right = p;
  } else {
    p->right = n->left;
    if (n->left)
      n->left->parent = p;
    n->left = p;
  }
  p->parent = n;
  replace_child (m, g, p, n);
}

int
map_insert_data (map *m, void *data) {
  if (!m || !data)
    return 0;
  const void *key = key_of (m, data);
  struct node *p = 0, **link = &m->root;
  while (*link) {
    p = *link;
    int c = m->cmp (key, key_of (m, p->data), m->cmp_arg);
    if (c == 0 && m->unique)
      return 0;
    link = c < 0 ? &p->left : &p->right;
  }
  struct node *n = calloc (1, sizeof (*n));
  if (!n)
    return 0;
  n->data = data;
  n->prio = next_prio ();
  n->parent = p;
  *link = n;
  while (n->parent && n->parent->prio < n->prio)
    rotate_up (m, n);
  m->size++;
  return 1;
}

/* Unlink by node pointer: never touches n->data. */
static void
unlink_node (map *m, struct node *n) {
  while (n->left || n->right) {
    struct node *c = !n->left ? n->right : (!n->right ? n->left : (n->left->prio > n->right->prio ? n->left : n->right));
    rotate_up (m, c);
  }
  replace_child (m, n->parent, n, 0);
  m->size--;
  n->dead = 1;
  if (m->traversals) {
    n->next_dead = m->dead_list;
    m->dead_list = n;
  } else
    free (n);
}

static void
end_traversal (map *m) {
  if (--m->traversals == 0)
    while (m->dead_list) {
      struct node *d = m->dead_list;
      m->dead_list = d->next_dead;
      free (d);
    }
}

size_t
map_find_key (map *m, const void *key, map_operator op, void *op_arg, map_selector sel, void *sel_arg) {
  if (!m || !key)
    return 0;
  struct node *n = m->root;
  while (n) {
    int c = m->cmp (key, key_of (m, n->data), m->cmp_arg);
    if (c == 0)
      break;
    n = c < 0 ? n->left : n->right;
  }
  if (!n || (sel && !sel (n->data, sel_arg)))
    return 0;
  if (op == get_one) { /* the scan path (aho_corasick.c:175) runs lock-free from many threads: strictly read-only */
    *(void **)op_arg = n->data;
    return 1;
  }
  if (op) {
    int remove = 0;
    m->traversals++;
    (void)op (n->data, op_arg, &remove);
    if (remove && !n->dead)
      unlink_node (m, n);
    end_traversal (m);
  }
  return 1;
}

size_t
map_traverse (map *m, map_operator op, void *op_arg, map_selector sel, void *sel_arg) {
  if (!m || !m->size)
    return 0;
  size_t n_snap = m->size, i = 0, applied = 0;
  struct node **snap = malloc (n_snap * sizeof (*snap));
  if (!snap)
    return 0;
  /* in-order walk using parent links */
  struct node *n = m->root;
  while (n->left)
    n = n->left;
  while (n) {
    snap[i++] = n;
    if (n->right) {
      n = n->right;
      while (n->left)
        n = n->left;
    } else {
      while (n->parent && n->parent->right == n)
        n = n->parent;
      n = n->parent;
    }
  }
  m->traversals++;
  for (i = 0; i < n_snap; i++) {
    struct node *e = snap[i];
    /* e may have been unlinked by a re-entrant operator: still visited, see the header comment */
    if (sel && !sel (e->data, sel_arg))
      continue;
    int remove = 0, cont = op ? op (e->data, op_arg, &remove) : 1;
    applied++;
    if (remove && !e->dead)
      unlink_node (m, e);
    if (!cont)
      break;
  }
  end_traversal (m);
  free (snap);
  return applied;
}
