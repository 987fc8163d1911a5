/* TEST INFRASTRUCTURE ONLY -- stand-in for github.com/farhiongit/minimaps ("map.h").
 *
 * The reference (aho_corasick.c:18) includes "map.h" from a sibling checkout of minimaps that is
 * not vendored, not version-pinned and not reachable (no network).  This header + map.c implement
 * ONLY the container contract the reference relies on, inferred from its 24 call sites
 * (aho_corasick.c:98,102,121,122,128,129,175,217,220,232,236,250,261,264,
 * 299,336,412,518,580).  It exists so that the UNMODIFIED reference source can be compiled into
 * oracle/_ref/ as the correctness oracle and CPU baseline.  Nothing in the product links it.
 */
#ifndef ORACLE_MAP_STANDIN_H
#define ORACLE_MAP_STANDIN_H

#include <stddef.h>

typedef struct map map;

/* Extracts the key of an element; when 0 the element pointer itself is the key (aho_corasick.c:98). */
typedef const void *(*map_key_extractor) (void *data);
/* Three-way comparison of two keys, with a user argument (aho_corasick.h:33 CMP_TYPE is compatible). */
typedef int (*map_key_comparator) (const void *key_a, const void *key_b, const void *cmp_arg);
/* Operator applied to an element.  Returns non-zero to continue, 0 to stop.  Setting *remove to 1
 * asks the container to unlink the element (the operator may already have freed it, aho_corasick.c:114). */
typedef int (*map_operator) (void *data, void *op_arg, int *remove);
/* Selector: non-zero if the element is to be operated on; 0 selector means "all". */
typedef int (*map_selector) (const void *data, void *sel_arg);

/* memcmp (a, b, *(const size_t *) cmp_arg) (aho_corasick.c:98). */
extern const map_key_comparator MAP_GENERIC_CMP;
/* *(void **) op_arg = data; stop (aho_corasick.c:175,299). */
extern const map_operator MAP_GET_ONE;
/* remove the element and continue (aho_corasick.c:128,217). */
extern const map_operator MAP_REMOVE_ALL;

map *map_create (map_key_extractor get_key, map_key_comparator cmp, const void *cmp_arg, int unique);
int map_destroy (map *m);                      /* non-zero on success; the map must be empty */
int map_insert_data (map *m, void *data);      /* non-zero on success */
size_t map_find_key (map *m, const void *key, map_operator op, void *op_arg, map_selector sel, void *sel_arg);
size_t map_traverse (map *m, map_operator op, void *op_arg, map_selector sel, void *sel_arg);
size_t map_size (const map *m);

#endif
