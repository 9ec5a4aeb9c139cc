/* oracle/port/orc_merkle.c -- Merkle tree restatement (TEST INFRASTRUCTURE).
 * merkle/merkle_tree.h:51-58 (hash2), :100-151 (MerkleTree), :62-98. */
#include <stdlib.h>
#include <string.h>

#include "orc.h"

/* merkle_tree.h:109-114: heap layout, leaves at [n,2n), node i = H(2i || 2i+1) */
void merkle_build(size_t n, uint8_t* nodes) {
  for (size_t i = n; i-- > 1;) {
    sha256 s;
    sha256_init(&s);
    sha256_update(&s, nodes + 32 * (2 * i), 32);
    sha256_update(&s, nodes + 32 * (2 * i + 1), 32);
    sha256_final(&s, nodes + 32 * i);
  }
}
/* merkle_tree.h:62-70 */
size_t merkle_tree_len(size_t n) {
  size_t r = 1;
  size_t pos = (n - 1);
  for (pos += n; pos > 1; pos >>= 1) ++r;
  return r;
}
/* merkle_tree.h:75-98 + :122-143 compressed multi-leaf proof */
size_t merkle_open(size_t n, const uint8_t* nodes, const size_t* pos, size_t np, uint8_t* path) {
  uint8_t* tree = (uint8_t*)calloc(2 * n, 1);
  for (size_t ip = 0; ip < np; ++ip) tree[pos[ip] + n] = 1;
  for (size_t i = n; i-- > 1;) tree[i] = (uint8_t)(tree[2 * i] || tree[2 * i + 1]);
  size_t sz = 0;
  for (size_t i = n; i-- > 1;) {
    if (tree[i]) {
      size_t child = 2 * i;
      if (tree[child]) child = 2 * i + 1;
      if (!tree[child]) {
        memcpy(path + 32 * sz, nodes + 32 * child, 32);
        ++sz;
      }
    }
  }
  free(tree);
  return sz;
}
