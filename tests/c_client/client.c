/* A plain C99 client of libgq.so: include/gq.h must be valid C (no C++ in the ABI), the library must link from C, and
 * the host-side entry points must work without a CUDA device; the compute entry points must fail with a message. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gq.h"

int main(void) {
  const uint8_t bases[] = {1, 2, 3, 4, 1, 1, 2, 2, 3, 3, 4, 4, 1, 2, 3, 4, 4, 3, /* read 2 */ 2, 2, 2};
  const uint64_t off[] = {0, 18, 18, 21};
  uint64_t words = 0;
  if (gq_packed_words(off, 3, &words) != 0 || words != (21 >> 4) + 3 + 1) return 1; /* (bases >> 4) + reads + 1 */
  uint32_t* packed = (uint32_t*)calloc(words + 1, 4);
  uint32_t word_off[4], len[3];
  if (gq_pack_reads(bases, off, 3, packed, word_off, len, 1) != 0) return 2;
  if (len[0] != 18 || len[1] != 0 || len[2] != 3 || word_off[0] != 0 || word_off[1] != 2 || word_off[2] != 3) return 3;
  if ((packed[0] & 0xFFu) != (0u | 1u << 2 | 2u << 4 | 3u << 6)) return 4; /* A C G T -> codes 0 1 2 3, base j at bits 2j */
  int n_dev = -1;
  gq_device_count(&n_dev);
  if (n_dev <= 0) { /* no GPU here: building an index must refuse, loudly */
    const uint32_t prg[] = {1, 2, 5, 3, 6, 4, 6, 1};
    gq_index* idx = NULL;
    if (gq_index_build(prg, 8, 2, 0, &idx) == 0) return 5;
    if (strstr(gq_last_error(), "CUDA") == NULL) return 6;
    uint32_t sa[9];
    if (gq_suffix_array(prg, 8, 0, sa, NULL) == 0) return 7;
  }
  free(packed);
  printf("c client ok (%d CUDA devices)\n", n_dev);
  return 0;
}
