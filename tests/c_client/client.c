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
  { /* the genotyping step is host code: a two-allele site with 9 reads on allele 1 and 1 read on allele 0 */
    const uint32_t prg[] = {1, 2, 5, 3, 6, 4, 6, 1};
    const uint16_t per_base[] = {1, 9};
    const uint32_t grouped[] = {0, 1, 1, 0, /* site 0: {0} x 1 */ 0, 9, 1, 1 /* {1} x 9 */};
    const double stats[3] = {9.0, 0.0, 0.001};
    double depth[2];
    uint64_t counts[2], bytes = 0;
    char* json;
    if (gq_read_depth_stats_host(prg, 8, per_base, 2, grouped, 8, depth, counts) != 0) return 8;
    if (depth[0] != 9.0 || counts[0] != 0 || counts[1] != 1) return 9;
    if (gq_level_genotype_json(prg, 8, per_base, 2, grouped, 8, stats, 1, "c", 42, 1, NULL, &bytes) != 0 || bytes < 100) return 10;
    json = (char*)malloc(bytes);
    if (gq_level_genotype_json(prg, 8, per_base, 2, grouped, 8, stats, 1, "c", 42, 2, json, &bytes) != 0) return 11;
    if (strstr(json, "\"ALS\":[\"G\",\"T\"]") == NULL || strstr(json, "\"GT\":[[1]]") == NULL) return 12;
    free(json);
    if (gq_level_genotype_json(prg, 6, per_base, 2, grouped, 8, stats, 1, "c", 42, 1, NULL, &bytes) == 0) return 13; /* a site with one allele */
  }
  free(packed);
  printf("c client ok (%d CUDA devices)\n", n_dev);
  return 0;
}
