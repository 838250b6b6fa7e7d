/* mb_synth.h — deterministic synthetic genomes of BASELINE.json configs C1..C5.
 *
 * Bench / test infrastructure, NOT part of the product ABI: built into its own libmbsynth.so (host only, no CUDA),
 * so that bench.py's reference arm can generate the same inputs without mapping libmauve_b200.so. */
#ifndef MB_SYNTH_H
#define MB_SYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Deterministic generator (splitmix64-seeded xoshiro256**, seed 0x4D415556 + config).  config 1..5 = BASELINE.json
 * configs C1..C5; scale divides every length (1 = full size).  Sequences are ASCII ACGT in host memory owned by
 * the handle.  Returns 0, or -1 (bad argument) / -2 (out of memory). */
typedef struct mb_synth mb_synth;
int mb_synth_create(int config, uint64_t scale, mb_synth** out);
uint32_t mb_synth_nseq(const mb_synth* s);
uint64_t mb_synth_len(const mb_synth* s, uint32_t i);
const uint8_t* mb_synth_seq(const mb_synth* s, uint32_t i);
void mb_synth_free(mb_synth* s);

#ifdef __cplusplus
}
#endif
#endif
