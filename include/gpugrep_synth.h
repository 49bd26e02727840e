/*
 * libgpugrep_synth.so - the synthetic-text generator of bench.py and the parity tests.
 * Test / bench infrastructure: not part of the drop-in boundary (include/gpugrep.h) and not linked into libgpugrep.so,
 * so that a process which must not load the product (bench.py --impl reference) can still produce the corpus.
 */
#ifndef GPUGREP_SYNTH_H
#define GPUGREP_SYNTH_H

#include <stddef.h>

#if defined(__GNUC__)
#define GPUGREP_SYNTH_API __attribute__((visibility("default")))
#else
#define GPUGREP_SYNTH_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Seeded synthetic syslog-shaped text: fills out[0,size) with complete '\n'-terminated lines (80-250 bytes, ~145 mean) and
 * returns the line count.  `plants`: optional indicator strings (each < 100 bytes), one appended to a line with
 * probability plant_ppm / 1e6. */
GPUGREP_SYNTH_API size_t gpugrep_synth_syslog(unsigned long long seed, char* out, size_t size, const char* const* plants,
                                              unsigned int nplants, unsigned int plant_ppm);

#ifdef __cplusplus
}
#endif
#endif /* GPUGREP_SYNTH_H */
