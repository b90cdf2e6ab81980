/* TEST INFRASTRUCTURE (oracle/): C restatement of retrieval_oracle.synth_rows - the counter-based synthetic
 * corpus generator of SURVEY.md §8d configs 3-5 - so that the CPU baseline legs of bench.py can build a
 * >= 1M-row fp32 sample in about a second instead of a minute of NumPy.  Bit-identical to the NumPy oracle
 * (tests/test_oracle.py::test_native_synth_rows_equal_numpy) and to csrc/rs_ingest.cu:synth_fill_kernel:
 * integer hashing plus one IEEE double sqrt, one division and one multiplication per element, each rounded once.
 * Never linked into or called by the product (rag_serving_system_b200/).
 *
 *   gcc -O2 -fopenmp -shared -fPIC oracle/synth_native.c -o oracle/_build/liboracle_native.so -lm
 *   (oracle/build_oracle.py does exactly this; -ffast-math must NOT be used)
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

static inline uint64_t mix64(uint64_t h, uint64_t seed) {
  uint64_t z = seed + (h + 1ull) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

/* out[n_rows][D] (float32, row-major) = rows row0 .. row0+n_rows-1 of the corpus with this seed */
void oracle_synth_rows(uint64_t seed, int64_t row0, int64_t n_rows, int32_t D, float* out) {
  const int64_t groups = D / 4;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_rows; ++i) {
    const uint64_t hbase = (uint64_t)(row0 + i) * (uint64_t)groups;
    float* dst = out + (size_t)i * (size_t)D;
    uint64_t sumsq = 0;
    for (int64_t g = 0; g < groups; ++g) {
      const uint64_t z = mix64(hbase + (uint64_t)g, seed);
      for (int e = 0; e < 4; ++e) {
        const int64_t v = (int64_t)((z >> (16 * e)) & 0xFFFFull) * 2 - 65535;
        sumsq += (uint64_t)(v * v);
      }
    }
    const double rinv = 1.0 / sqrt((double)sumsq);
    for (int64_t g = 0; g < groups; ++g) {
      const uint64_t z = mix64(hbase + (uint64_t)g, seed);
      for (int e = 0; e < 4; ++e) {
        const int64_t v = (int64_t)((z >> (16 * e)) & 0xFFFFull) * 2 - 65535;
        dst[4 * g + e] = (float)((double)v * rinv);
      }
    }
  }
}
