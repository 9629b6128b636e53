/* ORACLE / BENCH INFRASTRUCTURE ONLY — a fast host twin of pygrank_b200.synthetic.rmat_edges_host (the numpy
 * recipe documented there: counter-based splitmix64 bits compared with integer thresholds), so that the CPU
 * baseline legs of bench.py can build RMAT-22 inputs in seconds instead of minutes.  The reference has no
 * generator of its own (its graphs are downloaded, /root/reference/pygrank/benchmarks/download.py:62-72).
 * tests/test_oracle.py requires bit-identical edges to the numpy recipe. */
#include <stdint.h>

static inline uint64_t mix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void oracle_rmat_edges(int scale, int64_t first_edge, int64_t num_edges, uint64_t seed, uint32_t t1, uint32_t t2,
                       uint32_t t3, int32_t *src, int32_t *dst) {
    const uint64_t seed_hash = mix64(seed);
    for (int64_t i = 0; i < num_edges; ++i) {
        const uint64_t e = (uint64_t)(first_edge + i);
        const uint64_t s = mix64(seed_hash ^ (e * 0xD1342543DE82EF95ull));
        uint32_t us = 0, ud = 0;
        uint64_t h = 0;
        for (int lvl = 0; lvl < scale; ++lvl) {
            uint32_t u;
            if ((lvl & 1) == 0) {
                h = mix64(s + (uint64_t)(lvl >> 1) * 0x9E3779B97F4A7C15ull);
                u = (uint32_t)(h >> 32);
            } else {
                u = (uint32_t)(h & 0xFFFFFFFFull);
            }
            const uint32_t sbit = u >= t2;
            const uint32_t dbit = ((u >= t1) && (u < t2)) || (u >= t3);
            us = (us << 1) | sbit;
            ud = (ud << 1) | dbit;
        }
        src[i] = (int32_t)us;
        dst[i] = (int32_t)ud;
    }
}
