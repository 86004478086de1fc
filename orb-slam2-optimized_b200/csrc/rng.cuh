// rng.cuh -- the reference's minimal-set index stream, restated (host + device).
//
// Reference: DUtils::Random::RandomInt (Thirdparty/DBoW2/DUtils/Random.cpp:47-50)
//   int d = max - min + 1;  return int(((double)rand()/((double)RAND_MAX + 1.0)) * d) + min;
// over libc rand(), and the draw-without-replacement idiom of
// src/PnPsolver.cpp:125-138 / src/Sim3Solver.cpp:136-149 / src/MLPnPsolver.cpp:76-96.
//
// glibc's rand() is random_r() on the default TYPE_3 state: an additive feedback
// generator  r[i] = r[i-3] + r[i-31]  (mod 2^32), output r[i] >> 1, seeded by the
// Park-Miller LCG 16807*x mod (2^31-1) and 310 discarded outputs; seed 0 is replaced by 1.
// That published algorithm is restated here as integer arithmetic so that each problem owns
// a private stream (the reference shares one unseeded global stream between threads,
// SURVEY F9/Q8) and so that tables can be produced on the device.  tests/ pin it against
// the real libc rand() and the known answers in SURVEY section 4.
#pragma once
#include <cstdint>

namespace rsac {

struct GlibcRand {
    int32_t r[34];
    int idx;   // next position in the circular history (mod 34)

    __host__ __device__ void seed(uint32_t s)
    {
        if (s == 0) s = 1;
        r[0] = (int32_t)s;
        for (int i = 1; i < 31; ++i) {
            // 16807 * r[i-1] % 2147483647 without overflow (Schrage), as glibc does
            const int32_t hi = r[i - 1] / 127773;
            const int32_t lo = r[i - 1] % 127773;
            int32_t word = 16807 * lo - 2836 * hi;
            if (word < 0) word += 2147483647;
            r[i] = word;
        }
        for (int i = 31; i < 34; ++i) r[i] = r[i - 31];
        idx = 0;   // r[idx] is the oldest entry (i-34); history holds entries i-34 .. i-1
        for (int i = 34; i < 344; ++i) (void)next_raw();
    }

    // advances the recurrence by one and returns the full 32-bit word
    __host__ __device__ uint32_t next_raw()
    {
        // history is circular over 34 slots: slot idx = entry (i-34); entry (i-31) is idx+3, entry (i-3) is idx+31
        const int a = (idx + 3) % 34, b = (idx + 31) % 34;
        const uint32_t v = (uint32_t)r[a] + (uint32_t)r[b];
        r[idx] = (int32_t)v;
        idx = (idx + 1) % 34;
        return v;
    }

    // libc rand(): 31-bit output
    __host__ __device__ int32_t next() { return (int32_t)(next_raw() >> 1); }

    // DUtils::Random::RandomInt (Random.cpp:47-50); RAND_MAX = 2147483647
    __host__ __device__ int random_int(int min, int max)
    {
        const int d = max - min + 1;
        return (int)(((double)next() / ((double)2147483647 + 1.0)) * d) + min;
    }
};

// One RANSAC iteration's draw of k distinct indices from [0, n): the reference copies
// mvAllIndices and swap-removes; only the <= k touched slots are tracked here.
template <int KMAX>
__host__ __device__ inline void draw_minimal_set(GlibcRand& g, int n, int k, uint32_t* out)
{
    int pos[KMAX];      // overwritten positions
    uint32_t val[KMAX]; // value now stored there
    int nov = 0;
    int size = n;
    for (int i = 0; i < k; ++i) {
        const int randi = g.random_int(0, size - 1);
        // idx = avail[randi]
        uint32_t idx = (uint32_t)randi;
        for (int j = nov - 1; j >= 0; --j)
            if (pos[j] == randi) { idx = val[j]; break; }
        out[i] = idx;
        // avail[randi] = avail.back(); pop_back()
        const int last = size - 1;
        uint32_t lastv = (uint32_t)last;
        for (int j = nov - 1; j >= 0; --j)
            if (pos[j] == last) { lastv = val[j]; break; }
        pos[nov] = randi;
        val[nov] = lastv;
        ++nov;
        --size;
    }
}

}  // namespace rsac
