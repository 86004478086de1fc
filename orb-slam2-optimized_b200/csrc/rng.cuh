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
#include "common.cuh"

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

#ifdef __CUDACC__
// ---- minimal-set tables from per-problem seeds: one warp per problem ----
// The additive-feedback recurrence r[i] = r[i-31] + r[i-3] is advanced 31 values at a time: with the last
// 31 values w[0..30] on lanes 0..30, the next 31 are y[j] = w[j] + y[j-3] (y[-3..-1] = w[28..30]) -- three
// interleaved running sums, i.e. a stride-3 inclusive scan (4 shuffle steps) plus a carry.  Integer
// arithmetic mod 2^32, so the result is the serial stream bit for bit.  The 310 discarded outputs are
// exactly 10 such batches.  Every RANSAC iteration consumes exactly min_set values (PnPsolver.cpp:125-138),
// so once the raw stream is in shared memory the draws of different iterations are independent: lane <->
// iteration, `draw_from` restates draw_minimal_set over a given slice of the stream.
constexpr int kRngWarps = 4;          // problems per CTA
constexpr int kRngBuf = 32 * 8 + 32;  // raw values buffered per warp: one tile of 32 iterations x <= 8 draws + a partial batch

template <int KMAX>
__device__ inline void draw_from(const uint32_t* raw, int n, int k, uint32_t* out)
{
    int pos[KMAX];
    uint32_t val[KMAX];
    int nov = 0;
    int size = n;
    for (int i = 0; i < k; ++i) {
        // DUtils::Random::RandomInt(0, size-1) over rand() = raw >> 1 (Random.cpp:47-50)
        const int d = size;
        const int randi = (int)(((double)(int32_t)(raw[i] >> 1) / ((double)2147483647 + 1.0)) * d);
        uint32_t idx = (uint32_t)randi;
        for (int j = nov - 1; j >= 0; --j)
            if (pos[j] == randi) { idx = val[j]; break; }
        out[i] = idx;
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

static __global__ void __launch_bounds__(kRngWarps * 32) rng_tables_kernel(const ProblemMeta* metas, int C, uint32_t* tables)
{
    __shared__ uint32_t s_buf[kRngWarps][kRngBuf];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = blockIdx.x * kRngWarps + warp;
    if (p >= C) return;
    const ProblemMeta& m = metas[p];
    const int n = m.n, k = m.min_set, H = m.H;
    if (n < k || H <= 0) return;
    uint32_t* buf = s_buf[warp];

    // seeding (glibc srandom_r, TYPE_3): r[0] = seed, r[i] = 16807 r[i-1] mod (2^31-1), i < 31; r[31..33] = r[0..2].
    // window w = r[3..33] on lanes 0..30
    if (lane == 0) {
        uint32_t sd = m.seed;
        if (sd == 0) sd = 1;
        int32_t r = (int32_t)sd;
        buf[0] = (uint32_t)r;
        for (int i = 1; i < 31; ++i) {
            const int32_t hi = r / 127773, lo = r % 127773;
            int32_t word = 16807 * lo - 2836 * hi;
            if (word < 0) word += 2147483647;
            r = word;
            buf[i] = (uint32_t)r;
        }
    }
    __syncwarp();
    uint32_t w = 0;
    if (lane < 28) w = buf[3 + lane];
    else if (lane < 31) w = buf[lane - 28];
    __syncwarp();

    auto next_batch = [&]() -> uint32_t {       // lanes 0..30 return r[i+31]; w is advanced
        const uint32_t carry = __shfl_sync(0xffffffffu, w, 28 + lane % 3);
        uint32_t v = w;
#pragma unroll
        for (int d = 3; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += t;
        }
        v += carry;
        w = v;
        return v;
    };
    for (int b = 0; b < 10; ++b) (void)next_batch();   // the 310 discarded outputs

    uint32_t* out = tables + m.table_off;
    int have = 0;
    for (int h0 = 0; h0 < H; h0 += 32) {
        const int need = min(32, H - h0) * k;
        while (have < need) {
            const uint32_t y = next_batch();
            if (lane < 31) buf[have + lane] = y;
            have += 31;
        }
        __syncwarp();
        const int h = h0 + lane;
        if (h < H) {
            uint32_t idx[8];
            draw_from<8>(buf + lane * k, n, k, idx);
            uint32_t* o = out + (size_t)h * k;
            for (int i = 0; i < k; ++i) o[i] = idx[i];
        }
        __syncwarp();
        const int rem = have - need;            // < 31: move the unconsumed tail to the front
        uint32_t keep = 0;
        if (lane < rem) keep = buf[need + lane];
        __syncwarp();
        if (lane < rem) buf[lane] = keep;
        have = rem;
        __syncwarp();
    }
}

#endif  // __CUDACC__

}  // namespace rsac
