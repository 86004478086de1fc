/*
 * oracle/orc_rng.c -- TEST INFRASTRUCTURE ONLY (see orc.h).
 *
 * The reference draws minimal sets from libc rand() through
 * DUtils::Random::RandomInt (Thirdparty/DBoW2/DUtils/Random.cpp:47-50).  The
 * oracle calls the real libc rand() so that its stream IS the reference's
 * stream; the product restates glibc's TYPE_3 generator on its own and is
 * compared with this in tests.
 */
#include <stdlib.h>
#include "orc.h"

/* Random.cpp:33-36 */
void orc_rng_seed(unsigned seed) { srand(seed); }

/* Random.cpp:47-50 */
int orc_random_int(int min, int max)
{
    int d = max - min + 1;
    return (int)(((double)rand() / ((double)RAND_MAX + 1.0)) * d) + min;
}

/* PnPsolver.cpp:125-138 / Sim3Solver.cpp:136-149 / MLPnPsolver.cpp:76-96:
 *   vAvailableIndices = mvAllIndices;
 *   for i < k: randi = RandomInt(0, size-1); idx = avail[randi];
 *              avail[randi] = avail.back(); avail.pop_back();
 * restated literally, including the per-iteration copy. */
void orc_index_table(unsigned seed, int n, int k, int H, uint32_t *out)
{
    uint32_t *avail = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    srand(seed);
    for (int h = 0; h < H; ++h) {
        int size = n;
        for (int i = 0; i < n; ++i) avail[i] = (uint32_t)i;
        for (int i = 0; i < k; ++i) {
            int randi = orc_random_int(0, size - 1);
            out[(size_t)h * k + i] = avail[randi];
            avail[randi] = avail[size - 1];
            --size;
        }
    }
    free(avail);
}
