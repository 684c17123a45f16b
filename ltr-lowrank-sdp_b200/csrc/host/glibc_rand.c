/*
 * glibc_rand.c -- the value stream of glibc's srand()/rand(), restated without the library's lock.
 *
 * The reference seeds its starting point with srand(925) and fills every factor with
 * rand()/RAND_MAX - rand()/RAND_MAX (lorads_solver.c:529-539, 625-669, 864-906), so iteration-count
 * parity needs that exact stream.  glibc's rand() is the TYPE_3 additive-feedback generator of
 * random_r (degree 31, separation 3): the state is seeded by the Lehmer recurrence
 * x <- 16807 x mod (2^31 - 1), the first 310 outputs are discarded, and every call does
 * state[f] += state[r] (mod 2^32) and returns state[f] >> 1.  Calling the library costs a lock per
 * number (10-20 ns); at n = 1e7, rank 8 the three factors need 4.8e8 numbers, i.e. seconds of start-up
 * for nothing.  tests/test_host_cpu.py checks this stream against the C library's, number for number.
 */
#include <stdint.h>

#include "lorads_host.h"

static uint32_t g_state[31];
static int g_f = 3, g_r = 0;

void lh_srand(unsigned int seed)
{
    int32_t word = seed == 0 ? 1 : (int32_t)seed;
    g_state[0] = (uint32_t)word;
    for (int i = 1; i < 31; ++i) {
        /* 16807 * word mod 2147483647 without overflow (Schrage) */
        const int32_t hi = word / 127773, lo = word % 127773;
        word = 16807 * lo - 2836 * hi;
        if (word < 0) word += 2147483647;
        g_state[i] = (uint32_t)word;
    }
    g_f = 3;
    g_r = 0;
    for (int i = 0; i < 310; ++i) (void)lh_rand();
}

int lh_rand(void)
{
    const uint32_t v = (g_state[g_f] += g_state[g_r]);
    if (++g_f == 31) g_f = 0;
    if (++g_r == 31) g_r = 0;
    return (int)(v >> 1);
}

/* n elements of rand()/RAND_MAX - rand()/RAND_MAX in memory order (LORADS_RANDOM_rk_MAT, lorads_solver.c:529-539) */
void lh_random_fill(double *a, int64_t n)
{
    uint32_t s[31];
    for (int i = 0; i < 31; ++i) s[i] = g_state[i];
    int f = g_f, r = g_r;
    for (int64_t i = 0; i < n; ++i) {
        uint32_t v = (s[f] += s[r]);
        if (++f == 31) f = 0;
        if (++r == 31) r = 0;
        const double x = (double)(int)(v >> 1) / 2147483647;
        v = (s[f] += s[r]);
        if (++f == 31) f = 0;
        if (++r == 31) r = 0;
        a[i] = x - (double)(int)(v >> 1) / 2147483647;
    }
    for (int i = 0; i < 31; ++i) g_state[i] = s[i];
    g_f = f;
    g_r = r;
}
