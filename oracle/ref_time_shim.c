/* TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Deterministic replacement for time(2) that is linked INTO the compiled
 * reference module (oracle/_ref/<variant>/tetris_env*.so) with -Bsymbolic-functions,
 * so the reference's PythonHandle::seed() (PythonHandle.cpp:68-71: `seed1 = time(NULL)`)
 * draws its 16-bit seed from a value the parity harness chooses instead of
 * the wall clock.  No reference source is modified for this.
 */
#include <time.h>

static long long g_oracle_time = 1234567;

__attribute__((visibility("default"))) void oracle_set_time(long long t) { g_oracle_time = t; }
__attribute__((visibility("default"))) long long oracle_get_time(void) { return g_oracle_time; }

__attribute__((visibility("default"))) time_t time(time_t *out) {
    if (out) *out = (time_t)g_oracle_time;
    return (time_t)g_oracle_time;
}
