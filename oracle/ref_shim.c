/*
 * ref_shim.c — wraps the reference's vendored public-domain PRNG sources (included from where they lie under
 * /root/reference/Source/Externals/xoshiro; nothing is copied) so that their file-static state can be seeded.
 * Output: oracle/_ref/libxoshiro_ref.so.  Used only by tests/ to pin the oracle's RNG core, exactly like the
 * reference's own test (Source/Tools/FalcorTest/Tests/Sampling/PseudorandomTests.cpp:40-54,94-155).
 */
#include <stdint.h>

#define next ref_splitmix64_next
#define x ref_splitmix64_state
#include REF_SPLITMIX
#undef x
#undef next
void ref_splitmix64_seed(uint64_t v) { ref_splitmix64_state = v; }

#define next ref_xoshiro128ss_next
#include REF_XOSHIRO
#undef next
void ref_xoshiro128ss_seed(const uint32_t* v) { for (int i = 0; i < 4; i++) s[i] = v[i]; }
