// Force-included before the reference's db_query_4.cpp: gcc >= 8 ships _mm256_set_m128i,
// which collides with the static helper the reference defines at simd_scan.hpp:120.
#include <immintrin.h>
#include <x86intrin.h>
#define _mm256_set_m128i qadc_ref_mm256_set_m128i
