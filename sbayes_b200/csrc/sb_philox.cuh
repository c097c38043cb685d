// sb_philox.cuh -- the counter-based generator every kernel of the library draws from
#pragma once
#include "sb_platform.cuh"
#include <stdint.h>
namespace sbp {

// -----------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011); stream addressing identical to oracle/philox.py
// -----------------------------------------------------------------------------------------
struct Philox4 { uint32_t w[4]; };

SB_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 o; o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3; return o;
}
SB_HD double u01_double(uint32_t w0, uint32_t w1) {
    return ((double)(w0 >> 5) * 67108864.0 + (double)(w1 >> 6)) / 9007199254740992.0;
}
SB_DEV double u01_float(uint32_t w0) { return (double)(w0 >> 8) / 16777216.0; }


}  // namespace sbp
