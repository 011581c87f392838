// trig.cuh -- glibc's sinf / cosf, restated so that the DEVICE can produce the bits the reference's host code gets.
//
// The lattice axis tables of FastMatch are cosf / sinf of the candidate angles computed by the host libm
// (Subsystem_1/main.c:434-435); CUDA's own sinf / cosf differ from glibc's in the last bit for many inputs, which is
// why every other path of this library takes those tables from the host.  A per-scan loop that stays on the device
// (b200slam_scan_chain_*) needs them on the device.
//
// glibc >= 2.28 (this image: 2.39) computes both functions in DOUBLE precision from one table
// (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h: reduce_fast + sinf_poly; x86-64 selects a build with or
// without FMA contraction at load time):
//     |y| < pi/4 :  polynomial in x = (double) y directly
//     |y| < 120  :  n = round(x * 2/pi) (fixed point: hpi_inv is scaled by 2^24), x -= n * pi/2, the quadrant picks
//                   the sine or the cosine polynomial and the sign
// The constants below are the 14 doubles of glibc's __sincosf_table.  tools/trig_check.c compares this restatement
// (host build, with and without FMA) with the running libm on EVERY float of |y| < 120, both functions:
// 0 mismatches of 4.5e9 with FMA on this image's CPU; the two variants differ from each other on 34 inputs, all of
// |y| > 17.2.  The device version therefore serves |y| <= 16 (either host variant gives these bits) and the caller
// falls back to the host for anything beyond; b200slam_scan_chain_begin additionally checks a sample against the
// running libm before it trusts the device tables.
#pragma once

#ifdef __CUDACC__
#define B200SLAM_HD __host__ __device__ __forceinline__
#else
#define B200SLAM_HD static inline
#endif

#include <math.h>
#include <stdint.h>
#include <string.h>

namespace glibc_trig {

constexpr float MAX_ABS = 16.0f;          // the restatement is variant-independent below 17.2

B200SLAM_HD double tfma(double a, double b, double c)
{
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
B200SLAM_HD double tmul(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}

// n even: sine polynomial, n odd: cosine polynomial (sincosf.h: sinf_poly).  neg: the second table (quadrants 2, 3):
// its cosine coefficients are negated, the sine ones are not.
B200SLAM_HD float poly(double x, double x2, bool neg, int n)
{
    const double C0 = 1.0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
                 C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        const double x3 = tmul(x, x2);
        const double s1 = tfma(x2, S3, S2);
        const double x7 = tmul(x3, x2);
        const double s = tfma(x3, S1, x);
        return (float)tfma(x7, s1, s);
    }
    const double sg = neg ? -1.0 : 1.0;
    const double x4 = tmul(x2, x2);
    const double c2 = tfma(x2, sg * C4, sg * C3);
    const double c1 = tfma(x2, sg * C1, sg * C0);
    const double x6 = tmul(x4, x2);
    const double c = tfma(x4, sg * C2, c1);
    return (float)tfma(x6, c2, c);
}

// is_cos = 0: sinf(y), 1: cosf(y).  Valid (== glibc) for |y| <= MAX_ABS.
B200SLAM_HD float sincos(float y, int is_cos)
{
    uint32_t u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint(y);
#else
    memcpy(&u, &y, 4);
#endif
    const uint32_t top = (u >> 20) & 0x7ffu;                   // abstop12
    double x = (double)y;
    if (top < 0x3f4u) {                                        // |y| < pi/4
        if (top < 0x398u) return is_cos ? 1.0f : y;            // |y| < 2^-12
        return poly(x, tmul(x, x), false, is_cos);
    }
    const double HPI_INV = 0x1.45f306dc9c883p+23, HPI = 0x1.921fb54442d18p+0;
    const double r = tmul(x, HPI_INV);
    const int n = ((int32_t)r + 0x800000) >> 24;               // reduce_fast: (int32_t) truncates, |r| < 2^31 for |y| < 120
    x = tfma(-(double)n, HPI, x);
    const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;      // sign[n & 3] = {1, -1, -1, 1}
    return poly(tmul(x, s), tmul(x, x), (n & 2) != 0, is_cos ? n ^ 1 : n);
}

}  // namespace glibc_trig
