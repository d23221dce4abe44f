// float_thresholds.h -- comparisons of a float against a double literal without a float -> double conversion.
// Included by surfel.cu; compiled for the host and checked over ALL 2^32 float bit patterns by tests/test_float_thresholds.py.
#pragma once
#ifndef __CUDACC__
#define MSL_HD
#else
#define MSL_HD __host__ __device__
#endif

// The reference compares floats against double literals ((double)x < 0.4 ...).  A float <-> double conversion is the slowest
// arithmetic instruction of the superpixel kernels (F2F: 15.5 per clock per SM, tools/fp64_throughput.cu), and these
// comparisons need none: for a float x and a double c,  (double)x < c  <=>  x < RU(c)  and  (double)x >= c  <=>  x >= RU(c)
// with RU(c) the smallest float >= c;  (double)x > c  <=>  x > RD(c)  and  (double)x <= c  <=>  x <= RD(c)  with RD(c) the
// largest float <= c (every float below RU(c) is below c, RU(c) itself is not; NaN compares false on both sides).
MSL_HD constexpr float f_ulp(float a) {  // ulp of a normal float a > 0
    float u = 1.0f;
    while (u > a) u *= 0.5f;
    while (u * 2.0f <= a) u *= 2.0f;
    return u * 0x1p-23f;
}
MSL_HD constexpr float f_up_pos(double c) { return (double)(float)c >= c ? (float)c : (float)c + f_ulp((float)c); }
MSL_HD constexpr float f_dn_pos(double c) { return (double)(float)c <= c ? (float)c : (float)c - f_ulp((float)c); }  // (c is no power of two here)
MSL_HD constexpr float f_up(double c) { return c >= 0 ? f_up_pos(c) : -f_dn_pos(-c); }
MSL_HD constexpr float f_dn(double c) { return c >= 0 ? f_dn_pos(c) : -f_up_pos(-c); }
static_assert((double)f_up(0.4) >= 0.4 && (double)f_dn(0.4) <= 0.4 && f_up(0.4) - f_dn(0.4) == 0x1p-25f, "0.4 lies between two adjacent floats");
static_assert(f_up(0.4) == 0x1.99999ap-2f && f_dn(0.01) == 0x1.47ae14p-7f && f_up(0.01) == 0x1.47ae16p-7f && f_dn(0.1) == 0x1.999998p-4f, "");
static_assert(f_up(-0.4) == -f_dn(0.4) && f_dn(-0.4) == -f_up(0.4), "");
// (the threshold is a constexpr local: evaluated by the compiler, a literal in the device code; -DMSL_DOUBLE_COMPARES builds the
// reference's own form, for A/B timing -- tools/gpu_r3v.sh)
#ifdef MSL_DOUBLE_COMPARES
#define D_LT(x, c) ((double)(x) < (c))
#define D_GT(x, c) ((double)(x) > (c))
#define D_GE(x, c) ((double)(x) >= (c))
#define D_LE(x, c) ((double)(x) <= (c))
#else
#define D_LT(x, c) ([&] { constexpr float t_ = f_up(c); return (x) < t_; }())
#define D_GT(x, c) ([&] { constexpr float t_ = f_dn(c); return (x) > t_; }())
#define D_GE(x, c) ([&] { constexpr float t_ = f_up(c); return (x) >= t_; }())
#define D_LE(x, c) ([&] { constexpr float t_ = f_dn(c); return (x) <= t_; }())
#endif
