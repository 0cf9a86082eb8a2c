// atan2f as the reference's platform computes it. imageProjection.cpp:547 calls atan2 on two floats, i.e. glibc's atan2f, which up
// to glibc 2.39 (ROS Noetic ships 2.31) is the fdlibm algorithm: argument reduction to four intervals and an odd polynomial, all in
// IEEE single precision — within 1 ulp, NOT correctly rounded (it differs from a double atan2 narrowed to float on 16 % of the points
// of the shipped lidar sweep). Restated here from the published algorithm so that the column a return lands in is the reference's
// even when its angle sits on a column edge: pure float add / multiply / divide, so the same source gives the same bits on the host
// and on the device (the library is built without FMA contraction). tests/test_oracle_pins.py compiles this header with gcc and
// compares it with the C library's atan2f bit for bit (40 M random inputs when it was written: 0 mismatches).
#pragma once
#include <stdint.h>
#include <string.h>
#ifdef __CUDACC__
#define AT_HD __host__ __device__
#else
#define AT_HD
#endif
AT_HD static inline int32_t at_bits(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
AT_HD static inline float at_float(int32_t i) { float f; memcpy(&f, &i, 4); return f; }

AT_HD static inline float port_atanf(float x) {
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f, 9.0908870101e-02f, -7.6918758452e-02f,
                          6.6610731184e-02f, -5.8335702866e-02f, 4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};
    const float one = 1.0f, huge = 1.0e30f;
    float w, s1, s2, z;
    int32_t ix, hx, id;
    hx = at_bits(x);
    ix = hx & 0x7fffffff;
    if (ix >= 0x4c000000) {                 /* |x| >= 2^25 */
        if (ix > 0x7f800000) return x + x;  /* NaN */
        if (hx > 0) return atanhi[3] + atanlo[3];
        else return -atanhi[3] - atanlo[3];
    }
    if (ix < 0x3ee00000) {                  /* |x| < 0.4375 */
        if (ix < 0x31000000) {              /* |x| < 2^-29 */
            if (huge + x > one) return x;
        }
        id = -1;
    } else {
        x = at_float(ix);                   /* fabsf */
        if (ix < 0x3f980000) {              /* |x| < 1.1875 */
            if (ix < 0x3f300000) { id = 0; x = ((float)2.0 * x - one) / ((float)2.0 + x); }    /* 7/16 <= |x| < 11/16 */
            else { id = 1; x = (x - one) / (x + one); }                                         /* 11/16 <= |x| < 19/16 */
        } else {
            if (ix < 0x401c0000) { id = 2; x = (x - (float)1.5) / (one + (float)1.5 * x); }     /* |x| < 2.4375 */
            else { id = 3; x = -(float)1.0 / x; }                                                /* 2.4375 <= |x| < 2^25 */
        }
    }
    z = x * x;
    w = z * z;
    s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    z = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return (hx < 0) ? -z : z;
}

AT_HD static inline float port_atan2f(float y, float x) {
    const float tiny = 1.0e-30f, zero = 0.0f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    float z;
    int32_t k, m, hx, hy, ix, iy;
    hx = at_bits(x); ix = hx & 0x7fffffff;
    hy = at_bits(y); iy = hy & 0x7fffffff;
    if ((ix > 0x7f800000) || (iy > 0x7f800000)) return x + y;        /* NaN */
    if (hx == 0x3f800000) return port_atanf(y);                       /* x = 1.0 */
    m = ((hy >> 31) & 1) | ((hx >> 30) & 2);                          /* 2*sign(x) + sign(y) */
    if (iy == 0) {
        switch (m) {
        case 0: case 1: return y;
        case 2: return pi + tiny;
        case 3: return -pi - tiny;
        }
    }
    if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) {
            case 0: return pi_o_4 + tiny;
            case 1: return -pi_o_4 - tiny;
            case 2: return (float)3.0 * pi_o_4 + tiny;
            case 3: return (float)-3.0 * pi_o_4 - tiny;
            }
        } else {
            switch (m) {
            case 0: return zero;
            case 1: return -zero;
            case 2: return pi + tiny;
            case 3: return -pi - tiny;
            }
        }
    }
    if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    k = (iy - ix) >> 23;
    if (k > 60) z = pi_o_2 + (float)0.5 * pi_lo;                      /* |y/x| > 2^60 */
    else if (hx < 0 && k < -60) z = 0.0f;                             /* |y|/x < -2^60 */
    else { float q = y / x; z = port_atanf(at_float(at_bits(q) & 0x7fffffff)); }     /* safe to do y/x */
    switch (m) {
    case 0: return z;
    case 1: return at_float(at_bits(z) ^ (int32_t)0x80000000);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}
