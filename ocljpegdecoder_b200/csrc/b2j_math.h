// b2j_math.h -- integer arithmetic of the decode path, shared by the CUDA kernels and by the
// host-side unit checks (tests compile this header with g++ to compare every function with the
// oracle before GPU time is spent). Everything here is exact integer math; nothing is a CPU
// implementation of the product path.
//
// Reference semantics restated (paths relative to the reference repo):
//   * EXTEND of entropy-coded magnitudes ............ decoder.cpp:72-92
//   * zig-zag order .................................. zigzag.h:15-40
//   * Chen-Wang 8-point IDCT, rows then columns ...... cpuIDCT8x8.cpp:36-127
//   * YCbCr -> BGRA with truncation and clamping ..... decoder.cpp:367-370, macro.h:121-145
#ifndef B2J_MATH_H_INCLUDED
#define B2J_MATH_H_INCLUDED

#include <stdint.h>

#if defined(__CUDACC__)
#define B2J_HD __host__ __device__ __forceinline__
#else
#define B2J_HD static inline
#endif

namespace b2j {

// Natural (row-major) index of the i-th coefficient in scan order (the walk of zigzag.h:15-40
// yields the standard JPEG order).
#define B2J_ZIGZAG_TABLE                                                                       \
    {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,  \
     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,  \
     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63}

// JPEG EXTEND. `top` holds the value bits left-aligned in a 32-bit word (bit 31 = first bit),
// nbits in 0..16. A leading 0 bit means a negative number: v - (2^nbits - 1).
B2J_HD int32_t extend_top(uint32_t top, int nbits)
{
    const uint32_t v = (top >> 1) >> (31 - nbits);             // nbits == 0 -> 0
    const uint32_t neg = (uint32_t)((int32_t)(~top) >> 31);    // all ones when the first bit is 0
    return (int32_t)v - (int32_t)(neg & ((1u << nbits) - 1u));
}

// ---- Chen-Wang constants: 2048*sqrt(2)*cos(k*pi/16) ----
enum { IW1 = 2841, IW2 = 2676, IW3 = 2408, IW5 = 1609, IW6 = 1108, IW7 = 565 };

// One row, in registers. b0..b7 in, results out (cpuIDCT8x8.cpp:36-80; the all-zero-AC early out
// there is arithmetically the same as this general path).
B2J_HD void idct_row(int32_t &b0, int32_t &b1, int32_t &b2, int32_t &b3, int32_t &b4, int32_t &b5,
                     int32_t &b6, int32_t &b7)
{
    int32_t x0 = b0 * 2048 + 128, x1 = b4 * 2048, x2 = b6, x3 = b2, x4 = b1, x5 = b7, x6 = b5, x7 = b3, x8;
    x8 = IW7 * (x4 + x5);
    x4 = x8 + (IW1 - IW7) * x4;
    x5 = x8 - (IW1 + IW7) * x5;
    x8 = IW3 * (x6 + x7);
    x6 = x8 - (IW3 - IW5) * x6;
    x7 = x8 - (IW3 + IW5) * x7;
    x8 = x0 + x1;
    x0 -= x1;
    x1 = IW6 * (x3 + x2);
    x2 = x1 - (IW2 + IW6) * x2;
    x3 = x1 + (IW2 - IW6) * x3;
    x1 = x4 + x6;
    x4 -= x6;
    x6 = x5 + x7;
    x5 -= x7;
    x7 = x8 + x3;
    x8 -= x3;
    x3 = x0 + x2;
    x0 -= x2;
    x2 = (181 * (x4 + x5) + 128) >> 8;
    x4 = (181 * (x4 - x5) + 128) >> 8;
    b0 = (x7 + x1) >> 8;
    b1 = (x3 + x2) >> 8;
    b2 = (x0 + x4) >> 8;
    b3 = (x8 + x6) >> 8;
    b4 = (x8 - x6) >> 8;
    b5 = (x0 - x4) >> 8;
    b6 = (x3 - x2) >> 8;
    b7 = (x7 - x1) >> 8;
}

B2J_HD int32_t iclip(int32_t v) // the iclp table of cpuIDCT8x8.cpp:13-23: [-256, 255]
{
    return v < -256 ? -256 : (v > 255 ? 255 : v);
}

// One column (cpuIDCT8x8.cpp:82-127) up to and including the final >>14, WITHOUT the iclp clip
// (the kernel clips two values at a time after packing them to int16 pairs).
B2J_HD void idct_col_noclip(int32_t &b0, int32_t &b1, int32_t &b2, int32_t &b3, int32_t &b4, int32_t &b5,
                            int32_t &b6, int32_t &b7)
{
    int32_t x0 = b0 * 256 + 8192, x1 = b4 * 256, x2 = b6, x3 = b2, x4 = b1, x5 = b7, x6 = b5, x7 = b3, x8;
    x8 = IW7 * (x4 + x5) + 4;
    x4 = (x8 + (IW1 - IW7) * x4) >> 3;
    x5 = (x8 - (IW1 + IW7) * x5) >> 3;
    x8 = IW3 * (x6 + x7) + 4;
    x6 = (x8 - (IW3 - IW5) * x6) >> 3;
    x7 = (x8 - (IW3 + IW5) * x7) >> 3;
    x8 = x0 + x1;
    x0 -= x1;
    x1 = IW6 * (x3 + x2) + 4;
    x2 = (x1 - (IW2 + IW6) * x2) >> 3;
    x3 = (x1 + (IW2 - IW6) * x3) >> 3;
    x1 = x4 + x6;
    x4 -= x6;
    x6 = x5 + x7;
    x5 -= x7;
    x7 = x8 + x3;
    x8 -= x3;
    x3 = x0 + x2;
    x0 -= x2;
    x2 = (181 * (x4 + x5) + 128) >> 8;
    x4 = (181 * (x4 - x5) + 128) >> 8;
    b0 = (x7 + x1) >> 14;
    b1 = (x3 + x2) >> 14;
    b2 = (x0 + x4) >> 14;
    b3 = (x8 + x6) >> 14;
    b4 = (x8 - x6) >> 14;
    b5 = (x0 - x4) >> 14;
    b6 = (x3 - x2) >> 14;
    b7 = (x7 - x1) >> 14;
}

// One column with the reference's clip to [-256,255].
B2J_HD void idct_col(int32_t &b0, int32_t &b1, int32_t &b2, int32_t &b3, int32_t &b4, int32_t &b5,
                     int32_t &b6, int32_t &b7)
{
    idct_col_noclip(b0, b1, b2, b3, b4, b5, b6, b7);
    b0 = iclip(b0); b1 = iclip(b1); b2 = iclip(b2); b3 = iclip(b3);
    b4 = iclip(b4); b5 = iclip(b5); b6 = iclip(b6); b7 = iclip(b7);
}

// Full 8x8 block held as 64 scalars (a thread's registers on the device).
B2J_HD void idct_8x8(int32_t v[64])
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 8; r++)
        idct_row(v[8 * r + 0], v[8 * r + 1], v[8 * r + 2], v[8 * r + 3], v[8 * r + 4], v[8 * r + 5], v[8 * r + 6], v[8 * r + 7]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < 8; c++)
        idct_col(v[c], v[8 + c], v[16 + c], v[24 + c], v[32 + c], v[40 + c], v[48 + c], v[56 + c]);
}

// ---- colour ----------------------------------------------------------------------------
// The reference evaluates, in IEEE double and left to right,
//     R = (int)(Y + 1.402*V + 128), G = (int)(Y - 0.34414*U - 0.71414*V + 128), B = (int)(Y + 1.772*U + 128)
// then clamps to [0,255]. For Y,U,V in [-256,255] (the IDCT clip range) these are integer
// functions: with c = 128 + floor(k*chroma) the result is clamp(Y + c), because truncation and
// floor agree on non-negative sums and negative sums clamp to 0 either way. The three chroma
// offsets below are exact over the whole input range (tests/test_math.py enumerates all 2^27
// triples against the double formula):
//     r_off(V)   = 128 + floor(701*V/500)        = 128 + ((91881*V) >> 16)
//     b_off(U)   = 128 + floor(443*U/250)        = 128 + ((116130*U + 64) >> 16)
//     g_off(U,V) = 128 + floor(-(34414*U + 71414*V)/100000)
// with ONE exception that comes from double rounding in the reference: U = -200, V = 200 makes
// the chroma part exactly -74, the double sum lands just below the integer for Y >= 188 and the
// reference truncates to one less. g_fix() carries that case.
B2J_HD int32_t csc_r_off(int32_t V) { return 128 + ((91881 * V) >> 16); }
B2J_HD int32_t csc_b_off(int32_t U) { return 128 + ((116130 * U + 64) >> 16); }
B2J_HD int32_t csc_g_off(int32_t U, int32_t V)
{
    // 27,100,000 = 271 * 100000 >= max |34414*U + 71414*V|, so the dividend stays positive and the
    // division is a plain unsigned divide by a constant (mul.hi + shift on the device).
    const uint32_t x = (uint32_t)(27100000 - (34414 * U + 71414 * V));
    return 128 - 271 + (int32_t)(x / 100000u);
}
B2J_HD int32_t csc_g_fix(int32_t Y, int32_t U, int32_t V) { return (U == -200 && V == 200 && Y >= 188) ? 1 : 0; }

// The kernel keeps the IDCT samples biased by +256 (0..511, unsigned 16 bit: the bias rides on the DC
// term of the column pass for free and makes the clip a single packed min). The same offsets for
// biased inputs Yb = Y+256, Ub = U+256, Vb = V+256, such that Yb + off_b == Y + off:
//     r: 128 + floor(91881*(Vb-256) / 65536) - 256, and likewise b and g, constants folded.
B2J_HD int32_t csc_r_off_b(int32_t Vb) { return (91881 * Vb - 91881 * 256 - 128 * 65536) >> 16; }
B2J_HD int32_t csc_b_off_b(int32_t Ub) { return (116130 * Ub + 64 - 116130 * 256 - 128 * 65536) >> 16; }
B2J_HD int32_t csc_g_off_b(int32_t Ub, int32_t Vb)
{
    const uint32_t x = (uint32_t)(27100000 + 256 * (34414 + 71414) - (34414 * Ub + 71414 * Vb));
    return (int32_t)(x / 100000u) - 271 + 128 - 256;
}
B2J_HD int32_t csc_g_fix_b(int32_t Yb, int32_t Ub, int32_t Vb) { return (Ub == 56 && Vb == 456 && Yb >= 444) ? 1 : 0; }

B2J_HD uint32_t clamp255(int32_t n) { return n < 0 ? 0u : (n > 255 ? 255u : (uint32_t)n); }

// One pixel, the scalar form (edges, odd widths, and the host-side exhaustive check).
// Returns the little-endian BGRA word: B | G<<8 | R<<16, alpha 0 (macro.h:141-145).
B2J_HD uint32_t csc_pixel(int32_t Y, int32_t U, int32_t V)
{
    const uint32_t r = clamp255(Y + csc_r_off(V));
    const uint32_t g = clamp255(Y + csc_g_off(U, V) - csc_g_fix(Y, U, V));
    const uint32_t b = clamp255(Y + csc_b_off(U));
    return (r << 16) | (g << 8) | b;
}

// One pixel from biased samples (what the colour phase of the kernel computes).
B2J_HD uint32_t csc_pixel_biased(int32_t Yb, int32_t Ub, int32_t Vb)
{
    const uint32_t r = clamp255(Yb + csc_r_off_b(Vb));
    const uint32_t g = clamp255(Yb + csc_g_off_b(Ub, Vb) - csc_g_fix_b(Yb, Ub, Vb));
    const uint32_t b = clamp255(Yb + csc_b_off_b(Ub));
    return (r << 16) | (g << 8) | b;
}

} // namespace b2j
#endif
