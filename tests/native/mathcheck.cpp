// tests/native/mathcheck.cpp -- TEST INFRASTRUCTURE. Exposes the product's host+device arithmetic
// header (b2j_math.h) and the host LUT builder to the CPU test-suite, so the exact integer
// formulas the kernels use are checked against the oracle before any GPU time is spent. Nothing
// here is a decode path: the functions are the same inline code the kernels compile.
#include <stdint.h>
#include <string.h>
#include <vector>

#include "b2j_internal.h"
#include "b2j_math.h"

extern "C" {

void chk_idct(int32_t *v64) { b2j::idct_8x8(v64); }

uint32_t chk_csc_pixel(int32_t y, int32_t u, int32_t v) { return b2j::csc_pixel(y, u, v); }

int32_t chk_extend(uint32_t top, int nbits) { return b2j::extend_top(top, nbits); }

// Exhaustive comparison of csc_pixel() with a callback-free restatement of the reference's
// double formula (decoder.cpp:367-370) over Y,U,V in [-256,255]^3. Returns the mismatch count and
// the first mismatching triple.
static inline uint32_t ref_pixel(int Y, int U, int V)
{
    const int r = (int)(Y + 1.402 * V + 128), g = (int)(Y - 0.34414 * U - 0.71414 * V + 128), b = (int)(Y + 1.772 * U + 128);
    return (b2j::clamp255(r) << 16) | (b2j::clamp255(g) << 8) | b2j::clamp255(b);
}
long chk_csc_exhaustive(int32_t first_bad[3])
{
    long bad = 0;
    for (int y = -256; y < 256; y++)
        for (int u = -256; u < 256; u++)
            for (int v = -256; v < 256; v++)
                if (b2j::csc_pixel(y, u, v) != ref_pixel(y, u, v) || b2j::csc_pixel_biased(y + 256, u + 256, v + 256) != ref_pixel(y, u, v))
                {
                    if (!bad) { first_bad[0] = y; first_bad[1] = u; first_bad[2] = v; }
                    bad++;
                }
    return bad;
}

// Builds the two-level LUT of one table and decodes `peek` (32 bits, MSB first) with it, the way
// lut_lookup() in kernels.cu does. Returns len | size<<8 | run<<16 (0 = no codeword), -1 on build failure.
int chk_lut_decode(const uint8_t counts[16], const uint8_t *symbols, int is_dc, uint32_t peek, int *n_entries)
{
    std::vector<uint16_t> t;
    if (!b2j::build_huff_lut(counts, symbols, is_dc != 0, t, b2j::kLutMaxEntries)) return -1;
    if (n_entries) *n_entries = (int)t.size();
    const uint32_t K = is_dc ? b2j::kLutBitsDc : b2j::kLutBits;
    uint32_t e = t[peek >> (32 - K)];
    if ((e & 63u) < 32u)
    {
        if (e == 0) return 0;
        const uint32_t nb = e & 63u, off = (e >> 6) * b2j::kLutSubAlign;
        e = t[(1u << K) + off + ((peek << K) >> (32u - nb))];
        if (e == 0) return 0;
    }
    const int len = (int)(e & 31u);
    int size, run;
    if (is_dc) { size = (int)((e >> 6) & 31u); run = 0; }
    else
    {
        size = (int)((e >> 6) & 15u); run = (int)(e >> 10);
        if (run == (int)b2j::kRunEob) run = 0;   // the end-of-block symbol 0x00
    }
    return len | (size << 8) | (run << 16);
}

int chk_zigzag(int i)
{
    static const uint8_t zz[64] = B2J_ZIGZAG_TABLE;
    return zz[i];
}

} // extern "C"
