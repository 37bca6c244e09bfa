// decoder_b2j.cpp -- the reference's decoder API (decoder.h:4-7) on top of the B200 C ABI.
//
// Drop-in for the reference's decoder.cpp + oclDCT8x8.cpp + cpuIDCT8x8.cpp: parser.cpp's
// load_jpg() calls the four functions below in the same order, with the same arguments and the
// same `true` = success convention (parser.cpp:365-397). Everything underneath runs on the GPU
// through include/b2j.h; there is no CPU decode here -- without a device every call returns false.
//
//   is_supported_file   decoder.cpp:18-70   same checks; B2J_GATE=extended additionally admits 4:2:2/4:4:0
//   decode_init         decoder.cpp:161-219 geometry + mcu_data allocation; GPU context instead of clidct_create/build
//   decode_huffman_data decoder.cpp:262-365 reads the entropy-coded bytes from `strm`, uploads them, runs the
//                                           device pipeline, fills jpg.mcu_data with the reference's int32 tap,
//                                           leaves `strm` positioned on the EOI marker like read_more_data() does
//   decode_mcu_data     decoder.cpp:397-523 downloads the BGRA image and writes the 32-bit top-down BMP
//                                           (bmp_create, decoder.cpp:372-395) to the same fixed path
#ifdef B2J_USE_REFERENCE_HEADERS
#include "stdafx.h"
#include "macro.h"
#include "jpeg.h"
#include "decoder.h"
#include "idct.h"
#else
#include "refabi.h"
#endif

#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <vector>

#include "../../../include/b2j.h"

static_assert(sizeof(SOF0) == 15 && sizeof(SOS) == 10 && sizeof(APP0) == 16 && sizeof(DRI) == 2, "packed JPEG structs (jpeg.h)");
static_assert(sizeof(coef_t) == 4, "coef_t is int (jpeg.h:57)");

namespace {

b2j_ctx *g_ctx = nullptr;
b2j_batch *g_batch = nullptr;
b2j_image_desc g_desc;
std::vector<uint8_t> g_file;   // [0, scan_size) = the bytes decode_huffman_data() read from the stream

const char *out_path()
{
    const char *p = getenv("B2J_OUTPUT_BMP");
    return p ? p : "m:\\output.bmp";   // decoder.cpp:420
}

int gate_mode()
{
    const char *g = getenv("B2J_GATE");
    return (g && (g[0] == 'e' || g[0] == '1')) ? B2J_GATE_EXTENDED : B2J_GATE_REFERENCE;
}

bool fail(const char *what, int rc)
{
    printf("[X] %s: %s (%s)\n", what, b2j_strerror(rc), b2j_last_error());
    return false;
}

void drop_batch()
{
    if (g_batch) { b2j_batch_destroy(g_batch); g_batch = nullptr; }
}

// JPG_DATA -> b2j_image_desc. Returns false on table ids the C ABI does not carry (Th > 3).
bool to_desc(const JPG_DATA &jpg, b2j_image_desc &d)
{
    memset(&d, 0, sizeof(d));
    d.width = jpg.frame_info.img_width;
    d.height = jpg.frame_info.img_height;
    d.restart_interval = jpg.dri_info.restart_interval;
    for (int c = 0; c < 3; c++)
    {
        d.sampling[c] = jpg.frame_info.channel_info[c].sampling_factor;
        d.quant_id[c] = jpg.frame_info.channel_info[c].quant_tbl_id;
        d.huff_id[c] = jpg.scan_info.channel_data[c].huff_tbl_id;
        if ((d.huff_id[c] >> 4) > 3 || (d.huff_id[c] & 0xF) > 3 || d.quant_id[c] > 3) return false;
        d.blks_per_mcu[c] = jpg.blks_per_mcu[c];
    }
    for (int t = 0; t < 4; t++)
    {
        if (!jpg.quantization_table[t]) continue;
        d.quant_present[t] = 1;
        for (int k = 0; k < 64; k++) d.quant[t][k] = (uint16_t)jpg.quantization_table[t][k];
    }
    for (int tc = 0; tc < 2; tc++)
        for (int th = 0; th < 4; th++)
        {
            const HUFFMAN_TABLE *t = jpg.huffman_table[(tc << 4) | th];
            if (!t) continue;
            const int slot = tc * 4 + th;
            d.huff_present[slot] = 1;
            for (int n = 0; n < t->num_codeword; n++)
            {
                const size_t l = strlen(t->codeword[n]);   // the reference keeps codes as ASCII strings
                if (l < 1 || l > 16) return false;
                d.huff_counts[slot][l - 1]++;
                d.huff_symbols[slot][n] = t->value[n];
            }
        }
    d.mcu_width = jpg.mcu_width; d.mcu_height = jpg.mcu_height;
    d.mcu_count_w = jpg.mcu_count_w; d.mcu_count_h = jpg.mcu_count_h; d.mcu_count = jpg.mcu_count;
    d.tot_blks_per_mcu = jpg.tot_blks_per_mcu;
    d.blk_count = jpg.blk_count;
    d.color_space = (uint8_t)jpg.color_space;
    return true;
}

} // namespace

#ifdef B2J_USE_REFERENCE_HEADERS
// The two initialisers the reference's main() calls (main.cpp:21-22, idct.h:4,9). The clip table of
// cpuIDCT8x8.cpp lives in the kernel now; device discovery is b2j_create().
void Initialize_Fast_IDCT() {}
int Initialize_OpenCL_IDCT()
{
    puts("[ ] Initializing CUDA environment (b2j)");
    if (b2j_device_count() <= 0) { puts("[X] no CUDA device"); return 1; }
    return 0;
}
#endif

bool is_supported_file(const JPG_DATA &jpg)
{
    if (jpg.frame_info.bit_depth != 8) { puts("[X] unsupported bit depth"); return false; }
    if (jpg.frame_info.num_channels != 3 || jpg.scan_info.num_channels != 3) { puts("[X] unsupported number of components"); return false; }
    if (jpg.frame_info.img_width <= 0 || jpg.frame_info.img_height <= 0) { puts("[X] invalid dimensions"); return false; }
    for (int c = 0; c < 3; c++)
    {
        const int q = jpg.frame_info.channel_info[c].quant_tbl_id;
        if (q > 3 || !jpg.quantization_table[q]) { puts("[X] corrupted file. missing quantization table."); return false; }
        const int td = jpg.scan_info.channel_data[c].huff_tbl_id >> 4, ta = jpg.scan_info.channel_data[c].huff_tbl_id & 0xF;
        if (!jpg.huffman_table[td]) { puts("[X] corrupted file. missing huffman table for DC component."); return false; }
        if (!jpg.huffman_table[ta | 0x10]) { puts("[X] corrupted file. missing huffman table for AC component."); return false; }
    }
    const int y = jpg.frame_info.channel_info[0].sampling_factor;
    const bool chroma_1x1 = jpg.frame_info.channel_info[1].sampling_factor == 0x11 && jpg.frame_info.channel_info[2].sampling_factor == 0x11;
    if (chroma_1x1 && (y == 0x22 || y == 0x11)) return true;
    if (chroma_1x1 && gate_mode() == B2J_GATE_EXTENDED && (y == 0x21 || y == 0x12)) return true;
    puts("[X] sorry, currently only supports 8-bit YUV 4:1:1 or 4:4:4 format");
    return false;
}

bool decode_init(JPG_DATA &jpg)
{
    int mh = 0, mv = 0;
    jpg.tot_blks_per_mcu = 0;
    for (int c = 0; c < jpg.frame_info.num_channels; c++)
    {
        const int h = jpg.frame_info.channel_info[c].sampling_factor >> 4, v = jpg.frame_info.channel_info[c].sampling_factor & 0xF;
        if (h > mh) mh = h;
        if (v > mv) mv = v;
        jpg.blks_per_mcu[c] = h * v;
        jpg.tot_blks_per_mcu += h * v;
    }
    jpg.mcu_width = 8 * mh;
    jpg.mcu_height = 8 * mv;
    const int y = jpg.frame_info.channel_info[0].sampling_factor;
    jpg.color_space = y == 0x22 ? YUV411 : (y == 0x11 ? YUV444 : Other);
    jpg.mcu_count_w = (jpg.frame_info.img_width - 1) / jpg.mcu_width + 1;
    jpg.mcu_count_h = (jpg.frame_info.img_height - 1) / jpg.mcu_height + 1;
    jpg.mcu_count = jpg.mcu_count_w * jpg.mcu_count_h;
    jpg.blk_count = jpg.tot_blks_per_mcu * jpg.mcu_count;
    jpg.mcu_data = new coef_t[(size_t)jpg.blk_count][64];
    printf("[ ] %d * %d = %d MCUs in total, %d blocks per MCU.\n", jpg.mcu_count_w, jpg.mcu_count_h, jpg.mcu_count, jpg.tot_blks_per_mcu);
    printf("[ ] %d blocks in total.\n", jpg.blk_count);
    printf("[ ] MCU Size: %u px * %u px\n", jpg.mcu_width, jpg.mcu_height);
    if (!g_ctx)
    {
        puts("[C] b2j_create()");
        const char *dev = getenv("B2J_DEVICE");
        const int rc = b2j_create(dev ? atoi(dev) : 0, &g_ctx);
        if (rc != B2J_OK) return fail("b2j_create", rc);
    }
    return true;
}

bool decode_huffman_data(const JPG_DATA &jpg, FILE *const fp)
{
    if (!g_ctx) return false;
    drop_batch();
    if (!to_desc(jpg, g_desc)) { puts("[X] table ids outside the supported range"); return false; }
    // the stream sits on the first entropy-coded byte (parser.cpp:384); take everything to EOF
    const long start = ftell(fp);
    g_file.clear();
    uint8_t buf[65536];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), fp)) > 0) g_file.insert(g_file.end(), buf, buf + got);
    g_desc.scan_offset = 0;
    g_desc.scan_size = g_file.size();
    const uint8_t *files[1] = {g_file.data()};
    const size_t lens[1] = {g_file.size()};
    int rc = b2j_batch_create(g_ctx, 1, &g_desc, files, lens, &g_batch);
    if (rc != B2J_OK) return fail("b2j_batch_create", rc);
    puts("[C] b2j_batch_upload() + b2j_batch_decode()");
    const clock_t t0 = clock();
    if ((rc = b2j_batch_upload(g_batch, nullptr)) != B2J_OK) return fail("b2j_batch_upload", rc);
    b2j_stage_times tm;
    if ((rc = b2j_batch_decode_timed(g_batch, nullptr, &tm)) != B2J_OK) return fail("b2j_batch_decode", rc);
    printf("Time elapsed on the device: pre-pass %.3f ms, huffman %.3f ms, idct+colour %.3f ms (host clock %ld)\n",
           tm.prepass_ms, tm.huffman_ms, tm.idct_ms, (long)(clock() - t0));
    int32_t status = 0;
    if ((rc = b2j_batch_status(g_batch, nullptr, &status)) != B2J_OK) return fail("b2j_batch_status", rc);
    if (status)
    {
        if (status & B2J_ST_RST_MISMATCH) puts("[X] expected RSTn");
        printf("[X] data corrupted. (status 0x%x)\n", status);
        return false;
    }
    // the reference's tap: int32, natural order, dequantised (decoder.cpp:338-342)
    if ((rc = b2j_batch_read_coefs(g_batch, nullptr, 0, &jpg.mcu_data[0][0])) != B2J_OK) return fail("b2j_batch_read_coefs", rc);
    // leave the stream on the marker that ends the scan, as read_more_data() does (decoder.cpp:112-115)
    size_t i = 0;
    const size_t n = g_file.size();
    for (; i + 1 < n; i++)
        if (g_file[i] == 0xFF && g_file[i + 1] != 0x00 && g_file[i + 1] != 0xFF && (g_file[i + 1] & 0xF8) != 0xD0) break;
    fseek(fp, start + (long)(i + 1 < n ? i : n), SEEK_SET);
    return true;
}

bool decode_mcu_data(const JPG_DATA &jpg, FILE *const)
{
    if (!g_batch) return false;
    const size_t image_size = (size_t)jpg.frame_info.img_width * jpg.frame_info.img_height * 4;
    std::vector<uint8_t> image(image_size);
    puts("[C] b2j_batch_read_pixels()");
    const int rc = b2j_batch_read_pixels(g_batch, nullptr, 0, image.data());
    if (rc != B2J_OK) { drop_batch(); return fail("b2j_batch_read_pixels", rc); }
    // 54-byte header, 32 bpp, top-down (negative height), then tightly packed BGRA rows (decoder.cpp:372-395)
    FILE *bmp = fopen(out_path(), "wb");
    bool ok = bmp != nullptr;
    if (ok)
    {
        uint8_t h[54];
        memset(h, 0, sizeof(h));
        const uint32_t off = 54, size = off + (uint32_t)image_size;
        const int32_t w = jpg.frame_info.img_width, neg_h = -(int32_t)jpg.frame_info.img_height;
        const uint16_t planes = 1, bpp = 32;
        const uint32_t hdr = 40;
        h[0] = 'B'; h[1] = 'M';
        memcpy(h + 2, &size, 4); memcpy(h + 10, &off, 4); memcpy(h + 14, &hdr, 4);
        memcpy(h + 18, &w, 4); memcpy(h + 22, &neg_h, 4); memcpy(h + 26, &planes, 2); memcpy(h + 28, &bpp, 2);
        ok = fwrite(h, sizeof(h), 1, bmp) == 1 && fwrite(image.data(), image_size, 1, bmp) == 1;
        if (!ok) puts("[X] Write file error");
        fclose(bmp);
    }
    drop_batch();
    return ok;
}
