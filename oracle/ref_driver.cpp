// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// A thin C-ABI harness around the UNMODIFIED reference CPU path. It is compiled together with
// the reference's own translation units where they lie under /root/reference/src
// (bitstream.cpp huffman.cpp cpuIDCT8x8.cpp decoder.cpp parser.cpp, -DUSE_CPU_ONLY) by
// oracle/build_ref.sh; the only output is oracle/_ref/libjpegref.so (git-ignored).
//
// What it does: re-drives the reference's externally linked read_* / decode_* functions in the
// same order load_jpg() does (reference parser.cpp:272-419) so that the two taps the parity
// tests need become reachable:
//   * the coefficient tap  = JPG_DATA::mcu_data right after decode_huffman_data()
//                            (reference decoder.cpp:262-365), int32[blk_count][64]
//   * the pixel tap        = the body of the BMP decode_mcu_data() writes
//                            (reference decoder.cpp:397-523), BGRA, tight pitch
// The reference keeps JPG_DATA on load_jpg()'s stack, so load_jpg() itself cannot expose them.
// No reference code is copied here: the marker loop below is a re-statement of the control flow
// of load_jpg() that calls the reference's own segment readers.
#include "stdafx.h"
#include "macro.h"
#include "jpeg.h"
#include "decoder.h"
#include "idct.h"

#include <unistd.h>
#include <fcntl.h>
#include <time.h>
#include <new>

// external-linkage functions of reference parser.cpp (no header declares them)
bool read_soi(JPG_DATA &jpg, FILE * const strm);
bool read_dqt(JPG_DATA &jpg, FILE * const strm, size_t len);
bool read_sof(JPG_DATA &jpg, FILE * const strm, size_t len);
bool read_sos(JPG_DATA &jpg, FILE * const strm, size_t len);
bool read_dri(JPG_DATA &jpg, FILE * const strm, size_t len);
bool read_dht(JPG_DATA &jpg, FILE * const strm, size_t len);

namespace {

struct RefHandle
{
    JPG_DATA jpg;
    FILE *fp;
    bool huffman_done;
    bool pixels_done;
};

double now_s()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

// The reference prints progress with printf/puts on every call; keep test logs readable.
struct Silence
{
    int saved;
    explicit Silence(bool on) : saved(-1)
    {
        if (!on) return;
        fflush(stdout);
        saved = dup(1);
        int nul = open("/dev/null", O_WRONLY);
        if (nul >= 0) { dup2(nul, 1); close(nul); }
    }
    ~Silence()
    {
        if (saved < 0) return;
        fflush(stdout);
        dup2(saved, 1);
        close(saved);
    }
};

bool g_quiet = true;
bool g_idct_ready = false;

uint16_t be16(const uint8_t b[2]) { return (uint16_t)((b[0] << 8) | b[1]); }

} // namespace

extern "C" {

struct ref_info_t
{
    int32_t width, height;
    int32_t mcu_width, mcu_height;
    int32_t mcu_count_w, mcu_count_h;
    int32_t blks_per_mcu[3];
    int32_t tot_blks_per_mcu;
    int32_t blk_count;
    int32_t restart_interval;
    int32_t sampling[3];
    int64_t scan_offset; // file offset of the first entropy-coded byte
};

void ref_set_quiet(int quiet) { g_quiet = quiet != 0; }

// flags bit0: skip is_supported_file() (needed for 4:2:2, which the reference gate at
//             decoder.cpp:58-69 rejects although its CPU decode loops are generic).
// Returns NULL when the reference would have refused or failed to reach the scan.
void *ref_open(const uint8_t *file, size_t len, int flags, ref_info_t *info)
{
    Silence quiet(g_quiet);
    if (!g_idct_ready) { Initialize_Fast_IDCT(); g_idct_ready = true; }

    RefHandle *h = new (std::nothrow) RefHandle;
    if (!h) return nullptr;
    memset(&h->jpg, 0, sizeof(h->jpg));
    h->huffman_done = h->pixels_done = false;
    h->fp = fmemopen(const_cast<uint8_t *>(file), len, "rb");
    if (!h->fp) { delete h; return nullptr; }

    bool ready = false;
    uint8_t tag[2] = {0, 0};
    uint8_t lenb[2];
    do
    {
        if (!read_soi(h->jpg, h->fp)) break;
        // APPn segments directly after SOI are skipped (load_jpg, parser.cpp:295-322)
        tag[1] = 0;
        while (1 == fread(tag, 2, 1, h->fp) && tag[1] >= 0xE0 && tag[1] <= 0xEF)
        {
            if (1 != fread(lenb, 2, 1, h->fp)) break;
            fseek(h->fp, (long)be16(lenb) - 2, SEEK_CUR);
            tag[1] = 0;
        }
        // table / frame / scan segments (load_jpg, parser.cpp:323-414)
        bool stop = false;
        while (!stop && tag[1] != 0)
        {
            if (1 != fread(lenb, 2, 1, h->fp)) break;
            const size_t seglen = (uint16_t)(be16(lenb) - 2);
            switch (tag[1])
            {
            case 0xDB: if (!read_dqt(h->jpg, h->fp, seglen)) stop = true; break;
            case 0xC0: if (!read_sof(h->jpg, h->fp, seglen)) stop = true; break;
            case 0xC4: if (!read_dht(h->jpg, h->fp, seglen)) stop = true; break;
            case 0xDD: if (!read_dri(h->jpg, h->fp, seglen)) stop = true; break;
            case 0xDA:
                stop = true;
                if (!read_sos(h->jpg, h->fp, seglen)) break;
                if (!(flags & 1) && !is_supported_file(h->jpg)) break;
                if (!decode_init(h->jpg)) break;
                ready = true;
                break;
            default: // SOF1-3, EOI, COM, anything else: the reference stops parsing here
                stop = true;
                break;
            }
            if (stop) break;
            if (1 != fread(tag, 2, 1, h->fp)) break;
        }
    } while (0);

    if (!ready)
    {
        fclose(h->fp);
        delete h;
        return nullptr;
    }
    if (info)
    {
        const JPG_DATA &j = h->jpg;
        info->width = j.frame_info.img_width;
        info->height = j.frame_info.img_height;
        info->mcu_width = j.mcu_width;
        info->mcu_height = j.mcu_height;
        info->mcu_count_w = j.mcu_count_w;
        info->mcu_count_h = j.mcu_count_h;
        for (int i = 0; i < 3; i++)
        {
            info->blks_per_mcu[i] = j.blks_per_mcu[i];
            info->sampling[i] = j.frame_info.channel_info[i].sampling_factor;
        }
        info->tot_blks_per_mcu = j.tot_blks_per_mcu;
        info->blk_count = j.blk_count;
        info->restart_interval = j.dri_info.restart_interval;
        info->scan_offset = ftell(h->fp);
    }
    return h;
}

// Runs the reference's decode_huffman_data() and copies the coefficient tap
// (int32[blk_count][64], natural order, dequantised). Returns 1 ok, 0 reference failed.
int ref_huffman(void *handle, int32_t *coef_out, double *seconds)
{
    RefHandle *h = (RefHandle *)handle;
    if (!h || h->huffman_done) return 0;
    Silence quiet(g_quiet);
    const double t0 = now_s();
    const bool ok = decode_huffman_data(h->jpg, h->fp);
    const double t1 = now_s();
    h->huffman_done = true;
    if (seconds) *seconds = t1 - t0;
    if (ok && coef_out)
        memcpy(coef_out, h->jpg.mcu_data, sizeof(int32_t) * 64 * (size_t)h->jpg.blk_count);
    return ok ? 1 : 0;
}

// Runs the reference's decode_mcu_data() (Fast_IDCT + upsample + YUV_to_RGB32 + BMP write) with
// `workdir` as the current directory (the reference writes the fixed relative name
// "m:\output.bmp"), then reads the W*H*4 pixel bytes back. Returns 1 ok, 0 failed.
int ref_pixels(void *handle, const char *workdir, uint8_t *bgra_out, double *seconds)
{
    RefHandle *h = (RefHandle *)handle;
    if (!h || !h->huffman_done || h->pixels_done) return 0;
    char old_cwd[4096];
    if (!getcwd(old_cwd, sizeof(old_cwd))) return 0;
    if (chdir(workdir) != 0) return 0;
    bool ok;
    double t0, t1;
    {
        Silence quiet(g_quiet);
        t0 = now_s();
        ok = decode_mcu_data(h->jpg, h->fp);
        t1 = now_s();
    }
    h->pixels_done = true;
    if (seconds) *seconds = t1 - t0;
    if (ok && bgra_out)
    {
        FILE *bmp = fopen("m:\\output.bmp", "rb");
        if (!bmp) ok = false;
        else
        {
            const size_t body = (size_t)h->jpg.frame_info.img_width * h->jpg.frame_info.img_height * 4;
            if (fseek(bmp, 54, SEEK_SET) != 0 || 1 != fread(bgra_out, body, 1, bmp)) ok = false;
            fclose(bmp);
        }
    }
    if (chdir(old_cwd) != 0) ok = false;
    return ok ? 1 : 0;
}

// Frees what the reference itself leaks (parser.cpp:65,189,228; decoder.cpp:193).
void ref_close(void *handle)
{
    RefHandle *h = (RefHandle *)handle;
    if (!h) return;
    if (h->fp) fclose(h->fp);
    delete[] h->jpg.mcu_data;
    for (int i = 0; i < 4; i++) delete[] h->jpg.quantization_table[i];
    for (int i = 0; i < 32; i++)
    {
        if (!h->jpg.huffman_table[i]) continue;
        for (int n = 0; n < h->jpg.huffman_table[i]->num_codeword; n++)
            delete[] h->jpg.huffman_table[i]->codeword[n];
        delete h->jpg.huffman_table[i];
    }
    delete h;
}

// Direct access to the reference's arithmetic kernels for unit-level parity checks.
void ref_fast_idct(int32_t *block64)
{
    if (!g_idct_ready) { Initialize_Fast_IDCT(); g_idct_ready = true; }
    Fast_IDCT(block64);
}

} // extern "C"

// reference decoder.cpp:367 (external linkage, no header)
uint32_t YUV_to_RGB32(coef_t Y, coef_t U, coef_t V);
extern "C" uint32_t ref_yuv_to_rgb32(int32_t Y, int32_t U, int32_t V) { return YUV_to_RGB32(Y, U, V); }
