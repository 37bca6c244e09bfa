// idct_b2j.cpp -- the reference's device backend (idct.h:9-18) on top of the B200 C ABI.
//
// Drop-in for the reference's oclDCT8x8.cpp (+ idct8x8.cl): the reference's OWN decoder.cpp, built without
// USE_CPU_ONLY (decoder.cpp:11), entropy-decodes on the CPU and drives the ten functions below in this order
// (decoder.cpp:202-217, 352-357, 402-416, 518-521):
//
//   Initialize_OpenCL_IDCT            main.cpp:22            oclDCT8x8.cpp:25-74    -> b2j_create
//   clidct_create                     decoder.cpp:204        oclDCT8x8.cpp:76-110   -> (context already there)
//   clidct_allocate_memory            decoder.cpp:207        oclDCT8x8.cpp:112-144  -> b2j_idct_create (geometry)
//   clidct_build                      decoder.cpp:211        oclDCT8x8.cpp:167-273  -> checks the colour space
//   clidct_transfer_data_to_device    decoder.cpp:356        oclDCT8x8.cpp:146-165  -> b2j_idct_upload
//   clidct_run                        decoder.cpp:406        oclDCT8x8.cpp:275-299  -> b2j_idct_run
//   clidct_wait_for_completion        decoder.cpp:408        oclDCT8x8.cpp:301-304  -> (b2j_idct_run has synchronised)
//   clidct_retrieve_image_from_device decoder.cpp:414        oclDCT8x8.cpp:186-205  -> b2j_idct_read_pixels
//   clidct_retrieve_data_from_device  decoder.cpp:415        oclDCT8x8.cpp:177-184  -> b2j_idct_read_coefs
//   clidct_clean_up                   decoder.cpp:520        oclDCT8x8.cpp:306-341  -> b2j_idct_destroy
//
// Same `true` = success convention. The pixels are those of the reference's CPU path (cpuIDCT8x8 + YUV_to_RGB32,
// bit-exact), not the OpenCL kernel's float colour. One image in flight, like the reference.
#ifdef B2J_USE_REFERENCE_HEADERS
#include "stdafx.h"
#include "macro.h"
#include "jpeg.h"
#include "idct.h"
#else
#include "refabi.h"
int Initialize_OpenCL_IDCT();
bool clidct_create();
bool clidct_allocate_memory(const int total_blocks, const size_t image_width, const size_t image_height, const int mcu_width, const int mcu_height);
bool clidct_transfer_data_to_device(const int block_data_src[1][64], const int offset, const int count);
bool clidct_build(ColorSpace colorspace);
bool clidct_run(ColorSpace colorspace);
bool clidct_retrieve_data_from_device(int block_data_dest[1][64]);
bool clidct_retrieve_image_from_device(void *img_data_dest, const size_t img_width, const size_t img_height);
bool clidct_wait_for_completion();
bool clidct_clean_up();
#endif

#include <stddef.h>
#include <stdio.h>

#include "../../../include/b2j.h"

namespace {

b2j_ctx *g_ictx = nullptr;
b2j_idct *g_idct = nullptr;
int g_luma_h = 0, g_luma_v = 0;
size_t g_width = 0, g_height = 0;

bool ifail(const char *what, int rc)
{
    printf("[X] %s: %s (%s)\n", what, b2j_strerror(rc), b2j_last_error());
    return false;
}

} // namespace

int Initialize_OpenCL_IDCT()
{
    if (g_ictx) return 0;
    const int rc = b2j_create(0, &g_ictx);
    if (rc != B2J_OK) { ifail("Initialize_OpenCL_IDCT (b2j_create)", rc); return -1; }
    return 0;
}

bool clidct_create()
{
    return g_ictx != nullptr || Initialize_OpenCL_IDCT() == 0;
}

bool clidct_allocate_memory(const int total_blocks, const size_t image_width, const size_t image_height, const int mcu_width, const int mcu_height)
{
    if (!g_ictx) return false;
    if (g_idct) { b2j_idct_destroy(g_idct); g_idct = nullptr; }
    g_luma_h = mcu_width / 8; g_luma_v = mcu_height / 8;
    g_width = image_width; g_height = image_height;
    const int rc = b2j_idct_create(g_ictx, (int)image_width, (int)image_height, g_luma_h, g_luma_v, &g_idct);
    if (rc != B2J_OK) return ifail("clidct_allocate_memory", rc);
    if (b2j_idct_blk_count(g_idct) != total_blocks)
    {
        printf("[X] clidct_allocate_memory: %d blocks announced, the geometry has %d\n", total_blocks, b2j_idct_blk_count(g_idct));
        return false;
    }
    return true;
}

bool clidct_build(ColorSpace colorspace)
{
    // the reference compiles one of two kernels here (oclDCT8x8.cpp:167-273); the geometry already fixes ours
    if (colorspace == YUV444) return g_luma_h == 1 && g_luma_v == 1;
    if (colorspace == YUV411) return g_luma_h == 2 && g_luma_v == 2;
    return false;
}

bool clidct_transfer_data_to_device(const int block_data_src[1][64], const int offset, const int count)
{
    if (!g_idct) return false;
    const int rc = b2j_idct_upload(g_idct, &block_data_src[0][0], offset, count);
    return rc == B2J_OK ? true : ifail("clidct_transfer_data_to_device", rc);
}

bool clidct_run(ColorSpace)
{
    if (!g_idct) return false;
    const int rc = b2j_idct_run(g_idct);
    return rc == B2J_OK ? true : ifail("clidct_run", rc);
}

bool clidct_wait_for_completion() { return g_idct != nullptr; }

bool clidct_retrieve_data_from_device(int block_data_dest[1][64])
{
    if (!g_idct) return false;
    const int rc = b2j_idct_read_coefs(g_idct, &block_data_dest[0][0]);
    return rc == B2J_OK ? true : ifail("clidct_retrieve_data_from_device", rc);
}

bool clidct_retrieve_image_from_device(void *img_data_dest, const size_t img_width, const size_t img_height)
{
    if (!g_idct || img_width != g_width || img_height != g_height) return false;
    const int rc = b2j_idct_read_pixels(g_idct, static_cast<uint8_t *>(img_data_dest));
    return rc == B2J_OK ? true : ifail("clidct_retrieve_image_from_device", rc);
}

bool clidct_clean_up()
{
    if (g_idct) { b2j_idct_destroy(g_idct); g_idct = nullptr; }
    return true;
}
