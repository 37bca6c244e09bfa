// refabi.h -- the host-side types the reference's decoder API is written against, declared here so
// that the shim builds stand-alone. Field names, order and packing mirror the reference's
// src/jpeg.h:3-81, src/macro.h:114-119 and src/decoder.h:4-7 because the shim is a drop-in for
// decoder.cpp + oclDCT8x8.cpp: reference translation units (parser.cpp, main.cpp) compiled against
// THEIR jpeg.h must be link- and layout-compatible with it. When the shim is built inside the
// reference tree, define B2J_USE_REFERENCE_HEADERS and the reference's own headers are used
// instead of this file; the static_asserts in decoder_b2j.cpp check that both
// agree on every offset the shim touches.
#ifndef B2J_REFABI_H_INCLUDED
#define B2J_REFABI_H_INCLUDED

#include <stdint.h>
#include <stdio.h>

enum ColorSpace { YUV444, YUV411, Other };   // macro.h:114-119

typedef int coef_t;                           // jpeg.h:57

#pragma pack(push, 1)
struct APP0                                   // jpeg.h:4-15 (16 bytes)
{
    uint16_t len;
    uint8_t id[5];
    uint16_t ver;
    uint8_t res_unit;
    uint16_t res_x, res_y;
    uint8_t thumbnail_width, thumbnail_height;
};
struct SOF0                                   // jpeg.h:17-29 (15 bytes, as stored in the file)
{
    uint8_t bit_depth;
    uint16_t img_height, img_width;           // byte-swapped to host order by read_sof()
    uint8_t num_channels;
    struct { uint8_t id, sampling_factor, quant_tbl_id; } channel_info[3];
};
struct SOS                                    // jpeg.h:31-40 (10 bytes)
{
    uint8_t num_channels;
    struct { uint8_t id, huff_tbl_id; } channel_data[3];
    uint8_t reserved[3];
};
struct DRI { uint16_t restart_interval; };    // jpeg.h:42-45
#pragma pack(pop)

struct HUFFMAN_TABLE                          // jpeg.h:50-55
{
    int num_codeword;
    const char *codeword[256];                // ASCII '0'/'1' strings in DHT order (parser.cpp:221-257)
    uint8_t value[256];
};

struct JPG_DATA                               // jpeg.h:59-81
{
    APP0 app0;
    coef_t *quantization_table[4];            // 64 ints each, file (zig-zag) order
    SOF0 frame_info;
    HUFFMAN_TABLE *huffman_table[32];         // index (Tc<<4)|Th
    void *thumbnail;
    SOS scan_info;
    DRI dri_info;
    ColorSpace color_space;
    int mcu_width, mcu_height;                // pixels
    int mcu_count_w, mcu_count_h, mcu_count;
    coef_t (*mcu_data)[64];                   // the coefficient tap
    int blks_per_mcu[4];
    int tot_blks_per_mcu;
    int blk_count;
};

// decoder.h:4-7
bool is_supported_file(const JPG_DATA &jpg);
bool decode_init(JPG_DATA &jpg);
bool decode_huffman_data(const JPG_DATA &jpg, FILE *const strm);
bool decode_mcu_data(const JPG_DATA &jpg, FILE *const strm);

#endif
