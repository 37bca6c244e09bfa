// b2jdec_main.cpp -- stand-alone command line with the reference's argv contract (main.cpp:10-39:
// "prog file1 [file2 ...]", each file decoded to the BMP path) for use WITHOUT the reference tree.
// It parses with b2j_parse_header() and drives the same four decoder.h functions the reference's
// load_jpg() drives (parser.cpp:359-400).
#include "refabi.h"

#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../../include/b2j.h"

static bool load_jpg(const char *path)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) { puts("Couldn't open file."); return false; }
    std::vector<uint8_t> file;
    uint8_t buf[65536];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), fp)) > 0) file.insert(file.end(), buf, buf + got);
    b2j_image_desc d;
    const char *g = getenv("B2J_GATE");
    const int rc = b2j_parse_header(file.data(), file.size(), (g && (g[0] == 'e' || g[0] == '1')) ? B2J_GATE_EXTENDED : B2J_GATE_REFERENCE, &d);
    if (rc != B2J_OK) { printf("[X] this file is not supported (%s)\n", b2j_strerror(rc)); fclose(fp); return true; }
    // JPG_DATA as the reference's readers would have filled it
    JPG_DATA jpg;
    memset(&jpg, 0, sizeof(jpg));
    jpg.frame_info.bit_depth = 8;
    jpg.frame_info.img_width = (uint16_t)d.width;
    jpg.frame_info.img_height = (uint16_t)d.height;
    jpg.frame_info.num_channels = jpg.scan_info.num_channels = 3;
    jpg.dri_info.restart_interval = (uint16_t)d.restart_interval;
    for (int c = 0; c < 3; c++)
    {
        jpg.frame_info.channel_info[c].id = (uint8_t)(c + 1);
        jpg.frame_info.channel_info[c].sampling_factor = d.sampling[c];
        jpg.frame_info.channel_info[c].quant_tbl_id = d.quant_id[c];
        jpg.scan_info.channel_data[c].huff_tbl_id = d.huff_id[c];
    }
    for (int t = 0; t < 4; t++)
        if (d.quant_present[t])
        {
            jpg.quantization_table[t] = new coef_t[64];
            for (int k = 0; k < 64; k++) jpg.quantization_table[t][k] = d.quant[t][k];
        }
    for (int slot = 0; slot < 8; slot++)
        if (d.huff_present[slot])
        {
            HUFFMAN_TABLE *t = new HUFFMAN_TABLE;
            memset(t, 0, sizeof(*t));
            unsigned code = 0;
            for (int l = 1; l <= 16; l++)
            {
                for (int k = 0; k < d.huff_counts[slot][l - 1]; k++)
                {
                    char *s = new char[20];
                    for (int b = 0; b < l; b++) s[b] = ((code >> (l - 1 - b)) & 1) ? '1' : '0';
                    s[l] = 0;
                    t->codeword[t->num_codeword] = s;
                    t->value[t->num_codeword] = d.huff_symbols[slot][t->num_codeword];
                    t->num_codeword++;
                    code++;
                }
                code <<= 1;
            }
            jpg.huffman_table[((slot >> 2) << 4) | (slot & 3)] = t;
        }
    bool ok = is_supported_file(jpg) && decode_init(jpg);
    if (ok)
    {
        fseek(fp, (long)d.scan_offset, SEEK_SET);
        ok = decode_huffman_data(jpg, fp);
        if (!ok) puts("[X] decode_huffman_data() failed");
    }
    if (ok)
    {
        ok = decode_mcu_data(jpg, fp);
        puts(ok ? "[ ] decoding completed." : "[X] decode_mcu_data() failed");
    }
    fclose(fp);
    return true;
}

int main(int argc, char **argv)
{
    if (argc <= 1) { printf("Usage: %s file1 [file2 file3 ...]\n", argv[0]); return 0; }
    for (int i = 1; i < argc; i++)
    {
        printf("Processing %s\n", argv[i]);
        load_jpg(argv[i]);
    }
    return 0;
}
