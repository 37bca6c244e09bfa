// kernels.cu -- the sm_100a kernels of the baseline-JPEG decode path and their launchers.
//
//   k_scan_count / k_scan_chunks / k_unstuff_write
//       pre-pass over the raw entropy-coded bytes: removes byte stuffing and fill bytes, finds the
//       RSTn markers, checks their numbering and emits the start offset of every restart interval
//       in the cleaned stream. Replaces read_more_data<>() (reference decoder.cpp:94-159) and the
//       marker handling of decode_huffman_data() (decoder.cpp:289-307).
//   k_huff_decode
//       one lane per decode segment (restart interval): Huffman decode with shared-memory LUTs,
//       DC prediction, de-zig-zag into a per-lane block staged in shared memory, cooperative
//       128-bit flush of whole 128-byte blocks into the int16 coefficient plane. Replaces
//       decode_huffman_block()/decode_huffman_data() (decoder.cpp:221-365), BitStream's cached
//       reader (bitstream.h:311-365) and HuffmanTree::findCodeInCache() (huffman.h:277-314).
//   k_idct_csc
//       per tile of 192 blocks: coefficient tile staged in swizzled shared memory (TMA tensor
//       load, or a plain vector-load variant), thread-per-block dequantise + Chen-Wang IDCT in
//       registers, then chroma replication + exact integer YCbCr->BGRA and 128-bit coalesced
//       stores. Replaces Fast_IDCT (cpuIDCT8x8.cpp:25-127), the upsample/colour loop of
//       decode_mcu_data() (decoder.cpp:429-495) and the OpenCL kernels idct8x8.cl:157-222.
//   k_expand_coefs
//       the reference's coefficient tap (int32, dequantised) from the int16 plane, for parity.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "b2j_internal.h"
#include "b2j_math.h"
#include "kernels.h"

namespace b2j {

__constant__ uint8_t c_zigzag2[64];   // 2 * natural index of scan position i (byte offset in a block)

static const uint8_t h_zigzag[64] = B2J_ZIGZAG_TABLE;

cudaError_t init_constants()
{
    uint8_t z2[64];
    for (int i = 0; i < 64; i++) z2[i] = (uint8_t)(2 * h_zigzag[i]);
    return cudaMemcpyToSymbol(c_zigzag2, z2, sizeof(z2));
}

// =====================================================================================
// Pre-pass: byte classification shared by the count and the write kernel.
// For byte b with predecessor p and successor n (all inside one image's scan):
//   b == FF : kept iff n == 00 (stuffed data byte). Otherwise it is a marker prefix / fill byte
//             and is dropped; if n is none of 00, FF, D0..D7 the scan ends here (EOI or any
//             other marker) -- "terminator".
//   b != FF : dropped iff p == FF and b is 00 (stuffing) or D0..D7 (RSTn, recorded as a marker);
//             kept otherwise.
// The reference keeps the Dn byte in its stream and consumes it at the interval switch
// (decoder.cpp:136-141, 296-297); here it is dropped and becomes a segment boundary instead.
struct ScanFlags { uint32_t keep, mark, term; };   // one bit per byte of the thread's 16

__device__ __forceinline__ ScanFlags classify16(const uint8_t *__restrict__ scan, uint32_t raw_len, uint32_t pos, uint32_t bytes[4])
{
    ScanFlags f = {0u, 0u, 0u};
    if (pos >= raw_len) { bytes[0] = bytes[1] = bytes[2] = bytes[3] = 0; return f; }
    const uint4 v = *reinterpret_cast<const uint4 *>(scan + pos);   // scan base and pos are 16 B aligned
    bytes[0] = v.x; bytes[1] = v.y; bytes[2] = v.z; bytes[3] = v.w;
    uint32_t p = pos ? scan[pos - 1] : 0u;
    const uint32_t after = (pos + 16 < raw_len) ? scan[pos + 16] : 0xD9u;   // running off the end acts like EOI
#pragma unroll
    for (int j = 0; j < 16; j++)
    {
        const uint32_t b = (bytes[j >> 2] >> (8 * (j & 3))) & 0xFFu;
        uint32_t n;
        if (j < 15) n = (bytes[(j + 1) >> 2] >> (8 * ((j + 1) & 3))) & 0xFFu; else n = after;
        const bool inside = pos + j < raw_len;
        if (pos + j + 1 >= raw_len) n = 0xD9u;
        const bool n_rst = (n & 0xF8u) == 0xD0u;
        const bool b_rst = (b & 0xF8u) == 0xD0u;
        bool keep, mark = false, term = false;
        if (b == 0xFFu) { keep = (n == 0x00u); term = !(n == 0x00u || n == 0xFFu || n_rst); }
        else { mark = (p == 0xFFu) && b_rst; keep = !((p == 0xFFu) && (b == 0x00u || b_rst)); }
        if (inside)
        {
            f.keep |= (uint32_t)keep << j;
            f.mark |= (uint32_t)mark << j;
            f.term |= (uint32_t)term << j;
        }
        p = b;
    }
    return f;
}

// Chunk-local position of the first terminator, or 0xFFFF.
__device__ __forceinline__ uint32_t block_first_term(uint32_t term_bits, uint32_t tid, uint32_t *s_min)
{
    if (tid == 0) *s_min = kNoTerm;
    __syncthreads();
    if (term_bits) atomicMin(s_min, tid * 16u + (uint32_t)(__ffs(term_bits) - 1));
    __syncthreads();
    return *s_min;
}

__device__ __forceinline__ uint32_t mask_below(uint32_t tid, uint32_t limit)   // bits of this thread's bytes that lie below `limit`
{
    const uint32_t lo = tid * 16u;
    if (limit >= lo + 16u) return 0xFFFFu;
    if (limit <= lo) return 0u;
    return (1u << (limit - lo)) - 1u;
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_count(const uint8_t *__restrict__ raw, const ImgDev *__restrict__ imgs, const uint32_t *__restrict__ chunk_img,
             uint32_t *__restrict__ chunk_cnt, uint32_t *__restrict__ chunk_term)
{
    __shared__ uint32_t s_min;
    __shared__ uint32_t s_warp[kScanThreads / 32];
    const uint32_t c = blockIdx.x, tid = threadIdx.x;
    const ImgDev &im = imgs[chunk_img[c]];
    const uint32_t pos = (c - im.chunk_first) * kScanChunkBytes + tid * 16u;
    uint32_t bytes[4];
    const ScanFlags f = classify16(raw + im.raw_off, im.raw_len, pos, bytes);
    const uint32_t term = block_first_term(f.term, tid, &s_min);
    const uint32_t live = mask_below(tid, term);
    uint32_t packed = __popc(f.keep & live) | (__popc(f.mark & live) << 16);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) packed += __shfl_xor_sync(0xFFFFFFFFu, packed, o);
    if ((tid & 31) == 0) s_warp[tid >> 5] = packed;
    __syncthreads();
    if (tid == 0)
    {
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < kScanThreads / 32; w++) tot += s_warp[w];
        chunk_cnt[c] = tot;
        chunk_term[c] = term;
    }
}

// One warp per image: exclusive scan of the chunk counts, truncated at the first terminator.
__global__ void __launch_bounds__(128)
k_scan_chunks(const ImgDev *__restrict__ imgs, int n_images, const uint32_t *__restrict__ chunk_cnt,
              const uint32_t *__restrict__ chunk_term, uint32_t *__restrict__ chunk_base_keep,
              uint32_t *__restrict__ chunk_base_mark, uint32_t *__restrict__ clean_len,
              uint32_t *__restrict__ seg_start, int32_t *__restrict__ status)
{
    const uint32_t lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n_images) return;
    const ImgDev &im = imgs[i];
    uint32_t run_keep = 0, run_mark = 0;
    bool dead = false;
    for (uint32_t base = 0; base < im.n_chunks; base += 32)
    {
        const uint32_t k = base + lane;
        const bool valid = k < im.n_chunks;
        uint32_t cnt = valid ? chunk_cnt[im.chunk_first + k] : 0u;
        const bool has_term = valid && chunk_term[im.chunk_first + k] != kNoTerm;
        const uint32_t tb = __ballot_sync(0xFFFFFFFFu, has_term);
        const uint32_t first = tb ? (uint32_t)(__ffs(tb) - 1) : 32u;
        const bool my_dead = dead || lane > first;
        if (my_dead) cnt = 0;
        uint32_t keep = cnt & 0xFFFFu, mark = cnt >> 16;
        uint32_t ik = keep, imk = mark;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, ik, o), b = __shfl_up_sync(0xFFFFFFFFu, imk, o);
            if (lane >= (uint32_t)o) { ik += a; imk += b; }
        }
        if (valid)
        {
            chunk_base_keep[im.chunk_first + k] = my_dead ? kChunkDead : run_keep + ik - keep;
            chunk_base_mark[im.chunk_first + k] = run_mark + imk - mark;
        }
        run_keep += __shfl_sync(0xFFFFFFFFu, ik, 31);
        run_mark += __shfl_sync(0xFFFFFFFFu, imk, 31);
        if (tb) dead = true;
    }
    if (lane == 0)
    {
        clean_len[i] = run_keep;
        seg_start[im.seg_first] = 0;
        // a missing RSTn makes the reference fail with "expected RSTn" (decoder.cpp:298-302)
        if (im.has_dri && run_mark + 1 < im.n_segs) atomicOr(&status[i], B2J_ST_RST_MISMATCH);
    }
}

__global__ void __launch_bounds__(kScanThreads)
k_unstuff_write(const uint8_t *__restrict__ raw, uint8_t *__restrict__ clean, const ImgDev *__restrict__ imgs,
                const uint32_t *__restrict__ chunk_img, const uint32_t *__restrict__ chunk_term,
                const uint32_t *__restrict__ chunk_base_keep, const uint32_t *__restrict__ chunk_base_mark,
                uint32_t *__restrict__ seg_start, int32_t *__restrict__ status)
{
    __shared__ uint32_t s_warp[kScanThreads / 32];
    const uint32_t c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t base_keep = chunk_base_keep[c];
    if (base_keep == kChunkDead) return;   // behind the end of the scan
    const uint32_t img_idx = chunk_img[c];
    const ImgDev &im = imgs[img_idx];
    const uint32_t pos = (c - im.chunk_first) * kScanChunkBytes + tid * 16u;
    uint32_t bytes[4];
    const ScanFlags f = classify16(raw + im.raw_off, im.raw_len, pos, bytes);
    const uint32_t live = mask_below(tid, chunk_term[c]);
    const uint32_t keep = f.keep & live, mark = f.mark & live;
    const uint32_t mine = __popc(keep) | (__popc(mark) << 16);
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += a;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) before += (w < (int)warp) ? s_warp[w] : 0u;
    const uint32_t excl = before + incl - mine;
    uint32_t out = base_keep + (excl & 0xFFFFu);       // offset in this image's clean stream
    uint32_t rank = chunk_base_mark[c] + (excl >> 16); // ordinal of the next RSTn in the image
    uint8_t *dst = clean + im.raw_off;
    if (!(keep | mark)) return;
#pragma unroll
    for (int j = 0; j < 16; j++)
    {
        const uint32_t b = (bytes[j >> 2] >> (8 * (j & 3))) & 0xFFu;
        if ((mark >> j) & 1u)
        {
            if (im.has_dri && rank + 1 < im.n_segs)
            {
                seg_start[im.seg_first + rank + 1] = out;
                if (b != 0xD0u + (rank & 7u)) atomicOr(&status[img_idx], B2J_ST_RST_MISMATCH);   // decoder.cpp:298
            }
            rank++;
        }
        if ((keep >> j) & 1u) dst[out++] = (uint8_t)b;
    }
}

// =====================================================================================
// Huffman decode.
struct BitReader
{
    uint32_t cur, nxt;     // big-endian words: `cur` holds the bit at `bitpos`
    uint32_t raw2;         // the word after `nxt`, still in memory byte order (swapped when shifted in)
    uint32_t bitpos;       // 0..31 after refill()
    const uint32_t *wp;    // next word to fetch

    __device__ __forceinline__ void init(const uint8_t *p)
    {
        const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
        wp = reinterpret_cast<const uint32_t *>(p - a);
        cur = __byte_perm(__ldg(wp), 0, 0x0123);
        nxt = __byte_perm(__ldg(wp + 1), 0, 0x0123);
        raw2 = __ldg(wp + 2);
        wp += 3;
        bitpos = a * 8u;
    }
    __device__ __forceinline__ uint32_t peek() const { return __funnelshift_l(nxt, cur, bitpos); }
    __device__ __forceinline__ void refill()
    {
        if (bitpos >= 32u)
        {
            cur = nxt;
            nxt = __byte_perm(raw2, 0, 0x0123);
            raw2 = __ldg(wp);
            wp++;
            bitpos -= 32u;
        }
    }
};

// One symbol from a two-level LUT held in shared memory. Returns the leaf entry (0 = no codeword).
__device__ __forceinline__ uint32_t lut_lookup(const uint16_t *__restrict__ tab, uint32_t peek)
{
    uint32_t e = tab[peek >> (32 - kLutBits)];
    if (e & kLutEscape)
    {
        const uint32_t nb = e & 15u, off = (e >> 4) & 0x7FFu;
        e = tab[(1u << kLutBits) + off + ((peek << kLutBits) >> (32u - nb))];
    }
    return e;
}

__global__ void __launch_bounds__(kHuffThreads)
k_huff_decode(const uint8_t *__restrict__ clean, const ImgDev *__restrict__ imgs, const HuffCtaDev *__restrict__ ctas,
              const uint32_t *__restrict__ seg_start, const uint32_t *__restrict__ clean_len,
              const uint16_t *__restrict__ luts, int16_t *__restrict__ coef, int32_t *__restrict__ status)
{
    extern __shared__ __align__(16) uint8_t smem[];
    // [ per-lane block slots: kHuffThreads * 128 B ][ LUT set ]
    uint8_t *s_slots = smem;
    uint16_t *s_lut = reinterpret_cast<uint16_t *>(smem + kHuffThreads * 128);
    __shared__ uint8_t s_zz2[64];   // lanes sit at different scan positions: shared, not constant, memory

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const HuffCtaDev cta = ctas[blockIdx.x];
    const ImgDev &im = imgs[cta.img];

    // stage the LUT set (16-byte granules) and clear the slots
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(luts + im.lut_off);
        uint4 *dst = reinterpret_cast<uint4 *>(s_lut);
        for (uint32_t k = tid; k < im.lut_len / 8; k += kHuffThreads) dst[k] = __ldg(src + k);
        uint4 *z = reinterpret_cast<uint4 *>(s_slots);
        for (uint32_t k = tid; k < kHuffThreads * 8; k += kHuffThreads) z[k] = make_uint4(0, 0, 0, 0);
        if (tid < 64) s_zz2[tid] = c_zigzag2[tid];
    }
    __syncthreads();

    const uint32_t tot = im.tot_blks, ny = im.ny_blks;
    const uint32_t seg = cta.seg_first + tid;
    bool active = seg < im.n_segs;
    uint32_t start = 0, end = 0, nblk = 0, blk0 = 0;
    bool check_end = false;
    if (active)
    {
        start = seg_start[im.seg_first + seg];
        const uint32_t mcu0 = seg * im.restart_interval;
        const uint32_t nmcu = min(im.restart_interval, im.mcu_count - mcu0);
        nblk = nmcu * tot;
        blk0 = im.blk_first + mcu0 * tot;
        if (seg + 1 < im.n_segs) { end = seg_start[im.seg_first + seg + 1]; check_end = (end != kSegInvalid); }
        if (!check_end) end = clean_len[cta.img];
    }
    // a lane whose start is unknown (missing RSTn) still emits zero blocks so the plane is defined
    const bool decodable = active && start != kSegInvalid;

    BitReader br;
    const uint8_t *base = clean + im.raw_off;
    br.init(base + (decodable ? start : 0u));
    const uint32_t *wp0 = br.wp;
    const uint32_t bit0 = br.bitpos;

    int32_t dc0 = 0, dc1 = 0, dc2 = 0;
    int32_t err = 0;
    bool dead = !decodable;

    // byte address (shared window) of this lane's slot, with the chunk swizzle folded in:
    // coefficient n lives at slot + ((n>>3) ^ (lane&7))*16 + (n&7)*2 == (slot | lane_xor) ^ (2n)
    const uint32_t slot_key = tid * 128u + ((lane & 7u) << 4);

    const uint32_t max_nblk = __reduce_max_sync(0xFFFFFFFFu, nblk);
    uint32_t bi = 0;   // block index inside the MCU (uniform across the warp: segments start on MCU boundaries)
    for (uint32_t b = 0; b < max_nblk; b++)
    {
        const uint32_t comp = bi < ny ? 0u : (bi - ny + 1u);
        const uint16_t *dc_tab = s_lut + s_lut[comp];
        const uint16_t *ac_tab = s_lut + s_lut[3 + comp];
        const bool mine = b < nblk;
        if (mine && !dead)
        {
            // ---- DC (decoder.cpp:226-233)
            uint32_t pk = br.peek();
            uint32_t e = lut_lookup(dc_tab, pk);
            uint32_t len = e & 31u, size = (e >> 5) & 31u;
            if (len == 0) { err |= B2J_ST_BAD_CODE; dead = true; }
            else
            {
                const int32_t diff = extend_top(pk << len, (int)size);
                br.bitpos += len + size;
                br.refill();
                int32_t dcv;
                if (comp == 0) { dc0 += diff; dcv = dc0; }
                else if (comp == 1) { dc1 += diff; dcv = dc1; }
                else { dc2 += diff; dcv = dc2; }
                if (dcv != (int32_t)(int16_t)dcv) err |= B2J_ST_DC_RANGE;
                *reinterpret_cast<int16_t *>(s_slots + slot_key) = (int16_t)dcv;
                // ---- AC (decoder.cpp:236-258)
                uint32_t pos = 1;
                bool eob = false;
                while (pos < 64u)
                {
                    pk = br.peek();
                    e = lut_lookup(ac_tab, pk);
                    len = e & 31u;
                    if (len == 0) { err |= B2J_ST_BAD_CODE; dead = true; eob = true; break; }
                    size = (e >> 5) & 31u;
                    const uint32_t run = (e >> 10) & 15u;
                    const int32_t v = extend_top(pk << len, (int)size);
                    br.bitpos += len + size;
                    pos += run;
                    if ((e >> 5) == 0u) { eob = true; break; }   // run == 0 && size == 0
                    if (size != 0u && pos < 64u)
                        *reinterpret_cast<int16_t *>(s_slots + (slot_key ^ (uint32_t)s_zz2[pos])) = (int16_t)v;
                    pos++;   // past the stored coefficient, or the extra zero of a size-0 run (decoder.cpp:247-252)
                    br.refill();
                }
                if (eob) br.refill();
                else if (pos > 64u) { err |= B2J_ST_BLOCK_OVERFLOW; dead = true; }   // decoder.cpp:259
            }
        }
        __syncwarp();
        // ---- cooperative flush: 8 lanes move one 128-byte block, 4 blocks per step
        const uint32_t my_dst = mine ? (blk0 + b) : 0xFFFFFFFFu;
        const uint32_t warp_slot0 = (tid & ~31u) * 128u;
#pragma unroll
        for (int it = 0; it < 8; it++)
        {
            const uint32_t j = (uint32_t)it * 4u + (lane >> 3);   // source lane
            const uint32_t dst = __shfl_sync(0xFFFFFFFFu, my_dst, j);
            const uint32_t ch = lane & 7u;
            if (dst != 0xFFFFFFFFu)
            {
                uint4 *sp = reinterpret_cast<uint4 *>(s_slots + warp_slot0 + j * 128u + ((ch ^ (j & 7u)) << 4));
                const uint4 v = *sp;
                *sp = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4 *>(coef + (size_t)dst * 64 + ch * 8) = v;
            }
        }
        __syncwarp();
        bi = (bi + 1 == tot) ? 0u : bi + 1;
    }

    if (decodable && !dead)
    {
        // bits consumed since `start`; a restart interval must end exactly at its marker
        // (decoder.cpp:296-302 aligns to the byte boundary and expects RSTn there)
        const uint64_t bits = (uint64_t)(br.wp - wp0) * 32u + br.bitpos - bit0;
        const uint64_t used = (bits + 7u) >> 3;
        const uint64_t avail = (uint64_t)end - start;
        if (used > avail) err |= B2J_ST_OVERRUN;
        else if (check_end && used != avail) err |= B2J_ST_SEGMENT_END;
    }
    if (err) atomicOr(&status[cta.img], err);
}

// =====================================================================================
// IDCT + upsample + colour.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

// Packed clamp(a + b, 0, 255) on two int16 lanes.
__device__ __forceinline__ uint32_t addclamp2(uint32_t a, uint32_t b)
{
    return __viaddmin_s16x2_relu(a, b, 0x00FF00FFu);
}

// Converts RV rows x 4 pixels. y2[r][0..1]: packed int16 luma pairs of row r; cb2/cr2: the chroma
// samples of the two pixel pairs (already replicated horizontally) as packed pairs. Writes the
// 4 BGRA words per row.
template <int RH, int RV>
__device__ __forceinline__ void csc_rows(const uint32_t (*y2)[2], const uint32_t cb2[2], const uint32_t cr2[2], uint32_t (*out)[4])
{
#pragma unroll
    for (int h = 0; h < 2; h++)   // pixel pairs (0,1) and (2,3)
    {
        const int32_t u0 = (int16_t)(cb2[h] & 0xFFFFu), u1 = (int32_t)cb2[h] >> 16;
        const int32_t v0 = (int16_t)(cr2[h] & 0xFFFFu), v1 = (int32_t)cr2[h] >> 16;
        uint32_t ro, go, bo;
        bool special;
        if (RH == 2)
        {   // both pixels of the pair share one chroma sample
            ro = (uint32_t)(csc_r_off(v0) & 0xFFFF) * 0x10001u;
            go = (uint32_t)(csc_g_off(u0, v0) & 0xFFFF) * 0x10001u;
            bo = (uint32_t)(csc_b_off(u0) & 0xFFFF) * 0x10001u;
            special = (u0 == -200 && v0 == 200);
        }
        else
        {
            ro = (uint32_t)(csc_r_off(v0) & 0xFFFF) | ((uint32_t)csc_r_off(v1) << 16);
            go = (uint32_t)(csc_g_off(u0, v0) & 0xFFFF) | ((uint32_t)csc_g_off(u1, v1) << 16);
            bo = (uint32_t)(csc_b_off(u0) & 0xFFFF) | ((uint32_t)csc_b_off(u1) << 16);
            special = (u0 == -200 && v0 == 200) || (u1 == -200 && v1 == 200);
        }
#pragma unroll
        for (int r = 0; r < RV; r++)
        {
            const uint32_t yy = y2[r][h];
            const uint32_t R = addclamp2(yy, ro), B = addclamp2(yy, bo);
            uint32_t G = addclamp2(yy, go);
            if (special)
            {
                // the one double-rounding case of the reference (see b2j_math.h)
                const int32_t ya = (int16_t)(yy & 0xFFFFu), yb = (int32_t)yy >> 16;
                const uint32_t ga = clamp255(ya + csc_g_off(u0, v0) - csc_g_fix(ya, u0, v0));
                const uint32_t gb = clamp255(yb + csc_g_off(u1, v1) - csc_g_fix(yb, u1, v1));
                G = ga | (gb << 16);
            }
            const uint32_t bg = __byte_perm(B, G, 0x6240);          // B0 G0 B1 G1
            out[r][2 * h + 0] = __byte_perm(bg, R, 0x5410);         // B0 G0 R0 0
            out[r][2 * h + 1] = __byte_perm(bg, R, 0x7632);         // B1 G1 R1 0
        }
    }
}

// Layout constants of a tile for luma sampling RH x RV (chroma 1x1):
//   4:4:4 <1,1>  MCU 8x8,   3 blocks   | 4:2:0 <2,2>  MCU 16x16, 6 blocks
//   4:2:2 <2,1>  MCU 16x8,  4 blocks   | 4:4:0 <1,2>  MCU 8x16,  4 blocks
template <int RH, int RV>
struct TileGeom
{
    static constexpr uint32_t ny = RH * RV, tot = RH * RV + 2, yh = RH;
    static constexpr uint32_t mcu_w = 8 * RH, mcu_h = 8 * RV;
    static constexpr uint32_t mcus = kTileBlocks / tot;   // MCUs per tile
    static constexpr uint32_t xg = mcu_w / 4;             // 4-pixel groups per MCU row
};

template <int RH, int RV>
__device__ __forceinline__ void csc_phase(const uint8_t *__restrict__ s_tile, const uint2 *__restrict__ s_mcu_xy,
                                          const ImgDev &im, uint32_t n_mcus, uint8_t *__restrict__ pix, uint32_t tid)
{
    using G = TileGeom<RH, RV>;
    // work item = 4 pixels x RV rows; consecutive threads walk along x through the MCUs of the tile,
    // then down the 8 row groups every layout has (mcu_h / RV == 8)
    const uint32_t per_rg = n_mcus * G::xg;
    const bool vec_ok = (im.width & 3u) == 0u;
    uint32_t rg = 0, rem = tid;
    while (true)
    {
        while (rem >= per_rg) { rem -= per_rg; rg++; }
        if (rg >= 8u) break;
        const uint32_t m = rem / G::xg, x4 = rem % G::xg;
        const uint2 mxy = s_mcu_xy[m];
        const uint32_t xin = x4 * 4u, yin0 = rg * RV;
        const uint32_t px = mxy.x * G::mcu_w + xin;
        const uint32_t py0 = mxy.y * G::mcu_h + yin0;
        rem += kTileBlocks;
        if (px >= im.width || py0 >= im.height) continue;
        const uint32_t row0 = m * G::tot;
        // luma
        uint32_t y2[RV][2];
#pragma unroll
        for (int r = 0; r < RV; r++)
        {
            const uint32_t yin = yin0 + r;
            const uint32_t srow = row0 + (yin >> 3) * G::yh + (xin >> 3);
            const uint32_t off = srow * 128u + (((yin & 7u) ^ (srow & 7u)) << 4) + (xin & 7u) * 2u;
            const uint2 v = *reinterpret_cast<const uint2 *>(s_tile + off);
            y2[r][0] = v.x; y2[r][1] = v.y;
        }
        // chroma (pixel replication, decoder.cpp:478-480)
        uint32_t cb2[2], cr2[2];
        {
            const uint32_t cx = xin / RH, cy = rg;   // yin / RV == rg in every layout
            const uint32_t brow = row0 + G::ny, rrow = brow + 1u;
            const uint32_t boff = brow * 128u + ((cy ^ (brow & 7u)) << 4) + cx * 2u;
            const uint32_t roff = rrow * 128u + ((cy ^ (rrow & 7u)) << 4) + cx * 2u;
            if (RH == 2)
            {
                const uint32_t b = *reinterpret_cast<const uint32_t *>(s_tile + boff);
                const uint32_t r = *reinterpret_cast<const uint32_t *>(s_tile + roff);
                cb2[0] = __byte_perm(b, 0, 0x1010); cb2[1] = __byte_perm(b, 0, 0x3232);
                cr2[0] = __byte_perm(r, 0, 0x1010); cr2[1] = __byte_perm(r, 0, 0x3232);
            }
            else
            {
                const uint2 b = *reinterpret_cast<const uint2 *>(s_tile + boff);
                const uint2 r = *reinterpret_cast<const uint2 *>(s_tile + roff);
                cb2[0] = b.x; cb2[1] = b.y; cr2[0] = r.x; cr2[1] = r.y;
            }
        }
        uint32_t out[RV][4];
        csc_rows<RH, RV>(y2, cb2, cr2, out);
#pragma unroll
        for (int r = 0; r < RV; r++)
        {
            const uint32_t py = py0 + r;
            if (py >= im.height) break;
            uint8_t *dst = pix + im.pix_off + ((size_t)py * im.width + px) * 4u;
            if (vec_ok)   // width % 4 == 0 -> the group is whole and 16-byte aligned
                *reinterpret_cast<uint4 *>(dst) = make_uint4(out[r][0], out[r][1], out[r][2], out[r][3]);
            else
            {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (px + k < im.width) reinterpret_cast<uint32_t *>(dst)[k] = out[r][k];
            }
        }
    }
}

struct TileSmem
{
    // coefficient tile: kTileBlocks rows of 128 B; the 16-byte chunk c of row r sits at chunk c ^ (r & 7)
    // (the TMA SWIZZLE_128B pattern; the non-TMA variant stores with the same XOR)
    alignas(1024) uint8_t tile[kTileBlocks * 128];
    alignas(16) uint16_t qt[3][64];
    uint2 mcu_xy[kTileBlocks / 3];
    alignas(8) uint64_t bar;
};

template <int RH, int RV, bool USE_TMA>
__device__ __forceinline__ void tile_body(TileSmem &sm, const CUtensorMap *tmap, const int16_t *__restrict__ coef, const ImgDev &im,
                                          const TileDev tile, const uint16_t *__restrict__ qtabs, uint8_t *__restrict__ pix,
                                          int32_t *__restrict__ status)
{
    using G = TileGeom<RH, RV>;
    const uint32_t tid = threadIdx.x;
    const uint32_t row_first = im.blk_first + tile.mcu_first * G::tot;
    const uint32_t n_mcus = min(G::mcus, im.mcu_count - tile.mcu_first);

    if (USE_TMA)
    {
        if (tid == 0)
        {
            mbar_init(&sm.bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0)
        {
            mbar_expect_tx(&sm.bar, kTileBlocks * 128);
            tma_load_2d(sm.tile, tmap, 0, (int)row_first, &sm.bar);
        }
    }
    else
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(coef + (size_t)row_first * 64);
        uint4 v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = __ldg(src + (uint32_t)k * kTileBlocks + tid);   // the plane is padded by one tile
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
            const uint32_t idx = (uint32_t)k * kTileBlocks + tid;   // 16-byte granule: row = idx>>3, chunk = idx&7
            const uint32_t r = idx >> 3, c = idx & 7u;
            *reinterpret_cast<uint4 *>(sm.tile + r * 128u + ((c ^ (r & 7u)) << 4)) = v[k];
        }
    }
    // quantisers (natural order, per component) and MCU coordinates while the tile is in flight
    if (tid < 96) reinterpret_cast<uint32_t *>(&sm.qt[0][0])[tid] = __ldg(reinterpret_cast<const uint32_t *>(qtabs + (size_t)tile.img * 192) + tid);
    if (tid < n_mcus)
    {
        const uint32_t gm = tile.mcu_first + tid;
        const uint32_t my = gm / im.mcu_count_w;
        sm.mcu_xy[tid] = make_uint2(gm - my * im.mcu_count_w, my);
    }
    if (USE_TMA)
    {
        uint32_t spins = 0;
        while (!mbar_try_wait(&sm.bar, 0))
        {
            if (++spins > (1u << 22)) { if (tid == 0) atomicOr(&status[tile.img], 0x4000); break; }   // never hang the GPU
        }
    }
    __syncthreads();

    // ---- thread-per-block dequantise + IDCT, in place (decoder.cpp:340, cpuIDCT8x8.cpp:25-127)
    {
        const uint32_t bi = tid % G::tot;
        const uint32_t comp = bi < G::ny ? 0u : (bi - G::ny + 1u);
        const uint16_t *q = sm.qt[comp];
        uint8_t *rowp = sm.tile + tid * 128u;
        const uint32_t sw = tid & 7u;
        int32_t v[64];
#pragma unroll
        for (int r = 0; r < 8; r++)
        {
            const uint4 cw = *reinterpret_cast<const uint4 *>(rowp + ((r ^ sw) << 4));
            const uint4 qw = *reinterpret_cast<const uint4 *>(q + 8 * r);
            const uint32_t c4[4] = {cw.x, cw.y, cw.z, cw.w}, q4[4] = {qw.x, qw.y, qw.z, qw.w};
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                v[8 * r + 2 * k + 0] = (int32_t)(int16_t)(c4[k] & 0xFFFFu) * (int32_t)(q4[k] & 0xFFFFu);
                v[8 * r + 2 * k + 1] = ((int32_t)c4[k] >> 16) * (int32_t)(q4[k] >> 16);
            }
        }
        idct_8x8(v);
#pragma unroll
        for (int r = 0; r < 8; r++)
        {
            uint4 o;
            o.x = (uint32_t)(v[8 * r + 0] & 0xFFFF) | ((uint32_t)v[8 * r + 1] << 16);
            o.y = (uint32_t)(v[8 * r + 2] & 0xFFFF) | ((uint32_t)v[8 * r + 3] << 16);
            o.z = (uint32_t)(v[8 * r + 4] & 0xFFFF) | ((uint32_t)v[8 * r + 5] << 16);
            o.w = (uint32_t)(v[8 * r + 6] & 0xFFFF) | ((uint32_t)v[8 * r + 7] << 16);
            *reinterpret_cast<uint4 *>(rowp + ((r ^ sw) << 4)) = o;
        }
    }
    __syncthreads();

    // ---- chroma replication + colour + store (decoder.cpp:443-495, 367-370)
    csc_phase<RH, RV>(sm.tile, sm.mcu_xy, im, n_mcus, pix, tid);
}

template <bool USE_TMA>
__global__ void __launch_bounds__(kTileBlocks)
k_idct_csc(const __grid_constant__ CUtensorMap tmap, const int16_t *__restrict__ coef, const ImgDev *__restrict__ imgs,
           const TileDev *__restrict__ tiles, const uint16_t *__restrict__ qtabs, uint8_t *__restrict__ pix, int32_t *__restrict__ status)
{
    __shared__ TileSmem sm;
    const TileDev tile = tiles[blockIdx.x];
    const ImgDev &im = imgs[tile.img];
    switch (im.mode)   // uniform per CTA
    {
    case kMode444: tile_body<1, 1, USE_TMA>(sm, &tmap, coef, im, tile, qtabs, pix, status); break;
    case kMode420: tile_body<2, 2, USE_TMA>(sm, &tmap, coef, im, tile, qtabs, pix, status); break;
    case kMode422: tile_body<2, 1, USE_TMA>(sm, &tmap, coef, im, tile, qtabs, pix, status); break;
    default:       tile_body<1, 2, USE_TMA>(sm, &tmap, coef, im, tile, qtabs, pix, status); break;
    }
}

// =====================================================================================
// Coefficient tap: int16 quantised plane -> the reference's int32 dequantised mcu_data layout.
__global__ void __launch_bounds__(256)
k_expand_coefs(const int16_t *__restrict__ coef, const uint16_t *__restrict__ qtab, uint32_t blk_count,
               uint32_t tot, uint32_t ny, int32_t *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)blk_count * 64) return;
    const uint32_t blk = (uint32_t)(i >> 6), k = (uint32_t)(i & 63);
    const uint32_t bi = blk % tot;
    const uint32_t comp = bi < ny ? 0u : (bi - ny + 1u);
    out[i] = (int32_t)coef[i] * (int32_t)qtab[comp * 64 + k];
}

// =====================================================================================
// Launchers (host).
size_t huff_smem_bytes(uint32_t max_lut_len) { return (size_t)kHuffThreads * 128 + (size_t)max_lut_len * 2; }

cudaError_t configure_kernels(uint32_t max_lut_len)
{
    return cudaFuncSetAttribute(k_huff_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)huff_smem_bytes(max_lut_len));
}

void launch_prepass(const DecodeArgs &a, cudaStream_t s)
{
    if (a.n_chunks == 0) return;
    k_scan_count<<<a.n_chunks, kScanThreads, 0, s>>>(a.raw, a.imgs, a.chunk_img, a.chunk_cnt, a.chunk_term);
    k_scan_chunks<<<(a.n_images + 3) / 4, 128, 0, s>>>(a.imgs, a.n_images, a.chunk_cnt, a.chunk_term, a.chunk_base_keep,
                                                       a.chunk_base_mark, a.clean_len, a.seg_start, a.status);
    k_unstuff_write<<<a.n_chunks, kScanThreads, 0, s>>>(a.raw, a.clean, a.imgs, a.chunk_img, a.chunk_term, a.chunk_base_keep,
                                                        a.chunk_base_mark, a.seg_start, a.status);
}

void launch_huffman(const DecodeArgs &a, cudaStream_t s)
{
    if (a.n_huff_ctas == 0) return;
    k_huff_decode<<<a.n_huff_ctas, kHuffThreads, huff_smem_bytes(a.max_lut_len), s>>>(a.clean, a.imgs, a.huff_ctas, a.seg_start,
                                                                                      a.clean_len, a.luts, a.coef, a.status);
}

void launch_idct(const DecodeArgs &a, cudaStream_t s)
{
    if (a.n_tiles == 0) return;
    if (a.use_tma)
        k_idct_csc<true><<<a.n_tiles, kTileBlocks, 0, s>>>(*a.tmap, a.coef, a.imgs, a.tiles, a.qtabs, a.pixels, a.status);
    else
        k_idct_csc<false><<<a.n_tiles, kTileBlocks, 0, s>>>(*a.tmap, a.coef, a.imgs, a.tiles, a.qtabs, a.pixels, a.status);
}

void launch_expand(const int16_t *coef, const uint16_t *qtab, uint32_t blk_count, uint32_t tot, uint32_t ny, int32_t *out, cudaStream_t s)
{
    const size_t n = (size_t)blk_count * 64;
    if (!n) return;
    k_expand_coefs<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(coef, qtab, blk_count, tot, ny, out);
}

} // namespace b2j
