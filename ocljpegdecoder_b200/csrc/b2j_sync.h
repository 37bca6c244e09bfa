// b2j_sync.h -- the self-synchronisation of one "chunk" of a stream without restart markers, written once for
// the device (kernels.cu: k_sync_chunks / k_sync_sweep, one CTA per chunk, one thread per lane, phases separated by
// __syncthreads) and for a host emulation (tests/native/synccheck.cpp: the same phases run lane after lane), so that
// the round / merge / bookkeeping logic can be checked on a machine without a GPU.
//
// Replaces the serial scan loop of decode_huffman_data() (reference decoder.cpp:286-346) by a parallel search for
// the decoder state at every sub-sequence border.
//
// A chunk is kSyncLanes consecutive sub-sequences (kSubBytes * 8 bits each) of one image's clean stream plus up to
// kSyncPre sub-sequences in front of them (the "pre-lanes"), which are walked only to give the first lane of the
// chunk a state that is in step with the true decoder.
//   round 0   every lane walks its own sub-sequence from a guessed state (its border, block 0, DC expected), in two
//             halves; the state at the first symbol at or behind the middle plus what the second half contributes
//             are kept as a checkpoint;
//   round r   a lane whose predecessor's exit state is not the entry state it last started from walks again from
//             that exit state, up to the middle: a decoder that started from a wrong state falls into step with the
//             true one within a few dozen symbols, so nearly always it arrives exactly at the checkpoint -- then
//             round 0's second half stands and the exit state does not change; otherwise it walks on to the end.
//             Rounds repeat until no lane's entry state changes any more: lane k of the chunk is final after at most
//             k rounds, in practice after two or three. Everything lives in shared memory; no global round trips,
//             no launches between rounds.
//   output    per lane the record the decode kernel needs (blocks started, DC sums, first block start), per chunk
//             the entry state it assumed, its exit state and its totals.
// A chunk's result is right if the entry state of its first lane is right. For the first chunk of an image that is
// the true start state; for every other chunk it is the exit state of its pre-lanes, i.e. right unless the guess
// failed to fall into step within kSyncPre sub-sequences. k_sync_sweep compares every chunk's assumed entry with its
// predecessor's exit and re-runs the (rare) chunks that disagree, in order, with the entry forced: correctness never
// depends on luck.
#ifndef B2J_SYNC_H_INCLUDED
#define B2J_SYNC_H_INCLUDED

#include <stdint.h>

#include "b2j_internal.h"

#ifdef __CUDACC__
#define B2J_HD __host__ __device__ __forceinline__
#else
#define B2J_HD inline
#endif

#ifndef __CUDACC__
// host emulation (tests/native/synccheck.cpp): the two CUDA vector types this header uses
struct uint2 { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 v = {x, y}; return v; }
#endif

// host emulation only: counts steps: whole AC groups [0], first symbol of a group [1], through the decode tables [2], DC [3]
#if defined(B2J_WALK_STATS) && !defined(__CUDACC__)
extern uint64_t g_b2j_walk_steps[4];
#define B2J_WALK_COUNT(k) (g_b2j_walk_steps[k]++)
#else
#define B2J_WALK_COUNT(k) ((void)0)
#endif

namespace b2j {

// ---- arithmetic shared by device and host ---------------------------------------------------------------------
// high word of (hi:lo) << (sh & 31)
B2J_HD uint32_t fsh_l(uint32_t lo, uint32_t hi, uint32_t sh)
{
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, sh);
#else
    sh &= 31u;
    return sh ? (hi << sh) | (lo >> (32u - sh)) : hi;
#endif
}
B2J_HD uint32_t bswap32(uint32_t v)
{
#ifdef __CUDA_ARCH__
    return __byte_perm(v, 0, 0x0123);
#else
    return __builtin_bswap32(v);
#endif
}
// Value bits -> signed coefficient (JPEG EXTEND, decoder.cpp:72-82). v: the bits left-aligned, size = their number
// (0 -> 0). A leading 1 bit means positive.
B2J_HD int32_t extend_bits(uint32_t v, uint32_t size)
{
    const uint32_t u = fsh_l(v, 0u, size);                   // the value bits as a number: v >> (32 - size)
    const uint32_t neg = (uint32_t)((int32_t)~v >> 31);      // all ones when the leading bit is 0
    return (int32_t)(u - fsh_l(neg, 0u, size));              // negative: u - (2^size - 1)
}

// The image's clean stream as 32-bit words in global memory (device) / host memory.
struct StreamWords
{
    const uint32_t *w;
    B2J_HD uint32_t get(uint32_t i) const
    {
#ifdef __CUDA_ARCH__
        return __ldg(w + i);
#else
        return w[i];
#endif
    }
};

// MSB-first bit reader: a 64-bit window (cur:nxt) and one raw word of look-ahead (requested one refill early).
// Same scheme as BitStream's cached reader (bitstream.h:311-365), 32 bits at a time.
struct WalkReader
{
    uint32_t cur, nxt, raw, bitpos, wi;
    StreamWords src;
    B2J_HD void init(const StreamWords &s, uint32_t byte_off)
    {
        src = s;
        wi = byte_off >> 2;
        cur = bswap32(src.get(wi));
        nxt = bswap32(src.get(wi + 1u));
        raw = src.get(wi + 2u);
        wi += 3u;
        bitpos = (byte_off & 3u) * 8u;
    }
    B2J_HD uint32_t peek() const { return fsh_l(nxt, cur, bitpos); }
    B2J_HD void skip(uint32_t n)   // n <= 32
    {
        bitpos += n;
        if (bitpos >= 32u)
        {
            cur = nxt;
            nxt = bswap32(raw);
            raw = src.get(wi);
            wi++;
            bitpos -= 32u;
        }
    }
};

struct WalkState { uint32_t p, c, z; };

struct WalkResult
{
    uint32_t p, cz, nblk, fs, fc;
    int32_t dc0, dc1, dc2;
};

// Decode tables of one image as the walk sees them. LUT policy: at(i) / at32(i) read the u16 / u32 at u16-offset i of
// the LUT set (shared memory on the device), hdr(i) the i-th header word, ctab(c) the walk's per-block-index table.
//
// ctab(c), c = block index inside the MCU: DC walk table | AC walk table << 15 (u16 offsets in the set, < 32768) |
// component << 30. Built once per image by walk_ctab_entry(), so that the switch to the next block of the MCU is one
// lookup instead of a chain of compares and selects.
template <class LUT>
B2J_HD uint32_t walk_ctab_entry(const LUT &lut, uint32_t c, uint32_t ny, uint32_t nu)
{
    const uint32_t comp = (c >= ny ? 1u : 0u) + (c >= ny + nu ? 1u : 0u);
    return lut.hdr(6 + (int)comp) | lut.hdr(9 + (int)comp) << 15 | comp << 30;
}

// One symbol through the two-level decode tables (entry format: b2j_internal.h). Returns the leaf, 0 = no codeword.
template <class LUT>
B2J_HD uint32_t lookup_symbol(const LUT &lut, uint32_t tab, uint32_t pk, uint32_t bits)
{
    uint32_t e = lut.at(tab + (pk >> (32u - bits)));
    if (!(e & 32u) && e != 0u)
    {
        const uint32_t nb = e & 63u, off = (e >> 6) * kLutSubAlign;
        e = lut.at(tab + (1u << bits) + off + ((pk << bits) >> (32u - nb)));
    }
    return (e & 32u) ? e : 0u;
}

// Walks the stream from state `s` while the next symbol starts before bit `limit`. No output but the result:
// exit state, blocks started (DC symbols met), their DC sums per component and the first block start.
// On a valid stream, from a true state, the walk follows the reference's decoder (decoder.cpp:221-260 inside the
// loops of decoder.cpp:286-346) and ends in its state at the first symbol boundary at or behind `limit`, however the
// symbols were grouped on the way: an AC walk-table step covers several symbols at once, and is taken only while the
// walk is more than kWalkBitsAc - 1 bits in front of `limit` (every symbol of a group starts inside the index window)
// and the block cannot fill up in front of the group's last symbol; otherwise the first symbol of the group is taken
// alone, and where the walk table has no entry, one symbol goes through the decode tables.
// From any state -- true or guessed -- the result is a function of that state alone, which is all the
// synchronisation needs (a guessed walk that reaches a state of the true walk continues exactly like it).
// Shape: ONE loop, one table lookup per turn whatever the lane is at (DC and AC walk tables share the entry format),
// because the lanes of a warp sit at unrelated places of their blocks: what only some lanes need -- the DC
// bookkeeping, the switch to the next block -- is kept short, since every lane pays for it on every turn.
template <class LUT>
B2J_HD WalkResult walk_stream(const StreamWords &stream, const LUT &lut, WalkState s, uint32_t limit, uint32_t tot)
{
    WalkResult r;
    r.nblk = 0; r.fs = kSubNone; r.fc = 0; r.dc0 = r.dc1 = r.dc2 = 0;
    uint32_t p = s.p, c = s.c, z = s.z;
    bool bad = false;
    if (p < limit)
    {
        WalkReader br;
        br.init(stream, p >> 3);
        br.bitpos += p & 7u;
        const uint32_t lim_group = limit > (uint32_t)kWalkBitsAc ? limit - (uint32_t)(kWalkBitsAc - 1) : 0u;
        uint32_t ct = lut.ctab(c);
        uint32_t tdc = ct & 0x7FFFu, tac = (ct >> 15) & 0x7FFFu, comp = ct >> 30;
        while (p < limit)
        {
            const uint32_t pk = br.peek();
            const bool dc = z == 0u;
            const uint32_t e = lut.at32((dc ? tdc : tac) + 2u * (pk >> (dc ? 32u - (uint32_t)kWalkBitsDc : 32u - (uint32_t)kWalkBitsAc)));
            uint32_t nb, f;
            if (e != 0u)
            {
                // the whole group where it fits, else its first symbol (a DC entry is its own first symbol)
                const bool group = p < lim_group && z + ((e >> 5) & 63u) <= 64u;
                const uint32_t ee = group ? e : e >> 12;
                nb = ee & 31u;             // bits consumed
                f = (ee >> 5) & 127u;      // AC: scan positions advanced, >= 64 with the end-of-block flag; DC: category
                B2J_WALK_COUNT(dc ? 3 : (group ? 0 : 1));
            }
            else
            {
                B2J_WALK_COUNT(2);
                const uint32_t e1 = dc ? lookup_symbol(lut, lut.hdr((int)comp), pk, (uint32_t)kLutBitsDc)
                                       : lookup_symbol(lut, lut.hdr(3 + (int)comp), pk, (uint32_t)kLutBits);
                if (e1 == 0u) { bad = true; break; }
                const uint32_t size = (e1 >> 6) & (dc ? 31u : 15u);
                nb = (e1 & 31u) + size;
                f = dc ? size : (e1 >> 10) + 1u;   // AC: zero run + the coefficient, or the extra zero of a size-0 run; EOB: run 63
            }
            uint32_t zn = z + f;
            if (dc)
            {
                // a DC code starts a block: count it, remember the first one, add its difference to the component's sum
                const int32_t diff = extend_bits(pk << (nb - f), f);
                if (r.nblk == 0u) { r.fs = p; r.fc = c; }
                r.nblk++;
                if (comp == 0u) r.dc0 += diff;
                if (comp == 1u) r.dc1 += diff;
                if (comp == 2u) r.dc2 += diff;
                zn = 1u;
            }
            z = zn;
            p += nb;
            br.skip(nb);
            if (z >= 64u)
            {
                // the block is complete: next block of the MCU
                z = 0u;
                c = (c + 1u == tot) ? 0u : c + 1u;
                ct = lut.ctab(c);
                tdc = ct & 0x7FFFu; tac = (ct >> 15) & 0x7FFFu; comp = ct >> 30;
            }
        }
    }
    // A non-code can only be met by a walk that started from a wrong guess (or in a corrupt stream, which the final
    // decode flags): hand the next lane the same guess a first-round walk would use instead of a dead state, so
    // that a wrong guess never poisons the records downstream.
    r.p = bad ? (p > limit ? p : limit) : p;
    r.cz = bad ? 0u : (c | (z << 8));
    return r;
}

// Per-lane scratch of a chunk (shared memory on the device). Index = lane = thread.
struct SyncShared
{
    SubRec cur[kHuffThreads];       // the lane's record as of the latest round (exit state, totals, first block start);
                                    // inside a round, for a lane that goes on to its second half: what its first half found
    SubMid mid[kHuffThreads];       // round 0: state at the middle + the second half's contribution
    uint2 r0_exit[kHuffThreads];    // round 0: exit state (p, cz) -- stands whenever a later walk meets the checkpoint
    uint2 entry_used[kHuffThreads]; // the entry state (p, cz) cur[] was computed from
    uint32_t nq;                    // lanes in q[]
    uint8_t q[kHuffThreads];        // this round's lanes that did not meet their checkpoint: their second halves are walked
                                    // by the first nq threads, so that the few of them fill warps instead of thinning out four
};

// What a chunk is, for every lane alike.
struct SyncChunk
{
    uint32_t first;        // first OUTPUT sub-sequence of the chunk (lane kSyncPre)
    uint32_t n_sub;        // sub-sequences of the image
    uint32_t bits;         // length of the image's clean stream in bits
    uint32_t first_lane;   // first active lane: kSyncPre unless pre-lanes are in use (then max(0, kSyncPre - first))
    bool forced;           // the entry state of lane first_lane is given (true start of the image, or the sweep)
    uint2 forced_entry;    // (p, cz)
};

B2J_HD bool sync_lane_active(const SyncChunk &ch, uint32_t t)
{
    return t >= ch.first_lane && ch.first + t - (uint32_t)kSyncPre < ch.n_sub;
}
B2J_HD uint32_t sync_lane_sub(const SyncChunk &ch, uint32_t t) { return ch.first + t - (uint32_t)kSyncPre; }

B2J_HD void sync_set(SubRec &d, const WalkResult &r)
{
    d.p = r.p; d.cz = r.cz; d.nblk = r.nblk; d.dc[0] = r.dc0; d.dc[1] = r.dc1; d.dc[2] = r.dc2; d.fs = r.fs; d.fc = r.fc;
}

struct SyncBounds { uint32_t md, hi; };
B2J_HD SyncBounds sync_bounds(const SyncChunk &ch, uint32_t t)
{
    const uint32_t lo = sync_lane_sub(ch, t) * (uint32_t)(kSubBytes * 8);
    SyncBounds b;
    b.hi = lo + (uint32_t)(kSubBytes * 8) < ch.bits ? lo + (uint32_t)(kSubBytes * 8) : ch.bits;
    b.md = lo + (uint32_t)(kSubBytes * 4) < b.hi ? lo + (uint32_t)(kSubBytes * 4) : b.hi;
    return b;
}

// First half of lane t's sub-sequence from `entry`. Round 0: the state reached becomes the lane's checkpoint. Later
// rounds: if the walk arrives exactly at the checkpoint, round 0's second half stands -- the record is finished and
// true is returned. Otherwise what the first half found is parked in cur[t] for sync_lane_second().
template <class W>
B2J_HD bool sync_lane_first(const W &w, const SyncChunk &ch, SyncShared &sh, uint32_t t, uint2 entry, bool round0)
{
    const SyncBounds b = sync_bounds(ch, t);
    const WalkState s0 = {entry.x, entry.y & 0xFFu, entry.y >> 8};
    const WalkResult ra = w.walk(s0, b.md);
    sh.entry_used[t] = entry;
    SubRec &c = sh.cur[t];
    if (round0) { sh.mid[t].p = ra.p; sh.mid[t].cz = ra.cz; }
    else
    {
        const SubMid &m = sh.mid[t];
        if (ra.p == m.p && ra.cz == m.cz)
        {
            // in step at the middle: first half of this walk + second half of round 0; the exit state of round 0 stands
            c.p = sh.r0_exit[t].x; c.cz = sh.r0_exit[t].y;
            c.nblk = ra.nblk + m.nblk;
            c.dc[0] = ra.dc0 + m.dc[0]; c.dc[1] = ra.dc1 + m.dc[1]; c.dc[2] = ra.dc2 + m.dc[2];
            c.fs = ra.fs != kSubNone ? ra.fs : m.fs;
            c.fc = ra.fs != kSubNone ? ra.fc : m.fc;
            return true;
        }
    }
    sync_set(c, ra);
    return false;
}

// Second half of lane t's sub-sequence, from where sync_lane_first() stopped.
template <class W>
B2J_HD void sync_lane_second(const W &w, const SyncChunk &ch, SyncShared &sh, uint32_t t, bool round0)
{
    const SyncBounds b = sync_bounds(ch, t);
    SubRec &c = sh.cur[t];
    const WalkState s1 = {c.p, c.cz & 0xFFu, c.cz >> 8};
    const WalkResult rb = w.walk(s1, b.hi);
    if (round0)
    {
        SubMid &m = sh.mid[t];
        m.nblk = rb.nblk; m.dc[0] = rb.dc0; m.dc[1] = rb.dc1; m.dc[2] = rb.dc2; m.fs = rb.fs; m.fc = rb.fc;
        sh.r0_exit[t] = make_uint2(rb.p, rb.cz);
    }
    const bool first_seen = c.nblk != 0u;   // fs is only meaningful when a block started
    c.p = rb.p; c.cz = rb.cz;
    c.dc[0] += rb.dc0; c.dc[1] += rb.dc1; c.dc[2] += rb.dc2;
    if (!first_seen) { c.fs = rb.fs; c.fc = rb.fc; }
    c.nblk += rb.nblk;
}

// Round 0 of lane t: both halves from the guessed state (its border, block 0, DC expected), or from the given state.
template <class W>
B2J_HD void sync_phase_round0(const W &w, const SyncChunk &ch, SyncShared &sh, uint32_t t)
{
    if (!sync_lane_active(ch, t)) return;
    const bool given = ch.forced && t == ch.first_lane;
    const uint2 entry = given ? ch.forced_entry : make_uint2(sync_lane_sub(ch, t) * (uint32_t)(kSubBytes * 8), 0u);
    sync_lane_first(w, ch, sh, t, entry, true);
    sync_lane_second(w, ch, sh, t, true);
}

// Phase "need" of lane t: does the lane have to walk again? Reads the predecessor's exit state BEFORE the round
// (a barrier separates this phase from the next one) and hands it back in `entry`.
B2J_HD bool sync_phase_need(const SyncChunk &ch, const SyncShared &sh, uint32_t t, uint2 &entry)
{
    entry = make_uint2(0u, 0u);
    if (!sync_lane_active(ch, t) || t == ch.first_lane) return false;
    entry = make_uint2(sh.cur[t - 1].p, sh.cur[t - 1].cz);
    return entry.x != sh.entry_used[t].x || entry.y != sh.entry_used[t].y;
}

} // namespace b2j
#endif
