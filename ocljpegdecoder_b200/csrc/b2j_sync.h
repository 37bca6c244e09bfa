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
// predecessor's exit and repairs the (rare) chunks that disagree, in order: their sub-sequences are walked again from
// the true entry state until a walk falls into step with the old records (sync_repair_chunk): correctness never
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

// host emulation only: counts steps: whole AC groups [0], first symbol of a group [1], through the decode tables [2], DC [3], escapes / wide DC [4]
#if defined(B2J_WALK_STATS) && !defined(__CUDACC__)
extern uint64_t g_b2j_walk_steps[5];
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
// bits 8-15 of a word
B2J_HD uint32_t byte1(uint32_t v)
{
#ifdef __CUDA_ARCH__
    return __byte_perm(v, 0, 0x4441);
#else
    return (v >> 8) & 0xFFu;
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

// What the walk needs to know about block c of the MCU (c = block index inside the MCU), one lookup at every block
// switch instead of a chain of compares and selects. Built once per image by walk_ctab_entry().
struct WalkCtab
{
    uint32_t tdc, tac;     // table handles (LUT::tab) of the block's DC and AC walk tables
    uint32_t next;         // block index of the next block of the MCU (wraps to 0)
    uint32_t comp;         // component of the block (the one-symbol path picks its decode tables by it)
    uint32_t m1, m2;       // all ones when the block belongs to component 1 / 2 (DC sums without compares)
    uint32_t pad0, pad1;
};

// Decode tables of one image as the walk sees them. LUT policy:
//   tab(o)     handle of the walk table at u16-offset o of the LUT set (device: its shared-window byte address)
//   ld(h, i)   32-bit entry i of the walk table with handle h
//   ctab(c)    the WalkCtab of block c
//   at(i)      the u16 at u16-offset i of the set, hdr(i) the i-th header word (one-symbol path through the decode tables)
template <class LUT>
B2J_HD WalkCtab walk_ctab_entry(const LUT &lut, uint32_t c, uint32_t ny, uint32_t nu, uint32_t tot)
{
    const uint32_t comp = (c >= ny ? 1u : 0u) + (c >= ny + nu ? 1u : 0u);
    WalkCtab t;
    t.tdc = lut.tab(lut.hdr(6 + (int)comp));
    t.tac = lut.tab(lut.hdr(9 + (int)comp));
    t.next = c + 1u == tot ? 0u : c + 1u;
    t.comp = comp;
    t.m1 = comp == 1u ? 0xFFFFFFFFu : 0u;
    t.m2 = comp == 2u ? 0xFFFFFFFFu : 0u;
    t.pad0 = t.pad1 = 0u;
    return t;
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
// walk is more than kWalkBits - 1 bits in front of `limit` (every symbol of a group starts inside the index window)
// and the block cannot fill up in front of the group's last symbol; otherwise the first symbol of the group is taken
// alone, and where the walk table has no entry, one symbol goes through the decode tables.
// From any state -- true or guessed -- the result is a function of that state alone, which is all the
// synchronisation needs (a guessed walk that reaches a state of the true walk continues exactly like it).
// Shape: ONE loop, one table lookup per turn whatever the lane is at (DC and AC walk tables share index width and
// entry layout; the table in use is a handle that changes behind a DC symbol and at a block switch), because the lanes
// of a warp sit at unrelated places of their blocks: what only some lanes need -- the DC bookkeeping, the switch to
// the next block -- is predicated and short, since the warp pays for it on nearly every turn. The bit position p is
// the only reader state: the 64-bit window (cur:nxt) is word aligned, p & 31 is the shift.
template <class LUT>
B2J_HD WalkResult walk_stream(const StreamWords &stream, const LUT &lut, WalkState s, uint32_t limit)
{
    WalkResult r;
    r.nblk = 0; r.fs = kSubNone; r.fc = 0; r.dc0 = r.dc1 = r.dc2 = 0;
    uint32_t p = s.p, c = s.c, z = s.z;
    bool bad = false;
    if (p < limit)
    {
        uint32_t wi = p >> 5;
        uint32_t cur = bswap32(stream.get(wi)), nxt = bswap32(stream.get(wi + 1u)), raw = stream.get(wi + 2u);
        wi += 3u;
        // two stretches: groups while every symbol of a group starts in front of `limit` (more than kWalkBits - 1
        // bits in front of it), then single symbols up to `limit`
        const uint32_t lim_group = limit > (uint32_t)kWalkBits ? limit - (uint32_t)(kWalkBits - 1) : 0u;
        WalkCtab ct = lut.ctab(c);
        uint32_t tb = z == 0u ? ct.tdc : ct.tac;
        uint32_t fs = z == 0u ? p : kSubNone;     // where the first block of the walk starts (valid if one does)
        r.fc = z == 0u ? c : ct.next;
        int32_t dall = 0, d1 = 0, d2 = 0;
        uint32_t nblk = 0;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
        for (uint32_t stretch = 0; stretch < 2u; stretch++)
        {
        const bool groups = stretch == 0u;
        const uint32_t lim = groups ? lim_group : limit;
        // an AC entry's first symbol is taken alone when the group does not fit the block (position + needed in 65..128)
        // or, in the second stretch, always (position + needed in 1..127); the special entries (needed > 64) stay whole
        const uint32_t alone_lo = groups ? 65u : 1u, alone_span = groups ? 64u : 127u;
        while (p < lim)
        {
            const uint32_t pk = fsh_l(nxt, cur, p);
            const bool dc = z == 0u;
            uint32_t e = lut.ld(tb, pk >> (dc ? 32u - (uint32_t)kWalkBitsDc : 32u - (uint32_t)kWalkBits));
            // the whole group where it fits, else its first symbol (a DC entry is one symbol, its high half the difference)
            const bool whole = dc || z + byte1(e) - alone_lo >= alone_span;
            const int32_t diff = (int32_t)e >> 16;
            if (!whole) e >>= 16;
            uint32_t nb = e & 31u;                // bits consumed
            uint32_t adv = byte1(e);              // scan positions advanced (a DC symbol: 1)
            uint32_t eob = e & 0x80u;             // the set ends with the end-of-block symbol
            int32_t dv = diff;
            if (adv > 64u)
            {
                B2J_WALK_COUNT(4);
                if (adv == kWalkEscape)
                {
                    // AC code longer than the index: one symbol from the prefix's sub-table
                    const uint32_t e2 = lut.ld(tb, (e >> 16) + ((pk << (uint32_t)kWalkBits) >> (32u - nb)));
                    nb = e2 & 31u; adv = byte1(e2); eob = e2 & 0x80u;
                }
                else if (adv == kWalkDcWide)
                {
                    // DC value bits past the index
                    const uint32_t size = (e >> 16) & 31u;
                    dv = extend_bits(pk << nb, size);
                    nb += size; adv = 1u;
                }
                if (adv > 64u)
                {
                    // no walk-table entry: one symbol through the decode tables, or no codeword at all
                    B2J_WALK_COUNT(2);
                    const uint32_t e1 = dc ? lookup_symbol(lut, lut.hdr((int)ct.comp), pk, (uint32_t)kLutBitsDc)
                                           : lookup_symbol(lut, lut.hdr(3 + (int)ct.comp), pk, (uint32_t)kLutBits);
                    if (e1 == 0u)
                    {
                        // a block whose DC code is no codeword still counts as started: the decode lane meets the same bits and flags the image
                        if (dc) nblk++;
                        bad = true; stretch = 2u; break;
                    }
                    const uint32_t len = e1 & 31u, size = (e1 >> 6) & (dc ? 31u : 15u), run = e1 >> 10;
                    nb = len + size;
                    eob = (!dc && run == kRunEob) ? 0x80u : 0u;
                    adv = (dc || eob) ? 1u : run + 1u;   // AC: zero run + the coefficient, or the extra zero of a size-0 run
                    dv = extend_bits(pk << len, size);
                }
            }
            else B2J_WALK_COUNT(dc ? 3 : (whole ? 0 : 1));
            if (dc)
            {
                // a DC code starts a block: count it, add its difference to the sums; the AC table takes over
                nblk++;
                dall += dv; d1 += dv & (int32_t)ct.m1; d2 += dv & (int32_t)ct.m2;
                tb = ct.tac;
            }
            z += adv;
            const uint32_t pn = p + nb;
            if ((p ^ pn) & 32u)
            {
                cur = nxt;
                nxt = bswap32(raw);
                raw = stream.get(wi);
                wi++;
            }
            p = pn;
            if (eob != 0u || z >= 64u)
            {
                // the block is complete: next block of the MCU
                z = 0u;
                c = ct.next;
                ct = lut.ctab(c);
                tb = ct.tdc;
                fs = fs < p ? fs : p;
            }
        }
        }
        r.nblk = nblk;
        r.fs = nblk ? fs : kSubNone;
        r.dc0 = dall - d1 - d2; r.dc1 = d1; r.dc2 = d2;
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

// Repair of a chunk whose assumed entry state was not its predecessor's exit state (k_sync_sweep; one thread): the
// sub-sequences of the chunk are walked again in order from the true entry state until a walk leaves its sub-sequence
// in the state the old record left it in -- from there on the old records stand, since every record is a function of
// its entry state alone. tot4: the chunk's totals (blocks, DC sums), kept in step with the records. Returns the
// chunk's exit state in exit_pcz (unchanged when the walks merged inside the chunk).
template <class W>
B2J_HD void sync_repair_chunk(const W &w, uint32_t chunk, uint32_t n_sub, uint32_t bits, SubRec *rec_img, uint2 entry,
                              uint32_t tot4[4], uint2 &exit_pcz)
{
    const uint32_t first = chunk * (uint32_t)kSyncLanes;
    const uint32_t end = first + (uint32_t)kSyncLanes < n_sub ? first + (uint32_t)kSyncLanes : n_sub;
    for (uint32_t j = first; j < end; j++)
    {
        const SubRec old = rec_img[j];
        const uint32_t lo = j * (uint32_t)(kSubBytes * 8);
        const uint32_t hi = lo + (uint32_t)(kSubBytes * 8) < bits ? lo + (uint32_t)(kSubBytes * 8) : bits;
        const WalkState s0 = {entry.x, entry.y & 0xFFu, entry.y >> 8};
        const WalkResult r = w.walk(s0, hi);
        tot4[0] += r.nblk - old.nblk;
        tot4[1] += (uint32_t)(r.dc0 - old.dc[0]); tot4[2] += (uint32_t)(r.dc1 - old.dc[1]); tot4[3] += (uint32_t)(r.dc2 - old.dc[2]);
        sync_set(rec_img[j], r);
        if (r.p == old.p && r.cz == old.cz) return;   // in step with the old walks: everything behind stands
        entry = make_uint2(r.p, r.cz);
    }
    exit_pcz = entry;
}

} // namespace b2j
#endif
