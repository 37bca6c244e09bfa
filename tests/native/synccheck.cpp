// tests/native/synccheck.cpp -- TEST INFRASTRUCTURE. Host emulation of the self-synchronising entropy path:
// the walk (walk_stream, with the multi-symbol walk tables) and the chunk-wise synchronisation rounds of
// ocljpegdecoder_b200/csrc/b2j_sync.h are the same inline code the kernels compile; here their phases run lane after
// lane on the CPU and every sub-sequence record is compared with a sequential, table-free one-symbol-at-a-time walk
// of the same stream (the reference's scan order, decoder.cpp:221-346). Nothing here is a decode path.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define B2J_WALK_STATS 1
#include "b2j_internal.h"
#include "b2j_sync.h"

uint64_t g_b2j_walk_steps[5];

using namespace b2j;

namespace {

struct HostLut
{
    const uint16_t *v;
    WalkCtab ct[16];
    uint32_t at(uint32_t i) const { return v[i]; }
    uint32_t tab(uint32_t off16) const { return off16; }
    uint32_t ld(uint32_t h, uint32_t i) const { return (uint32_t)v[h + 2 * i] | (uint32_t)v[h + 2 * i + 1] << 16; }
    uint32_t hdr(int i) const { return v[i]; }
    WalkCtab ctab(uint32_t c) const { return ct[c]; }
};

struct HostWalk
{
    StreamWords stream;
    HostLut lut;
    uint32_t tot, ny, nu;
    void init() { for (uint32_t c = 0; c < tot && c < 16; c++) lut.ct[c] = walk_ctab_entry(lut, c, ny, nu, tot); }
    WalkResult walk(WalkState s, uint32_t limit) const { return walk_stream(stream, lut, s, limit); }
};

// ---- the independent check: canonical codes matched bit by bit, one symbol per step ----------------------------
struct Canon
{
    struct Code { uint32_t code; int len, sym; };
    std::vector<Code> codes;
    void build(const uint8_t counts[16], const uint8_t *symbols)
    {
        uint32_t code = 0;
        int n = 0;
        for (int l = 1; l <= 16; l++)
        {
            for (int i = 0; i < counts[l - 1]; i++) { codes.push_back({code, l, symbols[n]}); n++; code++; }
            code <<= 1;
        }
    }
};

struct PlainBits
{
    const uint8_t *p;
    uint32_t bit(uint64_t i) const { return (p[i >> 3] >> (7 - (i & 7))) & 1u; }
    uint32_t bits(uint64_t i, int n) const { uint32_t v = 0; for (int k = 0; k < n; k++) v = (v << 1) | bit(i + k); return v; }
};

// returns the symbol and its code length, or -1
int canon_symbol(const Canon &t, const PlainBits &b, uint64_t pos, int *len)
{
    uint32_t acc = 0;
    size_t k = 0;
    for (int l = 1; l <= 16; l++)
    {
        acc = (acc << 1) | b.bit(pos + l - 1);
        while (k < t.codes.size() && t.codes[k].len < l) k++;
        for (size_t j = k; j < t.codes.size() && t.codes[j].len == l; j++)
            if (t.codes[j].code == acc) { *len = l; return t.codes[j].sym; }
    }
    return -1;
}

struct RefImage
{
    Canon dc[3], ac[3];
    uint32_t tot, ny, nu;
};

WalkResult ref_walk(const RefImage &im, const PlainBits &b, WalkState s, uint32_t limit)
{
    WalkResult r;
    r.nblk = 0; r.fs = kSubNone; r.fc = 0; r.dc0 = r.dc1 = r.dc2 = 0;
    uint32_t p = s.p, c = s.c, z = s.z;
    bool bad = false;
    while (p < limit)
    {
        const uint32_t comp = (c >= im.ny ? 1u : 0u) + (c >= im.ny + im.nu ? 1u : 0u);
        int len = 0;
        if (z == 0)
        {
            const int sym = canon_symbol(im.dc[comp], b, p, &len);
            if (sym < 0 || sym > 16)
            {
                if (r.nblk == 0) { r.fs = p; r.fc = c; }
                r.nblk++;   // counted as started (the decode lane flags it)
                bad = true; break;
            }
            const uint32_t v = sym ? b.bits(p + len, sym) : 0u;
            const int32_t diff = sym == 0 ? 0 : ((v >> (sym - 1)) ? (int32_t)v : (int32_t)v + 1 - (1 << sym));   // decoder.cpp:72-82
            if (getenv("B2J_SYNCCHECK_TRACE")) fprintf(stderr, "  ref dc p %u len %d sym %d diff %d comp %u\n", p, len, sym, diff, comp);
            if (r.nblk == 0) { r.fs = p; r.fc = c; }
            r.nblk++;
            if (comp == 0) r.dc0 += diff; else if (comp == 1) r.dc1 += diff; else r.dc2 += diff;
            p += (uint32_t)(len + sym);
            z = 1;
        }
        else
        {
            const int sym = canon_symbol(im.ac[comp], b, p, &len);
            if (sym < 0) { bad = true; break; }
            const int run = sym >> 4, size = sym & 15;
            p += (uint32_t)(len + size);
            if (sym == 0) z = 64; else z += (uint32_t)run + 1u;   // decoder.cpp:241-256
        }
        if (z >= 64) { z = 0; c = (c + 1 == im.tot) ? 0 : c + 1; }
    }
    r.p = bad ? (p > limit ? p : limit) : p;
    r.cz = bad ? 0u : (c | (z << 8));
    return r;
}

bool same(const SubRec &a, const WalkResult &r)
{
    return a.p == r.p && a.cz == r.cz && a.nblk == r.nblk && a.dc[0] == r.dc0 && a.dc[1] == r.dc1 && a.dc[2] == r.dc2 &&
           (r.nblk == 0 || (a.fs == r.fs && a.fc == r.fc));
}

} // namespace

extern "C" {

// Checks one baseline JPEG without restart markers. Returns 0 when everything agrees, a negative code otherwise.
// stats[0] sub-sequences, [1] chunks, [2] lanes that walked again in round 1, [3] lanes that missed the checkpoint in
// round 1, [4] rounds of the slowest chunk, [5] chunks the sweep re-ran, AC steps of the walks of part 1 (two passes over the
// stream): [6] whole groups, [7] first symbol of a group, [8] through the decode tables, [9] sub-sequences whose walk from the true state disagrees (must be 0).
// pre_lanes < 0: kSyncPre. force_sweep: pre-lanes off, so that every chunk border goes through the sweep.
int b2j_synccheck(const uint8_t *file, size_t len, int gate, int force_sweep, uint64_t *stats)
{
    memset(stats, 0, 10 * sizeof(uint64_t));
    g_b2j_walk_steps[0] = g_b2j_walk_steps[1] = g_b2j_walk_steps[2] = g_b2j_walk_steps[3] = g_b2j_walk_steps[4] = 0;
    b2j_image_desc d;
    int rc = b2j_parse_header(file, len, gate, &d);
    if (rc != B2J_OK) return -100 + rc;
    if (d.restart_interval != 0) return -2;
    // unstuff (decoder.cpp:94-159): FF00 -> FF, stop at the first marker
    std::vector<uint8_t> clean;
    for (size_t i = d.scan_offset; i < len; i++)
    {
        if (file[i] != 0xFF) { clean.push_back(file[i]); continue; }
        if (i + 1 < len && file[i + 1] == 0x00) { clean.push_back(0xFF); i++; continue; }
        if (i + 1 < len && file[i + 1] == 0xFF) continue;
        break;
    }
    const uint32_t clean_len = (uint32_t)clean.size();
    clean.resize(clean.size() + 64, 0);
    while (clean.size() & 3) clean.push_back(0);
    std::vector<uint16_t> set;
    if (!build_lut_set(d, set)) return -3;

    RefImage ref;
    ref.tot = (uint32_t)d.tot_blks_per_mcu; ref.ny = (uint32_t)d.blks_per_mcu[0]; ref.nu = (uint32_t)d.blks_per_mcu[1];
    if (d.tot_blks_per_mcu == 1) { ref.nu = 0; }
    for (int c = 0; c < 3; c++)
    {
        ref.dc[c].build(d.huff_counts[d.huff_id[c] >> 4], d.huff_symbols[d.huff_id[c] >> 4]);
        ref.ac[c].build(d.huff_counts[4 + (d.huff_id[c] & 15)], d.huff_symbols[4 + (d.huff_id[c] & 15)]);
    }
    const PlainBits pb = {clean.data()};
    const uint32_t bits = clean_len * 8u;
    const uint32_t n_sub = (bits + kSubBytes * 8 - 1) / (kSubBytes * 8);
    stats[0] = n_sub;

    HostWalk w;
    w.stream.w = reinterpret_cast<const uint32_t *>(clean.data());
    w.lut.v = set.data();
    w.tot = ref.tot; w.ny = ref.ny; w.nu = ref.nu ? ref.nu : 1u;
    w.init();

    // ---- 1. the truth, sequentially; and the walk from the true state of every sub-sequence, whole and in halves
    std::vector<WalkResult> truth(n_sub);
    WalkState st = {0, 0, 0};
    for (uint32_t s = 0; s < n_sub; s++)
    {
        const uint32_t lo = s * (uint32_t)(kSubBytes * 8), hi = lo + kSubBytes * 8 < bits ? lo + kSubBytes * 8 : bits;
        const uint32_t md = lo + kSubBytes * 4 < hi ? lo + kSubBytes * 4 : hi;
        truth[s] = ref_walk(ref, pb, st, hi);
        const WalkResult whole = w.walk(st, hi);
        const WalkResult ra_ref = ref_walk(ref, pb, st, md), ra = w.walk(st, md);
        const WalkState sm = {ra.p, ra.cz & 0xFFu, ra.cz >> 8};
        const WalkResult rb = w.walk(sm, hi);
        SubRec a;
        sync_set(a, whole);
        bool ok = same(a, truth[s]);
        ok = ok && ra.p == ra_ref.p && ra.cz == ra_ref.cz && rb.p == truth[s].p && rb.cz == truth[s].cz &&
             ra.nblk + rb.nblk == truth[s].nblk && ra.dc0 + rb.dc0 == truth[s].dc0;
        if (!ok)
        {
            if (!stats[9] && getenv("B2J_SYNCCHECK_VERBOSE"))
                fprintf(stderr, "sub %u entry (%u,%u,%u) hi %u md %u: truth p %u cz %x nblk %u dc %d | whole p %u cz %x nblk %u dc %d | ra %u %x ref %u %x rb %u %x\n", s, st.p, st.c, st.z,
                        hi, md, truth[s].p, truth[s].cz, truth[s].nblk, truth[s].dc0, whole.p, whole.cz, whole.nblk, whole.dc0, ra.p, ra.cz, ra_ref.p, ra_ref.cz, rb.p, rb.cz);
            stats[9]++;
        }
        st.p = truth[s].p; st.c = truth[s].cz & 0xFFu; st.z = truth[s].cz >> 8;
    }
    if (stats[9]) return -4;

    stats[6] = g_b2j_walk_steps[0]; stats[7] = g_b2j_walk_steps[1]; stats[8] = g_b2j_walk_steps[2];
    // ---- 2. the chunk-wise synchronisation, as the kernels run it
    const uint32_t n_chunks = (n_sub + kSyncLanes - 1) / kSyncLanes;
    stats[1] = n_chunks;
    std::vector<SubRec> recs(n_sub);
    std::vector<uint2> chunk_entry(n_chunks), chunk_exit(n_chunks);
    SyncShared *sh = new SyncShared;
    auto run_chunk = [&](uint32_t k, bool forced, uint2 entry) {
        SyncChunk ch;
        ch.first = k * kSyncLanes; ch.n_sub = n_sub; ch.bits = bits;
        const bool pre = !forced && !force_sweep && k > 0;
        ch.first_lane = pre ? (ch.first >= (uint32_t)kSyncPre ? 0u : kSyncPre - ch.first) : (uint32_t)kSyncPre;
        ch.forced = forced || k == 0;
        ch.forced_entry = k == 0 ? make_uint2(0u, 0u) : entry;
        memset(sh, 0xEE, sizeof(*sh));
        for (uint32_t t = 0; t < (uint32_t)kHuffThreads; t++) sync_phase_round0(w, ch, *sh, t);
        for (uint32_t round = 1;; round++)
        {
            // the phases of a round, each for every lane before the next one starts (barriers on the device)
            bool need[kHuffThreads], any = false;
            uint2 ent[kHuffThreads];
            sh->nq = 0;
            for (uint32_t t = 0; t < (uint32_t)kHuffThreads; t++) { need[t] = sync_phase_need(ch, *sh, t, ent[t]); any = any || need[t]; }
            if (!any) break;
            if (round > stats[4]) stats[4] = round;
            for (uint32_t t = 0; t < (uint32_t)kHuffThreads; t++)
                if (need[t])
                {
                    const bool met = sync_lane_first(w, ch, *sh, t, ent[t], false);
                    if (!met) sh->q[sh->nq++] = (uint8_t)t;
                    if (round == 1 && !forced) { stats[2]++; if (!met) stats[3]++; }
                }
            for (uint32_t i = 0; i < sh->nq; i++) sync_lane_second(w, ch, *sh, sh->q[i], false);
            if (round > (uint32_t)kHuffThreads + 1) return false;   // cannot happen: lane k is final after k rounds
        }
        uint32_t last = kSyncPre;
        for (uint32_t t = kSyncPre; t < (uint32_t)kHuffThreads; t++)
            if (sync_lane_active(ch, t)) { recs[sync_lane_sub(ch, t)] = sh->cur[t]; last = t; }
        chunk_entry[k] = sh->entry_used[kSyncPre];
        chunk_exit[k] = make_uint2(sh->cur[last].p, sh->cur[last].cz);
        return true;
    };
    for (uint32_t k = 0; k < n_chunks; k++)
        if (!run_chunk(k, false, make_uint2(0, 0))) { delete sh; return -5; }
    // the sweep: in order, repair what started from a state its predecessor did not end in (sync_repair_chunk, as the kernel)
    for (uint32_t k = 1; k < n_chunks; k++)
        if (chunk_entry[k].x != chunk_exit[k - 1].x || chunk_entry[k].y != chunk_exit[k - 1].y)
        {
            stats[5]++;
            uint32_t t4[4] = {0, 0, 0, 0};
            sync_repair_chunk(w, k, n_sub, bits, recs.data(), chunk_exit[k - 1], t4, chunk_exit[k]);
            chunk_entry[k] = chunk_exit[k - 1];
        }
    delete sh;
    int bad = 0;
    for (uint32_t s = 0; s < n_sub; s++)
        if (!same(recs[s], truth[s])) bad++;
    return bad ? -6 : 0;
}

} // extern "C"
