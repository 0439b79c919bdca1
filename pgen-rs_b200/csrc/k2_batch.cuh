// k2_batch.cuh — K2 for launches whose lines are short enough that a CTA can stage a batch of
// B consecutive lines in shared memory (the gather-heavy chr22 shape: 250 kept samples of
// 2 504 -> 1 041-byte lines; short keep-all lines).  Like k2_core.cuh the source compiles
// both as sm_100a device code and, with PGB_HOSTSIM, as a thread-by-thread host simulation
// used ONLY by tests/.
//
// What it replaces: the same loop body as k2_core.cuh (/root/reference/src/pfile.rs:156-192),
// organised the way BASELINE.json's north_star (b)+(c) describes it.  The kernel is persistent
// (CTA c takes batches c, c + gridDim, ...), warp-specialised and software-pipelined over two or
// three shared-memory input stages (pgb_kernels.cu: k2_batch_kernel):
//   1. INPUT (producer warp, one lane per line of the batch): the pgb_line_meta of the next batch
//      is already in registers; the lane writes the stage's line table and hands the
//      16-byte-aligned covering range of its record (records + rec_off: pfile.rs:165-170 as a TMA
//      transfer) and the batch's prefix bytes to the bulk-copy engine (cp.async.bulk.shared.global,
//      completion on the stage's `full` mbarrier);
//   2. GATHER (consumer warps, a warp per line; pfile.rs:171-175): the kept samples of the staged
//      record are compacted into a packed 2-bit "virtual record" of ceil(K/4) bytes.  One lane per
//      OUTPUT byte; the lane's four source positions (byte offset, bit shift) come from the K0
//      index list and stay in registers (or shared memory) for the whole kernel — 4 LDS.U8 +
//      shift/mask per byte, no warp reductions;
//   3. FORMAT (pfile.rs:177-190): the keep-all chunk body runs on the virtual record, two 16-byte
//      chunks per lane and iteration;
//   4. OUTPUT: either straight to global memory (16-byte .cs stores for the aligned chunks, byte
//      stores for prefixes, newlines and the <= 15 + 15 ragged GT bytes of a line), or into an
//      image of the warp's contiguous output range in shared memory that leaves with ONE bulk
//      async store per warp and batch (cp.async.bulk.global.shared::cta; the <= 15 bytes on either
//      end are byte stores).  The consumer warps never wait for one another.
// Every output byte still has exactly one writer and no padding is emitted.
//
// Text decode uses a 16-entry table (a nibble = two genotypes -> 8 bytes of text, 128 bytes
// in all): lanes reading the same entry broadcast and different entries live in different
// banks, so the two LDS.64 of a chunk are bank-conflict-free by construction.
#pragma once
#include "k2_core.cuh"

#define K2B_THREADS 288 // 8 consumer warps + 1 producer warp
#define K2B_WARPS 8     // consumer warps
#define K2B_CONSUMERS 256
#define K2B_TAB_LINES 128u // a stage table: 8 x 16 bytes of per-warp headers, then 32 bytes per line

struct pgb_k2b_params {
    const uint8_t *records;
    const pgb_line_meta *meta;
    const uint8_t *prefix_blob;
    const uint32_t *kidx; // nullptr => keep all
    uint8_t *out;
    uint64_t n_lines;
    uint32_t n_batches;
    uint32_t K;
    uint32_t R;        // record bytes
    uint32_t B;        // lines per batch, <= 32 (one producer lane per line)
    uint32_t rowcap;   // shared-memory bytes per staged record: align16(R + 31)
    uint32_t pcap;     // shared-memory bytes per staged prefix: align16(max prefix + 31)
    uint32_t vcap;     // shared-memory bytes per virtual record (gather): align16(ceil(K/4) + 2)
    uint32_t wcap;     // shared-memory bytes of one consumer warp's part of an output image: align128(LPW * max_line + 32)
    uint32_t outcap;   // shared-memory bytes of one output image: K2B_WARPS * wcap
    uint32_t sfx;      // bytes appended to every prefix after its blob bytes (little-endian), e.g. "\tGT"
    uint32_t sfx_len;  // 0..4; pfx_len includes it
    uint32_t kidx_vec; // kidx is 16-byte aligned
    uint32_t images;     // 0: no image — the consumer warps store to global memory directly; 1: an image per warp,
                         // bulk-stored (the warp waits for the drain before the next batch); 2: two images per warp
    uint32_t stages;     // input stages (records, prefixes, table): 2 or 3
};

#if defined(PGB_HOSTSIM)
#define PGB_HD static inline
#define PGB_HDM inline
#else
#define PGB_HD __host__ __device__ __forceinline__
#define PGB_HDM __host__ __device__ __forceinline__
#endif

// Shared memory: [0,64) the mbarriers (full[s] at 8 s: a stage's copies have landed; empty[s] at 32 + 8 s: the
// consumer warps are done with a stage), the text table, then per stage s a table (16 bytes per consumer warp: body offset and
// image byte range of the warp's lines; 32 bytes per line, see k2b_produce), the staged records and prefixes;
// the gather plan; one virtual record per consumer warp; two images (each one part per consumer warp).
struct pgb_k2b_layout {
    uint32_t lut, tab0, tabsz, rows0, rowsz, pst0, pstsz, plan, vrec, outb0, outsz, total;
    // offsets of stage s / image i (plain arithmetic: indexing an array member with a run-time stage would put
    // the struct in local memory)
    PGB_HDM uint32_t tab(uint32_t s) const { return tab0 + s * tabsz; }
    PGB_HDM uint32_t rows(uint32_t s) const { return rows0 + s * rowsz; }
    PGB_HDM uint32_t pst(uint32_t s) const { return pst0 + s * pstsz; }
    PGB_HDM uint32_t outb(uint32_t i) const { return outb0 + i * outsz; }
};

PGB_HD uint32_t pgb_k2b_align(uint32_t x, uint32_t a) { return (x + a - 1u) & ~(a - 1u); }

PGB_HD pgb_k2b_layout pgb_k2b_smem_layout(uint32_t B, uint32_t rowcap, uint32_t pcap, uint32_t vcap, uint32_t outcap,
                                          bool gather, uint32_t images = 2, uint32_t stages = 2) {
    pgb_k2b_layout L;
    L.lut = 64;
    L.tabsz = K2B_TAB_LINES + 32u * B;
    L.rowsz = B * rowcap;
    L.pstsz = B * pcap;
    L.tab0 = L.lut + 128;
    L.rows0 = L.tab0 + stages * L.tabsz;
    L.pst0 = L.rows0 + stages * L.rowsz;
    L.plan = L.pst0 + stages * L.pstsz;
    L.vrec = L.plan + (gather ? vcap * 16u : 0u); // 16 bytes of plan per virtual-record byte (vcap >= ceil(K/4))
    L.outb0 = pgb_k2b_align(L.vrec + (gather ? K2B_WARPS * vcap : 0u), 128);
    L.outsz = images > 1 ? outcap : 0u;
    L.total = L.outb0 + images * outcap;
    return L;
}

struct pgb_u2 {
    uint32_t x, y;
};

#if defined(PGB_HOSTSIM)
PGB_DEV void k2b_sts16(uint8_t *d, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    uint32_t t[4] = {x, y, z, w};
    memcpy(d, t, 16);
}
PGB_DEV pgb_u2 k2b_lds8(const uint8_t *s) {
    pgb_u2 r;
    memcpy(&r, s, 8);
    return r;
}
PGB_DEV uint32_t pgb_funnel_l(uint32_t lo, uint32_t hi, uint32_t sh) {
    sh &= 31u;
    return sh ? ((hi << sh) | (lo >> (32u - sh))) : hi;
}
PGB_DEV uint64_t k2b_ld_u64(const uint64_t *p) { return *p; }
// bulk copies global -> "shared": 16-byte-aligned sources, sizes multiples of 16
PGB_DEV void k2b_stage_load(uint8_t *, uint8_t *rdst, const uint8_t *rsrc, uint32_t rbytes, uint8_t *pdst,
                            const uint8_t *psrc, uint32_t pbytes) {
    if (rbytes) memcpy(rdst, rsrc, rbytes);
    if (pbytes) memcpy(pdst, psrc, pbytes);
}
PGB_DEV void k2b_arrive(uint8_t *) {}
#else
PGB_DEV void k2b_sts16(uint8_t *d, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    *reinterpret_cast<uint4 *>(d) = make_uint4(x, y, z, w);
}
PGB_DEV pgb_u2 k2b_lds8(const uint8_t *s) {
    const uint2 v = *reinterpret_cast<const uint2 *>(s);
    pgb_u2 r = {v.x, v.y};
    return r;
}
PGB_DEV uint32_t pgb_funnel_l(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_l(lo, hi, sh); }
PGB_DEV uint64_t k2b_ld_u64(const uint64_t *p) { return __ldg(reinterpret_cast<const unsigned long long *>(p)); }
PGB_DEV uint32_t k2b_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
PGB_DEV void k2b_arrive(uint8_t *mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(k2b_smem_addr(mbar)) : "memory");
}
// One line of the next batch: arm the stage's mbarrier with the byte count, then hand the
// 16-byte-aligned covering ranges of its record and of its prefix to the bulk-copy engine
// (SASS: SYNCS.ARRIVE.TRANS64 + UBLKCP).
PGB_DEV void k2b_stage_load(uint8_t *mbar, uint8_t *rdst, const uint8_t *rsrc, uint32_t rbytes, uint8_t *pdst,
                            const uint8_t *psrc, uint32_t pbytes) {
    const uint32_t bar = k2b_smem_addr(mbar);
    if (rbytes + pbytes == 0) {
        k2b_arrive(mbar);
        return;
    }
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(rbytes + pbytes) : "memory");
    if (rbytes)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         k2b_smem_addr(rdst)),
                     "l"(rsrc), "r"(rbytes), "r"(bar)
                     : "memory");
    if (pbytes)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         k2b_smem_addr(pdst)),
                     "l"(psrc), "r"(pbytes), "r"(bar)
                     : "memory");
}
#endif

// Lines of a batch per consumer warp: warp w formats lines [w * LPW, (w + 1) * LPW) — a contiguous byte range of
// the output, which the warp stores on its own (no CTA-wide barrier anywhere in the steady state).
PGB_HD uint32_t k2b_lines_per_warp(uint32_t B) { return (B + K2B_WARPS - 1u) / K2B_WARPS; }

// ---- PRODUCER (one warp): stage table + record/prefix fetch of one batch.  Lane <-> line i0 + lane (B <= 32);
//      `m` is that line's pgb_line_meta, `first_off` the body offset of the first line of the lane's consumer
//      warp (a shuffle on the device), `next_off` the body offset of the line after the lane's, `base` / `end_off`
//      the body offsets of the batch's first line / end, `pfx0` the prefix offset of its first line.  Every lane
//      arrives once on the stage's `full` barrier (after its table stores: the arrive releases them to the
//      consumers).  `packed`: K1 marked the prefixes as packed back to back (pgb_line_meta::reserved bit 0), so
//      the batch's prefixes are one byte range and travel in ONE bulk copy, issued by lane 0.
//      Table, per consumer warp (16 bytes): body offset of its first line (u64), image offsets of its first byte
//      and past its last.  Per line (32 bytes): [0] image offset of the line start, [1] prefix bytes to copy,
//      [2] shared-memory offset of the staged prefix, [3] image offset of the newline, [4] shared-memory offset of
//      the staged record span, [5] image offset of the GT text, [6]/[7] image offsets of the first / past-the-last
//      aligned chunk.  Image offsets count from the stage's image; a warp's part starts at w * wcap and keeps the
//      16-byte phase of the global address it will be stored to. ----
PGB_DEV void k2b_produce(const pgb_k2b_params &p, uint8_t *smem, const pgb_k2b_layout &L, uint32_t stage, uint32_t nbl,
                         uint32_t lane, const pgb_line_meta &m, uint64_t first_off, uint64_t next_off, uint64_t base,
                         uint64_t end_off, uint64_t pfx0, bool packed, uint32_t span_lo, uint32_t span_len) {
    uint8_t *mbar = smem + 8u * stage;
    uint8_t *tab = smem + L.tab(stage);
    const uint8_t *p0 = p.prefix_blob + pfx0;
    const uint32_t pph0 = (uint32_t)(uintptr_t)p0 & 15u;
    uint8_t *pdst = nullptr;
    const uint8_t *psrc = nullptr;
    uint32_t pbytes = 0;
    if (lane == 0) {
        // all prefix bytes of the batch: its size minus the GT text, newlines and suffixes
        const uint32_t D = (uint32_t)(end_off - base) - nbl * (4u * p.K + 1u + p.sfx_len);
        if (packed && D) {
            pdst = smem + L.pst(stage);
            psrc = p0 - pph0;
            pbytes = (pph0 + D + 15u) & ~15u;
        }
    }
    const uint32_t LPW = k2b_lines_per_warp(p.B);
    if (lane < K2B_WARPS && lane * LPW >= nbl) reinterpret_cast<uint32_t *>(tab + 16u * lane)[2] = 0xFFFFFFFFu; // idle warp
    if (lane >= nbl) {
        k2b_arrive(mbar);
        return;
    }
    const uint32_t w = lane / LPW;
    const uint32_t phase = (uint32_t)((uint64_t)(uintptr_t)p.out + first_off) & 15u;
    const uint32_t o_ls = w * p.wcap + phase + (uint32_t)(m.line_off - first_off);
    if (lane == w * LPW) { // first line of a consumer warp
        *reinterpret_cast<uint64_t *>(tab + 16u * w) = first_off;
        reinterpret_cast<uint32_t *>(tab + 16u * w)[2] = o_ls;
    }
    if (lane + 1u == nbl || lane + 1u == (w + 1u) * LPW) // its last line
        reinterpret_cast<uint32_t *>(tab + 16u * w)[3] = o_ls + (uint32_t)(next_off - m.line_off);
    uint32_t *d = reinterpret_cast<uint32_t *>(tab + K2B_TAB_LINES + 32u * lane);
    const uint32_t dlen = m.pfx_len - p.sfx_len;
    const uint32_t o_gs = o_ls + m.pfx_len, o_ge = o_gs + 4u * p.K;
    const uint8_t *src = p.records + m.rec_off + span_lo;
    const uint32_t ph = (uint32_t)(uintptr_t)src & 15u;
    uint32_t pst_off;
    if (packed) {
        pst_off = pph0 + (uint32_t)(m.pfx_off - pfx0);
    } else {
        const uint8_t *q = p.prefix_blob + m.pfx_off;
        const uint32_t pph = dlen ? (uint32_t)(uintptr_t)q & 15u : 0u;
        pst_off = lane * p.pcap + pph;
        if (dlen) {
            pdst = smem + L.pst(stage) + lane * p.pcap;
            psrc = q - pph;
            pbytes = (pph + dlen + 15u) & ~15u;
        }
    }
    d[0] = o_ls;
    d[1] = dlen;
    d[2] = L.pst(stage) + pst_off;
    d[3] = o_ge;
    d[4] = L.rows(stage) + lane * p.rowcap + ph;
    d[5] = o_gs;
    d[6] = (o_gs + 15u) & ~15u;
    d[7] = o_ge & ~15u;
    k2b_stage_load(mbar, smem + L.rows(stage) + lane * p.rowcap, src - ph, span_len ? (ph + span_len + 15u) & ~15u : 0u, pdst,
                   psrc, pbytes);
}

// The gather plan of virtual-record byte j: its four kept samples as (source byte relative to the staged span) << 5
// | (left shift bringing the sample's two bits to [6 + 2k, 8 + 2k)).  The same for every variant (the kept-sample
// list does not depend on the variant, pfile.rs:128,171): built once per CTA into shared memory.
PGB_DEV void k2b_build_plan(const pgb_k2b_params &p, uint8_t *smem, const pgb_k2b_layout &L, uint32_t tid, uint32_t span_lo) {
    const uint32_t nb = (p.K + 3u) >> 2;
    uint32_t *plan = reinterpret_cast<uint32_t *>(smem + L.plan);
    for (uint32_t j = tid; j < nb; j += K2B_CONSUMERS) {
        uint32_t s0, s1, s2, s3;
        if (p.kidx_vec) {
            const pgb_u4 v = pgb_ld128(p.kidx + 4u * j); // kidx carries 8 entries of padding
            s0 = v.x; s1 = v.y; s2 = v.z; s3 = v.w;
        } else {
            s0 = pgb_ld32(p.kidx + 4u * j); s1 = pgb_ld32(p.kidx + 4u * j + 1); s2 = pgb_ld32(p.kidx + 4u * j + 2);
            s3 = pgb_ld32(p.kidx + 4u * j + 3);
        }
        if (4u * j + 1u >= p.K) s1 = s0; // fields past K (last byte only): any staged byte will do
        if (4u * j + 2u >= p.K) s2 = s0;
        if (4u * j + 3u >= p.K) s3 = s0;
        plan[4u * j + 0u] = ((s0 >> 2) - span_lo) << 5 | (6u - (s0 & 3u) * 2u);
        plan[4u * j + 1u] = ((s1 >> 2) - span_lo) << 5 | (8u - (s1 & 3u) * 2u);
        plan[4u * j + 2u] = ((s2 >> 2) - span_lo) << 5 | (10u - (s2 & 3u) * 2u);
        plan[4u * j + 3u] = ((s3 >> 2) - span_lo) << 5 | (12u - (s3 & 3u) * 2u);
    }
}

// One byte of a virtual record from the staged record span and the byte's plan entry.
PGB_DEV uint32_t k2b_compact_byte(const uint8_t *row, const pgb_u4 &e) {
    const uint32_t acc = (pgb_funnel_l(0u, row[e.x >> 5], e.x) & 0x00C0u) | (pgb_funnel_l(0u, row[e.y >> 5], e.y) & 0x0300u) |
                         (pgb_funnel_l(0u, row[e.z >> 5], e.z) & 0x0C00u) | (pgb_funnel_l(0u, row[e.w >> 5], e.w) & 0x3000u);
    return acc >> 6;
}

// One byte of a line's GT text from its (virtual) record: g = offset from the start of the text.
PGB_DEV uint32_t k2b_gt_byte(const uint8_t *vrec, const uint8_t *lut, uint32_t g) {
    const uint32_t f = g >> 2;
    const uint32_t code = ((uint32_t)vrec[f >> 2] >> ((f & 3u) * 2u)) & 3u;
    return lut[code * 8u + (g & 3u)]; // table entry `code`: the text word of genotype `code` comes first
}

// Where a consumer warp puts its text: its part of the shared-memory image (IMG; the image leaves by bulk store),
// or global memory directly (16-byte .cs stores for the aligned chunks, byte stores for prefixes and ragged ends).
// Offsets are image offsets in both cases; `g_al` is the global address of image offset 0 (16-byte aligned).
template <bool IMG>
PGB_DEV void k2b_put8(uint8_t *outb, uint64_t g_al, uint32_t off, uint32_t v) {
    if (IMG) outb[off] = (uint8_t)v;
    else pgb_st8(g_al + off, v);
}
template <bool IMG>
PGB_DEV void k2b_put16(uint8_t *outb, uint64_t g_al, uint32_t off, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    if (IMG) k2b_sts16(outb + off, x, y, z, w);
    else pgb_st16(g_al + off, x, y, z, w, 1);
}

// ---- CONSUMER, one line, first half (a warp per line, no CTA-wide synchronisation): prefix bytes
//      (pfile.rs:157-161) and the newline (pfile.rs:190); GATHER (pfile.rs:171-175): lane <-> bytes lane,
//      lane + 32, ... of the warp's virtual record, 4 LDS.U8 + shift/mask per byte.  `preg0` / `preg1`: the lane's
//      plan entries for bytes lane and lane + 32, held in registers for the whole kernel when ceil(K/4) <= 64. ----
template <bool GATHER, bool IMG>
PGB_DEV void k2b_line_gather(const pgb_k2b_params &p, uint8_t *smem, const pgb_k2b_layout &L, uint32_t stage, uint32_t img,
                             uint64_t g_al, uint32_t l, uint32_t warp, uint32_t lane, const pgb_u4 &preg0,
                             const pgb_u4 &preg1) {
    const pgb_u4 d = pgb_lds4(reinterpret_cast<const pgb_u4 *>(smem + L.tab(stage) + K2B_TAB_LINES + 32u * l));
    uint8_t *outb = smem + L.outb(img);
    const uint8_t *src = smem + d.z;
    for (uint32_t x = lane; x < d.y; x += 64) { // two bytes per lane and iteration: both loads before the stores
        const bool two = x + 32u < d.y;
        const uint32_t b0 = src[x], b1 = two ? src[x + 32u] : 0u;
        k2b_put8<IMG>(outb, g_al, d.x + x, b0);
        if (two) k2b_put8<IMG>(outb, g_al, d.x + x + 32u, b1);
    }
    if (lane < p.sfx_len) k2b_put8<IMG>(outb, g_al, d.x + d.y + lane, (p.sfx >> (8u * lane)) & 0xFFu);
    if (lane == 31) k2b_put8<IMG>(outb, g_al, d.w, '\n');
    if (!GATHER) return;
    const uint8_t *row = smem + reinterpret_cast<const uint32_t *>(smem + L.tab(stage) + K2B_TAB_LINES + 32u * l)[4];
    const uint32_t nb = (p.K + 3u) >> 2;
    uint8_t *vrec = smem + L.vrec + warp * p.vcap;
    const pgb_u4 *plan = reinterpret_cast<const pgb_u4 *>(smem + L.plan);
    if (nb <= 64u) { // the whole plan is in registers: bytes lane and lane + 32
        if (lane < nb) vrec[lane] = (uint8_t)k2b_compact_byte(row, preg0);
        if (lane + 32u < nb) vrec[lane + 32u] = (uint8_t)k2b_compact_byte(row, preg1);
        return;
    }
    for (uint32_t j = lane; j < nb; j += 32) vrec[j] = (uint8_t)k2b_compact_byte(row, pgb_lds4(plan + j));
}

// ---- CONSUMER, one line, second half: FORMAT (pfile.rs:177-188) — the GT text ----
template <bool GATHER, bool IMG>
PGB_DEV void k2b_line_format(const pgb_k2b_params &p, uint8_t *smem, const pgb_k2b_layout &L, uint32_t stage, uint32_t img,
                             uint64_t g_al, uint32_t l, uint32_t warp, uint32_t lane) {
    const pgb_u4 d = pgb_lds4(reinterpret_cast<const pgb_u4 *>(smem + L.tab(stage) + K2B_TAB_LINES + 16u + 32u * l));
    const uint8_t *lut = smem + L.lut;
    uint8_t *outb = smem + L.outb(img);
    const uint8_t *vrec = GATHER ? smem + L.vrec + warp * p.vcap : smem + d.x;
    const uint32_t o_gs = d.y, b0 = d.z, b1 = d.w;
    if (b0 >= b1) { // no aligned chunk inside the text: byte by byte
        const uint32_t K4 = 4u * p.K;
        for (uint32_t g = lane; g < K4; g += 32) k2b_put8<IMG>(outb, g_al, o_gs + g, k2b_gt_byte(vrec, lut, g));
        return;
    }
    const uint32_t delta = b0 - o_gs; // (A - o_gs) & 15 for any 16-aligned A
    const uint32_t r8 = (delta & 3u) * 8u, sh = (delta >> 2) * 2u;
    const uint8_t *vp = vrec + lane;
    // two chunks per lane and iteration (512 bytes apart): their loads are issued together, so a warp has two
    // independent LDS -> ALU -> LDS -> ALU -> store chains in flight instead of one
    for (uint32_t A = b0 + 16u * lane; A < b1; A += 1024u, vp += 64) {
        const bool two = A + 512u < b1;
        const uint32_t a0 = vp[0], a1 = vp[1], c0 = two ? vp[32] : 0u, c1 = two ? vp[33] : 0u;
        const uint32_t w = pgb_prmt(a0, a1, 0x1140u) >> sh, v = pgb_prmt(c0, c1, 0x1140u) >> sh; // 10 code bits: 5 fields
        const pgb_u2 e01 = k2b_lds8(lut + ((w & 15u) << 3)), e23 = k2b_lds8(lut + ((w & 0xF0u) >> 1));
        const pgb_u2 f01 = k2b_lds8(lut + ((v & 15u) << 3)), f23 = k2b_lds8(lut + ((v & 0xF0u) >> 1));
        const uint32_t W4 = pgb_prmt(0x2E313030u, 0x00002F09u, ((w >> 4) & 0x30u) | 0x0504u);
        const uint32_t V4 = pgb_prmt(0x2E313030u, 0x00002F09u, ((v >> 4) & 0x30u) | 0x0504u);
        k2b_put16<IMG>(outb, g_al, A, pgb_funnel_r(e01.x, e01.y, r8), pgb_funnel_r(e01.y, e23.x, r8),
                       pgb_funnel_r(e23.x, e23.y, r8), pgb_funnel_r(e23.y, W4, r8));
        if (two)
            k2b_put16<IMG>(outb, g_al, A + 512u, pgb_funnel_r(f01.x, f01.y, r8), pgb_funnel_r(f01.y, f23.x, r8),
                           pgb_funnel_r(f23.x, f23.y, r8), pgb_funnel_r(f23.y, V4, r8));
    }
    // <= 15 bytes of text in front of the first chunk (lanes 0-15) and behind the last one (lanes 16-31)
    const uint32_t x = lane < 16 ? o_gs + lane : b1 + (lane - 16u);
    const uint32_t end = lane < 16 ? b0 : o_gs + 4u * p.K;
    if (x < end) k2b_put8<IMG>(outb, g_al, x, k2b_gt_byte(vrec, lut, x - o_gs));
}

// The 16-entry text table: nibble -> the text words of its two genotypes.
PGB_DEV void k2b_build_lut(uint8_t *smem, const pgb_k2b_layout &L, uint32_t tid) {
    if (tid < 16u) {
        uint32_t *e = reinterpret_cast<uint32_t *>(smem + L.lut) + 2u * tid;
        e[0] = pgb_gt_word(tid & 3u);
        e[1] = pgb_gt_word(tid >> 2);
    }
}

// Ragged ends of a consumer warp's byte range (everything else leaves by bulk store): image bytes [ws, we) <->
// global [g_al + ws, g_al + we), g_al + (a multiple of 16) 16-byte aligned.  One warp.
PGB_DEV void k2b_store_edges(uint64_t g_al, const uint8_t *outb, uint32_t ws, uint32_t we, uint32_t lane) {
    const uint32_t h0 = (ws + 15u) & ~15u, h1 = we & ~15u;
    if (h0 >= h1) { // no aligned chunk: fewer than 31 bytes
        const uint32_t x = ws + lane;
        if (x < we) pgb_st8(g_al + x, outb[x]);
        return;
    }
    const uint32_t x = lane < 16 ? ws + lane : h1 + (lane - 16u);
    const uint32_t end = lane < 16 ? h0 : we;
    if (x < end) pgb_st8(g_al + x, outb[x]);
}
