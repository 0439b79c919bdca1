// k2_batch.cuh — K2 for launches whose lines are short enough that a CTA can stage a batch of
// B consecutive lines in shared memory (the gather-heavy chr22 shape: 250 kept samples of
// 2 504 -> 1 041-byte lines; short keep-all lines).  Like k2_core.cuh the source compiles
// both as sm_100a device code and, with PGB_HOSTSIM, as a thread-by-thread host simulation
// used ONLY by tests/.
//
// What it replaces: the same loop body as k2_core.cuh (/root/reference/src/pfile.rs:156-192),
// organised the way BASELINE.json's north_star (b)+(c) describes it.  The kernel is persistent
// (CTA c takes batches c, c + gridDim, ...) and software-pipelined over two shared-memory stages:
//   1. INPUT, one batch ahead: thread l hands the 16-byte-aligned covering ranges of line l's
//      record (records + rec_off: pfile.rs:165-170 as a TMA transfer) and of its prefix bytes to
//      the bulk-copy engine (cp.async.bulk.shared.global, completion on the stage's mbarrier);
//      the pgb_line_meta of the batch after that is already in registers, so neither the index
//      read nor the record read is ever waited for in the steady state;
//   2. GATHER (pfile.rs:171-175): the kept samples of each staged record are compacted into a
//      packed 2-bit "virtual record" of ceil(K/4) bytes.  One thread per OUTPUT byte; the
//      thread's four source positions (byte offset, bit shift) come from the K0 index list and
//      stay in registers for the whole kernel — 4 LDS.U8 + shift/mask per byte, no warp
//      reductions;
//   3. FORMAT (pfile.rs:177-190): the keep-all chunk body runs on the virtual record and writes
//      the text with 16-byte shared-memory stores into an image of the batch's output bytes
//      (lines are back to back in the VCF, so a batch is ONE contiguous byte range); prefixes,
//      the <= 15 + 15 GT bytes around each line's aligned body and the newlines are byte stores
//      — into shared memory, not global;
//   4. OUTPUT: the 16-byte-aligned interior of the image leaves with ONE bulk async store
//      (cp.async.bulk.global.shared::cta) that drains while the next batch is formatted into the
//      other image; the <= 15 bytes on either end are byte stores.
// Every output byte still has exactly one writer and no padding is emitted.
//
// Text decode uses a 16-entry table (a nibble = two genotypes -> 8 bytes of text, 128 bytes
// in all): lanes reading the same entry broadcast and different entries live in different
// banks, so the two LDS.64 of a chunk are bank-conflict-free by construction.
#pragma once
#include "k2_core.cuh"

#define K2B_THREADS 256
#define K2B_WARPS 8
#define K2B_MAX_LINES 128

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
    uint32_t B;        // lines per batch, <= K2B_MAX_LINES
    uint32_t rowcap;   // shared-memory bytes per staged record: align16(R + 31)
    uint32_t pcap;     // shared-memory bytes per staged prefix: align16(max prefix + 31)
    uint32_t vcap;     // shared-memory bytes per virtual record (gather): align16(ceil(K/4) + 2)
    uint32_t outcap;   // shared-memory bytes of one output image: align128(B * max_line + 32)
    uint32_t sfx;      // bytes appended to every prefix after its blob bytes (little-endian), e.g. "\tGT"
    uint32_t sfx_len;  // 0..4; pfx_len includes it
    uint32_t kidx_vec; // kidx is 16-byte aligned
    uint32_t store_mode; // 0 bulk async store, 1 16-byte st.global by all threads (A/B comparisons)
};

// Shared memory: [0,8) [8,16) the two stage mbarriers, the text table, then per stage s: a table
// (body offset of the batch; per line: image offset of the line start and of the GT text, offsets of
// the staged record span and prefix), the staged records and prefixes; the virtual records; two images.
struct pgb_k2b_layout {
    uint32_t lut, tab[2], rows[2], pst[2], vrec, outb[2], total;
    uint32_t t_ols, t_ogs, t_rbase, t_pbase; // offsets inside a stage table
};

#if defined(PGB_HOSTSIM)
#define PGB_HD static inline
#else
#define PGB_HD __host__ __device__ __forceinline__
#endif

PGB_HD uint32_t pgb_k2b_align(uint32_t x, uint32_t a) { return (x + a - 1u) & ~(a - 1u); }

PGB_HD pgb_k2b_layout pgb_k2b_smem_layout(uint32_t B, uint32_t rowcap, uint32_t pcap, uint32_t vcap, uint32_t outcap,
                                          bool gather) {
    pgb_k2b_layout L;
    L.lut = 16;
    L.t_ols = 8;
    L.t_ogs = L.t_ols + 4u * (B + 1u);
    L.t_rbase = L.t_ogs + 4u * B;
    L.t_pbase = L.t_rbase + 4u * B;
    const uint32_t tabsz = pgb_k2b_align(L.t_pbase + 4u * B, 16);
    L.tab[0] = L.lut + 128;
    L.tab[1] = L.tab[0] + tabsz;
    L.rows[0] = L.tab[1] + tabsz;
    L.rows[1] = L.rows[0] + B * rowcap;
    L.pst[0] = L.rows[1] + B * rowcap;
    L.pst[1] = L.pst[0] + B * pcap;
    L.vrec = L.pst[1] + B * pcap;
    L.outb[0] = pgb_k2b_align(L.vrec + (gather ? B * vcap : 0u), 128);
    L.outb[1] = L.outb[0] + outcap;
    L.total = L.outb[1] + outcap;
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

// ---- INPUT: stage table + record/prefix fetch of one batch.  Thread t <-> line i0 + t (thread nbl carries
//      the end of the batch); `m` is that line's pgb_line_meta, `base` the batch's offset in the body.
//      Threads 0..B-1 each arrive once on the stage's mbarrier. ----
PGB_DEV void k2b_phase_issue(const pgb_k2b_params &p, uint8_t *smem, const pgb_k2b_layout &L, uint32_t stage, uint32_t nbl,
                             uint32_t tid, const pgb_line_meta &m, uint64_t base, uint32_t span_lo, uint32_t span_len) {
    uint8_t *mbar = smem + 8u * stage;
    if (tid > nbl) {
        if (tid < p.B) k2b_arrive(mbar);
        return;
    }
    uint8_t *tab = smem + L.tab[stage];
    const uint32_t phase = (uint32_t)((uint64_t)(uintptr_t)p.out + base) & 15u;
    const uint32_t o_ls = phase + (uint32_t)(m.line_off - base);
    reinterpret_cast<uint32_t *>(tab + L.t_ols)[tid] = o_ls;
    if (tid == 0) *reinterpret_cast<uint64_t *>(tab) = base;
    if (tid == nbl) {
        if (tid < p.B) k2b_arrive(mbar);
        return;
    }
    reinterpret_cast<uint32_t *>(tab + L.t_ogs)[tid] = o_ls + m.pfx_len;
    const uint8_t *src = p.records + m.rec_off + span_lo;
    const uint32_t ph = (uint32_t)(uintptr_t)src & 15u;
    reinterpret_cast<uint32_t *>(tab + L.t_rbase)[tid] = tid * p.rowcap + ph;
    const uint32_t dlen = m.pfx_len - p.sfx_len;
    const uint8_t *psrc = p.prefix_blob + m.pfx_off;
    const uint32_t pph = dlen ? (uint32_t)(uintptr_t)psrc & 15u : 0u;
    reinterpret_cast<uint32_t *>(tab + L.t_pbase)[tid] = tid * p.pcap + pph;
    k2b_stage_load(mbar, smem + L.rows[stage] + tid * p.rowcap, src - ph, span_len ? (ph + span_len + 15u) & ~15u : 0u,
                   smem + L.pst[stage] + tid * p.pcap, psrc - pph, dlen ? (pph + dlen + 15u) & ~15u : 0u);
}

// ---- prefixes (pfile.rs:157-161) and newlines (pfile.rs:190) into the image; warp <-> line ----
PGB_DEV void k2b_phase_prefix(const pgb_k2b_params &p, uint8_t *smem, const pgb_k2b_layout &L, uint32_t stage, uint32_t nbl,
                              uint32_t warp, uint32_t lane) {
    const uint8_t *tab = smem + L.tab[stage];
    const uint32_t *ols = reinterpret_cast<const uint32_t *>(tab + L.t_ols);
    const uint32_t *ogs = reinterpret_cast<const uint32_t *>(tab + L.t_ogs);
    const uint32_t *pbase = reinterpret_cast<const uint32_t *>(tab + L.t_pbase);
    uint8_t *outb = smem + L.outb[stage];
    for (uint32_t l = warp; l < nbl; l += K2B_WARPS) {
        const uint32_t o_ls = ols[l], dlen = ogs[l] - o_ls - p.sfx_len;
        const uint8_t *src = smem + L.pst[stage] + pbase[l];
        for (uint32_t x = lane; x < dlen; x += 32) outb[o_ls + x] = src[x];
        if (lane < p.sfx_len) outb[o_ls + dlen + lane] = (uint8_t)(p.sfx >> (8u * lane));
        if (lane == 31) outb[ols[l + 1] - 1u] = '\n';
    }
}

// The gather plan of one thread: the four kept samples of virtual-record byte j as (source byte relative to
// the staged span) | (left shift bringing the sample's two bits to [6 + 2k, 8 + 2k)) << 24.  Constant for the
// whole kernel: the kept-sample list is the same for every variant (pfile.rs:128,171).
struct pgb_k2b_plan {
    uint32_t e[4];
};

PGB_DEV pgb_k2b_plan k2b_load_plan(const pgb_k2b_params &p, uint32_t j, uint32_t span_lo) {
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
    pgb_k2b_plan pl;
    pl.e[0] = ((s0 >> 2) - span_lo) | (6u - (s0 & 3u) * 2u) << 24;
    pl.e[1] = ((s1 >> 2) - span_lo) | (8u - (s1 & 3u) * 2u) << 24;
    pl.e[2] = ((s2 >> 2) - span_lo) | (10u - (s2 & 3u) * 2u) << 24;
    pl.e[3] = ((s3 >> 2) - span_lo) | (12u - (s3 & 3u) * 2u) << 24;
    return pl;
}

PGB_DEV uint32_t k2b_compact_byte(const uint8_t *row, const pgb_k2b_plan &pl) {
    const uint32_t acc = (((uint32_t)row[pl.e[0] & 0xFFFFFFu] << (pl.e[0] >> 24)) & 0x00C0u) |
                         (((uint32_t)row[pl.e[1] & 0xFFFFFFu] << (pl.e[1] >> 24)) & 0x0300u) |
                         (((uint32_t)row[pl.e[2] & 0xFFFFFFu] << (pl.e[2] >> 24)) & 0x0C00u) |
                         (((uint32_t)row[pl.e[3] & 0xFFFFFFu] << (pl.e[3] >> 24)) & 0x3000u);
    return acc >> 6;
}

// Thread <-> byte mapping of the compaction: W threads per line (a power of two >= ceil(K/4), at most the CTA),
// K2B_THREADS / W lines at a time.
PGB_DEV uint32_t k2b_compact_width(uint32_t nb) {
    uint32_t W = 32;
    while (W < nb && W < K2B_THREADS) W <<= 1;
    return W;
}

// ---- GATHER: staged records -> packed virtual records, one thread per output byte.  `plan0` is the plan of
//      byte (tid & (W-1)) when ceil(K/4) <= K2B_THREADS (the common case: loaded once per kernel). ----
PGB_DEV void k2b_phase_compact(const pgb_k2b_params &p, uint8_t *smem, const pgb_k2b_layout &L, uint32_t stage, uint32_t nbl,
                               uint32_t tid, uint32_t span_lo, const pgb_k2b_plan &plan0) {
    const uint32_t nb = (p.K + 3u) >> 2;
    const uint32_t W = k2b_compact_width(nb);
    const uint32_t LP = K2B_THREADS / W, jl = tid & (W - 1u), l0 = tid / W;
    const uint32_t *rbase = reinterpret_cast<const uint32_t *>(smem + L.tab[stage] + L.t_rbase);
    const uint8_t *rows = smem + L.rows[stage];
    uint8_t *vrec = smem + L.vrec;
    if (nb <= K2B_THREADS) {
        if (jl < nb)
            for (uint32_t l = l0; l < nbl; l += LP) vrec[l * p.vcap + jl] = (uint8_t)k2b_compact_byte(rows + rbase[l], plan0);
        return;
    }
    for (uint32_t j = jl; j < nb; j += W) { // long kept lists: the plan is re-read per byte column
        const pgb_k2b_plan pl = k2b_load_plan(p, j, span_lo);
        for (uint32_t l = l0; l < nbl; l += LP) vrec[l * p.vcap + j] = (uint8_t)k2b_compact_byte(rows + rbase[l], pl);
    }
}

// One byte of a line's GT text from its (virtual) record: g = offset from the start of the text.
PGB_DEV uint32_t k2b_gt_byte(const uint8_t *vrec, const uint8_t *lut, uint32_t g) {
    const uint32_t f = g >> 2;
    const uint32_t code = ((uint32_t)vrec[f >> 2] >> ((f & 3u) * 2u)) & 3u;
    return lut[code * 8u + (g & 3u)]; // table entry `code`: the text word of genotype `code` comes first
}

// ---- FORMAT: GT text of every line into the image; warp <-> line ----
template <bool GATHER>
PGB_DEV void k2b_phase_format(const pgb_k2b_params &p, uint8_t *smem, const pgb_k2b_layout &L, uint32_t stage, uint32_t nbl,
                              uint32_t warp, uint32_t lane) {
    const uint8_t *tab = smem + L.tab[stage];
    const uint32_t *ogs = reinterpret_cast<const uint32_t *>(tab + L.t_ogs);
    const uint32_t *rbase = reinterpret_cast<const uint32_t *>(tab + L.t_rbase);
    const uint8_t *lut = smem + L.lut;
    uint8_t *outb = smem + L.outb[stage];
    const uint32_t K4 = 4u * p.K;
    for (uint32_t l = warp; l < nbl; l += K2B_WARPS) {
        const uint8_t *vrec = GATHER ? smem + L.vrec + l * p.vcap : smem + L.rows[stage] + rbase[l];
        const uint32_t o_gs = ogs[l], o_ge = o_gs + K4;
        const uint32_t b0 = (o_gs + 15u) & ~15u, b1 = o_ge & ~15u;
        if (b0 >= b1) { // no aligned chunk inside the text: byte by byte
            for (uint32_t g = lane; g < K4; g += 32) outb[o_gs + g] = (uint8_t)k2b_gt_byte(vrec, lut, g);
            continue;
        }
        const uint32_t delta = b0 - o_gs; // (A - o_gs) & 15 for any 16-aligned A
        const uint32_t r8 = (delta & 3u) * 8u, sh = (delta >> 2) * 2u;
        const uint8_t *vp = vrec + lane;
        for (uint32_t A = b0 + 16u * lane; A < b1; A += 512u, vp += 32) {
            const uint32_t w = pgb_prmt(vp[0], vp[1], 0x1140u) >> sh; // 10 code bits: 5 fields
            const pgb_u2 e01 = k2b_lds8(lut + ((w & 15u) << 3)), e23 = k2b_lds8(lut + ((w & 0xF0u) >> 1));
            const uint32_t W4 = pgb_prmt(0x2E313030u, 0x00002F09u, ((w >> 4) & 0x30u) | 0x0504u);
            k2b_sts16(outb + A, pgb_funnel_r(e01.x, e01.y, r8), pgb_funnel_r(e01.y, e23.x, r8),
                      pgb_funnel_r(e23.x, e23.y, r8), pgb_funnel_r(e23.y, W4, r8));
        }
        // <= 15 bytes of text in front of the first chunk (lanes 0-15) and behind the last one (lanes 16-31)
        const uint32_t x = lane < 16 ? o_gs + lane : b1 + (lane - 16u);
        const uint32_t end = lane < 16 ? b0 : o_ge;
        if (x < end) outb[x] = (uint8_t)k2b_gt_byte(vrec, lut, x - o_gs);
    }
}

// The 16-entry text table: nibble -> the text words of its two genotypes.
PGB_DEV void k2b_build_lut(uint8_t *smem, const pgb_k2b_layout &L, uint32_t tid) {
    if (tid < 16u) {
        uint32_t *e = reinterpret_cast<uint32_t *>(smem + L.lut) + 2u * tid;
        e[0] = pgb_gt_word(tid & 3u);
        e[1] = pgb_gt_word(tid >> 2);
    }
}

// Ragged ends of the batch's byte range in global memory (everything else leaves by bulk store):
// image bytes [phase, phase + T) <-> global [g_al + phase, ...), g_al 16-byte aligned.  One warp.
PGB_DEV void k2b_store_edges(uint64_t g_al, const uint8_t *outb, uint32_t phase, uint32_t T, uint32_t lane) {
    const uint32_t h0 = phase ? 16u : 0u, h1 = (phase + T) & ~15u;
    if (h0 >= h1) { // no aligned chunk: T < 31
        const uint32_t x = phase + lane;
        if (x < phase + T) pgb_st8(g_al + x, outb[x]);
        return;
    }
    const uint32_t x = lane < 16 ? phase + lane : h1 + (lane - 16u);
    const uint32_t end = lane < 16 ? h0 : phase + T;
    if (x < end) pgb_st8(g_al + x, outb[x]);
}
