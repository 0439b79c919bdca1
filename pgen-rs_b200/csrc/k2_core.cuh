// k2_core.cuh — per-lane body of K2 (decode + gather + format), written so that the
// identical source compiles (a) as the sm_100a device code of libpgb200.so and (b) with
// PGB_HOSTSIM as a lane-by-lane host simulation used ONLY by tests/ to validate the
// addressing/byte logic before spending GPU time.  The simulation is never linked into
// the product library.
//
// What it replaces: the per-variant loop body of Pfile::output_vcf,
// /root/reference/src/pfile.rs:156-192 —
//   :157-161 prefix fields + "GT"       -> prefix bytes copied from prefix_blob
//   :165-170 record offset/seek/read    -> meta.rec_off (device record index from K1)
//   :171-175 per-sample 2-bit extract   -> body chunks (keep-all: two row bytes per 16 B
//                                          of text; gather: kidx lookups)
//   :177-188 code -> "\t0/0" ...        -> shared-memory LUT: record byte -> 4 text words
//   :190     "\n"                       -> tail bytes
//
// Geometry of one output line at absolute address a_ls (P = prefix bytes, K kept samples):
//   [a_ls, a_gs)  prefix      a_gs = a_ls + P
//   [a_gs, a_ge)  GT text     a_ge = a_gs + 4K      (field f at a_gs + 4f: '\t' x '/' y)
//   [a_ge, a_le)  '\n'        a_le = a_ge + 1
// The warp writes   [a_ls, a_gs)  prefix, byte stores
//                   [a_gs, b0)    <= 15 bytes of GT text, byte stores (b0 = align_up(a_gs, 16))
//                   body = [b0, b1)  16-byte stores, 512 B per warp instruction, every chunk
//                                    wholly inside the GT text
//                   [b1, a_le)    <= 15 bytes of GT text + '\n' (b1 = align_down(a_ge, 16))
// so every output byte has exactly one writer and no padding is ever emitted (the VCF
// must be bit-exact).  For a 16-byte chunk at A:  q = A - a_gs, field f0 = q >> 2, byte
// phase r = q & 3; (q & 15) is the same for every chunk of a line, so in the keep-all
// case the chunk needs row bytes (q >> 4) and (q >> 4) + 1 shifted by a per-line constant.
//
// Instruction budget (round-1 ncu: the first version was ALU-bound at ~100 instructions per
// 16-byte chunk): the chunk body is 2 LDG.U8, PRMT+SHF (the 10 code bits w), LOP+IMAD+LDS.128
// (four text words from the LUT), SHF+LOP+PRMT (the 5th field's word), 4 SHF (byte re-phasing),
// STG.128.  Entry points: pgb_k2_line (a whole line per warp, 32-bit line-relative offsets;
// launches whose lines fit in one tile) and pgb_k2_item (a 16 KiB tile of a wide line); both
// end in pgb_k2_body.
#pragma once
#include <stdint.h>

#include "pgb200.h"

struct pgb_k2_params {
    const uint8_t *records;
    const pgb_line_meta *meta;
    const uint8_t *prefix_blob;
    const uint32_t *kidx; // nullptr => keep all
    uint8_t *out;
    uint64_t n_lines;
    uint32_t K;
    uint32_t n_tiles;    // tiles per line
    uint32_t tile_bytes; // multiple of 512
    uint32_t row_bytes_hint; // keep-all: bytes of a record worth prefetching (R + 1)
    uint32_t kidx_vec;       // kidx is 16-byte aligned: index reads may be vectorised
    uint32_t sfx;            // bytes appended to every prefix after its prefix_blob bytes (little-endian), e.g. "\tGT"
    uint32_t sfx_len;        // 0..4; pgb_line_meta::pfx_len includes it
};

struct pgb_u4 {
    uint32_t x, y, z, w;
};

#if defined(PGB_HOSTSIM)
#include <string.h>
#define PGB_DEV static inline
PGB_DEV uint32_t pgb_prmt(uint32_t x, uint32_t y, uint32_t s) {
    uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int k = 0; k < 4; k++) {
        uint32_t sel = (s >> (4 * k)) & 7u;
        r |= (uint32_t)((v >> (8 * sel)) & 0xFF) << (8 * k);
    }
    return r;
}
PGB_DEV uint32_t pgb_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    sh &= 31u;
    return sh ? ((lo >> sh) | (hi << (32u - sh))) : lo;
}
PGB_DEV uint32_t pgb_ld8(const uint8_t *p) { return *p; }
PGB_DEV uint32_t pgb_ld32(const uint32_t *p) { return *p; }
PGB_DEV pgb_u4 pgb_ld128(const uint32_t *p) { pgb_u4 r = {p[0], p[1], p[2], p[3]}; return r; }
PGB_DEV pgb_line_meta pgb_ld_meta(const pgb_line_meta *p) { return *p; }
PGB_DEV void pgb_st8(uint64_t a, uint32_t v) { *(uint8_t *)(uintptr_t)a = (uint8_t)v; }
PGB_DEV void pgb_st16(uint64_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w, int) {
    uint32_t t[4] = {x, y, z, w};
    memcpy((void *)(uintptr_t)a, t, 16);
}
PGB_DEV pgb_u4 pgb_lds4(const pgb_u4 *e) { return *e; }
#else
#define PGB_DEV __device__ __forceinline__
PGB_DEV uint32_t pgb_prmt(uint32_t x, uint32_t y, uint32_t s) { return __byte_perm(x, y, s); }
PGB_DEV uint32_t pgb_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
PGB_DEV uint32_t pgb_ld8(const uint8_t *p) { return __ldg(p); }
PGB_DEV uint32_t pgb_ld32(const uint32_t *p) { return __ldg(p); }
PGB_DEV pgb_u4 pgb_ld128(const uint32_t *p) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    pgb_u4 r = {v.x, v.y, v.z, v.w};
    return r;
}
PGB_DEV pgb_line_meta pgb_ld_meta(const pgb_line_meta *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    pgb_line_meta m;
    m.line_off = ((uint64_t)a.y << 32) | a.x;
    m.rec_off = ((uint64_t)a.w << 32) | a.z;
    m.pfx_off = ((uint64_t)b.y << 32) | b.x;
    m.pfx_len = b.z;
    m.reserved = b.w;
    return m;
}
PGB_DEV void pgb_st8(uint64_t a, uint32_t v) { *reinterpret_cast<uint8_t *>(a) = (uint8_t)v; }
PGB_DEV void pgb_st16(uint64_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w, int hint) {
    if (hint == 1) {
        asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
    } else if (hint == 2) {
        asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(a), "r"(x), "r"(y), "r"(z), "r"(w)
                     : "memory");
    } else {
        asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
    }
}
PGB_DEV pgb_u4 pgb_lds4(const pgb_u4 *e) {
    uint4 v = *reinterpret_cast<const uint4 *>(e);
    pgb_u4 r = {v.x, v.y, v.z, v.w};
    return r;
}
#endif

// Little-endian text word of one genotype field: '\t', a, '/', b  (pfile.rs:177-188).
//   code 0 -> "\t0/0" 0x302F3009   1 -> "\t0/1" 0x312F3009
//   code 2 -> "\t1/1" 0x312F3109   3 -> "\t./." 0x2E2F2E09
// Byte pool {09,'0','1','.'} + {'/'}; the first PRMT turns the code into the selector
// (0x1410, 0x2410, 0x2420, 0x3430), the second applies it.
PGB_DEV uint32_t pgb_gt_word(uint32_t c) {
    uint32_t sel = pgb_prmt(0x30201010u, 0x34242414u, c * 0x11u + 0x40u);
    return pgb_prmt(0x2E313009u, 0x0000002Fu, sel);
}

// The four text words of a packed record byte (samples 4j..4j+3) — one smem LUT entry.
PGB_DEV pgb_u4 pgb_lut_entry(uint32_t byte) {
    pgb_u4 e;
    e.x = pgb_gt_word(byte & 3u);
    e.y = pgb_gt_word((byte >> 2) & 3u);
    e.z = pgb_gt_word((byte >> 4) & 3u);
    e.w = pgb_gt_word((byte >> 6) & 3u);
    return e;
}

// 2-bit code of sample s in a record (pfile.rs:172-175).
PGB_DEV uint32_t pgb_code(const uint8_t *row, uint32_t s) { return (pgb_ld8(row + (s >> 2)) >> ((s & 3u) * 2u)) & 3u; }

// One byte of the GT region: g = offset from a_gs (g == 4K is the newline).  Used only for
// the <= 15 + 16 bytes around the 16-byte-aligned body (or for whole lines too short to have one).
template <bool GATHER>
PGB_DEV uint32_t pgb_gt_byte(const pgb_k2_params &p, const uint8_t *row, uint64_t K4, uint64_t g) {
    if (g >= K4) return '\n';
    const uint32_t f = (uint32_t)(g >> 2);
    const uint32_t s = GATHER ? pgb_ld32(p.kidx + f) : f;
    return (pgb_gt_word(pgb_code(row, s)) >> (((uint32_t)g & 3u) * 8u)) & 0xFFu;
}

// Shared-memory table (built once per CTA by the kernel, per call by the host simulation):
//   lut4[byte * REPL + g]   the four text words of a packed record byte; REPL interleaved copies
//                           so that the 8 lanes of a quarter-warp (the unit of an LDS.128) can each
//                           use copy g = lane & (REPL-1) and never collide on a bank (REPL = 8).
// `l4` is the lane's view of it (lut4 + g).  w holds the chunk's 10 code bits (5 fields): bits
// 0-7 index the table; of the 5th field only '\t', its first character and '/' can fall
// inside the chunk (byte phase r <= 3), so its word comes from one PRMT on the table
// {'0','0','1','.'} | {'\t','/'} instead of a second lookup.
template <int REPL>
PGB_DEV void pgb_emit_chunk(uint64_t A, uint32_t w, uint32_t r8, const pgb_u4 *l4, int hint) {
    const pgb_u4 e = pgb_lds4(l4 + (w & 0xFFu) * REPL);
    const uint32_t W4 = pgb_prmt(0x2E313030u, 0x00002F09u, ((w >> 4) & 0x30u) | 0x0504u);
    pgb_st16(A, pgb_funnel_r(e.x, e.y, r8), pgb_funnel_r(e.y, e.z, r8), pgb_funnel_r(e.z, e.w, r8),
             pgb_funnel_r(e.w, W4, r8), hint);
}

// The 10 code bits (5 fields) of a 16-byte chunk.  Keep-all: `src` points at the record byte
// holding the chunk's first field (row + (q >> 4)), sh = 2-bit phase.  Gather: `src` points at
// the chunk's first kept-sample index (kidx + (q >> 2); kidx is padded by 8 entries).
template <bool GATHER, bool VEC = false>
PGB_DEV uint32_t pgb_chunk_codes(const uint8_t *row, const void *src, uint32_t sh) {
    if (!GATHER) {
        const uint8_t *b = (const uint8_t *)src;
        return pgb_prmt(pgb_ld8(b), pgb_ld8(b + 1), 0x1140u) >> sh;
    }
    const uint32_t *ki = (const uint32_t *)src;
    uint32_t s0, s1, s2, s3, s4;
    if (VEC) {
        // the chunk's first field f0 is congruent to j = sh / 2 modulo 4 for every chunk of the line, so
        // the five indices sit at offset j of two aligned 16-byte reads (kidx is 16-byte aligned)
        const uint32_t j = sh >> 1; // warp-uniform
        const pgb_u4 a = pgb_ld128(ki - j), b = pgb_ld128(ki - j + 4);
        switch (j) {
        case 0: s0 = a.x; s1 = a.y; s2 = a.z; s3 = a.w; s4 = b.x; break;
        case 1: s0 = a.y; s1 = a.z; s2 = a.w; s3 = b.x; s4 = b.y; break;
        case 2: s0 = a.z; s1 = a.w; s2 = b.x; s3 = b.y; s4 = b.z; break;
        default: s0 = a.w; s1 = b.x; s2 = b.y; s3 = b.z; s4 = b.w; break;
        }
    } else {
        s0 = pgb_ld32(ki); s1 = pgb_ld32(ki + 1); s2 = pgb_ld32(ki + 2); s3 = pgb_ld32(ki + 3);
        s4 = pgb_ld32(ki + 4);
    }
    return pgb_code(row, s0) | (pgb_code(row, s1) << 2) | (pgb_code(row, s2) << 4) | (pgb_code(row, s3) << 6) |
           (pgb_code(row, s4) << 8);
}

// The aligned body [lo, hi) of a line's GT text (16-byte chunks, every one wholly inside the
// text): short bodies chunk-indexed, long ones in 512-byte-aligned software-pipelined rows.
template <bool GATHER, int HINT, int REPL>
PGB_DEV void pgb_k2_body(const pgb_k2_params &p, const uint8_t *row, uint64_t a_gs, uint64_t lo, uint64_t hi,
                         uint32_t lane, const pgb_u4 *lut4) {
    constexpr int UNROLL = 4;

    const uint32_t delta = (uint32_t)(0ull - a_gs) & 15u; // (A - a_gs) & 15 for any 16-aligned A
    const uint32_t r8 = (delta & 3u) * 8u;                // byte phase inside a field, in bits
    const uint32_t sh = (delta >> 2) * 2u;                // 2-bit phase inside a record byte
    const pgb_u4 *l4 = lut4 + (lane & (uint32_t)(REPL - 1));

    if (hi - lo <= 2048u) {
        // Short body (short lines, e.g. a 10 % sample subset of 2 504): chunk-indexed instead of
        // 512-byte-row-aligned, so that 62 chunks take two warp iterations, not three.
        const uint32_t n_chunks = (uint32_t)(hi - lo) >> 4;
        const uint32_t qlo = (uint32_t)(lo - a_gs);
        for (uint32_t c = lane; c < n_chunks; c += 64) {
            const uint32_t qa = qlo + 16u * c, qb = qa + 512u;
            const bool two = c + 32 < n_chunks;
            const void *sa = GATHER ? (const void *)(p.kidx + (qa >> 2)) : (const void *)(row + (qa >> 4));
            const void *sb = GATHER ? (const void *)(p.kidx + (qb >> 2)) : (const void *)(row + (qb >> 4));
            uint32_t wa, wb = 0u;
            if (GATHER && p.kidx_vec) {
                wa = pgb_chunk_codes<GATHER, true>(row, sa, sh);
                if (two) wb = pgb_chunk_codes<GATHER, true>(row, sb, sh);
            } else {
                wa = pgb_chunk_codes<GATHER>(row, sa, sh);
                if (two) wb = pgb_chunk_codes<GATHER>(row, sb, sh);
            }
            pgb_emit_chunk<REPL>(lo + 16ull * c, wa, r8, l4, HINT);
            if (two) pgb_emit_chunk<REPL>(lo + 16ull * c + 512u, wb, r8, l4, HINT);
        }
        return;
    }
    uint64_t rowA = lo & ~511ull;             // warp-uniform: start of the current 512-byte output row
    uint64_t A = rowA + (uint64_t)lane * 16u; // this lane's chunk in it
    // source of the chunk's codes; one output row further = 32 record bytes / 128 kept samples
    constexpr uint32_t STEP = GATHER ? 128u * 4u : 32u;
    const int64_t q0 = (int64_t)(A - a_gs);   // may be negative only for a chunk that is skipped
    const uint8_t *src = GATHER ? (const uint8_t *)p.kidx + (q0 >> 2) * 4 : row + (q0 >> 4);
    if (!GATHER) {
        // Software-pipelined groups of UNROLL rows: the record bytes of group g+1 are requested
        // before group g is formatted, so a line's DRAM/L2 latency is paid once, not once per
        // group (round-1 ncu: 69 % of the stall samples were the first use of a loaded byte).
        // mask bit u = this lane has a chunk in row u of the group; 0xF on the fast path.
        const uint32_t lo_rel = (uint32_t)(lo - rowA), hi_rel = (uint32_t)(hi - rowA); // tile <= 128 MiB
        uint32_t rel = lane * 16u;
        uint32_t cb[2 * UNROLL], nb[2 * UNROLL];
        uint32_t cm = 0, nm;
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const uint32_t ru = rel + 512u * u;
            cb[2 * u] = cb[2 * u + 1] = 0;
            if (ru >= lo_rel && ru < hi_rel) {
                cm |= 1u << u;
                cb[2 * u] = pgb_ld8(src + STEP * u);
                cb[2 * u + 1] = pgb_ld8(src + STEP * u + 1);
            }
        }
        for (;;) {
            const uint32_t rel_n = rel + 512u * UNROLL;
            const bool more = (rel_n & ~511u) < hi_rel; // warp-uniform
            nm = 0;
            if (more) {
                if ((rel_n | 511u) + 512u * (UNROLL - 1) < hi_rel) { // every lane has all rows
                    nm = (1u << UNROLL) - 1u;
#pragma unroll
                    for (int u = 0; u < UNROLL; u++) {
                        nb[2 * u] = pgb_ld8(src + STEP * (UNROLL + u));
                        nb[2 * u + 1] = pgb_ld8(src + STEP * (UNROLL + u) + 1);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < UNROLL; u++) {
                        nb[2 * u] = nb[2 * u + 1] = 0;
                        if (rel_n + 512u * u < hi_rel) {
                            nm |= 1u << u;
                            nb[2 * u] = pgb_ld8(src + STEP * (UNROLL + u));
                            nb[2 * u + 1] = pgb_ld8(src + STEP * (UNROLL + u) + 1);
                        }
                    }
                }
            }
            if (cm == (1u << UNROLL) - 1u) {
#pragma unroll
                for (int u = 0; u < UNROLL; u++)
                    pgb_emit_chunk<REPL>(A + 512ull * u, pgb_prmt(cb[2 * u], cb[2 * u + 1], 0x1140u) >> sh, r8, l4, HINT);
            } else {
#pragma unroll
                for (int u = 0; u < UNROLL; u++)
                    if (cm & (1u << u))
                        pgb_emit_chunk<REPL>(A + 512ull * u, pgb_prmt(cb[2 * u], cb[2 * u + 1], 0x1140u) >> sh, r8, l4,
                                             HINT);
            }
            if (!more) break;
#pragma unroll
            for (int u = 0; u < 2 * UNROLL; u++) cb[u] = nb[u];
            cm = nm;
            rel = rel_n;
            A += 512ull * UNROLL;
            src += STEP * UNROLL;
        }
        return;
    }
    if (rowA < lo) { // the first row starts in front of the body
        if (A >= lo && A < hi) pgb_emit_chunk<REPL>(A, pgb_chunk_codes<GATHER>(row, src, sh), r8, l4, HINT);
        rowA += 512; A += 512; src += STEP;
    }
    for (; rowA + 512ull * UNROLL <= hi; rowA += 512ull * UNROLL, A += 512ull * UNROLL, src += STEP * UNROLL) {
        uint32_t w[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) w[u] = pgb_chunk_codes<GATHER>(row, src + STEP * u, sh);
#pragma unroll
        for (int u = 0; u < UNROLL; u++) pgb_emit_chunk<REPL>(A + 512ull * u, w[u], r8, l4, HINT);
    }
    for (; rowA < hi; rowA += 512, A += 512, src += STEP)
        if (A < hi) pgb_emit_chunk<REPL>(A, pgb_chunk_codes<GATHER>(row, src, sh), r8, l4, HINT);
}

// One (line, tile) item of a launch whose lines are cut into tiles (wide lines: 500 000 samples =
// 2 MB of text): every range is clipped to the tile [t0, t1).
template <bool GATHER, int HINT, int REPL>
PGB_DEV void pgb_k2_item(const pgb_k2_params &p, const pgb_line_meta &m, uint32_t tile, uint32_t lane,
                         const pgb_u4 *lut4) {
    const uint32_t P = m.pfx_len;
    const uint64_t K4 = 4ull * p.K;
    const uint64_t a_ls = (uint64_t)(uintptr_t)p.out + m.line_off;
    const uint64_t a_gs = a_ls + P;
    const uint64_t a_ge = a_gs + K4;
    const uint64_t a_le = a_ge + 1;
    const uint64_t t0 = (a_ls & ~511ull) + (uint64_t)tile * p.tile_bytes;
    if (t0 >= a_le) return;
    const uint64_t t1 = t0 + p.tile_bytes;
    uint64_t b0 = (a_gs + 15ull) & ~15ull, b1 = a_ge & ~15ull;
    if (b0 >= b1) { b0 = a_le; b1 = a_le; } // no aligned chunk inside the GT text: bytes only
    const uint8_t *row = p.records + m.rec_off;

    { // prefix bytes (pfile.rs:157-161): P - sfx_len bytes of the blob, then the suffix
        const uint8_t *pfx = p.prefix_blob + m.pfx_off;
        const uint64_t D = P - p.sfx_len;
        const uint64_t lo = a_ls > t0 ? a_ls : t0, hi = a_gs < t1 ? a_gs : t1;
        for (uint64_t a = lo + lane; a < hi; a += 32) {
            const uint64_t x = a - a_ls;
            pgb_st8(a, x < D ? pgb_ld8(pfx + x) : (p.sfx >> (8u * (uint32_t)(x - D))) & 0xFFu);
        }
    }
    if (b0 < b1) {
        // <= 15 GT bytes in front of the first aligned chunk (lanes 0-15) and <= 15 GT bytes + the
        // newline (pfile.rs:190) behind the last one (lanes 16-31), in one pass
        const uint64_t a = lane < 16 ? a_gs + lane : b1 + (lane - 16u);
        const uint64_t end = lane < 16 ? b0 : a_le;
        if (a < end && a >= t0 && a < t1) pgb_st8(a, pgb_gt_byte<GATHER>(p, row, K4, a - a_gs));
    } else { // no aligned chunk inside the GT text: the whole GT region byte by byte
        const uint64_t lo = a_gs > t0 ? a_gs : t0, hi = a_le < t1 ? a_le : t1;
        for (uint64_t a = lo + lane; a < hi; a += 32) pgb_st8(a, pgb_gt_byte<GATHER>(p, row, K4, a - a_gs));
    }
    const uint64_t lo = b0 > t0 ? b0 : t0, hi = b1 < t1 ? b1 : t1;
    if (lo < hi) pgb_k2_body<GATHER, HINT, REPL>(p, row, a_gs, lo, hi, lane, lut4);
}

// Prefix bytes (pfile.rs:157-161): P bytes from src to the line start a_ls (al = a_ls & 15).
// Short prefixes go byte by byte; from 64 bytes on (real 1000G .pvar rows average 163) the
// destination-aligned 16-byte chunks are copied with aligned 32-bit source reads re-phased by the
// source/destination misalignment (constant per line), and only the <= 15 bytes on either side are
// byte stores.
PGB_DEV void pgb_copy_prefix(const uint8_t *src, uint64_t a_ls, uint32_t al, uint32_t PS, uint32_t sfx, uint32_t sfx_len,
                             uint32_t lane, int hint) {
    // the last sfx_len bytes of the prefix are the constant suffix ("\tGT" when src is a raw .pvar row)
    const uint32_t P = PS - sfx_len;
    if (lane < sfx_len) pgb_st8(a_ls + P + lane, (sfx >> (8u * lane)) & 0xFFu);
    if (P < 64u) {
        for (uint32_t x = lane; x < P; x += 32) pgb_st8(a_ls + x, pgb_ld8(src + x));
        return;
    }
    const uint32_t xa0 = (16u - al) & 15u;          // first 16-byte-aligned destination offset
    const uint32_t xa1 = ((al + P) & ~15u) - al;    // end of the last whole chunk (P >= 64: xa1 > xa0)
    { // the ragged ends: lanes 0-15 in front, lanes 16-31 behind
        const uint32_t x = lane < 16 ? lane : xa1 + (lane - 16u);
        const uint32_t end = lane < 16 ? xa0 : P;
        if (x < end) pgb_st8(a_ls + x, pgb_ld8(src + x));
    }
    const uint32_t mis = (uint32_t)(uintptr_t)(src + xa0) & 3u; // same for every chunk of the line
    const uint32_t sh8 = mis * 8u;
    for (uint32_t x = xa0 + 16u * lane; x < xa1; x += 512u) {
        const uint32_t *w = (const uint32_t *)(src + x - mis); // 4-byte aligned
        const uint32_t w0 = pgb_ld32(w), w1 = pgb_ld32(w + 1), w2 = pgb_ld32(w + 2), w3 = pgb_ld32(w + 3);
        const uint32_t w4 = mis ? pgb_ld32(w + 4) : 0u; // its first byte is inside the prefix when mis != 0
        pgb_st16(a_ls + x, pgb_funnel_r(w0, w1, sh8), pgb_funnel_r(w1, w2, sh8), pgb_funnel_r(w2, w3, sh8),
                 pgb_funnel_r(w3, w4, sh8), hint);
    }
}

// One-byte GT lookup with line-relative 32-bit offsets (g = offset from the start of the GT text).
template <bool GATHER>
PGB_DEV uint32_t pgb_gt_byte32(const pgb_k2_params &p, const uint8_t *row, uint32_t K4, uint32_t g) {
    if (g >= K4) return '\n';
    const uint32_t f = g >> 2;
    const uint32_t s = GATHER ? pgb_ld32(p.kidx + f) : f;
    return (pgb_gt_word(pgb_code(row, s)) >> ((g & 3u) * 8u)) & 0xFFu;
}

// A whole line in one item (launches whose every line fits in tile_bytes: the chr22 shapes).
// Same bytes as pgb_k2_item, but every position is a 32-bit offset x from the line start
// (address = a_ls + x), there is no tile to clip against, and the short body keeps one 64-bit
// store pointer per lane — the per-line bookkeeping is what bounds short lines (ncu: 395 warp
// instructions per 1 041-byte line on the gather workload before this path existed).
template <bool GATHER, int HINT, int REPL>
PGB_DEV void pgb_k2_line(const pgb_k2_params &p, const pgb_line_meta &m, uint32_t lane, const pgb_u4 *lut4) {
    const uint32_t P = m.pfx_len;
    const uint32_t K4 = 4u * p.K; // < tile_bytes <= 128 MiB
    const uint64_t a_ls = (uint64_t)(uintptr_t)p.out + m.line_off;
    const uint32_t al = (uint32_t)a_ls & 15u;
    const uint32_t x_gs = P, x_ge = P + K4, x_le = x_ge + 1u;
    const uint8_t *row = p.records + m.rec_off;
    pgb_copy_prefix(p.prefix_blob + m.pfx_off, a_ls, al, P, p.sfx, p.sfx_len, lane, HINT);
    // [xb0, xb1): the part of the GT text made of whole 16-byte-aligned chunks
    const uint32_t yb0 = (al + x_gs + 15u) & ~15u, yb1 = (al + x_ge) & ~15u; // relative to a_ls - al
    const uint32_t xb0 = yb0 - al, xb1 = yb1 - al;                          // (xb1 is only used when yb0 < yb1)
    if (yb0 >= yb1) { // none: the GT text and the newline byte by byte
        for (uint32_t x = x_gs + lane; x < x_le; x += 32) pgb_st8(a_ls + x, pgb_gt_byte32<GATHER>(p, row, K4, x - x_gs));
        return;
    }
    { // <= 15 GT bytes in front (lanes 0-15), <= 15 GT bytes + the newline behind (lanes 16-31)
        const uint32_t x = lane < 16 ? x_gs + lane : xb1 + (lane - 16u);
        const uint32_t end = lane < 16 ? xb0 : x_le;
        if (x < end) pgb_st8(a_ls + x, pgb_gt_byte32<GATHER>(p, row, K4, x - x_gs));
    }
    const uint32_t n_chunks = (xb1 - xb0) >> 4;
    if (n_chunks > 128u) { // long body: the row-aligned pipelined path
        pgb_k2_body<GATHER, HINT, REPL>(p, row, a_ls + x_gs, a_ls + xb0, a_ls + xb1, lane, lut4);
        return;
    }
    const uint32_t delta = (0u - (al + x_gs)) & 15u; // (A - a_gs) & 15 for any 16-aligned A
    const uint32_t r8 = (delta & 3u) * 8u;
    const uint32_t sh = (delta >> 2) * 2u;
    const pgb_u4 *l4 = lut4 + (lane & (uint32_t)(REPL - 1));
    uint64_t A = a_ls + xb0 + 16u * lane; // this lane's first chunk
    uint32_t q = xb0 - x_gs + 16u * lane; // its GT offset
    for (uint32_t c = lane; c < n_chunks; c += 64, A += 1024, q += 1024) {
        const bool two = c + 32 < n_chunks;
        const void *sa = GATHER ? (const void *)(p.kidx + (q >> 2)) : (const void *)(row + (q >> 4));
        const void *sb = GATHER ? (const void *)(p.kidx + ((q + 512u) >> 2)) : (const void *)(row + ((q + 512u) >> 4));
        uint32_t wa, wb = 0u;
        if (GATHER && p.kidx_vec) {
            wa = pgb_chunk_codes<GATHER, true>(row, sa, sh);
            if (two) wb = pgb_chunk_codes<GATHER, true>(row, sb, sh);
        } else {
            wa = pgb_chunk_codes<GATHER>(row, sa, sh);
            if (two) wb = pgb_chunk_codes<GATHER>(row, sb, sh);
        }
        pgb_emit_chunk<REPL>(A, wa, r8, l4, HINT);
        if (two) pgb_emit_chunk<REPL>(A + 512u, wb, r8, l4, HINT);
    }
}
