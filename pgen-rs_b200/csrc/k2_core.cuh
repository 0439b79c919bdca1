// k2_core.cuh — per-lane body of K2 (decode + gather + format), written so that the
// identical source compiles (a) as the sm_100a device code of libpgb200.so and (b) with
// PGB_HOSTSIM as a lane-by-lane host simulation used ONLY by tests/ to validate the
// addressing/byte logic before spending GPU time.  The simulation is never linked into
// the product library.
//
// What it replaces: the per-variant loop body of Pfile::output_vcf,
// /root/reference/src/pfile.rs:156-192 —
//   :157-161 prefix fields + "GT"       -> head bytes copied from prefix_blob
//   :165-170 record offset/seek/read    -> meta.rec_off (device record index from K1)
//   :171-175 per-sample 2-bit extract   -> body chunks (keep-all: two row bytes per 16 B
//                                          of text; gather: kidx lookups)
//   :177-188 code -> "\t0/0" ...        -> gt_word() via two PRMTs (or a smem LUT)
//   :190     "\n"                       -> tail bytes
//
// Geometry of one output line at absolute address a_ls (P = prefix bytes, K kept samples):
//   [a_ls, a_gs)  prefix      a_gs = a_ls + P
//   [a_gs, a_ge)  GT text     a_ge = a_gs + 4K      (field f at a_gs + 4f: '\t' x '/' y)
//   [a_ge, a_le)  '\n'        a_le = a_ge + 1
// The warp writes   head = [a_ls, b0)   byte stores   (b0 = align_up(a_gs, 16))
//                   body = [b0, b1)     16-byte stores, 512 B per warp instruction, every
//                                       chunk wholly inside the GT text
//                   tail = [b1, a_le)   byte stores   (b1 = align_down(a_ge, 16))
// so every output byte has exactly one writer and no padding is ever emitted (the VCF
// must be bit-exact).  For a 16-byte chunk at A:  q = A - a_gs, field f0 = q >> 2, byte
// phase r = q & 3; (q & 15) is the same for every chunk of a line, so in the keep-all
// case the chunk needs row bytes (q >> 4) and (q >> 4) + 1 shifted by a per-line constant.
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
    int store_hint;      // 0 default, 1 .cs (streaming), 2 L1::no_allocate
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
PGB_DEV pgb_line_meta pgb_ld_meta(const pgb_line_meta *p) { return *p; }
PGB_DEV void pgb_st8(uint64_t a, uint32_t v) { *(uint8_t *)(uintptr_t)a = (uint8_t)v; }
PGB_DEV void pgb_st16(uint64_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w, int) {
    uint32_t t[4] = {x, y, z, w};
    memcpy((void *)(uintptr_t)a, t, 16);
}
PGB_DEV pgb_u4 pgb_lds_lut(const pgb_u4 *lut, uint32_t i) { return lut[i]; }
#else
#define PGB_DEV __device__ __forceinline__
PGB_DEV uint32_t pgb_prmt(uint32_t x, uint32_t y, uint32_t s) { return __byte_perm(x, y, s); }
PGB_DEV uint32_t pgb_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
PGB_DEV uint32_t pgb_ld8(const uint8_t *p) { return __ldg(p); }
PGB_DEV uint32_t pgb_ld32(const uint32_t *p) { return __ldg(p); }
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
PGB_DEV pgb_u4 pgb_lds_lut(const pgb_u4 *lut, uint32_t i) {
    uint4 v = *reinterpret_cast<const uint4 *>(lut + i);
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

// Byte `pos` of a line (generic, used for head/tail only).
template <bool GATHER>
PGB_DEV uint32_t pgb_line_byte(const pgb_k2_params &p, const uint8_t *row, const uint8_t *pfx, uint32_t P, uint64_t K4,
                               uint64_t pos) {
    if (pos < P) return pgb_ld8(pfx + pos);
    uint64_t g = pos - P;
    if (g >= K4) return '\n';
    uint32_t ch = (uint32_t)g & 3u;
    if (ch == 0u) return '\t';
    if (ch == 2u) return '/';
    uint32_t f = (uint32_t)(g >> 2);
    uint32_t s = GATHER ? pgb_ld32(p.kidx + f) : f;
    uint32_t c = pgb_code(row, s);
    if (c == 3u) return '.';
    if (ch == 1u) return c == 2u ? '1' : '0';
    return c == 0u ? '0' : '1';
}

template <bool GATHER, int UNROLL, bool LUT>
PGB_DEV void pgb_k2_item(const pgb_k2_params &p, uint64_t line, uint32_t tile, uint32_t lane, const pgb_u4 *lut) {
    const pgb_line_meta m = pgb_ld_meta(p.meta + line);
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
    if (b0 >= b1) { b0 = a_le; b1 = a_le; }
    const uint8_t *row = p.records + m.rec_off;
    const uint8_t *pfx = p.prefix_blob + m.pfx_off;

    { // head: prefix (+ up to 15 bytes of GT text, or the whole line when there is no body)
        const uint64_t lo = a_ls > t0 ? a_ls : t0, hi = b0 < t1 ? b0 : t1;
        for (uint64_t a = lo + lane; a < hi; a += 32) pgb_st8(a, pgb_line_byte<GATHER>(p, row, pfx, P, K4, a - a_ls));
    }
    { // tail: last partial chunk of GT text + '\n'
        const uint64_t lo = b1 > t0 ? b1 : t0, hi = a_le < t1 ? a_le : t1;
        for (uint64_t a = lo + lane; a < hi; a += 32) pgb_st8(a, pgb_line_byte<GATHER>(p, row, pfx, P, K4, a - a_ls));
    }
    const uint64_t lo = b0 > t0 ? b0 : t0, hi = b1 < t1 ? b1 : t1;
    if (lo >= hi) return;

    const uint32_t delta = (uint32_t)(0ull - a_gs) & 15u; // (A - a_gs) & 15 for any 16-aligned A
    const uint32_t r8 = (delta & 3u) * 8u;                // byte phase inside a field, in bits
    const uint32_t sh = (delta >> 2) * 2u;                // 2-bit phase inside a record byte
    for (uint64_t A = (lo & ~511ull) + (uint64_t)lane * 16u; A < hi; A += 512ull * UNROLL) {
        uint32_t w[UNROLL];
        bool ok[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const uint64_t Au = A + 512ull * u;
            ok[u] = Au >= lo && Au < hi;
            w[u] = 0;
            if (ok[u]) {
                const uint64_t q = Au - a_gs;
                if (!GATHER) {
                    const uint8_t *b = row + (q >> 4);
                    w[u] = (pgb_ld8(b) | (pgb_ld8(b + 1) << 8)) >> sh;
                } else {
                    const uint32_t *ki = p.kidx + (q >> 2);
                    uint32_t s0 = pgb_ld32(ki), s1 = pgb_ld32(ki + 1), s2 = pgb_ld32(ki + 2), s3 = pgb_ld32(ki + 3);
                    uint32_t s4 = pgb_ld32(ki + 4); // kidx is padded by 8 entries
                    w[u] = pgb_code(row, s0) | (pgb_code(row, s1) << 2) | (pgb_code(row, s2) << 4) |
                           (pgb_code(row, s3) << 6) | (pgb_code(row, s4) << 8);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            if (!ok[u]) continue;
            uint32_t W0, W1, W2, W3;
            if (LUT) {
                pgb_u4 e = pgb_lds_lut(lut, w[u] & 0xFFu);
                W0 = e.x; W1 = e.y; W2 = e.z; W3 = e.w;
            } else {
                W0 = pgb_gt_word(w[u] & 3u);
                W1 = pgb_gt_word((w[u] >> 2) & 3u);
                W2 = pgb_gt_word((w[u] >> 4) & 3u);
                W3 = pgb_gt_word((w[u] >> 6) & 3u);
            }
            const uint32_t W4 = pgb_gt_word((w[u] >> 8) & 3u);
            pgb_st16(A + 512ull * u, pgb_funnel_r(W0, W1, r8), pgb_funnel_r(W1, W2, r8), pgb_funnel_r(W2, W3, r8),
                     pgb_funnel_r(W3, W4, r8), p.store_hint);
        }
    }
}
