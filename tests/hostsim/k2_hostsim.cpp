// k2_hostsim.cpp — TEST-ONLY lane-by-lane host simulation of K2 (pgen-rs_b200/csrc/
// k2_core.cuh compiled with PGB_HOSTSIM).  Lets the CPU test suite check the kernel's
// addressing and byte logic against the oracle for every alignment phase without a GPU.
// It is not part of libpgb200.so and is never used by the product.
#define PGB_HOSTSIM 1
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../pgen-rs_b200/csrc/k2_batch.cuh"
#include "../../pgen-rs_b200/csrc/k2_core.cuh"

extern "C" int sim_format_lines_sfx(const uint8_t *records, uint64_t pitch, const uint32_t *var_row, uint64_t n_lines,
                                    const uint8_t *prefix_blob, const uint64_t *prefix_off, const uint32_t *kidx,
                                    uint32_t K, uint8_t *out, int variant, uint32_t sfx, uint32_t sfx_len);

// What K1 computes.  sfx_len > 0: the last sfx_len bytes of every prefix are the constant suffix `sfx`
// and prefix_blob holds only the bytes in front of it (prefix_off[i+1] - prefix_off[i] of them).
static uint32_t sim_meta(std::vector<pgb_line_meta> &meta, uint64_t pitch, const uint32_t *var_row, uint64_t n_lines,
                         const uint64_t *prefix_off, uint32_t K, uint32_t sfx_len, uint32_t packed = 1) {
    uint64_t off = 0;
    uint32_t maxp = 0;
    for (uint64_t i = 0; i < n_lines; i++) {
        uint64_t P = prefix_off[i + 1] - prefix_off[i] + sfx_len;
        meta[i].line_off = off;
        meta[i].rec_off = (var_row ? var_row[i] : i) * pitch;
        meta[i].pfx_off = prefix_off[i] - prefix_off[0];
        meta[i].pfx_len = (uint32_t)P;
        meta[i].reserved = packed;
        if (P > maxp) maxp = (uint32_t)P;
        off += P + 4ull * K + 1;
    }
    meta[n_lines].line_off = off;
    return maxp;
}

// The batch path (k2_batch.cuh): the persistent kernel's loop run for `grid` CTAs one after the other,
// phase by phase, with a loop over the CTA's threads standing in for each __syncthreads()-delimited
// phase; bulk copies are memcpy()s (so the two-stage pipeline degenerates to its data flow: the stage
// indices, tables and images alternate exactly as on the device).
extern "C" int sim_format_lines_batch(const uint8_t *records, uint64_t pitch, uint32_t R, const uint32_t *var_row,
                                      uint64_t n_lines, const uint8_t *prefix_blob, const uint64_t *prefix_off,
                                      const uint32_t *kidx, uint32_t K, uint8_t *out, uint32_t B, uint32_t sfx,
                                      uint32_t sfx_len, int flags) {
    // flags: bit 0 kidx_vec, bit 1 prefixes NOT marked as packed (one bulk copy per line instead of per batch)
    std::vector<pgb_line_meta> meta(n_lines + 1);
    const uint32_t maxp = sim_meta(meta, pitch, var_row, n_lines, prefix_off, K, sfx_len, (flags & 2) ? 0u : 1u);
    const bool gather = kidx != nullptr;
    pgb_k2b_params p;
    p.records = records;
    p.meta = meta.data();
    p.prefix_blob = prefix_blob + prefix_off[0];
    p.kidx = kidx;
    p.out = out;
    p.n_lines = n_lines;
    p.n_batches = (uint32_t)((n_lines + B - 1) / B);
    p.K = K;
    p.R = R;
    p.B = B;
    p.rowcap = pgb_k2b_align(R + 31u, 16);
    p.pcap = pgb_k2b_align(maxp - (maxp < sfx_len ? maxp : sfx_len) + 31u, 16);
    p.vcap = gather ? pgb_k2b_align((K + 3u) / 4u + 2u, 16) : 0u;
    const uint64_t max_line = (uint64_t)maxp + 4ull * K + 1ull;
    p.images = (flags & 16) ? 0u : (flags & 4) ? 1u : 2u;
    p.wcap = p.images ? pgb_k2b_align((uint32_t)(k2b_lines_per_warp(B) * max_line + 32u), 128) : 0u;
    p.outcap = K2B_WARPS * p.wcap;
    p.sfx = sfx;
    p.sfx_len = sfx_len;
    p.kidx_vec = (flags & 1) ? 1u : 0u;
    p.stages = (flags & 8) ? 3u : 2u;
    const pgb_k2b_layout L = pgb_k2b_smem_layout(p.B, p.rowcap, p.pcap, p.vcap, p.outcap, gather, p.images, p.stages);
    std::vector<uint8_t> smem_store(L.total + 256);
    uint8_t *smem = smem_store.data() + ((128 - ((uintptr_t)smem_store.data() & 127)) & 127);
    uint32_t span_lo = 0, span_len = K ? R : 0u;
    if (gather && K) {
        span_lo = kidx[0] >> 2;
        span_len = (kidx[K - 1] >> 2) + 1u - span_lo;
    }
    auto lines_of = [&](uint64_t b) -> uint32_t {
        if (b >= p.n_batches) return 0u;
        const uint64_t left = n_lines - b * B;
        return left < B ? (uint32_t)left : B;
    };
    const uint32_t LPW = k2b_lines_per_warp(B);
    auto issue = [&](uint64_t b, uint32_t stage) { // the producer warp
        const uint32_t nbl = lines_of(b);
        if (!nbl) return;
        for (uint32_t lane = 0; lane < 32; lane++) {
            pgb_line_meta m = {};
            uint64_t first_off = 0, next_off = 0;
            if (lane < nbl) {
                m = meta[b * B + lane];
                first_off = meta[b * B + lane / LPW * LPW].line_off;
                next_off = meta[b * B + lane + 1].line_off;
            }
            k2b_produce(p, smem, L, stage, nbl, lane, m, first_off, next_off, meta[b * B].line_off, meta[b * B + nbl].line_off,
                        meta[b * B].pfx_off, (meta[b * B].reserved & 1u) != 0, span_lo, span_len);
        }
    };
    const uint32_t grid = p.n_batches < 3 ? p.n_batches : 3; // a few "CTAs", several batches each
    for (uint32_t cta = 0; cta < grid; cta++) {
        memset(smem, 0xCD, L.total); // stale shared memory
        for (uint32_t t = 0; t < K2B_THREADS; t++) k2b_build_lut(smem, L, t);
        if (gather && K)
            for (uint32_t t = 0; t < K2B_CONSUMERS; t++) k2b_build_plan(p, smem, L, t, span_lo);
        uint64_t bt = cta;
        for (uint32_t s = 0; s + 1 < p.stages; s++) issue(bt + (uint64_t)s * grid, s); // the producer runs ahead
        for (uint32_t n = 0;; n++, bt += grid) {
            const uint32_t stage = n % p.stages, img = p.images > 1 ? n & 1u : 0u;
            const uint32_t nbl = lines_of(bt);
            if (!nbl) break;
            issue(bt + (uint64_t)(p.stages - 1) * grid, (n + p.stages - 1) % p.stages);
            for (uint32_t warp = 0; warp < K2B_WARPS; warp++) {
                uint32_t h[4];
                memcpy(h, smem + L.tab(stage) + 16u * warp, 16);
                const uint32_t ws = h[2], we = h[3];
                if ((ws == 0xFFFFFFFFu) != (warp * LPW >= nbl)) return -2;
                const uint64_t g_al = (uint64_t)(uintptr_t)p.out + (((uint64_t)h[1] << 32) | h[0]) - ws;
                if (ws != 0xFFFFFFFFu && (g_al & 15u)) return -1;
                const uint32_t nb = (K + 3u) >> 2;
                const uint32_t l1 = (warp + 1u) * LPW < nbl ? (warp + 1u) * LPW : nbl;
                for (uint32_t l = warp * LPW; l < l1; l++) {
                    for (uint32_t lane = 0; lane < 32; lane++) {
                        pgb_u4 preg[2] = {{0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}};
                        if (gather && nb <= 64u) {
                            const pgb_u4 *plan = reinterpret_cast<const pgb_u4 *>(smem + L.plan);
                            if (lane < nb) preg[0] = plan[lane];
                            if (lane + 32u < nb) preg[1] = plan[lane + 32u];
                        }
                        if (gather) {
                            if (p.images) k2b_line_gather<true, true>(p, smem, L, stage, img, g_al, l, warp, lane, preg[0], preg[1]);
                            else k2b_line_gather<true, false>(p, smem, L, stage, img, g_al, l, warp, lane, preg[0], preg[1]);
                        } else {
                            if (p.images) k2b_line_gather<false, true>(p, smem, L, stage, img, g_al, l, warp, lane, preg[0], preg[1]);
                            else k2b_line_gather<false, false>(p, smem, L, stage, img, g_al, l, warp, lane, preg[0], preg[1]);
                        }
                    }
                    for (uint32_t lane = 0; lane < 32; lane++) {
                        if (gather) {
                            if (p.images) k2b_line_format<true, true>(p, smem, L, stage, img, g_al, l, warp, lane);
                            else k2b_line_format<true, false>(p, smem, L, stage, img, g_al, l, warp, lane);
                        } else {
                            if (p.images) k2b_line_format<false, true>(p, smem, L, stage, img, g_al, l, warp, lane);
                            else k2b_line_format<false, false>(p, smem, L, stage, img, g_al, l, warp, lane);
                        }
                    }
                }
                if (ws == 0xFFFFFFFFu || !p.images) continue;
                const uint8_t *outb = smem + L.outb(img);
                const uint32_t h0 = (ws + 15u) & ~15u, h1 = we & ~15u;
                if (h0 < h1) memcpy((void *)(uintptr_t)(g_al + h0), outb + h0, h1 - h0); // the warp's bulk store
                for (uint32_t lane = 0; lane < 32; lane++) k2b_store_edges(g_al, outb, ws, we, lane);
            }
        }
    }
    return 0;
}

extern "C" int sim_format_lines(const uint8_t *records, uint64_t pitch, const uint32_t *var_row, uint64_t n_lines,
                                const uint8_t *prefix_blob, const uint64_t *prefix_off, const uint32_t *kidx,
                                uint32_t K, uint8_t *out, int variant) {
    return sim_format_lines_sfx(records, pitch, var_row, n_lines, prefix_blob, prefix_off, kidx, K, out, variant, 0, 0);
}

extern "C" int sim_format_lines_sfx(const uint8_t *records, uint64_t pitch, const uint32_t *var_row, uint64_t n_lines,
                                    const uint8_t *prefix_blob, const uint64_t *prefix_off, const uint32_t *kidx,
                                    uint32_t K, uint8_t *out, int variant, uint32_t sfx, uint32_t sfx_len) {
    std::vector<pgb_line_meta> meta(n_lines + 1);
    const uint32_t maxp = sim_meta(meta, pitch, var_row, n_lines, prefix_off, K, sfx_len);
    pgb_k2_params p;
    p.records = records;
    p.meta = meta.data();
    p.prefix_blob = prefix_blob + prefix_off[0];
    p.kidx = kidx;
    p.out = out;
    p.n_lines = n_lines;
    p.K = K;
    p.sfx = sfx;
    p.sfx_len = sfx_len;
    const int hint = variant & 0xF;
    const int lsel = (variant >> 4) & 0xF;
    const int single = lsel == 1 || (lsel == 0 && kidx == nullptr); // 1 => 8 LUT copies
    const int tsel = (variant >> 12) & 0xF;
    p.tile_bytes = tsel ? (4096u << tsel) : 16384u;
    const uint64_t max_line = (uint64_t)maxp + 4ull * K + 1ull;
    p.n_tiles = (uint32_t)((max_line + 511ull + p.tile_bytes - 1) / p.tile_bytes);
    p.row_bytes_hint = 0;
    p.kidx_vec = kidx && (variant & 1) == 0 ? 1u : 0u; // exercise both index-read forms
    // the tables the kernel builds in shared memory
    const int repl = single ? 8 : 1;
    std::vector<pgb_u4> lut4(256 * repl);
    for (uint32_t b = 0; b < 256; b++)
        for (int g = 0; g < repl; g++) lut4[b * repl + g] = pgb_lut_entry(b);
    const bool g = kidx != nullptr;
    for (uint64_t line = 0; line < n_lines; line++)
        for (uint32_t tile = 0; tile < p.n_tiles; tile++)
            for (uint32_t lane = 0; lane < 32; lane++) {
                const bool one = p.n_tiles == 1;
#define SIM(G, R)                                                                        \
    (one ? pgb_k2_line<G, 0, R>(p, meta[line], lane, lut4.data())                        \
         : pgb_k2_item<G, 0, R>(p, meta[line], tile, lane, lut4.data()))
                if (!single) {
                    if (g) SIM(true, 1); else SIM(false, 1);
                } else {
                    if (g) SIM(true, 8); else SIM(false, 8);
                }
#undef SIM
            }
    (void)hint;
    return 0;
}
