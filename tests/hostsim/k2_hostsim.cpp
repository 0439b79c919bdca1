// k2_hostsim.cpp — TEST-ONLY lane-by-lane host simulation of K2 (pgen-rs_b200/csrc/
// k2_core.cuh compiled with PGB_HOSTSIM).  Lets the CPU test suite check the kernel's
// addressing and byte logic against the oracle for every alignment phase without a GPU.
// It is not part of libpgb200.so and is never used by the product.
#define PGB_HOSTSIM 1
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../pgen-rs_b200/csrc/k2_core.cuh"

extern "C" int sim_format_lines(const uint8_t *records, uint64_t pitch, const uint32_t *var_row, uint64_t n_lines,
                                const uint8_t *prefix_blob, const uint64_t *prefix_off, const uint32_t *kidx,
                                uint32_t K, uint8_t *out, int variant) {
    std::vector<pgb_line_meta> meta(n_lines + 1);
    uint64_t off = 0;
    uint32_t maxp = 0;
    for (uint64_t i = 0; i < n_lines; i++) {
        uint64_t P = prefix_off[i + 1] - prefix_off[i];
        meta[i].line_off = off;
        meta[i].rec_off = (var_row ? var_row[i] : i) * pitch;
        meta[i].pfx_off = prefix_off[i] - prefix_off[0];
        meta[i].pfx_len = (uint32_t)P;
        meta[i].reserved = 0;
        if (P > maxp) maxp = (uint32_t)P;
        off += P + 4ull * K + 1;
    }
    meta[n_lines].line_off = off;
    pgb_k2_params p;
    p.records = records;
    p.meta = meta.data();
    p.prefix_blob = prefix_blob + prefix_off[0];
    p.kidx = kidx;
    p.out = out;
    p.n_lines = n_lines;
    p.K = K;
    p.store_hint = variant & 0xF;
    const int decode = (variant >> 4) & 0xF;
    int unroll = (variant >> 8) & 0xF;
    if (!unroll) unroll = 4;
    const int tsel = (variant >> 12) & 0xF;
    p.tile_bytes = tsel ? (4096u << tsel) : 16384u;
    const uint64_t max_line = (uint64_t)maxp + 4ull * K + 1ull;
    p.n_tiles = (uint32_t)((max_line + 511ull + p.tile_bytes - 1) / p.tile_bytes);
    pgb_u4 lut[256];
    for (uint32_t b = 0; b < 256; b++) lut[b] = pgb_lut_entry(b);
    for (uint64_t line = 0; line < n_lines; line++)
        for (uint32_t tile = 0; tile < p.n_tiles; tile++)
            for (uint32_t lane = 0; lane < 32; lane++) {
                const bool g = kidx != nullptr, l = decode == 1;
#define C(G, U, L) pgb_k2_item<G, U, L>(p, line, tile, lane, lut)
                if (unroll == 1) { if (g) { if (l) C(true, 1, true); else C(true, 1, false); } else { if (l) C(false, 1, true); else C(false, 1, false); } }
                else if (unroll == 2) { if (g) { if (l) C(true, 2, true); else C(true, 2, false); } else { if (l) C(false, 2, true); else C(false, 2, false); } }
                else { if (g) { if (l) C(true, 4, true); else C(true, 4, false); } else { if (l) C(false, 4, true); else C(false, 4, false); } }
#undef C
            }
    return 0;
}
