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
