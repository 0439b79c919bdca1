// pgen10.cpp — standard-format (.pgen storage mode 0x10) header walk and record index.
//
// Mirrors what Pgen::from_file_path computes (/root/reference/src/pgen.rs:18-137,140-258)
// — header-format byte, variant-block offset table, main-header-body geometry — and goes one
// step further by producing the per-variant record index the reference only prints pieces
// of:  off[v] = block_off[v / 65536] + sum of the lengths of the earlier records in v's block.
//
// Known defects of the reference walker that are NOT reproduced:
//   * last block size is taken as M % 65536, which is 0 when M is a multiple of 65536
//     (pgen.rs:200-204); here it is M - 65536 * b;
//   * record lengths are read byte by byte (pgen.rs:236-240); here as little-endian
//     integers of record_length_bytes bytes;
//   * allele-count bytes are ignored in the body size (pgen.rs:116-133); here included.
#include <errno.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../csrc/pgb_internal.h"

extern "C" int pgb_pgen10_index(const char *pgen_path, pgb_pgen10_info *info, uint64_t *rec_off, uint8_t *rec_type,
                                uint32_t *rec_len) {
    pgb_clear_error();
    if (!pgen_path || !info) return PGB_E_ARG;
    memset(info, 0, sizeof *info);
    FILE *f = fopen(pgen_path, "rb");
    if (!f) { pgb_set_error("open %s: %s", pgen_path, strerror(errno)); return PGB_E_IO; }
    struct Closer { FILE *f; ~Closer() { fclose(f); } } closer{f};
    unsigned char h[12];
    if (fread(h, 1, 12, f) != 12) { pgb_set_error("short header"); return PGB_E_IO; }
    if (h[0] != 0x6C || h[1] != 0x1B) return PGB_E_MAGIC; // pgen.rs:28-30
    info->storage_mode = h[2];                            // pgen.rs:32-38 (printed, not asserted, there)
    if (h[2] != 0x10) { pgb_set_error("storage mode 0x%02x is not the standard format 0x10", h[2]); return PGB_E_MODE; }
    const uint32_t M = (uint32_t)h[3] | (uint32_t)h[4] << 8 | (uint32_t)h[5] << 16 | (uint32_t)h[6] << 24;
    const uint32_t N = (uint32_t)h[7] | (uint32_t)h[8] << 8 | (uint32_t)h[9] << 16 | (uint32_t)h[10] << 24;
    info->n_variants = M;
    info->n_samples = N;
    const uint8_t fmt = h[11]; // pgen.rs:50-67
    info->header_format = fmt;
    const uint32_t mode = fmt & 0xF;
    if (mode / 4 > 1) { pgb_set_error("unsupported record storage mode %u", mode); return PGB_E_FLAGS; }
    info->record_type_bits = mode / 4 == 0 ? 4 : 8;
    info->record_length_bytes = (uint8_t)(mode % 4 + 1);
    info->allele_count_bytes = (fmt >> 4) & 3;
    info->provisional_ref_storage = (fmt >> 6) & 3;
    if (info->provisional_ref_storage != 1) { // assert at pgen.rs:66
        pgb_set_error("provisional-REF storage %u (reference asserts 1)", info->provisional_ref_storage);
        return PGB_E_FLAGS;
    }
    const uint32_t B = (uint32_t)(((uint64_t)M + 65535) / 65536); // pgen.rs:100-102
    info->variant_block_count = B;
    info->variant_block_offsets_offset = 12;
    info->main_header_body_offset = 12 + 8ull * B; // pgen.rs:104-110
    std::vector<uint64_t> block_off(B);
    for (uint32_t b = 0; b < B; b++) {
        unsigned char o[8];
        if (fread(o, 1, 8, f) != 8) { pgb_set_error("short block-offset table"); return PGB_E_IO; }
        uint64_t v = 0;
        for (int k = 7; k >= 0; k--) v = v << 8 | o[k];
        if (b && v <= block_off[b - 1]) { // pgen.rs:140-165 asserts strictly ascending
            pgb_set_error("variant block offsets not ascending at block %u", b);
            return PGB_E_FLAGS;
        }
        block_off[b] = v;
    }
    uint64_t body = 0;
    for (uint32_t b = 0; b < B; b++) {
        const uint64_t cnt = b + 1 < B ? 65536 : (uint64_t)M - 65536ull * b;
        body += (cnt * info->record_type_bits + 7) / 8 + cnt * info->record_length_bytes + cnt * info->allele_count_bytes;
    }
    info->main_header_body_size = body;
    info->variant_records_offset = info->main_header_body_offset + body; // pgen.rs:135-137
    if (B && block_off[0] != info->variant_records_offset) {
        pgb_set_error("first block offset %llu != end of header %llu", (unsigned long long)block_off[0],
                      (unsigned long long)info->variant_records_offset);
        return PGB_E_FLAGS;
    }
    if (!rec_off && !rec_type && !rec_len) return PGB_OK;
    std::vector<unsigned char> buf;
    uint64_t v = 0;
    for (uint32_t b = 0; b < B; b++) {
        const uint64_t cnt = b + 1 < B ? 65536 : (uint64_t)M - 65536ull * b;
        const uint64_t tbytes = (cnt * info->record_type_bits + 7) / 8, lbytes = cnt * info->record_length_bytes;
        const uint64_t abytes = cnt * info->allele_count_bytes;
        buf.resize(tbytes + lbytes + abytes);
        if (fread(buf.data(), 1, buf.size(), f) != buf.size()) { pgb_set_error("short main header body"); return PGB_E_IO; }
        uint64_t off = block_off[b];
        for (uint64_t k = 0; k < cnt; k++, v++) {
            uint8_t t = info->record_type_bits == 8 ? buf[k] : (uint8_t)((buf[k / 2] >> ((k & 1) * 4)) & 0xF);
            uint32_t len = 0;
            for (int j = info->record_length_bytes - 1; j >= 0; j--) len = len << 8 | buf[tbytes + k * info->record_length_bytes + j];
            if (rec_off) rec_off[v] = off;
            if (rec_type) rec_type[v] = t;
            if (rec_len) rec_len[v] = len;
            off += len;
        }
        if (rec_off && b + 1 == B) rec_off[M] = off;
    }
    if (rec_off && B == 0) rec_off[0] = info->variant_records_offset;
    return PGB_OK;
}
