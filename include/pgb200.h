/*
 * pgb200.h — C ABI of libpgb200.so: the B200-native replacement for the genotype
 * export hot path of teoremma/pgen-rs (decode 2-bit hardcalls -> gather kept samples
 * -> VCF GT text).
 *
 * The reference has no FFI boundary of its own: the seam is the Rust method
 * Pfile::output_vcf (src/pfile.rs:104-194).  A Rust maintainer binds the functions
 * below with an `extern "C"` block (INTEGRATION.md shows the exact shim) and calls
 * pgb_export_gt_vcf in place of the loop at src/pfile.rs:149-192.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary;
 *   - every function returns PGB_OK (0) or a negative pgb_status; nothing unwinds
 *     across the boundary (the Rust side re-creates the reference's panic semantics
 *     with .unwrap(), exit code 101);
 *   - the caller owns every input array for the duration of the call; calls are
 *     synchronous; a call may use internal threads and CUDA streams;
 *   - there is NO CPU fallback: without a usable CUDA device the export functions
 *     return PGB_E_NO_DEVICE.
 */
#ifndef PGB200_H
#define PGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGB_ABI_VERSION 1

typedef enum pgb_status {
    PGB_OK = 0,
    PGB_E_IO = -1,         /* File::open/seek/read_exact .unwrap()        src/pfile.rs:41,149,169,170 */
    PGB_E_MAGIC = -2,      /* assert_eq!(buf, [0x6C, 0x1B])               src/pfile.rs:44-47          */
    PGB_E_MODE = -3,       /* assert!(storage_mode == 0x02)               src/pfile.rs:49-53          */
    PGB_E_FLAGS = -4,      /* assert_eq!(buf, [0x40])                     src/pfile.rs:65-69          */
    PGB_E_ARG = -5,        /* malformed arguments (NULL, unsorted or duplicate sample list, ...)      */
    PGB_E_RANGE = -6,      /* variant/sample index outside the file       src/pfile.rs:170,173        */
    PGB_E_NO_DEVICE = -7,  /* no usable CUDA device — there is no CPU fallback                       */
    PGB_E_CUDA = -8,       /* a CUDA runtime call or kernel failed (see pgb_last_error)               */
    PGB_E_NOMEM = -9,      /* host or device allocation failed                                       */
    PGB_E_NO_HEADER = -10, /* .pvar/.psam without a leading '#' line      src/pfile.rs:217            */
    PGB_E_NO_IID = -11,    /* "IID not among the headers of ..."          src/pfile.rs:125-126        */
    PGB_E_CSV = -12,       /* ragged row / unsupported quoting (csv::Error -> unwrap, src/main.rs:123) */
    PGB_E_EXPR = -13,      /* evalexpr error (.unwrap() at src/pfile.rs:94,97,328)                    */
    PGB_E_SPACE = -14      /* caller-provided output buffer too small                                */
} pgb_status;

/* Opaque handle: owns the .pgen descriptor (or a borrowed host image) and the parsed
 * 12-byte header.  Replaces `struct Pfile` (src/pfile.rs:19-23) for the hot path; it also
 * caches per-device pinned/device buffers between calls. */
typedef struct pgb_file pgb_file;

typedef struct pgb_stats {
    uint64_t n_lines;        /* kept variants written                                   */
    uint64_t n_kept_samples; /* K                                                       */
    uint64_t genotypes;      /* n_lines * K                                             */
    uint64_t bytes_out;      /* VCF body bytes produced                                 */
    uint64_t bytes_h2d;      /* host->device bytes copied                               */
    uint64_t bytes_d2h;      /* device->host bytes copied                               */
    uint64_t kernel_launches;/* kernels of this library launched by the call            */
    double device_ms;        /* summed K0+K1+K2 device time (CUDA events), max over devices */
    double e2e_ms;           /* wall time of the call                                   */
    int32_t n_devices;
    int32_t n_chunks;
} pgb_stats;

/* ---- container: src/pfile.rs:38-76 (Pfile::from_prefix), :196-200 (record size) ---- */

/* Opens PATH (a .pgen), checks magic / storage mode 0x02 / flag byte 0x40 in the
 * reference's order. */
int pgb_open(const char *pgen_path, pgb_file **out);
/* Same checks on a host-memory image of a .pgen (header + records).  The image is
 * borrowed, not copied; if it is page-locked (cudaHostAlloc/cudaHostRegister) it is
 * DMA'd from directly. */
int pgb_open_mem(const void *pgen_image, uint64_t image_bytes, pgb_file **out);
/* EXTENSION beyond the reference (whose Pfile asserts storage mode 0x02): opens a standard-format
 * .pgen (storage mode 0x10).  The header walk of src/pgen.rs:18-258 (pgb_pgen10_index) becomes the
 * record index that K1 turns into the device-side record index; the export functions then accept
 * the variants stored as plain 2-bit hardcall records (record type 0, ceil(2N/8) bytes) and fail
 * with PGB_E_MODE for a kept variant stored in any compressed form. */
int pgb_open_standard(const char *pgen_path, pgb_file **out);
void pgb_dims(const pgb_file *f, uint32_t *n_variants, uint32_t *n_samples, uint32_t *record_bytes);
void pgb_close(pgb_file *f);

/* variant_record_size (src/pfile.rs:196-200) and the record offset of src/pfile.rs:165,
 * computed in u64 (the reference's u32 product wraps for files > 4 GiB). */
uint32_t pgb_record_bytes(uint32_t n_samples);
uint64_t pgb_record_offset(uint64_t var_idx, uint32_t record_bytes);

/* ---- the hot path: replaces the loop at src/pfile.rs:149-192 ---- */

/* Writes, for every kept variant i (in the order given):
 *     prefix_blob[prefix_off[i] .. prefix_off[i+1])      bytes of src/pfile.rs:157-161
 *     ("\t" GT(code(var_idx[i], s)))  for s in sam_idx    src/pfile.rs:171-188
 *     "\n"                                                src/pfile.rs:190
 * to out_fd at its current position (the VCF header of src/pfile.rs:139-146 has already
 * been written by the caller).
 *   var_idx  : n_var file row indices (what filter_metadata returns, src/pfile.rs:127);
 *              NULL => rows 0..n_var-1.
 *   sam_idx  : n_sam strictly ascending sample indices (src/pfile.rs:128); NULL => all
 *              samples (n_sam ignored).  Non-NULL with n_sam == 0 => no samples.
 *   prefix_off: n_var+1 ascending offsets into prefix_blob.
 *   device_ids/n_devices: CUDA devices to shard contiguous variant ranges over
 *              (NULL/0 => device 0).
 *   stats    : optional. */
int pgb_export_gt_vcf(pgb_file *f, const uint32_t *var_idx, uint64_t n_var, const uint32_t *sam_idx, uint64_t n_sam,
                      const uint8_t *prefix_blob, const uint64_t *prefix_off, int out_fd, const int *device_ids,
                      int n_devices, pgb_stats *stats);

/* Same, into a caller-provided host buffer (page-locked buffers are DMA'd into
 * directly).  *out_len receives the body size; PGB_E_SPACE if out_cap is too small. */
int pgb_export_gt_vcf_mem(pgb_file *f, const uint32_t *var_idx, uint64_t n_var, const uint32_t *sam_idx,
                          uint64_t n_sam, const uint8_t *prefix_blob, const uint64_t *prefix_off, uint8_t *out_buf,
                          uint64_t out_cap, uint64_t *out_len, const int *device_ids, int n_devices, pgb_stats *stats);

/* The same export with the line prefixes BUILT ON THE DEVICE from the raw .pvar image: kept variant i's prefix is
 * pvar_text[row_off[i] .. row_off[i] + row_len[i]) — the row as it is in the file, without its line terminator:
 * for a table without csv quoting that is every field followed by '\t' except the last (src/pfile.rs:157-160) —
 * followed by "\tGT" (src/pfile.rs:160-161), which K2 appends.  No per-row host pass builds a prefix blob; the
 * host only hands over the offsets its .pvar reader already has.  row_off need not be ascending. */
int pgb_export_gt_vcf_rows(pgb_file *f, const uint32_t *var_idx, uint64_t n_var, const uint32_t *sam_idx, uint64_t n_sam,
                           const uint8_t *pvar_text, uint64_t pvar_bytes, const uint64_t *row_off, const uint32_t *row_len,
                           int out_fd, const int *device_ids, int n_devices, pgb_stats *stats);
int pgb_export_gt_vcf_rows_mem(pgb_file *f, const uint32_t *var_idx, uint64_t n_var, const uint32_t *sam_idx,
                               uint64_t n_sam, const uint8_t *pvar_text, uint64_t pvar_bytes, const uint64_t *row_off,
                               const uint32_t *row_len, uint8_t *out_buf, uint64_t out_cap, uint64_t *out_len,
                               const int *device_ids, int n_devices, pgb_stats *stats);

/* The multi-GPU partition pgb_export_gt_vcf uses: the kept-variant list cut into n_shards
 * contiguous ranges balanced by output bytes.  line_begin (n_shards+1 entries) receives the
 * range boundaries, byte_begin (optional, n_shards+1) the body offset at which each range
 * starts; shard g owns lines [line_begin[g], line_begin[g+1]).  CPU only. */
int pgb_shard_plan(uint64_t n_var, uint64_t n_kept_samples, const uint64_t *prefix_off, int n_shards,
                   uint64_t *line_begin, uint64_t *byte_begin);

/* The per-device staging buffers, streams and events the export functions use are cached
 * process-wide across calls and handles; this frees the ones no export is using. */
void pgb_release_buffers(void);

/* Body size in bytes for the given selection: sum(P_i) + n_var * (4K + 1). */
uint64_t pgb_body_bytes(uint64_t n_var, uint64_t n_kept_samples, const uint64_t *prefix_off);

/* ---- host mirror of the reference's Pfile API (src/pfile.rs:78-194), C++ inside ---- */

/* Pfile::output_vcf(sam_query, var_query, filename) (src/pfile.rs:104-194): evaluates the
 * two include-expressions on the CPU (src/pfile.rs:312-335), writes the header, then
 * calls pgb_export_gt_vcf.  NULL query => keep all. */
int pgb_pfile_output_vcf(const char *pfile_prefix, const char *sam_query, const char *var_query, const char *out_path,
                         const int *device_ids, int n_devices, pgb_stats *stats);

/* Pfile::query_metadata (src/pfile.rs:78-102) on the .pvar (samples == 0) or .psam
 * (samples != 0); lines go to out_fd. */
int pgb_pfile_query(const char *pfile_prefix, const char *fstring, const char *query, int samples, int out_fd);

/* CPU-only planning stage of output_vcf (no GPU needed): the selections, VCF header and
 * line prefixes that pgb_export_gt_vcf consumes.  Free with pgb_plan_free. */
typedef struct pgb_plan pgb_plan;
int pgb_plan_vcf(const char *pfile_prefix, const char *sam_query, const char *var_query, pgb_plan **out);
void pgb_plan_free(pgb_plan *p);
uint64_t pgb_plan_n_var(const pgb_plan *p);
uint64_t pgb_plan_n_sam(const pgb_plan *p);
const uint32_t *pgb_plan_var_idx(const pgb_plan *p);
const uint32_t *pgb_plan_sam_idx(const pgb_plan *p);
const uint8_t *pgb_plan_header(const pgb_plan *p, uint64_t *len);
/* the raw .pvar image and the kept rows in it: what output_vcf hands to pgb_export_gt_vcf_rows */
const uint8_t *pgb_plan_pvar_text(const pgb_plan *p, uint64_t *len);
const uint64_t *pgb_plan_row_off(const pgb_plan *p);
const uint32_t *pgb_plan_row_len(const pgb_plan *p);
/* the finished prefixes (row + "\tGT") as a blob with n_var + 1 offsets: built on first use, for callers of
 * pgb_export_gt_vcf and for tests; output_vcf itself no longer builds it */
const uint8_t *pgb_plan_prefix_blob(const pgb_plan *p, uint64_t *len);
const uint64_t *pgb_plan_prefix_off(const pgb_plan *p);

/* Standard-format (.pgen storage mode 0x10) header walk — what Pgen::from_file_path
 * (src/pgen.rs:18-137) computes, plus a correct per-variant record index
 * (off[v] = block_off[v / 65536] + sum of earlier record lengths in the block).
 * rec_off/rec_type/rec_len may be NULL; otherwise they hold n_variants entries
 * (rec_off: n_variants + 1). */
typedef struct pgb_pgen10_info {
    uint32_t n_variants, n_samples;
    uint8_t storage_mode, header_format;
    uint8_t record_type_bits, record_length_bytes, allele_count_bytes, provisional_ref_storage;
    uint32_t variant_block_count;
    uint64_t variant_block_offsets_offset, main_header_body_offset, main_header_body_size, variant_records_offset;
} pgb_pgen10_info;
int pgb_pgen10_index(const char *pgen_path, pgb_pgen10_info *info, uint64_t *rec_off, uint8_t *rec_type,
                     uint32_t *rec_len);

/* ---- device-resident entry points (all pointers are DEVICE pointers on the current
 *      device; `stream` is a cudaStream_t passed as void*) ---- */

/* Per-line metadata produced by K1 and consumed by K2 (32 bytes, 16-byte aligned). */
typedef struct pgb_line_meta {
    uint64_t line_off; /* byte offset of the line in the output buffer (exclusive prefix sum) */
    uint64_t rec_off;  /* device record index: byte offset of the variant's record from `records` */
    uint64_t pfx_off;  /* byte offset of the line prefix in prefix_blob                         */
    uint32_t pfx_len;  /* P_v                                                                   */
    uint32_t reserved; /* bit 0: the launch's prefixes are packed back to back in prefix_blob (K1 sets it
                          unless it was given explicit prefix lengths); other bits 0                */
} pgb_line_meta;

/* K0: sample keep-mask (one byte per sample, non-zero = keep) -> ascending kept-index
 * list via warp ballot/popc.  kidx must hold n_samples + 8 entries; *count (device)
 * receives K. */
int pgb_dev_compact_samples(const uint8_t *keep_mask, uint32_t n_samples, uint32_t *kidx, uint32_t *count,
                            void *stream);

/* Scratch bytes K1 needs for n_lines lines. */
uint64_t pgb_dev_index_scratch_bytes(uint64_t n_lines);

/* K1: record index + line-length exclusive prefix sum.
 *   var_row   : n_lines row indices relative to `records` (NULL => 0..n_lines-1)
 *   prefix_off: n_lines+1 offsets (rebased so that prefix_off[0] addresses prefix_blob[0] is
 *               not required: pfx_off = prefix_off[i] - prefix_base)
 *   meta      : n_lines+1 entries out (entry n_lines carries the total in line_off)
 *   pitch     : bytes between consecutive records in `records` (>= record_bytes). */
int pgb_dev_index_lines(const uint32_t *var_row, const uint64_t *prefix_off, uint64_t prefix_base, uint64_t n_lines,
                        uint32_t n_kept, uint64_t pitch, pgb_line_meta *meta, void *scratch, void *stream);

/* K1 with an explicit record index: rec_off[i] = byte offset of line i's record from `records`
 * (what the standard-format header walk yields, src/pgen.rs:100-258) instead of row * pitch. */
int pgb_dev_index_lines_off(const uint64_t *rec_off, const uint64_t *prefix_off, uint64_t prefix_base, uint64_t n_lines,
                            uint32_t n_kept, pgb_line_meta *meta, void *scratch, void *stream);

/* K2: decode + gather + format.  kidx NULL => all n_samples samples (n_kept must equal
 * n_samples); otherwise n_kept entries (+8 padding; vectorised reads when 16-byte aligned).
 * `records` must be readable for 16 bytes past its last record, and the 128-byte-aligned
 * blocks containing its first and last byte must lie inside the allocation (true for any
 * pointer into a cudaMalloc allocation): K2 prefetches record slices by cache line.
 * variant = 0 selects the measured-best configuration; other values are tuning knobs
 * (store hint, LUT copies, items per warp, tile size — see pgb_kernels.cu). */
int pgb_dev_format_lines(const uint8_t *records, const pgb_line_meta *meta, uint64_t n_lines,
                         const uint8_t *prefix_blob, const uint32_t *kidx, uint32_t n_kept, uint32_t max_prefix_len,
                         uint8_t *out, int variant, void *stream);

/* K1, general form: the record index is rec_off[i] when rec_off is non-NULL, else var_row[i] * pitch
 * (var_row NULL => i * pitch).  The prefix of line i is prefix_len[i] bytes at prefix_off[i] when
 * prefix_len is non-NULL (rows of a raw .pvar image: prefix_off then holds n_lines entries and need not
 * be contiguous), else prefix_off[i+1] - prefix_off[i] bytes; suffix_len (0..4) constant bytes are
 * appended to every prefix by K2 and counted in pgb_line_meta::pfx_len. */
int pgb_dev_index_lines_ex(const uint32_t *var_row, const uint64_t *rec_off, uint64_t pitch, const uint64_t *prefix_off,
                           const uint32_t *prefix_len, uint32_t suffix_len, uint64_t prefix_base, uint64_t n_lines,
                           uint32_t n_kept, pgb_line_meta *meta, void *scratch, void *stream);

/* K2, general form.  record_bytes = ceil(2 * n_samples / 8) of the file (needed to size the shared-memory
 * staging of the batch path when kidx is non-NULL; 0 = unknown => per-line path).  `suffix`: suffix_len
 * (0..4) bytes, little-endian, written after the prefix_blob bytes of every line — "\tGT" (0x0054 4709,
 * length 3) turns a raw .pvar row into the line prefix of src/pfile.rs:157-161 on the device.
 * The batch path fetches records and prefixes with bulk copies of whole 16-byte blocks: besides the 16 readable
 * bytes after the last record, the 16-byte blocks containing the first and the last byte of every prefix must be
 * readable (true for any pointer into a cudaMalloc allocation that does not end inside the last block). */
int pgb_dev_format_lines_ex(const uint8_t *records, uint32_t record_bytes, const pgb_line_meta *meta, uint64_t n_lines,
                            const uint8_t *prefix_blob, uint32_t suffix, uint32_t suffix_len, const uint32_t *kidx,
                            uint32_t n_kept, uint32_t max_prefix_len, uint8_t *out, int variant, void *stream);

/* Synthetic records (tools/synth.py documents the integer hash): rows row0..row0+n_rows-1. */
int pgb_dev_synth_records(uint8_t *records, uint64_t pitch, uint64_t seed, uint64_t row0, uint64_t n_rows,
                          uint32_t n_samples, void *stream);

/* A cheaper synthetic generator for the biobank shape (tools/synth.py: synth_records_fast): one 32-bit hash per
 * four record bytes, uniform over the four genotype codes, padding samples 0. */
int pgb_dev_synth_records_fast(uint8_t *records, uint64_t pitch, uint32_t seed, uint64_t row0, uint64_t n_rows,
                               uint32_t n_samples, void *stream);

/* Store-only calibration kernel (16-byte streaming stores of a constant). */
int pgb_dev_fill(uint8_t *dst, uint64_t bytes, int variant, void *stream);

/* Store-only calibration with K2-like write patterns (tools/fill_sweep.py): segments of
 * seg_rows x 512 bytes, `coop` (1/2/4/8) warps of a CTA per segment in bursts of `burst` rows,
 * n_ctas persistent CTAs; seg_rows == 0 => grid-stride fill with `burst` strided stores per thread. */
int pgb_dev_fill_pattern(uint8_t *dst, uint64_t bytes, uint32_t seg_rows, uint32_t coop, uint32_t burst,
                         uint32_t n_ctas, int hint, void *stream);

int pgb_device_count(void);
const char *pgb_strerror(int status);
/* Thread-local detail of the last failure on this thread (empty string if none). */
const char *pgb_last_error(void);
int pgb_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PGB200_H */
