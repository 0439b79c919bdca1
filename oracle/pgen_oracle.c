/*
 * pgen_oracle.c — CPU restatement of pgen-rs's genotype export path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pgen-rs_b200/ may include, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker or as the timed
 * CPU baseline — never as the product.
 *
 * PARITY UNPINNED BY THE REFERENCE: teoremma/pgen-rs ships no tests, no golden
 * VCFs (*.vcf is git-ignored) and its .pgen fixtures are absent; it is Rust and
 * cannot be built in this image (no cargo/rustc, un-vendored clap/csv/evalexpr).
 * This restatement is pinned instead by (1) hand-derived known-answer vectors
 * from src/pfile.rs:172-183 (tests/golden/), (2) an independent numpy
 * implementation (oracle/oracle_np.py) and (3) the reference's shipped
 * basic1.pvar/.psam metadata with the sizes/counts its logs state.
 *
 * Every function cites the reference lines it follows (paths relative to the
 * reference checkout, i.e. src/pfile.rs unless stated).
 *
 * Two I/O modes for the body loop:
 *   ORC_IO_FAITHFUL  mirrors the reference's syscall pattern: one fresh
 *                    allocation + lseek + read per variant (pfile.rs:168-170),
 *                    an 8 KiB BufWriter (pfile.rs:137) and two buffered
 *                    appends per genotype (pfile.rs:186-187).  This is the mode
 *                    bench.py times as the CPU baseline.
 *   ORC_IO_BULK      same bytes, whole-record formatting into a large buffer
 *                    (used by tests for speed).
 */
#define _GNU_SOURCE
#define _FILE_OFFSET_BITS 64
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

#define ORC_OK 0
#define ORC_E_IO (-1)        /* File::open(..).unwrap(), read_exact(..).unwrap() */
#define ORC_E_MAGIC (-2)     /* assert_eq!(buf, [0x6C, 0x1B])     pfile.rs:47 */
#define ORC_E_MODE (-3)      /* assert!(storage_mode == 0x02)     pfile.rs:53 */
#define ORC_E_FLAGS (-4)     /* assert_eq!(buf, [0x40])           pfile.rs:69 */
#define ORC_E_NO_HEADER (-5) /* header_lines.pop().unwrap()       pfile.rs:217 */
#define ORC_E_NO_IID (-6)    /* panic!("IID not among ...")       pfile.rs:125 */
#define ORC_E_RAGGED (-7)    /* csv UnequalLengths (flexible(false)) */
#define ORC_E_QUOTE (-8)     /* input needs csv quoting rules: rejected, not guessed */
#define ORC_E_RANGE (-9)     /* slice index / read_exact past EOF  pfile.rs:170,173 */
#define ORC_E_NOMEM (-10)

#define ORC_FLAG_FAITHFUL_U32 1 /* record offset wraps like release Rust (pfile.rs:165) */
#define ORC_FLAG_IO_BULK 2      /* default is the faithful I/O pattern */

/* ---------------------------------------------------------------- header -- */

/* Pfile::from_prefix, pfile.rs:38-76: 2-byte magic, 1-byte storage mode
 * (must be 0x02), M u32 LE, N u32 LE, 1 flag byte (must be 0x40). */
int orc_read_pgen_header(const char *pgen_path, uint32_t *num_variants, uint32_t *num_samples) {
    FILE *f = fopen(pgen_path, "rb");
    if (!f) return ORC_E_IO;
    unsigned char h[12];
    size_t got = fread(h, 1, 12, f);
    fclose(f);
    if (got < 2) return ORC_E_IO;
    if (h[0] != 0x6C || h[1] != 0x1B) return ORC_E_MAGIC;
    if (got < 3) return ORC_E_IO;
    if (h[2] != 0x02) return ORC_E_MODE;
    if (got < 12) return ORC_E_IO;
    *num_variants = (uint32_t)h[3] | ((uint32_t)h[4] << 8) | ((uint32_t)h[5] << 16) | ((uint32_t)h[6] << 24);
    *num_samples = (uint32_t)h[7] | ((uint32_t)h[8] << 8) | ((uint32_t)h[9] << 16) | ((uint32_t)h[10] << 24);
    if (h[11] != 0x40) return ORC_E_FLAGS;
    return ORC_OK;
}

/* variant_record_size, pfile.rs:196-200 (u32 arithmetic, as written). */
uint32_t orc_record_size(uint32_t num_samples) {
    uint32_t bit_size = num_samples * 2u;
    return (bit_size / 8u) + ((bit_size % 8u == 0u) ? 0u : 1u);
}

/* record offset, pfile.rs:165: 12 + (var_idx as u32 * record_size) as u64.
 * faithful_u32 != 0 reproduces the release-build wrap; 0 is the u64 fix that
 * the GPU path (and parity for files > 4 GiB) is defined against. */
uint64_t orc_record_offset(uint64_t var_idx, uint32_t record_size, int faithful_u32) {
    if (faithful_u32) return 12u + (uint64_t)((uint32_t)var_idx * record_size);
    return 12u + var_idx * (uint64_t)record_size;
}

/* 2-bit extract, pfile.rs:172-175.  (static inline twin for the hot loop: an exported
 * symbol of a -fPIC library is interposable and would not be inlined, which made the
 * timed baseline several times slower than the release-mode Rust it stands in for.) */
static inline unsigned decode_(const unsigned char *record, uint64_t sam_idx) {
    uint64_t sample_offset = sam_idx / 4;
    unsigned host_byte = record[sample_offset];
    unsigned in_byte_offset = (unsigned)(sam_idx % 4);
    return (host_byte >> (in_byte_offset * 2)) & 0x3u;
}
unsigned orc_decode(const unsigned char *record, uint64_t sam_idx) { return decode_(record, sam_idx); }

/* genotype text, pfile.rs:177-183. */
static inline const char *gt_text_(unsigned code) {
    switch (code) {
    case 0: return "0/0";
    case 1: return "0/1";
    case 2: return "1/1";
    default: return "./.";
    }
}
const char *orc_gt_text(unsigned code) { return gt_text_(code); }

/* --------------------------------------------------------- metadata files -- */

typedef struct {
    char *data;
    size_t len;
} orc_buf;

static int slurp(const char *path, orc_buf *b) {
    FILE *f = fopen(path, "rb");
    if (!f) return ORC_E_IO;
    fseeko(f, 0, SEEK_END);
    off_t n = ftello(f);
    fseeko(f, 0, SEEK_SET);
    b->data = (char *)malloc((size_t)n + 1);
    if (!b->data) { fclose(f); return ORC_E_NOMEM; }
    b->len = fread(b->data, 1, (size_t)n, f);
    b->data[b->len] = 0;
    fclose(f);
    return ORC_OK;
}

/* BufRead::read_line: bytes up to and including the next '\n' (or EOF). */
static size_t line_end(const orc_buf *b, size_t pos) {
    while (pos < b->len) {
        if (b->data[pos++] == '\n') break;
    }
    return pos;
}

/* read_pvar_header, pfile.rs:202-220.  Leading lines that start with '#':
 * all but the last are returned verbatim (with their line endings) as
 * [0, *comments_len); the last one is [*col_begin, *col_end).  */
static int leading_hash_lines(const orc_buf *b, size_t *comments_len, size_t *col_begin, size_t *col_end,
                              size_t *first_data_begin, size_t *first_data_end) {
    size_t pos = 0, prev_begin = 0, prev_end = 0;
    int have = 0;
    for (;;) {
        size_t e = line_end(b, pos);
        /* an empty read (EOF) does not start with '#': loop ends, pfile.rs:210-214 */
        if (e > pos && b->data[pos] == '#') {
            prev_begin = pos;
            prev_end = e;
            have = 1;
            pos = e;
        } else {
            *first_data_begin = pos;
            *first_data_end = e;
            break;
        }
    }
    if (!have) return ORC_E_NO_HEADER;
    *comments_len = prev_begin;
    *col_begin = prev_begin;
    *col_end = prev_end;
    return ORC_OK;
}

/* find_metadata_file_header_start, pfile.rs:248-268.
 * current_pos is the stream position after reading the first non-'#' line;
 * offset = len(that line) + len(previous line) - 1.  With at least one '#'
 * line this lands just after the '#' of the column line.  With none, prev_buf
 * is empty and the result is byte 1 (the reference's own quirk, kept). */
static size_t metadata_header_start(const orc_buf *b) {
    size_t pos = 0, prev_len = 0, cur_len = 0;
    for (;;) {
        size_t e = line_end(b, pos);
        prev_len = cur_len;
        cur_len = e - pos;
        int is_hash = (cur_len > 0 && b->data[pos] == '#');
        pos = e;
        if (!is_hash) {
            size_t offset = cur_len + prev_len - 1; /* may wrap when both are 0, as u64 would panic */
            return pos - offset;
        }
    }
}

/* A parsed tab-delimited table: what csv::ReaderBuilder::new().delimiter(b'\t')
 * .has_headers(true) yields (pfile.rs:270-283) for inputs that need no quoting:
 * record terminators \n, \r\n, \r; empty lines skipped; every record must have
 * as many fields as the first one (flexible(false)); a '"' anywhere in the
 * table region is rejected (ORC_E_QUOTE) rather than guessed. */
typedef struct {
    orc_buf file;
    size_t n_cols;
    size_t n_rows;  /* data records, header excluded */
    /* field k of record r (r = 0 is the header record): */
    size_t *f_begin; /* (n_rows+1) * n_cols */
    size_t *f_len;
} orc_table;

static void table_free(orc_table *t) {
    free(t->file.data);
    free(t->f_begin);
    free(t->f_len);
    memset(t, 0, sizeof *t);
}

static int table_parse(const char *path, orc_table *t) {
    memset(t, 0, sizeof *t);
    int rc = slurp(path, &t->file);
    if (rc) return rc;
    const orc_buf *b = &t->file;
    size_t pos = metadata_header_start(b);
    if (pos > b->len) pos = b->len;
    size_t cap = 1024, nf = 0;
    size_t *fb = (size_t *)malloc(cap * sizeof(size_t)), *fl = (size_t *)malloc(cap * sizeof(size_t));
    if (!fb || !fl) return ORC_E_NOMEM;
    size_t n_rec = 0, n_cols = 0;
    while (pos < b->len) {
        /* skip empty lines (csv-core skips records that are only a terminator) */
        if (b->data[pos] == '\n' || b->data[pos] == '\r') { pos++; continue; }
        size_t rec_fields = 0;
        size_t fstart = pos;
        for (;;) {
            char c = (pos < b->len) ? b->data[pos] : '\n';
            if (c == '"') { free(fb); free(fl); return ORC_E_QUOTE; }
            if (c == '\t' || c == '\n' || c == '\r' || pos >= b->len) {
                if (nf == cap) {
                    cap *= 2;
                    fb = (size_t *)realloc(fb, cap * sizeof(size_t));
                    fl = (size_t *)realloc(fl, cap * sizeof(size_t));
                    if (!fb || !fl) return ORC_E_NOMEM;
                }
                fb[nf] = fstart;
                fl[nf] = pos - fstart;
                nf++;
                rec_fields++;
                if (c == '\t') { pos++; fstart = pos; continue; }
                /* terminator: \r\n counts as one */
                if (pos < b->len) {
                    if (c == '\r' && pos + 1 < b->len && b->data[pos + 1] == '\n') pos += 2;
                    else pos++;
                }
                break;
            }
            pos++;
        }
        if (n_rec == 0) n_cols = rec_fields;
        else if (rec_fields != n_cols) { free(fb); free(fl); return ORC_E_RAGGED; }
        n_rec++;
    }
    t->n_cols = n_cols;
    t->n_rows = n_rec ? n_rec - 1 : 0;
    t->f_begin = fb;
    t->f_len = fl;
    return ORC_OK;
}

static const char *tfield(const orc_table *t, size_t rec, size_t col, size_t *len) {
    size_t k = rec * t->n_cols + col;
    *len = t->f_len[k];
    return t->file.data + t->f_begin[k];
}

/* ------------------------------------------------------------ BufWriter -- */

/* std::io::BufWriter with the default 8 KiB capacity (pfile.rs:137).
 * write(): flush first if the piece does not fit; pieces >= capacity go
 * straight to the fd; otherwise append. */
#define BW_CAP 8192
typedef struct {
    int fd;
    size_t n;
    int err;
    unsigned char buf[BW_CAP];
} bufwriter;

static void bw_flush(bufwriter *w) {
    size_t off = 0;
    while (off < w->n) {
        ssize_t k = write(w->fd, w->buf + off, w->n - off);
        if (k < 0) {
            if (errno == EINTR) continue;
            w->err = 1;
            break;
        }
        off += (size_t)k;
    }
    w->n = 0;
}

/* BufWriter::write: the common case (the bytes fit in the spare capacity) is a plain copy,
 * inlined so that the constant-length appends of pfile.rs:186-187 become direct stores, as
 * they do in release-mode Rust; everything else goes through the out-of-line slow path. */
static __attribute__((noinline)) void bw_write_slow(bufwriter *w, const void *p, size_t len) {
    if (w->n + len > BW_CAP) bw_flush(w);
    if (len >= BW_CAP) {
        const unsigned char *q = (const unsigned char *)p;
        while (len) {
            ssize_t k = write(w->fd, q, len);
            if (k < 0) {
                if (errno == EINTR) continue;
                w->err = 1;
                return;
            }
            q += k;
            len -= (size_t)k;
        }
        return;
    }
    memcpy(w->buf + w->n, p, len);
    w->n += len;
}

static inline __attribute__((always_inline)) void bw_write(bufwriter *w, const void *p, size_t len) {
    if (__builtin_expect(w->n + len < BW_CAP, 1)) {
        memcpy(w->buf + w->n, p, len);
        w->n += len;
    } else {
        bw_write_slow(w, p, len);
    }
}

/* ----------------------------------------------------------- output_vcf -- */

static int is_rust_ascii_ws(unsigned char c) {
    /* str::trim() strips char::is_whitespace; the ASCII members are these six.
     * Non-ASCII White_Space code points are not handled (none occur in .pvar
     * column lines). */
    return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r';
}

/* Pfile::output_vcf, pfile.rs:104-194, with the two filter_metadata results
 * (pfile.rs:127-128, 312-335) supplied as ascending index lists.
 * n_var < 0 / n_sam < 0 mean "no query" = keep every row in file order. */
int orc_output_vcf(const char *prefix, const int64_t *var_idx, int64_t n_var, const int64_t *sam_idx, int64_t n_sam,
                   const char *out_path, int flags) {
    char path[4096];
    uint32_t M = 0, N = 0;
    snprintf(path, sizeof path, "%s.pgen", prefix);
    int rc = orc_read_pgen_header(path, &M, &N);
    if (rc) return rc;
    uint32_t R = orc_record_size(N);

    /* read_pvar_header, pfile.rs:110 */
    orc_table pvar, psam;
    snprintf(path, sizeof path, "%s.pvar", prefix);
    rc = table_parse(path, &pvar);
    if (rc) return rc;
    size_t comments_len, col_b, col_e, d_b, d_e;
    rc = leading_hash_lines(&pvar.file, &comments_len, &col_b, &col_e, &d_b, &d_e);
    if (rc) { table_free(&pvar); return rc; }

    /* psam_reader + IID lookup, pfile.rs:111-126 */
    snprintf(path, sizeof path, "%s.psam", prefix);
    rc = table_parse(path, &psam);
    if (rc) { table_free(&pvar); return rc; }
    size_t iid_col = (size_t)-1;
    for (size_t c = 0; c < psam.n_cols; c++) {
        size_t l;
        const char *s = tfield(&psam, 0, c, &l);
        if (l == 3 && memcmp(s, "IID", 3) == 0) { iid_col = c; break; }
    }
    if (iid_col == (size_t)-1) { table_free(&pvar); table_free(&psam); return ORC_E_NO_IID; }

    int64_t nv = n_var < 0 ? (int64_t)pvar.n_rows : n_var;
    int64_t ns = n_sam < 0 ? (int64_t)psam.n_rows : n_sam;
    for (int64_t i = 0; i < nv; i++) {
        int64_t v = n_var < 0 ? i : var_idx[i];
        if (v < 0 || (size_t)v >= pvar.n_rows) { table_free(&pvar); table_free(&psam); return ORC_E_RANGE; }
    }
    for (int64_t i = 0; i < ns; i++) {
        int64_t s = n_sam < 0 ? i : sam_idx[i];
        if (s < 0 || (size_t)s >= psam.n_rows) { table_free(&pvar); table_free(&psam); return ORC_E_RANGE; }
        /* record_buf[sample_offset] out of range panics, pfile.rs:173 */
        if ((uint64_t)s / 4 >= R) { table_free(&pvar); table_free(&psam); return ORC_E_RANGE; }
    }

    int ofd = open(out_path, O_WRONLY | O_CREAT | O_TRUNC, 0644); /* File::create, pfile.rs:136 */
    if (ofd < 0) { table_free(&pvar); table_free(&psam); return ORC_E_IO; }
    bufwriter *w = (bufwriter *)malloc(sizeof *w);
    w->fd = ofd; w->n = 0; w->err = 0;

    /* header, pfile.rs:139-146 */
    bw_write(w, "##fileformat=VCFv4.2\n", 21);
    bw_write(w, "##source=pgen-rs\n", 17);
    bw_write(w, pvar.file.data, comments_len);
    {
        size_t b = col_b, e = col_e;
        while (b < e && is_rust_ascii_ws((unsigned char)pvar.file.data[b])) b++;
        while (e > b && is_rust_ascii_ws((unsigned char)pvar.file.data[e - 1])) e--;
        bw_write(w, pvar.file.data + b, e - b);
    }
    bw_write(w, "\tFORMAT\t", 8);
    for (int64_t i = 0; i < ns; i++) { /* sam_ids joined with '\t', pfile.rs:130-134 */
        int64_t s = n_sam < 0 ? i : sam_idx[i];
        size_t l;
        const char *id = tfield(&psam, (size_t)s + 1, iid_col, &l);
        if (i) bw_write(w, "\t", 1);
        bw_write(w, id, l);
    }
    bw_write(w, "\n", 1);

    /* body, pfile.rs:149-192 */
    snprintf(path, sizeof path, "%s.pgen", prefix);
    int pfd = open(path, O_RDONLY);
    if (pfd < 0) { free(w); close(ofd); table_free(&pvar); table_free(&psam); return ORC_E_IO; }
    int faithful = (flags & ORC_FLAG_FAITHFUL_U32) != 0;
    int bulk = (flags & ORC_FLAG_IO_BULK) != 0;
    unsigned char *linebuf = NULL;
    unsigned char *recbuf_bulk = NULL;
    if (bulk) {
        linebuf = (unsigned char *)malloc((size_t)ns * 4 + 1);
        recbuf_bulk = (unsigned char *)malloc(R ? R : 1);
    }
    rc = ORC_OK;
    for (int64_t i = 0; i < nv && rc == ORC_OK; i++) {
        int64_t v = n_var < 0 ? i : var_idx[i];
        for (size_t c = 0; c < pvar.n_cols; c++) { /* pfile.rs:157-160 */
            size_t l;
            const char *s = tfield(&pvar, (size_t)v + 1, c, &l);
            bw_write(w, s, l);
            bw_write(w, "\t", 1);
        }
        bw_write(w, "GT", 2); /* pfile.rs:161 */
        uint64_t off = orc_record_offset((uint64_t)v, R, faithful);
        unsigned char *rec = bulk ? recbuf_bulk : (unsigned char *)calloc(R ? R : 1, 1); /* vec![0u8; R] */
        if (!rec) { rc = ORC_E_NOMEM; break; }
        if (!bulk) {
            if (lseek(pfd, (off_t)off, SEEK_SET) < 0) rc = ORC_E_IO; /* pfile.rs:169 */
        }
        size_t got = 0;
        while (rc == ORC_OK && got < R) { /* read_exact, pfile.rs:170 */
            ssize_t k = bulk ? pread(pfd, rec + got, R - got, (off_t)(off + got)) : read(pfd, rec + got, R - got);
            if (k < 0) { if (errno == EINTR) continue; rc = ORC_E_IO; break; }
            if (k == 0) { rc = ORC_E_RANGE; break; } /* UnexpectedEof → unwrap panic */
            got += (size_t)k;
        }
        if (rc == ORC_OK) {
            if (!bulk) {
                for (int64_t j = 0; j < ns; j++) { /* pfile.rs:171-188 */
                    int64_t s = n_sam < 0 ? j : sam_idx[j];
                    unsigned code = orc_decode(rec, (uint64_t)s);
                    const char *gt = orc_gt_text(code);
                    bw_write(w, "\t", 1);
                    bw_write(w, gt, 3);
                }
            } else {
                unsigned char *q = linebuf;
                for (int64_t j = 0; j < ns; j++) {
                    int64_t s = n_sam < 0 ? j : sam_idx[j];
                    const char *gt = gt_text_(decode_(rec, (uint64_t)s));
                    q[0] = '\t'; q[1] = (unsigned char)gt[0]; q[2] = (unsigned char)gt[1]; q[3] = (unsigned char)gt[2];
                    q += 4;
                }
                bw_write(w, linebuf, (size_t)(q - linebuf));
            }
            bw_write(w, "\n", 1); /* pfile.rs:190 */
        }
        if (!bulk) free(rec);
    }
    bw_flush(w); /* BufWriter drop */
    if (w->err && rc == ORC_OK) rc = ORC_E_IO;
    free(w);
    free(linebuf);
    free(recbuf_bulk);
    close(pfd);
    close(ofd);
    table_free(&pvar);
    table_free(&psam);
    return rc;
}

/* Body-only variant used by bench.py / tests when the caller already holds the
 * selections and prefixes (exactly the inputs the C-ABI pgb_export_gt_vcf
 * takes): same loop as pfile.rs:156-192 with prefix bytes supplied as a blob
 * (the bytes pfile.rs:157-161 would write).  sam_idx == NULL ⇒ all samples,
 * var_idx == NULL ⇒ rows 0..n_var-1.  Appends to out_fd. */
int orc_export_body(const char *pgen_path, const uint32_t *var_idx, uint64_t n_var, const uint32_t *sam_idx,
                    uint64_t n_sam, const uint8_t *prefix_blob, const uint64_t *prefix_off, int out_fd, int flags) {
    uint32_t M = 0, N = 0;
    int rc = orc_read_pgen_header(pgen_path, &M, &N);
    if (rc) return rc;
    uint32_t R = orc_record_size(N);
    int pfd = open(pgen_path, O_RDONLY);
    if (pfd < 0) return ORC_E_IO;
    bufwriter *w = (bufwriter *)malloc(sizeof *w);
    w->fd = out_fd; w->n = 0; w->err = 0;
    int faithful = (flags & ORC_FLAG_FAITHFUL_U32) != 0;
    int bulk = (flags & ORC_FLAG_IO_BULK) != 0;
    uint64_t ns = sam_idx ? n_sam : N;
    unsigned char *linebuf = bulk ? (unsigned char *)malloc((size_t)ns * 4 + 1) : NULL;
    unsigned char *recbuf_bulk = bulk ? (unsigned char *)malloc(R ? R : 1) : NULL;
    for (uint64_t i = 0; i < n_var && rc == ORC_OK; i++) {
        uint64_t v = var_idx ? var_idx[i] : i;
        bw_write(w, prefix_blob + prefix_off[i], (size_t)(prefix_off[i + 1] - prefix_off[i]));
        uint64_t off = orc_record_offset(v, R, faithful);
        unsigned char *rec = bulk ? recbuf_bulk : (unsigned char *)calloc(R ? R : 1, 1);
        if (!bulk && lseek(pfd, (off_t)off, SEEK_SET) < 0) rc = ORC_E_IO;
        size_t got = 0;
        while (rc == ORC_OK && got < R) {
            ssize_t k = bulk ? pread(pfd, rec + got, R - got, (off_t)(off + got)) : read(pfd, rec + got, R - got);
            if (k < 0) { if (errno == EINTR) continue; rc = ORC_E_IO; break; }
            if (k == 0) { rc = ORC_E_RANGE; break; }
            got += (size_t)k;
        }
        if (rc == ORC_OK) {
            if (!bulk) {
                for (uint64_t j = 0; j < ns; j++) {
                    uint64_t s = sam_idx ? sam_idx[j] : j;
                    if (s / 4 >= R) { rc = ORC_E_RANGE; break; }
                    const char *gt = gt_text_(decode_(rec, s));
                    bw_write(w, "\t", 1);
                    bw_write(w, gt, 3);
                }
            } else {
                unsigned char *q = linebuf;
                for (uint64_t j = 0; j < ns; j++) {
                    uint64_t s = sam_idx ? sam_idx[j] : j;
                    if (s / 4 >= R) { rc = ORC_E_RANGE; break; }
                    const char *gt = gt_text_(decode_(rec, s));
                    q[0] = '\t'; q[1] = (unsigned char)gt[0]; q[2] = (unsigned char)gt[1]; q[3] = (unsigned char)gt[2];
                    q += 4;
                }
                bw_write(w, linebuf, (size_t)(q - linebuf));
            }
            bw_write(w, "\n", 1);
        }
        if (!bulk) free(rec);
    }
    bw_flush(w);
    if (w->err && rc == ORC_OK) rc = ORC_E_IO;
    free(w);
    free(linebuf);
    free(recbuf_bulk);
    close(pfd);
    return rc;
}

/* In-memory single-line helper for known-answer tests: formats the GT fields of
 * one record for the given sample list (pfile.rs:171-188) into out (4 bytes per
 * sample).  Returns bytes written. */
uint64_t orc_format_gt_fields(const unsigned char *record, const uint32_t *sam_idx, uint64_t n_sam, uint32_t num_samples,
                              unsigned char *out) {
    uint64_t ns = sam_idx ? n_sam : num_samples;
    for (uint64_t j = 0; j < ns; j++) {
        uint64_t s = sam_idx ? sam_idx[j] : j;
        const char *gt = gt_text_(decode_(record, s));
        out[4 * j] = '\t';
        memcpy(out + 4 * j + 1, gt, 3);
    }
    return ns * 4;
}

const char *orc_strerror(int rc) {
    switch (rc) {
    case ORC_OK: return "ok";
    case ORC_E_IO: return "I/O error (reference: unwrap panic)";
    case ORC_E_MAGIC: return "bad magic (pfile.rs:47)";
    case ORC_E_MODE: return "storage mode != 0x02 (pfile.rs:53)";
    case ORC_E_FLAGS: return "header byte 11 != 0x40 (pfile.rs:69)";
    case ORC_E_NO_HEADER: return "no leading '#' line (pfile.rs:217)";
    case ORC_E_NO_IID: return "IID not among the headers (pfile.rs:125)";
    case ORC_E_RAGGED: return "ragged row (csv UnequalLengths)";
    case ORC_E_QUOTE: return "input needs csv quoting rules; rejected";
    case ORC_E_RANGE: return "index out of range (pfile.rs:170,173)";
    case ORC_E_NOMEM: return "out of memory";
    default: return "unknown";
    }
}

#ifdef ORC_MAIN
/* CLI: pgen_oracle <prefix> <out.vcf> [--var-idx FILE] [--sam-idx FILE] [--faithful-u32] [--bulk]
 * index files: one decimal index per line, ascending. */
static int64_t *read_idx(const char *path, int64_t *n) {
    FILE *f = fopen(path, "r");
    if (!f) { *n = -2; return NULL; }
    size_t cap = 1024, k = 0;
    int64_t *a = (int64_t *)malloc(cap * sizeof *a);
    long long x;
    while (fscanf(f, "%lld", &x) == 1) {
        if (k == cap) { cap *= 2; a = (int64_t *)realloc(a, cap * sizeof *a); }
        a[k++] = x;
    }
    fclose(f);
    *n = (int64_t)k;
    return a;
}
int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s <prefix> <out.vcf> [--var-idx F] [--sam-idx F] [--faithful-u32] [--bulk]\n", argv[0]);
        return 2;
    }
    int64_t *vi = NULL, *si = NULL, nv = -1, ns = -1;
    int flags = 0;
    for (int i = 3; i < argc; i++) {
        if (!strcmp(argv[i], "--var-idx") && i + 1 < argc) vi = read_idx(argv[++i], &nv);
        else if (!strcmp(argv[i], "--sam-idx") && i + 1 < argc) si = read_idx(argv[++i], &ns);
        else if (!strcmp(argv[i], "--faithful-u32")) flags |= ORC_FLAG_FAITHFUL_U32;
        else if (!strcmp(argv[i], "--bulk")) flags |= ORC_FLAG_IO_BULK;
    }
    if (nv == -2 || ns == -2) { fprintf(stderr, "cannot read index file\n"); return 2; }
    int rc = orc_output_vcf(argv[1], vi, nv, si, ns, argv[2], flags);
    if (rc) { fprintf(stderr, "pgen_oracle: %s\n", orc_strerror(rc)); return 101; }
    return 0;
}
#endif
