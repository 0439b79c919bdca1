// meta.cpp — see meta.hpp.
#include "meta.hpp"

#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <thread>

#include "pgb200.h"

namespace pgb {

MetaTable::MetaTable(const std::string &path) : path_(path) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw MetaError{PGB_E_IO, "open " + path + ": " + strerror(errno)};
    struct stat st;
    size_t have = 0;
    if (fstat(fd, &st) == 0 && st.st_size > 0) data_.resize((size_t)st.st_size);
    for (;;) {
        if (have == data_.size()) data_.resize(data_.size() + (1u << 20));
        ssize_t k = ::read(fd, &data_[have], data_.size() - have);
        if (k < 0) {
            if (errno == EINTR) continue;
            ::close(fd);
            throw MetaError{PGB_E_IO, "read " + path + ": " + strerror(errno)};
        }
        if (k == 0) break;
        have += (size_t)k;
    }
    ::close(fd);
    data_.resize(have);
}

unsigned MetaTable::worker_threads(size_t bytes) {
    size_t min_bytes = 8u << 20; // below this the threads cost more than they save
    if (const char *e = getenv("PGB_HOST_PAR_MIN_BYTES")) min_bytes = (size_t)strtoull(e, nullptr, 0);
    if (bytes < min_bytes) return 1;
    unsigned hw = std::thread::hardware_concurrency();
    if (const char *e = getenv("PGB_HOST_THREADS")) hw = (unsigned)atoi(e);
    return std::max(1u, std::min(hw, 32u));
}

namespace {
// BufRead::read_line span: through the next '\n' or to EOF.
size_t next_line(const std::string &d, size_t pos) {
    size_t j = d.find('\n', pos);
    return j == std::string::npos ? d.size() : j + 1;
}
} // namespace

void MetaTable::leading_header(std::string_view *comments, std::string_view *column_line) const {
    size_t pos = 0, last_b = 0, last_e = 0;
    bool have = false;
    for (;;) {
        size_t e = next_line(data_, pos);
        if (e > pos && data_[pos] == '#') {
            last_b = pos;
            last_e = e;
            have = true;
            pos = e;
        } else break;
    }
    if (!have) throw MetaError{PGB_E_NO_HEADER, path_ + ": no leading '#' line"};
    *comments = std::string_view(data_.data(), last_b);
    *column_line = std::string_view(data_.data() + last_b, last_e - last_b);
}

void MetaTable::parse() {
    if (parsed_) return;
    // find_metadata_file_header_start (pfile.rs:248-268): stream position after the first
    // non-'#' line, minus (len(that line) + len(previous line) - 1).
    size_t pos = 0, prev_len = 0, cur_len = 0;
    for (;;) {
        size_t e = next_line(data_, pos);
        prev_len = cur_len;
        cur_len = e - pos;
        bool hash = cur_len > 0 && data_[pos] == '#';
        pos = e;
        if (!hash) break;
    }
    if (cur_len + prev_len == 0) throw MetaError{PGB_E_CSV, path_ + ": empty metadata file"};
    size_t start = pos - (cur_len + prev_len - 1);
    if (start > data_.size()) start = data_.size();

    // header record (serial), then the data records in line-aligned blocks, one per worker
    size_t hdr_end = start;
    {
        // the header is the first non-empty record
        size_t p = start;
        while (p < data_.size() && (data_[p] == '\n' || data_[p] == '\r')) p++;
        size_t e = p;
        while (e < data_.size() && data_[e] != '\n' && data_[e] != '\r') e++;
        if (p < data_.size()) {
            size_t fs = p;
            for (size_t i = p; i <= e; i++) {
                if (i == e || data_[i] == '\t') {
                    headers_.emplace_back(data_.data() + fs, i - fs);
                    fs = i + 1;
                }
            }
            for (const std::string &h : headers_)
                if (h.find('"') != std::string::npos) throw MetaError{PGB_E_CSV, path_ + ": quoted fields are not supported"};
        }
        hdr_end = e;
    }
    const size_t n = data_.size(), n_cols = headers_.size();
    if (n_cols == 0) { parsed_ = true; return; }
    const unsigned T = worker_threads(n - hdr_end);
    std::vector<size_t> cut(T + 1);
    cut[0] = hdr_end;
    cut[T] = n;
    for (unsigned t = 1; t < T; t++) {
        size_t p = hdr_end + (n - hdr_end) / T * t;
        if (p < cut[t - 1]) p = cut[t - 1];
        const void *q = p < n ? memchr(data_.data() + p, '\n', n - p) : nullptr;
        cut[t] = q ? (size_t)((const char *)q - data_.data()) + 1 : n;
    }
    std::vector<Block> blk(T);
    if (T == 1) {
        parse_block(cut[0], cut[1], n_cols, 1, &blk[0]);
    } else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; t++) th.emplace_back([&, t] { parse_block(cut[t], cut[t + 1], n_cols, 0, &blk[t]); });
        for (auto &x : th) x.join();
    }
    size_t total = 0;
    for (unsigned t = 0; t < T; t++) {
        if (!blk[t].error.empty()) {
            // record numbers are only exact for a serial parse: redo serially for the reference's message
            if (T > 1) {
                Block all;
                parse_block(cut[0], cut[T], n_cols, 1, &all);
                throw MetaError{PGB_E_CSV, all.error.empty() ? blk[t].error : all.error};
            }
            throw MetaError{PGB_E_CSV, blk[t].error};
        }
        total += blk[t].rows.size();
    }
    rows_.resize(total);
    fstart_.resize(total * n_cols);
    {
        std::vector<std::thread> th;
        size_t at = 0;
        for (unsigned t = 0; t < T; t++) {
            const size_t dst = at;
            at += blk[t].rows.size();
            auto copy = [this, &blk, t, dst, n_cols] {
                if (blk[t].rows.empty()) return;
                memcpy(&rows_[dst], blk[t].rows.data(), blk[t].rows.size() * sizeof(RowSpan));
                memcpy(&fstart_[dst * n_cols], blk[t].fstart.data(), blk[t].fstart.size() * sizeof(uint32_t));
            };
            if (T == 1) copy(); else th.emplace_back(copy);
        }
        for (auto &x : th) x.join();
    }
    parsed_ = true;
}

// Splits [begin, end) into records (terminators \n, \r\n, \r; empty lines skipped) and fields
// (tab).  A record whose field count differs from n_cols is csv's UnequalLengths error.
void MetaTable::parse_block(size_t begin, size_t end, size_t n_cols, size_t first_record_no, Block *out) const {
    const char *d = data_.data();
    size_t pos = begin, rec_no = first_record_no;
    out->rows.reserve((end - begin) / 64 + 16);
    out->fstart.reserve(((end - begin) / 64 + 16) * n_cols);
    while (pos < end) {
        char c = d[pos];
        if (c == '\n' || c == '\r') { pos++; continue; } // empty line
        const size_t rs = pos;
        size_t nf = 0;
        out->fstart.push_back(0);
        nf = 1;
        for (; pos < end; pos++) {
            c = d[pos];
            if (c == '\t') {
                out->fstart.push_back((uint32_t)(pos + 1 - rs));
                nf++;
            } else if (c == '\n' || c == '\r') {
                break;
            } else if (c == '"') {
                out->error = path_ + ": quoted fields are not supported";
                return;
            }
        }
        if (pos - rs > 0xfffffff0ull) { out->error = path_ + ": record too long"; return; }
        out->rows.push_back(RowSpan{(uint64_t)rs, (uint32_t)(pos - rs)});
        if (nf != n_cols) {
            out->error = path_ + ": record " + std::to_string(rec_no) + " has " + std::to_string(nf) + " fields, expected " +
                         std::to_string(n_cols);
            return;
        }
        rec_no++;
        if (pos < end) pos += (d[pos] == '\r' && pos + 1 < data_.size() && d[pos + 1] == '\n') ? 2 : 1;
    }
}

void MetaTable::row(size_t r, std::vector<std::string_view> *out) const {
    out->resize(headers_.size());
    for (size_t c = 0; c < headers_.size(); c++) (*out)[c] = field(r, c);
}

} // namespace pgb
