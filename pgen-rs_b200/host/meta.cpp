// meta.cpp — see meta.hpp.
#include "meta.hpp"

#include <errno.h>
#include <stdio.h>
#include <string.h>

#include "pgb200.h"

namespace pgb {

MetaTable::MetaTable(const std::string &path) : path_(path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) throw MetaError{PGB_E_IO, "open " + path + ": " + strerror(errno)};
    char buf[1 << 16];
    size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) data_.append(buf, k);
    bool bad = ferror(f);
    fclose(f);
    if (bad) throw MetaError{PGB_E_IO, "read " + path};
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

    std::vector<Span> rec;
    size_t n_rec = 0, n_cols = 0;
    const size_t n = data_.size();
    pos = start;
    while (pos < n) {
        char c = data_[pos];
        if (c == '\n' || c == '\r') { pos++; continue; } // empty line
        rec.clear();
        size_t fstart = pos;
        for (;;) {
            bool at_end = pos >= n;
            c = at_end ? '\n' : data_[pos];
            if (c == '"') throw MetaError{PGB_E_CSV, path_ + ": quoted fields are not supported"};
            if (c == '\t' || c == '\n' || c == '\r') {
                if (pos - fstart > 0xffffffffull) throw MetaError{PGB_E_CSV, "field too long"};
                rec.push_back(Span{(uint64_t)fstart, (uint32_t)(pos - fstart)});
                if (c == '\t') { pos++; fstart = pos; continue; }
                if (!at_end) pos += (c == '\r' && pos + 1 < n && data_[pos + 1] == '\n') ? 2 : 1;
                break;
            }
            pos++;
        }
        if (n_rec == 0) {
            n_cols = rec.size();
            for (const Span &s : rec) headers_.emplace_back(data_.data() + s.off, s.len);
        } else {
            if (rec.size() != n_cols)
                throw MetaError{PGB_E_CSV, path_ + ": record " + std::to_string(n_rec) + " has " + std::to_string(rec.size()) +
                                               " fields, expected " + std::to_string(n_cols)};
            fields_.insert(fields_.end(), rec.begin(), rec.end());
        }
        n_rec++;
    }
    n_rows_ = n_rec ? n_rec - 1 : 0;
    parsed_ = true;
}

void MetaTable::row(size_t r, std::vector<std::string_view> *out) const {
    out->resize(headers_.size());
    for (size_t c = 0; c < headers_.size(); c++) (*out)[c] = field(r, c);
}

} // namespace pgb
