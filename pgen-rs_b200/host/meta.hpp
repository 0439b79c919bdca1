// meta.hpp — .pvar / .psam readers: the CPU-side producers of the hot path's inputs.
//
// Mirrors /root/reference/src/pfile.rs:
//   read_pvar_header                 :202-220
//   find_metadata_file_header_start  :248-268
//   metadata_file_reader             :270-283  (csv 1.3.0: tab delimiter, has_headers(true),
//                                               flexible(false), terminators \n \r\n \r, empty
//                                               lines skipped)
// csv quoting is not reproduced: a '"' in the table region is rejected (PGB_E_CSV) rather
// than guessed at (the shipped fixtures and all synthetic inputs contain none).
#pragma once
#include <stdint.h>

#include <string>
#include <string_view>
#include <vector>

namespace pgb {

struct MetaError {
    int status; // pgb_status
    std::string msg;
};

class MetaTable {
  public:
    // Reads the whole file.  Throws MetaError.
    explicit MetaTable(const std::string &path);

    const std::string &path() const { return path_; }
    const std::string &data() const { return data_; }
    // read_pvar_header: leading '#' lines except the last, verbatim; and the last one.
    // Throws MetaError(PGB_E_NO_HEADER) when the file has no leading '#' line.
    void leading_header(std::string_view *comments, std::string_view *column_line) const;

    // Parses the table (lazily called by the accessors below).  Throws MetaError.
    void parse();
    const std::vector<std::string> &headers() const { return headers_; }
    size_t n_cols() const { return headers_.size(); }
    size_t n_rows() const { return rows_.size(); }
    std::string_view field(size_t row, size_t col) const {
        const RowSpan &r = rows_[row];
        const uint32_t *fs = &fstart_[row * headers_.size()];
        const uint32_t b = fs[col];
        const uint32_t e = col + 1 < headers_.size() ? fs[col + 1] - 1 : r.len;
        return std::string_view(data_.data() + r.off + b, e - b);
    }
    // The record's text without its terminator: the fields joined by '\t', exactly as in the file.
    std::string_view row_text(size_t row) const { return std::string_view(data_.data() + rows_[row].off, rows_[row].len); }
    void row(size_t r, std::vector<std::string_view> *out) const;
    // Position of a record in data(): what the device needs to build the line prefix from the raw image.
    uint64_t row_offset(size_t row) const { return rows_[row].off; }
    uint32_t row_length(size_t row) const { return rows_[row].len; }
    // Hands the file image over (the table's accessors are dead afterwards).
    std::string take_data() { return std::move(data_); }

    // Worker threads used for parsing / filtering large tables (1 for small inputs).
    static unsigned worker_threads(size_t bytes);

  private:
    struct RowSpan {
        uint64_t off; // first byte of the record in data_
        uint32_t len; // bytes up to (not including) the terminator
    };
    struct Block { // parse result of one line-aligned slice of the file
        std::vector<RowSpan> rows;
        std::vector<uint32_t> fstart;
        std::string error;
    };
    void parse_block(size_t begin, size_t end, size_t n_cols, size_t first_record_no, Block *out) const;

    std::string path_;
    std::string data_;
    std::vector<std::string> headers_;
    std::vector<RowSpan> rows_;     // data records (header excluded)
    std::vector<uint32_t> fstart_;  // n_rows x n_cols field starts relative to the record
    bool parsed_ = false;
};

} // namespace pgb
