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
    size_t n_rows() const { return n_rows_; }
    std::string_view field(size_t row, size_t col) const {
        const Span &s = fields_[row * headers_.size() + col];
        return std::string_view(data_.data() + s.off, s.len);
    }
    void row(size_t r, std::vector<std::string_view> *out) const;

  private:
    struct Span {
        uint64_t off;
        uint32_t len;
    };
    std::string path_;
    std::string data_;
    std::vector<std::string> headers_;
    std::vector<Span> fields_;
    size_t n_rows_ = 0;
    bool parsed_ = false;
};

} // namespace pgb
