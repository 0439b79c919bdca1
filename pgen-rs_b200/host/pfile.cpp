// pfile.cpp — see pfile.hpp.
#include "pfile.hpp"

#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <memory>
#include <thread>

#include "expr.hpp"

namespace pgb {

namespace {

void write_all(int fd, const char *p, size_t n) {
    while (n) {
        ssize_t k = ::write(fd, p, n);
        if (k < 0) {
            if (errno == EINTR) continue;
            throw PfileError{PGB_E_IO, std::string("write: ") + strerror(errno)};
        }
        p += k;
        n -= (size_t)k;
    }
}

// str::trim(): the ASCII members of char::is_whitespace.
bool rust_ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }

} // namespace

Pfile Pfile::from_prefix(const std::string &pfile_prefix) {
    Pfile p;
    p.pfile_prefix = pfile_prefix;
    pgb_file *f = nullptr;
    int rc = pgb_open(p.pgen_path().c_str(), &f); // same checks, same order as pfile.rs:44-69
    if (rc != PGB_OK) throw PfileError{rc, std::string(pgb_strerror(rc)) + ": " + p.pgen_path() + " " + pgb_last_error()};
    pgb_dims(f, &p.num_variants, &p.num_samples, nullptr);
    pgb_close(f);
    return p;
}

MetaTable Pfile::pvar_reader() const {
    try {
        return MetaTable(pvar_path());
    } catch (const MetaError &e) {
        throw PfileError{e.status, e.msg};
    }
}

MetaTable Pfile::psam_reader() const {
    try {
        return MetaTable(psam_path());
    } catch (const MetaError &e) {
        throw PfileError{e.status, e.msg};
    }
}

std::vector<uint32_t> filter_metadata(MetaTable &table, const std::optional<std::string> &query) {
    std::vector<uint32_t> kept;
    try {
        table.parse();
        const size_t n = table.n_rows();
        if (n > 0xffffffffull) throw PfileError{PGB_E_ARG, "too many rows"};
        if (!query) { // map_or(true, ..), pfile.rs:321
            kept.resize(n);
            for (size_t i = 0; i < n; i++) kept[i] = (uint32_t)i;
            return kept;
        }
        // The reference re-parses the expression for every row (pfile.rs:328), so a bad
        // expression only surfaces when there is at least one row.
        if (n == 0) return kept;
        const unsigned T = std::min<size_t>(n, MetaTable::worker_threads(table.data().size()));
        std::vector<std::vector<uint32_t>> part(T);
        std::vector<std::string> err(T);
        std::vector<size_t> err_row(T, SIZE_MAX);
        auto work = [&](unsigned t) {
            const size_t a = n / T * t, b = t + 1 == T ? n : n / T * (t + 1);
            size_t i = a;
            try {
                Expr ex(*query, table.headers()); // one parsed expression per worker
                std::vector<std::string_view> row;
                for (; i < b; i++) {
                    table.row(i, &row);
                    if (ex.eval_boolean(row)) part[t].push_back((uint32_t)i);
                }
            } catch (const ExprError &e) {
                err[t] = e.msg;
                err_row[t] = i;
            }
        };
        if (T == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < T; t++) th.emplace_back(work, t);
            for (auto &x : th) x.join();
        }
        for (unsigned t = 0; t < T; t++) // the first failing row in file order is the one the reference panics on
            if (err_row[t] != SIZE_MAX) throw ExprError{err[t]};
        size_t total = 0;
        for (auto &v : part) total += v.size();
        kept.reserve(total);
        for (auto &v : part) kept.insert(kept.end(), v.begin(), v.end());
    } catch (const MetaError &e) {
        throw PfileError{e.status, e.msg};
    } catch (const ExprError &e) {
        throw PfileError{PGB_E_EXPR, e.msg};
    }
    return kept;
}

void Pfile::query_metadata(MetaTable &reader, const std::optional<std::string> &query, const std::string &f_string,
                           int out_fd) const {
    try {
        reader.parse();
        std::unique_ptr<Expr> q, fs;
        std::vector<std::string_view> row;
        std::string buf;
        for (size_t i = 0; i < reader.n_rows(); i++) {
            reader.row(i, &row);
            bool keep = true;
            if (query) {
                if (!q) q = std::make_unique<Expr>(*query, reader.headers());
                keep = q->eval_boolean(row);
            }
            if (keep) {
                if (!fs) fs = std::make_unique<Expr>(f_string, reader.headers());
                buf += fs->eval_string(row);
                buf.push_back('\n'); // println!, pfile.rs:98
                if (buf.size() > (1u << 16)) { write_all(out_fd, buf.data(), buf.size()); buf.clear(); }
            }
        }
        write_all(out_fd, buf.data(), buf.size());
    } catch (const MetaError &e) {
        throw PfileError{e.status, e.msg};
    } catch (const ExprError &e) {
        throw PfileError{PGB_E_EXPR, e.msg};
    }
}

VcfPlan Pfile::plan_vcf(const std::optional<std::string> &sam_query, const std::optional<std::string> &var_query) const {
    VcfPlan plan;
    try {
        // read_pvar_header, pfile.rs:110 — panics first when the .pvar has no '#' line
        MetaTable pvar = pvar_reader();
        std::string_view comments, column_line;
        pvar.leading_header(&comments, &column_line);
        // psam headers + IID column, pfile.rs:111-126
        MetaTable psam = psam_reader();
        psam.parse();
        int iid = -1;
        for (size_t c = 0; c < psam.n_cols(); c++)
            if (psam.headers()[c] == "IID") { iid = (int)c; break; }
        if (iid < 0) throw PfileError{PGB_E_NO_IID, "IID not among the headers of " + psam_path()};
        plan.var_idx = filter_metadata(pvar, var_query); // pfile.rs:127
        plan.sam_idx = filter_metadata(psam, sam_query); // pfile.rs:128

        // header, pfile.rs:139-146
        std::string &h = plan.header;
        h += "##fileformat=VCFv4.2\n##source=pgen-rs\n";
        h.append(comments.data(), comments.size());
        size_t b = 0, e = column_line.size();
        while (b < e && rust_ws(column_line[b])) b++;
        while (e > b && rust_ws(column_line[e - 1])) e--;
        h.append(column_line.data() + b, e - b);
        h += "\tFORMAT\t";
        for (size_t k = 0; k < plan.sam_idx.size(); k++) {
            if (k) h.push_back('\t');
            std::string_view id = psam.field(plan.sam_idx[k], (size_t)iid);
            h.append(id.data(), id.size());
        }
        h.push_back('\n');

        // line prefixes, pfile.rs:157-161: every field + '\t', then "GT".  Without csv quoting a record's fields
        // joined by '\t' are its text in the file, so the device copies the row from the raw image and appends
        // "\tGT"; the host only records where each kept row sits.
        const size_t nv = plan.var_idx.size();
        plan.row_off.resize(nv);
        plan.row_len.resize(nv);
        for (size_t k = 0; k < nv; k++) {
            plan.row_off[k] = pvar.row_offset(plan.var_idx[k]);
            plan.row_len[k] = pvar.row_length(plan.var_idx[k]);
        }
        plan.pvar_text = pvar.take_data();
    } catch (const MetaError &e) {
        throw PfileError{e.status, e.msg};
    }
    return plan;
}

void VcfPlan::materialize_prefixes() {
    const size_t nv = var_idx.size();
    if (prefix_off.size() == nv + 1) return;
    prefix_off.resize(nv + 1);
    uint64_t total = 0;
    for (size_t k = 0; k < nv; k++) {
        prefix_off[k] = total;
        total += (uint64_t)row_len[k] + 3;
    }
    prefix_off[nv] = total;
    prefix_blob.resize(total);
    uint8_t *base = prefix_blob.data();
    const unsigned T = (unsigned)std::max<size_t>(1, std::min<size_t>(nv, MetaTable::worker_threads(total)));
    auto fill = [&](unsigned t) {
        const size_t a = nv / T * t, b = t + 1 == T ? nv : nv / T * (t + 1);
        for (size_t k = a; k < b; k++) {
            uint8_t *w = base + prefix_off[k];
            memcpy(w, pvar_text.data() + row_off[k], row_len[k]);
            w += row_len[k];
            w[0] = '\t';
            w[1] = 'G';
            w[2] = 'T';
        }
    };
    if (T == 1) fill(0);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; t++) th.emplace_back(fill, t);
        for (auto &x : th) x.join();
    }
}

void Pfile::output_vcf(const std::optional<std::string> &sam_query, const std::optional<std::string> &var_query,
                       const std::string &filename, const int *device_ids, int n_devices, pgb_stats *stats) const {
    VcfPlan plan = plan_vcf(sam_query, var_query);
    int fd = ::open(filename.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644); // File::create, pfile.rs:136
    if (fd < 0) throw PfileError{PGB_E_IO, "create " + filename + ": " + strerror(errno)};
    struct Closer {
        int fd;
        ~Closer() { ::close(fd); }
    } closer{fd};
    write_all(fd, plan.header.data(), plan.header.size());
    pgb_file *f = nullptr;
    int rc = pgb_open(pgen_path().c_str(), &f); // File::open(self.pgen_path()), pfile.rs:149
    if (rc != PGB_OK) throw PfileError{rc, std::string(pgb_strerror(rc)) + ": " + pgb_last_error()};
    // an empty std::vector may hand out nullptr, which the C ABI reads as "all samples"
    static const uint32_t no_samples[1] = {0};
    const uint32_t *sam = plan.sam_idx.empty() ? no_samples : plan.sam_idx.data();
    rc = pgb_export_gt_vcf_rows(f, plan.var_idx.data(), plan.var_idx.size(), sam, plan.sam_idx.size(),
                                (const uint8_t *)plan.pvar_text.data(), plan.pvar_text.size(), plan.row_off.data(),
                                plan.row_len.data(), fd, device_ids, n_devices, stats);
    std::string detail = pgb_last_error();
    pgb_close(f);
    if (rc != PGB_OK) throw PfileError{rc, std::string(pgb_strerror(rc)) + ": " + detail};
}

} // namespace pgb
