// pfile.hpp — C++ mirror of the reference's `Pfile` (/root/reference/src/pfile.rs:19-194):
// same method names, argument meaning and failure points, with the genotype loop
// (pfile.rs:149-192) delegated to libpgb200's CUDA path through the C ABI.
//
// Rust is not available in this image, so this class plays the role of the patched
// pfile.rs; INTEGRATION.md shows the equivalent Rust shim.
#pragma once
#include <stdint.h>

#include <optional>
#include <string>
#include <vector>

#include "meta.hpp"
#include "pgb200.h"

namespace pgb {

struct PfileError {
    int status; // pgb_status
    std::string msg;
};

// What filter_metadata (pfile.rs:312-335) returns, flattened for the device: kept row
// indices in ascending file order.
std::vector<uint32_t> filter_metadata(MetaTable &table, const std::optional<std::string> &query);

// Everything output_vcf computes on the CPU before the genotype loop.
struct VcfPlan {
    std::vector<uint32_t> var_idx, sam_idx;
    std::string header;              // pfile.rs:139-146
    // The line prefixes (pfile.rs:157-161) are built on the device: the plan only carries the raw .pvar image
    // and, per kept variant, where its row sits in it.
    std::string pvar_text;
    std::vector<uint64_t> row_off;
    std::vector<uint32_t> row_len;
    // The finished prefixes as a blob + n_var + 1 offsets (what pgb_export_gt_vcf takes); built only on demand.
    void materialize_prefixes();
    std::vector<uint8_t> prefix_blob;
    std::vector<uint64_t> prefix_off;
};

class Pfile {
  public:
    std::string pfile_prefix;
    uint32_t num_variants = 0;
    uint32_t num_samples = 0;

    // Pfile::from_prefix, pfile.rs:38-76.  Throws PfileError (the reference panics).
    static Pfile from_prefix(const std::string &pfile_prefix);

    std::string pgen_path() const { return pfile_prefix + ".pgen"; }
    std::string psam_path() const { return pfile_prefix + ".psam"; }
    std::string pvar_path() const { return pfile_prefix + ".pvar"; }

    // pvar_reader / psam_reader, pfile.rs:285-287,308-310.
    MetaTable pvar_reader() const;
    MetaTable psam_reader() const;

    // query_metadata, pfile.rs:78-102: one line per kept row to out_fd.
    void query_metadata(MetaTable &reader, const std::optional<std::string> &query, const std::string &f_string,
                        int out_fd) const;

    // CPU half of output_vcf (pfile.rs:110-146 + the prefix bytes of :157-161).
    VcfPlan plan_vcf(const std::optional<std::string> &sam_query, const std::optional<std::string> &var_query) const;

    // output_vcf, pfile.rs:104-194.
    void output_vcf(const std::optional<std::string> &sam_query, const std::optional<std::string> &var_query,
                    const std::string &filename, const int *device_ids = nullptr, int n_devices = 0,
                    pgb_stats *stats = nullptr) const;
};

} // namespace pgb
