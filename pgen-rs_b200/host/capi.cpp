// capi.cpp — extern "C" wrappers over the C++ host mirror (pfile.hpp) so that tests and
// foreign hosts can drive Pfile::output_vcf / query_metadata / the CPU planning stage
// without C++ types crossing the boundary.
#include <new>
#include <optional>
#include <string>

#include "../csrc/pgb_internal.h"
#include "pfile.hpp"

struct pgb_plan {
    pgb::VcfPlan plan;
};

namespace {
std::optional<std::string> opt(const char *s) { return s ? std::optional<std::string>(s) : std::nullopt; }

template <typename F>
int guarded(F &&fn) {
    pgb_clear_error();
    try {
        fn();
        return PGB_OK;
    } catch (const pgb::PfileError &e) {
        pgb_set_error("%s", e.msg.c_str());
        return e.status;
    } catch (const std::bad_alloc &) {
        return PGB_E_NOMEM;
    }
}
} // namespace

extern "C" int pgb_pfile_output_vcf(const char *prefix, const char *sam_query, const char *var_query,
                                    const char *out_path, const int *device_ids, int n_devices, pgb_stats *stats) {
    if (!prefix || !out_path) return PGB_E_ARG;
    return guarded([&] {
        pgb::Pfile p = pgb::Pfile::from_prefix(prefix);
        p.output_vcf(opt(sam_query), opt(var_query), out_path, device_ids, n_devices, stats);
    });
}

extern "C" int pgb_pfile_query(const char *prefix, const char *fstring, const char *query, int samples, int out_fd) {
    if (!prefix || !fstring) return PGB_E_ARG;
    return guarded([&] {
        pgb::Pfile p = pgb::Pfile::from_prefix(prefix);
        pgb::MetaTable t = samples ? p.psam_reader() : p.pvar_reader();
        p.query_metadata(t, opt(query), fstring, out_fd);
    });
}

extern "C" int pgb_plan_vcf(const char *prefix, const char *sam_query, const char *var_query, pgb_plan **out) {
    if (!prefix || !out) return PGB_E_ARG;
    *out = nullptr;
    return guarded([&] {
        pgb::Pfile p = pgb::Pfile::from_prefix(prefix);
        pgb_plan *pl = new pgb_plan();
        try {
            pl->plan = p.plan_vcf(opt(sam_query), opt(var_query));
        } catch (...) {
            delete pl;
            throw;
        }
        *out = pl;
    });
}

extern "C" void pgb_plan_free(pgb_plan *p) { delete p; }
extern "C" uint64_t pgb_plan_n_var(const pgb_plan *p) { return p ? p->plan.var_idx.size() : 0; }
extern "C" uint64_t pgb_plan_n_sam(const pgb_plan *p) { return p ? p->plan.sam_idx.size() : 0; }
extern "C" const uint32_t *pgb_plan_var_idx(const pgb_plan *p) { return p ? p->plan.var_idx.data() : nullptr; }
extern "C" const uint32_t *pgb_plan_sam_idx(const pgb_plan *p) { return p ? p->plan.sam_idx.data() : nullptr; }
extern "C" const uint8_t *pgb_plan_header(const pgb_plan *p, uint64_t *len) {
    if (!p) return nullptr;
    if (len) *len = p->plan.header.size();
    return (const uint8_t *)p->plan.header.data();
}
extern "C" const uint8_t *pgb_plan_pvar_text(const pgb_plan *p, uint64_t *len) {
    if (!p) return nullptr;
    if (len) *len = p->plan.pvar_text.size();
    return (const uint8_t *)p->plan.pvar_text.data();
}
extern "C" const uint64_t *pgb_plan_row_off(const pgb_plan *p) { return p ? p->plan.row_off.data() : nullptr; }
extern "C" const uint32_t *pgb_plan_row_len(const pgb_plan *p) { return p ? p->plan.row_len.data() : nullptr; }
extern "C" const uint8_t *pgb_plan_prefix_blob(const pgb_plan *p, uint64_t *len) {
    if (!p) return nullptr;
    const_cast<pgb_plan *>(p)->plan.materialize_prefixes();
    if (len) *len = p->plan.prefix_blob.size();
    return p->plan.prefix_blob.data();
}
extern "C" const uint64_t *pgb_plan_prefix_off(const pgb_plan *p) {
    if (!p) return nullptr;
    const_cast<pgb_plan *>(p)->plan.materialize_prefixes();
    return p->plan.prefix_off.data();
}
