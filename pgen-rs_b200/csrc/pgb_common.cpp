// pgb_common.cpp — status strings, thread-local error detail, record geometry.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "pgb_internal.h"

static thread_local char t_err[512];

extern "C" void pgb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
}
extern "C" void pgb_clear_error(void) { t_err[0] = 0; }
extern "C" const char *pgb_last_error(void) { return t_err; }
extern "C" int pgb_abi_version(void) { return PGB_ABI_VERSION; }

extern "C" const char *pgb_strerror(int s) {
    switch (s) {
    case PGB_OK: return "ok";
    case PGB_E_IO: return "I/O error";
    case PGB_E_MAGIC: return "not a .pgen file (magic != 6C 1B)";
    case PGB_E_MODE: return "unsupported .pgen storage mode (only 0x02 fixed-width hardcalls)";
    case PGB_E_FLAGS: return "unsupported .pgen header flags (byte 11 != 0x40)";
    case PGB_E_ARG: return "invalid argument";
    case PGB_E_RANGE: return "variant or sample index out of range";
    case PGB_E_NO_DEVICE: return "no usable CUDA device (libpgb200 has no CPU fallback)";
    case PGB_E_CUDA: return "CUDA error";
    case PGB_E_NOMEM: return "out of memory";
    case PGB_E_NO_HEADER: return "metadata file has no leading '#' header line";
    case PGB_E_NO_IID: return "IID not among the headers of the .psam";
    case PGB_E_CSV: return "malformed metadata table";
    case PGB_E_EXPR: return "expression error";
    case PGB_E_SPACE: return "output buffer too small";
    default: return "unknown status";
    }
}

// variant_record_size, /root/reference/src/pfile.rs:196-200 (u32 arithmetic as written).
extern "C" uint32_t pgb_record_bytes(uint32_t n_samples) {
    uint32_t bit_size = n_samples * 2u;
    return bit_size / 8u + (bit_size % 8u == 0u ? 0u : 1u);
}

// /root/reference/src/pfile.rs:165 with the product taken in u64 (the reference's u32
// multiply wraps once var_idx * R >= 2^32; see DESIGN.md "divergences").
extern "C" uint64_t pgb_record_offset(uint64_t var_idx, uint32_t record_bytes) {
    return 12ull + var_idx * (uint64_t)record_bytes;
}

extern "C" uint64_t pgb_body_bytes(uint64_t n_var, uint64_t n_kept, const uint64_t *prefix_off) {
    uint64_t pfx = (prefix_off && n_var) ? prefix_off[n_var] - prefix_off[0] : 0;
    return pfx + n_var * (4ull * n_kept + 1ull);
}

extern "C" int pgb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
