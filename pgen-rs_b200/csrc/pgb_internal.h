// pgb_internal.h — shared by the translation units of libpgb200.so (not installed).
#pragma once
#include <stdint.h>

#include "pgb200.h"

#ifdef __cplusplus
extern "C" {
#endif
// printf-style; stores a thread-local message returned by pgb_last_error().
void pgb_set_error(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
void pgb_clear_error(void);
#ifdef __cplusplus
}
#endif
