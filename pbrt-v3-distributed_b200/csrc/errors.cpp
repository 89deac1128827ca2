// errors.cpp -- thread-local last-error string of the C ABI (b200pt_last_error).
#include <cstdarg>
#include <cstdio>

#include "b200pt_internal.h"

static thread_local char g_last_error[512] = "";

int b200pt_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *b200pt_last_error(void) { return g_last_error; }
extern "C" int b200pt_abi_version(void) { return B200PT_ABI_VERSION; }
