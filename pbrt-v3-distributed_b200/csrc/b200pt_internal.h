// b200pt_internal.h -- shared between the translation units of libb200pt.so.
#ifndef B200PT_INTERNAL_H
#define B200PT_INTERNAL_H

#include "../../include/b200pt.h"

// Records `msg` (printf-style) as the calling thread's last error and returns `code`.
int b200pt_fail(int code, const char *fmt, ...);

#endif
