#include "common.cuh"

#include <cstring>

namespace s2i {

static thread_local char g_err[1024] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

const char* last_error() { return g_err; }

}  // namespace s2i
