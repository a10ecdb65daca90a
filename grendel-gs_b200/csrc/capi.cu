// C-ABI plumbing: error text, version, compile-time tile constants.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void gs_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *gs_last_error(void) { return g_err; }

extern "C" const char *gs_version(void) { return "grendel-gs_b200 sm_100a r1"; }

// _C.get_block_XY -- /root/reference/arguments/__init__.py:254-257
extern "C" int gs_get_block_xy(int *block_x, int *block_y, int *one_dim_block_size) {
    GS_REQUIRE(block_x && block_y && one_dim_block_size, "null pointer");
    *block_x = GS_BLOCK_X;
    *block_y = GS_BLOCK_Y;
    *one_dim_block_size = GS_ONE_DIM_BLOCK_SIZE;
    return GS_OK;
}
