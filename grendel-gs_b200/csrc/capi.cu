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

// ---- per-stage device timing -------------------------------------------------------------------
#include <vector>
bool g_gs_profile_on = false;
int g_gs_debug_flags = 0;
extern "C" int gs_debug_set(int flags) {
    const int old = g_gs_debug_flags;
    g_gs_debug_flags = flags;
    return old;
}
namespace {
struct StageRec {
    std::vector<cudaEvent_t> begin, end;  // event pool, reused across reads
    size_t used = 0;
};
StageRec g_stage[GS_STAGE_NUM];
const char *kStageNames[GS_STAGE_NUM] = {
    "10 preprocess", "21-24 count local tiles", "30 InclusiveSum", "40 duplicateWithKeys", "50 SortPairs",
    "60 identifyTileRanges", "70 render", "b10 render", "b20 preprocess", "loss forward", "loss backward",
    "get_local2j_ids_bool", "all2all pack", "all2all unpack"};
}  // namespace

void gs_prof_mark(int stage, bool begin, cudaStream_t stream) {
    StageRec &r = g_stage[stage];
    if (begin) {
        if (r.used == r.begin.size()) {
            cudaEvent_t a, b;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
            r.begin.push_back(a);
            r.end.push_back(b);
        }
        cudaEventRecord(r.begin[r.used], stream);
    } else {
        if (r.used < r.end.size()) cudaEventRecord(r.end[r.used++], stream);
    }
}

extern "C" int gs_profile_enable(int on) {
    g_gs_profile_on = on != 0;
    if (on)
        for (int s = 0; s < GS_STAGE_NUM; s++) g_stage[s].used = 0;
    return GS_OK;
}

extern "C" int gs_profile_read(int stage, double *total_ms, int64_t *launches) {
    GS_REQUIRE(stage >= 0 && stage < GS_STAGE_NUM && total_ms && launches, "stage");
    StageRec &r = g_stage[stage];
    double tot = 0.0;
    for (size_t k = 0; k < r.used; k++) {
        GS_CUDA_TRY(cudaEventSynchronize(r.end[k]));
        float ms = 0.f;
        GS_CUDA_TRY(cudaEventElapsedTime(&ms, r.begin[k], r.end[k]));
        tot += ms;
    }
    *total_ms = tot;
    *launches = (int64_t)r.used;
    r.used = 0;
    return GS_OK;
}

extern "C" const char *gs_profile_stage_name(int stage) {
    return (stage >= 0 && stage < GS_STAGE_NUM) ? kStageNames[stage] : "?";
}
