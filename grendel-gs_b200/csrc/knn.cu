// simple_knn._C.distCUDA2 (/root/reference/scene/gaussian_model.py:20,163-166): mean SQUARED distance of every point to
// its 3 nearest neighbours, used once at start-up to initialise the Gaussian scales (log sqrt of it).  The reference's
// module is an absent submodule (bkerbl/simple-knn @ 44f7642, .SUBMODULES.json:24-28); its published algorithm is an
// EXACT 3-NN (Morton order + box pruning only skip work), self excluded by index, duplicates counted at distance 0.
// This is the exact answer by tiled brute force: every thread owns one query point and streams all points through
// shared memory.  O(N^2) but init-time only and compute-trivial: ~8 instructions per pair, i.e. ~1 s for 2 M points on
// a B200.  (A Morton / box-pruned version only matters beyond ~10 M initial points.)
// NOT yet run on a device (written after round 1's GPU budget was spent); the shim keeps its torch path until then.
#include <cfloat>

#include "common.cuh"

#define KN_THREADS 256

__global__ void __launch_bounds__(KN_THREADS)
k_knn3_mean_dist2(int N, const float *__restrict__ pts, float *__restrict__ out) {
    __shared__ float sx[KN_THREADS], sy[KN_THREADS], sz[KN_THREADS];
    const int i = blockIdx.x * KN_THREADS + threadIdx.x;
    const bool valid = i < N;
    const float qx = valid ? pts[3 * (size_t)i] : 0.f, qy = valid ? pts[3 * (size_t)i + 1] : 0.f,
                qz = valid ? pts[3 * (size_t)i + 2] : 0.f;
    float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;  // b0 <= b1 <= b2
    for (int base = 0; base < N; base += KN_THREADS) {
        const int j = base + threadIdx.x;
        __syncthreads();
        if (j < N) { sx[threadIdx.x] = pts[3 * (size_t)j]; sy[threadIdx.x] = pts[3 * (size_t)j + 1]; sz[threadIdx.x] = pts[3 * (size_t)j + 2]; }
        __syncthreads();
        const int cnt = min(KN_THREADS, N - base);
        for (int t = 0; t < cnt; t++) {
            const float dx = qx - sx[t], dy = qy - sy[t], dz = qz - sz[t];
            const float d = (base + t == i) ? FLT_MAX : dx * dx + dy * dy + dz * dz;   // self excluded by index
            if (d < b2) {
                if (d < b1) {
                    b2 = b1;
                    if (d < b0) { b1 = b0; b0 = d; } else { b1 = d; }
                } else {
                    b2 = d;
                }
            }
        }
    }
    if (valid) {
        // fewer than 3 other points: average the neighbours that exist (N == 1: 0)
        const int k = min(3, N - 1);
        float s = 0.f;
        if (k >= 1) s += b0;
        if (k >= 2) s += b1;
        if (k >= 3) s += b2;
        out[i] = k > 0 ? s / (float)k : 0.f;
    }
}

// points: (N,3) fp32; mean_dist2: (N) fp32.
extern "C" int gs_knn3_mean_dist2(int N, const float *points, float *mean_dist2, void *stream) {
    GS_REQUIRE(N >= 0, "N");
    if (N == 0) return GS_OK;
    GS_REQUIRE(points && mean_dist2, "null pointer");
    k_knn3_mean_dist2<<<(N + KN_THREADS - 1) / KN_THREADS, KN_THREADS, 0, (cudaStream_t)stream>>>(N, points, mean_dist2);
    GS_LAUNCH_CHECK();
    return GS_OK;
}
