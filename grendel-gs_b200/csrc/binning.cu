// Tile binning: CUDA stages 21-24 (local tile count), 30 (scan), 40 (duplicate with keys),
// 50 (radix sort), 60 (tile ranges) of /root/reference/analyze_statistic.py:1976-1980, i.e. the
// first half of GaussianRasterizer.render_gaussians (/root/reference/gaussian_renderer/__init__.py:1271).
//
// All of this is integer / byte work bound by HBM traffic.  The scan and the sort are CUB device
// primitives (the reference's stages 30 and 50 are CUB calls too); the sort only touches the
// 32 + ceil(log2 T) key bits that can differ.
#include <cub/cub.cuh>

#include "common.cuh"

#define BIN_THREADS 256

// Stages 21-24: per-splat number of LOCAL tiles its rectangle touches + the packed 48-byte record the
// blend kernels gather:
//   r0 = (mx, my, a', b')      a' = -A/2, b' = -B, c' = -C/2  so that  power = a'dx^2 + b'dx dy + c'dy^2
//   r1 = (c', opacity, thr, red)     thr = ln(1/(255*opacity)) - margin: power < thr  =>  alpha < 1/255
//   r2 = (green, blue, ex, ey)       half extents of the bounding box of {power >= thr} (+0.5 px slack),
//                                    used by the blend kernels to cull splats per 4x4 pixel block
#define GS_THR_MARGIN 0.02f
__global__ void __launch_bounds__(BIN_THREADS)
k_count_tiles(int P, int W, int H, const float *__restrict__ means2D, const float *__restrict__ conic_opacity,
              const float *__restrict__ rgb, const int32_t *__restrict__ radii,
              const uint8_t *__restrict__ compute_locally, uint32_t *__restrict__ touched, float *__restrict__ rec) {
    const int i = blockIdx.x * BIN_THREADS + threadIdx.x;
    if (i >= P) return;
    const int gx = (W + GS_BLOCK_X - 1) / GS_BLOCK_X, gy = (H + GS_BLOCK_Y - 1) / GS_BLOCK_Y;
    const int r = radii[i];
    uint32_t n = 0;
    const float2 m = *reinterpret_cast<const float2 *>(means2D + 2 * i);
    if (r > 0) {
        int x0, y0, x1, y1;
        gs_get_rect(m.x, m.y, r, gx, gy, x0, y0, x1, y1);
        for (int y = y0; y < y1; y++) {
            const uint8_t *row = compute_locally + y * gx;
            for (int x = x0; x < x1; x++) n += row[x] ? 1u : 0u;
        }
    }
    touched[i] = n;
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
    if (n > 0) {
        const float4 co = *reinterpret_cast<const float4 *>(conic_opacity + 4 * i);
        const float thr = -logf(255.0f * fmaxf(co.w, 1e-30f)) - GS_THR_MARGIN;
        // {d : A dx^2 + 2B dx dy + C dy^2 <= -2 thr} has half extents sqrt(t C/det), sqrt(t A/det)
        const float t = -2.f * thr, det = co.x * co.z - co.y * co.y;
        float ex = -1.f, ey = -1.f;  // never contributes
        if (t > 0.f) {
            if (det > 0.f && co.x > 0.f && co.z > 0.f) {
                ex = sqrtf(t * co.z / det) * 1.001f + 0.5f;
                ey = sqrtf(t * co.x / det) * 1.001f + 0.5f;
            } else {
                ex = ey = 3.0e38f;  // degenerate conic: never cull
            }
        }
        r0 = make_float4(m.x, m.y, -0.5f * co.x, -co.y);
        r1 = make_float4(-0.5f * co.z, co.w, thr, rgb[3 * i]);
        r2 = make_float4(rgb[3 * i + 1], rgb[3 * i + 2], ex, ey);
    }
    float4 *o = reinterpret_cast<float4 *>(rec + (size_t)GS_REC_FLOATS * i);
    o[0] = r0; o[1] = r1; o[2] = r2;
}

// Stage 40: one (tile << 32 | depth bits, splat id) pair per (splat, local tile), in splat order.
__global__ void __launch_bounds__(BIN_THREADS)
k_duplicate(int P, int W, int H, const float *__restrict__ means2D, const float *__restrict__ depths,
            const int32_t *__restrict__ radii, const uint8_t *__restrict__ compute_locally,
            const uint32_t *__restrict__ offsets, uint64_t *__restrict__ keys, uint32_t *__restrict__ ids) {
    const int i = blockIdx.x * BIN_THREADS + threadIdx.x;
    if (i >= P) return;
    const int r = radii[i];
    if (r <= 0) return;
    uint32_t off = (i == 0) ? 0u : offsets[i - 1];
    if (offsets[i] == off) return;
    const int gx = (W + GS_BLOCK_X - 1) / GS_BLOCK_X, gy = (H + GS_BLOCK_Y - 1) / GS_BLOCK_Y;
    const float2 m = *reinterpret_cast<const float2 *>(means2D + 2 * i);
    int x0, y0, x1, y1;
    gs_get_rect(m.x, m.y, r, gx, gy, x0, y0, x1, y1);
    const uint64_t dbits = (uint64_t)__float_as_uint(depths[i]);
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
            const int t = y * gx + x;
            if (!compute_locally[t]) continue;
            keys[off] = ((uint64_t)(uint32_t)t << 32) | dbits;
            ids[off] = (uint32_t)i;
            off++;
        }
}

// Stage 60: [start,end) of every tile in the sorted list.
__global__ void __launch_bounds__(BIN_THREADS)
k_tile_ranges(int64_t R, const uint64_t *__restrict__ keys, uint32_t *__restrict__ ranges) {
    const int64_t k = (int64_t)blockIdx.x * BIN_THREADS + threadIdx.x;
    if (k >= R) return;
    const uint32_t t = (uint32_t)(keys[k] >> 32);
    if (k == 0 || t != (uint32_t)(keys[k - 1] >> 32)) ranges[2 * t] = (uint32_t)k;
    if (k == R - 1 || t != (uint32_t)(keys[k + 1] >> 32)) ranges[2 * t + 1] = (uint32_t)(k + 1);
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" size_t gs_render_count_temp_bytes(int P) {
    size_t scan = 0;
    cub::DeviceScan::InclusiveSum(nullptr, scan, (const uint32_t *)nullptr, (uint32_t *)nullptr, P > 0 ? P : 1);
    return align_up((size_t)(P > 0 ? P : 1) * sizeof(uint32_t), 256) + align_up(scan, 256) + 256;
}

extern "C" int gs_render_count(int P, int image_height, int image_width, const float *means2D,
                               const float *conic_opacity, const float *rgb, const int32_t *radii,
                               const uint8_t *compute_locally, uint32_t *offsets, float *rec, void *temp,
                               size_t temp_bytes, int64_t *R_host, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GS_REQUIRE(P >= 0 && image_height > 0 && image_width > 0, "sizes");
    GS_REQUIRE(R_host != nullptr, "R_host");
    *R_host = 0;
    if (P == 0) return GS_OK;
    GS_REQUIRE(means2D && conic_opacity && rgb && radii && compute_locally && offsets && rec && temp, "null pointer");
    GS_REQUIRE(((uintptr_t)rec & 15) == 0 && ((uintptr_t)conic_opacity & 15) == 0 && ((uintptr_t)means2D & 7) == 0,
               "alignment");
    if (temp_bytes < gs_render_count_temp_bytes(P)) {
        gs_set_error("gs_render_count: temp too small (%zu < %zu)", temp_bytes, gs_render_count_temp_bytes(P));
        return GS_ENOMEM;
    }
    uint32_t *touched = (uint32_t *)temp;
    char *scan_temp = (char *)temp + align_up((size_t)P * sizeof(uint32_t), 256);
    size_t scan_bytes = temp_bytes - align_up((size_t)P * sizeof(uint32_t), 256);
    const int grid = (P + BIN_THREADS - 1) / BIN_THREADS;
    {
        GsStageTimer timer(GS_STAGE_COUNT_TILES, stream);
        k_count_tiles<<<grid, BIN_THREADS, 0, stream>>>(P, image_width, image_height, means2D, conic_opacity, rgb, radii,
                                                        compute_locally, touched, rec);
        GS_LAUNCH_CHECK();
    }
    {
        GsStageTimer timer(GS_STAGE_SCAN, stream);
        GS_CUDA_TRY(cub::DeviceScan::InclusiveSum(scan_temp, scan_bytes, touched, offsets, P, stream));
    }
    uint32_t last = 0;
    GS_CUDA_TRY(cudaMemcpyAsync(&last, offsets + (P - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    GS_CUDA_TRY(cudaStreamSynchronize(stream));
    *R_host = (int64_t)last;
    return GS_OK;
}

static int key_bits(int T) {
    int b = 0;
    while ((1ll << b) < (long long)T) b++;
    return 32 + (b > 0 ? b : 1);
}

extern "C" size_t gs_render_sort_temp_bytes(int64_t R) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                    (const uint32_t *)nullptr, (uint32_t *)nullptr, R > 0 ? R : 1, 0, 64);
    return align_up(bytes, 256) + 256;
}

// Implemented in blend.cu
int gs_launch_blend_forward(int64_t R, int H, int W, const float *rec, const float *bg, const uint8_t *compute_locally,
                            const uint32_t *ranges, const uint32_t *ids_sorted, float *image, float *final_T,
                            uint32_t *n_contrib, int64_t *stats, cudaStream_t stream);

extern "C" int gs_render_forward(int P, int64_t R, int image_height, int image_width, const float *means2D,
                                 const float *depths, const int32_t *radii, const uint8_t *compute_locally,
                                 const uint32_t *offsets, const float *rec, const float *bg, uint64_t *keys_unsorted,
                                 uint32_t *ids_unsorted, uint64_t *keys_sorted, uint32_t *ids_sorted, void *sort_temp,
                                 size_t sort_temp_bytes, uint32_t *ranges, float *image, float *final_T,
                                 uint32_t *n_contrib, int64_t *stats, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GS_REQUIRE(P >= 0 && R >= 0 && image_height > 0 && image_width > 0, "sizes");
    GS_REQUIRE(R < (1ll << 32), "more than 2^32 splat-tile instances");
    GS_REQUIRE(compute_locally && bg && ranges && image && final_T && n_contrib, "null pointer");
    const int gx = (image_width + GS_BLOCK_X - 1) / GS_BLOCK_X, gy = (image_height + GS_BLOCK_Y - 1) / GS_BLOCK_Y;
    const int T = gx * gy;
    GS_CUDA_TRY(cudaMemsetAsync(ranges, 0, sizeof(uint32_t) * 2 * (size_t)T, stream));
    if (R > 0) {
        GS_REQUIRE(means2D && depths && radii && offsets && rec && keys_unsorted && ids_unsorted && keys_sorted &&
                       ids_sorted && sort_temp,
                   "null pointer");
        if (sort_temp_bytes < gs_render_sort_temp_bytes(R)) {
            gs_set_error("gs_render_forward: sort temp too small");
            return GS_ENOMEM;
        }
        {
            GsStageTimer timer(GS_STAGE_DUPLICATE, stream);
            k_duplicate<<<(P + BIN_THREADS - 1) / BIN_THREADS, BIN_THREADS, 0, stream>>>(
                P, image_width, image_height, means2D, depths, radii, compute_locally, offsets, keys_unsorted, ids_unsorted);
            GS_LAUNCH_CHECK();
        }
        {
            GsStageTimer timer(GS_STAGE_SORT, stream);
            GS_CUDA_TRY(cub::DeviceRadixSort::SortPairs(sort_temp, sort_temp_bytes, keys_unsorted, keys_sorted,
                                                        ids_unsorted, ids_sorted, R, 0, key_bits(T), stream));
        }
        {
            GsStageTimer timer(GS_STAGE_RANGES, stream);
            k_tile_ranges<<<(unsigned)((R + BIN_THREADS - 1) / BIN_THREADS), BIN_THREADS, 0, stream>>>(R, keys_sorted, ranges);
            GS_LAUNCH_CHECK();
        }
    }
    return gs_launch_blend_forward(R, image_height, image_width, rec, bg, compute_locally, ranges, ids_sorted, image,
                                   final_T, n_contrib, stats, stream);
}
