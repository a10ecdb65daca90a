// Splat -> strip-owner routing and all-to-all staging.
//   gs_get_local2j_ids_bool        /root/reference/gaussian_renderer/workload_division.py:721-744
//   gs_get_local2j_ids_bool_rects  workload_division.py:471-484 (legacy "adjust_mode6")
//   gs_mask_scan/pack/unpack       replace the per-(destination, camera) nonzero() + index_select + cat
//                                  glue of gaussian_renderer/__init__.py:590-607,651-658
// Pure integer / byte work, HBM bound: 12 B read and world_size bytes written per splat.
#include <cub/cub.cuh>

#include "common.cuh"

#define DT_THREADS 256
#define ROW_FLOATS 11

__global__ void __launch_bounds__(DT_THREADS)
k_local2j(int P, int W, int H, int world_size, const float *__restrict__ means2D, const int32_t *__restrict__ radii,
          const int32_t *__restrict__ strategy, uint8_t *__restrict__ out) {
    const int i = blockIdx.x * DT_THREADS + threadIdx.x;
    if (i >= P) return;
    const int gx = (W + GS_BLOCK_X - 1) / GS_BLOCK_X, gy = (H + GS_BLOCK_Y - 1) / GS_BLOCK_Y;
    const int r = radii[i];
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    if (r > 0) {
        const float2 m = *reinterpret_cast<const float2 *>(means2D + 2 * i);
        gs_get_rect(m.x, m.y, r, gx, gy, x0, y0, x1, y1);
    }
    for (int j = 0; j < world_size; j++) {
        const int lo_j = strategy[j], hi_j = strategy[j + 1];
        bool hit = false;
        if (x1 > x0)
            for (int y = y0; y < y1 && !hit; y++) hit = max(y * gx + x0, lo_j) < min(y * gx + x1, hi_j);
        out[(size_t)i * world_size + j] = hit ? 1 : 0;
    }
}

__global__ void __launch_bounds__(DT_THREADS)
k_local2j_rects(int P, int W, int H, int world_size, const float *__restrict__ means2D,
                const int32_t *__restrict__ radii, const int32_t *__restrict__ rects, uint8_t *__restrict__ out) {
    const int i = blockIdx.x * DT_THREADS + threadIdx.x;
    if (i >= P) return;
    const int gx = (W + GS_BLOCK_X - 1) / GS_BLOCK_X, gy = (H + GS_BLOCK_Y - 1) / GS_BLOCK_Y;
    const int r = radii[i];
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    if (r > 0) {
        const float2 m = *reinterpret_cast<const float2 *>(means2D + 2 * i);
        gs_get_rect(m.x, m.y, r, gx, gy, x0, y0, x1, y1);
    }
    for (int j = 0; j < world_size; j++) {
        const int32_t *q = rects + 4 * j;
        const bool hit = r > 0 && max(y0, q[0]) < min(y1, q[1]) && max(x0, q[2]) < min(x1, q[3]);
        out[(size_t)i * world_size + j] = hit ? 1 : 0;
    }
}

extern "C" int gs_get_local2j_ids_bool(int P, int image_height, int image_width, int world_size, const float *means2D,
                                       const int32_t *radii, const int32_t *strategy, uint8_t *out, void *stream) {
    GS_REQUIRE(P >= 0 && world_size > 0 && image_height > 0 && image_width > 0, "sizes");
    if (P == 0) return GS_OK;
    GS_REQUIRE(means2D && radii && strategy && out, "null pointer");
    k_local2j<<<(P + DT_THREADS - 1) / DT_THREADS, DT_THREADS, 0, (cudaStream_t)stream>>>(
        P, image_width, image_height, world_size, means2D, radii, strategy, out);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

extern "C" int gs_get_local2j_ids_bool_rects(int P, int image_height, int image_width, int world_size,
                                             const float *means2D, const int32_t *radii, const int32_t *rects,
                                             uint8_t *out, void *stream) {
    GS_REQUIRE(P >= 0 && world_size > 0 && image_height > 0 && image_width > 0, "sizes");
    if (P == 0) return GS_OK;
    GS_REQUIRE(means2D && radii && rects && out, "null pointer");
    k_local2j_rects<<<(P + DT_THREADS - 1) / DT_THREADS, DT_THREADS, 0, (cudaStream_t)stream>>>(
        P, image_width, image_height, world_size, means2D, radii, rects, out);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

// ---- all-to-all staging ---------------------------------------------------------------------
struct MaskColumn {
    const uint8_t *mask;
    int world_size, column;
    __host__ __device__ int32_t operator()(int i) const { return mask[(size_t)i * world_size + column] ? 1 : 0; }
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" size_t gs_mask_scan_temp_bytes(int P) {
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const int32_t *)nullptr, (int32_t *)nullptr, P > 0 ? P : 1);
    return align_up(b, 256) + 256;
}

__global__ void k_mask_count(int P, int world_size, int column, const uint8_t *__restrict__ mask,
                             const int32_t *__restrict__ pos, int32_t *__restrict__ count) {
    // total = exclusive position of the last splat + its own flag
    *count = pos[P - 1] + (mask[(size_t)(P - 1) * world_size + column] ? 1 : 0);
}

extern "C" int gs_mask_scan(int P, int world_size, int column, const uint8_t *mask, int32_t *pos, int32_t *count,
                            void *temp, size_t temp_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GS_REQUIRE(P >= 0 && world_size > 0 && column >= 0 && column < world_size, "sizes");
    GS_REQUIRE(count != nullptr, "count");
    if (P == 0) {
        GS_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int32_t), stream));
        return GS_OK;
    }
    GS_REQUIRE(mask && pos && temp, "null pointer");
    cub::CountingInputIterator<int> idx(0);
    cub::TransformInputIterator<int32_t, MaskColumn, cub::CountingInputIterator<int>> it(idx, MaskColumn{mask, world_size, column});
    GS_CUDA_TRY(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, it, pos, P, stream));
    k_mask_count<<<1, 1, 0, stream>>>(P, world_size, column, mask, pos, count);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

__global__ void __launch_bounds__(DT_THREADS)
k_pack_rows(int P, int world_size, int column, const uint8_t *__restrict__ mask, const int32_t *__restrict__ pos,
            const float *__restrict__ means2D, const float *__restrict__ rgb, const float *__restrict__ conic_opacity,
            const int32_t *__restrict__ radii, const float *__restrict__ depths, float *__restrict__ out) {
    const int i = blockIdx.x * DT_THREADS + threadIdx.x;
    if (i >= P || !mask[(size_t)i * world_size + column]) return;
    float *o = out + (size_t)pos[i] * ROW_FLOATS;
    const float4 co = *reinterpret_cast<const float4 *>(conic_opacity + 4 * i);
    o[0] = means2D[2 * i]; o[1] = means2D[2 * i + 1];
    o[2] = rgb[3 * i]; o[3] = rgb[3 * i + 1]; o[4] = rgb[3 * i + 2];
    o[5] = co.x; o[6] = co.y; o[7] = co.z; o[8] = co.w;
    o[9] = (float)radii[i];
    o[10] = depths[i];
}

extern "C" int gs_pack_rows(int P, int world_size, int column, const uint8_t *mask, const int32_t *pos,
                            const float *means2D, const float *rgb, const float *conic_opacity, const int32_t *radii,
                            const float *depths, float *out, void *stream) {
    GS_REQUIRE(P >= 0 && world_size > 0 && column >= 0 && column < world_size, "sizes");
    if (P == 0) return GS_OK;
    GS_REQUIRE(mask && pos && means2D && rgb && conic_opacity && radii && depths && out, "null pointer");
    k_pack_rows<<<(P + DT_THREADS - 1) / DT_THREADS, DT_THREADS, 0, (cudaStream_t)stream>>>(
        P, world_size, column, mask, pos, means2D, rgb, conic_opacity, radii, depths, out);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

__global__ void __launch_bounds__(DT_THREADS)
k_unpack_rows(int n, const float *__restrict__ rows, float *__restrict__ means2D, float *__restrict__ rgb,
              float *__restrict__ conic_opacity, int32_t *__restrict__ radii, float *__restrict__ depths) {
    const int i = blockIdx.x * DT_THREADS + threadIdx.x;
    if (i >= n) return;
    const float *r = rows + (size_t)i * ROW_FLOATS;
    means2D[2 * i] = r[0]; means2D[2 * i + 1] = r[1];
    rgb[3 * i] = r[2]; rgb[3 * i + 1] = r[3]; rgb[3 * i + 2] = r[4];
    *reinterpret_cast<float4 *>(conic_opacity + 4 * i) = make_float4(r[5], r[6], r[7], r[8]);
    radii[i] = (int32_t)r[9];
    depths[i] = r[10];
}

extern "C" int gs_unpack_rows(int n, const float *rows, float *means2D, float *rgb, float *conic_opacity,
                              int32_t *radii, float *depths, void *stream) {
    GS_REQUIRE(n >= 0, "n");
    if (n == 0) return GS_OK;
    GS_REQUIRE(rows && means2D && rgb && conic_opacity && radii && depths, "null pointer");
    k_unpack_rows<<<(n + DT_THREADS - 1) / DT_THREADS, DT_THREADS, 0, (cudaStream_t)stream>>>(n, rows, means2D, rgb,
                                                                                             conic_opacity, radii, depths);
    GS_LAUNCH_CHECK();
    return GS_OK;
}
