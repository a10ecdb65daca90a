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
    GsStageTimer timer(GS_STAGE_LOCAL2J, (cudaStream_t)stream);
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
// One camera's routing mask (P, ncols) -> dense 11-float rows per destination column, splat order kept
// (the order the reference gets from nonzero(): gaussian_renderer/__init__.py:590-607, workload_division.py:741).
// Row = means2D(2) rgb(3) conic_opacity(4) radius-as-float depth   (gaussian_renderer/__init__.py:651-658)
// Gradient row = d means2D(2) d rgb(3) d conic_opacity(4).
#define GRAD_FLOATS 9
#define MAX_COLS 16

struct MaskColMajor {  // element e = c*P + i of the column-major traversal of the (P, ncols) mask
    const uint8_t *mask;
    int P, ncols;
    __host__ __device__ int32_t operator()(int e) const {
        const int c = e / P, i = e - c * P;
        return mask[(size_t)i * ncols + c] ? 1 : 0;
    }
};

struct Cols { int32_t v[MAX_COLS]; };

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" size_t gs_route_scan_temp_bytes(int P, int ncols) {
    size_t b = 0;
    const int n = (P > 0 ? P : 1) * (ncols > 0 ? ncols : 1);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const int32_t *)nullptr, (int32_t *)nullptr, n);
    return align_up(b, 256) + 256;
}

__global__ void k_colstart(int P, int ncols, const uint8_t *__restrict__ mask, const int32_t *__restrict__ gpos,
                           int32_t *__restrict__ colstart) {
    const int c = threadIdx.x;
    if (c < ncols) colstart[c] = gpos[(size_t)c * P];
    if (c == ncols) colstart[ncols] = gpos[(size_t)ncols * P - 1] + (mask[(size_t)(P - 1) * ncols + (ncols - 1)] ? 1 : 0);
}

// gpos: (ncols*P) int32, exclusive rank of every (column, splat) in the column-major flag stream;
// colstart: (ncols+1) int32, colstart[c] = gpos of column c's first element, colstart[ncols] = total flags;
// so column c holds colstart[c+1]-colstart[c] rows and splat i's row inside it is gpos[c*P+i]-colstart[c].
extern "C" int gs_route_scan(int P, int ncols, const uint8_t *mask, int32_t *gpos, int32_t *colstart, void *temp,
                             size_t temp_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GS_REQUIRE(P >= 0 && ncols > 0 && ncols <= MAX_COLS, "sizes (at most 16 destination columns)");
    GS_REQUIRE(colstart != nullptr, "colstart");
    if (P == 0) {
        GS_CUDA_TRY(cudaMemsetAsync(colstart, 0, sizeof(int32_t) * (ncols + 1), stream));
        return GS_OK;
    }
    GS_REQUIRE(mask && gpos && temp, "null pointer");
    GS_REQUIRE((long long)P * ncols < (1ll << 31), "P*ncols overflows int32");
    cub::CountingInputIterator<int> idx(0);
    cub::TransformInputIterator<int32_t, MaskColMajor, cub::CountingInputIterator<int>> it(idx, MaskColMajor{mask, P, ncols});
    GsStageTimer timer(GS_STAGE_PACK, stream);
    GS_CUDA_TRY(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, it, gpos, P * ncols, stream));
    k_colstart<<<1, 32, 0, stream>>>(P, ncols, mask, gpos, colstart);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

__global__ void __launch_bounds__(DT_THREADS)
k_pack_rows(int P, int ncols, const uint8_t *__restrict__ mask, const int32_t *__restrict__ gpos,
            const int32_t *__restrict__ colstart, Cols dst_off, const float *__restrict__ means2D,
            const float *__restrict__ rgb, const float *__restrict__ conic_opacity, const int32_t *__restrict__ radii,
            const float *__restrict__ depths, float *__restrict__ out) {
    const int i = blockIdx.x * DT_THREADS + threadIdx.x;
    if (i >= P) return;
    bool any = false;
    for (int c = 0; c < ncols; c++) any |= mask[(size_t)i * ncols + c] != 0;
    if (!any) return;
    const float2 m = *reinterpret_cast<const float2 *>(means2D + 2 * i);
    const float4 co = *reinterpret_cast<const float4 *>(conic_opacity + 4 * i);
    const float r0 = rgb[3 * i], r1 = rgb[3 * i + 1], r2 = rgb[3 * i + 2];
    const float rad = (float)radii[i], dep = depths[i];
    for (int c = 0; c < ncols; c++) {
        if (!mask[(size_t)i * ncols + c]) continue;
        float *o = out + (size_t)(dst_off.v[c] + gpos[(size_t)c * P + i] - colstart[c]) * ROW_FLOATS;
        o[0] = m.x; o[1] = m.y; o[2] = r0; o[3] = r1; o[4] = r2;
        o[5] = co.x; o[6] = co.y; o[7] = co.z; o[8] = co.w; o[9] = rad; o[10] = dep;
    }
}

// dst_off_host: HOST array (ncols): first row of column c inside `out`.
extern "C" int gs_pack_rows(int P, int ncols, const uint8_t *mask, const int32_t *gpos, const int32_t *colstart,
                            const int32_t *dst_off_host, const float *means2D, const float *rgb,
                            const float *conic_opacity, const int32_t *radii, const float *depths, float *out,
                            void *stream) {
    GS_REQUIRE(P >= 0 && ncols > 0 && ncols <= MAX_COLS, "sizes");
    if (P == 0) return GS_OK;
    GS_REQUIRE(mask && gpos && colstart && dst_off_host && means2D && rgb && conic_opacity && radii && depths && out,
               "null pointer");
    Cols off;
    for (int c = 0; c < MAX_COLS; c++) off.v[c] = c < ncols ? dst_off_host[c] : 0;
    GsStageTimer timer(GS_STAGE_PACK, (cudaStream_t)stream);
    k_pack_rows<<<(P + DT_THREADS - 1) / DT_THREADS, DT_THREADS, 0, (cudaStream_t)stream>>>(
        P, ncols, mask, gpos, colstart, off, means2D, rgb, conic_opacity, radii, depths, out);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

// Received rows of one camera = up to 16 segments of the recv buffer (one per source rank), concatenated.
struct Segs { int32_t off[MAX_COLS]; int32_t start[MAX_COLS + 1]; int n; };

__device__ __forceinline__ int seg_row(const Segs &s, int r) {
    int k = 0;
    while (k + 1 < s.n && r >= s.start[k + 1]) k++;
    return s.off[k] + (r - s.start[k]);
}

__global__ void __launch_bounds__(DT_THREADS)
k_unpack_rows(int n, Segs segs, const float *__restrict__ rows, float *__restrict__ means2D, float *__restrict__ rgb,
              float *__restrict__ conic_opacity, int32_t *__restrict__ radii, float *__restrict__ depths) {
    const int i = blockIdx.x * DT_THREADS + threadIdx.x;
    if (i >= n) return;
    const float *r = rows + (size_t)seg_row(segs, i) * ROW_FLOATS;
    means2D[2 * i] = r[0]; means2D[2 * i + 1] = r[1];
    rgb[3 * i] = r[2]; rgb[3 * i + 1] = r[3]; rgb[3 * i + 2] = r[4];
    *reinterpret_cast<float4 *>(conic_opacity + 4 * i) = make_float4(r[5], r[6], r[7], r[8]);
    radii[i] = (int32_t)r[9];
    depths[i] = r[10];
}

static int make_segs(int nseg, const int32_t *off_host, const int32_t *len_host, Segs &s) {
    GS_REQUIRE(nseg > 0 && nseg <= MAX_COLS && off_host && len_host, "segments");
    s.n = nseg;
    int acc = 0;
    for (int k = 0; k < MAX_COLS; k++) {
        s.off[k] = k < nseg ? off_host[k] : 0;
        s.start[k] = acc;
        if (k < nseg) acc += len_host[k];
    }
    s.start[MAX_COLS] = acc;
    return acc;
}

// seg_off_host / seg_len_host: HOST arrays (nseg): first row and row count of each segment in `rows`.
extern "C" int gs_unpack_rows(int nseg, const int32_t *seg_off_host, const int32_t *seg_len_host, const float *rows,
                              float *means2D, float *rgb, float *conic_opacity, int32_t *radii, float *depths,
                              void *stream) {
    Segs s;
    const int n = make_segs(nseg, seg_off_host, seg_len_host, s);
    if (n < 0) return n;
    if (n == 0) return GS_OK;
    GS_REQUIRE(rows && means2D && rgb && conic_opacity && radii && depths, "null pointer");
    GsStageTimer timer(GS_STAGE_UNPACK, (cudaStream_t)stream);
    k_unpack_rows<<<(n + DT_THREADS - 1) / DT_THREADS, DT_THREADS, 0, (cudaStream_t)stream>>>(n, s, rows, means2D, rgb,
                                                                                             conic_opacity, radii, depths);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

// backward of unpack: the camera's (n,2)(n,3)(n,4) gradients -> 9-float rows laid out like the recv buffer
__global__ void __launch_bounds__(DT_THREADS)
k_pack_grad_rows(int n, Segs segs, const float *__restrict__ d_means2D, const float *__restrict__ d_rgb,
                 const float *__restrict__ d_conic, float *__restrict__ rows) {
    const int i = blockIdx.x * DT_THREADS + threadIdx.x;
    if (i >= n) return;
    float *r = rows + (size_t)seg_row(segs, i) * GRAD_FLOATS;
    const float4 co = *reinterpret_cast<const float4 *>(d_conic + 4 * i);
    r[0] = d_means2D[2 * i]; r[1] = d_means2D[2 * i + 1];
    r[2] = d_rgb[3 * i]; r[3] = d_rgb[3 * i + 1]; r[4] = d_rgb[3 * i + 2];
    r[5] = co.x; r[6] = co.y; r[7] = co.z; r[8] = co.w;
}

extern "C" int gs_pack_grad_rows(int nseg, const int32_t *seg_off_host, const int32_t *seg_len_host,
                                 const float *d_means2D, const float *d_rgb, const float *d_conic_opacity, float *rows,
                                 void *stream) {
    Segs s;
    const int n = make_segs(nseg, seg_off_host, seg_len_host, s);
    if (n < 0) return n;
    if (n == 0) return GS_OK;
    GS_REQUIRE(rows && d_means2D && d_rgb && d_conic_opacity, "null pointer");
    GsStageTimer timer(GS_STAGE_UNPACK, (cudaStream_t)stream);
    k_pack_grad_rows<<<(n + DT_THREADS - 1) / DT_THREADS, DT_THREADS, 0, (cudaStream_t)stream>>>(n, s, d_means2D, d_rgb,
                                                                                                d_conic_opacity, rows);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

// backward of pack: every local splat sums the gradient rows that came back from the columns it was sent to
// (one thread per splat: no atomics; unflagged splats get exact zeros).
__global__ void __launch_bounds__(DT_THREADS)
k_scatter_grad_rows(int P, int ncols, const uint8_t *__restrict__ mask, const int32_t *__restrict__ gpos,
                    const int32_t *__restrict__ colstart, Cols src_off, const float *__restrict__ rows,
                    float *__restrict__ d_means2D, float *__restrict__ d_rgb, float *__restrict__ d_conic) {
    const int i = blockIdx.x * DT_THREADS + threadIdx.x;
    if (i >= P) return;
    float a[GRAD_FLOATS];
#pragma unroll
    for (int q = 0; q < GRAD_FLOATS; q++) a[q] = 0.f;
    for (int c = 0; c < ncols; c++) {
        if (!mask[(size_t)i * ncols + c]) continue;
        const float *r = rows + (size_t)(src_off.v[c] + gpos[(size_t)c * P + i] - colstart[c]) * GRAD_FLOATS;
#pragma unroll
        for (int q = 0; q < GRAD_FLOATS; q++) a[q] += r[q];
    }
    d_means2D[2 * i] = a[0]; d_means2D[2 * i + 1] = a[1];
    d_rgb[3 * i] = a[2]; d_rgb[3 * i + 1] = a[3]; d_rgb[3 * i + 2] = a[4];
    *reinterpret_cast<float4 *>(d_conic + 4 * i) = make_float4(a[5], a[6], a[7], a[8]);
}

extern "C" int gs_scatter_grad_rows(int P, int ncols, const uint8_t *mask, const int32_t *gpos, const int32_t *colstart,
                                    const int32_t *src_off_host, const float *rows, float *d_means2D, float *d_rgb,
                                    float *d_conic_opacity, void *stream) {
    GS_REQUIRE(P >= 0 && ncols > 0 && ncols <= MAX_COLS, "sizes");
    if (P == 0) return GS_OK;
    GS_REQUIRE(mask && gpos && colstart && src_off_host && d_means2D && d_rgb && d_conic_opacity, "null pointer");
    Cols off;
    for (int c = 0; c < MAX_COLS; c++) off.v[c] = c < ncols ? src_off_host[c] : 0;
    GsStageTimer timer(GS_STAGE_PACK, (cudaStream_t)stream);
    k_scatter_grad_rows<<<(P + DT_THREADS - 1) / DT_THREADS, DT_THREADS, 0, (cudaStream_t)stream>>>(
        P, ncols, mask, gpos, colstart, off, rows, d_means2D, d_rgb, d_conic_opacity);
    GS_LAUNCH_CHECK();
    return GS_OK;
}
