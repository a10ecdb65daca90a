// Tile-mask and tile gather/scatter helpers of the reference's LEGACY image-distribution path
// (SURVEY.md section 8a rows L3/L4 -- never called by the shipped trainer, provided so the drop-in package
// exports the extension's whole surface):
//   _C.get_touched_locally                       /root/reference/gaussian_renderer/loss_distribution.py:136-141
//   _C.get_pixels_compute_locally_and_in_rect    loss_distribution.py:205-213, 919-927, 1913-1921, 2135-2143
//   load_image_tiles_by_pos / merge_image_tiles_by_pos   loss_distribution.py:168-175, 188-195
// Byte / copy work, HBM bound, tiny.
#include "common.cuh"

#define LG_THREADS 256

// out[t] = some locally computed tile lies within `ext` tiles of tile t (the loss window is 11 < 16 pixels, so
// the reference always passes ext = 1)
__global__ void k_touched_locally(int ty, int tx, int ext, const uint8_t *__restrict__ cl, uint8_t *__restrict__ out) {
    const int t = blockIdx.x * LG_THREADS + threadIdx.x;
    if (t >= ty * tx) return;
    const int y = t / tx, x = t % tx;
    bool hit = false;
    for (int yy = max(0, y - ext); yy <= min(ty - 1, y + ext) && !hit; yy++)
        for (int xx = max(0, x - ext); xx <= min(tx - 1, x + ext); xx++) hit |= cl[yy * tx + xx] != 0;
    out[t] = hit ? 1 : 0;
}

extern "C" int gs_get_touched_locally(int tile_y, int tile_x, int extension_distance, const uint8_t *compute_locally,
                                      uint8_t *out, void *stream) {
    GS_REQUIRE(tile_y > 0 && tile_x > 0 && extension_distance >= 0 && compute_locally && out, "arguments");
    const int n = tile_y * tile_x;
    k_touched_locally<<<(n + LG_THREADS - 1) / LG_THREADS, LG_THREADS, 0, (cudaStream_t)stream>>>(
        tile_y, tile_x, extension_distance, compute_locally, out);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

// out[(y - min_y), (x - min_x)] = compute_locally[tile of pixel (y, x)], false outside the image
__global__ void k_pixels_in_rect(int H, int W, int tx, const uint8_t *__restrict__ cl, int min_y, int min_x, int rh, int rw,
                                 uint8_t *__restrict__ out) {
    const int k = blockIdx.x * LG_THREADS + threadIdx.x;
    if (k >= rh * rw) return;
    const int y = min_y + k / rw, x = min_x + k % rw;
    out[k] = (y >= 0 && y < H && x >= 0 && x < W && cl[(y / GS_BLOCK_Y) * tx + x / GS_BLOCK_X]) ? 1 : 0;
}

extern "C" int gs_get_pixels_compute_locally_and_in_rect(int image_height, int image_width, const uint8_t *compute_locally,
                                                         int min_y, int max_y, int min_x, int max_x, uint8_t *out,
                                                         void *stream) {
    GS_REQUIRE(image_height > 0 && image_width > 0 && max_y >= min_y && max_x >= min_x && compute_locally, "arguments");
    const int rh = max_y - min_y, rw = max_x - min_x;
    if (rh * rw == 0) return GS_OK;
    GS_REQUIRE(out != nullptr, "out");
    const int tx = (image_width + GS_BLOCK_X - 1) / GS_BLOCK_X;
    k_pixels_in_rect<<<(rh * rw + LG_THREADS - 1) / LG_THREADS, LG_THREADS, 0, (cudaStream_t)stream>>>(
        image_height, image_width, tx, compute_locally, min_y, min_x, rh, rw, out);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

// tiles[n][c][16][16] <-> image_rect[c][rh][rw]; pos = (n,2) int64 GLOBAL tile (y, x); the rect's origin is the
// pixel (rect_min_y, rect_min_x).  GATHER reads the rect (0 outside it / outside the image); the other direction
// accumulates with atomicAdd (it is also the adjoint of GATHER).
template <bool GATHER>
__global__ void k_tiles(int n, const int64_t *__restrict__ pos, float *__restrict__ rect, int rh, int rw, int rect_min_y,
                        int rect_min_x, int H, int W, float *__restrict__ tiles) {
    const int t = blockIdx.x;  // one CTA (256 threads = 16x16 pixels) per tile, loop over channels
    if (t >= n) return;
    const int ly = threadIdx.x / GS_BLOCK_X, lx = threadIdx.x % GS_BLOCK_X;
    const int gy = (int)pos[2 * t] * GS_BLOCK_Y + ly, gxp = (int)pos[2 * t + 1] * GS_BLOCK_X + lx;
    const int y = gy - rect_min_y, x = gxp - rect_min_x;
    const bool ok = gy >= 0 && gy < H && gxp >= 0 && gxp < W && y >= 0 && y < rh && x >= 0 && x < rw;
    for (int c = 0; c < 3; c++) {
        float *tp = tiles + (((size_t)t * 3 + c) * GS_BLOCK_Y + ly) * GS_BLOCK_X + lx;
        float *ip = rect + ((size_t)c * rh + (ok ? y : 0)) * rw + (ok ? x : 0);
        if (GATHER) *tp = ok ? *ip : 0.f;
        else if (ok) atomicAdd(ip, *tp);
    }
}

extern "C" int gs_image_tiles_gather(int n, const int64_t *pos, const float *image_rect, int rect_h, int rect_w,
                                     int rect_min_y, int rect_min_x, int image_height, int image_width, float *tiles,
                                     void *stream) {
    GS_REQUIRE(n >= 0 && rect_h >= 0 && rect_w >= 0, "sizes");
    if (n == 0) return GS_OK;
    GS_REQUIRE(pos && image_rect && tiles, "null pointer");
    k_tiles<true><<<n, GS_BLOCK_X * GS_BLOCK_Y, 0, (cudaStream_t)stream>>>(n, pos, const_cast<float *>(image_rect), rect_h,
                                                                           rect_w, rect_min_y, rect_min_x, image_height,
                                                                           image_width, tiles);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

// image_rect must be zero-initialised (or hold values to accumulate onto) by the caller.
extern "C" int gs_image_tiles_scatter_add(int n, const int64_t *pos, const float *tiles, int rect_h, int rect_w,
                                          int rect_min_y, int rect_min_x, int image_height, int image_width,
                                          float *image_rect, void *stream) {
    GS_REQUIRE(n >= 0 && rect_h >= 0 && rect_w >= 0, "sizes");
    if (n == 0) return GS_OK;
    GS_REQUIRE(pos && image_rect && tiles, "null pointer");
    k_tiles<false><<<n, GS_BLOCK_X * GS_BLOCK_Y, 0, (cudaStream_t)stream>>>(n, pos, image_rect, rect_h, rect_w, rect_min_y,
                                                                            rect_min_x, image_height, image_width,
                                                                            const_cast<float *>(tiles));
    GS_LAUNCH_CHECK();
    return GS_OK;
}
