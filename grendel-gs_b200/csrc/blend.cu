// Per-tile alpha blending: CUDA stages "70 render" (+81-83 statistics) and "b10 render"
// (/root/reference/analyze_statistic.py:1981-1987) -- the second half of
// GaussianRasterizer.render_gaussians and its autograd backward
// (/root/reference/gaussian_renderer/__init__.py:1271-1282, train_internal.py:195).
//
// One CTA = one 16x16 tile (BLOCK_X/Y are observable through _C.get_block_XY and baked into the
// reference's strip arithmetic, loss_distribution.py:2321-2330).  Both kernels are bound by instruction
// issue over (pixel, splat) pairs, not by HBM (ncu: issue slots ~85 % busy, DRAM ~2 %; algorithmic
// traffic is only 40 B / 76 B per instance), so the design minimises issued instructions
// per USEFUL pair:
//   * the tile is cut into sixteen 4x4 pixel blocks; every HALF-warp owns one block and walks its own
//     candidate list, so one warp instruction advances two (block, splat) pairs and ~60 % of the lanes do
//     useful work in the blend path (an 8x4 block per full warp reached 37 %);
//   * candidates come from a 16-bit "which 4x4 blocks can this splat reach" mask computed at staging time
//     from the bounding box of the splat's {alpha >= 1/255} ellipse (conservative: per-pixel results are
//     unchanged); a warp ballots 32 masks at a time;
//   * the alpha < 1/255 test is done on the exponent (power < ln(1/(255 o)) - margin) so rejected pairs never
//     reach MUFU.EX2; finished / out-of-image pixels carry NaN coordinates so they fail it for free;
//   * sorted splat ids are turned into 48-byte packed records (3 x float4, built by k_count_tiles), gathered
//     once per (splat, tile) into shared memory;
//   * backward (one 8x4 block per full warp, see the note above k_blend_bwd): per-pixel weight m = dL/dG * G; its
//     six moments (sum m, m dx, m dy, m dx^2, m dx dy, m dy^2) and three colour sums are reduced over the warp by a
//     9-value transposing butterfly (14 shuffles instead of 45), stored in per-warp private shared-memory slots
//     (no shared atomics: sm_100 has no native fp32 ATOMS.ADD -- it compiles to a CAS loop), then ONE thread per
//     splat sums the warps, applies the splat's constants and issues ONE set of 9 global RED.ADD per
//     (splat, tile) -- instead of 9 atomics per (splat, pixel) in the classical design.  The per-pixel state is
//     just (T, behind-colour B): lanes that skip a splat run the same instructions with alpha = 0, which makes
//     every update a no-op, so the blend path has no per-lane branches or conditional moves.
#include "common.cuh"

#define BL_THREADS 256
#define BL_WARPS (BL_THREADS / 32)
#define BL_BLOCKS 16  // 4x4 pixel blocks per tile
#define FW_CHUNK 256

#define ALPHA_MIN (1.0f / 255.0f)
#define ALPHA_MAX 0.99f
#define T_EPS 0.0001f
#define FULL 0xffffffffu

// exp(x) for x <= 0 as one FMUL + MUFU.EX2 (results below 2^-126 flush to 0: far under the 1/255 alpha floor)
GS_D float gs_exp_neg(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}

// thread -> (4x4 block, pixel).  Warp w holds blocks 2w and 2w+1 (horizontal neighbours), one per half-warp.
struct Where { int blk, px, py, half, l16; };
GS_D Where where_am_i(int tile, int gx) {
    Where p;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    p.half = lane >> 4;
    p.l16 = lane & 15;
    const int by = w >> 1, bx = (w & 1) * 2 + p.half;
    p.blk = by * 4 + bx;
    p.px = (tile % gx) * GS_BLOCK_X + bx * 4 + (p.l16 & 3);
    p.py = (tile / gx) * GS_BLOCK_Y + by * 4 + (p.l16 >> 2);
    return p;
}

// bit (by*4+bx) set <=> the splat's bounding box (centre m, half extents e) overlaps 4x4 block (bx,by)
GS_D uint32_t block_mask16(float mx, float my, float ex, float ey, float X0, float Y0) {
    if (ex < 0.f) return 0u;
    const float xl = mx - ex - X0, xh = mx + ex - X0, yl = my - ey - Y0, yh = my + ey - Y0;
    uint32_t xm = 0u, m = 0u;
#pragma unroll
    for (int b = 0; b < 4; b++)
        if (xh >= 4.f * b && xl <= 4.f * b + 3.f) xm |= 1u << b;
#pragma unroll
    for (int b = 0; b < 4; b++)
        if (yh >= 4.f * b && yl <= 4.f * b + 3.f) m |= xm << (4 * b);
    return m;
}

// staged splat: one 48-byte slot so a single address feeds all three shared-memory loads
struct __align__(16) SRec { float4 a; float4 b; float4 c; };

template <bool STATS>
__global__ void __launch_bounds__(BL_THREADS)
k_blend_fwd(int W, int H, int tiles_per_view, const float4 *__restrict__ rec, const float *__restrict__ bg,
            const uint8_t *__restrict__ compute_locally, const uint2 *__restrict__ ranges,
            const uint32_t *__restrict__ ids, float *__restrict__ image, float *__restrict__ final_T,
            uint32_t *__restrict__ n_contrib, unsigned long long *__restrict__ stats) {
    __shared__ SRec s_rec[FW_CHUNK];
    __shared__ uint16_t s_cull[FW_CHUNK];
    __shared__ unsigned long long s_stats[3];
    const int gx = (W + GS_BLOCK_X - 1) / GS_BLOCK_X;
    // blockIdx.x = view * tiles_per_view + tile: the views' tile arrays and images are concatenated (GsViews)
    const int view = blockIdx.x / tiles_per_view, tile = blockIdx.x - view * tiles_per_view;
    const int lane = threadIdx.x & 31;
    const Where me = where_am_i(tile, gx);
    const int px = me.px, py = me.py;
    const bool inside = px < W && py < H;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)py * W + px;
    image += (size_t)view * 3 * HW;
    final_T += (size_t)view * HW;
    n_contrib += (size_t)view * HW;
    if (!compute_locally[blockIdx.x]) {  // non-local tiles must read exactly 0 (loss_distribution.py:1875)
        if (inside) { image[pix] = 0.f; image[HW + pix] = 0.f; image[2 * HW + pix] = 0.f; }
        return;
    }
    const uint2 range = ranges[blockIdx.x];
    const int total = (int)(range.y - range.x);
    const float X0 = (float)((tile % gx) * GS_BLOCK_X), Y0 = (float)((tile / gx) * GS_BLOCK_Y);
    const float qnan = __int_as_float(0x7fc00000);
    float pxf = inside ? (float)px : qnan, pyf = (float)py;  // NaN coordinates: the pixel never passes a test
    float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t last = 0, blended = 0, considered = 0;
    bool done = !inside;
    const int blkA = me.blk - me.half;  // block of lanes 0-15; lanes 16-31 own blkA + 1
    for (int base = 0; base < total; base += FW_CHUNK) {
        if (__syncthreads_count(done) == BL_THREADS) break;
        const int cnt = min(FW_CHUNK, total - base);
        if ((int)threadIdx.x < cnt) {
            const uint32_t g = ids[range.x + base + threadIdx.x];
            const float4 *r = rec + (size_t)3 * g;
            const float4 a = __ldg(r), b = __ldg(r + 1), c = __ldg(r + 2);
            s_rec[threadIdx.x].a = a; s_rec[threadIdx.x].b = b; s_rec[threadIdx.x].c = c;
            s_cull[threadIdx.x] = (uint16_t)block_mask16(a.x, a.y, c.z, c.w, X0, Y0);
        }
        __syncthreads();
        for (int g0 = 0; g0 < cnt; g0 += 32) {
            if (__all_sync(FULL, done)) break;  // checked once per 32 entries; finished pixels are NaN anyway
            const int jj = g0 + lane;
            const uint32_t m = jj < cnt ? (uint32_t)s_cull[jj] : 0u;
            const uint32_t cA = __ballot_sync(FULL, (m >> blkA) & 1u), cB = __ballot_sync(FULL, (m >> (blkA + 1)) & 1u);
            uint32_t mine = me.half ? cB : cA;  // this half-warp's candidates among the 32 entries
            while (__any_sync(FULL, mine != 0u)) {
                const bool has = mine != 0u;
                const int j = g0 + (has ? __ffs(mine) - 1 : 0);
                mine &= mine - 1u;
                const SRec *sr = &s_rec[j];
                const float4 a = sr->a, b = sr->b;
                const float dx = a.x - pxf, dy = a.y - pyf;
                const float power = dx * (a.z * dx + a.w * dy) + b.x * dy * dy;
                const bool ok = has && power >= b.z;  // false for NaN (finished / outside pixels)
                if (!__any_sync(FULL, ok)) continue;
                if (ok && power <= 0.f) {
                    const float alpha = fminf(ALPHA_MAX, b.y * gs_exp_neg(power));
                    if (alpha >= ALPHA_MIN) {
                        const float test_T = T * (1.f - alpha);
                        if (test_T < T_EPS) {
                            done = true;
                            pxf = qnan;
                            if (STATS) considered = (uint32_t)(base + j + 1);
                        } else {
                            const float2 gb = *reinterpret_cast<const float2 *>(&sr->c);
                            const float w = alpha * T;
                            C0 += b.w * w; C1 += gb.x * w; C2 += gb.y * w;
                            T = test_T;
                            last = (uint32_t)(base + j + 1);
                            if (STATS) blended++;
                        }
                    }
                }
            }
        }
    }
    if (inside) {
        image[pix] = C0 + T * bg[0];
        image[HW + pix] = C1 + T * bg[1];
        image[2 * HW + pix] = C2 + T * bg[2];
        final_T[pix] = T;
        n_contrib[pix] = last;
        if (considered == 0) considered = (uint32_t)total;
    }
    if (STATS) {  // stages 81-83: sums of tile-list length / entries walked / entries blended
        if (threadIdx.x < 3) s_stats[threadIdx.x] = 0ull;
        __syncthreads();
        unsigned long long v0 = inside ? (unsigned long long)total : 0ull, v1 = inside ? considered : 0u,
                           v2 = inside ? blended : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_xor_sync(FULL, v0, o);
            v1 += __shfl_xor_sync(FULL, v1, o);
            v2 += __shfl_xor_sync(FULL, v2, o);
        }
        if (lane == 0) { atomicAdd(&s_stats[0], v0); atomicAdd(&s_stats[1], v1); atomicAdd(&s_stats[2], v2); }
        __syncthreads();
        if (threadIdx.x < 3) atomicAdd(&stats[3 * view + threadIdx.x], s_stats[threadIdx.x]);
    }
}

// ---- backward -------------------------------------------------------------------------------------------
// The backward keeps one 8x4 pixel block per FULL warp (8-bit cull mask, 128-entry chunks): the half-warp /
// 4x4 layout that helps the forward was measured slower here (2.18 vs 1.77 ms on c2) -- sixteen private
// partial-sum slots force 64-entry chunks, and the extra barriers + 64-thread flush cost more (barrier stalls
// 3.1 warps per issue) than the better lane utilisation returns.
#define BW8_CHUNK 128
#define BW8_STRIDE (BW8_CHUNK + 1)

GS_D void pixel_of_thread(int tile, int gx, int &px, int &py) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    px = (tile % gx) * GS_BLOCK_X + (w & 1) * 8 + (lane & 7);
    py = (tile / gx) * GS_BLOCK_Y + (w >> 1) * 4 + (lane >> 3);
}

// bit w set <=> the splat's bounding box (centre m, half extents e) overlaps warp w's 8x4 pixel block
GS_D uint32_t block_mask(float mx, float my, float ex, float ey, float X0, float Y0) {
    if (ex < 0.f) return 0u;
    const float xl = mx - ex - X0, xh = mx + ex - X0, yl = my - ey - Y0, yh = my + ey - Y0;
    const uint32_t xm = ((xh >= 0.f && xl <= 7.f) ? 1u : 0u) | ((xh >= 8.f && xl <= 15.f) ? 2u : 0u);
    uint32_t m = 0u;
#pragma unroll
    for (int wy = 0; wy < 4; wy++)
        if (yh >= 4.f * wy && yl <= 4.f * wy + 3.f) m |= xm << (2 * wy);
    return m;
}

// 9-value warp reduction.  After the call every lane holds in v[0] the warp total of value
// (lane >> 2) & 7, and in v[8] the warp total of value 8.
GS_D void warp_reduce9(float v[9], int lane) {
    {
        const bool h = lane & 16;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float send = h ? v[i] : v[i + 4], keep = h ? v[i + 4] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool h = lane & 8;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const float send = h ? v[i] : v[i + 2], keep = h ? v[i + 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool h = lane & 4;
        const float send = h ? v[0] : v[1], keep = h ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[8] += __shfl_xor_sync(0xffffffffu, v[8], o);
}

__global__ void __launch_bounds__(BL_THREADS)
k_blend_bwd(int W, int H, int tiles_per_view, const float4 *__restrict__ rec, const float *__restrict__ bg,
            const uint8_t *__restrict__ compute_locally, const uint2 *__restrict__ ranges,
            const uint32_t *__restrict__ ids, const float *__restrict__ final_T,
            const uint32_t *__restrict__ n_contrib, const float *__restrict__ dL_dimage,
            float *__restrict__ d_means2D, float *__restrict__ d_conic_opacity, float *__restrict__ d_rgb) {
    __shared__ SRec s_rec[BW8_CHUNK];
    __shared__ uint32_t s_id[BW8_CHUNK];
    __shared__ uint8_t s_cull[BW8_CHUNK];
    __shared__ float s_acc[BL_WARPS][9][BW8_STRIDE];
    __shared__ uint32_t s_mask[BL_WARPS][BW8_CHUNK / 32];
    __shared__ uint32_t s_max[BL_WARPS];
    const int gx = (W + GS_BLOCK_X - 1) / GS_BLOCK_X;
    if (!compute_locally[blockIdx.x]) return;
    const int view = blockIdx.x / tiles_per_view, tile = blockIdx.x - view * tiles_per_view;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int px, py;
    pixel_of_thread(tile, gx, px, py);
    const bool inside = px < W && py < H;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)py * W + px;
    final_T += (size_t)view * HW;
    n_contrib += (size_t)view * HW;
    dL_dimage += (size_t)view * 3 * HW;
    const uint2 range = ranges[blockIdx.x];
    const float X0 = (float)((tile % gx) * GS_BLOCK_X), Y0 = (float)((tile / gx) * GS_BLOCK_Y);
    const float pxf = (float)px, pyf = (float)py;
    const float T_final = inside ? final_T[pix] : 0.f;
    const uint32_t last = inside ? n_contrib[pix] : 0u;
    float dp0 = 0.f, dp1 = 0.f, dp2 = 0.f;
    if (inside) { dp0 = dL_dimage[pix]; dp1 = dL_dimage[HW + pix]; dp2 = dL_dimage[2 * HW + pix]; }
    const float bgdot = bg[0] * dp0 + bg[1] * dp1 + bg[2] * dp2;
    const float ddelx_dx = 0.5f * (float)W, ddely_dy = 0.5f * (float)H;
    // entries past the tile's deepest last-contributor cannot matter
    uint32_t m = last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) s_max[warp] = m;
    __syncthreads();
    uint32_t n_total = 0;
#pragma unroll
    for (int w = 0; w < BL_WARPS; w++) n_total = max(n_total, s_max[w]);
    const uint32_t wlast = m;  // deepest entry this warp's pixels reach
    // B = colour of everything behind the current splat (back-to-front recurrence B += alpha (c - B)); with alpha
    // forced to 0 for lanes that do not blend this splat every state update below is a no-op for them, so the
    // blend path needs no per-lane branches or conditional moves.
    float T = T_final, B0 = 0.f, B1 = 0.f, B2 = 0.f;
    const int n_chunks = ((int)n_total + BW8_CHUNK - 1) / BW8_CHUNK;
    for (int c = n_chunks - 1; c >= 0; c--) {
        const int base = c * BW8_CHUNK;
        const int cnt = min(BW8_CHUNK, (int)n_total - base);
        __syncthreads();  // previous chunk's flush has finished reading shared memory
        if ((int)threadIdx.x < cnt) {
            const uint32_t g = ids[range.x + base + threadIdx.x];
            s_id[threadIdx.x] = g;
            const float4 *r = rec + (size_t)3 * g;
            const float4 a = __ldg(r), b = __ldg(r + 1), cc = __ldg(r + 2);
            s_rec[threadIdx.x].a = a; s_rec[threadIdx.x].b = b; s_rec[threadIdx.x].c = cc;
            s_cull[threadIdx.x] = (uint8_t)block_mask(a.x, a.y, cc.z, cc.w, X0, Y0);
        }
        uint32_t wmask = 0u;  // lane q holds bits [32q, 32q+32) of "this warp produced a partial for entry j"
        __syncthreads();
        if ((uint32_t)base < wlast) {
            for (int g0 = (cnt - 1) & ~31; g0 >= 0; g0 -= 32) {
                // lane l inspects entry g0 + 31 - l, so the LOWEST set bit of the ballot is the DEEPEST candidate and
                // the walk (back to front) pops bits with the cheap x & (x - 1)
                const int jj = g0 + 31 - lane;
                uint32_t cand = __ballot_sync(0xffffffffu, jj < cnt && ((s_cull[jj] >> warp) & 1));
                uint32_t mybits = 0u;
                const int last_rel = (int)last - base;  // entry j of this chunk is live for this pixel iff j < last_rel
                while (cand) {
                    const int b31 = 32 - __ffs(cand);   // = 31 - (index of the lowest set bit)
                    const int j = g0 + b31;
                    cand &= cand - 1u;
                    const SRec *sr = &s_rec[j];
                    const float4 a = sr->a, b = sr->b;
                    const float dx = a.x - pxf, dy = a.y - pyf;
                    const float power = dx * (a.z * dx + a.w * dy) + b.x * dy * dy;
                    bool ok = (j < last_rel) && power >= b.z;
                    if (!__any_sync(0xffffffffu, ok)) continue;
                    const float G = gs_exp_neg(power);
                    const float alpha = fminf(ALPHA_MAX, b.y * G);
                    ok = ok && power <= 0.f && alpha >= ALPHA_MIN;
                    if (!__any_sync(0xffffffffu, ok)) continue;
                    // Per-pixel weight m = dL/dG * G; the per-splat gradients are its moments over the pixels
                    // (S0, Sx, Sy, Sxx, Sxy, Syy) plus three colour sums; they are combined with the splat's
                    // constants once per (splat, tile) in the flush below.
                    float v[9];
                    {
                        const float2 gb = *reinterpret_cast<const float2 *>(&sr->c);
                        const float ae = ok ? alpha : 0.f;   // effective alpha: 0 = this lane skips the splat
                        float inv;                            // 1/(1-ae), 1-ae in [0.01, 1]: one MUFU.RCP
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(1.f - ae));
                        T = T * inv;                          // transmittance in front of this splat
                        const float d0 = b.w - B0, d1 = gb.x - B1, d2 = gb.y - B2;
                        const float dL_dalpha = (d0 * dp0 + d1 * dp1 + d2 * dp2) * T - (T_final * inv) * bgdot;
                        const float mw = ok ? b.y * dL_dalpha * G : 0.f;
                        const float dch = ae * T;
                        B0 += ae * d0; B1 += ae * d1; B2 += ae * d2;
                        const float mx_ = mw * dx, my_ = mw * dy;
                        v[0] = mx_; v[1] = my_; v[2] = mx_ * dx; v[3] = mx_ * dy; v[4] = my_ * dy; v[5] = mw;
                        v[6] = dch * dp0; v[7] = dch * dp1; v[8] = dch * dp2;
                    }
                    warp_reduce9(v, lane);
                    if ((lane & 3) == 0) s_acc[warp][lane >> 2][j] = v[0];
                    if (lane == 1) s_acc[warp][8][j] = v[8];
                    mybits |= 1u << b31;
                }
                if (lane == (g0 >> 5)) wmask = mybits;
            }
        }
        if (lane < BW8_CHUNK / 32) s_mask[warp][lane] = wmask;
        __syncthreads();
        if ((int)threadIdx.x < cnt) {
            const int j = threadIdx.x;
            float s[9];
#pragma unroll
            for (int q = 0; q < 9; q++) s[q] = 0.f;
            bool any = false;
#pragma unroll
            for (int w = 0; w < BL_WARPS; w++) {
                if ((s_mask[w][j >> 5] >> (j & 31)) & 1u) {
                    any = true;
#pragma unroll
                    for (int q = 0; q < 9; q++) s[q] += s_acc[w][q][j];
                }
            }
            if (any) {
                const uint32_t g = s_id[j];
                const float4 a = s_rec[j].a, b = s_rec[j].b;  // (mx,my,a',b') (c',opacity,thr,red); A=-2a' B=-b' C=-2c'
                // d power/d mean = (2a'dx + b'dy, 2c'dy + b'dx); dL/dmeans2D is per NDC unit: * (W/2, H/2)
                atomicAdd(d_means2D + 2 * (size_t)g, (2.f * a.z * s[0] + a.w * s[1]) * ddelx_dx);
                atomicAdd(d_means2D + 2 * (size_t)g + 1, (2.f * b.x * s[1] + a.w * s[0]) * ddely_dy);
                atomicAdd(d_conic_opacity + 4 * (size_t)g, -0.5f * s[2]);
                atomicAdd(d_conic_opacity + 4 * (size_t)g + 1, -s[3]);
                atomicAdd(d_conic_opacity + 4 * (size_t)g + 2, -0.5f * s[4]);
                atomicAdd(d_conic_opacity + 4 * (size_t)g + 3, __fdividef(s[5], b.y));
                atomicAdd(d_rgb + 3 * (size_t)g, s[6]);
                atomicAdd(d_rgb + 3 * (size_t)g + 1, s[7]);
                atomicAdd(d_rgb + 3 * (size_t)g + 2, s[8]);
            }
        }
    }
}

int gs_launch_blend_forward(int num_views, int64_t R, int H, int W, const float *rec, const float *bg,
                            const uint8_t *compute_locally, const uint32_t *ranges, const uint32_t *ids_sorted,
                            float *image, float *final_T, uint32_t *n_contrib, int64_t *stats, cudaStream_t stream) {
    (void)R;
    const int gx = (W + GS_BLOCK_X - 1) / GS_BLOCK_X, gy = (H + GS_BLOCK_Y - 1) / GS_BLOCK_Y;
    const int T1 = gx * gy;
    if (stats) GS_CUDA_TRY(cudaMemsetAsync(stats, 0, 3 * sizeof(int64_t) * (size_t)num_views, stream));
    GsStageTimer timer(GS_STAGE_BLEND_FWD, stream);
    auto *st = reinterpret_cast<unsigned long long *>(stats);
    if (stats)
        k_blend_fwd<true><<<T1 * num_views, BL_THREADS, 0, stream>>>(
            W, H, T1, reinterpret_cast<const float4 *>(rec), bg, compute_locally, reinterpret_cast<const uint2 *>(ranges),
            ids_sorted, image, final_T, n_contrib, st);
    else
        k_blend_fwd<false><<<T1 * num_views, BL_THREADS, 0, stream>>>(
            W, H, T1, reinterpret_cast<const float4 *>(rec), bg, compute_locally, reinterpret_cast<const uint2 *>(ranges),
            ids_sorted, image, final_T, n_contrib, st);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

extern "C" int gs_render_backward(int P, int64_t R, int image_height, int image_width, const float *rec,
                                  const float *bg, const uint8_t *compute_locally, const uint32_t *ranges,
                                  const uint32_t *ids_sorted, const float *final_T, const uint32_t *n_contrib,
                                  const float *dL_dimage, float *dL_dmeans2D, float *dL_dconic_opacity, float *dL_drgb,
                                  void *stream) {
    return gs_render_backward_batched(1, P, R, image_height, image_width, rec, bg, compute_locally, ranges, ids_sorted,
                                      final_T, n_contrib, dL_dimage, dL_dmeans2D, dL_dconic_opacity, dL_drgb, stream);
}

extern "C" int gs_render_backward_batched(int num_views, int P, int64_t R, int image_height, int image_width,
                                          const float *rec, const float *bg, const uint8_t *compute_locally,
                                          const uint32_t *ranges, const uint32_t *ids_sorted, const float *final_T,
                                          const uint32_t *n_contrib, const float *dL_dimage, float *dL_dmeans2D,
                                          float *dL_dconic_opacity, float *dL_drgb, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GS_REQUIRE(num_views >= 1 && num_views <= GS_MAX_VIEWS, "num_views must be in [1, GS_MAX_VIEWS]");
    GS_REQUIRE(P >= 0 && R >= 0 && image_height > 0 && image_width > 0, "sizes");
    if (P == 0) return GS_OK;
    GS_REQUIRE(dL_dmeans2D && dL_dconic_opacity && dL_drgb, "null output");
    GS_CUDA_TRY(cudaMemsetAsync(dL_dmeans2D, 0, sizeof(float) * 2 * (size_t)P, stream));
    GS_CUDA_TRY(cudaMemsetAsync(dL_dconic_opacity, 0, sizeof(float) * 4 * (size_t)P, stream));
    GS_CUDA_TRY(cudaMemsetAsync(dL_drgb, 0, sizeof(float) * 3 * (size_t)P, stream));
    if (R == 0) return GS_OK;
    GS_REQUIRE(rec && bg && compute_locally && ranges && ids_sorted && final_T && n_contrib && dL_dimage, "null input");
    const int gx = (image_width + GS_BLOCK_X - 1) / GS_BLOCK_X, gy = (image_height + GS_BLOCK_Y - 1) / GS_BLOCK_Y;
    GsStageTimer timer(GS_STAGE_BLEND_BWD, stream);
    k_blend_bwd<<<gx * gy * num_views, BL_THREADS, 0, stream>>>(
        image_width, image_height, gx * gy, reinterpret_cast<const float4 *>(rec), bg, compute_locally,
        reinterpret_cast<const uint2 *>(ranges), ids_sorted, final_T, n_contrib, dL_dimage, dL_dmeans2D, dL_dconic_opacity,
        dL_drgb);
    GS_LAUNCH_CHECK();
    return GS_OK;
}
