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

// ---- backward, EXPERIMENTAL variant (off by default; gs_debug_set(GS_DEBUG_BWD_WHT_64 / _128)) ---------------------
// Same walk, same per-pixel arithmetic; only the lane -> warp reduction of the six moments differs.  SASS of
// k_blend_bwd: 110 instructions per fully processed (warp, splat) pair, 51 of them forming and reducing the nine sums
// (14 SHFL + 14 FSEL + 14 FADD + 9 FMUL).  Here the per-pixel weight m alone goes through a 5-stage Walsh-Hadamard
// butterfly (5 SHFL + 5 FFMA): afterwards lane L holds W_L = sum_l m_l (-1)^popc(L & l).  A pixel's offset inside the
// warp's 8x4 block is a polynomial of degree <= 2 in the lane bits (lx = b0 + 2 b1 + 4 b2, ly = b3 + 2 b4), so all six
// moments about the block origin are linear combinations of the 16 coefficients with popc(L) <= 2:
//     sum m b_i     = (W_0 - W_{e_i}) / 2          sum m b_i b_j = (W_0 - W_{e_i} - W_{e_j} + W_{e_i + e_j}) / 4
// The 16 lanes that hold them store them (one STS), the flush thread of the splat forms the moments and shifts them to
// the splat centre.  The three colour sums take a 4-value transposing butterfly (6 SHFL).  92-98 instead of 111
// instructions per pair (profiles/r1_sass_blend_bwd.md), at the price of 19 instead of 9 floats of shared memory per
// (warp, entry) -- hence the CHUNK parameter: 64 entries fit 4 CTAs per SM, 128 entries halve the barriers but leave 2
// CTAs per SM.  To be timed on the device.
#define BWH_NVAL 19

template <int CHUNK>
// 44 KB / 83 KB of shared memory per CTA allow 5 / 2 CTAs per SM; the 64-entry variant is built for 4: squeezing it into
// 48 registers for 5 costs 23 more instructions per pair (114 vs 91) in rematerialised addresses
__global__ void __launch_bounds__(BL_THREADS, CHUNK == 64 ? 4 : 2)
k_blend_bwd_wht(int W, int H, int tiles_per_view, const float4 *__restrict__ rec, const float *__restrict__ bg,
                const uint8_t *__restrict__ compute_locally, const uint2 *__restrict__ ranges,
                const uint32_t *__restrict__ ids, const float *__restrict__ final_T,
                const uint32_t *__restrict__ n_contrib, const float *__restrict__ dL_dimage,
                float *__restrict__ d_means2D, float *__restrict__ d_conic_opacity, float *__restrict__ d_rgb) {
    constexpr int STRIDE = CHUNK + 1;
    extern __shared__ float s_acc[];  // [BL_WARPS][BWH_NVAL][STRIDE]
    __shared__ SRec s_rec[CHUNK];
    __shared__ uint32_t s_id[CHUNK];
    __shared__ uint8_t s_cull[CHUNK];
    __shared__ uint32_t s_mask[BL_WARPS][CHUNK / 32];
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
    // where this lane's Walsh-Hadamard coefficient goes: 0 <- W_0, 1+i <- W_{e_i}, 6+pair(i,j) <- W_{e_i+e_j}
    int slot = -1;
    {
        const int pc = __popc(lane);
        if (pc == 0) slot = 0;
        else if (pc == 1) slot = __ffs(lane);  // 1 + bit index
        else if (pc == 2) {
            const int i = __ffs(lane) - 1, j = 31 - __clz(lane);
            slot = 6 + (i * (9 - i)) / 2 + (j - i - 1);
        }
    }
    // sign of this lane in butterfly stage s: -1 if bit s of the lane is set
    float sg[5];
#pragma unroll
    for (int s = 0; s < 5; s++) sg[s] = ((lane >> s) & 1) ? -1.f : 1.f;
    const bool h16 = lane & 16, h8 = lane & 8;
    uint32_t m = last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) s_max[warp] = m;
    __syncthreads();
    uint32_t n_total = 0;
#pragma unroll
    for (int w = 0; w < BL_WARPS; w++) n_total = max(n_total, s_max[w]);
    const uint32_t wlast = m;
    float *acc_w = s_acc + (size_t)warp * BWH_NVAL * STRIDE;
    // shared-window addresses of this lane's two store rows, formed once (entry j adds 4 j bytes)
    const uint32_t st_w = gs_smem_u32(acc_w + (slot >= 0 ? slot : 0) * STRIDE);
    const uint32_t st_q = gs_smem_u32(acc_w + (16 + ((lane >> 3) & 3)) * STRIDE);
    const bool has_w = slot >= 0, has_q = (lane & 7) == 0 && lane < 24;  // colour totals end in lanes 0, 8, 16
    float T = T_final, B0 = 0.f, B1 = 0.f, B2 = 0.f;
    const int n_chunks = ((int)n_total + CHUNK - 1) / CHUNK;
    for (int c = n_chunks - 1; c >= 0; c--) {
        const int base = c * CHUNK;
        const int cnt = min(CHUNK, (int)n_total - base);
        __syncthreads();  // previous chunk's flush has finished reading shared memory
        if ((int)threadIdx.x < cnt) {
            const uint32_t g = ids[range.x + base + threadIdx.x];
            s_id[threadIdx.x] = g;
            const float4 *r = rec + (size_t)3 * g;
            const float4 a = __ldg(r), b = __ldg(r + 1), cc = __ldg(r + 2);
            s_rec[threadIdx.x].a = a; s_rec[threadIdx.x].b = b; s_rec[threadIdx.x].c = cc;
            s_cull[threadIdx.x] = (uint8_t)block_mask(a.x, a.y, cc.z, cc.w, X0, Y0);
        }
        uint32_t wmask = 0u;
        __syncthreads();
        if ((uint32_t)base < wlast) {
            for (int g0 = (cnt - 1) & ~31; g0 >= 0; g0 -= 32) {
                const int jj = g0 + 31 - lane;
                uint32_t cand = __ballot_sync(0xffffffffu, jj < cnt && ((s_cull[jj] >> warp) & 1));
                uint32_t mybits = 0u;
                const int last_rel = (int)last - base;
                while (cand) {
                    const int b31 = 32 - __ffs(cand);
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
                    const float2 gb = *reinterpret_cast<const float2 *>(&sr->c);
                    const float ae = ok ? alpha : 0.f;
                    float inv;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(1.f - ae));
                    T = T * inv;
                    const float d0 = b.w - B0, d1 = gb.x - B1, d2 = gb.y - B2;
                    const float dL_dalpha = (d0 * dp0 + d1 * dp1 + d2 * dp2) * T - (T_final * inv) * bgdot;
                    float w = ok ? b.y * dL_dalpha * G : 0.f;   // per-pixel weight m = dL/dG * G
                    const float dch = ae * T;
                    B0 += ae * d0; B1 += ae * d1; B2 += ae * d2;
                    // Walsh-Hadamard butterfly of m over the 32 lanes
#pragma unroll
                    for (int s = 0; s < 5; s++) w = __fmaf_rn(sg[s], w, __shfl_xor_sync(0xffffffffu, w, 1 << s));
                    // colour sums: 4-value transposing butterfly of (c0, c1, c2, 0)
                    const float c0 = dch * dp0, c1 = dch * dp1, c2 = dch * dp2;
                    float r0 = (h16 ? c2 : c0) + __shfl_xor_sync(0xffffffffu, h16 ? c0 : c2, 16);
                    float r1 = (h16 ? 0.f : c1) + __shfl_xor_sync(0xffffffffu, h16 ? c1 : 0.f, 16);
                    float q = (h8 ? r1 : r0) + __shfl_xor_sync(0xffffffffu, h8 ? r0 : r1, 8);
                    q += __shfl_xor_sync(0xffffffffu, q, 4);
                    q += __shfl_xor_sync(0xffffffffu, q, 2);
                    q += __shfl_xor_sync(0xffffffffu, q, 1);
                    if (has_w) asm volatile("st.shared.f32 [%0], %1;" ::"r"(st_w + 4u * (uint32_t)j), "f"(w) : "memory");
                    if (has_q) asm volatile("st.shared.f32 [%0], %1;" ::"r"(st_q + 4u * (uint32_t)j), "f"(q) : "memory");
                    mybits |= 1u << b31;
                }
                if (lane == (g0 >> 5)) wmask = mybits;
            }
        }
        if (lane < CHUNK / 32) s_mask[warp][lane] = wmask;
        __syncthreads();
        if ((int)threadIdx.x < cnt) {
            const int j = threadIdx.x;
            const float4 a = s_rec[j].a, b = s_rec[j].b;
            float s[9];
#pragma unroll
            for (int q = 0; q < 9; q++) s[q] = 0.f;
            bool any = false;
#pragma unroll
            for (int w = 0; w < BL_WARPS; w++) {
                if ((s_mask[w][j >> 5] >> (j & 31)) & 1u) {
                    any = true;
                    const float *cw = s_acc + (size_t)w * BWH_NVAL * STRIDE + j;
                    float cf[BWH_NVAL];
#pragma unroll
                    for (int q = 0; q < BWH_NVAL; q++) cf[q] = cw[q * STRIDE];
                    // with sum m b_i = (W_0 - E_i)/2 and sum m b_i b_j = (W_0 - E_i - E_j + P_ij)/4 substituted into
                    // lx = b0 + 2 b1 + 4 b2, ly = b3 + 2 b4 and collected per coefficient
                    // (E_i = cf[1+i]; P_ij = cf[6 + i(9-i)/2 + j-i-1]: P01 6, P02 7, P03 8, P04 9, P12 10, P13 11, P14 12,
                    //  P23 13, P24 14, P34 15):
                    const float W0 = cf[0], E0 = cf[1], E1 = cf[2], E2 = cf[3], E3 = cf[4], E4 = cf[5];
                    const float Lx = 3.5f * W0 - 0.5f * E0 - E1 - 2.f * E2;                         // sum m lx
                    const float Ly = 1.5f * W0 - 0.5f * E3 - E4;                                    // sum m ly
                    const float Lxx = 17.5f * W0 - 3.5f * E0 - 7.f * E1 - 14.f * E2 + cf[6] + 2.f * cf[7] + 4.f * cf[10];
                    const float Lyy = 3.5f * W0 - 1.5f * E3 - 3.f * E4 + cf[15];
                    const float Lxy = 0.25f * (21.f * W0 - 3.f * E0 - 6.f * E1 - 12.f * E2 - 7.f * E3 - 14.f * E4 + cf[8] +
                                               2.f * cf[9] + 2.f * cf[11] + 4.f * cf[12] + 4.f * cf[13] + 8.f * cf[14]);
                    // offsets of the splat centre from this warp's block origin: dx = ux - lx, dy = uy - ly
                    const float ux = a.x - (X0 + (float)((w & 1) * 8)), uy = a.y - (Y0 + (float)((w >> 1) * 4));
                    s[0] += ux * W0 - Lx;
                    s[1] += uy * W0 - Ly;
                    s[2] += ux * (ux * W0 - 2.f * Lx) + Lxx;
                    s[3] += ux * (uy * W0 - Ly) - uy * Lx + Lxy;
                    s[4] += uy * (uy * W0 - 2.f * Ly) + Lyy;
                    s[5] += W0;
                    s[6] += cf[16]; s[7] += cf[17]; s[8] += cf[18];
                }
            }
            if (any) {
                const uint32_t g = s_id[j];
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

template <int CHUNK>
static int launch_bwd_wht(int grid, cudaStream_t stream, int W, int H, int T1, const float4 *rec, const float *bg,
                          const uint8_t *cl, const uint2 *ranges, const uint32_t *ids, const float *final_T,
                          const uint32_t *n_contrib, const float *dimg, float *d_m2, float *d_co, float *d_rgb) {
    const size_t dyn = sizeof(float) * BL_WARPS * BWH_NVAL * (CHUNK + 1);
    GS_CUDA_TRY(cudaFuncSetAttribute(k_blend_bwd_wht<CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    k_blend_bwd_wht<CHUNK><<<grid, BL_THREADS, dyn, stream>>>(W, H, T1, rec, bg, cl, ranges, ids, final_T, n_contrib, dimg,
                                                              d_m2, d_co, d_rgb);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

// ---- backward, EXPERIMENTAL variant 2 (off by default; gs_debug_set(GS_DEBUG_BWD_AUTO)) ---------------------------
// "Warp-autonomous": no CTA barrier and no shared partial-sum slots at all.  Every warp walks the tile's list on its
// own for its two 4x4 pixel blocks (one per half-warp, the forward's layout: ~60 % useful lanes instead of 37 %), 32
// entries at a time: lane l loads entry l's record into a warp-private staging row and computes its block mask, two
// ballots give each half-warp its candidates, and each candidate pair is reduced over the 16 lanes of its half (9-value
// transposing butterfly, 12 SHFL for BOTH halves together) and leaves the SM at once as ONE predicated RED.ADD
// instruction: the lane pair that ends up holding a sum adds "its" terms of the gradient (even lane / odd lane take
// the two outputs a sum feeds -- e.g. sum(m dx) feeds dL/dmean_x with 2a' and dL/dmean_y with b').  That is 11 global
// atomics per (4x4 block, splat) instead of 9 per (tile, splat), in exchange for: no flush pass, no barrier stalls (the
// default kernel idles ~20 % of its issue slots waiting at 3 barriers per chunk), and ~75 instead of 110 issued
// instructions per 8x4-block-equivalent pair.  Records are re-read by all 8 warps of a tile (L1 hits).  Whether the
// L2 atomic units keep up (~1.8e8 RED per c2 backward) is what the device run has to tell.
GS_D void half_reduce9(float v[9], int l16) {
    {
        const bool h = l16 & 8;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float send = h ? v[i] : v[i + 4], keep = h ? v[i + 4] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool h = l16 & 4;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const float send = h ? v[i] : v[i + 2], keep = h ? v[i + 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    {
        const bool h = l16 & 2;
        const float send = h ? v[0] : v[1], keep = h ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);  // lanes 2q, 2q+1 of the half: total of value q
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v[8] += __shfl_xor_sync(0xffffffffu, v[8], o);
}

__global__ void __launch_bounds__(BL_THREADS)
k_blend_bwd_auto(int W, int H, int tiles_per_view, const float4 *__restrict__ rec, const float *__restrict__ bg,
                 const uint8_t *__restrict__ compute_locally, const uint2 *__restrict__ ranges,
                 const uint32_t *__restrict__ ids, const float *__restrict__ final_T,
                 const uint32_t *__restrict__ n_contrib, const float *__restrict__ dL_dimage,
                 float *__restrict__ d_means2D, float *__restrict__ d_conic_opacity, float *__restrict__ d_rgb) {
    __shared__ SRec s_stage[BL_WARPS][32];
    const int gx = (W + GS_BLOCK_X - 1) / GS_BLOCK_X;
    if (!compute_locally[blockIdx.x]) return;
    const int view = blockIdx.x / tiles_per_view, tile = blockIdx.x - view * tiles_per_view;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Where me = where_am_i(tile, gx);
    const int px = me.px, py = me.py, l16 = me.l16;
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
    const int last = inside ? (int)n_contrib[pix] : 0;
    float dp0 = 0.f, dp1 = 0.f, dp2 = 0.f;
    if (inside) { dp0 = dL_dimage[pix]; dp1 = dL_dimage[HW + pix]; dp2 = dL_dimage[2 * HW + pix]; }
    const float bgdot = bg[0] * dp0 + bg[1] * dp1 + bg[2] * dp2;
    const float ddx = 0.5f * (float)W, ddy = 0.5f * (float)H;  // dL/dmeans2D is per NDC unit
    // What this lane adds after the reduction (it holds the sum with index (l16 >> 1) & 7; odd lanes take the second
    // output of that sum): target array, element, and the coefficient  kA a' + kB b' + kC c' + kK + kO / opacity.
    float kA = 0.f, kB = 0.f, kC = 0.f, kK = 0.f, kO = 0.f;
    float *out = nullptr;
    int stride = 0;
    switch (l16) {
        case 0: out = d_means2D; stride = 2; kA = 2.f * ddx; break;            // sum m dx   -> mean_x
        case 1: out = d_means2D + 1; stride = 2; kB = ddy; break;              // sum m dx   -> mean_y
        case 2: out = d_means2D; stride = 2; kB = ddx; break;                  // sum m dy   -> mean_x
        case 3: out = d_means2D + 1; stride = 2; kC = 2.f * ddy; break;        // sum m dy   -> mean_y
        case 4: out = d_conic_opacity; stride = 4; kK = -0.5f; break;          // sum m dx^2
        case 5: out = d_rgb + 2; stride = 3; kK = 1.f; break;                  // blue (value 8, held by every lane)
        case 6: out = d_conic_opacity + 1; stride = 4; kK = -1.f; break;       // sum m dx dy
        case 8: out = d_conic_opacity + 2; stride = 4; kK = -0.5f; break;      // sum m dy^2
        case 10: out = d_conic_opacity + 3; stride = 4; kO = 1.f; break;       // sum m      -> opacity
        case 12: out = d_rgb; stride = 3; kK = 1.f; break;                     // red
        case 14: out = d_rgb + 1; stride = 3; kK = 1.f; break;                 // green
        default: break;
    }
    const bool takes_v8 = l16 == 5;
    const uint32_t stride4 = 4u * (uint32_t)stride;  // bytes between the target elements of consecutive splats
    uint32_t m = (uint32_t)last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    const int wlast = (int)m;  // deepest entry any pixel of this warp reaches
    const int blkA = me.blk - me.half;
    const uint32_t half_lanes = me.half ? 0xffff0000u : 0x0000ffffu;
    SRec *stage = s_stage[warp];
    float T = T_final, B0 = 0.f, B1 = 0.f, B2 = 0.f;
    for (int g0 = (wlast - 1) & ~31; g0 >= 0 && wlast > 0; g0 -= 32) {
        // stage entries g0 .. g0+31 (lane l <- entry g0 + l)
        const int e = g0 + lane;
        uint32_t gid = 0u, m16 = 0u;
        __syncwarp();  // the previous batch's reads of the staging row are done
        if (e < wlast) {
            gid = ids[range.x + e];
            const float4 *r = rec + (size_t)3 * gid;
            const float4 a = __ldg(r), b = __ldg(r + 1), cc = __ldg(r + 2);
            stage[lane].a = a; stage[lane].b = b; stage[lane].c = cc;
            m16 = block_mask16(a.x, a.y, cc.z, cc.w, X0, Y0);
        }
        __syncwarp();
        const uint32_t cA = __ballot_sync(0xffffffffu, (m16 >> blkA) & 1u),
                       cB = __ballot_sync(0xffffffffu, (m16 >> (blkA + 1)) & 1u);
        uint32_t mine = me.half ? cB : cA;  // this half-warp's candidates; bit l <-> entry g0 + l
        const int last_rel = last - g0;     // entry g0 + j is live for this pixel iff j < last_rel
        while (__any_sync(0xffffffffu, mine != 0u)) {
            const bool has = mine != 0u;
            const int j = has ? 31 - __clz(mine) : 0;  // deepest remaining candidate of this half
            mine &= ~(1u << j);                         // (no-op when !has: mine == 0)
            const SRec *sr = &stage[j];
            const float4 a = sr->a, b = sr->b;
            const float dx = a.x - pxf, dy = a.y - pyf;
            const float power = dx * (a.z * dx + a.w * dy) + b.x * dy * dy;
            bool ok = (has & (j < last_rel)) & (power >= b.z);   // '&': evaluate everything once, no short circuit
            if (!__any_sync(0xffffffffu, ok)) continue;
            const float G = gs_exp_neg(power);
            const float alpha = fminf(ALPHA_MAX, b.y * G);
            ok = ok & (power <= 0.f) & (alpha >= ALPHA_MIN);
            const uint32_t okb = __ballot_sync(0xffffffffu, ok);
            if (okb == 0u) continue;
            float v[9];
            {
                const float2 gb = *reinterpret_cast<const float2 *>(&sr->c);
                const float ae = ok ? alpha : 0.f;
                float inv;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(1.f - ae));
                T = T * inv;
                const float d0 = b.w - B0, d1 = gb.x - B1, d2 = gb.y - B2;
                const float dL_dalpha = (d0 * dp0 + d1 * dp1 + d2 * dp2) * T - (T_final * inv) * bgdot;
                const float mw = ok ? b.y * dL_dalpha * G : 0.f;
                const float dch = ae * T;
                B0 += ae * d0; B1 += ae * d1; B2 += ae * d2;
                const float mx_ = mw * dx, my_ = mw * dy;
                v[0] = mx_; v[1] = my_; v[2] = mx_ * dx; v[3] = mx_ * dy; v[4] = my_ * dy; v[5] = mw;
                v[6] = dch * dp0; v[7] = dch * dp1; v[8] = dch * dp2;
            }
            half_reduce9(v, l16);
            const uint32_t g = __shfl_sync(0xffffffffu, gid, j);  // splat id of this half's entry
            if (out != nullptr && (okb & half_lanes) != 0u) {
                float inv_o;  // opacity >= 1/255 here (alpha >= 1/255 was reached): no denormal handling needed
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv_o) : "f"(b.y));
                const float coef = kA * a.z + kB * a.w + kC * b.x + kK + kO * inv_o;
                float *dst = reinterpret_cast<float *>(reinterpret_cast<char *>(out) + (size_t)g * stride4);
                atomicAdd(dst, coef * (takes_v8 ? v[8] : v[0]));
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
    if (g_gs_debug_flags & (GS_DEBUG_BWD_WHT_64 | GS_DEBUG_BWD_WHT_128 | GS_DEBUG_BWD_AUTO)) {  // experimental, see above
        const float4 *r4 = reinterpret_cast<const float4 *>(rec);
        const uint2 *rg = reinterpret_cast<const uint2 *>(ranges);
        if (g_gs_debug_flags & GS_DEBUG_BWD_AUTO) {
            k_blend_bwd_auto<<<gx * gy * num_views, BL_THREADS, 0, stream>>>(
                image_width, image_height, gx * gy, r4, bg, compute_locally, rg, ids_sorted, final_T, n_contrib, dL_dimage,
                dL_dmeans2D, dL_dconic_opacity, dL_drgb);
            GS_LAUNCH_CHECK();
            return GS_OK;
        }
        if (g_gs_debug_flags & GS_DEBUG_BWD_WHT_64)
            return launch_bwd_wht<64>(gx * gy * num_views, stream, image_width, image_height, gx * gy, r4, bg,
                                      compute_locally, rg, ids_sorted, final_T, n_contrib, dL_dimage, dL_dmeans2D,
                                      dL_dconic_opacity, dL_drgb);
        return launch_bwd_wht<128>(gx * gy * num_views, stream, image_width, image_height, gx * gy, r4, bg, compute_locally,
                                   rg, ids_sorted, final_T, n_contrib, dL_dimage, dL_dmeans2D, dL_dconic_opacity, dL_drgb);
    }
    k_blend_bwd<<<gx * gy * num_views, BL_THREADS, 0, stream>>>(
        image_width, image_height, gx * gy, reinterpret_cast<const float4 *>(rec), bg, compute_locally,
        reinterpret_cast<const uint2 *>(ranges), ids_sorted, final_T, n_contrib, dL_dimage, dL_dmeans2D, dL_dconic_opacity,
        dL_drgb);
    GS_LAUNCH_CHECK();
    return GS_OK;
}
