/*
 * gs_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the tile-based differentiable Gaussian rasterizer that
 * nyu-systems/Grendel-GS calls through `diff_gaussian_rasterization` (its CUDA source is an
 * un-vendored submodule: /root/reference/.SUBMODULES.json:9-15 pins
 * nyu-systems/diff-gaussian-rasterization @ eea4b699).  Because that source is absent, the
 * arithmetic below restates the PUBLISHED 3DGS algorithm (Kerbl et al., SIGGRAPH 2023 -- the
 * fork states it forks it, /root/reference/README.md:100) and is anchored on the reference's
 * in-tree call sites and conventions; each function cites the file:line it follows.
 *
 * PARITY STATUS: "parity unpinned" for the rasterizer proper -- the reference ships no tests,
 * golden vectors or fixtures for this path.  The pieces that DO exist
 * in-tree as Python (SH basis utils/sh_utils.py:57-120, camera matrices
 * utils/graphics_utils.py:42-76, L1/SSIM utils/loss_utils.py:88-132) are pinned against golden
 * vectors generated from those files (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (grendel-gs_b200/) never links, imports or executes it.
 *
 * Build: see oracle/Makefile.  -DGSO_REAL_IS_DOUBLE builds the fp64 variant used for
 * finite-difference checks of the hand-written backward.  -ffp-contract=off is REQUIRED: the
 * integer-deciding chain (radius, tile rect, depth key) is written as an explicit sequence of
 * IEEE fp32 operations that the CUDA kernels repeat op for op, which is what makes tile
 * indices bit-exact.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef GSO_REAL_IS_DOUBLE
typedef double real;
#define R_SQRT sqrt
#define R_EXP exp
#define R_CEIL ceil
#define R_FABS fabs
#else
typedef float real;
#define R_SQRT sqrtf
#define R_EXP expf
#define R_CEIL ceilf
#define R_FABS fabsf
#endif
#define RC(x) ((real)(x))

#define BLOCK_X 16
#define BLOCK_Y 16

static inline real r_min(real a, real b) { return a < b ? a : b; }
static inline real r_max(real a, real b) { return a > b ? a : b; }
static inline int i_min(int a, int b) { return a < b ? a : b; }
static inline int i_max(int a, int b) { return a > b ? a : b; }

/* SH basis constants: /root/reference/utils/sh_utils.py:26-55 */
static const real SH_C0 = RC(0.28209479177387814);
static const real SH_C1 = RC(0.4886025119029199);
static const real SH_C2[5] = {RC(1.0925484305920792), RC(-1.0925484305920792), RC(0.31539156525252005),
                              RC(-1.0925484305920792), RC(0.5462742152960396)};
static const real SH_C3[7] = {RC(-0.5900435899266435), RC(2.890611442640554), RC(-0.4570457994644658),
                              RC(0.3731763325901154), RC(-0.4570457994644658), RC(1.445305721320277),
                              RC(-0.5900435899266435)};

int gso_real_bytes(void) { return (int)sizeof(real); }

void gso_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* Compile-time tile constants the reference reads through _C.get_block_XY()
 * (/root/reference/arguments/__init__.py:254-257; utils/general_utils.py:78-86). */
void gso_get_block_xy(int *bx, int *by, int *one_dim) { *bx = BLOCK_X; *by = BLOCK_Y; *one_dim = 256; }

/* Tile rectangle touched by a splat of integer pixel radius r centred at (px,py).
 * Tiles are 16x16 (utils/general_utils.py:78-93: TILE_X = ceil(W/16)). */
static inline void get_rect(real px, real py, int r, int gx, int gy, int *x0, int *y0, int *x1, int *y1) {
    real rr = (real)r;
    *x0 = i_min(gx, i_max(0, (int)((px - rr) / RC(BLOCK_X))));
    *y0 = i_min(gy, i_max(0, (int)((py - rr) / RC(BLOCK_Y))));
    *x1 = i_min(gx, i_max(0, (int)((px + rr + RC(BLOCK_X - 1)) / RC(BLOCK_X))));
    *y1 = i_min(gy, i_max(0, (int)((py + rr + RC(BLOCK_Y - 1)) / RC(BLOCK_Y))));
}

/* Rotation matrix of a (w,x,y,z) quaternion, as /root/reference/utils/general_utils.py:416-438
 * (the operator receives rotations already normalised: scene/gaussian_model.py:114-115). */
static inline void quat_to_R(const real *q, real R[9]) {
    real r = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = RC(1) - RC(2) * (y * y + z * z); R[1] = RC(2) * (x * y - r * z); R[2] = RC(2) * (x * z + r * y);
    R[3] = RC(2) * (x * y + r * z); R[4] = RC(1) - RC(2) * (x * x + z * z); R[5] = RC(2) * (y * z - r * x);
    R[6] = RC(2) * (x * z - r * y); R[7] = RC(2) * (y * z + r * x); R[8] = RC(1) - RC(2) * (x * x + y * y);
}

/* Sigma = (R S)(R S)^T, /root/reference/utils/general_utils.py:441-451 (build_scaling_rotation)
 * with the operator's scale_modifier applied; 6 unique entries xx,xy,xz,yy,yz,zz. */
static inline void cov3d_from(const real *scale, real mod, const real *q, real L[9], real S6[6]) {
    real R[9];
    quat_to_R(q, R);
    real s0 = mod * scale[0], s1 = mod * scale[1], s2 = mod * scale[2];
    L[0] = R[0] * s0; L[1] = R[1] * s1; L[2] = R[2] * s2;
    L[3] = R[3] * s0; L[4] = R[4] * s1; L[5] = R[5] * s2;
    L[6] = R[6] * s0; L[7] = R[7] * s1; L[8] = R[8] * s2;
    S6[0] = L[0] * L[0] + L[1] * L[1] + L[2] * L[2];
    S6[1] = L[0] * L[3] + L[1] * L[4] + L[2] * L[5];
    S6[2] = L[0] * L[6] + L[1] * L[7] + L[2] * L[8];
    S6[3] = L[3] * L[3] + L[4] * L[4] + L[5] * L[5];
    S6[4] = L[3] * L[6] + L[4] * L[7] + L[5] * L[8];
    S6[5] = L[6] * L[6] + L[7] * L[7] + L[8] * L[8];
}

/* SH -> RGB, basis of /root/reference/utils/sh_utils.py:57-120, then +0.5 and clamp at 0
 * (the reference's gsplat path shows the same post-processing: gaussian_renderer/__init__.py:1141).
 * sh is (16,3) per Gaussian (scene/gaussian_model.py:122-125). Returns the clamp mask bits. */
static inline void sh_basis(int deg, real x, real y, real z, real b[16]) {
    b[0] = SH_C0;
    if (deg > 0) {
        b[1] = -SH_C1 * y; b[2] = SH_C1 * z; b[3] = -SH_C1 * x;
        if (deg > 1) {
            real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = SH_C2[0] * xy; b[5] = SH_C2[1] * yz; b[6] = SH_C2[2] * (RC(2) * zz - xx - yy);
            b[7] = SH_C2[3] * xz; b[8] = SH_C2[4] * (xx - yy);
            if (deg > 2) {
                b[9] = SH_C3[0] * y * (RC(3) * xx - yy);
                b[10] = SH_C3[1] * xy * z;
                b[11] = SH_C3[2] * y * (RC(4) * zz - xx - yy);
                b[12] = SH_C3[3] * z * (RC(2) * zz - RC(3) * xx - RC(3) * yy);
                b[13] = SH_C3[4] * x * (RC(4) * zz - xx - yy);
                b[14] = SH_C3[5] * z * (xx - yy);
                b[15] = SH_C3[6] * x * (xx - RC(3) * yy);
            }
        }
    }
}

/* Exposed for the golden-vector test against utils/sh_utils.py:eval_sh. */
void gso_eval_sh(int n, int deg, const real *sh /* n,16,3 */, const real *dirs /* n,3 unit */, real *out /* n,3 */) {
    int ncoef = (deg + 1) * (deg + 1);
    for (int i = 0; i < n; i++) {
        real b[16];
        sh_basis(deg, dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], b);
        for (int c = 0; c < 3; c++) {
            real acc = 0;
            for (int k = 0; k < ncoef; k++) acc += b[k] * sh[(i * 16 + k) * 3 + c];
            out[3 * i + c] = acc;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Stage "10 preprocess" (/root/reference/analyze_statistic.py:1973), called at
 * /root/reference/gaussian_renderer/__init__.py:949-956.
 * viewmatrix / projmatrix are the row-vector (transposed) 4x4 of scene/cameras.py:84-99,
 * row-major in memory, so point p maps to [p,1] @ M.
 * Outputs for culled Gaussians are all zero; radii == 0 marks them (__init__.py:971-973).
 * ------------------------------------------------------------------------------------------ */
void gso_preprocess_forward(int P, int D, const real *means3D, const real *scales, real scale_modifier,
                            const real *rotations, const real *opacities, const real *shs,
                            const real *V, const real *PM, const real *campos, int W, int H, real tanfovx,
                            real tanfovy, real *means2D, real *depths, int32_t *radii, real *cov3D,
                            real *conic_opacity, real *rgb, uint8_t *clamped) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    const real fx = (real)W / (RC(2) * tanfovx), fy = (real)H / (RC(2) * tanfovy);
    const int ncoef = (D + 1) * (D + 1);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        means2D[2 * i] = means2D[2 * i + 1] = 0; depths[i] = 0; radii[i] = 0;
        for (int k = 0; k < 6; k++) cov3D[6 * i + k] = 0;
        for (int k = 0; k < 4; k++) conic_opacity[4 * i + k] = 0;
        for (int k = 0; k < 3; k++) rgb[3 * i + k] = 0;
        clamped[i] = 0;
        const real px = means3D[3 * i], py = means3D[3 * i + 1], pz = means3D[3 * i + 2];
        /* view space */
        const real tx = V[0] * px + V[4] * py + V[8] * pz + V[12];
        const real ty = V[1] * px + V[5] * py + V[9] * pz + V[13];
        const real tz = V[2] * px + V[6] * py + V[10] * pz + V[14];
        if (tz <= RC(0.2)) continue; /* near cull */
        /* clip space */
        const real hx = PM[0] * px + PM[4] * py + PM[8] * pz + PM[12];
        const real hy = PM[1] * px + PM[5] * py + PM[9] * pz + PM[13];
        const real hw = PM[3] * px + PM[7] * py + PM[11] * pz + PM[15];
        const real pw = RC(1) / (hw + RC(0.0000001));
        const real ndcx = hx * pw, ndcy = hy * pw;
        /* 3D covariance */
        real L[9], S[6];
        cov3d_from(scales + 3 * i, scale_modifier, rotations + 4 * i, L, S);
        /* EWA projection with 1.3x guard-band clamp and 0.3 px low-pass */
        const real limx = RC(1.3) * tanfovx, limy = RC(1.3) * tanfovy;
        const real txtz = tx / tz, tytz = ty / tz;
        const real cx = r_min(limx, r_max(-limx, txtz)) * tz;
        const real cy = r_min(limy, r_max(-limy, tytz)) * tz;
        const real J00 = fx / tz, J02 = -(fx * cx) / (tz * tz);
        const real J11 = fy / tz, J12 = -(fy * cy) / (tz * tz);
        /* W[j][i] = V[i][j]: rotation rows of the world->view map */
        const real T00 = J00 * V[0] + J02 * V[2], T01 = J00 * V[4] + J02 * V[6], T02 = J00 * V[8] + J02 * V[10];
        const real T10 = J11 * V[1] + J12 * V[2], T11 = J11 * V[5] + J12 * V[6], T12 = J11 * V[9] + J12 * V[10];
        const real u00 = S[0] * T00 + S[1] * T01 + S[2] * T02;
        const real u01 = S[1] * T00 + S[3] * T01 + S[4] * T02;
        const real u02 = S[2] * T00 + S[4] * T01 + S[5] * T02;
        const real u10 = S[0] * T10 + S[1] * T11 + S[2] * T12;
        const real u11 = S[1] * T10 + S[3] * T11 + S[4] * T12;
        const real u12 = S[2] * T10 + S[4] * T11 + S[5] * T12;
        const real a = T00 * u00 + T01 * u01 + T02 * u02 + RC(0.3);
        const real b = T00 * u10 + T01 * u11 + T02 * u12;
        const real c = T10 * u10 + T11 * u11 + T12 * u12 + RC(0.3);
        const real det = a * c - b * b;
        if (det == RC(0)) continue;
        const real det_inv = RC(1) / det;
        const real mid = RC(0.5) * (a + c);
        const real disc = R_SQRT(r_max(RC(0.1), mid * mid - det));
        const real lam = r_max(mid + disc, mid - disc);
        const int rad = (int)R_CEIL(RC(3) * R_SQRT(lam));
        const real ix = ((ndcx + RC(1)) * (real)W - RC(1)) * RC(0.5);
        const real iy = ((ndcy + RC(1)) * (real)H - RC(1)) * RC(0.5);
        int x0, y0, x1, y1;
        get_rect(ix, iy, rad, gx, gy, &x0, &y0, &x1, &y1);
        if ((x1 - x0) * (y1 - y0) == 0) continue;
        /* colour */
        real dx = px - campos[0], dy = py - campos[1], dz = pz - campos[2];
        const real len = R_SQRT(dx * dx + dy * dy + dz * dz);
        dx = dx / len; dy = dy / len; dz = dz / len;
        real bas[16];
        sh_basis(D, dx, dy, dz, bas);
        uint8_t cm = 0;
        for (int ch = 0; ch < 3; ch++) {
            real acc = 0;
            for (int k = 0; k < ncoef; k++) acc += bas[k] * shs[(i * 16 + k) * 3 + ch];
            acc += RC(0.5);
            if (acc < RC(0)) { cm |= (uint8_t)(1u << ch); acc = 0; }
            rgb[3 * i + ch] = acc;
        }
        clamped[i] = cm;
        means2D[2 * i] = ix; means2D[2 * i + 1] = iy;
        depths[i] = tz; radii[i] = rad;
        for (int k = 0; k < 6; k++) cov3D[6 * i + k] = S[k];
        conic_opacity[4 * i + 0] = c * det_inv;
        conic_opacity[4 * i + 1] = -b * det_inv;
        conic_opacity[4 * i + 2] = a * det_inv;
        conic_opacity[4 * i + 3] = opacities[i];
    }
}

/* ------------------------------------------------------------------------------------------
 * Stage "b20 preprocess" backward (/root/reference/analyze_statistic.py:1990): autograd of
 * preprocess_gaussians (gaussian_renderer/__init__.py:949-958).
 * dL_dmeans2D is expressed PER NDC UNIT, i.e. (dL/dpix_x * W/2, dL/dpix_y * H/2): that is the
 * quantity the reference's densification reads from means2D.grad -- its gsplat path rescales
 * true pixel gradients by (W/2, H/2) to match it (scene/gaussian_model.py:1053-1063).
 * dL_dconic_opacity = (dA, dB, dC, dOpacity) are true partial derivatives wrt the four stored
 * numbers of conic_opacity.
 * ------------------------------------------------------------------------------------------ */
void gso_preprocess_backward(int P, int D, const real *means3D, const real *scales, real scale_modifier,
                             const real *rotations, const real *opacities, const real *shs, const real *V,
                             const real *PM, const real *campos, int W, int H, real tanfovx, real tanfovy,
                             const int32_t *radii, const uint8_t *clamped, const real *dL_dmeans2D,
                             const real *dL_dconic_opacity, const real *dL_drgb, real *dL_dmeans3D,
                             real *dL_dscales, real *dL_drotations, real *dL_dopacities, real *dL_dshs) {
    (void)opacities;
    const real fx = (real)W / (RC(2) * tanfovx), fy = (real)H / (RC(2) * tanfovy);
    const int ncoef = (D + 1) * (D + 1);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        real *gm = dL_dmeans3D + 3 * i, *gs = dL_dscales + 3 * i, *gq = dL_drotations + 4 * i;
        gm[0] = gm[1] = gm[2] = 0; gs[0] = gs[1] = gs[2] = 0; gq[0] = gq[1] = gq[2] = gq[3] = 0;
        dL_dopacities[i] = 0;
        for (int k = 0; k < 48; k++) dL_dshs[48 * i + k] = 0;
        if (radii[i] <= 0) continue;
        const real px = means3D[3 * i], py = means3D[3 * i + 1], pz = means3D[3 * i + 2];
        const real tx = V[0] * px + V[4] * py + V[8] * pz + V[12];
        const real ty = V[1] * px + V[5] * py + V[9] * pz + V[13];
        const real tz = V[2] * px + V[6] * py + V[10] * pz + V[14];
        real L[9], S[6];
        cov3d_from(scales + 3 * i, scale_modifier, rotations + 4 * i, L, S);
        const real limx = RC(1.3) * tanfovx, limy = RC(1.3) * tanfovy;
        const real txtz = tx / tz, tytz = ty / tz;
        const real cx = r_min(limx, r_max(-limx, txtz)) * tz;
        const real cy = r_min(limy, r_max(-limy, tytz)) * tz;
        const real xmul = (txtz < -limx || txtz > limx) ? RC(0) : RC(1);
        const real ymul = (tytz < -limy || tytz > limy) ? RC(0) : RC(1);
        const real J00 = fx / tz, J02 = -(fx * cx) / (tz * tz);
        const real J11 = fy / tz, J12 = -(fy * cy) / (tz * tz);
        const real T0[3] = {J00 * V[0] + J02 * V[2], J00 * V[4] + J02 * V[6], J00 * V[8] + J02 * V[10]};
        const real T1[3] = {J11 * V[1] + J12 * V[2], J11 * V[5] + J12 * V[6], J11 * V[9] + J12 * V[10]};
        const real Sm[3][3] = {{S[0], S[1], S[2]}, {S[1], S[3], S[4]}, {S[2], S[4], S[5]}};
        real u0[3], u1[3];
        for (int r = 0; r < 3; r++) {
            u0[r] = Sm[r][0] * T0[0] + Sm[r][1] * T0[1] + Sm[r][2] * T0[2];
            u1[r] = Sm[r][0] * T1[0] + Sm[r][1] * T1[1] + Sm[r][2] * T1[2];
        }
        const real a = T0[0] * u0[0] + T0[1] * u0[1] + T0[2] * u0[2] + RC(0.3);
        const real b = T0[0] * u1[0] + T0[1] * u1[1] + T0[2] * u1[2];
        const real c = T1[0] * u1[0] + T1[1] * u1[1] + T1[2] * u1[2] + RC(0.3);
        const real det = a * c - b * b;
        /* conic (A,B,C) = (c,-b,a)/det -> cov2D (a,b,c) */
        const real dA = dL_dconic_opacity[4 * i], dB = dL_dconic_opacity[4 * i + 1], dC = dL_dconic_opacity[4 * i + 2];
        const real d2 = RC(1) / (det * det);
        real dLda = 0, dLdb = 0, dLdc = 0;
        if (det != RC(0)) {
            dLda = d2 * (-c * c * dA + b * c * dB - b * b * dC);
            dLdb = d2 * (RC(2) * b * c * dA - (det + RC(2) * b * b) * dB + RC(2) * a * b * dC);
            dLdc = d2 * (-b * b * dA + a * b * dB - a * a * dC);
        }
        /* cov2D -> Sigma (symmetric parameterisation: off-diagonals counted once) */
        real G[3][3];
        for (int r = 0; r < 3; r++)
            for (int s = 0; s < 3; s++)
                G[r][s] = T0[r] * T0[s] * dLda + RC(0.5) * (T0[r] * T1[s] + T0[s] * T1[r]) * dLdb + T1[r] * T1[s] * dLdc;
        /* cov2D -> T -> J -> t */
        real dT0[3], dT1[3];
        for (int r = 0; r < 3; r++) {
            dT0[r] = RC(2) * dLda * u0[r] + dLdb * u1[r];
            dT1[r] = RC(2) * dLdc * u1[r] + dLdb * u0[r];
        }
        /* W[l][j] = V[4*j + l] */
        const real dJ00 = dT0[0] * V[0] + dT0[1] * V[4] + dT0[2] * V[8];
        const real dJ02 = dT0[0] * V[2] + dT0[1] * V[6] + dT0[2] * V[10];
        const real dJ11 = dT1[0] * V[1] + dT1[1] * V[5] + dT1[2] * V[9];
        const real dJ12 = dT1[0] * V[2] + dT1[1] * V[6] + dT1[2] * V[10];
        const real tzi = RC(1) / tz, tzi2 = tzi * tzi, tzi3 = tzi2 * tzi;
        real dtx = xmul * (-fx * tzi2) * dJ02;
        real dty = ymul * (-fy * tzi2) * dJ12;
        real dtz = -fx * tzi2 * dJ00 - fy * tzi2 * dJ11 + RC(2) * fx * cx * tzi3 * dJ02 + RC(2) * fy * cy * tzi3 * dJ12;
        real gmx = V[0] * dtx + V[1] * dty + V[2] * dtz;
        real gmy = V[4] * dtx + V[5] * dty + V[6] * dtz;
        real gmz = V[8] * dtx + V[9] * dty + V[10] * dtz;
        /* means2D (NDC units) -> mean3D through the perspective divide */
        {
            const real hx = PM[0] * px + PM[4] * py + PM[8] * pz + PM[12];
            const real hy = PM[1] * px + PM[5] * py + PM[9] * pz + PM[13];
            const real hw = PM[3] * px + PM[7] * py + PM[11] * pz + PM[15];
            const real pw = RC(1) / (hw + RC(0.0000001));
            const real g0 = dL_dmeans2D[2 * i], g1 = dL_dmeans2D[2 * i + 1];
            const real dhx = g0 * pw, dhy = g1 * pw, dhw = -(g0 * hx + g1 * hy) * pw * pw;
            gmx += PM[0] * dhx + PM[1] * dhy + PM[3] * dhw;
            gmy += PM[4] * dhx + PM[5] * dhy + PM[7] * dhw;
            gmz += PM[8] * dhx + PM[9] * dhy + PM[11] * dhw;
        }
        /* colour -> SH coefficients and view direction */
        {
            real vx = px - campos[0], vy = py - campos[1], vz = pz - campos[2];
            const real len = R_SQRT(vx * vx + vy * vy + vz * vz);
            const real x = vx / len, y = vy / len, z = vz / len;
            real bas[16];
            sh_basis(D, x, y, z, bas);
            real dc[3];
            for (int ch = 0; ch < 3; ch++) dc[ch] = ((clamped[i] >> ch) & 1) ? RC(0) : dL_drgb[3 * i + ch];
            for (int k = 0; k < ncoef; k++)
                for (int ch = 0; ch < 3; ch++) dL_dshs[(i * 16 + k) * 3 + ch] = bas[k] * dc[ch];
            /* d(colour . dc)/d(dir): derivative of the basis polynomials */
            real ddx = 0, ddy = 0, ddz = 0;
            if (D > 0) {
                real s[16];
                for (int k = 0; k < 16; k++) {
                    s[k] = 0;
                    if (k < ncoef) for (int ch = 0; ch < 3; ch++) s[k] += shs[(i * 16 + k) * 3 + ch] * dc[ch];
                }
                ddx += -SH_C1 * s[3]; ddy += -SH_C1 * s[1]; ddz += SH_C1 * s[2];
                if (D > 1) {
                    const real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    ddx += SH_C2[0] * y * s[4] + SH_C2[2] * RC(2) * -x * s[6] + SH_C2[3] * z * s[7] + SH_C2[4] * RC(2) * x * s[8];
                    ddy += SH_C2[0] * x * s[4] + SH_C2[1] * z * s[5] + SH_C2[2] * RC(2) * -y * s[6] + SH_C2[4] * RC(2) * -y * s[8];
                    ddz += SH_C2[1] * y * s[5] + SH_C2[2] * RC(2) * RC(2) * z * s[6] + SH_C2[3] * x * s[7];
                    if (D > 2) {
                        ddx += SH_C3[0] * s[9] * RC(3) * RC(2) * xy + SH_C3[1] * s[10] * yz + SH_C3[2] * s[11] * -RC(2) * xy +
                               SH_C3[3] * s[12] * -RC(3) * RC(2) * xz + SH_C3[4] * s[13] * (-RC(3) * xx + RC(4) * zz - yy) +
                               SH_C3[5] * s[14] * RC(2) * xz + SH_C3[6] * s[15] * RC(3) * (xx - yy);
                        ddy += SH_C3[0] * s[9] * RC(3) * (xx - yy) + SH_C3[1] * s[10] * xz +
                               SH_C3[2] * s[11] * (-RC(3) * yy + RC(4) * zz - xx) + SH_C3[3] * s[12] * -RC(3) * RC(2) * yz +
                               SH_C3[4] * s[13] * -RC(2) * xy + SH_C3[5] * s[14] * -RC(2) * yz +
                               SH_C3[6] * s[15] * -RC(3) * RC(2) * xy;
                        ddz += SH_C3[1] * s[10] * xy + SH_C3[2] * s[11] * RC(4) * RC(2) * yz +
                               SH_C3[3] * s[12] * RC(3) * (RC(2) * zz - xx - yy) + SH_C3[4] * s[13] * RC(4) * RC(2) * xz +
                               SH_C3[5] * s[14] * (xx - yy);
                    }
                }
            }
            /* through the normalisation v/|v| */
            const real dot = x * ddx + y * ddy + z * ddz;
            gmx += (ddx - x * dot) / len;
            gmy += (ddy - y * dot) / len;
            gmz += (ddz - z * dot) / len;
        }
        gm[0] = gmx; gm[1] = gmy; gm[2] = gmz;
        /* Sigma = L L^T, L = R diag(mod*s) */
        {
            real R[9];
            quat_to_R(rotations + 4 * i, R);
            const real s[3] = {scale_modifier * scales[3 * i], scale_modifier * scales[3 * i + 1], scale_modifier * scales[3 * i + 2]};
            real dLm[9]; /* dL/dL = 2 G L */
            for (int r = 0; r < 3; r++)
                for (int k = 0; k < 3; k++)
                    dLm[3 * r + k] = RC(2) * (G[r][0] * L[k] + G[r][1] * L[3 + k] + G[r][2] * L[6 + k]);
            real dR[9];
            for (int k = 0; k < 3; k++) {
                gs[k] = scale_modifier * (dLm[k] * R[k] + dLm[3 + k] * R[3 + k] + dLm[6 + k] * R[6 + k]);
                dR[k] = dLm[k] * s[k]; dR[3 + k] = dLm[3 + k] * s[k]; dR[6 + k] = dLm[6 + k] * s[k];
            }
            const real r = rotations[4 * i], x = rotations[4 * i + 1], y = rotations[4 * i + 2], z = rotations[4 * i + 3];
            gq[0] = RC(2) * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
            gq[1] = RC(2) * (y * dR[1] + z * dR[2] + y * dR[3] - RC(2) * x * dR[4] - r * dR[5] + z * dR[6] + r * dR[7] - RC(2) * x * dR[8]);
            gq[2] = RC(2) * (-RC(2) * y * dR[0] + x * dR[1] + r * dR[2] + x * dR[3] + z * dR[5] - r * dR[6] + z * dR[7] - RC(2) * y * dR[8]);
            gq[3] = RC(2) * (-RC(2) * z * dR[0] - r * dR[1] + x * dR[2] + r * dR[3] - RC(2) * z * dR[4] + y * dR[5] + x * dR[6] + y * dR[7]);
        }
        dL_dopacities[i] = dL_dconic_opacity[4 * i + 3];
    }
}

/* ------------------------------------------------------------------------------------------
 * _C.get_local2j_ids_bool (/root/reference/gaussian_renderer/workload_division.py:721-744):
 * out[i][j] = splat i's tile rectangle intersects rank j's flattened tile-id range
 * [strategy[j], strategy[j+1]).  `rank` is accepted for signature parity and unused.
 * ------------------------------------------------------------------------------------------ */
void gso_get_local2j_ids_bool(int P, int H, int W, int world_size, const real *means2D, const int32_t *radii,
                              const int32_t *strategy, uint8_t *out) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    for (int i = 0; i < P; i++) {
        for (int j = 0; j < world_size; j++) out[i * world_size + j] = 0;
        if (radii[i] <= 0) continue;
        int x0, y0, x1, y1;
        get_rect(means2D[2 * i], means2D[2 * i + 1], radii[i], gx, gy, &x0, &y0, &x1, &y1);
        if (x1 <= x0) continue;
        for (int y = y0; y < y1; y++) {
            const int lo = y * gx + x0, hi = y * gx + x1;
            for (int j = 0; j < world_size; j++)
                if (i_max(lo, strategy[j]) < i_min(hi, strategy[j + 1])) out[i * world_size + j] = 1;
        }
    }
}

/* _C.get_local2j_ids_bool_adjust_mode6 (workload_division.py:471-484): rank j owns the tile
 * rectangle rects[j] = (y_l, y_r, x_l, x_r). */
void gso_get_local2j_ids_bool_rects(int P, int H, int W, int world_size, const real *means2D, const int32_t *radii,
                                    const int32_t *rects, uint8_t *out) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    for (int i = 0; i < P; i++) {
        for (int j = 0; j < world_size; j++) out[i * world_size + j] = 0;
        if (radii[i] <= 0) continue;
        int x0, y0, x1, y1;
        get_rect(means2D[2 * i], means2D[2 * i + 1], radii[i], gx, gy, &x0, &y0, &x1, &y1);
        for (int j = 0; j < world_size; j++) {
            const int32_t *r = rects + 4 * j;
            if (i_max(y0, r[0]) < i_min(y1, r[1]) && i_max(x0, r[2]) < i_min(x1, r[3])) out[i * world_size + j] = 1;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * render_gaussians forward, stages 21-24,30,40,50,60,70,81-83
 * (/root/reference/analyze_statistic.py:1976-1984; call site gaussian_renderer/__init__.py:1271-1282).
 * compute_locally is the (TILE_Y,TILE_X) bool mask of workload_division.py:773-787.
 * ------------------------------------------------------------------------------------------ */

/* stages 21-24 + 30: per-splat count of LOCAL tiles touched, inclusive scan; returns R. */
int64_t gso_render_count(int P, int H, int W, const real *means2D, const int32_t *radii, const uint8_t *compute_locally,
                         uint32_t *tiles_touched, uint32_t *offsets) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    uint64_t run = 0;
    for (int i = 0; i < P; i++) {
        uint32_t n = 0;
        if (radii[i] > 0) {
            int x0, y0, x1, y1;
            get_rect(means2D[2 * i], means2D[2 * i + 1], radii[i], gx, gy, &x0, &y0, &x1, &y1);
            for (int y = y0; y < y1; y++)
                for (int x = x0; x < x1; x++) n += compute_locally[y * gx + x] ? 1u : 0u;
        }
        tiles_touched[i] = n;
        run += n;
        offsets[i] = (uint32_t)run;
    }
    return (int64_t)run;
}

typedef struct { uint64_t key; uint32_t val; } kv_t;

static void radix_sort_kv(kv_t *a, kv_t *tmp, int64_t n, int bits) {
    /* LSD radix sort, 8 bits per pass: stable, like cub::DeviceRadixSort (stage "50"). */
    for (int shift = 0; shift < bits; shift += 8) {
        int64_t cnt[257];
        memset(cnt, 0, sizeof(cnt));
        for (int64_t i = 0; i < n; i++) cnt[((a[i].key >> shift) & 0xFF) + 1]++;
        for (int k = 0; k < 256; k++) cnt[k + 1] += cnt[k];
        for (int64_t i = 0; i < n; i++) tmp[cnt[(a[i].key >> shift) & 0xFF]++] = a[i];
        kv_t *t = a; a = tmp; tmp = t;
    }
    /* caller passes bits as a multiple of 16 so the result ends in the original buffer */
}

/* stages 40,50,60: (tile<<32 | depth bits, splat id) pairs, stable sort, per-tile [start,end). */
void gso_render_bin(int P, int64_t R, int H, int W, const real *means2D, const float *depths_f32, const int32_t *radii,
                    const uint8_t *compute_locally, const uint32_t *offsets, uint64_t *keys_sorted,
                    uint32_t *ids_sorted, uint32_t *ranges /* T,2 */) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    const int T = gx * gy;
    kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * (size_t)(R > 0 ? R : 1));
    kv_t *tmp = (kv_t *)malloc(sizeof(kv_t) * (size_t)(R > 0 ? R : 1));
    for (int i = 0; i < P; i++) {
        if (radii[i] <= 0) continue;
        uint64_t off = (i == 0) ? 0 : offsets[i - 1];
        int x0, y0, x1, y1;
        get_rect(means2D[2 * i], means2D[2 * i + 1], radii[i], gx, gy, &x0, &y0, &x1, &y1);
        uint32_t dbits;
        memcpy(&dbits, &depths_f32[i], 4);
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) {
                if (!compute_locally[y * gx + x]) continue;
                kv[off].key = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
                kv[off].val = (uint32_t)i;
                off++;
            }
    }
    radix_sort_kv(kv, tmp, R, 64);
    for (int t = 0; t < 2 * T; t++) ranges[t] = 0;
    for (int64_t k = 0; k < R; k++) {
        keys_sorted[k] = kv[k].key;
        ids_sorted[k] = kv[k].val;
        uint32_t tile = (uint32_t)(kv[k].key >> 32);
        if (k == 0 || tile != (uint32_t)(kv[k - 1].key >> 32)) ranges[2 * tile] = (uint32_t)k;
        if (k == R - 1 || tile != (uint32_t)(kv[k + 1].key >> 32)) ranges[2 * tile + 1] = (uint32_t)(k + 1);
    }
    free(kv); free(tmp);
}

/* stage 70 (+81-83): front-to-back alpha blend of every local tile. Non-local tiles stay
 * exactly 0 (/root/reference/gaussian_renderer/loss_distribution.py:1875).
 * stats[0..2] = sums over rendered pixels of (tile-list length, entries walked, entries blended):
 * the n_render / n_consider / n_contrib of gaussian_renderer/__init__.py:1271. */
void gso_render_blend_forward(int H, int W, const real *means2D, const real *conic_opacity, const real *rgb,
                              const real *bg, const uint8_t *compute_locally, const uint32_t *ranges,
                              const uint32_t *ids_sorted, real *out_color /* 3,H,W */, real *final_T /* H,W */,
                              uint32_t *n_contrib /* H,W */, int64_t *stats) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    const size_t HW = (size_t)H * W;
    for (size_t k = 0; k < 3 * HW; k++) out_color[k] = 0;
    for (size_t k = 0; k < HW; k++) { final_T[k] = 0; n_contrib[k] = 0; }
    int64_t s0 = 0, s1 = 0, s2 = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : s0, s1, s2)
    for (int t = 0; t < gx * gy; t++) {
        if (!compute_locally[t]) continue;
        const int ty = t / gx, tx = t % gx;
        const uint32_t beg = ranges[2 * t], end = ranges[2 * t + 1];
        for (int py = ty * BLOCK_Y; py < i_min(H, (ty + 1) * BLOCK_Y); py++)
            for (int px = tx * BLOCK_X; px < i_min(W, (tx + 1) * BLOCK_X); px++) {
                const real pfx = (real)px, pfy = (real)py;
                real Tr = RC(1), C0 = 0, C1 = 0, C2 = 0;
                uint32_t contributor = 0, last = 0, blended = 0;
                for (uint32_t k = beg; k < end; k++) {
                    contributor++;
                    const uint32_t g = ids_sorted[k];
                    const real dx = means2D[2 * g] - pfx, dy = means2D[2 * g + 1] - pfy;
                    const real *co = conic_opacity + 4 * g;
                    const real power = RC(-0.5) * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > RC(0)) continue;
                    const real alpha = r_min(RC(0.99), co[3] * R_EXP(power));
                    if (alpha < RC(1.0) / RC(255.0)) continue;
                    const real test_T = Tr * (RC(1) - alpha);
                    if (test_T < RC(0.0001)) break;
                    const real w = alpha * Tr;
                    C0 += rgb[3 * g] * w; C1 += rgb[3 * g + 1] * w; C2 += rgb[3 * g + 2] * w;
                    Tr = test_T;
                    last = contributor;
                    blended++;
                }
                const size_t pix = (size_t)py * W + px;
                out_color[pix] = C0 + Tr * bg[0];
                out_color[HW + pix] = C1 + Tr * bg[1];
                out_color[2 * HW + pix] = C2 + Tr * bg[2];
                final_T[pix] = Tr;
                n_contrib[pix] = last;
                s0 += (int64_t)(end - beg); s1 += contributor; s2 += blended;
            }
    }
    if (stats) { stats[0] = s0; stats[1] = s1; stats[2] = s2; }
}

/* stage "b10 render" backward (/root/reference/analyze_statistic.py:1987): reverse walk of every
 * local tile. Gradient conventions documented at gso_preprocess_backward. */
void gso_render_blend_backward(int P, int H, int W, const real *means2D, const real *conic_opacity, const real *rgb,
                               const real *bg, const uint8_t *compute_locally, const uint32_t *ranges,
                               const uint32_t *ids_sorted, const real *final_T, const uint32_t *n_contrib,
                               const real *dL_dpix /* 3,H,W */, real *dL_dmeans2D, real *dL_dconic_opacity,
                               real *dL_drgb) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    const size_t HW = (size_t)H * W;
    for (int k = 0; k < 2 * P; k++) dL_dmeans2D[k] = 0;
    for (int k = 0; k < 4 * P; k++) dL_dconic_opacity[k] = 0;
    for (int k = 0; k < 3 * P; k++) dL_drgb[k] = 0;
    const real ddelx_dx = RC(0.5) * (real)W, ddely_dy = RC(0.5) * (real)H;
#pragma omp parallel for schedule(dynamic, 4)
    for (int t = 0; t < gx * gy; t++) {
        if (!compute_locally[t]) continue;
        const int ty = t / gx, tx = t % gx;
        const uint32_t beg = ranges[2 * t];
        for (int py = ty * BLOCK_Y; py < i_min(H, (ty + 1) * BLOCK_Y); py++)
            for (int px = tx * BLOCK_X; px < i_min(W, (tx + 1) * BLOCK_X); px++) {
                const size_t pix = (size_t)py * W + px;
                const real pfx = (real)px, pfy = (real)py;
                const real T_final = final_T[pix];
                real Tr = T_final;
                const real dp[3] = {dL_dpix[pix], dL_dpix[HW + pix], dL_dpix[2 * HW + pix]};
                const real bgdot = bg[0] * dp[0] + bg[1] * dp[1] + bg[2] * dp[2];
                real accum[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0;
                for (int64_t k = (int64_t)beg + n_contrib[pix] - 1; k >= (int64_t)beg; k--) {
                    const uint32_t g = ids_sorted[k];
                    const real dx = means2D[2 * g] - pfx, dy = means2D[2 * g + 1] - pfy;
                    const real *co = conic_opacity + 4 * g;
                    const real power = RC(-0.5) * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > RC(0)) continue;
                    const real Gv = R_EXP(power);
                    const real alpha = r_min(RC(0.99), co[3] * Gv);
                    if (alpha < RC(1.0) / RC(255.0)) continue;
                    Tr = Tr / (RC(1) - alpha);
                    const real dchannel_dcolor = alpha * Tr;
                    real dL_dalpha = 0;
                    real gc[3];
                    for (int ch = 0; ch < 3; ch++) {
                        const real c = rgb[3 * g + ch];
                        accum[ch] = last_alpha * last_color[ch] + (RC(1) - last_alpha) * accum[ch];
                        last_color[ch] = c;
                        dL_dalpha += (c - accum[ch]) * dp[ch];
                        gc[ch] = dchannel_dcolor * dp[ch];
                    }
                    dL_dalpha *= Tr;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (RC(1) - alpha)) * bgdot;
                    const real dL_dG = co[3] * dL_dalpha;
                    const real gdx = Gv * dx, gdy = Gv * dy;
                    const real dG_ddelx = -gdx * co[0] - gdy * co[1];
                    const real dG_ddely = -gdy * co[2] - gdx * co[1];
                    const real v[9] = {dL_dG * dG_ddelx * ddelx_dx, dL_dG * dG_ddely * ddely_dy,
                                       RC(-0.5) * gdx * dx * dL_dG, -gdx * dy * dL_dG, RC(-0.5) * gdy * dy * dL_dG,
                                       Gv * dL_dalpha, gc[0], gc[1], gc[2]};
                    real *dst[9] = {&dL_dmeans2D[2 * g], &dL_dmeans2D[2 * g + 1], &dL_dconic_opacity[4 * g],
                                    &dL_dconic_opacity[4 * g + 1], &dL_dconic_opacity[4 * g + 2],
                                    &dL_dconic_opacity[4 * g + 3], &dL_drgb[3 * g], &dL_drgb[3 * g + 1], &dL_drgb[3 * g + 2]};
                    for (int q = 0; q < 9; q++) {
#pragma omp atomic
                        *dst[q] += v[q];
                    }
                }
            }
    }
}

/* ------------------------------------------------------------------------------------------
 * Per-strip loss of the live path: /root/reference/gaussian_renderer/loss_distribution.py:2536-2585
 * with utils/loss_utils.py:88-132 (masked L1 + 11x11 sigma=1.5 SSIM, zero padding at the strip
 * edges) and the batch formula of loss_distribution.py:2627-2629.
 * img, gt: (3, rows, W) for the strip; n_pixels_total = H*W of the FULL image
 * (utils.get_num_pixels()).  Outputs Ll1, ssim (both already divided by 3*n_pixels_total) and
 * dL/dimg for loss = (1-lambda)*Ll1 + lambda*(1-ssim).
 * ------------------------------------------------------------------------------------------ */
static void conv11(const real *src, real *dst, int rows, int W, const real *g) {
    real *tmp = (real *)malloc(sizeof(real) * (size_t)rows * W);
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < W; x++) {
            real acc = 0;
            for (int k = -5; k <= 5; k++) { int xx = x + k; if (xx >= 0 && xx < W) acc += g[k + 5] * src[(size_t)y * W + xx]; }
            tmp[(size_t)y * W + x] = acc;
        }
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < W; x++) {
            real acc = 0;
            for (int k = -5; k <= 5; k++) { int yy = y + k; if (yy >= 0 && yy < rows) acc += g[k + 5] * tmp[(size_t)yy * W + x]; }
            dst[(size_t)y * W + x] = acc;
        }
    free(tmp);
}

void gso_loss(int rows, int W, double n_pixels_total, real lambda_dssim, const real *img, const real *gt, real *out_l1,
              real *out_ssim, real *dL_dimg) {
    real g[11];
    { double s = 0, e[11]; for (int k = 0; k < 11; k++) { e[k] = exp(-((k - 5) * (k - 5)) / (2.0 * 1.5 * 1.5)); s += e[k]; }
      for (int k = 0; k < 11; k++) g[k] = (real)(float)(e[k] / s); }
    const size_t n = (size_t)rows * W;
    const real C1 = RC(0.01) * RC(0.01), C2 = RC(0.03) * RC(0.03);
    const real norm = (real)(1.0 / (3.0 * n_pixels_total));
    double l1 = 0, ss = 0;
    real *buf = (real *)malloc(sizeof(real) * n * 10);
    real *mu1 = buf, *mu2 = buf + n, *e11 = buf + 2 * n, *e22 = buf + 3 * n, *e12 = buf + 4 * n, *p = buf + 5 * n,
         *dm1 = buf + 6 * n, *d11 = buf + 7 * n, *d12 = buf + 8 * n, *tmp = buf + 9 * n;
    for (int ch = 0; ch < 3; ch++) {
        const real *x = img + ch * n, *y = gt + ch * n;
        real *gout = dL_dimg + ch * n;
        conv11(x, mu1, rows, W, g); conv11(y, mu2, rows, W, g);
        for (size_t k = 0; k < n; k++) p[k] = x[k] * x[k];
        conv11(p, e11, rows, W, g);
        for (size_t k = 0; k < n; k++) p[k] = y[k] * y[k];
        conv11(p, e22, rows, W, g);
        for (size_t k = 0; k < n; k++) p[k] = x[k] * y[k];
        conv11(p, e12, rows, W, g);
        for (size_t k = 0; k < n; k++) {
            const real m1 = mu1[k], m2 = mu2[k];
            const real s1 = e11[k] - m1 * m1, s2 = e22[k] - m2 * m2, s12 = e12[k] - m1 * m2;
            const real A = RC(2) * m1 * m2 + C1, B = RC(2) * s12 + C2, Cc = m1 * m1 + m2 * m2 + C1, Dd = s1 + s2 + C2;
            const real map = (A * B) / (Cc * Dd);
            ss += (double)map;
            l1 += (double)R_FABS(x[k] - y[k]);
            /* d map / d(mu1, E[x^2], E[xy]) */
            const real dmap_dm1 = (RC(2) * m2 * B) / (Cc * Dd) - (A * B * RC(2) * m1) / (Cc * Cc * Dd)
                                  + (A * (RC(-2) * m2)) / (Cc * Dd) * RC(1) /* dB/dm1 = -2 m2 */
                                  - (A * B) / (Cc * Dd * Dd) * (RC(-2) * m1) /* dD/dm1 = -2 m1 */;
            const real dmap_d11 = -(A * B) / (Cc * Dd * Dd);
            const real dmap_d12 = (RC(2) * A) / (Cc * Dd);
            dm1[k] = dmap_dm1; d11[k] = dmap_d11; d12[k] = dmap_d12;
        }
        /* the window is symmetric, so the adjoint of each zero-padded convolution is itself */
        conv11(dm1, tmp, rows, W, g);
        for (size_t k = 0; k < n; k++) gout[k] = tmp[k];
        conv11(d11, tmp, rows, W, g);
        for (size_t k = 0; k < n; k++) gout[k] += RC(2) * x[k] * tmp[k];
        conv11(d12, tmp, rows, W, g);
        for (size_t k = 0; k < n; k++) gout[k] += y[k] * tmp[k];
        for (size_t k = 0; k < n; k++) {
            const real d = x[k] - y[k];
            const real sgn = d > 0 ? RC(1) : (d < 0 ? RC(-1) : RC(0));
            gout[k] = norm * ((RC(1) - lambda_dssim) * sgn - lambda_dssim * gout[k]);
        }
    }
    free(buf);
    *out_l1 = (real)(l1 * (double)norm);
    *out_ssim = (real)(ss * (double)norm);
}
