// Batched fused-activation projection: ALL B cameras of a step in one launch (forward and backward).
//
// The reference loops over the cameras of a batch in Python and calls the operator once per camera
// (/root/reference/gaussian_renderer/__init__.py:919-963), re-reading every Gaussian's 236-byte parameter record
// B times forward and B times backward.  Here a thread loads its Gaussian ONCE (activations, 3D covariance, the
// SH block via TMA), then loops over the B cameras in registers; the backward accumulates the B cameras'
// contributions in registers / shared memory and writes each parameter gradient once.  HBM traffic per step drops
// from B x (236 + 45) to 236 + 45 B bytes per Gaussian forward (and likewise backward), and the per-camera Python /
// launch overhead -- the scaling limiter at 8 GPUs x 8 views -- disappears.
//
// Compiled -fmad=false like preprocess.cu: per camera the operation sequence is exactly that of the single-camera
// kernels, so radii / tile rectangles / depth keys stay bit-exact with the oracle.
//
// cams: (B, 40) floats per camera: viewmatrix[16], projmatrix[16] (transposed storage, scene/cameras.py:84-99),
// campos[3], tanfovx, tanfovy, 3 pad.
#include "preprocess_common.cuh"

#define DC_FLOATS 3
#define REST_FLOATS 45
#define CAM_FLOATS 40
#define PB_FWD_THREADS 128
#define PB_BWD_THREADS 64

GS_D float sigmoidf_b(float x) { return 1.0f / (1.0f + expf(-x)); }

GS_D float4 normalize4b(const float4 r, float &denom) {
    const float n = sqrtf(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w);
    denom = fmaxf(n, 1e-12f);
    return make_float4(r.x / denom, r.y / denom, r.z / denom, r.w / denom);
}

GS_D void load_cam_row(Cam &c, float &tanfovx, float &tanfovy, const float *__restrict__ row) {
#pragma unroll
    for (int k = 0; k < 16; k++) { c.V[k] = __ldg(row + k); c.PM[k] = __ldg(row + 16 + k); }
    c.cp[0] = __ldg(row + 32); c.cp[1] = __ldg(row + 33); c.cp[2] = __ldg(row + 34);
    tanfovx = __ldg(row + 35); tanfovy = __ldg(row + 36);
}

GS_D float &sh_slot(float *s_dc, float *s_rest, int t, int f) {
    return f < DC_FLOATS ? s_dc[t * DC_FLOATS + f] : s_rest[t * REST_FLOATS + (f - DC_FLOATS)];
}

// stage the CTA's SH block (dc + rest) in shared memory: two TMA bulk copies, or plain loads for a ragged tail
template <int THREADS>
GS_D bool stage_sh(float *s_dc, float *s_rest, uint64_t *bar, const float *f_dc, const float *f_rest, int base, int nvalid) {
    const bool tma_ok = (nvalid & 3) == 0;
    if (threadIdx.x == 0) {
        gs_mbar_init(bar, 1);
        gs_fence_mbar_init();
    }
    __syncthreads();
    if (tma_ok) {
        if (threadIdx.x == 0) {
            const uint32_t b_dc = (uint32_t)nvalid * DC_FLOATS * 4u, b_rest = (uint32_t)nvalid * REST_FLOATS * 4u;
            gs_mbar_arrive_expect_tx(bar, b_dc + b_rest);
            gs_bulk_g2s(s_dc, f_dc + (size_t)base * DC_FLOATS, b_dc, bar);
            gs_bulk_g2s(s_rest, f_rest + (size_t)base * REST_FLOATS, b_rest, bar);
        }
    } else {
        for (int k = threadIdx.x; k < nvalid * DC_FLOATS; k += THREADS) s_dc[k] = f_dc[(size_t)base * DC_FLOATS + k];
        for (int k = threadIdx.x; k < nvalid * REST_FLOATS; k += THREADS) s_rest[k] = f_rest[(size_t)base * REST_FLOATS + k];
    }
    return tma_ok;
}

__global__ void __launch_bounds__(PB_FWD_THREADS)
k_preprocess_fwd_batched(int B, int P, int D, const float *__restrict__ xyz, const float *__restrict__ f_dc,
                         const float *__restrict__ f_rest, const float *__restrict__ scaling, float mod,
                         const float *__restrict__ rotation, const float *__restrict__ opacity,
                         const float *__restrict__ cams, int W, int H, float *__restrict__ means2D,
                         float *__restrict__ depths, int32_t *__restrict__ radii, float *__restrict__ conic_opacity,
                         float *__restrict__ rgb, uint8_t *__restrict__ clamped) {
    __shared__ __align__(128) float s_sh[PB_FWD_THREADS * SH_FLOATS];
    __shared__ __align__(8) uint64_t s_bar;
    float *s_dc = s_sh, *s_rest = s_sh + PB_FWD_THREADS * DC_FLOATS;
    const int base = blockIdx.x * PB_FWD_THREADS;
    const int nvalid = min(PB_FWD_THREADS, P - base);
    const bool tma_ok = stage_sh<PB_FWD_THREADS>(s_dc, s_rest, &s_bar, f_dc, f_rest, base, nvalid);
    const int i = base + threadIdx.x;
    const int gx = (W + GS_BLOCK_X - 1) / GS_BLOCK_X, gy = (H + GS_BLOCK_Y - 1) / GS_BLOCK_Y;
    float3 p = make_float3(0.f, 0.f, 0.f), sc = p;
    float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
    float op = 0.f;
    Proj pr;
    if (i < P) {
        p = make_float3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        sc = make_float3(expf(scaling[3 * i]), expf(scaling[3 * i + 1]), expf(scaling[3 * i + 2]));
        float den;
        q = normalize4b(*reinterpret_cast<const float4 *>(rotation + 4 * i), den);
        op = sigmoidf_b(opacity[i]);
        cov3d_from(sc, mod, q, pr.L, pr.S);  // camera independent
    }
    if (tma_ok) gs_mbar_wait(&s_bar, 0); else __syncthreads();
    if (i >= P) return;
    for (int k = 0; k < B; k++) {
        Cam cam;
        float tanfovx, tanfovy;
        load_cam_row(cam, tanfovx, tanfovy, cams + (size_t)k * CAM_FLOATS);
        float ix = 0.f, iy = 0.f, depth = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
        int rad = 0;
        uint8_t cm = 0;
        float4 co = make_float4(0.f, 0.f, 0.f, 0.f);
        const float fx = (float)W / (2.f * tanfovx), fy = (float)H / (2.f * tanfovy);
        if (project<true>(cam, p, sc, mod, q, fx, fy, tanfovx, tanfovy, pr)) {
            const float det_inv = 1.0f / pr.det;
            const float mid = 0.5f * (pr.a + pr.c);
            const float disc = sqrtf(fmaxf(0.1f, mid * mid - pr.det));
            const float lam = fmaxf(mid + disc, mid - disc);
            const int r_ = (int)ceilf(3.f * sqrtf(lam));
            const float ndcx = pr.hx * pr.pw, ndcy = pr.hy * pr.pw;
            const float ix_ = ((ndcx + 1.f) * (float)W - 1.f) * 0.5f;
            const float iy_ = ((ndcy + 1.f) * (float)H - 1.f) * 0.5f;
            int x0, y0, x1, y1;
            gs_get_rect(ix_, iy_, r_, gx, gy, x0, y0, x1, y1);
            if ((x1 - x0) * (y1 - y0) != 0) {
                ix = ix_; iy = iy_; rad = r_; depth = pr.tz;
                co = make_float4(pr.c * det_inv, -pr.b * det_inv, pr.a * det_inv, op);
                float dx = p.x - cam.cp[0], dy = p.y - cam.cp[1], dz = p.z - cam.cp[2];
                const float len = sqrtf(dx * dx + dy * dy + dz * dz);
                dx = dx / len; dy = dy / len; dz = dz / len;
                float bas[16];
                sh_basis(D, dx, dy, dz, bas);
                float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int f = 0; f < SH_FLOATS; f++) acc[f % 3] += bas[f / 3] * sh_slot(s_dc, s_rest, threadIdx.x, f);
                c0 = acc[0] + 0.5f; c1 = acc[1] + 0.5f; c2 = acc[2] + 0.5f;
                if (c0 < 0.f) { cm |= 1; c0 = 0.f; }
                if (c1 < 0.f) { cm |= 2; c1 = 0.f; }
                if (c2 < 0.f) { cm |= 4; c2 = 0.f; }
            }
        }
        const size_t o = (size_t)k * P + i;
        *reinterpret_cast<float2 *>(means2D + 2 * o) = make_float2(ix, iy);
        depths[o] = depth;
        radii[o] = rad;
        *reinterpret_cast<float4 *>(conic_opacity + 4 * o) = co;
        rgb[3 * o] = c0; rgb[3 * o + 1] = c1; rgb[3 * o + 2] = c2;
        clamped[o] = cm;
    }
}

__global__ void __launch_bounds__(PB_BWD_THREADS)
k_preprocess_bwd_batched(int B, int P, int D, const float *__restrict__ xyz, const float *__restrict__ f_dc,
                         const float *__restrict__ f_rest, const float *__restrict__ scaling, float mod,
                         const float *__restrict__ rotation, const float *__restrict__ opacity,
                         const float *__restrict__ cams, int W, int H, const int32_t *__restrict__ radii,
                         const uint8_t *__restrict__ clamped, const float *__restrict__ g_means2D,
                         const float *__restrict__ g_conic_opacity, const float *__restrict__ g_rgb,
                         float *__restrict__ d_xyz, float *__restrict__ d_dc, float *__restrict__ d_rest,
                         float *__restrict__ d_scaling, float *__restrict__ d_rotation, float *__restrict__ d_opacity) {
    __shared__ __align__(128) float s_sh[PB_BWD_THREADS * SH_FLOATS];   // SH coefficients of the CTA's splats
    __shared__ __align__(128) float s_g[PB_BWD_THREADS * SH_FLOATS];    // dL/dSH accumulated over the cameras
    __shared__ __align__(8) uint64_t s_bar;
    float *s_dc = s_sh, *s_rest = s_sh + PB_BWD_THREADS * DC_FLOATS;
    float *g_dc = s_g, *g_rest = s_g + PB_BWD_THREADS * DC_FLOATS;
    const int base = blockIdx.x * PB_BWD_THREADS;
    const int nvalid = min(PB_BWD_THREADS, P - base);
    const bool tma_ok = stage_sh<PB_BWD_THREADS>(s_dc, s_rest, &s_bar, f_dc, f_rest, base, nvalid);
    const int t = threadIdx.x;
    const int i = base + t;
#pragma unroll
    for (int f = 0; f < SH_FLOATS; f++) sh_slot(g_dc, g_rest, t, f) = 0.f;
    float3 p = make_float3(0.f, 0.f, 0.f), sc = p;
    float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
    float den = 1.f, op = 0.f;
    Proj pr;
    if (i < P) {
        p = make_float3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        sc = make_float3(expf(scaling[3 * i]), expf(scaling[3 * i + 1]), expf(scaling[3 * i + 2]));
        q = normalize4b(*reinterpret_cast<const float4 *>(rotation + 4 * i), den);
        op = sigmoidf_b(opacity[i]);
        cov3d_from(sc, mod, q, pr.L, pr.S);
    }
    if (tma_ok) gs_mbar_wait(&s_bar, 0); else __syncthreads();
    float gmx = 0.f, gmy = 0.f, gmz = 0.f, gop = 0.f;
    float G[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // symmetric dL/dSigma: xx, xy, xz, yy, yz, zz
    if (i < P) {
        for (int k = 0; k < B; k++) {
            const size_t o = (size_t)k * P + i;
            if (radii[o] <= 0) continue;
            Cam cam;
            float tanfovx, tanfovy;
            load_cam_row(cam, tanfovx, tanfovy, cams + (size_t)k * CAM_FLOATS);
            const float fx = (float)W / (2.f * tanfovx), fy = (float)H / (2.f * tanfovy);
            project<true>(cam, p, sc, mod, q, fx, fy, tanfovx, tanfovy, pr);
            const float *V = cam.V, *PM = cam.PM;
            const float4 gco = *reinterpret_cast<const float4 *>(g_conic_opacity + 4 * o);
            gop += gco.w;
            const float a = pr.a, b = pr.b, c = pr.c, det = pr.det;
            float dLda = 0.f, dLdb = 0.f, dLdc = 0.f;
            if (det != 0.f) {
                const float d2 = 1.0f / (det * det);
                dLda = d2 * (-c * c * gco.x + b * c * gco.y - b * b * gco.z);
                dLdb = d2 * (2.f * b * c * gco.x - (det + 2.f * b * b) * gco.y + 2.f * a * b * gco.z);
                dLdc = d2 * (-b * b * gco.x + a * b * gco.y - a * a * gco.z);
            }
            float dT0[3], dT1[3];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                dT0[r] = 2.f * dLda * pr.u0[r] + dLdb * pr.u1[r];
                dT1[r] = 2.f * dLdc * pr.u1[r] + dLdb * pr.u0[r];
            }
            const float dJ00 = dT0[0] * V[0] + dT0[1] * V[4] + dT0[2] * V[8];
            const float dJ02 = dT0[0] * V[2] + dT0[1] * V[6] + dT0[2] * V[10];
            const float dJ11 = dT1[0] * V[1] + dT1[1] * V[5] + dT1[2] * V[9];
            const float dJ12 = dT1[0] * V[2] + dT1[1] * V[6] + dT1[2] * V[10];
            const float tzi = 1.0f / pr.tz, tzi2 = tzi * tzi, tzi3 = tzi2 * tzi;
            const float dtx = pr.xmul * (-fx * tzi2) * dJ02;
            const float dty = pr.ymul * (-fy * tzi2) * dJ12;
            const float dtz = -fx * tzi2 * dJ00 - fy * tzi2 * dJ11 + 2.f * fx * pr.cx * tzi3 * dJ02 +
                              2.f * fy * pr.cy * tzi3 * dJ12;
            gmx += V[0] * dtx + V[1] * dty + V[2] * dtz;
            gmy += V[4] * dtx + V[5] * dty + V[6] * dtz;
            gmz += V[8] * dtx + V[9] * dty + V[10] * dtz;
            {
                const float2 g2 = *reinterpret_cast<const float2 *>(g_means2D + 2 * o);
                const float dhx = g2.x * pr.pw, dhy = g2.y * pr.pw;
                const float dhw = -(g2.x * pr.hx + g2.y * pr.hy) * pr.pw * pr.pw;
                gmx += PM[0] * dhx + PM[1] * dhy + PM[3] * dhw;
                gmy += PM[4] * dhx + PM[5] * dhy + PM[7] * dhw;
                gmz += PM[8] * dhx + PM[9] * dhy + PM[11] * dhw;
            }
            // dL/dSigma of this camera (symmetric), accumulated; Sigma -> (scale, rotation) runs once after the loop
            {
                const int rr[6] = {0, 0, 0, 1, 1, 2}, ss[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
                for (int e = 0; e < 6; e++) {
                    const int r = rr[e], s = ss[e];
                    G[e] += pr.T0[r] * pr.T0[s] * dLda + 0.5f * (pr.T0[r] * pr.T1[s] + pr.T0[s] * pr.T1[r]) * dLdb +
                            pr.T1[r] * pr.T1[s] * dLdc;
                }
            }
            // colour -> SH coefficients and view direction
            {
                float vx = p.x - cam.cp[0], vy = p.y - cam.cp[1], vz = p.z - cam.cp[2];
                const float len = sqrtf(vx * vx + vy * vy + vz * vz);
                const float x = vx / len, y = vy / len, z = vz / len;
                float bas[16];
                sh_basis(D, x, y, z, bas);
                const uint8_t cm = clamped[o];
                float dc[3];
                dc[0] = (cm & 1) ? 0.f : g_rgb[3 * o];
                dc[1] = (cm & 2) ? 0.f : g_rgb[3 * o + 1];
                dc[2] = (cm & 4) ? 0.f : g_rgb[3 * o + 2];
                const int ncoef = (D + 1) * (D + 1);
                float s[16];
#pragma unroll
                for (int kk = 0; kk < 16; kk++) s[kk] = 0.f;
#pragma unroll
                for (int f = 0; f < SH_FLOATS; f++) {
                    const int kk = f / 3, ch = f % 3;
                    if (kk < ncoef) {
                        s[kk] += sh_slot(s_dc, s_rest, t, f) * dc[ch];
                        sh_slot(g_dc, g_rest, t, f) += bas[kk] * dc[ch];
                    }
                }
                float ddx = 0.f, ddy = 0.f, ddz = 0.f;
                if (D > 0) {
                    ddx += -c_SH_C1 * s[3]; ddy += -c_SH_C1 * s[1]; ddz += c_SH_C1 * s[2];
                    if (D > 1) {
                        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                        ddx += c_SH_C2[0] * y * s[4] + c_SH_C2[2] * 2.f * -x * s[6] + c_SH_C2[3] * z * s[7] +
                               c_SH_C2[4] * 2.f * x * s[8];
                        ddy += c_SH_C2[0] * x * s[4] + c_SH_C2[1] * z * s[5] + c_SH_C2[2] * 2.f * -y * s[6] +
                               c_SH_C2[4] * 2.f * -y * s[8];
                        ddz += c_SH_C2[1] * y * s[5] + c_SH_C2[2] * 4.f * z * s[6] + c_SH_C2[3] * x * s[7];
                        if (D > 2) {
                            ddx += c_SH_C3[0] * s[9] * 6.f * xy + c_SH_C3[1] * s[10] * yz + c_SH_C3[2] * s[11] * -2.f * xy +
                                   c_SH_C3[3] * s[12] * -6.f * xz + c_SH_C3[4] * s[13] * (-3.f * xx + 4.f * zz - yy) +
                                   c_SH_C3[5] * s[14] * 2.f * xz + c_SH_C3[6] * s[15] * 3.f * (xx - yy);
                            ddy += c_SH_C3[0] * s[9] * 3.f * (xx - yy) + c_SH_C3[1] * s[10] * xz +
                                   c_SH_C3[2] * s[11] * (-3.f * yy + 4.f * zz - xx) + c_SH_C3[3] * s[12] * -6.f * yz +
                                   c_SH_C3[4] * s[13] * -2.f * xy + c_SH_C3[5] * s[14] * -2.f * yz +
                                   c_SH_C3[6] * s[15] * -6.f * xy;
                            ddz += c_SH_C3[1] * s[10] * xy + c_SH_C3[2] * s[11] * 8.f * yz +
                                   c_SH_C3[3] * s[12] * 3.f * (2.f * zz - xx - yy) + c_SH_C3[4] * s[13] * 8.f * xz +
                                   c_SH_C3[5] * s[14] * (xx - yy);
                        }
                    }
                }
                const float dot = x * ddx + y * ddy + z * ddz;
                gmx += (ddx - x * dot) / len;
                gmy += (ddy - y * dot) / len;
                gmz += (ddz - z * dot) / len;
            }
        }
        // Sigma = L L^T, L = R diag(mod*s): once, with the accumulated dL/dSigma
        float gs0, gs1, gs2;
        float4 gq;
        {
            const float Gm[3][3] = {{G[0], G[1], G[2]}, {G[1], G[3], G[4]}, {G[2], G[4], G[5]}};
            float R[9];
            quat_to_R(q, R);
            const float s[3] = {mod * sc.x, mod * sc.y, mod * sc.z};
            float dLm[9];
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int k = 0; k < 3; k++)
                    dLm[3 * r + k] = 2.f * (Gm[r][0] * pr.L[k] + Gm[r][1] * pr.L[3 + k] + Gm[r][2] * pr.L[6 + k]);
            float dR[9];
            float gsk[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                gsk[k] = mod * (dLm[k] * R[k] + dLm[3 + k] * R[3 + k] + dLm[6 + k] * R[6 + k]);
                dR[k] = dLm[k] * s[k]; dR[3 + k] = dLm[3 + k] * s[k]; dR[6 + k] = dLm[6 + k] * s[k];
            }
            gs0 = gsk[0] * sc.x; gs1 = gsk[1] * sc.y; gs2 = gsk[2] * sc.z;  // exp'
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            float4 dq;
            dq.x = 2.f * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
            dq.y = 2.f * (y * dR[1] + z * dR[2] + y * dR[3] - 2.f * x * dR[4] - r * dR[5] + z * dR[6] + r * dR[7] -
                          2.f * x * dR[8]);
            dq.z = 2.f * (-2.f * y * dR[0] + x * dR[1] + r * dR[2] + x * dR[3] + z * dR[5] - r * dR[6] + z * dR[7] -
                          2.f * y * dR[8]);
            dq.w = 2.f * (-2.f * z * dR[0] - r * dR[1] + x * dR[2] + r * dR[3] - 2.f * z * dR[4] + y * dR[5] +
                          x * dR[6] + y * dR[7]);
            const float dot = q.x * dq.x + q.y * dq.y + q.z * dq.z + q.w * dq.w;  // normalize'
            gq = make_float4((dq.x - q.x * dot) / den, (dq.y - q.y * dot) / den, (dq.z - q.z * dot) / den,
                             (dq.w - q.w * dot) / den);
        }
        d_xyz[3 * i] = gmx; d_xyz[3 * i + 1] = gmy; d_xyz[3 * i + 2] = gmz;
        d_scaling[3 * i] = gs0; d_scaling[3 * i + 1] = gs1; d_scaling[3 * i + 2] = gs2;
        *reinterpret_cast<float4 *>(d_rotation + 4 * i) = gq;
        d_opacity[i] = gop * op * (1.f - op);  // sigmoid'
    }
    gs_fence_proxy_async_smem();
    __syncthreads();
    if (tma_ok) {
        if (threadIdx.x == 0) {
            gs_bulk_s2g(d_dc + (size_t)base * DC_FLOATS, g_dc, (uint32_t)nvalid * DC_FLOATS * 4u);
            gs_bulk_s2g(d_rest + (size_t)base * REST_FLOATS, g_rest, (uint32_t)nvalid * REST_FLOATS * 4u);
            gs_bulk_commit();
            gs_bulk_wait_read0();
        }
    } else {
        for (int k = threadIdx.x; k < nvalid * DC_FLOATS; k += PB_BWD_THREADS) d_dc[(size_t)base * DC_FLOATS + k] = g_dc[k];
        for (int k = threadIdx.x; k < nvalid * REST_FLOATS; k += PB_BWD_THREADS) d_rest[(size_t)base * REST_FLOATS + k] = g_rest[k];
    }
}

#define PB_MAX_CAMS 64

extern "C" int gs_preprocess_forward_batched(int B, int P, int sh_degree, const float *xyz, const float *features_dc,
                                             const float *features_rest, const float *scaling, float scale_modifier,
                                             const float *rotation, const float *opacity, const float *cams,
                                             int image_width, int image_height, float *means2D, float *depths,
                                             int32_t *radii, float *conic_opacity, float *rgb, uint8_t *clamped,
                                             void *stream) {
    GS_REQUIRE(B > 0 && B <= PB_MAX_CAMS && P >= 0, "sizes");
    GS_REQUIRE(sh_degree >= 0 && sh_degree <= 3, "sh_degree must be 0..3");
    if (P == 0) return GS_OK;
    GS_REQUIRE(image_width > 0 && image_height > 0, "image size");
    GS_REQUIRE(xyz && features_dc && features_rest && scaling && rotation && opacity && cams && means2D && depths &&
                   radii && conic_opacity && rgb && clamped, "null pointer");
    GS_REQUIRE(((uintptr_t)features_dc & 15) == 0 && ((uintptr_t)features_rest & 15) == 0 &&
                   ((uintptr_t)rotation & 15) == 0 && ((uintptr_t)conic_opacity & 15) == 0 && ((uintptr_t)means2D & 7) == 0,
               "16-byte alignment");
    const int grid = (P + PB_FWD_THREADS - 1) / PB_FWD_THREADS;
    GsStageTimer timer(GS_STAGE_PREPROCESS_FWD, (cudaStream_t)stream);
    k_preprocess_fwd_batched<<<grid, PB_FWD_THREADS, 0, (cudaStream_t)stream>>>(
        B, P, sh_degree, xyz, features_dc, features_rest, scaling, scale_modifier, rotation, opacity, cams, image_width,
        image_height, means2D, depths, radii, conic_opacity, rgb, clamped);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

extern "C" int gs_preprocess_backward_batched(int B, int P, int sh_degree, const float *xyz, const float *features_dc,
                                              const float *features_rest, const float *scaling, float scale_modifier,
                                              const float *rotation, const float *opacity, const float *cams,
                                              int image_width, int image_height, const int32_t *radii,
                                              const uint8_t *clamped, const float *dL_dmeans2D,
                                              const float *dL_dconic_opacity, const float *dL_drgb, float *dL_dxyz,
                                              float *dL_dfeatures_dc, float *dL_dfeatures_rest, float *dL_dscaling,
                                              float *dL_drotation, float *dL_dopacity, void *stream) {
    GS_REQUIRE(B > 0 && B <= PB_MAX_CAMS && P >= 0, "sizes");
    GS_REQUIRE(sh_degree >= 0 && sh_degree <= 3, "sh_degree must be 0..3");
    if (P == 0) return GS_OK;
    GS_REQUIRE(xyz && features_dc && features_rest && scaling && rotation && opacity && cams && radii && clamped &&
                   dL_dmeans2D && dL_dconic_opacity && dL_drgb && dL_dxyz && dL_dfeatures_dc && dL_dfeatures_rest &&
                   dL_dscaling && dL_drotation && dL_dopacity, "null pointer");
    GS_REQUIRE(((uintptr_t)features_dc & 15) == 0 && ((uintptr_t)features_rest & 15) == 0 &&
                   ((uintptr_t)dL_dfeatures_dc & 15) == 0 && ((uintptr_t)dL_dfeatures_rest & 15) == 0 &&
                   ((uintptr_t)rotation & 15) == 0 && ((uintptr_t)dL_drotation & 15) == 0 &&
                   ((uintptr_t)dL_dconic_opacity & 15) == 0 && ((uintptr_t)dL_dmeans2D & 7) == 0,
               "16-byte alignment");
    const int grid = (P + PB_BWD_THREADS - 1) / PB_BWD_THREADS;
    GsStageTimer timer(GS_STAGE_PREPROCESS_BWD, (cudaStream_t)stream);
    k_preprocess_bwd_batched<<<grid, PB_BWD_THREADS, 0, (cudaStream_t)stream>>>(
        B, P, sh_degree, xyz, features_dc, features_rest, scaling, scale_modifier, rotation, opacity, cams, image_width,
        image_height, radii, clamped, dL_dmeans2D, dL_dconic_opacity, dL_drgb, dL_dxyz, dL_dfeatures_dc,
        dL_dfeatures_rest, dL_dscaling, dL_drotation, dL_dopacity);
    GS_LAUNCH_CHECK();
    return GS_OK;
}
