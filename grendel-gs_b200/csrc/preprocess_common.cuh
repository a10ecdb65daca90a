// Device helpers shared by preprocess.cu and preprocess_raw.cu (both compiled with -fmad=false; see
// preprocess.cu for why).  SH basis, quaternion -> rotation, covariance and the EWA projection of one splat.
#pragma once
#include "common.cuh"

#define PP_THREADS 128
#define SH_FLOATS 48

static __device__ __constant__ float c_SH_C0 = 0.28209479177387814f;
static __device__ __constant__ float c_SH_C1 = 0.4886025119029199f;
static __device__ __constant__ float c_SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                            -1.0925484305920792f, 0.5462742152960396f};
static __device__ __constant__ float c_SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                            0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                            -0.5900435899266435f};

struct Cam {
    float V[16];
    float PM[16];
    float cp[3];
};

GS_D void load_cam(Cam &c, const float *viewmatrix, const float *projmatrix, const float *campos) {
#pragma unroll
    for (int k = 0; k < 16; k++) { c.V[k] = __ldg(viewmatrix + k); c.PM[k] = __ldg(projmatrix + k); }
    c.cp[0] = __ldg(campos); c.cp[1] = __ldg(campos + 1); c.cp[2] = __ldg(campos + 2);
}

GS_D void quat_to_R(const float4 q, float R[9]) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - r * z); R[2] = 2.f * (x * z + r * y);
    R[3] = 2.f * (x * y + r * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - r * x);
    R[6] = 2.f * (x * z - r * y); R[7] = 2.f * (y * z + r * x); R[8] = 1.f - 2.f * (x * x + y * y);
}

// Sigma = (R S)(R S)^T; L = R diag(mod*s) row-major, S6 = xx,xy,xz,yy,yz,zz
GS_D void cov3d_from(const float3 sc, float mod, const float4 q, float L[9], float S[6]) {
    float R[9];
    quat_to_R(q, R);
    const float s0 = mod * sc.x, s1 = mod * sc.y, s2 = mod * sc.z;
    L[0] = R[0] * s0; L[1] = R[1] * s1; L[2] = R[2] * s2;
    L[3] = R[3] * s0; L[4] = R[4] * s1; L[5] = R[5] * s2;
    L[6] = R[6] * s0; L[7] = R[7] * s1; L[8] = R[8] * s2;
    S[0] = L[0] * L[0] + L[1] * L[1] + L[2] * L[2];
    S[1] = L[0] * L[3] + L[1] * L[4] + L[2] * L[5];
    S[2] = L[0] * L[6] + L[1] * L[7] + L[2] * L[8];
    S[3] = L[3] * L[3] + L[4] * L[4] + L[5] * L[5];
    S[4] = L[3] * L[6] + L[4] * L[7] + L[5] * L[8];
    S[5] = L[6] * L[6] + L[7] * L[7] + L[8] * L[8];
}

// SH basis of /root/reference/utils/sh_utils.py:57-120
GS_D void sh_basis(int deg, float x, float y, float z, float b[16]) {
    b[0] = c_SH_C0;
#pragma unroll
    for (int k = 1; k < 16; k++) b[k] = 0.f;
    if (deg > 0) {
        b[1] = -c_SH_C1 * y; b[2] = c_SH_C1 * z; b[3] = -c_SH_C1 * x;
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = c_SH_C2[0] * xy; b[5] = c_SH_C2[1] * yz; b[6] = c_SH_C2[2] * (2.f * zz - xx - yy);
            b[7] = c_SH_C2[3] * xz; b[8] = c_SH_C2[4] * (xx - yy);
            if (deg > 2) {
                b[9] = c_SH_C3[0] * y * (3.f * xx - yy);
                b[10] = c_SH_C3[1] * xy * z;
                b[11] = c_SH_C3[2] * y * (4.f * zz - xx - yy);
                b[12] = c_SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                b[13] = c_SH_C3[4] * x * (4.f * zz - xx - yy);
                b[14] = c_SH_C3[5] * z * (xx - yy);
                b[15] = c_SH_C3[6] * x * (xx - 3.f * yy);
            }
        }
    }
}

// Projection of one splat; returns false when culled. Shared by forward and backward so both see
// identical intermediates.
struct Proj {
    float tx, ty, tz;          // view space
    float hx, hy, hw, pw;      // clip space and 1/(w+eps)
    float L[9], S[6];
    float cx, cy, xmul, ymul;  // guard-band clamped view x,y and their gradient gates
    float J00, J02, J11, J12;
    float T0[3], T1[3], u0[3], u1[3];
    float a, b, c, det;
};

// COV_READY: o.L / o.S (camera independent) were filled by the caller -- the batched kernels compute them once per
// splat and project it into B cameras.
template <bool COV_READY = false>
GS_D bool project(const Cam &cam, const float3 p, const float3 sc, float mod, const float4 q, float fx, float fy,
                  float tanfovx, float tanfovy, Proj &o) {
    const float *V = cam.V, *PM = cam.PM;
    o.tx = V[0] * p.x + V[4] * p.y + V[8] * p.z + V[12];
    o.ty = V[1] * p.x + V[5] * p.y + V[9] * p.z + V[13];
    o.tz = V[2] * p.x + V[6] * p.y + V[10] * p.z + V[14];
    if (o.tz <= 0.2f) return false;
    o.hx = PM[0] * p.x + PM[4] * p.y + PM[8] * p.z + PM[12];
    o.hy = PM[1] * p.x + PM[5] * p.y + PM[9] * p.z + PM[13];
    o.hw = PM[3] * p.x + PM[7] * p.y + PM[11] * p.z + PM[15];
    o.pw = 1.0f / (o.hw + 0.0000001f);
    if (!COV_READY) cov3d_from(sc, mod, q, o.L, o.S);
    const float limx = 1.3f * tanfovx, limy = 1.3f * tanfovy;
    const float txtz = o.tx / o.tz, tytz = o.ty / o.tz;
    o.cx = fminf(limx, fmaxf(-limx, txtz)) * o.tz;
    o.cy = fminf(limy, fmaxf(-limy, tytz)) * o.tz;
    o.xmul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    o.ymul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    o.J00 = fx / o.tz; o.J02 = -(fx * o.cx) / (o.tz * o.tz);
    o.J11 = fy / o.tz; o.J12 = -(fy * o.cy) / (o.tz * o.tz);
    o.T0[0] = o.J00 * V[0] + o.J02 * V[2]; o.T0[1] = o.J00 * V[4] + o.J02 * V[6]; o.T0[2] = o.J00 * V[8] + o.J02 * V[10];
    o.T1[0] = o.J11 * V[1] + o.J12 * V[2]; o.T1[1] = o.J11 * V[5] + o.J12 * V[6]; o.T1[2] = o.J11 * V[9] + o.J12 * V[10];
    const float *S = o.S;
    o.u0[0] = S[0] * o.T0[0] + S[1] * o.T0[1] + S[2] * o.T0[2];
    o.u0[1] = S[1] * o.T0[0] + S[3] * o.T0[1] + S[4] * o.T0[2];
    o.u0[2] = S[2] * o.T0[0] + S[4] * o.T0[1] + S[5] * o.T0[2];
    o.u1[0] = S[0] * o.T1[0] + S[1] * o.T1[1] + S[2] * o.T1[2];
    o.u1[1] = S[1] * o.T1[0] + S[3] * o.T1[1] + S[4] * o.T1[2];
    o.u1[2] = S[2] * o.T1[0] + S[4] * o.T1[1] + S[5] * o.T1[2];
    o.a = o.T0[0] * o.u0[0] + o.T0[1] * o.u0[1] + o.T0[2] * o.u0[2] + 0.3f;
    o.b = o.T0[0] * o.u1[0] + o.T0[1] * o.u1[1] + o.T0[2] * o.u1[2];
    o.c = o.T1[0] * o.u1[0] + o.T1[1] * o.u1[1] + o.T1[2] * o.u1[2] + 0.3f;
    o.det = o.a * o.c - o.b * o.b;
    return o.det != 0.f;
}

