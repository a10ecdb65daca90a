// Densification step in three launches: select -> scan -> gather.
// /root/reference/scene/gaussian_model.py:1005-1044 (densify_and_prune) runs densify_and_clone (:973-1003),
// densify_and_split (:922-971) and two prune_points (:816-835) as ~150 torch kernels: every step boolean-indexes or
// concatenates all six parameters and both Adam moments (cat_tensors_to_optimizer :837-881, _prune_optimizer :789-814)
// and reads a count back to the host each time.  The END STATE of that sequence is, in this order:
//     [ originals that are neither split nor pruned | clones | split children, copy 1 | split children, copy 2 ]
// (each block in index order, the final opacity / world-size prune applied to every block), with the Adam moments of
// the survivors carried over and those of new Gaussians zero.  So:
//   1. k_densify_flags: per Gaussian, the five decisions (keep, clone, child 1, child 2, "was selected for split") as
//      five byte planes -- the concatenation of the first four planes IS the output order;
//   2. one exclusive scan over the 5 P flags: a flag's rank is its output row (planes 0-3) or its index into the
//      split's block of normal draws (plane 4); the plane totals go back to the host once;
//   3. k_densify_gather: every tensor (6 parameters, 12 moments, send_to_gpui_cnt) is read once and written to its
//      output rows, with the split's two transforms applied on the fly (position: R(q) (s * z) + x, scale:
//      log(s / 1.6)).
// Byte / index work plus a few transcendentals: HBM bound, ~(3 x 236 + 4 W) B read and written per Gaussian.
#include <cub/cub.cuh>

#include "common.cuh"

#define DN_THREADS 256
#define DN_PLANES 5

struct PlaneToInt {
    const uint8_t *f;
    __host__ __device__ int32_t operator()(int e) const { return f[e]; }
};

static size_t dn_align(size_t v) { return (v + 255) / 256 * 256; }

static size_t dn_scan_bytes(int P) {
    size_t b = 0;
    const int n = DN_PLANES * (P > 0 ? P : 1);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const int32_t *)nullptr, (int32_t *)nullptr, n);
    return dn_align(b);
}

// temp layout: flags (5 P bytes) | pos (5 P int32) | totals (8 int32) | scan scratch
extern "C" size_t gs_densify_temp_bytes(int P) {
    const size_t n = (size_t)DN_PLANES * (size_t)(P > 0 ? P : 1);
    return dn_align(n) + dn_align(4 * n) + 256 + dn_scan_bytes(P) + 256;
}

__global__ void __launch_bounds__(DN_THREADS)
k_densify_flags(int P, const float *__restrict__ accum, const float *__restrict__ denom,
                const float *__restrict__ scaling, const float *__restrict__ opacity, float max_grad, float min_opacity,
                float dense_thr, float big_thr, int use_screen, uint8_t *__restrict__ flags) {
    const int i = blockIdx.x * DN_THREADS + threadIdx.x;
    if (i >= P) return;
    float grad = __fdiv_rn(accum[i], denom[i]);          // grads = xyz_gradient_accum / denom      (:1018)
    if (isnan(grad)) grad = 0.f;                          // grads[grads.isnan()] = 0.0              (:1019)
    const float s0 = expf(scaling[3 * i]), s1 = expf(scaling[3 * i + 1]), s2 = expf(scaling[3 * i + 2]);
    const float smax = fmaxf(s0, fmaxf(s1, s2));
    const bool hot = fabsf(grad) >= max_grad;             // torch.norm(grads, dim=-1) >= grad_threshold (:975-977)
    const bool sel_clone = hot && smax <= dense_thr;      // (:978-982)
    const bool sel_split = grad >= max_grad && smax > dense_thr;   // (:927-932); clones have zero padded_grad
    const float opa = __fdiv_rn(1.f, 1.f + expf(-opacity[i]));
    const bool faint = opa < min_opacity;                 // (:1026)
    const bool prune_orig = faint || (use_screen && smax > big_thr);   // (:1037-1040)
    // a child's scale is stored as log(s / (0.8 N)) (:949-951) and read back through exp (:110-111)
    const float c0 = expf(logf(__fdiv_rn(s0, 1.6f))), c1 = expf(logf(__fdiv_rn(s1, 1.6f))), c2 = expf(logf(__fdiv_rn(s2, 1.6f)));
    const bool prune_child = faint || (use_screen && fmaxf(c0, fmaxf(c1, c2)) > big_thr);
    flags[i] = (!sel_split && !prune_orig) ? 1 : 0;
    flags[(size_t)P + i] = (sel_clone && !prune_orig) ? 1 : 0;
    const uint8_t child = (sel_split && !prune_child) ? 1 : 0;
    flags[(size_t)2 * P + i] = child;
    flags[(size_t)3 * P + i] = child;
    flags[(size_t)4 * P + i] = sel_split ? 1 : 0;
}

__global__ void k_densify_totals(int P, const uint8_t *__restrict__ flags, const int32_t *__restrict__ pos,
                                 int32_t *__restrict__ totals) {
    const int c = threadIdx.x;
    if (c < DN_PLANES) totals[c] = pos[(size_t)c * P];   // first output row / first draw of plane c
    if (c == DN_PLANES) totals[DN_PLANES] = pos[(size_t)DN_PLANES * P - 1] + flags[(size_t)DN_PLANES * P - 1];
}

// counts_host (HOST, 6 int32): rows kept, clones, children copy 1, children copy 2, S = Gaussians selected for the split
// (the split consumes 2 S rows of `noise`), and the new number of Gaussians.  Synchronises `stream`.
extern "C" int gs_densify_select(int P, const float *xyz_gradient_accum, const float *denom, const float *scaling_raw,
                                 const float *opacity_raw, float max_grad, float min_opacity, float extent,
                                 float percent_dense, int use_screen_size, void *temp, size_t temp_bytes,
                                 int32_t *counts_host, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GS_REQUIRE(P > 0, "P");
    GS_REQUIRE((long long)DN_PLANES * P < (1ll << 31), "5 P overflows int32");
    GS_REQUIRE(xyz_gradient_accum && denom && scaling_raw && opacity_raw && temp && counts_host, "null pointer");
    if (temp_bytes < gs_densify_temp_bytes(P)) {
        gs_set_error("gs_densify_select: temp too small");
        return GS_ENOMEM;
    }
    const size_t n = (size_t)DN_PLANES * P;
    uint8_t *flags = (uint8_t *)temp;
    int32_t *pos = (int32_t *)((char *)temp + dn_align(n));
    int32_t *totals = (int32_t *)((char *)pos + dn_align(4 * n));
    void *scratch = (char *)totals + 256;
    size_t scratch_bytes = dn_scan_bytes(P);
    // the thresholds are Python doubles in the reference, rounded to fp32 when compared with fp32 tensors
    const float dense_thr = (float)((double)percent_dense * (double)extent), big_thr = (float)(0.1 * (double)extent);
    k_densify_flags<<<(P + DN_THREADS - 1) / DN_THREADS, DN_THREADS, 0, stream>>>(
        P, xyz_gradient_accum, denom, scaling_raw, opacity_raw, max_grad, min_opacity, dense_thr, big_thr,
        use_screen_size ? 1 : 0, flags);
    GS_LAUNCH_CHECK();
    cub::CountingInputIterator<int> idx(0);
    cub::TransformInputIterator<int32_t, PlaneToInt, cub::CountingInputIterator<int>> it(idx, PlaneToInt{flags});
    GS_CUDA_TRY(cub::DeviceScan::ExclusiveSum(scratch, scratch_bytes, it, pos, (int)n, stream));
    k_densify_totals<<<1, 32, 0, stream>>>(P, flags, pos, totals);
    GS_LAUNCH_CHECK();
    int32_t t[DN_PLANES + 1];
    GS_CUDA_TRY(cudaMemcpyAsync(t, totals, sizeof(t), cudaMemcpyDeviceToHost, stream));
    GS_CUDA_TRY(cudaStreamSynchronize(stream));
    counts_host[0] = t[1] - t[0];   // kept originals
    counts_host[1] = t[2] - t[1];   // clones
    counts_host[2] = t[3] - t[2];   // children, copy 1
    counts_host[3] = t[4] - t[3];   // children, copy 2
    counts_host[4] = t[5] - t[4];   // S
    counts_host[5] = t[4];          // new number of Gaussians
    return GS_OK;
}

#define DN_MAX_TENSORS 24
enum { DN_COPY = 0, DN_XYZ = 1, DN_SCALING = 2, DN_MOMENT = 3 };

struct DnTensors {
    const float *src[DN_MAX_TENSORS];
    float *dst[DN_MAX_TENSORS];
    int width[DN_MAX_TENSORS];
    int kind[DN_MAX_TENSORS];
};

__global__ void __launch_bounds__(DN_THREADS)
k_densify_gather(int P, int S, const DnTensors t, const float *__restrict__ scaling,
                 const float *__restrict__ rotation, const float *__restrict__ noise, const uint8_t *__restrict__ flags,
                 const int32_t *__restrict__ pos, int split_base) {
    const int k = blockIdx.y;
    const int d = t.width[k], kind = t.kind[k];
    const long long e = (long long)blockIdx.x * DN_THREADS + threadIdx.x;
    if (e >= (long long)P * d) return;
    const int i = (int)(e / d), col = (int)(e - (long long)i * d);
    const float v = t.src[k][e];
    float *dst = t.dst[k];
    if (flags[i]) dst[(size_t)pos[i] * d + col] = v;                                             // survivor: as is
    if (flags[(size_t)P + i]) dst[(size_t)pos[(size_t)P + i] * d + col] = kind == DN_MOMENT ? 0.f : v;   // clone
    if (flags[(size_t)2 * P + i]) {                                                               // two children
        const int rank = pos[(size_t)4 * P + i] - split_base;   // index among the S Gaussians selected for the split
        float c[2] = {v, v};
        if (kind == DN_MOMENT) {
            c[0] = c[1] = 0.f;
        } else if (kind == DN_SCALING) {
            c[0] = c[1] = logf(__fdiv_rn(expf(v), 1.6f));       // scaling_inverse_activation(get_scaling / (0.8 N))
        } else if (kind == DN_XYZ) {
            // new_xyz = R(q) (s * z) + xyz, R of utils/general_utils.py:416-438 on the RAW quaternion
            const float q0 = rotation[4 * i], q1 = rotation[4 * i + 1], q2 = rotation[4 * i + 2], q3 = rotation[4 * i + 3];
            const float nrm = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
            const float w = q0 / nrm, x = q1 / nrm, y = q2 / nrm, z = q3 / nrm;
            float r0, r1, r2;   // row `col` of R
            if (col == 0) { r0 = 1.f - 2.f * (y * y + z * z); r1 = 2.f * (x * y - w * z); r2 = 2.f * (x * z + w * y); }
            else if (col == 1) { r0 = 2.f * (x * y + w * z); r1 = 1.f - 2.f * (x * x + z * z); r2 = 2.f * (y * z - w * x); }
            else { r0 = 2.f * (x * z - w * y); r1 = 2.f * (y * z + w * x); r2 = 1.f - 2.f * (x * x + y * y); }
            const float s0 = expf(scaling[3 * i]), s1 = expf(scaling[3 * i + 1]), s2 = expf(scaling[3 * i + 2]);
#pragma unroll
            for (int copy = 0; copy < 2; copy++) {
                const float *zz = noise + 3 * ((size_t)copy * S + rank);   // stds.repeat(N,1): copy-major blocks of S rows
                c[copy] = (r0 * (s0 * zz[0]) + r1 * (s1 * zz[1]) + r2 * (s2 * zz[2])) + v;   // v = xyz[i][col]
            }
        }
        dst[(size_t)pos[(size_t)2 * P + i] * d + col] = c[0];
        dst[(size_t)pos[(size_t)3 * P + i] * d + col] = c[1];
    }
}

// After gs_densify_select (same temp, untouched; S and new_P are its counts[4] and counts[5]).  src_host / dst_host:
// HOST arrays of num_tensors device pointers to (P, width) inputs and (new_P, width) outputs of 4-byte elements; kind:
// 0 copy (f_dc, f_rest, opacity, rotation, send_to_gpui_cnt), 1 position, 2 log-scale, 3 Adam moment (zero for new
// Gaussians).  noise: (2 S, 3) standard-normal draws (torch.normal's role at scene/gaussian_model.py:936-938), may be
// NULL if S == 0.
extern "C" int gs_densify_gather(int P, int S, int new_P, int num_tensors, const void *const *src_host,
                                 void *const *dst_host, const int32_t *width_host, const int32_t *kind_host,
                                 const float *scaling_raw, const float *rotation_raw, const float *noise, const void *temp,
                                 void *stream) {
    GS_REQUIRE(P > 0 && S >= 0 && new_P >= 0 && num_tensors > 0 && num_tensors <= DN_MAX_TENSORS, "sizes");
    GS_REQUIRE(src_host && dst_host && width_host && kind_host && scaling_raw && rotation_raw && temp, "null pointer");
    GS_REQUIRE(S == 0 || noise != nullptr, "noise");
    const size_t n = (size_t)DN_PLANES * P;
    const uint8_t *flags = (const uint8_t *)temp;
    const int32_t *pos = (const int32_t *)((const char *)temp + dn_align(n));
    DnTensors t;
    int widest = 0;
    for (int k = 0; k < DN_MAX_TENSORS; k++) {
        const bool v = k < num_tensors;
        t.src[k] = v ? (const float *)src_host[k] : nullptr;
        t.dst[k] = v ? (float *)dst_host[k] : nullptr;
        t.width[k] = v ? width_host[k] : 0;
        t.kind[k] = v ? kind_host[k] : 0;
        if (!v) continue;
        GS_REQUIRE(t.src[k] && t.dst[k] && t.width[k] > 0 && t.kind[k] >= DN_COPY && t.kind[k] <= DN_MOMENT, "tensor table");
        GS_REQUIRE((t.kind[k] != DN_XYZ && t.kind[k] != DN_SCALING) || t.width[k] == 3, "position / scale rows have 3 elements");
        widest = t.width[k] > widest ? t.width[k] : widest;
    }
    if (new_P == 0) return GS_OK;
    // pos is ONE exclusive scan over all five planes: plane 4 (the draw indices) starts where the output rows end
    const int split_base = new_P;
    const long long blocks = ((long long)P * widest + DN_THREADS - 1) / DN_THREADS;
    GS_REQUIRE(blocks < (1ll << 31), "too many elements");
    dim3 grid((unsigned)blocks, (unsigned)num_tensors);
    k_densify_gather<<<grid, DN_THREADS, 0, (cudaStream_t)stream>>>(P, S, t, scaling_raw, rotation_raw, noise, flags, pos,
                                                                      split_base);
    GS_LAUNCH_CHECK();
    return GS_OK;
}
