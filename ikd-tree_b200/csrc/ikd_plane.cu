// Plane fit on the k nearest neighbours, run on the device right behind the kNN kernel.
//
// The known caller of Nearest_Search (FAST-LIO2 laserMapping.cpp, h_share_model -> esti_plane; external to the
// reference tree, so this is an extension of the path, not a restatement of reference code) takes the 5 neighbours
// of every scan point, solves the 5x3 least-squares problem A n = -1 (rows of A = neighbour coordinates) with a
// column-pivoted Householder QR in fp32, normalises n to (a, b, c, d = 1/|n|), rejects the plane when any neighbour
// lies further than `threshold` from it, and evaluates the query's signed distance a*x + b*y + c*z + d. Doing this
// here keeps the neighbours in HBM: 21 bytes per query return to the host instead of 8k + 4.
//
// Arithmetic is fp32, fixed operation order, no FMA (-fmad=false; IEEE sqrt/div), so the CPU checker used by the
// tests reproduces it bit for bit. One thread per query; the 5x3 system lives in
// registers (all loops are unrolled over the template K).
#include "ikd_host.h"

namespace ikd {

template <int K>
__device__ __forceinline__ bool fit_plane(float (&A)[K][3], float thr, float (&pl)[4]) {
    float P[K][3];
    float b[K];
#pragma unroll
    for (int j = 0; j < K; j++) {
        b[j] = -1.f;
#pragma unroll
        for (int c = 0; c < 3; c++) P[j][c] = A[j][c];
    }
    int perm[3] = {0, 1, 2};
    bool ok = true;
#pragma unroll
    for (int s = 0; s < 3; s++) {
        // pivot: remaining column with the largest squared norm over rows s..K-1 (first one on ties)
        int best = s;
        float bestn = -1.f;
#pragma unroll
        for (int c = s; c < 3; c++) {
            float n = 0.f;
#pragma unroll
            for (int j = s; j < K; j++) n = n + A[j][c] * A[j][c];
            if (n > bestn) { bestn = n; best = c; }
        }
#pragma unroll
        for (int c = s + 1; c < 3; c++) {
            if (c == best) {
#pragma unroll
                for (int j = 0; j < K; j++) { float tmp = A[j][s]; A[j][s] = A[j][c]; A[j][c] = tmp; }
                int tp = perm[s]; perm[s] = perm[c]; perm[c] = tp;
            }
        }
        // Householder reflector for rows s..K-1 of column s: v = (1, ess), H = I - tau v v^T, H x = (beta, 0..)
        float c0 = A[s][s];
        float tail = 0.f;
#pragma unroll
        for (int j = s + 1; j < K; j++) tail = tail + A[j][s] * A[j][s];
        float beta = c0, tau = 0.f;
        float ess[K];
#pragma unroll
        for (int j = 0; j < K; j++) ess[j] = 0.f;
        if (tail != 0.f) {
            beta = sqrtf(c0 * c0 + tail);
            if (c0 >= 0.f) beta = -beta;
            float den = c0 - beta;
#pragma unroll
            for (int j = s + 1; j < K; j++) ess[j] = A[j][s] / den;
            tau = (beta - c0) / beta;
        }
        A[s][s] = beta;
        if (!(beta != 0.f)) ok = false;  // zero or NaN pivot: rank deficient
#pragma unroll
        for (int c = s + 1; c < 3; c++) {
            float w = A[s][c];
#pragma unroll
            for (int j = s + 1; j < K; j++) w = w + ess[j] * A[j][c];
            w = w * tau;
            A[s][c] = A[s][c] - w;
#pragma unroll
            for (int j = s + 1; j < K; j++) A[j][c] = A[j][c] - ess[j] * w;
        }
        {
            float w = b[s];
#pragma unroll
            for (int j = s + 1; j < K; j++) w = w + ess[j] * b[j];
            w = w * tau;
            b[s] = b[s] - w;
#pragma unroll
            for (int j = s + 1; j < K; j++) b[j] = b[j] - ess[j] * w;
        }
    }
    // back substitution on the 3x3 upper triangle
    float x2 = b[2] / A[2][2];
    float x1 = (b[1] - A[1][2] * x2) / A[1][1];
    float x0 = ((b[0] - A[0][1] * x1) - A[0][2] * x2) / A[0][0];
    float nv[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 3; a++) {  // undo the column permutation (static indexing keeps nv in registers)
        if (perm[0] == a) nv[a] = x0;
        if (perm[1] == a) nv[a] = x1;
        if (perm[2] == a) nv[a] = x2;
    }
    float nn = sqrtf((nv[0] * nv[0] + nv[1] * nv[1]) + nv[2] * nv[2]);
    pl[0] = nv[0] / nn;
    pl[1] = nv[1] / nn;
    pl[2] = nv[2] / nn;
    pl[3] = 1.f / nn;
#pragma unroll
    for (int c = 0; c < 4; c++)
        if (!(fabsf(pl[c]) <= 3.0e38f)) ok = false;  // inf / NaN
    if (ok) {
#pragma unroll
        for (int j = 0; j < K; j++) {
            float r = ((pl[0] * P[j][0] + pl[1] * P[j][1]) + pl[2] * P[j][2]) + pl[3];
            if (fabsf(r) > thr) ok = false;
        }
    }
    return ok;
}

template <int K>
__global__ void __launch_bounds__(128)
plane_fit_kernel(const float4* __restrict__ q, int nq, const int32_t* __restrict__ idx, const float* __restrict__ sqd,
                 const int32_t* __restrict__ cnt, const float4* __restrict__ pid_xyz, float max_kth_sqdist, float thr,
                 float4* __restrict__ out_plane, float* __restrict__ out_resid, uint8_t* __restrict__ out_valid) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    float4 plane = make_float4(0.f, 0.f, 0.f, 0.f);
    float resid = 0.f;
    uint8_t valid = 0;
    // same gate as the caller: k neighbours found and the k-th one close enough
    if (cnt[i] == K && sqd[(size_t)i * K + (K - 1)] <= max_kth_sqdist) {
        float A[K][3];
#pragma unroll
        for (int j = 0; j < K; j++) {
            float4 p = __ldg(&pid_xyz[idx[(size_t)i * K + j]]);
            A[j][0] = p.x; A[j][1] = p.y; A[j][2] = p.z;
        }
        float pl[4];
        bool ok = fit_plane<K>(A, thr, pl);
        float4 v = q[i];
        plane = make_float4(pl[0], pl[1], pl[2], pl[3]);
        resid = ((pl[0] * v.x + pl[1] * v.y) + pl[2] * v.z) + pl[3];
        if (!(fabsf(resid) <= 3.0e38f)) { ok = false; resid = 0.f; plane = make_float4(0.f, 0.f, 0.f, 0.f); }
        valid = ok ? 1 : 0;
    }
    out_plane[i] = plane;
    out_resid[i] = resid;
    out_valid[i] = valid;
}

int plane_fit_launch(ikd_tree* t, const float4* q_dev, int64_t nq, int k, const int32_t* idx, const float* sqd,
                     const int32_t* cnt, float max_kth_sqdist, float thr, float* out_plane, float* out_resid,
                     uint8_t* out_valid, cudaStream_t s) {
    if (nq <= 0) return IKD_OK;
    if (k < IKD_PLANE_MIN_K || k > IKD_PLANE_MAX_K) {
        set_error("plane fit needs %d <= k <= %d (got %d)", IKD_PLANE_MIN_K, IKD_PLANE_MAX_K, k);
        return IKD_ERR_ARG;
    }
    const int n = (int)nq, TPB = 128, grid = (n + TPB - 1) / TPB;
    const float4* px = t->pid_xyz.as<float4>();
    float4* op = (float4*)out_plane;
#define IKD_PLANE_CASE(KK)                                                                                             \
    case KK:                                                                                                           \
        IKD_LAUNCH plane_fit_kernel<KK><<<grid, TPB, 0, s>>>(q_dev, n, idx, sqd, cnt, px, max_kth_sqdist, thr, op, out_resid, \
                                                             out_valid);                                               \
        break;
    switch (k) {
        IKD_PLANE_CASE(3) IKD_PLANE_CASE(4) IKD_PLANE_CASE(5) IKD_PLANE_CASE(6) IKD_PLANE_CASE(7) IKD_PLANE_CASE(8)
    }
#undef IKD_PLANE_CASE
    IKD_CUDA(cudaGetLastError());
    return IKD_OK;
}

}  // namespace ikd
