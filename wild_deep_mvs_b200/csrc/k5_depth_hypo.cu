// K5: the per-pixel part of CVP-MVSNet's calDepthHypo (models/CVP_MVSNet/models/modules.py:131-226), fp64 like the
// reference: for every pixel of the up-sampled depth map, the depth change that moves its projection into the FIRST
// source view by one pixel along the epipolar line.  The reference does this with ~150 PyTorch launches per level
// (fp64 matmuls over 3 x HW matrices, batched 2x2 inverses, det, boolean indexing); here it is one kernel -- a few
// hundred fp64 FLOPs per pixel, one 4-byte read and one 8-byte write.  The level's interval is the median of the
// valid |delta| (a sort on the device, done by the caller: cvpmvsnet.cal_depth_hypo).
#include "common.cuh"

namespace mvsb200 {

__device__ bool k5_inverse4(const double *a, double *out)
{
    double m[4][8];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            m[i][j] = a[i * 4 + j];
            m[i][4 + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < 4; c++) {
        int p = c;
        for (int r = c + 1; r < 4; r++)
            if (fabs(m[r][c]) > fabs(m[p][c])) p = r;
        if (m[p][c] == 0.0) return false;
        if (p != c)
            for (int j = 0; j < 8; j++) {
                double t = m[c][j];
                m[c][j] = m[p][j];
                m[p][j] = t;
            }
        const double piv = 1.0 / m[c][c];
        for (int j = 0; j < 8; j++) m[c][j] *= piv;
        for (int r = 0; r < 4; r++) {
            if (r == c) continue;
            const double f = m[r][c];
            for (int j = 0; j < 8; j++) m[r][j] -= f * m[c][j];
        }
    }
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) out[i * 4 + j] = m[i][4 + j];
    return true;
}

__device__ bool k5_inverse3(const double *a, double *o)
{
    const double c0 = a[4] * a[8] - a[5] * a[7], c1 = a[5] * a[6] - a[3] * a[8], c2 = a[3] * a[7] - a[4] * a[6];
    const double det = a[0] * c0 + a[1] * c1 + a[2] * c2;
    if (det == 0.0) return false;
    const double id = 1.0 / det;
    o[0] = c0 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    o[3] = c1 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    o[6] = c2 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
    return true;
}

__device__ void k5_mul3(const double *a, const double *b, double *o)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) o[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}

// s_m: [0,9) inv(K_ref)   [9,25) inv(E_ref)   [25,37) K_src * E_src[:3,:4]   [37,46) A = (K_ref R_ref) inv(K_src R_src)
__global__ void __launch_bounds__(256) k5_depth_delta_kernel(const float *__restrict__ depth, const float *__restrict__ ref_in,
                                                             const float *__restrict__ src_in, const float *__restrict__ ref_ex,
                                                             const float *__restrict__ src_ex, int H, int W,
                                                             double *__restrict__ abs_delta)
{
    __shared__ double s_m[46];
    __shared__ int s_ok;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) {
        double Kr[9], Ks[9], Er[16], Es[16], t[9], u[9];
        for (int i = 0; i < 9; i++) { Kr[i] = (double)ref_in[b * 9 + i]; Ks[i] = (double)src_in[b * 9 + i]; }
        for (int i = 0; i < 16; i++) { Er[i] = (double)ref_ex[b * 16 + i]; Es[i] = (double)src_ex[b * 16 + i]; }
        bool ok = k5_inverse3(Kr, s_m) && k5_inverse4(Er, s_m + 9);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 4; j++) s_m[25 + i * 4 + j] = Ks[i * 3] * Es[j] + Ks[i * 3 + 1] * Es[4 + j] + Ks[i * 3 + 2] * Es[8 + j];
        double Rr[9], Rs[9];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) { Rr[i * 3 + j] = Er[i * 4 + j]; Rs[i * 3 + j] = Es[i * 4 + j]; }
        k5_mul3(Ks, Rs, t);
        ok = ok && k5_inverse3(t, u);
        k5_mul3(Kr, Rr, t);
        k5_mul3(t, u, s_m + 37);
        s_ok = ok ? 1 : 0;
    }
    __syncthreads();
    const long long HW = (long long)H * W;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    double res = inf;   // +inf marks an invalid pixel (sorted to the end, not counted)
    if (s_ok) {
        const double X[3] = {(double)(pix % W), (double)(pix / W), 1.0};
        const float d1f = __ldg(depth + (long long)b * HW + pix);
        const double Dv[2] = {(double)d1f, (double)(d1f + 1.0f)};   // D2 = D1 + 1 is formed in fp32 (modules.py:163)
        double Q[2][3], z[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            double ray[3], P[4], Pw[4];
#pragma unroll
            for (int i = 0; i < 3; i++) ray[i] = s_m[i * 3] * (X[0] * Dv[k]) + s_m[i * 3 + 1] * (X[1] * Dv[k]) + s_m[i * 3 + 2] * (X[2] * Dv[k]);
#pragma unroll
            for (int i = 0; i < 4; i++) Pw[i] = s_m[9 + i * 4] * ray[0] + s_m[9 + i * 4 + 1] * ray[1] + s_m[9 + i * 4 + 2] * ray[2] + s_m[9 + i * 4 + 3];
#pragma unroll
            for (int i = 0; i < 3; i++) P[i] = s_m[25 + i * 4] * Pw[0] + s_m[25 + i * 4 + 1] * Pw[1] + s_m[25 + i * 4 + 2] * Pw[2] + s_m[25 + i * 4 + 3] * Pw[3];
            z[k] = P[2];
#pragma unroll
            for (int i = 0; i < 3; i++) Q[k][i] = P[i] / z[k];
        }
        const double dx = Q[1][0] - Q[0][0], dy = Q[1][1] - Q[0][1], dz = Q[1][2] - Q[0][2];
        const double nrm = sqrt(dx * dx + dy * dy + dz * dz);
        const double inv_n = 1.0 / fmax(nrm, 1e-8);
        const double X3[3] = {Q[0][0] + dx * inv_n, Q[0][1] + dy * inv_n, Q[0][2] + dz * inv_n};
        const double *A = s_m + 37;
        double t1[3], t2[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            t1[i] = z[0] * (A[i * 3] * Q[0][0] + A[i * 3 + 1] * Q[0][1] + A[i * 3 + 2] * Q[0][2]);
            t2[i] = A[i * 3] * X3[0] + A[i * 3 + 1] * X3[1] + A[i * 3 + 2] * X3[2];
        }
        // rows 1,2 of [X | tmp2] * (delta, .)^T = rows 1,2 of tmp1
        const double det = X[1] * t2[2] - t2[1] * X[2];
        const bool valid = nrm > 1e-8 && z[0] > 1e-8 && z[1] > 1e-8 && fabs(det) > 1e-8;
        if (valid) res = fabs((t1[1] * t2[2] - t2[1] * t1[2]) / det);
    }
    abs_delta[(long long)b * HW + pix] = res;
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_cvp_depth_delta(const float *ref_depth, const float *ref_in, const float *src_in, const float *ref_ex,
                                       const float *src_ex, int B, int H, int W, double *abs_delta, mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(ref_depth && ref_in && src_in && ref_ex && src_ex && abs_delta, "cvp_depth_delta: null pointer");
    MVSB200_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, "cvp_depth_delta: bad shape B=%d H=%d W=%d", B, H, W);
    const long long HW = (long long)H * W;
    dim3 grid((unsigned)((HW + 255) / 256), (unsigned)B);
    k5_depth_delta_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ref_depth, ref_in, src_in, ref_ex, src_ex, H, W, abs_delta);
    return check_launch("k5_depth_delta_kernel");
}
