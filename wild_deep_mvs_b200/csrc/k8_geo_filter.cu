// K8: geometric-consistency filter of a depth map against its source views (evaluation/filtering.py:59-84 of the
// reference; row f3 of SURVEY.md 8 -- the consumer of the gathered depth maps).  Per reference pixel and source view:
// unproject with the reference depth, project into the source, sample the source's depth map (bilinear, zero padding,
// the reference's (size-1)-normalised grid read with align_corners=False), unproject with the sampled depth, project
// back; then the three tests of the reference (re-projection error, relative depth difference, triangulation angle)
// and the per-pixel vote over the sources.  The reference runs this on the CPU with ~25 tensor-wide PyTorch ops per
// scene view; here it is one kernel, one thread per pixel, the cameras in shared memory, no intermediate tensors.
#include "common.cuh"

namespace mvsb200 {

constexpr int K8_THREADS = 256;

struct K8Params {
    const float *depth;
    const float *src_depth[MVSB200_MAX_SRC];
    int src_h[MVSB200_MAX_SRC], src_w[MVSB200_MAX_SRC];
    const float *K, *R, *t;      // [1+N,3,3], [1+N,3,3], [1+N,3]
    unsigned char *mask_depth, *mask_disp, *geo_mask, *votes;
    int N, H, W, need;
    float depth_threshold, max_reproj_error, min_tri_angle;
};

// per view in shared memory: K[9] R[9] t[3] Kinv[9] c[3] (c = R^T t)  -> 33 floats
constexpr int K8_CAM = 33;

__device__ void k8_inverse3(const float *a, float *o)
{
    const double A[9] = {a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]};
    const double c0 = A[4] * A[8] - A[5] * A[7], c1 = A[5] * A[6] - A[3] * A[8], c2 = A[3] * A[7] - A[4] * A[6];
    const double id = 1.0 / (A[0] * c0 + A[1] * c1 + A[2] * c2);
    o[0] = (float)(c0 * id); o[1] = (float)((A[2] * A[7] - A[1] * A[8]) * id); o[2] = (float)((A[1] * A[5] - A[2] * A[4]) * id);
    o[3] = (float)(c1 * id); o[4] = (float)((A[0] * A[8] - A[2] * A[6]) * id); o[5] = (float)((A[2] * A[3] - A[0] * A[5]) * id);
    o[6] = (float)(c2 * id); o[7] = (float)((A[1] * A[6] - A[0] * A[7]) * id); o[8] = (float)((A[0] * A[4] - A[1] * A[3]) * id);
}

__device__ __forceinline__ void k8_mul(const float *m, float x, float y, float z, float &ox, float &oy, float &oz)
{
    ox = m[0] * x + m[1] * y + m[2] * z;
    oy = m[3] * x + m[4] * y + m[5] * z;
    oz = m[6] * x + m[7] * y + m[8] * z;
}
__device__ __forceinline__ void k8_mul_t(const float *m, float x, float y, float z, float &ox, float &oy, float &oz)
{   // m^T v
    ox = m[0] * x + m[3] * y + m[6] * z;
    oy = m[1] * x + m[4] * y + m[7] * z;
    oz = m[2] * x + m[5] * y + m[8] * z;
}

__global__ void __launch_bounds__(K8_THREADS) k8_geo_filter_kernel(const K8Params p)
{
    __shared__ float s_cam[(MVSB200_MAX_SRC + 1) * K8_CAM];
    for (int v = threadIdx.x; v <= p.N; v += K8_THREADS) {
        float *c = s_cam + v * K8_CAM;
        for (int i = 0; i < 9; i++) { c[i] = p.K[v * 9 + i]; c[9 + i] = p.R[v * 9 + i]; }
        for (int i = 0; i < 3; i++) c[18 + i] = p.t[v * 3 + i];
        k8_inverse3(c, c + 21);
        k8_mul_t(c + 9, c[18], c[19], c[20], c[30], c[31], c[32]);
    }
    __syncthreads();
    const long long pix = (long long)blockIdx.x * K8_THREADS + threadIdx.x;
    if (pix >= (long long)p.H * p.W) return;
    const float fx = (float)(pix % p.W), fy = (float)(pix / p.W);
    const float d0 = __ldg(p.depth + pix);
    const float *c0 = s_cam;
    // unproject (utils_3D.py:131): ((x, y, 1) d) inv(K0)^T - t0^T) R0   ==   R0^T (inv(K0) (x d, y d, d) - t0)
    float ax, ay, az, X, Y, Z;
    k8_mul(c0 + 21, fx * d0, fy * d0, d0, ax, ay, az);
    k8_mul_t(c0 + 9, ax - c0[18], ay - c0[19], az - c0[20], X, Y, Z);
    const float r1x = X + c0[30], r1y = Y + c0[31], r1z = Z + c0[32];               // ray from the reference centre (:304)
    const float n1 = fmaxf(sqrtf(r1x * r1x + r1y * r1y + r1z * r1z), 1e-12f);
    int n_depth = 0, n_disp = 0, n_geo = 0;
    for (int i = 0; i < p.N; i++) {
        const float *c = s_cam + (1 + i) * K8_CAM;
        // project_all (:70): K_i (R_i X + t_i)
        float qx, qy, qz, ux, uy, dsrc;
        k8_mul(c + 9, X, Y, Z, qx, qy, qz);
        k8_mul(c, qx + c[18], qy + c[19], qz + c[20], ux, uy, dsrc);
        const float zc = fmaxf(dsrc, 1e-6f);
        const float px = ux / zc, py = uy / zc;
        // normalize (:267-268) with (size - 1), grid_sample with align_corners = False, zero padding (filtering.py:66-69)
        const int Ws = p.src_w[i], Hs = p.src_h[i];
        const float gx = 2.f * px / (float)(Ws - 1) - 1.f, gy = 2.f * py / (float)(Hs - 1) - 1.f;
        const float ix = ((gx + 1.f) * (float)Ws - 1.f) / 2.f, iy = ((gy + 1.f) * (float)Hs - 1.f) / 2.f;
        float wd = 0.f;
        if (ix == ix && iy == iy && fabsf(ix) < 1e9f && fabsf(iy) < 1e9f) {
            const float x0 = floorf(ix), y0 = floorf(iy);
            const int xi = (int)x0, yi = (int)y0;
            const float wx1 = ix - x0, wy1 = iy - y0;
            const float *sd = p.src_depth[i];
#pragma unroll
            for (int dy = 0; dy < 2; dy++)
#pragma unroll
                for (int dx = 0; dx < 2; dx++) {
                    const int xx = xi + dx, yy = yi + dy;
                    if ((unsigned)xx < (unsigned)Ws && (unsigned)yy < (unsigned)Hs)
                        wd += __ldg(sd + (long long)yy * Ws + xx) * ((dx ? wx1 : 1.f - wx1) * (dy ? wy1 : 1.f - wy1));
                }
        }
        // unproj_all (:178) with the sampled depth, then project into the reference (:104-107)
        float bx, by, bz, sx, sy, sz, vx, vy, vz, wx, wy, wz;
        k8_mul(c + 21, px * wd, py * wd, wd, bx, by, bz);
        k8_mul_t(c + 9, bx - c[18], by - c[19], bz - c[20], sx, sy, sz);
        k8_mul(c0 + 9, sx, sy, sz, vx, vy, vz);
        k8_mul(c0, vx + c0[18], vy + c0[19], vz + c0[20], wx, wy, wz);
        const float drep = wz + 1e-6f;
        const float ex = wx / drep - fx, ey = wy / drep - fy;
        const bool valid_disp = sqrtf(ex * ex + ey * ey) < p.max_reproj_error;                                  // :73
        const bool mask_depth = (fabsf(drep - d0) < fmaxf(drep, d0) * p.depth_threshold) && drep > 0.f && dsrc > 0.f;   // :75-76
        const float r2x = X + c[30], r2y = Y + c[31], r2z = Z + c[32];                                          // :305
        const float n2 = fmaxf(sqrtf(r2x * r2x + r2y * r2y + r2z * r2z), 1e-12f);
        const float cs = fminf(fmaxf((r1x * r2x + r1y * r2y + r1z * r2z) / n1 / n2, -1.f), 1.f);
        const bool tri = acosf(cs) / 3.14159265358979323846f * 180.f > p.min_tri_angle;                          // :78
        n_depth += mask_depth;
        n_disp += valid_disp;
        n_geo += (mask_depth && valid_disp && tri);
    }
    p.mask_depth[pix] = n_depth >= p.need;
    p.mask_disp[pix] = n_disp >= p.need;
    p.geo_mask[pix] = n_geo >= p.need;
    if (p.votes) {
        const long long hw = (long long)p.H * p.W;
        p.votes[pix] = (unsigned char)n_depth; p.votes[hw + pix] = (unsigned char)n_disp; p.votes[2 * hw + pix] = (unsigned char)n_geo;
    }
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_geometric_filter(const float *depth, int H, int W, const float *const *src_depth, const int *src_h,
                                        const int *src_w, int N, const float *K, const float *R, const float *t,
                                        float depth_threshold, float max_reproj_error, float min_tri_angle, int num_consistent,
                                        unsigned char *mask_depth, unsigned char *mask_disp, unsigned char *geo_mask,
                                        unsigned char *votes, mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(depth && src_depth && src_h && src_w && K && R && t && mask_depth && mask_disp && geo_mask, "geometric_filter: null pointer");
    MVSB200_REQUIRE(H > 0 && W > 0 && N >= 1 && N <= MVSB200_MAX_SRC, "geometric_filter: bad shape H=%d W=%d N=%d", H, W, N);
    K8Params p;
    p.depth = depth;
    for (int i = 0; i < N; i++) {
        MVSB200_REQUIRE(src_depth[i] && src_h[i] > 1 && src_w[i] > 1, "geometric_filter: source %d invalid", i);
        p.src_depth[i] = src_depth[i]; p.src_h[i] = src_h[i]; p.src_w[i] = src_w[i];
    }
    p.K = K; p.R = R; p.t = t;
    p.mask_depth = mask_depth; p.mask_disp = mask_disp; p.geo_mask = geo_mask; p.votes = votes;
    p.N = N; p.H = H; p.W = W; p.need = num_consistent - 1;
    p.depth_threshold = depth_threshold; p.max_reproj_error = max_reproj_error; p.min_tri_angle = min_tri_angle;
    const long long blocks = ((long long)H * W + K8_THREADS - 1) / K8_THREADS;
    MVSB200_REQUIRE(blocks < (1ll << 31), "geometric_filter: image too large");
    k8_geo_filter_kernel<<<(unsigned)blocks, K8_THREADS, 0, (cudaStream_t)stream>>>(p);
    return check_launch("k8_geo_filter_kernel");
}
