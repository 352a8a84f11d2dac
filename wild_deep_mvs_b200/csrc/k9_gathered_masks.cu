// K9: the consumer of the all-gathered depth maps (row f3 of SURVEY.md 8): the geometric half of
// `masked_photometricloss` (models/trainer.py:240-278 of the reference) with `get_flow_from_depthmap` (:209-219),
// `flows_from_single_depthmap` (utils/utils_3D.py:190-211) and `normalize` (:243-268) folded in.
// Rank r holds the depth map it computed with view r as the reference plus, after the ONE all-gather of the path, the
// maps of all N views.  Per reference pixel and source view s:
//   X      = P_ref^-1 (x d, y d, d, 1)                              d = this rank's depth at the pixel
//   q      = P_s X,  depth_src = q.z,  flow = q.xy / max(q.z, 1e-6)
//   g      = 2 flow / (size - 1) - 1;   q.z <= 0 -> (-10, -10);   clamp to [-10, 10]                (the sampling grid)
//   inside = -1 < g.x < 1 and -1 < g.y < 1
//   wd     = grid_sample(gathered[s], g, bilinear, zeros, align_corners=False)
//   mask   = inside and |depth_src - wd| / max(wd, 1e-8) < geom_clamping
//   warped = grid_sample(imgs[s], g, ...)   (optional: what the SSIM term of the loss compares with the reference image)
// The reference runs ~30 tensor-wide ops and materialises the [b, N-1, h, w, 2] grid plus five same-sized temporaries;
// here it is one kernel, one thread per (pixel, source view), the 4x4 matrices in shared memory.
// The masks are a no-gradient quantity in the reference as well (a comparison); the differentiable image warp of the
// training loss stays with autograd -- this kernel is the inference / masking path.
#include "common.cuh"

namespace mvsb200 {

constexpr int K9_THREADS = 256;

struct K9Params {
    const float *ref_depth;     // [B,H,W]
    const float *gathered;      // [B,N,H,W]
    const float *proj;          // [B,N,4,4]
    const float *inv_ref;       // [B,4,4] = proj[:, ref]^-1
    const float *imgs;          // [B,N,C,H,W] or null
    float *grid;                // [B,N-1,H,W,2] or null
    float *depth_src;           // [B,N-1,H,W] or null
    float *warped_depth;        // [B,N-1,H,W] or null
    float *warped;              // [B,N-1,C,H,W] or null
    unsigned char *mask;        // [B,N-1,H,W]
    unsigned char *inside;      // [B,N-1,H,W] or null
    int B, N, C, H, W, ref;
    float geom_clamping;
};

// grid_sample(bilinear, zeros, align_corners=False) of one channel plane at normalised (gx, gy): ATen's order nw, ne, sw, se
__device__ __forceinline__ float k9_sample(const float *plane, int H, int W, int x0, int y0, float w00, float w01, float w10, float w11)
{
    const bool xa = x0 >= 0 && x0 < W, xb = x0 + 1 >= 0 && x0 + 1 < W, ya = y0 >= 0 && y0 < H, yb = y0 + 1 >= 0 && y0 + 1 < H;
    float r = 0.f;
    if (xa && ya) r += __ldg(plane + (long long)y0 * W + x0) * w00;
    if (xb && ya) r += __ldg(plane + (long long)y0 * W + x0 + 1) * w01;
    if (xa && yb) r += __ldg(plane + (long long)(y0 + 1) * W + x0) * w10;
    if (xb && yb) r += __ldg(plane + (long long)(y0 + 1) * W + x0 + 1) * w11;
    return r;
}

__global__ void __launch_bounds__(K9_THREADS) k9_gathered_masks_kernel(const K9Params p)
{
    __shared__ float s_inv[16];
    __shared__ float s_proj[16];
    const int b = blockIdx.z, j = blockIdx.y;                 // j: index among the sources (all views but `ref`)
    const int s = j < p.ref ? j : j + 1;
    if (threadIdx.x < 16) {
        s_inv[threadIdx.x] = p.inv_ref[b * 16 + threadIdx.x];
        s_proj[threadIdx.x] = p.proj[((long long)b * p.N + s) * 16 + threadIdx.x];
    }
    __syncthreads();
    const long long HW = (long long)p.H * p.W;
    const long long pix = (long long)blockIdx.x * K9_THREADS + threadIdx.x;
    if (pix >= HW) return;
    const int y = (int)(pix / p.W), x = (int)(pix % p.W);
    const float d = __ldg(p.ref_depth + b * HW + pix);
    // add_hom(add_hom(grid) * depth) @ inv_proj^T, then @ proj^T: two 4x4 products in the reference's order
    const float h0 = (float)x * d, h1 = (float)y * d, h2 = d;
    float X[4], q[3];
#pragma unroll
    for (int r = 0; r < 4; r++) X[r] = s_inv[r * 4] * h0 + s_inv[r * 4 + 1] * h1 + s_inv[r * 4 + 2] * h2 + s_inv[r * 4 + 3];
#pragma unroll
    for (int r = 0; r < 3; r++) q[r] = s_proj[r * 4] * X[0] + s_proj[r * 4 + 1] * X[1] + s_proj[r * 4 + 2] * X[2] + s_proj[r * 4 + 3] * X[3];
    const float zc = fmaxf(q[2], 1e-6f);
    float gx = 2.f * (q[0] / zc) / (float)(p.W - 1) - 1.f;
    float gy = 2.f * (q[1] / zc) / (float)(p.H - 1) - 1.f;
    if (q[2] <= 0.f) gx = gy = -10.f;
    // torch.clamp propagates NaN (fminf / fmaxf would not): a NaN coordinate stays NaN, every comparison below is then
    // false (mask 0) and the sample reads nothing
    const bool bad = (gx != gx) || (gy != gy);
    gx = (gx != gx) ? gx : fminf(fmaxf(gx, -10.f), 10.f);
    gy = (gy != gy) ? gy : fminf(fmaxf(gy, -10.f), 10.f);
    const bool inside = gx < 1.f && gy < 1.f && gx > -1.f && gy > -1.f;

    // align_corners=False: ix = ((g + 1) * size - 1) / 2
    const float ix = ((gx + 1.f) * (float)p.W - 1.f) / 2.f, iy = ((gy + 1.f) * (float)p.H - 1.f) / 2.f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = bad ? -2 : (int)fx, y0 = bad ? -2 : (int)fy;
    const float w00 = (fx + 1.f - ix) * (fy + 1.f - iy), w01 = (ix - fx) * (fy + 1.f - iy);
    const float w10 = (fx + 1.f - ix) * (iy - fy), w11 = (ix - fx) * (iy - fy);

    const long long o = ((long long)b * (p.N - 1) + j) * HW + pix;
    const float wd = k9_sample(p.gathered + ((long long)b * p.N + s) * HW, p.H, p.W, x0, y0, w00, w01, w10, w11);
    const float diff = fabsf(q[2] - wd) / fmaxf(wd, 1e-8f);
    p.mask[o] = (inside && diff < p.geom_clamping) ? 1 : 0;
    if (p.inside) p.inside[o] = inside ? 1 : 0;
    if (p.grid) { p.grid[o * 2] = gx; p.grid[o * 2 + 1] = gy; }
    if (p.depth_src) p.depth_src[o] = q[2];
    if (p.warped_depth) p.warped_depth[o] = wd;
    if (p.warped) {
        for (int c = 0; c < p.C; c++)
            p.warped[(((long long)b * (p.N - 1) + j) * p.C + c) * HW + pix] =
                k9_sample(p.imgs + (((long long)b * p.N + s) * p.C + c) * HW, p.H, p.W, x0, y0, w00, w01, w10, w11);
    }
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_gathered_masks(int B, int N, int C, int H, int W, int ref, const float *ref_depth, const float *gathered,
                                      const float *proj, const float *inv_ref, const float *imgs, float geom_clamping,
                                      unsigned char *mask, unsigned char *inside, float *grid, float *depth_src, float *warped_depth,
                                      float *warped, mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(ref_depth && gathered && proj && inv_ref && mask, "gathered_masks: null pointer");
    MVSB200_REQUIRE(B > 0 && B <= 65535 && N >= 2 && N <= 65535 && H > 1 && W > 1, "gathered_masks: bad shape B=%d N=%d H=%d W=%d", B, N, H, W);
    MVSB200_REQUIRE(ref >= 0 && ref < N, "gathered_masks: reference view %d not in [0,%d)", ref, N);
    MVSB200_REQUIRE(!warped || (imgs && C > 0), "gathered_masks: warped images need imgs and C > 0");
    K9Params p;
    p.ref_depth = ref_depth; p.gathered = gathered; p.proj = proj; p.inv_ref = inv_ref; p.imgs = imgs;
    p.grid = grid; p.depth_src = depth_src; p.warped_depth = warped_depth; p.warped = warped; p.mask = mask; p.inside = inside;
    p.B = B; p.N = N; p.C = C; p.H = H; p.W = W; p.ref = ref; p.geom_clamping = geom_clamping;
    const long long HW = (long long)H * W;
    dim3 grid_dim((unsigned)((HW + K9_THREADS - 1) / K9_THREADS), (unsigned)(N - 1), (unsigned)B);
    k9_gathered_masks_kernel<<<grid_dim, K9_THREADS, 0, (cudaStream_t)stream>>>(p);
    return check_launch("k9_gathered_masks_kernel");
}
