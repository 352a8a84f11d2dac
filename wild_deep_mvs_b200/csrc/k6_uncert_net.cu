// K6: Vis-MVSNet's UncertNet (models/VisMVSNet/model_cas.py:77-98) as ONE kernel: the per-pair entropy map goes through
// conv 1->8 + BN + ReLU, conv 8->8 + BN + ReLU (+ the input broadcast over the channels, :95) and the 8->1 head.  Three
// tiny 2-D convolutions on [S*B, H, W] maps were four launches (one of them a shape-agnostic fallback kernel) and
// 11 % of the cascade; fused, the two 8-channel intermediates of a 16x16 tile live in shared memory and the 752
// weights / folded BN constants are launch parameters read from the constant bank by fully unrolled FMAs.
#include "common.cuh"

namespace mvsb200 {

constexpr int U_T = 16;                       // output tile edge
constexpr int U_E0 = U_T + 6, U_E1 = U_T + 4, U_E2 = U_T + 2;

struct UncertParams {
    const float *ent;
    float *out;
    int H, W, tiles_x, tiles_y;
    float w1[9 * 8], s1[8], b1[8];            // [tap][co]
    float w2[9 * 8 * 8], s2[8], b2[8];        // [tap][ci][co]
    float wh[9 * 8];                          // [tap][ci]
};

__global__ void __launch_bounds__(256) k6_uncert_net_kernel(const __grid_constant__ UncertParams p)
{
    __shared__ float s_x[U_E0 * U_E0];
    __shared__ __align__(16) float s_y1[U_E1 * U_E1 * 8];
    __shared__ __align__(16) float s_y2[U_E2 * U_E2 * 8];
    int t = blockIdx.x;
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    const int n = t / p.tiles_y;
    const int x0 = tx * U_T, y0 = ty * U_T;
    const float *img = p.ent + (long long)n * p.H * p.W;

    for (int i = threadIdx.x; i < U_E0 * U_E0; i += 256) {
        const int gy = y0 - 3 + i / U_E0, gx = x0 - 3 + i % U_E0;
        s_x[i] = ((unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W) ? __ldg(img + (long long)gy * p.W + gx) : 0.f;
    }
    __syncthreads();
    // layer 1 on the (T+4)^2 halo region; positions outside the image are the zero padding layer 2 sees
    for (int i = threadIdx.x; i < U_E1 * U_E1; i += 256) {
        const int ly = i / U_E1, lx = i % U_E1;
        const int gy = y0 - 2 + ly, gx = x0 - 2 + lx;
        float a[8];
#pragma unroll
        for (int c = 0; c < 8; c++) a[c] = 0.f;
#pragma unroll
        for (int dy = 0; dy < 3; dy++)
#pragma unroll
            for (int dx = 0; dx < 3; dx++) {
                const float v = s_x[(ly + dy) * U_E0 + lx + dx];
#pragma unroll
                for (int c = 0; c < 8; c++) a[c] = fmaf(v, p.w1[(dy * 3 + dx) * 8 + c], a[c]);
            }
        const bool in = (unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W;
#pragma unroll
        for (int c = 0; c < 8; c++) s_y1[i * 8 + c] = in ? fmaxf(fmaf(a[c], p.s1[c], p.b1[c]), 0.f) : 0.f;
    }
    __syncthreads();
    // layer 2 (+ the input, broadcast over the channels) on the (T+2)^2 halo region
    for (int i = threadIdx.x; i < U_E2 * U_E2; i += 256) {
        const int ly = i / U_E2, lx = i % U_E2;
        const int gy = y0 - 1 + ly, gx = x0 - 1 + lx;
        float a[8];
#pragma unroll
        for (int c = 0; c < 8; c++) a[c] = 0.f;
#pragma unroll
        for (int dy = 0; dy < 3; dy++)
#pragma unroll
            for (int dx = 0; dx < 3; dx++) {
                const float4 u0 = *reinterpret_cast<const float4 *>(&s_y1[((ly + dy) * U_E1 + lx + dx) * 8]);
                const float4 u1 = *reinterpret_cast<const float4 *>(&s_y1[((ly + dy) * U_E1 + lx + dx) * 8 + 4]);
                const float u[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
                for (int ci = 0; ci < 8; ci++)
#pragma unroll
                    for (int c = 0; c < 8; c++) a[c] = fmaf(u[ci], p.w2[((dy * 3 + dx) * 8 + ci) * 8 + c], a[c]);
            }
        const bool in = (unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W;
        const float xin = s_x[(ly + 2) * U_E0 + lx + 2];
#pragma unroll
        for (int c = 0; c < 8; c++) s_y2[i * 8 + c] = in ? fmaxf(fmaf(a[c], p.s2[c], p.b2[c]), 0.f) + xin : 0.f;
    }
    __syncthreads();
    // head
    {
        const int ly = threadIdx.x / U_T, lx = threadIdx.x % U_T;
        const int gy = y0 + ly, gx = x0 + lx;
        float acc = 0.f;
#pragma unroll
        for (int dy = 0; dy < 3; dy++)
#pragma unroll
            for (int dx = 0; dx < 3; dx++) {
                const float4 u0 = *reinterpret_cast<const float4 *>(&s_y2[((ly + dy) * U_E2 + lx + dx) * 8]);
                const float4 u1 = *reinterpret_cast<const float4 *>(&s_y2[((ly + dy) * U_E2 + lx + dx) * 8 + 4]);
                const float *w = &p.wh[(dy * 3 + dx) * 8];
                acc = fmaf(u0.x, w[0], acc); acc = fmaf(u0.y, w[1], acc); acc = fmaf(u0.z, w[2], acc); acc = fmaf(u0.w, w[3], acc);
                acc = fmaf(u1.x, w[4], acc); acc = fmaf(u1.y, w[5], acc); acc = fmaf(u1.z, w[6], acc); acc = fmaf(u1.w, w[7], acc);
            }
        if (gy < p.H && gx < p.W) p.out[((long long)n * p.H + gy) * p.W + gx] = acc;
    }
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_vis_uncert_net(const float *entropy, int N, int H, int W, const float *params_host, float *out,
                                      mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(entropy && params_host && out, "vis_uncert_net: null pointer");
    MVSB200_REQUIRE(N > 0 && H > 0 && W > 0, "vis_uncert_net: bad shape N=%d H=%d W=%d", N, H, W);
    UncertParams p;
    p.ent = entropy; p.out = out; p.H = H; p.W = W;
    p.tiles_x = (W + U_T - 1) / U_T;
    p.tiles_y = (H + U_T - 1) / U_T;
    const float *q = params_host;
    for (int i = 0; i < 72; i++) p.w1[i] = *q++;
    for (int i = 0; i < 8; i++) p.s1[i] = *q++;
    for (int i = 0; i < 8; i++) p.b1[i] = *q++;
    for (int i = 0; i < 576; i++) p.w2[i] = *q++;
    for (int i = 0; i < 8; i++) p.s2[i] = *q++;
    for (int i = 0; i < 8; i++) p.b2[i] = *q++;
    for (int i = 0; i < 72; i++) p.wh[i] = *q++;
    const long long blocks = (long long)p.tiles_x * p.tiles_y * N;
    MVSB200_REQUIRE(blocks < (1ll << 31), "vis_uncert_net: too many tiles");
    k6_uncert_net_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("k6_uncert_net_kernel");
}
