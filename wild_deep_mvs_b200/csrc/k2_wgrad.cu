// K2 weight gradient (row f2 of SURVEY.md 8): the correlation every 3x3x3 layer of the regularisers needs in training,
//
//     R[ca][cb][t] = sum over voxels v of  a[v, ca] * b[s*v + t - 1, cb]          t = (dz, dy, dx) in {0,1,2}^3
//
// with `a` the low-resolution side and `b` the high-resolution side of the layer (zero outside the volume):
//   Conv3d(k=3, p=1, stride s), weight [Cout,Cin,3,3,3]           a = grad_out, b = input     -> R = dW
//   ConvTranspose3d(k=3, p=1, stride s, output_padding s-1), weight [Cin,Cout,3,3,3]
//                                                                  a = input,    b = grad_out -> R = dW
// (autograd of models/MVSNet/module.py:41-48, MVSNet/model.py:59-72, VisMVSNet/nn_utils.py:123-278, CVP net.py:50-74).
//
// First version, on the CUDA cores in fp32: a block owns an 8 x 32 tile of (ca, cb) pairs -- one pair per thread, 27
// tap accumulators in registers -- and walks rows (b, z, y) of `a`; the row of `a` and the nine rows of `b` it touches
// are staged in shared memory in x segments, so a warp reads 32 consecutive cb of one staged position (conflict free)
// and one broadcast value of `a` per FMA group.  Blocks add their partial sums into R with atomics (R zeroed by the
// caller).  The dgrad of the same layers needs no kernel of its own: it is a forward call of the K2 engines with the
// weights re-packed (ops.conv3d_input_grad).
#include "common.cuh"

namespace mvsb200 {

constexpr int WG_CA = 8, WG_CB = 32, WG_THREADS = WG_CA * WG_CB;

struct WgradParams {
    const float *a, *b;
    float *r;
    int B, Da, Ha, Wa, Db, Hb, Wb, Ca, Cb, stride;
    int tiles_a, tiles_b;
    long long rows;
};

template <int S>
__global__ void __launch_bounds__(WG_THREADS) k2_wgrad_kernel(const WgradParams p)
{
    constexpr int WG_XS = S == 1 ? 32 : 16;              // x positions of `a` per staged segment (38 KB of b either way)
    constexpr int XB = S * (WG_XS - 1) + 3;              // staged x extent of `b` for one segment
    __shared__ float s_a[WG_XS][WG_CA];
    __shared__ float s_b[9][XB][WG_CB];
    const int ta = blockIdx.y % p.tiles_a, tb = blockIdx.y / p.tiles_a;
    const int ca0 = ta * WG_CA, cb0 = tb * WG_CB;
    const int la = threadIdx.x / WG_CB, lb = threadIdx.x % WG_CB;
    float acc[27];
#pragma unroll
    for (int t = 0; t < 27; t++) acc[t] = 0.f;

    for (long long row = blockIdx.x; row < p.rows; row += gridDim.x) {
        const int ya = (int)(row % p.Ha);
        const int za = (int)((row / p.Ha) % p.Da);
        const int bb = (int)(row / ((long long)p.Ha * p.Da));
        for (int x0 = 0; x0 < p.Wa; x0 += WG_XS) {
            __syncthreads();
            // stage a[bb, za, ya, x0 .. x0+XS) for the tile's channels (zero beyond the row / the channel count)
            for (int i = threadIdx.x; i < WG_XS * WG_CA; i += WG_THREADS) {
                const int c = i % WG_CA, x = i / WG_CA;
                float v = 0.f;
                if (x0 + x < p.Wa && ca0 + c < p.Ca)
                    v = __ldg(p.a + ((((long long)bb * p.Da + za) * p.Ha + ya) * p.Wa + x0 + x) * p.Ca + ca0 + c);
                s_a[x][c] = v;
            }
            // stage the nine (dz, dy) rows of b, x from S*x0 - 1
            for (int i = threadIdx.x; i < 9 * XB * WG_CB; i += WG_THREADS) {
                const int c = i % WG_CB, x = (i / WG_CB) % XB, rr = i / (WG_CB * XB);
                const int zb = S * za + rr / 3 - 1, yb = S * ya + rr % 3 - 1, xb = S * x0 - 1 + x;
                float v = 0.f;
                if ((unsigned)zb < (unsigned)p.Db && (unsigned)yb < (unsigned)p.Hb && (unsigned)xb < (unsigned)p.Wb && cb0 + c < p.Cb)
                    v = __ldg(p.b + ((((long long)bb * p.Db + zb) * p.Hb + yb) * p.Wb + xb) * p.Cb + cb0 + c);
                s_b[rr][x][c] = v;
            }
            __syncthreads();
            const int nx = min(WG_XS, p.Wa - x0);
            for (int x = 0; x < nx; x++) {
                const float av = s_a[x][la];
#pragma unroll
                for (int rr = 0; rr < 9; rr++)
#pragma unroll
                    for (int dx = 0; dx < 3; dx++) acc[rr * 3 + dx] = fmaf(av, s_b[rr][S * x + dx][lb], acc[rr * 3 + dx]);
            }
        }
    }
    if (ca0 + la < p.Ca && cb0 + lb < p.Cb) {
        float *dst = p.r + ((long long)(ca0 + la) * p.Cb + cb0 + lb) * 27;
#pragma unroll
        for (int t = 0; t < 27; t++) atomicAdd(dst + t, acc[t]);
    }
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_conv3d_wgrad(const float *a, const float *b, int B, int Da, int Ha, int Wa, int Ca, int Db, int Hb, int Wb,
                                    int Cb, int stride, float *r, mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(a && b && r, "conv3d_wgrad: null pointer");
    MVSB200_REQUIRE(B > 0 && Da > 0 && Ha > 0 && Wa > 0 && Ca > 0 && Db > 0 && Hb > 0 && Wb > 0 && Cb > 0,
                    "conv3d_wgrad: bad shape");
    MVSB200_REQUIRE(stride == 1 || stride == 2, "conv3d_wgrad: stride=%d", stride);
    MVSB200_REQUIRE(Da == (Db + stride - 1) / stride && Ha == (Hb + stride - 1) / stride && Wa == (Wb + stride - 1) / stride,
                    "conv3d_wgrad: a is %dx%dx%d, b is %dx%dx%d, stride %d", Da, Ha, Wa, Db, Hb, Wb, stride);
    WgradParams p;
    p.a = a; p.b = b; p.r = r;
    p.B = B; p.Da = Da; p.Ha = Ha; p.Wa = Wa; p.Db = Db; p.Hb = Hb; p.Wb = Wb; p.Ca = Ca; p.Cb = Cb; p.stride = stride;
    p.tiles_a = (Ca + WG_CA - 1) / WG_CA;
    p.tiles_b = (Cb + WG_CB - 1) / WG_CB;
    p.rows = (long long)B * Da * Ha;
    const int ctiles = p.tiles_a * p.tiles_b;
    long long gx = (148ll * 4 + ctiles - 1) / ctiles;      // about four blocks per SM over all channel tiles
    if (gx > p.rows) gx = p.rows;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)ctiles);
    cudaStream_t st = (cudaStream_t)stream;
    if (stride == 1) k2_wgrad_kernel<1><<<grid, WG_THREADS, 0, st>>>(p);
    else k2_wgrad_kernel<2><<<grid, WG_THREADS, 0, st>>>(p);
    return check_launch("k2_wgrad_kernel");
}
