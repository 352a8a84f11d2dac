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
// On the CUDA cores in fp32.  A thread owns NCA channels of `a`, one channel of `b` and all 27 taps (27 * NCA accumulators
// in registers): per staged x position it reads 27 values of `b` and NCA values of `a` from shared memory for 27 * NCA
// FMAs (one pair per thread is bound 4 : 1 by shared-memory reads; NCA = 2 halves that and still fits two blocks per SM).  The 256 threads of a block split into (x lanes) x (groups of NCA `a`
// channels) x (lanes over the `b` channels); the split is chosen per layer so that layers with few channel pairs (the
// 8 -> 1 head, 32 -> 8) spread over x instead of idling.  A block walks rows (b, z, y) of `a`; the row of `a` and the
// nine rows of `b` it touches are staged in x segments, double buffered with cp.async (segment q + 1 lands while q is consumed).  Blocks add their partial sums into R with atomics (R zeroed
// by the caller).  The dgrad of the same layers needs no kernel of its own: it is a forward call of the K2 engines
// with the weights re-packed (ops.conv3d_input_grad).
#include "common.cuh"
#include <cstdlib>

namespace mvsb200 {

constexpr int WG_THREADS = 256, WG_CAT = 32, WG_CBL = 32;   // largest channel tile of a block: 32 (a) x 32 (b)

struct WgradParams {
    const float *a, *b;
    float *r;
    int B, Da, Ha, Wa, Db, Hb, Wb, Ca, Cb;
    int cbl, cag, xl;          // lanes over b channels, groups of NCA a channels, x lanes: cbl * cag * xl == 256
    int b_vec4;                // Cb % 4 == 0 and b 16-byte aligned: stage b in 16-byte pieces
    int cbl_sh, cat_sh;        // log2(cbl), log2(cag * NCA): every split is a power of two, index arithmetic by shifts
    int tiles_a, tiles_b;
    long long rows;
};

// 4-byte asynchronous copy into shared memory, zero filled when !valid: the staging phase keeps dozens of loads per
// thread in flight without holding registers (the register-tiled kernel runs at one block per SM)
__device__ __forceinline__ void wg_cp_async4(float *smem, const float *gmem, bool valid)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sa), "l"(gmem), "r"(bytes) : "memory");
}

__device__ __forceinline__ void wg_cp_async16(float *smem, const float *gmem, bool valid)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(bytes) : "memory");
}

template <int S> struct WgTile {
    static constexpr int XS = S == 1 ? 32 : 16;          // x positions of `a` per staged segment (38 KB of b either way)
    static constexpr int XB = S * (XS - 1) + 3;          // staged x extent of `b` for one segment
    static constexpr int SA = XS * WG_CAT, SB = 9 * XB * WG_CBL;     // floats per buffer: a [x][cat], b [(dz,dy)][x][cbl]
    static constexpr size_t SMEM = (size_t)2 * (SA + SB) * sizeof(float);
};

template <int S, int NCA>
__global__ void __launch_bounds__(WG_THREADS) k2_wgrad_kernel(const WgradParams p)
{
    constexpr int WG_XS = WgTile<S>::XS, XB = WgTile<S>::XB, SA = WgTile<S>::SA, SB = WgTile<S>::SB;
    extern __shared__ __align__(16) float wg_smem[];     // two buffers: segment q + 1 is staged while segment q is consumed
    const int cbl = p.cbl, cat = p.cag * NCA;
    const int ta = blockIdx.y % p.tiles_a, tb = blockIdx.y / p.tiles_a;
    const int ca0 = ta * cat, cb0 = tb * cbl;
    const int csh = p.cbl_sh, ash = p.cat_sh;
    const int lb = threadIdx.x & (cbl - 1), g = (threadIdx.x >> csh) & (p.cag - 1), xlane = threadIdx.x / (cbl * p.cag);
    float acc[NCA][27];
#pragma unroll
    for (int j = 0; j < NCA; j++)
#pragma unroll
        for (int t = 0; t < 27; t++) acc[j][t] = 0.f;

    const int nsx = (p.Wa + WG_XS - 1) / WG_XS;
    const long long my_rows = blockIdx.x < p.rows ? (p.rows - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long nseg = my_rows * nsx;

    // stage segment q (row blockIdx.x + (q / nsx) * gridDim.x, x from (q % nsx) * XS) into buffer buf
    auto stage = [&](long long q, int buf) {
        const long long row = blockIdx.x + (q / nsx) * gridDim.x;
        const int x0 = (int)(q % nsx) * WG_XS;
        const int ya = (int)(row % p.Ha);
        const int za = (int)((row / p.Ha) % p.Da);
        const int bb = (int)(row / ((long long)p.Ha * p.Da));
        float *s_a = wg_smem + buf * (SA + SB), *s_b = s_a + SA;
        for (int i = threadIdx.x; i < WG_XS * cat; i += WG_THREADS) {      // a[bb, za, ya, x0 .. x0+XS), the tile's channels
            const int c = i & (cat - 1), x = i >> ash;
            const bool ok = x0 + x < p.Wa && ca0 + c < p.Ca;
            wg_cp_async4(&s_a[x * cat + c], ok ? p.a + ((((long long)bb * p.Da + za) * p.Ha + ya) * p.Wa + x0 + x) * p.Ca + ca0 + c : p.a, ok);
        }
        if (p.b_vec4) {      // Cb % 4 == 0: 16-byte pieces (4 channels), a quarter of the copies and of their index arithmetic
            const int q4 = cbl >> 2, qsh = csh - 2;
            for (int i = threadIdx.x; i < 9 * XB * q4; i += WG_THREADS) {
                const int c = (i & (q4 - 1)) << 2, x = (i >> qsh) % XB, rr = (i >> qsh) / XB;
                const int zb = S * za + rr / 3 - 1, yb = S * ya + rr % 3 - 1, xb = S * x0 - 1 + x;
                const bool ok = (unsigned)zb < (unsigned)p.Db && (unsigned)yb < (unsigned)p.Hb && (unsigned)xb < (unsigned)p.Wb && cb0 + c < p.Cb;
                wg_cp_async16(&s_b[(rr * XB + x) * cbl + c], ok ? p.b + ((((long long)bb * p.Db + zb) * p.Hb + yb) * p.Wb + xb) * p.Cb + cb0 + c : p.b, ok);
            }
        } else {
            for (int i = threadIdx.x; i < 9 * XB * cbl; i += WG_THREADS) {     // the nine (dz, dy) rows of b, x from S*x0 - 1
                const int c = i & (cbl - 1), x = (i >> csh) % XB, rr = (i >> csh) / XB;
                const int zb = S * za + rr / 3 - 1, yb = S * ya + rr % 3 - 1, xb = S * x0 - 1 + x;
                const bool ok = (unsigned)zb < (unsigned)p.Db && (unsigned)yb < (unsigned)p.Hb && (unsigned)xb < (unsigned)p.Wb && cb0 + c < p.Cb;
                wg_cp_async4(&s_b[i], ok ? p.b + ((((long long)bb * p.Db + zb) * p.Hb + yb) * p.Wb + xb) * p.Cb + cb0 + c : p.b, ok);
            }
        }
    };

    if (nseg > 0) stage(0, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (long long q = 0; q < nseg; q++) {
        const int buf = (int)(q & 1);
        if (q + 1 < nseg) stage(q + 1, buf ^ 1);
        asm volatile("cp.async.commit_group;\n cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        const float *s_a = wg_smem + buf * (SA + SB), *s_b = s_a + SA;
        const int x0 = (int)(q % nsx) * WG_XS;
        const int nx = min(WG_XS, p.Wa - x0);
        for (int x = xlane; x < nx; x += p.xl) {
            float av[NCA];
#pragma unroll
            for (int j = 0; j < NCA; j++) av[j] = s_a[x * cat + g * NCA + j];
            const float *sb = s_b + (S * x) * cbl + lb;
#pragma unroll
            for (int rr = 0; rr < 9; rr++)
#pragma unroll
                for (int dx = 0; dx < 3; dx++) {
                    const float bv = sb[(rr * XB + dx) * cbl];
#pragma unroll
                    for (int j = 0; j < NCA; j++) acc[j][rr * 3 + dx] = fmaf(av[j], bv, acc[j][rr * 3 + dx]);
                }
        }
        __syncthreads();   // the buffer just read is refilled by the next iteration's prefetch
    }
    if (cb0 + lb < p.Cb) {
#pragma unroll
        for (int j = 0; j < NCA; j++) {
            const int ca = ca0 + g * NCA + j;
            if (ca >= p.Ca) continue;
            float *dst = p.r + ((long long)ca * p.Cb + cb0 + lb) * 27;
#pragma unroll
            for (int t = 0; t < 27; t++) atomicAdd(dst + t, acc[j][t]);
        }
    }
}

template <int S, int NCA>
static int launch_wgrad(const WgradParams &p, dim3 grid, cudaStream_t st)
{
    if (int rc = ensure_dynamic_smem(k2_wgrad_kernel<S, NCA>, WgTile<S>::SMEM, "conv3d_wgrad")) return rc;
    k2_wgrad_kernel<S, NCA><<<grid, WG_THREADS, WgTile<S>::SMEM, st>>>(p);
    return check_launch("k2_wgrad_kernel");
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
    p.B = B; p.Da = Da; p.Ha = Ha; p.Wa = Wa; p.Db = Db; p.Hb = Hb; p.Wb = Wb; p.Ca = Ca; p.Cb = Cb;
    // thread split: lanes over b channels (8, 16 or 32), groups of NCA a channels (tile of at most 32), the rest over x
    // NCA = 2 measured best on B200 (two blocks per SM overlap staging and FMAs; NCA = 4 runs one block of 156 registers
    // per SM and waits on its own staging): conv0 of MVSNet at the cfg1 volume 1.47 ms against 1.84 (NCA = 4), 1.33 (NCA = 1)
    int nca = Ca >= 2 ? 2 : 1;
    if (const char *e = getenv("MVSB200_WG_NCA")) { const int v = atoi(e); if ((v == 2 || v == 4) && Ca >= v) nca = v; }
    p.cbl = Cb <= 8 ? 8 : Cb <= 16 ? 16 : 32;
    int cag = (Ca + nca - 1) / nca;
    if (cag > WG_CAT / nca) cag = WG_CAT / nca;
    while (cag & (cag - 1)) cag++;                       // power of two, so that the split divides 256
    if (cag * p.cbl > WG_THREADS) cag = WG_THREADS / p.cbl;
    p.cag = cag;
    p.xl = WG_THREADS / (p.cbl * p.cag);
    p.cbl_sh = p.cbl == 8 ? 3 : p.cbl == 16 ? 4 : 5;
    p.b_vec4 = (Cb % 4 == 0 && ((unsigned long long)b & 15) == 0) ? 1 : 0;
    p.cat_sh = 0;
    while ((1 << p.cat_sh) < cag * nca) p.cat_sh++;
    p.tiles_a = (Ca + cag * nca - 1) / (cag * nca);
    p.tiles_b = (Cb + p.cbl - 1) / p.cbl;
    p.rows = (long long)B * Da * Ha;
    const int ctiles = p.tiles_a * p.tiles_b;
    long long gx = (148ll * (nca == 1 ? 2 : 4) + ctiles - 1) / ctiles;      // two to four blocks per SM over all channel tiles
    if (gx > p.rows) gx = p.rows;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)ctiles);
    cudaStream_t st = (cudaStream_t)stream;
    if (stride == 1) return nca == 4 ? launch_wgrad<1, 4>(p, grid, st) : nca == 2 ? launch_wgrad<1, 2>(p, grid, st) : launch_wgrad<1, 1>(p, grid, st);
    return nca == 4 ? launch_wgrad<2, 4>(p, grid, st) : nca == 2 ? launch_wgrad<2, 2>(p, grid, st) : launch_wgrad<2, 1>(p, grid, st);
}
