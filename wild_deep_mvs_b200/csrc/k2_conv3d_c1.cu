// K2 single-output-channel head: 3x3x3 convolution Cin -> 1 (MVSNet `prob`, Vis-MVSNet `final_conv`, CVP `prob0`:
// models/MVSNet/model.py:72, models/VisMVSNet/model_cas.py:44,61, CVP_MVSNet/models/net.py:67) on the CUDA cores, fp32.
//
// A 1-column GEMM wastes 7/8 of the narrowest tensor-core tile, and the layer is tiny in FLOPs (216 FMA per voxel for
// Cin = 8) but reads a full-resolution volume: it belongs on the FMA pipe, next to HBM.  One thread owns TWO vertically
// adjacent (y,x) output columns of a 16 x 32 tile (Cin = 8; one column of an 8 x 32 tile for wider inputs) and marches along z with three rolling accumulators per column (the
// planes z-1, z, z+1 that the current input plane contributes to), so every input plane is staged in shared memory
// exactly once per tile (cp.async, double buffered, channel-quad-major so the 16-byte reads of a warp are conflict
// free), the two columns share two of their three tap rows (24 instead of 36 shared-memory reads per pair) and every
// weight fetched from the constant bank feeds two FMAs.  The 27*Cin weights travel as KERNEL PARAMETERS, no weight
// loads at all.  The kernel is bound by instruction issue, so the per-plane staging does no index arithmetic either:
// a thread's (global offset, shared slot) pairs are the same for every plane and are computed once.
#include "common.cuh"

namespace mvsb200 {

constexpr int C1_TX = 32, C1_THREADS = 8 * C1_TX, C1_EX = C1_TX + 2;
#ifndef C1_NBUF
#define C1_NBUF 3      // staged plane buffers per CTA (>= 3: the plane being refilled is never the one being read)
#endif
#ifndef C1_ROWS8
#define C1_ROWS8 2     // output columns per thread for Cin = 8 (experiment switch: 4 = 9 instead of 12 shared loads per output, 2 CTAs / SM)
#endif
#ifndef C1_MIN_BLOCKS
#define C1_MIN_BLOCKS 3
#endif
// ROWS = vertically adjacent output columns per thread (rows ROWS*ly .. ROWS*ly + ROWS-1 of an 8*ROWS x 32 tile): 2 for
// Cin = 8; wider inputs keep 1 (two columns of 16+ channels do not fit the register file at 3 CTAs / SM)
template <int CIN> struct C1Tile {
    static constexpr int ROWS = CIN == 8 ? C1_ROWS8 : 1;
    static constexpr int TY = 8 * ROWS, EY = TY + 2, NPOS = EY * C1_EX;
    static constexpr int NBUF = C1_NBUF;
};

template <int CIN> struct C1Params {
    const float *x;
    float *y;
    int B, D, H, W, zseg, nseg, tiles_x, tiles_y;
    float scale, bias;
    int relu;
    alignas(16) float w[27 * CIN];   // [tap][ci]: the four weights of a channel quad are one 16-byte uniform load
};

__device__ __forceinline__ void cp_async16_zfill(void *smem, const void *gmem, bool valid)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int CIN>
__global__ void __launch_bounds__(C1_THREADS, C1_MIN_BLOCKS) k2_conv3d_c1_kernel(const __grid_constant__ C1Params<CIN> p)
{
    constexpr int C4 = CIN / 4, ROWS = C1Tile<CIN>::ROWS, C1_TY = C1Tile<CIN>::TY, C1_NPOS = C1Tile<CIN>::NPOS;
    extern __shared__ __align__(16) float4 c1_smem[];   // [C1_NBUF][C4][NPOS]
    int t = blockIdx.x;
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y; t /= p.tiles_y;
    const int sg = t % p.nseg;
    const int b = t / p.nseg;
    const int x0 = tx * C1_TX, y0 = ty * C1_TY;
    const int zb = sg * p.zseg, ze = min(zb + p.zseg, p.D);
    const int lx = threadIdx.x % C1_TX, ly = threadIdx.x / C1_TX;

    // The pieces (16 bytes = 4 channels of one halo position) this thread stages are the same for every plane: element
    // offset inside the plane (-1: outside the volume, zero filled) and shared-memory slot (-1: no piece).
    constexpr int NIT = (C1_NPOS * C4 + C1_THREADS - 1) / C1_THREADS;
    int goff[NIT], soff[NIT];
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int i = threadIdx.x + it * C1_THREADS;
        const int c4 = i % C4, pos = i / C4;     // consecutive threads: consecutive 16-byte pieces in global memory
        const int gy = y0 - 1 + pos / C1_EX, gx = x0 - 1 + pos % C1_EX;
        const bool ok = (unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W;
        goff[it] = ok ? (gy * p.W + gx) * CIN + c4 * 4 : -1;
        soff[it] = i < C1_NPOS * C4 ? c4 * C1_NPOS + pos : -1;
    }
    // stage input plane z (with its 1-voxel (y,x) halo, zero filled outside the volume) into buffer `buf`
    auto stage = [&](int z, int buf) {
        if ((unsigned)z < (unsigned)p.D) {
            const float *plane = p.x + ((long long)b * p.D + z) * p.H * p.W * CIN;
            float4 *dst = c1_smem + buf * C4 * C1_NPOS;
#pragma unroll
            for (int it = 0; it < NIT; it++)
                if (soff[it] >= 0) cp_async16_zfill(dst + soff[it], plane + max(goff[it], 0), goff[it] >= 0);
        }
        cp_async_commit();
    };

    // acc[j][k]: output column j of this thread, output plane z-1+k, while input plane z is consumed -- as TWO partial sums
    // (.x: channels 0, 1 of every quad, .y: channels 2, 3 ... see below), added when the plane is complete
    float2 acc[ROWS][3];
#pragma unroll
    for (int j = 0; j < ROWS; j++) acc[j][0] = acc[j][1] = acc[j][2] = make_float2(0.f, 0.f);
    // C1_NBUF plane buffers: the copies of planes z+1 ... z+NBUF-1 are in flight while plane z is consumed (with two
    // buffers the loop waited a full L2 / HBM round trip per plane: halving its FMA instructions did not move its time)
    constexpr int NBUF = C1Tile<CIN>::NBUF;
    static_assert(NBUF >= 3, "the single barrier per plane needs the refilled buffer to differ from the one being read");
#pragma unroll
    for (int i = 0; i < NBUF - 1; i++) {
        if (zb - 1 + i <= ze) stage(zb - 1 + i, i); else cp_async_commit();
    }
    for (int z = zb - 1, it = 0; z <= ze; z++, it++) {
        const int buf = it % NBUF;
        cp_async_wait<NBUF - 2>();      // plane z has landed (this thread's pieces) ...
        __syncthreads();                // ... for every thread; and every thread is done reading plane z-1,
        // whose buffer is the one refilled now: ONE barrier per plane (needs NBUF >= 3)
        if (z + NBUF - 1 <= ze) stage(z + NBUF - 1, (it + NBUF - 1) % NBUF); else cp_async_commit();
        if ((unsigned)z < (unsigned)p.D) {
            const float4 *sp = c1_smem + buf * C4 * C1_NPOS + (ROWS * ly) * C1_EX + lx;
#pragma unroll
            for (int r = 0; r < ROWS + 2; r++)      // halo row ROWS*ly + r is tap row r - j of output column j
#pragma unroll
                for (int dx = 0; dx < 3; dx++)
#pragma unroll
                    for (int c4 = 0; c4 < C4; c4++) {
                        const float4 v = sp[c4 * C1_NPOS + r * C1_EX + dx];
#pragma unroll
                        for (int j = 0; j < ROWS; j++) {
                            const int dy = r - j;
                            if (dy < 0 || dy > 2) continue;
#pragma unroll
                            for (int k = 0; k < 3; k++) {   // kz = 2 - k feeds output plane z-1+k
                                // packed fp32 FMAs (FFMA2): channel pairs (x, y) and (z, w) of the staged vector against the
                                // matching weight pairs, which arrive as ONE uniform 16-byte load per (tap, quad) and feed
                                // both columns of the thread: half the FMA instructions of the scalar form
                                const int w = (((2 - k) * 3 + dy) * 3 + dx) * CIN + c4 * 4;
                                acc[j][k] = __ffma2_rn(make_float2(v.x, v.y), make_float2(p.w[w], p.w[w + 1]), acc[j][k]);
                                acc[j][k] = __ffma2_rn(make_float2(v.z, v.w), make_float2(p.w[w + 2], p.w[w + 3]), acc[j][k]);
                            }
                        }
                    }
        }
        const int zo = z - 1;   // complete now
        if (zo >= zb && zo < ze) {
            const int oy = y0 + ROWS * ly, ox = x0 + lx;
            if (ox < p.W) {
                float *dst = p.y + (((long long)b * p.D + zo) * p.H + oy) * p.W + ox;
#pragma unroll
                for (int j = 0; j < ROWS; j++) {
                    float r = fmaf(acc[j][0].x + acc[j][0].y, p.scale, p.bias);
                    if (p.relu) r = fmaxf(r, 0.f);
                    if (oy + j < p.H) dst[(long long)j * p.W] = r;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < ROWS; j++) { acc[j][0] = acc[j][1]; acc[j][1] = acc[j][2]; acc[j][2] = make_float2(0.f, 0.f); }
    }
}

template <int CIN>
static int launch_c1(const mvsb200_conv3d_desc *d, const float *x, const float *w_host, float scale, float bias, float *y, cudaStream_t st)
{
    C1Params<CIN> p;
    p.x = x; p.y = y;
    p.B = d->B; p.D = d->D; p.H = d->H; p.W = d->W;
    p.scale = scale; p.bias = bias; p.relu = d->relu;
    for (int i = 0; i < 27 * CIN; i++) p.w[i] = w_host[i];
    p.tiles_x = (d->W + C1_TX - 1) / C1_TX;
    p.tiles_y = (d->H + C1Tile<CIN>::TY - 1) / C1Tile<CIN>::TY;
    // depth segments: trade the z-halo (2 extra planes per segment) against the balance of the last wave of CTAs
    const size_t smem = (size_t)C1Tile<CIN>::NBUF * (CIN / 4) * C1Tile<CIN>::NPOS * sizeof(float4);
    const long long base = (long long)p.tiles_x * p.tiles_y * d->B;
    int resident = (int)((220 * 1024) / smem);
    if (resident > C1_MIN_BLOCKS) resident = C1_MIN_BLOCKS;   // __launch_bounds__(C1_THREADS, C1_MIN_BLOCKS)
    const long long slots = 148ll * resident;
    double best = -1.0;
    p.zseg = d->D;
    for (int zseg = d->D < 4 ? d->D : 4; zseg <= d->D; zseg++) {
        const int nseg = (d->D + zseg - 1) / zseg;
        const long long ctas = base * nseg;
        const double eff = (double)ctas / (double)((ctas + slots - 1) / slots * slots) * (double)d->D / (double)(nseg * (zseg + 2));
        if (eff > best) { best = eff; p.zseg = zseg; }
    }
    p.nseg = (d->D + p.zseg - 1) / p.zseg;
    const long long blocks = base * p.nseg;
    if (blocks >= (1ll << 31)) {
        set_error("conv3d_c1: volume too large");
        return MVSB200_E_INVALID;
    }
    if (smem > 48 * 1024)
        if (int rc = ensure_dynamic_smem(k2_conv3d_c1_kernel<CIN>, smem, "conv3d_c1")) return rc;
    k2_conv3d_c1_kernel<CIN><<<(unsigned)blocks, C1_THREADS, smem, st>>>(p);
    return check_launch("k2_conv3d_c1_kernel");
}

static bool c1_shape_ok(const mvsb200_conv3d_desc *d)
{
    return d->kd == 3 && d->kh == 3 && d->kw == 3 && d->stride == 1 && !d->transposed && d->Cout == 1 && d->Cin2 == 0 &&
           (d->Cin == 8 || d->Cin == 16 || d->Cin == 24 || d->Cin == 32) && d->skip_mode == MVSB200_SKIP_NONE;
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_conv3d_c1_supported(const mvsb200_conv3d_desc *d) { return d && c1_shape_ok(d) ? 1 : 0; }

extern "C" int mvsb200_conv3d_c1(const mvsb200_conv3d_desc *d, const float *x, const float *w_host, float scale, float bias,
                                 float *y, mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(d && x && w_host && y, "conv3d_c1: null pointer");
    MVSB200_REQUIRE(c1_shape_ok(d), "conv3d_c1: needs a 3x3x3 stride-1 conv with Cout == 1, Cin in {8,16,24,32}, one input, no skip");
    MVSB200_REQUIRE(d->B > 0 && d->D > 0 && d->H > 0 && d->W > 0, "conv3d_c1: bad shape B=%d D=%d H=%d W=%d", d->B, d->D, d->H, d->W);
    MVSB200_REQUIRE((long long)d->H * d->W * d->Cin < (1ll << 31), "conv3d_c1: plane too large (%dx%dx%d)", d->H, d->W, d->Cin);
    cudaStream_t st = (cudaStream_t)stream;
    switch (d->Cin) {
    case 8: return launch_c1<8>(d, x, w_host, scale, bias, y, st);
    case 16: return launch_c1<16>(d, x, w_host, scale, bias, y, st);
    case 24: return launch_c1<24>(d, x, w_host, scale, bias, y, st);
    default: return launch_c1<32>(d, x, w_host, scale, bias, y, st);
    }
}
