// K2 single-output-channel head: 3x3x3 convolution Cin -> 1 (MVSNet `prob`, Vis-MVSNet `final_conv`, CVP `prob0`:
// models/MVSNet/model.py:72, models/VisMVSNet/model_cas.py:44,61, CVP_MVSNet/models/net.py:67) on the CUDA cores, fp32.
//
// A 1-column GEMM wastes 7/8 of the narrowest tensor-core tile, and the layer is tiny in FLOPs (216 FMA per voxel for
// Cin = 8) but reads a full-resolution volume: it belongs on the FMA pipe, next to HBM.  One thread owns one (y,x)
// output column of an 8 x 32 tile and marches along z with three rolling accumulators (the planes z-1, z, z+1 that
// the current input plane contributes to), so every input plane is staged in shared memory exactly once per tile
// (cp.async, double buffered, channel-quad-major so the 16-byte reads of a warp are conflict free) and every staged
// value is used for three FMAs.  The 27*Cin weights travel as KERNEL PARAMETERS: the fully unrolled FMAs take them
// straight from the constant bank, no weight loads at all.
#include "common.cuh"

namespace mvsb200 {

constexpr int C1_TY = 8, C1_TX = 32, C1_THREADS = C1_TY * C1_TX;
constexpr int C1_EY = C1_TY + 2, C1_EX = C1_TX + 2, C1_NPOS = C1_EY * C1_EX;

template <int CIN> struct C1Params {
    const float *x;
    float *y;
    int B, D, H, W, zseg, nseg, tiles_x, tiles_y;
    float scale, bias;
    int relu;
    float w[27 * CIN];   // [tap][ci]
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
__global__ void __launch_bounds__(C1_THREADS) k2_conv3d_c1_kernel(const __grid_constant__ C1Params<CIN> p)
{
    constexpr int C4 = CIN / 4;
    extern __shared__ __align__(16) float4 c1_smem[];   // [2][C4][NPOS]
    int t = blockIdx.x;
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y; t /= p.tiles_y;
    const int sg = t % p.nseg;
    const int b = t / p.nseg;
    const int x0 = tx * C1_TX, y0 = ty * C1_TY;
    const int zb = sg * p.zseg, ze = min(zb + p.zseg, p.D);
    const int lx = threadIdx.x % C1_TX, ly = threadIdx.x / C1_TX;

    // stage input plane z (with its 1-voxel (y,x) halo, zero filled outside the volume) into buffer `buf`
    auto stage = [&](int z, int buf) {
        if ((unsigned)z < (unsigned)p.D) {
            const float *plane = p.x + ((long long)b * p.D + z) * p.H * p.W * CIN;
            for (int i = threadIdx.x; i < C1_NPOS * C4; i += C1_THREADS) {
                const int c4 = i % C4, pos = i / C4;     // consecutive threads: consecutive 16-byte pieces in global memory
                const int gy = y0 - 1 + pos / C1_EX, gx = x0 - 1 + pos % C1_EX;
                const bool ok = (unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W;
                const float *src = ok ? plane + ((long long)gy * p.W + gx) * CIN + c4 * 4 : plane;
                cp_async16_zfill(&c1_smem[(buf * C4 + c4) * C1_NPOS + pos], src, ok);
            }
        }
        cp_async_commit();
    };

    float a0 = 0.f, a1 = 0.f, a2 = 0.f;   // accumulators of output planes z-1, z, z+1 while input plane z is consumed
    stage(zb - 1, 0);
    for (int z = zb - 1, it = 0; z <= ze; z++, it++) {
        const int buf = it & 1;
        if (z < ze) stage(z + 1, buf ^ 1); else cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if ((unsigned)z < (unsigned)p.D) {
            const float4 *sp = c1_smem + buf * C4 * C1_NPOS + ly * C1_EX + lx;
#pragma unroll
            for (int dy = 0; dy < 3; dy++)
#pragma unroll
                for (int dx = 0; dx < 3; dx++)
#pragma unroll
                    for (int c4 = 0; c4 < C4; c4++) {
                        const float4 v = sp[c4 * C1_NPOS + dy * C1_EX + dx];
                        const int w2 = ((2 * 3 + dy) * 3 + dx) * CIN + c4 * 4;   // kz = 2 -> output plane z-1
                        const int w1 = ((1 * 3 + dy) * 3 + dx) * CIN + c4 * 4;   // kz = 1 -> output plane z
                        const int w0 = ((0 * 3 + dy) * 3 + dx) * CIN + c4 * 4;   // kz = 0 -> output plane z+1
                        a0 = fmaf(v.x, p.w[w2], a0); a0 = fmaf(v.y, p.w[w2 + 1], a0); a0 = fmaf(v.z, p.w[w2 + 2], a0); a0 = fmaf(v.w, p.w[w2 + 3], a0);
                        a1 = fmaf(v.x, p.w[w1], a1); a1 = fmaf(v.y, p.w[w1 + 1], a1); a1 = fmaf(v.z, p.w[w1 + 2], a1); a1 = fmaf(v.w, p.w[w1 + 3], a1);
                        a2 = fmaf(v.x, p.w[w0], a2); a2 = fmaf(v.y, p.w[w0 + 1], a2); a2 = fmaf(v.z, p.w[w0 + 2], a2); a2 = fmaf(v.w, p.w[w0 + 3], a2);
                    }
        }
        const int zo = z - 1;   // complete now
        if (zo >= zb && zo < ze) {
            const int oy = y0 + ly, ox = x0 + lx;
            if (oy < p.H && ox < p.W) {
                float r = fmaf(a0, p.scale, p.bias);
                if (p.relu) r = fmaxf(r, 0.f);
                p.y[(((long long)b * p.D + zo) * p.H + oy) * p.W + ox] = r;
            }
        }
        a0 = a1; a1 = a2; a2 = 0.f;
        __syncthreads();   // the buffer just read is refilled by the next iteration's prefetch
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
    p.tiles_y = (d->H + C1_TY - 1) / C1_TY;
    // depth segments: enough CTAs to fill the machine several times over, at most ~1/8 z-halo overhead when possible
    const long long base = (long long)p.tiles_x * p.tiles_y * d->B;
    int nseg = (int)((148 * 8 + base - 1) / base);
    if (nseg > (d->D + 15) / 16) nseg = (d->D + 15) / 16;
    if (nseg < 1) nseg = 1;
    p.zseg = (d->D + nseg - 1) / nseg;
    p.nseg = (d->D + p.zseg - 1) / p.zseg;
    const long long blocks = base * p.nseg;
    if (blocks >= (1ll << 31)) {
        set_error("conv3d_c1: volume too large");
        return MVSB200_E_INVALID;
    }
    const size_t smem = (size_t)2 * (CIN / 4) * C1_NPOS * sizeof(float4);
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
    cudaStream_t st = (cudaStream_t)stream;
    switch (d->Cin) {
    case 8: return launch_c1<8>(d, x, w_host, scale, bias, y, st);
    case 16: return launch_c1<16>(d, x, w_host, scale, bias, y, st);
    case 24: return launch_c1<24>(d, x, w_host, scale, bias, y, st);
    default: return launch_c1<32>(d, x, w_host, scale, bias, y, st);
    }
}
