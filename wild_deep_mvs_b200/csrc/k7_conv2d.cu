// K7: 2-D convolution + folded BatchNorm + ReLU over channels-last maps, fp32 on the CUDA cores -- the building block of
// MVSNet's FeatureNet (models/MVSNet/model.py:21-41, ConvBnReLU of module.py:9-17): 3x3 stride-1 and 5x5 stride-2
// layers with 3..32 channels at image resolution.  (Row f1 of SURVEY.md 8: the extractor sits inside the same
// `forward` and, once the hot path is down to ~1.3 ms, takes more time than it when run as cuDNN calls on 8-channel
// tensors.)
//
// A CTA computes a 32-pixel-wide tile for all output channels.  A thread owns 8 output channels and FOUR pixels that
// lie 8 apart in x, so the lanes of a warp read ADJACENT pixels of the staged input (16-byte reads, conflict free) and
// every weight vector it loads is used for four pixels (128 FMAs per 12 shared-memory loads).  The input tile with
// its halo is staged channel-quad-major, the layer's weights [tap][ci][co] once per CTA.
#include "common.cuh"
#include <cstdint>

namespace mvsb200 {

constexpr int C2_THREADS = 256, C2_TW = 32, C2_PX = 4, C2_CT = 8;

struct C2Params {
    const float *x, *w, *scale, *bias;
    float *y;
    int B, H, W, Ho, Wo, relu, tiles_x, tiles_y;
};

template <int CIN, int COUT, int K, int S> struct C2Cfg {
    static constexpr int CG = COUT / C2_CT;                    // output-channel groups of 8
    static constexpr int TH = C2_THREADS * C2_PX / (C2_TW * CG);   // tile rows
    static constexpr int IH = (TH - 1) * S + K, IW = (C2_TW - 1) * S + K, NPOS = IH * IW;
    static constexpr int C4 = CIN / 4;
    // Layers whose whole filter + input tile exceed 64 KB of shared memory stage their weights one filter ROW at a time
    // (the 5x5 layers: 51 KB of weights next to an 81 KB input tile left ONE CTA per SM for 16 -> 32; 32 -> 32 3x3: 80 KB,
    // two per SM for a grid of 2.7 CTAs per SM); the others keep the whole filter resident
    static constexpr bool ROW_WEIGHTS = ((size_t)C4 * NPOS * 16 + (size_t)K * K * CIN * COUT * 4 > 64 * 1024);
    static constexpr int W_FLOATS = (ROW_WEIGHTS ? K : K * K) * CIN * COUT;
    static constexpr size_t SMEM = (size_t)C4 * NPOS * 16 + (size_t)W_FLOATS * 4;
    static_assert(CIN % 4 == 0 && COUT % C2_CT == 0 && TH >= 1, "unsupported channel counts");
};

template <int CIN, int COUT, int K, int S>
__global__ void __launch_bounds__(C2_THREADS) k7_conv2d_kernel(const C2Params p)
{
    using T = C2Cfg<CIN, COUT, K, S>;
    constexpr int CG = T::CG, C4 = T::C4, IW = T::IW, NPOS = T::NPOS, PAD = K / 2;
    extern __shared__ __align__(16) float4 c2_smem[];
    float4 *s_in = c2_smem;                                        // [C4][NPOS]
    float *s_w = reinterpret_cast<float *>(c2_smem + C4 * NPOS);   // [K*K][CIN][COUT]
    int t = blockIdx.x;
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    const int b = t / p.tiles_y;
    const int ox0 = tx * C2_TW, oy0 = ty * T::TH;
    const int ix0 = ox0 * S - PAD, iy0 = oy0 * S - PAD;

    if (!T::ROW_WEIGHTS)
        for (int i = threadIdx.x; i < T::W_FLOATS / 4; i += C2_THREADS)
            reinterpret_cast<float4 *>(s_w)[i] = __ldg(reinterpret_cast<const float4 *>(p.w) + i);
    const float *img = p.x + (long long)b * p.H * p.W * CIN;
    for (int i = threadIdx.x; i < NPOS * C4; i += C2_THREADS) {
        const int c4 = i % C4, pos = i / C4;   // consecutive threads: consecutive 16-byte pieces of a pixel
        const int gy = iy0 + pos / IW, gx = ix0 + pos % IW;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W) v = ldg4(img + ((long long)gy * p.W + gx) * CIN + c4 * 4);
        s_in[c4 * NPOS + pos] = v;
    }
    __syncthreads();

    // thread -> (output-channel group g, row ly, x phase qx); its pixels are x = qx + 8 j
    const int g = threadIdx.x % CG;
    const int q = threadIdx.x / CG;
    const int qx = q % (C2_TW / C2_PX), ly = q / (C2_TW / C2_PX);
    // accumulators as channel PAIRS: packed fp32 FMAs (FFMA2: two IEEE fp32 FMAs per instruction, same rounding as the
    // scalar form) with the weight pair as the vector operand -- the registers an LDS.128 fills are already pairs -- and
    // the input value as the scalar
    float2 acc[C2_PX][C2_CT / 2];
#pragma unroll
    for (int j = 0; j < C2_PX; j++)
#pragma unroll
        for (int c = 0; c < C2_CT / 2; c++) acc[j][c] = make_float2(0.f, 0.f);

#pragma unroll 1
    for (int ky = 0; ky < K; ky++) {
        if (T::ROW_WEIGHTS) {
            if (ky > 0) __syncthreads();      // the previous row's weights have been consumed
            for (int i = threadIdx.x; i < T::W_FLOATS / 4; i += C2_THREADS)
                reinterpret_cast<float4 *>(s_w)[i] = __ldg(reinterpret_cast<const float4 *>(p.w) + (size_t)ky * (T::W_FLOATS / 4) + i);
            __syncthreads();
        }
#pragma unroll
        for (int kx = 0; kx < K; kx++) {
            const float *wt = s_w + (((T::ROW_WEIGHTS ? 0 : ky * K) + kx) * CIN) * COUT + g * C2_CT;
            const int base = (ly * S + ky) * IW + qx * S + kx;
#pragma unroll
            for (int c4 = 0; c4 < C4; c4++) {
                float4 in[C2_PX];
#pragma unroll
                for (int j = 0; j < C2_PX; j++) in[j] = s_in[c4 * NPOS + base + j * 8 * S];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float4 w0 = *reinterpret_cast<const float4 *>(wt + (c4 * 4 + k) * COUT);
                    const float4 w1 = *reinterpret_cast<const float4 *>(wt + (c4 * 4 + k) * COUT + 4);
                    const float2 wv[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
#pragma unroll
                    for (int j = 0; j < C2_PX; j++) {
                        const float v = (k == 0) ? in[j].x : (k == 1) ? in[j].y : (k == 2) ? in[j].z : in[j].w;
                        const float2 vv = make_float2(v, v);
#pragma unroll
                        for (int c = 0; c < C2_CT / 2; c++) acc[j][c] = __ffma2_rn(vv, wv[c], acc[j][c]);
                    }
                }
            }
        }
    }

    const int oy = oy0 + ly;
    if (oy >= p.Ho) return;
    float sc[C2_CT], bi[C2_CT];
#pragma unroll
    for (int c = 0; c < C2_CT; c++) {
        sc[c] = p.scale ? __ldg(p.scale + g * C2_CT + c) : 1.f;
        bi[c] = p.bias ? __ldg(p.bias + g * C2_CT + c) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < C2_PX; j++) {
        const int ox = ox0 + qx + 8 * j;
        if (ox >= p.Wo) continue;
        float r[C2_CT];
#pragma unroll
        for (int c = 0; c < C2_CT; c++) {
            r[c] = fmaf((c & 1) ? acc[j][c / 2].y : acc[j][c / 2].x, sc[c], bi[c]);
            if (p.relu) r[c] = fmaxf(r[c], 0.f);
        }
        float *dst = p.y + (((long long)b * p.Ho + oy) * p.Wo + ox) * COUT + g * C2_CT;
        st4(dst, make_float4(r[0], r[1], r[2], r[3]));
        st4(dst + 4, make_float4(r[4], r[5], r[6], r[7]));
    }
}

template <int CIN, int COUT, int K, int S>
static int launch_c2(C2Params p, cudaStream_t st)
{
    using T = C2Cfg<CIN, COUT, K, S>;
    p.tiles_x = (p.Wo + C2_TW - 1) / C2_TW;
    p.tiles_y = (p.Ho + T::TH - 1) / T::TH;
    const long long blocks = (long long)p.tiles_x * p.tiles_y * p.B;
    if (blocks >= (1ll << 31)) {
        set_error("conv2d: image too large");
        return MVSB200_E_INVALID;
    }
    if (T::SMEM > 48 * 1024)
        if (int rc = ensure_dynamic_smem(k7_conv2d_kernel<CIN, COUT, K, S>, T::SMEM, "conv2d")) return rc;
    k7_conv2d_kernel<CIN, COUT, K, S><<<(unsigned)blocks, C2_THREADS, T::SMEM, st>>>(p);
    return check_launch("k7_conv2d_kernel");
}

template <int K, int S>
static int dispatch_c2(const C2Params &p, int cin, int cout, cudaStream_t st)
{
#define C2_CASE(ci, co) if (cin == ci && cout == co) return launch_c2<ci, co, K, S>(p, st)
    C2_CASE(4, 8); C2_CASE(8, 8); C2_CASE(8, 16); C2_CASE(16, 16); C2_CASE(16, 32); C2_CASE(32, 32);
    C2_CASE(4, 16); C2_CASE(4, 32); C2_CASE(8, 32); C2_CASE(16, 8); C2_CASE(32, 16); C2_CASE(32, 8);
#undef C2_CASE
    set_error("conv2d: unsupported channel counts Cin=%d Cout=%d (Cin in {4,8,16,32}, Cout in {8,16,32})", cin, cout);
    return MVSB200_E_INVALID;
}

// K7b: y = act(y * scale[c] + bias[c] (+ residual)) in place over a channels-last map (one pass over the map instead of the separate
// broadcast-add and activation kernels a convolution library call without a fused epilogue is followed by):
// the epilogue of CVP-MVSNet's `conv` = Conv2d(bias) + LeakyReLU(0.1) (models/CVP_MVSNet/models/modules.py:24-28).
// slope = 0 is ReLU, slope = 1 no activation.  One thread per 16 bytes; C % 4 == 0, so a vector holds channels c .. c+3.
__global__ void __launch_bounds__(256) k7_bias_act_kernel(float4 *y, long long n4, int C4, const float4 *scale, const float4 *bias,
                                                          const float4 *residual, float slope, float *amax)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    float vmax = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int c4 = (int)(i % C4);
        float4 v = y[i];
        const float4 b = bias ? __ldg(bias + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (scale) {
            const float4 s = __ldg(scale + c4);
            v.x = fmaf(v.x, s.x, b.x); v.y = fmaf(v.y, s.y, b.y); v.z = fmaf(v.z, s.z, b.z); v.w = fmaf(v.w, s.w, b.w);
        } else {
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        if (residual) {
            const float4 r = __ldg(residual + i);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
        v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
        y[i] = v;
        vmax = fmaxf(fmaxf(vmax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    if (amax) {   // max |y| for the z-march conv engine's operand scale (a zeroed device scalar)
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, m));
        if ((threadIdx.x & 31) == 0 && vmax > 0.f) atomicMax(reinterpret_cast<unsigned int *>(amax), __float_as_uint(vmax));
    }
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_bias_act(float *y, long long n_pixels, int C, const float *scale, const float *bias, const float *residual,
                                float slope, float *amax, mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(y && (bias || residual), "bias_act: null pointer");
    MVSB200_REQUIRE(n_pixels >= 0 && C > 0 && C % 4 == 0, "bias_act: C=%d must be a positive multiple of 4", C);
    MVSB200_REQUIRE(((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(scale) |
                      reinterpret_cast<uintptr_t>(residual)) & 15) == 0, "bias_act: y, scale, bias and residual must be 16-byte aligned");
    if (n_pixels == 0) return MVSB200_OK;
    const long long n4 = n_pixels * (C / 4);
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    long long blocks = (n4 + 255) / 256;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;      // grid-stride: 16 resident-sized waves of 256 threads per SM
    k7_bias_act_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4 *>(y), n4, C / 4,
                                                                          reinterpret_cast<const float4 *>(scale),
                                                                          reinterpret_cast<const float4 *>(bias),
                                                                          reinterpret_cast<const float4 *>(residual), slope, amax);
    return check_launch("k7_bias_act_kernel");
}

extern "C" int mvsb200_conv2d(int B, int H, int W, int Cin, int Cout, int k, int stride, int relu, const float *x, const float *w,
                              const float *scale, const float *bias, float *y, mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(x && w && y, "conv2d: null pointer");
    MVSB200_REQUIRE(B > 0 && H > 0 && W > 0, "conv2d: bad shape B=%d H=%d W=%d", B, H, W);
    MVSB200_REQUIRE((k == 3 && stride == 1) || (k == 5 && stride == 2), "conv2d: supported (kernel, stride): (3,1) and (5,2); got (%d,%d)", k, stride);
    C2Params p;
    p.x = x; p.w = w; p.scale = scale; p.bias = bias; p.y = y;
    p.B = B; p.H = H; p.W = W; p.relu = relu;
    p.Ho = (H + 2 * (k / 2) - k) / stride + 1;
    p.Wo = (W + 2 * (k / 2) - k) / stride + 1;
    p.tiles_x = p.tiles_y = 0;
    cudaStream_t st = (cudaStream_t)stream;
    return k == 3 ? dispatch_c2<3, 1>(p, Cin, Cout, st) : dispatch_c2<5, 2>(p, Cin, Cout, st);
}
