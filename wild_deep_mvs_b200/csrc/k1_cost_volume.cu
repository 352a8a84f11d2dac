// K1: fused homography warp + cross-view aggregation (variance / softmin / group correlation).
//
// Layout: features NHWC so that each bilinear tap is one contiguous C*4-byte vector; a group of
// LPV = C/4 lanes owns one reference pixel, each lane one float4 of channels.  A warp therefore covers
// 32/LPV adjacent pixels and every tap is a 16-byte load per lane, fully coalesced per pixel.
// The depth axis is walked in chunks of DCH hypotheses whose running aggregates stay in registers while
// the source views are visited one after the other (view-outer, depth-inner), so nothing C-wide except
// the final cost volume is ever written.  The projection / tap geometry of a (pixel, hypothesis, view) does not
// depend on the channel, so the LPV lanes of a pixel split the chunk's hypotheses between them and exchange the
// packed taps by warp shuffle; the four taps are re-loaded only when the 2x2 cell changes between hypotheses.
//
// Sampling semantics follow the reference bit for bit in structure (see oracle/mvs_oracle.c):
//   MVS geometry  models/MVSNet/module.py:138-155   (integer pixel grid, z<=0 -> (-10,-10), clamp +-10)
//   VIS geometry  models/VisMVSNet/homography.py:77-121 (pixel centres +0.5, normalise by size, clamp +-1.1)
//   grid_sample(bilinear, zeros, align_corners=True): per-tap zero padding.
#include "common.cuh"

namespace mvsb200 {

constexpr int K1_DCH = 8;       // hypotheses per thread
constexpr int K1_THREADS = 256;

struct K1Params {
    const float *ref;
    const float *src[MVSB200_MAX_SRC];
    int src_h[MVSB200_MAX_SRC], src_w[MVSB200_MAX_SRC];
    const float *warp;
    const float *depth;
    const float *interval;
    const float *temp;
    float *out;
    float *out_amax;
    long long out_view_stride;
    int B, S, D, H, W, depth_mode;
};

struct Taps {
    long long o00;   // element offset of the north-west tap (pixel index, not yet times C)
    float w00, w01, w10, w11;
    int dx, dy;      // offsets (in pixels) to the east / south taps
};

// Normalised grid coordinate -> four taps with per-tap zero padding folded into the weights.
__device__ __forceinline__ Taps make_taps(float gx, float gy, int Hs, int Ws)
{
    float ix = ((gx + 1.f) / 2.f) * (float)(Ws - 1);
    float iy = ((gy + 1.f) / 2.f) * (float)(Hs - 1);
    float fx = floorf(ix), fy = floorf(iy);
    int x0 = (int)fx, y0 = (int)fy;
    float x1 = fx + 1.f, y1 = fy + 1.f;
    Taps t;
    t.w00 = (x1 - ix) * (y1 - iy);
    t.w01 = (ix - fx) * (y1 - iy);
    t.w10 = (x1 - ix) * (iy - fy);
    t.w11 = (ix - fx) * (iy - fy);
    bool xin0 = (x0 >= 0) & (x0 < Ws), xin1 = (x0 + 1 >= 0) & (x0 + 1 < Ws);
    bool yin0 = (y0 >= 0) & (y0 < Hs), yin1 = (y0 + 1 >= 0) & (y0 + 1 < Hs);
    if (!(xin0 & yin0)) t.w00 = 0.f;
    if (!(xin1 & yin0)) t.w01 = 0.f;
    if (!(xin0 & yin1)) t.w10 = 0.f;
    if (!(xin1 & yin1)) t.w11 = 0.f;
    // clamp the addresses into the map; out-of-range taps carry weight 0
    const int xa = min(max(x0, 0), Ws - 1), xb = min(max(x0 + 1, 0), Ws - 1);
    const int ya = min(max(y0, 0), Hs - 1), yb = min(max(y0 + 1, 0), Hs - 1);
    t.dx = xb - xa;
    t.dy = yb - ya;
    t.o00 = (long long)ya * Ws + xa;
    return t;
}

// x / d for a divisor whose correctly rounded reciprocal r is known: one residual correction of x*r gives the
// correctly rounded quotient for normal operands (the fast path of the IEEE division sequence, without its range
// check and slow-path call -- the epilogue divides 64 values per thread by V and V^2).
__device__ __forceinline__ float div_by(float x, float d, float r)
{
    const float q = x * r;
    return fmaf(fmaf(-q, d, x), r, q);
}

// Packed taps as they travel between lanes: cell = clamped north-west pixel index with the east/south steps in the
// two top bits (maps are far below 2^29 pixels), plus the four zero-padding-folded weights.
struct PackedTaps {
    int cell;
    float w00, w01, w10, w11;
};

__device__ __forceinline__ PackedTaps pack_taps(const Taps &t)
{
    PackedTaps q;
    q.cell = (int)t.o00 | (t.dx << 29) | (t.dy << 30);
    q.w00 = t.w00; q.w01 = t.w01; q.w10 = t.w10; q.w11 = t.w11;
    return q;
}

template <int C, int GEOM, int AGG>
__global__ void __launch_bounds__(K1_THREADS) k1_cost_volume_kernel(const K1Params p)
{
    constexpr int LPV = C / 4;              // lanes per voxel
    constexpr int VPB = K1_THREADS / LPV;   // pixels per block
    constexpr int KPL = K1_DCH / LPV;       // hypotheses whose geometry each lane of a pixel group computes
    static_assert(K1_DCH % LPV == 0, "depth chunk must split evenly over the lanes of a pixel group");
    __shared__ float s_warp[MVSB200_MAX_SRC * 16];

    const int b = blockIdx.z;
    const int d0 = blockIdx.y * K1_DCH;
    const long long HW = (long long)p.H * p.W;
    for (int i = threadIdx.x; i < p.S * 16; i += K1_THREADS) s_warp[i] = p.warp[(long long)b * p.S * 16 + i];
    __syncthreads();

    const int sub = threadIdx.x % LPV;
    const long long pix_raw = (long long)blockIdx.x * VPB + threadIdx.x / LPV;
    const bool active = pix_raw < HW;                 // inactive groups shadow the last pixel and store nothing,
    const long long pix = active ? pix_raw : HW - 1;  // so every shuffle below runs with the full warp
    const int y = (int)(pix / p.W), x = (int)(pix % p.W);

    const float4 r = ldg4(p.ref + ((long long)b * HW + pix) * C + sub * 4);
    const float interval = (p.depth_mode >= MVSB200_DEPTH_START) ? __ldg(p.interval + b) : 0.f;

    // The projection of a (pixel, hypothesis, view) is the same for every channel: lane `sub` of the pixel group
    // computes it for hypotheses sub, sub+LPV, ... of the chunk and the group shares the taps by shuffle.
    float dv[KPL];
#pragma unroll
    for (int j = 0; j < KPL; j++) {
        int d = min(d0 + j * LPV + sub, p.D - 1);
        dv[j] = hypothesis(p.depth_mode, p.depth, interval, b, d, p.D, HW, pix);
    }

    float4 acc1[K1_DCH], acc2[K1_DCH];
    float sum_exp[K1_DCH];
#pragma unroll
    for (int k = 0; k < K1_DCH; k++) {
        if (AGG == MVSB200_AGG_VARIANCE || AGG == MVSB200_AGG_VARIANCE_MEAN) {
            acc1[k] = r;
            acc2[k] = make_float4(r.x * r.x, r.y * r.y, r.z * r.z, r.w * r.w);
        } else {
            acc1[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            acc2[k] = acc1[k];
        }
        sum_exp[k] = 0.f;
    }
    const float temp = (AGG == MVSB200_AGG_SOFTMIN) ? __ldg(p.temp) : 0.f;
    float vmax = 0.f;   // max |stored value| of this thread (abs-max tracking for the z-march conv engine)

    for (int s = 0; s < p.S; s++) {
        const float *wp = s_warp + s * 16;
        const int Hs = p.src_h[s], Ws = p.src_w[s];
        const float *map = p.src[s] + (long long)b * Hs * Ws * C;   // warp-uniform base, 32-bit element offsets below
        float ax, ay, az, np_ = 0.f;
        if (GEOM == MVSB200_GEOM_MVS) {
            const float fx = (float)x, fy = (float)y;
            ax = wp[0] * fx + wp[1] * fy + wp[2];
            ay = wp[3] * fx + wp[4] * fy + wp[5];
            az = wp[6] * fx + wp[7] * fy + wp[8];
        } else {
            const float fx = (float)x + 0.5f, fy = (float)y + 0.5f;
            ax = wp[0] * fx + wp[1] * fy + wp[2];
            ay = wp[3] * fx + wp[4] * fy + wp[5];
            az = wp[6] * fx + wp[7] * fy + wp[8];
            np_ = wp[12] * fx + wp[13] * fy + wp[14];
        }
        const float bx = wp[9], by = wp[10], bz = wp[11];
        PackedTaps own[KPL];
#pragma unroll
        for (int j = 0; j < KPL; j++) {
            float gx, gy;
            if (GEOM == MVSB200_GEOM_MVS) {
                float qx = ax * dv[j] + bx, qy = ay * dv[j] + by, qz = az * dv[j] + bz;
                float px = qx / qz, py = qy / qz;
                if (qz <= 0.f) px = -10.f, py = -10.f;
                gx = clampf(px / ((float)(Ws - 1) / 2.f) - 1.f, -10.f, 10.f);
                gy = clampf(py / ((float)(Hs - 1) / 2.f) - 1.f, -10.f, 10.f);
            } else {
                float f = np_ / (dv[j] + 1e-9f);
                float qx = ax - bx * f, qy = ay - by * f, qz = az - bz * f;
                float zc = fmaxf(qz, 1e-9f);
                float u = qx / zc, v = qy / zc;
                if (!(qz > 0.f)) u = -10.f, v = -10.f;
                gx = clampf((u / (float)Ws) * 2.f - 1.f, -1.1f, 1.1f);
                gy = clampf((v / (float)Hs) * 2.f - 1.f, -1.1f, 1.1f);
            }
            // NaN coordinates (degenerate cameras) sample nothing
            if (!(gx == gx) || !(gy == gy)) gx = gy = -10.f;
            own[j] = pack_taps(make_taps(gx, gy, Hs, Ws));
        }

        // Consecutive hypotheses of a pixel move along the epipolar line by a fraction of a pixel, so they mostly
        // fall into the same 2x2 tap cell: the four 16-byte taps stay in registers until the cell changes.
        int cur_cell = -1;
        float4 ta = make_float4(0.f, 0.f, 0.f, 0.f), tb = ta, tc = ta, td = ta;
#pragma unroll
        for (int k = 0; k < K1_DCH; k++) {
            const int owner = k % LPV, j = k / LPV;
            const int cell = __shfl_sync(0xffffffffu, own[j].cell, owner, LPV);
            const float w00 = __shfl_sync(0xffffffffu, own[j].w00, owner, LPV);
            const float w01 = __shfl_sync(0xffffffffu, own[j].w01, owner, LPV);
            const float w10 = __shfl_sync(0xffffffffu, own[j].w10, owner, LPV);
            const float w11 = __shfl_sync(0xffffffffu, own[j].w11, owner, LPV);
            if (cell != cur_cell) {
                cur_cell = cell;
                const unsigned o00 = (unsigned)(cell & 0x1fffffff) * C + sub * 4;
                const unsigned ox = ((cell >> 29) & 1) * C, oy = ((cell >> 30) & 1) * (unsigned)(Ws * C);
                ta = ldg4(map + o00);
                tb = ldg4(map + (o00 + ox));
                tc = ldg4(map + (o00 + oy));
                td = ldg4(map + (o00 + oy + ox));
            }
            float4 w;
            // accumulation order nw, ne, sw, se (ATen grid_sampler_2d)
            w.x = ta.x * w00; w.y = ta.y * w00; w.z = ta.z * w00; w.w = ta.w * w00;
            w.x += tb.x * w01; w.y += tb.y * w01; w.z += tb.z * w01; w.w += tb.w * w01;
            w.x += tc.x * w10; w.y += tc.y * w10; w.z += tc.z * w10; w.w += tc.w * w10;
            w.x += td.x * w11; w.y += td.y * w11; w.z += td.z * w11; w.w += td.w * w11;
            if (AGG == MVSB200_AGG_VARIANCE || AGG == MVSB200_AGG_VARIANCE_MEAN) {
                acc1[k].x += w.x; acc1[k].y += w.y; acc1[k].z += w.z; acc1[k].w += w.w;
                acc2[k].x += w.x * w.x; acc2[k].y += w.y * w.y; acc2[k].z += w.z * w.z; acc2[k].w += w.w * w.w;
            } else if (AGG == MVSB200_AGG_SOFTMIN) {
                float4 df = make_float4(w.x - r.x, w.y - r.y, w.z - r.z, w.w - r.w);
                df.x *= df.x; df.y *= df.y; df.z *= df.z; df.w *= df.w;
                float ssd = (df.x + df.y) + (df.z + df.w);
#pragma unroll
                for (int m = LPV / 2; m >= 1; m >>= 1) ssd += __shfl_xor_sync(0xffffffffu, ssd, m);
                float e = expf(-temp * ssd);
                sum_exp[k] += e;
                acc1[k].x += df.x * e; acc1[k].y += df.y * e; acc1[k].z += df.z * e; acc1[k].w += df.w * e;
            } else {  // GROUPCORR: one lane == one group of 4 channels; one output volume per source view
                float g = r.x * w.x;
                g += r.y * w.y;
                g += r.z * w.z;
                g += r.w * w.w;
                if (active && d0 + k < p.D) {
                    vmax = fmaxf(vmax, fabsf(g));
                    __stcs(p.out + s * p.out_view_stride + (((long long)b * p.D + d0 + k) * HW + pix) * LPV + sub, g);
                }
            }
        }
    }

    const float V = (float)(p.S + 1), V2 = V * V;
    const float rV = 1.f / V, rV2 = 1.f / V2;
#pragma unroll
    for (int k = 0; k < K1_DCH; k++) {
        if (AGG == MVSB200_AGG_GROUPCORR || !active || d0 + k >= p.D) break;
        float4 o;
        if (AGG == MVSB200_AGG_VARIANCE) {
            o.x = div_by(acc2[k].x, V, rV) - div_by(acc1[k].x * acc1[k].x, V2, rV2);
            o.y = div_by(acc2[k].y, V, rV) - div_by(acc1[k].y * acc1[k].y, V2, rV2);
            o.z = div_by(acc2[k].z, V, rV) - div_by(acc1[k].z * acc1[k].z, V2, rV2);
            o.w = div_by(acc2[k].w, V, rV) - div_by(acc1[k].w * acc1[k].w, V2, rV2);
        } else if (AGG == MVSB200_AGG_VARIANCE_MEAN) {
            const float mx = div_by(acc1[k].x, V, rV), my = div_by(acc1[k].y, V, rV), mz = div_by(acc1[k].z, V, rV),
                        mw = div_by(acc1[k].w, V, rV);
            o.x = div_by(acc2[k].x, V, rV) - mx * mx;
            o.y = div_by(acc2[k].y, V, rV) - my * my;
            o.z = div_by(acc2[k].z, V, rV) - mz * mz;
            o.w = div_by(acc2[k].w, V, rV) - mw * mw;
        } else {
            const float den = sum_exp[k] + 1e-6f, rden = 1.f / den;
            o.x = div_by(acc1[k].x, den, rden); o.y = div_by(acc1[k].y, den, rden);
            o.z = div_by(acc1[k].z, den, rden); o.w = div_by(acc1[k].w, den, rden);
        }
        vmax = fmaxf(fmaxf(vmax, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
        st4_stream(p.out + (((long long)b * p.D + d0 + k) * HW + pix) * C + sub * 4, o);
    }
    if (p.out_amax) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, m));
        if ((threadIdx.x & 31) == 0 && vmax > 0.f) atomicMax(reinterpret_cast<unsigned int *>(p.out_amax), __float_as_uint(vmax));
    }
}

template <int C, int GEOM>
static int launch_agg(const K1Params &p, int agg, dim3 grid, cudaStream_t st)
{
    switch (agg) {
    case MVSB200_AGG_VARIANCE: k1_cost_volume_kernel<C, GEOM, MVSB200_AGG_VARIANCE><<<grid, K1_THREADS, 0, st>>>(p); break;
    case MVSB200_AGG_VARIANCE_MEAN: k1_cost_volume_kernel<C, GEOM, MVSB200_AGG_VARIANCE_MEAN><<<grid, K1_THREADS, 0, st>>>(p); break;
    case MVSB200_AGG_SOFTMIN: k1_cost_volume_kernel<C, GEOM, MVSB200_AGG_SOFTMIN><<<grid, K1_THREADS, 0, st>>>(p); break;
    case MVSB200_AGG_GROUPCORR: k1_cost_volume_kernel<C, GEOM, MVSB200_AGG_GROUPCORR><<<grid, K1_THREADS, 0, st>>>(p); break;
    default: set_error("build_cost_volume: unknown aggregation %d", agg); return MVSB200_E_INVALID;
    }
    return check_launch("k1_cost_volume_kernel");
}

template <int C>
static int launch_geom(const K1Params &p, int geom, int agg, dim3 grid, cudaStream_t st)
{
    if (geom == MVSB200_GEOM_MVS) return launch_agg<C, MVSB200_GEOM_MVS>(p, agg, grid, st);
    if (geom == MVSB200_GEOM_VIS) return launch_agg<C, MVSB200_GEOM_VIS>(p, agg, grid, st);
    set_error("build_cost_volume: unknown geometry %d", geom);
    return MVSB200_E_INVALID;
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_build_cost_volume(const mvsb200_cost_volume_desc *d, const float *ref,
                                         const float *const *src, const float *warp, const float *depth,
                                         const float *interval, const float *temp, float *out, float *out_amax,
                                         mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(d && ref && src && warp && depth && out, "build_cost_volume: null pointer");
    MVSB200_REQUIRE(d->B > 0 && d->D > 0 && d->H > 0 && d->W > 0, "build_cost_volume: bad shape B=%d D=%d H=%d W=%d",
                    d->B, d->D, d->H, d->W);
    MVSB200_REQUIRE(d->S >= 1 && d->S <= MVSB200_MAX_SRC, "build_cost_volume: S=%d not in [1,%d]", d->S, MVSB200_MAX_SRC);
    MVSB200_REQUIRE(d->C == 8 || d->C == 16 || d->C == 32, "build_cost_volume: C=%d (supported: 8, 16, 32)", d->C);
    MVSB200_REQUIRE(d->depth_mode >= 0 && d->depth_mode <= 3, "build_cost_volume: depth_mode=%d", d->depth_mode);
    MVSB200_REQUIRE(d->depth_mode < MVSB200_DEPTH_START || interval, "build_cost_volume: interval is null");
    MVSB200_REQUIRE(d->agg != MVSB200_AGG_SOFTMIN || temp, "build_cost_volume: softmin needs temp");
    MVSB200_REQUIRE(d->agg != MVSB200_AGG_GROUPCORR || d->groups * 4 == d->C,
                    "build_cost_volume: group correlation needs C == 4*groups (C=%d groups=%d)", d->C, d->groups);
    MVSB200_REQUIRE(d->B <= 65535, "build_cost_volume: B too large");
    K1Params p;
    p.ref = ref;
    for (int s = 0; s < d->S; s++) {
        MVSB200_REQUIRE(src[s] && d->src_h[s] > 0 && d->src_w[s] > 0, "build_cost_volume: source %d invalid", s);
        MVSB200_REQUIRE((long long)d->src_h[s] * d->src_w[s] < (1ll << 29) && (long long)d->src_h[s] * d->src_w[s] * d->C < (1ll << 31),
                        "build_cost_volume: source %d too large (%dx%d)", s, d->src_h[s], d->src_w[s]);
        p.src[s] = src[s];
        p.src_h[s] = d->src_h[s];
        p.src_w[s] = d->src_w[s];
    }
    p.warp = warp; p.depth = depth; p.interval = interval; p.temp = temp; p.out = out; p.out_amax = out_amax;
    p.out_view_stride = d->out_view_stride;
    p.B = d->B; p.S = d->S; p.D = d->D; p.H = d->H; p.W = d->W; p.depth_mode = d->depth_mode;
    const long long HW = (long long)d->H * d->W;
    const int vpb = K1_THREADS / (d->C / 4);
    dim3 grid((unsigned)((HW + vpb - 1) / vpb), (unsigned)((d->D + K1_DCH - 1) / K1_DCH), (unsigned)d->B);
    MVSB200_REQUIRE(grid.y <= 65535, "build_cost_volume: D too large");
    cudaStream_t st = (cudaStream_t)stream;
    switch (d->C) {
    case 8: return launch_geom<8>(p, d->geom, d->agg, grid, st);
    case 16: return launch_geom<16>(p, d->geom, d->agg, grid, st);
    default: return launch_geom<32>(p, d->geom, d->agg, grid, st);
    }
}
