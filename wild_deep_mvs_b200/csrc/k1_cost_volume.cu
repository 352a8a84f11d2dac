// K1: fused homography warp + cross-view aggregation (variance / softmin / group correlation).
//
// Layout: features NHWC so that each bilinear tap is one contiguous C*4-byte vector; a group of
// LPV = C/8 lanes owns one reference pixel, each lane 8 contiguous channels moved by ONE 256-bit access (LDG.256 /
// STG.256, sm_100): a tap of a C = 32 pixel is one 128-byte line read by its four lanes, so a warp-wide load touches
// 32/LPV lines in one instruction -- the L1 data pipe (wavefronts = instructions x lines touched) is the unit this
// kernel loads most.  A warp covers 32/LPV adjacent pixels.  All per-channel arithmetic runs as packed fp32 pairs
// (FFMA2), and 8 channels per lane halve the shuffles per channel.
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
#include "k1_common.cuh"
#include <cstdlib>

namespace mvsb200 {

template <int C, int GEOM, int AGG>
__global__ void __launch_bounds__(K1_THREADS, 4) k1_cost_volume_kernel(const K1Params p)
{
    constexpr int LPV = C / 8;              // lanes per voxel, each owning 8 channels (two 16-byte vectors per tap)
    constexpr int VPB = K1_THREADS / LPV;   // pixels per block
    constexpr int KPL = K1_DCH / LPV;       // hypotheses whose geometry each lane of a pixel group computes
    static_assert(K1_DCH % LPV == 0, "depth chunk must split evenly over the lanes of a pixel group");
    __shared__ float s_warp[MVSB200_MAX_SRC * 16];

    const int b = blockIdx.z;
    const long long HW = (long long)p.H * p.W;
    for (int i = threadIdx.x; i < p.S * 16; i += K1_THREADS) s_warp[i] = p.warp[(long long)b * p.S * 16 + i];
    __syncthreads();

    // A block owns a (32/LPV) x 4 pixel tile (one warp per row: the taps of vertically adjacent pixels share cache
    // lines) and walks `chunks` depth chunks one after the other (the taps of consecutive chunks are the same or
    // neighbouring pixels, so they are served by L1).
    constexpr int TW = 32 / LPV, TH = K1_THREADS / 32;
    const int sub = threadIdx.x % LPV;
    const int tiles_x = (p.W + TW - 1) / TW;
    const int x_raw = (int)(blockIdx.x % tiles_x) * TW + (threadIdx.x % 32) / LPV;
    const int y_raw = (int)(blockIdx.x / tiles_x) * TH + threadIdx.x / 32;
    const bool active = x_raw < p.W && y_raw < p.H;        // inactive groups shadow a border pixel and store nothing,
    const int x = min(x_raw, p.W - 1), y = min(y_raw, p.H - 1);   // so every shuffle below runs with the full warp
    const long long pix = (long long)y * p.W + x;

    const float *refp = p.ref + ((long long)b * HW + pix) * C + sub * 8;
    const float interval = (p.depth_mode >= MVSB200_DEPTH_START) ? __ldg(p.interval + b) : 0.f;
    const float temp = (AGG == MVSB200_AGG_SOFTMIN) ? __ldg(p.temp) : 0.f;
    float vmax = 0.f;   // max |stored value| of this thread (abs-max tracking for the z-march conv engine)

    const int d_end = min(p.D, (int)(blockIdx.y + 1) * p.chunks * K1_DCH);
    for (int d0 = blockIdx.y * p.chunks * K1_DCH; d0 < d_end; d0 += K1_DCH) {
    const F8 r = ld8(refp);
    // The projection of a (pixel, hypothesis, view) is the same for every channel: lane `sub` of the pixel group
    // computes it for hypotheses sub, sub+LPV, ... of the chunk and the group shares the taps by shuffle.
    float dv[KPL];
#pragma unroll
    for (int j = 0; j < KPL; j++) {
        int d = min(d0 + j * LPV + sub, p.D - 1);
        dv[j] = hypothesis(p.depth_mode, p.depth, interval, b, d, p.D, HW, pix);
    }

    float2 acc1[K1_DCH][4], acc2[K1_DCH][4];
    float sum_exp[K1_DCH];
#pragma unroll
    for (int k = 0; k < K1_DCH; k++) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (AGG == MVSB200_AGG_VARIANCE || AGG == MVSB200_AGG_VARIANCE_MEAN) {
                acc1[k][q] = r.v[q];
                acc2[k][q] = __fmul2_rn(r.v[q], r.v[q]);
            } else {
                acc1[k][q] = make_float2(0.f, 0.f);
                acc2[k][q] = acc1[k][q];
            }
        }
        sum_exp[k] = 0.f;
    }
    for (int s = 0; s < p.S; s++) {
        const float *wp = s_warp + s * 16;
        const int Hs = p.src_h[s], Ws = p.src_w[s];
        // this lane's 8 channels of pixel 0 of the map; pinned so that the tap loads below do not re-derive it
        const float *mapl = p.src[s] + (long long)b * Hs * Ws * C + sub * 8;
        asm volatile("" : "+l"(mapl));
        const unsigned row_bytes = (unsigned)(Ws * C) * 4u;
        PackedTaps own[KPL];
        view_taps<GEOM, KPL>(p, s, wp, x, y, dv, own);

        // Consecutive hypotheses of a pixel move along the epipolar line by a fraction of a pixel, so they mostly
        // fall into the same 2x2 tap cell: the four 32-byte taps stay in registers until the cell changes.
        int cur_cell = 0;
        F8 ta, tb, tc, td;
        if (K1_DBG(p) & 4) ta = tb = tc = td = r;
#pragma unroll
        for (int k = 0; k < K1_DCH; k++) {
            const int owner = k % LPV, j = k / LPV;
            const int cell = __shfl_sync(0xffffffffu, own[j].cell, owner, LPV);
            const float w00 = __shfl_sync(0xffffffffu, own[j].w00, owner, LPV);
            const float w01 = __shfl_sync(0xffffffffu, own[j].w01, owner, LPV);
            const float w10 = __shfl_sync(0xffffffffu, own[j].w10, owner, LPV);
            const float w11 = __shfl_sync(0xffffffffu, own[j].w11, owner, LPV);
            if (!(K1_DBG(p) & 4) && (k == 0 || (cell != cur_cell && !(K1_DBG(p) & 2)))) {
                cur_cell = cell;
                const char *q0 = reinterpret_cast<const char *>(mapl) + (unsigned long long)(unsigned)cell * (C * 4);
                const char *q1 = q0 + row_bytes;
                ta = ld8(reinterpret_cast<const float *>(q0));
                tb = ld8(reinterpret_cast<const float *>(q0) + C);
                tc = ld8(reinterpret_cast<const float *>(q1));
                td = ld8(reinterpret_cast<const float *>(q1) + C);
            }
            const float2 p00 = make_float2(w00, w00), p01 = make_float2(w01, w01), p10 = make_float2(w10, w10), p11 = make_float2(w11, w11);
            float2 w[4];
            // accumulation order nw, ne, sw, se (ATen grid_sampler_2d)
#pragma unroll
            for (int q = 0; q < 4; q++) {
                w[q] = __fmul2_rn(ta.v[q], p00);
                w[q] = __ffma2_rn(tb.v[q], p01, w[q]);
                w[q] = __ffma2_rn(tc.v[q], p10, w[q]);
                w[q] = __ffma2_rn(td.v[q], p11, w[q]);
            }
            if (AGG == MVSB200_AGG_VARIANCE || AGG == MVSB200_AGG_VARIANCE_MEAN) {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    acc1[k][q] = __fadd2_rn(acc1[k][q], w[q]);
                    acc2[k][q] = __ffma2_rn(w[q], w[q], acc2[k][q]);
                }
            } else if (AGG == MVSB200_AGG_SOFTMIN) {
                float2 df[4];
                const float2 m1 = make_float2(-1.f, -1.f);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    df[q] = __ffma2_rn(r.v[q], m1, w[q]);   // w - r (the product is exact)
                    df[q] = __fmul2_rn(df[q], df[q]);
                }
                float ssd = ((df[0].x + df[0].y) + (df[1].x + df[1].y)) + ((df[2].x + df[2].y) + (df[3].x + df[3].y));
#pragma unroll
                for (int m = LPV / 2; m >= 1; m >>= 1) ssd += __shfl_xor_sync(0xffffffffu, ssd, m);
                const float e = expf(-temp * ssd);
                sum_exp[k] += e;
                const float2 ee = make_float2(e, e);
#pragma unroll
                for (int q = 0; q < 4; q++) acc1[k][q] = __ffma2_rn(df[q], ee, acc1[k][q]);
            } else {  // GROUPCORR: groups of 4 channels, two per lane; one output volume per source view
                float g0 = r.v[0].x * w[0].x;
                g0 += r.v[0].y * w[0].y;
                g0 += r.v[1].x * w[1].x;
                g0 += r.v[1].y * w[1].y;
                float g1 = r.v[2].x * w[2].x;
                g1 += r.v[2].y * w[2].y;
                g1 += r.v[3].x * w[3].x;
                g1 += r.v[3].y * w[3].y;
                if (active && d0 + k < p.D) {
                    vmax = fmaxf(vmax, fmaxf(fabsf(g0), fabsf(g1)));
                    __stcs(reinterpret_cast<float2 *>(p.out + s * p.out_view_stride + (((long long)b * p.D + d0 + k) * HW + pix) * (2 * LPV) + sub * 2),
                           make_float2(g0, g1));
                }
            }
        }
    }

    const float V = (float)(p.S + 1), V2 = V * V;
    const float rV = 1.f / V, rV2 = 1.f / V2;
    // x / d for a divisor whose correctly rounded reciprocal r is known (see div_by), on channel pairs
    auto div2 = [](float2 xx, float d, float rr) {
        const float2 q = __fmul2_rn(xx, make_float2(rr, rr));
        const float2 t = __ffma2_rn(q, make_float2(-d, -d), xx);
        return __ffma2_rn(t, make_float2(rr, rr), q);
    };
#pragma unroll
    for (int k = 0; k < K1_DCH; k++) {
        if (AGG == MVSB200_AGG_GROUPCORR || !active || d0 + k >= p.D) break;
        float2 o[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (AGG == MVSB200_AGG_VARIANCE) {
                const float2 m2 = div2(acc2[k][q], V, rV), m11 = div2(__fmul2_rn(acc1[k][q], acc1[k][q]), V2, rV2);
                o[q] = __ffma2_rn(m11, make_float2(-1.f, -1.f), m2);
            } else if (AGG == MVSB200_AGG_VARIANCE_MEAN) {
                const float2 m = div2(acc1[k][q], V, rV), m2 = div2(acc2[k][q], V, rV);
                o[q] = __ffma2_rn(__fmul2_rn(m, m), make_float2(-1.f, -1.f), m2);
            } else {
                const float den = sum_exp[k] + 1e-6f, rden = 1.f / den;
                o[q] = div2(acc1[k][q], den, rden);
            }
            vmax = fmaxf(vmax, fmaxf(fabsf(o[q].x), fabsf(o[q].y)));
        }
        float *dst = p.out + (((long long)b * p.D + d0 + k) * HW + pix) * C + sub * 8;
        if (!(K1_DBG(p) & 1)) st8_stream(dst, o);
    }
    }   // depth chunks
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
    K1Params p;
    p.ref = ref;
    if (int rc = k1_fill_sources(p, d, src, "build_cost_volume")) return rc;
    p.warp = warp; p.depth = depth; p.interval = interval; p.temp = temp; p.out = out; p.out_amax = out_amax;
    p.out_view_stride = d->out_view_stride;
    p.B = d->B; p.S = d->S; p.D = d->D; p.H = d->H; p.W = d->W; p.depth_mode = d->depth_mode;
    dim3 grid;
    if (int rc = k1_grid(p, d, grid, "build_cost_volume")) return rc;
#ifdef MVSB200_K1_EXPERIMENTS
    {
        const char *e = getenv("MVSB200_K1_DEBUG");
        p.dbg = e ? atoi(e) : 0;
    }
#endif
    cudaStream_t st = (cudaStream_t)stream;
    switch (d->C) {
    case 8: return launch_geom<8>(p, d->geom, d->agg, grid, st);
    case 16: return launch_geom<16>(p, d->geom, d->agg, grid, st);
    default: return launch_geom<32>(p, d->geom, d->agg, grid, st);
    }
}
