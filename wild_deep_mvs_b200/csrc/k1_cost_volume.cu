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
#include "common.cuh"
#include <cstdlib>

namespace mvsb200 {

constexpr int K1_DCH = 4;       // hypotheses per thread (x 8 channels x {M1, M2} = 64 accumulator registers)
constexpr int K1_THREADS = 256;
#ifdef MVSB200_K1_EXPERIMENTS
#define K1_DBG(p) ((p).dbg)
#else
#define K1_DBG(p) 0
#endif

struct K1Params {
    const float *ref;
    const float *src[MVSB200_MAX_SRC];
    int src_h[MVSB200_MAX_SRC], src_w[MVSB200_MAX_SRC];
    // divisor that normalises a projected coordinate and its correctly rounded reciprocal (host, IEEE):
    // MVS (Ws-1)/2, (Hs-1)/2 (module.py:151-152); VIS Ws, Hs (homography.py:93-94)
    float nx[MVSB200_MAX_SRC], rnx[MVSB200_MAX_SRC], ny[MVSB200_MAX_SRC], rny[MVSB200_MAX_SRC];
    const float *warp;
    const float *depth;
    const float *interval;
    const float *temp;
    float *out;
    float *out_amax;
    long long out_view_stride;
    int B, S, D, H, W, depth_mode;
    int chunks;   // depth chunks (of K1_DCH hypotheses) a block walks
#ifdef MVSB200_K1_EXPERIMENTS
    int dbg;      // MVSB200_K1_DEBUG experiment mask (profiles/k1_variants.py); compiled out of the product
#endif
};

// The four taps of a sample as they travel between the lanes of a pixel group: `cell` = index of the north-west pixel
// of the 2x2 block that is LOADED, and the weights of its four pixels.  The block is clamped into the map
// (xa in [0, Ws-2], ya in [0, Hs-2]) so that its pixels sit at fixed offsets (+1 pixel, +1 row): where the sample's own
// 2x2 cell straddles the border, the weights of its in-range taps move onto the loaded pixels they coincide with and
// the out-of-range taps (zero padding) drop out.  A moved weight multiplies the same feature value and the vacated slot
// contributes an exact 0, so the sum nw, ne, sw, se (ATen grid_sampler_2d order) is bit for bit the reference's.
struct PackedTaps {
    int cell;
    float w00, w01, w10, w11;
};

// Normalised grid coordinate -> packed taps (grid_sample bilinear, zeros padding, align_corners=True).
__device__ __forceinline__ PackedTaps make_taps(float gx, float gy, int Hs, int Ws)
{
    const float ix = ((gx + 1.f) / 2.f) * (float)(Ws - 1);
    const float iy = ((gy + 1.f) / 2.f) * (float)(Hs - 1);
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float x1 = fx + 1.f, y1 = fy + 1.f;
    float w00 = (x1 - ix) * (y1 - iy);
    float w01 = (ix - fx) * (y1 - iy);
    float w10 = (x1 - ix) * (iy - fy);
    float w11 = (ix - fx) * (iy - fy);
    const int xa = min(max(x0, 0), Ws - 2), ya = min(max(y0, 0), Hs - 2);
    const int sx = x0 - xa, sy = y0 - ya;   // 0 inside; -1 / +1: one column (row) of the cell is still in the map
    if (sx != 0) {
        const float l0 = (sx == -1) ? w01 : 0.f, l1 = (sx == -1) ? w11 : 0.f;
        const float r0 = (sx == 1) ? w00 : 0.f, r1 = (sx == 1) ? w10 : 0.f;
        w00 = l0; w10 = l1; w01 = r0; w11 = r1;
    }
    if (sy != 0) {
        const float u0 = (sy == -1) ? w10 : 0.f, u1 = (sy == -1) ? w11 : 0.f;
        const float d0 = (sy == 1) ? w00 : 0.f, d1 = (sy == 1) ? w01 : 0.f;
        w00 = u0; w01 = u1; w10 = d0; w11 = d1;
    }
    PackedTaps t;
    t.cell = ya * Ws + xa;
    t.w00 = w00; t.w01 = w01; t.w10 = w10; t.w11 = w11;
    return t;
}

// x / d for a divisor whose reciprocal r is known to within an ulp: one residual correction of x*r gives the correctly
// rounded quotient for normal operands when r is the correctly rounded reciprocal (the epilogue's V, V^2 and the
// coordinate normalisers, all computed once), and the IEEE quotient up to rare last-bit ties when r comes from
// rcp_nr (the fast path of the division sequence without its range check and slow-path call; a zero, denormal or
// non-finite divisor yields NaN or a huge value, which the callers' clamps turn into an out-of-map sample exactly
// as the reference's +-inf would).
__device__ __forceinline__ float div_by(float x, float d, float r)
{
    const float q = x * r;
    return fmaf(fmaf(-q, d, x), r, q);
}
__device__ __forceinline__ float rcp_nr(float d)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return fmaf(fmaf(-d, r, 1.f), r, r);
}

// Packed fp32 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2: two IEEE fp32 operations per instruction, same rounding
// as the scalar forms) over the channel pairs a lane owns.
struct F8 {   // 8 channels of one lane
    float2 v[4];
};
__device__ __forceinline__ F8 ld8(const float *p)   // 32-byte aligned, read-only path
{
    F8 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.v[0].x), "=f"(r.v[0].y), "=f"(r.v[1].x), "=f"(r.v[1].y), "=f"(r.v[2].x), "=f"(r.v[2].y), "=f"(r.v[3].x), "=f"(r.v[3].y)
        : "l"(p));
    return r;
}
// streaming 256-bit store: the written volume is consumed by a later kernel, not by this one
__device__ __forceinline__ void st8_stream(float *p, const float2 (&o)[4])
{
    asm volatile("st.global.cs.v8.f32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
                 :: "f"(o[0].x), "f"(o[0].y), "f"(o[1].x), "f"(o[1].y), "f"(o[2].x), "f"(o[2].y), "f"(o[3].x), "f"(o[3].y), "l"(p)
                 : "memory");
}

template <int C, int GEOM, int AGG>
__global__ void __launch_bounds__(K1_THREADS, 2) k1_cost_volume_kernel(const K1Params p)
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

    // A block owns a (32/LPV) x 8 pixel tile (one warp per row: the taps of vertically adjacent pixels share cache
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
        const float nx = p.nx[s], rnx = p.rnx[s], ny = p.ny[s], rny = p.rny[s];
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
                const float qx = ax * dv[j] + bx, qy = ay * dv[j] + by, qz = az * dv[j] + bz;
                const float rz = rcp_nr(qz);
                float px = div_by(qx, qz, rz), py = div_by(qy, qz, rz);
                if (qz <= 0.f) px = -10.f, py = -10.f;
                gx = clampf(div_by(px, nx, rnx) - 1.f, -10.f, 10.f);
                gy = clampf(div_by(py, ny, rny) - 1.f, -10.f, 10.f);
            } else {
                const float dd = dv[j] + 1e-9f;
                const float f = div_by(np_, dd, rcp_nr(dd));
                const float qx = ax - bx * f, qy = ay - by * f, qz = az - bz * f;
                const float zc = fmaxf(qz, 1e-9f), rz = rcp_nr(zc);
                float u = div_by(qx, zc, rz), v = div_by(qy, zc, rz);
                if (!(qz > 0.f)) u = -10.f, v = -10.f;
                gx = clampf(div_by(u, nx, rnx) * 2.f - 1.f, -1.1f, 1.1f);
                gy = clampf(div_by(v, ny, rny) * 2.f - 1.f, -1.1f, 1.1f);
            }
            // NaN coordinates (degenerate cameras) sample nothing
            if (!(gx == gx) || !(gy == gy)) gx = gy = -10.f;
            own[j] = make_taps(gx, gy, Hs, Ws);
        }

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
    MVSB200_REQUIRE(d->B <= 65535, "build_cost_volume: B too large");
    K1Params p;
    p.ref = ref;
    for (int s = 0; s < d->S; s++) {
        MVSB200_REQUIRE(src[s] && d->src_h[s] > 1 && d->src_w[s] > 1, "build_cost_volume: source %d invalid (maps are at least 2x2)", s);
        MVSB200_REQUIRE((long long)d->src_h[s] * d->src_w[s] < (1ll << 29) && (long long)d->src_h[s] * d->src_w[s] * d->C < (1ll << 31),
                        "build_cost_volume: source %d too large (%dx%d)", s, d->src_h[s], d->src_w[s]);
        p.src[s] = src[s];
        p.src_h[s] = d->src_h[s];
        p.src_w[s] = d->src_w[s];
        p.nx[s] = d->geom == MVSB200_GEOM_MVS ? (float)(d->src_w[s] - 1) / 2.f : (float)d->src_w[s];
        p.ny[s] = d->geom == MVSB200_GEOM_MVS ? (float)(d->src_h[s] - 1) / 2.f : (float)d->src_h[s];
        p.rnx[s] = 1.f / p.nx[s];
        p.rny[s] = 1.f / p.ny[s];
    }
    p.warp = warp; p.depth = depth; p.interval = interval; p.temp = temp; p.out = out; p.out_amax = out_amax;
    p.out_view_stride = d->out_view_stride;
    p.B = d->B; p.S = d->S; p.D = d->D; p.H = d->H; p.W = d->W; p.depth_mode = d->depth_mode;
    // pixel tiles of (32/LPV) x 8; a block walks as many depth chunks as still leaves >= 8 blocks per SM in the grid
    const int tw = 32 / (d->C / 8), th = K1_THREADS / 32;
    const long long tiles = (long long)((d->W + tw - 1) / tw) * ((d->H + th - 1) / th);
    const int nchunk = (d->D + K1_DCH - 1) / K1_DCH;
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    int chunks = 1;
    while (chunks < 8 && chunks * 2 <= nchunk && tiles * d->B * ((nchunk + chunks * 2 - 1) / (chunks * 2)) >= 8ll * sms) chunks *= 2;
    p.chunks = chunks;
#ifdef MVSB200_K1_EXPERIMENTS
    {
        const char *e = getenv("MVSB200_K1_DEBUG");
        p.dbg = e ? atoi(e) : 0;
    }
#endif
    MVSB200_REQUIRE(tiles < (1ll << 31), "build_cost_volume: image too large");
    dim3 grid((unsigned)tiles, (unsigned)((nchunk + chunks - 1) / chunks), (unsigned)d->B);
    cudaStream_t st = (cudaStream_t)stream;
    switch (d->C) {
    case 8: return launch_geom<8>(p, d->geom, d->agg, grid, st);
    case 16: return launch_geom<16>(p, d->geom, d->agg, grid, st);
    default: return launch_geom<32>(p, d->geom, d->agg, grid, st);
    }
}
