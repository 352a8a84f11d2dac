// K1 backward (row f2 of SURVEY.md 8): gradient of the fused warp + aggregation with respect to the FEATURE maps.
//
// The reference builds the sampling grid under torch.no_grad() (models/MVSNet/module.py:127, VisMVSNet/homography.py:25,
// 110), so no gradient reaches the cameras or the depth hypotheses: the backward of `build_cost_volume` is
//   dL/dsrc_s[tap]  += bilinear weight * dL/dwarped_s      (grid_sampler_2d_backward: a 4-tap scatter-add)
//   dL/dref          = sum over depth of the aggregation's own derivative
// plus dL/dtemp for the soft-min aggregation (models/MVSNet/model.py:94-95,141-173).
//
// Same decomposition as the forward kernel (k1_cost_volume.cu): a lane group owns a pixel, a lane 8 channels, a thread
// walks 4 hypotheses; the warped features are RECOMPUTED from the taps (never stored by the forward), in two sweeps over
// the source views where the aggregation needs a cross-view quantity first (variance: M1; soft-min: the normaliser and
// the output).  The scatter uses 128-bit vector reductions (red.global.add.v4.f32), two per tap and lane, skipping taps
// whose weight is zero (zero padding, out-of-map samples).
//
//   variance      out = M2/V - (M1/V)^2               dL/dw_s = (2 g / V) (w_s - M1/V);  dL/dr likewise
//   soft-min      out = sum_s e_s d_s / (sum_s e_s + 1e-6), d_s = (r - w_s)^2, e_s = exp(-temp sum_c d_s)
//   group corr.   out_s[g] = sum_{c in g} r_c w_s,c    dL/dw_s,c = G_s[g] r_c;  dL/dr_c = sum_s,d G_s[g] w_s,c
#include "k1_common.cuh"
#include <cstdint>

namespace mvsb200 {

struct K1BwdParams {
    const float *gout;                 // gradient of the forward output, same layout
    float *gref;                       // [B,H,W,C], zeroed by the caller (blocks add their depth range)
    float *gsrc[MVSB200_MAX_SRC];      // [B,Hs,Ws,C] each, zeroed by the caller
    float *gtemp;                      // soft-min: one float, zeroed by the caller
};

__device__ __forceinline__ void red4(float *p, float2 a, float2 b)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y) : "memory");
}

// weight * gradient of one warped sample into the four pixels of its tap block (this lane's 8 channels)
__device__ __forceinline__ void scatter_taps(float *g00, unsigned row_floats, int C, float w00, float w01, float w10, float w11,
                                             const float2 (&gw)[4])
{
    auto one = [&](float *dst, float wt) {
        if (wt != 0.f) {
            const float2 ww = make_float2(wt, wt);
            red4(dst, __fmul2_rn(gw[0], ww), __fmul2_rn(gw[1], ww));
            red4(dst + 4, __fmul2_rn(gw[2], ww), __fmul2_rn(gw[3], ww));
        }
    };
    one(g00, w00);
    one(g00 + C, w01);
    one(g00 + row_floats, w10);
    one(g00 + row_floats + C, w11);
}

// One sweep over the hypotheses of the chunk in one source view: shuffles the taps of hypothesis k to the lanes of the
// pixel, keeps the 2x2 tap block in registers while the cell does not change, interpolates this lane's 8 channels in
// the reference's order (nw, ne, sw, se) and hands them to f(k, cell, w00, w01, w10, w11, warped).
template <int C, int LPV, int KPL, typename F>
__device__ __forceinline__ void sweep_view(const float *mapl, unsigned row_bytes, const PackedTaps (&own)[KPL], F &f)
{
    int cur_cell = 0;
    F8 ta, tb, tc, td;
#pragma unroll
    for (int k = 0; k < K1_DCH; k++) {
        const int owner = k % LPV, j = k / LPV;
        const int cell = __shfl_sync(0xffffffffu, own[j].cell, owner, LPV);
        const float w00 = __shfl_sync(0xffffffffu, own[j].w00, owner, LPV);
        const float w01 = __shfl_sync(0xffffffffu, own[j].w01, owner, LPV);
        const float w10 = __shfl_sync(0xffffffffu, own[j].w10, owner, LPV);
        const float w11 = __shfl_sync(0xffffffffu, own[j].w11, owner, LPV);
        if (k == 0 || cell != cur_cell) {
            cur_cell = cell;
            const char *q0 = reinterpret_cast<const char *>(mapl) + (unsigned long long)(unsigned)cell * (C * 4);
            const char *q1 = q0 + row_bytes;
            ta = ld8(reinterpret_cast<const float *>(q0));
            tb = ld8(reinterpret_cast<const float *>(q0) + C);
            tc = ld8(reinterpret_cast<const float *>(q1));
            td = ld8(reinterpret_cast<const float *>(q1) + C);
        }
        const float2 p00 = make_float2(w00, w00), p01 = make_float2(w01, w01), p10 = make_float2(w10, w10), p11 = make_float2(w11, w11);
        F8 w;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            w.v[q] = __fmul2_rn(ta.v[q], p00);
            w.v[q] = __ffma2_rn(tb.v[q], p01, w.v[q]);
            w.v[q] = __ffma2_rn(tc.v[q], p10, w.v[q]);
            w.v[q] = __ffma2_rn(td.v[q], p11, w.v[q]);
        }
        f(k, cell, w00, w01, w10, w11, w);
    }
}

template <int C, int GEOM, int AGG>
__global__ void __launch_bounds__(K1_THREADS, 4) k1_backward_kernel(const K1Params p, const K1BwdParams bp)
{
    constexpr int LPV = C / 8;
    constexpr int KPL = K1_DCH / LPV;
    constexpr bool VAR = AGG == MVSB200_AGG_VARIANCE || AGG == MVSB200_AGG_VARIANCE_MEAN;
    __shared__ float s_warp[MVSB200_MAX_SRC * 16];

    const int b = blockIdx.z;
    const long long HW = (long long)p.H * p.W;
    for (int i = threadIdx.x; i < p.S * 16; i += K1_THREADS) s_warp[i] = p.warp[(long long)b * p.S * 16 + i];
    __syncthreads();

    constexpr int TW = 32 / LPV, TH = K1_THREADS / 32;
    const int sub = threadIdx.x % LPV;
    const int tiles_x = (p.W + TW - 1) / TW;
    const int x_raw = (int)(blockIdx.x % tiles_x) * TW + (threadIdx.x % 32) / LPV;
    const int y_raw = (int)(blockIdx.x / tiles_x) * TH + threadIdx.x / 32;
    const bool active = x_raw < p.W && y_raw < p.H;
    const int x = min(x_raw, p.W - 1), y = min(y_raw, p.H - 1);
    const long long pix = (long long)y * p.W + x;

    const F8 r = ld8(p.ref + ((long long)b * HW + pix) * C + sub * 8);
    const float interval = (p.depth_mode >= MVSB200_DEPTH_START) ? __ldg(p.interval + b) : 0.f;
    const float temp = (AGG == MVSB200_AGG_SOFTMIN) ? __ldg(p.temp) : 0.f;
    const float V = (float)(p.S + 1);
    float2 gr[4];
#pragma unroll
    for (int q = 0; q < 4; q++) gr[q] = make_float2(0.f, 0.f);
    float gtemp = 0.f;

    auto map_of = [&](int s) { return p.src[s] + (long long)b * p.src_h[s] * p.src_w[s] * C + sub * 8; };
    auto gsrc_at = [&](int s, int cell) { return bp.gsrc[s] + ((long long)b * p.src_h[s] * p.src_w[s] + cell) * C + sub * 8; };
    auto group_sum = [&](float v) {   // sum over the lanes of the pixel group (all 8*LPV channels)
#pragma unroll
        for (int m = LPV / 2; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        return v;
    };
    auto sum8 = [](const float2 (&a)[4]) { return ((a[0].x + a[0].y) + (a[1].x + a[1].y)) + ((a[2].x + a[2].y) + (a[3].x + a[3].y)); };

    const int d_end = min(p.D, (int)(blockIdx.y + 1) * p.chunks * K1_DCH);
    for (int d0 = blockIdx.y * p.chunks * K1_DCH; d0 < d_end; d0 += K1_DCH) {
        float dv[KPL];
#pragma unroll
        for (int j = 0; j < KPL; j++) {
            int d = min(d0 + j * LPV + sub, p.D - 1);
            dv[j] = hypothesis(p.depth_mode, p.depth, interval, b, d, p.D, HW, pix);
        }
        bool valid[K1_DCH];
#pragma unroll
        for (int k = 0; k < K1_DCH; k++) valid[k] = active && d0 + k < p.D;

        if (AGG == MVSB200_AGG_GROUPCORR) {
            for (int s = 0; s < p.S; s++) {
                PackedTaps own[KPL];
                view_taps<GEOM, KPL>(p, s, s_warp + s * 16, x, y, dv, own);
                const unsigned row_floats = (unsigned)(p.src_w[s] * C);
                auto fn0 = [&](int k, int cell, float w00, float w01, float w10, float w11, const F8 &ww) {
                    const float2 (&w)[4] = ww.v;
                    if (!valid[k]) return;
                    const float2 G = __ldg(reinterpret_cast<const float2 *>(
                        bp.gout + s * p.out_view_stride + (((long long)b * p.D + d0 + k) * HW + pix) * (2 * LPV) + sub * 2));
                    const float2 g0 = make_float2(G.x, G.x), g1 = make_float2(G.y, G.y);
                    float2 gw[4] = {__fmul2_rn(r.v[0], g0), __fmul2_rn(r.v[1], g0), __fmul2_rn(r.v[2], g1), __fmul2_rn(r.v[3], g1)};
                    gr[0] = __ffma2_rn(w[0], g0, gr[0]);
                    gr[1] = __ffma2_rn(w[1], g0, gr[1]);
                    gr[2] = __ffma2_rn(w[2], g1, gr[2]);
                    gr[3] = __ffma2_rn(w[3], g1, gr[3]);
                    scatter_taps(gsrc_at(s, cell), row_floats, C, w00, w01, w10, w11, gw);
                };
                sweep_view<C, LPV, KPL>(map_of(s), (unsigned)(p.src_w[s] * C) * 4u, own, fn0);
            }
            continue;
        }

        // ---- first sweep: the cross-view quantity ------------------------------------------------------------------
        float2 acc[K1_DCH][4];      // variance: M1;  soft-min: sum_s e_s d_s
        float sum_exp[K1_DCH];
#pragma unroll
        for (int k = 0; k < K1_DCH; k++) {
#pragma unroll
            for (int q = 0; q < 4; q++) acc[k][q] = VAR ? r.v[q] : make_float2(0.f, 0.f);
            sum_exp[k] = 0.f;
        }
        for (int s = 0; s < p.S; s++) {
            PackedTaps own[KPL];
            view_taps<GEOM, KPL>(p, s, s_warp + s * 16, x, y, dv, own);
            auto fn1 = [&](int k, int, float, float, float, float, const F8 &ww) {
                const float2 (&w)[4] = ww.v;
                if (VAR) {
#pragma unroll
                    for (int q = 0; q < 4; q++) acc[k][q] = __fadd2_rn(acc[k][q], w[q]);
                } else {
                    float2 df[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        df[q] = __ffma2_rn(r.v[q], make_float2(-1.f, -1.f), w[q]);
                        df[q] = __fmul2_rn(df[q], df[q]);
                    }
                    const float e = expf(-temp * group_sum(sum8(df)));
                    sum_exp[k] += e;
#pragma unroll
                    for (int q = 0; q < 4; q++) acc[k][q] = __ffma2_rn(df[q], make_float2(e, e), acc[k][q]);
                }
            };
            sweep_view<C, LPV, KPL>(map_of(s), (unsigned)(p.src_w[s] * C) * 4u, own, fn1);
        }

        // ---- coefficients of the second sweep ----------------------------------------------------------------------
        // variance:  dL/dw = c0 w - c1  with c0 = 2 g / V, c1 = c0 M1 / V       (kept in g[], acc[])
        // soft-min:  gz = g / Z (kept in g[]), az = sum_c gz out_c              (Z = sum_exp + 1e-6)
        float2 g[K1_DCH][4];
        float az[K1_DCH];
#pragma unroll
        for (int k = 0; k < K1_DCH; k++) {
            F8 go;
#pragma unroll
            for (int q = 0; q < 4; q++) go.v[q] = make_float2(0.f, 0.f);
            if (valid[k]) go = ld8(bp.gout + (((long long)b * p.D + d0 + k) * HW + pix) * C + sub * 8);
            if (VAR) {
                const float2 c = make_float2(2.f / V, 2.f / V), iv = make_float2(1.f / V, 1.f / V);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    g[k][q] = __fmul2_rn(go.v[q], c);
                    acc[k][q] = __fmul2_rn(g[k][q], __fmul2_rn(acc[k][q], iv));
                    gr[q] = __fadd2_rn(gr[q], __ffma2_rn(g[k][q], r.v[q], __fmul2_rn(acc[k][q], make_float2(-1.f, -1.f))));
                }
                az[k] = 0.f;
            } else {
                const float rz = 1.f / (sum_exp[k] + 1e-6f);
                float2 t[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    g[k][q] = __fmul2_rn(go.v[q], make_float2(rz, rz));
                    t[q] = __fmul2_rn(g[k][q], __fmul2_rn(acc[k][q], make_float2(rz, rz)));   // gz * out
                }
                az[k] = group_sum(sum8(t));
            }
        }

        // ---- second sweep: per-view gradients, scattered into the source maps ----------------------------------------
        for (int s = 0; s < p.S; s++) {
            PackedTaps own[KPL];
            view_taps<GEOM, KPL>(p, s, s_warp + s * 16, x, y, dv, own);
            const unsigned row_floats = (unsigned)(p.src_w[s] * C);
            auto fn2 = [&](int k, int cell, float w00, float w01, float w10, float w11, const F8 &ww) {
                    const float2 (&w)[4] = ww.v;
                float2 gw[4];
                if (VAR) {
#pragma unroll
                    for (int q = 0; q < 4; q++) gw[q] = __ffma2_rn(g[k][q], w[q], __fmul2_rn(acc[k][q], make_float2(-1.f, -1.f)));
                } else {
                    float2 df[4], d2[4], t[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        df[q] = __ffma2_rn(r.v[q], make_float2(-1.f, -1.f), w[q]);   // w - r
                        d2[q] = __fmul2_rn(df[q], df[q]);
                        t[q] = __fmul2_rn(g[k][q], d2[q]);
                    }
                    const float ssd = group_sum(sum8(d2));
                    const float e = expf(-temp * ssd);
                    const float de = group_sum(sum8(t)) - az[k];            // dL/de_s
                    if (sub == 0) gtemp = fmaf(de, -ssd * e, gtemp);
                    const float k0 = 2.f * e, k1 = -2.f * temp * e * de;   // dL/dd_s,c = e gz_c - temp e de; times 2 (w - r)
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const float2 gd = __ffma2_rn(g[k][q], make_float2(k0, k0), make_float2(k1, k1));
                        gw[q] = __fmul2_rn(df[q], gd);
                        gr[q] = __ffma2_rn(gw[q], make_float2(-1.f, -1.f), gr[q]);
                    }
                }
                if (valid[k]) scatter_taps(gsrc_at(s, cell), row_floats, C, w00, w01, w10, w11, gw);
            };
            sweep_view<C, LPV, KPL>(map_of(s), (unsigned)(p.src_w[s] * C) * 4u, own, fn2);
        }
    }

    if (active) {
        float *dst = bp.gref + ((long long)b * HW + pix) * C + sub * 8;
        red4(dst, gr[0], gr[1]);
        red4(dst + 4, gr[2], gr[3]);
    }
    if (AGG == MVSB200_AGG_SOFTMIN) {
        if (!active) gtemp = 0.f;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) gtemp += __shfl_xor_sync(0xffffffffu, gtemp, m);
        if ((threadIdx.x & 31) == 0 && gtemp != 0.f) atomicAdd(bp.gtemp, gtemp);
    }
}

template <int C, int GEOM>
static int launch_bwd_agg(const K1Params &p, const K1BwdParams &bp, int agg, dim3 grid, cudaStream_t st)
{
    switch (agg) {
    case MVSB200_AGG_VARIANCE:
    case MVSB200_AGG_VARIANCE_MEAN: k1_backward_kernel<C, GEOM, MVSB200_AGG_VARIANCE><<<grid, K1_THREADS, 0, st>>>(p, bp); break;
    case MVSB200_AGG_SOFTMIN: k1_backward_kernel<C, GEOM, MVSB200_AGG_SOFTMIN><<<grid, K1_THREADS, 0, st>>>(p, bp); break;
    case MVSB200_AGG_GROUPCORR: k1_backward_kernel<C, GEOM, MVSB200_AGG_GROUPCORR><<<grid, K1_THREADS, 0, st>>>(p, bp); break;
    default: set_error("build_cost_volume_backward: unknown aggregation %d", agg); return MVSB200_E_INVALID;
    }
    return check_launch("k1_backward_kernel");
}

template <int C>
static int launch_bwd_geom(const K1Params &p, const K1BwdParams &bp, int geom, int agg, dim3 grid, cudaStream_t st)
{
    if (geom == MVSB200_GEOM_MVS) return launch_bwd_agg<C, MVSB200_GEOM_MVS>(p, bp, agg, grid, st);
    if (geom == MVSB200_GEOM_VIS) return launch_bwd_agg<C, MVSB200_GEOM_VIS>(p, bp, agg, grid, st);
    set_error("build_cost_volume_backward: unknown geometry %d", geom);
    return MVSB200_E_INVALID;
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_build_cost_volume_backward(const mvsb200_cost_volume_desc *d, const float *ref, const float *const *src,
                                                  const float *warp, const float *depth, const float *interval,
                                                  const float *temp, const float *grad_out, float *grad_ref,
                                                  float *const *grad_src, float *grad_temp, mvsb200_stream_t stream)
{
    const char *what = "build_cost_volume_backward";
    MVSB200_REQUIRE(d && ref && src && warp && depth && grad_out && grad_ref && grad_src, "%s: null pointer", what);
    // 256-bit loads of the maps and of grad_out, 128-bit vector reductions into the gradient maps
    MVSB200_REQUIRE(((reinterpret_cast<uintptr_t>(ref) | reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(grad_ref)) & 31) == 0,
                    "%s: ref, grad_out and grad_ref must be 32-byte aligned", what);
    for (int s = 0; s < d->S && s < MVSB200_MAX_SRC; s++)
        MVSB200_REQUIRE(src[s] && grad_src[s] && ((reinterpret_cast<uintptr_t>(src[s]) | reinterpret_cast<uintptr_t>(grad_src[s])) & 31) == 0,
                        "%s: src[%d] and grad_src[%d] must be non-null and 32-byte aligned", what, s, s);
    MVSB200_REQUIRE(d->B > 0 && d->D > 0 && d->H > 0 && d->W > 0, "%s: bad shape B=%d D=%d H=%d W=%d", what, d->B, d->D, d->H, d->W);
    MVSB200_REQUIRE(d->S >= 1 && d->S <= MVSB200_MAX_SRC, "%s: S=%d not in [1,%d]", what, d->S, MVSB200_MAX_SRC);
    MVSB200_REQUIRE(d->C == 8 || d->C == 16 || d->C == 32, "%s: C=%d (supported: 8, 16, 32)", what, d->C);
    MVSB200_REQUIRE(d->depth_mode >= 0 && d->depth_mode <= 3, "%s: depth_mode=%d", what, d->depth_mode);
    MVSB200_REQUIRE(d->depth_mode < MVSB200_DEPTH_START || interval, "%s: interval is null", what);
    MVSB200_REQUIRE(d->agg != MVSB200_AGG_SOFTMIN || (temp && grad_temp), "%s: softmin needs temp and grad_temp", what);
    MVSB200_REQUIRE(d->agg != MVSB200_AGG_GROUPCORR || d->groups * 4 == d->C, "%s: group correlation needs C == 4*groups (C=%d groups=%d)",
                    what, d->C, d->groups);
    K1Params p;
    K1BwdParams bp;
    p.ref = ref;
    if (int rc = k1_fill_sources(p, d, src, what)) return rc;
    for (int s = 0; s < d->S; s++) {
        MVSB200_REQUIRE(grad_src[s], "%s: grad_src[%d] is null", what, s);
        bp.gsrc[s] = grad_src[s];
    }
    p.warp = warp; p.depth = depth; p.interval = interval; p.temp = temp; p.out = nullptr; p.out_amax = nullptr;
    p.out_view_stride = d->out_view_stride;
    p.B = d->B; p.S = d->S; p.D = d->D; p.H = d->H; p.W = d->W; p.depth_mode = d->depth_mode;
#ifdef MVSB200_K1_EXPERIMENTS
    p.dbg = 0;
#endif
    bp.gout = grad_out; bp.gref = grad_ref; bp.gtemp = grad_temp;
    dim3 grid;
    if (int rc = k1_grid(p, d, grid, what)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (d->C) {
    case 8: return launch_bwd_geom<8>(p, bp, d->geom, d->agg, grid, st);
    case 16: return launch_bwd_geom<16>(p, bp, d->geom, d->agg, grid, st);
    default: return launch_bwd_geom<32>(p, bp, d->geom, d->agg, grid, st);
    }
}
