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
#include <cstdint>
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


// ------------------------------------------------------------------------------------------------------------------
// K1-M, the depth-marching kernel (1, 2 or 4 source views, C = 16 / 32; chosen for long sweeps over scalar hypotheses, see
// mvsb200_build_cost_volume at the end of this file: the pixel-tile kernel above is the faster one for short sweeps).
//
// What bounds this path is the traffic from L1 into the register file (128 B / clk / SM): every (pixel, hypothesis, view)
// needs four C-wide taps, 512 bytes for C = 32 (profiles/k1_r2_hypothesis_major.md).  Consecutive hypotheses of a pixel move
// a fraction of a source pixel along the epipolar line, so the kernel above keeps the taps of a view in registers while the
// 2x2 cell is unchanged -- but it walks the depth axis in chunks of 4 hypotheses with the views outside, and the forced
// reload at the start of every (chunk, view) is the larger half of its loads.  Here a pixel is owned by L = C/4 lanes, each
// holding FOUR channels of the 2x2 windows of ALL S views (S x 16 registers), and marches along a long depth segment with
// the views inside: a window is re-loaded only when its cell really changes.  One hypothesis is accumulated at a time
// (8 accumulator registers), finished and stored, so nothing else is live.
// The channel-independent tap geometry of the 4 x S (hypothesis, view) pairs of a mini-chunk is computed once, spread
// over the L lanes of the pixel -- a lane always serves the same view, whose constants it keeps in registers, with the
// per-pixel part of the projection hoisted out of the depth loop -- and handed to the other lanes through shared memory
// (one LDS.128 + one LDS.32 per (hypothesis, view): 5 wavefronts of the LSU pipe, this kernel's busiest unit after the
// change; broadcast 128-bit shared loads cost 4 wavefronts each, so packed-pair duplicates are made with moves instead).
// The pixels that share a warp are stacked ACROSS the epipolar direction so that they change cells together, and the
// window re-load is a real warp-uniform branch (the compiler otherwise if-converts it into ~56 always-issued instructions
// per hypothesis, see K1M_REAL_BRANCH_*).
// ------------------------------------------------------------------------------------------------------------------
constexpr int K1M_THREADS = 128;
// Experiment switches (profiles/k1_ab.py builds the variants with -D; the defaults are what measured best on B200):
#ifndef K1M_MIN_BLOCKS
#define K1M_MIN_BLOCKS 4 // blocks per SM the register allocation aims at: 5 (96 registers, spilled view constants) 0.262 -> 0.300 ms
#endif
#ifndef K1M_GROUP
#define K1M_GROUP 1      // views whose re-load checks precede their arithmetic together (inline mode): 2 / 4 -> 0.267 / 0.287 ms
#endif
// Window re-loads one hypothesis ahead of the arithmetic (see the kernel): group correlation 0.112 -> 0.103 ms, variance-mean
// with per-pixel hypotheses 0.487 -> 0.462 ms, but cfg2's variance 0.262 -> 0.269 ms (its arithmetic block is the longest and
// already covers the load latency): on for the first two, off otherwise.
#ifndef K1M_AHEAD
#define K1M_AHEAD(AGG) ((AGG) == MVSB200_AGG_GROUPCORR || (AGG) == MVSB200_AGG_VARIANCE_MEAN)
#endif
// ptxas if-converts a short conditional block into predicated instructions that are issued whether or not any lane needs
// them; a block that ends in a (never taken) backward branch on a launch parameter the compiler knows nothing about stays a
// branch: `if (warp-uniform condition) K1M_REAL_BRANCH_BEGIN { ... } K1M_REAL_BRANCH_END(positive_kernel_argument);`
#define K1M_REAL_BRANCH_BEGIN do
#define K1M_REAL_BRANCH_END(positive) while ((positive) < 0)
constexpr int K1M_HC = 4;        // hypotheses per mini-chunk (geometry is shared per mini-chunk)

// shared-memory record accesses by 32-bit shared address (+ compile-time offset folded into the instruction)
__device__ __forceinline__ float4 lds128(unsigned a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds32(unsigned a)
{
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(unsigned a, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(unsigned a, int v) { asm volatile("st.shared.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }

struct K1MView {                 // per source view, staged in shared memory once per block
    const float *base;           // first pixel of the view's map for batch item b
    unsigned row_bytes;
    int Hs, Ws;
    float nx, rnx, ny, rny;
    float wp[16];
};

template <int C, int GEOM, int AGG, int S>
__global__ void __launch_bounds__(K1M_THREADS, K1M_MIN_BLOCKS) k1m_cost_volume_kernel(const K1Params p, const int seg)
{
    constexpr int L = C / 4;                // lanes per pixel, each owning 4 channels (one 16-byte vector per tap)
    constexpr int PXW = 32 / L;             // adjacent pixels per warp (along x or along y, see `stack_y`)
    constexpr int R = L / S;                // hypotheses whose geometry one round of the L lanes covers
    constexpr int ROUNDS = (K1M_HC + R - 1) / R;
    constexpr int TH = K1M_THREADS / 32;
    constexpr int NREC = K1M_HC * S;        // (hypothesis, view) tap records of a pixel per mini-chunk
    static_assert(L % S == 0 && R >= 1 && (K1M_HC % R == 0 || R > K1M_HC), "views must divide the lanes of a pixel");
    __shared__ K1MView s_view[S];
    // tap records {w00, w01, w10, w11} + cell.  A broadcast LDS.128 costs the LSU pipe four wavefronts however few distinct
    // addresses it has (one per quarter warp) and that pipe, shared with the tap loads, is this kernel's busiest unit: the
    // record is therefore ONE 16-byte vector + one 4-byte word (5 wavefronts), the packed-pair weights are made with moves
    // per pixel: NREC weight vectors, then NREC cells, padded to an odd number of 16-byte vectors (conflict-free across the pixels of a warp) -- ONE shared
    // base address per lane (kept in a register) + immediate offsets
    constexpr int REC_VECS = ((NREC * 20 + 15) / 16) | 1;   // odd number of 16-byte vectors: no bank conflicts between pixels
    constexpr int REC_BYTES = REC_VECS * 16;
    constexpr int REC_CELL = NREC * 16;     // byte offset of the cells inside a pixel's record block
    __shared__ __align__(16) unsigned char s_rec[TH][PXW][REC_BYTES];

    const int b = blockIdx.z;
    const long long HW = (long long)p.H * p.W;
    if (threadIdx.x < S) {
        const int s = threadIdx.x;
        K1MView &v = s_view[s];
        v.Hs = p.src_h[s]; v.Ws = p.src_w[s];
        v.base = p.src[s] + (long long)b * v.Hs * v.Ws * C;
        v.row_bytes = (unsigned)(v.Ws * C) * 4u;
        v.nx = p.nx[s]; v.rnx = p.rnx[s]; v.ny = p.ny[s]; v.rny = p.rny[s];
#pragma unroll
        for (int i = 0; i < 16; i++) v.wp[i] = p.warp[((long long)b * S + s) * 16 + i];
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int sub = lane % L, pxi = lane / L;
    const int sv = sub % S, h_off = sub / S;         // this lane's geometry duty: view sv, hypothesis h_off of every round
    // Which PXW pixels share a warp?  A warp re-loads a window whenever ANY of its pixels changes cell, so its pixels should
    // change cells TOGETHER: neighbours ACROSS the epipolar direction see (almost) the same sub-pixel motion per hypothesis,
    // neighbours ALONG it cross cell borders at different hypotheses.  d(u,v)/d(depth) of a projected point is parallel to
    // (ax bz - az bx, ay bz - az by) whatever the depth; it is evaluated at the image centre, summed over the views (one
    // decision per launch, the same in every block): mostly horizontal motion -> the pixels of a warp are stacked along y.
    float ex = 0.f, ey = 0.f;
    {
        const float cx = (float)(p.W / 2) + (GEOM == MVSB200_GEOM_VIS ? 0.5f : 0.f), cy = (float)(p.H / 2) + (GEOM == MVSB200_GEOM_VIS ? 0.5f : 0.f);
#pragma unroll
        for (int s = 0; s < S; s++) {
            const float *wp = s_view[s].wp;
            const float cax = wp[0] * cx + wp[1] * cy + wp[2], cay = wp[3] * cx + wp[4] * cy + wp[5], caz = wp[6] * cx + wp[7] * cy + wp[8];
            ex += fabsf(cax * wp[11] - caz * wp[9]);
            ey += fabsf(cay * wp[11] - caz * wp[10]);
        }
    }
    const bool stack_y = ex >= ey;
    const int tw = stack_y ? TH : PXW, th = stack_y ? PXW : TH;      // tile = tw x th pixels
    const int tiles_x = (p.W + tw - 1) / tw;
    if ((long long)blockIdx.x >= (long long)tiles_x * ((p.H + th - 1) / th)) return;   // the grid covers the larger of the two tilings
    const int x_raw = (int)(blockIdx.x % tiles_x) * tw + (stack_y ? wrp : pxi);
    const int y_raw = (int)(blockIdx.x / tiles_x) * th + (stack_y ? pxi : wrp);
    const bool active = x_raw < p.W && y_raw < p.H;          // inactive lanes shadow a border pixel and store nothing
    const int x = min(x_raw, p.W - 1), y = min(y_raw, p.H - 1);
    const long long pix = (long long)y * p.W + x;
    const int d_begin = (int)blockIdx.y * seg, d_end = min(p.D, d_begin + seg);

    // ---- this lane's view: constants in registers, the per-pixel part of the projection hoisted ----
    const K1MView &mv = s_view[sv];
    const int Hs = mv.Hs, Ws = mv.Ws;
    const float Wm1f = (float)(Ws - 1), Hm1f = (float)(Hs - 1);
    const float nx = mv.nx, rnx = mv.rnx, ny = mv.ny, rny = mv.rny;
    const float fx = (GEOM == MVSB200_GEOM_MVS) ? (float)x : (float)x + 0.5f, fy = (GEOM == MVSB200_GEOM_MVS) ? (float)y : (float)y + 0.5f;
    const float ax = mv.wp[0] * fx + mv.wp[1] * fy + mv.wp[2];
    const float ay = mv.wp[3] * fx + mv.wp[4] * fy + mv.wp[5];
    const float az = mv.wp[6] * fx + mv.wp[7] * fy + mv.wp[8];
    const float bx = mv.wp[9], by = mv.wp[10], bz = mv.wp[11];
    const float np_ = (GEOM == MVSB200_GEOM_VIS) ? mv.wp[12] * fx + mv.wp[13] * fy + mv.wp[14] : 0.f;

    const float interval = (p.depth_mode >= MVSB200_DEPTH_START) ? __ldg(p.interval + b) : 0.f;
    const float temp = (AGG == MVSB200_AGG_SOFTMIN) ? __ldg(p.temp) : 0.f;
    const float V = (float)(S + 1);
    const float rV = 1.f / V, rV2 = 1.f / (V * V);
    const float4 r4 = ldg4(p.ref + ((long long)b * HW + pix) * C + sub * 4);
    const float2 r0 = make_float2(r4.x, r4.y), r1 = make_float2(r4.z, r4.w);
    float vmax = 0.f;   // max |stored value| of this thread (abs-max tracking for the z-march conv engine)

    // the 2x2 windows of the S views (this lane's 4 channels) and the cells they hold
    float4 ta[S], tb[S], tc[S], td[S];
    int cur[S];
#pragma unroll
    for (int s = 0; s < S; s++) cur[s] = -2;

    // block-uniform view bases (batch item b) and row pitches in 16-byte vectors: launch parameters -> uniform registers
    const float4 *vbase[S];
    unsigned vrow[S];
#pragma unroll
    for (int s = 0; s < S; s++) {
        vbase[s] = reinterpret_cast<const float4 *>(p.src[s] + (long long)b * p.src_h[s] * p.src_w[s] * C);
        vrow[s] = (unsigned)p.src_w[s] * L;
    }
    // this lane's 16-byte vector of the volume, advanced one plane per hypothesis
    float4 *outp = reinterpret_cast<float4 *>(p.out + (((long long)b * p.D + d_begin) * HW + pix) * C + sub * 4);
    const long long plane_vecs = HW * L;
    float *outg = p.out + (((long long)b * p.D + d_begin) * HW + pix) * L + sub;   // group correlation: one float per lane, plane and view

    unsigned rec = (unsigned)__cvta_generic_to_shared(&s_rec[wrp][pxi][0]);
    unsigned rec_st = rec + (h_off * S + sv) * 16, rec_st_cell = rec + REC_CELL + (h_off * S + sv) * 4;   // the records this lane writes
    asm volatile("" : "+r"(rec), "+r"(rec_st), "+r"(rec_st_cell));   // pinned: the compiler otherwise re-derives the addresses before every access
    // the four depth modes of common.cuh hypothesis() as ONE expression: value(d) = dbase[d * dstride] + interval * d
    // (a mode without an interval adds 0 * d, one with a start value reads the same word for every d)
    const float *dbase;
    long long dstride = 0;
    switch (p.depth_mode) {
    case MVSB200_DEPTH_VALUES: dbase = p.depth + (long long)b * p.D; dstride = 1; break;
    case MVSB200_DEPTH_VOLUME: dbase = p.depth + (long long)b * p.D * HW + pix; dstride = HW; break;
    case MVSB200_DEPTH_START: dbase = p.depth + b; break;
    default: dbase = p.depth + (long long)b * HW + pix; break;
    }
    for (int k0 = d_begin; k0 < d_end; k0 += K1M_HC) {
        __syncwarp();   // the previous mini-chunk's records have been read
        // ---- geometry of the mini-chunk's (hypothesis, view) pairs, one per lane and round ----
#pragma unroll
        for (int rd = 0; rd < ROUNDS; rd++) {
            const int k = rd * R + h_off;
            if (R > K1M_HC && k >= K1M_HC) break;        // more lanes than (hypothesis, view) pairs: the rest sit the round out
            const int d = min(k0 + k, p.D - 1);
            const float dv = __ldg(dbase + (long long)d * dstride) + interval * (float)d;   // common.cuh hypothesis(), branch-free
            float gx, gy;
            if (GEOM == MVSB200_GEOM_MVS) {
                const float qx = ax * dv + bx, qy = ay * dv + by, qz = az * dv + bz;
                const float rz = rcp_nr(qz);
                float px = div_by(qx, qz, rz), py = div_by(qy, qz, rz);
                if (qz <= 0.f) px = -10.f, py = -10.f;
                gx = clampf(div_by(px, nx, rnx) - 1.f, -10.f, 10.f);
                gy = clampf(div_by(py, ny, rny) - 1.f, -10.f, 10.f);
            } else {
                const float dd = dv + 1e-9f;
                const float f = div_by(np_, dd, rcp_nr(dd));
                const float qx = ax - bx * f, qy = ay - by * f, qz = az - bz * f;
                const float zc = fmaxf(qz, 1e-9f), rz = rcp_nr(zc);
                float u = div_by(qx, zc, rz), w = div_by(qy, zc, rz);
                if (!(qz > 0.f)) u = -10.f, w = -10.f;
                gx = clampf(div_by(u, nx, rnx) * 2.f - 1.f, -1.1f, 1.1f);
                gy = clampf(div_by(w, ny, rny) * 2.f - 1.f, -1.1f, 1.1f);
            }
            // grid_sample(bilinear, zeros, align_corners=True).  A sample whose 2x2 cell lies inside the map (the test is
            // false for NaN coordinates) needs none of make_taps' border handling: its weights and cell directly.
            PackedTaps t;
            {
                const float ix = ((gx + 1.f) / 2.f) * Wm1f, iy = ((gy + 1.f) / 2.f) * Hm1f;
                if (ix >= 0.f && ix < Wm1f && iy >= 0.f && iy < Hm1f) {
                    const float fx0 = floorf(ix), fy0 = floorf(iy);
                    const float x1 = fx0 + 1.f, y1 = fy0 + 1.f;
                    t.w00 = (x1 - ix) * (y1 - iy);
                    t.w01 = (ix - fx0) * (y1 - iy);
                    t.w10 = (x1 - ix) * (iy - fy0);
                    t.w11 = (ix - fx0) * (iy - fy0);
                    t.cell = (int)fy0 * Ws + (int)fx0;
                } else {
                    if (!(gx == gx) || !(gy == gy)) gx = gy = -10.f;   // NaN coordinates (degenerate cameras) sample nothing
                    t = make_taps(gx, gy, Hs, Ws);
                }
            }
            sts128(rec_st + rd * (R * S * 16), make_float4(t.w00, t.w01, t.w10, t.w11));
            sts32(rec_st_cell + rd * (R * S * 4), t.cell);
        }
        __syncwarp();

        // Window re-loads of hypothesis k (all S views): a real, warp-uniform branch per view (the compiler would otherwise
        // if-convert the block and issue its predicated instructions for every (hypothesis, view) whether or not any lane
        // re-loads).  They run one hypothesis AHEAD of the arithmetic -- after the taps of hypothesis k-1 have been consumed,
        // before its epilogue and store -- so the loads are in flight while other instructions issue, and the arithmetic of
        // the S views below is one straight-line block (8 independent FMA chains) that the branches do not cut.
#define K1M_RELOAD_VIEW(s, cell)                                                                                 \
        if (__any_sync(0xffffffffu, (cell) != cur[s])) K1M_REAL_BRANCH_BEGIN {                                   \
            if ((cell) != cur[s]) {                                                                              \
                /* block-uniform view base (uniform registers) + this lane's 16-byte vector index */             \
                const float4 *q0 = vbase[s] + ((unsigned)(cell) * L + sub);                                      \
                const float4 *q1 = q0 + vrow[s];                                                                 \
                ta[s] = __ldg(q0);                                                                               \
                tb[s] = __ldg(q0 + L);                                                                           \
                tc[s] = __ldg(q1);                                                                               \
                td[s] = __ldg(q1 + L);                                                                           \
                cur[s] = (cell);                                                                                 \
            }                                                                                                    \
        } K1M_REAL_BRANCH_END(seg);
#define K1M_RELOAD(k)                                                                                            \
        {                                                                                                        \
            int cells[S];                                                                                        \
            _Pragma("unroll") for (int s = 0; s < S; s++) cells[s] = lds32(rec + REC_CELL + ((k) * S + s) * 4);   \
            _Pragma("unroll") for (int s = 0; s < S; s++) { K1M_RELOAD_VIEW(s, cells[s]) }                       \
        }
        if (K1M_AHEAD(AGG)) K1M_RELOAD(0)
#pragma unroll
        for (int k = 0; k < K1M_HC; k++) {
            const int d = k0 + k;
            if (d >= d_end) break;                       // warp-uniform
            float2 m1a, m1b, m2a, m2b;
            float sum_exp = 0.f;
            if (AGG == MVSB200_AGG_VARIANCE || AGG == MVSB200_AGG_VARIANCE_MEAN) {
                m1a = r0; m1b = r1;
                m2a = __fmul2_rn(r0, r0); m2b = __fmul2_rn(r1, r1);
            } else {
                m1a = m1b = m2a = m2b = make_float2(0.f, 0.f);
            }
            int cells[S];                                // all S cell loads in flight before the first vote waits on one
            if (!K1M_AHEAD(AGG)) {
#pragma unroll
                for (int s = 0; s < S; s++) cells[s] = lds32(rec + REC_CELL + (k * S + s) * 4);
            }
#pragma unroll
            for (int s = 0; s < S; s++) {
                const float4 wt = lds128(rec + (k * S + s) * 16);             // {w00, w01, w10, w11}
                if (!K1M_AHEAD(AGG) && s % K1M_GROUP == 0) {     // the re-load checks of K1M_GROUP views, then their arithmetic as one block
#pragma unroll
                    for (int u = s; u < s + K1M_GROUP && u < S; u++) { K1M_RELOAD_VIEW(u, cells[u]) }
                }
                // accumulation order nw, ne, sw, se (ATen grid_sampler_2d)
                const float2 p00 = make_float2(wt.x, wt.x), p01 = make_float2(wt.y, wt.y), p10 = make_float2(wt.z, wt.z), p11 = make_float2(wt.w, wt.w);
                float2 wa = __fmul2_rn(make_float2(ta[s].x, ta[s].y), p00), wb = __fmul2_rn(make_float2(ta[s].z, ta[s].w), p00);
                wa = __ffma2_rn(make_float2(tb[s].x, tb[s].y), p01, wa); wb = __ffma2_rn(make_float2(tb[s].z, tb[s].w), p01, wb);
                wa = __ffma2_rn(make_float2(tc[s].x, tc[s].y), p10, wa); wb = __ffma2_rn(make_float2(tc[s].z, tc[s].w), p10, wb);
                wa = __ffma2_rn(make_float2(td[s].x, td[s].y), p11, wa); wb = __ffma2_rn(make_float2(td[s].z, td[s].w), p11, wb);
                if (AGG == MVSB200_AGG_VARIANCE || AGG == MVSB200_AGG_VARIANCE_MEAN) {
                    m1a = __fadd2_rn(m1a, wa); m1b = __fadd2_rn(m1b, wb);
                    m2a = __ffma2_rn(wa, wa, m2a); m2b = __ffma2_rn(wb, wb, m2b);
                } else if (AGG == MVSB200_AGG_SOFTMIN) {
                    const float2 m1 = make_float2(-1.f, -1.f);
                    float2 da = __ffma2_rn(r0, m1, wa), db = __ffma2_rn(r1, m1, wb);   // w - r (the product is exact)
                    da = __fmul2_rn(da, da); db = __fmul2_rn(db, db);
                    float ssd = (da.x + da.y) + (db.x + db.y);
#pragma unroll
                    for (int m = L / 2; m >= 1; m >>= 1) ssd += __shfl_xor_sync(0xffffffffu, ssd, m);
                    const float ex = expf(-temp * ssd);
                    sum_exp += ex;
                    const float2 ee = make_float2(ex, ex);
                    m1a = __ffma2_rn(da, ee, m1a); m1b = __ffma2_rn(db, ee, m1b);
                } else {  // GROUPCORR: a lane's 4 channels are one group; one output volume per source view
                    float g = r0.x * wa.x;
                    g += r0.y * wa.y;
                    g += r1.x * wb.x;
                    g += r1.y * wb.y;
                    vmax = fmaxf(vmax, fabsf(g));
                    if (active) __stcs(outg + s * p.out_view_stride, g);   // this lane's group of plane d; + the view's volume (uniform)
                }
            }
            if (K1M_AHEAD(AGG) && k + 1 < K1M_HC && d + 1 < d_end) K1M_RELOAD(k + 1)
            if (AGG != MVSB200_AGG_GROUPCORR && active) {
                float2 oa, ob;
                if (AGG == MVSB200_AGG_VARIANCE) {           // M2/V - M1^2/V^2 (models/MVSNet/model.py:134)
                    const float2 nv2 = make_float2(-rV2, -rV2), v1 = make_float2(rV, rV);
                    oa = __ffma2_rn(m2a, v1, __fmul2_rn(__fmul2_rn(m1a, m1a), nv2));
                    ob = __ffma2_rn(m2b, v1, __fmul2_rn(__fmul2_rn(m1b, m1b), nv2));
                } else if (AGG == MVSB200_AGG_VARIANCE_MEAN) {   // M2/V - (M1/V)^2 (CVP_MVSNet/models/net.py:152)
                    const float2 v1 = make_float2(rV, rV), neg = make_float2(-1.f, -1.f);
                    const float2 ma = __fmul2_rn(m1a, v1), mb = __fmul2_rn(m1b, v1);
                    oa = __ffma2_rn(m2a, v1, __fmul2_rn(__fmul2_rn(ma, ma), neg));
                    ob = __ffma2_rn(m2b, v1, __fmul2_rn(__fmul2_rn(mb, mb), neg));
                } else {
                    const float rden = 1.f / (sum_exp + 1e-6f);
                    oa = __fmul2_rn(m1a, make_float2(rden, rden));
                    ob = __fmul2_rn(m1b, make_float2(rden, rden));
                }
                vmax = fmaxf(fmaxf(vmax, fmaxf(fabsf(oa.x), fabsf(oa.y))), fmaxf(fabsf(ob.x), fabsf(ob.y)));
                __stcs(outp, make_float4(oa.x, oa.y, ob.x, ob.y));
            }
            outp += plane_vecs;
            outg += plane_vecs;       // (floats: a plane of the L-group volume has HW * L of them)
        }
#undef K1M_RELOAD
#undef K1M_RELOAD_VIEW
    }
    if (p.out_amax) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, m));
        if (lane == 0 && vmax > 0.f) atomicMax(reinterpret_cast<unsigned int *>(p.out_amax), __float_as_uint(vmax));
    }
}

// Host: pixel tiles of (128 / C) x 4; the depth axis is cut into as few segments as still give the grid ~8 waves of blocks
// (a segment starts with one forced window load per view, so longer is better).  Measured at cfg2 (profiles/k1m_nseg_round2k.log):
// 3, 4 (the choice below) and 5 segments tie at 0.260 ms, 6 / 8 / 12 cost 0.264 / 0.271 / 0.285 ms.
static bool k1m_supported(const mvsb200_cost_volume_desc *d)
{
    return (d->C == 16 || d->C == 32) && (d->S == 1 || d->S == 2 || d->S == 4);
}
static int k1m_grid(const mvsb200_cost_volume_desc *d, dim3 &grid, int &seg, const char *what)
{
    const int pxw = 32 / (d->C / 4), th = K1M_THREADS / 32;
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    // the kernel tiles the image pxw x th or th x pxw (see `stack_y`): enough blocks for either
    const long long tiles_a = (long long)((d->W + pxw - 1) / pxw) * ((d->H + th - 1) / th);
    const long long tiles_b = (long long)((d->W + th - 1) / th) * ((d->H + pxw - 1) / pxw);
    const long long tiles = tiles_a > tiles_b ? tiles_a : tiles_b;
    MVSB200_REQUIRE(tiles < (1ll << 31), "%s: image too large", what);
    const int chunks = (d->D + K1M_HC - 1) / K1M_HC;
    long long nseg = (32ll * sms + tiles * d->B - 1) / (tiles * d->B);
    if (const char *e = getenv("MVSB200_K1M_NSEG")) { if (atoi(e) > 0) nseg = atoi(e); }   // experiment switch (profiles/k1_ab.py)
    if (nseg > chunks) nseg = chunks;
    if (nseg < 1) nseg = 1;
    seg = (int)((chunks + nseg - 1) / nseg) * K1M_HC;
    nseg = (d->D + seg - 1) / seg;
    MVSB200_REQUIRE(d->B <= 65535 && nseg <= 65535, "%s: B or D too large", what);
    grid = dim3((unsigned)tiles, (unsigned)nseg, (unsigned)d->B);
    return MVSB200_OK;
}

template <int C, int GEOM, int S>
static int launch_m_agg(const K1Params &p, int agg, dim3 grid, int seg, cudaStream_t st)
{
    switch (agg) {
    case MVSB200_AGG_VARIANCE: k1m_cost_volume_kernel<C, GEOM, MVSB200_AGG_VARIANCE, S><<<grid, K1M_THREADS, 0, st>>>(p, seg); break;
    case MVSB200_AGG_VARIANCE_MEAN: k1m_cost_volume_kernel<C, GEOM, MVSB200_AGG_VARIANCE_MEAN, S><<<grid, K1M_THREADS, 0, st>>>(p, seg); break;
    case MVSB200_AGG_SOFTMIN: k1m_cost_volume_kernel<C, GEOM, MVSB200_AGG_SOFTMIN, S><<<grid, K1M_THREADS, 0, st>>>(p, seg); break;
    case MVSB200_AGG_GROUPCORR: k1m_cost_volume_kernel<C, GEOM, MVSB200_AGG_GROUPCORR, S><<<grid, K1M_THREADS, 0, st>>>(p, seg); break;
    default: set_error("build_cost_volume: unknown aggregation %d", agg); return MVSB200_E_INVALID;
    }
    return check_launch("k1m_cost_volume_kernel");
}
template <int C, int GEOM>
static int launch_m_views(const K1Params &p, int S, int agg, dim3 grid, int seg, cudaStream_t st)
{
    if (S == 1) return launch_m_agg<C, GEOM, 1>(p, agg, grid, seg, st);
    if (S == 2) return launch_m_agg<C, GEOM, 2>(p, agg, grid, seg, st);
    return launch_m_agg<C, GEOM, 4>(p, agg, grid, seg, st);
}
template <int C>
static int launch_m_geom(const K1Params &p, int geom, int agg, dim3 grid, int seg, cudaStream_t st)
{
    if (geom == MVSB200_GEOM_MVS) return launch_m_views<C, MVSB200_GEOM_MVS>(p, p.S, agg, grid, seg, st);
    if (geom == MVSB200_GEOM_VIS) return launch_m_views<C, MVSB200_GEOM_VIS>(p, p.S, agg, grid, seg, st);
    set_error("build_cost_volume: unknown geometry %d", geom);
    return MVSB200_E_INVALID;
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
    // feature maps and the volume are moved with 256-bit (pixel-major kernel) / 128-bit (depth-marching kernel) accesses
    MVSB200_REQUIRE(((reinterpret_cast<uintptr_t>(ref) | reinterpret_cast<uintptr_t>(out)) & 31) == 0,
                    "build_cost_volume: ref and out must be 32-byte aligned");
    for (int s = 0; s < d->S && s < MVSB200_MAX_SRC; s++)
        MVSB200_REQUIRE(src[s] && (reinterpret_cast<uintptr_t>(src[s]) & 31) == 0, "build_cost_volume: src[%d] must be non-null and 32-byte aligned", s);
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
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid;
    // Which kernel: the depth-marching one pays off where a block marches a LONG run of hypotheses whose samples move by a
    // fraction of a source pixel (scalar hypotheses, D >= 96: cfg2 0.259 vs 0.310 ms); for short sweeps and per-pixel
    // hypotheses -- Vis-MVSNet's stages (D = 16 ... 64 around the previous stage's depth), CVP-MVSNet's refinement levels
    // (D = 8, one source pixel per hypothesis by construction), MVSNet-s (D = 48) -- its start-up (the geometry of a
    // mini-chunk, the first windows of every view) is not amortised and the pixel-tile kernel wins (profiles/k1_ab.py:
    // Vis stage 3 0.95 vs 1.24 ms, CVP level 0 0.98 vs 1.86 ms, cfg1 0.067 vs 0.085 ms).  MVSB200_K1 = v1 | m forces one (A/B runs).
    const char *k1_env = getenv("MVSB200_K1");   // read per call: the parity tests run both kernels in one process
    const int forced = !k1_env ? 0 : (k1_env[0] == 'v' ? 1 : (k1_env[0] == 'm' ? 2 : 0));
    const bool long_march = d->depth_mode != MVSB200_DEPTH_VOLUME && d->depth_mode != MVSB200_DEPTH_START_MAP && d->D >= 96;
    if (k1m_supported(d) && forced != 1 && (forced == 2 || long_march)) {   // depth-marching kernel
        int seg = 0;
        if (int rc = k1m_grid(d, grid, seg, "build_cost_volume")) return rc;
        p.chunks = 0;
        return d->C == 16 ? launch_m_geom<16>(p, d->geom, d->agg, grid, seg, st) : launch_m_geom<32>(p, d->geom, d->agg, grid, seg, st);
    }
    if (int rc = k1_grid(p, d, grid, "build_cost_volume")) return rc;
#ifdef MVSB200_K1_EXPERIMENTS
    {
        const char *e = getenv("MVSB200_K1_DEBUG");
        p.dbg = e ? atoi(e) : 0;
    }
#endif
    switch (d->C) {
    case 8: return launch_geom<8>(p, d->geom, d->agg, grid, st);
    case 16: return launch_geom<16>(p, d->geom, d->agg, grid, st);
    default: return launch_geom<32>(p, d->geom, d->agg, grid, st);
    }
}
