// K3: softmax over the depth axis + regression heads, and K4: visibility-weighted fusion.
//
// K3 reads the score volume [B,D,H,W] with a block of 32 pixels x 8 depth lanes: consecutive lanes of a warp sit on
// consecutive pixels (every load of a depth slice is one coalesced 128-byte line) and the 8 warps of a block split the
// depth axis (d = warp, warp + 8, ...), combining their partial max / normaliser / expectations through shared
// memory.  The three sweeps run out of registers (a lane's share of the column is loaded once, see NV below).  (One thread per pixel, the first
// version, left 20 k threads walking D serially: 60 us of pure latency for 16 MB.)
#include "common.cuh"

namespace mvsb200 {

constexpr int K3_PIX = 32, K3_DL = 8, K3_THREADS = K3_PIX * K3_DL;

struct K3Params {
    const float *score, *depth, *interval;
    float *depth_out, *conf_out, *entropy_out, *prob_out;
    int B, D, depth_mode, conf_mode;
    long long HW;
};

// DL = depth lanes per pixel: 8 for deep volumes (MVSNet's D = 192: the lanes split the sweep and combine through shared
// memory), 1 for the shallow ones of the cascades (Vis-MVSNet stages D = 16 ... 64, CVP refinement levels D = 8): with a
// handful of hypotheses per lane the partial-sum exchange and its three barriers cost more than the sweep -- a thread then
// owns its pixel, a block 256 pixels.
//
// NV > 0: the lane's share of the column (at most NV scores: d = dl, dl + DL, ...) is loaded ONCE into registers -- all
// loads of the sweep in flight together -- and the three sweeps run out of registers; NV = 0 streams the column three
// times (columns deeper than NV * DL).  Same operations in the same order either way: the results are bit-identical.
// Measured (profiles/k3_nv_round2k.log): NV = 24 with five resident blocks per SM (48 registers) beats the streaming kernel by
// ~10 % on every shape; NV = 32 at 63 registers (four blocks: the 640 blocks of cfg2 then need a second wave) is slower than it.
#ifndef K3_NV
#define K3_NV 24
#endif
#ifndef K3_MIN_BLOCKS
#define K3_MIN_BLOCKS 5
#endif
template <int DL, int NV>
__global__ void __launch_bounds__(K3_THREADS, K3_MIN_BLOCKS) k3_depth_regress_kernel(const K3Params p)
{
    constexpr int K3_DL = DL, K3_PIX = K3_THREADS / DL;
    __shared__ float red[4][K3_DL][K3_PIX];
    const int b = blockIdx.y, lane = threadIdx.x % K3_PIX, dl = threadIdx.x / K3_PIX;
    const long long pix_raw = (long long)blockIdx.x * K3_PIX + lane;
    const bool active = pix_raw < p.HW;
    const long long pix = active ? pix_raw : p.HW - 1;
    const float *s = p.score + (long long)b * p.D * p.HW + pix;
    const int D = p.D;

    // sweep(f): f(i, d, score) for the lane's hypotheses d = dl + i * DL, in order
    float v[NV > 0 ? NV : 1];
    if (NV > 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            const int d = dl + i * K3_DL;
            if (d >= D) break;
            v[i] = __ldg(s + d * p.HW);
        }
    }
#define K3_SWEEP(BODY)                                                                        \
    if (NV > 0) {                                                                             \
        _Pragma("unroll") for (int i = 0; i < NV; i++) {                                      \
            const int d = dl + i * K3_DL;                                                     \
            if (d >= D) break;                                                                \
            const float sc = v[i];                                                            \
            BODY                                                                              \
        }                                                                                     \
    } else {                                                                                  \
        for (int d = dl; d < D; d += K3_DL) {                                                 \
            const float sc = __ldg(s + d * p.HW);                                             \
            BODY                                                                              \
        }                                                                                     \
    }
    float mx = -INFINITY;
    K3_SWEEP(mx = fmaxf(mx, sc);)
    red[0][dl][lane] = mx;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K3_DL; k++) mx = fmaxf(mx, red[0][k][lane]);
    float sum = 0.f;
    K3_SWEEP(sum += expf(sc - mx);)
    red[1][dl][lane] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int k = 0; k < K3_DL; k++) sum += red[1][k][lane];

    const float interval = (p.depth_mode >= MVSB200_DEPTH_START) ? __ldg(p.interval + b) : 0.f;
    float e_idx = 0.f, e_dep = 0.f, ent = 0.f;
    float *prob = (p.prob_out && active) ? p.prob_out + (long long)b * D * p.HW + pix : nullptr;
    K3_SWEEP(
        const float pr = expf(sc - mx) / sum;
        if (prob) prob[d * p.HW] = pr;
        e_idx += pr * (float)d;
        if (p.depth_mode == MVSB200_DEPTH_VALUES) e_dep += pr * __ldg(p.depth + (long long)b * D + d);
        else if (p.depth_mode == MVSB200_DEPTH_VOLUME) e_dep += pr * __ldg(p.depth + ((long long)b * D + d) * p.HW + pix);
        if (p.entropy_out) ent += -pr * logf(fminf(fmaxf(pr, 1e-9f), 1.f));
    )
#undef K3_SWEEP
    __syncthreads();   // red[0] / red[1] are reused below
    red[0][dl][lane] = e_idx;
    red[2][dl][lane] = e_dep;
    red[3][dl][lane] = ent;
    __syncthreads();
    if (dl != 0 || !active) return;
    e_idx = e_dep = ent = 0.f;
#pragma unroll
    for (int k = 0; k < K3_DL; k++) { e_idx += red[0][k][lane]; e_dep += red[2][k][lane]; ent += red[3][k][lane]; }
    if (p.depth_mode == MVSB200_DEPTH_START) e_dep = e_idx * interval + __ldg(p.depth + b);
    else if (p.depth_mode == MVSB200_DEPTH_START_MAP) e_dep = e_idx * interval + __ldg(p.depth + (long long)b * p.HW + pix);
    p.depth_out[(long long)b * p.HW + pix] = e_dep;
    if (p.entropy_out) p.entropy_out[(long long)b * p.HW + pix] = ent;
    if (p.conf_out) {
        float c = 0.f;
        if (p.conf_mode == MVSB200_CONF_SUM4) {
            const int idx = (int)e_idx;  // .long() truncation, models/MVSNet/model.py:214
            for (int j = idx - 1; j <= idx + 2; j++)
                if (j >= 0 && j < D) c += expf(__ldg(s + j * p.HW) - mx) / sum;
        } else {
            // |d - E[d]| <= 2 touches at most 5 bins
            const int lo = max((int)ceilf(e_idx - 2.f), 0), hi = min((int)floorf(e_idx + 2.f), D - 1);
            for (int j = lo; j <= hi; j++)
                if (fabsf((float)j - e_idx) <= 2.f) c += expf(__ldg(s + j * p.HW) - mx) / sum;
        }
        p.conf_out[(long long)b * p.HW + pix] = c;
    }
}

// K3 backward (row f2): depth = sum_d p_d h_d with p = softmax(score)  =>  dL/dscore_d = g p_d (h_d - depth).
// The confidence / entropy outputs carry no gradient in the reference (torch.no_grad, models/MVSNet/model.py:211).
// Same block shape as the forward kernel; the statistics are recomputed instead of stored.
struct K3BwdParams {
    const float *score, *depth, *interval, *gdepth;
    float *gscore, *ghyp;
    int B, D, depth_mode;
    long long HW;
};

__global__ void __launch_bounds__(K3_THREADS) k3_depth_regress_backward_kernel(const K3BwdParams p)
{
    __shared__ float red[3][K3_DL][K3_PIX];
    const int b = blockIdx.y, lane = threadIdx.x & 31, dl = threadIdx.x >> 5;
    const long long pix_raw = (long long)blockIdx.x * K3_PIX + lane;
    const bool active = pix_raw < p.HW;
    const long long pix = active ? pix_raw : p.HW - 1;
    const float *s = p.score + (long long)b * p.D * p.HW + pix;
    const int D = p.D;
    const float interval = (p.depth_mode >= MVSB200_DEPTH_START) ? __ldg(p.interval + b) : 0.f;
    auto hyp = [&](int d) { return hypothesis(p.depth_mode, p.depth, interval, b, d, D, p.HW, pix); };

    float mx = -INFINITY;
    for (int d = dl; d < D; d += K3_DL) mx = fmaxf(mx, __ldg(s + d * p.HW));
    red[0][dl][lane] = mx;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K3_DL; k++) mx = fmaxf(mx, red[0][k][lane]);
    float sum = 0.f, e_dep = 0.f;
    for (int d = dl; d < D; d += K3_DL) {
        const float e = expf(__ldg(s + d * p.HW) - mx);
        sum += e;
        e_dep += e * hyp(d);
    }
    red[1][dl][lane] = sum;
    red[2][dl][lane] = e_dep;
    __syncthreads();
    sum = e_dep = 0.f;
#pragma unroll
    for (int k = 0; k < K3_DL; k++) { sum += red[1][k][lane]; e_dep += red[2][k][lane]; }
    if (!active) return;
    e_dep /= sum;
    const float g = __ldg(p.gdepth + (long long)b * p.HW + pix) / sum;
    float *gs = p.gscore + (long long)b * D * p.HW + pix;
    float *gh = p.ghyp ? p.ghyp + (long long)b * D * p.HW + pix : nullptr;   // per-voxel hypotheses: dL/dh_d = g p_d
    for (int d = dl; d < D; d += K3_DL) {
        const float gp = g * expf(__ldg(s + d * p.HW) - mx);
        gs[d * p.HW] = gp * (hyp(d) - e_dep);
        if (gh) gh[d * p.HW] = gp;
    }
}

constexpr int K4_THREADS = 256;

struct K4Params {
    const float *interm[MVSB200_MAX_SRC];
    const float *uncert[MVSB200_MAX_SRC];
    float *fused;
    int S, D, G;
    long long HW, total4;  // total4 = B*D*HW*G/4
};

// One thread per float4 of the fused volume; the per-pixel weights exp(-u_s) are recomputed per voxel
// (S expf per float4 is noise next to the S+1 16-byte memory accesses).
__global__ void __launch_bounds__(K4_THREADS) k4_vis_fuse_kernel(const K4Params p)
{
    const long long i = (long long)blockIdx.x * K4_THREADS + threadIdx.x;
    if (i >= p.total4) return;
    const long long vox = (i * 4) / p.G;              // b*D*HW + d*HW + pix
    const long long b = vox / ((long long)p.D * p.HW);
    const long long pix = vox % p.HW;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float wsum = 0.f;
    for (int s = 0; s < p.S; s++) {
        const float w = expf(-__ldg(p.uncert[s] + b * p.HW + pix));
        const float4 v = __ldcs(reinterpret_cast<const float4 *>(p.interm[s]) + i);
        wsum = wsum + w;
        acc.x = acc.x + v.x * w; acc.y = acc.y + v.y * w; acc.z = acc.z + v.z * w; acc.w = acc.w + v.w * w;
    }
    acc.x /= wsum; acc.y /= wsum; acc.z /= wsum; acc.w /= wsum;
    reinterpret_cast<float4 *>(p.fused)[i] = acc;
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_depth_regress(const float *score, int B, int D, int H, int W, int depth_mode,
                                     const float *depth, const float *interval, int conf_mode, float *depth_out,
                                     float *conf_out, float *entropy_out, float *prob_out, mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(score && depth && depth_out, "depth_regress: null pointer");
    MVSB200_REQUIRE(B > 0 && B <= 65535 && D > 0 && H > 0 && W > 0, "depth_regress: bad shape B=%d D=%d H=%d W=%d", B, D, H, W);
    MVSB200_REQUIRE(depth_mode >= 0 && depth_mode <= 3, "depth_regress: depth_mode=%d", depth_mode);
    MVSB200_REQUIRE(depth_mode < MVSB200_DEPTH_START || interval, "depth_regress: interval is null");
    MVSB200_REQUIRE(conf_mode >= 0 && conf_mode <= 2, "depth_regress: conf_mode=%d", conf_mode);
    MVSB200_REQUIRE(conf_mode == MVSB200_CONF_NONE || conf_out, "depth_regress: conf_out is null");
    K3Params p;
    p.score = score; p.depth = depth; p.interval = interval;
    p.depth_out = depth_out; p.conf_out = conf_mode ? conf_out : nullptr; p.entropy_out = entropy_out; p.prob_out = prob_out;
    p.B = B; p.D = D; p.depth_mode = depth_mode; p.conf_mode = conf_mode;
    p.HW = (long long)H * W;
    // one thread per pixel needs enough pixels to fill the machine while each thread walks D serially
    if (D <= 32 || (D <= 64 && (long long)B * p.HW >= (1ll << 18))) {
        dim3 grid((unsigned)((p.HW + K3_THREADS - 1) / K3_THREADS), (unsigned)B);
        if (D <= K3_NV) k3_depth_regress_kernel<1, K3_NV><<<grid, K3_THREADS, 0, (cudaStream_t)stream>>>(p);
        else k3_depth_regress_kernel<1, 0><<<grid, K3_THREADS, 0, (cudaStream_t)stream>>>(p);
    } else {
        dim3 grid((unsigned)((p.HW + K3_PIX - 1) / K3_PIX), (unsigned)B);
        if (D <= K3_NV * K3_DL) k3_depth_regress_kernel<K3_DL, K3_NV><<<grid, K3_THREADS, 0, (cudaStream_t)stream>>>(p);
        else k3_depth_regress_kernel<K3_DL, 0><<<grid, K3_THREADS, 0, (cudaStream_t)stream>>>(p);
    }
    return check_launch("k3_depth_regress_kernel");
}

extern "C" int mvsb200_vis_fuse(const float *const *interm, const float *const *uncert, int S, int B, int D, int H,
                                int W, int G, float *fused, mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(interm && uncert && fused, "vis_fuse: null pointer");
    MVSB200_REQUIRE(S >= 1 && S <= MVSB200_MAX_SRC, "vis_fuse: S=%d", S);
    MVSB200_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && G > 0 && G % 4 == 0, "vis_fuse: bad shape (G must be a multiple of 4)");
    K4Params p;
    for (int s = 0; s < S; s++) {
        MVSB200_REQUIRE(interm[s] && uncert[s], "vis_fuse: null view %d", s);
        p.interm[s] = interm[s];
        p.uncert[s] = uncert[s];
    }
    p.fused = fused; p.S = S; p.D = D; p.G = G;
    p.HW = (long long)H * W;
    p.total4 = (long long)B * D * p.HW * G / 4;
    long long blocks = (p.total4 + K4_THREADS - 1) / K4_THREADS;
    MVSB200_REQUIRE(blocks < (1ll << 31), "vis_fuse: volume too large");
    k4_vis_fuse_kernel<<<(unsigned)blocks, K4_THREADS, 0, (cudaStream_t)stream>>>(p);
    return check_launch("k4_vis_fuse_kernel");
}

extern "C" int mvsb200_depth_regress_backward(const float *score, int B, int D, int H, int W, int depth_mode, const float *depth,
                                              const float *interval, const float *grad_depth, float *grad_score,
                                              float *grad_hyp, mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(score && depth && grad_depth && grad_score, "depth_regress_backward: null pointer");
    MVSB200_REQUIRE(!grad_hyp || depth_mode == MVSB200_DEPTH_VOLUME, "depth_regress_backward: grad_hyp needs per-voxel hypotheses (DEPTH_VOLUME)");
    MVSB200_REQUIRE(B > 0 && B <= 65535 && D > 0 && H > 0 && W > 0, "depth_regress_backward: bad shape B=%d D=%d H=%d W=%d", B, D, H, W);
    MVSB200_REQUIRE(depth_mode >= 0 && depth_mode <= 3, "depth_regress_backward: depth_mode=%d", depth_mode);
    MVSB200_REQUIRE(depth_mode < MVSB200_DEPTH_START || interval, "depth_regress_backward: interval is null");
    K3BwdParams p;
    p.score = score; p.depth = depth; p.interval = interval; p.gdepth = grad_depth; p.gscore = grad_score; p.ghyp = grad_hyp;
    p.B = B; p.D = D; p.depth_mode = depth_mode; p.HW = (long long)H * W;
    dim3 grid((unsigned)((p.HW + K3_PIX - 1) / K3_PIX), (unsigned)B);
    k3_depth_regress_backward_kernel<<<grid, K3_THREADS, 0, (cudaStream_t)stream>>>(p);
    return check_launch("k3_depth_regress_backward_kernel");
}
