// K2 (fp32 CUDA-core path): 3-D convolution / transposed convolution on channels-last volumes with the
// BN scale+bias, ReLU and skip-add of the reference's ConvBnReLU3D / deconv blocks fused in the epilogue.
//
// One 128-thread CTA owns a TZ x TY x 32 tile of output voxels (for the transposed conv: of INPUT
// positions, one CTA per output-parity class) and COUT_T output channels.  Lanes run along x, so with
// the input tile staged in shared memory as [channel-quad][position][4 floats] every warp-level
// LDS.128 is conflict free, and the weights of the current (tap, channel-quad) are warp-uniform
// broadcast loads.  Each thread accumulates VX voxels x COUT_T channels in registers while the input
// channels stream through shared memory CK at a time.
//
// Transposed conv (k=3, stride 2, padding 1, output_padding 1) is evaluated in gather form per output
// parity (pz,py,px): even coordinates take tap 1 at input i, odd ones take tap 0 at i+1 and tap 2 at i.
#include "common.cuh"

namespace mvsb200 {

constexpr int K2_THREADS = 128;
constexpr int MODE_S1 = 0, MODE_S2 = 1, MODE_DECONV = 2;

template <int MODE> struct TileCfg;
template <> struct TileCfg<MODE_S1> { static constexpr int TZ = 2, TY = 8, TX = 32, VX = 4, CK = 8; };
template <> struct TileCfg<MODE_S2> { static constexpr int TZ = 2, TY = 4, TX = 32, VX = 2, CK = 4; };
template <> struct TileCfg<MODE_DECONV> { static constexpr int TZ = 2, TY = 8, TX = 32, VX = 4, CK = 8; };

struct K2Params {
    const float *x, *x2, *w, *scale, *bias, *skip;
    float *y;
    int B, D, H, W, Do, Ho, Wo;
    int Cin1, Cin2, Cout;
    int kd, kh, kw, pd, ph, pw;
    int relu, skip_mode;
    int tiles_x, tiles_y, tiles_z;
};

template <int MODE>
__host__ __device__ inline void tile_extents(int kd, int kh, int kw, int &ez, int &ey, int &ex)
{
    using T = TileCfg<MODE>;
    if (MODE == MODE_S1) { ez = T::TZ + kd - 1; ey = T::TY + kh - 1; ex = T::TX + kw - 1; }
    else if (MODE == MODE_S2) { ez = (T::TZ - 1) * 2 + kd; ey = (T::TY - 1) * 2 + kh; ex = (T::TX - 1) * 2 + kw; }
    else { ez = T::TZ + 1; ey = T::TY + 1; ex = T::TX + 1; }
}

template <int MODE, int COUT_T>
__global__ void __launch_bounds__(K2_THREADS) k2_conv3d_kernel(const K2Params p)
{
    using T = TileCfg<MODE>;
    constexpr int CK = T::CK, VX = T::VX, NQ = CK / 4;
    extern __shared__ float4 smem4[];
    __shared__ int s_tap_off[27], s_tap_w[27];
    __shared__ int s_ntaps;

    int ez, ey, ex;
    tile_extents<MODE>(p.kd, p.kh, p.kw, ez, ey, ex);
    const int npos = ez * ey * ex;
    float4 *s_in = smem4;                                       // [NQ][npos]
    float *s_w = reinterpret_cast<float *>(smem4 + NQ * npos);  // [ntaps][CK][COUT_T]

    const int tid = threadIdx.x;
    int t = blockIdx.x;
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y; t /= p.tiles_y;
    const int tz = t % p.tiles_z;
    const int b = t / p.tiles_z;
    const int co0 = blockIdx.y * COUT_T;
    const int z0 = tz * T::TZ, y0 = ty * T::TY, x0 = tx * T::TX;
    int pz = 0, py = 0, px = 0;
    if (MODE == MODE_DECONV) { pz = blockIdx.z >> 2; py = (blockIdx.z >> 1) & 1; px = blockIdx.z & 1; }

    int iz0, iy0, ix0;
    if (MODE == MODE_S1) { iz0 = z0 - p.pd; iy0 = y0 - p.ph; ix0 = x0 - p.pw; }
    else if (MODE == MODE_S2) { iz0 = 2 * z0 - p.pd; iy0 = 2 * y0 - p.ph; ix0 = 2 * x0 - p.pw; }
    else { iz0 = z0; iy0 = y0; ix0 = x0; }

    if (tid == 0) {
        int n = 0;
        if (MODE == MODE_DECONV) {
            // parity 0: (tap 1, +0);  parity 1: (tap 0, +1), (tap 2, +0)
            for (int a = 0; a < 3; a++) {
                if ((a == 1) != (pz == 0)) continue;
                for (int bb = 0; bb < 3; bb++) {
                    if ((bb == 1) != (py == 0)) continue;
                    for (int e = 0; e < 3; e++) {
                        if ((e == 1) != (px == 0)) continue;
                        int oz = (a == 0), oy = (bb == 0), ox = (e == 0);
                        s_tap_off[n] = (oz * ey + oy) * ex + ox;
                        s_tap_w[n] = (a * 3 + bb) * 3 + e;
                        n++;
                    }
                }
            }
        } else {
            for (int a = 0; a < p.kd; a++)
                for (int bb = 0; bb < p.kh; bb++)
                    for (int e = 0; e < p.kw; e++) {
                        s_tap_off[n] = (a * ey + bb) * ex + e;
                        s_tap_w[n] = (a * p.kh + bb) * p.kw + e;
                        n++;
                    }
        }
        s_ntaps = n;
    }

    // per-thread voxel slots
    int base[VX], lz[VX], ly[VX], lx[VX];
#pragma unroll
    for (int v = 0; v < VX; v++) {
        int lin = v * K2_THREADS + tid;
        lx[v] = lin % T::TX;
        ly[v] = (lin / T::TX) % T::TY;
        lz[v] = lin / (T::TX * T::TY);
        const int m = (MODE == MODE_S2) ? 2 : 1;
        base[v] = ((lz[v] * m) * ey + ly[v] * m) * ex + lx[v] * m;
    }

    float acc[VX][COUT_T];
#pragma unroll
    for (int v = 0; v < VX; v++)
#pragma unroll
        for (int c = 0; c < COUT_T; c++) acc[v][c] = 0.f;

    const int CinT = p.Cin1 + p.Cin2;
    __syncthreads();
    const int ntaps = s_ntaps;

    for (int c0 = 0; c0 < CinT; c0 += CK) {
        if (c0) __syncthreads();
        const float *src;
        int cs, cstride;
        if (c0 < p.Cin1) { src = p.x; cs = c0; cstride = p.Cin1; }
        else { src = p.x2; cs = c0 - p.Cin1; cstride = p.Cin2; }
        // ---- stage the input tile (zero padded) ----
        for (int i = tid; i < npos * NQ; i += K2_THREADS) {
            const int q = i % NQ, pos = i / NQ;
            const int ix = pos % ex, r = pos / ex;
            const int iy = r % ey, iz = r / ey;
            const int gz = iz0 + iz, gy = iy0 + iy, gx = ix0 + ix;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gz >= 0 && gz < p.D && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)
                v = ldg4(src + ((((long long)b * p.D + gz) * p.H + gy) * p.W + gx) * cstride + cs + q * 4);
            s_in[q * npos + pos] = v;
        }
        // ---- stage the weights of this channel chunk: s_w[t][ci][co] ----
        for (int i = tid; i < ntaps * CK * COUT_T; i += K2_THREADS) {
            const int co = i % COUT_T, r = i / COUT_T;
            const int ci = r % CK, tt = r / CK;
            s_w[i] = __ldg(p.w + ((long long)s_tap_w[tt] * CinT + c0 + ci) * p.Cout + co0 + co);
        }
        __syncthreads();
        // ---- accumulate ----
        for (int tt = 0; tt < ntaps; tt++) {
            const int off = s_tap_off[tt];
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                float4 a[VX];
#pragma unroll
                for (int v = 0; v < VX; v++) a[v] = s_in[q * npos + base[v] + off];
                const float *wq = s_w + (tt * CK + q * 4) * COUT_T;
                if (COUT_T == 1) {
                    const float4 w4 = *reinterpret_cast<const float4 *>(wq);
#pragma unroll
                    for (int v = 0; v < VX; v++) {
                        acc[v][0] += a[v].x * w4.x;
                        acc[v][0] += a[v].y * w4.y;
                        acc[v][0] += a[v].z * w4.z;
                        acc[v][0] += a[v].w * w4.w;
                    }
                } else {
#pragma unroll
                    for (int ci = 0; ci < 4; ci++) {
                        float wv[COUT_T];
#pragma unroll
                        for (int c4 = 0; c4 < COUT_T / 4; c4++) {
                            const float4 w4 = *reinterpret_cast<const float4 *>(wq + ci * COUT_T + c4 * 4);
                            wv[c4 * 4 + 0] = w4.x; wv[c4 * 4 + 1] = w4.y; wv[c4 * 4 + 2] = w4.z; wv[c4 * 4 + 3] = w4.w;
                        }
#pragma unroll
                        for (int v = 0; v < VX; v++) {
                            const float av = (ci == 0) ? a[v].x : (ci == 1) ? a[v].y : (ci == 2) ? a[v].z : a[v].w;
#pragma unroll
                            for (int c = 0; c < COUT_T; c++) acc[v][c] += av * wv[c];
                        }
                    }
                }
            }
        }
    }

    // ---- epilogue: y = act(acc*scale + bias [+ skip]) [+ skip] ----
    float sc[COUT_T], bi[COUT_T];
#pragma unroll
    for (int c = 0; c < COUT_T; c++) {
        sc[c] = p.scale ? __ldg(p.scale + co0 + c) : 1.f;
        bi[c] = p.bias ? __ldg(p.bias + co0 + c) : 0.f;
    }
#pragma unroll
    for (int v = 0; v < VX; v++) {
        int oz = z0 + lz[v], oy = y0 + ly[v], ox = x0 + lx[v];
        if (MODE == MODE_DECONV) {
            if (oz >= p.D || oy >= p.H || ox >= p.W) continue;
            oz = 2 * oz + pz; oy = 2 * oy + py; ox = 2 * ox + px;
        }
        if (oz >= p.Do || oy >= p.Ho || ox >= p.Wo) continue;
        const long long o = ((((long long)b * p.Do + oz) * p.Ho + oy) * p.Wo + ox) * p.Cout + co0;
        float r[COUT_T];
#pragma unroll
        for (int c = 0; c < COUT_T; c++) r[c] = acc[v][c] * sc[c] + bi[c];
        if (p.skip_mode == MVSB200_SKIP_BEFORE_RELU) {
#pragma unroll
            for (int c = 0; c < COUT_T; c++) r[c] += __ldg(p.skip + o + c);
        }
        if (p.relu) {
#pragma unroll
            for (int c = 0; c < COUT_T; c++) r[c] = fmaxf(r[c], 0.f);
        }
        if (p.skip_mode == MVSB200_SKIP_AFTER_RELU) {
#pragma unroll
            for (int c = 0; c < COUT_T; c++) r[c] += __ldg(p.skip + o + c);
        }
        if (COUT_T == 1) p.y[o] = r[0];
        else {
#pragma unroll
            for (int c4 = 0; c4 < COUT_T / 4; c4++)
                st4(p.y + o + c4 * 4, make_float4(r[c4 * 4], r[c4 * 4 + 1], r[c4 * 4 + 2], r[c4 * 4 + 3]));
        }
    }
}

// Shape-agnostic fallback (one thread per output element) for layers the tiled kernel does not cover
// (Cin not a multiple of 4: the 1->8 conv of UncertNet, models/VisMVSNet/model_cas.py:81-85).
__global__ void k2_conv3d_naive_kernel(const K2Params p, int stride, int transposed)
{
    const long long n = (long long)p.B * p.Do * p.Ho * p.Wo * p.Cout;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int co = (int)(i % p.Cout);
    long long r = i / p.Cout;
    const int ox = (int)(r % p.Wo); r /= p.Wo;
    const int oy = (int)(r % p.Ho); r /= p.Ho;
    const int oz = (int)(r % p.Do);
    const int b = (int)(r / p.Do);
    const int CinT = p.Cin1 + p.Cin2;
    float acc = 0.f;
    for (int a = 0; a < p.kd; a++)
        for (int bb = 0; bb < p.kh; bb++)
            for (int e = 0; e < p.kw; e++) {
                int iz, iy, ix;
                if (!transposed) {
                    iz = oz * stride + a - p.pd; iy = oy * stride + bb - p.ph; ix = ox * stride + e - p.pw;
                } else {
                    int zn = oz + p.pd - a, yn = oy + p.ph - bb, xn = ox + p.pw - e;
                    if (zn < 0 || yn < 0 || xn < 0 || (zn % stride) || (yn % stride) || (xn % stride)) continue;
                    iz = zn / stride; iy = yn / stride; ix = xn / stride;
                }
                if (iz < 0 || iz >= p.D || iy < 0 || iy >= p.H || ix < 0 || ix >= p.W) continue;
                const long long vox = (((long long)b * p.D + iz) * p.H + iy) * p.W + ix;
                const float *wt = p.w + ((long long)((a * p.kh + bb) * p.kw + e) * CinT) * p.Cout + co;
                for (int c = 0; c < p.Cin1; c++) acc += __ldg(p.x + vox * p.Cin1 + c) * __ldg(wt + (long long)c * p.Cout);
                for (int c = 0; c < p.Cin2; c++)
                    acc += __ldg(p.x2 + vox * p.Cin2 + c) * __ldg(wt + (long long)(p.Cin1 + c) * p.Cout);
            }
    float v = acc * (p.scale ? __ldg(p.scale + co) : 1.f) + (p.bias ? __ldg(p.bias + co) : 0.f);
    if (p.skip_mode == MVSB200_SKIP_BEFORE_RELU) v += __ldg(p.skip + i);
    if (p.relu) v = fmaxf(v, 0.f);
    if (p.skip_mode == MVSB200_SKIP_AFTER_RELU) v += __ldg(p.skip + i);
    p.y[i] = v;
}

// 1x1x1 convolution (the stride-2 projection shortcut of Vis-MVSNet's BasicBlock, nn_utils.py:150-158): a per-voxel
// Cin x Cout matrix product.  One thread per (output voxel, 4 output channels); the weights sit in shared memory,
// the input voxel is a handful of 16-byte loads that the threads of a voxel share through L1.
__global__ void __launch_bounds__(256) k2_pointwise_kernel(const K2Params p, int stride)
{
    extern __shared__ float s_wpt[];   // [Cin][Cout]
    for (int i = threadIdx.x; i < p.Cin1 * p.Cout; i += blockDim.x) s_wpt[i] = __ldg(p.w + i);
    __syncthreads();
    const int q4 = p.Cout >> 2;
    const long long n = (long long)p.B * p.Do * p.Ho * p.Wo * q4;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c4 = (int)(i % q4);
    long long r = i / q4;
    const int ox = (int)(r % p.Wo); r /= p.Wo;
    const int oy = (int)(r % p.Ho); r /= p.Ho;
    const int oz = (int)(r % p.Do);
    const int b = (int)(r / p.Do);
    const float *xin = p.x + ((((long long)b * p.D + (long long)oz * stride) * p.H + (long long)oy * stride) * p.W + (long long)ox * stride) * p.Cin1;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < p.Cin1; c += 4) {
        const float4 v = ldg4(xin + c);
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float4 w = *reinterpret_cast<const float4 *>(&s_wpt[(c + k) * p.Cout + c4 * 4]);
            acc.x = fmaf(vv[k], w.x, acc.x); acc.y = fmaf(vv[k], w.y, acc.y); acc.z = fmaf(vv[k], w.z, acc.z); acc.w = fmaf(vv[k], w.w, acc.w);
        }
    }
    const long long o = (i / q4) * p.Cout + c4 * 4;
    float y[4] = {acc.x, acc.y, acc.z, acc.w};
    if (p.scale) { const float4 s4 = ldg4(p.scale + c4 * 4); y[0] *= s4.x; y[1] *= s4.y; y[2] *= s4.z; y[3] *= s4.w; }
    if (p.bias) { const float4 b4 = ldg4(p.bias + c4 * 4); y[0] += b4.x; y[1] += b4.y; y[2] += b4.z; y[3] += b4.w; }
    float4 sk = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.skip_mode != MVSB200_SKIP_NONE) sk = ldg4(p.skip + o);
    if (p.skip_mode == MVSB200_SKIP_BEFORE_RELU) { y[0] += sk.x; y[1] += sk.y; y[2] += sk.z; y[3] += sk.w; }
    if (p.relu) { y[0] = fmaxf(y[0], 0.f); y[1] = fmaxf(y[1], 0.f); y[2] = fmaxf(y[2], 0.f); y[3] = fmaxf(y[3], 0.f); }
    if (p.skip_mode == MVSB200_SKIP_AFTER_RELU) { y[0] += sk.x; y[1] += sk.y; y[2] += sk.z; y[3] += sk.w; }
    st4(p.y + o, make_float4(y[0], y[1], y[2], y[3]));
}

template <int MODE, int COUT_T>
static int launch_tiled(K2Params p, cudaStream_t st)
{
    using T = TileCfg<MODE>;
    int ez, ey, ex;
    tile_extents<MODE>(p.kd, p.kh, p.kw, ez, ey, ex);
    const int npos = ez * ey * ex;
    const int ntaps_max = (MODE == MODE_DECONV) ? 8 : p.kd * p.kh * p.kw;
    const size_t smem = (size_t)(T::CK / 4) * npos * sizeof(float4) + (size_t)ntaps_max * T::CK * COUT_T * sizeof(float);
    // tiles run over output voxels (conv) or input positions (transposed conv)
    const int nz = (MODE == MODE_DECONV) ? p.D : p.Do, ny = (MODE == MODE_DECONV) ? p.H : p.Ho,
              nx = (MODE == MODE_DECONV) ? p.W : p.Wo;
    p.tiles_z = (nz + T::TZ - 1) / T::TZ;
    p.tiles_y = (ny + T::TY - 1) / T::TY;
    p.tiles_x = (nx + T::TX - 1) / T::TX;
    const long long tiles = (long long)p.tiles_x * p.tiles_y * p.tiles_z * p.B;
    if (tiles >= (1ll << 31)) {
        set_error("conv3d: volume too large");
        return MVSB200_E_INVALID;
    }
    if (int rc = ensure_dynamic_smem(k2_conv3d_kernel<MODE, COUT_T>, 100 * 1024, "conv3d")) return rc;
    dim3 grid((unsigned)tiles, (unsigned)(p.Cout / COUT_T), (MODE == MODE_DECONV) ? 8 : 1);
    k2_conv3d_kernel<MODE, COUT_T><<<grid, K2_THREADS, smem, st>>>(p);
    return check_launch("k2_conv3d_kernel");
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_conv3d_out_shape(const mvsb200_conv3d_desc *d, int *Do, int *Ho, int *Wo)
{
    MVSB200_REQUIRE(d, "conv3d_out_shape: null descriptor");
    MVSB200_REQUIRE((d->kd == 1 || d->kd == 3) && (d->kh == 1 || d->kh == 3) && (d->kw == 1 || d->kw == 3),
                    "conv3d: kernel extents must be 1 or 3 (got %dx%dx%d)", d->kd, d->kh, d->kw);
    MVSB200_REQUIRE(d->stride == 1 || d->stride == 2, "conv3d: stride must be 1 or 2 (got %d)", d->stride);
    int o[3];
    const int in[3] = {d->D, d->H, d->W}, k[3] = {d->kd, d->kh, d->kw};
    for (int i = 0; i < 3; i++) {
        const int pad = k[i] / 2;
        if (d->transposed) o[i] = (in[i] - 1) * d->stride - 2 * pad + k[i] + (d->stride - 1);
        else o[i] = (in[i] + 2 * pad - k[i]) / d->stride + 1;
    }
    if (Do) *Do = o[0];
    if (Ho) *Ho = o[1];
    if (Wo) *Wo = o[2];
    return MVSB200_OK;
}

extern "C" int mvsb200_conv3d(const mvsb200_conv3d_desc *d, const float *x, const float *x2, const float *w,
                              const float *scale, const float *bias, const float *skip, float *y,
                              mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(d && x && w && y, "conv3d: null pointer");
    MVSB200_REQUIRE(d->B > 0 && d->D > 0 && d->H > 0 && d->W > 0, "conv3d: bad shape B=%d D=%d H=%d W=%d", d->B, d->D, d->H, d->W);
    MVSB200_REQUIRE(d->Cin > 0 && d->Cin2 >= 0 && d->Cout > 0, "conv3d: bad channels Cin=%d Cin2=%d Cout=%d", d->Cin, d->Cin2, d->Cout);
    MVSB200_REQUIRE(d->Cin2 == 0 || x2, "conv3d: Cin2=%d but x2 is null", d->Cin2);
    MVSB200_REQUIRE(d->skip_mode >= 0 && d->skip_mode <= 2, "conv3d: skip_mode=%d", d->skip_mode);
    MVSB200_REQUIRE(d->skip_mode == MVSB200_SKIP_NONE || skip, "conv3d: skip_mode=%d but skip is null", d->skip_mode);
    K2Params p;
    int rc = mvsb200_conv3d_out_shape(d, &p.Do, &p.Ho, &p.Wo);
    if (rc) return rc;
    MVSB200_REQUIRE(p.Do > 0 && p.Ho > 0 && p.Wo > 0, "conv3d: empty output");
    if (d->transposed)
        MVSB200_REQUIRE(d->kd == 3 && d->kh == 3 && d->kw == 3 && d->stride == 2,
                        "conv3d: transposed conv supports k=3 stride=2 only (stride-1 transposed convs are packed as flipped convs)");
    p.x = x; p.x2 = x2; p.w = w; p.scale = scale; p.bias = bias; p.skip = skip; p.y = y;
    p.B = d->B; p.D = d->D; p.H = d->H; p.W = d->W;
    p.Cin1 = d->Cin; p.Cin2 = d->Cin2; p.Cout = d->Cout;
    p.kd = d->kd; p.kh = d->kh; p.kw = d->kw;
    p.pd = d->kd / 2; p.ph = d->kh / 2; p.pw = d->kw / 2;
    p.relu = d->relu; p.skip_mode = d->skip_mode;
    p.tiles_x = p.tiles_y = p.tiles_z = 0;
    cudaStream_t st = (cudaStream_t)stream;

    if (d->kd == 1 && d->kh == 1 && d->kw == 1 && !d->transposed && d->Cin2 == 0 && d->Cin % 4 == 0 && d->Cout % 4 == 0 &&
        (size_t)d->Cin * d->Cout * sizeof(float) <= 48 * 1024) {
        const long long n = (long long)p.B * p.Do * p.Ho * p.Wo * (p.Cout / 4);
        MVSB200_REQUIRE(n < (1ll << 31) * 256, "conv3d: volume too large");
        k2_pointwise_kernel<<<(unsigned)((n + 255) / 256), 256, (size_t)d->Cin * d->Cout * sizeof(float), st>>>(p, d->stride);
        return check_launch("k2_pointwise_kernel");
    }
    const bool tiled_ok = (d->Cin % 8 == 0) && (d->Cin2 % 8 == 0) && (d->Cout == 1 || d->Cout % 8 == 0);
    if (!tiled_ok) {
        const long long n = (long long)p.B * p.Do * p.Ho * p.Wo * p.Cout;
        MVSB200_REQUIRE(n < (1ll << 31) * 256, "conv3d: volume too large");
        k2_conv3d_naive_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, d->stride, d->transposed);
        return check_launch("k2_conv3d_naive_kernel");
    }
    if (d->transposed) return d->Cout == 1 ? launch_tiled<MODE_DECONV, 1>(p, st) : launch_tiled<MODE_DECONV, 8>(p, st);
    if (d->stride == 2) return d->Cout == 1 ? launch_tiled<MODE_S2, 1>(p, st) : launch_tiled<MODE_S2, 8>(p, st);
    return d->Cout == 1 ? launch_tiled<MODE_S1, 1>(p, st) : launch_tiled<MODE_S1, 8>(p, st);
}
