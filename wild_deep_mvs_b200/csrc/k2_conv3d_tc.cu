// K2 on the 5th-generation tensor cores: 3x3x3 convolution / stride-2 convolution / stride-2 transposed
// convolution over channels-last volumes as an implicit GEMM issued with tcgen05.mma (kind::tf32), accumulators
// in TMEM, BN scale+bias / ReLU / skip-add fused into the TMEM->register epilogue.
//
// GEMM view: rows = voxels of an output tile, N = output channels (16 or 32 per CTA, zero padded), K = 8 input
// channels per MMA, one MMA per (filter tap, 8-channel chunk, 128-voxel row tile).
//
// The trick that makes the GEMM implicit: the input halo block of a tile is staged in shared memory as a dense
// local grid, linear index l = (lz*EY + ly)*EX + lx, in the SWIZZLE_NONE K-major canonical layout
// [k-chunk][row][16 bytes] (umma.cuh).  Output rows use the SAME pitches (r = (oz*EY + oy)*EX + ox), so the
// A operand of filter tap (dz,dy,dx) is the window of 128 staged rows starting at r0 + (dz*EY + dy)*EX + dx:
// every tap is just a different descriptor start address into one staged block -- no im2col, no re-staging.
// Rows whose (oy,ox) fall into the halo columns produce garbage accumulators that the epilogue never stores
// (each D row depends on its own A row only).
//
// Stride 2 is handled by staging the eight parity sub-grids of the input one after the other (space-to-depth):
// inside a parity class a stride-2 tap is again a unit-stride window.  The transposed conv (k3, s2, p1, op1) is
// evaluated in gather form per output parity class (one class per blockIdx.z): 1/2/4/8 taps each.
//
// Precision: the reference is fp32.  Default is 3xTF32: activations and weights are split into a tf32-exact high
// part and a remainder, D += a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (error ~2^-21 relative, fp32-equivalent);
// NPROD = 1 is the plain single-pass TF32 mode.
#include "common.cuh"
#include "umma.cuh"

namespace mvsb200 {

using namespace umma;

constexpr int TC_THREADS = 256;
constexpr int TC_S1 = 0, TC_S2 = 1, TC_DECONV = 2;

template <int MODE> struct TcTile {
    static constexpr int TZ = 2, TY = 7, TX = 32;
    static constexpr int HALO = (MODE == TC_S1) ? 2 : 1;
    static constexpr int EZ = TZ + HALO, EY = TY + HALO, EX = TX + HALO;
    static constexpr int NSUB = (MODE == TC_S2) ? 8 : 1;
    static constexpr int P = EY * EX;                       // plane pitch (rows)
    static constexpr int SUBR = EZ * P;                     // staged rows per sub-block
    static constexpr int ROWS_PLANE = (TY - 1) * EX + TX;   // rows of one output plane that can be valid
    static constexpr int MT_PLANE = (ROWS_PLANE + 127) / 128;
    static constexpr int NMT = TZ * MT_PLANE;               // 128-row MMA tiles per CTA
    static constexpr int MAXSHIFT = HALO * (P + EX + 1);
    static constexpr int R_NEED = (TZ - 1) * P + MT_PLANE * 128 + MAXSHIFT;
    static constexpr int R_ALLOC = ((R_NEED > SUBR ? R_NEED : SUBR) + 7) / 8 * 8;
    static constexpr int MAX_TAPS = (MODE == TC_S1) ? 27 : 8;  // taps per stage
};

// ---- per-dimension tap tables ---------------------------------------------------------------------------------
// variant v: S1 -> 0; S2 -> parity of the staged input sub-grid; DECONV -> parity of the output class.
// Option j of a variant = (filter index k, row shift in that dimension).
__host__ __device__ constexpr int dim_opts(int mode, int v) { return mode == TC_S1 ? 3 : (mode == TC_S2 ? (v == 0 ? 2 : 1) : (v == 0 ? 1 : 2)); }
__host__ __device__ constexpr int dim_k(int mode, int v, int j)
{
    return mode == TC_S1 ? j : (mode == TC_S2 ? (v == 0 ? 2 * j : 1) : (v == 0 ? 1 : (j == 0 ? 2 : 0)));
}
__host__ __device__ constexpr int dim_shift(int mode, int v, int j) { return mode == TC_S1 ? j : (mode == TC_S2 ? (v == 0 ? j : 0) : (v == 0 ? 0 : j)); }
__host__ __device__ constexpr int taps_of(int mode, int v3)
{
    return dim_opts(mode, (v3 >> 2) & 1) * dim_opts(mode, (v3 >> 1) & 1) * dim_opts(mode, v3 & 1);
}
// number of tap blocks before (class cl, chunk c, sub s) in the packed weight buffer of one N block
__host__ __device__ constexpr int pack_block_offset(int mode, int cl, int c, int s, int nch)
{
    if (mode == TC_S1) return c * 27;
    if (mode == TC_S2) {
        int pre = 0;
        for (int i = 0; i < s; i++) pre += taps_of(mode, i);
        return c * 27 + pre;
    }
    int pre = 0;
    for (int i = 0; i < cl; i++) pre += taps_of(mode, i);
    return pre * nch + c * taps_of(mode, cl);
}

struct TcParams {
    const float *x, *x2, *wp, *scale, *bias, *skip;
    float *y;
    int B, D, H, W, Do, Ho, Wo;
    int Cin1, Cin2, Cout;
    int relu, skip_mode;
    int tiles_x, tiles_y, tiles_z;
};

template <int MODE, int NT, int NPROD>
__global__ void __launch_bounds__(TC_THREADS, 2) k2_conv3d_tc_kernel(const TcParams p)
{
    using T = TcTile<MODE>;
    constexpr int NHL = (NPROD == 3) ? 2 : 1;
    constexpr int RA = T::R_ALLOC;
    constexpr int NCOLS = (T::NMT * NT <= 32) ? 32 : (T::NMT * NT <= 64) ? 64 : (T::NMT * NT <= 128) ? 128 : (T::NMT * NT <= 256) ? 256 : 512;
    static_assert(T::NMT * NT <= 512, "accumulators exceed TMEM");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *sA = reinterpret_cast<float4 *>(smem_raw);   // [hl][k-chunk][RA]
    float4 *sB = sA + NHL * 2 * RA;                       // [tap][hl(2)][k-chunk][NT]
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int t = blockIdx.x;
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y; t /= p.tiles_y;
    const int tz = t % p.tiles_z;
    const int b = t / p.tiles_z;
    const int nb = blockIdx.y;
    const int cl = (MODE == TC_DECONV) ? (int)blockIdx.z : 0;
    const int z0 = tz * T::TZ, y0 = ty * T::TY, x0 = tx * T::TX;   // tile origin (output coords; DECONV: input coords)
    const int nch = (p.Cin1 + p.Cin2) >> 3;
    const uint32_t bar = smem_u32(&s_bar);

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), NCOLS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = s_tmem;

    const float *wp_nb = p.wp + (size_t)nb * 27 * nch * (16 * NT);
    uint32_t phase = 0;
    bool first_stage = true;

    for (int c = 0; c < nch; c++) {
        const float *src;
        int cs, cstride;
        if (c * 8 < p.Cin1) { src = p.x; cs = c * 8; cstride = p.Cin1; }
        else { src = p.x2; cs = c * 8 - p.Cin1; cstride = p.Cin2; }
#pragma unroll 1
        for (int s = 0; s < T::NSUB; s++) {
            if (!first_stage) {   // the previous stage's MMAs still read sA / sB
                mbar_wait(bar, phase);
                phase ^= 1;
            }
            // ---- stage the input block: global (fp32, channels-last) -> hi/lo tf32 planes in shared memory ----
            const int pz = (s >> 2) & 1, py = (s >> 1) & 1, px = s & 1;
#pragma unroll 2
            for (int i = tid; i < T::SUBR; i += TC_THREADS) {
                const int lx = i % T::EX, r = i / T::EX;
                const int ly = r % T::EY, lz = r / T::EY;
                int gz, gy, gx;
                if (MODE == TC_S1) { gz = z0 - 1 + lz; gy = y0 - 1 + ly; gx = x0 - 1 + lx; }
                else if (MODE == TC_S2) { gz = 2 * (z0 + lz) - 1 + pz; gy = 2 * (y0 + ly) - 1 + py; gx = 2 * (x0 + lx) - 1 + px; }
                else { gz = z0 + lz; gy = y0 + ly; gx = x0 + lx; }
                float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                if ((unsigned)gz < (unsigned)p.D && (unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W) {
                    const float *q = src + ((((long long)b * p.D + gz) * p.H + gy) * p.W + gx) * cstride + cs;
                    v0 = ldg4(q);
                    v1 = ldg4(q + 4);
                }
                if (NPROD == 3) {
                    float4 h0, l0, h1, l1;
                    split_tf32(v0.x, h0.x, l0.x); split_tf32(v0.y, h0.y, l0.y); split_tf32(v0.z, h0.z, l0.z); split_tf32(v0.w, h0.w, l0.w);
                    split_tf32(v1.x, h1.x, l1.x); split_tf32(v1.y, h1.y, l1.y); split_tf32(v1.z, h1.z, l1.z); split_tf32(v1.w, h1.w, l1.w);
                    sA[i] = h0; sA[RA + i] = h1; sA[2 * RA + i] = l0; sA[3 * RA + i] = l1;
                } else {
                    sA[i] = make_float4(round_tf32(v0.x), round_tf32(v0.y), round_tf32(v0.z), round_tf32(v0.w));
                    sA[RA + i] = make_float4(round_tf32(v1.x), round_tf32(v1.y), round_tf32(v1.z), round_tf32(v1.w));
                }
            }
            // ---- stage the weights of this (class, chunk, sub-grid): contiguous in the packed buffer ----
            const int v3 = (MODE == TC_S2) ? s : cl;
            const int ntaps = taps_of(MODE, v3);
            {
                const float4 *wsrc = reinterpret_cast<const float4 *>(wp_nb) + (size_t)pack_block_offset(MODE, cl, c, s, nch) * (4 * NT);
                for (int i = tid; i < ntaps * 4 * NT; i += TC_THREADS) sB[i] = __ldg(wsrc + i);
            }
            fence_proxy_async_smem();
            __syncthreads();
            // ---- one thread issues every MMA of the stage ----
            if (warp == 0) {
                if (lane == 0) {
                    tc_fence_after_sync();
                    constexpr uint32_t idesc = idesc_tf32(128, NT);
                    const uint64_t adesc = smem_desc(smem_u32(sA), RA * 16, 128);
                    const uint64_t bdesc = smem_desc(smem_u32(sB), NT * 16, 128);
                    const int vz = (v3 >> 2) & 1, vy = (v3 >> 1) & 1, vx = v3 & 1;
                    const int nz = dim_opts(MODE, vz), ny = dim_opts(MODE, vy), nx = dim_opts(MODE, vx);
                    int tt = 0;
#pragma unroll 1
                    for (int jz = 0; jz < nz; jz++)
#pragma unroll 1
                        for (int jy = 0; jy < ny; jy++)
#pragma unroll 1
                            for (int jx = 0; jx < nx; jx++, tt++) {
                                const int shift = (dim_shift(MODE, vz, jz) * T::EY + dim_shift(MODE, vy, jy)) * T::EX + dim_shift(MODE, vx, jx);
                                const uint64_t b_hi = bdesc + (uint64_t)(tt * 4 * NT), b_lo = b_hi + 2 * NT;
                                const uint32_t acc0 = (first_stage && tt == 0) ? 0u : 1u;
#pragma unroll
                                for (int mt = 0; mt < T::NMT; mt++) {
                                    const int row0 = (mt / T::MT_PLANE) * T::P + (mt % T::MT_PLANE) * 128 + shift;
                                    const uint64_t a_hi = adesc + (uint64_t)row0, a_lo = a_hi + 2 * RA;
                                    const uint32_t d = tmem + mt * NT;
                                    if (NPROD == 3) {
                                        mma_tf32(d, a_lo, b_hi, idesc, acc0);
                                        mma_tf32(d, a_hi, b_lo, idesc, 1u);
                                        mma_tf32(d, a_hi, b_hi, idesc, 1u);
                                    } else {
                                        mma_tf32(d, a_hi, b_hi, idesc, acc0);
                                    }
                                }
                            }
                    mma_commit(bar);
                }
                __syncwarp();
            }
            first_stage = false;
        }
    }
    mbar_wait(bar, phase);
    tc_fence_after_sync();

    // ---- epilogue: TMEM -> registers -> y = act(acc*scale + bias [+ skip]) [+ skip] ----
    const int wq = warp & 3;
    const int ncol = min(NT, p.Cout - nb * NT);   // real output channels of this N block (1, 8, 16 or 32)
    const int co0 = nb * NT;
    for (int mt = warp >> 2; mt < T::NMT; mt += TC_THREADS / 128) {
        float v[NT];
        const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + mt * NT;
        if (ncol <= 8) tmem_ld8(taddr, v);
        else {
#pragma unroll
            for (int j = 0; j < NT; j += 16) tmem_ld16(taddr + j, v + j);
        }
        tmem_ld_wait();
        const int pr = (mt % T::MT_PLANE) * 128 + wq * 32 + lane;
        const int oy_l = pr / T::EX, ox_l = pr % T::EX;
        int oz = z0 + mt / T::MT_PLANE, oy = y0 + oy_l, ox = x0 + ox_l;
        bool ok = oy_l < T::TY && ox_l < T::TX;
        if (MODE == TC_DECONV) {
            ok = ok && oz < p.D && oy < p.H && ox < p.W;
            oz = 2 * oz + ((cl >> 2) & 1); oy = 2 * oy + ((cl >> 1) & 1); ox = 2 * ox + (cl & 1);
        }
        ok = ok && oz < p.Do && oy < p.Ho && ox < p.Wo;
        if (!ok) continue;
        const long long o = ((((long long)b * p.Do + oz) * p.Ho + oy) * p.Wo + ox) * p.Cout + co0;
        if (ncol == 1) {
            float r = v[0] * (p.scale ? __ldg(p.scale) : 1.f) + (p.bias ? __ldg(p.bias) : 0.f);
            if (p.skip_mode == MVSB200_SKIP_BEFORE_RELU) r += __ldg(p.skip + o);
            if (p.relu) r = fmaxf(r, 0.f);
            if (p.skip_mode == MVSB200_SKIP_AFTER_RELU) r += __ldg(p.skip + o);
            p.y[o] = r;
            continue;
        }
#pragma unroll
        for (int c4 = 0; c4 < NT / 4; c4++) {
            if (c4 * 4 >= ncol) break;
            float r[4] = {v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]};
            if (p.scale) {
                const float4 sc = ldg4(p.scale + co0 + c4 * 4);
                r[0] *= sc.x; r[1] *= sc.y; r[2] *= sc.z; r[3] *= sc.w;
            }
            if (p.bias) {
                const float4 bi = ldg4(p.bias + co0 + c4 * 4);
                r[0] += bi.x; r[1] += bi.y; r[2] += bi.z; r[3] += bi.w;
            }
            float4 sk = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.skip_mode != MVSB200_SKIP_NONE) sk = ldg4(p.skip + o + c4 * 4);
            if (p.skip_mode == MVSB200_SKIP_BEFORE_RELU) { r[0] += sk.x; r[1] += sk.y; r[2] += sk.z; r[3] += sk.w; }
            if (p.relu) { r[0] = fmaxf(r[0], 0.f); r[1] = fmaxf(r[1], 0.f); r[2] = fmaxf(r[2], 0.f); r[3] = fmaxf(r[3], 0.f); }
            if (p.skip_mode == MVSB200_SKIP_AFTER_RELU) { r[0] += sk.x; r[1] += sk.y; r[2] += sk.z; r[3] += sk.w; }
            st4(p.y + o + c4 * 4, make_float4(r[0], r[1], r[2], r[3]));
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, NCOLS);
}

// ---- weight packing: tap-major [27][Cin][Cout] -> per (N block, class, chunk, sub, tap) blocks of
//      [hi/lo][k-chunk][NT][4] in the canonical B layout, zero padded to NT output channels ----
struct TcPackParams {
    const float *w;
    float *wp;
    int Cin, Cout, nch, nblocks, mode, NT;
};

__global__ void k2_tc_pack_kernel(const TcPackParams p)
{
    // one block per (nb, cl, c, s); threads over (tap, hl, q, n, e)
    int id = blockIdx.x;
    const int nsub = (p.mode == TC_S2) ? 8 : 1, ncl = (p.mode == TC_DECONV) ? 8 : 1;
    const int s = id % nsub; id /= nsub;
    const int c = id % p.nch; id /= p.nch;
    const int cl = id % ncl;
    const int nb = id / ncl;
    const int v3 = (p.mode == TC_S2) ? s : cl;
    const int vz = (v3 >> 2) & 1, vy = (v3 >> 1) & 1, vx = v3 & 1;
    const int ny = dim_opts(p.mode, vy), nx = dim_opts(p.mode, vx);
    const int ntaps = taps_of(p.mode, v3);
    const int blk = 16 * p.NT;   // floats per tap block
    float *dst = p.wp + ((size_t)nb * 27 * p.nch + pack_block_offset(p.mode, cl, c, s, p.nch)) * blk;
    for (int i = threadIdx.x; i < ntaps * blk; i += blockDim.x) {
        const int e = i & 3;
        int r = i >> 2;
        const int n = r % p.NT; r /= p.NT;
        const int q = r & 1; r >>= 1;
        const int hl = r & 1;
        const int tt = r >> 1;
        const int jx = tt % nx, jy = (tt / nx) % ny, jz = tt / (nx * ny);
        const int k = (dim_k(p.mode, vz, jz) * 3 + dim_k(p.mode, vy, jy)) * 3 + dim_k(p.mode, vx, jx);
        const int ci = c * 8 + q * 4 + e, co = nb * p.NT + n;
        float val = 0.f;
        if (co < p.Cout) val = p.w[((size_t)k * p.Cin + ci) * p.Cout + co];
        float hi, lo;
        split_tf32(val, hi, lo);
        dst[i] = hl ? lo : hi;
    }
}

static bool tc_shape_ok(const mvsb200_conv3d_desc *d)
{
    const int cin = d->Cin + d->Cin2;
    return d->kd == 3 && d->kh == 3 && d->kw == 3 && d->Cin % 8 == 0 && d->Cin2 % 8 == 0 && cin >= 8 && cin <= 128 &&
           (d->Cout == 1 || d->Cout == 8 || d->Cout == 16 || d->Cout % 32 == 0) && d->Cout <= 128 &&
           (d->stride == 1 || d->stride == 2) && (!d->transposed || d->stride == 2);
}

static int tc_mode(const mvsb200_conv3d_desc *d) { return d->transposed ? TC_DECONV : (d->stride == 2 ? TC_S2 : TC_S1); }
static int tc_nt(const mvsb200_conv3d_desc *d) { return d->Cout <= 16 ? 16 : 32; }

template <int MODE, int NT, int NPROD>
static int launch_tc(TcParams p, cudaStream_t st)
{
    using T = TcTile<MODE>;
    constexpr int NHL = (NPROD == 3) ? 2 : 1;
    const size_t smem = (size_t)(NHL * 2 * T::R_ALLOC + T::MAX_TAPS * 4 * NT) * sizeof(float4);
    const int nz = (MODE == TC_DECONV) ? p.D : p.Do, ny = (MODE == TC_DECONV) ? p.H : p.Ho, nx = (MODE == TC_DECONV) ? p.W : p.Wo;
    p.tiles_z = (nz + T::TZ - 1) / T::TZ;
    p.tiles_y = (ny + T::TY - 1) / T::TY;
    p.tiles_x = (nx + T::TX - 1) / T::TX;
    const long long tiles = (long long)p.tiles_x * p.tiles_y * p.tiles_z * p.B;
    if (tiles >= (1ll << 31)) {
        set_error("conv3d_tc: volume too large");
        return MVSB200_E_INVALID;
    }
    static bool attr_set = false;  // per template instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k2_conv3d_tc_kernel<MODE, NT, NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("conv3d_tc: cudaFuncSetAttribute(%zu bytes): %s", smem, cudaGetErrorString(e));
            return MVSB200_E_CUDA;
        }
        attr_set = true;
    }
    dim3 grid((unsigned)tiles, (unsigned)((p.Cout + NT - 1) / NT), (MODE == TC_DECONV) ? 8 : 1);
    k2_conv3d_tc_kernel<MODE, NT, NPROD><<<grid, TC_THREADS, smem, st>>>(p);
    return check_launch("k2_conv3d_tc_kernel");
}

template <int MODE>
static int launch_tc_mode(const TcParams &p, int nt, int nprod, cudaStream_t st)
{
    if (nt == 16) return nprod == 3 ? launch_tc<MODE, 16, 3>(p, st) : launch_tc<MODE, 16, 1>(p, st);
    return nprod == 3 ? launch_tc<MODE, 32, 3>(p, st) : launch_tc<MODE, 32, 1>(p, st);
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_conv3d_tc_supported(const mvsb200_conv3d_desc *d)
{
    return d && tc_shape_ok(d) ? 1 : 0;
}

extern "C" long long mvsb200_conv3d_tc_packed_floats(const mvsb200_conv3d_desc *d)
{
    if (!d || !tc_shape_ok(d)) return 0;
    const int nt = tc_nt(d), nblocks = (d->Cout + nt - 1) / nt, nch = (d->Cin + d->Cin2) / 8;
    return (long long)nblocks * 27 * nch * 16 * nt;
}

extern "C" int mvsb200_conv3d_tc_pack(const mvsb200_conv3d_desc *d, const float *w, float *packed, mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(d && w && packed, "conv3d_tc_pack: null pointer");
    MVSB200_REQUIRE(tc_shape_ok(d), "conv3d_tc_pack: layer not supported by the tensor-core engine (need k=3, Cin%%8==0, Cout in {1,8,16,32k})");
    TcPackParams p;
    p.w = w; p.wp = packed;
    p.Cin = d->Cin + d->Cin2; p.Cout = d->Cout; p.nch = p.Cin / 8;
    p.NT = tc_nt(d); p.nblocks = (d->Cout + p.NT - 1) / p.NT; p.mode = tc_mode(d);
    const int nsub = (p.mode == TC_S2) ? 8 : 1, ncl = (p.mode == TC_DECONV) ? 8 : 1;
    k2_tc_pack_kernel<<<p.nblocks * ncl * p.nch * nsub, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("k2_tc_pack_kernel");
}

extern "C" int mvsb200_conv3d_tc(const mvsb200_conv3d_desc *d, const float *x, const float *x2, const float *packed,
                                 const float *scale, const float *bias, const float *skip, float *y, int precision,
                                 mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(d && x && packed && y, "conv3d_tc: null pointer");
    MVSB200_REQUIRE(tc_shape_ok(d), "conv3d_tc: layer not supported by the tensor-core engine (need k=3, Cin%%8==0, Cout in {1,8,16,32k})");
    MVSB200_REQUIRE(d->B > 0 && d->D > 0 && d->H > 0 && d->W > 0, "conv3d_tc: bad shape B=%d D=%d H=%d W=%d", d->B, d->D, d->H, d->W);
    MVSB200_REQUIRE(d->Cin2 == 0 || x2, "conv3d_tc: Cin2=%d but x2 is null", d->Cin2);
    MVSB200_REQUIRE(d->skip_mode >= 0 && d->skip_mode <= 2, "conv3d_tc: skip_mode=%d", d->skip_mode);
    MVSB200_REQUIRE(d->skip_mode == MVSB200_SKIP_NONE || skip, "conv3d_tc: skip_mode=%d but skip is null", d->skip_mode);
    MVSB200_REQUIRE(precision == MVSB200_PRECISION_3XTF32 || precision == MVSB200_PRECISION_TF32, "conv3d_tc: precision=%d", precision);
    TcParams p;
    int rc = mvsb200_conv3d_out_shape(d, &p.Do, &p.Ho, &p.Wo);
    if (rc) return rc;
    MVSB200_REQUIRE(p.Do > 0 && p.Ho > 0 && p.Wo > 0, "conv3d_tc: empty output");
    p.x = x; p.x2 = x2; p.wp = packed; p.scale = scale; p.bias = bias; p.skip = skip; p.y = y;
    p.B = d->B; p.D = d->D; p.H = d->H; p.W = d->W;
    p.Cin1 = d->Cin; p.Cin2 = d->Cin2; p.Cout = d->Cout;
    p.relu = d->relu; p.skip_mode = d->skip_mode;
    p.tiles_x = p.tiles_y = p.tiles_z = 0;
    const int nprod = precision == MVSB200_PRECISION_3XTF32 ? 3 : 1;
    cudaStream_t st = (cudaStream_t)stream;
    switch (tc_mode(d)) {
    case TC_S1: return launch_tc_mode<TC_S1>(p, tc_nt(d), nprod, st);
    case TC_S2: return launch_tc_mode<TC_S2>(p, tc_nt(d), nprod, st);
    default: return launch_tc_mode<TC_DECONV>(p, tc_nt(d), nprod, st);
    }
}
