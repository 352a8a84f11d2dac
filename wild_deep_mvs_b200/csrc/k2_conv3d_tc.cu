// K2 on the 5th-generation tensor cores: 3x3x3 convolution / stride-2 convolution / stride-2 transposed
// convolution over channels-last volumes as an implicit GEMM issued with tcgen05.mma (kind::tf32), accumulators
// in TMEM, BN scale+bias / ReLU / skip-add fused into the TMEM->register epilogue.
//
// The input halo block of a tile is staged ONCE in shared memory as a dense local grid, linear index
// l = (lz*EY + ly)*EX + lx, in the SWIZZLE_NONE K-major canonical layout [k-chunk][row][16 bytes] (umma.cuh).
// GEMM rows use the SAME pitches, so the A operand of a filter tap is just a window of 128 staged rows starting
// at a shifted row: a tap is a descriptor start address, never an im2col copy.
//
// With Cout as small as 8 the MMA is limited by the tensor core's shared-memory reads of A (4 KB per 128x8 tile),
// not by math, so the kernel minimises A reads per voxel:
//   * the (dz,dy) taps are row-shifted windows (9 per 8-channel chunk), but the three dx taps are folded into the
//     N dimension: P[row, dx, co] = sum_{dz,dy,ci} A[row + (dz*EY+dy)*EX, ci] * w[dz,dy,dx,ci,co], and the epilogue
//     forms out[c] = P[c,0] + P[c+1,1] + P[c+2,2] by exchanging neighbour rows through shared memory;
//   * the 3xTF32 split needs a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (+ a_lo*w_lo): w_hi and w_lo are also concatenated
//     on N, so the four products cost two A reads (a_hi, a_lo) instead of three, and the epilogue adds the two
//     column halves.
// Columns of one accumulator: ((xshift * NHL) + hl) * CT + co.
//
// Stride 2 stages the eight parity sub-grids of the input one after the other (space-to-depth): inside a parity
// class a stride-2 tap is again a unit-stride window.  The transposed conv (k3, s2, p1, op1) is evaluated in gather
// form per output parity class (one class per blockIdx.z): 1/2/4/8 taps each.
//
// Precision: the reference is fp32.  Default (NPROD = 3) is the error-compensated split above (fp32-equivalent);
// NPROD = 1 is plain single-pass TF32.
#include "common.cuh"
#include "umma.cuh"

namespace mvsb200 {

using namespace umma;

constexpr int TC_THREADS = 256;
constexpr int TC_S1 = 0, TC_S2 = 1, TC_DECONV = 2;

__host__ __device__ constexpr int tc_ncols(int mode, int ct, int nprod)
{
    return (((mode == TC_S1 ? 3 : 2) * (nprod == 3 ? 2 : 1) * ct) + 15) / 16 * 16;
}

template <int MODE, int CT, int NPROD> struct TcTile {
    static constexpr int NXS = (MODE == TC_S1) ? 3 : 2;     // x shifts folded into N
    static constexpr int NHL = (NPROD == 3) ? 2 : 1;
    static constexpr int NCOLS = tc_ncols(MODE, CT, NPROD);  // accumulator columns per 128-row tile (MMA N)
    static constexpr int TZ = (NCOLS * 4 <= 256) ? 2 : 1, TY = 7, TX = 32;
    static constexpr int HALO = (MODE == TC_S1) ? 2 : 1;
    static constexpr int EZ = TZ + HALO, EY = TY + HALO, EX = TX + HALO;
    static constexpr int NSUB = (MODE == TC_S2) ? 8 : 1;
    static constexpr int P = EY * EX;                       // plane pitch (rows)
    static constexpr int SUBR = EZ * P;                     // staged rows per sub-block
    static constexpr int ROWS_PLANE = (TY - 1) * EX + TX + (NXS - 1);  // rows of a plane some output depends on
    static constexpr int MT_PLANE = (ROWS_PLANE + 127) / 128;
    static constexpr int NMT = TZ * MT_PLANE;               // 128-row MMA tiles per CTA
    static constexpr int MAXSHIFT = HALO * (P + EX);        // largest (z,y) window shift
    static constexpr int R_NEED = (TZ - 1) * P + MT_PLANE * 128 + MAXSHIFT;
    // +4: the two k-chunk planes land 64 bytes apart modulo 128, so the paired staging stores are conflict free
    static constexpr int R_ALLOC = ((R_NEED > SUBR ? R_NEED : SUBR) + 7) / 8 * 8 + 4;
    static constexpr int MAX_WIN = (MODE == TC_S1) ? 9 : 4;  // (z,y) windows per stage
    static constexpr int TMEM_COLS = (NMT * NCOLS <= 32) ? 32 : (NMT * NCOLS <= 64) ? 64 : (NMT * NCOLS <= 128) ? 128 : (NMT * NCOLS <= 256) ? 256 : 512;
    static constexpr int XROWS = NMT * 128 + 8;              // rows of the epilogue exchange buffer
    static constexpr int A_F4 = NHL * 2 * R_ALLOC, B_F4 = MAX_WIN * 2 * NCOLS, X_F4 = (NXS - 1) * (CT / 4) * XROWS;
    static constexpr int SMEM_F4 = (A_F4 + B_F4 > X_F4) ? A_F4 + B_F4 : X_F4;
    static_assert(NMT * NCOLS <= 512, "accumulators exceed TMEM");
    static_assert(MT_PLANE * 128 <= P + 128, "row tiles of a plane must not run far into the next plane");
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
__host__ __device__ constexpr int windows_of(int mode, int v3) { return dim_opts(mode, (v3 >> 2) & 1) * dim_opts(mode, (v3 >> 1) & 1); }
__host__ __device__ constexpr int windows_per_chunk(int mode) { return mode == TC_S1 ? 9 : 18; }
// number of window blocks before (class cl, chunk c, sub s) in the packed weight buffer of one N block
__host__ __device__ constexpr int pack_block_offset(int mode, int cl, int c, int s, int nch)
{
    if (mode == TC_S1) return c * 9;
    if (mode == TC_S2) {
        int pre = 0;
        for (int i = 0; i < s; i++) pre += windows_of(mode, i);
        return c * 18 + pre;
    }
    int pre = 0;
    for (int i = 0; i < cl; i++) pre += windows_of(mode, i);
    return pre * nch + c * windows_of(mode, cl);
}

struct TcParams {
    const float *x, *x2, *wp, *scale, *bias, *skip;
    float *y;
    int B, D, H, W, Do, Ho, Wo;
    int Cin1, Cin2, Cout;
    int relu, skip_mode;
    int tiles_x, tiles_y, tiles_z;
};

template <int MODE, int CT, int NPROD>
__global__ void __launch_bounds__(TC_THREADS, (TcTile<MODE, CT, NPROD>::TMEM_COLS > 256) ? 1 : 2) k2_conv3d_tc_kernel(const TcParams p)
{
    using T = TcTile<MODE, CT, NPROD>;
    constexpr int NHL = T::NHL, RA = T::R_ALLOC, NC = T::NCOLS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *sA = reinterpret_cast<float4 *>(smem_raw);   // [hl][k-chunk][RA]
    float4 *sB = sA + T::A_F4;                            // [window][k-chunk][NC]
    float4 *sX = reinterpret_cast<float4 *>(smem_raw);   // epilogue exchange, overlays A/B once the MMAs are done
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
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), T::TMEM_COLS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = s_tmem;

    const float *wp_nb = p.wp + (size_t)nb * windows_per_chunk(MODE) * nch * (8 * NC);
    uint32_t phase = 0;

    // Software pipeline: the global loads of stage st+1 are issued into registers right after the MMAs of stage st
    // have been launched, so their latency overlaps the tensor-core work; the registers are split into tf32 hi/lo
    // and stored to shared memory once the MMAs (which read that memory) have completed.
    // Lane pairs share a voxel: even lanes fetch channels [0,4) of the chunk, odd lanes [4,8), so a warp-level
    // LDG.128 covers 16 voxels x 32 contiguous bytes (one sector each).
    constexpr int NPF = (T::SUBR + TC_THREADS / 2 - 1) / (TC_THREADS / 2);   // staged voxels per lane pair
    constexpr int NPB = (T::MAX_WIN * 2 * NC + TC_THREADS - 1) / TC_THREADS;  // weight float4s per thread
    const int half = tid & 1;
    float4 preA[NPF], preB[NPB];
    const int nstages = nch * T::NSUB;

    auto prefetch = [&](int st) {
        const int c = st / T::NSUB, s = st % T::NSUB;
        const float *src;
        int cs, cstride;
        if (c * 8 < p.Cin1) { src = p.x; cs = c * 8; cstride = p.Cin1; }
        else { src = p.x2; cs = c * 8 - p.Cin1; cstride = p.Cin2; }
        const int pz = (s >> 2) & 1, py = (s >> 1) & 1, px = s & 1;
#pragma unroll
        for (int k = 0; k < NPF; k++) {
            const int i = (tid >> 1) + k * (TC_THREADS / 2);
            const int lx = i % T::EX, r = i / T::EX;
            const int ly = r % T::EY, lz = r / T::EY;
            int gz, gy, gx;
            if (MODE == TC_S1) { gz = z0 - 1 + lz; gy = y0 - 1 + ly; gx = x0 - 1 + lx; }
            else if (MODE == TC_S2) { gz = 2 * (z0 + lz) - 1 + pz; gy = 2 * (y0 + ly) - 1 + py; gx = 2 * (x0 + lx) - 1 + px; }
            else { gz = z0 + lz; gy = y0 + ly; gx = x0 + lx; }
            preA[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < T::SUBR && (unsigned)gz < (unsigned)p.D && (unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W)
                preA[k] = ldg4(src + ((((long long)b * p.D + gz) * p.H + gy) * p.W + gx) * cstride + cs + half * 4);
        }
        // weights of this (class, chunk, sub-grid): contiguous in the packed buffer
        const int nwin = windows_of(MODE, (MODE == TC_S2) ? s : cl);
        const float4 *wsrc = reinterpret_cast<const float4 *>(wp_nb) + (size_t)pack_block_offset(MODE, cl, c, s, nch) * (2 * NC);
#pragma unroll
        for (int k = 0; k < NPB; k++) {
            const int i = tid + k * TC_THREADS;
            if (i < nwin * 2 * NC) preB[k] = __ldg(wsrc + i);
        }
    };

    prefetch(0);
#pragma unroll 1
    for (int st = 0; st < nstages; st++) {
        const int s = st % T::NSUB;
        if (st > 0) {   // the previous stage's MMAs still read sA / sB
            mbar_wait(bar, phase);
            phase ^= 1;
        }
        // ---- registers -> hi/lo tf32 planes in shared memory ----
#pragma unroll
        for (int k = 0; k < NPF; k++) {
            const int i = (tid >> 1) + k * (TC_THREADS / 2);
            if (i < T::SUBR) {
                const float4 v = preA[k];
                if (NPROD == 3) {
                    float4 h, l;
                    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
                    sA[half * RA + i] = h; sA[(2 + half) * RA + i] = l;
                } else {
                    sA[half * RA + i] = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
                }
            }
        }
        const int v3 = (MODE == TC_S2) ? s : cl;
        const int vz = (v3 >> 2) & 1, vy = (v3 >> 1) & 1;
        const int nz = dim_opts(MODE, vz), ny = dim_opts(MODE, vy);
#pragma unroll
        for (int k = 0; k < NPB; k++) {
            const int i = tid + k * TC_THREADS;
            if (i < nz * ny * 2 * NC) sB[i] = preB[k];
        }
        fence_proxy_async_smem();
        __syncthreads();
        // ---- one thread issues every MMA of the stage ----
        if (warp == 0) {
            if (lane == 0) {
                tc_fence_after_sync();
                constexpr uint32_t idesc = idesc_tf32(128, NC);
                const uint64_t adesc = smem_desc(smem_u32(sA), RA * 16, 128);
                const uint64_t bdesc = smem_desc(smem_u32(sB), NC * 16, 128);
                int win = 0;
#pragma unroll 1
                for (int jz = 0; jz < nz; jz++)
#pragma unroll 1
                    for (int jy = 0; jy < ny; jy++, win++) {
                        const int shift = (dim_shift(MODE, vz, jz) * T::EY + dim_shift(MODE, vy, jy)) * T::EX;
                        const uint64_t bd = bdesc + (uint64_t)(win * 2 * NC);
                        const uint32_t acc0 = (st == 0 && win == 0) ? 0u : 1u;
#pragma unroll
                        for (int mt = 0; mt < T::NMT; mt++) {
                            const int row0 = (mt / T::MT_PLANE) * T::P + (mt % T::MT_PLANE) * 128 + shift;
                            const uint64_t a_hi = adesc + (uint64_t)row0;
                            const uint32_t d = tmem + mt * NC;
                            mma_tf32(d, a_hi, bd, idesc, acc0);
                            if (NPROD == 3) mma_tf32(d, a_hi + 2 * RA, bd, idesc, 1u);
                        }
                    }
                mma_commit(bar);
            }
            __syncwarp();
        }
        if (st + 1 < nstages) prefetch(st + 1);
    }
    mbar_wait(bar, phase);
    tc_fence_after_sync();

    // ---- epilogue 1: TMEM -> registers, add the hi/lo column halves, keep the x-shift-0 partials, publish the
    //      shifted ones to shared memory (sX overlays the operand buffers: every MMA has completed) ----
    constexpr int MPT = (T::NMT + 1) / 2;   // row tiles per thread (warps 0-3 take even tiles, 4-7 odd ones)
    constexpr int C4 = CT / 4;
    const int wq = warp & 3;
    float q0[MPT][CT];
#pragma unroll
    for (int m = 0; m < MPT; m++) {
        const int mt = (warp >> 2) + 2 * m;
        if (mt < T::NMT) {   // warp-uniform
            const int xrow = mt * 128 + wq * 32 + lane;
#pragma unroll
            for (int xs = 0; xs < T::NXS; xs++) {
                float v[NHL * CT];
                const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + mt * NC + xs * NHL * CT;
#pragma unroll
                for (int j = 0; j < NHL * CT; j += 8) tmem_ld8(taddr + j, v + j);
                tmem_ld_wait();
                if (NHL == 2) {
#pragma unroll
                    for (int k = 0; k < CT; k++) v[k] += v[CT + k];
                }
                if (xs == 0) {
#pragma unroll
                    for (int k = 0; k < CT; k++) q0[m][k] = v[k];
                } else {
#pragma unroll
                    for (int c4 = 0; c4 < C4; c4++)
                        sX[((xs - 1) * C4 + c4) * T::XROWS + xrow] = make_float4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();

    // ---- epilogue 2: out[c] = P[c, 0] + P[c+1, 1] (+ P[c+2, 2]);  y = act(out*scale + bias [+ skip]) [+ skip] ----
    const int ncol = min(CT, p.Cout - nb * CT);   // real output channels of this N block (1, 8, 16 or 32)
    const int co0 = nb * CT;
#pragma unroll
    for (int m = 0; m < MPT; m++) {
        const int mt = (warp >> 2) + 2 * m;
        if (mt >= T::NMT) continue;
        const int pr = (mt % T::MT_PLANE) * 128 + wq * 32 + lane;
        const int xrow = mt * 128 + wq * 32 + lane;
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
#pragma unroll
        for (int c4 = 0; c4 < C4; c4++) {
            if (c4 * 4 >= ncol) break;
            float r[4] = {q0[m][c4 * 4], q0[m][c4 * 4 + 1], q0[m][c4 * 4 + 2], q0[m][c4 * 4 + 3]};
#pragma unroll
            for (int xs = 1; xs < T::NXS; xs++) {
                const float4 nbv = sX[((xs - 1) * C4 + c4) * T::XROWS + xrow + xs];
                r[0] += nbv.x; r[1] += nbv.y; r[2] += nbv.z; r[3] += nbv.w;
            }
            if (ncol == 1) {
                float y1 = r[0] * (p.scale ? __ldg(p.scale) : 1.f) + (p.bias ? __ldg(p.bias) : 0.f);
                if (p.skip_mode == MVSB200_SKIP_BEFORE_RELU) y1 += __ldg(p.skip + o);
                if (p.relu) y1 = fmaxf(y1, 0.f);
                if (p.skip_mode == MVSB200_SKIP_AFTER_RELU) y1 += __ldg(p.skip + o);
                p.y[o] = y1;
                break;
            }
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
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, T::TMEM_COLS);
}

// ---- weight packing: tap-major [27][Cin][Cout] -> per (N block, class, chunk, sub, (z,y) window) blocks of
//      [k-chunk][NCOLS][4] in the canonical B layout; column = ((xshift * NHL) + hl) * CT + co, zero padded ----
struct TcPackParams {
    const float *w;
    float *wp;
    int Cin, Cout, nch, mode, CT, NC, nhl;
};

__global__ void k2_tc_pack_kernel(const TcPackParams p)
{
    // one block per (nb, cl, c, s); threads over (window, q, column, e)
    int id = blockIdx.x;
    const int nsub = (p.mode == TC_S2) ? 8 : 1, ncl = (p.mode == TC_DECONV) ? 8 : 1;
    const int s = id % nsub; id /= nsub;
    const int c = id % p.nch; id /= p.nch;
    const int cl = id % ncl;
    const int nb = id / ncl;
    const int v3 = (p.mode == TC_S2) ? s : cl;
    const int vz = (v3 >> 2) & 1, vy = (v3 >> 1) & 1, vx = v3 & 1;
    const int ny = dim_opts(p.mode, vy), nx = dim_opts(p.mode, vx);
    const int nwin = windows_of(p.mode, v3);
    const int blk = 8 * p.NC;   // floats per window block
    float *dst = p.wp + ((size_t)nb * windows_per_chunk(p.mode) * p.nch + pack_block_offset(p.mode, cl, c, s, p.nch)) * blk;
    for (int i = threadIdx.x; i < nwin * blk; i += blockDim.x) {
        const int e = i & 3;
        int r = i >> 2;
        const int col = r % p.NC; r /= p.NC;
        const int q = r & 1;
        const int win = r >> 1;
        const int jy = win % ny, jz = win / ny;
        const int n = col % p.CT, g = col / p.CT;    // g = xshift * nhl + hl
        const int hl = g % p.nhl, xs = g / p.nhl;
        float val = 0.f;
        int jx = -1;
        for (int j = 0; j < nx; j++)
            if (dim_shift(p.mode, vx, j) == xs) jx = j;
        const int ci = c * 8 + q * 4 + e, co = nb * p.CT + n;
        if (jx >= 0 && co < p.Cout) {
            const int k = (dim_k(p.mode, vz, jz) * 3 + dim_k(p.mode, vy, jy)) * 3 + dim_k(p.mode, vx, jx);
            val = p.w[((size_t)k * p.Cin + ci) * p.Cout + co];
        }
        float hi, lo;
        split_tf32(val, hi, lo);
        dst[i] = (p.nhl == 2) ? (hl ? lo : hi) : hi;
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
static int tc_ct(const mvsb200_conv3d_desc *d) { return d->Cout <= 8 ? 8 : (d->Cout <= 16 ? 16 : 32); }

template <int MODE, int CT, int NPROD>
static int launch_tc(TcParams p, cudaStream_t st)
{
    using T = TcTile<MODE, CT, NPROD>;
    const size_t smem = (size_t)T::SMEM_F4 * sizeof(float4);
    const int nz = (MODE == TC_DECONV) ? p.D : p.Do, ny = (MODE == TC_DECONV) ? p.H : p.Ho, nx = (MODE == TC_DECONV) ? p.W : p.Wo;
    p.tiles_z = (nz + T::TZ - 1) / T::TZ;
    p.tiles_y = (ny + T::TY - 1) / T::TY;
    p.tiles_x = (nx + T::TX - 1) / T::TX;
    const long long tiles = (long long)p.tiles_x * p.tiles_y * p.tiles_z * p.B;
    if (tiles >= (1ll << 31)) {
        set_error("conv3d_tc: volume too large");
        return MVSB200_E_INVALID;
    }
    if (int rc = ensure_dynamic_smem(k2_conv3d_tc_kernel<MODE, CT, NPROD>, smem, "conv3d_tc")) return rc;
    dim3 grid((unsigned)tiles, (unsigned)((p.Cout + CT - 1) / CT), (MODE == TC_DECONV) ? 8 : 1);
    k2_conv3d_tc_kernel<MODE, CT, NPROD><<<grid, TC_THREADS, smem, st>>>(p);
    return check_launch("k2_conv3d_tc_kernel");
}

template <int MODE>
static int launch_tc_mode(const TcParams &p, int ct, int nprod, cudaStream_t st)
{
    if (ct == 8) return nprod == 3 ? launch_tc<MODE, 8, 3>(p, st) : launch_tc<MODE, 8, 1>(p, st);
    if (ct == 16) return nprod == 3 ? launch_tc<MODE, 16, 3>(p, st) : launch_tc<MODE, 16, 1>(p, st);
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
    const int ct = tc_ct(d), nblocks = (d->Cout + ct - 1) / ct, nch = (d->Cin + d->Cin2) / 8, mode = tc_mode(d);
    // both precisions are packed behind each other: [3xTF32 layout][TF32 layout]
    return (long long)nblocks * windows_per_chunk(mode) * nch * 8 * (tc_ncols(mode, ct, 3) + tc_ncols(mode, ct, 1));
}

extern "C" int mvsb200_conv3d_tc_pack(const mvsb200_conv3d_desc *d, const float *w, float *packed, mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(d && w && packed, "conv3d_tc_pack: null pointer");
    MVSB200_REQUIRE(tc_shape_ok(d), "conv3d_tc_pack: layer not supported by the tensor-core engine (need k=3, Cin%%8==0, Cout in {1,8,16,32k})");
    TcPackParams p;
    p.w = w;
    p.Cin = d->Cin + d->Cin2; p.Cout = d->Cout; p.nch = p.Cin / 8;
    p.CT = tc_ct(d); p.mode = tc_mode(d);
    const int nblocks = (d->Cout + p.CT - 1) / p.CT;
    const int nsub = (p.mode == TC_S2) ? 8 : 1, ncl = (p.mode == TC_DECONV) ? 8 : 1;
    float *dst = packed;
    for (int nprod = 3; nprod >= 1; nprod -= 2) {
        p.wp = dst; p.nhl = nprod == 3 ? 2 : 1; p.NC = tc_ncols(p.mode, p.CT, nprod);
        k2_tc_pack_kernel<<<nblocks * ncl * p.nch * nsub, 256, 0, (cudaStream_t)stream>>>(p);
        dst += (size_t)nblocks * windows_per_chunk(p.mode) * p.nch * 8 * p.NC;
    }
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
    const int mode = tc_mode(d), ct = tc_ct(d);
    const float *wp = packed;
    if (precision == MVSB200_PRECISION_TF32)   // the single-pass layout sits behind the split one
        wp += (size_t)((d->Cout + ct - 1) / ct) * windows_per_chunk(mode) * ((d->Cin + d->Cin2) / 8) * 8 * tc_ncols(mode, ct, 3);
    p.x = x; p.x2 = x2; p.wp = wp; p.scale = scale; p.bias = bias; p.skip = skip; p.y = y;
    p.B = d->B; p.D = d->D; p.H = d->H; p.W = d->W;
    p.Cin1 = d->Cin; p.Cin2 = d->Cin2; p.Cout = d->Cout;
    p.relu = d->relu; p.skip_mode = d->skip_mode;
    p.tiles_x = p.tiles_y = p.tiles_z = 0;
    const int nprod = precision == MVSB200_PRECISION_3XTF32 ? 3 : 1;
    cudaStream_t st = (cudaStream_t)stream;
    switch (mode) {
    case TC_S1: return launch_tc_mode<TC_S1>(p, ct, nprod, st);
    case TC_S2: return launch_tc_mode<TC_S2>(p, ct, nprod, st);
    default: return launch_tc_mode<TC_DECONV>(p, ct, nprod, st);
    }
}
