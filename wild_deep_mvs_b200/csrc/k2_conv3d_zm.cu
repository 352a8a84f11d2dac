// K2 "z-march" engine: 3x3x3 convolution / stride-2 convolution / stride-2 transposed convolution over
// channels-last volumes on the 5th-generation tensor cores (tcgen05.mma kind::f16, TMEM accumulators), as a
// persistent, warp-specialised kernel that walks a (y,x) tile column along the depth axis.
//
// Why another engine (see profiles/ncu_r1f.md): the tile-at-a-time kernel of k2_conv3d_tc.cu is bound by the shared
// memory pipe -- tensor-core operand reads (62 % of peak) plus the staging stores (18 %) -- not by math or HBM.
// This engine attacks exactly those bytes:
//   * every input plane of a tile column is staged ONCE and feeds the three output planes it touches (the old
//     kernel re-staged each plane twice as z halo), accumulators of the planes in flight live in TMEM;
//   * operands are split into two fp16 pieces instead of two tf32 pieces (x*s = h + l, |l| <= 2^-11 |h|, s a power
//     of two taken from the tensor's tracked abs-max so h never overflows): half the operand bytes per element and
//     twice the MMA rate.  The two pieces of the activation are the two K core matrices of ONE kind::f16 MMA
//     (K = 16 = 8 channels x {h, l}); the weight pieces sit on the N axis:
//         k-plane 0 (x_h):  [ w_h | w_l ]        k-plane 1 (x_l):  [ w_h | 0 ]
//     so one MMA yields x_h*w_h + x_l*w_h (columns "hl=0") and x_h*w_l (columns "hl=1"); the dropped x_l*w_l term is
//     2^-22 relative.  Products are exact in the fp32 accumulator: the result is fp32-equivalent (~1e-6).
//   * producer warps (global -> registers -> fp16 split -> shared), one MMA-issuing thread and epilogue warps
//     (TMEM -> registers -> BN/ReLU/skip -> global) run concurrently through mbarrier pipelines; the packed weights
//     of the layer stay resident in shared memory for the lifetime of the CTA.
// As in k2_conv3d_tc.cu a (dz,dy) tap is a shifted 128-row window of the staged plane (descriptor start address) and
// the dx taps are folded into the MMA N dimension, re-aligned in the epilogue.
//
// Accumulator columns of one 128-row tile and output plane: ((xs * 2) + hl) * CT + co.
//
// z fold: the weight blocks of the z taps an input plane uses are concatenated on N in the order of the output planes
// they feed (S1: kz = 2,1,0 -> planes p-2, p-1, p), and the TMEM slots of consecutive output planes are adjacent, so
// one MMA per (dy window, row tile) updates all of them: A is read from shared memory once instead of three times and
// the instruction count drops 3x (an N = 48 MMA costs ~80 cycles of issue/operand latency for 24 cycles of math).
// The MMA is split only where the slot ring wraps and, in the first window of a plane, between the planes that
// accumulate and the plane that starts (overwrite).
#include <cuda.h>        // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace mvsb200 {

using namespace umma;

constexpr int ZM_S1 = 0, ZM_S2 = 1, ZM_DECONV = 2;
constexpr int ZM_EPI_WARPS = 8, ZM_PROD_WARPS = 9;   // 19 warps with the MMA and planner warps: at most 5 per scheduler, 96 registers each
constexpr int ZM_THREADS = (ZM_EPI_WARPS + ZM_PROD_WARPS + 3) * 32;   // 640: + MMA issuer warp + planner warp + TMA loader warp
constexpr int ZM_WARP_MMA = ZM_EPI_WARPS + ZM_PROD_WARPS, ZM_WARP_PLAN = ZM_WARP_MMA + 1, ZM_WARP_LOAD = ZM_WARP_MMA + 2;
constexpr int ZM_MAXRAW = 8;                                           // raw (fp32, TMA-written) unit buffers in the ring
constexpr int ZM_NPLAN = 4;                                            // plans the planner may run ahead of the issuer
constexpr int ZM_PROD_GROUP = 96;                                      // producer threads working on one unit
constexpr int ZM_NGROUPS = ZM_PROD_WARPS * 32 / ZM_PROD_GROUP;         // units being filled concurrently
constexpr int ZM_HEADER_HALVES = 8;                                    // 16-byte header in front of the packed weights

template <int MODE, int CT> struct ZmCfg {
    static constexpr int NXS = (MODE == ZM_S1) ? 3 : 2;       // x taps folded into N
    static constexpr int NC = NXS * 2 * CT;                   // accumulator columns per 128-row tile (MMA N)
    static constexpr int HALO = (MODE == ZM_S1) ? 2 : 1;
    static constexpr int MT = (CT == 8) ? 2 : 1;              // 128-row MMA tiles per plane
    static constexpr int TY = 7, TX = (CT == 8) ? 32 : 16;
    static constexpr int EY = TY + HALO, EX = TX + HALO;
    static constexpr int ROWS = EY * EX;                      // staged rows per (plane, chunk, sub-grid)
    static constexpr int R_NEED = MT * 128 + HALO * EX;       // rows the shifted windows may touch
    // rounded up to 2 mod 4: consecutive stage buffers (2*RA*16 bytes) then start 64 bytes apart modulo 128, so the
    // 8-byte stores of a producer warp, which fan out over the stage buffers of a unit, are bank-conflict free
    static constexpr int RA = ((R_NEED > ROWS ? R_NEED : ROWS) + 1) / 4 * 4 + 2;
    static constexpr int NSUB = (MODE == ZM_S2) ? 4 : 1;      // (py,px) parity sub-grids staged per plane
    static constexpr int NPXL = (MODE == ZM_S2) ? 2 : 1;      // x-parity weight variants resident per CTA
    static constexpr int NACC = (512 / (MT * NC) >= 8) ? 8 : 5;   // accumulator planes in TMEM
    // TMEM column of (row tile mt, accumulator slot s): mt * NACC * NC + s * NC -- the slots of consecutive output
    // planes are adjacent, so ONE MMA of N = 3 * NC accumulates into all three planes an input plane contributes to
    static constexpr int MT_COLS = NACC * NC;
    static constexpr int XROWS = MT * 128 + 8;
    static constexpr int STAGE_BYTES = 2 * RA * 16;
    static constexpr int UNIT_STAGES = 4;                     // stage buffers of one pipeline slot (= one producer unit)
    static constexpr int UNIT_BYTES = UNIT_STAGES * STAGE_BYTES;
    static constexpr int WBLOCK_BYTES = 2 * NC * 16;          // one (chunk, px, kz, ky) weight block
    static constexpr int WREGION16 = 9 * WBLOCK_BYTES / 16;   // 16-byte units of one (variant, chunk) weight region
    static constexpr int X_BYTES = (NXS - 1) * (CT / 4) * XROWS * 16;   // x-shift exchange buffer of one epilogue team
    static constexpr int NTEAMS = (MT == 2) ? 1 : 2;
    static constexpr int XBUF = (NTEAMS == 1) ? 2 : 1;       // exchange buffers per team (one team: alternate per plane)
    static_assert(TY * EX <= MT * 128, "plane tile does not fit the MMA row tiles");
    static_assert(MT * MT_COLS <= 512, "accumulators exceed TMEM");
    static_assert(NC % 16 == 0 && NC <= 256, "invalid MMA N");
};

struct ZmParams {
    const float *x, *x2, *scale, *bias, *skip;
    const __half *wp;            // packed weights of this layer (header + blocks)
    float *y;
    const float *x_amax, *x2_amax;
    float *y_amax;
    int B, D, H, W, Do, Ho, Wo;
    int Cin1, Cin2, Cout;
    int relu, skip_mode;
    int tiles_x, tiles_y, per, total, nseg, nstages;   // work decomposition: see ZmWalk
    int g1, g2;              // chunks (of 8 channels) per producer unit for x and for x2 (one TMA box each)
    int nraw, raw_bytes;     // ring of raw fp32 unit buffers the TMA loader fills
    int unit_bytes;          // bytes of one converted unit buffer = max(g1, g2) * NPX stage buffers
    int x_cstride, x_coff;   // x may be a channel slice of a wider tensor: voxel pitch and first channel (floats)
    int profile;
    int wide;                // Cout % 8 == 0 with y (and skip) 32-byte aligned: the epilogue moves 8 channels per 256-bit access
    int pdl;                 // launched as a programmatic dependent launch: x / x2 / skip / amax reads follow griddepcontrol.wait
};

// ---- small PTX helpers local to this engine -------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Waiters that are not on the critical path back off between polls so that their spinning does not take issue slots
// from the MMA-issuing thread (bounded like umma::mbar_wait: a pipeline bug must trap, not hang the GPU).
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
#pragma unroll 1
    for (uint32_t i = 0; i < (1u << 24); i++) {
        __nanosleep(64);
        if (mbar_try_wait(bar, parity)) return;
    }
    __trap();
}
__device__ __forceinline__ bool elect_one_sync()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N)
{
    // kind::f16: [4,6) D format 1 = f32 | [7,10) A format 0 = f16 | [10,13) B format 0 = f16 | N >> 3 | M >> 4
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Power-of-two scale that maps a tensor with the given abs-max into [2^13, 2^14): h = fp16(x*s) cannot overflow and
// the absolute rounding floor of the l piece (2^-25) is 2^-38 of the abs-max.  Returns s; *inv = 1/s.
__device__ __forceinline__ float pow2_scale(float amax, float *inv)
{
    if (!(amax > 0.f) || !(amax < 3.0e38f)) {
        *inv = 1.f;
        return 1.f;
    }
    int e = (int)((__float_as_uint(amax) >> 23) & 0xffu) - 127;   // floor(log2(amax)) for normal numbers
    e = max(-100, min(100, e));
    *inv = __uint_as_float((uint32_t)(127 + e - 13) << 23);
    return __uint_as_float((uint32_t)(127 + 13 - e) << 23);
}

// z-march tables.  Local input plane p of a segment -> (local output plane q, z tap kz) contributions.
//   S1      z_in = zb - 1 + p      q = p - kz                       kz = 0,1,2
//   S2      z_in = 2 zb - 1 + p    p even: (q = p/2, kz 0), (q = p/2 - 1, kz 2);  p odd: (q = (p-1)/2, kz 1)
//   DECONV  z_in = zb + p          (q = 2p - 1, kz 0), (q = 2p, kz 1), (q = 2p + 1, kz 2)     [gather form]
template <int MODE> __device__ __forceinline__ int zm_zin(int zb, int p)
{
    return MODE == ZM_S1 ? zb - 1 + p : (MODE == ZM_S2 ? 2 * zb - 1 + p : zb + p);
}
template <int MODE> __device__ __forceinline__ int zm_ncontrib(int p) { return MODE == ZM_S2 ? ((p & 1) ? 1 : 2) : 3; }
template <int MODE> __device__ __forceinline__ void zm_contrib(int p, int j, int &q, int &kz)
{
    if (MODE == ZM_S1) { kz = 2 - j; q = p - kz; }                    // finishing plane first
    else if (MODE == ZM_S2) {
        if (p & 1) { kz = 1; q = (p - 1) >> 1; }
        else if (j == 0) { kz = 2; q = (p >> 1) - 1; }
        else { kz = 0; q = p >> 1; }
    } else { kz = j; q = 2 * p - 1 + j; }
}
// planes of one segment: nq output planes -> number of input planes walked
template <int MODE> __device__ __forceinline__ int zm_nplanes(int nq) { return MODE == ZM_S1 ? nq + 2 : (MODE == ZM_S2 ? 2 * nq + 1 : nq / 2 + 1); }
// output planes that are complete once input plane p has been consumed: [qlo, qhi)
template <int MODE> __device__ __forceinline__ void zm_complete(int p, int nq, int &qlo, int &qhi)
{
    if (MODE == ZM_S1) { qlo = p - 2; qhi = p - 1; }
    else if (MODE == ZM_S2) { if (p & 1) { qlo = qhi = 0; } else { qlo = (p >> 1) - 1; qhi = p >> 1; } }
    else { qlo = 2 * p - 1; qhi = 2 * p + 1; }
    qlo = max(qlo, 0);
    qhi = min(qhi, nq);
}
// z fold tables: the output planes plane p contributes to are consecutive, q = zm_qfirst(p) + i, i < zm_nblk(p), and
// block i of the packed weights of plane type zm_ztype(p) holds filter z index zm_zblock_kz(type, i).
template <int MODE> __host__ __device__ constexpr int zm_ztype(int p) { return MODE == ZM_S2 ? (p & 1) : 0; }
__host__ __device__ constexpr int zm_type_nblk(int mode, int type) { return mode == ZM_S2 ? (type ? 1 : 2) : 3; }
__host__ __device__ constexpr int zm_zblock_kz(int mode, int type, int i)
{
    return mode == ZM_S1 ? 2 - i : (mode == ZM_S2 ? (type ? 1 : (i == 0 ? 2 : 0)) : i);
}
template <int MODE> __device__ __forceinline__ int zm_qfirst(int p)
{
    return MODE == ZM_S1 ? p - 2 : (MODE == ZM_S2 ? ((p & 1) ? (p - 1) >> 1 : (p >> 1) - 1) : 2 * p - 1);
}
// 16-byte units from the start of a (variant, chunk) weight region to the [k-plane][nblk * NC][16 B] matrix of (type, ky)
__host__ __device__ constexpr int zm_wmat_off16(int mode, int type, int ky, int nc)
{
    return mode == ZM_S2 ? (type ? 12 * nc + ky * 2 * nc : ky * 4 * nc) : ky * 6 * nc;
}

// y-dimension tap options of a staged block: variant v (S1: 0; S2: parity of the staged sub-grid; DECONV: parity of the
// output class) -> option j = (filter index, row shift)
__host__ __device__ constexpr int zm_dim_opts(int mode, int v) { return mode == ZM_S1 ? 3 : (mode == ZM_S2 ? (v == 0 ? 2 : 1) : (v == 0 ? 1 : 2)); }
__host__ __device__ constexpr int zm_dim_k(int mode, int v, int j)
{
    return mode == ZM_S1 ? j : (mode == ZM_S2 ? (v == 0 ? 2 * j : 1) : (v == 0 ? 1 : (j == 0 ? 2 : 0)));
}
__host__ __device__ constexpr int zm_dim_shift(int mode, int v, int j) { return mode == ZM_S1 ? j : (mode == ZM_S2 ? (v == 0 ? j : 0) : (v == 0 ? 0 : j)); }

// Optional in-kernel cycle accounting (debug aid, enabled by mvsb200_debug_zm_profile): CTA (0,0) adds the cycles
// its MMA warp, first producer warp and first epilogue warp spend in each kind of wait to this buffer.
//   [0] mma total  [1] mma wait full  [2] mma wait acc-empty  [3] mma issue   [4] stages
//   [8] prod total [9] prod wait empty [10] prod units
//   [16] epi total [17] epi wait acc-full [18] epi named barriers [19] epi planes
__device__ unsigned long long g_zm_prof[32];
static int g_zm_prof_on = 0;

struct ZmTile {
    int b, zb, nq, y0, x0, cls;
};

// Work decomposition over the (tile column, z) index space of a layer -- columns = (x tile, y tile, batch item), z counted in
// output planes (S1, S2) or input planes (DECONV); per output-parity class for DECONV (class = blockIdx.x % 4, so the class
// and with it the resident weight variant is constant per CTA).  Two schemes, chosen per launch by the host:
//   * lockstep (nseg > 0): z is cut into nseg equal segments and the CTAs take the (segment, column) tiles round-robin,
//     neighbouring columns of the SAME segment at the same time -- the in-plane halo rows a tile shares with its
//     neighbours are then in L2 when the neighbour asks for them (a volume that does not fit L2 is read once);
//   * flat (nseg == 0): the index space is flattened column-major and cut into EQUAL contiguous runs, one per CTA, walked
//     as one z-march per column touched.  Every CTA gets the same number of planes whatever the ratio of columns to
//     SMs (50 columns on 148 SMs: 148 CTAs x 33 planes instead of 100 x 48) -- for volumes whose lockstep tiling leaves
//     SMs idle, which are the ones small enough for L2 to absorb the halo re-reads.
// All warp roles of a CTA walk the same sequence.
struct ZmWalk {
    int flat, hi, cls;
};
template <int MODE> __device__ __forceinline__ ZmWalk zm_walk_begin(const ZmParams &p)
{
    ZmWalk w;
    int c = blockIdx.x;
    w.cls = 0;
    if (MODE == ZM_DECONV) { w.cls = c & 3; c >>= 2; }
    if (p.nseg > 0) {          // lockstep: tiles c, c + G, c + 2G, ... of the (segment, column) grid
        w.flat = c;
        w.hi = p.total;        // = segments x columns
    } else {                   // flat: the contiguous run [c * per, (c + 1) * per) of the (column, z) index space
        w.flat = c * p.per;
        w.hi = min(w.flat + p.per, p.total);
    }
    return w;
}
template <int MODE, int CT> __device__ __forceinline__ bool zm_walk_next(const ZmParams &p, ZmWalk &w, ZmTile &t)
{
    using T = ZmCfg<MODE, CT>;
    if (w.flat >= w.hi) return false;
    const int ztot = (MODE == ZM_DECONV) ? p.D : p.Do;
    int col, z, n;
    if (p.nseg > 0) {
        const int ncol = p.total / p.nseg;
        const int sg = w.flat / ncol;
        col = w.flat - sg * ncol;
        z = sg * p.per;
        n = min(p.per, ztot - z);
        w.flat += (MODE == ZM_DECONV) ? (int)(gridDim.x >> 2) : (int)gridDim.x;
    } else {
        col = w.flat / ztot;
        z = w.flat - col * ztot;
        n = min(ztot - z, w.hi - w.flat);
        w.flat += n;
    }
    const int tx = col % p.tiles_x; col /= p.tiles_x;
    const int ty = col % p.tiles_y;
    t.b = col / p.tiles_y;
    t.x0 = tx * T::TX;
    t.y0 = ty * T::TY;
    t.zb = z;
    t.nq = (MODE == ZM_DECONV) ? 2 * n : n;
    t.cls = w.cls;
    return true;
}


// Pipeline stages of one input plane, in issue order:  for py: for chunk c: for px  (py, px: parity sub-grids of S2).
// The producers fill them in UNITS of G chunks x NPX sub-grids = one contiguous run of G*32 bytes per voxel (the
// whole 128-byte voxel record when Cin >= 32), so that a warp-level 16-byte load covers whole cache lines.
__device__ __forceinline__ int zm_unit_chunks(int c, int nch1, int g1, int g2)
{
    return c < nch1 ? g1 : g2;   // constant per tensor: one TMA box shape per tensor map
}

// ---- TMA (cp.async.bulk.tensor) --------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// One 5-D box {channels, x, y, z, b} of the fp32 volume -> dense shared memory [y][x][channels]; elements outside the
// tensor (the conv's zero padding: x, y = -1 / W, H) arrive as zeros; completion is signalled on `bar` (transaction bytes).
__device__ __forceinline__ void tma_load_5d(uint32_t dst_smem, const CUtensorMap *tm, uint32_t bar, int c, int x, int y, int z, int b)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c), "r"(x), "r"(y), "r"(z), "r"(b)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tm)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// 256-bit global accesses (sm_100): 8 channels of a voxel in ONE instruction -- the L1 data pipe counts wavefronts =
// instructions x lines touched, and the epilogue's stores / skip loads touch a line per one or two threads (the strided
// outputs of a transposed layer's parity class: one per thread), so two 128-bit accesses cost twice the wavefronts of one
__device__ __forceinline__ void stg256(float *p, const float4 &a, const float4 &b)
{
    asm volatile("st.global.v8.f32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
                 :: "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w), "l"(p) : "memory");
}
__device__ __forceinline__ void ldg256(const float *p, float4 &a, float4 &b)
{
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

#ifndef ZM_WIDE_OK
#define ZM_WIDE_OK true
#endif
#define ZM_T0() (prof ? clock64() : 0ll)
#define ZM_ACC(var, t0) do { if (prof) var += clock64() - (t0); } while (0)

template <int MODE, int CT>
__global__ void __launch_bounds__(ZM_THREADS, 1)
k2_conv3d_zm_kernel(const ZmParams p, const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_x2)
{
    const bool prof = p.profile && blockIdx.x == 0 && blockIdx.y == 0;
    using T = ZmCfg<MODE, CT>;
    constexpr int NC = T::NC, RA = T::RA, MT = T::MT, NACC = T::NACC, EX = T::EX;
    constexpr int MAXUB = 8;
    constexpr int NPY = (MODE == ZM_S2) ? 2 : 1, NPX = NPY;
    constexpr int TEAM_WARPS = 4 * MT, NTEAMS = ZM_EPI_WARPS / TEAM_WARPS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_full[MAXUB], s_empty[MAXUB], s_accfull[8], s_accempty[8];
    __shared__ __align__(8) unsigned long long s_rawfull[ZM_MAXRAW], s_rawempty[ZM_MAXRAW];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) unsigned long long s_planfull[ZM_NPLAN], s_planempty[ZM_NPLAN];
    __shared__ __align__(16) uint32_t s_plan[ZM_NPLAN][8];   // planner -> MMA issuer, see the planner warp
    __shared__ __align__(16) float s_sc[CT], s_bi[CT];   // epilogue: y = acc * s_sc + s_bi (operand un-scaling and BN folded)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = blockIdx.y;
    const int nch = (p.Cin1 + p.Cin2) >> 3, nch1 = p.Cin1 >> 3;
    const int NUB = p.nstages;   // unit buffers in the ring (the pipeline hands over whole units: 1 wait + 1 commit per unit)
    const int wblocks = T::NPXL * nch * 9;
    unsigned char *sW = smem_raw;
    unsigned char *sA = sW + (size_t)wblocks * T::WBLOCK_BYTES;
    float4 *sXall = reinterpret_cast<float4 *>(sA + (size_t)NUB * p.unit_bytes);
    unsigned char *sRaw = reinterpret_cast<unsigned char *>(sXall) + (size_t)T::NTEAMS * T::XBUF * T::X_BYTES;
    sRaw += (128u - (smem_u32(sRaw) & 127u)) & 127u;   // TMA destinations are 128-byte aligned (the plan reserves the slack)

    if (tid == 0) {
        for (int i = 0; i < NUB; i++) { mbar_init(smem_u32(&s_full[i]), ZM_PROD_GROUP / 32); mbar_init(smem_u32(&s_empty[i]), 1); }
        for (int i = 0; i < NACC; i++) { mbar_init(smem_u32(&s_accfull[i]), 1); mbar_init(smem_u32(&s_accempty[i]), TEAM_WARPS); }
        for (int i = 0; i < p.nraw; i++) { mbar_init(smem_u32(&s_rawfull[i]), 1); mbar_init(smem_u32(&s_rawempty[i]), ZM_PROD_GROUP / 32); }
        for (int i = 0; i < ZM_NPLAN; i++) { mbar_init(smem_u32(&s_planfull[i]), 1); mbar_init(smem_u32(&s_planempty[i]), 1); }
        fence_mbar_init();
    }
    if (warp == ZM_WARP_MMA) tmem_alloc(smem_u32(&s_tmem), 512);
    if (warp == ZM_WARP_LOAD && lane == 0) {
        tma_prefetch_desc(&tm_x);
        if (p.x2) tma_prefetch_desc(&tm_x2);
    }
    // Programmatic dependent launch: let the NEXT kernel of the stream be scheduled as soon as every CTA of this grid is
    // resident (its CTAs need this SM's shared memory, so they start as this grid's CTAs retire and run their own
    // prologue while the slower CTAs of this grid finish).
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // staged rows beyond the plane tile are only ever read by discarded GEMM rows; give them a defined value once
    for (int i = tid; i < NUB * p.unit_bytes / 16; i += ZM_THREADS) reinterpret_cast<uint4 *>(sA)[i] = make_uint4(0, 0, 0, 0);
    if (warp >= ZM_EPI_WARPS && warp < ZM_EPI_WARPS + ZM_PROD_WARPS) {
        // resident weights of this CTA: S1 one variant, S2 both x-parity variants, DECONV the variant of the CTA's
        // output class (the grid is a multiple of 4 wide and the class is the fastest tile index, so every tile of a
        // CTA has class blockIdx.x % 4).  Layer parameters do not depend on the previous kernel: staged before the wait.
        const int ptid = tid - ZM_EPI_WARPS * 32;
        const size_t wblock_halves = T::WBLOCK_BYTES / 2;
        const __half *wsrc = p.wp + ZM_HEADER_HALVES + (size_t)nb * ((MODE == ZM_S1 ? 1 : 2) * nch * 9) * wblock_halves;
        if (MODE == ZM_DECONV) wsrc += (size_t)(blockIdx.x & 1) * (nch * 9) * wblock_halves;
        const uint4 *src = reinterpret_cast<const uint4 *>(wsrc);
        uint4 *dst = reinterpret_cast<uint4 *>(sW);
        const int n16 = wblocks * T::WBLOCK_BYTES / 16;
        for (int i = ptid; i < n16; i += ZM_PROD_WARPS * 32) dst[i] = __ldg(src + i);
    }
    // everything below reads what the previous kernel(s) of the stream wrote (activations, their abs-max scalars)
    if (p.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    // operand scales
    float amax = __ldg(p.x_amax);
    if (p.x2) amax = fmaxf(amax, __ldg(p.x2_amax));
    float inv_sx;
    const float sx = pow2_scale(amax, &inv_sx);
    if (tid < CT) {
        const int co = nb * CT + tid;
        const float unscale = inv_sx * __ldg(reinterpret_cast<const float *>(p.wp) + 1);   // header: {w_scale, w_inv_scale, -, -}
        s_sc[tid] = unscale * ((p.scale && co < p.Cout) ? __ldg(p.scale + co) : 1.f);
        s_bi[tid] = (p.bias && co < p.Cout) ? __ldg(p.bias + co) : 0.f;
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = s_tmem;

    if (warp < ZM_EPI_WARPS) {
        // =================================== epilogue warps ===================================
        // MT == 2: one team of 8 warps, warps 0-3 take row tile 0 and warps 4-7 row tile 1 of every plane;
        // MT == 1: two teams of 4 warps taking alternate planes.  A warp reads the TMEM lanes of quadrant warp % 4.
        const int quad = warp & 3, team = (MT == 2) ? 0 : (warp >> 2), mt = (MT == 2) ? (warp >> 2) : 0;
        float4 *sX0 = sXall + (size_t)team * (T::XBUF * T::X_BYTES / 16);   // XBUF exchange buffers per team, alternating per plane
        int xpar = 0;
        constexpr int C4 = CT / 4;
        const int ncol = min(CT, p.Cout - nb * CT), co0 = nb * CT;
        const int pr = mt * 128 + quad * 32 + lane;           // GEMM row of this thread within the plane tile
        const int oy_l = pr / EX, ox_l = pr % EX;
        float vmax = 0.f;
        int qg = 0;
        long long pe_tot = ZM_T0(), pe_wait = 0, pe_bar = 0, pe_n = 0;
        ZmTile t;
        for (ZmWalk walk = zm_walk_begin<MODE>(p); zm_walk_next<MODE, CT>(p, walk, t);) {
            const int ry = (t.cls >> 1) & 1, rx = t.cls & 1;
            int oy = t.y0 + oy_l, ox = t.x0 + ox_l;
            bool ok_yx = oy_l < T::TY && ox_l < T::TX;
            if (MODE == ZM_DECONV) {
                ok_yx = ok_yx && oy < p.H && ox < p.W;
                oy = 2 * oy + ry; ox = 2 * ox + rx;
            }
            ok_yx = ok_yx && oy < p.Ho && ox < p.Wo;
            // The skip operand does not depend on the accumulator: it is fetched one of this team's planes AHEAD, so its
            // (HBM) latency overlaps the previous plane's epilogue instead of sitting between "accumulator ready" and
            // "store" (which is what bounded the up-sampling layers: the epilogue was 80 % busy, mostly in this wait).
            const long long plane_stride = (long long)p.Ho * p.Wo * p.Cout;
            const long long o_yx = (((long long)t.b * p.Do * p.Ho + oy) * p.Wo + ox) * p.Cout + co0;
            const int oz0 = (MODE == ZM_DECONV) ? 2 * t.zb : t.zb;
            auto load_skip = [&](int q, float4 (&sk)[C4]) {
                const bool want = q < t.nq && ok_yx && oz0 + q < p.Do && p.skip_mode != MVSB200_SKIP_NONE;
#pragma unroll
                for (int c4 = 0; c4 < C4; c4 += 2) {
                    sk[c4] = sk[c4 + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!want) continue;
                    const float *src = p.skip + o_yx + (oz0 + q) * plane_stride + c4 * 4;
                    if (p.wide && ZM_WIDE_OK) {
                        ldg256(src, sk[c4], sk[c4 + 1]);
                    } else {
                        if (c4 * 4 < ncol) sk[c4] = ldg4(src);
                        if (c4 * 4 + 4 < ncol) sk[c4 + 1] = ldg4(src + 4);
                    }
                }
            };
            constexpr bool AHEAD = (CT == 8);   // CT = 16: a second skip buffer would spill; its two teams overlap instead
            float4 sk[C4], sk_next[AHEAD ? C4 : 1];
            if (AHEAD) load_skip((NTEAMS == 2 && (qg & 1) != team) ? 1 : 0, reinterpret_cast<float4 (&)[C4]>(sk_next));   // first plane of this team in the tile
            for (int q = 0; q < t.nq; q++, qg++) {
                if (NTEAMS == 2 && (qg & 1) != team) continue;
                const int slot = qg % NACC;
                const int oz = oz0 + q;
                const bool ok = ok_yx && oz < p.Do;
                const long long o = o_yx + oz * plane_stride;
                if (AHEAD) {
#pragma unroll
                    for (int c4 = 0; c4 < C4; c4++) sk[c4] = sk_next[c4 < (AHEAD ? C4 : 1) ? c4 : 0];
                    load_skip(q + NTEAMS, reinterpret_cast<float4 (&)[C4]>(sk_next));
                } else {
                    load_skip(q, sk);
                }
                long long tq = ZM_T0();
                mbar_wait_relaxed(smem_u32(&s_accfull[slot]), (uint32_t)(qg / NACC) & 1u);
                ZM_ACC(pe_wait, tq);
                pe_n++;
                tc_fence_after_sync();
                float4 *sX = sX0 + xpar * (T::X_BYTES / 16);
                if (T::XBUF == 2) xpar ^= 1;
                float r0[CT];
                const uint32_t taddr0 = tmem + ((uint32_t)(quad * 32) << 16) + mt * T::MT_COLS + slot * NC;
                if (CT == 8) {
                    // all TMEM loads of the row in flight at once, one wait
                    float v[T::NXS][16];
#pragma unroll
                    for (int xs = 0; xs < T::NXS; xs++) tmem_ld16(taddr0 + xs * 16, v[xs]);
                    tmem_ld_wait();
#pragma unroll
                    for (int xs = 0; xs < T::NXS; xs++) {
#pragma unroll
                        for (int k = 0; k < 8; k++) v[xs][k] += v[xs][8 + k];
                        if (xs == 0) {
#pragma unroll
                            for (int k = 0; k < 8; k++) r0[k] = v[0][k];
                        } else {
#pragma unroll
                            for (int c4 = 0; c4 < 2; c4++)
                                sX[((xs - 1) * C4 + c4) * T::XROWS + pr] = make_float4(v[xs][c4 * 4], v[xs][c4 * 4 + 1], v[xs][c4 * 4 + 2], v[xs][c4 * 4 + 3]);
                        }
                    }
                } else {
#pragma unroll
                    for (int xs = 0; xs < T::NXS; xs++) {
                        float v[2 * CT];
                        const uint32_t taddr = taddr0 + xs * 2 * CT;
#pragma unroll
                        for (int j = 0; j < 2 * CT; j += 16) tmem_ld16(taddr + j, v + j);
                        tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < CT; k++) v[k] += v[CT + k];
                        if (xs == 0) {
#pragma unroll
                            for (int k = 0; k < CT; k++) r0[k] = v[k];
                        } else {
#pragma unroll
                            for (int c4 = 0; c4 < C4; c4++)
                                sX[((xs - 1) * C4 + c4) * T::XROWS + pr] = make_float4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
                        }
                    }
                }
                // the accumulator plane is in registers / shared memory: hand the TMEM slot back to the MMA thread
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&s_accempty[slot]));
                tq = ZM_T0();
                if (team == 0) named_bar_sync(1, TEAM_WARPS * 32); else named_bar_sync(2, TEAM_WARPS * 32);
                ZM_ACC(pe_bar, tq);
                if (ok) {
                    float4 first = make_float4(0.f, 0.f, 0.f, 0.f);   // the even channel quad, held for the 256-bit store
#pragma unroll
                    for (int c4 = 0; c4 < C4; c4++) {
                        if (c4 * 4 >= ncol) break;
                        float r[4] = {r0[c4 * 4], r0[c4 * 4 + 1], r0[c4 * 4 + 2], r0[c4 * 4 + 3]};
#pragma unroll
                        for (int xs = 1; xs < T::NXS; xs++) {
                            const float4 nbv = sX[((xs - 1) * C4 + c4) * T::XROWS + pr + xs];
                            r[0] += nbv.x; r[1] += nbv.y; r[2] += nbv.z; r[3] += nbv.w;
                        }
                        const float4 sc = *reinterpret_cast<const float4 *>(&s_sc[c4 * 4]), bi = *reinterpret_cast<const float4 *>(&s_bi[c4 * 4]);
                        r[0] = fmaf(r[0], sc.x, bi.x); r[1] = fmaf(r[1], sc.y, bi.y);
                        r[2] = fmaf(r[2], sc.z, bi.z); r[3] = fmaf(r[3], sc.w, bi.w);
                        if (p.skip_mode == MVSB200_SKIP_BEFORE_RELU) { r[0] += sk[c4].x; r[1] += sk[c4].y; r[2] += sk[c4].z; r[3] += sk[c4].w; }
                        if (p.relu) { r[0] = fmaxf(r[0], 0.f); r[1] = fmaxf(r[1], 0.f); r[2] = fmaxf(r[2], 0.f); r[3] = fmaxf(r[3], 0.f); }
                        if (p.skip_mode == MVSB200_SKIP_AFTER_RELU) { r[0] += sk[c4].x; r[1] += sk[c4].y; r[2] += sk[c4].z; r[3] += sk[c4].w; }
                        vmax = fmaxf(fmaxf(vmax, fmaxf(fabsf(r[0]), fabsf(r[1]))), fmaxf(fabsf(r[2]), fabsf(r[3])));
                        const float4 cur = make_float4(r[0], r[1], r[2], r[3]);
                        if (p.wide && ZM_WIDE_OK) {       // (ncol is 8 or 16) one 256-bit store per 8 channels
                            if ((c4 & 1) == 0) first = cur; else stg256(p.y + o + (c4 - 1) * 4, first, cur);
                        } else {
                            st4(p.y + o + c4 * 4, cur);
                        }
                    }
                }
                // One team: no second barrier -- the next plane writes the OTHER exchange buffer, and the barrier of that
                // plane orders this plane's reads before the writes of the plane after it.  Two teams: single buffer.
                if (T::XBUF == 1) { if (team == 0) named_bar_sync(1, TEAM_WARPS * 32); else named_bar_sync(2, TEAM_WARPS * 32); }
            }
        }
        if (p.y_amax) {
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, m));
            if (lane == 0 && vmax > 0.f) atomicMax(reinterpret_cast<unsigned int *>(p.y_amax), __float_as_uint(vmax));
        }
        if (prof && tid == 0) {
            g_zm_prof[16] += clock64() - pe_tot; g_zm_prof[17] += pe_wait; g_zm_prof[18] += pe_bar; g_zm_prof[19] += pe_n;
        }
    } else if (warp < ZM_EPI_WARPS + ZM_PROD_WARPS) {
        // =================================== converter warps ===================================
        // The TMA loader warp streams raw fp32 units (one 5-D box each: G chunks x the haloed plane tile, zero padding
        // filled in by the TMA unit) into the raw ring; a group of 96 threads turns a raw unit into the scaled fp16 h / l
        // pieces in the canonical K-major layout the MMAs read.  The converters never touch global memory: their loads
        // are conflict-free LDS.128 at a fixed stride, so a unit costs ~25 instructions per 16-byte piece and no
        // memory latency -- the bytes in flight are set by the depth of the raw ring, not by the thread count.
        const int ptid = tid - ZM_EPI_WARPS * 32;
        const int group = ptid / ZM_PROD_GROUP, gt = ptid % ZM_PROD_GROUP;
        constexpr int XV = EX * NPX;          // voxels of one staged row run (contiguous in x)
        constexpr int NVOX = T::EY * XV;
        constexpr int BATCH = 8;              // 16-byte pieces converted per loop trip (loads first, then the arithmetic)
        int un = 0;               // units so far (all groups count all units)
        int ub = 0;               // ring position of the current unit's buffer
        uint32_t uphase = 1;      // parity to wait for on its empty barrier (first pass: free)
        long long pp_tot = ZM_T0(), pp_wait = 0, pp_n = 0;
        const uint32_t raw0 = smem_u32(sRaw);
        ZmTile t;
        for (ZmWalk walk = zm_walk_begin<MODE>(p); zm_walk_next<MODE, CT>(p, walk, t);) {
            const int np = zm_nplanes<MODE>(t.nq);
            for (int pl = 0; pl < np; pl++) {
                const int gz = zm_zin<MODE>(t.zb, pl);
                if ((unsigned)gz >= (unsigned)p.D) continue;   // an all-zero plane contributes nothing: no stage at all
                for (int py = 0; py < NPY; py++) {
                    for (int c0 = 0; c0 < nch;) {
                        const int G = zm_unit_chunks(c0, nch1, p.g1, p.g2);
                        const int my_ub = ub;
                        const uint32_t my_phase = uphase;
                        if (++ub == NUB) { ub = 0; uphase ^= 1u; }
                        c0 += G;
                        const int u = un++;
                        if ((u % ZM_NGROUPS) != group) continue;
                        const int rs = u % p.nraw;
                        const uint32_t rphase = (uint32_t)(u / p.nraw) & 1u;
                        const int lgp = (G == 4) ? 3 : (G == 2 ? 2 : 1);     // log2(16-byte pieces per voxel)
                        // A thread's pieces are idx = gt + k * ZM_PROD_GROUP; the group size is a multiple of the pieces
                        // per voxel, so the piece within the voxel -- and with it the chunk, the 8-byte half and (S1,
                        // DECONV) the stage buffer -- is the same for every k; the raw unit is dense [voxel][piece],
                        // so piece idx sits at byte idx * 16.
                        const int piece = gt & ((1 << lgp) - 1);
                        const int vstep = ZM_PROD_GROUP >> lgp;
                        uint32_t dst_px[NPX];                                  // shared address of row 0 in the stage(s) of this piece
#pragma unroll
                        for (int px = 0; px < NPX; px++)
                            dst_px[px] = smem_u32(sA) + my_ub * p.unit_bytes + ((piece >> 1) * NPX + px) * T::STAGE_BYTES + (piece & 1) * 8;
                        {
                            const long long tq = ZM_T0();
                            mbar_wait_relaxed(smem_u32(&s_rawfull[rs]), rphase);      // the TMA box has landed
                            mbar_wait_relaxed(smem_u32(&s_empty[my_ub]), my_phase);   // the MMAs that read this unit buffer are done
                            ZM_ACC(pp_wait, tq);
                            pp_n++;
                        }
                        const uint32_t src0 = raw0 + rs * p.raw_bytes + gt * 16;
                        for (int v0 = gt >> lgp, k0 = 0; v0 < NVOX; v0 += BATCH * vstep, k0 += BATCH) {
                            float4 v[BATCH];
#pragma unroll
                            for (int k = 0; k < BATCH; k++) {
                                v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (v0 + k * vstep < NVOX)
                                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                                 : "=f"(v[k].x), "=f"(v[k].y), "=f"(v[k].z), "=f"(v[k].w)
                                                 : "r"(src0 + (k0 + k) * (ZM_PROD_GROUP * 16)));
                            }
#pragma unroll
                            for (int k = 0; k < BATCH; k++) {
                                const int vx = v0 + k * vstep;
                                if (vx >= NVOX) break;
                                const int lxv = vx % XV, ly = vx / XV;
                                const int px = (NPX == 2) ? (lxv & 1) : 0, lx = (NPX == 2) ? (lxv >> 1) : lxv;
                                const float f0 = v[k].x * sx, f1 = v[k].y * sx, f2 = v[k].z * sx, f3 = v[k].w * sx;
                                const __half2 h01 = __floats2half2_rn(f0, f1), h23 = __floats2half2_rn(f2, f3);
                                const float2 g01 = __half22float2(h01), g23 = __half22float2(h23);
                                const __half2 l01 = __floats2half2_rn(f0 - g01.x, f1 - g01.y), l23 = __floats2half2_rn(f2 - g23.x, f3 - g23.y);
                                const uint32_t dst = ((NPX == 2 && px) ? dst_px[NPX - 1] : dst_px[0]) + (ly * EX + lx) * 16;
                                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(*reinterpret_cast<const uint32_t *>(&h01)),
                                             "r"(*reinterpret_cast<const uint32_t *>(&h23)) : "memory");
                                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst + RA * 16), "r"(*reinterpret_cast<const uint32_t *>(&l01)),
                                             "r"(*reinterpret_cast<const uint32_t *>(&l23)) : "memory");
                            }
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(smem_u32(&s_rawempty[rs]));   // raw slot read: the loader may refill it
                            mbar_arrive(smem_u32(&s_full[my_ub]));    // converted unit ready for the MMAs
                        }
                    }
                }
            }
        }
        if (prof && ptid == 0) { g_zm_prof[8] += clock64() - pp_tot; g_zm_prof[9] += pp_wait; g_zm_prof[10] += pp_n; }
    } else if (warp == ZM_WARP_LOAD) {
        // =================================== TMA loader ===================================
        // One thread walks the units in pipeline order and keeps the raw ring full: wait until the converters have
        // drained a slot, post the transaction size on its "full" barrier and issue the box load.
        if (lane == 0) {
            constexpr int XV = EX * NPX;
            const uint32_t raw0 = smem_u32(sRaw);
            int u = 0;
            ZmTile t;
            for (ZmWalk walk = zm_walk_begin<MODE>(p); zm_walk_next<MODE, CT>(p, walk, t);) {
                const int np = zm_nplanes<MODE>(t.nq);
                const int xbase = (MODE == ZM_S1) ? t.x0 - 1 : (MODE == ZM_S2 ? 2 * t.x0 - 1 : t.x0);
                for (int pl = 0; pl < np; pl++) {
                    const int gz = zm_zin<MODE>(t.zb, pl);
                    if ((unsigned)gz >= (unsigned)p.D) continue;
                    for (int py = 0; py < NPY; py++) {
                        const int ybase = (MODE == ZM_S1) ? t.y0 - 1 : (MODE == ZM_S2 ? 2 * t.y0 - 1 + py : t.y0);
                        for (int c0 = 0; c0 < nch;) {
                            const int G = zm_unit_chunks(c0, nch1, p.g1, p.g2);
                            const int rs = u % p.nraw;
                            mbar_wait_relaxed(smem_u32(&s_rawempty[rs]), ((uint32_t)(u / p.nraw) & 1u) ^ 1u);   // first pass: free
                            const uint32_t bar = smem_u32(&s_rawfull[rs]);
                            mbar_arrive_expect_tx(bar, (uint32_t)(G * 32 * T::EY * XV));
                            if (c0 < nch1) tma_load_5d(raw0 + rs * p.raw_bytes, &tm_x, bar, c0 * 8, xbase, ybase, gz, t.b);
                            else tma_load_5d(raw0 + rs * p.raw_bytes, &tm_x2, bar, (c0 - nch1) * 8, xbase, ybase, gz, t.b);
                            c0 += G;
                            u++;
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == ZM_WARP_MMA) {
        // =================================== MMA issuer ===================================
        // ONE thread issues every MMA of the CTA and the queue between it and the tensor core is shallow, so whatever
        // this warp does besides issuing idles the tensor core.  It therefore does nothing else: which accumulator slots
        // an input plane feeds, how they are cut into runs of adjacent TMEM columns, the waits for drained slots and
        // the list of output planes to hand to the epilogue come ready-made from the planner warp (s_plan); ring
        // positions and phases are carried incrementally, the tap loops are fully unrolled, a descriptor is a 32-bit
        // low word (start address | LBO) plus a constant high word.  The whole warp runs the (warp-uniform) control
        // flow; only the tcgen05 instructions are predicated on the elected lane, which keeps descriptors in uniform
        // registers instead of a per-instruction leader-election loop.
        {
            const uint32_t a_hi = (uint32_t)(smem_desc(0, RA * 16, 128) >> 32), b_hi = (uint32_t)(smem_desc(0, NC * 16, 128) >> 32);
            const uint32_t a_lo0 = (uint32_t)smem_desc(smem_u32(sA), RA * 16, 128), b_lo0 = (uint32_t)smem_desc(smem_u32(sW), 0, 128);
            const uint32_t full0 = smem_u32(&s_full[0]), empty0 = smem_u32(&s_empty[0]);
            const uint32_t accfull0 = smem_u32(&s_accfull[0]);
            const uint32_t planfull0 = smem_u32(&s_planfull[0]), planempty0 = smem_u32(&s_planempty[0]);
            int ub = 0;              // unit-buffer ring position
            uint32_t uphase = 0;     // parity of the current pass over the ring
            int pi = 0;              // plan ring position
            uint32_t pphase = 0;
            long long pm_tot = ZM_T0(), pm_full = 0, pm_acc = 0, pm_issue = 0, pm_n = 0;
            ZmTile t;
            for (ZmWalk walk = zm_walk_begin<MODE>(p); zm_walk_next<MODE, CT>(p, walk, t);) {
                const int np = zm_nplanes<MODE>(t.nq);
                const int vy_cls = (t.cls >> 1) & 1;
                for (int pl = 0; pl < np; pl++) {
                    // ---- the plan of this input plane ----
                    long long tq = ZM_T0();
                    mbar_wait(planfull0 + pi * 8, pphase);
                    ZM_ACC(pm_acc, tq);
                    const uint4 pa = *reinterpret_cast<const uint4 *>(&s_plan[pi][0]);
                    const uint4 pb = *reinterpret_cast<const uint4 *>(&s_plan[pi][4]);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(planempty0 + pi * 8);
                    if (++pi == ZM_NPLAN) { pi = 0; pphase ^= 1u; }
                    const uint32_t flags = pa.x;
                    if (flags & 1u) {   // the plane lies inside the volume: its units are in the pipeline
                        const int nf = (flags >> 8) & 0xf, nr = (flags >> 12) & 0xf;
                        const int ztype = zm_ztype<MODE>(pl), nblk = zm_type_nblk(MODE, ztype);
                        const uint32_t b_lbo = (uint32_t)(nblk * NC) << 16;   // LBO field: k-plane stride of the folded B matrix
                        // run word: TMEM column | B row offset << 10 | N << 20 | accumulate << 31
#define ZM_RUN(w, d, bo, id) const uint32_t d = (w) & 0x3ffu, bo = ((w) >> 10) & 0x3ffu, id = idesc_f16(128, ((w) >> 20) & 0x1ffu)
                        ZM_RUN(pa.y, fd0, fb0, fi0); ZM_RUN(pa.z, fd1, fb1, fi1); ZM_RUN(pa.w, fd2, fb2, fi2);
                        ZM_RUN(pb.x, rd0, rb0, id0); ZM_RUN(pb.y, rd1, rb1, id1);
#undef ZM_RUN
                        const uint32_t fa0 = pa.y >> 31, fa1 = pa.z >> 31, fa2 = pa.w >> 31;
                        tc_fence_after_sync();
                        bool first = true;
                        for (int py = 0; py < NPY; py++) {
                            const int vy = (MODE == ZM_S2) ? py : vy_cls;
                            for (int c0 = 0; c0 < nch;) {
                                const int G = zm_unit_chunks(c0, nch1, p.g1, p.g2);
                                tq = ZM_T0();
                                mbar_wait(full0 + ub * 8, uphase);
                                ZM_ACC(pm_full, tq);
                                pm_n++;
                                tc_fence_after_sync();
                                tq = ZM_T0();
                                if (elect_one_sync()) {
                                    uint32_t a_lo = a_lo0 + ub * (p.unit_bytes / 16);
                                    for (int cc = 0; cc < G; cc++) {
#pragma unroll
                                        for (int px = 0; px < NPX; px++, a_lo += T::STAGE_BYTES / 16) {
                                            const uint32_t b_c = b_lo0 + b_lbo + (px * nch + c0 + cc) * T::WREGION16;
#pragma unroll
                                            for (int jy = 0; jy < 3; jy++) {
                                                if (jy >= zm_dim_opts(MODE, vy)) continue;
                                                const uint32_t b_y = b_c + zm_wmat_off16(MODE, ztype, zm_dim_k(MODE, vy, jy), NC);
                                                const uint32_t a_y = a_lo + zm_dim_shift(MODE, vy, jy) * EX;
                                                const uint64_t ad0 = ((uint64_t)a_hi << 32) | a_y, ad1 = ((uint64_t)a_hi << 32) | (a_y + 128);
                                                if (first) {
                                                    first = false;
                                                    mma_f16(tmem + fd0, ad0, ((uint64_t)b_hi << 32) | (b_y + fb0), fi0, fa0);
                                                    if (MT == 2) mma_f16(tmem + T::MT_COLS + fd0, ad1, ((uint64_t)b_hi << 32) | (b_y + fb0), fi0, fa0);
                                                    if (nf > 1) {
                                                        mma_f16(tmem + fd1, ad0, ((uint64_t)b_hi << 32) | (b_y + fb1), fi1, fa1);
                                                        if (MT == 2) mma_f16(tmem + T::MT_COLS + fd1, ad1, ((uint64_t)b_hi << 32) | (b_y + fb1), fi1, fa1);
                                                    }
                                                    if (nf > 2) {
                                                        mma_f16(tmem + fd2, ad0, ((uint64_t)b_hi << 32) | (b_y + fb2), fi2, fa2);
                                                        if (MT == 2) mma_f16(tmem + T::MT_COLS + fd2, ad1, ((uint64_t)b_hi << 32) | (b_y + fb2), fi2, fa2);
                                                    }
                                                } else {
                                                    mma_f16(tmem + rd0, ad0, ((uint64_t)b_hi << 32) | (b_y + rb0), id0, 1u);
                                                    if (MT == 2) mma_f16(tmem + T::MT_COLS + rd0, ad1, ((uint64_t)b_hi << 32) | (b_y + rb0), id0, 1u);
                                                    if (nr > 1) {
                                                        mma_f16(tmem + rd1, ad0, ((uint64_t)b_hi << 32) | (b_y + rb1), id1, 1u);
                                                        if (MT == 2) mma_f16(tmem + T::MT_COLS + rd1, ad1, ((uint64_t)b_hi << 32) | (b_y + rb1), id1, 1u);
                                                    }
                                                }
                                            }
                                        }
                                    }
                                    mma_commit(empty0 + ub * 8);
                                }
                                first = false;
                                __syncwarp();
                                ZM_ACC(pm_issue, tq);
                                if (++ub == NUB) { ub = 0; uphase ^= 1u; }
                                c0 += G;
                            }
                        }
                    }
                    // output planes that are complete now: hand them to the epilogue
                    const int ncommit = (flags >> 16) & 0xf;
                    if (ncommit > 0 && elect_one_sync()) {
                        mma_commit(accfull0 + (pb.z & 0xffu) * 8);
                        if (ncommit > 1) mma_commit(accfull0 + ((pb.z >> 8) & 0xffu) * 8);
                    }
                    __syncwarp();
                }
            }
            if (prof && lane == 0) {
                g_zm_prof[0] += clock64() - pm_tot; g_zm_prof[1] += pm_full; g_zm_prof[2] += pm_acc; g_zm_prof[3] += pm_issue; g_zm_prof[4] += pm_n;
            }
        }
        __syncwarp();
    } else if (warp == ZM_WARP_PLAN) {
        // =================================== planner warp ===================================
        // Runs ahead of the MMA issuer (ring of ZM_NPLAN plans).  Per input plane: the output planes q = qf + i,
        // i in [i0, i1), it feeds sit in consecutive accumulator ring positions; cut them into runs of adjacent TMEM
        // columns -- runs "F" for the first window of the plane (also cut between planes that accumulate and planes
        // that start: their first MMA overwrites), runs "R" for every other window -- wait until the epilogue has
        // drained the slots that start, and list the output planes that are complete after this plane.
        const uint32_t accempty0 = smem_u32(&s_accempty[0]);
        int qbase = 0;           // output planes of the tiles already planned (accumulator ring position = qbase + q)
        uint32_t started = 0;    // bit per accumulator slot: the plane in it has received its first MMA
        int pi = 0;
        uint32_t pphase = 1;     // first pass over the plan ring: free
        ZmTile t;
        for (ZmWalk walk = zm_walk_begin<MODE>(p); zm_walk_next<MODE, CT>(p, walk, t);) {
            const int np = zm_nplanes<MODE>(t.nq);
            for (int pl = 0; pl < np; pl++) {
                const int gz = zm_zin<MODE>(t.zb, pl);
                uint32_t flags = 0, fw[3] = {0, 0, 0}, rw[2] = {0, 0}, commits = 0;
                if ((unsigned)gz < (unsigned)p.D) {
                    const int nblk = zm_type_nblk(MODE, zm_ztype<MODE>(pl));
                    const int qf = zm_qfirst<MODE>(pl);
                    const int i0 = max(0, -qf), i1 = min(nblk, t.nq - qf);
                    int nf = 0, nr = 0;
                    uint32_t fcur_n = 0, fcur_a = 0, rcur_n = 0;
                    for (int i = i0; i < i1; i++) {
                        const int qg = qbase + qf + i, aslot = qg % NACC;
                        const uint32_t st = (started >> aslot) & 1u;
                        if (!st) mbar_wait_relaxed(accempty0 + aslot * 8, ((uint32_t)(qg / NACC) & 1u) ^ 1u);   // drained by the epilogue?
                        // (an MMA is at most 256 columns wide)
                        if (nr == 0 || aslot == 0 || rcur_n + NC > 256) { rw[nr] = (uint32_t)(aslot * NC) | ((uint32_t)(i * NC) << 10); nr++; rcur_n = 0; }
                        rcur_n += NC;
                        rw[nr - 1] = (rw[nr - 1] & 0x000fffffu) | (rcur_n << 20);
                        if (nf == 0 || aslot == 0 || st != fcur_a || fcur_n + NC > 256) {
                            fw[nf] = (uint32_t)(aslot * NC) | ((uint32_t)(i * NC) << 10) | (st << 31);
                            nf++; fcur_n = 0; fcur_a = st;
                        }
                        fcur_n += NC;
                        fw[nf - 1] = (fw[nf - 1] & 0x800fffffu) | (fcur_n << 20);
                        started |= 1u << aslot;
                    }
                    flags = 1u | ((uint32_t)nf << 8) | ((uint32_t)nr << 12);
                }
                int qlo, qhi;
                zm_complete<MODE>(pl, t.nq, qlo, qhi);
                int nc = 0;
                for (int q = qlo; q < qhi; q++, nc++) {
                    const int aslot = (qbase + q) % NACC;
                    commits |= (uint32_t)aslot << (8 * nc);
                    started &= ~(1u << aslot);
                }
                flags |= (uint32_t)nc << 16;
                mbar_wait_relaxed(smem_u32(&s_planempty[pi]), pphase);
                if (lane == 0) {
                    *reinterpret_cast<uint4 *>(&s_plan[pi][0]) = make_uint4(flags, fw[0], fw[1], fw[2]);
                    *reinterpret_cast<uint4 *>(&s_plan[pi][4]) = make_uint4(rw[0], rw[1], commits, 0u);
                    mbar_arrive(smem_u32(&s_planfull[pi]));
                }
                __syncwarp();
                if (++pi == ZM_NPLAN) { pi = 0; pphase ^= 1u; }
            }
            qbase += t.nq;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == ZM_WARP_MMA) tmem_dealloc(tmem, 512);
}

// ---- weight packing ----------------------------------------------------------------------------------------------
// layout: header {w_scale, w_inv_scale, 0, 0} (fp32, 16 bytes) then
//   [N block][x-parity variant (S2: px, DECONV: rx; S1: single)][chunk] regions of 9 * 2 * NC * 16 bytes, each holding
//   per (plane type, ky) one matrix [k-plane 2][z blocks * NC][8 halves] at zm_wmat_off16 (z fold, see the top)
struct ZmPackParams {
    const float *w;
    __half *wp;
    int Cin, Cout, nch, mode, CT, NC, nvar;
};

__global__ void k2_zm_absmax_kernel(const float *x, long long n, float *out)
{
    float m = 0.f;
    const long long n4 = n >> 2;
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x4 + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(__ldg(x + i)));
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, k));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int *>(out), __float_as_uint(m));
}

__global__ void k2_zm_header_kernel(float *hdr)
{
    // hdr[0] holds the abs-max of the weights on entry
    float inv;
    const float s = pow2_scale(hdr[0], &inv);
    hdr[0] = s; hdr[1] = inv; hdr[2] = 0.f; hdr[3] = 0.f;
}

__host__ __device__ constexpr int zm_kx_of(int mode, int var, int xs)
{
    // filter x index multiplied with the rows shifted by xs, or -1
    return mode == ZM_S1 ? xs
         : mode == ZM_S2 ? (var == 0 ? (xs == 0 ? 0 : 2) : (xs == 0 ? 1 : -1))
                         : (var == 0 ? (xs == 0 ? 1 : -1) : (xs == 0 ? 2 : 0));
}

__global__ void k2_zm_pack_kernel(const ZmPackParams p)
{
    // one block per (nb, var, c, plane type, ky); threads over (k-plane, z block, column, element)
    int id = blockIdx.x;
    const int ntype = (p.mode == ZM_S2) ? 2 : 1;
    const int ky = id % 3; id /= 3;
    const int type = id % ntype; id /= ntype;
    const int c = id % p.nch; id /= p.nch;
    const int var = id % p.nvar;
    const int nb = id / p.nvar;
    const int nblk = zm_type_nblk(p.mode, type);
    const float s = reinterpret_cast<const float *>(p.wp)[0];
    const size_t region = (size_t)((nb * p.nvar + var) * p.nch + c) * (9 * 2 * p.NC * 8);   // halves
    __half *dst = p.wp + ZM_HEADER_HALVES + region + (size_t)zm_wmat_off16(p.mode, type, ky, p.NC) * 8;
    const int rows = nblk * p.NC;
    for (int i = threadIdx.x; i < 2 * rows * 8; i += blockDim.x) {
        const int e = i & 7;
        int r = i >> 3;
        const int row = r % rows;
        const int kp = r / rows;
        const int col = row % p.NC, blk = row / p.NC;
        const int kz = zm_zblock_kz(p.mode, type, blk);
        const int co_l = col % p.CT, g = col / p.CT;
        const int hl = g & 1, xs = g >> 1;
        const int kx = zm_kx_of(p.mode, var, xs);
        const int ci = c * 8 + e, co = nb * p.CT + co_l;
        float val = 0.f;
        if (kx >= 0 && co < p.Cout) val = p.w[((size_t)((kz * 3 + ky) * 3 + kx) * p.Cin + ci) * p.Cout + co] * s;
        const __half h = __float2half_rn(val);
        const __half l = __float2half_rn(val - __half2float(h));
        dst[i] = (kp == 0) ? (hl ? l : h) : (hl ? __float2half_rn(0.f) : h);
    }
}

static int zm_mode(const mvsb200_conv3d_desc *d) { return d->transposed ? ZM_DECONV : (d->stride == 2 ? ZM_S2 : ZM_S1); }
static int zm_ct(const mvsb200_conv3d_desc *d) { return d->Cout == 8 ? 8 : 16; }
static int zm_nvar(int mode) { return mode == ZM_S1 ? 1 : 2; }

// ---- shared-memory plan ---------------------------------------------------------------------------------------------
// resident weights | NUB converted unit buffers | x-shift exchange buffers | NRAW raw (TMA) unit buffers
struct ZmDims {
    int npx, npxl, ey, xv, stage_bytes, wblock_bytes, x_bytes_total;
};
template <int MODE, int CT> static ZmDims zm_dims_of()
{
    using T = ZmCfg<MODE, CT>;
    ZmDims d;
    d.npx = (MODE == ZM_S2) ? 2 : 1;
    d.npxl = T::NPXL;
    d.ey = T::EY;
    d.xv = T::EX * d.npx;
    d.stage_bytes = T::STAGE_BYTES;
    d.wblock_bytes = T::WBLOCK_BYTES;
    d.x_bytes_total = T::NTEAMS * T::XBUF * T::X_BYTES;
    return d;
}
static ZmDims zm_dims(int mode, int ct)
{
    if (mode == ZM_S1) return ct == 8 ? zm_dims_of<ZM_S1, 8>() : zm_dims_of<ZM_S1, 16>();
    if (mode == ZM_S2) return ct == 8 ? zm_dims_of<ZM_S2, 8>() : zm_dims_of<ZM_S2, 16>();
    return ct == 8 ? zm_dims_of<ZM_DECONV, 8>() : zm_dims_of<ZM_DECONV, 16>();
}

struct ZmPlan {
    int g1, g2, nub, nraw, raw_bytes, unit_bytes;
    size_t smem;   // dynamic shared memory; 0: the layer does not fit this engine
};
// Chunks per unit are a power of two that divides the tensor's chunk count (one TMA box shape per tensor), as large as
// the shared memory left by the resident weights allows; then the raw ring -- the bytes in flight from HBM -- gets
// what is left (up to ZM_MAXRAW slots), the converted ring one more buffer if there is room.
static ZmPlan zm_plan(int mode, int ct, int nch1, int nch2)
{
    const ZmDims d = zm_dims(mode, ct);
    ZmPlan pl = {};
    const size_t limit = 227 * 1024 - 2560;   // static shared memory (barriers, plans, scale / bias) + alignment slack
    const size_t fixed = (size_t)d.npxl * (nch1 + nch2) * 9 * d.wblock_bytes + d.x_bytes_total;
    auto pow2_div = [](int n, int gmax) { int g = 1; while (g * 2 <= gmax && n % (g * 2) == 0) g *= 2; return g; };
    for (int gmax = 4 / d.npx; gmax >= 1; gmax /= 2) {
        const int g1 = pow2_div(nch1, gmax), g2 = nch2 ? pow2_div(nch2, gmax) : 0, gm = g1 > g2 ? g1 : g2;
        const size_t unit = (size_t)gm * d.npx * d.stage_bytes;
        const size_t raw = ((size_t)gm * 32 * d.ey * d.xv + 127) / 128 * 128;
        // at least as many buffers as converter groups in either ring: a group waits on an "empty" barrier by phase
        // parity, which is only unambiguous if it can never be two completions behind
        if (fixed + ZM_NGROUPS * (unit + raw) > limit) continue;
        int nub = ZM_NGROUPS, nraw = ZM_NGROUPS;
        auto fits = [&](int a, int b) { return fixed + a * unit + b * raw <= limit; };
        while (nraw < 6 && fits(nub, nraw + 1)) nraw++;
        if (nub < 4 && fits(nub + 1, nraw)) nub++;
        while (nraw < ZM_MAXRAW && fits(nub, nraw + 1)) nraw++;
        pl.g1 = g1; pl.g2 = g2; pl.nub = nub; pl.nraw = nraw;
        pl.raw_bytes = (int)raw; pl.unit_bytes = (int)unit;
        pl.smem = fixed + nub * unit + nraw * raw + 128;   // + slack to align the raw ring to 128 bytes
        return pl;
    }
    return pl;
}

// ---- tensor maps ------------------------------------------------------------------------------------------------------
typedef CUresult (*ZmEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static ZmEncodeTiled zm_encode_fn()
{
    static ZmEncodeTiled fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return reinterpret_cast<ZmEncodeTiled>(f);
    }();
    return fn;
}
// fp32 volume [B, D, H, W, cstride] (channels [coff, coff + C) of every voxel) as a rank-5 tensor {C, W, H, D, B};
// box = {g * 8 channels, xv voxels, ey rows (every ystep-th row), 1, 1}; out-of-range elements read as zero.
static int zm_make_tmap(CUtensorMap *tm, const float *x, int coff, int C, int cstride, int B, int D, int H, int W, int g, int xv, int ey,
                        int ystep, const char *what)
{
    ZmEncodeTiled enc = zm_encode_fn();
    if (!enc) {
        set_error("%s: cuTensorMapEncodeTiled is not available from this driver", what);
        return MVSB200_E_CUDA;
    }
    const cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
    const cuuint64_t vox = (cuuint64_t)cstride * 4;
    const cuuint64_t gstr[4] = {vox, vox * W, vox * W * H, vox * W * H * D};
    const cuuint32_t box[5] = {(cuuint32_t)(g * 8), (cuuint32_t)xv, (cuuint32_t)(ey * ystep), 1, 1};
    const cuuint32_t estr[5] = {1, 1, (cuuint32_t)ystep, 1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float *>(x + coff), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("%s: cuTensorMapEncodeTiled failed (%d) for a [%d,%d,%d,%d,%d/%d] volume, box %dx%dx%d", what, (int)r, B, D, H, W, C, cstride,
                  g * 8, xv, ey);
        return MVSB200_E_INVALID;
    }
    return MVSB200_OK;
}

static bool zm_shape_ok(const mvsb200_conv3d_desc *d)
{
    const int cin = d->Cin + d->Cin2;
    if (!(d->kd == 3 && d->kh == 3 && d->kw == 3 && d->Cin % 8 == 0 && d->Cin2 % 8 == 0 && cin >= 8)) return false;
    if (!(d->Cout == 8 || (d->Cout % 16 == 0 && d->Cout <= 256))) return false;
    if (!((d->stride == 1 && !d->transposed) || d->stride == 2)) return false;
    return zm_plan(zm_mode(d), zm_ct(d), d->Cin / 8, d->Cin2 / 8).smem > 0;
}

template <int MODE, int CT>
static int launch_zm(ZmParams p, int sm_count, cudaStream_t st)
{
    using T = ZmCfg<MODE, CT>;
    const ZmPlan plan = zm_plan(MODE, CT, p.Cin1 / 8, p.Cin2 / 8);
    p.nstages = plan.nub;
    p.g1 = plan.g1; p.g2 = plan.g2 ? plan.g2 : 1;
    p.nraw = plan.nraw; p.raw_bytes = plan.raw_bytes; p.unit_bytes = plan.unit_bytes;
    const size_t smem = plan.smem;
    const ZmDims dm = zm_dims(MODE, CT);
    CUtensorMap tm_x, tm_x2;
    const int ystep = (MODE == ZM_S2) ? 2 : 1;
    if (int rc = zm_make_tmap(&tm_x, p.x, p.x_coff, p.Cin1, p.x_cstride, p.B, p.D, p.H, p.W, plan.g1, dm.xv, dm.ey, ystep, "conv3d_zm")) return rc;
    tm_x2 = tm_x;
    if (p.x2)
        if (int rc = zm_make_tmap(&tm_x2, p.x2, 0, p.Cin2, p.Cin2, p.B, p.D, p.H, p.W, plan.g2, dm.xv, dm.ey, ystep, "conv3d_zm")) return rc;
    const int nzt = (MODE == ZM_DECONV) ? p.D : p.Do, ny = (MODE == ZM_DECONV) ? p.H : p.Ho, nx = (MODE == ZM_DECONV) ? p.W : p.Wo;
    p.tiles_y = (ny + T::TY - 1) / T::TY;
    p.tiles_x = (nx + T::TX - 1) / T::TX;
    const int nblocks = (p.Cout + CT - 1) / CT;
    const int classes = (MODE == ZM_DECONV) ? 4 : 1;
    int ctas = sm_count / nblocks;
    if (ctas < classes) ctas = classes;
    ctas = ctas / classes * classes;                    // CTAs per class x classes (blockIdx.x % 4 is the class)
    const long long total = (long long)p.tiles_x * p.tiles_y * p.B * nzt;   // planes of one class
    if (total * classes >= (1ll << 30)) {
        set_error("conv3d_zm: volume too large");
        return MVSB200_E_INVALID;
    }
    // lockstep candidate: segment count that best trades the z halo against filling all CTAs for whole rounds
    const long long ncol = (long long)p.tiles_x * p.tiles_y * p.B;
    const int cpc = ctas / classes;                    // CTAs per class
    const int halo = (MODE == ZM_S1) ? 2 : 1;
    double best = -1.0;
    int best_nseg = 1;
    for (int nseg = 1; nseg <= nzt; nseg++) {
        const int zseg = (nzt + nseg - 1) / nseg;
        if ((long long)(nseg - 1) * zseg >= nzt) continue;
        const long long tiles = ncol * nseg;
        const long long rounds = (tiles + cpc - 1) / cpc;
        const double eff = (double)tiles / (double)(rounds * cpc) * (double)zseg / (double)(zseg + halo) *
                           (1.0 - 0.02 * (double)rounds / (double)(rounds + 8));
        if (eff > best) { best = eff; best_nseg = nseg; }
        if (tiles > 64ll * cpc) break;
    }
    long long used;
    // lockstep only pays for inputs that do not fit L2 (B200: 126 MB; halo re-reads of a smaller volume are L2 hits under
    // either scheme) and only if its tiling keeps the CTAs busy
    const double in_bytes = (double)p.B * p.D * p.H * p.W * (p.Cin1 + p.Cin2) * 4.0;
    if (best >= 0.88 && in_bytes >= 64e6) {
        p.nseg = best_nseg;
        p.per = (nzt + best_nseg - 1) / best_nseg;
        p.total = (int)(ncol * best_nseg);
        used = (p.total < cpc ? p.total : cpc) * (long long)classes;
    } else {
        // flat: equal runs of planes per CTA; a run shorter than 4 planes would spend most of its time on the z halo and the
        // pipeline fill of its segment, so small volumes use fewer CTAs instead
        long long per = (total + cpc - 1) / cpc;
        const long long per_min = nzt < 4 ? nzt : 4;
        if (per < per_min) per = per_min;
        p.nseg = 0;
        p.per = (int)per;
        p.total = (int)total;
        used = (total + per - 1) / per * classes;
    }
    if (int rc = ensure_dynamic_smem(k2_conv3d_zm_kernel<MODE, CT>, smem, "conv3d_zm")) return rc;
    dim3 grid((unsigned)used, (unsigned)nblocks, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(ZM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = p.pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k2_conv3d_zm_kernel<MODE, CT>, p, tm_x, tm_x2);
    if (e != cudaSuccess) {
        set_error("k2_conv3d_zm_kernel: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return MVSB200_E_CUDA;
    }
    return check_launch("k2_conv3d_zm_kernel");
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_absmax(const float *x, long long n, float *amax, mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(x && amax && n >= 0, "absmax: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(amax, 0, sizeof(float), st);
    if (e != cudaSuccess) {
        set_error("absmax: cudaMemsetAsync: %s", cudaGetErrorString(e));
        return MVSB200_E_CUDA;
    }
    if (n == 0) return MVSB200_OK;
    MVSB200_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "absmax: x must be 16-byte aligned");
    long long blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    k2_zm_absmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n, amax);
    return check_launch("k2_zm_absmax_kernel");
}

// Debug aid (not part of the drop-in surface): enable != 0 switches the in-kernel cycle accounting on and zeroes the
// counters; with `out` non-null the 32 counters are copied to host memory (synchronises the device).
extern "C" MVSB200_API int mvsb200_debug_zm_profile(int enable, unsigned long long *out)
{
    if (out && cudaMemcpyFromSymbol(out, g_zm_prof, sizeof(unsigned long long) * 32) != cudaSuccess) {
        set_error("debug_zm_profile: cudaMemcpyFromSymbol failed");
        return MVSB200_E_CUDA;
    }
    g_zm_prof_on = enable;
    if (enable) {
        static const unsigned long long zeros[32] = {0};
        if (cudaMemcpyToSymbol(g_zm_prof, zeros, sizeof(zeros)) != cudaSuccess) {
            set_error("debug_zm_profile: cudaMemcpyToSymbol failed");
            return MVSB200_E_CUDA;
        }
    }
    return MVSB200_OK;
}

extern "C" int mvsb200_conv3d_zm_supported(const mvsb200_conv3d_desc *d)
{
    return d && zm_shape_ok(d) ? 1 : 0;
}

extern "C" long long mvsb200_conv3d_zm_packed_bytes(const mvsb200_conv3d_desc *d)
{
    if (!d || !zm_shape_ok(d)) return 0;
    const int ct = zm_ct(d), nblocks = (d->Cout + ct - 1) / ct, nch = (d->Cin + d->Cin2) / 8, mode = zm_mode(d);
    const int nc = (mode == ZM_S1 ? 3 : 2) * 2 * ct;
    return 16 + (long long)nblocks * zm_nvar(mode) * nch * 9 * (2 * nc * 16);
}

extern "C" int mvsb200_conv3d_zm_pack(const mvsb200_conv3d_desc *d, const float *w, void *packed, mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(d && w && packed, "conv3d_zm_pack: null pointer");
    MVSB200_REQUIRE(zm_shape_ok(d), "conv3d_zm_pack: layer not supported by the z-march engine");
    MVSB200_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "conv3d_zm_pack: packed must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    ZmPackParams p;
    p.w = w; p.wp = reinterpret_cast<__half *>(packed);
    p.Cin = d->Cin + d->Cin2; p.Cout = d->Cout; p.nch = p.Cin / 8;
    p.mode = zm_mode(d); p.CT = zm_ct(d); p.nvar = zm_nvar(p.mode);
    p.NC = (p.mode == ZM_S1 ? 3 : 2) * 2 * p.CT;
    const int nblocks = (d->Cout + p.CT - 1) / p.CT;
    int rc = mvsb200_absmax(w, 27ll * p.Cin * p.Cout, reinterpret_cast<float *>(packed), stream);
    if (rc) return rc;
    k2_zm_header_kernel<<<1, 1, 0, st>>>(reinterpret_cast<float *>(packed));
    k2_zm_pack_kernel<<<nblocks * p.nvar * p.nch * (p.mode == ZM_S2 ? 2 : 1) * 3, 256, 0, st>>>(p);
    return check_launch("k2_zm_pack_kernel");
}

extern "C" int mvsb200_conv3d_zm(const mvsb200_conv3d_desc *d, const float *x, const float *x2, const void *packed,
                                 const float *scale, const float *bias, const float *skip, float *y,
                                 const float *x_amax, const float *x2_amax, float *y_amax, mvsb200_stream_t stream)
{
    return mvsb200_conv3d_zm_slice(d, x, d ? d->Cin : 0, 0, x2, packed, scale, bias, skip, y, x_amax, x2_amax, y_amax, stream);
}

extern "C" int mvsb200_conv3d_zm_slice(const mvsb200_conv3d_desc *d, const float *x, int x_channels, int x_first_channel,
                                       const float *x2, const void *packed, const float *scale, const float *bias,
                                       const float *skip, float *y, const float *x_amax, const float *x2_amax, float *y_amax,
                                       mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(d && x && packed && y && x_amax, "conv3d_zm: null pointer");
    MVSB200_REQUIRE(x_first_channel >= 0 && x_first_channel % 4 == 0 && x_channels % 4 == 0 && x_first_channel + d->Cin <= x_channels,
                    "conv3d_zm: channel slice [%d, %d) of %d channels", x_first_channel, x_first_channel + d->Cin, x_channels);
    MVSB200_REQUIRE(zm_shape_ok(d), "conv3d_zm: layer not supported by the z-march engine");
    MVSB200_REQUIRE(d->B > 0 && d->D > 0 && d->H > 0 && d->W > 0, "conv3d_zm: bad shape B=%d D=%d H=%d W=%d", d->B, d->D, d->H, d->W);
    MVSB200_REQUIRE(d->Cin2 == 0 || (x2 && x2_amax), "conv3d_zm: Cin2=%d but x2 / x2_amax is null", d->Cin2);
    MVSB200_REQUIRE(d->skip_mode >= 0 && d->skip_mode <= 2, "conv3d_zm: skip_mode=%d", d->skip_mode);
    MVSB200_REQUIRE(d->skip_mode == MVSB200_SKIP_NONE || skip, "conv3d_zm: skip_mode=%d but skip is null", d->skip_mode);
    ZmParams p;
    int rc = mvsb200_conv3d_out_shape(d, &p.Do, &p.Ho, &p.Wo);
    if (rc) return rc;
    MVSB200_REQUIRE(p.Do > 0 && p.Ho > 0 && p.Wo > 0, "conv3d_zm: empty output");
    p.x = x; p.x2 = d->Cin2 ? x2 : nullptr; p.wp = reinterpret_cast<const __half *>(packed);
    p.scale = scale; p.bias = bias; p.skip = skip; p.y = y;
    p.x_amax = x_amax; p.x2_amax = x2_amax; p.y_amax = y_amax;
    p.B = d->B; p.D = d->D; p.H = d->H; p.W = d->W;
    p.Cin1 = d->Cin; p.Cin2 = d->Cin2; p.Cout = d->Cout;
    p.relu = d->relu; p.skip_mode = d->skip_mode;
    p.tiles_x = p.tiles_y = p.per = p.total = p.nseg = p.nstages = 0;
    p.g1 = p.g2 = p.nraw = p.raw_bytes = p.unit_bytes = 0;
    p.x_cstride = x_channels; p.x_coff = x_first_channel;
    p.profile = g_zm_prof_on;
    p.wide = (d->Cout % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 31) == 0 && (!skip || (reinterpret_cast<uintptr_t>(skip) & 31) == 0)) ? 1 : 0;
    p.pdl = d->static_params ? 1 : 0;
    int sm_count = 0, dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0) sm_count = 148;
    cudaStream_t st = (cudaStream_t)stream;
    const int mode = zm_mode(d), ct = zm_ct(d);
    if (mode == ZM_S1) return ct == 8 ? launch_zm<ZM_S1, 8>(p, sm_count, st) : launch_zm<ZM_S1, 16>(p, sm_count, st);
    if (mode == ZM_S2) return ct == 8 ? launch_zm<ZM_S2, 8>(p, sm_count, st) : launch_zm<ZM_S2, 16>(p, sm_count, st);
    return ct == 8 ? launch_zm<ZM_DECONV, 8>(p, sm_count, st) : launch_zm<ZM_DECONV, 16>(p, sm_count, st);
}
