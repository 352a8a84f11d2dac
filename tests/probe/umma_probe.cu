// GPU probe (not a pytest): pins down the tcgen05 shared-memory descriptor semantics libmvsb200 relies on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I wild_deep_mvs_b200/csrc -o .scratch/umma_probe tests/probe/umma_probe.cu
// For each variant it runs D[128x16] = sum over `nacc` MMAs of A[window] * B^T and prints the max abs error.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "umma.cuh"
using namespace mvsb200::umma;

constexpr int R = 512;   // staged rows
constexpr int N = 16;

struct Cfg { int swap_lbo_sbo; int shift0; int shift1; int nacc; int col0; };

__global__ void probe_kernel(const float *A, const float *B, float *D, Cfg c)
{
    __shared__ __align__(128) float sA[2 * R * 4];   // [chunk][row][4]
    __shared__ __align__(128) float sB[2 * N * 4];   // [chunk][n][4]
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < R * 8; i += blockDim.x) { int r = i / 8, k = i % 8; sA[(k / 4) * R * 4 + r * 4 + (k % 4)] = A[i]; }
    for (int i = tid; i < N * 8; i += blockDim.x) { int n = i / 8, k = i % 8; sB[(k / 4) * N * 4 + n * 4 + (k % 4)] = B[i]; }
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (tid < 32) tmem_alloc(smem_u32(&tmem_slot), 64);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = idesc_tf32(128, N);
        uint32_t lboA = R * 16, sbo = 128, lboB = N * 16;
        for (int i = 0; i < c.nacc; i++) {
            int s = (i == 0) ? c.shift0 : c.shift1;
            uint64_t da = c.swap_lbo_sbo ? smem_desc(smem_u32(sA) + s * 16, sbo, lboA) : smem_desc(smem_u32(sA) + s * 16, lboA, sbo);
            uint64_t db = c.swap_lbo_sbo ? smem_desc(smem_u32(sB), sbo, lboB) : smem_desc(smem_u32(sB), lboB, sbo);
            mma_tf32(tmem + c.col0, da, db, idesc, i > 0);
        }
        mma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after_sync();
    float v[16];
    const int warp = tid / 32, lane = tid % 32;
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c.col0, v);
    tmem_ld_wait();
    for (int n = 0; n < N; n++) D[(warp * 32 + lane) * N + n] = v[n];
    tc_fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 64);
}

int main()
{
    std::vector<float> A(R * 8), B(N * 8), D(128 * N);
    for (int r = 0; r < R; r++) for (int k = 0; k < 8; k++) A[r * 8 + k] = (float)((r * 7 + k * 3) % 31 - 15) * 0.25f;
    for (int n = 0; n < N; n++) for (int k = 0; k < 8; k++) B[n * 8 + k] = (float)((n * 5 + k) % 13 - 6) * 0.5f;
    // rounding probe: row 300 k=0 has sub-tf32 bits
    A[300 * 8 + 0] = 1.0f + ldexpf(1.f, -11) + ldexpf(1.f, -12);
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    Cfg cfgs[] = {{0, 0, 0, 1, 0}, {0, 1, 0, 1, 0}, {0, 3, 0, 1, 0}, {0, 37, 0, 1, 16}, {0, 5, 200, 2, 32}, {0, 300, 0, 1, 0}};
    for (auto &c : cfgs) {
        cudaMemset(dD, 0, D.size() * 4);
        probe_kernel<<<1, 128>>>(dA, dB, dD, c);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant swap=%d shift=%d/%d nacc=%d: CUDA error %s\n", c.swap_lbo_sbo, c.shift0, c.shift1, c.nacc, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; m++) for (int n = 0; n < N; n++) {
            double want = 0;
            for (int i = 0; i < c.nacc; i++) { int s = (i == 0) ? c.shift0 : c.shift1; for (int k = 0; k < 8; k++) want += (double)A[(s + m) * 8 + k] * B[n * 8 + k]; }
            maxerr = fmax(maxerr, fabs(want - D[m * N + n]));
        }
        printf("variant swap=%d shift=%d/%d nacc=%d col0=%d: max abs err %.6g   D[0][0..3]=%g %g %g %g\n", c.swap_lbo_sbo, c.shift0, c.shift1, c.nacc, c.col0, maxerr, D[0], D[1], D[2], D[3]);
        if (c.shift0 == 300) {
            // row 0 of the window is staged row 300: D[0][n] includes a*B[n][0]; isolate the rounding behaviour
            double t = 0, r = 0; float a = A[300 * 8];
            float at = 1.0f, ar = 1.0f + ldexpf(1.f, -10);
            for (int k = 1; k < 8; k++) { t += (double)A[300 * 8 + k] * B[1 * 8 + k]; }
            r = t + (double)ar * B[1 * 8]; t += (double)at * B[1 * 8];
            printf("  tf32 conversion of %.9g: D=%.9g  (truncate -> %.9g, round-nearest -> %.9g)\n", a, D[1], t, r);
        }
    }
    printf("probe done\n");
    return 0;
}
