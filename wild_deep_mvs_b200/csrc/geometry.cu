// Error plumbing, device query and the two geometry prologue kernels of mvsb200.h.
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace mvsb200 {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void clear_error() { g_err[0] = 0; }

int ensure_dynamic_smem_impl(const void *kernel, size_t bytes, const char *what)
{
    struct Entry { const void *kernel; int dev; size_t bytes; };
    static Entry table[256];
    static int n = 0;
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    Entry *e = nullptr;
    for (int i = 0; i < n; i++)
        if (table[i].kernel == kernel && table[i].dev == dev) e = &table[i];
    if (e && e->bytes >= bytes) return MVSB200_OK;
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (err != cudaSuccess) {
        set_error("%s: cudaFuncSetAttribute(%zu bytes): %s", what, bytes, cudaGetErrorString(err));
        return MVSB200_E_CUDA;
    }
    if (!e && n < 256) e = &table[n++];
    if (e) { e->kernel = kernel; e->dev = dev; e->bytes = bytes; }
    return MVSB200_OK;
}

// ---- small dense fp64 helpers (one thread does a whole camera pair) -----------------------------

__device__ bool inverse4(const double *a, double *out)
{
    double m[4][8];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            m[i][j] = a[i * 4 + j];
            m[i][4 + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < 4; c++) {
        int p = c;
        for (int r = c + 1; r < 4; r++)
            if (fabs(m[r][c]) > fabs(m[p][c])) p = r;
        if (m[p][c] == 0.0) return false;
        if (p != c)
            for (int j = 0; j < 8; j++) {
                double t = m[c][j];
                m[c][j] = m[p][j];
                m[p][j] = t;
            }
        double piv = 1.0 / m[c][c];
        for (int j = 0; j < 8; j++) m[c][j] *= piv;
        for (int r = 0; r < 4; r++) {
            if (r == c) continue;
            double f = m[r][c];
            for (int j = 0; j < 8; j++) m[r][j] -= f * m[c][j];
        }
    }
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) out[i * 4 + j] = m[i][4 + j];
    return true;
}

__device__ void mul3(const double *a, const double *b, double *o)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) o[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}

__device__ bool inverse3(const double *a, double *o)
{
    double c0 = a[4] * a[8] - a[5] * a[7], c1 = a[5] * a[6] - a[3] * a[8], c2 = a[3] * a[7] - a[4] * a[6];
    double det = a[0] * c0 + a[1] * c1 + a[2] * c2;
    if (det == 0.0) return false;
    double id = 1.0 / det;
    o[0] = c0 * id;
    o[1] = (a[2] * a[7] - a[1] * a[8]) * id;
    o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    o[3] = c1 * id;
    o[4] = (a[0] * a[8] - a[2] * a[6]) * id;
    o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    o[6] = c2 * id;
    o[7] = (a[1] * a[6] - a[0] * a[7]) * id;
    o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
    return true;
}

// proj = src_proj @ inv(ref_proj); a singular reference camera yields NaNs (as torch.inverse would raise;
// the host wrapper checks finiteness only in debug mode to stay sync-free).
__global__ void mvs_relative_proj_kernel(const float *__restrict__ ref_proj, const float *__restrict__ src_proj,
                                         float *__restrict__ warp, int B, int S)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * S) return;
    int b = i / S;
    double r[16], s[16], inv[16];
    for (int k = 0; k < 16; k++) {
        r[k] = (double)ref_proj[b * 16 + k];
        s[k] = (double)src_proj[(long long)i * 16 + k];
    }
    bool ok = inverse4(r, inv);
    float *o = warp + (long long)i * 16;
    const float nanv = __int_as_float(0x7fc00000);
    for (int row = 0; row < 3; row++)
        for (int col = 0; col < 4; col++) {
            double acc = 0.0;
            for (int k = 0; k < 4; k++) acc += s[row * 4 + k] * inv[k * 4 + col];
            float v = ok ? (float)acc : nanv;
            if (col < 3) o[row * 3 + col] = v;
            else o[9 + row] = v;
        }
    o[12] = o[13] = o[14] = o[15] = 0.f;
}

__global__ void vis_homography_params_kernel(const float *__restrict__ ref_cam, const float *__restrict__ src_cam,
                                             float scale, float *__restrict__ warp, int B, int S)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * S) return;
    int b = i / S;
    const float *rc = ref_cam + (long long)b * 32;
    const float *sc = src_cam + (long long)i * 32;
    double Rr[9], Rs[9], Kr[9], Ks[9], tr[3], ts[3];
    for (int a = 0; a < 3; a++) {
        for (int c = 0; c < 3; c++) {
            Rr[a * 3 + c] = rc[a * 4 + c];
            Rs[a * 3 + c] = sc[a * 4 + c];
            Kr[a * 3 + c] = rc[16 + a * 4 + c];
            Ks[a * 3 + c] = sc[16 + a * 4 + c];
        }
        tr[a] = rc[a * 4 + 3];
        ts[a] = sc[a * 4 + 3];
    }
    // scale_camera multiplies fx, fy, cx, cy in fp32 (preproc.py:76-83)
    Kr[0] = (double)((float)Kr[0] * scale); Kr[4] = (double)((float)Kr[4] * scale);
    Kr[2] = (double)((float)Kr[2] * scale); Kr[5] = (double)((float)Kr[5] * scale);
    Ks[0] = (double)((float)Ks[0] * scale); Ks[4] = (double)((float)Ks[4] * scale);
    Ks[2] = (double)((float)Ks[2] * scale); Ks[5] = (double)((float)Ks[5] * scale);
    double Kri[9];
    bool ok = inverse3(Kr, Kri);
    double RrT[9];
    for (int a = 0; a < 3; a++)
        for (int c = 0; c < 3; c++) RrT[a * 3 + c] = Rr[c * 3 + a];
    double cr[3], cs[3], crel[3];
    for (int a = 0; a < 3; a++) {
        cr[a] = -(Rr[a] * tr[0] + Rr[3 + a] * tr[1] + Rr[6 + a] * tr[2]);
        cs[a] = -(Rs[a] * ts[0] + Rs[3 + a] * ts[1] + Rs[6 + a] * ts[2]);
        crel[a] = cs[a] - cr[a];
    }
    double M[9], KsRs[9], A[9];
    mul3(RrT, Kri, M);   // R_r^T K_r^-1
    mul3(Ks, Rs, KsRs);  // K_s R_s
    mul3(KsRs, M, A);
    float *o = warp + (long long)i * 16;
    const float nanv = __int_as_float(0x7fc00000);
    for (int k = 0; k < 9; k++) o[k] = ok ? (float)A[k] : nanv;
    for (int a = 0; a < 3; a++) {
        double bb = KsRs[a * 3] * crel[0] + KsRs[a * 3 + 1] * crel[1] + KsRs[a * 3 + 2] * crel[2];
        double nn = Rr[6] * M[a] + Rr[7] * M[3 + a] + Rr[8] * M[6 + a];  // fronto direction R_r[2,:] times M
        o[9 + a] = ok ? (float)bb : nanv;
        o[12 + a] = ok ? (float)nn : nanv;
    }
    o[15] = 0.f;
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" {

int mvsb200_abi_version(void) { return MVSB200_ABI_VERSION; }

const char *mvsb200_last_error(void) { return g_err; }

int mvsb200_device_info(int *sm_count, int *cc_major, int *cc_minor)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("cudaGetDevice: %s", cudaGetErrorString(e));
        return MVSB200_E_NODEVICE;
    }
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) {
        set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
        return MVSB200_E_CUDA;
    }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return MVSB200_OK;
}

int mvsb200_mvs_relative_proj(const float *ref_proj, const float *src_proj, float *warp, int B, int S,
                              mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(ref_proj && src_proj && warp, "mvs_relative_proj: null pointer");
    MVSB200_REQUIRE(B > 0 && S > 0, "mvs_relative_proj: B=%d S=%d", B, S);
    int n = B * S;
    mvs_relative_proj_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(ref_proj, src_proj, warp, B, S);
    return check_launch("mvs_relative_proj");
}

int mvsb200_vis_homography_params(const float *ref_cam, const float *src_cam, float scale, float *warp, int B,
                                  int S, mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(ref_cam && src_cam && warp, "vis_homography_params: null pointer");
    MVSB200_REQUIRE(B > 0 && S > 0 && scale > 0.f, "vis_homography_params: B=%d S=%d scale=%g", B, S, scale);
    int n = B * S;
    vis_homography_params_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(ref_cam, src_cam, scale, warp, B, S);
    return check_launch("vis_homography_params");
}

}  // extern "C"
