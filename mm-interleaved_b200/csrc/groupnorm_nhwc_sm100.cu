// groupnorm_nhwc_sm100.cu -- GroupNorm (+ optional SiLU) on NHWC activations, the normalisation in front of every
// UNet convolution of the denoise step (diffusers ResnetBlock2D / Transformer2DModel / conv_norm_out, called from
// the reference's patched forward, utils/monkey_patch/sd_unet_forward_monkey_patch.py:235-366).
//
// Why it exists: torch's CUDA group_norm always returns an NCHW tensor, so a channels-last UNet pays a layout
// round trip around every convolution; this kernel keeps the activations NHWC for the implicit-GEMM convolution
// (conv_igemm_sm100.cu) and folds the SiLU in.  Pure bandwidth work: two passes (statistics, apply), the second
// one hitting L2 for the map sizes of SD (<= 42 MB per tensor at batch 16).
//
// Thread mapping (both kernels): thread t owns the 16-byte channel vector t % (C/VEC) of every k-th pixel, so the
// per-channel scale/shift (apply) and the partial sums (statistics) live in registers and consecutive threads
// read consecutive 16-byte vectors of a pixel row (coalesced).  Statistics are fp32 sum / sum-of-squares, reduced
// per group through shared then global atomics; the bf16 pipeline's rounding points (norm -> T, silu -> T) are
// kept so a bf16 run tracks diffusers' bf16 run.
#include "common.cuh"

namespace mmfs {

template <typename T> __device__ __forceinline__ float gn_rnd(float x) { return to_op(from_op<T>(x)); }
template <> __device__ __forceinline__ float gn_rnd<float>(float x) { return x; }

template <typename T>
__global__ void __launch_bounds__(1024) gn_stats_kernel(const T *__restrict__ x, float *__restrict__ stats, int HW, int C,
                                                         int G, int ppb, int cvecs, int lanes) {
    constexpr int VEC = 16 / (int)sizeof(T);
    extern __shared__ float s_acc[];                     // [G][2]
    const int b = blockIdx.y, p0 = blockIdx.x * ppb, p1 = min(HW, p0 + ppb);
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    const int cv = threadIdx.x % cvecs, pl = threadIdx.x / cvecs;
    float s[VEC], ss[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) s[k] = ss[k] = 0.f;
    const T *base = x + ((size_t)b * HW) * C + (size_t)cv * VEC;
    if (pl < lanes) {
        int p = p0 + pl;
        for (; p + 3 * lanes < p1; p += 4 * lanes) {     // four independent 16-byte loads in flight
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ldg_nc_v4(base + (size_t)(p + u * lanes) * C);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float f[VEC];
                Vec16<T>::unpack(v[u], f);
#pragma unroll
                for (int k = 0; k < VEC; ++k) { s[k] += f[k]; ss[k] = fmaf(f[k], f[k], ss[k]); }
            }
        }
        for (; p < p1; p += lanes) {
            float f[VEC];
            Vec16<T>::unpack(ldg_nc_v4(base + (size_t)p * C), f);
#pragma unroll
            for (int k = 0; k < VEC; ++k) { s[k] += f[k]; ss[k] = fmaf(f[k], f[k], ss[k]); }
        }
        const int cg = C / G;
        int g = (cv * VEC) / cg;
        float a = 0.f, q = 0.f;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {                  // flush run-wise: a vector spans at most a few groups
            const int gk = (cv * VEC + k) / cg;
            if (gk != g) { atomicAdd(&s_acc[2 * g], a); atomicAdd(&s_acc[2 * g + 1], q); a = q = 0.f; g = gk; }
            a += s[k]; q += ss[k];
        }
        atomicAdd(&s_acc[2 * g], a); atomicAdd(&s_acc[2 * g + 1], q);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(&stats[(size_t)b * 2 * G + i], s_acc[i]);
}

template <typename T>
__global__ void __launch_bounds__(1024) gn_apply_kernel(const T *__restrict__ x, const T *__restrict__ gamma,
                                                         const T *__restrict__ beta, const float *__restrict__ stats,
                                                         T *__restrict__ y, int HW, int C, int G, int ppb, int cvecs, int lanes,
                                                         float eps, int silu) {
    constexpr int VEC = 16 / (int)sizeof(T);
    const int b = blockIdx.y, p0 = blockIdx.x * ppb, p1 = min(HW, p0 + ppb);
    const int cv = threadIdx.x % cvecs, pl = threadIdx.x / cvecs;
    if (pl >= lanes) return;
    const int cg = C / G;
    const float inv_n = 1.f / ((float)cg * (float)HW);
    float sc[VEC], sh[VEC];
    {
        float gm[VEC], bt[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) { gm[k] = 1.f; bt[k] = 0.f; }
        if (gamma) Vec16<T>::unpack(*reinterpret_cast<const uint4 *>(gamma + cv * VEC), gm);
        if (beta) Vec16<T>::unpack(*reinterpret_cast<const uint4 *>(beta + cv * VEC), bt);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const int g = (cv * VEC + k) / cg;
            const float mean = stats[((size_t)b * G + g) * 2] * inv_n;
            const float var = fmaxf(stats[((size_t)b * G + g) * 2 + 1] * inv_n - mean * mean, 0.f);
            const float r = rsqrtf(var + eps);
            sc[k] = r * gm[k];
            sh[k] = bt[k] - mean * sc[k];
        }
    }
    const size_t off = ((size_t)b * HW) * C + (size_t)cv * VEC;
    const T *xb = x + off;
    T *yb = y + off;
    auto one = [&](const uint4 &v, size_t p) {
        float f[VEC], o[VEC];
        Vec16<T>::unpack(v, f);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            float n = gn_rnd<T>(fmaf(f[k], sc[k], sh[k]));
            if (silu) n = __fdividef(n, 1.f + __expf(-n));
            o[k] = n;
        }
        stg_v4(yb + p * C, Vec16<T>::pack(o));
    };
    int p = p0 + pl;
    for (; p + 3 * lanes < p1; p += 4 * lanes) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ldg_nc_v4(xb + (size_t)(p + u * lanes) * C);
#pragma unroll
        for (int u = 0; u < 4; ++u) one(v[u], (size_t)(p + u * lanes));
    }
    for (; p < p1; p += lanes) one(ldg_nc_v4(xb + (size_t)p * C), (size_t)p);
}

template <typename T>
static int gn_launch(const void *x, const void *gamma, const void *beta, void *y, float *stats, int B, int HW, int C, int G,
                     float eps, int silu, cudaStream_t st) {
    constexpr int VEC = 16 / (int)sizeof(T);
    const int cvecs = C / VEC;
    const int lanes = max(1, min(HW, 512 / cvecs));
    const int threads = cvecs * lanes;
    const int target_blocks = max(1, (num_sms() * 8) / B);               // ~8 CTAs' worth of work per SM over the batch
    int ppb = max(lanes * 4, (HW + target_blocks - 1) / target_blocks);
    ppb = min(ppb, HW);
    const int chunks = (HW + ppb - 1) / ppb;
    MMFS_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * (size_t)B * G, st));
    dim3 grid(chunks, B);
    gn_stats_kernel<T><<<grid, threads, 2 * G * sizeof(float), st>>>((const T *)x, stats, HW, C, G, ppb, cvecs, lanes);
    gn_apply_kernel<T><<<grid, threads, 0, st>>>((const T *)x, (const T *)gamma, (const T *)beta, stats, (T *)y, HW, C, G, ppb,
                                                 cvecs, lanes, eps, silu);
    MMFS_CUDA(cudaGetLastError());
    return MMFS_OK;
}

}  // namespace mmfs

extern "C" int mmfs_groupnorm_nhwc(const void *x, const void *gamma, const void *beta, void *y, float *stats, int B, int HW,
                                   int C, int G, float eps, int silu, int dtype, void *stream) {
    using namespace mmfs;
    MMFS_CHECK_ARG(x && y && stats, "groupnorm: null pointer");
    MMFS_CHECK_ARG(B > 0 && HW > 0 && C > 0 && G > 0 && C % G == 0 && B <= 65535, "groupnorm: bad sizes B=%d HW=%d C=%d G=%d", B, HW, C, G);
    const size_t es = dtype_size(dtype);
    if (dtype == MMFS_F64 || es == 0 || C % (16 / (int)es) != 0 || C / (16 / (int)es) > 1024) {
        set_error("groupnorm: unsupported dtype %d / channel count %d (need C %% (16/sizeof) == 0, C*sizeof <= 16 KiB)", dtype, C);
        return MMFS_EUNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case MMFS_F32: return gn_launch<float>(x, gamma, beta, y, stats, B, HW, C, G, eps, silu, st);
        case MMFS_F16: return gn_launch<__half>(x, gamma, beta, y, stats, B, HW, C, G, eps, silu, st);
        default: return gn_launch<__nv_bfloat16>(x, gamma, beta, y, stats, B, HW, C, G, eps, silu, st);
    }
}
