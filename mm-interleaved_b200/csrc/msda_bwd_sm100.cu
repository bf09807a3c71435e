// msda_bwd_sm100.cu -- multi-scale deformable attention BACKWARD for sm_100a (training path).
//
// Replaces ms_deform_attn_cuda_backward (ops/src/cuda/ms_deform_attn_cuda.cu:84-166) and the nine col2im kernels
// of ops/src/cuda/ms_deform_im2col_cuda.cuh:304-923 (+ launcher :959-1330) with ONE kernel that reuses the forward's
// machinery (sampler_common.cuh): a warp per (b, q, m) row, index math once per point, 16-byte gathers.
//   grad_value        scatter: red.global.add.v4.f32 of go[c] * (lerp_k * a) into an fp32 buffer -- like the reference,
//                     atomics => summation order is not deterministic (cuh:128-155); the reference too accumulates
//                     half inputs in fp32 and casts back (cu:122-129, 156-160; done by the Python shim here)
//   grad_sampling_loc one writer per (b,q,m,l,p): a*W*(-hh d1 + hh d2 - lh d3 + lh d4), a*H*(-hw d1 - lw d2 + hw d3 + lw d4)
//   grad_attn_weight  one writer: sum_k lerp_k d_k         with d_k = <grad_out row, value row of corner k> (cuh:90-162)
// The channel dot products d_k are reduced with warp shuffles inside each slot and handed to the point's lane
// through shared memory; no block-size-specialised reduction variants are needed.
#include "sampler_common.cuh"

namespace mmfs {

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

constexpr int kDotStride = 33;

template <typename T, int D>
__global__ void __launch_bounds__(32 * kWarpsPerCta, 3)
msda_bwd_rows_kernel(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ starts,
                     const T *__restrict__ loc, const T *__restrict__ attn, const T *__restrict__ grad_out,
                     float *__restrict__ grad_value, float *__restrict__ grad_loc, float *__restrict__ grad_attn,
                     int S, int M, int L, int Lq, int P, int p_shift, int rows_per_warp, int qtiles, long ntiles,
                     int ctas_per_sm, int nsm) {
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int LPR = D / VEC;
    constexpr int RPI = 32 / LPR;
    constexpr int NIT = 128 / RPI;
    extern __shared__ int4 s_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int4 *s_lvl = s_dyn;
    const int per_warp_bytes = kTapsPerWarp * (int)sizeof(Tap) + 4 * kDotStride * 4 + 16;
    char *wbase = reinterpret_cast<char *>(s_dyn + L) + warp * (((per_warp_bytes + 15) / 16) * 16);
    Tap *taps = reinterpret_cast<Tap *>(wbase);
    float *dots = reinterpret_cast<float *>(wbase + kTapsPerWarp * sizeof(Tap));   // [corner][point] skewed
    for (int l = threadIdx.x; l < L; l += blockDim.x)
        s_lvl[l] = make_int4((int)shapes[2 * l], (int)shapes[2 * l + 1], (int)starts[l], 0);
    __syncthreads();

    const int LP = L * P;
    const long long row_bytes = (long long)M * D * (int)sizeof(T);
    const int slot = lane / LPR, sub = lane % LPR;
    RowWalk walk;
    walk.itiles = (int)ntiles; walk.igrid = (int)gridDim.x; walk.qtiles = qtiles; walk.M = M; walk.Lq = Lq;
    walk.rows_per_warp = rows_per_warp; walk.warp = warp;

    for (RowCursor cur = walk.first(ctas_per_sm, nsm, 1); cur.ok; cur = walk.next(cur)) {
        const int b = cur.b, m = cur.m, q = cur.q;
        const size_t qm = ((size_t)b * Lq + q) * M + m;
        const T *locp = loc + qm * (size_t)LP * 2;
        const T *attp = attn + qm * (size_t)LP;
        const char *slab = reinterpret_cast<const char *>(value + ((size_t)b * S * M + m) * D);
        const char *vbase = slab + sub * 16;
        float *gv_base = grad_value + ((size_t)b * S * M + m) * D + sub * VEC;   // fp32 twin of the value slab
        const long long zero_off = reinterpret_cast<const char *>(g_zero_row) - slab;
        float go[VEC];   // this lane's channels of the incoming gradient row (same for every slot)
        Vec16<T>::unpack(*reinterpret_cast<const uint4 *>(grad_out + qm * D + sub * VEC), go);

        for (int p0 = 0; p0 < LP; p0 += 32) {
            const int j = p0 + lane;
            bool live = false;
            PointGeom<float> g;
            g.in_range = false; g.h_low = g.w_low = 0; g.lh = g.lw = 0.f;
            float a = 0.f;
            int4 lv = make_int4(1, 1, 0, 0);
            if (j < LP) {
                a = to_op(attp[j]);
                const float x = to_op(locp[2 * j]), y = to_op(locp[2 * j + 1]);
                lv = s_lvl[(p_shift >= 0) ? (j >> p_shift) : (j / P)];
                g = point_geom(x, y, lv.x, lv.y);
                live = g.in_range;
            }
            const unsigned livemask = __ballot_sync(0xffffffffu, live);
            __syncwarp();
            if (livemask != 0u) {
                emit_taps(taps, lane, live, g, a, lv.x, lv.y, lv.z, row_bytes, zero_off);
#pragma unroll
                for (int k = 0; k < 4; ++k) dots[k * kDotStride + lane] = 0.f;
                __syncwarp();
#pragma unroll 4
                for (int it = 0; it < NIT; ++it) {
                    const int tix = it * RPI + slot;              // tap = point * 4 + corner
                    const int pt = tix >> 2, corner = tix & 3;
                    if (((livemask >> pt) & 1u) == 0u && RPI <= 4) continue;   // whole instruction dead (RPI <= 4: one point)
                    Tap t;
                    *reinterpret_cast<uint4 *>(&t) = *reinterpret_cast<const uint4 *>(&taps[corner * kTapStride + pt]);
                    const bool hit = t.off != zero_off;           // fetched corner of a live point
                    float f[VEC];
                    Vec16<T>::unpack(ldg_nc_v4(vbase + t.off), f);
                    float d = 0.f;
#pragma unroll
                    for (int k = 0; k < VEC; ++k) d = fmaf(go[k], f[k], d);
#pragma unroll
                    for (int o = 1; o < LPR; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                    if (hit) {
                        float *gp = reinterpret_cast<float *>(reinterpret_cast<char *>(gv_base) + (t.off / (long long)sizeof(T)) * 4);
#pragma unroll
                        for (int k = 0; k < VEC; k += 4)
                            red_add_v4(gp + k, go[k] * t.w0, go[k + 1] * t.w0, go[k + 2] * t.w0, go[k + 3] * t.w0);
                        if (sub == 0) dots[corner * kDotStride + pt] = d;
                    }
                }
                __syncwarp();
            }
            // ---- per-point gradients (lane = point) ----------------------------------------
            if (j < LP) {
                float gx = 0.f, gy = 0.f, gw = 0.f;
                if (live) {
                    const float d1 = dots[0 * kDotStride + lane], d2 = dots[1 * kDotStride + lane];
                    const float d3 = dots[2 * kDotStride + lane], d4 = dots[3 * kDotStride + lane];
                    const float lh = g.lh, lw = g.lw, hh = 1.f - lh, hw = 1.f - lw;
                    gw = hh * hw * d1 + hh * lw * d2 + lh * hw * d3 + lh * lw * d4;            // cuh:150-156
                    gx = (float)lv.y * a * (-hh * d1 + hh * d2 - lh * d3 + lh * d4);              // width  * grad_w_weight
                    gy = (float)lv.x * a * (-hw * d1 - lw * d2 + hw * d3 + lw * d4);              // height * grad_h_weight
                }
                grad_attn[qm * LP + j] = gw;
                grad_loc[(qm * LP + j) * 2] = gx;
                grad_loc[(qm * LP + j) * 2 + 1] = gy;
            }
            __syncwarp();
        }
    }
}

template <typename T, int D>
static int launch_bwd(const void *value, const int64_t *shapes, const int64_t *starts, const void *loc, const void *attn,
                      const void *grad_out, float *gv, float *gl, float *ga, int N, int S, int M, int L, int Lq, int P,
                      cudaStream_t st) {
    int p_shift = -1;
    if ((P & (P - 1)) == 0) { p_shift = 0; while ((1 << p_shift) < P) ++p_shift; }
    const size_t per_warp = ((kTapsPerWarp * sizeof(Tap) + 4 * kDotStride * 4 + 16 + 15) / 16) * 16;
    const size_t smem = (size_t)L * sizeof(int4) + kWarpsPerCta * per_warp;
    auto kern = msda_bwd_rows_kernel<T, D>;
    if (smem > 48 * 1024) MMFS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ctas = 0;
    MMFS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, 32 * kWarpsPerCta, smem));
    if (ctas < 1) { set_error("msda_backward: kernel does not fit"); return MMFS_EUNSUPPORTED; }
    const int nsm = num_sms();
    int rpw = 8;
    while (rpw > 1 && (long)N * M * ((Lq + kWarpsPerCta * rpw - 1) / (kWarpsPerCta * rpw)) < 2L * nsm * ctas) rpw >>= 1;
    const int qtiles = (Lq + kWarpsPerCta * rpw - 1) / (kWarpsPerCta * rpw);
    const long ntiles = (long)N * M * qtiles;
    if (ntiles > 0x3fffffffL) { set_error("msda_backward: too many tiles"); return MMFS_EUNSUPPORTED; }
    const long full = (long)nsm * ctas;
    kern<<<(unsigned)(ntiles < full ? ntiles : full), 32 * kWarpsPerCta, smem, st>>>(
        (const T *)value, shapes, starts, (const T *)loc, (const T *)attn, (const T *)grad_out, gv, gl, ga,
        S, M, L, Lq, P, p_shift, rpw, qtiles, ntiles, ctas, nsm);
    MMFS_CUDA(cudaGetLastError());
    return MMFS_OK;
}

template <typename T>
static int dispatch_bwd(int D, const void *value, const int64_t *shapes, const int64_t *starts, const void *loc, const void *attn,
                        const void *grad_out, float *gv, float *gl, float *ga, int N, int S, int M, int L, int Lq, int P,
                        cudaStream_t st) {
    switch (D) {
        case 32: return launch_bwd<T, 32>(value, shapes, starts, loc, attn, grad_out, gv, gl, ga, N, S, M, L, Lq, P, st);
        case 64: return launch_bwd<T, 64>(value, shapes, starts, loc, attn, grad_out, gv, gl, ga, N, S, M, L, Lq, P, st);
        case 128: return launch_bwd<T, 128>(value, shapes, starts, loc, attn, grad_out, gv, gl, ga, N, S, M, L, Lq, P, st);
        default: set_error("msda_backward: head size %d unsupported (32/64/128)", D); return MMFS_EUNSUPPORTED;
    }
}

}  // namespace mmfs

using namespace mmfs;

extern "C" int mmfs_msda_backward(const void *value, const int64_t *shapes, const int64_t *starts, const void *loc,
                                  const void *attn, const void *grad_out, float *grad_value, float *grad_loc,
                                  float *grad_attn, int N, int S, int M, int D, int L, int Lq, int P, int dtype, void *stream) {
    MMFS_CHECK_ARG(N >= 0 && Lq >= 0 && S > 0 && M > 0 && D > 0 && L > 0 && P > 0, "msda_backward: bad dimension");
    if (N == 0 || Lq == 0) return MMFS_OK;
    MMFS_CHECK_ARG(value && shapes && starts && loc && attn && grad_out && grad_value && grad_loc && grad_attn,
                   "msda_backward: null pointer argument");
    MMFS_CHECK_ARG(((uintptr_t)value | (uintptr_t)grad_out | (uintptr_t)grad_value) % 16 == 0, "msda_backward: 16-byte alignment required");
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case MMFS_F32: return dispatch_bwd<float>(D, value, shapes, starts, loc, attn, grad_out, grad_value, grad_loc, grad_attn, N, S, M, L, Lq, P, st);
        case MMFS_F16: return dispatch_bwd<__half>(D, value, shapes, starts, loc, attn, grad_out, grad_value, grad_loc, grad_attn, N, S, M, L, Lq, P, st);
        case MMFS_BF16: return dispatch_bwd<__nv_bfloat16>(D, value, shapes, starts, loc, attn, grad_out, grad_value, grad_loc, grad_attn, N, S, M, L, Lq, P, st);
        default: set_error("msda_backward: dtype %d unsupported (f32/f16/bf16)", dtype); return MMFS_EUNSUPPORTED;
    }
}
