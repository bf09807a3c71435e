// msda_fwd_sm100.cu -- multi-scale deformable attention forward for sm_100a.
//
// Replaces the reference's thread-per-output-scalar kernel
//   ops/src/cuda/ms_deform_im2col_cuda.cuh:240-302 (ms_deformable_im2col_gpu_kernel)
//   ops/src/cuda/ms_deform_im2col_cuda.cuh:36-87   (ms_deform_attn_im2col_bilinear)
// with a B200-first design (nothing here is translated from the reference kernel):
//
//   * one WARP per output row (b, q, m); every warp-level fetch moves 512 B: the warp is
//     split into RPI = 32 / (D*sizeof(T)/16) "slots", each slot fetches one 16-byte-
//     per-lane value row, so for D = 64 bf16 a single LDG.128 gathers the four
//     bilinear corners of a sampling point (4 fully used 128-byte lines);
//   * index math is done ONCE per (point, corner) -- lane = (point-in-chunk, corner) --
//     instead of once per channel, and handed to the fetching lanes through a 256-byte
//     per-warp shared-memory mailbox (or warp shuffles, template switch);
//   * CTAs are laid out (b, m, q-tile) with the q-tile fastest, so all warps of a CTA
//     -- and neighbouring CTAs -- walk the SAME head's value slab: it stays L1/L2
//     resident (the reference interleaves 2 heads per 128-thread block);
//   * all N batch entries in one launch, 64-bit base addressing, fp32 accumulation,
//     one rounding at the store; sampling locations / weights are streamed past L1;
//   * fetches of taps that cannot contribute (outside the map, masked image => weight
//     exactly 0) are predicated off, whole 8-point chunks are skipped warp-uniformly.
//
// The sampling-point index math (point_geom) is the single statement shared by the
// forward kernels and by the index-stream kernel that the parity tests compare
// bit-for-bit with oracle/msda_ref.c.
#include "common.cuh"

namespace mmfs {

// ------------------------------------------------------------------------------------
// Index math of one sampling point: cuh:287-291 (pixel coordinates, in-range predicate)
// and cuh:41-48 (floor, lerp fractions).
// ------------------------------------------------------------------------------------
template <typename OP> struct PointGeom {
    bool in_range;
    int h_low, w_low;
    OP lh, lw;
};

__device__ __forceinline__ PointGeom<float> point_geom(float x, float y, int H, int W) {
    // cuh:287-288  `loc_h * spatial_h - 0.5`: the product is an opmath (float) multiply
    // rounded on its own; the double literal then forces a separate subtraction (exact
    // in double, rounded once to float) -- equivalent to an un-fused float subtract.
    // __fmul_rn/__fsub_rn are never contracted into an FMA by nvcc.
    const float h_im = __fsub_rn(__fmul_rn(y, (float)H), 0.5f);
    const float w_im = __fsub_rn(__fmul_rn(x, (float)W), 0.5f);
    PointGeom<float> g;
    g.in_range = (h_im > -1.f) && (w_im > -1.f) && (h_im < (float)H) && (w_im < (float)W);  // cuh:291
    const float hf = floorf(h_im), wf = floorf(w_im);                                       // cuh:41-42
    g.h_low = (int)hf;
    g.w_low = (int)wf;
    g.lh = h_im - hf;  // cuh:46 (h - h_low; hf is integral, the subtraction is exact)
    g.lw = w_im - wf;
    return g;
}

__device__ __forceinline__ PointGeom<double> point_geom(double x, double y, int H, int W) {
    // double dispatch (cu:65): same source expression as the reference, so nvcc applies the
    // same contraction it applies there.
    const double h_im = y * H - 0.5;
    const double w_im = x * W - 0.5;
    PointGeom<double> g;
    g.in_range = (h_im > -1) && (w_im > -1) && (h_im < H) && (w_im < W);
    const double hf = floor(h_im), wf = floor(w_im);
    g.h_low = (int)hf;
    g.w_low = (int)wf;
    g.lh = h_im - hf;
    g.lw = w_im - wf;
    return g;
}

// corner k = 0..3 <-> reference v1..v4: (h_low,w_low) (h_low,w_high) (h_high,w_low) (h_high,w_high)
// validity predicates exactly as cuh:59,65,71,77.
__device__ __forceinline__ bool corner_valid(int corner, int h_low, int w_low, int H, int W) {
    const bool okh = (corner & 2) ? (h_low + 1 <= H - 1) : (h_low >= 0);
    const bool okw = (corner & 1) ? (w_low + 1 <= W - 1) : (w_low >= 0);
    return okh && okw;
}

// ------------------------------------------------------------------------------------
// Fast path: one warp per (b, q, m); D * sizeof(T) in {64,128,256,512} bytes.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void fma2(float &a0, float &a1, float w0, float w1, float v0, float v1) {
    // Blackwell packed fp32 FMA (fma.rn.f32x2): two accumulator updates per issue slot.
    unsigned long long acc, vv, ww;
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(vv) : "f"(v0), "f"(v1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ww) : "f"(w0), "f"(w1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(ww), "l"(vv));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc));
}

template <typename T> struct PointLoad;  // raw (x, y, a) of one sampling point -> fp32
template <> struct PointLoad<float> {
    __device__ __forceinline__ static void load(const float *locp, const float *attp, int j, float &x, float &y, float &a) {
        const uint2 xy = ldg_stream_v2(locp + 2 * (size_t)j);
        x = __uint_as_float(xy.x); y = __uint_as_float(xy.y);
        a = __uint_as_float(ldg_stream_u32(attp + j));
    }
};
template <> struct PointLoad<__nv_bfloat16> {
    __device__ __forceinline__ static void load(const __nv_bfloat16 *locp, const __nv_bfloat16 *attp, int j, float &x, float &y, float &a) {
        const uint32_t xy = ldg_stream_u32(locp + 2 * (size_t)j);
        x = __uint_as_float(xy << 16); y = __uint_as_float(xy & 0xffff0000u);
        a = __uint_as_float(((uint32_t)ldg_stream_u16(attp + j)) << 16);
    }
};
template <> struct PointLoad<__half> {
    __device__ __forceinline__ static void load(const __half *locp, const __half *attp, int j, float &x, float &y, float &a) {
        const uint32_t xy = ldg_stream_u32(locp + 2 * (size_t)j);
        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&xy));
        x = f.x; y = f.y;
        const uint16_t aw = ldg_stream_u16(attp + j);
        a = __half2float(*reinterpret_cast<const __half *>(&aw));
    }
};

// 512 bytes of zeros: taps that must not contribute (outside the map, invalid corner, masked
// image) are pointed here with weight 0, so the gather loop needs no predicates and a
// non-finite `value` entry can never leak through a 0 * inf product.
__device__ uint4 g_zero_row[32];

struct __align__(16) Tap {  // mailbox record handed from the index-math lane to the fetching slot
    long long off;          // byte offset from this lane's value base (or to g_zero_row)
    float w0, w1;           // lerp weight * attention weight, duplicated for fma.rn.f32x2
};

template <typename T, int D, bool SMEM_XCHG>
__global__ void __launch_bounds__(256, 3)
msda_fwd_warp_kernel(const T *__restrict__ value, const int64_t *__restrict__ shapes,
                     const int64_t *__restrict__ starts, const T *__restrict__ loc,
                     const T *__restrict__ attn, T *__restrict__ out,
                     long nrows, int S, int M, int L, int Lq, int P, int p_shift, unsigned flags, int qtiles, int mapping) {
    constexpr int VEC = 16 / (int)sizeof(T);  // channels per lane
    constexpr int LPR = D / VEC;              // lanes per value row
    constexpr int RPI = 32 / LPR;             // rows (taps) fetched per warp instruction
    static_assert(D % VEC == 0 && LPR >= 1 && LPR <= 32 && (LPR & (LPR - 1)) == 0, "unsupported D");

    extern __shared__ int4 s_dyn[];
    int4 *s_lvl = s_dyn;                                             // [L] {H, W, start, -}
    Tap *s_box = reinterpret_cast<Tap *>(s_dyn + L);                 // [warps][2][32] mailbox
    for (int l = threadIdx.x; l < L; l += blockDim.x)
        s_lvl[l] = make_int4((int)shapes[2 * l], (int)shapes[2 * l + 1], (int)starts[l], 0);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    int b, m, q;
    if (mapping == 0) {  // (b, m, q-tile): all warps of a CTA share one head slab
        const int qt = blockIdx.x % qtiles;
        const int bm = blockIdx.x / qtiles;
        m = bm % M; b = bm / M;
        q = qt * wpc + warp;
    } else {             // reference-like order (b, q, m): for A/B measurements only
        const long gw = (long)blockIdx.x * wpc + warp;
        if (gw >= nrows) return;
        m = (int)(gw % M);
        const long bq = gw / M;
        q = (int)(bq % Lq); b = (int)(bq / Lq);
    }
    if (q >= Lq) return;  // no block-level sync below this point

    const int LP = L * P;
    const size_t qm = ((size_t)b * Lq + q) * M + m;
    const T *locp = loc + qm * (size_t)LP * 2;
    const T *attp = attn + qm * (size_t)LP;
    const int slot = lane / LPR;
    // per-lane gather base, materialised in a register pair (one 64-bit add per fetch)
    const char *vbase = reinterpret_cast<const char *>(value + ((size_t)b * S * M + m) * D) + (lane % LPR) * 16;
    asm volatile("" : "+l"(vbase));
    const long long row_bytes = (long long)M * D * (int)sizeof(T);
    // what a lane in phase 1 must publish so that `vbase(of the fetching lane) + off` lands in
    // g_zero_row: the fetching lane's (lane % LPR) * 16 is part of ITS vbase, so subtract the
    // slab origin only.
    const long long zero_off = reinterpret_cast<const char *>(g_zero_row) -
                               reinterpret_cast<const char *>(value + ((size_t)b * S * M + m) * D);
    const bool strict = flags & MMFS_MSDA_STRICT;

    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;

    const int pt = lane >> 2, corner = lane & 3;  // phase-1 role: (point in chunk, corner)
    float nx = 0.f, ny = 0.f, na = 0.f;
    if (pt < LP) PointLoad<T>::load(locp, attp, pt, nx, ny, na);
    Tap *box = s_box + warp * 64;
    int parity = 0;

    for (int j0 = 0; j0 < LP; j0 += 8) {
        const float x = nx, y = ny, a = na;
        const int j = j0 + pt;
        if (j + 8 < LP) PointLoad<T>::load(locp, attp, j + 8, nx, ny, na);  // software prefetch

        // ---- phase 1: one (point, corner) tap per lane ---------------------------------
        Tap tap;
        tap.off = zero_off;
        tap.w0 = 0.f;
        bool live = false;
        if (j < LP) {
            const int l = (p_shift >= 0) ? (j >> p_shift) : (j / P);
            const int4 lv = s_lvl[l];
            const PointGeom<float> g = point_geom(x, y, lv.x, lv.y);
            live = g.in_range && corner_valid(corner, g.h_low, g.w_low, lv.x, lv.y) && (strict || a != 0.f);
            if (live) {
                const int hc = g.h_low + (corner >> 1), wc = g.w_low + (corner & 1);
                const int row = lv.z + hc * lv.y + wc;  // row of the (S, M*D) slab of batch entry b
                tap.off = (long long)row * row_bytes;
                const float fh = (corner & 2) ? g.lh : 1.f - g.lh;   // cuh:48, 83
                const float fw = (corner & 1) ? g.lw : 1.f - g.lw;
                tap.w0 = fh * fw * a;
            }
        }
        tap.w1 = tap.w0;
        if (__ballot_sync(0xffffffffu, live) == 0u) continue;  // nothing to fetch (masked image)

        // ---- phase 2: slot s fetches tap (it*RPI + s); 16 bytes per lane ---------------
        Tap *mybox = box + parity * 32;   // double-buffered: one __syncwarp per chunk
        parity ^= 1;
        if (SMEM_XCHG) {
            *reinterpret_cast<uint4 *>(mybox + lane) = *reinterpret_cast<const uint4 *>(&tap);  // STS.128
            __syncwarp();
        }
        constexpr int NIT = 32 / RPI;
        constexpr int G = NIT < 4 ? NIT : 4;  // fetches in flight per lane
#pragma unroll
        for (int g0 = 0; g0 < NIT; g0 += G) {
            Tap t[G];
            uint4 v[G];
#pragma unroll
            for (int it = 0; it < G; ++it) {
                const int src = (g0 + it) * RPI + slot;
                if (SMEM_XCHG) {
                    *reinterpret_cast<uint4 *>(&t[it]) = *reinterpret_cast<const uint4 *>(mybox + src);  // LDS.128
                } else {
                    t[it].off = __shfl_sync(0xffffffffu, tap.off, src);
                    t[it].w0 = __shfl_sync(0xffffffffu, tap.w0, src);
                    t[it].w1 = t[it].w0;
                }
            }
#pragma unroll
            for (int it = 0; it < G; ++it) v[it] = ldg_nc_v4(vbase + t[it].off);
#pragma unroll
            for (int it = 0; it < G; ++it) {
                float f[VEC];
                Vec16<T>::unpack(v[it], f);
#pragma unroll
                for (int k = 0; k < VEC; k += 2) fma2(acc[k], acc[k + 1], t[it].w0, t[it].w1, f[k], f[k + 1]);
            }
        }
    }

    // ---- epilogue: sum the RPI slots, one rounding, 16-byte stores ---------------------
#pragma unroll
    for (int off = LPR; off < 32; off <<= 1)
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
    if (lane < LPR) {
        T *op = out + qm * D + lane * VEC;
        stg_v4(op, Vec16<T>::pack(acc));
    }
}

// ------------------------------------------------------------------------------------
// Generic path (any D, f64): one thread per output scalar, index math per thread.
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
msda_fwd_generic_kernel(const T *__restrict__ value, const int64_t *__restrict__ shapes,
                        const int64_t *__restrict__ starts, const T *__restrict__ loc,
                        const T *__restrict__ attn, T *__restrict__ out,
                        long total, int S, int M, int D, int L, int Lq, int P, unsigned flags) {
    using OP = typename OpMath<T>::type;
    const bool strict = flags & MMFS_MSDA_STRICT;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % D);
        const long qm = idx / D;
        const int m = (int)(qm % M);
        const int b = (int)(qm / M / Lq);
        const T *vb = value + ((size_t)b * S * M + m) * D + c;
        const T *locp = loc + (size_t)qm * L * P * 2;
        const T *attp = attn + (size_t)qm * L * P;
        OP col = 0;
        for (int l = 0; l < L; ++l) {
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1], st = (int)starts[l];
            for (int p = 0; p < P; ++p) {
                const int j = l * P + p;
                const OP x = to_op(locp[2 * j]), y = to_op(locp[2 * j + 1]), a = to_op(attp[j]);
                const PointGeom<OP> g = point_geom(x, y, H, W);
                if (!g.in_range || (!strict && a == (OP)0)) continue;
                OP val = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!corner_valid(k, g.h_low, g.w_low, H, W)) continue;
                    const int hc = g.h_low + (k >> 1), wc = g.w_low + (k & 1);
                    const OP fh = (k & 2) ? g.lh : (OP)1 - g.lh;
                    const OP fw = (k & 1) ? g.lw : (OP)1 - g.lw;
                    val += fh * fw * to_op(vb[(size_t)(st + hc * W + wc) * M * D]);
                }
                col += val * a;
            }
        }
        out[idx] = from_op<T>(col);
    }
}

// ------------------------------------------------------------------------------------
// Index stream (parity instrumentation; same point_geom / corner_valid as above).
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
msda_index_stream_kernel(const int64_t *__restrict__ shapes, const int64_t *__restrict__ starts,
                         const T *__restrict__ loc, int32_t *__restrict__ idx,
                         long total, int M, int D, int L, int P) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int p = (int)(i % P);
        const int l = (int)((i / P) % L);
        const int m = (int)((i / P / L) % M);
        (void)p;
        const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
        const float x = to_op(loc[2 * i]), y = to_op(loc[2 * i + 1]);
        const PointGeom<float> g = point_geom(x, y, H, W);
        int32_t rec[8] = {0, 0, 0, 0, -1, -1, -1, -1};
        if (g.in_range) {
            rec[0] = 1; rec[1] = g.h_low; rec[2] = g.w_low;
            const int w_stride = M * D, h_stride = W * w_stride;            // cuh:50-51
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (corner_valid(k, g.h_low, g.w_low, H, W)) {
                    rec[3] |= 1 << k;
                    rec[4 + k] = (g.h_low + (k >> 1)) * h_stride + (g.w_low + (k & 1)) * w_stride + m * D;
                }
        }
        int4 *o = reinterpret_cast<int4 *>(idx + i * 8);
        o[0] = make_int4(rec[0], rec[1], rec[2], rec[3]);
        o[1] = make_int4(rec[4], rec[5], rec[6], rec[7]);
    }
}

// ------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------
static int g_warps_per_cta = 0;  // 0 = automatic
static int g_mapping = 0;        // bit0: 0 (b,m,q-tile) / 1 (b,q,m); bit1: shuffle exchange instead of mailbox

template <typename T, int D, bool X>
static int launch_warp(const void *value, const int64_t *shapes, const int64_t *starts, const void *loc,
                       const void *attn, void *out, int N, int S, int M, int L, int Lq, int P,
                       unsigned flags, cudaStream_t st) {
    int p_shift = -1;
    if ((P & (P - 1)) == 0) { p_shift = 0; while ((1 << p_shift) < P) ++p_shift; }
    // warps per CTA: as large as possible (L1 locality on the head slab) while still giving
    // every SM several CTAs; tiny problems (decode, Lq = 1) fall to 1-2 warps per CTA.
    int wpc = g_warps_per_cta;
    if (wpc <= 0) {
        const long want = 4L * num_sms();
        wpc = 8;
        while (wpc > 1 && ((long)N * M * ((Lq + wpc - 1) / wpc) < want || wpc / 2 >= Lq)) wpc >>= 1;
    }
    const int mapping = g_mapping & 1;
    const int qtiles = (Lq + wpc - 1) / wpc;
    const long nrows = (long)N * Lq * M;
    const long nb = mapping == 0 ? (long)N * M * qtiles : (nrows + wpc - 1) / wpc;
    if (nb > 0x7fffffffL) { set_error("msda: grid too large (%ld CTAs)", nb); return MMFS_EUNSUPPORTED; }
    const size_t smem = (size_t)L * sizeof(int4) + (X ? (size_t)wpc * 64 * sizeof(Tap) : 0);
    if (smem > 48 * 1024) { set_error("msda: too many levels (%d)", L); return MMFS_EUNSUPPORTED; }
    msda_fwd_warp_kernel<T, D, X><<<dim3((unsigned)nb), dim3(32 * wpc), smem, st>>>(
        (const T *)value, shapes, starts, (const T *)loc, (const T *)attn, (T *)out,
        nrows, S, M, L, Lq, P, p_shift, flags, qtiles, mapping);
    MMFS_CUDA(cudaGetLastError());
    return MMFS_OK;
}

template <typename T>
static int launch_generic(const void *value, const int64_t *shapes, const int64_t *starts, const void *loc,
                          const void *attn, void *out, int N, int S, int M, int D, int L, int Lq, int P,
                          unsigned flags, cudaStream_t st) {
    const long total = (long)N * Lq * M * D;
    const long blocks = (total + 255) / 256;
    const int grid = (int)(blocks < 148L * 32 ? blocks : 148L * 32);
    msda_fwd_generic_kernel<T><<<grid, 256, 0, st>>>((const T *)value, shapes, starts, (const T *)loc,
                                                       (const T *)attn, (T *)out, total, S, M, D, L, Lq, P, flags);
    MMFS_CUDA(cudaGetLastError());
    return MMFS_OK;
}

template <typename T>
static int dispatch_d(const void *value, const int64_t *shapes, const int64_t *starts, const void *loc,
                      const void *attn, void *out, int N, int S, int M, int D, int L, int Lq, int P,
                      unsigned flags, cudaStream_t st) {
    const bool shfl = g_mapping & 2;
#define MMFS_CASE(DD)                                                                                     \
    case DD:                                                                                              \
        return shfl ? launch_warp<T, DD, false>(value, shapes, starts, loc, attn, out, N, S, M, L, Lq, P, flags, st) \
                    : launch_warp<T, DD, true>(value, shapes, starts, loc, attn, out, N, S, M, L, Lq, P, flags, st);
    const bool aligned16 = ((uintptr_t)value % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                           ((uintptr_t)loc % 8 == 0) && ((uintptr_t)attn % 4 == 0);
    if (aligned16 && (long)S * M * D * (long)sizeof(T) < (1L << 32)) {
        switch (D) {
            MMFS_CASE(32)
            MMFS_CASE(64)
            MMFS_CASE(128)
            default: break;
        }
    }
#undef MMFS_CASE
    return launch_generic<T>(value, shapes, starts, loc, attn, out, N, S, M, D, L, Lq, P, flags, st);
}

}  // namespace mmfs

using namespace mmfs;

extern "C" int mmfs_msda_set_tuning(int warps_per_cta, int mapping) {
    if (warps_per_cta < 0 || warps_per_cta > 8 || (warps_per_cta & (warps_per_cta - 1))) {
        set_error("mmfs_msda_set_tuning: warps_per_cta must be 0 or a power of two <= 8");
        return MMFS_EINVAL;
    }
    g_warps_per_cta = warps_per_cta;
    g_mapping = mapping;
    return MMFS_OK;
}

static int check_msda_args(const void *value, const int64_t *shapes, const int64_t *starts, const void *loc,
                           const void *attn, const void *out, int N, int S, int M, int D, int L, int Lq, int P, int dtype) {
    MMFS_CHECK_ARG(N >= 0 && Lq >= 0, "msda: negative batch (%d) or query count (%d)", N, Lq);
    MMFS_CHECK_ARG(S > 0 && M > 0 && D > 0 && L > 0 && P > 0,
                   "msda: non-positive dimension S=%d M=%d D=%d L=%d P=%d", S, M, D, L, P);
    MMFS_CHECK_ARG(dtype_size(dtype) != 0, "msda: unknown dtype code %d", dtype);
    if (N == 0 || Lq == 0) return MMFS_OK;
    MMFS_CHECK_ARG(value && shapes && starts && loc && attn && out, "msda: null pointer argument");
    if ((long)S * M * D >= (1L << 31) || (long)Lq * M * L * P * 2 >= (1L << 31) * 64L) {
        set_error("msda: per-batch slab too large for 32-bit row offsets (S=%d M=%d D=%d)", S, M, D);
        return MMFS_EUNSUPPORTED;
    }
    return MMFS_OK;
}

extern "C" int mmfs_msda_forward(const void *value, const int64_t *shapes, const int64_t *starts,
                                 const void *loc, const void *attn, void *out,
                                 int N, int S, int M, int D, int L, int Lq, int P,
                                 int dtype, unsigned flags, void *stream) {
    int rc = check_msda_args(value, shapes, starts, loc, attn, out, N, S, M, D, L, Lq, P, dtype);
    if (rc != MMFS_OK || N == 0 || Lq == 0) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case MMFS_F32: return dispatch_d<float>(value, shapes, starts, loc, attn, out, N, S, M, D, L, Lq, P, flags, st);
        case MMFS_F16: return dispatch_d<__half>(value, shapes, starts, loc, attn, out, N, S, M, D, L, Lq, P, flags, st);
        case MMFS_BF16: return dispatch_d<__nv_bfloat16>(value, shapes, starts, loc, attn, out, N, S, M, D, L, Lq, P, flags, st);
        case MMFS_F64: return launch_generic<double>(value, shapes, starts, loc, attn, out, N, S, M, D, L, Lq, P, flags, st);
    }
    return MMFS_EINVAL;
}

extern "C" int mmfs_msda_index_stream(const int64_t *shapes, const int64_t *starts, const void *loc,
                                      int32_t *idx, int N, int M, int D, int L, int Lq, int P,
                                      int dtype, void *stream) {
    MMFS_CHECK_ARG(N >= 0 && Lq >= 0 && M > 0 && D > 0 && L > 0 && P > 0, "msda_index_stream: bad dimension");
    MMFS_CHECK_ARG(dtype == MMFS_F32 || dtype == MMFS_F16 || dtype == MMFS_BF16,
                   "msda_index_stream: dtype %d not supported (f32/f16/bf16 only)", dtype);
    const long total = (long)N * Lq * M * L * P;
    if (total == 0) return MMFS_OK;
    MMFS_CHECK_ARG(shapes && starts && loc && idx, "msda_index_stream: null pointer argument");
    cudaStream_t st = (cudaStream_t)stream;
    const long blocks = (total + 255) / 256;
    const int grid = (int)(blocks < 148L * 32 ? blocks : 148L * 32);
    switch (dtype) {
        case MMFS_F32: msda_index_stream_kernel<float><<<grid, 256, 0, st>>>(shapes, starts, (const float *)loc, idx, total, M, D, L, P); break;
        case MMFS_F16: msda_index_stream_kernel<__half><<<grid, 256, 0, st>>>(shapes, starts, (const __half *)loc, idx, total, M, D, L, P); break;
        default: msda_index_stream_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(shapes, starts, (const __nv_bfloat16 *)loc, idx, total, M, D, L, P); break;
    }
    MMFS_CUDA(cudaGetLastError());
    return MMFS_OK;
}
