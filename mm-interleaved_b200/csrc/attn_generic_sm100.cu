// attn_generic_sm100.cu -- bandwidth-oriented softmax(QK^T * scale + mask) V for the cases the
// tensor-core kernel (attn_fwd_sm100.cu) does not take: single-token decode over a KV cache
// (q_len = 1: a GEMV-shaped, HBM-bound problem -- no tensor cores needed) and small / odd shapes
// (head sizes other than 64/128, a few query rows).
//
// Semantics follow the reference's eager attention, LlamaAttention.forward
// (decoders/modeling_llama_mmfs.py:246-264): scores = (q * hd^-0.5) k^T + causal/padding mask,
// fp32 softmax, P V; and CLIPXAttention.forward (encoders/vit_adapter/xattn.py:47-141) when
// causal = 0 and no key mask.  Layout is the projection GEMMs' own (B, T, H, hd) -- no transposes.
// A query row whose keys are ALL masked (a left-padding position) returns zeros; the reference's
// finfo.min clamp makes such rows attend uniformly to every key, but those rows are padding and
// their outputs are never consumed ("Attention masks").
//
// One warp per (b, h, query row).  Keys are processed in chunks of 32*KPL: phase 1 gives every
// lane whole keys (row-contiguous 16-byte loads, q broadcast from shared memory), phase 2 gives
// every lane channels (coalesced V rows), online softmax across chunks.
#include "common.cuh"

namespace mmfs {

constexpr int kAttnChunk = 256;   // keys per chunk (8 per lane)
constexpr int kAttnWarps = 4;

template <typename T>
__global__ void __launch_bounds__(32 * kAttnWarps)
attn_generic_kernel(const T *__restrict__ q, const T *__restrict__ k, const T *__restrict__ v, T *__restrict__ out,
                    const uint8_t *__restrict__ key_mask, long n_rows, int H, int Tq, int Tkv, int hd,
                    long q_bs, long q_ts, long k_bs, long k_ts, long v_bs, long v_ts, long o_bs, long o_ts,
                    float scale, int causal, int past) {
    extern __shared__ float s_dyn_f[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *s_q = s_dyn_f + warp * (hd + kAttnChunk);
    float *s_p = s_q + hd;
    const int cpl = (hd + 31) / 32;   // channels per lane (<= 8)

    for (long row = (long)blockIdx.x * kAttnWarps + warp; row < n_rows; row += (long)gridDim.x * kAttnWarps) {
        const int i = (int)(row % Tq);
        const int h = (int)((row / Tq) % H);
        const int b = (int)(row / Tq / H);
        const T *qp = q + b * q_bs + i * q_ts + (long)h * hd;
        __syncwarp();
        for (int d = lane; d < hd; d += 32) s_q[d] = to_op(qp[d]) * scale;
        __syncwarp();
        const int last_key = causal ? min(Tkv - 1, past + i) : Tkv - 1;
        float m_run = -INFINITY, l_run = 0.f;
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.f;

        for (int j0 = 0; j0 <= last_key; j0 += kAttnChunk) {
            // phase 1: scores of up to kAttnChunk keys, lane owns keys j0 + lane + 32*t
            float cmax = -INFINITY;
            for (int t = 0; t < kAttnChunk / 32; ++t) {
                const int j = j0 + lane + 32 * t;
                float s = -INFINITY;
                if (j <= last_key && (key_mask == nullptr || key_mask[(long)b * Tkv + j])) {
                    const T *kp = k + b * k_bs + j * k_ts + (long)h * hd;
                    float dot = 0.f;
                    for (int d = 0; d < hd; ++d) dot += s_q[d] * to_op(kp[d]);
                    s = dot;
                }
                s_p[lane + 32 * t] = s;
                cmax = fmaxf(cmax, s);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
            const float m_new = fmaxf(m_run, cmax);
            if (m_new == -INFINITY) { __syncwarp(); continue; }       // nothing visible yet
            const float corr = __expf(m_run - m_new);                 // m_run = -inf -> 0
            float csum = 0.f;
            for (int t = 0; t < kAttnChunk / 32; ++t) {
                const float p = __expf(s_p[lane + 32 * t] - m_new);   // masked (-inf) -> 0
                s_p[lane + 32 * t] = p;
                csum += p;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
            l_run = l_run * corr + csum;
            m_run = m_new;
            __syncwarp();
            // phase 2: acc[c] += p_j * v[j][channel], lane owns channels lane + 32*c
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] *= corr;
            const int jn = min(kAttnChunk, last_key - j0 + 1);
            for (int jj = 0; jj < jn; ++jj) {
                const float p = s_p[jj];
                if (p == 0.f) continue;                                // warp-uniform (same smem word)
                const T *vp = v + b * v_bs + (long)(j0 + jj) * v_ts + (long)h * hd;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < cpl && lane + 32 * c < hd) acc[c] += p * to_op(vp[lane + 32 * c]);
            }
            __syncwarp();
        }
        T *op = out + b * o_bs + i * o_ts + (long)h * hd;
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (c < cpl && lane + 32 * c < hd) op[lane + 32 * c] = from_op<T>(acc[c] * inv);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Decode (q_len = 1 over a KV cache): split-KV ("flash decoding").  The row-per-warp kernel above gives a decode step
// B*H = 160 warps for the whole GPU, each walking 2k keys serially with scalar loads (3.4 ms per layer at the cfg-3
// cache).  Here grid = (ceil(Tkv / 256), H, B): every CTA reduces 256 keys of one (b, h) -- lane = key for the
// scores (16-byte loads along the key row, q broadcast from shared memory), lane = channels for P V (coalesced V
// rows) -- and writes an (m, l, acc[hd]) partial; a second kernel merges the partials.  HBM-bound: K and V are read
// exactly once.
// ------------------------------------------------------------------------------------------------------------
constexpr int kDecKeys = 256;     // keys per CTA
constexpr int kDecWarps = 4;      // 64 keys per warp, two passes of 32

template <typename T>
__global__ void __launch_bounds__(32 * kDecWarps)
attn_decode_split_kernel(const T *__restrict__ q, const T *__restrict__ k, const T *__restrict__ v,
                         const uint8_t *__restrict__ key_mask, float *__restrict__ part, int H, int Tkv, int hd,
                         long q_bs, long k_bs, long k_ts, long v_bs, long v_ts, float scale, int last_key) {
    constexpr int VEC = 16 / (int)sizeof(T);
    extern __shared__ float s_dec[];                  // q[hd] | per-warp partials [kDecWarps][hd + 2]
    float *s_q = s_dec, *s_red = s_dec + hd;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z, n_split = gridDim.x;
    const int cpl = hd / 32;                          // channels per lane in the P V phase (host: hd % 32 == 0, <= 8)
    const T *qp = q + b * q_bs + (long)h * hd;
    for (int d = threadIdx.x; d < hd; d += blockDim.x) s_q[d] = to_op(qp[d]) * scale;
    __syncthreads();

    float m_run = -INFINITY, l_run = 0.f;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    const int kbase = split * kDecKeys + warp * (kDecKeys / kDecWarps);
#pragma unroll 1
    for (int it = 0; it < kDecKeys / kDecWarps / 32; ++it) {
        const int j0 = kbase + it * 32;
        if (j0 > last_key) break;                     // warp-uniform
        const int j = j0 + lane;
        float sc = -INFINITY;
        if (j <= last_key && (key_mask == nullptr || key_mask[(long)b * Tkv + j])) {
            const T *kp = k + b * k_bs + (long)j * k_ts + (long)h * hd;
            float dot = 0.f;
            for (int d0 = 0; d0 < hd; d0 += VEC) {
                float f[VEC];
                Vec16<T>::unpack(ldg_nc_v4(kp + d0), f);
#pragma unroll
                for (int e = 0; e < VEC; e += 4) {
                    const float4 qq = *reinterpret_cast<const float4 *>(s_q + d0 + e);
                    dot = fmaf(f[e], qq.x, dot); dot = fmaf(f[e + 1], qq.y, dot);
                    dot = fmaf(f[e + 2], qq.z, dot); dot = fmaf(f[e + 3], qq.w, dot);
                }
            }
            sc = dot;
        }
        float cmax = sc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
        const float m_new = fmaxf(m_run, cmax);
        if (m_new == -INFINITY) continue;             // nothing visible in this pass (warp-uniform)
        const float corr = __expf(m_run - m_new);     // m_run = -inf -> 0
        const float pj = __expf(sc - m_new);          // masked (-inf) -> 0
        float psum = pj;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
        l_run = l_run * corr + psum;
        m_run = m_new;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] *= corr;
        const T *vb = v + b * v_bs + (long)h * hd + lane * cpl;
        for (int jj = 0; jj < 32; ++jj) {
            const float pw = __shfl_sync(0xffffffffu, pj, jj);
            if (pw == 0.f) continue;                  // warp-uniform
            const T *vp = vb + (long)(j0 + jj) * v_ts;
            bool done = false;
            if constexpr (sizeof(T) == 2) {
                if (cpl == 4) {                        // hd = 128, 16-bit: one 8-byte load per lane, 256 B per warp
                    const uint2 raw = *reinterpret_cast<const uint2 *>(vp);
                    float f[8];
                    Vec16<T>::unpack(make_uint4(raw.x, raw.y, 0u, 0u), f);
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[c] = fmaf(pw, f[c], acc[c]);
                    done = true;
                } else if (cpl == 2) {                 // hd = 64
                    const uint32_t raw = *reinterpret_cast<const uint32_t *>(vp);
                    float f[8];
                    Vec16<T>::unpack(make_uint4(raw, 0u, 0u, 0u), f);
                    acc[0] = fmaf(pw, f[0], acc[0]); acc[1] = fmaf(pw, f[1], acc[1]);
                    done = true;
                }
            }
            if (!done) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < cpl) acc[c] = fmaf(pw, to_op(vp[c]), acc[c]);
            }
        }
    }
    // merge the four warps of the CTA
    float *mine = s_red + warp * (hd + 2);
    if (lane == 0) { mine[hd] = m_run; mine[hd + 1] = l_run; }
#pragma unroll
    for (int c = 0; c < 8; ++c)
        if (c < cpl) mine[lane * cpl + c] = acc[c];
    __syncthreads();
    float m_all = -INFINITY;
#pragma unroll
    for (int w = 0; w < kDecWarps; ++w) m_all = fmaxf(m_all, s_red[w * (hd + 2) + hd]);
    float *dst = part + (((long)b * H + h) * n_split + split) * (hd + 2);
    for (int d = threadIdx.x; d < hd + 2; d += blockDim.x) {
        float r = 0.f;
        if (d == hd) r = m_all;
        else if (m_all != -INFINITY) {
#pragma unroll
            for (int w = 0; w < kDecWarps; ++w) {
                const float mw = s_red[w * (hd + 2) + hd];
                if (mw != -INFINITY) r += __expf(mw - m_all) * s_red[w * (hd + 2) + (d < hd ? d : hd + 1)];
            }
        }
        dst[d] = r;                                   // d < hd: acc; d == hd: m; d == hd + 1: l
    }
}

// hd = 128, 16-bit elements (the Llama decode step): every load is a fully used 16-byte vector.
//   Q K^T : 8 lanes per key (2 x LDG.128 each = the key's 256 bytes), 4 keys per warp step, 3-shuffle reduction;
//   P V   : 16 lanes per key (LDG.128 = 8 channels each), 2 keys per warp step, 4 steps in flight.
// A warp reduces 64 keys in ONE pass (no running rescale) and writes its own (m, l, acc[128]) partial: the merge kernel
// sees kDecWarps partials per CTA.  (The generic kernel above reads K with 32 different rows per warp instruction --
// half of every 32-byte sector per load -- and V with 8-byte loads, one key per step: 60 us per layer at the cfg-3 cache.)
template <typename T>
__global__ void __launch_bounds__(32 * kDecWarps)
attn_decode_split128_kernel(const T *__restrict__ q, const T *__restrict__ k, const T *__restrict__ v,
                            const uint8_t *__restrict__ key_mask, float *__restrict__ part, int H, int Tkv,
                            long q_bs, long k_bs, long k_ts, long v_bs, long v_ts, float scale, int last_key) {
    constexpr int HD = 128, KPW = kDecKeys / kDecWarps;          // 64 keys per warp
    __shared__ float s_p[kDecWarps][KPW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int n_part = gridDim.x * kDecWarps;
    float *dst = part + (((long)b * H + h) * n_part + split * kDecWarps + warp) * (HD + 2);
    const int k0 = split * kDecKeys + warp * KPW;
    if (k0 > last_key) {                                          // nothing for this warp: an empty partial
        if (lane == 0) { dst[HD] = -INFINITY; dst[HD + 1] = 0.f; }
        return;
    }
    // this lane's 16 channels of q (pre-scaled): channels (lane & 7) * 16 ..
    const int sub = lane & 7, grp = lane >> 3;
    float qv[16];
    {
        const T *qp = q + b * q_bs + (long)h * HD + sub * 16;
        float f[8];
        Vec16<T>::unpack(ldg_nc_v4(qp), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) qv[i] = f[i] * scale;
        Vec16<T>::unpack(ldg_nc_v4(qp + 8), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) qv[8 + i] = f[i] * scale;
    }
    const T *kb = k + b * k_bs + (long)h * HD + sub * 16;
    // ---- scores: 16 steps of 4 keys ------------------------------------------------------------------------
#pragma unroll 4
    for (int st = 0; st < KPW / 4; ++st) {
        const int j = k0 + st * 4 + grp;
        const bool ok = j <= last_key && (key_mask == nullptr || key_mask[(long)b * Tkv + j]);
        float dot = 0.f;
        if (ok) {
            const T *kp = kb + (long)j * k_ts;
            float f[8];
            Vec16<T>::unpack(ldg_nc_v4(kp), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) dot = fmaf(f[i], qv[i], dot);
            Vec16<T>::unpack(ldg_nc_v4(kp + 8), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) dot = fmaf(f[i], qv[8 + i], dot);
        }
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        dot += __shfl_xor_sync(0xffffffffu, dot, 4);
        if (sub == 0) s_p[warp][st * 4 + grp] = ok ? dot : -INFINITY;
    }
    __syncwarp();
    // ---- softmax over the warp's 64 keys ---------------------------------------------------------------------
    const float s0 = s_p[warp][lane], s1 = s_p[warp][lane + 32];
    float m = fmaxf(s0, s1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (m == -INFINITY) {                                         // every key of this warp is masked
        if (lane == 0) { dst[HD] = -INFINITY; dst[HD + 1] = 0.f; }
        return;
    }
    const float p0 = __expf(s0 - m), p1 = __expf(s1 - m);       // masked (-inf) -> 0
    float l = p0 + p1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    __syncwarp();
    s_p[warp][lane] = p0;
    s_p[warp][lane + 32] = p1;
    __syncwarp();
    // ---- P V: 32 steps of 2 keys, 16 lanes x 8 channels per key -----------------------------------------------
    const int half = lane >> 4, ch = (lane & 15) * 8;
    const T *vb = v + b * v_bs + (long)h * HD + ch;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll 4
    for (int st = 0; st < KPW / 2; ++st) {
        const int jj = st * 2 + half;
        const float pw = s_p[warp][jj];
        if (pw != 0.f) {                                          // masked / past-the-end keys are never loaded
            float f[8];
            Vec16<T>::unpack(ldg_nc_v4(vb + (long)(k0 + jj) * v_ts), f);
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = fmaf(pw, f[c], acc[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 16);
    if (lane < 16) {        // a partial is hd + 2 = 130 floats: 8-byte aligned, not 16
#pragma unroll
        for (int c = 0; c < 8; c += 2) *reinterpret_cast<float2 *>(dst + ch + c) = make_float2(acc[c], acc[c + 1]);
    }
    if (lane == 0) { dst[HD] = m; dst[HD + 1] = l; }
}

template <typename T>
__global__ void attn_decode_merge_kernel(const float *__restrict__ part, T *__restrict__ out, int H, int hd, int n_split,
                                         long o_bs) {
    const int h = blockIdx.x, b = blockIdx.y;
    const float *p0 = part + (((long)b * H + h) * n_split) * (hd + 2);
    float m = -INFINITY;
    for (int s = 0; s < n_split; ++s) m = fmaxf(m, p0[s * (hd + 2) + hd]);
    for (int d = threadIdx.x; d < hd; d += blockDim.x) {
        float num = 0.f, den = 0.f;
        if (m != -INFINITY)
            for (int s = 0; s < n_split; ++s) {
                const float ms = p0[s * (hd + 2) + hd];
                if (ms == -INFINITY) continue;
                const float w = __expf(ms - m);
                num = fmaf(w, p0[s * (hd + 2) + d], num);
                den = fmaf(w, p0[s * (hd + 2) + hd + 1], den);
            }
        out[b * o_bs + (long)h * hd + d] = from_op<T>(den > 0.f ? num / den : 0.f);   // fully masked row -> zeros
    }
}

template <typename T>
static int launch_attn_decode(const void *q, const void *k, const void *v, void *out, const uint8_t *key_mask, float *scratch,
                              int B, int H, int Tkv, int hd, long q_bs, long k_bs, long k_ts, long v_bs, long v_ts, long o_bs,
                              float scale, int last_key, cudaStream_t st) {
    const int n_split = (last_key + kDecKeys) / kDecKeys;           // keys 0 .. last_key
    dim3 grid(n_split, H, B);
    if constexpr (sizeof(T) == 2) {
        if (hd == 128 && ((uintptr_t)q % 16 == 0) && (q_bs % 8 == 0) && ((uintptr_t)scratch % 16 == 0)) {
            attn_decode_split128_kernel<T><<<grid, 32 * kDecWarps, 0, st>>>((const T *)q, (const T *)k, (const T *)v, key_mask, scratch, H,
                                                                         Tkv, q_bs, k_bs, k_ts, v_bs, v_ts, scale, last_key);
            attn_decode_merge_kernel<T><<<dim3(H, B), 128, 0, st>>>(scratch, (T *)out, H, hd, n_split * kDecWarps, o_bs);
            MMFS_CUDA(cudaGetLastError());
            return MMFS_OK;
        }
    }
    const size_t smem = (size_t)(hd + kDecWarps * (hd + 2)) * sizeof(float);
    attn_decode_split_kernel<T><<<grid, 32 * kDecWarps, smem, st>>>((const T *)q, (const T *)k, (const T *)v, key_mask, scratch, H,
                                                                   Tkv, hd, q_bs, k_bs, k_ts, v_bs, v_ts, scale, last_key);
    attn_decode_merge_kernel<T><<<dim3(H, B), hd < 128 ? 64 : 128, 0, st>>>(scratch, (T *)out, H, hd, n_split, o_bs);
    MMFS_CUDA(cudaGetLastError());
    return MMFS_OK;
}

template <typename T>
static int launch_attn_generic(const void *q, const void *k, const void *v, void *out, const uint8_t *key_mask,
                               int B, int H, int Tq, int Tkv, int hd, long q_bs, long q_ts, long k_bs, long k_ts,
                               long v_bs, long v_ts, long o_bs, long o_ts, float scale, int causal, int past, cudaStream_t st) {
    const long n_rows = (long)B * H * Tq;
    const long blocks = (n_rows + kAttnWarps - 1) / kAttnWarps;
    const int grid = (int)(blocks < 148L * 16 ? blocks : 148L * 16);
    const size_t smem = (size_t)kAttnWarps * (hd + kAttnChunk) * sizeof(float);
    attn_generic_kernel<T><<<grid, 32 * kAttnWarps, smem, st>>>((const T *)q, (const T *)k, (const T *)v, (T *)out, key_mask,
                                                              n_rows, H, Tq, Tkv, hd, q_bs, q_ts, k_bs, k_ts, v_bs, v_ts,
                                                              o_bs, o_ts, scale, causal, past);
    MMFS_CUDA(cudaGetLastError());
    return MMFS_OK;
}

}  // namespace mmfs

using namespace mmfs;

extern "C" int mmfs_attn_generic(const void *q, const void *k, const void *v, void *out, const uint8_t *key_mask,
                                 int B, int H, int Tq, int Tkv, int hd,
                                 long q_bs, long q_ts, long k_bs, long k_ts, long v_bs, long v_ts, long o_bs, long o_ts,
                                 float scale, int causal, int past, int dtype, void *stream) {
    MMFS_CHECK_ARG(B >= 0 && H > 0 && Tq >= 0 && Tkv > 0 && hd > 0 && hd <= 256, "attn_generic: bad shape (hd <= 256)");
    if (B == 0 || Tq == 0) return MMFS_OK;
    MMFS_CHECK_ARG(q && k && v && out, "attn_generic: null pointer argument");
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case MMFS_F32: return launch_attn_generic<float>(q, k, v, out, key_mask, B, H, Tq, Tkv, hd, q_bs, q_ts, k_bs, k_ts, v_bs, v_ts, o_bs, o_ts, scale, causal, past, st);
        case MMFS_F16: return launch_attn_generic<__half>(q, k, v, out, key_mask, B, H, Tq, Tkv, hd, q_bs, q_ts, k_bs, k_ts, v_bs, v_ts, o_bs, o_ts, scale, causal, past, st);
        case MMFS_BF16: return launch_attn_generic<__nv_bfloat16>(q, k, v, out, key_mask, B, H, Tq, Tkv, hd, q_bs, q_ts, k_bs, k_ts, v_bs, v_ts, o_bs, o_ts, scale, causal, past, st);
        default: set_error("attn_generic: dtype %d unsupported", dtype); return MMFS_EINVAL;
    }
}

extern "C" long mmfs_attn_decode_scratch_floats(int B, int H, int Tkv, int hd) {
    return (long)B * H * ((Tkv + kDecKeys - 1) / kDecKeys) * kDecWarps * (hd + 2);   // one partial per warp (hd 128 path)
}

extern "C" int mmfs_attn_decode(const void *q, const void *k, const void *v, void *out, const uint8_t *key_mask, float *scratch,
                                int B, int H, int Tkv, int hd, long q_bs, long k_bs, long k_ts, long v_bs, long v_ts, long o_bs,
                                float scale, int causal, int past, int dtype, void *stream) {
    MMFS_CHECK_ARG(B >= 0 && H > 0 && Tkv > 0 && hd > 0, "attn_decode: bad shape");
    if (B == 0) return MMFS_OK;
    MMFS_CHECK_ARG(q && k && v && out && scratch, "attn_decode: null pointer argument");
    const size_t es = dtype_size(dtype);
    if (dtype == MMFS_F64 || es == 0 || hd % 32 != 0 || hd > 256 || B > 65535 || H > 65535 ||
        ((uintptr_t)k | (uintptr_t)v) % 16 != 0 || (k_bs * es) % 16 != 0 || (k_ts * es) % 16 != 0 ||
        (v_bs * es) % 16 != 0 || (v_ts * es) % 16 != 0 || (hd * es) % 16 != 0) {
        set_error("attn_decode: needs f32/f16/bf16, hd %% 32 == 0 (<= 256), 16-byte aligned K / V rows");
        return MMFS_EUNSUPPORTED;
    }
    const int last_key = causal ? (past < Tkv - 1 ? past : Tkv - 1) : Tkv - 1;     // the single query row sits at position `past`
    MMFS_CHECK_ARG(last_key >= 0, "attn_decode: negative past");
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case MMFS_F32: return launch_attn_decode<float>(q, k, v, out, key_mask, scratch, B, H, Tkv, hd, q_bs, k_bs, k_ts, v_bs, v_ts, o_bs, scale, last_key, st);
        case MMFS_F16: return launch_attn_decode<__half>(q, k, v, out, key_mask, scratch, B, H, Tkv, hd, q_bs, k_bs, k_ts, v_bs, v_ts, o_bs, scale, last_key, st);
        default: return launch_attn_decode<__nv_bfloat16>(q, k, v, out, key_mask, scratch, B, H, Tkv, hd, q_bs, k_bs, k_ts, v_bs, v_ts, o_bs, scale, last_key, st);
    }
}
