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
