// llama_ops_sm100.cu -- HBM-bound element-wise / row-wise kernels of the Llama-MMFS decoder layer.
//
//   mmfs_rmsnorm      LlamaRMSNorm.forward               decoders/modeling_llama_mmfs.py:53-70
//   mmfs_rope_qk      apply_rotary_pos_emb / rotate_half decoders/modeling_llama_mmfs.py:158-172
//   mmfs_swiglu       LlamaMLP: act_fn(gate) * up        decoders/modeling_llama_mmfs.py:188-189
//   mmfs_layernorm    nn.LayerNorm (CLIP / Q-Former / MMFSBlock norms)
//
// Each mimics the rounding points of the reference's tensor pipeline in the storage type T (a
// tensor op in bf16 rounds its result to bf16), so a bf16 run tracks the reference's bf16 run and
// an fp32 run tracks its fp32 run.  All are pure bandwidth kernels: 16-byte vector loads/stores,
// one pass over the data, warp-shuffle + shared-memory row reductions, grid = rows.
#include "common.cuh"

namespace mmfs {

template <typename T> __device__ __forceinline__ float rnd(float x) { return to_op(from_op<T>(x)); }
template <> __device__ __forceinline__ float rnd<float>(float x) { return x; }

__device__ __forceinline__ float block_sum(float v, float *s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = (lane < nw) ? s_red[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    __syncthreads();
    return t;
}

// ---- RMSNorm: y = w * cast_T(x * rsqrt(mean(x^2) + eps)) -------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) rmsnorm_kernel(const T *__restrict__ x, const T *__restrict__ w,
                                                       T *__restrict__ y, int cols, float eps) {
    constexpr int VEC = 16 / (int)sizeof(T);
    __shared__ float s_red[32];
    const size_t row = blockIdx.x;
    const T *xr = x + row * cols;
    T *yr = y + row * cols;
    const int nvec = cols / VEC;
    float ss = 0.f;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        float f[VEC];
        Vec16<T>::unpack(*reinterpret_cast<const uint4 *>(xr + i * VEC), f);
#pragma unroll
        for (int k = 0; k < VEC; ++k) ss += f[k] * f[k];
    }
    for (int i = nvec * VEC + threadIdx.x; i < cols; i += blockDim.x) { const float v = to_op(xr[i]); ss += v * v; }
    const float var = block_sum(ss, s_red) / (float)cols;      // variance in fp32 (:62)
    const float r = rsqrtf(var + eps);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {     // second pass hits L1/L2
        float f[VEC], g[VEC], o[VEC];
        Vec16<T>::unpack(*reinterpret_cast<const uint4 *>(xr + i * VEC), f);
        Vec16<T>::unpack(*reinterpret_cast<const uint4 *>(w + i * VEC), g);
#pragma unroll
        for (int k = 0; k < VEC; ++k) o[k] = g[k] * rnd<T>(f[k] * r);   // cast to weight dtype, then weight * (:65-69)
        *reinterpret_cast<uint4 *>(yr + i * VEC) = Vec16<T>::pack(o);
    }
    for (int i = nvec * VEC + threadIdx.x; i < cols; i += blockDim.x)
        yr[i] = from_op<T>(to_op(w[i]) * rnd<T>(to_op(xr[i]) * r));
}

// ---- LayerNorm (biased variance, fp32 statistics) ------------------------------------------------------
// Warp per row: the row lives in registers (16-byte vector loads), mean and centred variance are two warp
// reductions, one pass over HBM.  Rows of up to 32 * VEC * kLnChunks elements (2048 bf16 / 1024 fp32) take this
// path -- every LayerNorm on the interleaved path (64 ... 1280 columns); longer or unaligned rows use the block kernel.
constexpr int kLnChunks = 8;

template <typename T>
__global__ void __launch_bounds__(256) layernorm_warp_kernel(const T *__restrict__ x, const T *__restrict__ w,
                                                              const T *__restrict__ b, T *__restrict__ y, long rows, int cols, float eps) {
    constexpr int VEC = 16 / (int)sizeof(T);
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const T *xr = x + row * cols;
    T *yr = y + row * cols;
    const int nvec = cols / VEC;                 // host guarantees cols % VEC == 0 and nvec <= 32 * kLnChunks
    float v[kLnChunks][VEC];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kLnChunks; ++c) {
        const int i = lane + 32 * c;
        if (i < nvec) {
            Vec16<T>::unpack(*reinterpret_cast<const uint4 *>(xr + i * VEC), v[c]);
#pragma unroll
            for (int k = 0; k < VEC; ++k) s += v[c][k];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)cols;
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < kLnChunks; ++c)
        if (lane + 32 * c < nvec) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) { const float d = v[c][k] - mean; ss += d * d; }
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float r = rsqrtf(ss / (float)cols + eps);
#pragma unroll
    for (int c = 0; c < kLnChunks; ++c) {
        const int i = lane + 32 * c;
        if (i < nvec) {
            float g[VEC], bb[VEC], o[VEC];
            if (w) Vec16<T>::unpack(*reinterpret_cast<const uint4 *>(w + i * VEC), g);
            if (b) Vec16<T>::unpack(*reinterpret_cast<const uint4 *>(b + i * VEC), bb);
#pragma unroll
            for (int k = 0; k < VEC; ++k) o[k] = (v[c][k] - mean) * r * (w ? g[k] : 1.f) + (b ? bb[k] : 0.f);
            *reinterpret_cast<uint4 *>(yr + i * VEC) = Vec16<T>::pack(o);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) layernorm_kernel(const T *__restrict__ x, const T *__restrict__ w,
                                                         const T *__restrict__ b, T *__restrict__ y, int cols, float eps) {
    __shared__ float s_red[32];
    const size_t row = blockIdx.x;
    const T *xr = x + row * cols;
    T *yr = y + row * cols;
    float s = 0.f;
    for (int i = threadIdx.x; i < cols; i += blockDim.x) s += to_op(xr[i]);
    const float mean = block_sum(s, s_red) / (float)cols;
    float ss = 0.f;
    for (int i = threadIdx.x; i < cols; i += blockDim.x) { const float d = to_op(xr[i]) - mean; ss += d * d; }
    const float r = rsqrtf(block_sum(ss, s_red) / (float)cols + eps);
    for (int i = threadIdx.x; i < cols; i += blockDim.x) {
        const float n = (to_op(xr[i]) - mean) * r;
        yr[i] = from_op<T>(n * (w ? to_op(w[i]) : 1.f) + (b ? to_op(b[i]) : 0.f));
    }
}

// ---- RoPE on q and k, in place, (B, T, H, hd) layout (the GEMM output layout: no transposes) -------
//   q_embed = q * cos + rotate_half(q) * sin, rotate_half(x) = cat(-x2, x1)       (:158-172)
template <typename T>
__global__ void __launch_bounds__(256) rope_qk_kernel(T *__restrict__ q, T *__restrict__ k, const float *__restrict__ cos_t,
                                                       const float *__restrict__ sin_t, const int64_t *__restrict__ pos,
                                                       long n_tok, int H, int hd, int q_stride, int k_stride, int pos_per_batch,
                                                       int T_len) {
    // one thread per (token, head, q-or-k, chunk of VEC rotation pairs): two 16-byte loads (x1 | x2 halves),
    // two 16-byte stores; cos/sin rows are fp32 and L1/L2 resident (one row per token)
    constexpr int VEC = 16 / (int)sizeof(T);
    const int half = hd >> 1;
    const int chunks = half / VEC;
    const long total = n_tok * H * 2 * chunks;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % chunks);
        long r = idx / chunks;
        const int which = (int)(r & 1); r >>= 1;
        const int h = (int)(r % H);
        const long tok = r / H;
        const long p = pos[pos_per_batch ? tok : (tok % T_len)];
        T *x = (which ? k + tok * k_stride : q + tok * q_stride) + (size_t)h * hd + c * VEC;
        float x1[VEC], x2[VEC], o1[VEC], o2[VEC];
        Vec16<T>::unpack(*reinterpret_cast<const uint4 *>(x), x1);
        Vec16<T>::unpack(*reinterpret_cast<const uint4 *>(x + half), x2);
        const float *cp = cos_t + p * hd + c * VEC, *sp = sin_t + p * hd + c * VEC;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float cs = rnd<T>(cp[i]), sn = rnd<T>(sp[i]);          // tables cast to x.dtype (:141-144)
            o1[i] = rnd<T>(x1[i] * cs) + rnd<T>(-x2[i] * sn);
            o2[i] = rnd<T>(x2[i] * cs) + rnd<T>(x1[i] * sn);
        }
        *reinterpret_cast<uint4 *>(x) = Vec16<T>::pack(o1);
        *reinterpret_cast<uint4 *>(x + half) = Vec16<T>::pack(o2);
    }
}

// ---- SwiGLU: out = silu(gate) * up, gate|up stored as one (rows, 2*I) GEMM output -------------------
template <typename T>
__global__ void __launch_bounds__(256) swiglu_kernel(const T *__restrict__ gu, T *__restrict__ out, long rows, int I) {
    constexpr int VEC = 16 / (int)sizeof(T);
    const int nvec = I / VEC;
    const long total = rows * nvec;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long r = idx / nvec;
        const int c = (int)(idx % nvec) * VEC;
        float g[VEC], u[VEC], o[VEC];
        Vec16<T>::unpack(*reinterpret_cast<const uint4 *>(gu + r * 2 * I + c), g);
        Vec16<T>::unpack(*reinterpret_cast<const uint4 *>(gu + r * 2 * I + I + c), u);
#pragma unroll
        for (int k = 0; k < VEC; ++k) o[k] = rnd<T>(g[k] / (1.f + expf(-g[k]))) * u[k];   // act_fn(gate) is a tensor in T
        *reinterpret_cast<uint4 *>(out + r * I + c) = Vec16<T>::pack(o);
    }
}

template <typename T>
static int launch_all(int which, const void *a, const void *b, const void *c, void *d, const void *e, const void *f,
                      long n0, int i0, int i1, int i2, int i3, int i4, int i5, float eps, cudaStream_t st) {
    (void)i5;
    switch (which) {
        case 0:   // rmsnorm: a=x b=w d=y n0=rows i0=cols
            rmsnorm_kernel<T><<<(unsigned)n0, 256, 0, st>>>((const T *)a, (const T *)b, (T *)d, i0, eps);
            break;
        case 1: { // layernorm: a=x b=w c=bias d=y
            constexpr int VEC = 16 / (int)sizeof(T);
            const bool vec_ok = (i0 % VEC == 0) && (i0 / VEC <= 32 * kLnChunks) &&
                                (((uintptr_t)a | (uintptr_t)d | (uintptr_t)b | (uintptr_t)c) % 16 == 0);
            if (vec_ok)
                layernorm_warp_kernel<T><<<(unsigned)((n0 + 7) / 8), 256, 0, st>>>((const T *)a, (const T *)b, (const T *)c, (T *)d, n0, i0, eps);
            else
                layernorm_kernel<T><<<(unsigned)n0, 256, 0, st>>>((const T *)a, (const T *)b, (const T *)c, (T *)d, i0, eps);
            break;
        }
        case 2: { // rope: d=q (in place), a=k (in place, cast away const), e=cos f=sin c=pos; n0=tokens i0=H i1=hd i2=q_stride i3=k_stride i4=pos_per_batch i5=T
            const long total = n0 * i0 * 2 * ((i1 / 2) / (16 / (int)sizeof(T)));
            const int grid = (int)((total + 255) / 256 < 148L * 16 ? (total + 255) / 256 : 148L * 16);
            rope_qk_kernel<T><<<grid, 256, 0, st>>>((T *)d, (T *)const_cast<void *>(a), (const float *)e, (const float *)f,
                                                     (const int64_t *)c, n0, i0, i1, i2, i3, i4, i5);
            break;
        }
        case 3: { // swiglu: a=gate_up d=out n0=rows i0=I
            const long total = n0 * (i0 / (16 / (int)sizeof(T)));
            const int grid = (int)((total + 255) / 256 < 148L * 16 ? (total + 255) / 256 : 148L * 16);
            swiglu_kernel<T><<<grid, 256, 0, st>>>((const T *)a, (T *)d, n0, i0);
            break;
        }
    }
    MMFS_CUDA(cudaGetLastError());
    return MMFS_OK;
}

static int dispatch(int dtype, int which, const void *a, const void *b, const void *c, void *d, const void *e, const void *f,
                    long n0, int i0, int i1, int i2, int i3, int i4, int i5, float eps, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case MMFS_F32: return launch_all<float>(which, a, b, c, d, e, f, n0, i0, i1, i2, i3, i4, i5, eps, st);
        case MMFS_F16: return launch_all<__half>(which, a, b, c, d, e, f, n0, i0, i1, i2, i3, i4, i5, eps, st);
        case MMFS_BF16: return launch_all<__nv_bfloat16>(which, a, b, c, d, e, f, n0, i0, i1, i2, i3, i4, i5, eps, st);
        default: set_error("dtype %d not supported by this kernel", dtype); return MMFS_EINVAL;
    }
}

}  // namespace mmfs

using namespace mmfs;

extern "C" int mmfs_rmsnorm(const void *x, const void *weight, void *y, long rows, int cols, float eps, int dtype, void *stream) {
    MMFS_CHECK_ARG(rows >= 0 && cols > 0, "rmsnorm: bad shape");
    if (rows == 0) return MMFS_OK;
    MMFS_CHECK_ARG(x && weight && y, "rmsnorm: null pointer argument");
    MMFS_CHECK_ARG(((uintptr_t)x | (uintptr_t)weight | (uintptr_t)y) % 16 == 0 && (cols * dtype_size(dtype)) % 16 == 0,
                   "rmsnorm: rows must be 16-byte aligned");
    return dispatch(dtype, 0, x, weight, nullptr, y, nullptr, nullptr, rows, cols, 0, 0, 0, 0, 0, eps, stream);
}

extern "C" int mmfs_layernorm(const void *x, const void *weight, const void *bias, void *y, long rows, int cols, float eps,
                              int dtype, void *stream) {
    MMFS_CHECK_ARG(rows >= 0 && cols > 0, "layernorm: bad shape");
    if (rows == 0) return MMFS_OK;
    MMFS_CHECK_ARG(x && y, "layernorm: null pointer argument");
    return dispatch(dtype, 1, x, weight, bias, y, nullptr, nullptr, rows, cols, 0, 0, 0, 0, 0, eps, stream);
}

extern "C" int mmfs_rope_qk(void *q, void *k, const float *cos_table, const float *sin_table, const int64_t *position_ids,
                            long n_tokens, int T_len, int H, int hd, int q_stride, int k_stride, int pos_per_batch,
                            int dtype, void *stream) {
    MMFS_CHECK_ARG(n_tokens >= 0 && H > 0 && hd > 0 && hd % 2 == 0 && T_len > 0, "rope_qk: bad shape");
    if (n_tokens == 0) return MMFS_OK;
    MMFS_CHECK_ARG(q && k && cos_table && sin_table && position_ids, "rope_qk: null pointer argument");
    MMFS_CHECK_ARG((hd / 2) % (16 / (int)dtype_size(dtype)) == 0 && ((uintptr_t)q | (uintptr_t)k) % 16 == 0 &&
                   (q_stride * dtype_size(dtype)) % 16 == 0 && (k_stride * dtype_size(dtype)) % 16 == 0,
                   "rope_qk: head_dim/2 must be a multiple of the 16-byte vector and rows 16-byte aligned");
    return dispatch(dtype, 2, k, nullptr, position_ids, q, cos_table, sin_table, n_tokens, H, hd, q_stride, k_stride,
                    pos_per_batch, T_len, 0.f, stream);
}

extern "C" int mmfs_swiglu(const void *gate_up, void *out, long rows, int inter, int dtype, void *stream) {
    MMFS_CHECK_ARG(rows >= 0 && inter > 0, "swiglu: bad shape");
    if (rows == 0) return MMFS_OK;
    MMFS_CHECK_ARG(gate_up && out, "swiglu: null pointer argument");
    MMFS_CHECK_ARG(((uintptr_t)gate_up | (uintptr_t)out) % 16 == 0 && (inter * dtype_size(dtype)) % 16 == 0,
                   "swiglu: rows must be 16-byte aligned");
    return dispatch(dtype, 3, gate_up, nullptr, nullptr, out, nullptr, nullptr, rows, inter, 0, 0, 0, 0, 0, 0.f, stream);
}
