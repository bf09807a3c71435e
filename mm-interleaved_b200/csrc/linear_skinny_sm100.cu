// linear_skinny_sm100.cu -- y[M, N] = prologue(x)[M, K] . W[N, K]^T (+ residual) for M <= 8 rows: the dense linears of
// a DECODE step (LlamaAttention q/k/v and o_proj, LlamaMLP gate/up and down, decoders/modeling_llama_mmfs.py:175-189,
// 217-280, with the RMSNorm in front of them, :53-70, and the SwiGLU between them, :188-189, folded in).
//
// At M = batch (4 at cfg 5) every weight byte is used once per token: the op is a stream of W out of HBM (27 GB per
// token for the 13 B decoder), not a tensor-core problem.  What the round-2 decode profile showed for cuBLAS here:
// 5.9-6.0 TB/s on the two wide projections, 4.8 TB/s on the two N = 5120 ones (80-odd CTAs for 148 SMs), plus a
// separate RMSNorm / SwiGLU kernel in front of each pair.  This kernel:
//   * persistent, one CTA of 16 warps per SM.  The work is numbered in STEPS: step (block b, J) = the 32 weight rows of
//     block b x the 512 columns J*512 .. J*512+511, block-major; the steps are cut into equal contiguous ranges, one
//     per CTA ("stream-K" at 32 KB granularity): every SM streams the same number of bytes whatever N is;
//   * in a step warp w owns columns J*512 + w*32 .. +31: lane (g = lane / 4, t = lane % 4) owns, in each of the block's
//     four 8-row groups r, the 16 bytes W[32b + 8r + g][.. + 8t .. 8t+7] (per warp and group: 8 rows x 64 contiguous
//     bytes, every sector fully used) and feeds them, as they are, to two mma.sync.m16n8k16: the k index of a dot product
//     may be permuted freely as long as both operands use the same permutation, so the lane's four registers serve as
//     the B fragments of two MMAs whose A fragments are the matching 16 bytes of x row g.  No shuffles, no unpacking,
//     no shared-memory staging of W.  Rows 8-15 of the m16 tile are unused.  (The math is ~2 % of
//     the tensor pipe; the legacy mma path is used because the output tile is 8 x 8 and the operand arrives in
//     registers -- tcgen05's M = 128 tile, TMEM allocation and commit protocol buy nothing for an HBM-bound stream.)
//   * memory-level parallelism: every lane keeps 3-4 steps x 4 cp.async (16-byte, L2-only) copies of ITS OWN future
//     operands in flight into private 16-byte slots of a shared-memory ring and reads them back itself (no barrier: a
//     thread sees its own completed copies after cp.async.wait_group) -- 96-128 KB in flight per SM, across block
//     boundaries and across the reduction barriers.  Plain loads into a register ring do not get there: a warp has six
//     scoreboards, so twelve individually tracked loads serialise (measured at 8-row blocks: 1.7 TB/s re-issuing one
//     load per step, 3.6 TB/s in groups of four; cp.async: 4.3 TB/s but ISSUE-bound at 33 instructions per 512 bytes --
//     hence four row groups per step: one x fragment, one cursor update, one commit per 2 KB).
//     (The first version of this kernel staged W through a shared-memory ring filled by cp.async.bulk row pieces and
//     ldmatrix: 2.4-3.0 TB/s with 1 KB pieces, 5.2 TB/s with 5 KB pieces -- profiles/r02_skinny_linear_tma_ring_ab.log:
//     the 1-D bulk copies cost ~30 ns + 23 ns/KB each on the SM's copy engine, an asymptote below the cuBLAS kernels.)
//   * x is staged ONCE per CTA in shared memory by the consumer warps while the producer already streams W, through
//     the prologue: plain copy | RMSNorm(x) * weight (LlamaRMSNorm's rounding points: T(x * rstd), then * weight in T) |
//     SwiGLU of a [gate | up] row pair (T(silu(gate)) * up in T) -- the same arithmetic as the stand-alone kernels in
//     llama_ops_sm100.cu;
//   * a block that lies inside one CTA's range is finished there (sum over the 16 column-slice warps in fixed order,
//     + residual, one rounding to T); a block cut by a range boundary leaves fp32 partials in scratch and the LAST
//     CTA to arrive (a ticket per block) adds them in slot order: results do not depend on timing.  The tickets live
//     in caller-provided scratch that must be ZERO before the first call and is left zero by every call (no memset
//     node per linear in the decode graph); calls sharing a scratch buffer must be ordered (one stream).
// Roofline: HBM, N * K * sizeof(T) bytes per call.
#include "common.cuh"

namespace mmfs {

constexpr int kSkRG = 4;                                  // 8-row groups per block: the x fragment of a step serves all four
constexpr int kSkRows = 8 * kSkRG;                        // weight rows per block
constexpr int kSkWarps = 16;
constexpr int kSkThreads = kSkWarps * 32;
constexpr int kSkStepCols = kSkWarps * 32;                // 512 columns per step
constexpr int kSkPlane = kSkThreads * 16;                 // one row group of a ring slot: 16 bytes per lane
constexpr int kSkSlotBytes = kSkRG * kSkPlane;            // 32 KB per step
constexpr int kSkMaxM = 8;
constexpr int kSkMaxSmem = 227 * 1024;
constexpr int kSkTile = kSkRows * 8;                      // fp32 outputs of a block (32 weight rows x 8 x rows)
constexpr int kSkXPad = 64;                               // x rows K * 2 + 64 bytes apart: two rows per LDS.128 phase, no conflict
constexpr int kSkMaxBlocks = 1 << 16;                     // tickets [kSkMaxBlocks] sit at the head of scratch for every N

struct SkinnyArgs {
    const void *x, *w, *residual, *norm_w;
    void *y;
    float *part;                                          // [n_blocks][2][256] fp32 partial tiles
    unsigned *tickets;                                    // [n_blocks], zero on entry; the last arriver re-zeroes its ticket
    int M, N, K, prologue;
    float eps;
    int spb, total, q, rm;                                // steps per block (K / 512), total steps, steps per CTA (q, +1 for c < rm)
    int xp;                                               // bytes between x rows in shared memory
};

template <typename T> struct SkMma;
template <> struct SkMma<__nv_bfloat16> {
    __device__ __forceinline__ static void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};
template <> struct SkMma<__half> {
    __device__ __forceinline__ static void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};

// 16-byte asynchronous copy global -> shared through L2 only (SASS LDGSTS.E.BYPASS.128): completion is tracked by
// commit groups, not by the warp's six scoreboards, so the depth of the prefetch is a free parameter.
__device__ __forceinline__ void sk_cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void sk_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void sk_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void finisher_sync() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

template <typename T> __device__ __forceinline__ float sk_rnd(float x) { return to_op(from_op<T>(x)); }
__device__ __forceinline__ float sk_silu(float g) { return __fdividef(g, 1.f + __expf(-g)); }

__device__ __forceinline__ int sk_range_start(int c, int q, int rm) { return c * q + (c < rm ? c : rm); }
__device__ __forceinline__ int sk_cta_of(int f, int q, int rm) {          // the CTA whose range holds stage f
    const int cut = rm * (q + 1);
    return f < cut ? f / (q + 1) : rm + (f - cut) / q;
}

// The lane's 16 bytes of x are stored as words (0, 2, 1, 3): one LDS.128 = (X0, X2, X1, X3) is then, as it is, the A
// fragment of the first MMA (a0 = X0, a2 = X1; a1 / a3 feed the unused rows 8-15) against B = (W0, W1), and X2, X3 move
// into slots 0 / 2 of a second fragment against B = (W2, W3): two register moves per 512 bytes of W per warp.
__device__ __forceinline__ uint4 sk_perm(const uint4 &v) { return make_uint4(v.x, v.z, v.y, v.w); }

// ---- x -> shared memory through the prologue (all 512 threads) ----------------------------------------------------------
template <typename T>
__device__ __forceinline__ void sk_stage_x(const SkinnyArgs &a, uint8_t *xs, float *s_red, int tid) {
    constexpr int VEC = 8;
    const int nvec = a.K / VEC;                                          // 16-byte vectors per row
    const T *x = static_cast<const T *>(a.x);
    if (a.prologue == 0) {
        for (int m = 0; m < a.M; ++m)
            for (int j = tid; j < nvec; j += kSkThreads)
                *reinterpret_cast<uint4 *>(xs + (size_t)m * a.xp + j * 16) = sk_perm(ldg_nc_v4(x + (size_t)m * a.K + j * VEC));
    } else if (a.prologue == 1) {                                        // RMSNorm
        float ss[kSkMaxM];
#pragma unroll
        for (int m = 0; m < kSkMaxM; ++m) ss[m] = 0.f;
#pragma unroll
        for (int m = 0; m < kSkMaxM; ++m) {
            if (m >= a.M) break;
            for (int j = tid; j < nvec; j += kSkThreads) {
                float f[VEC];
                Vec16<T>::unpack(ldg_nc_v4(x + (size_t)m * a.K + j * VEC), f);
#pragma unroll
                for (int k = 0; k < VEC; ++k) ss[m] = fmaf(f[k], f[k], ss[m]);
            }
        }
#pragma unroll
        for (int m = 0; m < kSkMaxM; ++m) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss[m] += __shfl_xor_sync(0xffffffffu, ss[m], o);
            if ((tid & 31) == 0) s_red[(tid >> 5) * kSkMaxM + m] = ss[m];
        }
        __syncthreads();
        const T *nw = static_cast<const T *>(a.norm_w);
#pragma unroll
        for (int m = 0; m < kSkMaxM; ++m) {
            if (m >= a.M) break;
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < kSkWarps; ++w) tot += s_red[w * kSkMaxM + m];
            const float r = rsqrtf(tot / (float)a.K + a.eps);
            for (int j = tid; j < nvec; j += kSkThreads) {
                float f[VEC], g[VEC], o[VEC];
                Vec16<T>::unpack(ldg_nc_v4(x + (size_t)m * a.K + j * VEC), f);
                Vec16<T>::unpack(ldg_nc_v4(nw + j * VEC), g);
#pragma unroll
                for (int k = 0; k < VEC; ++k) o[k] = g[k] * sk_rnd<T>(f[k] * r);
                *reinterpret_cast<uint4 *>(xs + (size_t)m * a.xp + j * 16) = sk_perm(Vec16<T>::pack(o));
            }
        }
    } else {                                                             // SwiGLU of [gate | up] rows of 2K columns
        for (int m = 0; m < a.M; ++m)
            for (int j = tid; j < nvec; j += kSkThreads) {
                float g[VEC], u[VEC], o[VEC];
                Vec16<T>::unpack(ldg_nc_v4(x + (size_t)m * 2 * a.K + j * VEC), g);
                Vec16<T>::unpack(ldg_nc_v4(x + (size_t)m * 2 * a.K + a.K + j * VEC), u);
#pragma unroll
                for (int k = 0; k < VEC; ++k) o[k] = sk_rnd<T>(sk_silu(g[k])) * u[k];
                *reinterpret_cast<uint4 *>(xs + (size_t)m * a.xp + j * 16) = sk_perm(Vec16<T>::pack(o));
            }
    }
}

// One block's share is complete in this CTA: reduce the 16 column slices, then finish the block or leave a partial.
template <typename T>
__device__ __noinline__ void sk_block_end(const SkinnyArgs &a, float *red, int *s_flag, const float (&v)[kSkRG][2], int blk) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    __syncthreads();                                                          // the previous block's tile has been read
#pragma unroll
    for (int r = 0; r < kSkRG; ++r) {                                         // tile[weight row][x row]
        red[warp * kSkTile + (r * 8 + 2 * t) * 8 + g] = v[r][0];
        red[warp * kSkTile + (r * 8 + 2 * t + 1) * 8 + g] = v[r][1];
    }
    __syncthreads();
    if (tid >= kSkTile) return;
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kSkWarps; ++w) acc += red[w * kSkTile + tid];
    const int row = tid >> 3, m = tid & 7, c = blockIdx.x;
    const int first = blk * a.spb, last = first + a.spb - 1;                  // the block's step range
    const int c_first = sk_cta_of(first, a.q, a.rm), c_last = sk_cta_of(last, a.q, a.rm);
    bool finish = true;
    if (c_first != c_last) {                                                  // cut by a range boundary: partials + ticket
        const int n_part = c_last - c_first + 1;                              // == 2 (host: every range holds >= spb steps)
        float *pp = a.part + ((size_t)blk * 2) * kSkTile;
        pp[(c - c_first) * kSkTile + tid] = acc;
        __threadfence();
        finisher_sync();
        if (tid == 0) *s_flag = atomicAdd(a.tickets + blk, 1u) == (unsigned)(n_part - 1);
        finisher_sync();
        finish = *s_flag != 0;
        if (finish) {
            __threadfence();
            acc = 0.f;
            for (int p = 0; p < n_part; ++p) acc += __ldcg(pp + p * kSkTile + tid);   // slot order: timing-independent
            if (tid == 0) a.tickets[blk] = 0u;                                // zero on entry, zero on exit
        }
        finisher_sync();                                                      // s_flag may be rewritten by the next block
    }
    if (finish && m < a.M) {
        const size_t o = (size_t)m * a.N + (size_t)blk * kSkRows + row;
        if (a.residual != nullptr) acc += to_op(static_cast<const T *>(a.residual)[o]);
        static_cast<T *>(a.y)[o] = from_op<T>(acc);
    }
}

template <typename T, int DEPTH>
__global__ void __launch_bounds__(kSkThreads, 1) linear_skinny_kernel(const __grid_constant__ SkinnyArgs a) {
    extern __shared__ __align__(128) uint8_t s_dyn[];
    // layout: W ring [DEPTH][4 row groups][512 lanes][16 B] | x rows | reduction tile [16][256] fp32 | rmsnorm scratch | flag
    uint8_t *ring = s_dyn;
    uint8_t *xs = ring + DEPTH * kSkSlotBytes;
    float *red = reinterpret_cast<float *>(xs + (size_t)a.M * a.xp);
    float *s_red = red + kSkWarps * kSkTile;
    int *s_flag = reinterpret_cast<int *>(s_red + kSkWarps * kSkMaxM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int c = blockIdx.x;
    const int f0 = sk_range_start(c, a.q, a.rm), n_my = sk_range_start(c + 1, a.q, a.rm) - f0;

    // ---- load cursor: the lane's 16 bytes of step (blk, J), row group r: W[32 blk + 8 r + g][J*512 + warp*32 + 8t ..]
    const size_t row_bytes = (size_t)a.K * 2, group_bytes = 8 * row_bytes;
    int l_j = f0 % a.spb, l_left = n_my;                                      // steps still to request
    const uint8_t *lp = static_cast<const uint8_t *>(a.w) + ((size_t)(f0 / a.spb) * kSkRows + g) * row_bytes +
                        ((size_t)l_j * kSkStepCols + warp * 32 + t * 8) * 2;
    const uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(ring) + tid * 16;
    uint32_t l_slot = ring_addr;
    auto request = [&]() {                                                    // one commit group per step, empty past the end
        if (l_left > 0) {
#pragma unroll
            for (int r = 0; r < kSkRG; ++r) sk_cp_async16(l_slot + r * kSkPlane, lp + r * group_bytes);
            --l_left;
            lp += kSkStepCols * 2;
            if (++l_j == a.spb) { l_j = 0; lp += (kSkRows - 1) * row_bytes; }   // same row of the next block
        }
        sk_commit();
        l_slot += kSkSlotBytes;
        if (l_slot == ring_addr + DEPTH * kSkSlotBytes) l_slot = ring_addr;
    };
#pragma unroll
    for (int i = 0; i < DEPTH; ++i) request();

    sk_stage_x<T>(a, xs, s_red, tid);                     // overlaps the first DEPTH steps of copies
    __syncthreads();

    // ---- compute cursor ------------------------------------------------------------------------------------------------
    // Operand roles: the x rows are the M side (row g of the tile = x row g; rows >= M repeat row 0 and rows 8-15 are
    // whatever the unused registers hold -- an output row depends on its own operand row only, and those outputs are
    // never stored), the 8 weight rows of a group are the N side: the lane's 16 bytes (W0, W1, W2, W3) are the B
    // fragments of two MMAs as they are.
    int blk = f0 / a.spb, j = f0 - blk * a.spb;
    const uint8_t *xq = xs + (size_t)(g < a.M ? g : 0) * a.xp + (warp * 32 + t * 8) * 2 + (size_t)j * (kSkStepCols * 2);
    const uint8_t *wq = ring + tid * 16, *wq_end = wq + DEPTH * kSkSlotBytes;
    const uint8_t *wc = wq;
    float d[kSkRG][2][4];                                                     // [group][chain]: d[.][.][0..1] = x row g x weight rows 2t, 2t+1
#pragma unroll
    for (int r = 0; r < kSkRG; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) d[r][0][k] = d[r][1][k] = 0.f;
    uint32_t a1[4] = {0u, 0u, 0u, 0u};                                        // slots 1, 3 stay as they are
    int it = 0;
    while (it < n_my) {
        const int seg = min(n_my - it, a.spb - j);                            // steps of this block in my range
        for (int s = 0; s < seg; ++s) {
            sk_wait_group<DEPTH - 1>();                                       // the copies of this step have landed
            const uint4 xv = *reinterpret_cast<const uint4 *>(xq);            // (X0, X2, X1, X3)
            const uint32_t a0[4] = {xv.x, xv.y, xv.z, xv.w};
            a1[0] = xv.y;
            a1[2] = xv.w;
#pragma unroll
            for (int r = 0; r < kSkRG; ++r) {
                const uint4 wv = *reinterpret_cast<const uint4 *>(wc + r * kSkPlane);
                SkMma<T>::mma(d[r][0], a0, wv.x, wv.y);
                SkMma<T>::mma(d[r][1], a1, wv.z, wv.w);
            }
            request();                                                        // refills the slot just read (issued after its readers)
            xq += kSkStepCols * 2;
            wc += kSkSlotBytes;
            if (wc == wq_end) wc = wq;
        }
        it += seg;
        j += seg;
        float v[kSkRG][2];
#pragma unroll
        for (int r = 0; r < kSkRG; ++r) {
            v[r][0] = d[r][0][0] + d[r][1][0];
            v[r][1] = d[r][0][1] + d[r][1][1];
#pragma unroll
            for (int k = 0; k < 4; ++k) d[r][0][k] = d[r][1][k] = 0.f;
        }
        sk_block_end<T>(a, red, s_flag, v, blk);
        if (j == a.spb) { j = 0; ++blk; xq -= (size_t)a.spb * (kSkStepCols * 2); }
    }
}

static inline size_t skinny_smem(int depth, int M, int K) {
    return (size_t)depth * kSkSlotBytes + (size_t)M * (K * 2 + kSkXPad) + (kSkWarps * kSkTile + kSkWarps * kSkMaxM) * sizeof(float) + 16;
}

template <typename T, int DEPTH>
static int launch_skinny_depth(const SkinnyArgs &a, int grid, cudaStream_t st) {
    auto kern = linear_skinny_kernel<T, DEPTH>;
    static bool attr_set[kMaxDevices] = {};
    const int dev = current_device();
    if (dev < 0 || dev >= kMaxDevices || !attr_set[dev]) {
        MMFS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSkMaxSmem));
        if (dev >= 0 && dev < kMaxDevices) attr_set[dev] = true;
    }
    kern<<<grid, kSkThreads, skinny_smem(DEPTH, a.M, a.K), st>>>(a);
    MMFS_CUDA(cudaGetLastError());
    return MMFS_OK;
}

template <typename T>
static int launch_skinny(SkinnyArgs a, cudaStream_t st) {
    const int n_blocks = a.N / kSkRows;
    a.xp = a.K * 2 + kSkXPad;
    a.spb = a.K / kSkStepCols;
    a.total = n_blocks * a.spb;
    // at most one CTA per block: every range then holds >= spb steps, so a block is cut by at most ONE range boundary
    // (two partial tiles per block in scratch)
    int grid = num_sms();
    if (grid > n_blocks) grid = n_blocks;
    a.q = a.total / grid;
    a.rm = a.total % grid;
    if (skinny_smem(4, a.M, a.K) <= (size_t)kSkMaxSmem) return launch_skinny_depth<T, 4>(a, grid, st);
    if (skinny_smem(3, a.M, a.K) <= (size_t)kSkMaxSmem) return launch_skinny_depth<T, 3>(a, grid, st);
    if (skinny_smem(2, a.M, a.K) <= (size_t)kSkMaxSmem) return launch_skinny_depth<T, 2>(a, grid, st);
    set_error("linear_skinny: M = %d rows of K = %d do not fit shared memory next to the copy ring", a.M, a.K);
    return MMFS_EUNSUPPORTED;
}

}  // namespace mmfs

using namespace mmfs;

extern "C" long mmfs_linear_skinny_scratch_floats(int N) {
    const long n_blocks = (N + kSkRows - 1) / kSkRows;
    // tickets at a FIXED place (the head) whatever N is -- a buffer shared between calls of different N must never see
    // one call's partial tiles where another call expects zero tickets -- then two partial tiles per block
    return kSkMaxBlocks + n_blocks * 2 * kSkTile;
}

extern "C" int mmfs_linear_skinny(const void *x, const void *w, void *y, const void *residual, const void *norm_weight,
                                  float *scratch, int M, int N, int K, int prologue, float eps, int dtype, void *stream) {
    MMFS_CHECK_ARG(M >= 0 && N > 0 && K > 0, "linear_skinny: bad shape");
    if (M == 0) return MMFS_OK;
    MMFS_CHECK_ARG(x && w && y && scratch, "linear_skinny: null pointer argument");
    MMFS_CHECK_ARG(prologue >= 0 && prologue <= 2 && (prologue != 1 || norm_weight != nullptr),
                   "linear_skinny: prologue 0 (none) / 1 (rmsnorm, needs norm_weight) / 2 (swiglu)");
    if (M > kSkMaxM || (dtype != MMFS_F16 && dtype != MMFS_BF16) || N % kSkRows != 0 || N / kSkRows > kSkMaxBlocks || K % kSkStepCols != 0 ||
        ((uintptr_t)x | (uintptr_t)w | (uintptr_t)y | (uintptr_t)scratch | (uintptr_t)norm_weight) % 16 != 0) {
        set_error("linear_skinny: needs M <= %d, f16 / bf16, N %% %d == 0, K %% %d == 0, 16-byte aligned pointers", kSkMaxM, kSkRows, kSkStepCols);
        return MMFS_EUNSUPPORTED;
    }
    SkinnyArgs a{};
    a.x = x; a.w = w; a.y = y; a.residual = residual; a.norm_w = norm_weight;
    a.tickets = reinterpret_cast<unsigned *>(scratch);
    a.part = scratch + kSkMaxBlocks;
    a.M = M; a.N = N; a.K = K; a.prologue = prologue; a.eps = eps;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == MMFS_F16 ? launch_skinny<__half>(a, st) : launch_skinny<__nv_bfloat16>(a, st);
}
