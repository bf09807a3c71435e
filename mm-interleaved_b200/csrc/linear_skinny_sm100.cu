// linear_skinny_sm100.cu -- y[M, N] = prologue(x)[M, K] . W[N, K]^T (+ residual) for M <= 8 rows: the dense linears of
// a DECODE step (LlamaAttention q/k/v and o_proj, LlamaMLP gate/up and down, decoders/modeling_llama_mmfs.py:175-189,
// 217-280, with the RMSNorm in front of them, :53-70, and the SwiGLU between them, :188-189, folded in).
//
// At M = batch (4 at cfg 5) every weight byte is used once per token: the op is a stream of W out of HBM (27 GB per
// token for the 13 B decoder), not a tensor-core problem.  What the round-2 decode profile showed for cuBLAS here:
// 5.9-6.0 TB/s on the two wide projections, 4.8 TB/s on the two N = 5120 ones (80-odd CTAs for 148 SMs), plus a
// separate RMSNorm / SwiGLU kernel in front of each pair.  This kernel:
//   * persistent, one CTA per SM; the (8-row block, KC-column K chunk) stages of W are numbered block-major and cut
//     into equal contiguous ranges, one per CTA ("stream-K"): every SM streams the same number of bytes whatever N is;
//   * warp 8 is the producer: per stage 8 bulk async copies (cp.async.bulk, one row piece of KC * 2 bytes each, SASS
//     UBLKCP) into a ring of shared-memory stages, completion on an mbarrier -- 100-160 KB in flight per SM with no
//     registers involved.  KC is the largest divisor of K that is a multiple of 256 and <= 2560 (2560 / 2304 columns =
//     5 / 4.5 KB per copy for the decoder's two K): the first version used 1 KB pieces and was ISSUE-bound on the copies
//     (16 per 16 KB stage, 2.4-3.0 TB/s).  Rows are stored KC * 2 + 16 bytes apart: ldmatrix is bank-conflict free;
//   * warps 0-7 are consumers: each takes a KC / 8-column slice of the stage, ldmatrix.x2 -> mma.sync.m16n8k16 with the
//     8 weight rows as (half of) the M side and the (<= 8) x rows as the N side, fp32 accumulators in registers across
//     the block's K chunks.  (The math is ~2 % of the tensor pipe; the legacy mma path is used because the operands are
//     already in shared memory in row-major pieces and the output tile is 8 x 8 -- tcgen05's M = 128 tile, TMEM
//     allocation and commit protocol buy nothing for an HBM-bound stream.)
//   * x is staged ONCE per CTA in shared memory by the consumer warps while the producer already streams W, through
//     the prologue: plain copy | RMSNorm(x) * weight (LlamaRMSNorm's rounding points: T(x * rstd), then * weight in T) |
//     SwiGLU of a [gate | up] row pair (T(silu(gate)) * up in T) -- the same arithmetic as the stand-alone kernels in
//     llama_ops_sm100.cu;
//   * a block that lies inside one CTA's range is finished there (sum over the 8 K-slice warps in fixed order,
//     + residual, one rounding to T); a block cut by a range boundary leaves fp32 partials in scratch and the LAST
//     CTA to arrive (a ticket per block) adds them in slot order: results do not depend on timing.  The tickets live
//     in caller-provided scratch that must be ZERO before the first call and is left zero by every call (no memset
//     node per linear in the decode graph); calls sharing a scratch buffer must be ordered (one stream).
// Roofline: HBM, N * K * sizeof(T) bytes per call.
#include "tc_common.cuh"

namespace mmfs {

constexpr int kSkRows = 8;                                // weight rows per block (rows 8-15 of the m16 tile are zero)
constexpr int kSkWarps = 8;                               // consumer warps: KC / 8 columns of the stage each
constexpr int kSkThreads = (kSkWarps + 1) * 32;           // + the producer warp
constexpr int kSkMaxM = 8;
constexpr int kSkMaxStages = 12;
constexpr int kSkMaxSmem = 227 * 1024;
constexpr int kSkTile = kSkRows * 8;                      // fp32 outputs of a block (8 weight rows x 8 x rows)
static int g_sk_kc_max = 2560, g_sk_ring_max = kSkMaxStages;

struct SkinnyArgs {
    const void *x, *w, *residual, *norm_w;
    void *y;
    float *part;                                          // [n_blocks][2][64] fp32 partial tiles
    unsigned *tickets;                                    // [n_blocks], zero on entry; the last arriver re-zeroes its ticket
    int M, N, K, prologue, ring;                          // ring = stages in shared memory
    float eps;
    int spb, total, q, rm;                                // stages per block, total stages, stages per CTA (q, +1 for c < rm)
    int xp;                                               // bytes between x rows in shared memory (K * 2 + 16)
    int kc, pitch, stage_bytes;                           // columns per stage, bytes between its rows, bytes per stage
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

__device__ __forceinline__ void ldmatrix_x2(uint32_t &r0, uint32_t &r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void sk_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                 "r"(dst), "l"(src), "r"(bytes), "r"(s_addr(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void finisher_sync() { asm volatile("bar.sync 2, 64;" ::: "memory"); }

template <typename T> __device__ __forceinline__ float sk_rnd(float x) { return to_op(from_op<T>(x)); }
__device__ __forceinline__ float sk_silu(float g) { return __fdividef(g, 1.f + __expf(-g)); }

__device__ __forceinline__ int sk_range_start(int c, int q, int rm) { return c * q + (c < rm ? c : rm); }
__device__ __forceinline__ int sk_cta_of(int f, int q, int rm) {          // the CTA whose range holds stage f
    const int cut = rm * (q + 1);
    return f < cut ? f / (q + 1) : rm + (f - cut) / q;
}

// ---- x -> shared memory through the prologue (256 consumer threads) -------------------------------------------------
template <typename T>
__device__ __forceinline__ void sk_stage_x(const SkinnyArgs &a, uint8_t *xs, float *s_red, int tid) {
    constexpr int VEC = 8;
    const int nvec = a.K / VEC;                                          // 16-byte vectors per row
    const T *x = static_cast<const T *>(a.x);
    if (a.prologue == 0) {
        for (int m = 0; m < a.M; ++m)
            for (int j = tid; j < nvec; j += 256)
                *reinterpret_cast<uint4 *>(xs + (size_t)m * a.xp + j * 16) = ldg_nc_v4(x + (size_t)m * a.K + j * VEC);
    } else if (a.prologue == 1) {                                        // RMSNorm
        float ss[kSkMaxM];
#pragma unroll
        for (int m = 0; m < kSkMaxM; ++m) ss[m] = 0.f;
#pragma unroll
        for (int m = 0; m < kSkMaxM; ++m) {
            if (m >= a.M) break;
            for (int j = tid; j < nvec; j += 256) {
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
        consumer_sync();
        const T *nw = static_cast<const T *>(a.norm_w);
#pragma unroll
        for (int m = 0; m < kSkMaxM; ++m) {
            if (m >= a.M) break;
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < kSkWarps; ++w) tot += s_red[w * kSkMaxM + m];
            const float r = rsqrtf(tot / (float)a.K + a.eps);
            for (int j = tid; j < nvec; j += 256) {
                float f[VEC], g[VEC], o[VEC];
                Vec16<T>::unpack(ldg_nc_v4(x + (size_t)m * a.K + j * VEC), f);
                Vec16<T>::unpack(ldg_nc_v4(nw + j * VEC), g);
#pragma unroll
                for (int k = 0; k < VEC; ++k) o[k] = g[k] * sk_rnd<T>(f[k] * r);
                *reinterpret_cast<uint4 *>(xs + (size_t)m * a.xp + j * 16) = Vec16<T>::pack(o);
            }
        }
    } else {                                                             // SwiGLU of [gate | up] rows of 2K columns
        for (int m = 0; m < a.M; ++m)
            for (int j = tid; j < nvec; j += 256) {
                float g[VEC], u[VEC], o[VEC];
                Vec16<T>::unpack(ldg_nc_v4(x + (size_t)m * 2 * a.K + j * VEC), g);
                Vec16<T>::unpack(ldg_nc_v4(x + (size_t)m * 2 * a.K + a.K + j * VEC), u);
#pragma unroll
                for (int k = 0; k < VEC; ++k) o[k] = sk_rnd<T>(sk_silu(g[k])) * u[k];
                *reinterpret_cast<uint4 *>(xs + (size_t)m * a.xp + j * 16) = Vec16<T>::pack(o);
            }
    }
}

template <typename T>
__global__ void __launch_bounds__(kSkThreads, 1) linear_skinny_kernel(const SkinnyArgs a) {
    extern __shared__ __align__(128) uint8_t s_dyn[];
    // layout: ring stages | x rows | reduction tiles [2][8][64] fp32 | rmsnorm scratch | barriers
    uint8_t *ring = s_dyn;
    uint8_t *xs = ring + (size_t)a.ring * a.stage_bytes;
    float *red = reinterpret_cast<float *>(xs + (size_t)a.M * a.xp);
    float *s_red = red + 2 * kSkWarps * kSkTile;
    uint64_t *full = reinterpret_cast<uint64_t *>(s_red + kSkWarps * kSkMaxM);
    uint64_t *empty = full + kSkMaxStages;
    int *s_flag = reinterpret_cast<int *>(empty + kSkMaxStages);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c = blockIdx.x;
    const int f0 = sk_range_start(c, a.q, a.rm), f1 = sk_range_start(c + 1, a.q, a.rm);
    const int n_my = f1 - f0;

    if (tid == 0) {
        for (int s = 0; s < a.ring; ++s) { bar_init(full + s, 1); bar_init(empty + s, kSkWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kSkWarps) {
        // ======================================= producer =======================================
        const uint8_t *w = static_cast<const uint8_t *>(a.w);
        int s = 0, round = 0;                                                 // ring slot, times the ring wrapped
        int blk = f0 / a.spb, kc = f0 - blk * a.spb;
        for (int it = 0; it < n_my; ++it) {
            if (round > 0) bar_wait(empty + s, (round - 1) & 1);
            if (lane == 0) bar_expect_tx(full + s, (uint32_t)(kSkRows * a.kc * 2));
            __syncwarp();
            if (lane < kSkRows)
                sk_bulk_g2s(s_addr(ring + (size_t)s * a.stage_bytes + lane * a.pitch),
                            w + ((size_t)(blk * kSkRows + lane) * a.K + (size_t)kc * a.kc) * 2, (uint32_t)(a.kc * 2), full + s);
            if (++s == a.ring) { s = 0; ++round; }
            if (++kc == a.spb) { kc = 0; ++blk; }
        }
        return;
    }

    // ========================================= consumers =========================================
    sk_stage_x<T>(a, xs, s_red, tid);
    // rows m >= M of the B operand are zero: nothing to stage (the fragment load below substitutes 0)
    consumer_sync();

    const int g = lane >> 2, t = lane & 3;
    const int wcols = a.kc >> 3;                                               // this warp's columns of a stage
    // ldmatrix.x2: lanes 0-7 -> rows 0-7 / k 0-7 (a0), lanes 8-15 -> rows 0-7 / k 8-15 (a2); lanes 16-31: ignored, valid
    const uint32_t a_lane_off = (uint32_t)((lane & 7) * a.pitch + (((lane >> 3) & 1) * 8 + warp * wcols) * 2);
    const uint32_t ring_addr = s_addr(ring);
    const uint8_t *xb = xs + (size_t)g * a.xp + (warp * wcols + 2 * t) * 2;   // this lane's B words: x[g][k + 2t], x[g][k + 2t + 8]
    const bool b_live = g < a.M;
    const int n_ch = wcols >> 4;                                               // k16 steps per warp per stage
    T *y = static_cast<T *>(a.y);
    const T *res = static_cast<const T *>(a.residual);

    float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};         // two chains: even / odd k16 steps
    int n_done = 0;                                                            // blocks finished by this CTA (red buffer parity)
    int s = 0, round = 0;
    int blk = f0 / a.spb, kc = f0 - blk * a.spb;
    for (int it = 0; it < n_my; ++it) {
        bar_wait(full + s, round & 1);
        const uint32_t st = ring_addr + (uint32_t)s * a.stage_bytes + a_lane_off;
        const uint8_t *xk = xb + (size_t)kc * a.kc * 2;
#pragma unroll 2
        for (int ch = 0; ch < n_ch; ch += 2) {
            uint32_t af[4] = {0u, 0u, 0u, 0u}, ag[4] = {0u, 0u, 0u, 0u};
            ldmatrix_x2(af[0], af[2], st + ch * 32);
            ldmatrix_x2(ag[0], ag[2], st + ch * 32 + 32);
            uint32_t b0 = 0u, b1 = 0u, b2 = 0u, b3 = 0u;
            if (b_live) {
                b0 = *reinterpret_cast<const uint32_t *>(xk + ch * 32);
                b1 = *reinterpret_cast<const uint32_t *>(xk + ch * 32 + 16);
                b2 = *reinterpret_cast<const uint32_t *>(xk + ch * 32 + 32);
                b3 = *reinterpret_cast<const uint32_t *>(xk + ch * 32 + 48);
            }
            SkMma<T>::mma(d0, af, b0, b1);
            SkMma<T>::mma(d1, ag, b2, b3);
        }
        __syncwarp();
        if (lane == 0) bar_arrive(empty + s);
        if (++s == a.ring) { s = 0; ++round; }

        const bool block_ends = (kc == a.spb - 1) || (it == n_my - 1);
        const int this_blk = blk;
        if (++kc == a.spb) { kc = 0; ++blk; }
        if (!block_ends) continue;
        // ---- this CTA's share of block `this_blk` is complete: reduce the 8 K slices ------------------------------
        float *rb = red + (n_done & 1) * kSkWarps * kSkTile + warp * kSkTile;
        rb[g * 8 + 2 * t] = d0[0] + d1[0];                                    // rows 8-15 of the tile (d[2], d[3]) are zero
        rb[g * 8 + 2 * t + 1] = d0[1] + d1[1];
        d0[0] = d0[1] = d0[2] = d0[3] = 0.f;
        d1[0] = d1[1] = d1[2] = d1[3] = 0.f;
        consumer_sync();
        if (tid < kSkTile) {
            const float *r0 = red + (n_done & 1) * kSkWarps * kSkTile + tid;
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < kSkWarps; ++w) v += r0[w * kSkTile];
            const int row = tid >> 3, m = tid & 7;
            const int first = this_blk * a.spb, last = first + a.spb - 1;     // the block's stage range
            const int c_first = sk_cta_of(first, a.q, a.rm), c_last = sk_cta_of(last, a.q, a.rm);
            bool finish = true;
            if (c_first != c_last) {                                          // cut by a range boundary: partials + ticket
                const int n_part = c_last - c_first + 1;                      // == 2 (host: every range holds >= spb stages)
                float *pp = a.part + ((size_t)this_blk * 2) * kSkTile;
                pp[(c - c_first) * kSkTile + tid] = v;
                __threadfence();
                finisher_sync();
                if (tid == 0) *s_flag = atomicAdd(a.tickets + this_blk, 1u) == (unsigned)(n_part - 1);
                finisher_sync();
                finish = *s_flag != 0;
                if (finish) {
                    __threadfence();
                    v = 0.f;
                    for (int p = 0; p < n_part; ++p) v += __ldcg(pp + p * kSkTile + tid);   // slot order: timing-independent
                    if (tid == 0) a.tickets[this_blk] = 0u;                   // zero on entry, zero on exit
                }
                finisher_sync();                                              // s_flag may be rewritten by the next block
            }
            if (finish && m < a.M) {
                const size_t o = (size_t)m * a.N + (size_t)this_blk * kSkRows + row;
                if (res != nullptr) v += to_op(res[o]);
                y[o] = from_op<T>(v);
            }
        }
        ++n_done;
    }
}

static inline size_t skinny_smem(int ring, int stage_bytes, int M, int K) {
    return (size_t)ring * stage_bytes + (size_t)M * (K * 2 + 16) + (2 * kSkWarps * kSkTile + kSkWarps * kSkMaxM) * sizeof(float) +
           2 * kSkMaxStages * sizeof(uint64_t) + 16;
}

// columns per stage: the largest divisor of K that is a multiple of 256 (8 warps x two k16 steps) and <= kc_max
static inline int skinny_kc(int K, int kc_max) {
    for (int d = (kc_max / 256) * 256; d >= 256; d -= 256)
        if (K % d == 0) return d;
    return 0;
}

template <typename T>
static int launch_skinny(SkinnyArgs a, cudaStream_t st) {
    auto kern = linear_skinny_kernel<T>;
    static bool attr_set[kMaxDevices] = {};
    const int dev = current_device();
    if (dev < 0 || dev >= kMaxDevices || !attr_set[dev]) {
        MMFS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSkMaxSmem));
        if (dev >= 0 && dev < kMaxDevices) attr_set[dev] = true;
    }
    const int n_blocks = a.N / kSkRows;
    a.xp = a.K * 2 + 16;
    // the stage: as many columns per copy as leave room for at least 3 stages next to x
    int kc = skinny_kc(a.K, g_sk_kc_max);
    while (kc > 0 && skinny_smem(3, kSkRows * (kc * 2 + 16), a.M, a.K) > (size_t)kSkMaxSmem) kc = skinny_kc(a.K, kc - 256);
    if (kc == 0) {
        set_error("linear_skinny: M = %d rows of K = %d do not fit shared memory next to a 3-stage weight ring", a.M, a.K);
        return MMFS_EUNSUPPORTED;
    }
    a.kc = kc;
    a.pitch = kc * 2 + 16;
    a.stage_bytes = kSkRows * a.pitch;
    a.spb = a.K / kc;
    a.total = n_blocks * a.spb;
    // at most one CTA per block: every range then holds >= spb stages, so a block is cut by at most ONE range boundary
    // (two partial tiles per block in scratch)
    int grid = num_sms();
    if (grid > n_blocks) grid = n_blocks;
    a.q = a.total / grid;
    a.rm = a.total % grid;
    int ring = g_sk_ring_max;
    while (ring > 3 && skinny_smem(ring, a.stage_bytes, a.M, a.K) > (size_t)kSkMaxSmem) --ring;
    if (ring > a.q + 1) ring = a.q + 1;                                    // no point in more stages than the CTA streams
    a.ring = ring;
    kern<<<grid, kSkThreads, skinny_smem(ring, a.stage_bytes, a.M, a.K), st>>>(a);
    MMFS_CUDA(cudaGetLastError());
    return MMFS_OK;
}

}  // namespace mmfs

using namespace mmfs;

extern "C" long mmfs_linear_skinny_scratch_floats(int N) {
    const long n_blocks = (N + kSkRows - 1) / kSkRows;
    return (n_blocks + 3) / 4 * 4 + n_blocks * 2 * kSkTile;                 // tickets, then two partial tiles per block
}

extern "C" int mmfs_linear_skinny_set_tuning(int kc_max, int ring_max) {
    MMFS_CHECK_ARG(kc_max == 0 || (kc_max >= 256 && kc_max <= 8192), "linear_skinny_set_tuning: kc_max 0 (default) or 256..8192");
    MMFS_CHECK_ARG(ring_max == 0 || (ring_max >= 3 && ring_max <= kSkMaxStages), "linear_skinny_set_tuning: ring_max 0 (default) or 3..12");
    g_sk_kc_max = kc_max == 0 ? 2560 : kc_max;
    g_sk_ring_max = ring_max == 0 ? kSkMaxStages : ring_max;
    return MMFS_OK;
}

extern "C" int mmfs_linear_skinny(const void *x, const void *w, void *y, const void *residual, const void *norm_weight,
                                  float *scratch, int M, int N, int K, int prologue, float eps, int dtype, void *stream) {
    MMFS_CHECK_ARG(M >= 0 && N > 0 && K > 0, "linear_skinny: bad shape");
    if (M == 0) return MMFS_OK;
    MMFS_CHECK_ARG(x && w && y && scratch, "linear_skinny: null pointer argument");
    MMFS_CHECK_ARG(prologue >= 0 && prologue <= 2 && (prologue != 1 || norm_weight != nullptr),
                   "linear_skinny: prologue 0 (none) / 1 (rmsnorm, needs norm_weight) / 2 (swiglu)");
    if (M > kSkMaxM || (dtype != MMFS_F16 && dtype != MMFS_BF16) || N % kSkRows != 0 || K % 256 != 0 ||
        ((uintptr_t)x | (uintptr_t)w | (uintptr_t)y | (uintptr_t)scratch | (uintptr_t)norm_weight) % 16 != 0) {
        set_error("linear_skinny: needs M <= %d, f16 / bf16, N %% %d == 0, K %% 256 == 0, 16-byte aligned pointers", kSkMaxM, kSkRows);
        return MMFS_EUNSUPPORTED;
    }
    SkinnyArgs a{};
    a.x = x; a.w = w; a.y = y; a.residual = residual; a.norm_w = norm_weight;
    const long n_blocks = N / kSkRows;
    a.tickets = reinterpret_cast<unsigned *>(scratch);
    a.part = scratch + (n_blocks + 3) / 4 * 4;
    a.M = M; a.N = N; a.K = K; a.prologue = prologue; a.eps = eps;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == MMFS_F16 ? launch_skinny<__half>(a, st) : launch_skinny<__nv_bfloat16>(a, st);
}
