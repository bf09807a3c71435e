/*
 * oracle/msda_ref.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * Scalar CPU restatement of the reference's multi-scale deformable attention
 * forward.  It follows, statement by statement,
 *
 *   /root/reference/mm_interleaved/models/utils/ops/src/cuda/ms_deform_im2col_cuda.cuh
 *       :36-87    ms_deform_attn_im2col_bilinear   (one bilinear tap set)
 *       :240-302  ms_deformable_im2col_gpu_kernel  (one thread per output scalar)
 *   /root/reference/mm_interleaved/models/utils/ops/src/cuda/ms_deform_attn_cuda.cu
 *       :21-81    host contract (shapes, zero-initialised output)
 *
 * and additionally emits the INTEGER index stream of every sampling point so the
 * CUDA path can be pinned bit-exactly on index math (in-range predicate, floor,
 * four corner-validity bits, four int offsets).
 *
 * opmath_t is float for float/half/bf16 inputs and double for double inputs
 * (at::opmath_type, cuh:32).  Half/bf16 inputs are passed in here already rounded
 * to their storage type and widened to float, exactly what `opmath_t x = data[i]`
 * does on the device (cuh:283-285).
 *
 * Build:  see oracle/Makefile  (gcc -O2 -ffp-contract=off: the blend is evaluated
 * as written, without FMA contraction, so this file is a deterministic statement
 * of the arithmetic; the device code is free to contract within the tolerance).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define IDX_FIELDS 8 /* in_range, h_low, w_low, valid_mask, ptr1, ptr2, ptr3, ptr4 */

/* ---- float (opmath = float) ------------------------------------------------ */

/* cuh:36-87.  bottom = value + (b*S + level_start)*M*D  (cuh:279) */
static float bilinear_f32(const float *bottom, int height, int width, int nheads,
                          int channels, float h, float w, int m, int c,
                          int32_t *idx /* may be NULL; 7 trailing fields */)
{
    const int h_low = (int)floorf(h);              /* cuh:41 */
    const int w_low = (int)floorf(w);              /* cuh:42 */
    const int h_high = h_low + 1;
    const int w_high = w_low + 1;

    const float lh = h - (float)h_low;             /* cuh:46 */
    const float lw = w - (float)w_low;
    const float hh = 1 - lh, hw = 1 - lw;

    const int w_stride = nheads * channels;        /* cuh:50 */
    const int h_stride = width * w_stride;
    const int h_low_ptr_offset = h_low * h_stride;
    const int h_high_ptr_offset = h_low_ptr_offset + h_stride;
    const int w_low_ptr_offset = w_low * w_stride;
    const int w_high_ptr_offset = w_low_ptr_offset + w_stride;
    const int base_ptr = m * channels + c;

    int valid = 0;
    int p1 = -1, p2 = -1, p3 = -1, p4 = -1;
    float v1 = 0;
    if (h_low >= 0 && w_low >= 0) {                /* cuh:59 */
        p1 = h_low_ptr_offset + w_low_ptr_offset + base_ptr;
        v1 = bottom[p1];
        valid |= 1;
    }
    float v2 = 0;
    if (h_low >= 0 && w_high <= width - 1) {       /* cuh:65 */
        p2 = h_low_ptr_offset + w_high_ptr_offset + base_ptr;
        v2 = bottom[p2];
        valid |= 2;
    }
    float v3 = 0;
    if (h_high <= height - 1 && w_low >= 0) {      /* cuh:71 */
        p3 = h_high_ptr_offset + w_low_ptr_offset + base_ptr;
        v3 = bottom[p3];
        valid |= 4;
    }
    float v4 = 0;
    if (h_high <= height - 1 && w_high <= width - 1) { /* cuh:77 */
        p4 = h_high_ptr_offset + w_high_ptr_offset + base_ptr;
        v4 = bottom[p4];
        valid |= 8;
    }
    if (idx) {
        idx[0] = h_low; idx[1] = w_low; idx[2] = valid;
        idx[3] = p1; idx[4] = p2; idx[5] = p3; idx[6] = p4;
    }
    const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw; /* cuh:83 */
    const float val = (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);          /* cuh:85 */
    return val;
}

/*
 * cuh:240-302 with the per-sample launch loop of cu:62-76 folded into `b`.
 * value (N,S,M,D)  shapes (L,2)=[H,W]  starts (L)  loc (N,Lq,M,L,P,2)=(x,y)
 * attn (N,Lq,M,L,P)  out (N,Lq,M*D) fp32, un-rounded opmath accumulator.
 * index_stream: NULL or int32 (N,Lq,M,L,P,IDX_FIELDS), offsets are for c = 0.
 */
void msda_ref_forward_f32(const float *value, const int64_t *shapes, const int64_t *starts,
                          const float *loc, const float *attn, float *out,
                          int32_t *index_stream,
                          int N, int S, int M, int D, int L, int Lq, int P,
                          long bq_begin, long bq_end)
{
    const int qid_stride = M * D;                               /* cuh:271 */
    /* [bq_begin, bq_end) is a slice of the flattened (b,q) space so that a host
       thread pool can split the work (libgomp is not in this image); every output
       scalar is still produced by the scalar statement sequence below, one writer
       per output (cuh:300). */
    (void)N;
    for (long bq = bq_begin; bq < bq_end; ++bq)
    for (int m = 0; m < M; ++m) {
        const int b = (int)(bq / Lq);
        const long sampling_index = bq * M + m;                 /* cuh:262 */
        for (int c = 0; c < D; ++c) {
            long data_weight_ptr = sampling_index * L * P;      /* cuh:269 */
            long data_loc_w_ptr = data_weight_ptr << 1;
            const long data_value_ptr_init_offset = (long)b * S * qid_stride;
            float col = 0;
            for (int l = 0; l < L; ++l) {
                const int level_start_id = (int)starts[l];      /* cuh:276 */
                const int spatial_h = (int)shapes[2 * l];
                const int spatial_w = (int)shapes[2 * l + 1];
                const float *data_value_ptr =
                    value + (data_value_ptr_init_offset + (long)level_start_id * qid_stride);
                for (int p = 0; p < P; ++p) {
                    const float loc_w = loc[data_loc_w_ptr];    /* cuh:283 */
                    const float loc_h = loc[data_loc_w_ptr + 1];
                    const float weight = attn[data_weight_ptr];
                    /* cuh:287-288: float*int -> float product, then "- 0.5" with a
                       double literal: evaluated in double, rounded once to opmath. */
                    const float h_im = (float)((double)(loc_h * (float)spatial_h) - 0.5);
                    const float w_im = (float)((double)(loc_w * (float)spatial_w) - 0.5);
                    int32_t *rec = NULL;
                    if (index_stream && c == 0) {
                        rec = index_stream + data_weight_ptr * IDX_FIELDS;
                        rec[0] = 0; rec[1] = 0; rec[2] = 0; rec[3] = 0;
                        rec[4] = rec[5] = rec[6] = rec[7] = -1;
                    }
                    if (h_im > -1 && w_im > -1 && h_im < spatial_h && w_im < spatial_w) { /* cuh:291 */
                        if (rec) rec[0] = 1;
                        col += bilinear_f32(data_value_ptr, spatial_h, spatial_w, M, D,
                                            h_im, w_im, m, c, rec ? rec + 1 : NULL) * weight;
                    }
                    data_weight_ptr += 1;
                    data_loc_w_ptr += 2;
                }
            }
            out[sampling_index * D + c] = col;                  /* cuh:300 */
        }
    }
}

/* ---- double (opmath = double; AT_DISPATCH_FLOATING_TYPES, cu:65) ----------- */

static double bilinear_f64(const double *bottom, int height, int width, int nheads,
                           int channels, double h, double w, int m, int c)
{
    const int h_low = (int)floor(h);
    const int w_low = (int)floor(w);
    const int h_high = h_low + 1;
    const int w_high = w_low + 1;
    const double lh = h - h_low, lw = w - w_low;
    const double hh = 1 - lh, hw = 1 - lw;
    const int w_stride = nheads * channels;
    const int h_stride = width * w_stride;
    const int h_low_ptr_offset = h_low * h_stride;
    const int h_high_ptr_offset = h_low_ptr_offset + h_stride;
    const int w_low_ptr_offset = w_low * w_stride;
    const int w_high_ptr_offset = w_low_ptr_offset + w_stride;
    const int base_ptr = m * channels + c;
    double v1 = 0, v2 = 0, v3 = 0, v4 = 0;
    if (h_low >= 0 && w_low >= 0) v1 = bottom[h_low_ptr_offset + w_low_ptr_offset + base_ptr];
    if (h_low >= 0 && w_high <= width - 1) v2 = bottom[h_low_ptr_offset + w_high_ptr_offset + base_ptr];
    if (h_high <= height - 1 && w_low >= 0) v3 = bottom[h_high_ptr_offset + w_low_ptr_offset + base_ptr];
    if (h_high <= height - 1 && w_high <= width - 1) v4 = bottom[h_high_ptr_offset + w_high_ptr_offset + base_ptr];
    const double w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
    return (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
}

void msda_ref_forward_f64(const double *value, const int64_t *shapes, const int64_t *starts,
                          const double *loc, const double *attn, double *out,
                          int N, int S, int M, int D, int L, int Lq, int P,
                          long bq_begin, long bq_end)
{
    const int qid_stride = M * D;
    (void)N;
    for (long bq = bq_begin; bq < bq_end; ++bq)
    for (int m = 0; m < M; ++m) {
        const int b = (int)(bq / Lq);
        const long sampling_index = bq * M + m;
        for (int c = 0; c < D; ++c) {
            long data_weight_ptr = sampling_index * L * P;
            long data_loc_w_ptr = data_weight_ptr << 1;
            double col = 0;
            for (int l = 0; l < L; ++l) {
                const int spatial_h = (int)shapes[2 * l];
                const int spatial_w = (int)shapes[2 * l + 1];
                const double *data_value_ptr =
                    value + ((long)b * S * qid_stride + (long)starts[l] * qid_stride);
                for (int p = 0; p < P; ++p) {
                    const double loc_w = loc[data_loc_w_ptr];
                    const double loc_h = loc[data_loc_w_ptr + 1];
                    const double weight = attn[data_weight_ptr];
                    const double h_im = loc_h * spatial_h - 0.5;
                    const double w_im = loc_w * spatial_w - 0.5;
                    if (h_im > -1 && w_im > -1 && h_im < spatial_h && w_im < spatial_w)
                        col += bilinear_f64(data_value_ptr, spatial_h, spatial_w, M, D,
                                            h_im, w_im, m, c) * weight;
                    data_weight_ptr += 1;
                    data_loc_w_ptr += 2;
                }
            }
            out[sampling_index * D + c] = col;
        }
    }
}

int msda_ref_idx_fields(void) { return IDX_FIELDS; }
