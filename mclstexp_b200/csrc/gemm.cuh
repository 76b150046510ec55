// Split-precision tcgen05 GEMM building block (gemm.cu).
#pragma once
#include "common.cuh"

namespace mclst {

// fp16 TilePack image(s) of one operand: hi = fp16(x), lo = fp16(x - hi) (lo may be null).
struct PackedOperand {
  uint8_t* hi;
  uint8_t* lo;
  int64_t rows_pad;   // multiple of 128 (A operand) or 256 (B operand)
  int nkb;            // K / 64 (padded)
  size_t bytes;       // bytes of one image (per batch entry)
  float* inv_scale;   // [batch][rows_pad] inverse of the per-row power-of-two factor, or null
};

struct GemmParams {
  const uint8_t *a_hi, *a_lo, *b_hi, *b_lo;
  size_t a_batch_bytes, b_batch_bytes;   // stride between batch entries of the packed operands
  int nkb, nseg, batch;
  int64_t M, N;
  float* c;
  int64_t ldc;
  size_t c_batch_elems;
  float alpha;
  const float* bias;        // [N] or null
  int act;                  // 0 none, 1 exact-erf GELU (applied after bias)
  const float* residual;    // same layout as c, added after the activation, or null
  // operand factors undone by the epilogue (gemm.cu header): per-row vectors and / or a tensor-wide
  // factor derived from a device max|x| word, applied amax_pow (1 or 2) times
  const float *a_scale, *b_scale;
  size_t a_scale_batch, b_scale_batch;
  const uint32_t* amax_bits;
  int amax_pow;
  // A = [image 1 | image 2] along K: the first a_nkb1 K blocks come from a_hi / a_lo (an image
  // with a_nkb1 blocks per row block), the remaining nkb - a_nkb1 from a2_hi / a2_lo.  0 = plain.
  const uint8_t *a2_hi, *a2_lo;
  int a_nkb1;
  // a1_mn != 0: image 1 is consumed TRANSPOSED (MN-major descriptors): it is stored [K', M] with
  // a1_mn = its K blocks per row block (its row-major width / 64); output row m then is image
  // column m and the contraction runs over the image's rows.  Needs a_nkb1 > 0.
  int a1_mn;
  // fused row log-sum-exp of the scaled product (loss): lse_part [2 * ceil(N/256)][lse_ld] receives
  // one (max, sum exp(x - max)) pair per row, tile column and 128-column half (-inf, 0 when that
  // half lies beyond N is NOT written: merge only the halves that exist); diag (nullable) [M]
  // receives element (m, m + diag_offset).  c may then be null: statistics only, nothing stored.
  float2* lse_part;
  int64_t lse_ld;
  float* diag;
  int64_t diag_offset;
};

size_t packed_operand_bytes(int64_t rows, int64_t k, bool is_b, int64_t* rows_pad, int* nkb);
PackedOperand take_operand(Arena& a, int64_t rows, int64_t k, bool is_b, bool split, int batch,
                           bool row_scaled = false);
// x [rows, cols] fp32 (ld) scaled by `scale`; transpose = false: operand rows = x rows, K = x
// cols; transpose = true: operand rows = x cols, K = x rows.  Written at K-block kb_offset of
// dst (dst.nkb blocks in total), nkb_mine blocks wide, zero padded.
// Scaling: inv_scale != null -> per-row power-of-two factors found by the kernel (their inverses
// are stored there; a transposed pack then scans whole columns, one block per 32 of them);
// else amax_bits != null -> the tensor-wide factor of that max|x| word; else none.
int launch_pack_split(const float* x, int64_t rows, int64_t cols, int64_t ld, bool transpose,
                      float scale, const PackedOperand& dst, int kb_offset, int nkb_mine,
                      const uint32_t* amax_bits, float* inv_scale, cudaStream_t st, int batch = 1,
                      int64_t x_batch_elems = 0);
// atomicMax of max|x| (as float bits) into *out (zeroed by the caller)
int launch_amax_bits(const float* x, int64_t rows, int64_t cols, int64_t ld, uint32_t* out, cudaStream_t st);
int launch_gemm_tn(const GemmParams& p, cudaStream_t st);
// C_z = act(alpha * op(A_z) op(B_z)^T + bias) + residual straight from row-major fp32 tensors:
// 3xTF32 split inside the kernel (gemm_tf32.cu); *_trans = 1: the operand is stored [K, rows].
bool gemm_tf32x3_aligned(const float* A, int64_t lda, int a_trans, int64_t a_batch, const float* B, int64_t ldb,
                         int b_trans, int64_t b_batch, int64_t K);
int launch_gemm_tf32x3(const float* A, int64_t lda, int a_trans, int64_t a_batch, const float* B, int64_t ldb,
                       int b_trans, int64_t b_batch, float* C, int64_t ldc, int64_t c_batch, int64_t M,
                       int64_t N, int64_t K, int batch, float alpha, const float* bias, int act,
                       const float* residual, cudaStream_t st);

}  // namespace mclst
