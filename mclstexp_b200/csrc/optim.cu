// Adam with L2 weight decay for the two position-embedding tables (reference: train.py:118-120,
// torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-3) over x_embed / y_embed,
// model.py:204-205: 2 x 65536 x G parameters) -- SURVEY.md section 8(f) rank 3.
//
// A training step touches at most B rows of each table, but dense Adam rewrites all 65536: with
// weight decay every row moves every step (g = wd * p even where the data gradient is zero), so
// rows cannot simply be skipped.  They can be DEFERRED: a row's trajectory between two steps
// that touch it depends only on its own (p, m, v) and on the per-step scalars, so it is replayed
// in registers when the row is next needed ("catch-up") instead of being streamed through HBM on
// every step.  The replay executes, step by step, exactly the arithmetic of the dense kernel in
// this file (same device function), so lazy == dense bit for bit; dense == torch.optim.Adam up
// to fp32 rounding (tests/test_optim_gpu.py).
//
//   coef[s]   per-step scalars of step s (1-based), appended by every step call
//   last[r]   number of steps already applied to row r
#include <algorithm>
#include <cmath>
#include "common.cuh"
#include "../../include/mclst_b200.h"

namespace mclst {

struct AdamCoef { float neg_step_size, bc2_sqrt, one_minus_b1, b2, one_minus_b2, eps, wd, pad; };

// One Adam step on one element; the op order follows torch/optim/adam.py (_single_tensor_adam):
// grad += wd*p; m.lerp_(grad, 1-b1); v = v*b2 + (1-b2)*g*g; p += -step_size * m / (sqrt(v)/bc2_sqrt + eps)
__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g_data, const AdamCoef& c) {
  const float g = __fmaf_rn(c.wd, p, g_data);
  m = __fmaf_rn(g - m, c.one_minus_b1, m);
  v = __fmaf_rn(c.one_minus_b2 * g, g, v * c.b2);
  const float denom = __fdiv_rn(__fsqrt_rn(v), c.bc2_sqrt) + c.eps;
  p = __fmaf_rn(c.neg_step_size, __fdiv_rn(m, denom), p);
}

constexpr int OP_THREADS = 256;

__global__ void __launch_bounds__(OP_THREADS)
adam_dense_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                  float* __restrict__ v, int64_t n, const AdamCoef* __restrict__ coef, int step) {
  const AdamCoef c = coef[step];
  for (int64_t i = (int64_t)blockIdx.x * OP_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * OP_THREADS) {
    float pi = p[i], mi = m[i], vi = v[i];
    adam_update(pi, mi, vi, g ? g[i] : 0.f, c);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

__global__ void coef_write_kernel(AdamCoef* coef, int step, AdamCoef c) { coef[step] = c; }

// first[row] = smallest token index whose position (column `col`) truncates to `row`
__global__ void __launch_bounds__(OP_THREADS)
row_first_kernel(const float* __restrict__ pos, int64_t ld_p, int col, int batch, int table_rows,
                 int* __restrict__ first, uint32_t* __restrict__ err) {
  const int b = blockIdx.x * OP_THREADS + threadIdx.x;
  if (b >= batch) return;
  const long long r = (long long)pos[b * ld_p + col];
  if (r < 0 || r >= table_rows) { atomicOr(err, 1u); return; }
  atomicMin(first + r, b);
}

// One block per token; only the first token of every distinct row works.  Replays the steps the
// row has missed (last[r]+1 .. upto, data gradient zero) and, when d_out != nullptr, applies step
// `upto + 1` with the row's gradient = sum of d_out over its tokens, in token order.
__global__ void __launch_bounds__(OP_THREADS)
lazy_rows_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v,
                 int* __restrict__ last, const int* __restrict__ first, const float* __restrict__ pos,
                 int64_t ld_p, int col, int batch, int table_rows, int G,
                 const float* __restrict__ d_out, int64_t ld_d, const AdamCoef* __restrict__ coef,
                 int upto_arg, const int* __restrict__ upto_dev) {
  // the step counter may live on the device: a captured CUDA graph replays this launch with the
  // counter's CURRENT value instead of the one baked into the launch arguments
  const int upto = upto_dev ? *upto_dev : upto_arg;
  __shared__ int dup[OP_THREADS];
  __shared__ int n_dup;
  const int b = blockIdx.x;
  const long long r = (long long)pos[b * ld_p + col];
  if (r < 0 || r >= table_rows || first[r] != b) return;
  const int from = last[r];
  if (d_out == nullptr && from >= upto) return;          // nothing to replay
  float* wr = w + r * (int64_t)G;
  float* mr = m + r * (int64_t)G;
  float* vr = v + r * (int64_t)G;
  for (int e0 = 0; e0 < G; e0 += OP_THREADS) {
    const int e = e0 + threadIdx.x;
    float pi = 0.f, mi = 0.f, vi = 0.f;
    if (e < G) { pi = wr[e]; mi = mr[e]; vi = vr[e]; }
    for (int s = from + 1; s <= upto; ++s) {
      const AdamCoef c = coef[s];
      if (e < G) adam_update(pi, mi, vi, 0.f, c);
    }
    if (d_out != nullptr) {
      // gradient of this row: its tokens in ascending order (deterministic), found in chunks
      float acc = 0.f;
      for (int t0 = b; t0 < batch; t0 += OP_THREADS) {
        __syncthreads();
        if (threadIdx.x == 0) n_dup = 0;
        __syncthreads();
        const int t = t0 + threadIdx.x;
        const bool mine = t < batch && (long long)pos[t * ld_p + col] == r;
        // ordered compaction of the matching tokens of this chunk
        const unsigned bal = __ballot_sync(0xffffffffu, mine);
        __shared__ int warp_cnt[OP_THREADS / 32];
        if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = __popc(bal);
        __syncthreads();
        int base = 0;
        for (int wv = 0; wv < (int)(threadIdx.x >> 5); ++wv) base += warp_cnt[wv];
        if (mine) dup[base + __popc(bal & ((1u << (threadIdx.x & 31)) - 1u))] = t;
        if (threadIdx.x == 0) {
          int tot = 0;
          for (int wv = 0; wv < OP_THREADS / 32; ++wv) tot += warp_cnt[wv];
          n_dup = tot;
        }
        __syncthreads();
        if (e < G)
          for (int i = 0; i < n_dup; ++i) acc += d_out[(int64_t)dup[i] * ld_d + e];
      }
      const AdamCoef c = coef[upto + 1];
      if (e < G) adam_update(pi, mi, vi, acc, c);
    }
    if (e < G) { wr[e] = pi; mr[e] = mi; vr[e] = vi; }
  }
  __syncthreads();
  if (threadIdx.x == 0) last[r] = d_out != nullptr ? upto + 1 : upto;
}

// every row to step `upto` (before reading the whole table: state_dict, evaluation)
__global__ void __launch_bounds__(OP_THREADS)
lazy_flush_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v,
                  int* __restrict__ last, int table_rows, int G, const AdamCoef* __restrict__ coef,
                  int upto) {
  for (int r = blockIdx.x; r < table_rows; r += gridDim.x) {
    const int from = last[r];
    __syncthreads();
    if (from >= upto) continue;
    float* wr = w + r * (int64_t)G;
    float* mr = m + r * (int64_t)G;
    float* vr = v + r * (int64_t)G;
    for (int e = threadIdx.x; e < G; e += OP_THREADS) {
      float pi = wr[e], mi = mr[e], vi = vr[e];
      for (int s = from + 1; s <= upto; ++s) {
        const AdamCoef c = coef[s];
        adam_update(pi, mi, vi, 0.f, c);
      }
      wr[e] = pi; mr[e] = mi; vr[e] = vi;
    }
    __syncthreads();
    if (threadIdx.x == 0) last[r] = upto;
  }
}

static AdamCoef make_coef(int step, double lr, double beta1, double beta2, double eps, double wd) {
  // the scalars torch computes in Python doubles (adam.py: bias_correction1/2, step_size)
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  AdamCoef c;
  c.neg_step_size = (float)(-(lr / bc1));
  c.bc2_sqrt = (float)sqrt(bc2);
  c.one_minus_b1 = (float)(1.0 - beta1);
  c.b2 = (float)beta2;
  c.one_minus_b2 = (float)(1.0 - beta2);
  c.eps = (float)eps;
  c.wd = (float)wd;
  c.pad = 0.f;
  return c;
}

}  // namespace mclst

using namespace mclst;

extern "C" size_t mclst_adam_coef_bytes(int max_steps) { return (size_t)(max_steps + 1) * sizeof(AdamCoef); }

extern "C" int mclst_adam_set_step(void* coef_table, int max_steps, int step, double lr, double beta1,
                                   double beta2, double eps, double weight_decay, mclst_stream_t stream) {
  MCLST_REQUIRE(coef_table && step >= 1 && step <= max_steps, MCLST_ERR_INVALID,
                "adam_set_step: step %d outside 1..%d", step, max_steps);
  MCLST_REQUIRE(lr >= 0 && beta1 >= 0 && beta1 < 1 && beta2 >= 0 && beta2 < 1 && eps >= 0 && weight_decay >= 0,
                MCLST_ERR_INVALID, "adam_set_step: bad hyper-parameters");
  coef_write_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((AdamCoef*)coef_table, step,
                                                       make_coef(step, lr, beta1, beta2, eps, weight_decay));
  MCLST_LAUNCH_CHECK();
  return 0;
}

extern "C" int mclst_adam_dense(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                                int64_t n, const void* coef_table, int step, mclst_stream_t stream) {
  MCLST_REQUIRE(param && exp_avg && exp_avg_sq && coef_table && n >= 0 && step >= 1, MCLST_ERR_INVALID,
                "adam_dense: bad args");
  if (n == 0) return 0;
  prof_mark((cudaStream_t)stream, "adam_dense");
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, OP_THREADS), (int64_t)sm_count() * 16);
  adam_dense_kernel<<<grid, OP_THREADS, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n,
                                                                   (const AdamCoef*)coef_table, step);
  MCLST_LAUNCH_CHECK();
  return 0;
}

extern "C" int mclst_adam_lazy_rows(float* table, float* exp_avg, float* exp_avg_sq, int* last_step,
                                    int* first_scratch, int table_rows, int genes, const float* position,
                                    int64_t ld_p, int column, int batch, const float* d_out, int64_t ld_d,
                                    const void* coef_table, int steps_done, const int* steps_done_dev,
                                    uint32_t* error_flag, mclst_stream_t stream) {
  MCLST_REQUIRE(table && exp_avg && exp_avg_sq && last_step && first_scratch && position && coef_table &&
                error_flag, MCLST_ERR_INVALID, "adam_lazy_rows: null pointer");
  MCLST_REQUIRE(table_rows >= 1 && genes >= 1 && batch >= 0 && (column == 0 || column == 1) && steps_done >= 0,
                MCLST_ERR_INVALID, "adam_lazy_rows: bad shape");
  if (batch == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  prof_mark(st, d_out ? "adam_lazy_step" : "adam_lazy_catch_up");
  MCLST_CUDA(cudaMemsetAsync(first_scratch, 0x7f, (size_t)table_rows * sizeof(int), st));
  row_first_kernel<<<(unsigned)ceil_div((int64_t)batch, OP_THREADS), OP_THREADS, 0, st>>>(
      position, ld_p, column, batch, table_rows, first_scratch, error_flag);
  MCLST_LAUNCH_CHECK();
  lazy_rows_kernel<<<batch, OP_THREADS, 0, st>>>(table, exp_avg, exp_avg_sq, last_step, first_scratch,
                                                 position, ld_p, column, batch, table_rows, genes, d_out,
                                                 ld_d, (const AdamCoef*)coef_table, steps_done, steps_done_dev);
  MCLST_LAUNCH_CHECK();
  return 0;
}

extern "C" int mclst_adam_lazy_flush(float* table, float* exp_avg, float* exp_avg_sq, int* last_step,
                                     int table_rows, int genes, const void* coef_table, int steps_done,
                                     mclst_stream_t stream) {
  MCLST_REQUIRE(table && exp_avg && exp_avg_sq && last_step && coef_table && table_rows >= 1 && genes >= 1 &&
                steps_done >= 0, MCLST_ERR_INVALID, "adam_lazy_flush: bad args");
  prof_mark((cudaStream_t)stream, "adam_lazy_flush");
  const unsigned grid = (unsigned)std::min<int64_t>(table_rows, (int64_t)sm_count() * 8);
  lazy_flush_kernel<<<grid, OP_THREADS, 0, (cudaStream_t)stream>>>(table, exp_avg, exp_avg_sq, last_step,
                                                                   table_rows, genes,
                                                                   (const AdamCoef*)coef_table, steps_done);
  MCLST_LAUNCH_CHECK();
  return 0;
}
