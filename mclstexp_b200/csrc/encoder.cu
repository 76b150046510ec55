// Memory-bound pieces of the spot encoder and projection heads (reference model.py:10-69,
// :151-168, :204-205, :230-236): position-embedding gather/add and its scatter-add backward,
// LayerNorm forward/backward, exact-erf GELU forward/backward, row softmax forward/backward,
// column sums (bias gradients).  Every kernel is a single pass over its operands with
// coalesced accesses; the contractions between them run in gemm.cu.
#include <algorithm>
#include "common.cuh"

namespace mclst {

constexpr int EN_THREADS = 256;

__device__ __forceinline__ float blk_sum(float v, float* sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < EN_THREADS / 32; ++w) r += sh[w];
  __syncthreads();
  return r;
}
__device__ __forceinline__ float blk_max(float v, float* sh) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int w = 1; w < EN_THREADS / 32; ++w) r = fmaxf(r, sh[w]);
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------- position embeddings
// out[b,:] = expr[b,:] + Wx[long(pos[b,0]),:] + Wy[long(pos[b,1]),:]   (model.py:230-235)
__global__ void __launch_bounds__(EN_THREADS)
embed_add_kernel(const float* __restrict__ expr, int64_t ld_e, const float* __restrict__ pos,
                 int64_t ld_p, const float* __restrict__ wx, const float* __restrict__ wy,
                 int table_rows, int G, float* __restrict__ out, int64_t ld_o,
                 uint32_t* __restrict__ err) {
  const int64_t b = blockIdx.x;
  // .long() truncates toward zero (model.py:230-231)
  const long long x = (long long)pos[b * ld_p], y = (long long)pos[b * ld_p + 1];
  if (x < 0 || x >= table_rows || y < 0 || y >= table_rows) {
    if (threadIdx.x == 0) atomicOr(err, 1u);    // nn.Embedding raises on out-of-range indices
    return;
  }
  const float* e = expr + b * ld_e;
  const float* a = wx + x * (int64_t)G;
  const float* c = wy + y * (int64_t)G;
  float* o = out + b * ld_o;
  for (int g = threadIdx.x; g < G; g += EN_THREADS) o[g] = e[g] + a[g] + c[g];
}

// dWx[long(pos[b,0]),:] += dh[b,:], dWy likewise (dense gradients, like nn.Embedding's default)
__global__ void __launch_bounds__(EN_THREADS)
embed_scatter_kernel(const float* __restrict__ dh, int64_t ld_d, const float* __restrict__ pos,
                     int64_t ld_p, int G, float* __restrict__ dwx, float* __restrict__ dwy) {
  const int64_t b = blockIdx.x;
  const long long x = (long long)pos[b * ld_p], y = (long long)pos[b * ld_p + 1];
  const float* d = dh + b * ld_d;
  float* ax = dwx + x * (int64_t)G;
  float* ay = dwy + y * (int64_t)G;
  for (int g = threadIdx.x; g < G; g += EN_THREADS) {
    const float v = d[g];
    atomicAdd(ax + g, v);
    atomicAdd(ay + g, v);
  }
}

// ---------------------------------------------------------------- LayerNorm
__global__ void __launch_bounds__(EN_THREADS)
layernorm_fwd_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, int C, float eps, float* __restrict__ y,
                     int64_t ld_y, float* __restrict__ mean, float* __restrict__ rstd) {
  __shared__ float sh[EN_THREADS / 32];
  const int64_t r = blockIdx.x;
  const float* p = x + r * ld_x;
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += EN_THREADS) s += p[c];
  const float mu = blk_sum(s, sh) / (float)C;
  float v = 0.f;
  for (int c = threadIdx.x; c < C; c += EN_THREADS) { const float d = p[c] - mu; v += d * d; }
  const float rs = rsqrtf(blk_sum(v, sh) / (float)C + eps);       // biased variance, like ATen
  float* o = y + r * ld_y;
  for (int c = threadIdx.x; c < C; c += EN_THREADS) o[c] = (p[c] - mu) * rs * gamma[c] + beta[c];
  if (threadIdx.x == 0) { mean[r] = mu; rstd[r] = rs; }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
__global__ void __launch_bounds__(EN_THREADS)
layernorm_bwd_dx_kernel(const float* __restrict__ dy, int64_t ld_dy, const float* __restrict__ x,
                        int64_t ld_x, const float* __restrict__ gamma, const float* __restrict__ mean,
                        const float* __restrict__ rstd, int C, float* __restrict__ dx, int64_t ld_dx) {
  __shared__ float sh[EN_THREADS / 32];
  const int64_t r = blockIdx.x;
  const float* pd = dy + r * ld_dy;
  const float* px = x + r * ld_x;
  const float mu = mean[r], rs = rstd[r];
  float s1 = 0.f, s2 = 0.f;
  for (int c = threadIdx.x; c < C; c += EN_THREADS) {
    const float g = pd[c] * gamma[c];
    s1 += g;
    s2 += g * (px[c] - mu) * rs;
  }
  s1 = blk_sum(s1, sh) / (float)C;
  s2 = blk_sum(s2, sh) / (float)C;
  float* o = dx + r * ld_dx;
  for (int c = threadIdx.x; c < C; c += EN_THREADS) {
    const float xh = (px[c] - mu) * rs;
    o[c] = rs * (pd[c] * gamma[c] - s1 - xh * s2);
  }
}

// dgamma[c] = sum_r dy*xhat, dbeta[c] = sum_r dy: deterministic two-stage column reduction.
// grid (ceil(C/32), nsplit); block 32 x 8; partial[nsplit][2][C]
__global__ void __launch_bounds__(256)
layernorm_bwd_param_partial_kernel(const float* __restrict__ dy, int64_t ld_dy,
                                   const float* __restrict__ x, int64_t ld_x,
                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                   int64_t R, int C, float* __restrict__ partial) {
  __shared__ float sg[8][33], sb[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ty = threadIdx.x >> 5;
  const int64_t rows_per = (R + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = blockIdx.y * rows_per, r1 = min(R, r0 + rows_per);
  float ag = 0.f, ab = 0.f;
  if (c < C)
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const float d = dy[r * ld_dy + c];
      ab += d;
      ag += d * (x[r * ld_x + c] - mean[r]) * rstd[r];
    }
  sg[ty][threadIdx.x & 31] = ag;
  sb[ty][threadIdx.x & 31] = ab;
  __syncthreads();
  if (ty == 0 && c < C) {
    float g = 0.f, b = 0.f;
    for (int i = 0; i < 8; ++i) { g += sg[i][threadIdx.x]; b += sb[i][threadIdx.x]; }
    partial[((size_t)blockIdx.y * 2 + 0) * C + c] = g;
    partial[((size_t)blockIdx.y * 2 + 1) * C + c] = b;
  }
}
__global__ void layernorm_bwd_param_final_kernel(const float* __restrict__ partial, int nsplit, int C,
                                                 float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float g = 0.f, b = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    g += partial[((size_t)s * 2 + 0) * C + c];
    b += partial[((size_t)s * 2 + 1) * C + c];
  }
  dgamma[c] = g;
  dbeta[c] = b;
}

// ---------------------------------------------------------------- GELU (exact erf)
__global__ void gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float v = x[i]; y[i] = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }
}
__global__ void gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                float* __restrict__ dx, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float v = x[i];
    const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * v * v);
    dx[i] = dy[i] * (cdf + v * pdf);
  }
}

// ---------------------------------------------------------------- softmax over rows (in place)
__global__ void __launch_bounds__(EN_THREADS)
softmax_fwd_kernel(float* __restrict__ s, int64_t ld, int C) {
  __shared__ float sh[EN_THREADS / 32];
  float* p = s + (int64_t)blockIdx.x * ld;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < C; c += EN_THREADS) m = fmaxf(m, p[c]);
  m = blk_max(m, sh);
  float t = 0.f;
  for (int c = threadIdx.x; c < C; c += EN_THREADS) { const float e = expf(p[c] - m); p[c] = e; t += e; }
  t = 1.f / blk_sum(t, sh);
  for (int c = threadIdx.x; c < C; c += EN_THREADS) p[c] *= t;
}
// ds = p * (dp - sum(dp * p)), written over dp
__global__ void __launch_bounds__(EN_THREADS)
softmax_bwd_kernel(const float* __restrict__ p, float* __restrict__ dp, int64_t ld, int C) {
  __shared__ float sh[EN_THREADS / 32];
  const float* pp = p + (int64_t)blockIdx.x * ld;
  float* pd = dp + (int64_t)blockIdx.x * ld;
  float t = 0.f;
  for (int c = threadIdx.x; c < C; c += EN_THREADS) t += pd[c] * pp[c];
  t = blk_sum(t, sh);
  for (int c = threadIdx.x; c < C; c += EN_THREADS) pd[c] = pp[c] * (pd[c] - t);
}

// Block-diagonal softmax for grouped attention (eval bank build, evel_her2st.py:24,47-69: the bank
// is embedded in consecutive batches of `group` spots and tokens only attend inside their batch).
// scores: [rows, cols] tiles of `cols` consecutive tokens (cols % group == 0); row r is token
// (r / cols) * cols + r % cols.  Inside the token's own group the softmax is taken over the valid
// tokens (< n_valid); every other column is set to 0 so that P V only mixes the group.
__global__ void __launch_bounds__(128)
softmax_blockdiag_kernel(float* __restrict__ s, int64_t ld, int cols, int group, int64_t n_valid) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  float* p = s + r * ld;
  const int64_t tile0 = (r / cols) * cols;            // first token of this tile
  const int i = (int)(r - tile0);
  const int g0 = (i / group) * group;
  const int g1 = (int)min((int64_t)(g0 + group), n_valid - tile0);   // valid columns of the group
  float m = -INFINITY;
  for (int c = g0 + lane; c < g1; c += 32) m = fmaxf(m, p[c]);
  m = warp_max(m);
  float t = 0.f;
  for (int c = g0 + lane; c < g1; c += 32) t += expf(p[c] - m);
  t = warp_sum(t);
  const float inv = t > 0.f ? 1.f / t : 0.f;
  for (int c = lane; c < cols; c += 32)
    p[c] = (c >= g0 && c < g1) ? expf(p[c] - m) * inv : 0.f;
}

// ---------------------------------------------------------------- column sums (bias grads)
// 32 columns per block, 32 row lanes, 8 independent loads in flight per thread (a serial walk of a
// 1024-row column measured 19 us per call, ten calls per training step); fixed summation order.
__global__ void __launch_bounds__(1024)
col_sum_kernel(const float* __restrict__ x, int64_t ld, int64_t R, int C, float* __restrict__ out) {
  __shared__ float sm[32][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ty = threadIdx.x >> 5;
  float a = 0.f;
  if (c < C) {
    const float* col = x + c;
    int64_t r = ty;
    for (; r + 7 * 32 < R; r += 8 * 32) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __ldg(col + (r + 32 * i) * ld);
#pragma unroll
      for (int i = 0; i < 8; ++i) a += v[i];
    }
    for (; r < R; r += 32) a += __ldg(col + r * ld);
  }
  sm[ty][threadIdx.x & 31] = a;
  __syncthreads();
  if (ty == 0 && c < C) {
    float t = 0.f;
    for (int i = 0; i < 32; ++i) t += sm[i][threadIdx.x];
    out[c] = t;
  }
}

}  // namespace mclst

using namespace mclst;

extern "C" int mclst_embed_add(const float* expression, int64_t ld_e, const float* position,
                               int64_t ld_p, const float* x_table, const float* y_table,
                               int table_rows, int batch, int genes, float* out, int64_t ld_o,
                               uint32_t* error_flag, mclst_stream_t stream) {
  MCLST_REQUIRE(expression && position && x_table && y_table && out && error_flag, MCLST_ERR_INVALID,
                "embed_add: null pointer");
  if (batch == 0) return 0;
  prof_mark((cudaStream_t)stream, "embed_add");
  embed_add_kernel<<<batch, EN_THREADS, 0, (cudaStream_t)stream>>>(
      expression, ld_e, position, ld_p, x_table, y_table, table_rows, genes, out, ld_o, error_flag);
  MCLST_LAUNCH_CHECK();
  return 0;
}

extern "C" int mclst_embed_add_backward(const float* d_out, int64_t ld_d, const float* position,
                                        int64_t ld_p, int table_rows, int batch, int genes,
                                        float* d_x_table, float* d_y_table, mclst_stream_t stream) {
  MCLST_REQUIRE(d_out && position && d_x_table && d_y_table, MCLST_ERR_INVALID,
                "embed_add_backward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  prof_mark(st, "embed_scatter");
  // dense gradients (nn.Embedding default, what torch.optim.Adam(weight_decay) expects)
  MCLST_CUDA(cudaMemsetAsync(d_x_table, 0, (size_t)table_rows * genes * sizeof(float), st));
  MCLST_CUDA(cudaMemsetAsync(d_y_table, 0, (size_t)table_rows * genes * sizeof(float), st));
  if (batch == 0) return 0;
  embed_scatter_kernel<<<batch, EN_THREADS, 0, st>>>(d_out, ld_d, position, ld_p, genes, d_x_table,
                                                    d_y_table);
  MCLST_LAUNCH_CHECK();
  return 0;
}

// Background zero fill with a BOUNDED grid.  The dense table gradients are 0.5 GB of zeros that
// nothing waits for until the very end of the backward pass; a full-size fill kernel (torch.zeros)
// launched early on a side stream took every CTA slot of the GPU for 2 x 37 us and the forward
// chain queued behind it (CUPTI timeline: the first LayerNorm started 33 us late).  `ctas` blocks
// trickle the stores out at ~30 GB/s each and leave the other SMs -- and the rest of their own --
// to the step.
__global__ void __launch_bounds__(256)
zero_fill_kernel(float4* __restrict__ p, size_t n16, float* __restrict__ tail, int n_tail) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p[i] = z;
  if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) tail[threadIdx.x] = 0.f;
}

extern "C" int mclst_zero_fill_background(float* x, int64_t n, int ctas, mclst_stream_t stream) {
  MCLST_REQUIRE(x && n >= 0 && ctas >= 1, MCLST_ERR_INVALID, "zero_fill_background: bad args");
  MCLST_REQUIRE(((uintptr_t)x & 15) == 0, MCLST_ERR_INVALID, "zero_fill_background: pointer not 16-byte aligned");
  if (n == 0) return 0;
  const size_t n16 = (size_t)n / 4;
  zero_fill_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(x), n16,
                                                                     x + n16 * 4, (int)((size_t)n - n16 * 4));
  MCLST_LAUNCH_CHECK();
  return 0;
}

extern "C" int mclst_embed_add_backward_accumulate(const float* d_out, int64_t ld_d, const float* position,
                                                   int64_t ld_p, int table_rows, int batch, int genes,
                                                   float* d_x_table, float* d_y_table,
                                                   mclst_stream_t stream) {
  MCLST_REQUIRE(d_out && position && d_x_table && d_y_table, MCLST_ERR_INVALID,
                "embed_add_backward_accumulate: null pointer");
  (void)table_rows;
  if (batch == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  prof_mark(st, "embed_scatter");
  embed_scatter_kernel<<<batch, EN_THREADS, 0, st>>>(d_out, ld_d, position, ld_p, genes, d_x_table,
                                                    d_y_table);
  MCLST_LAUNCH_CHECK();
  return 0;
}

extern "C" int mclst_layernorm_forward(const float* x, int64_t ld_x, const float* gamma,
                                       const float* beta, int64_t rows, int cols, float eps,
                                       float* y, int64_t ld_y, float* mean, float* rstd,
                                       mclst_stream_t stream) {
  MCLST_REQUIRE(x && gamma && beta && y && mean && rstd, MCLST_ERR_INVALID, "layernorm_forward: null");
  if (rows == 0) return 0;
  prof_mark((cudaStream_t)stream, "layernorm_fwd");
  layernorm_fwd_kernel<<<(unsigned)rows, EN_THREADS, 0, (cudaStream_t)stream>>>(
      x, ld_x, gamma, beta, cols, eps, y, ld_y, mean, rstd);
  MCLST_LAUNCH_CHECK();
  return 0;
}

extern "C" int mclst_layernorm_backward(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x,
                                        const float* gamma, const float* mean, const float* rstd,
                                        int64_t rows, int cols, float* dx, int64_t ld_dx,
                                        float* dgamma, float* dbeta, float* scratch,
                                        size_t scratch_floats, mclst_stream_t stream) {
  MCLST_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && scratch,
                MCLST_ERR_INVALID, "layernorm_backward: null");
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  prof_mark(st, "layernorm_bwd");
  layernorm_bwd_dx_kernel<<<(unsigned)rows, EN_THREADS, 0, st>>>(dy, ld_dy, x, ld_x, gamma, mean, rstd,
                                                               cols, dx, ld_dx);
  MCLST_LAUNCH_CHECK();
  int nsplit = (int)std::min<int64_t>(64, std::max<int64_t>(1, rows / 64));
  while (nsplit > 1 && (size_t)nsplit * 2 * cols > scratch_floats) nsplit >>= 1;
  MCLST_REQUIRE((size_t)nsplit * 2 * cols <= scratch_floats, MCLST_ERR_WORKSPACE,
                "layernorm_backward: scratch needs %d floats", 2 * cols);
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)nsplit);
  layernorm_bwd_param_partial_kernel<<<grid, 256, 0, st>>>(dy, ld_dy, x, ld_x, mean, rstd, rows, cols, scratch);
  MCLST_LAUNCH_CHECK();
  layernorm_bwd_param_final_kernel<<<(unsigned)ceil_div(cols, 256), 256, 0, st>>>(scratch, nsplit, cols, dgamma, dbeta);
  MCLST_LAUNCH_CHECK();
  return 0;
}

extern "C" int mclst_gelu_forward(const float* x, float* y, int64_t n, mclst_stream_t stream) {
  MCLST_REQUIRE(x && y, MCLST_ERR_INVALID, "gelu_forward: null");
  if (n == 0) return 0;
  prof_mark((cudaStream_t)stream, "gelu_fwd");
  gelu_fwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n);
  MCLST_LAUNCH_CHECK();
  return 0;
}
extern "C" int mclst_gelu_backward(const float* dy, const float* x, float* dx, int64_t n,
                                   mclst_stream_t stream) {
  MCLST_REQUIRE(dy && x && dx, MCLST_ERR_INVALID, "gelu_backward: null");
  if (n == 0) return 0;
  prof_mark((cudaStream_t)stream, "gelu_bwd");
  gelu_bwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, x, dx, n);
  MCLST_LAUNCH_CHECK();
  return 0;
}

extern "C" int mclst_softmax_forward(float* scores, int64_t ld, int64_t rows, int cols,
                                     mclst_stream_t stream) {
  MCLST_REQUIRE(scores, MCLST_ERR_INVALID, "softmax_forward: null");
  if (rows == 0) return 0;
  prof_mark((cudaStream_t)stream, "softmax_fwd");
  softmax_fwd_kernel<<<(unsigned)rows, EN_THREADS, 0, (cudaStream_t)stream>>>(scores, ld, cols);
  MCLST_LAUNCH_CHECK();
  return 0;
}
extern "C" int mclst_softmax_backward(const float* probs, float* d_probs_to_d_scores, int64_t ld,
                                      int64_t rows, int cols, mclst_stream_t stream) {
  MCLST_REQUIRE(probs && d_probs_to_d_scores, MCLST_ERR_INVALID, "softmax_backward: null");
  if (rows == 0) return 0;
  prof_mark((cudaStream_t)stream, "softmax_bwd");
  softmax_bwd_kernel<<<(unsigned)rows, EN_THREADS, 0, (cudaStream_t)stream>>>(probs, d_probs_to_d_scores, ld, cols);
  MCLST_LAUNCH_CHECK();
  return 0;
}

extern "C" int mclst_softmax_blockdiag(float* scores, int64_t ld, int64_t rows, int cols, int group,
                                       int64_t n_valid, mclst_stream_t stream) {
  MCLST_REQUIRE(scores, MCLST_ERR_INVALID, "softmax_blockdiag: null");
  MCLST_REQUIRE(cols >= 1 && group >= 1 && cols % group == 0 && rows % 4 == 0, MCLST_ERR_INVALID,
                "softmax_blockdiag: cols %d must be a multiple of group %d, rows a multiple of 4", cols, group);
  if (rows == 0) return 0;
  prof_mark((cudaStream_t)stream, "softmax_blockdiag");
  softmax_blockdiag_kernel<<<(unsigned)(rows / 4), 128, 0, (cudaStream_t)stream>>>(scores, ld, cols, group, n_valid);
  MCLST_LAUNCH_CHECK();
  return 0;
}

extern "C" int mclst_col_sum(const float* x, int64_t ld, int64_t rows, int cols, float* out,
                             mclst_stream_t stream) {
  MCLST_REQUIRE(x && out, MCLST_ERR_INVALID, "col_sum: null");
  prof_mark((cudaStream_t)stream, "col_sum");
  col_sum_kernel<<<(unsigned)ceil_div(cols, 32), 1024, 0, (cudaStream_t)stream>>>(x, ld, rows, cols, out);
  MCLST_LAUNCH_CHECK();
  return 0;
}
