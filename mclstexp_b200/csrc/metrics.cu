// Evaluation metrics on the predicted expression matrix, the step immediately downstream of
// retrieval: per-gene Pearson correlation (reference
// utils.py:52-65 get_R with scipy.stats.pearsonr, one Python call per gene), per-gene mean of the
// ground truth (top-50 highly-expressed-gene selection, evel_her2st.py:201-205) and the squared /
// absolute error sums behind sklearn's MSE / MAE (evel_her2st.py:214-221).  One pass over both
// [Q,G] matrices, float64 accumulation, coalesced along genes.
#include <algorithm>
#include "common.cuh"

namespace mclst {

template <bool PRED_F64, bool TRUE_F64>
__global__ void __launch_bounds__(256)
gene_metrics_partial_kernel(const void* __restrict__ truth_, int64_t ld_t, const void* __restrict__ pred_,
                            int64_t ld_p, int64_t Q, int G, double* __restrict__ partial /*[ns][7][G]*/) {
  __shared__ double sm[8][7][33];
  const int g = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ty = threadIdx.x >> 5;
  const int64_t per = (Q + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = blockIdx.y * per, r1 = min(Q, r0 + per);
  double a[7] = {0, 0, 0, 0, 0, 0, 0};
  if (g < G)
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const double x = TRUE_F64 ? reinterpret_cast<const double*>(truth_)[r * ld_t + g]
                                : (double)reinterpret_cast<const float*>(truth_)[r * ld_t + g];
      const double y = PRED_F64 ? reinterpret_cast<const double*>(pred_)[r * ld_p + g]
                                : (double)reinterpret_cast<const float*>(pred_)[r * ld_p + g];
      a[0] += x; a[1] += y; a[2] += x * x; a[3] += y * y; a[4] += x * y;
      a[5] += (x - y) * (x - y); a[6] += fabs(x - y);
    }
  for (int i = 0; i < 7; ++i) sm[ty][i][threadIdx.x & 31] = a[i];
  __syncthreads();
  if (ty == 0 && g < G)
    for (int i = 0; i < 7; ++i) {
      double t = 0;
      for (int w = 0; w < 8; ++w) t += sm[w][i][threadIdx.x];
      partial[((size_t)blockIdx.y * 7 + i) * G + g] = t;
    }
}

__global__ void gene_metrics_final_kernel(const double* __restrict__ partial, int ns, int64_t Q, int G,
                                          double* __restrict__ mean_true, double* __restrict__ pcc,
                                          double* __restrict__ sq_err, double* __restrict__ abs_err) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  double s[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < ns; ++k)
    for (int i = 0; i < 7; ++i) s[i] += partial[((size_t)k * 7 + i) * G + g];
  const double n = (double)Q;
  const double mx = s[0] / n, my = s[1] / n;
  const double vx = s[2] - n * mx * mx, vy = s[3] - n * my * my, cxy = s[4] - n * mx * my;
  mean_true[g] = mx;
  // scipy.stats.pearsonr returns NaN for a constant input
  const double tiny = 1e-13;
  double r = (vx <= tiny * fabs(s[2]) || vy <= tiny * fabs(s[3])) ? nan("") : cxy / sqrt(vx * vy);
  if (r > 1.0) r = 1.0;
  if (r < -1.0) r = -1.0;
  pcc[g] = r;
  sq_err[g] = s[5];
  abs_err[g] = s[6];
}

}  // namespace mclst

using namespace mclst;

extern "C" int mclst_gene_metrics_scratch_doubles(int genes, size_t* n) {
  MCLST_REQUIRE(n && genes >= 1, MCLST_ERR_INVALID, "gene_metrics_scratch: bad args");
  *n = (size_t)64 * 7 * genes;
  return 0;
}

extern "C" int mclst_gene_metrics(const void* truth, int64_t ld_true, int true_is_f64,
                                  const void* pred, int64_t ld_pred, int pred_is_f64,
                                  int64_t n_spots, int genes, double* mean_true, double* pcc,
                                  double* sq_err, double* abs_err, double* scratch,
                                  mclst_stream_t stream) {
  MCLST_REQUIRE(truth && pred && mean_true && pcc && sq_err && abs_err && scratch, MCLST_ERR_INVALID,
                "gene_metrics: null pointer");
  MCLST_REQUIRE(n_spots >= 1 && genes >= 1, MCLST_ERR_INVALID, "gene_metrics: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int ns = (int)std::min<int64_t>(64, std::max<int64_t>(1, n_spots / 64));
  dim3 grid((unsigned)ceil_div(genes, 32), (unsigned)ns);
  prof_mark(st, "gene_metrics");
#define GO(P, T) gene_metrics_partial_kernel<P, T><<<grid, 256, 0, st>>>(truth, ld_true, pred, ld_pred, n_spots, genes, scratch)
  if (pred_is_f64 && true_is_f64) GO(true, true);
  else if (pred_is_f64) GO(true, false);
  else if (true_is_f64) GO(false, true);
  else GO(false, false);
#undef GO
  MCLST_LAUNCH_CHECK();
  gene_metrics_final_kernel<<<(unsigned)ceil_div(genes, 128), 128, 0, st>>>(scratch, ns, n_spots, genes,
                                                                          mean_true, pcc, sq_err, abs_err);
  prof_mark(st, "end");
  MCLST_LAUNCH_CHECK();
  return 0;
}
