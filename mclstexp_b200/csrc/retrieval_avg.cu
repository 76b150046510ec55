// Top-k weighted expression average: the per-query Python loop of the reference
// (evel_her2st.py:175-187, evel_visium.py:194-205, evel_cscc.py:198-215,
// baselines/Bleep/BLEEP_inference.ipynb cell 5) as one HBM-bound kernel.
//
// One CTA per query.  Phase A: one warp per neighbour gathers the un-normalised
// spot_key row (coalesced) and reduces the L1 / squared-L2 distance to the query.
// Phase B: weights are normalised in shared memory; every thread owns a 128-bit
// column slice of the output and streams the k expression rows with independent
// vector loads in flight (ld.global.nc, no L1 allocation: each row is used once).
//
// Algorithmic bytes per query: k*G*e (expression rows) + k*D*4 (spot_key rows, only
// for the distance / emb_pred modes) + G*o (output) -- DESIGN.md section "gather".
#include "common.cuh"
#include "retrieval.cuh"

namespace mclst {

constexpr int AVG_THREADS = 256;
constexpr int AVG_KMAX = 1024;

__device__ __forceinline__ void store_out(void* out, int64_t off, float v, int out_is_f64) {
  if (out_is_f64) reinterpret_cast<double*>(out)[off] = (double)v;
  else reinterpret_cast<float*>(out)[off] = v;
}

// distances of the k neighbours of one query; result in dist_sm[0..k)
__device__ void neighbour_distances(const float* __restrict__ spot_key, int64_t ld_key,
                                    const float* __restrict__ qrow, int dim,
                                    const int64_t* __restrict__ idx, int k, int64_t index_offset,
                                    int64_t n_bank, int p, float* dist_sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = warp; j < k; j += nw) {
    const int64_t r = idx[j] - index_offset;
    float acc = 0.f;
    if (r >= 0 && r < n_bank) {
      const float* s = spot_key + r * ld_key;
      for (int d = lane; d < dim; d += 32) {
        const float df = __ldg(s + d) - __ldg(qrow + d);
        acc += (p == 1) ? fabsf(df) : df * df;
      }
      acc = warp_sum(acc);
    } else {
      acc = -1.f;   // not owned by this shard
    }
    if (lane == 0) dist_sm[j] = acc;   // p==1: L1 norm; p==2: SQUARED L2 norm
  }
}

template <bool VEC4, bool EXPR_F64>
__global__ void __launch_bounds__(AVG_THREADS)
weighted_average_kernel(const float* __restrict__ spot_key, int64_t n_bank, int64_t ld_key,
                        const void* __restrict__ expr_, int64_t ld_expr, int genes,
                        const float* __restrict__ query, int64_t ld_query, int dim,
                        const int64_t* __restrict__ indices, const float* __restrict__ values,
                        const float* __restrict__ distances,
                        int k, int64_t index_offset, int mode,
                        void* __restrict__ out_emb, void* __restrict__ out_expr, int out_is_f64) {
  __shared__ float w_sm[AVG_KMAX];
  __shared__ int64_t idx_sm[AVG_KMAX];
  __shared__ float red[2];
  const int64_t q = blockIdx.x;
  const int tid = threadIdx.x;
  const int64_t* idx = indices + q * k;
  for (int j = tid; j < k; j += AVG_THREADS) idx_sm[j] = idx[j] - index_offset;
  const float* qrow = query + q * ld_query;
  // ---- phase A: un-normalised weights ------------------------------------------------
  if (mode == MCLST_W_INV_SQ_L1 || mode == MCLST_W_INV_SQ_L2 || mode == MCLST_W_BLEEP_EXP) {
    if (distances) {     // L1 norm (inv_sq_l1) or L2 norm, e.g. from mclst_find_matches_dist
      for (int j = tid; j < k; j += AVG_THREADS) {
        const float d = distances[q * k + j];
        w_sm[j] = mode == MCLST_W_INV_SQ_L1 ? d : d * d;      // phase A keeps SQUARED L2 norms
      }
    } else {
      neighbour_distances(spot_key, ld_key, qrow, dim, idx, k, index_offset, n_bank,
                          mode == MCLST_W_INV_SQ_L1 ? 1 : 2, w_sm);
    }
  } else if (mode == MCLST_W_SIMILARITY) {
    for (int j = tid; j < k; j += AVG_THREADS) w_sm[j] = values[q * k + j];
  } else {
    for (int j = tid; j < k; j += AVG_THREADS) w_sm[j] = 1.f;
  }
  __syncthreads();
  if (tid < 32) {
    // one warp turns distances into weights and normalises them
    float zero_cnt = 0.f;
    if (mode == MCLST_W_INV_SQ_L1 || mode == MCLST_W_INV_SQ_L2)
      for (int j = tid; j < k; j += 32) zero_cnt += (w_sm[j] == 0.f) ? 1.f : 0.f;
    zero_cnt = warp_sum(zero_cnt);
    const float d2_best = w_sm[0];
    __syncwarp();
    float sum = 0.f;
    for (int j = tid; j < k; j += 32) {
      float w = w_sm[j];
      if (mode == MCLST_W_INV_SQ_L1) {
        // reciprocal(a**2), evel_her2st.py:182; a zero distance takes all the weight
        w = zero_cnt > 0.f ? (w == 0.f ? 1.f : 0.f) : __frcp_rn(w * w);
      } else if (mode == MCLST_W_INV_SQ_L2) {
        // w_sm holds the squared L2 norm: a = sqrt(.), reciprocal(a**2)
        const float a = __fsqrt_rn(w);
        w = zero_cnt > 0.f ? (w == 0.f ? 1.f : 0.f) : __frcp_rn(a * a);
      } else if (mode == MCLST_W_BLEEP_EXP) {
        w = expf(-(w - d2_best + 1.0f));        // nb cell 5 l.41-42
      }
      w_sm[j] = w;
      sum += w;
    }
    sum = warp_sum(sum);
    if (tid == 0) red[0] = sum;
  }
  __syncthreads();
  const float inv_sum = 1.0f / red[0];
  // ---- phase B: weighted sums --------------------------------------------------------
  if (out_emb) {
    for (int d = tid; d < dim; d += AVG_THREADS) {
      float acc = 0.f;
      for (int j = 0; j < k; ++j) acc += w_sm[j] * __ldg(spot_key + idx_sm[j] * ld_key + d);
      store_out(out_emb, q * dim + d, acc * inv_sum, out_is_f64);
    }
  }
  if (EXPR_F64) {
    const double* expr = reinterpret_cast<const double*>(expr_);
    for (int g = tid; g < genes; g += AVG_THREADS) {
      double acc = 0.0;
      for (int j = 0; j < k; ++j) acc += (double)w_sm[j] * expr[idx_sm[j] * ld_expr + g];
      const double r = acc * (double)inv_sum;
      if (out_is_f64) reinterpret_cast<double*>(out_expr)[q * genes + g] = r;
      else reinterpret_cast<float*>(out_expr)[q * genes + g] = (float)r;
    }
  } else if (VEC4) {
    const float* expr = reinterpret_cast<const float*>(expr_);
    const int g4n = genes >> 2;
    for (int g4 = tid; g4 < g4n; g4 += AVG_THREADS) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int j = 0;
      for (; j + 8 <= k; j += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          v[u] = ldg_stream(reinterpret_cast<const float4*>(expr + idx_sm[j + u] * ld_expr) + g4);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float w = w_sm[j + u];
          acc.x = fmaf(w, v[u].x, acc.x); acc.y = fmaf(w, v[u].y, acc.y);
          acc.z = fmaf(w, v[u].z, acc.z); acc.w = fmaf(w, v[u].w, acc.w);
        }
      }
      for (; j < k; ++j) {
        const float4 v = ldg_stream(reinterpret_cast<const float4*>(expr + idx_sm[j] * ld_expr) + g4);
        const float w = w_sm[j];
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
        acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
      const int64_t o = q * genes + 4 * g4;
      if (out_is_f64) {
        double* op = reinterpret_cast<double*>(out_expr) + o;
        reinterpret_cast<double2*>(op)[0] = make_double2((double)(acc.x * inv_sum), (double)(acc.y * inv_sum));
        reinterpret_cast<double2*>(op)[1] = make_double2((double)(acc.z * inv_sum), (double)(acc.w * inv_sum));
      } else {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(out_expr) + o)[0] =
            make_float4(acc.x * inv_sum, acc.y * inv_sum, acc.z * inv_sum, acc.w * inv_sum);
      }
    }
  } else {
    const float* expr = reinterpret_cast<const float*>(expr_);
    for (int g = tid; g < genes; g += AVG_THREADS) {
      float acc = 0.f;
      int j = 0;
      for (; j + 4 <= k; j += 4) {
        const float v0 = __ldg(expr + idx_sm[j] * ld_expr + g);
        const float v1 = __ldg(expr + idx_sm[j + 1] * ld_expr + g);
        const float v2 = __ldg(expr + idx_sm[j + 2] * ld_expr + g);
        const float v3 = __ldg(expr + idx_sm[j + 3] * ld_expr + g);
        acc = fmaf(w_sm[j], v0, acc); acc = fmaf(w_sm[j + 1], v1, acc);
        acc = fmaf(w_sm[j + 2], v2, acc); acc = fmaf(w_sm[j + 3], v3, acc);
      }
      for (; j < k; ++j) acc = fmaf(w_sm[j], __ldg(expr + idx_sm[j] * ld_expr + g), acc);
      store_out(out_expr, q * genes + g, acc * inv_sum, out_is_f64);
    }
  }
}

// ---- sharded-bank pieces ---------------------------------------------------------------
__global__ void __launch_bounds__(AVG_THREADS)
neighbour_distance_kernel(const float* __restrict__ spot_key, int64_t n_bank, int64_t ld_key,
                          const float* __restrict__ query, int64_t ld_query, int dim,
                          const int64_t* __restrict__ indices, int k, int64_t index_offset, int p,
                          float* __restrict__ out_dist, const int* __restrict__ qlist,
                          const int* __restrict__ qcount_ptr) {
  __shared__ float d_sm[AVG_KMAX];
  int64_t q = blockIdx.x;
  if (qlist) {                         // only the listed queries (exact-fallback rows)
    if ((int)blockIdx.x >= *qcount_ptr) return;
    q = qlist[blockIdx.x];
  }
  neighbour_distances(spot_key, ld_key, query + q * ld_query, dim, indices + q * k, k,
                      index_offset, n_bank, p, d_sm);
  __syncthreads();
  for (int j = threadIdx.x; j < k; j += AVG_THREADS) {
    float d = d_sm[j];
    if (d >= 0.f && p == 2) d = __fsqrt_rn(d);
    out_dist[q * k + j] = d;          // -1 for neighbours another shard owns
  }
}

template <bool VEC4, bool EXPR_F64>
__global__ void __launch_bounds__(AVG_THREADS)
weighted_gather_kernel(const void* __restrict__ expr_, int64_t n_bank, int64_t ld_expr, int genes,
                       const int64_t* __restrict__ indices, const float* __restrict__ weights,
                       int k, int64_t index_offset, float* __restrict__ out) {
  __shared__ float w_sm[AVG_KMAX];
  __shared__ int64_t idx_sm[AVG_KMAX];
  __shared__ int n_sm;
  const int64_t q = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid == 0) {
    int n = 0;                                    // compact the neighbours this shard owns
    for (int j = 0; j < k; ++j) {
      const int64_t r = indices[q * k + j] - index_offset;
      if (r >= 0 && r < n_bank) { idx_sm[n] = r; w_sm[n] = weights[q * k + j]; ++n; }
    }
    n_sm = n;
  }
  __syncthreads();
  const int n = n_sm;
  if (EXPR_F64) {
    const double* expr = reinterpret_cast<const double*>(expr_);
    for (int g = tid; g < genes; g += AVG_THREADS) {
      double acc = 0.0;
      for (int j = 0; j < n; ++j) acc += (double)w_sm[j] * expr[idx_sm[j] * ld_expr + g];
      out[q * genes + g] = (float)acc;
    }
  } else if (VEC4) {
    const float* expr = reinterpret_cast<const float*>(expr_);
    for (int g4 = tid; g4 < (genes >> 2); g4 += AVG_THREADS) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < n; ++j) {
        const float4 v = ldg_stream(reinterpret_cast<const float4*>(expr + idx_sm[j] * ld_expr) + g4);
        const float w = w_sm[j];
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
        acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
      reinterpret_cast<float4*>(out + q * genes)[g4] = acc;
    }
  } else {
    const float* expr = reinterpret_cast<const float*>(expr_);
    for (int g = tid; g < genes; g += AVG_THREADS) {
      float acc = 0.f;
      for (int j = 0; j < n; ++j) acc = fmaf(w_sm[j], __ldg(expr + idx_sm[j] * ld_expr + g), acc);
      out[q * genes + g] = acc;
    }
  }
}

}  // namespace mclst

using namespace mclst;

extern "C" int mclst_weighted_average(const float* spot_key, int64_t n_bank, int64_t ld_key,
                                      const void* expression_key, int64_t ld_expr, int genes,
                                      int expr_is_f64, const float* image_query, int64_t n_query,
                                      int64_t ld_query, int dim, const int64_t* indices,
                                      const float* values, const float* distances, int top_k,
                                      int64_t index_offset, int weight_mode, void* out_emb,
                                      void* out_expr, int out_is_f64, mclst_stream_t stream) {
  MCLST_REQUIRE(spot_key && expression_key && image_query && indices && out_expr, MCLST_ERR_INVALID,
                "weighted_average: null pointer");
  MCLST_REQUIRE(top_k >= 1 && top_k <= AVG_KMAX, MCLST_ERR_UNSUPPORTED,
                "weighted_average: top_k %d outside [1,%d]", top_k, AVG_KMAX);
  MCLST_REQUIRE(weight_mode >= 0 && weight_mode <= MCLST_W_BLEEP_EXP, MCLST_ERR_INVALID,
                "weighted_average: bad weight mode %d", weight_mode);
  MCLST_REQUIRE(weight_mode != MCLST_W_SIMILARITY || values, MCLST_ERR_INVALID,
                "weighted_average: similarity weights need values");
  MCLST_REQUIRE(genes >= 1 && dim >= 1, MCLST_ERR_INVALID, "weighted_average: bad shape");
  if (n_query == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec4 = !expr_is_f64 && (genes % 4 == 0) && (ld_expr % 4 == 0) &&
                    ((uintptr_t)expression_key % 16 == 0) && ((uintptr_t)out_expr % 16 == 0);
  dim3 grid((unsigned)n_query), block(AVG_THREADS);
#define LAUNCH(V, F)                                                                         \
  weighted_average_kernel<V, F><<<grid, block, 0, st>>>(                                     \
      spot_key, n_bank, ld_key, expression_key, ld_expr, genes, image_query, ld_query, dim,  \
      indices, values, distances, top_k, index_offset, weight_mode, out_emb, out_expr, out_is_f64)
  prof_mark(st, "weighted_average");
  if (expr_is_f64) LAUNCH(false, true);
  else if (vec4) LAUNCH(true, false);
  else LAUNCH(false, false);
#undef LAUNCH
  prof_mark(st, "end");
  MCLST_LAUNCH_CHECK();
  return 0;
}

extern "C" int mclst_neighbor_distances(const float* spot_key, int64_t n_bank, int64_t ld_key,
                                        const float* image_query, int64_t n_query,
                                        int64_t ld_query, int dim, const int64_t* indices,
                                        int top_k, int64_t index_offset, int p, float* out_dist,
                                        mclst_stream_t stream) {
  MCLST_REQUIRE(spot_key && image_query && indices && out_dist, MCLST_ERR_INVALID,
                "neighbor_distances: null pointer");
  MCLST_REQUIRE(top_k >= 1 && top_k <= AVG_KMAX && (p == 1 || p == 2), MCLST_ERR_UNSUPPORTED,
                "neighbor_distances: bad top_k/p");
  if (n_query == 0) return 0;
  prof_mark((cudaStream_t)stream, "neighbor_distances");
  neighbour_distance_kernel<<<(unsigned)n_query, AVG_THREADS, 0, (cudaStream_t)stream>>>(
      spot_key, n_bank, ld_key, image_query, ld_query, dim, indices, top_k, index_offset, p,
      out_dist, nullptr, nullptr);
  MCLST_LAUNCH_CHECK();
  return 0;
}

int mclst::launch_neighbor_distances(const float* spot_key, int64_t n_bank, int64_t ld_key,
                                     const float* query, int64_t n_query, int64_t ld_query, int dim,
                                     const int64_t* indices, int k, int64_t index_offset, int p,
                                     float* out_dist, const int* qlist, const int* qcount_ptr,
                                     cudaStream_t st) {
  if (n_query == 0) return 0;
  neighbour_distance_kernel<<<(unsigned)n_query, AVG_THREADS, 0, st>>>(
      spot_key, n_bank, ld_key, query, ld_query, dim, indices, k, index_offset, p, out_dist, qlist,
      qcount_ptr);
  MCLST_LAUNCH_CHECK();
  return 0;
}

extern "C" int mclst_weighted_gather(const void* expression_key, int64_t n_bank, int64_t ld_expr,
                                     int genes, int expr_is_f64, const int64_t* indices,
                                     const float* weights, int64_t n_query, int top_k,
                                     int64_t index_offset, float* out_partial,
                                     mclst_stream_t stream) {
  MCLST_REQUIRE(expression_key && indices && weights && out_partial, MCLST_ERR_INVALID,
                "weighted_gather: null pointer");
  MCLST_REQUIRE(top_k >= 1 && top_k <= AVG_KMAX, MCLST_ERR_UNSUPPORTED, "weighted_gather: top_k");
  if (n_query == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec4 = !expr_is_f64 && (genes % 4 == 0) && (ld_expr % 4 == 0) &&
                    ((uintptr_t)expression_key % 16 == 0) && ((uintptr_t)out_partial % 16 == 0);
  dim3 grid((unsigned)n_query), block(AVG_THREADS);
  prof_mark(st, "weighted_gather");
  if (expr_is_f64)
    weighted_gather_kernel<false, true><<<grid, block, 0, st>>>(expression_key, n_bank, ld_expr, genes, indices, weights, top_k, index_offset, out_partial);
  else if (vec4)
    weighted_gather_kernel<true, false><<<grid, block, 0, st>>>(expression_key, n_bank, ld_expr, genes, indices, weights, top_k, index_offset, out_partial);
  else
    weighted_gather_kernel<false, false><<<grid, block, 0, st>>>(expression_key, n_bank, ld_expr, genes, indices, weights, top_k, index_offset, out_partial);
  prof_mark(st, "end");
  MCLST_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- sharded-bank merge
// Candidate lists gathered from R ranks -> global top-k by (value desc, index asc), carrying
// the neighbour distance along.  One warp per query; entries sorted in shared memory.
namespace mclst {

__global__ void __launch_bounds__(128)
merge_topk_kernel(const float* __restrict__ vals, const int64_t* __restrict__ idx,
                  const float* __restrict__ dist, int R, int64_t Q, int k,
                  float* __restrict__ out_val, int64_t* __restrict__ out_idx,
                  float* __restrict__ out_dist, int per_warp) {
  extern __shared__ __align__(16) unsigned char mg_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (q >= Q) return;
  unsigned long long* key = reinterpret_cast<unsigned long long*>(mg_smem) + (size_t)warp * per_warp;
  unsigned short* slot = reinterpret_cast<unsigned short*>(
      reinterpret_cast<unsigned long long*>(mg_smem) + (size_t)(blockDim.x >> 5) * per_warp) +
      (size_t)warp * per_warp;
  const int M = R * k;
  int P = 1;
  while (P < M) P <<= 1;
  for (int i = lane; i < P; i += 32) {
    if (i < M) {
      const int r = i / k, j = i - r * k;
      const size_t src = ((size_t)r * Q + q) * k + j;
      const float v = vals[src];
      const uint32_t kv = (v != v) ? 0xffffffffu : f2ord(v);
      key[i] = ((unsigned long long)kv << 32) | (unsigned long long)(0xffffffffu - (uint32_t)idx[src]);
    } else {
      key[i] = 0ull;
    }
    slot[i] = (unsigned short)i;
  }
  __syncwarp();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = lane; t < (P >> 1); t += 32) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = key[lo], b = key[hi];
        if ((a < b) == desc) {
          key[lo] = b; key[hi] = a;
          const unsigned short s = slot[lo]; slot[lo] = slot[hi]; slot[hi] = s;
        }
      }
      __syncwarp();
    }
  }
  for (int j = lane; j < k; j += 32) {
    const int i = slot[j];
    const int r = i / k, jj = i - r * k;
    const size_t src = ((size_t)r * Q + q) * k + jj;
    out_val[q * k + j] = vals[src];
    out_idx[q * k + j] = idx[src];
    if (out_dist) out_dist[q * k + j] = dist[src];
  }
}

// normalised weights [Q,k] from neighbour distances (L1 norm, or L2 norm) / similarities
__global__ void __launch_bounds__(128)
weights_kernel(const float* __restrict__ dist, const float* __restrict__ values, int64_t Q, int k,
               int mode, float* __restrict__ w) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= Q) return;
  const float* d = dist ? dist + q * k : nullptr;
  const float* v = values ? values + q * k : nullptr;
  float* o = w + q * k;
  float zero_cnt = 0.f;
  if (mode == MCLST_W_INV_SQ_L1 || mode == MCLST_W_INV_SQ_L2)
    for (int j = lane; j < k; j += 32) zero_cnt += (d[j] == 0.f) ? 1.f : 0.f;
  zero_cnt = warp_sum(zero_cnt);
  const float d_best = (mode == MCLST_W_BLEEP_EXP) ? d[0] : 0.f;
  float sum = 0.f;
  for (int j = lane; j < k; j += 32) {
    float x;
    if (mode == MCLST_W_INV_SQ_L1 || mode == MCLST_W_INV_SQ_L2) {
      const float a = d[j];
      x = zero_cnt > 0.f ? (a == 0.f ? 1.f : 0.f) : __frcp_rn(a * a);
    } else if (mode == MCLST_W_SIMILARITY) {
      x = v[j];
    } else if (mode == MCLST_W_BLEEP_EXP) {
      x = expf(-(d[j] * d[j] - d_best * d_best + 1.0f));
    } else {
      x = 1.f;
    }
    o[j] = x;
    sum += x;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  for (int j = lane; j < k; j += 32) o[j] *= inv;
}

}  // namespace mclst

extern "C" int mclst_merge_topk(const float* values, const int64_t* indices, const float* distances,
                                int n_lists, int64_t n_query, int top_k, float* out_values,
                                int64_t* out_indices, float* out_distances, mclst_stream_t stream) {
  MCLST_REQUIRE(values && indices && out_values && out_indices, MCLST_ERR_INVALID, "merge_topk: null");
  MCLST_REQUIRE((distances == nullptr) == (out_distances == nullptr), MCLST_ERR_INVALID,
                "merge_topk: distances in and out go together");
  MCLST_REQUIRE(n_lists >= 1 && top_k >= 1 && (int64_t)n_lists * top_k <= 16384, MCLST_ERR_UNSUPPORTED,
                "merge_topk: n_lists * top_k = %lld > 16384", (long long)n_lists * top_k);
  if (n_query == 0) return 0;
  int per_warp = 1;
  while (per_warp < n_lists * top_k) per_warp <<= 1;
  int wpb = 4;
  while (wpb > 1 && (size_t)wpb * per_warp * 10 > 160 * 1024) wpb >>= 1;
  const size_t smem = (size_t)wpb * per_warp * 10;
  static size_t attr = 0;
  if (smem > attr && smem > 48 * 1024) {
    MCLST_CUDA(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  prof_mark((cudaStream_t)stream, "merge_topk");
  merge_topk_kernel<<<(unsigned)ceil_div(n_query, wpb), wpb * 32, smem, (cudaStream_t)stream>>>(
      values, indices, distances, n_lists, n_query, top_k, out_values, out_indices, out_distances, per_warp);
  MCLST_LAUNCH_CHECK();
  return 0;
}

extern "C" int mclst_neighbor_weights(const float* distances, const float* values, int64_t n_query,
                                      int top_k, int weight_mode, float* out_weights,
                                      mclst_stream_t stream) {
  MCLST_REQUIRE(out_weights, MCLST_ERR_INVALID, "neighbor_weights: null");
  MCLST_REQUIRE(weight_mode >= 0 && weight_mode <= MCLST_W_BLEEP_EXP, MCLST_ERR_INVALID, "neighbor_weights: mode");
  const bool need_d = weight_mode == MCLST_W_INV_SQ_L1 || weight_mode == MCLST_W_INV_SQ_L2 ||
                      weight_mode == MCLST_W_BLEEP_EXP;
  MCLST_REQUIRE(!need_d || distances, MCLST_ERR_INVALID, "neighbor_weights: distances required");
  MCLST_REQUIRE(weight_mode != MCLST_W_SIMILARITY || values, MCLST_ERR_INVALID, "neighbor_weights: values required");
  if (n_query == 0) return 0;
  prof_mark((cudaStream_t)stream, "neighbor_weights");
  weights_kernel<<<(unsigned)ceil_div(n_query, 4), 128, 0, (cudaStream_t)stream>>>(
      distances, values, n_query, top_k, weight_mode, out_weights);
  MCLST_LAUNCH_CHECK();
  return 0;
}
