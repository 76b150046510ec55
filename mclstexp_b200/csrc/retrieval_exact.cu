// Exact (brute-force) cosine top-k: the arithmetic statement of find_matches
// (reference evel_her2st.py:74-84) that every other path is verified against on
// the device, and the fallback for queries the tensor-core candidate pass cannot
// certify (sim_topk.cu).
//
//   key(q, s) = float32( (sum_d double(q_d) * double(s_d)) / (nq * ns) )
//
// with nq, ns = max(sqrt(sum_d double(x_d)^2), 1e-12) in float64 (F.normalize's eps 1e-12
// clamp): the float64 cosine of the raw float32 rows, rounded once to float32.  Ranking is (key descending, bank index ascending).  NaN ranks largest
// (ATen topk semantics).  See oracle/oracle.py::find_matches_spec.
#include "common.cuh"
#include "retrieval.cuh"

namespace mclst {

// ---------------------------------------------------------------- row norms
__global__ void __launch_bounds__(256)
row_norm_kernel(const float* __restrict__ x, int64_t rows, int64_t ld, int dim,
                double* __restrict__ nrm) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* p = x + r * ld;
  double ss = 0.0;
  for (int d = lane; d < dim; d += 32) {
    const double v = (double)__ldg(p + d);
    ss = fma(v, v, ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) nrm[r] = fmax(sqrt(ss), 1e-12);
}

int launch_row_norms(const float* x, int64_t rows, int64_t ld, int dim, double* nrm,
                     cudaStream_t st) {
  if (rows == 0) return 0;
  const int wpb = 8;
  row_norm_kernel<<<(unsigned)ceil_div(rows, wpb), wpb * 32, 0, st>>>(x, rows, ld, dim, nrm);
  MCLST_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- exact top-k
__device__ __forceinline__ uint32_t key_of(float v) {
  return (v != v) ? 0xffffffffu : f2ord(v);
}

constexpr int EX_THREADS = 256;
constexpr int EX_KMAX = 2048;   // sort buffer (power of two >= top_k)

// Block-wide radix select + ordered tie collection + bitonic sort over scores[0..N).
// Writes the k winners of this query; all threads must call.
__device__ void block_select_topk(const float* __restrict__ scores, int64_t N, int k,
                                  int64_t index_offset, int64_t* __restrict__ out_idx,
                                  float* __restrict__ out_val,
                                  unsigned long long* sel /*[EX_KMAX]*/, int* hist /*[256]*/,
                                  int* sh /*[16]*/) {
  const int tid = threadIdx.x;
  // 1) k-th largest key by 4 passes of 8-bit radix select
  uint32_t prefix = 0, mask = 0;
  int remaining = k;
  for (int pass = 3; pass >= 0; --pass) {
    hist[tid] = 0;                       // EX_THREADS == 256 bins
    __syncthreads();
    for (int64_t i = tid; i < N; i += EX_THREADS) {
      const uint32_t key = key_of(scores[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255u], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int cum = 0, b = 255;
      for (; b > 0; --b) {
        if (cum + hist[b] >= remaining) break;
        cum += hist[b];
      }
      sh[0] = b;
      sh[1] = remaining - cum;
    }
    __syncthreads();
    prefix |= (uint32_t)sh[0] << (8 * pass);
    mask |= 0xffu << (8 * pass);
    remaining = sh[1];
    __syncthreads();
  }
  const uint32_t T = prefix;          // exact k-th largest key
  const int quota = remaining;        // how many elements == T are taken (lowest indices first)
  const int n_gt = k - quota;         // elements strictly greater
  // 2) collect
  if (tid == 0) { sh[2] = 0; sh[3] = 0; }
  for (int i = tid; i < EX_KMAX; i += EX_THREADS) sel[i] = 0ull;
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  for (int64_t base = 0; base < N; base += EX_THREADS) {
    const int64_t i = base + tid;
    uint32_t key = 0;
    bool gt = false, eq = false;
    if (i < N) {
      key = key_of(scores[i]);
      gt = key > T;
      eq = key == T;
    }
    const unsigned long long packed =
        ((unsigned long long)key << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
    if (gt) sel[atomicAdd(&sh[2], 1)] = packed;
    const int eq_before = sh[3];
    if (eq_before < quota) {             // uniform: sh[3] only changes behind the barriers below
      const unsigned bal = __ballot_sync(0xffffffffu, eq);
      if (lane == 0) sh[4 + warp] = __popc(bal);
      __syncthreads();
      int off = eq_before;
      for (int w = 0; w < warp; ++w) off += sh[4 + w];
      off += __popc(bal & ((1u << lane) - 1u));
      if (eq && off < quota) sel[n_gt + off] = packed;
      __syncthreads();
      if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < EX_THREADS / 32; ++w) tot += sh[4 + w];
        sh[3] = eq_before + tot;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // 3) bitonic sort, descending, over the next power of two >= k
  int P = 1;
  while (P < k) P <<= 1;
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (P >> 1); t += EX_THREADS) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = sel[lo], b = sel[hi];
        if ((a < b) == desc) { sel[lo] = b; sel[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += EX_THREADS) {
    const unsigned long long e = sel[i];
    const uint32_t key = (uint32_t)(e >> 32);
    out_idx[i] = (int64_t)(0xffffffffu - (uint32_t)e) + index_offset;
    if (out_val) out_val[i] = (key == 0xffffffffu) ? __int_as_float(0x7fc00000) : ord2f(key);
  }
  __syncthreads();
}

template <int QB>
__global__ void __launch_bounds__(EX_THREADS)
exact_topk_kernel(const float* __restrict__ bank, int64_t N, int64_t ldb,
                  const double* __restrict__ bank_nrm,
                  const float* __restrict__ query, int64_t ldq,
                  const double* __restrict__ q_nrm, int dim,
                  const int* __restrict__ qlist, const int* __restrict__ qcount_ptr, int qcount,
                  int k, int64_t index_offset, float* __restrict__ scratch, int64_t n_pad,
                  int64_t* __restrict__ out_idx, float* __restrict__ out_val) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* qd = reinterpret_cast<double*>(smem_raw);                       // [QB][dim] raw query rows
  unsigned long long* sel = reinterpret_cast<unsigned long long*>(qd + (size_t)QB * dim);
  int* hist = reinterpret_cast<int*>(sel + EX_KMAX);
  int* sh = hist + 256;
  const int nq = qcount_ptr ? *qcount_ptr : qcount;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* my_scratch = scratch + (size_t)blockIdx.x * QB * n_pad;
  for (int g = blockIdx.x; g * QB < nq; g += gridDim.x) {
    int qs[QB];
    double qn[QB];
#pragma unroll
    for (int j = 0; j < QB; ++j) {
      const int slot = g * QB + j;
      qs[j] = slot < nq ? (qlist ? qlist[slot] : slot) : -1;
      qn[j] = qs[j] >= 0 ? q_nrm[qs[j]] : 1.0;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < QB; ++j) {
      if (qs[j] >= 0) {
        for (int d = tid; d < dim; d += EX_THREADS)
          qd[j * dim + d] = (double)query[(int64_t)qs[j] * ldq + d];
      } else {
        for (int d = tid; d < dim; d += EX_THREADS) qd[j * dim + d] = 0.0;
      }
    }
    __syncthreads();
    for (int64_t r = warp; r < N; r += EX_THREADS / 32) {
      const float* p = bank + r * ldb;
      const double bn = bank_nrm[r];
      double acc[QB];
#pragma unroll
      for (int j = 0; j < QB; ++j) acc[j] = 0.0;
      for (int d = lane; d < dim; d += 32) {
        const double s = (double)__ldg(p + d);
#pragma unroll
        for (int j = 0; j < QB; ++j) acc[j] = fma(s, qd[j * dim + d], acc[j]);
      }
#pragma unroll
      for (int j = 0; j < QB; ++j) {
        const double v = warp_sum(acc[j]);
        if (lane == 0) my_scratch[(size_t)j * n_pad + r] = (float)(v / (qn[j] * bn));
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < QB; ++j) {
      if (qs[j] < 0) continue;     // uniform across the block
      block_select_topk(my_scratch + (size_t)j * n_pad, N, k, index_offset,
                        out_idx + (int64_t)qs[j] * k,
                        out_val ? out_val + (int64_t)qs[j] * k : nullptr, sel, hist, sh);
    }
  }
}

constexpr int EX_QB = 4;

int exact_topk_ctas(int64_t n_bank, int64_t n_query) {
  // scratch is ctas * QB * n_pad floats; keep it under 256 MiB
  int64_t by_mem = (int64_t)(256ll << 20) / (int64_t)(EX_QB * 4 * align_up((size_t)n_bank, 64));
  int64_t ctas = ceil_div(n_query, EX_QB);
  if (ctas > 2 * sm_count()) ctas = 2 * sm_count();
  if (ctas > by_mem) ctas = by_mem;
  if (ctas < 1) ctas = 1;
  return (int)ctas;
}

size_t exact_topk_scratch_floats(int64_t n_bank, int64_t n_query) {
  return (size_t)exact_topk_ctas(n_bank, n_query) * EX_QB * align_up((size_t)n_bank, 64);
}

int launch_exact_topk(const float* bank, int64_t N, int64_t ldb, const double* bank_nrm,
                      const float* query, int64_t ldq, const double* q_nrm, int dim,
                      const int* qlist, const int* qcount_ptr, int qcount, int64_t n_query_cap,
                      int k, int64_t index_offset, float* scratch, int64_t* out_idx,
                      float* out_val, cudaStream_t st) {
  if (n_query_cap == 0 || N == 0) return 0;
  MCLST_REQUIRE(k <= EX_KMAX, MCLST_ERR_UNSUPPORTED, "top_k %d > %d", k, EX_KMAX);
  const int ctas = exact_topk_ctas(N, n_query_cap);
  const size_t smem = (size_t)EX_QB * dim * sizeof(double) + EX_KMAX * 8 + (256 + 16) * 4;
  MCLST_REQUIRE(smem <= 200 * 1024, MCLST_ERR_UNSUPPORTED, "dim %d too large", dim);
  static bool attr_set = false;
  if (!attr_set) {
    MCLST_CUDA(cudaFuncSetAttribute(exact_topk_kernel<EX_QB>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  exact_topk_kernel<EX_QB><<<ctas, EX_THREADS, smem, st>>>(
      bank, N, ldb, bank_nrm, query, ldq, q_nrm, dim, qlist, qcount_ptr, qcount, k, index_offset,
      scratch, (int64_t)align_up((size_t)N, 64), out_idx, out_val);
  MCLST_LAUNCH_CHECK();
  return 0;
}

}  // namespace mclst
