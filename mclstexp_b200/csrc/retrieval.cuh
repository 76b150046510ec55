// Internal declarations shared by the retrieval translation units.
#pragma once
#include "common.cuh"

namespace mclst {

int launch_row_norms(const float* x, int64_t rows, int64_t ld, int dim, float* nrm, cudaStream_t st);

int exact_topk_ctas(int64_t n_bank, int64_t n_query);
size_t exact_topk_scratch_floats(int64_t n_bank, int64_t n_query);
// qlist/qcount_ptr (device) select a subset of queries; when both are null queries
// 0..qcount-1 are processed.  n_query_cap bounds the subset size (grid / scratch sizing).
int launch_exact_topk(const float* bank, int64_t N, int64_t ldb, const float* bank_nrm,
                      const float* query, int64_t ldq, const float* q_nrm, int dim,
                      const int* qlist, const int* qcount_ptr, int qcount, int64_t n_query_cap,
                      int k, int64_t index_offset, float* scratch, int64_t* out_idx,
                      float* out_val, cudaStream_t st);

}  // namespace mclst
