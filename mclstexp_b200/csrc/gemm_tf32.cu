// fp32 matmul straight from the row-major fp32 tensors: 3xTF32 split INSIDE the kernel.
//
// The packed route (gemm.cu) pays one pack launch per product that reads and writes both operands
// before the tensor cores start.  For short contractions that pack is most of the time, and for a
// large batched operand that is used once (attention probabilities, 32 MB) it is an HBM round trip
// for nothing.  Here the operand tiles never exist in global memory:
//
//   8 worker warps  load a 128 x 32 (A) and 64 x 32 (B) fp32 K block from the row-major tensors
//                   (plain or transposed storage) into registers, one K block ahead, split every
//                   value into hi = x with the 13 low mantissa bits cleared (exactly a TF32 number)
//                   and lo = x - hi (exact in fp32), and write both as K-major SWIZZLE_128B tiles
//                   (128-byte rows = 32 floats) into a 4-stage shared-memory ring
//                   (fence.proxy.async + one mbarrier arrival per warp);
//   1 MMA warp      walks the loop in uniform control flow, one elected lane issues
//                   tcgen05.mma.kind::tf32: hi*hi + lo*hi + hi*lo into one fp32 TMEM accumulator
//                   (128 lanes x 64 columns); tcgen05.commit frees the stage;
//   the same 8 warps drain the accumulator (tcgen05.ld) through alpha / bias / exact-erf GELU /
//                   residual to global memory.
//
// TF32 has the fp32 exponent, so there is nothing to rescale (the fp16 split needs power-of-two
// row factors); the dropped lo*lo term and the 10-bit rounding of lo are 2^-21 relative to |a||b|.
// A transposed operand (stored [K, rows]) is read with the lanes along its contiguous dimension and
// transposed on the way into shared memory (4 consecutive K of one row form one 16-byte chunk), so
// both operands are always K-major for the tensor core and one shared-memory descriptor type
// serves all four storage combinations.
//
// What bounds it (clock64 stamps per K block from a -DTF_TIMING build, tools/tf32_timing.py,
// profiles/r2_tf32_kblock_timing.log): ~1200 cycles per K block on the worker side, of which
// 650-950 are the twelve 16-byte shared-memory stores per thread (48 KiB of hi / lo tiles per
// K block at ~64 B/clk: every STS.128 of a warp is four wavefronts), 200 the load issue, 100 fence +
// arrival; the MMA side needs 560 (issue) and waits for the workers.  Ablations agree: without the
// fence, with a deeper register ring, with the issue moved to an elected lane of a uniform warp
// (each worth 1-2 us of 30), nothing changes the slope, and 128 x 128 tiles are slower per CTA.
// Converting inside the kernel costs 0.18 B of shared-memory stores per MAC because every CTA
// re-converts its A block for its own 64 columns; the packed route converts every element once.
// mclst_matmul therefore routes here only the shapes where the pack launch costs more than that
// (gemm.cu: mclst_matmul).
#include <algorithm>
#include "common.cuh"
#include "gemm.cuh"
#include "umma.cuh"

namespace mclst {
using namespace ptx;

#ifdef TF_TIMING
__device__ long long g_tf_dbg[64 * 16];
#define TF_STAMP(kb, slot) do { if (blockIdx.x == 0 && (kb) < 64) g_tf_dbg[(kb) * 16 + (slot)] = clock64(); } while (0)
#else
#define TF_STAMP(kb, slot) do {} while (0)
#endif

constexpr int TF_BM = 128, TF_BK = 32;              // tile width BN = 64 or 128: template parameter
constexpr int TF_WORKERS = 256;                       // 8 warps: loaders/splitters, then the epilogue
constexpr int TF_THREADS = 32 + TF_WORKERS;           // warp 0: TMEM owner + MMA issuer
constexpr int TF_A_BYTES = TF_BM * TF_BK * 4;         // 16 KiB: one term (hi or lo) of the A block
// per tile width: B term bytes, stage = A hi | A lo | B hi | B lo, ring depth, dynamic shared memory
__host__ __device__ constexpr int tf_b_bytes(int bn) { return bn * TF_BK * 4; }
__host__ __device__ constexpr int tf_stage_bytes(int bn) { return 2 * TF_A_BYTES + 2 * tf_b_bytes(bn); }
__host__ __device__ constexpr int tf_stages(int bn) { return bn == 64 ? 4 : 3; }
__host__ __device__ constexpr int tf_smem(int bn) { return tf_stages(bn) * tf_stage_bytes(bn) + 1024 /*barriers*/ + 1024 /*align*/; }
constexpr int TF_A_CHUNKS = TF_BM * 8 / TF_WORKERS;   // 16-byte chunks per thread per K block: 4

struct Tf32Params {
  const float *A, *B;
  int64_t lda, ldb, a_batch, b_batch;
  int a_trans, b_trans;
  float* C;
  int64_t ldc, c_batch;
  int64_t M, N, K;
  int batch;
  float alpha;
  const float* bias;
  int act;
  const float* residual;
};

// kind::tf32 instruction descriptor: fp32 accumulator, A and B TF32, both K-major.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __noinline__ float gelu_erf_tf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// Slow path of an mbarrier wait, out of line (bounded: a pipeline bug traps instead of hanging).
__device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void mbar_wait_fast(uint64_t* bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}

// One thread's share of one operand K block: NCH chunks of 4 consecutive K values of one row.
//   plain storage   [rows, K]: chunk q = tid + 256 i -> row q / 8, chunk column q % 8 (8 lanes read
//                   one 128-byte row segment, float4 when aligned);
//   transposed      [K, rows]: row = tid % ROWS, chunk column tid / ROWS + (256 / ROWS) i (a warp
//                   reads 32 consecutive rows of one K index: coalesced scalar loads).
// The K loop is issue-bound (8 warps share 4 schedulers), so everything that does not change from
// K block to K block is computed once per tile (source pointers, row validity) or once per kernel
// (shared-memory offsets), and only the last K block of a tile checks the K bound.
template <int ROWS, int NCH, bool TRANS, bool VEC>
struct OperandLoader {
  int64_t ld, K;
  const float* src[NCH];  // chunk i at K block 0 of the current tile
  uint32_t soff[NCH];     // byte offset of chunk i inside a hi / lo tile (K-major SWIZZLE_128B)
  int kc[NCH];            // first K index of chunk i inside the K block
  bool rv[NCH];           // row inside the matrix

  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      int r, c;
      if (!TRANS) {
        const int q = (int)threadIdx.x - 32 + TF_WORKERS * i;
        r = q >> 3;
        c = q & 7;
      } else {
        const int t = (int)threadIdx.x - 32;
        r = t & (ROWS - 1);
        c = (t / ROWS) + (TF_WORKERS / ROWS) * i;
      }
      kc[i] = 4 * c;
      soff[i] = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
    }
  }
  __device__ __forceinline__ void set_tile(const float* origin, int64_t rows_left) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int r = (int)(soff[i] >> 7);
      rv[i] = r < rows_left;
      src[i] = origin + (TRANS ? (int64_t)kc[i] * ld + r : (int64_t)r * ld + kc[i]);
    }
  }
  template <bool CHECK>
  __device__ __forceinline__ void load(int kb, float (&v)[NCH][4]) const {
    const int64_t k0 = (int64_t)kb * TF_BK;
    const int64_t step = TRANS ? k0 * ld : k0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const float* s = src[i] + step;
      const int64_t k = k0 + kc[i];
      if (!TRANS && VEC) {      // aligned rows and K % 4 == 0: a chunk is wholly inside or wholly outside
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rv[i] && (!CHECK || k < K)) t = __ldg(reinterpret_cast<const float4*>(s));
        v[i][0] = t.x; v[i][1] = t.y; v[i][2] = t.z; v[i][3] = t.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          v[i][j] = (rv[i] && (!CHECK || k + j < K)) ? __ldg(s + (TRANS ? j * ld : j)) : 0.f;
      }
    }
  }
  // hi = x with the 13 low mantissa bits cleared, lo = x - hi (a non-finite x stays non-finite in
  // at least one of the two: inf -> (inf, NaN), NaN -> (NaN or inf, NaN))
  __device__ __forceinline__ void store(uint32_t hi, uint32_t lo, const float (&v)[NCH][4]) const {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      float h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        h[j] = __uint_as_float(__float_as_uint(v[i][j]) & 0xffffe000u);
        l[j] = v[i][j] - h[j];
      }
      sts128(hi + soff[i], h[0], h[1], h[2], h[3]);
      sts128(lo + soff[i], l[0], l[1], l[2], l[3]);
    }
  }
};

// One accumulator row (TMEM lane) x 32 columns per thread: alpha, bias, GELU, residual, store.  Kept
// out of line: the main loop is unrolled over the register ring and an inlined epilogue per ring
// slot pushed the kernel past the instruction cache (ncu: stall_no_instruction was the top reason).
__device__ __noinline__ void tf32_epilogue(const Tf32Params& p, uint32_t taddr, int64_t m, int64_t n0, int z) {
  uint32_t v[32];
  tmem_ld_32x32(taddr, v);
  tmem_ld_wait();
  tc_fence_before();
  if (m >= p.M || n0 >= p.N) return;
  float* crow = p.C + (int64_t)z * p.c_batch + m * p.ldc + n0;
  const float* rrow = p.residual ? p.residual + (int64_t)z * p.c_batch + m * p.ldc + n0 : nullptr;
  const bool vec = (p.ldc % 4 == 0) && ((uintptr_t)p.C % 16 == 0) && (p.c_batch % 4 == 0) &&
                   (!p.residual || (uintptr_t)p.residual % 16 == 0) && n0 + 32 <= p.N;
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[i] = __uint_as_float(v[j + i]) * p.alpha;
      if (p.bias && n0 + j + i < p.N) o[i] += __ldg(p.bias + n0 + j + i);
      if (p.act == 1) o[i] = gelu_erf_tf(o[i]);
    }
    if (vec) {
      float4 w = make_float4(o[0], o[1], o[2], o[3]);
      if (rrow) {
        const float4 rr = *reinterpret_cast<const float4*>(rrow + j);
        w.x += rr.x; w.y += rr.y; w.z += rr.z; w.w += rr.w;
      }
      *reinterpret_cast<float4*>(crow + j) = w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (n0 + j + i < p.N) crow[j + i] = o[i] + (rrow ? rrow[j + i] : 0.f);
    }
  }
}

template <bool A_T, bool B_T, bool VEC, int TF_BN>
__global__ void __launch_bounds__(TF_THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ Tf32Params p) {
  constexpr int TF_B_BYTES = tf_b_bytes(TF_BN), TF_STAGE_BYTES = tf_stage_bytes(TF_BN), TF_STAGES = tf_stages(TF_BN);
  constexpr int TF_B_CHUNKS = TF_BN * 8 / TF_WORKERS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + TF_STAGES * TF_STAGE_BYTES);
  uint64_t* bar_empty = bar_full + TF_STAGES;
  uint64_t* bar_tfull = bar_empty + TF_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (int)((p.K + TF_BK - 1) / TF_BK);
  const int mt = (int)((p.M + TF_BM - 1) / TF_BM), nt = (int)((p.N + TF_BN - 1) / TF_BN);
  const int64_t tiles = (int64_t)mt * nt * p.batch;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TF_STAGES; ++i) { mbar_init(&bar_full[i], TF_WORKERS / 32); mbar_init(&bar_empty[i], 1); }
    mbar_init(bar_tfull, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, TF_BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // The whole warp walks the loop (uniform control flow keeps the descriptors in uniform
    // registers) and ONE elected lane issues: with `if (lane == 0)` around the loop every 32-cycle
    // MMA cost ~100 cycles of single-thread descriptor arithmetic and R2UR moves, and the issuing
    // thread -- not the tensor pipe, not the loaders -- set the pace of the K loop.
    constexpr uint32_t idesc = make_idesc_tf32(TF_BM, TF_BN);
    const uint32_t leader = elect_one();
    const uint32_t smem_base = smem_u32(smem);
    // descriptor of a K-major SWIZZLE_128B tile at byte offset 0 of the ring; the start-address
    // field (bits 0..13, in 16-byte units) advances by plain addition
    const uint64_t desc0 = make_smem_desc_sw128(smem_base);
    uint32_t stage = 0, phase = 0;
    for (int64_t g = blockIdx.x; g < tiles; g += gridDim.x) {
      // (the accumulator is free: the workers publish the first K block of a tile only after
      // they have drained the previous tile)
      for (int kb = 0; kb < nkb; ++kb) {
        if (lane == 0) TF_STAMP(kb, 8);
        mbar_wait_fast(&bar_full[stage], phase);
        if (lane == 0) TF_STAMP(kb, 9);
        tc_fence_after();
        const uint64_t sd = desc0 + (uint64_t)((stage * TF_STAGE_BYTES) >> 4);
        if (leader) {
#pragma unroll
          for (int seg = 0; seg < 3; ++seg) {              // hi*hi, lo*hi, hi*lo
            const uint64_t ad = sd + (uint64_t)((seg == 1 ? TF_A_BYTES : 0) >> 4);
            const uint64_t bd = sd + (uint64_t)((2 * TF_A_BYTES + (seg == 2 ? TF_B_BYTES : 0)) >> 4);
#pragma unroll
            for (int k8 = 0; k8 < TF_BK / 8; ++k8)         // 8 TF32 = 32 bytes of K per instruction
              mma_tf32_ss(tmem_base, ad + (uint64_t)(k8 * 2), bd + (uint64_t)(k8 * 2), idesc,
                          (kb | seg | k8) != 0 ? 1u : 0u);
          }
          mma_commit(&bar_empty[stage]);
        }
        __syncwarp();
        if (lane == 0) TF_STAMP(kb, 10);
        if (++stage == TF_STAGES) { stage = 0; phase ^= 1; }
      }
      if (leader) mma_commit(bar_tfull);
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ workers
    OperandLoader<TF_BM, TF_A_CHUNKS, A_T, VEC> la;
    OperandLoader<TF_BN, TF_B_CHUNKS, B_T, VEC> lb;
    la.ld = p.lda; la.K = p.K;
    lb.ld = p.ldb; lb.K = p.K;
    la.init();
    lb.init();
    const uint32_t smem_base = smem_u32(smem);
    const int quad = warp & 3, half = (warp - 1) >> 2;    // TMEM lane quadrant of this warp, column half
    const int nkb_full = (int)(p.K / TF_BK);              // K blocks that need no K bound check
    uint32_t stage = 0, phase = 0, n_done = 0;
    float va[2][TF_A_CHUNKS][4], vb[2][TF_B_CHUNKS][4];   // double buffer: one K block of loads in flight
    auto load = [&](int kb, float (&a)[TF_A_CHUNKS][4], float (&b)[TF_B_CHUNKS][4]) {
      if (kb < nkb_full) { la.template load<false>(kb, a); lb.template load<false>(kb, b); }
      else { la.template load<true>(kb, a); lb.template load<true>(kb, b); }
    };
    auto publish = [&](const float (&a)[TF_A_CHUNKS][4], const float (&b)[TF_B_CHUNKS][4], int kbs) {
      if (threadIdx.x == 32) TF_STAMP(kbs, 1);
      mbar_wait_fast(&bar_empty[stage], phase ^ 1);
      if (threadIdx.x == 32) TF_STAMP(kbs, 2);
      const uint32_t st = smem_base + stage * TF_STAGE_BYTES;
      la.store(st, st + TF_A_BYTES, a);
      lb.store(st + 2 * TF_A_BYTES, st + 2 * TF_A_BYTES + TF_B_BYTES, b);
      if (threadIdx.x == 32) TF_STAMP(kbs, 3);
      fence_proxy_async_smem();                           // generic-proxy writes -> visible to the MMA
      if (threadIdx.x == 32) TF_STAMP(kbs, 4);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[stage]);       // one arrival per warp
      if (threadIdx.x == 32) TF_STAMP(kbs, 5);
      if (++stage == TF_STAGES) { stage = 0; phase ^= 1; }
    };
#pragma unroll 1
    for (int64_t g = blockIdx.x; g < tiles; g += gridDim.x) {
      const int mb = (int)(g % mt), nb = (int)((g / mt) % nt), z = (int)(g / ((int64_t)mt * nt));
      const int64_t m0 = (int64_t)mb * TF_BM, n0 = (int64_t)nb * TF_BN;
      la.set_tile(p.A + (int64_t)z * p.a_batch + (A_T ? m0 : m0 * p.lda), p.M - m0);
      lb.set_tile(p.B + (int64_t)z * p.b_batch + (B_T ? n0 : n0 * p.ldb), p.N - n0);
      load(0, va[0], vb[0]);
#pragma unroll 1
      for (int kb = 0; kb < nkb; kb += 2) {
        if (threadIdx.x == 32) TF_STAMP(kb, 0);
        if (kb + 1 < nkb) load(kb + 1, va[1], vb[1]);
        publish(va[0], vb[0], kb);
        if (kb + 1 < nkb) {
          if (threadIdx.x == 32) TF_STAMP(kb + 1, 0);
          if (kb + 2 < nkb) load(kb + 2, va[0], vb[0]);
          publish(va[1], vb[1], kb + 1);
        }
      }
      mbar_wait_fast(bar_tfull, n_done & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < TF_BN / 64; ++c)                // this warp's half of the columns, 32 at a time
        tf32_epilogue(p, tmem_base + ((uint32_t)(quad * 32) << 16) + half * (TF_BN / 2) + c * 32,
                      m0 + quad * 32 + lane, n0 + half * (TF_BN / 2) + c * 32, z);
      ++n_done;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TF_BN);
}

// float4 loads of the plain operands: 16-byte aligned rows, and K % 4 == 0 so that a 4-wide chunk is
// wholly inside or wholly outside the contraction
bool gemm_tf32x3_aligned(const float* A, int64_t lda, int a_trans, int64_t a_batch, const float* B, int64_t ldb,
                         int b_trans, int64_t b_batch, int64_t K) {
  const bool a_ok = a_trans || (lda % 4 == 0 && (uintptr_t)A % 16 == 0 && a_batch % 4 == 0);
  const bool b_ok = b_trans || (ldb % 4 == 0 && (uintptr_t)B % 16 == 0 && b_batch % 4 == 0);
  return a_ok && b_ok && K % 4 == 0;
}

int launch_gemm_tf32x3(const float* A, int64_t lda, int a_trans, int64_t a_batch, const float* B, int64_t ldb,
                       int b_trans, int64_t b_batch, float* C, int64_t ldc, int64_t c_batch, int64_t M,
                       int64_t N, int64_t K, int batch, float alpha, const float* bias, int act,
                       const float* residual, cudaStream_t st) {
  Tf32Params p{};
  p.A = A; p.B = B; p.lda = lda; p.ldb = ldb; p.a_batch = a_batch; p.b_batch = b_batch;
  p.a_trans = a_trans; p.b_trans = b_trans; p.C = C; p.ldc = ldc; p.c_batch = c_batch;
  p.M = M; p.N = N; p.K = K; p.batch = batch; p.alpha = alpha; p.bias = bias; p.act = act;
  p.residual = residual;
  prof_mark(st, "gemm_tf32");
  const int64_t tiles = ceil_div(M, TF_BM) * ceil_div(N, 64) * batch;
  const unsigned grid = (unsigned)std::min<int64_t>(tiles, sm_count());
  auto launch = [&](auto kern) -> int {
    MCLST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tf_smem(64)));
    kern<<<grid, TF_THREADS, tf_smem(64), st>>>(p);
    MCLST_LAUNCH_CHECK();
    return 0;
  };
  MCLST_REQUIRE(gemm_tf32x3_aligned(A, lda, a_trans, a_batch, B, ldb, b_trans, b_batch, K), MCLST_ERR_INVALID,
                "gemm_tf32x3: operands not 16-byte aligned");
  // (128 x 128 tiles were built and measured: 37 us against 28 us on the 1024 x 1000 x 1000 product --
  // one K block of register loads in flight per thread is what bounds the K loop, and a wider tile
  // only lengthens it)
  if (!a_trans && !b_trans) return launch(gemm_tf32x3_kernel<false, false, true, 64>);
  if (!a_trans && b_trans) return launch(gemm_tf32x3_kernel<false, true, true, 64>);
  if (a_trans && !b_trans) return launch(gemm_tf32x3_kernel<true, false, true, 64>);
  return launch(gemm_tf32x3_kernel<true, true, true, 64>);
}

}  // namespace mclst

#ifdef TF_TIMING
extern "C" int mclst_debug_tf32_timing(long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, mclst::g_tf_dbg, sizeof(long long) * (size_t)n);
}
#endif
