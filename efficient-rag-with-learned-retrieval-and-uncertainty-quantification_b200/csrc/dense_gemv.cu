// Dense scoring for tiny query batches (K2a): bandwidth-bound bf16 GEMV with fused top-k.
//
// Replaces DenseIndex.search -> collection.query (rag_uq/streaming_index.py:353-370): the
// reference asks ChromaDB's approximate HNSW for n_results neighbours in cosine space and
// returns 1 - distance; here every passage row is scored exactly.
//
// Data layout: passages [n_rows, dim] bf16 row-major (unit rows), queries [B, dim] bf16.
// Algorithmic bytes: n_rows * dim * 2 per call (queries and candidates are noise).
//
// Kernel shape: one resident wave (SMs x the blocks of 256 threads that fit per SM, at most 4).  A warp owns 4 rows at a
// time: each lane issues 4 x (dim/256) independent 16-byte streaming loads (L1 bypass), so
// a full SM keeps ~100 KB in flight, multiplies against the queries held in shared memory
// as fp32, and reduces with shuffles.  Scores go straight into the block's running top-k
// (one per query); the score vector never exists in memory.
#include <cstdlib>

#include <algorithm>

#include "common.cuh"
#include "topk.cuh"

namespace ragb {

constexpr int GV_THREADS = 256;
constexpr int GV_WARPS = GV_THREADS / 32;
constexpr int GV_ROWS_PER_WARP = 4;
constexpr int GV_TILE = 512;  // rows per block between two reserve() calls
constexpr int GV_MAX_DIM = 2048;

template <int NQ, int CHUNKS>  // CHUNKS = ceil(dim / 256): 16-byte loads per lane per row
__global__ void __launch_bounds__(GV_THREADS) gemv_topk_kernel(const uint4* __restrict__ passages, int64_t n_rows,
                                                               int dim, const __nv_bfloat16* __restrict__ queries,
                                                               int k, int capacity, int64_t id_base,
                                                               int64_t row_first, int64_t rows_per_block,
                                                               const float* __restrict__ seed_thr,
                                                               uint64_t* __restrict__ part_keys) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_q = reinterpret_cast<float*>(smem_raw);                                 // [NQ][dim]
  uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem_raw + sizeof(float) * NQ * dim);  // [NQ][capacity]
  __shared__ int s_count[NQ];
  __shared__ uint64_t s_thr[NQ];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < NQ * dim; i += GV_THREADS) s_q[i] = __bfloat162float(queries[i]);
  BlockTopK<GV_THREADS> tk[NQ];
#pragma unroll
  for (int qi = 0; qi < NQ; ++qi) {
    // a proven lower bound of the query's k-th best score (the k-th best of the sampled prefix) as the floor of the
    // selection: nothing below it is ever appended, so after the prefix a block appends a handful of candidates and
    // never sorts before the end (an equal score still passes: the id bits of a key are never all zero)
    uint64_t floor_key = 0ull;
    if (seed_thr != nullptr) {
      const float seed = __ldg(seed_thr + qi);
      if (seed > -INFINITY) floor_key = static_cast<uint64_t>(float_to_ordered(seed)) << 32;
    }
    tk[qi].init(s_keys + qi * capacity, &s_count[qi], &s_thr[qi], k, capacity, floor_key);
  }

  const int vec_per_row = dim >> 3;
  const int64_t begin = row_first + static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t end = min(n_rows, begin + rows_per_block);

  for (int64_t tile = begin; tile < end; tile += GV_TILE) {
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) tk[qi].reserve(GV_TILE);
    uint64_t thr[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) thr[qi] = s_thr[qi];

    bool appended = false;   // this thread appended to SOME query's list in this tile (reserve() must hear about it)
    for (int r0 = warp * GV_ROWS_PER_WARP; r0 < GV_TILE; r0 += GV_WARPS * GV_ROWS_PER_WARP) {
      const int64_t row0 = tile + r0;
      if (row0 >= end) break;
      uint4 v[GV_ROWS_PER_WARP][CHUNKS];
#pragma unroll
      for (int r = 0; r < GV_ROWS_PER_WARP; ++r) {
        const int64_t row = row0 + r;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          const int col = c * 32 + lane;
          if (row < end && col < vec_per_row)
            v[r][c] = ldg_stream_u4(passages + row * vec_per_row + col);
          else
            v[r][c] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      float acc[GV_ROWS_PER_WARP][NQ];
#pragma unroll
      for (int r = 0; r < GV_ROWS_PER_WARP; ++r)
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) acc[r][qi] = 0.0f;
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c) {
        const int col = c * 32 + lane;
        if (col < vec_per_row) {
          float e[GV_ROWS_PER_WARP][8];
#pragma unroll
          for (int r = 0; r < GV_ROWS_PER_WARP; ++r) unpack_bf16x8(v[r][c], e[r]);
#pragma unroll
          for (int qi = 0; qi < NQ; ++qi) {
            const float4 qa = *reinterpret_cast<const float4*>(s_q + qi * dim + col * 8);
            const float4 qb = *reinterpret_cast<const float4*>(s_q + qi * dim + col * 8 + 4);
#pragma unroll
            for (int r = 0; r < GV_ROWS_PER_WARP; ++r) {
              float s = acc[r][qi];
              s = fmaf(e[r][0], qa.x, s);
              s = fmaf(e[r][1], qa.y, s);
              s = fmaf(e[r][2], qa.z, s);
              s = fmaf(e[r][3], qa.w, s);
              s = fmaf(e[r][4], qb.x, s);
              s = fmaf(e[r][5], qb.y, s);
              s = fmaf(e[r][6], qb.z, s);
              s = fmaf(e[r][7], qb.w, s);
              acc[r][qi] = s;
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < GV_ROWS_PER_WARP; ++r)
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
          float s = acc[r][qi];
#pragma unroll
          for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
          acc[r][qi] = s;
        }
      // lane (r * NQ + qi) publishes one (row, query) score
      const int slot_r = lane / NQ, slot_q = lane % NQ;
      if (lane < GV_ROWS_PER_WARP * NQ) {
        float s = 0.0f;
#pragma unroll
        for (int r = 0; r < GV_ROWS_PER_WARP; ++r)
#pragma unroll
          for (int qi = 0; qi < NQ; ++qi)
            if (r == slot_r && qi == slot_q) s = acc[r][qi];
        const int64_t row = row0 + slot_r;
        if (row < end) {
          uint64_t t = 0;
#pragma unroll
          for (int qi = 0; qi < NQ; ++qi)
            if (qi == slot_q) t = thr[qi];
          const uint64_t key = make_key(s, static_cast<int32_t>(id_base + row));
          if (key > t) {
            const int slot = atomicAdd(&s_count[slot_q], 1);
            s_keys[slot_q * capacity + k + slot] = key;
            appended = true;
          }
        }
      }
    }
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) tk[qi].dirty |= appended;
  }
#pragma unroll
  for (int qi = 0; qi < NQ; ++qi) tk[qi].finish();
  for (int i = tid; i < NQ * k; i += GV_THREADS) {
    const int qi = i / k, j = i % k;
    part_keys[(static_cast<int64_t>(qi) * gridDim.x + blockIdx.x) * k + j] = s_keys[qi * capacity + j];
  }
}

// Plain score matrix for small shapes: one warp per (row), all queries of a block-column.
__global__ void __launch_bounds__(256) dense_scores_kernel(const uint4* __restrict__ passages, int64_t n_rows, int dim,
                                                           const __nv_bfloat16* __restrict__ queries, int n_queries,
                                                           float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int vec_per_row = dim >> 3;
  for (int q = blockIdx.y; q < n_queries; q += gridDim.y) {
    const uint4* qv = reinterpret_cast<const uint4*>(queries + static_cast<int64_t>(q) * dim);
    float s = 0.0f;
    for (int col = lane; col < vec_per_row; col += 32) {
      float e[8], f[8];
      unpack_bf16x8(__ldg(passages + row * vec_per_row + col), e);
      unpack_bf16x8(__ldg(qv + col), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(e[i], f[i], s);
    }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
    if (lane == 0) out[static_cast<int64_t>(q) * n_rows + row] = s;
  }
}

template <int NQ>
static int launch_gemv(const void* passages, int64_t row_first, int64_t n_rows, int dim, const void* queries, int k,
                       int64_t id_base, const float* seed_thr, uint64_t* part, int grid, cudaStream_t stream) {
  // scores rows [row_first, n_rows) with `grid` blocks
  const int capacity = topk_capacity(k);
  const size_t smem = sizeof(float) * NQ * dim + sizeof(uint64_t) * NQ * capacity;
  const int64_t rows_per_block = ceil_div64(n_rows - row_first, grid);
  const int chunks = (dim + 255) / 256;
#define RAGB_GEMV_CASE(C)                                                                                          \
  case C: {                                                                                                        \
    RAGB_CUDA(cudaFuncSetAttribute(gemv_topk_kernel<NQ, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                   static_cast<int>(smem)));                                                       \
    gemv_topk_kernel<NQ, C><<<grid, GV_THREADS, smem, stream>>>(static_cast<const uint4*>(passages), n_rows, dim,  \
                                                                 static_cast<const __nv_bfloat16*>(queries), k,    \
                                                                 capacity, id_base, row_first, rows_per_block,     \
                                                                 seed_thr, part);                                  \
    break;                                                                                                         \
  }
  switch (chunks) {
    RAGB_GEMV_CASE(1)
    RAGB_GEMV_CASE(2)
    RAGB_GEMV_CASE(3)
    RAGB_GEMV_CASE(4)
    RAGB_GEMV_CASE(8)
    default:
      set_error("ragb_dense_gemv_topk: dim=%d not supported (dim/256 rounded up must be 1,2,3,4 or 8)", dim);
      return RAGB_ELIMIT;
  }
#undef RAGB_GEMV_CASE
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

// Resident blocks per SM of the instantiation launch_gemv<NQ> picks for this shape (0 = unsupported shape).
template <int NQ>
static int gemv_blocks_per_sm(int dim, int k) {
  const size_t smem = sizeof(float) * NQ * dim + sizeof(uint64_t) * NQ * topk_capacity(k);
  int n = 0;
#define RAGB_GEMV_OCC(C)                                                                                               \
  case C:                                                                                                              \
    if (cudaFuncSetAttribute(gemv_topk_kernel<NQ, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,                     \
                             static_cast<int>(smem)) != cudaSuccess ||                                                 \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gemv_topk_kernel<NQ, C>, GV_THREADS, smem) != cudaSuccess)   \
      n = 0;                                                                                                           \
    break;
  switch ((dim + 255) / 256) {
    RAGB_GEMV_OCC(1)
    RAGB_GEMV_OCC(2)
    RAGB_GEMV_OCC(3)
    RAGB_GEMV_OCC(4)
    RAGB_GEMV_OCC(8)
    default: break;
  }
#undef RAGB_GEMV_OCC
  return n;
}

}  // namespace ragb

using namespace ragb;

extern "C" {

// workspace: [n_queries, <= 148 * 4, k] block lists, [n_queries, k] merged list of the sampled prefix, [n_queries] bounds
size_t ragb_dense_gemv_workspace_bytes(int32_t n_queries, int32_t k) {
  if (n_queries <= 0 || k <= 0) return 0;
  return static_cast<size_t>(n_queries) * (148 * 4 + 1 + MERGE_SPLIT_MAX) * k * sizeof(uint64_t) +
         static_cast<size_t>(n_queries) * sizeof(float);
}

// one pass over rows [row_first, row_last) for all queries (groups of 4 / 2 / 1: register budget); -> grid used
static int gemv_pass(const void* passages, int64_t row_first, int64_t row_last, int dim, const __nv_bfloat16* q, int n_queries,
                     int k, int64_t id_base, const float* seed_thr, uint64_t* part, int* grid_out, cudaStream_t stream) {
  // One wave of blocks, all resident: the rows are split evenly between the blocks, so a grid of SMs x 4 with only
  // 3 blocks per SM resident (76 registers at dim 768, batch 1) ran a quarter of the rows in a second wave at a
  // third of the occupancy (ncu at 1M rows: 0.30 ms, 20 of 64 warps active on average).  The grid is SMs x the
  // residency of the slowest-fitting instantiation this pass launches.
  int per_sm = 4;
  if (n_queries >= 4) per_sm = std::min(per_sm, gemv_blocks_per_sm<4>(dim, k));
  if ((n_queries & 3) >= 2) per_sm = std::min(per_sm, gemv_blocks_per_sm<2>(dim, k));
  if (n_queries & 1) per_sm = std::min(per_sm, gemv_blocks_per_sm<1>(dim, k));
  if (per_sm < 1) per_sm = 1;
  int grid = device_sm_count() * per_sm;
  if (grid > 148 * 4) grid = 148 * 4;
  const int64_t max_blocks = ceil_div64(row_last - row_first, GV_WARPS * GV_ROWS_PER_WARP);
  if (grid > max_blocks) grid = static_cast<int>(max_blocks);
  *grid_out = grid;
  int done = 0;
  int rc = RAGB_OK;
  while (done < n_queries && rc == RAGB_OK) {
    const int left = n_queries - done;
    uint64_t* dst = part + static_cast<int64_t>(done) * grid * k;
    const __nv_bfloat16* qg = q + static_cast<int64_t>(done) * dim;
    const float* sg = seed_thr != nullptr ? seed_thr + done : nullptr;
    if (left >= 4) {
      rc = launch_gemv<4>(passages, row_first, row_last, dim, qg, k, id_base, sg, dst, grid, stream);
      done += 4;
    } else if (left >= 2) {
      rc = launch_gemv<2>(passages, row_first, row_last, dim, qg, k, id_base, sg, dst, grid, stream);
      done += 2;
    } else {
      rc = launch_gemv<1>(passages, row_first, row_last, dim, qg, k, id_base, sg, dst, grid, stream);
      done += 1;
    }
  }
  return rc;
}

int ragb_dense_gemv_topk(const void* passages_bf16, int64_t n_rows, int32_t dim, const void* queries_bf16,
                         int32_t n_queries, int32_t k, int64_t id_base, float* out_score, int32_t* out_id,
                         void* workspace, size_t workspace_bytes, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(passages_bf16 && queries_bf16 && out_score && out_id && workspace, RAGB_EINVAL,
               "ragb_dense_gemv_topk: null pointer");
  RAGB_REQUIRE((reinterpret_cast<uintptr_t>(passages_bf16) & 15) == 0, RAGB_EINVAL,
               "ragb_dense_gemv_topk: passages must be 16-byte aligned");
  RAGB_REQUIRE(n_rows > 0 && dim > 0 && dim % 8 == 0 && dim <= GV_MAX_DIM, RAGB_EINVAL,
               "ragb_dense_gemv_topk: dim=%d must be a multiple of 8, at most %d", dim, GV_MAX_DIM);
  RAGB_REQUIRE(n_queries >= 1 && n_queries <= RAGB_GEMV_MAX_BATCH, RAGB_ELIMIT,
               "ragb_dense_gemv_topk: n_queries=%d outside [1,%d] (use ragb_dense_mma_topk)", n_queries,
               RAGB_GEMV_MAX_BATCH);
  RAGB_REQUIRE(k > 0 && k <= RAGB_MAX_TOPK, RAGB_ELIMIT, "ragb_dense_gemv_topk: k=%d outside [1,%d]", k, RAGB_MAX_TOPK);
  RAGB_REQUIRE(id_base >= 0 && id_base + n_rows < (1ll << 31), RAGB_ELIMIT, "ragb_dense_gemv_topk: ids must fit int32");
  RAGB_REQUIRE(workspace_bytes >= ragb_dense_gemv_workspace_bytes(n_queries, k), RAGB_ENOSPC,
               "ragb_dense_gemv_topk: workspace too small");
  uint64_t* part = static_cast<uint64_t*>(workspace);
  uint64_t* sample_keys = part + static_cast<size_t>(n_queries) * 148 * 4 * k;
  uint64_t* scratch = sample_keys + static_cast<size_t>(n_queries) * k;     // partial lists of the two-level merge
  float* thr = reinterpret_cast<float*>(scratch + static_cast<size_t>(n_queries) * MERGE_SPLIT_MAX * k);
  const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(queries_bf16);
  // Optional sampled prefix (as in ragb_dense_mma_topk): the k-th best score of the first 1/DIV of the rows is a proven
  // lower bound of the final k-th best and becomes the floor of every block's selection over the remaining rows, which
  // removes the warm-up sorts of the blocks.  OFF by default: measured at 1M rows / batch 1 the extra pass + merge cost
  // what the warm-up costs (0.345 ms with DIV = 32 against 0.342 ms without).
  static const int sample_div = [] {
    const char* e = getenv("RAGB_GEMV_SAMPLE_DIV");   // tuning aid: 0 = no sampled prefix
    return e ? atoi(e) : 0;
  }();
  int64_t prefix = (sample_div > 0 && n_rows >= 262144) ? (n_rows / sample_div) / 32 * 32 : 0;
  if (prefix < 4 * static_cast<int64_t>(k)) prefix = 0;
  int grid = 0;
  int rc;
  const float* seed = nullptr;
  if (prefix > 0) {
    rc = gemv_pass(passages_bf16, 0, prefix, dim, q, n_queries, k, id_base, nullptr, part, &grid, stream);
    if (rc != RAGB_OK) return rc;
    rc = launch_merge_keys_split(part, n_queries, grid, k, nullptr, 0, k, nullptr, nullptr, sample_keys, thr, scratch, stream);
    if (rc != RAGB_OK) return rc;
    seed = thr;
  }
  rc = gemv_pass(passages_bf16, prefix, n_rows, dim, q, n_queries, k, id_base, seed, part, &grid, stream);
  if (rc != RAGB_OK) return rc;
  return launch_merge_keys_split(part, n_queries, grid, k, prefix > 0 ? sample_keys : nullptr, k, k, out_score, out_id, nullptr,
                                 nullptr, scratch, stream);
}

int ragb_dense_scores(const void* passages_bf16, int64_t n_rows, int32_t dim, const void* queries_bf16,
                      int32_t n_queries, float* out_scores, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(passages_bf16 && queries_bf16 && out_scores, RAGB_EINVAL, "ragb_dense_scores: null pointer");
  RAGB_REQUIRE(n_rows > 0 && n_queries > 0 && dim > 0 && dim % 8 == 0, RAGB_EINVAL, "ragb_dense_scores: bad shape");
  RAGB_REQUIRE(((reinterpret_cast<uintptr_t>(passages_bf16) | reinterpret_cast<uintptr_t>(queries_bf16)) & 15) == 0,
               RAGB_EINVAL, "ragb_dense_scores: inputs must be 16-byte aligned");
  dim3 grid(static_cast<unsigned>(ceil_div64(n_rows, 8)), n_queries < 64 ? n_queries : 64);
  dense_scores_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint4*>(passages_bf16), n_rows, dim,
                                                static_cast<const __nv_bfloat16*>(queries_bf16), n_queries, out_scores);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

}  // extern "C"
