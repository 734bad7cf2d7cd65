// Stand-alone selection kernels: row-wise top-k, candidate-list merge, pool fusion.
#include "common.cuh"
#include "topk.cuh"

namespace ragb {

constexpr int SEL_THREADS = 256;
constexpr int SEL_CHUNK = 512;  // candidates offered between two reserve() calls

struct SelSmem {
  int count;
  uint64_t threshold;
};

// ---------------------------------------------------------------------------------------
// Row-wise top-k of a dense fp32 matrix.  grid = (n_rows, n_split); block s of row r scans
// columns [s*span, (s+1)*span) and emits its own k best as keys; n_split == 1 writes the
// final (score, index) directly.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SEL_THREADS) topk_rows_kernel(const float* __restrict__ scores, int64_t n_cols,
                                                                int64_t span, int k, int capacity,
                                                                uint64_t* __restrict__ part_keys,
                                                                float* __restrict__ out_score,
                                                                int32_t* __restrict__ out_index) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
  __shared__ SelSmem st;
  BlockTopK<SEL_THREADS> tk;
  tk.init(keys, &st.count, &st.threshold, k, capacity, 0ull);

  const int row = blockIdx.x;
  const int64_t begin = static_cast<int64_t>(blockIdx.y) * span;
  const int64_t end = min(begin + span, n_cols);
  const float* src = scores + static_cast<int64_t>(row) * n_cols;
  for (int64_t base = begin; base < end; base += SEL_CHUNK) {
    tk.reserve(SEL_CHUNK);
    const uint64_t thr = *tk.threshold;
#pragma unroll
    for (int j = 0; j < SEL_CHUNK / SEL_THREADS; ++j) {
      int64_t idx = base + j * SEL_THREADS + threadIdx.x;
      if (idx < end) tk.offer(make_key(__ldg(src + idx), static_cast<int32_t>(idx)), thr);
    }
  }
  tk.finish();
  if (part_keys != nullptr) {
    uint64_t* dst = part_keys + (static_cast<int64_t>(row) * gridDim.y + blockIdx.y) * k;
    for (int i = threadIdx.x; i < k; i += SEL_THREADS) dst[i] = keys[i];
  } else {
    for (int i = threadIdx.x; i < k; i += SEL_THREADS) {
      uint64_t key = keys[i];
      out_score[static_cast<int64_t>(row) * k + i] = key ? key_score(key) : 0.0f;
      out_index[static_cast<int64_t>(row) * k + i] = key ? key_id(key) : -1;
    }
  }
}

// Full descending sort of short rows (n_cols <= 8192): serves top-k requests with k beyond
// RAGB_MAX_TOPK, e.g. hybrid_rerank(top_k >= P) which the reference clamps to P (router.py:202).
__global__ void __launch_bounds__(SEL_THREADS) sort_rows_kernel(const float* __restrict__ scores, int n_cols, int k,
                                                                int sort_n, float* __restrict__ out_score,
                                                                int32_t* __restrict__ out_index) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
  const int row = blockIdx.x;
  const float* src = scores + static_cast<int64_t>(row) * n_cols;
  for (int i = threadIdx.x; i < sort_n; i += SEL_THREADS) keys[i] = i < n_cols ? make_key(src[i], i) : 0ull;
  __syncthreads();
  bitonic_sort_desc<SEL_THREADS>(keys, sort_n);
  for (int i = threadIdx.x; i < k; i += SEL_THREADS) {
    uint64_t key = i < sort_n ? keys[i] : 0ull;
    out_score[static_cast<int64_t>(row) * k + i] = key ? key_score(key) : 0.0f;
    out_index[static_cast<int64_t>(row) * k + i] = key ? key_id(key) : -1;
  }
}

// ---------------------------------------------------------------------------------------
// Merge candidate lists.  Input either packed keys [n_queries, n_lists, k_in] or separate
// (score, id) arrays with id -1 = empty.  One block per query.
// ---------------------------------------------------------------------------------------
// Optional extras (all nullable): ``extra_keys`` [n_queries, k_extra] one more list per query stored elsewhere (the
// sampled prefix of the seeded dense search); ``out_keys`` [n_queries, k_out] the merged list as packed keys instead
// of (score, id); ``out_thr`` [n_queries] the score of the k_out-th best candidate, -inf while fewer than k_out exist
// (a proven lower bound of the final k-th best score whenever the input is a subset of the candidates).
__global__ void __launch_bounds__(SEL_THREADS) topk_merge_kernel(const uint64_t* __restrict__ in_keys,
                                                                 const float* __restrict__ in_score,
                                                                 const int32_t* __restrict__ in_id, int n_lists,
                                                                 int k_in, int k_out, int capacity,
                                                                 float* __restrict__ out_score,
                                                                 int32_t* __restrict__ out_id,
                                                                 const uint64_t* __restrict__ extra_keys, int k_extra,
                                                                 uint64_t* __restrict__ out_keys,
                                                                 float* __restrict__ out_thr, int64_t query_stride,
                                                                 int64_t list_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
  __shared__ SelSmem st;
  BlockTopK<SEL_THREADS> tk;
  tk.init(keys, &st.count, &st.threshold, k_out, capacity, 0ull);

  const int q = blockIdx.x;
  const int64_t total = static_cast<int64_t>(n_lists) * k_in;
  const int64_t off = static_cast<int64_t>(q) * total;
  for (int64_t base = 0; base < total; base += SEL_CHUNK) {
    tk.reserve(SEL_CHUNK);
    const uint64_t thr = *tk.threshold;
#pragma unroll
    for (int j = 0; j < SEL_CHUNK / SEL_THREADS; ++j) {
      int64_t idx = base + j * SEL_THREADS + threadIdx.x;
      if (idx < total) {
        uint64_t key;
        if (in_keys != nullptr) {
          key = in_keys[off + idx];
        } else {
          // (score, id) arrays: contiguous [n_queries, n_lists, k_in] when the strides are 0, else element
          // (q, list, j) lives at q * query_stride + list * list_stride + j (the all-gathered exchange buffer)
          const int64_t at = list_stride == 0 ? off + idx
                                              : static_cast<int64_t>(q) * query_stride + (idx / k_in) * list_stride + idx % k_in;
          int32_t id = in_id[at];
          key = id >= 0 ? make_key(in_score[at], id) : 0ull;
        }
        tk.offer(key, thr);  // key 0 never beats a threshold
      }
    }
  }
  if (extra_keys != nullptr) {
    for (int base = 0; base < k_extra; base += SEL_CHUNK) {
      tk.reserve(SEL_CHUNK);
      const uint64_t thr = *tk.threshold;
      for (int idx = base + threadIdx.x; idx < min(k_extra, base + SEL_CHUNK); idx += SEL_THREADS)
        tk.offer(extra_keys[static_cast<int64_t>(q) * k_extra + idx], thr);
    }
  }
  tk.finish();
  for (int i = threadIdx.x; i < k_out; i += SEL_THREADS) {
    uint64_t key = keys[i];
    if (out_keys != nullptr) out_keys[static_cast<int64_t>(q) * k_out + i] = key;
    if (out_score != nullptr) {
      out_score[static_cast<int64_t>(q) * k_out + i] = key ? key_score(key) : 0.0f;
      out_id[static_cast<int64_t>(q) * k_out + i] = key ? key_id(key) : -1;
    }
  }
  if (out_thr != nullptr && threadIdx.x == 0) out_thr[q] = keys[k_out - 1] ? key_score(keys[k_out - 1]) : -INFINITY;
}

// First level of the two-level merge: block (q, s) merges lists [s * per, (s + 1) * per) of query q into
// out_keys[q, s, 0..k_out) (zero = empty).
__global__ void __launch_bounds__(SEL_THREADS) topk_merge_slices_kernel(const uint64_t* __restrict__ in_keys, int n_lists, int per,
                                                                        int k_in, int k_out, int capacity,
                                                                        uint64_t* __restrict__ out_keys) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
  __shared__ SelSmem st;
  BlockTopK<SEL_THREADS> tk;
  tk.init(keys, &st.count, &st.threshold, k_out, capacity, 0ull);
  const int q = blockIdx.x, s = blockIdx.y;
  const int l0 = min(n_lists, s * per), l1 = min(n_lists, l0 + per);
  const int64_t total = static_cast<int64_t>(l1 - l0) * k_in;
  const uint64_t* src = in_keys + (static_cast<int64_t>(q) * n_lists + l0) * k_in;
  for (int64_t base = 0; base < total; base += SEL_CHUNK) {
    tk.reserve(SEL_CHUNK);
    const uint64_t thr = *tk.threshold;
#pragma unroll
    for (int j = 0; j < SEL_CHUNK / SEL_THREADS; ++j) {
      const int64_t idx = base + j * SEL_THREADS + threadIdx.x;
      if (idx < total) tk.offer(src[idx], thr);
    }
  }
  tk.finish();
  uint64_t* dst = out_keys + (static_cast<int64_t>(q) * gridDim.y + s) * k_out;
  for (int i = threadIdx.x; i < k_out; i += SEL_THREADS) dst[i] = keys[i];
}

// ---------------------------------------------------------------------------------------
// Pool fusion, HybridRetriever.hybrid_search (rag_uq/streaming_index.py:484-523).
// One block per query, pool <= 256 so at most 512 union rows.
// ---------------------------------------------------------------------------------------
constexpr int FUSE_THREADS = 128;
constexpr int FUSE_MAX_POOL = 256;

__global__ void __launch_bounds__(FUSE_THREADS) hybrid_fuse_kernel(
    const float* __restrict__ bm25_score, const int32_t* __restrict__ bm25_id, const float* __restrict__ dense_score,
    const int32_t* __restrict__ dense_id, int pool, int k, int32_t* __restrict__ out_id, float* __restrict__ out_bm25,
    float* __restrict__ out_dense, float* __restrict__ out_hybrid) {
  __shared__ int32_t row_id[2 * FUSE_MAX_POOL];
  __shared__ float row_b[2 * FUSE_MAX_POOL];
  __shared__ float row_d[2 * FUSE_MAX_POOL];
  __shared__ uint64_t keys[2 * FUSE_MAX_POOL];
  __shared__ float red_b[FUSE_THREADS / 32], red_d[FUSE_THREADS / 32];
  __shared__ int red_any[FUSE_THREADS / 32];

  const int q = blockIdx.x;
  const int tid = threadIdx.x;
  const int64_t off = static_cast<int64_t>(q) * pool;
  const int rows = 2 * pool;
  int sort_n = 2;
  while (sort_n < rows) sort_n <<= 1;

  // rows [0, pool): the BM25 pool (dense = 0.0 until matched, :498-499)
  for (int i = tid; i < pool; i += FUSE_THREADS) {
    row_id[i] = bm25_id[off + i];
    row_b[i] = bm25_score[off + i];
    row_d[i] = 0.0f;
  }
  __syncthreads();
  // rows [pool, 2*pool): dense-only entries; dense entries also present in the BM25 pool
  // deposit their score on that row instead (dict union, :489)
  for (int j = tid; j < pool; j += FUSE_THREADS) {
    int32_t id = dense_id[off + j];
    float sd = dense_score[off + j];
    int hit = -1;
    if (id >= 0) {
      for (int i = 0; i < pool; ++i)
        if (row_id[i] == id) {
          hit = i;
          break;
        }
    }
    if (hit >= 0) {
      row_d[hit] = sd;
      row_id[pool + j] = -1;
    } else {
      row_id[pool + j] = id;
      row_d[pool + j] = sd;
    }
    row_b[pool + j] = 0.0f;
  }
  __syncthreads();

  // max over the union (:513-514): "max(...) or 1"
  float mb = -INFINITY, md = -INFINITY;
  int any = 0;
  for (int i = tid; i < rows; i += FUSE_THREADS) {
    if (row_id[i] >= 0) {
      mb = fmaxf(mb, row_b[i]);
      md = fmaxf(md, row_d[i]);
      any = 1;
    }
  }
  for (int s = 16; s > 0; s >>= 1) {
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, s));
    md = fmaxf(md, __shfl_xor_sync(0xffffffffu, md, s));
    any |= __shfl_xor_sync(0xffffffffu, any, s);
  }
  if ((tid & 31) == 0) {
    red_b[tid >> 5] = mb;
    red_d[tid >> 5] = md;
    red_any[tid >> 5] = any;
  }
  __syncthreads();
  mb = red_b[0];
  md = red_d[0];
  any = red_any[0];
  for (int w = 1; w < FUSE_THREADS / 32; ++w) {
    mb = fmaxf(mb, red_b[w]);
    md = fmaxf(md, red_d[w]);
    any |= red_any[w];
  }
  if (mb == 0.0f) mb = 1.0f;
  if (md == 0.0f) md = 1.0f;

  // hybrid = (b / max_b + d / max_d) / 2 (:517-519), sort descending (:521)
  for (int i = tid; i < sort_n; i += FUSE_THREADS) {
    uint64_t key = 0ull;
    if (i < rows && row_id[i] >= 0) {
      float h = (row_b[i] / mb + row_d[i] / md) * 0.5f;
      key = make_key(h, row_id[i]);
    }
    keys[i] = key;
  }
  __syncthreads();
  bitonic_sort_desc<FUSE_THREADS>(keys, sort_n);

  for (int j = tid; j < k; j += FUSE_THREADS) {
    uint64_t key = (any && j < sort_n) ? keys[j] : 0ull;
    int32_t id = -1;
    float sb = 0.0f, sd = 0.0f, sh = 0.0f;
    if (key) {
      id = key_id(key);
      sh = key_score(key);
      for (int i = 0; i < rows; ++i)
        if (row_id[i] == id) {
          sb = row_b[i];
          sd = row_d[i];
          break;
        }
    }
    const int64_t o = static_cast<int64_t>(q) * k + j;
    out_id[o] = id;
    out_bm25[o] = sb;
    out_dense[o] = sd;
    out_hybrid[o] = sh;
  }
}

// ---------------------------------------------------------------------------------------
// Retrieval uncertainty of a ranked list, docs/uncertainty_theory.md:48-56 of the reference:
//   U = std(s_top-k) + lambda * (1 - |s_1 - s_k|)      (population std, like numpy's default)
// over the valid entries (id >= 0) of each row.  One warp per query.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) retrieval_uncertainty_kernel(const float* __restrict__ score,
                                                                    const int32_t* __restrict__ id, int n_queries,
                                                                    int k, float lambda, float* __restrict__ out) {
  const int q = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (q >= n_queries) return;
  float sum = 0.0f, hi = -INFINITY, lo = INFINITY;
  int n = 0;
  for (int j = lane; j < k; j += 32) {
    if (id[static_cast<int64_t>(q) * k + j] >= 0) {
      const float v = score[static_cast<int64_t>(q) * k + j];
      sum += v;
      hi = fmaxf(hi, v);
      lo = fminf(lo, v);
      ++n;
    }
  }
  for (int s = 16; s > 0; s >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, s);
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, s));
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, s));
    n += __shfl_xor_sync(0xffffffffu, n, s);
  }
  const float mean = n > 0 ? sum / n : 0.0f;
  float ss = 0.0f;
  for (int j = lane; j < k; j += 32)
    if (id[static_cast<int64_t>(q) * k + j] >= 0) {
      const float d = score[static_cast<int64_t>(q) * k + j] - mean;
      ss = fmaf(d, d, ss);
    }
  for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
  if (lane == 0) out[q] = n > 0 ? sqrtf(ss / n) + lambda * (1.0f - fabsf(hi - lo)) : lambda;
}

// Host-side helpers shared with the scoring kernels ---------------------------------------
int launch_merge_keys(const uint64_t* keys, int n_queries, int n_lists, int k_in, int k_out, float* out_score,
                      int32_t* out_id, cudaStream_t stream) {
  return launch_merge_keys_ex(keys, n_queries, n_lists, k_in, nullptr, 0, k_out, out_score, out_id, nullptr, nullptr, stream);
}

int launch_merge_keys_ex(const uint64_t* keys, int n_queries, int n_lists, int k_in, const uint64_t* extra_keys, int k_extra,
                         int k_out, float* out_score, int32_t* out_id, uint64_t* out_keys, float* out_thr,
                         cudaStream_t stream) {
  const int capacity = topk_capacity(k_out);
  topk_merge_kernel<<<n_queries, SEL_THREADS, capacity * sizeof(uint64_t), stream>>>(
      keys, nullptr, nullptr, n_lists, k_in, k_out, capacity, out_score, out_id, extra_keys, k_extra, out_keys, out_thr, 0, 0);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

// Two-level merge for FEW queries with MANY lists (the GEMV path: 592 block lists for 1..8 queries).  One block per
// query would walk n_lists * k_in keys in 512-key rounds with a barrier each (~35 us for 30k keys); here `split` blocks
// per query merge a slice of the lists each into scratch[n_queries, split, k_out] and a second, tiny launch merges
// those (plus the optional extra list).  scratch: n_queries * MERGE_SPLIT_MAX * k_out keys.
int launch_merge_keys_split(const uint64_t* keys, int n_queries, int n_lists, int k_in, const uint64_t* extra_keys, int k_extra,
                            int k_out, float* out_score, int32_t* out_id, uint64_t* out_keys, float* out_thr, uint64_t* scratch,
                            cudaStream_t stream) {
  int split = n_lists / 24;                      // ~24 lists (1200 keys at k = 50) per first-level block
  if (split > MERGE_SPLIT_MAX) split = MERGE_SPLIT_MAX;
  if (split < 2 || k_in > k_out || static_cast<int64_t>(n_queries) * split > 148 * 8)
    return launch_merge_keys_ex(keys, n_queries, n_lists, k_in, extra_keys, k_extra, k_out, out_score, out_id, out_keys, out_thr,
                                stream);
  const int capacity = topk_capacity(k_out);
  const int per = (n_lists + split - 1) / split;
  topk_merge_slices_kernel<<<dim3(n_queries, split), SEL_THREADS, capacity * sizeof(uint64_t), stream>>>(
      keys, n_lists, per, k_in, k_out, capacity, scratch);
  RAGB_AFTER_LAUNCH(1);
  return launch_merge_keys_ex(scratch, n_queries, split, k_out, extra_keys, k_extra, k_out, out_score, out_id, out_keys, out_thr,
                              stream);
}

static int rows_split(int n_rows, int64_t n_cols) {
  // enough blocks to fill the machine (148 SMs x 8 resident blocks), each at least 4096 wide
  int64_t want = ceil_div64(148 * 8, n_rows);
  int64_t max_split = ceil_div64(n_cols, 4096);
  int64_t s = want < max_split ? want : max_split;
  return s < 1 ? 1 : static_cast<int>(s);
}

}  // namespace ragb

using namespace ragb;

extern "C" {

size_t ragb_topk_rows_workspace_bytes(int32_t n_rows, int64_t n_cols, int32_t k) {
  if (n_rows <= 0 || n_cols <= 0 || k <= 0 || k > RAGB_MAX_TOPK) return 0;
  return static_cast<size_t>(n_rows) * rows_split(n_rows, n_cols) * k * sizeof(uint64_t);
}

int ragb_topk_rows(const float* scores, int32_t n_rows, int64_t n_cols, int32_t k, float* out_score,
                   int32_t* out_index, void* workspace, size_t workspace_bytes, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(scores && out_score && out_index, RAGB_EINVAL, "ragb_topk_rows: null pointer");
  RAGB_REQUIRE(n_rows > 0 && n_cols > 0 && n_cols < (1ll << 31), RAGB_EINVAL, "ragb_topk_rows: bad shape");
  RAGB_REQUIRE(k > 0, RAGB_EINVAL, "ragb_topk_rows: k must be positive");
  if (k > RAGB_MAX_TOPK) {
    RAGB_REQUIRE(n_cols <= 8192, RAGB_ELIMIT, "ragb_topk_rows: k=%d > %d is only served for rows of at most 8192 columns",
                 k, RAGB_MAX_TOPK);
    int sort_n = 2;
    while (sort_n < n_cols) sort_n <<= 1;
    const size_t smem = sort_n * sizeof(uint64_t);
    RAGB_CUDA(cudaFuncSetAttribute(sort_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    sort_rows_kernel<<<n_rows, SEL_THREADS, smem, stream>>>(scores, static_cast<int>(n_cols), k, sort_n, out_score,
                                                            out_index);
    RAGB_AFTER_LAUNCH(1);
    return RAGB_OK;
  }
  const int split = rows_split(n_rows, n_cols);
  const int capacity = topk_capacity(k);
  const int64_t span = ceil_div64(n_cols, split);
  if (split == 1) {
    topk_rows_kernel<<<dim3(n_rows, 1), SEL_THREADS, capacity * sizeof(uint64_t), stream>>>(
        scores, n_cols, span, k, capacity, nullptr, out_score, out_index);
    RAGB_AFTER_LAUNCH(1);
    return RAGB_OK;
  }
  RAGB_REQUIRE(workspace && workspace_bytes >= ragb_topk_rows_workspace_bytes(n_rows, n_cols, k), RAGB_ENOSPC,
               "ragb_topk_rows: workspace too small");
  uint64_t* part = static_cast<uint64_t*>(workspace);
  topk_rows_kernel<<<dim3(n_rows, split), SEL_THREADS, capacity * sizeof(uint64_t), stream>>>(
      scores, n_cols, span, k, capacity, part, nullptr, nullptr);
  RAGB_AFTER_LAUNCH(1);
  return launch_merge_keys(part, n_rows, split, k, k, out_score, out_index, stream);
}

int ragb_topk_merge(const float* in_score, const int32_t* in_id, int32_t n_queries, int32_t n_lists, int32_t k_in,
                    int32_t k_out, float* out_score, int32_t* out_id, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(in_score && in_id && out_score && out_id, RAGB_EINVAL, "ragb_topk_merge: null pointer");
  RAGB_REQUIRE(n_queries > 0 && n_lists > 0 && k_in > 0, RAGB_EINVAL, "ragb_topk_merge: bad shape");
  RAGB_REQUIRE(k_out > 0 && k_out <= RAGB_MAX_TOPK, RAGB_ELIMIT, "ragb_topk_merge: k_out=%d outside [1,%d]", k_out,
               RAGB_MAX_TOPK);
  const int capacity = topk_capacity(k_out);
  topk_merge_kernel<<<n_queries, SEL_THREADS, capacity * sizeof(uint64_t), stream>>>(
      nullptr, in_score, in_id, n_lists, k_in, k_out, capacity, out_score, out_id, nullptr, 0, nullptr, nullptr, 0, 0);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

int ragb_topk_merge_strided(const float* in_score, const int32_t* in_id, int32_t n_queries, int32_t n_lists, int32_t k_in,
                            int64_t query_stride, int64_t list_stride, int32_t k_out, float* out_score, int32_t* out_id,
                            ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(in_score && in_id && out_score && out_id, RAGB_EINVAL, "ragb_topk_merge_strided: null pointer");
  RAGB_REQUIRE(n_queries > 0 && n_lists > 0 && k_in > 0 && query_stride > 0 && list_stride > 0, RAGB_EINVAL,
               "ragb_topk_merge_strided: bad shape");
  RAGB_REQUIRE(k_out > 0 && k_out <= RAGB_MAX_TOPK, RAGB_ELIMIT, "ragb_topk_merge_strided: k_out=%d outside [1,%d]", k_out,
               RAGB_MAX_TOPK);
  const int capacity = topk_capacity(k_out);
  topk_merge_kernel<<<n_queries, SEL_THREADS, capacity * sizeof(uint64_t), stream>>>(
      nullptr, in_score, in_id, n_lists, k_in, k_out, capacity, out_score, out_id, nullptr, 0, nullptr, nullptr, query_stride,
      list_stride);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

int ragb_retrieval_uncertainty(const float* score, const int32_t* id, int32_t n_queries, int32_t k, double lambda,
                               float* out, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(score && id && out, RAGB_EINVAL, "ragb_retrieval_uncertainty: null pointer");
  RAGB_REQUIRE(n_queries > 0 && k > 0, RAGB_EINVAL, "ragb_retrieval_uncertainty: empty shape");
  retrieval_uncertainty_kernel<<<(n_queries + 7) / 8, 256, 0, stream>>>(score, id, n_queries, k,
                                                                        static_cast<float>(lambda), out);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

int ragb_hybrid_fuse_topk(const float* bm25_score, const int32_t* bm25_id, const float* dense_score,
                          const int32_t* dense_id, int32_t n_queries, int32_t pool, int32_t k, int32_t* out_id,
                          float* out_bm25, float* out_dense, float* out_hybrid, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(bm25_score && bm25_id && dense_score && dense_id && out_id && out_bm25 && out_dense && out_hybrid,
               RAGB_EINVAL, "ragb_hybrid_fuse_topk: null pointer");
  RAGB_REQUIRE(n_queries > 0, RAGB_EINVAL, "ragb_hybrid_fuse_topk: n_queries must be positive");
  RAGB_REQUIRE(pool > 0 && pool <= FUSE_MAX_POOL, RAGB_ELIMIT, "ragb_hybrid_fuse_topk: pool=%d outside [1,%d]", pool,
               FUSE_MAX_POOL);
  RAGB_REQUIRE(k > 0 && k <= 2 * FUSE_MAX_POOL, RAGB_ELIMIT, "ragb_hybrid_fuse_topk: k=%d outside [1,%d]", k,
               2 * FUSE_MAX_POOL);
  hybrid_fuse_kernel<<<n_queries, FUSE_THREADS, 0, stream>>>(bm25_score, bm25_id, dense_score, dense_id, pool, k,
                                                             out_id, out_bm25, out_dense, out_hybrid);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

}  // extern "C"
