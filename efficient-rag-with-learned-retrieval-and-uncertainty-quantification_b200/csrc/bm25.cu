// Okapi BM25 over a term-major CSR inverted index (K1 in SURVEY.md).
//
// Replaces rank_bm25.BM25Okapi.get_scores + np.argsort as called from
// BM25Index.search (rag_uq/streaming_index.py:165-179) and the statistics built by
// BM25Okapi.__init__ (reached from streaming_index.py:142,220).
//
// Data layout in HBM (per shard):
//   term_off[V+1] int64, post_doc[nnz] int32 ascending per term, post_tf[nnz] uint16,
//   norm[N] float32 = k1*(1-b+b*len/avgdl), idf[V] float32.
// Algorithmic bytes per posting: 4 (doc) + 2 (tf) = 6, plus 4 per document per range for norm.
//
// Kernel shape.  grid = (queries, stripes); a block of 8 warps owns one stripe of
// consecutive documents for one query, each WARP owns a contiguous eighth of it and walks
// it in ranges of 256 documents.  Per range the warp keeps a 256-float accumulator in
// shared memory and streams, term after term, the postings that fall in the range:
// posting lists are sorted by document, so every warp simply continues reading where it
// stopped (one cursor per term), 128-byte coalesced, 1 to 8 independent 32-posting chunks in
// flight per pass depending on the list's density.  Within one 32-posting chunk all documents are distinct, so
// the accumulation is a plain shared-memory read-modify-write: no atomics anywhere.
// Terms whose next posting lies beyond the range are skipped without touching memory.
// The block-wide running top-k (topk.cuh) then filters the 8x256 scores: one barrier per
// range in steady state.
#include <climits>

#include "common.cuh"
#include "topk.cuh"

namespace ragb {

constexpr int BM_THREADS = 256;
constexpr int BM_WARPS = BM_THREADS / 32;
constexpr int BM_RANGE = 256;          // documents per warp range
constexpr int BM_MAX_TERMS = 64;
constexpr int BM_SEARCH = 4;  // posting lists searched concurrently while placing the cursors

struct Bm25Args {
  const int64_t* term_off;
  const int32_t* post_doc;
  const uint16_t* post_tf;
  const float* norm;
  const float* idf;
  const int32_t* q_terms;
  const int32_t* q_off;
  int64_t vocab;
  int64_t n_docs;
  int64_t id_base;
  int64_t stripe_docs;  // multiple of BM_WARPS * BM_RANGE
  float k1p1;
  int max_terms;
  int k;
  int capacity;
  uint64_t* part_keys;  // [queries, stripes, k]            (top-k mode)
  float* out_scores;    // [queries, n_docs]                (dense mode)
};

// Stream the postings of one term that fall below d1 into the warp's accumulator.
// U chunks of 32 postings are loaded per pass (all loads independent) plus one "peek" posting
// right behind them, so a pass that consumes everything it loaded still learns the next
// document without another round trip to memory.
template <int U>
__device__ __forceinline__ void stream_term(const int32_t* __restrict__ post_doc,
                                            const uint16_t* __restrict__ post_tf, int64_t& pos, const int64_t end,
                                            const int d0, const int d1, const float weight, float* accw,
                                            const float* nrmw, const int lane, int& next_doc) {
  while (true) {
    int doc[U];
    unsigned tf[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t idx = pos + u * 32 + lane;
      const bool in = idx < end;
      doc[u] = in ? __ldg(post_doc + idx) : INT_MAX;
      tf[u] = in ? static_cast<unsigned>(__ldg(post_tf + idx)) : 0u;
    }
    const int64_t peek_idx = pos + U * 32;
    int peek = INT_MAX;
    if (lane == 0 && peek_idx < end) peek = __ldg(post_doc + peek_idx);
    int taken = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool take = doc[u] < d1;
      if (take) {
        const float f = static_cast<float>(tf[u]);
        const int o = doc[u] - d0;
        accw[o] += weight * __fdividef(f, f + nrmw[o]);
      }
      taken += __popc(__ballot_sync(0xffffffffu, take));
    }
    pos += taken;
    if (taken < U * 32) {
      // sorted list: the taken postings are a prefix; the first one left is the next document
      int cand = INT_MAX;
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (u == (taken >> 5)) cand = doc[u];
      next_doc = __shfl_sync(0xffffffffu, cand, taken & 31);
      return;
    }
    peek = __shfl_sync(0xffffffffu, peek, 0);
    if (peek >= d1) {
      next_doc = peek;
      return;
    }
    __syncwarp();
  }
}

template <bool DENSE_OUT>
__global__ void __launch_bounds__(BM_THREADS, 4) bm25_kernel(const Bm25Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = a.max_terms;
  // carve shared memory
  unsigned char* sp = smem_raw;
  int64_t* s_pos = reinterpret_cast<int64_t*>(sp) + warp * mt;
  sp += sizeof(int64_t) * BM_WARPS * mt;
  int64_t* s_end = reinterpret_cast<int64_t*>(sp) + warp * mt;
  sp += sizeof(int64_t) * BM_WARPS * mt;
  uint64_t* s_keys = reinterpret_cast<uint64_t*>(sp);
  if (!DENSE_OUT) sp += sizeof(uint64_t) * a.capacity;
  float* accw = reinterpret_cast<float*>(sp) + warp * BM_RANGE;
  sp += sizeof(float) * BM_WARPS * BM_RANGE;
  float* nrmw = reinterpret_cast<float*>(sp) + warp * BM_RANGE;
  sp += sizeof(float) * BM_WARPS * BM_RANGE;
  float* s_wgt = reinterpret_cast<float*>(sp) + warp * mt;
  sp += sizeof(float) * BM_WARPS * mt;
  int* s_nxt = reinterpret_cast<int*>(sp) + warp * mt;
  sp += sizeof(int) * BM_WARPS * mt;
  unsigned char* s_dense = sp + warp * mt;

  __shared__ int s_count;
  __shared__ uint64_t s_threshold;
  __shared__ int s_pending[3];

  const int q = blockIdx.x;
  const int64_t stripe_begin = static_cast<int64_t>(blockIdx.y) * a.stripe_docs;
  const int64_t stripe_end = min(a.n_docs, stripe_begin + a.stripe_docs);
  const int64_t sub_docs = a.stripe_docs / BM_WARPS;
  const int64_t w_begin = min(stripe_end, stripe_begin + warp * sub_docs);
  const int64_t w_end = min(stripe_end, w_begin + sub_docs);
  const int n_iters = static_cast<int>(sub_docs / BM_RANGE);

  BlockTopK<BM_THREADS> tk;
  if (!DENSE_OUT) {
    if (tid < 3) s_pending[tid] = 0;
    tk.init(s_keys, &s_count, &s_threshold, a.k, a.capacity, positive_floor_key());
  }

  // ---- per-warp cursors: lower_bound(post_doc[term], w_begin) by a 32-ary search, 4 terms at a time
  const int qb = a.q_off[q];
  const int nt = min(a.q_off[q + 1] - qb, mt);
  int ntv = 0;  // valid terms kept (warp-uniform)
  for (int g = 0; g < nt; g += BM_SEARCH) {
    int64_t lo[BM_SEARCH], hi[BM_SEARCH], te[BM_SEARCH], ts[BM_SEARCH];
    float wg[BM_SEARCH];
#pragma unroll
    for (int j = 0; j < BM_SEARCH; ++j) {
      lo[j] = hi[j] = te[j] = ts[j] = 0;
      wg[j] = 0.0f;
      if (g + j < nt) {
        const int t = a.q_terms[qb + g + j];
        if (t >= 0 && t < a.vocab) {
          const float w = a.idf[t] * a.k1p1;
          if (w != 0.0f) {
            lo[j] = ts[j] = a.term_off[t];
            hi[j] = te[j] = a.term_off[t + 1];
            wg[j] = w;
          }
        }
      }
    }
    const int target = static_cast<int>(w_begin);
    bool more = true;
    while (more) {
      int probe[BM_SEARCH];
      int64_t chunk[BM_SEARCH];
#pragma unroll
      for (int j = 0; j < BM_SEARCH; ++j) {
        const int64_t len = hi[j] - lo[j];
        chunk[j] = 0;
        probe[j] = INT_MAX;
        if (len > 32) {
          chunk[j] = (len + 31) >> 5;
          int64_t idx = (lane + 1) * chunk[j] - 1;
          if (idx > len - 1) idx = len - 1;
          probe[j] = __ldg(a.post_doc + lo[j] + idx);
        }
      }
      more = false;
#pragma unroll
      for (int j = 0; j < BM_SEARCH; ++j) {
        if (chunk[j] > 0) {  // warp-uniform
          const int c = __popc(__ballot_sync(0xffffffffu, probe[j] < target));
          if (c == 32) {
            lo[j] = hi[j];
          } else {
            lo[j] += c * chunk[j];
            if (lo[j] + chunk[j] < hi[j]) hi[j] = lo[j] + chunk[j];
          }
          if (hi[j] - lo[j] > 32) more = true;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < BM_SEARCH; ++j) {
      if (wg[j] != 0.0f) {  // warp-uniform
        const int64_t idx = lo[j] + lane;
        const bool below = idx < hi[j] && __ldg(a.post_doc + idx) < target;
        const int64_t cur = lo[j] + __popc(__ballot_sync(0xffffffffu, below));
        int nd = INT_MAX;
        if (cur < te[j]) nd = __ldg(a.post_doc + cur);
        if (lane == 0) {
          s_pos[ntv] = cur;
          s_end[ntv] = te[j];
          s_wgt[ntv] = wg[j];
          s_nxt[ntv] = nd;
          // expected postings of this term per 256-document range
          const double per_range = static_cast<double>(te[j] - ts[j]) * BM_RANGE / static_cast<double>(a.n_docs);
          s_dense[ntv] = per_range < 24.0 ? 0 : (per_range < 56.0 ? 1 : (per_range < 120.0 ? 2 : 3));
        }
        ++ntv;
      }
    }
  }
  __syncwarp();

  int local_count = 0;  // replica of tk.count, identical in every thread
  for (int it = 0; it < n_iters; ++it) {
    const int64_t d0l = w_begin + static_cast<int64_t>(it) * BM_RANGE;
    const int d0 = static_cast<int>(d0l < w_end ? d0l : w_end);
    const int d1 = static_cast<int>(min(w_end, d0l + BM_RANGE));
    const int cnt = d1 > d0 ? d1 - d0 : 0;
#pragma unroll
    for (int j = lane; j < BM_RANGE; j += 32) {
      accw[j] = 0.0f;
      nrmw[j] = j < cnt ? __ldg(a.norm + d0 + j) : 1.0f;
    }
    __syncwarp();
    if (cnt > 0) {
      for (int ti = 0; ti < ntv; ++ti) {
        if (s_nxt[ti] >= d1) continue;  // nothing of this term in the range
        int64_t pos = s_pos[ti];
        const int64_t end = s_end[ti];
        const float w = s_wgt[ti];
        int next_doc;
        switch (s_dense[ti]) {  // chunks per pass sized to the term's density (warp-uniform)
          case 0: stream_term<1>(a.post_doc, a.post_tf, pos, end, d0, d1, w, accw, nrmw, lane, next_doc); break;
          case 1: stream_term<2>(a.post_doc, a.post_tf, pos, end, d0, d1, w, accw, nrmw, lane, next_doc); break;
          case 2: stream_term<4>(a.post_doc, a.post_tf, pos, end, d0, d1, w, accw, nrmw, lane, next_doc); break;
          default: stream_term<8>(a.post_doc, a.post_tf, pos, end, d0, d1, w, accw, nrmw, lane, next_doc); break;
        }
        if (lane == 0) {
          s_pos[ti] = pos;
          s_nxt[ti] = next_doc;
        }
        __syncwarp();
      }
    }
    if (DENSE_OUT) {
      float* dst = a.out_scores + static_cast<int64_t>(q) * a.n_docs + d0;
      for (int j = lane; j < cnt; j += 32) dst[j] = accw[j];
      __syncwarp();
    } else {
      // ---- block-wide selection: count what beats the threshold, one barrier, then append
      const uint64_t thr = s_threshold;
      const int32_t gid0 = static_cast<int32_t>(a.id_base + d0);
      unsigned mask = 0;
#pragma unroll
      for (int u = 0; u < BM_RANGE / 32; ++u) {
        const int j = u * 32 + lane;
        if (j < cnt && make_key(accw[j], gid0 + j) > thr) mask |= 1u << u;
      }
      int mine = __popc(mask);
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, s);
      if (lane == 0 && mine) atomicAdd(&s_pending[it % 3], mine);
      __syncthreads();
      const int pending = s_pending[it % 3];
      if (tid == 0) s_pending[(it + 2) % 3] = 0;
      if (local_count + pending <= tk.room()) {
#pragma unroll
        for (int u = 0; u < BM_RANGE / 32; ++u)
          if (mask & (1u << u)) {
            const int j = u * 32 + lane;
            const int slot = atomicAdd(&s_count, 1);
            s_keys[a.k + slot] = make_key(accw[j], gid0 + j);
          }
        local_count += pending;
      } else {
        // slow path (first ranges, adversarial data): 64 documents per warp at a time
        for (int ph = 0; ph < BM_RANGE / 64; ++ph) {
          tk.reserve(BM_WARPS * 64);
          const uint64_t t2 = s_threshold;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int j = ph * 64 + u * 32 + lane;
            if (j < cnt) tk.offer(make_key(accw[j], gid0 + j), t2);
          }
        }
        __syncthreads();
        local_count = s_count;
      }
    }
  }
  if (!DENSE_OUT) {
    tk.finish();
    uint64_t* dst = a.part_keys + (static_cast<int64_t>(q) * gridDim.y + blockIdx.y) * a.k;
    for (int i = tid; i < a.k; i += BM_THREADS) dst[i] = s_keys[i];
  }
}

// ---------------------------------------------------------------------------------------
// Statistics: idf with the epsilon floor (two deterministic passes), length norm.
// ---------------------------------------------------------------------------------------
constexpr int IDF_BLOCKS = 256;
constexpr int IDF_THREADS = 256;

__device__ __forceinline__ double raw_idf(int df, double n) { return log(n - df + 0.5) - log(df + 0.5); }

__global__ void __launch_bounds__(IDF_THREADS) idf_partial_kernel(const int32_t* __restrict__ df, int64_t vocab,
                                                                  double n, double* __restrict__ partial) {
  __shared__ double s_sum[IDF_THREADS];
  __shared__ double s_cnt[IDF_THREADS];
  double sum = 0.0, cnt = 0.0;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * IDF_THREADS + threadIdx.x; t < vocab;
       t += static_cast<int64_t>(IDF_BLOCKS) * IDF_THREADS) {
    const int d = df[t];
    if (d > 0) {
      sum += raw_idf(d, n);
      cnt += 1.0;
    }
  }
  s_sum[threadIdx.x] = sum;
  s_cnt[threadIdx.x] = cnt;
  __syncthreads();
  for (int s = IDF_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      s_sum[threadIdx.x] += s_sum[threadIdx.x + s];
      s_cnt[threadIdx.x] += s_cnt[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = s_sum[0];
    partial[2 * blockIdx.x + 1] = s_cnt[0];
  }
}

__global__ void __launch_bounds__(IDF_THREADS) idf_final_kernel(const int32_t* __restrict__ df, int64_t vocab, double n,
                                                                double epsilon, const double* __restrict__ partial,
                                                                float* __restrict__ idf_out) {
  __shared__ double s_sum[IDF_BLOCKS];
  __shared__ double s_cnt[IDF_BLOCKS];
  s_sum[threadIdx.x] = partial[2 * threadIdx.x];
  s_cnt[threadIdx.x] = partial[2 * threadIdx.x + 1];
  __syncthreads();
  for (int s = IDF_BLOCKS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      s_sum[threadIdx.x] += s_sum[threadIdx.x + s];
      s_cnt[threadIdx.x] += s_cnt[threadIdx.x + s];
    }
    __syncthreads();
  }
  const double average = s_cnt[0] > 0.0 ? s_sum[0] / s_cnt[0] : 0.0;
  const double floor_value = epsilon * average;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * IDF_THREADS + threadIdx.x; t < vocab;
       t += static_cast<int64_t>(gridDim.x) * IDF_THREADS) {
    const int d = df[t];
    double v = 0.0;
    if (d > 0) {
      v = raw_idf(d, n);
      if (v < 0.0) v = floor_value;
    }
    idf_out[t] = static_cast<float>(v);
  }
}

__global__ void norm_kernel(const int32_t* __restrict__ doc_len, int64_t n_docs, double avgdl, double k1, double b,
                            float* __restrict__ norm_out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n_docs) norm_out[i] = static_cast<float>(k1 * (1.0 - b + b * static_cast<double>(doc_len[i]) / avgdl));
}

static int bm25_stripes(int n_queries, int64_t n_docs, int64_t* stripe_docs_out) {
  const int64_t unit = static_cast<int64_t>(BM_WARPS) * BM_RANGE;
  const int64_t target_blocks = 148 * 6 * 2;
  int64_t stripes = ceil_div64(target_blocks, n_queries);
  const int64_t max_stripes = ceil_div64(n_docs, unit);
  if (stripes > max_stripes) stripes = max_stripes;
  if (stripes < 1) stripes = 1;
  int64_t stripe_docs = ceil_div64(ceil_div64(n_docs, stripes), unit) * unit;
  stripes = ceil_div64(n_docs, stripe_docs);
  *stripe_docs_out = stripe_docs;
  return static_cast<int>(stripes);
}

static size_t bm25_smem_bytes(int max_terms, int capacity, bool dense_out) {
  size_t b = 0;
  b += 2 * sizeof(int64_t) * BM_WARPS * max_terms;
  if (!dense_out) b += sizeof(uint64_t) * capacity;
  b += 2 * sizeof(float) * BM_WARPS * BM_RANGE;
  b += (sizeof(float) + sizeof(int) + 1) * BM_WARPS * max_terms;
  return (b + 15) & ~static_cast<size_t>(15);
}

static int bm25_common_checks(const char* who, const int64_t* term_off, const int32_t* post_doc,
                              const uint16_t* post_tf, const float* norm, const float* idf, int64_t vocab,
                              const int32_t* q_terms, const int32_t* q_off, int32_t n_queries, int64_t n_docs,
                              int32_t max_terms) {
  RAGB_REQUIRE(term_off && post_doc && post_tf && norm && idf && q_terms && q_off, RAGB_EINVAL, "%s: null pointer", who);
  RAGB_REQUIRE(vocab > 0 && n_queries > 0 && n_docs > 0, RAGB_EINVAL, "%s: empty shape", who);
  RAGB_REQUIRE(n_docs < (1ll << 31) - BM_RANGE, RAGB_ELIMIT, "%s: n_docs per shard must fit int32", who);
  RAGB_REQUIRE(max_terms >= 1 && max_terms <= BM_MAX_TERMS, RAGB_ELIMIT,
               "%s: max_query_terms=%d outside [1,%d]", who, max_terms, BM_MAX_TERMS);
  return RAGB_OK;
}

}  // namespace ragb

using namespace ragb;

extern "C" {

size_t ragb_bm25_idf_scratch_bytes(int64_t) { return 2 * IDF_BLOCKS * sizeof(double); }

int ragb_bm25_build_idf(const int32_t* df, int64_t vocab, int64_t corpus_size, double epsilon, float* idf_out,
                        void* scratch, size_t scratch_bytes, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(df && idf_out && scratch, RAGB_EINVAL, "ragb_bm25_build_idf: null pointer");
  RAGB_REQUIRE(vocab > 0 && corpus_size > 0, RAGB_EINVAL, "ragb_bm25_build_idf: empty shape");
  RAGB_REQUIRE(scratch_bytes >= ragb_bm25_idf_scratch_bytes(vocab), RAGB_ENOSPC, "ragb_bm25_build_idf: scratch too small");
  double* partial = static_cast<double*>(scratch);
  idf_partial_kernel<<<IDF_BLOCKS, IDF_THREADS, 0, stream>>>(df, vocab, static_cast<double>(corpus_size), partial);
  RAGB_AFTER_LAUNCH(1);
  int blocks = static_cast<int>(ceil_div64(vocab, IDF_THREADS));
  if (blocks > 148 * 8) blocks = 148 * 8;
  idf_final_kernel<<<blocks, IDF_THREADS, 0, stream>>>(df, vocab, static_cast<double>(corpus_size), epsilon, partial,
                                                       idf_out);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

int ragb_bm25_build_norm(const int32_t* doc_len, int64_t n_docs, double avgdl, double k1, double b, float* norm_out,
                         ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(doc_len && norm_out, RAGB_EINVAL, "ragb_bm25_build_norm: null pointer");
  RAGB_REQUIRE(n_docs > 0 && avgdl > 0.0, RAGB_EINVAL, "ragb_bm25_build_norm: empty shape");
  norm_kernel<<<static_cast<unsigned>(ceil_div64(n_docs, 256)), 256, 0, stream>>>(doc_len, n_docs, avgdl, k1, b, norm_out);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

size_t ragb_bm25_topk_workspace_bytes(int32_t n_queries, int64_t n_docs, int32_t k) {
  if (n_queries <= 0 || n_docs <= 0 || k <= 0) return 0;
  int64_t stripe_docs;
  const int stripes = bm25_stripes(n_queries, n_docs, &stripe_docs);
  return static_cast<size_t>(n_queries) * stripes * k * sizeof(uint64_t);
}

int ragb_bm25_score_topk(const int64_t* term_off, const int32_t* post_doc, const uint16_t* post_tf, const float* norm,
                         const float* idf, int64_t vocab, double k1, const int32_t* q_terms, const int32_t* q_off,
                         int32_t n_queries, int32_t max_query_terms, int64_t n_docs, int64_t id_base, int32_t k,
                         float* out_score, int32_t* out_id, void* workspace, size_t workspace_bytes,
                         ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = bm25_common_checks("ragb_bm25_score_topk", term_off, post_doc, post_tf, norm, idf, vocab, q_terms, q_off,
                              n_queries, n_docs, max_query_terms);
  if (rc != RAGB_OK) return rc;
  RAGB_REQUIRE(out_score && out_id && workspace, RAGB_EINVAL, "ragb_bm25_score_topk: null pointer");
  RAGB_REQUIRE(k > 0 && k <= RAGB_MAX_TOPK, RAGB_ELIMIT, "ragb_bm25_score_topk: k=%d outside [1,%d]", k, RAGB_MAX_TOPK);
  RAGB_REQUIRE(id_base >= 0 && id_base + n_docs < (1ll << 31), RAGB_ELIMIT, "ragb_bm25_score_topk: ids must fit int32");
  RAGB_REQUIRE(workspace_bytes >= ragb_bm25_topk_workspace_bytes(n_queries, n_docs, k), RAGB_ENOSPC,
               "ragb_bm25_score_topk: workspace too small");
  Bm25Args a{};
  a.term_off = term_off;
  a.post_doc = post_doc;
  a.post_tf = post_tf;
  a.norm = norm;
  a.idf = idf;
  a.q_terms = q_terms;
  a.q_off = q_off;
  a.vocab = vocab;
  a.n_docs = n_docs;
  a.id_base = id_base;
  a.k1p1 = static_cast<float>(k1 + 1.0);
  a.max_terms = max_query_terms;
  a.k = k;
  a.capacity = topk_capacity(k);
  a.part_keys = static_cast<uint64_t*>(workspace);
  a.out_scores = nullptr;
  const int stripes = bm25_stripes(n_queries, n_docs, &a.stripe_docs);
  const size_t smem = bm25_smem_bytes(a.max_terms, a.capacity, false);
  RAGB_CUDA(cudaFuncSetAttribute(bm25_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  bm25_kernel<false><<<dim3(n_queries, stripes), BM_THREADS, smem, stream>>>(a);
  RAGB_AFTER_LAUNCH(1);
  return launch_merge_keys(a.part_keys, n_queries, stripes, k, k, out_score, out_id, stream);
}

int ragb_bm25_scores(const int64_t* term_off, const int32_t* post_doc, const uint16_t* post_tf, const float* norm,
                     const float* idf, int64_t vocab, double k1, const int32_t* q_terms, const int32_t* q_off,
                     int32_t n_queries, int32_t max_query_terms, int64_t n_docs, float* out_scores,
                     ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = bm25_common_checks("ragb_bm25_scores", term_off, post_doc, post_tf, norm, idf, vocab, q_terms, q_off,
                              n_queries, n_docs, max_query_terms);
  if (rc != RAGB_OK) return rc;
  RAGB_REQUIRE(out_scores, RAGB_EINVAL, "ragb_bm25_scores: null pointer");
  Bm25Args a{};
  a.term_off = term_off;
  a.post_doc = post_doc;
  a.post_tf = post_tf;
  a.norm = norm;
  a.idf = idf;
  a.q_terms = q_terms;
  a.q_off = q_off;
  a.vocab = vocab;
  a.n_docs = n_docs;
  a.id_base = 0;
  a.k1p1 = static_cast<float>(k1 + 1.0);
  a.max_terms = max_query_terms;
  a.k = 1;
  a.capacity = 0;
  const int stripes = bm25_stripes(n_queries, n_docs, &a.stripe_docs);
  a.out_scores = out_scores;
  const size_t smem = bm25_smem_bytes(a.max_terms, 0, true);
  RAGB_CUDA(cudaFuncSetAttribute(bm25_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  bm25_kernel<true><<<dim3(n_queries, stripes), BM_THREADS, smem, stream>>>(a);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

}  // extern "C"
