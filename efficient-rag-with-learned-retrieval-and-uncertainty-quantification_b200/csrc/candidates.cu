// Exact scores of GIVEN (query, passage) pairs: the random-access companions of the two streaming scorers.
//
// Used by the threshold-algorithm form of full-fusion (RetrievalRouter.hybrid_rerank over all passages,
// rag_uq/router.py:179-202, without ever forming a [B, N] matrix): the streaming kernels deliver each side's exact
// ranked list, these kernels fill in the OTHER side's score of every listed passage, and a proven bound on everything
// unlisted ends the search (engine.full_fusion_topk).
//
//   bm25_score_docs   rank_bm25 get_scores (rag_uq/streaming_index.py:169) evaluated for chosen documents only, with
//                     the arithmetic AND the summation order of bm25_kernel (table terms in query order, then posting-
//                     list terms in query order, the two partial sums added last), so a document's score is bit-identical
//                     to what the streaming kernel computes for it.
//   dense_score_docs  the inner product of DenseIndex.search (streaming_index.py:353-370) for chosen rows, fp32
//                     accumulation (not the tensor core's summation order: equal to ~1 ulp, not bit for bit).
#include <climits>

#include "common.cuh"

namespace ragb {

constexpr int CD_THREADS = 256;
constexpr int CD_WARPS = CD_THREADS / 32;
constexpr int CD_MAX_TERMS = 64;

__device__ __forceinline__ float cd_fast_rcp(float x) {   // the expression bm25_kernel uses
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

struct ScoreDocsArgs {
  const int64_t* term_off;
  const int32_t* post_doc;
  const uint16_t* post_tf;
  const float* norm;
  const float* idf;
  const int32_t* q_terms;
  const int32_t* q_off;
  const uint8_t* dense_tf;
  const int32_t* dense_terms;
  int64_t dense_stride;
  int64_t vocab;
  int64_t n_docs;
  int64_t id_base;
  int n_dense;
  int max_terms;
  float k1p1;
  const int32_t* cand;   // [n_queries, n_cand] global ids, -1 = none
  int n_cand;
  float* out;            // [n_queries, n_cand]
};

// one warp per (query, candidate); lane t evaluates query term t (and t + 32), lane 0 sums in the kernel's order
__global__ void __launch_bounds__(CD_THREADS) bm25_score_docs_kernel(const ScoreDocsArgs a, int n_queries) {
  const int lane = threadIdx.x & 31;
  const int64_t pair = static_cast<int64_t>(blockIdx.x) * CD_WARPS + (threadIdx.x >> 5);
  if (pair >= static_cast<int64_t>(n_queries) * a.n_cand) return;
  const int q = static_cast<int>(pair / a.n_cand);
  const int32_t gid = a.cand[pair];
  if (gid < 0 || gid < a.id_base || gid >= a.id_base + a.n_docs) {
    if (lane == 0) a.out[pair] = 0.0f;
    return;
  }
  const int doc = static_cast<int>(gid - a.id_base);
  const float nrm = __ldg(a.norm + doc);
  const int qb = a.q_off[q];
  const int nt = min(a.q_off[q + 1] - qb, a.max_terms);
  float w[2], x[2];
  int kind[2];   // 0 = contributes nothing, 1 = table term, 2 = posting-list term that holds the document
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int ti = r * 32 + lane;
    w[r] = 0.0f;
    x[r] = 0.0f;
    kind[r] = 0;
    if (ti < nt) {
      const int t = a.q_terms[qb + ti];
      if (t >= 0 && t < a.vocab) {
        const float wt = a.idf[t] * a.k1p1;
        if (wt != 0.0f) {
          int lo = 0, hi = a.n_dense;   // lower_bound in the sorted table directory
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(a.dense_terms + mid) < t) lo = mid + 1; else hi = mid;
          }
          if (lo < a.n_dense && __ldg(a.dense_terms + lo) == t) {
            const unsigned tfb = __ldg(a.dense_tf + static_cast<int64_t>(lo) * a.dense_stride + doc);
            const float f = __uint_as_float(0x4B000000u | tfb) - 8388608.0f;
            w[r] = wt;
            x[r] = f * cd_fast_rcp(f + nrm);
            kind[r] = 1;
          } else {
            int64_t p0 = a.term_off[t], p1 = a.term_off[t + 1];   // lower_bound of doc in the term's posting list
            const int64_t end = p1;
            while (p0 < p1) {
              const int64_t mid = (p0 + p1) >> 1;
              if (__ldg(a.post_doc + mid) < doc) p0 = mid + 1; else p1 = mid;
            }
            if (p0 < end && __ldg(a.post_doc + p0) == doc) {
              const float f = static_cast<float>(__ldg(a.post_tf + p0));
              w[r] = wt;
              x[r] = f * cd_fast_rcp(f + nrm);
              kind[r] = 2;
            }
          }
        }
      }
    }
  }
  // the kernel's order: every table term in query order (an absent term adds w * 0), then the list terms that hold
  // the document in query order (the first one starts the sum), then table part + list part
  float table_sum = 0.0f, list_sum = 0.0f;
  bool any_list = false;
  for (int ti = 0; ti < nt; ++ti) {
    const int r = ti >> 5, src = ti & 31;
    const float wt = __shfl_sync(0xffffffffu, r == 0 ? w[0] : w[1], src);
    const float xt = __shfl_sync(0xffffffffu, r == 0 ? x[0] : x[1], src);
    const int kd = __shfl_sync(0xffffffffu, r == 0 ? kind[0] : kind[1], src);
    if (kd == 1) table_sum = fmaf(wt, xt, table_sum);
    if (kd == 2) {
      list_sum = fmaf(wt, xt, any_list ? list_sum : 0.0f);
      any_list = true;
    }
  }
  if (lane == 0) a.out[pair] = any_list ? table_sum + list_sum : table_sum;
}

// one warp per (query, candidate): fp32 dot product of two bf16 rows (dim % 8 == 0)
__global__ void __launch_bounds__(CD_THREADS) dense_score_docs_kernel(const uint4* __restrict__ passages, int64_t n_rows, int dim,
                                                                      const uint4* __restrict__ queries, int n_queries,
                                                                      const int32_t* __restrict__ cand, int n_cand,
                                                                      int64_t id_base, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t pair = static_cast<int64_t>(blockIdx.x) * CD_WARPS + (threadIdx.x >> 5);
  if (pair >= static_cast<int64_t>(n_queries) * n_cand) return;
  const int q = static_cast<int>(pair / n_cand);
  const int32_t gid = cand[pair];
  if (gid < 0 || gid < id_base || gid >= id_base + n_rows) {
    if (lane == 0) out[pair] = 0.0f;
    return;
  }
  const int vec = dim >> 3;
  const uint4* prow = passages + static_cast<int64_t>(gid - id_base) * vec;
  const uint4* qrow = queries + static_cast<int64_t>(q) * vec;
  float s = 0.0f;
  for (int c = lane; c < vec; c += 32) {
    float e[8], f[8];
    unpack_bf16x8(__ldg(prow + c), e);
    unpack_bf16x8(__ldg(qrow + c), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(e[i], f[i], s);
  }
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
  if (lane == 0) out[pair] = s;
}

}  // namespace ragb

using namespace ragb;

extern "C" {

int ragb_bm25_score_docs(const int64_t* term_off, const int32_t* post_doc, const uint16_t* post_tf, const float* norm,
                         const float* idf, int64_t vocab, double k1, const uint8_t* dense_tf, int64_t dense_stride,
                         const int32_t* dense_terms, int32_t n_dense, const int32_t* q_terms, const int32_t* q_off,
                         int32_t n_queries, int32_t max_query_terms, int64_t n_docs, int64_t id_base, const int32_t* cand_ids,
                         int32_t n_cand, float* out_scores, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(term_off && post_doc && post_tf && norm && idf && q_terms && q_off && cand_ids && out_scores, RAGB_EINVAL,
               "ragb_bm25_score_docs: null pointer");
  RAGB_REQUIRE(vocab > 0 && n_queries > 0 && n_docs > 0 && n_cand > 0, RAGB_EINVAL, "ragb_bm25_score_docs: empty shape");
  RAGB_REQUIRE(max_query_terms >= 1 && max_query_terms <= CD_MAX_TERMS, RAGB_ELIMIT,
               "ragb_bm25_score_docs: max_query_terms=%d outside [1,%d]", max_query_terms, CD_MAX_TERMS);
  RAGB_REQUIRE(n_dense >= 0 && (n_dense == 0 || (dense_tf && dense_terms && dense_stride >= n_docs)), RAGB_EINVAL,
               "ragb_bm25_score_docs: bad dense table");
  ScoreDocsArgs a{};
  a.term_off = term_off;
  a.post_doc = post_doc;
  a.post_tf = post_tf;
  a.norm = norm;
  a.idf = idf;
  a.q_terms = q_terms;
  a.q_off = q_off;
  a.dense_tf = dense_tf;
  a.dense_terms = dense_terms;
  a.dense_stride = dense_stride;
  a.vocab = vocab;
  a.n_docs = n_docs;
  a.id_base = id_base;
  a.n_dense = n_dense;
  a.max_terms = max_query_terms;
  a.k1p1 = static_cast<float>(k1 + 1.0);
  a.cand = cand_ids;
  a.n_cand = n_cand;
  a.out = out_scores;
  const int64_t pairs = static_cast<int64_t>(n_queries) * n_cand;
  bm25_score_docs_kernel<<<static_cast<unsigned>(ceil_div64(pairs, CD_WARPS)), CD_THREADS, 0, stream>>>(a, n_queries);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

int ragb_dense_score_docs(const void* passages_bf16, int64_t n_rows, int32_t dim, const void* queries_bf16,
                          int32_t n_queries, int64_t id_base, const int32_t* cand_ids, int32_t n_cand, float* out_scores,
                          ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(passages_bf16 && queries_bf16 && cand_ids && out_scores, RAGB_EINVAL, "ragb_dense_score_docs: null pointer");
  RAGB_REQUIRE(n_rows > 0 && n_queries > 0 && n_cand > 0 && dim > 0 && dim % 8 == 0, RAGB_EINVAL, "ragb_dense_score_docs: bad shape");
  RAGB_REQUIRE(((reinterpret_cast<uintptr_t>(passages_bf16) | reinterpret_cast<uintptr_t>(queries_bf16)) & 15) == 0,
               RAGB_EINVAL, "ragb_dense_score_docs: inputs must be 16-byte aligned");
  const int64_t pairs = static_cast<int64_t>(n_queries) * n_cand;
  dense_score_docs_kernel<<<static_cast<unsigned>(ceil_div64(pairs, CD_WARPS)), CD_THREADS, 0, stream>>>(
      static_cast<const uint4*>(passages_bf16), n_rows, dim, static_cast<const uint4*>(queries_bf16), n_queries, cand_ids, n_cand,
      id_base, out_scores);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

}  // extern "C"
