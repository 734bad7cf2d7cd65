/*
 * ragb200.h - C ABI of the B200-native retrieval-scoring hot path.
 *
 * One shared library (libragb200.so, sm_100a only).  Every entry point
 *   - is extern "C", takes plain pointers and sizes (no torch / C++ types),
 *   - returns 0 (RAGB_OK) or a negative RAGB_E* code; ragb_last_error() holds the text,
 *   - enqueues work on the caller's stream and never synchronises it,
 *   - never allocates device memory: the caller owns every buffer, including the
 *     workspace whose size the matching ragb_*_workspace_bytes() reports,
 *   - refuses to run (RAGB_EARCH) on a device whose compute capability is not 10.x.
 *     There is no CPU or generic-GPU fallback.
 *
 * The reference (manikya7022/Efficient-RAG-with-Learned-Retrieval-and-Uncertainty-
 * Quantification) is pure Python and has no FFI boundary of its own;
 * each function below cites the reference call site (file:line under the reference
 * root) whose arithmetic it replaces.  Candidate ids are int32 GLOBAL passage row
 * numbers (id_base + local row); -1 marks an empty slot.  Score ties are broken the
 * same way everywhere: higher score first, then lower id.
 */
#ifndef RAGB200_H_
#define RAGB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAGB_OK        0
#define RAGB_EINVAL  (-1)   /* bad argument (null pointer, size out of range, misalignment) */
#define RAGB_EARCH   (-2)   /* device is not sm_100 (B200); nothing was launched            */
#define RAGB_ECUDA   (-3)   /* a CUDA runtime / driver call failed; see ragb_last_error()   */
#define RAGB_ELIMIT  (-4)   /* argument exceeds a documented kernel limit                   */
#define RAGB_ENOSPC  (-5)   /* workspace too small                                          */

#define RAGB_MAX_TOPK          256    /* k accepted by every *_topk entry point             */
#define RAGB_MAX_QUERY_TERMS   64     /* tokens per query accepted by ragb_bm25_*           */
#define RAGB_GEMV_MAX_BATCH    8      /* queries per call of ragb_dense_gemv_topk           */
#define RAGB_ROUTER_MAX_HIDDEN 128    /* RouterConfig.hidden_dim accepted (router.py:37)    */

typedef void* ragb_stream_t;          /* a cudaStream_t */

/* ---- library ------------------------------------------------------------------------- */
int         ragb_abi_version(void);
const char* ragb_last_error(void);                 /* thread-local, valid until next call  */
int         ragb_device_check(int device);         /* RAGB_OK iff cc major == 10           */
int64_t     ragb_launch_count(void);               /* kernels launched by this library     */

/* ---- BM25 statistics : rank_bm25 BM25Okapi._calc_idf / _initialize
 *      (reached from rag_uq/streaming_index.py:142,220) ------------------------------- */
/* idf[t] = ln(N-df+.5) - ln(df+.5) in float64; terms with idf < 0 get epsilon * mean(idf)
 * (mean over terms with df > 0, negatives included); df == 0 -> 0.  Output float32.
 * scratch: >= ragb_bm25_idf_scratch_bytes(vocab) bytes. */
size_t ragb_bm25_idf_scratch_bytes(int64_t vocab);
int ragb_bm25_build_idf(const int32_t* df, int64_t vocab, int64_t corpus_size, double epsilon,
                        float* idf_out, void* scratch, size_t scratch_bytes, ragb_stream_t stream);
/* norm[d] = k1 * (1 - b + b * doc_len[d] / avgdl), evaluated in float64, stored float32 */
int ragb_bm25_build_norm(const int32_t* doc_len, int64_t n_docs, double avgdl, double k1, double b,
                         float* norm_out, ragb_stream_t stream);

/* ---- builders of the optional dense tf table and its impact bounds (what BM25Okapi.__init__ keeps as one
 *      dict per document, rank_bm25 _initialize, reached from rag_uq/streaming_index.py:142,220) ------------
 * ragb_bm25_term_max_tf: max_tf_out[i] = largest term frequency among the postings of terms[i] (0 = empty list);
 *   a term qualifies for a table row only if this is <= 255 on EVERY shard.
 * ragb_bm25_build_dense_table: table_out[i * stride + d] = tf of terms[i] in local document d, 0 where absent
 *   (the table is cleared first); stride % 256 == 0, >= n_docs; n_terms <= 1024.
 * ragb_bm25_build_impact_bounds: dense_imp_fp16_out[r * stride + d] = the smallest IEEE fp16 >=
 *   tf / (tf + norm[d]) * (1 + 2e-6), dense_max_imp_out[r] = the maximum of row r (see ragb_bm25_score_topk). */
int ragb_bm25_term_max_tf(const int64_t* term_off, const uint16_t* post_tf, const int32_t* terms, int32_t n_terms,
                          int32_t* max_tf_out, ragb_stream_t stream);
int ragb_bm25_build_dense_table(const int64_t* term_off, const int32_t* post_doc, const uint16_t* post_tf,
                                const int32_t* terms, int32_t n_terms, int64_t n_docs, uint8_t* table_out,
                                int64_t stride, ragb_stream_t stream);
int ragb_bm25_build_impact_bounds(const uint8_t* dense_tf, int64_t dense_stride, int32_t n_dense, const float* norm,
                                  int64_t n_docs, uint16_t* dense_imp_fp16_out, float* dense_max_imp_out,
                                  ragb_stream_t stream);
/* ragb_bm25_build_posting_impacts: post_imp_out[i] = tf / (tf + norm[post_doc[i]]) of posting i, float32, evaluated
 *   with the scoring kernel's own expression (so the search adds the very bits the tf + norm path computes): the
 *   per-document part of a posting's contribution, which rank_bm25 get_scores recomputes for every query
 *   (streaming_index.py:169).  Must be rebuilt whenever norm changes (global statistics).  See ragb_bm25_score_topk. */
int ragb_bm25_build_posting_impacts(const int32_t* post_doc, const uint16_t* post_tf, const float* norm, int64_t nnz,
                                    float* post_imp_out, ragb_stream_t stream);

/* ---- BM25 scoring : rank_bm25 BM25Okapi.get_scores + BM25Index.search
 *      (rag_uq/streaming_index.py:165-179) --------------------------------------------
 * Term-major CSR over the LOCAL shard: postings of term t are
 * post_doc/post_tf[term_off[t] .. term_off[t+1]) with post_doc ascending local rows.
 * Queries are ragged lists of term ids q_terms[q_off[i] .. q_off[i+1]); every OCCURRENCE
 * contributes; ids outside [0, vocab) are out-of-vocabulary and contribute 0.
 * Optional dense tf table: dense_tf[r * dense_stride + d] (uint8, 0 = term absent) holds the
 * term frequencies of term dense_terms[r] (sorted ascending) for r < n_dense <= 1024 (terms so frequent that a
 * byte per document beats a posting list; every tf of such a term must be <= 255, and the
 * choice must be the same on every shard).  Their posting lists stay in the CSR but are not
 * read.  dense_stride is a multiple of 256, >= n_docs; n_dense = 0 disables the table.
 * Optional impact bounds for the table terms (both or neither): dense_imp_fp16[r * dense_stride + d] = an IEEE
 * fp16 UPPER bound of tf / (tf + norm[d]) (0 where the term is absent; 16-byte aligned), dense_max_imp[r] = the
 * largest value of row r.  They only prune work (a tighter bound on what the table terms can add, and a cheap
 * fp16 pass that marks the documents worth scoring exactly); results are identical with and without them.
 * Optional impact cap of the table rows (all three or none; needs the table): dense_cap[r] and the ascending local
 * rows hi_doc[hi_off[r] .. hi_off[r + 1]) of EVERY document whose tf / (tf + norm) for row r exceeds dense_cap[r].
 * A document that is on no such "marker list" gets at most weight * dense_cap[r] from row r - a much tighter promise
 * than the row maximum, which a handful of documents set - so far more queries can skip the documents no posting
 * list touches; the listed documents are always scored exactly.  Pruning only: results are identical.
 * Optional baked impacts post_imp[nnz] (ragb_bm25_build_posting_impacts; NULL = none): with them a posting is
 * self-contained (document, impact) and the pruned phase of the search reads the pair instead of chaining a gather
 * of norm[doc], a convert and a reciprocal behind every load of postings.  Speed only: results are bit-identical
 * with and without them.
 * max_query_terms (<= RAGB_MAX_QUERY_TERMS) is the caller's bound on the longest query;
 * it sizes the per-warp cursor table and longer queries are cut to it.
 * score = sum idf[t] * tf * (k1 + 1) / (tf + norm[d]).   Only score > 0 is returned
 * (streaming_index.py:176); short lists are padded with id -1 / score 0. */
size_t ragb_bm25_topk_workspace_bytes(int32_t n_queries, int64_t n_docs, int32_t k);
int ragb_bm25_score_topk(const int64_t* term_off, const int32_t* post_doc, const uint16_t* post_tf,
                         const float* norm, const float* idf, int64_t vocab, double k1,
                         const uint8_t* dense_tf, int64_t dense_stride,
                         const int32_t* dense_terms, int32_t n_dense,
                         const uint16_t* dense_imp_fp16, const float* dense_max_imp,
                         const float* dense_cap, const int32_t* hi_off, const int32_t* hi_doc,
                         const float* post_imp,
                         const int32_t* q_terms, const int32_t* q_off, int32_t n_queries,
                         int32_t max_query_terms, int64_t n_docs, int64_t id_base, int32_t k,
                         const float* seed_thr, float* out_score, int32_t* out_id,
                         void* workspace, size_t workspace_bytes, ragb_stream_t stream);
/* Threshold seeding on its own: seed_out[q] = a PROVEN lower bound of query q's k-th best score over this shard
 * (0 = none).  ragb_bm25_score_topk computes it itself when seed_thr is null; a caller that shards the corpus
 *  calls this first, takes the MAXIMUM over all shards (one all-reduce) and passes the result as
 * seed_thr, so that every shard prunes against the best bound any shard has proven - the k-th best of the whole
 * corpus is at least the k-th best of any part of it.  Results are identical with any valid bound. */
int ragb_bm25_seed(const int64_t* term_off, const int32_t* post_doc, const uint16_t* post_tf,
                   const float* norm, const float* idf, int64_t vocab, double k1,
                   const uint8_t* dense_tf, int64_t dense_stride,
                   const int32_t* dense_terms, int32_t n_dense,
                   const int32_t* q_terms, const int32_t* q_off, int32_t n_queries,
                   int32_t max_query_terms, int64_t n_docs, int32_t k, float* seed_out, ragb_stream_t stream);
/* The same search in stages, for a caller that overlaps it with another kernel (HybridEngine, --overlap): the
 * documents are cut into ragb_bm25_stripe_count(n_queries, n_docs) stripes; ragb_bm25_score_part scores the stripes
 * [stripe_begin, stripe_end) into the workspace (the part with stripe_begin == 0 must come first: it also installs the
 * thresholds, from seed_thr or the seed kernel); ragb_bm25_score_finish merges all stripes once every one of them has
 * been scored.  min_smem_bytes pads the shared memory of every block (0 = natural size): a padding of ~96 KB leaves
 * room for exactly one such block next to a resident 4-stage ragb_dense_mma_* block on every SM.  Parts that are
 * scored later start from the thresholds the earlier parts have proven.  Results equal ragb_bm25_score_topk. */
int32_t ragb_bm25_stripe_count(int32_t n_queries, int64_t n_docs);
int ragb_bm25_score_part(const int64_t* term_off, const int32_t* post_doc, const uint16_t* post_tf,
                         const float* norm, const float* idf, int64_t vocab, double k1,
                         const uint8_t* dense_tf, int64_t dense_stride,
                         const int32_t* dense_terms, int32_t n_dense,
                         const uint16_t* dense_imp_fp16, const float* dense_max_imp,
                         const float* dense_cap, const int32_t* hi_off, const int32_t* hi_doc,
                         const float* post_imp,
                         const int32_t* q_terms, const int32_t* q_off, int32_t n_queries,
                         int32_t max_query_terms, int64_t n_docs, int64_t id_base, int32_t k,
                         const float* seed_thr, int32_t stripe_begin, int32_t stripe_end, int64_t min_smem_bytes,
                         void* workspace, size_t workspace_bytes, ragb_stream_t stream);
int ragb_bm25_score_finish(int32_t n_queries, int64_t n_docs, int32_t k, float* out_score, int32_t* out_id,
                           const void* workspace, size_t workspace_bytes, ragb_stream_t stream);
/* Same arithmetic, full score vectors (get_scores itself).
 * tiled = 0: out_scores[q * out_ld + d], out_ld >= n_docs (row-major).
 * tiled = 1: out_scores[((d / 256) * out_ld + q) * 256 + d % 256], out_ld >= n_queries: 256-document tiles
 *            with the query rows of a tile next to each other, ceil(n_docs / 256) * out_ld * 256 floats.
 *            This is the layout ragb_dense_mma_fused_topk reads (one contiguous piece per GEMM tile). */
int ragb_bm25_scores(const int64_t* term_off, const int32_t* post_doc, const uint16_t* post_tf,
                     const float* norm, const float* idf, int64_t vocab, double k1,
                     const uint8_t* dense_tf, int64_t dense_stride,
                     const int32_t* dense_terms, int32_t n_dense,
                     const int32_t* q_terms, const int32_t* q_off, int32_t n_queries,
                     int32_t max_query_terms, int64_t n_docs, float* out_scores, int64_t out_ld,
                     int32_t tiled, ragb_stream_t stream);

/* ---- dense scoring : DenseIndex.search (rag_uq/streaming_index.py:353-370) -----------
 * Exact inner product of bf16 query rows with bf16 passage rows (unit rows -> cosine,
 * i.e. the reference's 1 - distance), fp32 accumulation, fused per-query top-k.
 * passages [n_rows, dim] row-major bf16, 16-byte aligned; dim % 64 == 0. */
size_t ragb_dense_gemv_workspace_bytes(int32_t n_queries, int32_t k);
int ragb_dense_gemv_topk(const void* passages_bf16, int64_t n_rows, int32_t dim,
                         const void* queries_bf16, int32_t n_queries, int32_t k, int64_t id_base,
                         float* out_score, int32_t* out_id,
                         void* workspace, size_t workspace_bytes, ragb_stream_t stream);
/* tcgen05 / TMEM / TMA path for query batches.  n_queries is padded internally to a
 * multiple of 128 (queries_bf16 must hold n_queries rows; padding rows are zero-filled
 * by TMA out-of-bounds handling).  variant: 0 = A (queries) and B (passages) both
 * streamed through shared memory, 128-passage tiles; 1 = query slab resident in TMEM;
 * 2 = as 0 with 256-passage tiles; 3 = CTA pairs (tcgen05 cta_group::2, 256 queries x 256
 * passages per MMA, falls back to 2 for a single slab).  */
size_t ragb_dense_mma_workspace_bytes(int32_t n_queries, int32_t k);
int ragb_dense_mma_topk(const void* passages_bf16, int64_t n_rows, int32_t dim,
                        const void* queries_bf16, int32_t n_queries, int32_t k, int64_t id_base,
                        int32_t variant, float* out_score, int32_t* out_id,
                        void* workspace, size_t workspace_bytes, ragb_stream_t stream);
/* ragb_dense_mma_topk that ALSO reports the smallest score of every query: min_inout[n_queries] must hold +inf (or
 * a running minimum of other shards) on entry and receives min(min_inout[q], min over this shard's rows of the score);
 * rows past the end of the shard count as 0 (the result is a valid LOWER bound of the true minimum).  With the k-th
 * largest dense score it brackets the dense score of every passage outside the top-k list - what the stopping rule of
 * the threshold-algorithm full-fusion needs (engine.full_fusion_topk). */
int ragb_dense_mma_topk_min(const void* passages_bf16, int64_t n_rows, int32_t dim,
                            const void* queries_bf16, int32_t n_queries, int32_t k, int64_t id_base,
                            int32_t variant, float* out_score, int32_t* out_id, float* min_inout,
                            void* workspace, size_t workspace_bytes, ragb_stream_t stream);
/* The same search in its two phases, for callers that shard the corpus.  On large shards ragb_dense_mma_topk
 * first searches a sampled prefix (1/32 of the passage tiles): the k-th best score found there is a proven lower
 * bound of the final k-th best, and every candidate list of the remaining tiles starts from it instead of from
 * -inf.  ragb_dense_mma_sample runs the first phase (its merged list stays in the workspace) and reports the
 * bounds in thr_out[n_queries] (-inf = none); the caller may raise them to the MAXIMUM over all shards (one
 * all-reduce, shared with ragb_bm25_seed); ragb_dense_mma_seeded then searches the rest with thr and merges both
 * phases.  Same workspace (contents preserved between the calls), same arguments; results are identical with
 * any valid bounds. */
int ragb_dense_mma_sample(const void* passages_bf16, int64_t n_rows, int32_t dim,
                          const void* queries_bf16, int32_t n_queries, int32_t k, int64_t id_base,
                          int32_t variant, float* thr_out,
                          void* workspace, size_t workspace_bytes, ragb_stream_t stream);
int ragb_dense_mma_seeded(const void* passages_bf16, int64_t n_rows, int32_t dim,
                          const void* queries_bf16, int32_t n_queries, int32_t k, int64_t id_base,
                          int32_t variant, const float* thr, float* out_score, int32_t* out_id,
                          void* workspace, size_t workspace_bytes, ragb_stream_t stream);
/* Full-fusion mode: RetrievalRouter.hybrid_rerank (rag_uq/router.py:179-202) evaluated over ALL
 * passages of the shard inside the epilogue of the tcgen05 GEMM: for every (query, passage)
 *   fused = gate(bm25, dense) * dense + (1 - gate) * bm25,   gate = RetrievalRouter.forward with
 * running statistics (stats_initialized == True, router.py:130-132), dense = the accumulator in TMEM,
 * bm25 = the TILED matrix ragb_bm25_scores(..., tiled = 1) wrote with out_ld = bm25_rows
 * (32-byte aligned).  Output: the k best fused scores per query and their global ids; neither the
 * dense nor the fused score matrix ever exists in memory.
 * gate_bound_table[ib * n_d + id] (n_d a power of two, n_b * n_d <= 8192 cells, staged in shared
 * memory) packs two bfloat16 numbers, bits 0-15 = lo and bits 16-31 = hi, with lo <= gate(bm25, dense) <= hi over the
 * cell  bm25 in [ib, ib+1) * b_cap / n_b  x  dense in -d_hi + [id, id+1) * 2 d_hi / n_d; the last
 * bm25 row (which also receives bm25 >= b_cap and bm25 < 0) must hold lo = 0, hi = 1; d_hi must
 * bound |dense| for every pair (e.g. max passage norm x max query norm).
 * The table only prunes gate evaluations (a pair is skipped when
 * bm25 + (dense <= bm25 ? lo : hi) * (dense - bm25) is below the query's running k-th best fused
 * score), it never changes results: lo = 0, hi = 1 everywhere is valid (and slow).
 * rag_uq_b200.router.full_fusion_bounds derives a tight table from the router's weights.
 * counters: NULL or 2 device uint64 that receive (gate evaluations, list admissions).
 * Workspace: ragb_dense_mma_workspace_bytes(n_queries, k). */
int ragb_dense_mma_fused_topk(const void* passages_bf16, int64_t n_rows, int32_t dim,
                              const void* queries_bf16, int32_t n_queries, int32_t k, int64_t id_base,
                              const float* bm25_scores, int64_t bm25_rows,
                              const float* w1, const float* b1, const float* w2, const float* b2,
                              const float* stats, int32_t hidden,
                              const uint32_t* gate_bound_table, int32_t n_b, int32_t n_d, float b_cap, float d_hi,
                              float* out_score, int32_t* out_id, unsigned long long* counters,
                              void* workspace, size_t workspace_bytes, ragb_stream_t stream);
/* ---- exact scores of GIVEN (query, passage) pairs -------------------------------------------------------------
 * The random-access companions of the two streaming scorers, used by the threshold-algorithm form of full-fusion
 * (RetrievalRouter.hybrid_rerank over all passages, rag_uq/router.py:179-202, without a [B, N] matrix): each side's
 * streaming kernel delivers its exact ranked list, these fill in the OTHER side's score of every listed passage.
 * cand_ids [n_queries, n_cand]: global ids, -1 (or an id outside this shard) = none -> score 0.
 * ragb_bm25_score_docs: rank_bm25 get_scores (streaming_index.py:169) for the chosen documents, same arithmetic and
 *   summation order as ragb_bm25_score_topk: the scores are bit-identical to the streaming kernel's.
 * ragb_dense_score_docs: the bf16 inner product of DenseIndex.search (streaming_index.py:353-370) for the chosen rows,
 *   fp32 accumulation (equal to the tensor-core result to ~1 ulp, not bit for bit). */
int ragb_bm25_score_docs(const int64_t* term_off, const int32_t* post_doc, const uint16_t* post_tf,
                         const float* norm, const float* idf, int64_t vocab, double k1,
                         const uint8_t* dense_tf, int64_t dense_stride,
                         const int32_t* dense_terms, int32_t n_dense,
                         const int32_t* q_terms, const int32_t* q_off, int32_t n_queries,
                         int32_t max_query_terms, int64_t n_docs, int64_t id_base,
                         const int32_t* cand_ids, int32_t n_cand, float* out_scores, ragb_stream_t stream);
int ragb_dense_score_docs(const void* passages_bf16, int64_t n_rows, int32_t dim,
                          const void* queries_bf16, int32_t n_queries, int64_t id_base,
                          const int32_t* cand_ids, int32_t n_cand, float* out_scores, ragb_stream_t stream);
/* Plain score matrix out[n_queries, n_rows] fp32 (small shapes, tests, un-fused full-fusion mode). */
int ragb_dense_scores(const void* passages_bf16, int64_t n_rows, int32_t dim,
                      const void* queries_bf16, int32_t n_queries, float* out_scores,
                      ragb_stream_t stream);

/* ---- selection ------------------------------------------------------------------------
 * np.argsort(scores)[::-1][:k] (streaming_index.py:172) / torch.topk (router.py:202) */
size_t ragb_topk_rows_workspace_bytes(int32_t n_rows, int64_t n_cols, int32_t k);
int ragb_topk_rows(const float* scores, int32_t n_rows, int64_t n_cols, int32_t k,
                   float* out_score, int32_t* out_index,
                   void* workspace, size_t workspace_bytes, ragb_stream_t stream);
/* Merge n_lists candidate lists per query ([n_queries, n_lists, k_in], id -1 = empty) into
 * one top-k_out; used for intra-GPU stripes and for the lists all-gathered across GPUs. */
int ragb_topk_merge(const float* in_score, const int32_t* in_id, int32_t n_queries, int32_t n_lists,
                    int32_t k_in, int32_t k_out, float* out_score, int32_t* out_id,
                    ragb_stream_t stream);
/* ragb_topk_merge reading the lists in place from a strided buffer: element j of list l of query q is at
 * q * query_stride + l * list_stride + j (in elements) of in_score / in_id - e.g. the all-gathered exchange buffer
 * [ranks, queries, 2, pools] of a sharded search, merged without a transposing copy. */
int ragb_topk_merge_strided(const float* in_score, const int32_t* in_id, int32_t n_queries, int32_t n_lists,
                            int32_t k_in, int64_t query_stride, int64_t list_stride, int32_t k_out,
                            float* out_score, int32_t* out_id, ragb_stream_t stream);

/* ---- pool fusion : HybridRetriever.hybrid_search (rag_uq/streaming_index.py:484-523)
 * Union of the two pools by id, missing score = 0.0, each score divided by the union
 * maximum ("max(...) or 1"), averaged, sorted descending, cut to k.  Padded with id -1. */
int ragb_hybrid_fuse_topk(const float* bm25_score, const int32_t* bm25_id,
                          const float* dense_score, const int32_t* dense_id,
                          int32_t n_queries, int32_t pool, int32_t k,
                          int32_t* out_id, float* out_bm25, float* out_dense, float* out_hybrid,
                          ragb_stream_t stream);

/* ---- retrieval uncertainty : docs/uncertainty_theory.md:48-56 ("next" row N4) ---------
 * U = std(s_top-k) + lambda * (1 - |s_1 - s_k|) over the valid entries (id >= 0) of each of
 * the n_queries ranked lists [n_queries, k]; population std; an empty list gives lambda. */
int ragb_retrieval_uncertainty(const float* score, const int32_t* id, int32_t n_queries, int32_t k,
                               double lambda, float* out, ragb_stream_t stream);

/* ---- router gate : RetrievalRouter.forward / hybrid_rerank (rag_uq/router.py:100-202)
 * Inputs bm25 / dense [n_rows, n_cand] fp32.  w1 [hidden,3], b1 [hidden], w2 [hidden],
 * b2 [1], stats [4] = bm25_mean, bm25_std, dense_mean, dense_std - all DEVICE float32
 * (the reference's state-dict tensors, untouched).  norm_mode:
 *   1 = running statistics (stats_initialized == True, router.py:130-132)
 *   0 = statistics of this very call over the whole [n_rows, n_cand] input
 *       (router.py:133-136; torch.std is the unbiased estimator, one element -> NaN)
 *   2 = the same, per row: what the reference's evaluation loop computes when it calls
 *       the router once per query with a [1, P] tensor (experiments/run_evaluation.py:171-177)
 * out_gate = sigmoid(...) in (0,1); out_fused = gate*dense + (1-gate)*bm25 on the RAW
 * scores (router.py:199).  Either output may be NULL.  scratch: device memory of at least
 * ragb_router_scratch_bytes(n_rows, norm_mode) bytes (unused for norm_mode 1). */
size_t ragb_router_scratch_bytes(int32_t n_rows, int32_t norm_mode);
int ragb_router_forward(const float* bm25, const float* dense, int32_t n_rows, int32_t n_cand,
                        const float* w1, const float* b1, const float* w2, const float* b2,
                        const float* stats, int32_t hidden, int32_t norm_mode,
                        float* out_gate, float* out_fused, void* scratch, ragb_stream_t stream);

/* ---- MC-Dropout : nn.Dropout (router.py:78) sampled T times, aggregated as
 *      MCDropoutConfidence does (rag_uq/confidence.py:195-202, 258-264) ----------------
 * One Philox4x32-10 stream (curand layout) seeded (seed, offset); mask_layout 1
 * reproduces torch's fused CUDA dropout element->counter mapping for a [B*P, hidden]
 * tensor on a device with sm_count SMs, pass t using offset + t * increment;
 * mask_layout 0 is the library's own (subsequence = candidate, draw = t*hidden/4+u/4).
 * Outputs [n_queries, n_cand]: mean/std (population) of gate and fused score;
 * per query: variance (std of L2 distances of the T gate vectors to their centroid),
 * consensus sample index.  mask_dump (uint8 [T, B*P, hidden]) and gate_dump
 * (float [T, B, P]) may be NULL.  norm_mode and scratch as for ragb_router_forward. */
int ragb_router_mc_dropout(const float* bm25, const float* dense, int32_t n_queries, int32_t n_cand,
                           const float* w1, const float* b1, const float* w2, const float* b2,
                           const float* stats, int32_t hidden, int32_t norm_mode,
                           int32_t n_samples, double p_drop, uint64_t seed, uint64_t offset,
                           int32_t mask_layout, int32_t sm_count,
                           float* mean_gate, float* std_gate, float* mean_fused, float* std_fused,
                           float* variance, int32_t* consensus,
                           uint8_t* mask_dump, float* gate_dump, void* scratch, ragb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RAGB200_H_ */
