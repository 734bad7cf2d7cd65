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
// Kernel shape.  grid = (queries, stripes), enough stripes for ~64 blocks per SM (queries differ widely in
// cost once pruning works, fine stripes balance the SMs); a block of 8 warps owns one stripe of consecutive
// documents for one query, each WARP owns a contiguous eighth of it.  A warp goes through up to three phases:
//
//  1. Dense accumulator (1024-document super-ranges, the only phase of the get_scores variant).  Two ways a
//     term reaches the accumulator:
//     * the terms that occur in more than 1/24 of all documents (~220; they carry >95% of all postings) are
//       ALSO stored as a dense row of uint8 term frequencies; a lane loads its 8 bytes with one 64-bit load
//       per term and accumulates in registers - no document ids, no cursor, no compare;
//     * every other term streams its posting list: lists are sorted by document, so a warp continues reading
//       where it stopped (one cursor per term, placed once by a 32-ary search), 128-byte coalesced, several
//       independent 32-posting chunks per pass (1 to 8 by the list's density in the get_scores variant; ONE variant
//       per phase in the top-k kernel, whose ~125 KB of code made instruction fetch a measurable stall), into a
//       shared-memory accumulator that covers the super-range.  All documents inside one chunk are distinct, so the update is a plain
//       read-modify-write: no atomics on scores.  Terms whose next posting lies beyond the super-range are
//       skipped with one ballot.
//  2. fp16 bound pass (top-k variant, while thr is a sizeable fraction of what the table terms can add):
//     an fp16 UPPER bound of every document's table score (one 128-bit load + four HFMA2 per term per 8
//     documents) marks the few documents that may still reach thr; only those and the ones a posting list
//     touched get the exact arithmetic.  With it the exact table path of phase 1 is never taken at 10M x 1024.
//  3. Window mode (once thr exceeds the table bound): only documents a posting list touches matter - ~3% of
//     the corpus per query - so the warp covers up to 32k documents per visit and keeps the accumulator as an
//     open-addressing hash table in the same 4 KB of shared memory (see stream_term_hash).  With baked impacts
//     (Bm25Args::post_imp) a posting is a (document, impact) pair and nothing is gathered behind it.  MaxScore marks
//     only documents of essential (heavy, rare) lists, and marked documents whose list part plus the table bound
//     cannot reach thr are dropped before any table byte is gathered.
//
// Which terms use the table is decided from GLOBAL document frequencies, so every shard makes the same
// choice, and all phases accumulate with the same fused multiply-adds in the same order (table terms in query
// order, then list terms in query order): a document's score does not depend on the phase that computed it
// nor on how the corpus is sharded.  Every pruning step is exact (proven bounds only).
// Selection is warp-private too (WarpTopK in topk.cuh: threshold in a register, appends through a
// ballot prefix, rare warp-level bitonic merge), so the main loop has no block barrier at all; the
// 8 warp lists are folded once at the end of the block and the per-block lists of all stripes
// are merged by topk_merge_kernel.  bm25_seed_kernel proves a lower bound of each query's k-th best score
// from its posting lists so that most queries start directly in phase 3.
#include <cuda_fp16.h>

#include <climits>
#include <cstdlib>

#include "common.cuh"
#include "topk.cuh"

namespace ragb {

constexpr int BM_THREADS = 256;
constexpr int BM_WARPS = BM_THREADS / 32;
constexpr int BM_RANGE = 256;          // documents per warp range (8 per lane)
constexpr int BM_SUPER = 4;            // ranges per super-range: posting lists are streamed once per 1024 documents
constexpr int BM_SUPER_DOCS = BM_RANGE * BM_SUPER;
constexpr int BM_MAX_TERMS = 64;
constexpr int BM_MAX_DENSE = 1024;  // rows of the dense tf table (term ids sorted ascending)
#ifndef RAGB_BM_MIN_BLOCKS
#define RAGB_BM_MIN_BLOCKS 3
#endif
// Resident blocks per SM the register budget is sized for.  4 (64 registers) spilled 80 bytes in the window loop and
// only 23 of its 32 warps were resident on average anyway; 3 (80 registers, no spills): 14.0 -> 11.6 ms at 10M x 1024.
constexpr int BM_MIN_BLOCKS = RAGB_BM_MIN_BLOCKS;
constexpr int BM_SEARCH = 4;  // posting lists searched concurrently while placing the cursors
constexpr int BM_APPROX_MAX = 320;  // fp16 bound pass: fall back to the exact table path when more documents of a super-range pass

// debug counters (RAGB_BM25_DEBUG=1): [0] super-ranges full mode, [1] pruned, [2] pruned with postings,
// [3] documents scored in pruned mode, [4] queries seeded, [5] queries with seed > table bound,
// [6] super-ranges decided by the fp16 bound pass, [7] bound passes that fell back to the exact table path
__device__ unsigned long long g_bm25_dbg[8];
#ifdef RAGB_BM25_PROFILE
// Profiling build only (scripts/profile_bm25_queries.py): warp cycles per query and phase, [phase][query]:
// 0 set-up (term split, cursor placement), 1 window mode, 2 dense-accumulator / bound-pass super-ranges, 3 block fold,
// inside window mode: 4 posting streaming + hash inserts, 5 compaction, 6 scoring of the marked documents; 7 the part of
// the fold spent waiting for the slowest warp of the block.
constexpr int BM_PROF_QUERIES = 4096;
__device__ unsigned long long g_bm25_prof[8][BM_PROF_QUERIES];
#define BM_PROF_T(var) const long long var = clock64()
#define BM_PROF_ADD(phase, q, cycles) do { if (lane == 0 && (q) < BM_PROF_QUERIES) atomicAdd(&g_bm25_prof[phase][q], static_cast<unsigned long long>(cycles)); } while (0)
#else
#define BM_PROF_T(var)
#define BM_PROF_ADD(phase, q, cycles)
#endif

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
  int64_t stripe_docs;  // multiple of BM_WARPS * BM_SUPER_DOCS
  float k1p1;
  int max_terms;
  int k;
  int capacity;
  uint64_t* part_keys;  // [queries, stripes, k]            (top-k mode)
  float* out_scores;    // [queries, out_ld]                (dense mode)
  int64_t out_ld;       // row stride of out_scores in floats (>= n_docs); tiled: query rows per tile
  int out_tiled;        // 1: out_scores[(d / 256) * out_ld + q][d % 256] (256-document tiles, query rows inside a tile)
  const uint8_t* dense_tf;   // [n_dense, dense_stride] tf of the most frequent terms, 0 = absent
  const int32_t* dense_terms;  // [n_dense] term id of each row
  int64_t dense_stride;      // multiple of BM_RANGE, >= n_docs
  int n_dense;
  const __half* dense_imp;   // optional [n_dense, dense_stride]: fp16 UPPER bound of tf / (tf + norm) per (table term, document)
  const float* dense_maximp; // optional [n_dense]: row maxima of dense_imp (a term contributes at most weight * maximp)
  // optional impact cap of the table rows: every document of row r whose impact bound exceeds dense_cap[r] is listed in
  // hi_doc[hi_off[r] .. hi_off[r + 1]) (ascending); all others contribute at most weight * dense_cap[r]
  const float* dense_cap;
  const int32_t* hi_off;
  const int32_t* hi_doc;
  // optional [nnz]: post_imp[i] = tf / (tf + norm[doc]) of posting i, evaluated with the kernel's own expression (see
  // posting_impacts_kernel), so reading it gives the bits the tf + norm path computes.  With it the window phase
  // reads (document, impact) pairs and drops the dependent gather of norm[doc] and the reciprocal from its chain.
  const float* post_imp;
  int window_mode;           // > 0: hash-window mode for the pruned phase, aiming at this many postings per window
  float* seed_thr;           // [queries] proven lower bound of each query's k-th best score (0 = none)
  int stripe0;               // this launch covers stripes [stripe0, stripe0 + gridDim.y) of n_stripes
  int n_stripes;
  int debug;
};

// 1/x for x in the normal range (here x = tf + norm in [0.3, 7e4]): one MUFU.RCP, none of the
// range-scaling __fdividef wraps around it.  Both accumulation paths use this same expression, so
// a term scores identically whether it is read from the dense table or from its posting list.
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Stream the postings of one term that fall below d1 into the warp's accumulator.
// U chunks of 32 postings are loaded per pass (all loads independent) plus one "peek" posting
// right behind them, so a pass that consumes everything it loaded still learns the next
// document without another round trip to memory.
// MARK_ONLY: a marker list (documents of a table row above its impact cap): no term frequencies, nothing is added to
// the accumulator, the documents are only marked for exact scoring.
template <int U, bool MARK_ONLY = false>
__device__ __forceinline__ void stream_term(const int32_t* __restrict__ post_doc,
                                            const uint16_t* __restrict__ post_tf, int64_t& pos, const int64_t end,
                                            const int d0, const int d1, const float weight, float* accw,
                                            const float* __restrict__ norm, unsigned* touched, const int lane,
                                            int& next_doc) {
  while (true) {
    int doc[U];
    unsigned tf[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t idx = pos + u * 32 + lane;
      const bool in = idx < end;
      doc[u] = in ? __ldg(post_doc + idx) : INT_MAX;
      tf[u] = (in && !MARK_ONLY) ? static_cast<unsigned>(__ldg(post_tf + idx)) : 0u;
    }
    const int64_t peek_idx = pos + U * 32;
    int peek = INT_MAX;
    if (lane == 0 && peek_idx < end) peek = __ldg(post_doc + peek_idx);
    int taken = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool take = doc[u] < d1;
      if (take) {
        const int o = doc[u] - d0;
        if (!MARK_ONLY) {
          const float f = static_cast<float>(tf[u]);
          accw[o] = fmaf(weight, f * fast_rcp(f + __ldg(norm + doc[u])), accw[o]);   // same expression in stream_term_hash
        }
        if (touched != nullptr) atomicOr(touched + (o >> 5), 1u << (o & 31));
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

// ---------------------------------------------------------------------------------------
// Window mode (pruned phase).  Once thr > ub_table only documents that a posting list touches matter, and a
// query's lists touch ~3% of the documents: walking 1024-document super-ranges with a dense accumulator then
// spends its time on per-visit bookkeeping (~30 postings per visit).  Window mode covers up to 32k documents per
// visit and keeps the accumulator as an open-addressing hash table (document offset -> partial score) in the SAME
// 4 KB of shared memory: 512 slots, at most BM_HASH_MAX distinct documents (the window is halved and redone when
// it would overflow).  Terms are streamed in the same order and with the same fused multiply-add as in the dense
// accumulator, so every document gets bit-identical sums in both modes (and in any sharding).
// ---------------------------------------------------------------------------------------
constexpr int BM_HASH_SLOTS = 512;
constexpr int BM_HASH_MAX = 352;       // inserts stop being attempted beyond this: 352 + 4 * 32 < 512 keeps probing finite
constexpr int BM_WINDOW_MAX = 32768;   // documents per window (offsets must also stay well inside int32)

// Returns false when the table would overflow (nothing usable was changed: the caller restores the cursors).
template <int U, bool MARK_ONLY = false, bool IMP = false>   // IMP: post_tf is really post_imp (float): baked impacts   // IMP: post_tf is really post_imp (float): baked impacts
__device__ __forceinline__ bool stream_term_hash(const int32_t* __restrict__ post_doc,
                                                 const uint16_t* __restrict__ post_tf, int64_t& pos, const int64_t end,
                                                 const int d0, const int d1, const float weight, int* keys, float* vals,
                                                 unsigned* flags, const bool essential,
                                                 const float* __restrict__ norm, const int lane, int& next_doc,
                                                 int& n_keys, unsigned* mflags = nullptr) {
  static_assert(U * 32 + BM_HASH_MAX < BM_HASH_SLOTS, "a pass must fit behind the fill limit");
  while (true) {
    if (n_keys > BM_HASH_MAX) return false;
    int doc[U];
    unsigned tf[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t idx = pos + u * 32 + lane;
      const bool in = idx < end;
      doc[u] = in ? __ldg(post_doc + idx) : INT_MAX;
      if (IMP) tf[u] = (in && !MARK_ONLY) ? __ldg(reinterpret_cast<const unsigned*>(post_tf) + idx) : 0u;
      else tf[u] = (in && !MARK_ONLY) ? static_cast<unsigned>(__ldg(post_tf + idx)) : 0u;
    }
    const int64_t peek_idx = pos + U * 32;
    int peek = INT_MAX;
    if (lane == 0 && peek_idx < end) peek = __ldg(post_doc + peek_idx);
    int taken = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool take = doc[u] < d1;
      bool fresh = false;
      if (take) {
        float x = 0.0f;
        if (!MARK_ONLY) {
          if (IMP) {
            x = __uint_as_float(tf[u]);   // the bits the expression below produced when the impacts were baked
          } else {
            const float f = static_cast<float>(tf[u]);
            x = f * fast_rcp(f + __ldg(norm + doc[u]));
          }
        }
        const int key = doc[u] - d0;
        unsigned slot = (static_cast<unsigned>(key) * 0x9E3779B1u) >> 23;   // 9 bits
        int prev;
        while (true) {
          prev = atomicCAS(keys + slot, -1, key);
          if (prev == -1 || prev == key) break;
          slot = (slot + 1) & (BM_HASH_SLOTS - 1);
        }
        fresh = prev == -1;
        // documents of one list are distinct, so no other lane works on this slot's value right now
        if (MARK_ONLY) {
          if (fresh) vals[slot] = 0.0f;
          atomicOr(mflags + (slot >> 5), 1u << (slot & 31));   // scored exactly whatever its list part is
        } else {
          vals[slot] = fmaf(weight, x, fresh ? 0.0f : vals[slot]);
        }
        if (essential) atomicOr(flags + (slot >> 5), 1u << (slot & 31));
      }
      taken += __popc(__ballot_sync(0xffffffffu, take));
      n_keys += __popc(__ballot_sync(0xffffffffu, fresh));
    }
    pos += taken;
    if (taken < U * 32) {
      int cand = INT_MAX;
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (u == (taken >> 5)) cand = doc[u];
      next_doc = __shfl_sync(0xffffffffu, cand, taken & 31);
      return true;
    }
    peek = __shfl_sync(0xffffffffu, peek, 0);
    if (peek >= d1) {
      next_doc = peek;
      return true;
    }
    __syncwarp();
  }
}


// post_imp[i] = tf / (tf + norm[doc]) with the expression of stream_term / stream_term_hash (one MUFU.RCP, one
// multiply): the window phase adds weight * post_imp[i] and gets the very bits the tf + norm path computes, without
// the gather of norm[doc] behind every load of postings (a dependent round trip) and the convert / add / reciprocal.
__global__ void __launch_bounds__(256) posting_impacts_kernel(const int32_t* __restrict__ post_doc,
                                                              const uint16_t* __restrict__ post_tf,
                                                              const float* __restrict__ norm, int64_t nnz,
                                                              float* __restrict__ post_imp) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nnz;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float f = static_cast<float>(__ldg(post_tf + i));
    post_imp[i] = f * fast_rcp(f + __ldg(norm + __ldg(post_doc + i)));
  }
}

// per-warp shared-memory region: [accumulator / hash window 4 KB][touched bits 128 B][per-term arrays 49 B x max_terms]
constexpr int BM_REGION_BITS = sizeof(float) * BM_SUPER_DOCS;
constexpr int BM_REGION_TERMS = BM_REGION_BITS + sizeof(unsigned) * (BM_SUPER_DOCS / 32);
__host__ __device__ constexpr size_t bm25_warp_region_bytes(int max_terms) {
  return (static_cast<size_t>(BM_REGION_TERMS) + 49u * max_terms + 15u) & ~static_cast<size_t>(15);
}

template <bool DENSE_OUT, bool IMP = false>   // IMP: the window phase reads baked impacts (Bm25Args::post_imp)
__global__ void __launch_bounds__(BM_THREADS, DENSE_OUT ? 4 : BM_MIN_BLOCKS) bm25_kernel(const Bm25Args a) {   // get_scores streams: occupancy first
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // read once: a volatile asm is not re-executed, so the thread index (and what hangs on it) stays in registers
  // instead of being re-derived from S2R at every use under the register cap
  unsigned tid_u;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_u));
  const int tid = static_cast<int>(tid_u), warp = tid >> 5, lane = tid & 31;
  const int mt = a.max_terms;
  BM_PROF_T(prof_t0);
  // Carve shared memory, warp-major: everything a warp touches in its main loop lives in ONE region behind one base
  // pointer, at offsets that are multiples of max_terms (a kernel parameter, i.e. a constant-bank operand), so an
  // address costs one multiply-add from the base instead of a chain that starts at threadIdx (under the 80-register
  // budget the compiler re-materialised those chains - S2R included - all over the window loop: 7 % of the
  // instructions).  Only the candidate keys stay array-major: the fold at the end of the block reads all warps' lists.
  unsigned char* const wb = smem_raw + static_cast<size_t>(warp) * bm25_warp_region_bytes(mt);
  float* const sacc = reinterpret_cast<float*>(wb);                                  // posting-list contributions of a super-range
  unsigned* const s_bits = reinterpret_cast<unsigned*>(wb + BM_REGION_BITS);         // documents touched by a posting
  unsigned* const touched = DENSE_OUT ? nullptr : s_bits;
  int64_t* const s_pos = reinterpret_cast<int64_t*>(wb + BM_REGION_TERMS);           // cursor of each sparse term
  int64_t* const s_end = reinterpret_cast<int64_t*>(wb + BM_REGION_TERMS + 8 * mt);
  const uint8_t** const s_drow = reinterpret_cast<const uint8_t**>(wb + BM_REGION_TERMS + 16 * mt);  // dense-table row of each dense term
  const __half** const s_irow = reinterpret_cast<const __half**>(wb + BM_REGION_TERMS + 24 * mt);    // its row of fp16 impact bounds
  float* const s_wgt = reinterpret_cast<float*>(wb + BM_REGION_TERMS + 32 * mt);     // sparse term weights
  float* const s_dwgt = reinterpret_cast<float*>(wb + BM_REGION_TERMS + 36 * mt);    // dense-table term weights
  int* const s_nxt = reinterpret_cast<int*>(wb + BM_REGION_TERMS + 40 * mt);         // next document of each sparse term
  int* const s_tmp = reinterpret_cast<int*>(wb + BM_REGION_TERMS + 44 * mt);         // term ids while the cursors are placed
  unsigned char* const s_dense = wb + BM_REGION_TERMS + 48 * mt;                     // chunks per pass class of each sparse term
  unsigned char* sp = smem_raw + static_cast<size_t>(BM_WARPS) * bm25_warp_region_bytes(mt);
  uint64_t* s_keys = reinterpret_cast<uint64_t*>(sp) + warp * a.capacity;  // this warp's candidate keys
  if (!DENSE_OUT) sp += sizeof(uint64_t) * a.capacity * BM_WARPS;
  int* const s_dterms = reinterpret_cast<int*>(sp);           // table directory (block-wide, a.n_dense entries)

  const int q = blockIdx.x;
  const int stripe = a.stripe0 + static_cast<int>(blockIdx.y);
  const int64_t stripe_begin = static_cast<int64_t>(stripe) * a.stripe_docs;
  const int64_t stripe_end = min(a.n_docs, stripe_begin + a.stripe_docs);
  const int64_t sub_docs = a.stripe_docs / BM_WARPS;
  const int64_t w_begin = min(stripe_end, stripe_begin + warp * sub_docs);
  const int64_t w_end = min(stripe_end, w_begin + sub_docs);
  const int n_super = static_cast<int>(sub_docs / BM_SUPER_DOCS);

  for (int i = tid; i < a.n_dense; i += BM_THREADS) s_dterms[i] = a.dense_terms[i];
  WarpTopK tk;
  if (!DENSE_OUT) {
    tk.init(s_keys, a.k, a.capacity, positive_floor_key(), lane);
    if (a.seed_thr != nullptr) tk.raise(a.seed_thr[q]);
  }
  const float seed_dbg = (!DENSE_OUT && a.seed_thr != nullptr) ? a.seed_thr[q] : 0.0f;
  __syncthreads();

  // ---- split the query's terms: rows of the dense tf table vs posting lists -------------------
  const int qb = a.q_off[q];
  const int nt = min(a.q_off[q + 1] - qb, mt);
  int nd = 0, ns = 0;  // dense / sparse term counts (warp-uniform)
  float ub_lane = 0.0f, ubw_lane = 0.0f;  // this lane's share of the table-term bounds (tight / weights only)
  const bool use_cap = !DENSE_OUT && a.dense_cap != nullptr;
  for (int base = 0; base < nt; base += 32) {
    const int ti = base + lane;
    int t = -1, slot = -1;
    float w = 0.0f;
    if (ti < nt) {
      t = a.q_terms[qb + ti];
      if (t >= 0 && t < a.vocab) {
        w = a.idf[t] * a.k1p1;
        int lo_e = 0, hi_e = a.n_dense;  // lower_bound in the sorted table directory
        while (lo_e < hi_e) {
          const int mid = (lo_e + hi_e) >> 1;
          if (s_dterms[mid] < t) lo_e = mid + 1; else hi_e = mid;
        }
        if (lo_e < a.n_dense && s_dterms[lo_e] == t) slot = lo_e;
      }
    }
    const bool live = w != 0.0f;
    const unsigned md = __ballot_sync(0xffffffffu, live && slot >= 0);
    const unsigned ms = __ballot_sync(0xffffffffu, live && slot < 0);
    if (live && slot >= 0) {
      const int o = nd + __popc(md & ((1u << lane) - 1));
      s_drow[o] = a.dense_tf + static_cast<int64_t>(slot) * a.dense_stride;
      s_irow[o] = a.dense_imp != nullptr ? a.dense_imp + static_cast<int64_t>(slot) * a.dense_stride : nullptr;
      s_dwgt[o] = w;
      s_nxt[o] = slot;   // table row, until the marker lists have been queued below (s_nxt is rewritten afterwards)
      const float wp = fmaxf(w, 0.0f);
      ubw_lane += wp;
      // with an impact cap the row contributes at most w * cap to every document that is NOT on its marker list
      ub_lane += use_cap ? wp * a.dense_cap[slot] : (a.dense_maximp != nullptr ? wp * a.dense_maximp[slot] : wp);
    }
    if (live && slot < 0) s_tmp[ns + __popc(ms & ((1u << lane) - 1))] = t;
    nd += __popc(md);
    ns += __popc(ms);
  }
  __syncwarp();
  // Impact-capped table: the few documents of a row above its cap are listed separately; such a "marker list" is
  // queued like a posting list (negative id = -(row + 1)) that adds nothing and only marks its documents for exact
  // scoring.  Every other document gets at most w * cap from the row, which is what ub_table now promises.
  if (use_cap) {
    for (int i = lane; i < nd; i += 32) s_tmp[ns + i] = -(s_nxt[i] + 1);
    ns += nd;
    __syncwarp();
  }

  // A document without any posting-list term scores at most ub_table = sum of the positive weights
  // of the table terms (tf / (tf + norm) < 1).  Once the warp's k-th best score exceeds that bound,
  // only documents touched by a posting list can still qualify (the "essential lists" of MaxScore):
  // exact, and it turns ~10M evaluations per query into the ~80k documents its rarer terms occur in.
  // With the row maxima of the impact bounds (tf / (tf + norm) <= maximp < 1) the bound is 10-20% tighter.
  float ub_table = ub_lane, ub_weights = ubw_lane;
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) {
    ub_table += __shfl_xor_sync(0xffffffffu, ub_table, sh);
    ub_weights += __shfl_xor_sync(0xffffffffu, ub_weights, sh);
  }
  ub_table *= 1.000001f;  // the exact path multiplies by an approximate reciprocal (1 ulp)
  // fp16 bound pass: every HFMA2 rounds once, |error| <= 2^-11 * (partial sum <= ub_weights)
  const float approx_slack = static_cast<float>(nd) * ub_weights * (1.0f / 2048.0f) * 1.01f + 1e-3f;

  if (a.debug && tid == 0 && stripe == 0) {
    if (seed_dbg > 0.0f) atomicAdd(&g_bm25_dbg[4], 1ull);
    if (seed_dbg > ub_table) atomicAdd(&g_bm25_dbg[5], 1ull);
  }
  // ---- per-warp cursors: lower_bound(post_doc[term], w_begin) by a 32-ary search, 4 terms at a time
  int ntv = 0;  // sparse terms with a non-empty posting list (warp-uniform)
  float list_density = 0.0f;  // postings of all list terms per 1024 documents (warp-uniform)
  for (int g = 0; g < ns; g += BM_SEARCH) {
    int64_t lo[BM_SEARCH], hi[BM_SEARCH], te[BM_SEARCH], ts[BM_SEARCH];
    bool mk[BM_SEARCH];   // marker list (hi_doc) instead of a posting list (post_doc)
    float wg[BM_SEARCH];
#pragma unroll
    for (int j = 0; j < BM_SEARCH; ++j) {
      lo[j] = hi[j] = te[j] = ts[j] = 0;
      wg[j] = 0.0f;
      mk[j] = false;
      if (g + j < ns) {
        const int t = s_tmp[g + j];
        if (t >= 0) {
          lo[j] = ts[j] = a.term_off[t];
          hi[j] = te[j] = a.term_off[t + 1];
          if (hi[j] > lo[j]) wg[j] = a.idf[t] * a.k1p1;
        } else {                                   // marker list of table row -(t + 1)
          mk[j] = true;
          lo[j] = ts[j] = a.hi_off[-(t + 1)];
          hi[j] = te[j] = a.hi_off[-(t + 1) + 1];
          if (hi[j] > lo[j]) wg[j] = INFINITY;     // never non-essential; nothing is ever multiplied by it
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
          probe[j] = __ldg((mk[j] ? a.hi_doc : a.post_doc) + lo[j] + idx);
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
    __syncwarp();  // s_tmp[g..g+3] has been read by every lane before slots below are written
#pragma unroll
    for (int j = 0; j < BM_SEARCH; ++j) {
      if (wg[j] != 0.0f) {  // warp-uniform
        const int64_t idx = lo[j] + lane;
        const int32_t* const plist = mk[j] ? a.hi_doc : a.post_doc;
        const bool below = idx < hi[j] && __ldg(plist + idx) < target;
        const int64_t cur = lo[j] + __popc(__ballot_sync(0xffffffffu, below));
        int nx = INT_MAX;
        if (cur < te[j]) nx = __ldg(plist + cur);
        // expected postings of this term per super-range
        const double per_range = static_cast<double>(te[j] - ts[j]) * BM_SUPER_DOCS / static_cast<double>(a.n_docs);
        list_density += static_cast<float>(per_range);
        if (lane == 0) {
          s_pos[ntv] = cur;
          s_end[ntv] = te[j];
          s_wgt[ntv] = wg[j];
          s_nxt[ntv] = nx;
          // chunks per pass by density; class 4 = marker list (sparse by construction, mark-only)
          s_dense[ntv] = mk[j] ? 4 : (per_range < 24.0 ? 0 : (per_range < 56.0 ? 1 : (per_range < 120.0 ? 2 : 3)));
        }
        ++ntv;
      }
    }
  }
  __syncwarp();

  // ---- MaxScore over the posting-list terms.  Rank them by weight (ascending); psum = weight of a term plus
  // all lighter ones.  While  ub_table + psum(rank j) < thr  the j+1 lightest terms are NON-essential: a document
  // that only they (and table terms) touch cannot reach thr, so their postings still feed the accumulator but no
  // longer mark documents for scoring.  The heavy (rare) terms that stay essential have short lists: at 10M
  // passages this takes the documents scored per query from ~310k to a third of that.
  // Lane ti < ntv holds rank and psum of term ti; needs ntv <= 32 and no negative weight.
  int my_rank = 0;
  float my_psum = INFINITY;
  bool maxscore = !DENSE_OUT && ntv > 0 && ntv <= 32;
  if (maxscore) {
    const float wt = lane < ntv ? s_wgt[lane] : INFINITY;
    maxscore = !__any_sync(0xffffffffu, lane < ntv && !(wt > 0.0f));
    float ps = 0.0f;
    for (int j = 0; j < ntv; ++j) {
      const float wj = __shfl_sync(0xffffffffu, wt, j);
      const bool before = wj < wt || (wj == wt && j < lane);   // strict total order, duplicates by position
      my_rank += before ? 1 : 0;
      ps += (before || j == lane) ? wj : 0.0f;
    }
    if (lane < ntv) my_psum = ps;
  }

  // window mode (pruned phase): aim at ~256 postings per window
  bool window_ok = !DENSE_OUT && a.window_mode != 0 && ntv <= 32;
  int window = BM_SUPER_DOCS;
  {
    const float per_doc = fmaxf(list_density, 1e-3f) / BM_SUPER_DOCS;
    const int want = static_cast<int>(fminf(static_cast<float>(a.window_mode) / per_doc, static_cast<float>(BM_WINDOW_MAX)));
    window = max(BM_SUPER_DOCS, want / BM_SUPER_DOCS * BM_SUPER_DOCS);
  }

  // The admission threshold is shared between all blocks of a query through seed_thr[q] in L2: every block that
  // finishes publishes the k-th best score of its stripe (atomicMax below) - a proven lower bound of the query's
  // final k-th best - and every warp picks the current value up once per visit.  Blocks of later stripes therefore
  // start where the best earlier stripe ended instead of warming their own top-k up from the seed.  The value is
  // requested one visit ahead so its L2 latency hides behind the work of the visit.  It only prunes: results do
  // not depend on which blocks ran first.
  float pending_thr = 0.0f;
  const int j0 = lane * 8;  // the 8 documents of a range this lane owns
  BM_PROF_T(prof_t1);
  BM_PROF_ADD(0, blockIdx.x, prof_t1 - prof_t0);
#ifdef RAGB_BM25_PROFILE
  long long prof_win = 0, prof_stream = 0, prof_compact = 0, prof_score = 0, prof_visits = 0;
#endif
  for (int sup = 0; sup < n_super; ++sup) {
    const int64_t s0l = w_begin + static_cast<int64_t>(sup) * BM_SUPER_DOCS;
    const int s0 = static_cast<int>(s0l < w_end ? s0l : w_end);
    const int s1 = static_cast<int>(min(w_end, s0l + BM_SUPER_DOCS));
    if (s1 <= s0) break;  // warp-uniform; nothing below synchronises the block
    if (!DENSE_OUT) {
      tk.raise(pending_thr);
      pending_thr = __ldcg(a.seed_thr + q);
    }
    if (!DENSE_OUT && window_ok && tk.thr_score > ub_table) {
      // ================= window mode: the rest of this warp's documents (see stream_term_hash) =================
      int* const keys = reinterpret_cast<int*>(sacc);
      float* const vals = sacc + BM_HASH_SLOTS;
      unsigned* const flags = s_bits;                              // slot marked by an essential list
      unsigned* const mflags = s_bits + BM_HASH_SLOTS / 32;        // slot on a marker list: scored whatever its list part
      int64_t d0l = s0l;
      bool fell_back = false;
      BM_PROF_T(prof_w0);
      while (d0l < w_end) {
        const int d0 = static_cast<int>(d0l);
        const int d1 = static_cast<int>(min(w_end, d0l + window));
        tk.raise(pending_thr);
        pending_thr = __ldcg(a.seed_thr + q);
        unsigned act = __ballot_sync(0xffffffffu, lane < ntv && s_nxt[lane] < d1);
        if (act == 0u) {   // no posting of any list in this window: nothing can reach thr
          if (a.debug && lane == 0) atomicAdd(&g_bm25_dbg[1], 1ull);
          d0l += window;
          continue;
        }
        const int64_t save_pos = lane < ntv ? s_pos[lane] : 0;
        const int save_nxt = lane < ntv ? s_nxt[lane] : 0;
#pragma unroll
        for (int i = 0; i < BM_HASH_SLOTS / 128; ++i)
          reinterpret_cast<int4*>(keys)[i * 32 + lane] = make_int4(-1, -1, -1, -1);
        s_bits[lane] = 0u;   // flags (words 0-15) and marker flags (words 16-31)
        __syncwarp();
        int n_noness = 0;
        if (maxscore) {
          const float thr_safe = tk.thr_score - 2e-5f * fabsf(tk.thr_score);
          n_noness = __popc(__ballot_sync(0xffffffffu, lane < ntv && ub_table + my_psum < thr_safe));
        }
        int n_keys = 0;
        bool ok = true;
        BM_PROF_T(prof_s0);
        while (act != 0u && ok) {
          const int ti = __ffs(act) - 1;
          act &= act - 1u;
          int64_t pos = s_pos[ti];
          const int64_t end = s_end[ti];
          const float w = s_wgt[ti];
          const bool essential = !(n_noness > 0 && __shfl_sync(0xffffffffu, my_rank, ti) < n_noness);
          int next_doc = INT_MAX;
          const uint16_t* const ptf = IMP ? reinterpret_cast<const uint16_t*>(a.post_imp) : a.post_tf;
          // One instantiation for every posting list (two 32-posting chunks per pass) and one for the marker lists: the
          // density-specific variants (1 / 2 / 4 chunks per pass) measured 7 % SLOWER - the kernel is ~125 KB of code,
          // warps of a block sit in different phases, and ncu charged 11 % of the stalls to instruction fetch.
          if (s_dense[ti] == 4)   // warp-uniform
            ok = stream_term_hash<1, true>(a.hi_doc, nullptr, pos, end, d0, d1, 0.0f, keys, vals, flags, true, a.norm, lane, next_doc, n_keys, mflags);
          else
            ok = stream_term_hash<2, false, IMP>(a.post_doc, ptf, pos, end, d0, d1, w, keys, vals, flags, essential, a.norm, lane, next_doc, n_keys);
          if (ok && lane == 0) {
            s_pos[ti] = pos;
            s_nxt[ti] = next_doc;
          }
          __syncwarp();
        }
        if (!ok) {
          // too many distinct documents: put the cursors back and redo the window at half the size; a single
          // super-range that still overflows goes back to the dense accumulator for good
          if (lane < ntv) {
            s_pos[lane] = save_pos;
            s_nxt[lane] = save_nxt;
          }
          __syncwarp();
          if (window == BM_SUPER_DOCS) {
            fell_back = true;
            break;
          }
          window = max(BM_SUPER_DOCS, (window / 2) / BM_SUPER_DOCS * BM_SUPER_DOCS);
          continue;
        }
        // compact the marked slots to the front of the table (position <= slot, so in place) ...  A marked
        // document whose list part plus everything the table terms could add stays below thr is dropped here,
        // before any of its table bytes is gathered (most marked documents match a single light list term).
        const float need = (tk.thr_score - 2e-5f * fabsf(tk.thr_score)) - ub_table;
        BM_PROF_T(prof_s1);
        int n_valid = 0;
        for (int base = 0; base < BM_HASH_SLOTS; base += 32) {
          const int key = keys[base + lane];
          const float val = vals[base + lane];
          const bool valid = key != -1 && ((flags[base >> 5] >> lane) & 1u) != 0u &&
                             (val >= need || ((mflags[base >> 5] >> lane) & 1u) != 0u);
          const unsigned m = __ballot_sync(0xffffffffu, valid);
          if (valid) {
            const int j = n_valid + __popc(m & ((1u << lane) - 1u));
            keys[j] = key;
            vals[j] = val;
          }
          n_valid += __popc(m);
          __syncwarp();
        }
        if (a.debug && lane == 0) {
          atomicAdd(&g_bm25_dbg[2], 1ull);
          atomicAdd(&g_bm25_dbg[3], static_cast<unsigned long long>(n_valid));
        }
        // ... and score them: table terms first, in query order, then the list part - as in the dense mode
        BM_PROF_T(prof_s2);
        for (int i0 = 0; i0 < n_valid; i0 += 32) {
          const bool valid = i0 + lane < n_valid;
          float total = 0.0f;
          int doc = d0;
          if (valid) {
            doc = d0 + keys[i0 + lane];
            const float nrm = __ldg(a.norm + doc);
            // the byte gathers are L2 / DRAM round trips and this phase is latency-bound: issue those of up to
            // eight table terms together, then run the fused multiply-adds in query order as everywhere else
            for (int i = 0; i < nd; i += 8) {
              unsigned tfb[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) tfb[u] = (i + u < nd) ? __ldg(s_drow[i + u] + doc) : 0u;
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                if (i + u < nd) {
                  const float f = __uint_as_float(0x4B000000u | tfb[u]) - 8388608.0f;
                  total = fmaf(s_dwgt[i + u], f * fast_rcp(f + nrm), total);
                }
              }
            }
            total += vals[i0 + lane];
          }
          tk.offer(valid && total >= tk.thr_score, make_key(total, static_cast<int32_t>(a.id_base + doc)), lane);
        }
        __syncwarp();
        d0l += window;
#ifdef RAGB_BM25_PROFILE
        prof_stream += prof_s1 - prof_s0;
        prof_compact += prof_s2 - prof_s1;
        prof_score += clock64() - prof_s2;
        prof_visits += 1;
#endif
      }
#ifdef RAGB_BM25_PROFILE
      prof_win += clock64() - prof_w0;
#endif
      if (!fell_back) break;   // this warp is done
      window_ok = false;
      sup = static_cast<int>((d0l - w_begin) / BM_SUPER_DOCS) - 1;   // resume the dense-accumulator loop at d0l
      continue;
    }
    // ---- posting-list terms: streamed once per super-range into shared memory, and only when
    //      some list actually reaches into it (one ballot decides)
    unsigned active = 0;
    bool have_sparse = false;
    if (ntv > 0) {
      if (ntv <= 32) {
        active = __ballot_sync(0xffffffffu, lane < ntv && s_nxt[lane] < s1);
        have_sparse = active != 0;
      } else {
        have_sparse = true;
      }
    }
    // non-essential terms of this super-range (the threshold only rises, so the set only grows)
    // (n_noness > 0 implies thr > ub_table, i.e. the marked-documents path below)
    int n_noness = 0;
    if (maxscore) {
      const float thr_safe = tk.thr_score - 2e-5f * fabsf(tk.thr_score);
      n_noness = __popc(__ballot_sync(0xffffffffu, lane < ntv && ub_table + my_psum < thr_safe));
    }
    if (have_sparse) {
#pragma unroll
      for (int i = 0; i < BM_SUPER_DOCS / 128; ++i)
        *reinterpret_cast<float4*>(sacc + (i * 32 + lane) * 4) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (!DENSE_OUT) s_bits[lane] = 0u;
      __syncwarp();
      for (int ti = 0; ti < ntv; ++ti) {
        if (ntv <= 32) {
          if (active == 0) break;
          ti = __ffs(active) - 1;
          active &= active - 1;
        } else if (s_nxt[ti] >= s1) {
          continue;
        }
        int64_t pos = s_pos[ti];
        const int64_t end = s_end[ti];
        const float w = s_wgt[ti];
        int next_doc;
        // a non-essential term adds to the accumulator but marks nothing
        unsigned* const mark = (n_noness > 0 && __shfl_sync(0xffffffffu, my_rank, ti & 31) < n_noness) ? nullptr : touched;
        switch (s_dense[ti]) {  // chunks per pass sized to the term's density (warp-uniform)
          case 4: stream_term<1, true>(a.hi_doc, nullptr, pos, end, s0, s1, 0.0f, sacc, a.norm, touched, lane, next_doc); break;
          default:
            if (DENSE_OUT) {   // get_scores streams every list: chunks per pass by density
              switch (s_dense[ti]) {
                case 0: stream_term<1>(a.post_doc, a.post_tf, pos, end, s0, s1, w, sacc, a.norm, mark, lane, next_doc); break;
                case 1: stream_term<2>(a.post_doc, a.post_tf, pos, end, s0, s1, w, sacc, a.norm, mark, lane, next_doc); break;
                case 2: stream_term<4>(a.post_doc, a.post_tf, pos, end, s0, s1, w, sacc, a.norm, mark, lane, next_doc); break;
                default: stream_term<8>(a.post_doc, a.post_tf, pos, end, s0, s1, w, sacc, a.norm, mark, lane, next_doc); break;
              }
            } else {           // top-k search: this phase is the minority, one variant keeps the kernel's code small
              stream_term<4>(a.post_doc, a.post_tf, pos, end, s0, s1, w, sacc, a.norm, mark, lane, next_doc);
            }
            break;
        }
        if (lane == 0) {
          s_pos[ti] = pos;
          s_nxt[ti] = next_doc;
        }
        __syncwarp();
      }
    }
    // ---- which documents of this super-range have to be scored exactly?
    //  * thr > ub_table: only those a posting list touched (the "essential lists" of MaxScore);
    //  * otherwise, once the threshold is a sizeable fraction of the table bound: an fp16 UPPER bound of
    //    the table score of every document (one 128-bit load and four HFMA2 per term per 8 documents,
    //    12x fewer instructions than the exact path) marks the few that may still reach thr; the exact
    //    arithmetic then runs for the marked and the touched documents only.  Too many marks (the list
    //    is still warming up): fall through to the exact table path for the whole super-range.
    bool use_bits = false, bits_valid = have_sparse;
    if (!DENSE_OUT) {
      if (tk.thr_score > ub_table) {
        use_bits = true;
      } else if (a.dense_imp != nullptr && nd > 0 && tk.thr_score - approx_slack > 0.5f * ub_table) {   // n_noness == 0 here
        if (!have_sparse) {
          s_bits[lane] = 0u;
          __syncwarp();
        }
        const __half2 thr2 = __half2half2(__float2half_rd(tk.thr_score - approx_slack));
        // Terms outside, the (up to) four 256-document ranges of the super-range inside: the four 128-bit loads of a
        // term are independent, so a lane keeps four requests in flight instead of one (the pass streams the fp16
        // rows from DRAM and was latency-bound with a single load per lane outstanding).
        const int n_rng = (s1 - s0 + BM_RANGE - 1) / BM_RANGE;
        __half2 acc[BM_SUPER][4];
#pragma unroll
        for (int r = 0; r < BM_SUPER; ++r)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[r][j] = __float2half2_rn(0.0f);
        const int64_t off0 = static_cast<int64_t>(s0) + j0;
#pragma unroll 1
        for (int i = 0; i < nd; ++i) {
          const __half2 w2 = __half2half2(__float2half_ru(fmaxf(s_dwgt[i], 0.0f)));
          const __half* row = s_irow[i] + off0;
          uint4 v[BM_SUPER];
#pragma unroll
          for (int r = 0; r < BM_SUPER; ++r)
            v[r] = r < n_rng ? __ldg(reinterpret_cast<const uint4*>(row + r * BM_RANGE)) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int r = 0; r < BM_SUPER; ++r) {
            acc[r][0] = __hfma2(w2, *reinterpret_cast<const __half2*>(&v[r].x), acc[r][0]);
            acc[r][1] = __hfma2(w2, *reinterpret_cast<const __half2*>(&v[r].y), acc[r][1]);
            acc[r][2] = __hfma2(w2, *reinterpret_cast<const __half2*>(&v[r].z), acc[r][2]);
            acc[r][3] = __hfma2(w2, *reinterpret_cast<const __half2*>(&v[r].w), acc[r][3]);
          }
        }
#pragma unroll
        for (int r = 0; r < BM_SUPER; ++r) {
          if (r < n_rng) {
            unsigned m8 = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const unsigned mk = __hge2_mask(acc[r][j], thr2);   // 0xFFFF per half that passes
              m8 |= ((mk & 1u) | ((mk >> 15) & 2u)) << (2 * j);
            }
            const int left = s1 - (s0 + r * BM_RANGE) - j0;     // documents of this lane that exist
            if (left < 8) m8 &= left <= 0 ? 0u : ((1u << left) - 1u);
            if (m8) atomicOr(s_bits + r * (BM_RANGE / 32) + (lane >> 2), m8 << ((lane & 3) * 8));
          }
        }
        __syncwarp();
        int marks = __popc(s_bits[lane]);
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) marks += __shfl_xor_sync(0xffffffffu, marks, sh);
        if (marks <= BM_APPROX_MAX) {
          use_bits = true;
          bits_valid = true;
        }
        if (a.debug && lane == 0) atomicAdd(&g_bm25_dbg[use_bits ? 6 : 7], 1ull);
      }
    }
    if (a.debug && lane == 0) atomicAdd(&g_bm25_dbg[use_bits ? (have_sparse ? 2 : 1) : 0], 1ull);
    if (use_bits) {
      // ---- score only the marked documents (one bit each)
      if (bits_valid) {
        unsigned word = s_bits[lane];
        if (a.debug) atomicAdd(&g_bm25_dbg[3], static_cast<unsigned long long>(__popc(word)));
        while (__any_sync(0xffffffffu, word != 0u)) {
          const bool valid = word != 0u;
          float total = 0.0f;
          int doc = s0;
          if (valid) {
            const int o = lane * 32 + __ffs(word) - 1;
            word &= word - 1;
            doc = s0 + o;
            const float nrm = __ldg(a.norm + doc);
            for (int i = 0; i < nd; ++i) {  // same arithmetic, same order as the full path below
              const unsigned tfb = __ldg(s_drow[i] + doc);
              const float f = __uint_as_float(0x4B000000u | tfb) - 8388608.0f;
              total = fmaf(s_dwgt[i], f * fast_rcp(f + nrm), total);
            }
            if (have_sparse) total += sacc[o];
          }
          tk.offer(valid && total >= tk.thr_score, make_key(total, static_cast<int32_t>(a.id_base + doc)), lane);
        }
      }
      __syncwarp();
      continue;
    }
#pragma unroll 1
    for (int r = 0; r < BM_SUPER; ++r) {
      const int d0 = s0 + r * BM_RANGE;
      if (d0 >= s1) break;
      const int d1 = min(s1, d0 + BM_RANGE);
      const int cnt = d1 - d0;
      float ac[8], nr[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) ac[j] = 0.0f;
      if (cnt == BM_RANGE) {
        const float4 n0 = __ldg(reinterpret_cast<const float4*>(a.norm + d0 + j0));
        const float4 n1 = __ldg(reinterpret_cast<const float4*>(a.norm + d0 + j0 + 4));
        nr[0] = n0.x; nr[1] = n0.y; nr[2] = n0.z; nr[3] = n0.w;
        nr[4] = n1.x; nr[5] = n1.y; nr[6] = n1.z; nr[7] = n1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) nr[j] = (j0 + j) < cnt ? __ldg(a.norm + d0 + j0 + j) : 1.0f;
      }
      // ---- dense-table terms: 8 tf bytes per lane per term, accumulators stay in registers.
      // One term at a time with the next term's bytes already in flight.
      if (nd > 0) {
        const int64_t off = static_cast<int64_t>(d0) + j0;
        uint2 cur = __ldg(reinterpret_cast<const uint2*>(s_drow[0] + off));
        for (int i = 0; i < nd; ++i) {
          const float w = s_dwgt[i];
          uint2 nxt = make_uint2(0u, 0u);
          if (i + 1 < nd) nxt = __ldg(reinterpret_cast<const uint2*>(s_drow[i + 1] + off));
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const unsigned word = j < 4 ? cur.x : cur.y;
            // one PRMT builds 0x4B0000tt = 8388608.0f + tf; subtracting 2^23 gives tf exactly
            const float f = __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540u | (j & 3))) - 8388608.0f;
            ac[j] = fmaf(w, f * fast_rcp(f + nr[j]), ac[j]);
          }
          cur = nxt;
        }
      }
      if (have_sparse) {  // score = (table terms in query order) + (list terms in query order)
        const float4 r0 = *reinterpret_cast<const float4*>(sacc + r * BM_RANGE + j0);
        const float4 r1 = *reinterpret_cast<const float4*>(sacc + r * BM_RANGE + j0 + 4);
        ac[0] += r0.x; ac[1] += r0.y; ac[2] += r0.z; ac[3] += r0.w;
        ac[4] += r1.x; ac[5] += r1.y; ac[6] += r1.z; ac[7] += r1.w;
      }
      if (DENSE_OUT) {
        // d0 is a multiple of BM_RANGE = 256 = the tile width, so a warp range is one row segment of one tile
        float* dst = a.out_tiled
                         ? a.out_scores + (static_cast<int64_t>(d0 / BM_RANGE) * a.out_ld + q) * BM_RANGE + j0
                         : a.out_scores + static_cast<int64_t>(q) * a.out_ld + d0 + j0;
        if (cnt == BM_RANGE && (a.out_tiled || (a.out_ld & 3) == 0) && (reinterpret_cast<uintptr_t>(a.out_scores) & 15) == 0) {
          // d0 and j0 are multiples of 8: two 128-bit stores per lane, 32 contiguous bytes
          reinterpret_cast<float4*>(dst)[0] = make_float4(ac[0], ac[1], ac[2], ac[3]);
          reinterpret_cast<float4*>(dst)[1] = make_float4(ac[4], ac[5], ac[6], ac[7]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if ((j0 + j) < cnt) dst[j] = ac[j];
        }
      } else {
        // ---- warp-private selection: no barrier; in steady state 8 compares and one vote per range
        bool hot = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) hot |= (ac[j] >= tk.thr_score) && (j0 + j) < cnt;
        if (__any_sync(0xffffffffu, hot)) {
          const int32_t gid0 = static_cast<int32_t>(a.id_base + d0) + j0;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            tk.offer((ac[j] >= tk.thr_score) && (j0 + j) < cnt, make_key(ac[j], gid0 + j), lane);
        }
      }
    }
    __syncwarp();  // the next super-range clears sacc
  }
  BM_PROF_T(prof_t2);
#ifdef RAGB_BM25_PROFILE
  BM_PROF_ADD(1, blockIdx.x, prof_win);
  BM_PROF_ADD(4, blockIdx.x, prof_stream);
  BM_PROF_ADD(5, blockIdx.x, prof_compact);
  BM_PROF_ADD(6, blockIdx.x, prof_score);
  BM_PROF_ADD(2, blockIdx.x, prof_t2 - prof_t1 - prof_win);
#endif
  if (!DENSE_OUT) {
    // fold the 8 warp lists of this block into one (8k <= 2048 keys, one bitonic sort), so the
    // cross-stripe merge sees one list per block
    tk.flush(lane);
    __syncthreads();
#ifdef RAGB_BM25_PROFILE
    BM_PROF_ADD(7, blockIdx.x, clock64() - prof_t2);   // of the fold: the wait for the block's slowest warp
#endif
    uint64_t* all_keys = s_keys - warp * a.capacity;  // [BM_WARPS][capacity], each sorted in its first k slots
    int n = 2;
    while (n < BM_WARPS * a.k) n <<= 1;               // <= BM_WARPS * capacity because k <= capacity / 2
    uint64_t mine[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int e = r * BM_THREADS + tid;
      mine[r] = e < BM_WARPS * a.k ? all_keys[(e / a.k) * a.capacity + (e % a.k)] : 0ull;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int e = r * BM_THREADS + tid;
      if (e < n) all_keys[e] = mine[r];
    }
    __syncthreads();
    bitonic_sort_desc<BM_THREADS>(all_keys, n);
    uint64_t* dst = a.part_keys + (static_cast<int64_t>(q) * a.n_stripes + stripe) * a.k;
    for (int i = tid; i < a.k; i += BM_THREADS) dst[i] = all_keys[i];
    // publish this stripe's k-th best score (positive floats order like their bit patterns)
    if (tid == 0 && all_keys[a.k - 1] != 0ull) {
      const float kth = key_score(all_keys[a.k - 1]);
      if (kth > 0.0f) atomicMax(reinterpret_cast<int*>(a.seed_thr + q), __float_as_int(kth));
    }
  }
#ifdef RAGB_BM25_PROFILE
  BM_PROF_ADD(3, blockIdx.x, clock64() - prof_t2);
#endif
}

// ---------------------------------------------------------------------------------------
// Threshold seeding.  For one posting-list term e of the query, every document d in its list has
//   score(q, d) >= w_e * imp_e(d) + sum over table terms (exact)        (other list terms add >= 0)
// so the k-th largest of these lower bounds over (a prefix of) the list is a proven lower bound of
// the query's k-th best score: at least k distinct documents reach it.  The best bound over the
// query's list terms lets the main kernel start in pruned mode.  Skipped (bound 0) when a list
// term has a negative weight or no list is k documents long.
// ---------------------------------------------------------------------------------------
constexpr int SEED_THREADS = 256;
constexpr int SEED_DOCS = 1024;

__global__ void __launch_bounds__(SEED_THREADS) bm25_seed_kernel(const Bm25Args a) {
  __shared__ int s_dterms[BM_MAX_DENSE];
  __shared__ const uint8_t* s_row[BM_MAX_TERMS];
  __shared__ float s_roww[BM_MAX_TERMS];
  __shared__ int64_t s_lo[BM_MAX_TERMS];
  __shared__ int s_len[BM_MAX_TERMS];
  __shared__ float s_lw[BM_MAX_TERMS];
  __shared__ uint32_t s_lb[SEED_DOCS];
  __shared__ int s_nrow, s_nlist, s_negative;
  const int q = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < a.n_dense; i += SEED_THREADS) s_dterms[i] = a.dense_terms[i];
  if (tid == 0) s_nrow = s_nlist = s_negative = 0;
  __syncthreads();
  const int qb = a.q_off[q];
  const int nt = min(a.q_off[q + 1] - qb, a.max_terms);
  if (tid < nt) {
    const int t = a.q_terms[qb + tid];
    if (t >= 0 && t < a.vocab) {
      const float w = a.idf[t] * a.k1p1;
      int lo_e = 0, hi_e = a.n_dense;
      while (lo_e < hi_e) {
        const int mid = (lo_e + hi_e) >> 1;
        if (s_dterms[mid] < t) lo_e = mid + 1; else hi_e = mid;
      }
      if (w != 0.0f) {
        if (lo_e < a.n_dense && s_dterms[lo_e] == t) {
          const int o = atomicAdd(&s_nrow, 1);
          s_row[o] = a.dense_tf + static_cast<int64_t>(lo_e) * a.dense_stride;
          s_roww[o] = w;
        } else {
          if (w < 0.0f) s_negative = 1;
          const int64_t lo = a.term_off[t], hi = a.term_off[t + 1];
          if (hi - lo >= a.k) {
            const int o = atomicAdd(&s_nlist, 1);
            s_lo[o] = lo;
            s_len[o] = static_cast<int>(min(hi - lo, static_cast<int64_t>(SEED_DOCS)));
            s_lw[o] = w;
          }
        }
      }
    }
  }
  __syncthreads();
  float best = 0.0f;
  if (!s_negative) {
    const int nrow = s_nrow, nlist = s_nlist;
    for (int e = 0; e < nlist; ++e) {
      const int64_t lo = s_lo[e];
      const int n = s_len[e];
      const float w = s_lw[e];
      for (int i = tid; i < SEED_DOCS; i += SEED_THREADS) {
        uint32_t key = 0u;
        if (i < n) {
          const int doc = __ldg(a.post_doc + lo + i);
          const float nrm = __ldg(a.norm + doc);
          const float f = static_cast<float>(__ldg(a.post_tf + lo + i));
          float lb = w * (f * fast_rcp(f + nrm));
          for (int r = 0; r < nrow; ++r) {
            const float g = static_cast<float>(__ldg(s_row[r] + doc));
            lb = fmaf(s_roww[r], g * fast_rcp(g + nrm), lb);
          }
          key = float_to_ordered(lb);
        }
        s_lb[i] = key;
      }
      __syncthreads();
      // descending bitonic sort of the 1024 ordered bounds
      for (int size = 2; size <= SEED_DOCS; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          for (int i = tid; i < SEED_DOCS / 2; i += SEED_THREADS) {
            const int lo_i = 2 * i - (i & (stride - 1)), hi_i = lo_i + stride;
            const bool desc = ((lo_i & size) == 0);
            const uint32_t x = s_lb[lo_i], y = s_lb[hi_i];
            if ((x < y) == desc) {
              s_lb[lo_i] = y;
              s_lb[hi_i] = x;
            }
          }
          __syncthreads();
        }
      }
      const uint32_t kth = s_lb[a.k - 1];
      if (kth != 0u) best = fmaxf(best, ordered_to_float(kth));
      __syncthreads();
    }
  }
  // a hair below the bound: the main kernel sums the same terms in another order
  if (tid == 0) a.seed_thr[q] = best > 0.0f ? best * (1.0f - 8e-6f) : 0.0f;
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

// ---------------------------------------------------------------------------------------
// Builders of the dense tf table and its fp16 impact bounds (used by SparseShard.finalize).
// ---------------------------------------------------------------------------------------
constexpr int TB_THREADS = 256;
constexpr int TB_SPLIT = 64;   // blocks per term: grid = (TB_SPLIT, n_terms)

// max_tf[i] = largest term frequency in the posting list of terms[i] (0 for an empty list); max_tf is zeroed by the caller
__global__ void __launch_bounds__(TB_THREADS) term_max_tf_kernel(const int64_t* __restrict__ term_off,
                                                                 const uint16_t* __restrict__ post_tf,
                                                                 const int32_t* __restrict__ terms, int32_t* __restrict__ max_tf) {
  const int t = terms[blockIdx.y];
  const int64_t lo = term_off[t], hi = term_off[t + 1];
  int best = 0;
  for (int64_t i = lo + static_cast<int64_t>(blockIdx.x) * TB_THREADS + threadIdx.x; i < hi; i += static_cast<int64_t>(TB_SPLIT) * TB_THREADS)
    best = max(best, static_cast<int>(__ldg(post_tf + i)));
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, sh));
  if ((threadIdx.x & 31) == 0 && best > 0) atomicMax(max_tf + blockIdx.y, best);
}

// table[i][doc] = tf for every posting of terms[i]; the table was zeroed by the caller (0 = term absent)
__global__ void __launch_bounds__(TB_THREADS) dense_table_fill_kernel(const int64_t* __restrict__ term_off,
                                                                      const int32_t* __restrict__ post_doc,
                                                                      const uint16_t* __restrict__ post_tf,
                                                                      const int32_t* __restrict__ terms, uint8_t* __restrict__ table,
                                                                      int64_t stride) {
  const int t = terms[blockIdx.y];
  const int64_t lo = term_off[t], hi = term_off[t + 1];
  uint8_t* row = table + static_cast<int64_t>(blockIdx.y) * stride;
  for (int64_t i = lo + static_cast<int64_t>(blockIdx.x) * TB_THREADS + threadIdx.x; i < hi; i += static_cast<int64_t>(TB_SPLIT) * TB_THREADS)
    row[__ldg(post_doc + i)] = static_cast<uint8_t>(__ldg(post_tf + i));
}

// imp[r][d] = smallest fp16 >= tf / (tf + norm[d]) * (1 + 2e-6)  (the scoring kernels multiply by an approximate
// reciprocal: the factor keeps the bound above what they compute); max_imp[r] = row maximum (zeroed by the caller;
// non-negative floats order like their bit patterns, so one atomicMax on the bits per block suffices).
__global__ void __launch_bounds__(TB_THREADS) impact_bounds_kernel(const uint8_t* __restrict__ table, const float* __restrict__ norm,
                                                                   int64_t n_docs, int64_t stride, __half* __restrict__ imp,
                                                                   float* __restrict__ max_imp) {
  const int r = blockIdx.y;
  const uint8_t* row = table + static_cast<int64_t>(r) * stride;
  __half* out = imp + static_cast<int64_t>(r) * stride;
  float best = 0.0f;
  // 8 documents per thread per step: one 64-bit load of tf bytes, one 128-bit store of fp16 bounds (stride % 256 == 0)
  for (int64_t d0 = (static_cast<int64_t>(blockIdx.x) * TB_THREADS + threadIdx.x) * 8; d0 < stride;
       d0 += static_cast<int64_t>(gridDim.x) * TB_THREADS * 8) {
    const uint2 bytes = __ldg(reinterpret_cast<const uint2*>(row + d0));
    __align__(16) __half h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float f = static_cast<float>(((j < 4 ? bytes.x : bytes.y) >> (8 * (j & 3))) & 0xffu);
      const float nrm = (d0 + j) < n_docs ? __ldg(norm + d0 + j) : 1.0f;
      const float exact = __fmul_rn(__fdiv_rn(f, __fadd_rn(f, nrm)), 1.000002f);
      h[j] = __float2half_ru(exact);
      best = fmaxf(best, __half2float(h[j]));
    }
    *reinterpret_cast<uint4*>(out + d0) = *reinterpret_cast<const uint4*>(h);
  }
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, sh));
  if ((threadIdx.x & 31) == 0 && best > 0.0f) atomicMax(reinterpret_cast<unsigned*>(max_imp + r), __float_as_uint(best));
}

static int bm25_stripes(int n_queries, int64_t n_docs, int64_t* stripe_docs_out) {
  const int64_t unit = static_cast<int64_t>(BM_WARPS) * BM_SUPER_DOCS;
  static const int64_t per_sm = [] {
    const char* e = getenv("RAGB_BM25_BLOCKS_PER_SM");  // tuning aid: blocks aimed at per SM (more = finer balance, more set-up)
    return e ? static_cast<int64_t>(atoi(e)) : 64;   // measured at 10M x 1024: 12 -> 18.9 ms, 48 -> 14.5, 72 -> 14.0, 192 -> 14.3
                                                       // (queries differ a lot in cost once pruning works: finer stripes balance the SMs)
  }();
  const int64_t target_blocks = 148 * per_sm;
  int64_t stripes = ceil_div64(target_blocks, n_queries);
  // ... but no stripe below 256k documents (32k per warp): every warp places its cursors and splits the query's terms
  // anew, and on a small shard (1.25M documents at 8 GPUs) ten stripes made that set-up a third of the kernel
  // (measured at 8 GPUs: 10 stripes 2.17 ms, 4 stripes 2.04 ms, 2 stripes 2.25 ms)
  int64_t max_stripes = ceil_div64(n_docs, unit);
  const int64_t by_size = n_docs / 262144 > 0 ? n_docs / 262144 : 1;
  if (max_stripes > by_size && n_queries >= 64) max_stripes = by_size;
  if (stripes > max_stripes) stripes = max_stripes;
  if (stripes < 1) stripes = 1;
  int64_t stripe_docs = ceil_div64(ceil_div64(n_docs, stripes), unit) * unit;
  stripes = ceil_div64(n_docs, stripe_docs);
  *stripe_docs_out = stripe_docs;
  return static_cast<int>(stripes);
}

static size_t bm25_smem_bytes(int max_terms, int capacity, bool dense_out, int n_dense) {
  size_t b = BM_WARPS * bm25_warp_region_bytes(max_terms);
  if (!dense_out) b += sizeof(uint64_t) * capacity * BM_WARPS;
  b += sizeof(int) * n_dense;                                    // table directory
  return (b + 15) & ~static_cast<size_t>(15);
}

static int bm25_common_checks(const char* who, const int64_t* term_off, const int32_t* post_doc,
                              const uint16_t* post_tf, const float* norm, const float* idf, int64_t vocab,
                              const int32_t* q_terms, const int32_t* q_off, int32_t n_queries, int64_t n_docs,
                              int32_t max_terms, const uint8_t* dense_tf, int64_t dense_stride,
                              const int32_t* dense_terms, int32_t n_dense) {
  RAGB_REQUIRE(term_off && post_doc && post_tf && norm && idf && q_terms && q_off, RAGB_EINVAL, "%s: null pointer", who);
  RAGB_REQUIRE(vocab > 0 && n_queries > 0 && n_docs > 0, RAGB_EINVAL, "%s: empty shape", who);
  RAGB_REQUIRE(n_docs < (1ll << 31) - BM_RANGE, RAGB_ELIMIT, "%s: n_docs per shard must fit int32", who);
  RAGB_REQUIRE(max_terms >= 1 && max_terms <= BM_MAX_TERMS, RAGB_ELIMIT,
               "%s: max_query_terms=%d outside [1,%d]", who, max_terms, BM_MAX_TERMS);
  RAGB_REQUIRE(n_dense >= 0 && n_dense <= BM_MAX_DENSE, RAGB_ELIMIT, "%s: n_dense=%d outside [0,%d]", who, n_dense,
               BM_MAX_DENSE);
  if (n_dense > 0) {
    RAGB_REQUIRE(dense_tf && dense_terms, RAGB_EINVAL, "%s: dense table pointers missing", who);
    RAGB_REQUIRE(dense_stride >= n_docs && dense_stride % BM_RANGE == 0, RAGB_EINVAL,
                 "%s: dense_stride must be a multiple of %d covering n_docs", who, BM_RANGE);
    RAGB_REQUIRE((reinterpret_cast<uintptr_t>(dense_tf) & 15) == 0, RAGB_EINVAL, "%s: dense table must be 16-byte aligned", who);
  }
  RAGB_REQUIRE((reinterpret_cast<uintptr_t>(norm) & 15) == 0, RAGB_EINVAL, "%s: norm must be 16-byte aligned", who);
  return RAGB_OK;
}

}  // namespace ragb

using namespace ragb;

extern "C" {

// Debug aid (not part of the documented ABI): synchronously copy and clear the kernel counters.
int ragb_debug_bm25_counters(unsigned long long* out8) {
  unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(out8, g_bm25_dbg, sizeof(zero)) != cudaSuccess) return RAGB_ECUDA;
  if (cudaMemcpyToSymbol(g_bm25_dbg, zero, sizeof(zero)) != cudaSuccess) return RAGB_ECUDA;
  return RAGB_OK;
}

#ifdef RAGB_BM25_PROFILE
// Profiling build only: copy and clear the per-query phase cycles, out[8][4096].
int ragb_debug_bm25_profile(unsigned long long* out) {
  if (cudaMemcpyFromSymbol(out, g_bm25_prof, sizeof(unsigned long long) * 8 * BM_PROF_QUERIES) != cudaSuccess) return RAGB_ECUDA;
  void* p = nullptr;
  if (cudaGetSymbolAddress(&p, g_bm25_prof) != cudaSuccess) return RAGB_ECUDA;
  if (cudaMemset(p, 0, sizeof(unsigned long long) * 8 * BM_PROF_QUERIES) != cudaSuccess) return RAGB_ECUDA;
  return RAGB_OK;
}
#endif

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

int ragb_bm25_term_max_tf(const int64_t* term_off, const uint16_t* post_tf, const int32_t* terms, int32_t n_terms,
                          int32_t* max_tf_out, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(term_off && post_tf && terms && max_tf_out, RAGB_EINVAL, "ragb_bm25_term_max_tf: null pointer");
  RAGB_REQUIRE(n_terms > 0 && n_terms <= 65535, RAGB_ELIMIT, "ragb_bm25_term_max_tf: n_terms=%d outside [1,65535]", n_terms);
  RAGB_CUDA(cudaMemsetAsync(max_tf_out, 0, sizeof(int32_t) * n_terms, stream));
  term_max_tf_kernel<<<dim3(TB_SPLIT, n_terms), TB_THREADS, 0, stream>>>(term_off, post_tf, terms, max_tf_out);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

int ragb_bm25_build_dense_table(const int64_t* term_off, const int32_t* post_doc, const uint16_t* post_tf,
                                const int32_t* terms, int32_t n_terms, int64_t n_docs, uint8_t* table_out, int64_t stride,
                                ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(term_off && post_doc && post_tf && terms && table_out, RAGB_EINVAL, "ragb_bm25_build_dense_table: null pointer");
  RAGB_REQUIRE(n_terms > 0 && n_terms <= BM_MAX_DENSE, RAGB_ELIMIT, "ragb_bm25_build_dense_table: n_terms=%d outside [1,%d]",
               n_terms, BM_MAX_DENSE);
  RAGB_REQUIRE(stride >= n_docs && stride % BM_RANGE == 0, RAGB_EINVAL,
               "ragb_bm25_build_dense_table: stride must be a multiple of %d covering n_docs", BM_RANGE);
  RAGB_CUDA(cudaMemsetAsync(table_out, 0, static_cast<size_t>(n_terms) * stride, stream));
  dense_table_fill_kernel<<<dim3(TB_SPLIT, n_terms), TB_THREADS, 0, stream>>>(term_off, post_doc, post_tf, terms, table_out, stride);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

int ragb_bm25_build_impact_bounds(const uint8_t* dense_tf, int64_t dense_stride, int32_t n_dense, const float* norm,
                                  int64_t n_docs, uint16_t* dense_imp_fp16_out, float* dense_max_imp_out,
                                  ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(dense_tf && norm && dense_imp_fp16_out && dense_max_imp_out, RAGB_EINVAL, "ragb_bm25_build_impact_bounds: null pointer");
  RAGB_REQUIRE(n_dense > 0 && n_dense <= BM_MAX_DENSE, RAGB_ELIMIT, "ragb_bm25_build_impact_bounds: n_dense=%d outside [1,%d]",
               n_dense, BM_MAX_DENSE);
  RAGB_REQUIRE(dense_stride >= n_docs && dense_stride % BM_RANGE == 0, RAGB_EINVAL,
               "ragb_bm25_build_impact_bounds: dense_stride must be a multiple of %d covering n_docs", BM_RANGE);
  RAGB_REQUIRE(((reinterpret_cast<uintptr_t>(dense_tf) | reinterpret_cast<uintptr_t>(dense_imp_fp16_out)) & 15) == 0, RAGB_EINVAL,
               "ragb_bm25_build_impact_bounds: tables must be 16-byte aligned");
  RAGB_CUDA(cudaMemsetAsync(dense_max_imp_out, 0, sizeof(float) * n_dense, stream));
  int64_t blocks = ceil_div64(dense_stride, static_cast<int64_t>(TB_THREADS) * 8);
  if (blocks > 148 * 4) blocks = 148 * 4;
  impact_bounds_kernel<<<dim3(static_cast<unsigned>(blocks), n_dense), TB_THREADS, 0, stream>>>(
      dense_tf, norm, n_docs, dense_stride, reinterpret_cast<__half*>(dense_imp_fp16_out), dense_max_imp_out);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

int ragb_bm25_build_posting_impacts(const int32_t* post_doc, const uint16_t* post_tf, const float* norm, int64_t nnz,
                                    float* post_imp_out, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(post_doc && post_tf && norm && post_imp_out, RAGB_EINVAL, "ragb_bm25_build_posting_impacts: null pointer");
  RAGB_REQUIRE(nnz > 0, RAGB_EINVAL, "ragb_bm25_build_posting_impacts: empty index");
  int64_t blocks = ceil_div64(nnz, 256 * 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  posting_impacts_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(post_doc, post_tf, norm, nnz, post_imp_out);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

size_t ragb_bm25_topk_workspace_bytes(int32_t n_queries, int64_t n_docs, int32_t k) {
  if (n_queries <= 0 || n_docs <= 0 || k <= 0) return 0;
  int64_t stripe_docs;
  const int stripes = bm25_stripes(n_queries, n_docs, &stripe_docs);
  return static_cast<size_t>(n_queries) * stripes * k * sizeof(uint64_t) + static_cast<size_t>(n_queries) * sizeof(float);
}

int ragb_bm25_seed(const int64_t* term_off, const int32_t* post_doc, const uint16_t* post_tf, const float* norm,
                   const float* idf, int64_t vocab, double k1, const uint8_t* dense_tf, int64_t dense_stride,
                   const int32_t* dense_terms, int32_t n_dense, const int32_t* q_terms, const int32_t* q_off,
                   int32_t n_queries, int32_t max_query_terms, int64_t n_docs, int32_t k, float* seed_out,
                   ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = bm25_common_checks("ragb_bm25_seed", term_off, post_doc, post_tf, norm, idf, vocab, q_terms, q_off, n_queries,
                              n_docs, max_query_terms, dense_tf, dense_stride, dense_terms, n_dense);
  if (rc != RAGB_OK) return rc;
  RAGB_REQUIRE(seed_out, RAGB_EINVAL, "ragb_bm25_seed: null pointer");
  RAGB_REQUIRE(k > 0 && k <= RAGB_MAX_TOPK, RAGB_ELIMIT, "ragb_bm25_seed: k=%d outside [1,%d]", k, RAGB_MAX_TOPK);
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
  a.k1p1 = static_cast<float>(k1 + 1.0);
  a.max_terms = max_query_terms;
  a.dense_tf = dense_tf;
  a.dense_terms = dense_terms;
  a.dense_stride = dense_stride;
  a.n_dense = n_dense;
  a.k = k;
  a.seed_thr = seed_out;
  bm25_seed_kernel<<<n_queries, SEED_THREADS, 0, stream>>>(a);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

// Argument block shared by the one-call and the staged form of the top-k search.
static int bm25_topk_args(const char* who, Bm25Args& a, int* stripes_out, const int64_t* term_off, const int32_t* post_doc,
                          const uint16_t* post_tf, const float* norm, const float* idf, int64_t vocab, double k1,
                          const uint8_t* dense_tf, int64_t dense_stride, const int32_t* dense_terms, int32_t n_dense,
                          const uint16_t* dense_imp_fp16, const float* dense_max_imp, const float* dense_cap,
                          const int32_t* hi_off, const int32_t* hi_doc, const float* post_imp, const int32_t* q_terms,
                          const int32_t* q_off, int32_t n_queries, int32_t max_query_terms, int64_t n_docs, int64_t id_base,
                          int32_t k, void* workspace, size_t workspace_bytes) {
  int rc = bm25_common_checks(who, term_off, post_doc, post_tf, norm, idf, vocab, q_terms, q_off, n_queries, n_docs,
                              max_query_terms, dense_tf, dense_stride, dense_terms, n_dense);
  if (rc != RAGB_OK) return rc;
  RAGB_REQUIRE(workspace, RAGB_EINVAL, "%s: null pointer", who);
  RAGB_REQUIRE(k > 0 && k <= RAGB_MAX_TOPK, RAGB_ELIMIT, "%s: k=%d outside [1,%d]", who, k, RAGB_MAX_TOPK);
  RAGB_REQUIRE(id_base >= 0 && id_base + n_docs < (1ll << 31), RAGB_ELIMIT, "%s: ids must fit int32", who);
  RAGB_REQUIRE(workspace_bytes >= ragb_bm25_topk_workspace_bytes(n_queries, n_docs, k), RAGB_ENOSPC, "%s: workspace too small", who);
  a = Bm25Args{};
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
  a.dense_tf = dense_tf;
  a.dense_terms = dense_terms;
  a.dense_stride = dense_stride;
  a.n_dense = n_dense;
  RAGB_REQUIRE((dense_imp_fp16 == nullptr) == (dense_max_imp == nullptr), RAGB_EINVAL, "%s: dense_imp_fp16 and dense_max_imp go together", who);
  RAGB_REQUIRE(dense_imp_fp16 == nullptr || (n_dense > 0 && (reinterpret_cast<uintptr_t>(dense_imp_fp16) & 15) == 0),
               RAGB_EINVAL, "%s: dense_imp_fp16 needs the dense table and 16-byte alignment", who);
  a.dense_imp = reinterpret_cast<const __half*>(dense_imp_fp16);
  a.dense_maximp = dense_max_imp;
  RAGB_REQUIRE((dense_cap != nullptr) == (hi_off != nullptr) && (hi_off != nullptr) == (hi_doc != nullptr), RAGB_EINVAL,
               "%s: dense_cap, hi_off and hi_doc go together", who);
  RAGB_REQUIRE(dense_cap == nullptr || n_dense > 0, RAGB_EINVAL, "%s: an impact cap needs the dense table", who);
  a.dense_cap = dense_cap;
  a.hi_off = hi_off;
  a.hi_doc = hi_doc;
  a.post_imp = post_imp;
  a.k = k;
  a.capacity = warp_topk_capacity(k);
  a.part_keys = static_cast<uint64_t*>(workspace);
  a.out_scores = nullptr;
  const int stripes = bm25_stripes(n_queries, n_docs, &a.stripe_docs);
  a.n_stripes = stripes;
  a.seed_thr = reinterpret_cast<float*>(a.part_keys + static_cast<size_t>(n_queries) * stripes * k);
  static const int debug_flag = [] { const char* e = getenv("RAGB_BM25_DEBUG"); return e ? atoi(e) : 0; }();
  a.debug = debug_flag;
  // postings aimed at per window (0 turns window mode off); tuning aid, the default is what was measured best
  static const int window_flag = [] { const char* e = getenv("RAGB_BM25_WINDOW"); return e ? atoi(e) : 256; }();
  a.window_mode = window_flag;
  *stripes_out = stripes;
  return RAGB_OK;
}

// seeds for a search: the caller's bounds (copied, the kernel raises its private copy) or the seed kernel
static int bm25_init_seeds(const Bm25Args& a, const float* seed_thr, int n_queries, cudaStream_t stream) {
  if (seed_thr != nullptr) {
    RAGB_CUDA(cudaMemcpyAsync(a.seed_thr, seed_thr, sizeof(float) * n_queries, cudaMemcpyDeviceToDevice, stream));
  } else {
    bm25_seed_kernel<<<n_queries, SEED_THREADS, 0, stream>>>(a);
    RAGB_AFTER_LAUNCH(1);
  }
  return RAGB_OK;
}

// the scoring kernel over stripes [s0, s1); shared memory per block padded to at least min_smem bytes (0 = natural)
static int bm25_launch_stripes(Bm25Args a, int n_queries, int s0, int s1, size_t min_smem, cudaStream_t stream) {
  if (s1 <= s0) return RAGB_OK;
  static const int imp_flag = [] { const char* e = getenv("RAGB_BM25_IMPACTS"); return e ? atoi(e) : 1; }();   // 0: ignore post_imp
  const bool with_imp = a.post_imp != nullptr && a.window_mode != 0 && imp_flag != 0;
  size_t smem = bm25_smem_bytes(a.max_terms, a.capacity, false, a.n_dense);
  if (min_smem > smem) smem = min_smem;
  RAGB_REQUIRE(smem <= 200 * 1024, RAGB_ELIMIT, "ragb_bm25_score_part: shared-memory padding %zu too large", smem);
  a.stripe0 = s0;
  auto kernel = with_imp ? bm25_kernel<false, true> : bm25_kernel<false, false>;
  RAGB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  // the maximal carve-out only when a caller pads the blocks to share SMs with another kernel: it leaves 28 KB of L1,
  // and the kernel's gathers (norm, table bytes) want the L1 the default split gives them (measured: +8 % with it)
  RAGB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 min_smem > 0 ? static_cast<int>(cudaSharedmemCarveoutMaxShared)
                                              : static_cast<int>(cudaSharedmemCarveoutDefault)));
  kernel<<<dim3(n_queries, s1 - s0), BM_THREADS, smem, stream>>>(a);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

int ragb_bm25_score_topk(const int64_t* term_off, const int32_t* post_doc, const uint16_t* post_tf, const float* norm,
                         const float* idf, int64_t vocab, double k1, const uint8_t* dense_tf, int64_t dense_stride,
                         const int32_t* dense_terms, int32_t n_dense, const uint16_t* dense_imp_fp16,
                         const float* dense_max_imp, const float* dense_cap, const int32_t* hi_off,
                         const int32_t* hi_doc, const float* post_imp, const int32_t* q_terms, const int32_t* q_off,
                         int32_t n_queries, int32_t max_query_terms, int64_t n_docs, int64_t id_base, int32_t k,
                         const float* seed_thr, float* out_score, int32_t* out_id, void* workspace,
                         size_t workspace_bytes, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(out_score && out_id, RAGB_EINVAL, "ragb_bm25_score_topk: null pointer");
  Bm25Args a;
  int stripes = 0;
  int rc = bm25_topk_args("ragb_bm25_score_topk", a, &stripes, term_off, post_doc, post_tf, norm, idf, vocab, k1, dense_tf,
                          dense_stride, dense_terms, n_dense, dense_imp_fp16, dense_max_imp, dense_cap, hi_off, hi_doc, post_imp, q_terms, q_off,
                          n_queries,
                          max_query_terms, n_docs, id_base, k, workspace, workspace_bytes);
  if (rc != RAGB_OK) return rc;
  rc = bm25_init_seeds(a, seed_thr, n_queries, stream);
  if (rc != RAGB_OK) return rc;
  rc = bm25_launch_stripes(a, n_queries, 0, stripes, 0, stream);
  if (rc != RAGB_OK) return rc;
  return launch_merge_keys(a.part_keys, n_queries, stripes, k, k, out_score, out_id, stream);
}

int32_t ragb_bm25_stripe_count(int32_t n_queries, int64_t n_docs) {
  if (n_queries <= 0 || n_docs <= 0) return 0;
  int64_t stripe_docs;
  return bm25_stripes(n_queries, n_docs, &stripe_docs);
}

int ragb_bm25_score_part(const int64_t* term_off, const int32_t* post_doc, const uint16_t* post_tf, const float* norm,
                         const float* idf, int64_t vocab, double k1, const uint8_t* dense_tf, int64_t dense_stride,
                         const int32_t* dense_terms, int32_t n_dense, const uint16_t* dense_imp_fp16,
                         const float* dense_max_imp, const float* dense_cap, const int32_t* hi_off,
                         const int32_t* hi_doc, const float* post_imp, const int32_t* q_terms, const int32_t* q_off,
                         int32_t n_queries, int32_t max_query_terms, int64_t n_docs, int64_t id_base, int32_t k,
                         const float* seed_thr, int32_t stripe_begin, int32_t stripe_end, int64_t min_smem_bytes,
                         void* workspace, size_t workspace_bytes, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Bm25Args a;
  int stripes = 0;
  int rc = bm25_topk_args("ragb_bm25_score_part", a, &stripes, term_off, post_doc, post_tf, norm, idf, vocab, k1, dense_tf,
                          dense_stride, dense_terms, n_dense, dense_imp_fp16, dense_max_imp, dense_cap, hi_off, hi_doc, post_imp, q_terms, q_off,
                          n_queries,
                          max_query_terms, n_docs, id_base, k, workspace, workspace_bytes);
  if (rc != RAGB_OK) return rc;
  RAGB_REQUIRE(0 <= stripe_begin && stripe_begin <= stripe_end && stripe_end <= stripes, RAGB_EINVAL,
               "ragb_bm25_score_part: stripes [%d, %d) outside [0, %d]", stripe_begin, stripe_end, stripes);
  RAGB_REQUIRE(min_smem_bytes >= 0, RAGB_EINVAL, "ragb_bm25_score_part: negative shared-memory padding");
  if (stripe_begin == 0) {   // the part that starts the search also starts the thresholds
    rc = bm25_init_seeds(a, seed_thr, n_queries, stream);
    if (rc != RAGB_OK) return rc;
  }
  return bm25_launch_stripes(a, n_queries, stripe_begin, stripe_end, static_cast<size_t>(min_smem_bytes), stream);
}

int ragb_bm25_score_finish(int32_t n_queries, int64_t n_docs, int32_t k, float* out_score, int32_t* out_id,
                           const void* workspace, size_t workspace_bytes, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(out_score && out_id && workspace, RAGB_EINVAL, "ragb_bm25_score_finish: null pointer");
  RAGB_REQUIRE(n_queries > 0 && n_docs > 0 && k > 0 && k <= RAGB_MAX_TOPK, RAGB_EINVAL, "ragb_bm25_score_finish: bad shape");
  RAGB_REQUIRE(workspace_bytes >= ragb_bm25_topk_workspace_bytes(n_queries, n_docs, k), RAGB_ENOSPC,
               "ragb_bm25_score_finish: workspace too small");
  int64_t stripe_docs;
  const int stripes = bm25_stripes(n_queries, n_docs, &stripe_docs);
  return launch_merge_keys(static_cast<const uint64_t*>(workspace), n_queries, stripes, k, k, out_score, out_id, stream);
}

int ragb_bm25_scores(const int64_t* term_off, const int32_t* post_doc, const uint16_t* post_tf, const float* norm,
                     const float* idf, int64_t vocab, double k1, const uint8_t* dense_tf, int64_t dense_stride,
                     const int32_t* dense_terms, int32_t n_dense, const int32_t* q_terms, const int32_t* q_off,
                     int32_t n_queries, int32_t max_query_terms, int64_t n_docs, float* out_scores, int64_t out_ld,
                     int32_t tiled, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = bm25_common_checks("ragb_bm25_scores", term_off, post_doc, post_tf, norm, idf, vocab, q_terms, q_off,
                              n_queries, n_docs, max_query_terms, dense_tf, dense_stride, dense_terms, n_dense);
  if (rc != RAGB_OK) return rc;
  RAGB_REQUIRE(out_scores, RAGB_EINVAL, "ragb_bm25_scores: null pointer");
  RAGB_REQUIRE(tiled == 0 || tiled == 1, RAGB_EINVAL, "ragb_bm25_scores: tiled must be 0 or 1");
  RAGB_REQUIRE(tiled ? out_ld >= n_queries : out_ld >= n_docs, RAGB_EINVAL,
               "ragb_bm25_scores: out_ld=%lld too small (row-major: >= n_docs, tiled: >= n_queries)", static_cast<long long>(out_ld));
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
  a.dense_tf = dense_tf;
  a.dense_terms = dense_terms;
  a.dense_stride = dense_stride;
  a.n_dense = n_dense;
  a.k = 1;
  a.capacity = 0;
  const int stripes = bm25_stripes(n_queries, n_docs, &a.stripe_docs);
  a.out_scores = out_scores;
  a.out_ld = out_ld;
  a.out_tiled = tiled;
  const size_t smem = bm25_smem_bytes(a.max_terms, 0, true, a.n_dense);
  RAGB_CUDA(cudaFuncSetAttribute(bm25_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  bm25_kernel<true><<<dim3(n_queries, stripes), BM_THREADS, smem, stream>>>(a);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

}  // extern "C"
