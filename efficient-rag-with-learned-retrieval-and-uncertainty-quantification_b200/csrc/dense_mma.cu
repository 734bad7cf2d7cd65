// Batched dense scoring on the 5th-generation tensor cores (K2b in SURVEY.md):
// scores = Q[B,dim] x E[n_rows,dim]^T in bf16 with fp32 accumulation in TMEM, and a fused
// per-query running top-k in the epilogue so the [B, n_rows] score matrix never exists.
//
// Replaces DenseIndex.search -> collection.query (rag_uq/streaming_index.py:353-370) for
// query batches; the reference answers one query at a time through ChromaDB's HNSW.
//
// Roles (320 threads, 1 block per SM, persistent):
//   warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled tiles into a ring of
//               shared-memory stages, completion on mbarriers.
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (M = 128 queries).
//   warps 2..9  epilogue (two warps per TMEM lane quadrant, one per half of the accumulator
//               columns): tcgen05.ld of the accumulator; THREAD r OWNS QUERY r of the
//               slab for the whole kernel, so its admission threshold is a register and
//               the common case is "32 scores, one max, one compare".  Survivors are
//               appended to the thread's private list (global workspace, L2-resident); a full
//               list is compacted to its best k by the whole warp (shuffle bitonic sort).
//   The accumulator is double-buffered in TMEM, so the epilogue of tile i overlaps the
//   MMAs of tile i+1.
//
// Work split: block c serves query slab (c % n_slabs) and passage group (c / n_slabs); the
// blocks of one group read the same passages at the same time, so HBM sees each passage
// row once and L2 serves the other slabs.
//
// variant 0 (SS): A = query slab and B = passage tile both streamed through shared memory.
// variant 1 (TS): the query slab [128 x dim] is stored ONCE into TMEM (dim/2 columns of
//   packed bf16 pairs) and used as the A operand from there; only passages stream through
//   shared memory, which halves the L2 -> SM traffic per flop.  Needs dim <= 768.
// variant 2: like 0 with 256-passage tiles (N = 256 per tcgen05.mma, all 512 TMEM columns as two
//   accumulators): 25% less L2 -> SM traffic per flop, fewer pipeline stages.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "router.cuh"

namespace ragb {

constexpr int MM_BM = 128;
constexpr int MM_BK = 64;
constexpr int MM_THREADS = 320;  // TMA warp, MMA warp, 8 epilogue warps
constexpr int MM_A_STAGE_BYTES = MM_BM * MM_BK * 2;  // 16 KB
constexpr int MM_MAX_STAGES = 16;
constexpr int MM_MAX_SMEM = 227 * 1024;
constexpr int MM_PACE_TILES = 4;         // how far a block may run ahead of the slowest slab of its group
constexpr int MM_PROGRESS_BYTES = 4096;  // head of the workspace: progress counters (<= 148 blocks)

struct MmaArgs {
  const uint4* queries;  // raw pointer, used by variant 1 to fill TMEM
  int64_t n_rows;
  int64_t id_base;
  int n_queries;
  int dim;
  int k;
  int n_slabs;
  int n_groups;
  int tiles_per_group;
  int tile_lo, tile_hi;  // this launch covers passage tiles [tile_lo, tile_hi) (the sampled prefix, or the rest)
  int n_stages;
  const float* seed_thr;  // optional [n_queries]: a proven lower bound of every query's k-th best score; lists start there
  float* min_out;         // optional [n_queries], initialised to +inf by the caller: receives the SMALLEST score of every query
  uint64_t* part_keys;  // [n_queries, lists_per_query, k]; this launch fills slots [group * 2 + half]
  int lists_per_query;
  int* progress;        // [n_groups, n_slabs] tiles issued so far (soft pacing between the slabs of a group)
  uint64_t* lists;      // [blocks, 128, list capacity + 1] per-thread candidate lists
  int stage_limit;
  int pace_tiles;       // how far (in tiles) a block may run ahead of the slowest slab of its passage group
  int pace_mask;        // a block publishes its position and checks the others every (pace_mask + 1) tiles
  // ---- full-fusion epilogue only (FUSED kernels) ----
  const float* bm25;    // [ceil(n_rows / 256), bm25_rows, 256] fp32 BM25 scores of this shard's rows, 256-passage tiles
  int64_t bm25_ld;      // query rows per tile (>= n_queries)
  RouterWeights rw;     // running-statistics router (router.py:130-132)
  const uint32_t* ff_table;  // [ff_nb, ff_nd] fp16 pair (lo, hi): bounds of the gate on each (bm25, dense) cell
  int ff_nb, ff_nd;
  float ff_inv_wb;        // ff_nb / b_cap
  float ff_inv_wd;        // ff_nd / (2 * d_hi)
  float ff_d_hi;          // |dense| <= d_hi for every pair (caller's guarantee)
  unsigned long long* counters;  // optional [2]: gate evaluations, admissions (debug / bench)
  int ff_debug;           // RAGB_FF_DEBUG (timing attribution only): 1 = no bm25 loads, 2 = no bound scan
};

// ---- PTX wrappers -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000ll) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

// K-major, 128-byte swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // start address      bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                       // leading byte offset (unused, swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;               // stride byte offset  bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                       // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                       // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// Descending bitonic sort of 32*KPL keys held striped over a warp (element e = r*32 + lane).
template <int KPL>
__device__ __forceinline__ void warp_sort_desc(uint64_t (&v)[KPL], const int lane) {
  constexpr int N = 32 * KPL;
#pragma unroll
  for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
    for (int j = size >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int jr = j >> 5;
#pragma unroll
        for (int r = 0; r < KPL; ++r) {
          const int r2 = r ^ jr;
          if (r2 > r) {
            const bool desc = (((r * 32 + lane) & size) == 0);
            const uint64_t x = v[r], y = v[r2];
            const bool sw = (x < y) == desc;
            v[r] = sw ? y : x;
            v[r2] = sw ? x : y;
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < KPL; ++r) {
          const uint64_t other = __shfl_xor_sync(0xffffffffu, v[r], j);
          const bool desc = (((r * 32 + lane) & size) == 0);
          const bool take_max = (((lane & j) == 0) == desc);
          const uint64_t mx = v[r] > other ? v[r] : other;
          const uint64_t mn = v[r] > other ? other : v[r];
          v[r] = take_max ? mx : mn;
        }
      }
    }
  }
}

// Per-thread candidate list: LIST_CAP = 32*KPL slots, the best k sorted in front after a
// compaction, appended survivors behind them.  When a lane's list is full the WHOLE WARP sorts
// it (KPL keys per lane, shuffles only), keeps the best k and tightens that lane's threshold:
// an append is one shared-memory store, a compaction ~20 shuffle stages once per (cap - k) appends.
struct ListState {
  int cnt;
  float thr_score;
  uint64_t thr_key;
};

// Deliberately NOT inlined: it is called from inside the 32-way unrolled admission loop and
// inlining it there blows the instruction cache (measured: stall_no_inst dominated the kernel).
template <int KPL>
__device__ __noinline__ ListState compact_lists(uint64_t* warp_lists, unsigned lanes, const int lane, const int k,
                                                ListState st) {
  constexpr int CAP = 32 * KPL;
  constexpr int STRIDE = CAP + 1;
  __syncwarp();
  while (lanes) {
    const int src = __ffs(lanes) - 1;
    lanes &= lanes - 1;
    const int n_src = __shfl_sync(0xffffffffu, st.cnt, src);
    uint64_t* list = warp_lists + src * STRIDE;
    uint64_t v[KPL];
#pragma unroll
    for (int r = 0; r < KPL; ++r)
      v[r] = (r * 32 + lane) < n_src ? __ldcg(reinterpret_cast<const unsigned long long*>(list + r * 32 + lane)) : 0ull;
    warp_sort_desc<KPL>(v, lane);
#pragma unroll
    for (int r = 0; r < KPL; ++r) list[r * 32 + lane] = v[r];
    uint64_t kth = 0ull;
#pragma unroll
    for (int r = 0; r < KPL; ++r)
      if (r == ((k - 1) >> 5)) kth = v[r];
    kth = __shfl_sync(0xffffffffu, kth, (k - 1) & 31);
    if (lane == src) {
      st.cnt = n_src < k ? n_src : k;
      if (n_src >= k) {
        st.thr_key = kth;
        st.thr_score = key_score(kth);
      }
    }
  }
  __syncwarp();
  return st;
}

// Epilogue shared by the 1-CTA and the CTA-pair kernel.  8 warps: two per TMEM lane quadrant, each
// scanning one half of the accumulator columns (a single warp per scheduler could not keep up with
// the tensor core at k = 50: measured +3 ms).  Thread (quadrant, lane) owns query row 32*quadrant+lane
// of the slab for the whole kernel and keeps its candidates of "its" column half in a private list.
template <int BN, int KPL>
__device__ __forceinline__ void epilogue_scan(const MmaArgs& a, const uint32_t tmem_base, const uint32_t acc_col0,
                                              uint64_t* block_lists, const int warp, const int lane, const int slab,
                                              const int group, const int tile_begin, const int tile_end,
                                              uint64_t* bar_tmem_full, const uint32_t empty_addr0,
                                              const uint32_t empty_addr1, const bool remote_arrive) {
  constexpr int LIST_CAP = 32 * KPL;
  constexpr int LIST_STRIDE = LIST_CAP + 1;
  constexpr int HALF_CHUNKS = BN / 64;  // 32-column chunks per half
  const int quad = warp & 3;
  const int half = (warp - 2) >> 2;
  const int query = slab * MM_BM + quad * 32 + lane;
  const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
  uint64_t* warp_lists = block_lists + static_cast<size_t>(half * 4 + quad) * 32 * LIST_STRIDE;
  uint64_t* my_list = warp_lists + lane * LIST_STRIDE;
  ListState st{0, -INFINITY, 0ull};
  if (a.seed_thr != nullptr && query < a.n_queries) {
    // nothing below a proven lower bound of the query's k-th best score can end up in its top-k: start the list
    // there instead of at -inf (an equal score still passes: the key's id bits are never all zero)
    const float seed = __ldg(a.seed_thr + query);
    if (seed > -INFINITY) {
      st.thr_score = seed;
      st.thr_key = static_cast<uint64_t>(float_to_ordered(seed)) << 32;
    }
  }
  const bool track_min = a.min_out != nullptr;
  float lowest = INFINITY;
  uint32_t buf = 0, acc_phase = 0;
  for (int tile = tile_begin; tile < tile_end; ++tile) {
    mbar_wait(smem_u32(&bar_tmem_full[buf]), acc_phase);
    tc_fence_after();
    const int64_t row0 = static_cast<int64_t>(tile) * BN;
    const int valid = static_cast<int>(min(static_cast<int64_t>(BN), a.n_rows - row0));  // rows of this tile that exist
#pragma unroll 1
    for (int cc = 0; cc < HALF_CHUNKS; ++cc) {
      const int c = half * HALF_CHUNKS + cc;
      uint32_t v[32];
      __syncwarp();
      tc_ld32(tmem_base + lane_addr + acc_col0 + buf * BN + c * 32, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float m = __uint_as_float(v[0]);
#pragma unroll
      for (int i = 1; i < 32; ++i) m = fmaxf(m, __uint_as_float(v[i]));
      if (track_min) {   // block-uniform; rows past the end of the shard read as 0, which can only lower the minimum
        float mn = __uint_as_float(v[0]);
#pragma unroll
        for (int i = 1; i < 32; ++i) mn = fminf(mn, __uint_as_float(v[i]));
        lowest = fminf(lowest, mn);
      }
      if (__any_sync(0xffffffffu, m >= st.thr_score)) {
        // room for a whole chunk is guaranteed up front, so the admission loop has no votes in it
        const unsigned full = __ballot_sync(0xffffffffu, st.cnt > LIST_CAP - 32);
        if (full) st = compact_lists<KPL>(warp_lists, full, lane, a.k, st);
        // 32 different queries share the warp, so in the warm-up phase (and on small shards for the whole
        // kernel) some lane has a candidate in almost every chunk.  Walking all 32 columns with a compare,
        // a bounds test and a predicated append each cost ~25 instructions per column (47% of the samples at
        // 1.25M rows, k = 50).  Instead every lane records its candidates in a bit mask (2 instructions per
        // column) and the warp then serves one candidate per lane per round: a register select picks the
        // score, so nothing is indexed dynamically; the number of rounds is the largest candidate count of
        // any lane - one or two outside the first chunks of a list.
        const int32_t id0 = static_cast<int32_t>(a.id_base + row0) + c * 32;
        const int lim = valid - c * 32;
        uint32_t pm = 0u;
#pragma unroll
        for (int i = 0; i < 32; ++i) pm |= (__uint_as_float(v[i]) >= st.thr_score) ? (1u << i) : 0u;
        if (lim < 32) pm &= lim <= 0 ? 0u : ((1u << lim) - 1u);
        while (__any_sync(0xffffffffu, pm != 0u)) {
          const int sel = __ffs(pm) - 1;
          float sv = 0.0f;
#pragma unroll
          for (int i = 0; i < 32; ++i) sv = i == sel ? __uint_as_float(v[i]) : sv;
          if (pm != 0u) {
            const uint64_t key = make_key(sv, id0 + sel);
            if (key > st.thr_key) my_list[st.cnt++] = key;
            pm &= pm - 1u;
          }
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      const uint32_t bar = buf == 0 ? empty_addr0 : empty_addr1;
      if (remote_arrive)
        asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
      else
        mbar_arrive(bar);
    }
    buf ^= 1;
    if (buf == 0) acc_phase ^= 1;
  }
  // only lists that hold more than k candidates have to be cut down (a whole-warp sort per list: ~50 us per warp when all
  // 32 need it); the merge kernel takes its input as a bag, so a seeded list that admitted <= k candidates is written
  // out as it is - on a small shard that is most of them
  {
    const unsigned over = __ballot_sync(0xffffffffu, st.cnt > a.k);
    if (over) st = compact_lists<KPL>(warp_lists, over, lane, a.k, st);
  }
  const int cnt = st.cnt;
  if (query < a.n_queries) {
    uint64_t* dst = a.part_keys + (static_cast<int64_t>(query) * a.lists_per_query + group * 2 + half) * a.k;
    for (int j = 0; j < a.k; ++j) dst[j] = j < cnt ? __ldcg(reinterpret_cast<const unsigned long long*>(my_list + j)) : 0ull;
    if (track_min && lowest < INFINITY) {
      // float minimum through integer atomics on the raw bits: non-negative floats order like signed ints, negative
      // ones in reverse like unsigned ints, and every negative pattern is larger (unsigned) than every positive one
      if (lowest >= 0.0f) atomicMin(reinterpret_cast<int*>(a.min_out + query), __float_as_int(lowest));
      else atomicMax(reinterpret_cast<unsigned*>(a.min_out + query), __float_as_uint(lowest));
    }
  }
}

// =============================================================================================
// Full-fusion epilogue (RetrievalRouter.hybrid_rerank over ALL passages,
// rag_uq/router.py:179-202): the candidate score is
//     fused = g * dense + (1 - g) * bm25 = bm25 + g * (dense - bm25),   g = gate(bm25, dense)
// with dense taken from the TMEM accumulator and bm25 read from the [n_queries, n_rows] matrix the
// BM25 kernel wrote.  Evaluating the 3 -> H -> 1 gate for every (query, passage) pair would cost
// ~7x the GEMM on the CUDA cores, so the scan is driven by an EXACT bound instead.  The host hands in
// a table of gate bounds lo[ib][id] <= g(b, d) <= hi[ib][id] for every (b, d) of a cell (computed from
// the vertices of the gate's piecewise-linear pre-activation, router.py full_fusion_bounds; two
// bf16 per cell, 32 KB, kept in shared memory).  Then
//     fused <= b + (d <= b ? lo : hi) * (d - b)
// so a pair whose bound is below the thread's admission threshold cannot enter its top-k.  Per
// pair that is ~15 instructions and one shared-memory lookup.  The gate itself runs only for the
// survivors (~0.1 % of the pairs), evaluated by the whole warp: each lane owns hidden units
// lane, lane+32, ... with their weights in registers, the partial sums are folded by shuffles.
// =============================================================================================
constexpr int FF_TABLE_CELLS = 8192;   // gate-bound cells (n_b * n_d <= this), 4 bytes each: 32 KB of shared memory
constexpr int FF_MAX_UNITS = RAGB_ROUTER_MAX_HIDDEN / 32;

// 256-bit streaming load (sm_100): read once, keep out of L1, first in line for L2 eviction so the
// bm25 matrix does not push the passage tiles (which the other query slabs re-read) out of L2
__device__ __forceinline__ void ldg_stream_f8(const float* p, float* dst) {
  uint32_t r[8];
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
#pragma unroll
  for (int i = 0; i < 8; ++i) dst[i] = __uint_as_float(r[i]);
}

// bm25 values of 32 consecutive passages of one query.  The matrix is stored in 256-passage tiles with the
// query rows of a tile next to each other ([tile][query][256]): the 128 rows a block reads for one passage
// tile are one contiguous 128 KB piece (one DRAM / TLB page instead of 128 rows 4*n_rows bytes apart), and
// the last tile is padded, so the load is never out of bounds (columns beyond n_rows are masked by the caller).
__device__ __forceinline__ void load_bm25_chunk(const float* __restrict__ p, float (&b)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) ldg_stream_f8(p + 8 * i, b + 8 * i);
}

struct FusedBound {
  const uint32_t* table;  // shared memory, [n_b][n_d]: bf16 lo in bits 0-15, bf16 hi in bits 16-31
  float inv_wb, inv_wd, d_hi;
  uint32_t nb_max;
  int nd_max;  // n_b - 1, n_d - 1 (n_d is a power of two)
  int nd_shift;
  // Full-rate ALU only (no F2I / F2F): for 0.25 <= x < 2^22, fmaf(.., 8388607.5f) = x + (2^23 - 0.5) rounds to
  // 2^23 + floor(x), so the low mantissa bits are the cell index (a point exactly on a cell boundary may land in
  // either neighbour: both closed cells contain it).  The magic constant must be the LAST rounding step: the
  // offset of the dense axis is therefore added to d first (d + d_hi >= 0), never folded into the constant -
  // 2^23 - 0.5 + d_hi / wd is not representable and would turn the floor into a round-to-nearest.
  // x < 0.25 gives 2^23 - 0.5 (bit pattern just below 0x4B000000): as a signed difference that is -1 and clamps
  // to cell 0.  Negative, huge or NaN bm25 wrap to a large UNSIGNED value and clamp into the last row
  // (lo = 0, hi = 1: bound = max(b, d)); dense stays inside +-d_hi by contract and is clamped (not masked) so a
  // violation lands in an edge cell instead of wrapping to the opposite end of the table.
  __device__ __forceinline__ float operator()(const float b, const float d) const {
    const uint32_t ib = min(__float_as_uint(fmaf(b, inv_wb, 8388607.5f)) - 0x4B000000u, nb_max);
    const int id = min(max(static_cast<int>(__float_as_uint(fmaf(d + d_hi, inv_wd, 8388607.5f)) - 0x4B000000u), 0), nd_max);
    const uint32_t e = table[(ib << nd_shift) + id];
    const float diff = d - b;
    const float g = __uint_as_float(diff <= 0.0f ? (e << 16) : (e & 0xffff0000u));
    return fmaf(g, diff, b);
  }
};

// This lane's share of the gate MLP (hidden units lane, lane + 32, ...; absent units have zero weights).
struct LaneGate {
  float w0[FF_MAX_UNITS], w1[FF_MAX_UNITS], w2[FF_MAX_UNITS], b1[FF_MAX_UNITS], v[FF_MAX_UNITS];
  float b2, mean_b, den_b, mean_d, den_d;
  int units;
  __device__ __forceinline__ void load(const RouterWeights& rw, const int lane) {
    units = (rw.hidden + 31) >> 5;
#pragma unroll
    for (int u = 0; u < FF_MAX_UNITS; ++u) {
      const int j = u * 32 + lane;
      const bool ok = j < rw.hidden;
      w0[u] = ok ? __ldg(rw.w1 + 3 * j) : 0.0f;
      w1[u] = ok ? __ldg(rw.w1 + 3 * j + 1) : 0.0f;
      w2[u] = ok ? __ldg(rw.w1 + 3 * j + 2) : 0.0f;
      b1[u] = ok ? __ldg(rw.b1 + j) : 0.0f;
      v[u] = ok ? __ldg(rw.w2 + j) : 0.0f;
    }
    b2 = __ldg(rw.b2);
    mean_b = __ldg(rw.stats);
    den_b = __ldg(rw.stats + 1) + RT_EPS;
    mean_d = __ldg(rw.stats + 2);
    den_d = __ldg(rw.stats + 3) + RT_EPS;
  }
  // all 32 lanes call this with the same (xb, xd); everyone gets the gate
  __device__ __forceinline__ float operator()(const float xb, const float xd) const {
    const float bn = (xb - mean_b) / den_b;
    const float dn = (xd - mean_d) / den_d;
    const float df = dn - bn;
    float p = 0.0f;
#pragma unroll
    for (int u = 0; u < FF_MAX_UNITS; ++u) {
      if (u < units) {
        float h = fmaf(w2[u], df, fmaf(w1[u], dn, fmaf(w0[u], bn, b1[u])));
        h = h < 0.0f ? 0.0f : h;
        p = fmaf(v[u], h, p);
      }
    }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) p += __shfl_xor_sync(0xffffffffu, p, sh);
    return sigmoidf_exact(b2 + p);
  }
};

// Debug aid (tests): the bound exactly as the fused epilogue evaluates it, for arbitrary (bm25, dense) pairs.
__global__ void ff_bound_debug_kernel(const float* __restrict__ b, const float* __restrict__ d, int n,
                                      const uint32_t* __restrict__ table, int n_b, int n_d, float inv_wb, float inv_wd,
                                      float d_hi, float* __restrict__ out) {
  const FusedBound bound{table, inv_wb, inv_wd, d_hi, static_cast<uint32_t>(n_b - 1), n_d - 1, 31 - __clz(n_d)};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = bound(b[i], d[i]);
}

template <int BN, int KPL>
__device__ __forceinline__ void epilogue_scan_fused(const MmaArgs& a, const uint32_t* s_table, const uint32_t tmem_base,
                                                    const uint32_t acc_col0, uint64_t* block_lists, const int warp,
                                                    const int lane, const int slab, const int group,
                                                    const int tile_begin, const int tile_end, uint64_t* bar_tmem_full,
                                                    const uint32_t empty_addr0, const uint32_t empty_addr1,
                                                    const bool remote_arrive) {
  constexpr int LIST_CAP = 32 * KPL;
  constexpr int LIST_STRIDE = LIST_CAP + 1;
  constexpr int HALF_CHUNKS = BN / 64;  // 32-column chunks per half
  const int quad = warp & 3;
  const int half = (warp - 2) >> 2;
  const int query = slab * MM_BM + quad * 32 + lane;
  const bool live = query < a.n_queries;
  const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
  uint64_t* warp_lists = block_lists + static_cast<size_t>(half * 4 + quad) * 32 * LIST_STRIDE;
  uint64_t* my_list = warp_lists + lane * LIST_STRIDE;
  // this thread's row inside tile 0; tile t is t * bm25_ld * BN floats further (BN == 256 == the tile width)
  static_assert(BN == 256, "the bm25 matrix is stored in 256-passage tiles");
  const float* brow = a.bm25 + static_cast<int64_t>(live ? query : 0) * BN;
  const int64_t tile_stride = a.bm25_ld * BN;
  const FusedBound bound{s_table, a.ff_inv_wb, a.ff_inv_wd, a.ff_d_hi,
                         static_cast<uint32_t>(a.ff_nb - 1), a.ff_nd - 1, 31 - __clz(a.ff_nd)};
  LaneGate gate;
  gate.load(a.rw, lane);
  ListState st{0, -INFINITY, 0ull};
  // pruning threshold = admission threshold minus rounding slack; padding rows of the slab never pass
  float thr_cmp = live ? -INFINITY : INFINITY;
  unsigned long long n_eval = 0, n_admit = 0;
  uint32_t buf = 0, acc_phase = 0;
  float nb[32];  // bm25 values of the NEXT chunk, in flight while the current one is scanned
  if (tile_begin < tile_end) load_bm25_chunk(brow + tile_begin * tile_stride + half * HALF_CHUNKS * 32, nb);
  for (int tile = tile_begin; tile < tile_end; ++tile) {
    mbar_wait(smem_u32(&bar_tmem_full[buf]), acc_phase);
    tc_fence_after();
    const int64_t row0 = static_cast<int64_t>(tile) * BN;
    const int valid = static_cast<int>(min(static_cast<int64_t>(BN), a.n_rows - row0));
#pragma unroll 1
    for (int cc = 0; cc < HALF_CHUNKS; ++cc) {
      const int c = half * HALF_CHUNKS + cc;
      float b[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) b[i] = nb[i];
      {  // prefetch: next chunk of this tile, or the first chunk of the next tile
        const bool last = cc + 1 == HALF_CHUNKS;
        const int nt = last ? tile + 1 : tile;
        const int nc = last ? half * HALF_CHUNKS : c + 1;
        if (nt < tile_end && !(a.ff_debug & 1)) load_bm25_chunk(brow + nt * tile_stride + nc * 32, nb);
      }
      uint32_t v[32];
      __syncwarp();
      tc_ld32(tmem_base + lane_addr + acc_col0 + buf * BN + c * 32, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // one bit per pair of this lane's query that may still be admitted
      uint32_t pm = 0u;
#pragma unroll
      for (int i = 0; i < 32; ++i) pm |= (bound(b[i], __uint_as_float(v[i])) >= thr_cmp) ? (1u << i) : 0u;
      if (a.ff_debug & 2) pm = 0u;
      const int lim = valid - c * 32;
      if (lim < 32) pm &= lim <= 0 ? 0u : ((1u << lim) - 1u);
      if (__any_sync(0xffffffffu, pm != 0u)) {
        const unsigned full = __ballot_sync(0xffffffffu, st.cnt > LIST_CAP - 32);
        if (full) {
          st = compact_lists<KPL>(warp_lists, full, lane, a.k, st);
          if (live && st.cnt >= a.k) thr_cmp = st.thr_score - 1e-5f * fabsf(st.thr_score) - 1e-6f;
        }
        const int32_t id0 = static_cast<int32_t>(a.id_base + row0) + c * 32;
        unsigned active;
        while ((active = __ballot_sync(0xffffffffu, pm != 0u)) != 0u) {
          // every lane with work picks its next pair (register select, no dynamic indexing) ...
          const int sel = __ffs(pm) - 1;
          float bs = 0.0f, ds = 0.0f;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            bs = i == sel ? b[i] : bs;
            ds = i == sel ? __uint_as_float(v[i]) : ds;
          }
          pm &= pm - 1u;
          // ... and the warp evaluates the gate of one pair at a time
          while (active) {
            const int src = __ffs(active) - 1;
            active &= active - 1u;
            const float xb = __shfl_sync(0xffffffffu, bs, src);
            const float xd = __shfl_sync(0xffffffffu, ds, src);
            const float g = gate(xb, xd);
            if (lane == src) {
              const float h = fuse_scores(g, xb, xd);
              ++n_eval;
              if (h >= st.thr_score) {
                const uint64_t key = make_key(h, id0 + sel);
                if (key > st.thr_key) {
                  my_list[st.cnt++] = key;
                  ++n_admit;
                }
              }
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      const uint32_t bar = buf == 0 ? empty_addr0 : empty_addr1;
      if (remote_arrive)
        asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
      else
        mbar_arrive(bar);
    }
    buf ^= 1;
    if (buf == 0) acc_phase ^= 1;
  }
  st = compact_lists<KPL>(warp_lists, 0xffffffffu, lane, a.k, st);
  const int cnt = st.cnt;
  if (live) {
    uint64_t* dst = a.part_keys + (static_cast<int64_t>(query) * a.lists_per_query + group * 2 + half) * a.k;
    for (int j = 0; j < a.k; ++j) dst[j] = j < cnt ? __ldcg(reinterpret_cast<const unsigned long long*>(my_list + j)) : 0ull;
  }
  if (a.counters != nullptr) {
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) {
      n_eval += __shfl_xor_sync(0xffffffffu, n_eval, sh);
      n_admit += __shfl_xor_sync(0xffffffffu, n_admit, sh);
    }
    if (lane == 0) {
      atomicAdd(a.counters, n_eval);
      atomicAdd(a.counters + 1, n_admit);
    }
  }
}

template <int BN, bool A_IN_TMEM, int KPL, bool FUSED = false>
__global__ void __launch_bounds__(MM_THREADS, 1) dense_mma_kernel(const __grid_constant__ CUtensorMap tmap_q,
                                                                  const __grid_constant__ CUtensorMap tmap_e,
                                                                  const MmaArgs a) {
  constexpr int B_STAGE_BYTES = BN * MM_BK * 2;
  constexpr int STAGE_BYTES = (A_IN_TMEM ? 0 : MM_A_STAGE_BYTES) + B_STAGE_BYTES;
  constexpr uint32_t IDESC = make_idesc(MM_BM, BN);

  const int n_slabs = a.n_slabs;
  if (static_cast<int>(blockIdx.x) >= n_slabs * a.n_groups) return;
  const int slab = blockIdx.x % n_slabs;
  const int group = blockIdx.x / n_slabs;
  const int tile_begin = min(a.tile_hi, a.tile_lo + group * a.tiles_per_group);
  const int tile_end = min(a.tile_hi, tile_begin + a.tiles_per_group);
  const int n_kb = a.dim / MM_BK;
  const int a_cols = A_IN_TMEM ? a.dim / 2 : 0;  // TMEM columns holding the packed query slab
  const uint32_t tmem_cols = A_IN_TMEM ? 512u : (2 * BN <= 32 ? 32u : (2 * BN <= 64 ? 64u : (2 * BN <= 128 ? 128u : (2 * BN <= 256 ? 256u : 512u))));

  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* stages = base;
  uint32_t* s_table = reinterpret_cast<uint32_t*>(stages + static_cast<size_t>(a.n_stages) * STAGE_BYTES);  // FUSED only
  if constexpr (FUSED) {
    for (int i = threadIdx.x; i < a.ff_nb * a.ff_nd; i += MM_THREADS) s_table[i] = __ldg(a.ff_table + i);
  }
  // candidate lists live in the caller's workspace (L2-resident, touched ~once per thousand scores), so
  // shared memory holds only the operand ring and blocks of other kernels can share the SM
  uint64_t* lists = a.lists + static_cast<size_t>(blockIdx.x) * 2 * MM_BM * (32 * KPL + 1);
  __shared__ __align__(8) uint64_t bar_full[MM_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[MM_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_tmem_full[2];
  __shared__ __align__(8) uint64_t bar_tmem_empty[2];
  __shared__ __align__(8) uint64_t bar_a_ready;
  __shared__ uint32_t s_tmem_base;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.n_stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_tmem_full[b]), 1);
      mbar_init(smem_u32(&bar_tmem_empty[b]), 8);
    }
    mbar_init(smem_u32(&bar_a_ready), 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_e)) : "memory");
    if (!A_IN_TMEM) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_q)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  const uint32_t acc_col0 = static_cast<uint32_t>(a_cols);

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      volatile int* group_progress = a.progress + group * n_slabs;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        // Soft pacing: the blocks of a group stream the SAME passage tiles for different query slabs.
        // If one runs far ahead, the tiles it pulled into L2 are gone by the time the others arrive
        // and HBM traffic multiplies (measured 3.4x).  Every 4 tiles a block publishes its position
        // and briefly waits for the slowest slab; the wait is bounded, so it can never deadlock.
        const int it = tile - tile_begin;
        if (n_slabs > 1 && (it & a.pace_mask) == 0) {
          group_progress[slab] = it;
          for (int spins = 0; spins < 256; ++spins) {
            int slowest = it;
            for (int sl = 0; sl < n_slabs; ++sl) slowest = min(slowest, group_progress[sl]);
            if (it - slowest <= a.pace_tiles) break;
            __nanosleep(256);
          }
        }
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
          const uint32_t full = smem_u32(&bar_full[stage]);
          unsigned char* st = stages + static_cast<size_t>(stage) * STAGE_BYTES;
          mbar_expect_tx(full, STAGE_BYTES);
          if (!A_IN_TMEM) {
            tma_load_2d(smem_u32(st), &tmap_q, full, kb * MM_BK, slab * MM_BM);
            tma_load_2d(smem_u32(st + MM_A_STAGE_BYTES), &tmap_e, full, kb * MM_BK, tile * BN);
          } else {
            tma_load_2d(smem_u32(st), &tmap_e, full, kb * MM_BK, tile * BN);
          }
          if (++stage == static_cast<uint32_t>(a.n_stages)) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (n_slabs > 1) group_progress[slab] = 0x7fffffff;  // done: nobody waits for this block any more
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      if (A_IN_TMEM) {
        mbar_wait(smem_u32(&bar_a_ready), 0);
        tc_fence_after();
      }
      uint32_t stage = 0, phase = 0, buf = 0, acc_phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        mbar_wait(smem_u32(&bar_tmem_empty[buf]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc_col0 + buf * BN;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          unsigned char* st = stages + static_cast<size_t>(stage) * STAGE_BYTES;
          const uint32_t a_addr = smem_u32(st);
          const uint32_t b_addr = smem_u32(st + (A_IN_TMEM ? 0 : MM_A_STAGE_BYTES));
#pragma unroll
          for (int kk = 0; kk < MM_BK / 16; ++kk) {
            const uint32_t acc = (kb | kk) != 0 ? 1u : 0u;
            const uint64_t b_desc = make_sw128_desc(b_addr + kk * 32);
            if (A_IN_TMEM) {
              tc_mma_ts(d_tmem, tmem_base + kb * (MM_BK / 2) + kk * 8, b_desc, IDESC, acc);
            } else {
              tc_mma_ss(d_tmem, make_sw128_desc(a_addr + kk * 32), b_desc, IDESC, acc);
            }
          }
          tc_commit(smem_u32(&bar_empty[stage]));  // frees the stage once these MMAs retire
          if (++stage == static_cast<uint32_t>(a.n_stages)) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(smem_u32(&bar_tmem_full[buf]));
        buf ^= 1;
        if (buf == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ================= epilogue: thread owns one query =================
    if (A_IN_TMEM && warp < 6) {
      // store the thread's query row (packed bf16 pairs, K ascending) into its TMEM lane
      const int quad = warp & 3;
      const int query = slab * MM_BM + quad * 32 + lane;
      const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
      const uint4* qrow = a.queries + static_cast<int64_t>(query) * (a.dim / 8);
      for (int c = 0; c < a_cols / 32; ++c) {
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint4 x = make_uint4(0u, 0u, 0u, 0u);
          if (query < a.n_queries) x = __ldg(qrow + c * 8 + i);
          v[4 * i] = x.x;
          v[4 * i + 1] = x.y;
          v[4 * i + 2] = x.z;
          v[4 * i + 3] = x.w;
        }
        tc_st32(tmem_base + lane_addr + c * 32, v);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_a_ready));
    }
    if constexpr (FUSED)
      epilogue_scan_fused<BN, KPL>(a, s_table, tmem_base, acc_col0, lists, warp, lane, slab, group, tile_begin, tile_end,
                                   bar_tmem_full, smem_u32(&bar_tmem_empty[0]), smem_u32(&bar_tmem_empty[1]), false);
    else
      epilogue_scan<BN, KPL>(a, tmem_base, acc_col0, lists, warp, lane, slab, group, tile_begin, tile_end, bar_tmem_full,
                             smem_u32(&bar_tmem_empty[0]), smem_u32(&bar_tmem_empty[1]), false);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// =============================================================================================
// variant 3: CTA pair (cta_group::2).  Two blocks of a cluster form one 256 x 256 MMA tile: each
// holds the A rows of ITS 128-query slab and HALF of the passage tile (128 rows) in shared memory,
// the leader's single thread issues tcgen05.mma.cta_group::2 (M = 256), each block's TMEM receives
// its 128 x 256 half of the accumulator.  Per SM this moves 32 KB per k-block instead of 48 KB, so
// the same shared memory holds a 7-deep ring instead of 4 and L2 -> SM traffic per FLOP drops by a
// third.  TMA loads of both blocks complete on the LEADER's full barrier; tcgen05.commit multicasts
// "stage free" / "accumulator ready" to both blocks; both epilogues release the accumulator with a
// remote arrive on the leader's barrier.
// =============================================================================================
constexpr uint32_t MM_PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address
constexpr int MM2_BN = 256;
constexpr int MM2_STAGE_BYTES = MM_A_STAGE_BYTES + (MM2_BN / 2) * MM_BK * 2;  // 32 KB per block

__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void tc_mma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int KPL, bool FUSED = false>
__global__ void __launch_bounds__(MM_THREADS, 1) dense_mma_pair_kernel(const __grid_constant__ CUtensorMap tmap_q,
                                                                       const __grid_constant__ CUtensorMap tmap_e,
                                                                       const MmaArgs a) {
  constexpr int BN = MM2_BN;
  constexpr uint32_t IDESC = make_idesc(2 * MM_BM, BN);
  uint32_t cta_rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const bool leader = cta_rank == 0;
  const int pair = blockIdx.x >> 1;                 // cluster index
  const int n_pairs_slab = (a.n_slabs + 1) >> 1;    // slab pairs per passage group
  // every block of the grid belongs to a live pair (the host sizes the grid exactly)
  const int slab = 2 * (pair % n_pairs_slab) + static_cast<int>(cta_rank);
  const int group = pair / n_pairs_slab;
  const int tile_begin = min(a.tile_hi, a.tile_lo + group * a.tiles_per_group);
  const int tile_end = min(a.tile_hi, tile_begin + a.tiles_per_group);
  const int n_kb = a.dim / MM_BK;

  extern __shared__ unsigned char smem_dyn[];
  unsigned char* stages = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint32_t* s_table = reinterpret_cast<uint32_t*>(stages + static_cast<size_t>(a.n_stages) * MM2_STAGE_BYTES);  // FUSED only
  if constexpr (FUSED) {
    for (int i = threadIdx.x; i < a.ff_nb * a.ff_nd; i += MM_THREADS) s_table[i] = __ldg(a.ff_table + i);
  }
  uint64_t* lists = a.lists + static_cast<size_t>(blockIdx.x) * 2 * MM_BM * (32 * KPL + 1);
  __shared__ __align__(8) uint64_t bar_full[MM_MAX_STAGES];   // used in the leader only
  __shared__ __align__(8) uint64_t bar_empty[MM_MAX_STAGES];  // one per block, fed by the multicast commit
  __shared__ __align__(8) uint64_t bar_tmem_full[2];          // one per block
  __shared__ __align__(8) uint64_t bar_tmem_empty[2];         // leader only: 8 epilogue warps arrive
  __shared__ uint32_t s_tmem_base;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.n_stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_tmem_full[b]), 1);
      mbar_init(smem_u32(&bar_tmem_empty[b]), 16);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_e)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_q)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of both blocks initialised, TMEM allocated in both
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 0) {
    // ================= TMA producer (both blocks; bytes complete on the leader's barrier) =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      volatile int* group_progress = a.progress + group * n_pairs_slab;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int it = tile - tile_begin;
        if (leader && n_pairs_slab > 1 && (it & a.pace_mask) == 0) {   // soft pacing between the pairs of a group
          group_progress[pair % n_pairs_slab] = it;
          for (int spins = 0; spins < 256; ++spins) {
            int slowest = it;
            for (int sl = 0; sl < n_pairs_slab; ++sl) slowest = min(slowest, group_progress[sl]);
            if (it - slowest <= a.pace_tiles) break;
            __nanosleep(256);
          }
        }
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
          const uint32_t leader_full = smem_u32(&bar_full[stage]) & MM_PEER_MASK;
          unsigned char* st = stages + static_cast<size_t>(stage) * MM2_STAGE_BYTES;
          if (leader) mbar_expect_tx(smem_u32(&bar_full[stage]), 2 * MM2_STAGE_BYTES);
          tma_load_2d_2sm(smem_u32(st), &tmap_q, leader_full, kb * MM_BK, slab * MM_BM);
          tma_load_2d_2sm(smem_u32(st + MM_A_STAGE_BYTES), &tmap_e, leader_full, kb * MM_BK,
                          tile * BN + static_cast<int>(cta_rank) * (BN / 2));
          if (++stage == static_cast<uint32_t>(a.n_stages)) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (leader && n_pairs_slab > 1) group_progress[pair % n_pairs_slab] = 0x7fffffff;
    }
  } else if (warp == 1) {
    // ================= MMA issuer: one thread of the leader drives both tensor cores =================
    if (leader && lane == 0) {
      uint32_t stage = 0, phase = 0, buf = 0, acc_phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        mbar_wait(smem_u32(&bar_tmem_empty[buf]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          unsigned char* st = stages + static_cast<size_t>(stage) * MM2_STAGE_BYTES;
          const uint32_t a_addr = smem_u32(st);
          const uint32_t b_addr = smem_u32(st + MM_A_STAGE_BYTES);
#pragma unroll
          for (int kk = 0; kk < MM_BK / 16; ++kk)
            tc_mma_ss_pair(d_tmem, make_sw128_desc(a_addr + kk * 32), make_sw128_desc(b_addr + kk * 32), IDESC,
                           (kb | kk) != 0 ? 1u : 0u);
          tc_commit_pair(smem_u32(&bar_empty[stage]));   // frees the stage in BOTH blocks
          if (++stage == static_cast<uint32_t>(a.n_stages)) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit_pair(smem_u32(&bar_tmem_full[buf]));   // accumulator ready in BOTH blocks
        buf ^= 1;
        if (buf == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ================= epilogue: thread owns one query of this block's slab =================
    // both blocks release the accumulator on the LEADER's barrier
    if constexpr (FUSED)
      epilogue_scan_fused<MM2_BN, KPL>(a, s_table, tmem_base, 0u, lists, warp, lane, slab, group, tile_begin, tile_end,
                                       bar_tmem_full, smem_u32(&bar_tmem_empty[0]) & MM_PEER_MASK,
                                       smem_u32(&bar_tmem_empty[1]) & MM_PEER_MASK, true);
    else
      epilogue_scan<MM2_BN, KPL>(a, tmem_base, 0u, lists, warp, lane, slab, group, tile_begin, tile_end, bar_tmem_full,
                                 smem_u32(&bar_tmem_empty[0]) & MM_PEER_MASK, smem_u32(&bar_tmem_empty[1]) & MM_PEER_MASK,
                                 true);
  }

  tc_fence_before();
  cluster_sync_all();   // nobody frees TMEM or exits while the peer may still touch this block
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- host ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// [rows, dim] bf16 row-major -> 2-D map, box = 64 columns (128 bytes) x box_rows, 128B swizzle
static int make_map(CUtensorMap* map, const void* ptr, int64_t rows, int dim, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  RAGB_REQUIRE(fn, RAGB_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(dim), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(dim) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(MM_BK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t elem[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, elem,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RAGB_REQUIRE(r == CUDA_SUCCESS, RAGB_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return RAGB_OK;
}

static int mma_pace_mask() {
  static const int v = [] {
    const char* e = getenv("RAGB_MMA_PUBLISH");  // tuning aid: 1, 2 or 4 tiles between position updates
    const int n = e ? atoi(e) : 4;
    return n <= 1 ? 0 : (n == 2 ? 1 : 3);
  }();
  return v;
}

static int mma_pace_tiles() {
  static const int v = [] {
    const char* e = getenv("RAGB_MMA_PACE");  // tuning aid
    return e ? atoi(e) : MM_PACE_TILES;
  }();
  return v;
}

// fused: null (dense-only top-k) or the full-fusion fields of MmaArgs (bm25, router, bound table)
static void copy_fused_fields(MmaArgs& a, const MmaArgs* fused) {
  if (fused == nullptr) return;
  a.bm25 = fused->bm25;
  a.bm25_ld = fused->bm25_ld;
  a.rw = fused->rw;
  a.ff_table = fused->ff_table;
  a.ff_nb = fused->ff_nb;
  a.ff_nd = fused->ff_nd;
  a.ff_inv_wb = fused->ff_inv_wb;
  a.ff_inv_wd = fused->ff_inv_wd;
  a.ff_d_hi = fused->ff_d_hi;
  a.counters = fused->counters;
  a.ff_debug = fused->ff_debug;
}

// Passage groups of one launch over `tiles` tiles: as many as the SMs (or CTA pairs) left after one block per query
// slab (pair) allow.  Pure function of its arguments: the sampled and the seeded phase and the workspace layout
// all derive their list counts from it.
struct MmaPlan {
  int n_slabs, n_groups, tiles_per_group, blocks;
};
static MmaPlan mma_plan(bool pair, int n_queries, int tiles, int sms) {
  MmaPlan p{};
  p.n_slabs = (n_queries + MM_BM - 1) / MM_BM;
  const int units = pair ? (p.n_slabs + 1) / 2 : p.n_slabs;      // blocks (or CTA pairs) per passage group
  const int slots = pair ? sms / 2 : sms;
  p.n_groups = units > 0 ? slots / units : 0;
  if (p.n_groups > tiles) p.n_groups = tiles;
  if (p.n_groups < 1) p.n_groups = 1;
  p.tiles_per_group = (tiles + p.n_groups - 1) / p.n_groups;
  p.n_groups = (tiles + p.tiles_per_group - 1) / p.tiles_per_group;
  p.blocks = (pair ? 2 * units : units) * p.n_groups;
  return p;
}

// what one launch covers and where its lists go
struct MmaRange {
  int tile_lo, tile_hi;       // passage tiles [tile_lo, tile_hi)
  const float* seed_thr;      // optional [n_queries] proven lower bounds of the k-th best score
  float* min_out;             // optional [n_queries] running minimum of all scores (see ragb_dense_mma_topk_min)
  uint64_t* part;             // [n_queries, lists_per_query, k]
  int lists_per_query;        // >= 2 * groups of this launch
};

template <int BN, bool A_IN_TMEM, int KPL, bool FUSED = false>
static int launch_mma(const void* passages, int64_t n_rows, int dim, const void* queries, int n_queries, int k,
                      int64_t id_base, const MmaRange& r, int* progress, uint64_t* lists, int stage_limit,
                      cudaStream_t stream, const MmaArgs* fused = nullptr) {
  constexpr int STAGE_BYTES = (A_IN_TMEM ? 0 : MM_A_STAGE_BYTES) + BN * MM_BK * 2;
  CUtensorMap map_q, map_e;
  int rc = make_map(&map_q, queries, n_queries, dim, MM_BM);
  if (rc != RAGB_OK) return rc;
  rc = make_map(&map_e, passages, n_rows, dim, BN);
  if (rc != RAGB_OK) return rc;

  const int sms = device_sm_count();
  const MmaPlan plan = mma_plan(false, n_queries, r.tile_hi - r.tile_lo, sms);
  RAGB_REQUIRE(plan.n_slabs <= sms, RAGB_ELIMIT, "ragb_dense_mma_topk: n_queries=%d needs more than %d slabs of 128",
               n_queries, sms);
  RAGB_REQUIRE(2 * plan.n_groups <= r.lists_per_query, RAGB_EINVAL, "ragb_dense_mma_topk: internal list layout mismatch");
  MmaArgs a{};
  a.queries = static_cast<const uint4*>(queries);
  a.n_rows = n_rows;
  a.id_base = id_base;
  a.n_queries = n_queries;
  a.dim = dim;
  a.k = k;
  a.n_slabs = plan.n_slabs;
  a.n_groups = plan.n_groups;
  a.tiles_per_group = plan.tiles_per_group;
  a.tile_lo = r.tile_lo;
  a.tile_hi = r.tile_hi;
  a.seed_thr = r.seed_thr;
  a.min_out = r.min_out;
  a.part_keys = r.part;
  a.lists_per_query = r.lists_per_query;
  a.progress = progress;
  a.lists = lists;
  copy_fused_fields(a, fused);
  a.pace_tiles = mma_pace_tiles();
  a.pace_mask = mma_pace_mask();
  RAGB_CUDA(cudaMemsetAsync(progress, 0, MM_PROGRESS_BYTES, stream));
  // Ring depth: everything shared memory offers (RAGB_MMA_STAGES caps it, e.g. to leave room for
  // blocks of another kernel on the same SM when experimenting with two-stream overlap).
  constexpr int TABLE_BYTES = FUSED ? FF_TABLE_CELLS * static_cast<int>(sizeof(uint32_t)) : 0;
  int stages = static_cast<int>((MM_MAX_SMEM - 1024 - 256 - TABLE_BYTES) / STAGE_BYTES);
  if (stage_limit > 0 && stages > stage_limit) stages = stage_limit;
  if (stages > MM_MAX_STAGES) stages = MM_MAX_STAGES;
  a.n_stages = stages;
  const size_t smem = 1024 + static_cast<size_t>(stages) * STAGE_BYTES + TABLE_BYTES;
  RAGB_CUDA(cudaFuncSetAttribute(dense_mma_kernel<BN, A_IN_TMEM, KPL, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
  RAGB_CUDA(cudaFuncSetAttribute(dense_mma_kernel<BN, A_IN_TMEM, KPL, FUSED>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
  dense_mma_kernel<BN, A_IN_TMEM, KPL, FUSED><<<plan.blocks, MM_THREADS, smem, stream>>>(map_q, map_e, a);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

template <int KPL, bool FUSED = false>
static int launch_mma_pair(const void* passages, int64_t n_rows, int dim, const void* queries, int n_queries, int k,
                           int64_t id_base, const MmaRange& r, int* progress, uint64_t* lists, int stage_limit,
                           cudaStream_t stream, const MmaArgs* fused = nullptr) {
  CUtensorMap map_q, map_e;
  int rc = make_map(&map_q, queries, n_queries, dim, MM_BM);
  if (rc != RAGB_OK) return rc;
  rc = make_map(&map_e, passages, n_rows, dim, MM2_BN / 2);   // each block loads half a passage tile
  if (rc != RAGB_OK) return rc;
  const int sms = device_sm_count();
  const MmaPlan plan = mma_plan(true, n_queries, r.tile_hi - r.tile_lo, sms);
  RAGB_REQUIRE((plan.n_slabs + 1) / 2 <= sms / 2, RAGB_ELIMIT, "ragb_dense_mma_topk: n_queries=%d needs more than %d CTA pairs",
               n_queries, sms / 2);
  RAGB_REQUIRE(2 * plan.n_groups <= r.lists_per_query, RAGB_EINVAL, "ragb_dense_mma_topk: internal list layout mismatch");
  MmaArgs a{};
  a.queries = static_cast<const uint4*>(queries);
  a.n_rows = n_rows;
  a.id_base = id_base;
  a.n_queries = n_queries;
  a.dim = dim;
  a.k = k;
  a.n_slabs = plan.n_slabs;
  a.n_groups = plan.n_groups;
  a.tiles_per_group = plan.tiles_per_group;
  a.tile_lo = r.tile_lo;
  a.tile_hi = r.tile_hi;
  a.seed_thr = r.seed_thr;
  a.min_out = r.min_out;
  a.part_keys = r.part;
  a.lists_per_query = r.lists_per_query;
  a.progress = progress;
  a.lists = lists;
  copy_fused_fields(a, fused);
  a.pace_tiles = mma_pace_tiles();
  a.pace_mask = mma_pace_mask();
  RAGB_CUDA(cudaMemsetAsync(progress, 0, MM_PROGRESS_BYTES, stream));
  constexpr int TABLE_BYTES = FUSED ? FF_TABLE_CELLS * static_cast<int>(sizeof(uint32_t)) : 0;
  int stages = (MM_MAX_SMEM - 1024 - 256 - TABLE_BYTES) / MM2_STAGE_BYTES;
  if (stage_limit > 0 && stages > stage_limit) stages = stage_limit;
  if (stages > MM_MAX_STAGES) stages = MM_MAX_STAGES;
  a.n_stages = stages;
  const size_t smem = 1024 + static_cast<size_t>(stages) * MM2_STAGE_BYTES + TABLE_BYTES;
  RAGB_CUDA(cudaFuncSetAttribute(dense_mma_pair_kernel<KPL, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
  // the SM keeps its maximal shared-memory carve-out even when this kernel asks for less (RAGB_MMA_STAGES), so
  // blocks of another kernel (BM25 on a second stream) can become co-resident without reconfiguring the SM
  RAGB_CUDA(cudaFuncSetAttribute(dense_mma_pair_kernel<KPL, FUSED>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(plan.blocks);
  cfg.blockDim = dim3(MM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  // per-launch form of the carve-out preference (takes precedence over the function attribute): keep the SM at its
  // maximal shared-memory configuration so that a block of another kernel can be placed next to this one
  attr[1].id = cudaLaunchAttributePreferredSharedMemoryCarveout;
  attr[1].val.sharedMemCarveout = 100;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  RAGB_CUDA(cudaLaunchKernelEx(&cfg, dense_mma_pair_kernel<KPL, FUSED>, map_q, map_e, a));
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

static int mma_list_kpl(int k) {   // capacity 32*KPL >= k + 32
  static const int forced = [] {
    const char* e = getenv("RAGB_MMA_KPL");  // tuning aid only: a larger list is compacted less often
    return e ? atoi(e) : 0;
  }();
  const int kpl = k <= 32 ? 2 : (k <= 96 ? 4 : 8);
  return (forced == 4 || forced == 8) && forced > kpl ? forced : kpl;
}
static size_t mma_list_bytes(int k) {
  return static_cast<size_t>(148) * 2 * MM_BM * (32 * mma_list_kpl(k) + 1) * sizeof(uint64_t);
}
constexpr int MM_MAX_LISTS = 148 * 2;   // lists per query of one launch, at most (one group per SM, two column halves)

static int mma_stage_limit() {
  static const int v = [] {
    const char* e = getenv("RAGB_MMA_STAGES");  // tuning aid / overlap: leaves shared memory for co-resident blocks
    return e ? atoi(e) : 0;
  }();
  return v;
}

// Sampled prefix: the first 1/32 of the tiles is searched un-seeded; the k-th best score found there is a proven
// lower bound of the final k-th best, and every list of the remaining 31/32 starts from it.  Without it each of
// the 2 x groups lists of a query warms up from -inf on its own: 36 x k (1 + ln(n / k)) appends and a sort per
// (capacity - k) of them - 1.2 ms of the 2.8 ms at 1.25M rows and k = 50.  With it a query admits ~31 k candidates
// in total, spread over all its lists, and almost no list is compacted before the end.  Small shards skip the prefix.
static int mma_sample_tiles(int n_tiles) {
  static const int div = [] {
    const char* e = getenv("RAGB_MMA_SAMPLE_DIV");  // tuning aid: 0 = no sampled prefix
    return e ? atoi(e) : 32;
  }();
  if (div <= 0 || n_tiles < 512) return 0;
  return n_tiles / div;
}

// How the sampled prefix is searched.  On a large shard (mode 2) it is searched exactly, its merged list joins the final
// merge and the seeded phase covers only the remaining tiles: no tile is searched twice.  But un-seeded lists pay a fixed
// warm-up - four or five whole-warp sorts per query thread, ~0.4 ms per launch whatever the number of tiles - which on a
// small shard (1.25M rows: 1.5 ms of seeded search) is a quarter of the kernel.  There (mode 1) the prefix only ESTABLISHES
// the bound: every list keeps just its 8 best (tiny lists, one cheap sort), the k-th best of the 36 x 8 merged candidates is
// still a proven lower bound of the final k-th best (k real passages reach it) and in practice the sample's true k-th best;
// the seeded phase then covers ALL tiles (3 % more FLOPs) and the prefix list is not merged again.
static int mma_prefix_mode(int n_tiles) {
  if (mma_sample_tiles(n_tiles) == 0) return 0;
  static const int exact_from = [] {
    const char* e = getenv("RAGB_MMA_EXACT_PREFIX_TILES");  // tuning aid: shards of at least this many tiles use mode 2
    return e ? atoi(e) : 32768;   // measured (k = 50, 1024 queries): 5M rows 6.94 ms exact / 6.59 ms bound-only, 10M rows 13.33 / 13.42
  }();
  return n_tiles >= exact_from ? 2 : 1;
}
constexpr int MM_BOUND_ONLY_K = 8;

struct MmaWorkspace {
  int* progress;
  uint64_t* lists;
  uint64_t* sample_keys;   // [n_queries, k] merged result of the sampled prefix
  uint64_t* part;          // [n_queries, <= MM_MAX_LISTS, k]
};
static size_t mma_workspace_bytes(int n_queries, int k) {
  return MM_PROGRESS_BYTES + mma_list_bytes(k) + static_cast<size_t>(n_queries) * (MM_MAX_LISTS + 1) * k * sizeof(uint64_t);
}
static MmaWorkspace mma_carve(void* workspace, int n_queries, int k) {
  unsigned char* p = static_cast<unsigned char*>(workspace);
  MmaWorkspace w;
  w.progress = reinterpret_cast<int*>(p);
  w.lists = reinterpret_cast<uint64_t*>(p + MM_PROGRESS_BYTES);
  w.sample_keys = reinterpret_cast<uint64_t*>(p + MM_PROGRESS_BYTES + mma_list_bytes(k));
  w.part = w.sample_keys + static_cast<size_t>(n_queries) * k;
  return w;
}

// one un-fused launch over a tile range with the variant / list-capacity dispatch
static int mma_dispatch(int variant, const void* passages, int64_t n_rows, int dim, const void* queries, int n_queries,
                        int k, int64_t id_base, const MmaRange& r, const MmaWorkspace& w, cudaStream_t stream) {
  const int kpl = mma_list_kpl(k);
  // variant 4 = variant 3 with a 4-stage operand ring (129 KB instead of 225 KB of shared memory per block): measured
  // as fast as the full ring, and it leaves room for one block of another kernel on the same SM (two-stream overlap)
  const int sl = variant == 4 ? 4 : mma_stage_limit();
#define RAGB_MMA_ARGS passages, n_rows, dim, queries, n_queries, k, id_base, r, w.progress, w.lists, sl, stream
#define RAGB_MMA_DISPATCH(FN, ...)                                         \
  do {                                                                    \
    if (kpl == 2) return FN<__VA_ARGS__ 2>(RAGB_MMA_ARGS);                \
    if (kpl == 4) return FN<__VA_ARGS__ 4>(RAGB_MMA_ARGS);                \
    return FN<__VA_ARGS__ 8>(RAGB_MMA_ARGS);                              \
  } while (0)
#define RAGB_COMMA ,
  if (variant == 0) RAGB_MMA_DISPATCH(launch_mma, 128 RAGB_COMMA false RAGB_COMMA);
  if (variant == 1) RAGB_MMA_DISPATCH(launch_mma, 64 RAGB_COMMA true RAGB_COMMA);
  if (variant == 2 || n_queries <= MM_BM) RAGB_MMA_DISPATCH(launch_mma, 256 RAGB_COMMA false RAGB_COMMA);   // a lone slab has no partner
  RAGB_MMA_DISPATCH(launch_mma_pair, );
#undef RAGB_MMA_DISPATCH
#undef RAGB_COMMA
#undef RAGB_MMA_ARGS
}
static int mma_tile_rows(int variant) { return variant == 0 ? 128 : (variant == 1 ? 64 : 256); }
static bool mma_uses_pairs(int variant, int n_queries) { return variant >= 3 && n_queries > MM_BM; }

static int mma_common_checks(const char* who, const void* passages_bf16, int64_t n_rows, int32_t dim, const void* queries_bf16,
                             int32_t n_queries, int32_t k, int64_t id_base, int32_t variant, const void* workspace,
                             size_t workspace_bytes) {
  RAGB_REQUIRE(passages_bf16 && queries_bf16 && workspace, RAGB_EINVAL, "%s: null pointer", who);
  RAGB_REQUIRE(((reinterpret_cast<uintptr_t>(passages_bf16) | reinterpret_cast<uintptr_t>(queries_bf16)) & 15) == 0,
               RAGB_EINVAL, "%s: inputs must be 16-byte aligned", who);
  RAGB_REQUIRE(n_rows > 0 && n_queries > 0, RAGB_EINVAL, "%s: empty shape", who);
  RAGB_REQUIRE(dim >= MM_BK && dim % MM_BK == 0, RAGB_EINVAL, "%s: dim=%d must be a multiple of %d", who, dim, MM_BK);
  RAGB_REQUIRE(k > 0 && k <= 100, RAGB_ELIMIT, "%s: k=%d outside [1,100]", who, k);
  RAGB_REQUIRE(variant >= 0 && variant <= 4, RAGB_EINVAL, "%s: variant must be 0, 1, 2, 3 or 4", who);
  RAGB_REQUIRE(variant != 1 || dim <= 768, RAGB_ELIMIT, "%s: variant 1 keeps the query slab in TMEM and needs dim <= 768", who);
  RAGB_REQUIRE(id_base >= 0 && id_base + n_rows < (1ll << 31), RAGB_ELIMIT, "%s: ids must fit int32", who);
  RAGB_REQUIRE(workspace_bytes >= mma_workspace_bytes(n_queries, k), RAGB_ENOSPC, "%s: workspace too small", who);
  return RAGB_OK;
}

// phase 1: search the sampled prefix, leave its merged list in the workspace, report the per-query k-th best score
static int mma_sample_phase(const void* passages, int64_t n_rows, int dim, const void* queries, int n_queries, int k,
                            int64_t id_base, int variant, float* thr_out, const MmaWorkspace& w, cudaStream_t stream,
                            float* min_out = nullptr) {
  const int n_tiles = static_cast<int>(ceil_div64(n_rows, mma_tile_rows(variant)));
  const int ts = mma_sample_tiles(n_tiles);
  const int mode = mma_prefix_mode(n_tiles);
  if (mode == 0)   // no prefix: merging zero lists leaves an empty sample list and the bound -inf for every query
    return launch_merge_keys_ex(w.part, n_queries, 0, k, nullptr, 0, k, nullptr, nullptr, w.sample_keys, thr_out, stream);
  const int kp = (mode == 1 && k > MM_BOUND_ONLY_K) ? MM_BOUND_ONLY_K : k;   // bound-only prefix: short lists
  const MmaPlan plan = mma_plan(mma_uses_pairs(variant, n_queries), n_queries, ts, device_sm_count());
  const MmaRange r{0, ts, nullptr, mode == 2 ? min_out : nullptr, w.part, 2 * plan.n_groups};
  int rc = mma_dispatch(variant, passages, n_rows, dim, queries, n_queries, kp, id_base, r, w, stream);
  if (rc != RAGB_OK) return rc;
  // the k-th best of the merged candidates (-inf while there are fewer than k) is the bound in either mode
  return launch_merge_keys_ex(w.part, n_queries, 2 * plan.n_groups, kp, nullptr, 0, k, nullptr, nullptr, w.sample_keys, thr_out,
                              stream);
}

// phase 2: the rest of the tiles, every list seeded with thr (own or exchanged between shards), final merge
static int mma_seeded_phase(const void* passages, int64_t n_rows, int dim, const void* queries, int n_queries, int k,
                            int64_t id_base, int variant, const float* thr, float* out_score, int32_t* out_id,
                            const MmaWorkspace& w, cudaStream_t stream, float* min_out = nullptr) {
  const int n_tiles = static_cast<int>(ceil_div64(n_rows, mma_tile_rows(variant)));
  const int mode = mma_prefix_mode(n_tiles);
  const int first = mode == 2 ? mma_sample_tiles(n_tiles) : 0;    // modes 0 / 1: the seeded phase covers every tile
  const MmaPlan plan = mma_plan(mma_uses_pairs(variant, n_queries), n_queries, n_tiles - first, device_sm_count());
  const MmaRange r{first, n_tiles, thr, min_out, w.part, 2 * plan.n_groups};
  int rc = mma_dispatch(variant, passages, n_rows, dim, queries, n_queries, k, id_base, r, w, stream);
  if (rc != RAGB_OK) return rc;
  return launch_merge_keys_ex(w.part, n_queries, 2 * plan.n_groups, k, mode == 2 ? w.sample_keys : nullptr, k, k, out_score,
                              out_id, nullptr, nullptr, stream);
}

}  // namespace ragb

using namespace ragb;

extern "C" {

size_t ragb_dense_mma_workspace_bytes(int32_t n_queries, int32_t k) {
  if (n_queries <= 0 || k <= 0) return 0;
  return mma_workspace_bytes(n_queries, k) + static_cast<size_t>(n_queries) * sizeof(float);   // + own thresholds
}

int ragb_dense_mma_topk(const void* passages_bf16, int64_t n_rows, int32_t dim, const void* queries_bf16,
                        int32_t n_queries, int32_t k, int64_t id_base, int32_t variant, float* out_score,
                        int32_t* out_id, void* workspace, size_t workspace_bytes, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(out_score && out_id, RAGB_EINVAL, "ragb_dense_mma_topk: null pointer");
  RAGB_REQUIRE(workspace_bytes >= ragb_dense_mma_workspace_bytes(n_queries, k), RAGB_ENOSPC, "ragb_dense_mma_topk: workspace too small");
  int rc = mma_common_checks("ragb_dense_mma_topk", passages_bf16, n_rows, dim, queries_bf16, n_queries, k, id_base, variant,
                             workspace, workspace_bytes);
  if (rc != RAGB_OK) return rc;
  const MmaWorkspace w = mma_carve(workspace, n_queries, k);
  float* thr = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) + mma_workspace_bytes(n_queries, k));
  rc = mma_sample_phase(passages_bf16, n_rows, dim, queries_bf16, n_queries, k, id_base, variant, thr, w, stream);
  if (rc != RAGB_OK) return rc;
  return mma_seeded_phase(passages_bf16, n_rows, dim, queries_bf16, n_queries, k, id_base, variant, thr, out_score, out_id, w,
                          stream);
}

int ragb_dense_mma_topk_min(const void* passages_bf16, int64_t n_rows, int32_t dim, const void* queries_bf16,
                            int32_t n_queries, int32_t k, int64_t id_base, int32_t variant, float* out_score,
                            int32_t* out_id, float* min_inout, void* workspace, size_t workspace_bytes,
                            ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(out_score && out_id && min_inout, RAGB_EINVAL, "ragb_dense_mma_topk_min: null pointer");
  RAGB_REQUIRE(workspace_bytes >= ragb_dense_mma_workspace_bytes(n_queries, k), RAGB_ENOSPC, "ragb_dense_mma_topk_min: workspace too small");
  int rc = mma_common_checks("ragb_dense_mma_topk_min", passages_bf16, n_rows, dim, queries_bf16, n_queries, k, id_base, variant,
                             workspace, workspace_bytes);
  if (rc != RAGB_OK) return rc;
  const MmaWorkspace w = mma_carve(workspace, n_queries, k);
  float* thr = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) + mma_workspace_bytes(n_queries, k));
  rc = mma_sample_phase(passages_bf16, n_rows, dim, queries_bf16, n_queries, k, id_base, variant, thr, w, stream, min_inout);
  if (rc != RAGB_OK) return rc;
  return mma_seeded_phase(passages_bf16, n_rows, dim, queries_bf16, n_queries, k, id_base, variant, thr, out_score, out_id, w,
                          stream, min_inout);
}

int ragb_dense_mma_sample(const void* passages_bf16, int64_t n_rows, int32_t dim, const void* queries_bf16,
                          int32_t n_queries, int32_t k, int64_t id_base, int32_t variant, float* thr_out,
                          void* workspace, size_t workspace_bytes, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(thr_out, RAGB_EINVAL, "ragb_dense_mma_sample: null pointer");
  int rc = mma_common_checks("ragb_dense_mma_sample", passages_bf16, n_rows, dim, queries_bf16, n_queries, k, id_base, variant,
                             workspace, workspace_bytes);
  if (rc != RAGB_OK) return rc;
  return mma_sample_phase(passages_bf16, n_rows, dim, queries_bf16, n_queries, k, id_base, variant, thr_out,
                          mma_carve(workspace, n_queries, k), stream);
}

int ragb_dense_mma_seeded(const void* passages_bf16, int64_t n_rows, int32_t dim, const void* queries_bf16,
                          int32_t n_queries, int32_t k, int64_t id_base, int32_t variant, const float* thr,
                          float* out_score, int32_t* out_id, void* workspace, size_t workspace_bytes,
                          ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(thr && out_score && out_id, RAGB_EINVAL, "ragb_dense_mma_seeded: null pointer");
  int rc = mma_common_checks("ragb_dense_mma_seeded", passages_bf16, n_rows, dim, queries_bf16, n_queries, k, id_base, variant,
                             workspace, workspace_bytes);
  if (rc != RAGB_OK) return rc;
  return mma_seeded_phase(passages_bf16, n_rows, dim, queries_bf16, n_queries, k, id_base, variant, thr, out_score, out_id,
                          mma_carve(workspace, n_queries, k), stream);
}

int ragb_dense_mma_fused_topk(const void* passages_bf16, int64_t n_rows, int32_t dim, const void* queries_bf16,
                              int32_t n_queries, int32_t k, int64_t id_base, const float* bm25_scores, int64_t bm25_rows,
                              const float* w1, const float* b1, const float* w2, const float* b2, const float* stats,
                              int32_t hidden, const uint32_t* gate_bound_table, int32_t n_b, int32_t n_d, float b_cap,
                              float d_hi, float* out_score, int32_t* out_id, unsigned long long* counters,
                              void* workspace, size_t workspace_bytes, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(passages_bf16 && queries_bf16 && bm25_scores && out_score && out_id && workspace, RAGB_EINVAL,
               "ragb_dense_mma_fused_topk: null pointer");
  RAGB_REQUIRE(w1 && b1 && w2 && b2 && stats && gate_bound_table, RAGB_EINVAL,
               "ragb_dense_mma_fused_topk: router weights, running statistics and gate bound table are required");
  RAGB_REQUIRE(((reinterpret_cast<uintptr_t>(passages_bf16) | reinterpret_cast<uintptr_t>(queries_bf16)) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(bm25_scores) & 31) == 0,
               RAGB_EINVAL, "ragb_dense_mma_fused_topk: passages / queries must be 16-byte, bm25_scores 32-byte aligned");
  RAGB_REQUIRE(n_rows > 0 && n_queries > 0, RAGB_EINVAL, "ragb_dense_mma_fused_topk: empty shape");
  RAGB_REQUIRE(bm25_rows >= n_queries, RAGB_EINVAL,
               "ragb_dense_mma_fused_topk: bm25_rows=%lld must be >= n_queries", static_cast<long long>(bm25_rows));
  RAGB_REQUIRE(dim >= MM_BK && dim % MM_BK == 0, RAGB_EINVAL, "ragb_dense_mma_fused_topk: dim=%d must be a multiple of %d",
               dim, MM_BK);
  RAGB_REQUIRE(k > 0 && k <= 100, RAGB_ELIMIT, "ragb_dense_mma_fused_topk: k=%d outside [1,100]", k);
  RAGB_REQUIRE(hidden >= 4 && hidden <= RAGB_ROUTER_MAX_HIDDEN && hidden % 4 == 0, RAGB_ELIMIT,
               "ragb_dense_mma_fused_topk: hidden=%d must be a multiple of 4 in [4,%d]", hidden, RAGB_ROUTER_MAX_HIDDEN);
  RAGB_REQUIRE(n_b >= 2 && n_d >= 1 && (n_d & (n_d - 1)) == 0 && static_cast<int64_t>(n_b) * n_d <= FF_TABLE_CELLS,
               RAGB_ELIMIT, "ragb_dense_mma_fused_topk: gate bound table %d x %d: n_d must be a power of two, at most %d cells",
               n_b, n_d, FF_TABLE_CELLS);
  RAGB_REQUIRE(b_cap > 0.0f && d_hi > 0.0f, RAGB_EINVAL, "ragb_dense_mma_fused_topk: b_cap and d_hi must be positive");
  RAGB_REQUIRE(id_base >= 0 && id_base + n_rows < (1ll << 31), RAGB_ELIMIT, "ragb_dense_mma_fused_topk: ids must fit int32");
  RAGB_REQUIRE(workspace_bytes >= ragb_dense_mma_workspace_bytes(n_queries, k), RAGB_ENOSPC,
               "ragb_dense_mma_fused_topk: workspace too small");
  const MmaWorkspace w = mma_carve(workspace, n_queries, k);
  MmaArgs f{};
  f.bm25 = bm25_scores;
  f.bm25_ld = bm25_rows;
  f.rw = RouterWeights{w1, b1, w2, b2, stats, hidden};
  f.ff_table = gate_bound_table;
  f.ff_nb = n_b;
  f.ff_nd = n_d;
  f.ff_inv_wb = static_cast<float>(n_b) / b_cap;
  f.ff_inv_wd = static_cast<float>(n_d) / (2.0f * d_hi);
  f.ff_d_hi = d_hi;
  f.counters = counters;
  static const int ff_debug = [] {
    const char* e = getenv("RAGB_FF_DEBUG");  // timing attribution only: results are wrong when set
    return e ? atoi(e) : 0;
  }();
  f.ff_debug = ff_debug;
  int rc;
  const int kpl = mma_list_kpl(k);
  const bool pair = n_queries > MM_BM;   // a lone slab has no partner for a CTA pair
  const int n_tiles = static_cast<int>(ceil_div64(n_rows, MM2_BN));
  const MmaPlan plan = mma_plan(pair, n_queries, n_tiles, device_sm_count());
  const MmaRange r{0, n_tiles, nullptr, nullptr, w.part, 2 * plan.n_groups};
#define RAGB_FUSED_ARGS passages_bf16, n_rows, dim, queries_bf16, n_queries, k, id_base, r, w.progress, w.lists, 0, stream, &f
  if (!pair) {
    if (kpl == 2) rc = launch_mma<256, false, 2, true>(RAGB_FUSED_ARGS);
    else if (kpl == 4) rc = launch_mma<256, false, 4, true>(RAGB_FUSED_ARGS);
    else rc = launch_mma<256, false, 8, true>(RAGB_FUSED_ARGS);
  } else {
    if (kpl == 2) rc = launch_mma_pair<2, true>(RAGB_FUSED_ARGS);
    else if (kpl == 4) rc = launch_mma_pair<4, true>(RAGB_FUSED_ARGS);
    else rc = launch_mma_pair<8, true>(RAGB_FUSED_ARGS);
  }
#undef RAGB_FUSED_ARGS
  if (rc != RAGB_OK) return rc;
  return launch_merge_keys(w.part, n_queries, 2 * plan.n_groups, k, k, out_score, out_id, stream);
}

// Debug aid (not part of the documented ABI): evaluate the full-fusion bound of the fused epilogue for n
// (bm25, dense) pairs with the same table geometry ragb_dense_mma_fused_topk derives from (n_b, n_d, b_cap, d_hi).
int ragb_debug_fused_bound(const float* bm25, const float* dense, int32_t n, const uint32_t* gate_bound_table, int32_t n_b,
                           int32_t n_d, float b_cap, float d_hi, float* out, ragb_stream_t stream_) {
  RAGB_ENTRY();
  RAGB_REQUIRE(bm25 && dense && gate_bound_table && out && n > 0, RAGB_EINVAL, "ragb_debug_fused_bound: bad argument");
  RAGB_REQUIRE(n_b >= 2 && n_d >= 1 && (n_d & (n_d - 1)) == 0, RAGB_EINVAL, "ragb_debug_fused_bound: bad table shape");
  ff_bound_debug_kernel<<<148, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      bm25, dense, n, gate_bound_table, n_b, n_d, static_cast<float>(n_b) / b_cap, static_cast<float>(n_d) / (2.0f * d_hi),
      d_hi, out);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

}  // extern "C"
