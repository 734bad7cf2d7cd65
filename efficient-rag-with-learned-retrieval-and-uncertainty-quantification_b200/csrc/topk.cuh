// Block-level exact top-k over a stream of 64-bit candidate keys.
//
// Replaces np.argsort(scores)[::-1][:k] (rag_uq/streaming_index.py:172), torch.topk
// (rag_uq/router.py:202) and the Python sort in hybrid_search (streaming_index.py:521).
//
// Algorithm: threshold filter + deferred bitonic merge.  The block keeps its current best
// k keys sorted at keys[0..k) and appends every candidate whose key beats the running
// k-th best ("threshold") to keys[k..).  When the append region could overflow, the whole
// live prefix is bitonic-sorted (descending), the tail beyond k is cleared and the
// threshold tightens.  After the first few hundred candidates almost nothing passes the
// filter (expected appends ~ k ln(n/k)), so the sort cost is amortised to nothing.
#pragma once
#include "common.cuh"

namespace ragb {

// Sort keys[0..n) descending; n is a power of two; all NT threads of the block call this.
template <int NT>
__device__ __forceinline__ void bitonic_sort_desc(uint64_t* keys, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (n >> 1); i += NT) {
        int lo = 2 * i - (i & (stride - 1));
        int hi = lo + stride;
        bool descending = ((lo & size) == 0);
        uint64_t a = keys[lo], b = keys[hi];
        if ((a < b) == descending) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
      __syncthreads();
    }
  }
}

template <int NT>
struct BlockTopK {
  uint64_t* keys;      // shared, capacity entries, zero = empty
  int* count;          // shared, number of appended candidates
  uint64_t* threshold; // shared, candidates must be strictly greater
  int k;
  int capacity;        // power of two, > k
  uint64_t floor_key;
  int known;           // *count as of the last reserve() - the same value in every thread of the block
  bool dirty;          // this thread appended since the last reserve()

  __device__ __forceinline__ void init(uint64_t* keys_, int* count_, uint64_t* thr_, int k_, int capacity_,
                                       uint64_t floor_key_) {
    keys = keys_;
    count = count_;
    threshold = thr_;
    k = k_;
    capacity = capacity_;
    floor_key = floor_key_;
    known = 0;
    dirty = false;
    for (int i = threadIdx.x; i < capacity; i += NT) keys[i] = 0ull;
    if (threadIdx.x == 0) {
      *count = 0;
      *threshold = floor_key;
    }
    __syncthreads();
  }

  __device__ __forceinline__ int room() const { return capacity - k; }

  // Caller guarantees (through reserve) that the append region cannot overflow.
  __device__ __forceinline__ void offer(uint64_t key, uint64_t thr) {
    if (key > thr) {
      int slot = atomicAdd(count, 1);
      keys[k + slot] = key;
      dirty = true;
    }
  }

  // Block-wide: make room for `incoming` further appends.  Contains barriers.
  // The flush decision must be block-uniform (flush() has barriers inside), so it is taken on `known`, a register
  // copy of *count that every thread refreshes at the same point: the vote tells whether anybody appended since
  // the last call (after warm-up: almost never - one barrier per call); only then is *count re-read, and a second
  // barrier keeps faster warps from appending again before every thread has read it.  Reading *count right after a
  // single barrier would let a slow warp see appends of the NEXT round and disagree with the others.
  __device__ __forceinline__ void reserve(int incoming) {
    const int any = __syncthreads_or(dirty ? 1 : 0);   // also orders every earlier append before the read below
    dirty = false;
    if (any) {
      known = *count;
      __syncthreads();
    }
    if (known + incoming > room()) {
      flush();
      known = 0;
    }
  }

  // Block-wide: fold the appended candidates into the sorted top-k.  Contains barriers;
  // every thread must reach it after a barrier that ordered the last appends.
  __device__ __forceinline__ void flush() {
    int appended = *count;
    if (appended == 0) return;
    int n = 2;
    while (n < k + appended) n <<= 1;
    __syncthreads();  // everyone has read *count before thread 0 resets it
    bitonic_sort_desc<NT>(keys, n);
    for (int i = k + threadIdx.x; i < n; i += NT) keys[i] = 0ull;
    if (threadIdx.x == 0) {
      *count = 0;
      uint64_t kth = keys[k - 1];
      *threshold = kth > floor_key ? kth : floor_key;
    }
    __syncthreads();
  }

  // Final result for this block: keys[0..k) sorted best first, zero = empty.
  __device__ __forceinline__ void finish() {
    __syncthreads();
    flush();
  }
};

// ---------------------------------------------------------------------------------------------
// Warp-scope variant: same algorithm, no block barrier anywhere.  One warp owns keys[0..capacity)
// in shared memory; the threshold and the append count live in (warp-uniform) registers.
// ---------------------------------------------------------------------------------------------
// Sort the live prefix, clear what falls behind the best k, return the new threshold.  Kept out
// of line (and free of references) so the caller's state stays in registers.
static __device__ __noinline__ uint64_t warp_topk_flush(uint64_t* keys, int k, int count, uint64_t floor_key, int lane) {
  __syncwarp();
  int n = 2;
  while (n < k + count) n <<= 1;
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = lane; i < (n >> 1); i += 32) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool descending = ((lo & size) == 0);
        const uint64_t x = keys[lo], y = keys[hi];
        if ((x < y) == descending) {
          keys[lo] = y;
          keys[hi] = x;
        }
      }
      __syncwarp();
    }
  }
  for (int i = k + lane; i < n; i += 32) keys[i] = 0ull;
  const uint64_t kth = keys[k - 1];
  __syncwarp();
  return kth > floor_key ? kth : floor_key;
}

struct WarpTopK {
  uint64_t* keys;
  int k, capacity, count;
  uint64_t floor_key, thr;
  float thr_score;

  __device__ __forceinline__ void init(uint64_t* keys_, int k_, int capacity_, uint64_t floor_key_, int lane) {
    keys = keys_;
    k = k_;
    capacity = capacity_;
    floor_key = floor_key_;
    thr = floor_key_;
    thr_score = key_score(floor_key_);
    count = 0;
    for (int i = lane; i < capacity; i += 32) keys[i] = 0ull;
    __syncwarp();
  }

  // Warp-wide: fold the appended candidates into the sorted top-k and tighten the threshold.
  __device__ __forceinline__ void flush(int lane) {
    if (count == 0) return;
    const uint64_t t = warp_topk_flush(keys, k, count, floor_key, lane);
    if (t > thr) {   // never loosen: an external bound may already be tighter
      thr = t;
      thr_score = key_score(t);
    }
    count = 0;
  }

  // A proven lower bound of the final k-th best score (e.g. from another warp or a seeding pass):
  // nothing below it can end up in the result, so it may serve as admission threshold right away.
  __device__ __forceinline__ void raise(float score) {
    if (score > thr_score) {
      thr_score = score;
      thr = static_cast<uint64_t>(float_to_ordered(score)) << 32;
    }
  }

  // Warp-wide: every lane offers at most one key (valid says whether it has one).
  __device__ __forceinline__ void offer(bool valid, uint64_t key, int lane) {
    bool c = valid && key > thr;
    unsigned m = __ballot_sync(0xffffffffu, c);
    if (m == 0) return;
    if (count + __popc(m) > capacity - k) {
      flush(lane);
      c = valid && key > thr;
      m = __ballot_sync(0xffffffffu, c);
    }
    if (c) keys[k + count + __popc(m & ((1u << lane) - 1))] = key;
    count += __popc(m);
  }
};

// (a list needs room for k kept keys plus at least one round of 32 appended ones)
__host__ __device__ constexpr int warp_topk_capacity(int k) { return k <= 64 ? 128 : (k <= 100 ? 256 : 512); }

// Capacity used for a given k (entries of 8 bytes).
__host__ __device__ constexpr int topk_capacity(int k) { return k <= 64 ? 1024 : 2048; }

}  // namespace ragb
