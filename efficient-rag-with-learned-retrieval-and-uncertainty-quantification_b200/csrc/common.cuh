// Shared host/device helpers for libragb200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ragb200.h"

namespace ragb {

// ---- host side -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int require_b200();            // RAGB_OK iff current device is cc 10.x (cached per device)
int device_sm_count();
void note_launch(int n = 1);

#define RAGB_REQUIRE(cond, code, ...)            \
  do {                                           \
    if (!(cond)) {                               \
      ::ragb::set_error(__VA_ARGS__);            \
      return (code);                             \
    }                                            \
  } while (0)

#define RAGB_CUDA(call)                                                                 \
  do {                                                                                  \
    cudaError_t _e = (call);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::ragb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return RAGB_ECUDA;                                                                \
    }                                                                                   \
  } while (0)

#define RAGB_ENTRY()                         \
  do {                                       \
    int _rc = ::ragb::require_b200();        \
    if (_rc != RAGB_OK) return _rc;          \
  } while (0)

#define RAGB_AFTER_LAUNCH(n)                 \
  do {                                       \
    RAGB_CUDA(cudaGetLastError());           \
    ::ragb::note_launch(n);                  \
  } while (0)

// select.cu: merge [n_queries, n_lists, k_in] packed keys into (score, id)[n_queries, k_out]
int launch_merge_keys(const uint64_t* keys, int n_queries, int n_lists, int k_in, int k_out, float* out_score,
                      int32_t* out_id, cudaStream_t stream);
// same with one more list per query stored elsewhere (extra_keys [n_queries, k_extra], nullable) and optional outputs:
// out_keys [n_queries, k_out] packed keys, out_thr [n_queries] score of the k_out-th best (-inf if fewer); out_score /
// out_id may be null when out_keys is given
int launch_merge_keys_ex(const uint64_t* keys, int n_queries, int n_lists, int k_in, const uint64_t* extra_keys, int k_extra,
                         int k_out, float* out_score, int32_t* out_id, uint64_t* out_keys, float* out_thr,
                         cudaStream_t stream);

// two-level variant for few queries with many lists; scratch holds n_queries * MERGE_SPLIT_MAX * k_out keys
constexpr int MERGE_SPLIT_MAX = 32;
int launch_merge_keys_split(const uint64_t* keys, int n_queries, int n_lists, int k_in, const uint64_t* extra_keys, int k_extra,
                            int k_out, float* out_score, int32_t* out_id, uint64_t* out_keys, float* out_thr, uint64_t* scratch,
                            cudaStream_t stream);

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// ---- device side: ordered 64-bit candidate keys ---------------------------------------
// key = (order-preserving image of the fp32 score) << 32 | ~id.  A larger key is a better
// candidate: higher score first, then LOWER id.  key 0 means "empty slot".
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t make_key(float score, int32_t id) {
  return (static_cast<uint64_t>(float_to_ordered(score)) << 32) | static_cast<uint32_t>(~id);
}
__device__ __forceinline__ float key_score(uint64_t key) { return ordered_to_float(static_cast<uint32_t>(key >> 32)); }
__device__ __forceinline__ int32_t key_id(uint64_t key) { return static_cast<int32_t>(~static_cast<uint32_t>(key)); }
// smallest key that is NOT accepted when only strictly positive scores count
// (BM25Index.search keeps scores > 0, streaming_index.py:176)
__device__ __host__ constexpr uint64_t positive_floor_key() { return (0x80000000ull << 32) | 0xFFFFFFFFull; }

__device__ __forceinline__ void unpack_bf16x8(const uint4& v, float (&f)[8]) {
  f[0] = __uint_as_float(v.x << 16);
  f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16);
  f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16);
  f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16);
  f[7] = __uint_as_float(v.w & 0xffff0000u);
}

// streaming 128-bit load that does not pollute L1
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

}  // namespace ragb
