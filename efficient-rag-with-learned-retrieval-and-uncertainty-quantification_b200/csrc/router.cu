// RetrievalRouter gate, learned fusion and MC-Dropout sampling (K3 / K4 in SURVEY.md).
//
// Replaces RetrievalRouter._normalize_scores / forward / hybrid_rerank
// (rag_uq/router.py:100-202) and T stochastic passes of the nn.Dropout at router.py:78,
// aggregated with the arithmetic of MCDropoutConfidence (rag_uq/confidence.py:195-202,
// 258-264).  The gate is Linear(3,H) -> ReLU -> Dropout -> Linear(H,1) -> Sigmoid on the
// features [bn, dn, dn - bn]; the arithmetic keeps the reference's three products per hidden
// unit (no weight folding) so fp32 results stay within 1e-5 of torch.
#include <cmath>

#include "common.cuh"
#include "router.cuh"

namespace ragb {

constexpr int RT_THREADS = 256;
constexpr int RT_STAT_BLOCKS = 128;

// ---- statistics of the call's own scores (norm_mode 0 / 2) ------------------------------
// partial[b] = {sum_b, sumsq_b, sum_d, sumsq_d} in float64, fixed grid, fixed reduction order.
__global__ void __launch_bounds__(RT_THREADS) stats_partial_kernel(const float* __restrict__ bm25,
                                                                   const float* __restrict__ dense, int64_t n,
                                                                   double* __restrict__ partial) {
  __shared__ double s[4][RT_THREADS];
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * RT_THREADS + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * RT_THREADS) {
    const double x = bm25[i], y = dense[i];
    a0 += x;
    a1 += x * x;
    a2 += y;
    a3 += y * y;
  }
  s[0][threadIdx.x] = a0;
  s[1][threadIdx.x] = a1;
  s[2][threadIdx.x] = a2;
  s[3][threadIdx.x] = a3;
  __syncthreads();
  for (int st = RT_THREADS / 2; st > 0; st >>= 1) {
    if (threadIdx.x < st)
      for (int c = 0; c < 4; ++c) s[c][threadIdx.x] += s[c][threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x < 4) partial[4 * blockIdx.x + threadIdx.x] = s[threadIdx.x][0];
}

// stats_out[4] = mean_b, std_b (unbiased), mean_d, std_d as float32, exactly what
// x.mean() / x.std() hand to the fp32 normalisation at router.py:135-136.
__global__ void stats_final_kernel(const double* __restrict__ partial, int n_partials, int64_t n,
                                   float* __restrict__ stats_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double t[4] = {0, 0, 0, 0};
  for (int b = 0; b < n_partials; ++b)
    for (int c = 0; c < 4; ++c) t[c] += partial[4 * b + c];
  const double nn = static_cast<double>(n);
  for (int c = 0; c < 2; ++c) {
    const double mean = t[2 * c] / nn;
    double var = (t[2 * c + 1] - nn * mean * mean) / (nn - 1.0);  // n == 1 -> 0/0 = NaN like torch
    if (n > 1 && var < 0.0) var = 0.0;
    stats_out[2 * c] = static_cast<float>(mean);
    stats_out[2 * c + 1] = static_cast<float>(sqrt(var));
  }
}

// per-row statistics (norm_mode 2): one warp per row of P candidates
__global__ void __launch_bounds__(RT_THREADS) stats_rows_kernel(const float* __restrict__ bm25,
                                                                const float* __restrict__ dense, int n_rows, int p,
                                                                float* __restrict__ stats_rows) {
  const int row = blockIdx.x * (RT_THREADS / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int j = lane; j < p; j += 32) {
    const double x = bm25[static_cast<int64_t>(row) * p + j], y = dense[static_cast<int64_t>(row) * p + j];
    a0 += x;
    a1 += x * x;
    a2 += y;
    a3 += y * y;
  }
  for (int s = 16; s > 0; s >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, s);
    a1 += __shfl_xor_sync(0xffffffffu, a1, s);
    a2 += __shfl_xor_sync(0xffffffffu, a2, s);
    a3 += __shfl_xor_sync(0xffffffffu, a3, s);
  }
  if (lane == 0) {
    const double nn = p;
    const double mb = a0 / nn, md = a2 / nn;
    double vb = (a1 - nn * mb * mb) / (nn - 1.0), vd = (a3 - nn * md * md) / (nn - 1.0);
    if (p > 1 && vb < 0.0) vb = 0.0;
    if (p > 1 && vd < 0.0) vd = 0.0;
    float* o = stats_rows + 4 * static_cast<int64_t>(row);
    o[0] = static_cast<float>(mb);
    o[1] = static_cast<float>(sqrt(vb));
    o[2] = static_cast<float>(md);
    o[3] = static_cast<float>(sqrt(vd));
  }
}

// ---- gate ---------------------------------------------------------------------------------
// stats_ptr: [4] (modes 0/1) or [rows,4] (mode 2, row = element / p)
__global__ void __launch_bounds__(RT_THREADS) router_forward_kernel(const float* __restrict__ bm25,
                                                                    const float* __restrict__ dense, int64_t n,
                                                                    RouterWeights w, const float* __restrict__ stats_ptr,
                                                                    int per_row_p, float* __restrict__ out_gate,
                                                                    float* __restrict__ out_fused) {
  __shared__ float s_w1[RAGB_ROUTER_MAX_HIDDEN * 3];
  __shared__ float s_b1[RAGB_ROUTER_MAX_HIDDEN];
  __shared__ float s_w2[RAGB_ROUTER_MAX_HIDDEN];
  const int H = w.hidden;
  for (int i = threadIdx.x; i < 3 * H; i += RT_THREADS) s_w1[i] = w.w1[i];
  for (int i = threadIdx.x; i < H; i += RT_THREADS) {
    s_b1[i] = w.b1[i];
    s_w2[i] = w.w2[i];
  }
  __syncthreads();
  const float b2 = w.b2[0];
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * RT_THREADS + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * RT_THREADS) {
    const float* st = per_row_p > 0 ? stats_ptr + 4 * (i / per_row_p) : stats_ptr;
    const float xb = bm25[i], xd = dense[i];
    const float g = gate_eval(s_w1, s_b1, s_w2, b2, H, st, xb, xd);
    if (out_gate) out_gate[i] = g;
    if (out_fused) out_fused[i] = fuse_scores(g, xb, xd);  // router.py:199, raw scores
  }
}

// ---- MC-Dropout ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    if (r != 9) {
      k.x += 0x9E3779B9u;
      k.y += 0xBB67AE85u;
    }
  }
  return c;
}
__device__ __forceinline__ float uniform_from_bits(uint32_t x) {
  return static_cast<float>(x) * 2.3283064365386963e-10f + 1.1641532182693481e-10f;  // curand_uniform
}

struct McArgs {
  const float* bm25;
  const float* dense;
  int n_queries, n_cand, n_samples;
  RouterWeights w;
  const float* stats_ptr;
  int per_row_p;  // > 0: stats_ptr is [rows,4]
  float keep_prob, scale;
  uint64_t seed, offset;
  int mask_layout;
  uint64_t torch_threads_total, torch_increment;  // layout 1
  float *mean_gate, *std_gate, *mean_fused, *std_fused, *variance;
  int32_t* consensus;
  uint8_t* mask_dump;
  float* gate_dump;
};

__global__ void __launch_bounds__(RT_THREADS) router_mc_kernel(const McArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_gate = reinterpret_cast<float*>(smem_raw);  // [T][P]
  float* s_dist = s_gate + a.n_samples * a.n_cand;      // [T]
  __shared__ float s_w1[RAGB_ROUTER_MAX_HIDDEN * 3];
  __shared__ float s_b1[RAGB_ROUTER_MAX_HIDDEN];
  __shared__ float s_w2[RAGB_ROUTER_MAX_HIDDEN];
  const int H = a.w.hidden, T = a.n_samples, P = a.n_cand;
  const int q = blockIdx.x;
  for (int i = threadIdx.x; i < 3 * H; i += RT_THREADS) s_w1[i] = a.w.w1[i];
  for (int i = threadIdx.x; i < H; i += RT_THREADS) {
    s_b1[i] = a.w.b1[i];
    s_w2[i] = a.w.w2[i];
  }
  __syncthreads();
  const float b2 = a.w.b2[0];
  const float* st = a.per_row_p > 0 ? a.stats_ptr + 4 * q : a.stats_ptr;
  const uint2 key = make_uint2(static_cast<uint32_t>(a.seed), static_cast<uint32_t>(a.seed >> 32));

  // phase 1: one (sample, candidate) pair per thread
  for (int pair = threadIdx.x; pair < T * P; pair += RT_THREADS) {
    const int t = pair / P, p = pair % P;
    const int64_t cand = static_cast<int64_t>(q) * P + p;
    const float xb = a.bm25[cand], xd = a.dense[cand];
    const float bn = (xb - st[0]) / (st[1] + RT_EPS);
    const float dn = (xd - st[2]) / (st[3] + RT_EPS);
    const float df = dn - bn;
    float z = b2;
    for (int u = 0; u < H; u += 4) {
      uint64_t subseq, off4;
      if (a.mask_layout == 1) {
        const uint64_t quad = (static_cast<uint64_t>(cand) * H + u) >> 2;
        subseq = quad % a.torch_threads_total;
        off4 = ((a.offset + static_cast<uint64_t>(t) * a.torch_increment) >> 2) + quad / a.torch_threads_total;
      } else {
        subseq = static_cast<uint64_t>(cand);
        off4 = (a.offset >> 2) + static_cast<uint64_t>(t) * (H >> 2) + (u >> 2);
      }
      const uint4 ctr = make_uint4(static_cast<uint32_t>(off4), static_cast<uint32_t>(off4 >> 32),
                                   static_cast<uint32_t>(subseq), static_cast<uint32_t>(subseq >> 32));
      const uint4 bits = philox4x32_10(ctr, key);
      const uint32_t rb[4] = {bits.x, bits.y, bits.z, bits.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = u + i;
        const bool keep = uniform_from_bits(rb[i]) < a.keep_prob;
        float h = fmaf(s_w1[3 * j + 2], df, fmaf(s_w1[3 * j + 1], dn, fmaf(s_w1[3 * j], bn, s_b1[j])));
        h = h < 0.0f ? 0.0f : h;
        h = keep ? h * a.scale : 0.0f;
        z = fmaf(s_w2[j], h, z);
        if (a.mask_dump) a.mask_dump[(static_cast<int64_t>(t) * a.n_queries * P + cand) * H + j] = keep ? 1 : 0;
      }
    }
    const float g = sigmoidf_exact(z);
    s_gate[t * P + p] = g;
    if (a.gate_dump) a.gate_dump[(static_cast<int64_t>(t) * a.n_queries + q) * P + p] = g;
  }
  __syncthreads();

  // phase 2a: per candidate mean / population std over the T samples (confidence.py:200 uses ddof = 0)
  for (int p = threadIdx.x; p < P; p += RT_THREADS) {
    const int64_t cand = static_cast<int64_t>(q) * P + p;
    const float xb = a.bm25[cand], xd = a.dense[cand];
    float mg = 0.0f, mf = 0.0f;
    for (int t = 0; t < T; ++t) {
      const float g = s_gate[t * P + p];
      mg += g;
      mf += fuse_scores(g, xb, xd);
    }
    mg /= T;
    mf /= T;
    float vg = 0.0f, vf = 0.0f;
    for (int t = 0; t < T; ++t) {
      const float g = s_gate[t * P + p];
      const float f = fuse_scores(g, xb, xd);
      vg += (g - mg) * (g - mg);
      vf += (f - mf) * (f - mf);
    }
    a.mean_gate[cand] = mg;
    a.std_gate[cand] = sqrtf(vg / T);
    a.mean_fused[cand] = mf;
    a.std_fused[cand] = sqrtf(vf / T);
  }
  __syncthreads();
  // phase 2b: distance of every sample's gate vector to the centroid (confidence.py:196-199)
  for (int t = threadIdx.x; t < T; t += RT_THREADS) {
    float d2 = 0.0f;
    for (int p = 0; p < P; ++p) {
      const float diff = s_gate[t * P + p] - a.mean_gate[static_cast<int64_t>(q) * P + p];
      d2 = fmaf(diff, diff, d2);
    }
    s_dist[t] = sqrtf(d2);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.0f;
    int best = 0;
    for (int t = 0; t < T; ++t) {
      m += s_dist[t];
      if (s_dist[t] < s_dist[best]) best = t;
    }
    m /= T;
    float v = 0.0f;
    for (int t = 0; t < T; ++t) v += (s_dist[t] - m) * (s_dist[t] - m);
    a.variance[q] = sqrtf(v / T);  // "variance = float(distances.std())", confidence.py:200
    a.consensus[q] = best;         // argmin distance, confidence.py:248-249
  }
}

static int router_checks(const char* who, const float* w1, const float* b1, const float* w2, const float* b2,
                         const float* stats, int hidden, int norm_mode) {
  RAGB_REQUIRE(w1 && b1 && w2 && b2, RAGB_EINVAL, "%s: null weight pointer", who);
  RAGB_REQUIRE(hidden >= 4 && hidden <= RAGB_ROUTER_MAX_HIDDEN && hidden % 4 == 0, RAGB_ELIMIT,
               "%s: hidden=%d must be a multiple of 4 in [4,%d]", who, hidden, RAGB_ROUTER_MAX_HIDDEN);
  RAGB_REQUIRE(norm_mode >= 0 && norm_mode <= 2, RAGB_EINVAL, "%s: norm_mode must be 0, 1 or 2", who);
  RAGB_REQUIRE(norm_mode != 1 || stats, RAGB_EINVAL, "%s: running statistics required for norm_mode 1", who);
  return RAGB_OK;
}

// Produces the statistics pointer the kernels read; returns per_row_p through *per_row.
static int prepare_stats(const float* bm25, const float* dense, int64_t n, int rows, int p, int norm_mode,
                         const float* running, void* scratch, const float** stats_out, int* per_row,
                         cudaStream_t stream) {
  *per_row = 0;
  if (norm_mode == 1) {
    *stats_out = running;
    return RAGB_OK;
  }
  RAGB_REQUIRE(scratch, RAGB_EINVAL, "router: scratch required for batch statistics");
  if (norm_mode == 0) {
    double* partial = static_cast<double*>(scratch);
    float* st = reinterpret_cast<float*>(partial + 4 * RT_STAT_BLOCKS);
    stats_partial_kernel<<<RT_STAT_BLOCKS, RT_THREADS, 0, stream>>>(bm25, dense, n, partial);
    RAGB_AFTER_LAUNCH(1);
    stats_final_kernel<<<1, 32, 0, stream>>>(partial, RT_STAT_BLOCKS, n, st);
    RAGB_AFTER_LAUNCH(1);
    *stats_out = st;
    return RAGB_OK;
  }
  float* st = static_cast<float*>(scratch);
  stats_rows_kernel<<<(rows + 7) / 8, RT_THREADS, 0, stream>>>(bm25, dense, rows, p, st);
  RAGB_AFTER_LAUNCH(1);
  *stats_out = st;
  *per_row = p;
  return RAGB_OK;
}

}  // namespace ragb

using namespace ragb;

extern "C" {

size_t ragb_router_scratch_bytes(int32_t n_rows, int32_t norm_mode) {
  if (norm_mode == 2) return static_cast<size_t>(n_rows > 0 ? n_rows : 1) * 4 * sizeof(float);
  return 4 * RT_STAT_BLOCKS * sizeof(double) + 4 * sizeof(float);
}

int ragb_router_forward(const float* bm25, const float* dense, int32_t n_rows, int32_t n_cand, const float* w1,
                        const float* b1, const float* w2, const float* b2, const float* stats, int32_t hidden,
                        int32_t norm_mode, float* out_gate, float* out_fused, void* scratch, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(bm25 && dense && (out_gate || out_fused), RAGB_EINVAL, "ragb_router_forward: null pointer");
  RAGB_REQUIRE(n_rows > 0 && n_cand > 0, RAGB_EINVAL, "ragb_router_forward: empty shape");
  int rc = router_checks("ragb_router_forward", w1, b1, w2, b2, stats, hidden, norm_mode);
  if (rc != RAGB_OK) return rc;
  const int64_t n = static_cast<int64_t>(n_rows) * n_cand;
  const float* st = nullptr;
  int per_row = 0;
  rc = prepare_stats(bm25, dense, n, n_rows, n_cand, norm_mode, stats, scratch, &st, &per_row, stream);
  if (rc != RAGB_OK) return rc;
  RouterWeights w{w1, b1, w2, b2, stats, hidden};
  int64_t blocks = ceil_div64(n, RT_THREADS);
  if (blocks > 148 * 8) blocks = 148 * 8;
  router_forward_kernel<<<static_cast<unsigned>(blocks), RT_THREADS, 0, stream>>>(bm25, dense, n, w, st, per_row,
                                                                                  out_gate, out_fused);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

int ragb_router_mc_dropout(const float* bm25, const float* dense, int32_t n_queries, int32_t n_cand, const float* w1,
                           const float* b1, const float* w2, const float* b2, const float* stats, int32_t hidden,
                           int32_t norm_mode, int32_t n_samples, double p_drop, uint64_t seed, uint64_t offset,
                           int32_t mask_layout, int32_t sm_count, float* mean_gate, float* std_gate, float* mean_fused,
                           float* std_fused, float* variance, int32_t* consensus, uint8_t* mask_dump, float* gate_dump,
                           void* scratch, ragb_stream_t stream_) {
  RAGB_ENTRY();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RAGB_REQUIRE(bm25 && dense && mean_gate && std_gate && mean_fused && std_fused && variance && consensus, RAGB_EINVAL,
               "ragb_router_mc_dropout: null pointer");
  RAGB_REQUIRE(n_queries > 0 && n_cand > 0 && n_samples > 0, RAGB_EINVAL, "ragb_router_mc_dropout: empty shape");
  RAGB_REQUIRE(p_drop >= 0.0 && p_drop < 1.0, RAGB_EINVAL, "ragb_router_mc_dropout: p_drop must be in [0,1)");
  RAGB_REQUIRE(offset % 4 == 0, RAGB_EINVAL, "ragb_router_mc_dropout: Philox offset must be a multiple of 4");
  RAGB_REQUIRE(mask_layout == 0 || mask_layout == 1, RAGB_EINVAL, "ragb_router_mc_dropout: mask_layout must be 0 or 1");
  int rc = router_checks("ragb_router_mc_dropout", w1, b1, w2, b2, stats, hidden, norm_mode);
  if (rc != RAGB_OK) return rc;
  const size_t smem = sizeof(float) * (static_cast<size_t>(n_samples) * n_cand + n_samples);
  RAGB_REQUIRE(smem <= 200 * 1024, RAGB_ELIMIT, "ragb_router_mc_dropout: n_samples*n_cand=%lld too large",
               static_cast<long long>(n_samples) * n_cand);
  const int64_t n = static_cast<int64_t>(n_queries) * n_cand;
  const float* st = nullptr;
  int per_row = 0;
  rc = prepare_stats(bm25, dense, n, n_queries, n_cand, norm_mode, stats, scratch, &st, &per_row, stream);
  if (rc != RAGB_OK) return rc;
  McArgs a{};
  a.bm25 = bm25;
  a.dense = dense;
  a.n_queries = n_queries;
  a.n_cand = n_cand;
  a.n_samples = n_samples;
  a.w = RouterWeights{w1, b1, w2, b2, stats, hidden};
  a.stats_ptr = st;
  a.per_row_p = per_row;
  a.keep_prob = static_cast<float>(1.0 - p_drop);
  a.scale = static_cast<float>(1.0 / static_cast<double>(a.keep_prob));
  a.seed = seed;
  a.offset = offset;
  a.mask_layout = mask_layout;
  if (mask_layout == 1) {
    // torch fused dropout launch geometry for a float tensor of n*hidden elements (vector width 4)
    RAGB_REQUIRE(sm_count > 0, RAGB_EINVAL, "ragb_router_mc_dropout: sm_count required for mask_layout 1");
    const uint64_t elems = static_cast<uint64_t>(n) * hidden;
    const uint64_t block = 256;
    uint64_t grid = (elems + block - 1) / block;
    const uint64_t cap = static_cast<uint64_t>(sm_count) * (2048 / block);
    if (grid > cap) grid = cap;
    a.torch_threads_total = grid * block;
    a.torch_increment = ((elems - 1) / (block * grid * 4) + 1) * 4;
  }
  a.mean_gate = mean_gate;
  a.std_gate = std_gate;
  a.mean_fused = mean_fused;
  a.std_fused = std_fused;
  a.variance = variance;
  a.consensus = consensus;
  a.mask_dump = mask_dump;
  a.gate_dump = gate_dump;
  RAGB_CUDA(cudaFuncSetAttribute(router_mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  router_mc_kernel<<<n_queries, RT_THREADS, smem, stream>>>(a);
  RAGB_AFTER_LAUNCH(1);
  return RAGB_OK;
}

}  // extern "C"
