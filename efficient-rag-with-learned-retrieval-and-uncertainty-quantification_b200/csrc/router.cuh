// Gate arithmetic of RetrievalRouter (rag_uq/router.py:100-202) shared by the candidate-list
// kernels (router.cu) and the fused GEMM epilogue (dense_mma.cu), so that a (bm25, dense) pair
// gets bit-identical gate / fused values no matter which kernel evaluates it.
#pragma once
#include "common.cuh"

namespace ragb {

constexpr float RT_EPS = 1e-6f;  // router.py:112

struct RouterWeights {
  const float* w1;     // [H,3]
  const float* b1;     // [H]
  const float* w2;     // [H]
  const float* b2;     // [1]
  const float* stats;  // [4] bm25_mean, bm25_std, dense_mean, dense_std
  int hidden;
};

__device__ __forceinline__ float sigmoidf_exact(float x) { return 1.0f / (1.0f + expf(-x)); }

// features [bn, dn, dn - bn] (router.py:164-167) -> Linear(3,H) -> ReLU -> Linear(H,1) -> Sigmoid
// (router.py:67-85).  Three products per hidden unit, like the reference (no weight folding), so
// fp32 results stay within 1e-5 of torch.  w1 / b1 / w2 may point to shared or global memory.
__device__ __forceinline__ float gate_from_normalised(const float* w1, const float* b1, const float* w2, const float b2,
                                                      const int hidden, const float bn, const float dn) {
  const float df = dn - bn;
  float z = b2;
  for (int j = 0; j < hidden; ++j) {
    float h = fmaf(w1[3 * j + 2], df, fmaf(w1[3 * j + 1], dn, fmaf(w1[3 * j], bn, b1[j])));
    h = h < 0.0f ? 0.0f : h;  // ReLU that lets NaN through like torch.relu
    z = fmaf(w2[j], h, z);
  }
  return sigmoidf_exact(z);
}

// "(x - mean) / (std + 1e-6)" of router.py:130-136; st = bm25_mean, bm25_std, dense_mean, dense_std
__device__ __forceinline__ float gate_eval(const float* w1, const float* b1, const float* w2, const float b2,
                                           const int hidden, const float* st, const float xb, const float xd) {
  const float bn = (xb - st[0]) / (st[1] + RT_EPS);
  const float dn = (xd - st[2]) / (st[3] + RT_EPS);
  return gate_from_normalised(w1, b1, w2, b2, hidden, bn, dn);
}

// weights * dense + (1 - weights) * bm25 on the raw scores (router.py:199): two rounded products and
// one rounded sum like the three torch ops, never contracted into an FMA.
__device__ __forceinline__ float fuse_scores(const float g, const float xb, const float xd) {
  return __fadd_rn(__fmul_rn(g, xd), __fmul_rn(1.0f - g, xb));
}

}  // namespace ragb
