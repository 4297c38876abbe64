// Tree evaluation (allcal, codes/funcs.py:175-220) as a stack-machine interpreter over pre-order tokens, plus the
// per-chain Gram / X^T y / column-sum / max-abs reductions ylogLike and the intercept refit need
// (codes/funcs.py:1147-1162, codes/bsr_class.py:211-227).
//
// A tree is scanned right-to-left; `acc` holds the value of the subtree that starts at the current token, the
// stack holds finished subtrees still waiting for their binary parent.  All lanes of a warp run the same tree on
// different rows, so the opcode dispatch is warp-uniform.  Each thread carries R rows in registers so one token
// decode is amortised over R row evaluations.
#pragma once
#include <cfloat>
#include "bsr_common.cuh"

template <typename T> struct OpMath;

template <> struct OpMath<float> {
  static __device__ __forceinline__ float exp_guard(float x) { return (x <= 200.0f) ? expf(x) : 1e10f; }   // funcs.py:184-188
  static __device__ __forceinline__ float inv_guard(float x) { return (x == 0.0f) ? 0.0f : 1.0f / x; }     // funcs.py:191-195
  static __device__ __forceinline__ float sin_(float x) { return sinf(x); }
  static __device__ __forceinline__ float cos_(float x) { return cosf(x); }
  static __device__ __forceinline__ bool finite(float x) { return fabsf(x) <= FLT_MAX; }
};
template <> struct OpMath<double> {
  static __device__ __forceinline__ double exp_guard(double x) { return (x <= 200.0) ? exp(x) : 1e10; }
  static __device__ __forceinline__ double inv_guard(double x) { return (x == 0.0) ? 0.0 : 1.0 / x; }
  static __device__ __forceinline__ double sin_(double x) { return sin(x); }
  static __device__ __forceinline__ double cos_(double x) { return cos(x); }
  static __device__ __forceinline__ bool finite(double x) { return fabs(x) <= DBL_MAX; }
};

#define BSR_STACK (BSR_MAXN / 2 + 1)

// Evaluate one tree on R rows.  tk/ta/tb: tokens and lt parameters (shared memory), X: column-major data with
// leading dimension ld, rows[r]: row index of lane-row r (already clamped to a valid row).
template <typename T, int R>
__device__ __forceinline__ void eval_tree_rows(const uint32_t* tk, const T* ta, const T* tb, int m, const T* __restrict__ X,
                                               int64_t ld, const int64_t (&rows)[R], T (&acc)[R]) {
  T stk[BSR_STACK][R];
  int sp = 0;
  for (int i = m - 1; i >= 0; --i) {
    const uint32_t t = tk[i];
    const int o = tok_op(t);
    if (o == OP_LEAF) {
      if (i != m - 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) stk[sp][r] = acc[r];
        ++sp;
      }
      const T* col = X + (int64_t)tok_ft(t) * ld;
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = __ldg(col + rows[r]);
    } else if (o >= OP_ADD) {
      --sp;
      if (o == OP_ADD) {
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = acc[r] + stk[sp][r];
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = acc[r] * stk[sp][r];
      }
    } else {
      switch (o) {
        case OP_LT: {
          const T a = ta[i], b = tb[i];
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = a * acc[r] + b;
        } break;
        case OP_INV:
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = OpMath<T>::inv_guard(acc[r]);
          break;
        case OP_NEG:
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = -acc[r];
          break;
        case OP_SIN:
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = OpMath<T>::sin_(acc[r]);
          break;
        case OP_COS:
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = OpMath<T>::cos_(acc[r]);
          break;
        case OP_EXP:
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = OpMath<T>::exp_guard(acc[r]);
          break;
        case OP_SQUARE:
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = acc[r] * acc[r];
          break;
        default:  // OP_CUBIC
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = acc[r] * acc[r] * acc[r];
          break;
      }
    }
  }
}

// Layout of the per-chain reduction record produced by the eval kernel for P columns:
//   sums : G upper triangle (row-major, i<=j) [P(P+1)/2], col.y [P], col sums [P]
//   maxs : max|col| [P]
__host__ __device__ __forceinline__ int gram_n_sum(int P) { return P * (P + 1) / 2 + 2 * P; }
__host__ __device__ __forceinline__ int gram_idx(int P, int i, int j) {   // i <= j
  return i * P - i * (i - 1) / 2 + (j - i);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
