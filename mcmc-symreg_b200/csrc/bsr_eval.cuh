// Tree evaluation (allcal, codes/funcs.py:175-220) as a stack-machine interpreter over pre-order tokens, plus the
// per-chain Gram / X^T y / column-sum / max-abs reductions ylogLike and the intercept refit need
// (codes/funcs.py:1147-1162, codes/bsr_class.py:211-227).
//
// A tree is scanned right-to-left; `acc` holds the value of the subtree that starts at the current token, the
// stack holds finished subtrees still waiting for their binary parent.  All lanes of a warp run the same tree on
// different rows, so the opcode dispatch is warp-uniform.  Each thread owns R consecutive rows (one 16-byte vector
// load per leaf: float4 / double2, coalesced across the warp), so one token decode is amortised over R rows.
#pragma once
#include <cfloat>
#include "bsr_common.cuh"

// ---- per-operator arithmetic -----------------------------------------------------------------------------------
// float: MUFU-based (SFU) versions with an fp32 accuracy budget: sin/cos use a two-constant Cody-Waite reduction
// to [-pi, pi] followed by MUFU.SIN/COS (abs err ~5e-7 for |x| < 1e4; sin of a small reduced argument from its series,
// sin_reduced), exp is MUFU.EX2 on x*log2(e), inv is MUFU.RCP.  double: libm-accurate versions (precision=fp64 mode and the re-evaluation of out-of-range chains).
template <typename T> struct OpMath;

// sin of a reduced argument r in [-pi, pi].  MUFU.SIN carries an ABSOLUTE error of ~2e-7 whatever the argument (measured,
// scripts/sfu_accuracy.py: sin of |x| < 1e-6 comes back as 0, relative error 17 % at 1e-5, 1.8e-3 at 1e-3), and small
// arguments are common in the trees (x^2, x^3, x*y of inputs around 0), mostly under an inv: below 2^-5 the value is taken
// from r - r^3/6 instead (relative error r^4/120 < 8e-9), above it the MUFU result is good to 6e-6 relative or better
// until the next zero of the sine.
static __device__ __forceinline__ float sin_reduced(float r) {
  const float p = fmaf(r * r, -0.16666667f * r, r);
  return fabsf(r) < 0.03125f ? p : __sinf(r);
}

template <> struct OpMath<float> {
  static __device__ __forceinline__ float exp_guard(float x) { return (x <= 200.0f) ? __expf(x) : 1e10f; }   // funcs.py:184-188
  static __device__ __forceinline__ float inv_guard(float x) {                                                  // funcs.py:191-195
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return (x == 0.0f) ? 0.0f : r;
  }
  static __device__ __forceinline__ float reduce_2pi(float x) {
    const float k = rintf(x * 0.15915494309189535f);
    float r = fmaf(k, -6.2831854820251465f, x);
    return fmaf(k, 1.7484555e-7f, r);
  }
  static __device__ __forceinline__ float sin_(float x) { return sin_reduced(reduce_2pi(x)); }
  static __device__ __forceinline__ float cos_(float x) { return __cosf(reduce_2pi(x)); }
};
template <> struct OpMath<double> {
  static __device__ __forceinline__ double exp_guard(double x) { return (x <= 200.0) ? exp(x) : 1e10; }
  static __device__ __forceinline__ double inv_guard(double x) { return (x == 0.0) ? 0.0 : 1.0 / x; }
  static __device__ __forceinline__ double sin_(double x) { return sin(x); }
  static __device__ __forceinline__ double cos_(double x) { return cos(x); }
};

// Double RANGE with the fp32 accuracy budget of the fp32 mode: used where a column left the fp32 range (exp of
// 88.7 .. 200, powers of large values) and is interpreted again in double.  What those columns need is exponent range,
// not 53 bits: exp is MUFU.EX2 on the fraction plus an exponent add, sin / cos reduce the argument in double and hand
// the remainder to MUFU -- a few instructions instead of the libm routines and their slow paths for huge arguments.
struct OpMathWide {
  static __device__ __forceinline__ double exp_guard(double x) {                     // funcs.py:184-188
    if (!(x <= 200.0)) return 1e10;                                                  // NaN -> 1e10, like the reference's loop
    const double t = x * 1.4426950408889634;
    if (t < -1020.0) return 0.0;
    const double n = rint(t);
    const double p = (double)exp2f((float)(t - n));                                  // [0.707, 1.415]
    return __longlong_as_double(__double_as_longlong(p) + ((long long)n << 52));     // p * 2^n, n in [-1020, 289]
  }
  static __device__ __forceinline__ double inv_guard(double x) { return (x == 0.0) ? 0.0 : 1.0 / x; }
  static __device__ __forceinline__ float reduce_2pi(double x) {
    const double k = rint(x * 0.15915494309189535);
    double r = fma(k, -6.283185307179586, x);
    r = fma(k, -2.4492935982947064e-16, r);
    // |x| beyond ~2^52: the double holds no bit of the phase (libm returns some value in [-1, 1], numpy another).  Keep the
    // column finite like they do, with a phase taken from the low mantissa bits (deterministic, not a constant, so such
    // columns do not turn collinear with the intercept).  inf / NaN arguments stay NaN.
    if (!(fabs(r) <= 4.0))
      r = (fabs(x) <= DBL_MAX) ? ((double)(unsigned)__double2loint(x) * (1.0 / 4294967296.0) - 0.5) * 6.283185307179586 : x - x;
    return (float)r;
  }
  static __device__ __forceinline__ double sin_(double x) { return (double)sin_reduced(reduce_2pi(x)); }
  static __device__ __forceinline__ double cos_(double x) { return (double)__cosf(reduce_2pi(x)); }
};

// R consecutive values of T as one 16-byte vector
template <typename T> struct RowVec;
template <> struct RowVec<float> { static constexpr int R = 4; typedef float4 V; };
template <> struct RowVec<double> { static constexpr int R = 2; typedef double2 V; };

template <typename T, int R>
__device__ __forceinline__ void vec_load(const T* p, T (&v)[R]) {
  typedef typename RowVec<T>::V V;
  const V q = __ldg(reinterpret_cast<const V*>(p));
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = ((const T*)&q)[r];
}

// Pre-decoded token as staged in shared memory: opcode, element offset of the leaf's column (feature * ld), and the
// lt parameters in the evaluation type.
template <typename T>
struct __align__(8) EvTok {
  int op;
  uint32_t off;
  T a, b;
};

#define BSR_STACK (BSR_MAXN / 2 + 1)

// Evaluate one tree on the R rows starting at element `row0` of every column.
// M: arithmetic of the transcendental / guarded operators (OpMath<T>, or OpMathWide for double range at fp32 accuracy).
template <typename T, int R, typename M = OpMath<T> >
__device__ __forceinline__ void eval_tree_rows(const EvTok<T>* tk, int m, const T* __restrict__ X, uint32_t row0, T (&acc)[R]) {
  T stk[BSR_STACK][R];
  int sp = 0;
#pragma unroll 1
  for (int i = m - 1; i >= 0; --i) {
    const EvTok<T> t = tk[i];
    const int o = t.op;
    if (o == OP_LEAF) {
      if (i != m - 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) stk[sp][r] = acc[r];
        ++sp;
      }
      vec_load<T, R>(X + (size_t)t.off + row0, acc);
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
        case OP_LT:
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = t.a * acc[r] + t.b;
          break;
        case OP_INV:
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = M::inv_guard(acc[r]);
          break;
        case OP_NEG:
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = -acc[r];
          break;
        case OP_SIN:
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = M::sin_(acc[r]);
          break;
        case OP_COS:
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = M::cos_(acc[r]);
          break;
        case OP_EXP:
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = M::exp_guard(acc[r]);
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

// Same interpreter with NV row vectors per thread: one token decode (shared-memory load, opcode dispatch) is amortised
// over NV * R rows, and the NV independent vectors give the FP32 / SFU pipes instruction-level parallelism.  Vector u
// of the thread starts at element rowoff[u] of every column (the caller interleaves the vectors of a warp so that each
// 16-byte load is coalesced).
template <typename T, int R, int NV>
__device__ __forceinline__ void eval_tree_rows_nv(const EvTok<T>* tk, int m, const T* __restrict__ X, const uint32_t (&rowoff)[NV],
                                                  T (&acc)[NV][R]) {
  T stk[BSR_STACK][NV][R];
  int sp = 0;
  EvTok<T> nxt = tk[m - 1];
#pragma unroll 1
  for (int i = m - 1; i >= 0; --i) {
    const EvTok<T> t = nxt;
    nxt = tk[i > 0 ? i - 1 : 0];              // the next token is in flight while this one executes
    const int o = t.op;
    if (o == OP_LEAF) {
      if (i != m - 1) {
#pragma unroll
        for (int u = 0; u < NV; ++u)
#pragma unroll
          for (int r = 0; r < R; ++r) stk[sp][u][r] = acc[u][r];
        ++sp;
      }
      const T* col = X + (size_t)t.off;
#pragma unroll
      for (int u = 0; u < NV; ++u) vec_load<T, R>(col + rowoff[u], acc[u]);
    } else if (o >= OP_ADD) {
      --sp;
      if (o == OP_ADD) {
#pragma unroll
        for (int u = 0; u < NV; ++u)
#pragma unroll
          for (int r = 0; r < R; ++r) acc[u][r] = acc[u][r] + stk[sp][u][r];
      } else {
#pragma unroll
        for (int u = 0; u < NV; ++u)
#pragma unroll
          for (int r = 0; r < R; ++r) acc[u][r] = acc[u][r] * stk[sp][u][r];
      }
    } else {
#define BSR_UNARY(EXPR)                                   \
  _Pragma("unroll") for (int u = 0; u < NV; ++u)          \
  _Pragma("unroll") for (int r = 0; r < R; ++r) { const T x = acc[u][r]; acc[u][r] = (EXPR); }
      switch (o) {
        case OP_LT: BSR_UNARY(t.a * x + t.b) break;
        case OP_INV: BSR_UNARY(OpMath<T>::inv_guard(x)) break;
        case OP_NEG: BSR_UNARY(-x) break;
        case OP_SIN: BSR_UNARY(OpMath<T>::sin_(x)) break;
        case OP_COS: BSR_UNARY(OpMath<T>::cos_(x)) break;
        case OP_EXP: BSR_UNARY(OpMath<T>::exp_guard(x)) break;
        case OP_SQUARE: BSR_UNARY(x * x) break;
        default: BSR_UNARY(x * x * x) break;   // OP_CUBIC
      }
#undef BSR_UNARY
    }
  }
}

// Layout of the per-chain reduction record produced by the eval kernel for P columns:
//   sums : G upper triangle (row-major, i<=j) [P(P+1)/2], col.y [P], col sums [P]
//   maxs : max|col| [P]   (+inf marks a column with a non-finite value)
__host__ __device__ constexpr int gram_n_sum(int P) { return P * (P + 1) / 2 + 2 * P; }
__host__ __device__ constexpr int gram_idx(int P, int i, int j) {   // i <= j
  return i * P - i * (i - 1) / 2 + (j - i);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { T w = __shfl_xor_sync(0xffffffffu, v, o); v = v > w ? v : w; }
  return v;
}

// Per-thread Gram accumulators for PC columns (compile-time): G = C'C (upper triangle), C'y, column sums in fp64
// (products of fp32 values are exact in fp64), max|column| in T.  A non-finite value in column p makes G[p][p]
// non-finite, which is how out-of-range columns are detected -- no per-value checks.
template <typename T, int PC>
struct GramAcc {
  static constexpr int NG = PC * (PC + 1) / 2;
  double G[NG], Y[PC], S[PC];
  T M[PC];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < NG; ++i) G[i] = 0.0;
#pragma unroll
    for (int i = 0; i < PC; ++i) { Y[i] = 0.0; S[i] = 0.0; M[i] = (T)0; }
  }
};

// Column-cache modes of the fp32 pass: CM_PLAIN evaluates all PC trees; CM_FILL does the same and writes the K
// live columns (and the proposals) into the cache; CM_CACHED reads the K live columns from the cache, evaluates only
// the K proposals (writing them to the spare cache buffer) and accumulates only Gram entries involving a proposal.
// CM_MIXED (fp64 pass of a chain that left the fp32 range): only the columns flagged in `badmask` are interpreted in
// double; the others are read from the fp32 cache and widened.
enum : int { CM_PLAIN = 0, CM_FILL = 1, CM_CACHED = 2, CM_MIXED = 3 };

template <typename T> __device__ __forceinline__ void vec_store(T* p, const typename RowVec<T>::V& v) {
  *reinterpret_cast<typename RowVec<T>::V*>(p) = v;
}

// Evaluate the PC staged trees on all row vectors owned by this thread (vector q = pass * tpc + lane covers rows
// [q*R, q*R + R)) and accumulate the Gram record.  my_cv: this thread's staging slot for column p is my_cv[p * cvs].
// cp[p]: cache column of record column p (p < K: the live column of slot p, p >= K: the spare column that receives
// the proposal of slot p - K); only used when CM != CM_PLAIN.
// LOADALL: every column already sits in the cache (written by k_trees); nothing is interpreted here and the values go
// straight from global memory to registers.
template <typename T, int PC, int CM, bool LOADALL = false>
__device__ __forceinline__ void eval_chain_rows(GramAcc<T, PC>& acc, const EvTok<T>* s_tok, const int* s_m,
                                                typename RowVec<T>::V* my_cv, int cvs, const T* __restrict__ X,
                                                const double* __restrict__ y, uint32_t n, uint32_t v0, uint32_t v1, int lane, int tpc,
                                                T* const* cp, unsigned badmask = 0) {
  constexpr int R = RowVec<T>::R;
  constexpr int KH = PC / 2;
  typedef typename RowVec<T>::V V;
#pragma unroll 1
  for (uint32_t q = v0 + lane; q < v1; q += tpc) {   // row vectors [v0, v1) of this (chain, row split)
    const uint32_t row0 = q * R;
    V cv[PC];
    if (LOADALL) {
#pragma unroll
      for (int p = 0; p < PC; ++p) {
        if (s_m[p] > 0) cv[p] = *reinterpret_cast<const V*>(cp[p] + row0);
        else {
#pragma unroll
          for (int r = 0; r < R; ++r) ((T*)&cv[p])[r] = (T)0;
        }
      }
    } else {
#pragma unroll 1
    for (int p = 0; p < PC; ++p) {
      V pack;
      if (CM == CM_CACHED && p < KH) {
        pack = *reinterpret_cast<const V*>(cp[p] + row0);
      } else if (CM == CM_MIXED && !((badmask >> p) & 1u)) {
        // in-range column: widen the cached fp32 values (R == 2 here)
        if (s_m[p] > 0) {
          const float2 f = *reinterpret_cast<const float2*>(reinterpret_cast<const float* const*>(cp)[p] + row0);
          ((T*)&pack)[0] = (T)f.x; ((T*)&pack)[1] = (T)f.y;
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) ((T*)&pack)[r] = (T)0;
        }
      } else {
        T v[R];
        const int m = s_m[p];
        if (m > 0) {
          eval_tree_rows<T, R>(s_tok + p * BSR_MAXN, m, X, row0, v);
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) v[r] = (T)0;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) ((T*)&pack)[r] = v[r];
        if ((CM == CM_FILL || CM == CM_CACHED) && m > 0) vec_store<T>(cp[p] + row0, pack);
      }
      my_cv[p * cvs] = pack;
    }
#pragma unroll
    for (int i = 0; i < PC; ++i) cv[i] = my_cv[i * cvs];
    }
    double yv[R];                           // y stays fp64 (the y-terms of the Gram enter the SSE by cancellation)
#pragma unroll
    for (int r = 0; r < R; ++r) yv[r] = __ldg(y + row0 + r);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (row0 + r >= n) continue;          // ragged tail: rows past n are padding
      double v[PC];
#pragma unroll
      for (int i = 0; i < PC; ++i) v[i] = (double)((const T*)&cv[i])[r];
      const double yr = yv[r];
      int k = 0;
#pragma unroll
      for (int i = 0; i < PC; ++i) {
#pragma unroll
        for (int j = i; j < PC; ++j) {
          if (CM != CM_CACHED || j >= KH) acc.G[k] = fma(v[i], v[j], acc.G[k]);
          ++k;
        }
        if (CM != CM_CACHED || i >= KH) {
          acc.Y[i] = fma(v[i], yr, acc.Y[i]);
          acc.S[i] += v[i];
          const T xv = ((const T*)&cv[i])[r];
          const T av = xv < (T)0 ? -xv : xv;
          acc.M[i] = acc.M[i] > av ? acc.M[i] : av;
        }
      }
    }
  }
}

// Warp-level reduction of a GramAcc; lane 0 writes dst[0 .. n_sum + PC): the sums then the max-abs values (as
// doubles).  In CM_CACHED mode the live x live entries are not reduced (they come from the state's Gram cache).
template <typename T, int PC, int CM>
__device__ __forceinline__ void warp_reduce_store(const GramAcc<T, PC>& acc, double* dst, int wlane) {
  constexpr int KH = PC / 2;
  int q = 0;
#pragma unroll
  for (int i = 0; i < PC; ++i)
#pragma unroll
    for (int j = i; j < PC; ++j, ++q)
      if (CM != CM_CACHED || j >= KH) { double v = warp_sum(acc.G[q]); if (wlane == 0) dst[q] = v; }
#pragma unroll
  for (int i = 0; i < PC; ++i, ++q) if (CM != CM_CACHED || i >= KH) { double v = warp_sum(acc.Y[i]); if (wlane == 0) dst[q] = v; }
#pragma unroll
  for (int i = 0; i < PC; ++i, ++q) if (CM != CM_CACHED || i >= KH) { double v = warp_sum(acc.S[i]); if (wlane == 0) dst[q] = v; }
#pragma unroll
  for (int i = 0; i < PC; ++i, ++q) if (CM != CM_CACHED || i >= KH) { T v = warp_max<T>(acc.M[i]); if (wlane == 0) dst[q] = (double)v; }
}

// Layout of the per-chain Gram cache of the live columns: G(live, live) upper [K(K+1)/2], live'y, sums, max-abs [K each].
__host__ __device__ constexpr int sg_size(int K) { return K * (K + 1) / 2 + 3 * K; }
// Copy the live x live part of a P = 2K record from / to the state's Gram cache.
__device__ __forceinline__ void sg_to_record(const double* sg, double* rec, int K, int i0, int step) {
  const int P = 2 * K, n_sum = gram_n_sum(P), ng = P * (P + 1) / 2, kg = K * (K + 1) / 2;
  for (int e = i0; e < kg + 3 * K; e += step) {
    if (e < kg) {
      int i = 0, r = e;
      while (r >= K - i) { r -= K - i; ++i; }
      rec[gram_idx(P, i, i + r)] = sg[e];
    } else {
      const int w = (e - kg) / K, i = (e - kg) % K;
      rec[(w == 0 ? ng : (w == 1 ? ng + P : n_sum)) + i] = sg[e];
    }
  }
}

// Given the reduced record of P columns, mark non-finite columns (+inf in maxs) and return their bit mask.
__device__ __forceinline__ unsigned mark_bad_columns(const double* sums, double* maxs, int P, int p0 = 0) {
  unsigned bad = 0;
  for (int p = p0; p < P; ++p) {
    const double gpp = sums[gram_idx(P, p, p)];
    if (!(fabs(gpp) <= DBL_MAX) || !(maxs[p] <= DBL_MAX)) { bad |= 1u << p; maxs[p] = INFINITY; }
  }
  return bad;
}
