// __global__ kernels of the split (three-phase) sweep pipeline:
//   k_init_chains  prior initialisation                         codes/bsr_class.py:123-142
//   k_propose      Prop + auxProp + fStruc for K trees/chain    codes/funcs.py:1188-1210
//   k_eval         allcal of the 2K columns + Gram reductions   codes/funcs.py:1212-1224, 1147-1157
//   k_resolve      rank test, ylogLike, logR, accept, refit     codes/funcs.py:1226-1306, bsr_class.py:195-252
//   k_eval_trees / k_predict   allcal / BSR.predict for arbitrary trees
#pragma once
#include "bsr_common.cuh"
#include "bsr_eval.cuh"
#include "bsr_propose.cuh"
#include "bsr_rng.cuh"
#include "bsr_solve.cuh"

struct ProposeCtx {
  uint64_t seed;
  int64_t chain_offset;
  int64_t sweep;
  const double* tape;
  const int64_t* tape_off;
  int steps, step_base;
  double* rec;        // [C][rec_steps][rec_cap] recorded draws (MODE 2)
  int* rec_count;     // [C][rec_steps]
  int rec_steps, rec_cap, rec_base;
};

template <int MODE>
__global__ void k_init_chains(ChainState st, PriorTables pt, uint64_t seed, int64_t chain_offset) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= st.C * st.K) return;
  int c = g / st.K, k = g % st.K;
  Draws<MODE> dr;
  dr.init_philox(seed, (uint64_t)(chain_offset + c), (uint32_t)k, 0u);
  size_t slot = (size_t)g * BSR_MAXN;
  double sa, sb;
  init_tree<MODE>(pt, dr, st.tok[0] + slot, st.pa[0] + slot, st.pb[0] + slot, st.nn[0] + g, sa, sb);
  st.sa[g] = sa; st.sb[g] = sb;
  st.which[g] = 0; st.report_which[g] = 0;
  st.nn[1][g] = 0;
  if (k == 0) {
    Draws<MODE> ds;
    ds.init_philox(seed, (uint64_t)(chain_offset + c), (uint32_t)st.K, 0u);
    st.sigma[c] = ds.invgamma(1);                                             // bsr_class.py:123
  }
}

template <int MODE>
__global__ void k_propose(ChainState st, PriorTables pt, ProposeCtx pc) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= st.C * st.K) return;
  int c = g / st.K, k = g % st.K;
  if (st.done[c]) return;
  Draws<MODE> dr;
  dr.init_philox(pc.seed, (uint64_t)(pc.chain_offset + c), (uint32_t)(pc.sweep * st.K + k), 1u);
  if (MODE == 1) {
    size_t s = (size_t)c * pc.steps + pc.step_base + k;
    dr.init_tape(pc.tape, (int)pc.tape_off[s], (int)pc.tape_off[s + 1]);
  }
  if (MODE == 2 && pc.rec != nullptr && pc.rec_base + k < pc.rec_steps)
    dr.init_record(pc.rec + ((size_t)c * pc.rec_steps + pc.rec_base + k) * pc.rec_cap, pc.rec_cap);
  const int w = st.which[g];
  const size_t slot = (size_t)g * BSR_MAXN;
  PropInfo info;
  propose_one<MODE>(pt, st.tok[w] + slot, st.pa[w] + slot, st.pb[w] + slot, st.nn[w][g], st.sa[g], st.sb[g], dr,
                    st.tok[w ^ 1] + slot, st.pa[w ^ 1] + slot, st.pb[w ^ 1] + slot, st.nn[w ^ 1] + g, info);
  st.pinfo[g] = info;
  if (MODE == 2 && pc.rec != nullptr && pc.rec_base + k < pc.rec_steps)
    pc.rec_count[(size_t)c * pc.rec_steps + pc.rec_base + k] = dr.pos;
}

template <typename T>
struct EvalCtx {
  const T* X;      // column-major [d][ld]
  const T* y;      // [n]
  int64_t n, ld;
  double* sums;    // [C][n_sum]
  double* maxs;    // [C][P]
  int* need64;     // [C] set by the fp32 pass when a column left the fp32 range; consumed by the fp64 pass
  int only_flagged;  // fp64 pass: 1 = only chains with need64 set
  int init_only;     // evaluate the K live trees only
  int tpc;           // threads per chain: 32 (warp per chain) or blockDim.x (block per chain)
};

// Shared-memory footprint per chain group: tokens + lt parameters of the P trees, and per block the staged column
// values and the cross-warp reduction scratch.
template <typename T>
__host__ __device__ inline size_t eval_smem_bytes(int P, int R, int threads, int tpc) {
  int groups = threads / tpc;
  size_t per_group = (size_t)P * BSR_MAXN * (sizeof(uint32_t) + 2 * sizeof(T)) + (size_t)P * sizeof(int);
  per_group = (per_group + 15) / 16 * 16;
  size_t cv = (size_t)P * R * threads * sizeof(T);
  size_t red = (tpc > 32) ? (size_t)(threads / 32) * (gram_n_sum(P) + P + 1) * sizeof(double) : 0;
  return groups * per_group + cv + red + 64;
}

template <typename T, int KT, int R>
__global__ void __launch_bounds__(256) k_eval(ChainState st, EvalCtx<T> ec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K = (KT > 0) ? KT : st.K;
  const int P = 2 * K;
  const int tpc = ec.tpc;
  const int groups = blockDim.x / tpc;
  const int grp = threadIdx.x / tpc;
  const int lane = threadIdx.x % tpc;       // index inside the chain group
  const int c = blockIdx.x * groups + grp;
  const bool block_mode = tpc > 32;

  size_t per_group = (size_t)P * BSR_MAXN * (sizeof(uint32_t) + 2 * sizeof(T)) + (size_t)P * sizeof(int);
  per_group = (per_group + 15) / 16 * 16;
  unsigned char* gbase = smem_raw + (size_t)grp * per_group;
  T* s_a = reinterpret_cast<T*>(gbase);
  T* s_b = s_a + (size_t)P * BSR_MAXN;
  uint32_t* s_tok = reinterpret_cast<uint32_t*>(s_b + (size_t)P * BSR_MAXN);
  int* s_m = reinterpret_cast<int*>(s_tok + (size_t)P * BSR_MAXN);
  T* s_cv = reinterpret_cast<T*>(smem_raw + (size_t)groups * per_group);
  double* s_red = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(s_cv) + (size_t)P * R * blockDim.x * sizeof(T));

  bool active = (c < st.C);
  if (active && !ec.init_only && st.done[c]) active = false;
  if (active && ec.only_flagged && !ec.need64[c]) active = false;
  if (block_mode) { if (!active) return; }   // one chain per block: uniform exit
  else if (!active) return;                  // warp per chain: whole warp exits together

  // ---- stage the P trees of this chain in shared memory ----
  for (int p = 0; p < P; ++p) {
    const int k = (p < K) ? p : p - K;
    const int g = c * K + k;
    const int w = st.which[g] ^ (p < K ? 0 : 1);
    int m = st.nn[w][g];
    if (p >= K && (ec.init_only || (st.pinfo[g].flags & PF_CAPACITY))) m = 0;
    if (lane == 0) s_m[p] = m;
    const size_t slot = (size_t)g * BSR_MAXN;
    for (int j = lane; j < m; j += tpc) {
      s_tok[p * BSR_MAXN + j] = st.tok[w][slot + j];
      s_a[p * BSR_MAXN + j] = (T)st.pa[w][slot + j];
      s_b[p * BSR_MAXN + j] = (T)st.pb[w][slot + j];
    }
  }
  if (block_mode) __syncthreads(); else __syncwarp();

  constexpr int PC = (KT > 0) ? 2 * KT : 1;                 // compile-time column count (register Gram)
  constexpr int NG = (KT > 0) ? PC * (PC + 1) / 2 : 1;
  double accG[NG], accY[PC], accS[PC], accM[PC];
  // generic path (KT == 0): accumulators in local memory
  double genG[(KT > 0) ? 1 : (2 * BSR_MAXK) * (2 * BSR_MAXK + 1) / 2];
  double genY[(KT > 0) ? 1 : 2 * BSR_MAXK], genS[(KT > 0) ? 1 : 2 * BSR_MAXK], genM[(KT > 0) ? 1 : 2 * BSR_MAXK];
  if (KT > 0) {
#pragma unroll
    for (int i = 0; i < NG; ++i) accG[i] = 0.0;
#pragma unroll
    for (int i = 0; i < PC; ++i) { accY[i] = 0.0; accS[i] = 0.0; accM[i] = 0.0; }
  } else {
    for (int i = 0; i < P * (P + 1) / 2; ++i) genG[i] = 0.0;
    for (int i = 0; i < P; ++i) { genY[i] = 0.0; genS[i] = 0.0; genM[i] = 0.0; }
  }
  unsigned bad = 0;

  const int64_t n = ec.n;
  T* my_cv = s_cv + threadIdx.x;
  const int cvs = blockDim.x;   // stride between consecutive (p, r) entries
  for (int64_t base = 0; base < n; base += (int64_t)tpc * R) {
    int64_t rows[R];
    bool valid[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int64_t row = base + (int64_t)r * tpc + lane;
      valid[r] = row < n;
      rows[r] = valid[r] ? row : n - 1;
    }
    for (int p = 0; p < P; ++p) {
      T acc[R];
      const int m = s_m[p];
      if (m > 0) {
        eval_tree_rows<T, R>(s_tok + p * BSR_MAXN, s_a + p * BSR_MAXN, s_b + p * BSR_MAXN, m, ec.X, ec.ld, rows, acc);
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = (T)0;
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (!OpMath<T>::finite(acc[r])) bad |= (1u << p);
        my_cv[(p * R + r) * cvs] = acc[r];
      }
    }
    T yv[R];
#pragma unroll
    for (int r = 0; r < R; ++r) yv[r] = __ldg(ec.y + rows[r]);
    if (KT > 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (!valid[r]) continue;
        double v[PC];
#pragma unroll
        for (int i = 0; i < PC; ++i) v[i] = (double)my_cv[(i * R + r) * cvs];
        const double yr = (double)yv[r];
        int q = 0;
#pragma unroll
        for (int i = 0; i < PC; ++i) {
#pragma unroll
          for (int j = i; j < PC; ++j) { accG[q] = fma(v[i], v[j], accG[q]); ++q; }
          accY[i] = fma(v[i], yr, accY[i]);
          accS[i] += v[i];
          accM[i] = fmax(accM[i], fabs(v[i]));
        }
      }
    } else {
      for (int r = 0; r < R; ++r) {
        if (!valid[r]) continue;
        const double yr = (double)yv[r];
        int q = 0;
        for (int i = 0; i < P; ++i) {
          const double vi = (double)my_cv[(i * R + r) * cvs];
          for (int j = i; j < P; ++j) { genG[q] = fma(vi, (double)my_cv[(j * R + r) * cvs], genG[q]); ++q; }
          genY[i] = fma(vi, yr, genY[i]);
          genS[i] += vi;
          genM[i] = fmax(genM[i], fabs(vi));
        }
      }
    }
  }

  // ---- reduce over the chain group and write the record ----
  const int n_sum = gram_n_sum(P);
  double* out_s = ec.sums + (size_t)c * n_sum;
  double* out_m = ec.maxs + (size_t)c * P;
  const int wlane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nacc = n_sum + P;
  // reduce in the order [G..., Y..., S..., M...]
  auto get_acc = [&](int i) -> double {
    if (KT > 0) return 0.0;   // unused in the compile-time path
    const int ng = P * (P + 1) / 2;
    if (i < ng) return genG[i];
    if (i < ng + P) return genY[i - ng];
    if (i < ng + 2 * P) return genS[i - ng - P];
    return genM[i - ng - 2 * P];
  };
  unsigned bad_all = bad;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bad_all |= __shfl_xor_sync(0xffffffffu, bad_all, o);
  if (KT > 0) {
    double* dst = block_mode ? (s_red + (size_t)wid * (nacc + 1)) : nullptr;
    int q = 0;
#pragma unroll
    for (int i = 0; i < NG; ++i, ++q) { double v = warp_sum(accG[i]); if (wlane == 0) { if (block_mode) dst[q] = v; else out_s[q] = v; } }
#pragma unroll
    for (int i = 0; i < PC; ++i, ++q) { double v = warp_sum(accY[i]); if (wlane == 0) { if (block_mode) dst[q] = v; else out_s[q] = v; } }
#pragma unroll
    for (int i = 0; i < PC; ++i, ++q) { double v = warp_sum(accS[i]); if (wlane == 0) { if (block_mode) dst[q] = v; else out_s[q] = v; } }
#pragma unroll
    for (int i = 0; i < PC; ++i) {
      double v = warp_max(accM[i]);
      if ((bad_all >> i) & 1u) v = INFINITY;
      if (wlane == 0) { if (block_mode) dst[q + i] = v; else out_m[i] = v; }
    }
    if (block_mode && wlane == 0) dst[nacc] = (double)bad_all;
  } else {
    double* dst = block_mode ? (s_red + (size_t)wid * (nacc + 1)) : nullptr;
    for (int i = 0; i < nacc; ++i) {
      double a = get_acc(i);
      double v = (i < n_sum) ? warp_sum(a) : warp_max(a);
      if (i >= n_sum && ((bad_all >> (i - n_sum)) & 1u)) v = INFINITY;
      if (wlane == 0) { if (block_mode) dst[i] = v; else if (i < n_sum) out_s[i] = v; else out_m[i - n_sum] = v; }
    }
    if (block_mode && wlane == 0) dst[nacc] = (double)bad_all;
  }
  if (block_mode) {
    __syncthreads();
    const int nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < nacc; i += blockDim.x) {
      double v = (i < n_sum) ? 0.0 : 0.0;
      for (int w = 0; w < nw; ++w) {
        double x = s_red[(size_t)w * (nacc + 1) + i];
        v = (i < n_sum) ? v + x : fmax(v, x);
      }
      if (i < n_sum) out_s[i] = v; else out_m[i - n_sum] = v;
    }
    if (threadIdx.x == 0) {
      unsigned b = 0;
      for (int w = 0; w < nw; ++w) b |= (unsigned)s_red[(size_t)w * (nacc + 1) + nacc];
      bad_all = b;
    }
  }
  if ((block_mode ? threadIdx.x == 0 : wlane == 0)) {
    if (sizeof(T) == 4) { if (bad_all) ec.need64[c] = 1; }
    else if (ec.only_flagged) { ec.need64[c] = 0; st.counters[(size_t)c * BSR_N_COUNTERS + BSR_CNT_FP64_SWEEPS] += 1; }
  }
}

template <int MODE>
__global__ void k_resolve(ChainState st, ResolveCtx rc, const double* sums, const double* maxs, int init_only) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= st.C) return;
  const int P = 2 * st.K;
  resolve_chain<MODE>(st, rc, c, sums + (size_t)c * gram_n_sum(P), maxs + (size_t)c * P, init_only != 0);
}

// allcal for arbitrary trees: out[t][row] (float64), one block per tree.
template <typename T>
__global__ void k_eval_trees(const uint32_t* tok, const double* pa, const double* pb, const int* nn, const T* X, int64_t n,
                             int64_t ld, double* out) {
  __shared__ uint32_t s_tok[BSR_MAXN];
  __shared__ T s_a[BSR_MAXN], s_b[BSR_MAXN];
  const int t = blockIdx.x;
  const int m = nn[t];
  for (int j = threadIdx.x; j < m; j += blockDim.x) {
    s_tok[j] = tok[(size_t)t * BSR_MAXN + j];
    s_a[j] = (T)pa[(size_t)t * BSR_MAXN + j];
    s_b[j] = (T)pb[(size_t)t * BSR_MAXN + j];
  }
  __syncthreads();
  for (int64_t row = threadIdx.x; row < n; row += blockDim.x) {
    int64_t rows[1] = {row};
    T acc[1];
    eval_tree_rows<T, 1>(s_tok, s_a, s_b, m, X, ld, rows, acc);
    out[(size_t)t * n + row] = (double)acc[0];
  }
}

// BSR.predict (bsr_class.py:53-68): out[row] = beta0 + sum_k beta_k * tree_k(X[row]), float64 evaluation.
__global__ void k_predict(const uint32_t* tok, const double* pa, const double* pb, const int* nn, int K, const double* beta,
                          const double* X, int64_t n, int64_t ld, double* out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_a = reinterpret_cast<double*>(smem_raw);
  double* s_b = s_a + (size_t)K * BSR_MAXN;
  uint32_t* s_tok = reinterpret_cast<uint32_t*>(s_b + (size_t)K * BSR_MAXN);
  for (int j = threadIdx.x; j < K * BSR_MAXN; j += blockDim.x) {
    int k = j / BSR_MAXN, i = j % BSR_MAXN;
    if (i < nn[k]) { s_tok[j] = tok[j]; s_a[j] = pa[j]; s_b[j] = pb[j]; }
  }
  __syncthreads();
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (int64_t)gridDim.x * blockDim.x) {
    int64_t rows[1] = {row};
    double v = beta[0];
    for (int k = 0; k < K; ++k) {
      double acc[1];
      eval_tree_rows<double, 1>(s_tok + k * BSR_MAXN, s_a + k * BSR_MAXN, s_b + k * BSR_MAXN, nn[k], X, ld, rows, acc);
      v += beta[k + 1] * acc[0];
    }
    out[row] = v;
  }
}

template <typename TI, typename TO>
__global__ void k_convert(const TI* in, TO* out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (TO)in[i];
}

// row-major float64 (n x d) -> column-major T (d x ld)
template <typename TO>
__global__ void k_transpose_in(const double* in, TO* out, int64_t n, int d, int64_t ld) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * d; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / d;
    int col = (int)(i % d);
    out[(int64_t)col * ld + row] = (TO)in[i];
  }
}

// y statistics: sum(y), y'y in fp64 (single block, deterministic order per launch geometry)
__global__ void k_y_stats(const double* y, int64_t n, double* out2) {
  __shared__ double s1[32], s2[32];
  double a = 0.0, b = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { double v = y[i]; a += v; b = fma(v, v, b); }
  a = warp_sum(a); b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = a; s2[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double x = 0.0, z = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { x += s1[w]; z += s2[w]; }
    out2[0] = x; out2[1] = z;
  }
}

__global__ void k_count_done(const int* done, int C, int* out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  int v = (c < C) ? (done[c] != 0) : 0;
  unsigned b = __ballot_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(out, __popc(b));
}
