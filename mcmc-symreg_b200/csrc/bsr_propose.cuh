// Structural proposal + reversible-jump auxiliary step + tree prior, one thread per (chain, tree).
//
// Restates, on pre-order token arrays, what the reference does on pointer trees:
//   grow      codes/funcs.py:74-119      fStruc   codes/funcs.py:349-398
//   Prop      codes/funcs.py:406-923     auxProp  codes/funcs.py:935-1138
// Every index the reference draws (Term[pod], Nterm[pod], detcd[det_od], Tree[ins_ind]) is a position in
// genList order, i.e. a token index here; the seven moves are splices of token spans.  The reference's quirks
// (plus Q22: b->u Qinv divides by the node count) are reproduced on purpose.
#pragma once
#include "bsr_common.cuh"
#include "bsr_rng.cuh"

__device__ __forceinline__ int span_end(const uint32_t* tk, int i) {
  int need = 1;
  while (need > 0) { need += op_arity(tok_op(tk[i])) - 1; ++i; }
  return i;
}

// detcd (funcs.py:454-468): non-terminals, except a root all of whose children are terminal.
__device__ __forceinline__ int det_count(const uint32_t* tk, int m, int n_nonterm) {
  int a0 = op_arity(tok_op(tk[0]));
  bool excl = false;
  if (a0 == 1) excl = tok_op(tk[1]) == OP_LEAF;
  else if (a0 == 2) excl = tok_op(tk[1]) == OP_LEAF && tok_op(tk[span_end(tk, 1)]) == OP_LEAF;
  return n_nonterm - (excl ? 1 : 0);
}

// Which of the seven moves the draw `test` selects (funcs.py:475-483), from the counts of the current tree.  Mirrors
// the threshold chain of propose_one expression by expression (k_wclassify sorts proposals by it).
__device__ __forceinline__ int select_move(int L, int Nt, int D, double test) {
  const double p_stay = 0.25 * L / (L + 3);
  const double p_grow = (1 - p_stay) * fmin(1.0, 4.0 / (Nt + 2)) / 3;
  const double p_prune = (1 - p_stay) / 3 - p_grow;
  const double p_detr = (1 - p_stay) * (1.0 / 3) * D / (3 + D);
  const double p_trans = (1 - p_stay) / 3 - p_detr;
  const double p_rop = (1 - p_stay) / 6;
  if (test <= p_stay) return MV_STAY;
  if (test <= p_stay + p_grow) return MV_GROW;
  if (test <= p_stay + p_grow + p_prune) return MV_PRUNE;
  if (test <= p_stay + p_grow + p_prune + p_detr) return MV_DETR;
  if (test <= p_stay + p_grow + p_prune + p_detr + p_trans) return MV_TRANS;
  if (test <= p_stay + p_grow + p_prune + p_detr + p_trans + p_rop) return MV_ROP;
  return MV_RFEAT;
}

// log p(T,M) (first component of fStruc) of the token span tk[lo,hi) given per-slot depths.
static __device__ __noinline__ double fstruc_span(const PriorTables& pt, const uint32_t* tk, const uint8_t* dp, int lo, int hi) {
  double ll = 0.0;
#pragma unroll 1
  for (int j = lo; j < hi; ++j) {
    int o = tok_op(tk[j]), d = dp[j];
    if (o == OP_LEAF) ll += pt.log1m[d] - pt.lognf;
    else ll += (d == 0 ? 0.0 : pt.logsplit[d]) + pt.logw[tok_oi(tk[j])];
  }
  return ll;
}

// fStruc of a whole tree with depths recomputed from the root (upDepth, funcs.py:298-307):
// ll = log p(T,M), lp = log p(Theta | T, sigma_a, sigma_b) (funcs.py:372-376).
static __device__ __noinline__ void fstruc_tree(const PriorTables& pt, const uint32_t* tk, int m, const double* a, const double* b,
                                            double s_a, double s_b, double& ll, double& lp) {
  uint8_t pend[BSR_MAXN + 2];
  int sp = 0;
  pend[sp++] = 0;
  ll = 0.0; lp = 0.0;
  double cst = 0.0;
  bool have_cst = false;          // two double-precision logs: only trees with an lt node (one operator in ten) need them
#pragma unroll 1
  for (int j = 0; j < m; ++j) {
    int d = pend[--sp];
    int o = tok_op(tk[j]);
    if (o == OP_LEAF) {
      ll += pt.log1m[d] - pt.lognf;
    } else {
      ll += (d == 0 ? 0.0 : pt.logsplit[d]) + pt.logw[tok_oi(tk[j])];
      if (o == OP_LT) {
        if (!have_cst) { cst = -0.5 * log(2.0 * 3.141592653589793 * s_a) - 0.5 * log(2.0 * 3.141592653589793 * s_b); have_cst = true; }
        double da = a[j] - 1.0, db = b[j];
        lp -= da * da / (2.0 * s_a);
        lp -= db * db / (2.0 * s_b);
        lp += cst;
      }
      pend[sp++] = (uint8_t)(d + 1);
      if (o >= OP_ADD) pend[sp++] = (uint8_t)(d + 1);
    }
  }
}

// grow (funcs.py:74-119): append a subtree rooted at depth0 to dst starting at pos; at most `limit` tokens may
// be occupied in dst.  Accumulates log p(T,M) of the grown subtree in fs.  Draw order per node: depth>0: U,
// terminal => RI, RI (second kept) / operator => CH; depth 0: CH; lt => N(a), N(b) before descending.
template <int MODE>
__device__ __noinline__ int grow_tokens(const PriorTables& pt, int depth0, double s_a, double s_b, Draws<MODE>& dr, uint32_t* dst,
                           int pos, int limit, double& fs, bool& overflow) {
  uint8_t pend[BSR_MAXN + 2];
  int sp = 0;
  pend[sp++] = (uint8_t)depth0;
  fs = 0.0;
  const double sd_a = sqrt(s_a), sd_b = sqrt(s_b);
  while (sp > 0) {
    int d = pend[--sp];
    if (pos >= limit || d >= BSR_MAXN) { overflow = true; return pos; }
    bool terminal = false;
    int oi = 0;
    if (d > 0) {
      double test = dr.uniform();
      if (test > pt.psplit[d]) { (void)dr.randint(0, pt.n_feature); terminal = true; }
      else oi = dr.choice(pt);
    } else {
      oi = dr.choice(pt);
    }
    if (terminal) {
      int f = dr.randint(0, pt.n_feature);
      dst[pos++] = make_tok(OP_LEAF, 0, f);
      fs += pt.log1m[d] - pt.lognf;
    } else {
      int o = pt.ops[oi];
      if (o == OP_LT) { (void)dr.normal(1.0, sd_a); (void)dr.normal(0.0, sd_b); }   // overwritten by auxProp
      dst[pos++] = make_tok(o, oi, 0);
      fs += (d == 0 ? 0.0 : pt.logsplit[d]) + pt.logw[oi];
      pend[sp++] = (uint8_t)(d + 1);
      if (o >= OP_ADD) pend[sp++] = (uint8_t)(d + 1);
    }
  }
  return pos;
}

static __device__ __noinline__ double log_ig_pdf(double x, double a, double lgamma_a) {
  return -(a + 1.0) * log(x) - 1.0 / x - lgamma_a;   // log invgamma.pdf(x, a)
}
static __device__ __noinline__ double log_norm_pdf0(double x, double var) {   // log N(x; 0, sqrt(var))
  return -0.5 * x * x / var - 0.5 * log(var) - 0.9189385332046727;
}
static __device__ __noinline__ double norm_pdf(double x, double loc, double var) {
  double z = x - loc;
  return exp(-0.5 * z * z / var) / sqrt(2.0 * 3.141592653589793 * var);
}

// Token-array helpers of propose_one.  Out of line and not unrolled: they are called from ~30 places, and as inlined,
// unrolled lambdas they were 28 % of a 160 KB kernel (instruction-cache misses are what bounds the proposal kernels).
static __device__ __noinline__ int span_copy(uint32_t* dst, const uint32_t* src, int pos, int lo, int hi) {
#pragma unroll 1
  for (int j = lo; j < hi; ++j) dst[pos++] = src[j];
  return pos;
}
static __device__ __noinline__ void count_lt_leaf(const uint32_t* t, int n, int& n_lt, int& n_leaf) {
  int a = 0, b = 0;
#pragma unroll 1
  for (int j = 0; j < n; ++j) { const int o = tok_op(t[j]); a += (o == OP_LT); b += (o == OP_LEAF); }
  n_lt = a; n_leaf = b;
}
static __device__ __noinline__ int nth_node(const uint32_t* t, int n, int k, bool want_leaf) {   // k-th terminal / non-terminal in pre-order
#pragma unroll 1
  for (int j = 0; j < n; ++j) if ((tok_op(t[j]) == OP_LEAF) == want_leaf) { if (k == 0) return j; --k; }
  return 0;
}

// One proposal for one tree: Prop + (sigma ~ IG(4)) + auxProp + both prior terms (funcs.py:1188-1210,1241-1289).
// Writes the proposed tree to (ntok, na, nb, *nn_out) and the scalars logR needs to `info`.
template <int MODE>
__device__ void propose_one(const PriorTables& pt, const uint32_t* __restrict__ otok, const double* __restrict__ oa,
                            const double* __restrict__ ob, int m, double sa, double sb, Draws<MODE>& dr,
                            uint32_t* __restrict__ ntok, double* __restrict__ na, double* __restrict__ nb, int* nn_out,
                            PropInfo& info, const double* __restrict__ live_fs = nullptr) {
  uint32_t tk[BSR_MAXN], nt[BSR_MAXN];
  uint8_t sz[BSR_MAXN], dp[BSR_MAXN], lts[BSR_MAXN];
#pragma unroll 1
  for (int j = 0; j < m; ++j) tk[j] = otok[j];

  // subtree sizes (right-to-left) and depths (left-to-right)                     funcs.py:414-418
#pragma unroll 1
  for (int j = m - 1; j >= 0; --j) {
    int ar = op_arity(tok_op(tk[j]));
    int s = 1;
    if (ar >= 1) s += sz[j + 1];
    if (ar == 2) s += sz[j + 1 + sz[j + 1]];
    sz[j] = (uint8_t)s;
  }
  dp[0] = 0;
  int L = 0, T = 0;
#pragma unroll 1
  for (int j = 0; j < m; ++j) {
    int o = tok_op(tk[j]);
    int ar = op_arity(o);
    if (ar >= 1) dp[j + 1] = dp[j] + 1;
    if (ar == 2) dp[j + 1 + sz[j + 1]] = dp[j] + 1;
    if (o == OP_LEAF) ++T;
    if (o == OP_LT) lts[L++] = (uint8_t)j;
  }
  const int Nt = m - T;
  const int D = det_count(tk, m, Nt);
  const int nf = pt.n_feature;

  // move probabilities                                                          funcs.py:475-480
  const double p_stay = 0.25 * L / (L + 3);
  const double p_grow = (1 - p_stay) * fmin(1.0, 4.0 / (Nt + 2)) / 3;
  const double p_prune = (1 - p_stay) / 3 - p_grow;
  const double p_detr = (1 - p_stay) * (1.0 / 3) * D / (3 + D);
  const double p_trans = (1 - p_stay) / 3 - p_detr;
  const double p_rop = (1 - p_stay) / 6;

  const double test = dr.uniform();                                            // funcs.py:483
  int change = CH_NONE, move;
  double Q = 1.0, Qinv = 1.0;
  int mp = m;               // size of the proposed tree
  int changed_ln = -1;      // old slot of an lt node whose operator field the reference overwrites
  bool overflow = false;

  auto copy_span = [&](int pos, int lo, int hi) { return span_copy(nt, tk, pos, lo, hi); };
  auto count_new = [&](int& Lp, int& Tp) { count_lt_leaf(nt, mp, Lp, Tp); };
  auto nth = [&](int k, bool want_leaf) { return nth_node(tk, m, k, want_leaf); };

  if (test <= p_stay) {                                                        // stay  funcs.py:490-500
    move = MV_STAY;
    Q = Qinv = p_stay;
    copy_span(0, 0, m);
    const double sd_a = sqrt(sa), sd_b = sqrt(sb);
#pragma unroll 1
    for (int i = 0; i < L; ++i) { (void)dr.normal(1.0, sd_a); (void)dr.normal(1.0, sd_b); }   // Q5; overwritten below
  } else if (test <= p_stay + p_grow) {                                        // grow  funcs.py:503-536
    move = MV_GROW;
    int pod = dr.randint(0, T);
    int i = nth(pod, true);
    int pos = copy_span(0, 0, i);
    double fs;
    int tail = m - (i + 1);
    int pos2 = grow_tokens(pt, dp[i], sa, sb, dr, nt, pos, BSR_MAXN - tail, fs, overflow);
    if (!overflow) {
      bool root_leaf = tok_op(nt[pos]) == OP_LEAF;
      mp = copy_span(pos2, i + 1, m);
      if (root_leaf) { Q = Qinv = 1.0; }
      else {
        Q = p_grow * exp(fs) / T;
        int Lp, Tp; count_new(Lp, Tp);
        int Ntp = mp - Tp;
        double new_p = (1 - 0.25 * Lp / (Lp + 3)) * (1 - fmin(1.0, 4.0 / (Ntp + 2))) / 3;
        Qinv = new_p / (double)max(1, mp - Tp - 1);
        if (Lp > L) change = CH_EXPANSION;
      }
    }
  } else if (test <= p_stay + p_grow + p_prune) {                              // prune funcs.py:539-579
    move = MV_PRUNE;
    int pod = dr.randint(1, Nt);
    int i = nth(pod, false);
    double fs = fstruc_span(pt, tk, dp, i, i + sz[i]);
    int p_lt = 0;
#pragma unroll 1
    for (int j = i; j < i + sz[i]; ++j) p_lt += (tok_op(tk[j]) == OP_LT);
    if (p_lt > 0) change = CH_SHRINKAGE;
    if (tok_op(tk[i]) == OP_LT) changed_ln = i;
    int pos = copy_span(0, 0, i);
    nt[pos++] = make_tok(OP_LEAF, 0, dr.randint(0, nf));
    mp = copy_span(pos, i + sz[i], m);
    int Lp, Tp; count_new(Lp, Tp);
    int Ntp = mp - Tp;
    Q = p_prune / (double)((Nt - 1) * nf);
    double pg = 1 - 0.25 * Lp / (Lp + 3) * 0.75 * fmin(1.0, 4.0 / (Ntp + 2));   // literal precedence (Q8)
    Qinv = pg * exp(fs) / Tp;
  } else if (test <= p_stay + p_grow + p_prune + p_detr) {                     // detransform funcs.py:582-673
    move = MV_DETR;
    int det_od = dr.randint(0, D);
    // det_od-th candidate: non-terminals in pre-order, skipping an excluded root
    int i = 0;
    {
      bool root_excl = (D != Nt);
      int k = det_od;
#pragma unroll 1
      for (int j = 0; j < m; ++j) {
        if (tok_op(tk[j]) == OP_LEAF) continue;
        if (j == 0 && root_excl) continue;
        if (k == 0) { i = j; break; }
        --k;
      }
    }
    Q = p_detr / D;
    int keep_lo, keep_hi, cut_lo = -1, cut_hi = -1;
    if (op_arity(tok_op(tk[i])) == 1) { keep_lo = i + 1; keep_hi = i + sz[i]; }
    else {
      int l_lo = i + 1, l_hi = i + 1 + sz[i + 1], r_lo = l_hi, r_hi = i + sz[i];
      bool keep_left;
      if (i == 0 && tok_op(tk[l_lo]) == OP_LEAF) keep_left = false;            // funcs.py:597-599
      else if (i == 0 && tok_op(tk[r_lo]) == OP_LEAF) keep_left = true;        // funcs.py:600-602
      else { double aa = dr.uniform(); keep_left = (aa <= 0.5); Q = Q / 2; }   // funcs.py:603-611, 623-640
      if (keep_left) { keep_lo = l_lo; keep_hi = l_hi; cut_lo = r_lo; cut_hi = r_hi; }
      else { keep_lo = r_lo; keep_hi = r_hi; cut_lo = l_lo; cut_hi = l_hi; }
    }
    int pos = copy_span(0, 0, i);
    pos = copy_span(pos, keep_lo, keep_hi);
    mp = copy_span(pos, i + sz[i], m);
    int Lp, Tp; count_new(Lp, Tp);
    if (Lp < L) change = CH_SHRINKAGE;
    double new_pstay = 0.25 * Lp / (Lp + 3);
    int Dp = det_count(nt, mp, mp - Tp);
    double new_pdetr = (1 - new_pstay) * (1.0 / 3) * Dp / (Dp + 3);
    double new_ptr = (1 - new_pstay) / 3 - new_pdetr;
    Qinv = new_ptr * pt.w[tok_oi(tk[i])] / mp;
    if (cut_lo >= 0) Qinv = Qinv * exp(fstruc_span(pt, tk, dp, cut_lo, cut_hi));   // cut keeps OLD depths
  } else if (test <= p_stay + p_grow + p_prune + p_detr + p_trans) {           // transform funcs.py:679-786
    move = MV_TRANS;
    int i = dr.randint(0, m);
    int ins_oi = dr.choice(pt);
    int ins_op = pt.ops[ins_oi];
    if (m + 1 > BSR_MAXN) overflow = true;
    else {
      int pos = copy_span(0, 0, i);
      nt[pos++] = make_tok(ins_op, ins_oi, 0);
      pos = copy_span(pos, i, i + sz[i]);
      if (ins_op < OP_ADD) {
        if (ins_op == OP_LT) change = CH_EXPANSION;
        mp = copy_span(pos, i + sz[i], m);
        Q = p_trans * pt.w[ins_oi] / m;
      } else {
        double fs;
        int tail = m - (i + sz[i]);
        int pos2 = grow_tokens(pt, dp[i] + 1, sa, sb, dr, nt, pos, BSR_MAXN - tail, fs, overflow);
        if (!overflow) {
          mp = copy_span(pos2, i + sz[i], m);
          Q = p_trans * pt.w[ins_oi] * exp(fs) / m;
        }
      }
      if (!overflow) {
        int Lp, Tp; count_new(Lp, Tp);
        if (Lp > L) change = CH_EXPANSION;
        double new_pstay = 0.25 * Lp / (Lp + 3);
        int Dp = det_count(nt, mp, mp - Tp);
        double new_pdetr = (1 - new_pstay) * (1.0 / 3) * Dp / (Dp + 3);
        Qinv = new_pdetr / Dp;
        if (ins_op >= OP_ADD) {
          // new node at slot i: left child = old span (slot i+1), right child = grown subtree
          if (tok_op(nt[i + 1]) != OP_LEAF && tok_op(nt[i + 1 + sz[i]]) != OP_LEAF) Qinv = Qinv / 2;
        }
      }
    }
  } else if (test <= p_stay + p_grow + p_prune + p_detr + p_trans + p_rop) {   // reassignOperator funcs.py:791-903
    move = MV_ROP;
    int pod = dr.randint(0, Nt);
    int i = nth(pod, false);
    int last_op = tok_op(tk[i]), last_oi = tok_oi(tk[i]);                      // op_ind never refreshed (Q7)
    int new_oi = dr.choice(pt);
    int new_op = pt.ops[new_oi];
    uint32_t retag = make_tok(new_op, last_oi, 0);
    if (last_op < OP_ADD) {
      if (new_op < OP_ADD) {                                                   // u -> u  funcs.py:810-825
        copy_span(0, 0, m);
        nt[i] = retag;
        if (last_op == OP_LT) { if (new_op != OP_LT) { change = CH_SHRINKAGE; changed_ln = i; } }
        else if (new_op == OP_LT) change = CH_EXPANSION;
        Q = pt.w[new_oi];
        Qinv = pt.w[last_oi];
      } else {                                                                 // u -> b  funcs.py:827-860
        if (last_op == OP_LT) changed_ln = i;
        int pos = copy_span(0, 0, i + sz[i]);
        nt[i] = retag;
        double fs;
        int tail = m - (i + sz[i]);
        int pos2 = grow_tokens(pt, dp[i] + 1, sa, sb, dr, nt, pos, BSR_MAXN - tail, fs, overflow);
        if (!overflow) {
          mp = copy_span(pos2, i + sz[i], m);
          Q = p_rop * exp(fs) * pt.w[new_oi] / Nt;
          int Lp, Tp; count_new(Lp, Tp);
          double new_p0 = (double)Lp / (4 * (Lp + 3));
          Qinv = 0.125 * (1 - new_p0) * pt.w[last_oi] / (mp - Tp);
          if (Lp > L) change = CH_EXPANSION; else if (Lp < L) change = CH_SHRINKAGE;
        }
      }
    } else {
      if (new_op < OP_ADD) {                                                   // b -> u  funcs.py:867-894
        int lo = i + 1 + sz[i + 1], hi = i + sz[i];
        int p_lt = 0;
#pragma unroll 1
        for (int j = lo; j < hi; ++j) p_lt += (tok_op(tk[j]) == OP_LT);
        if (p_lt > 1) change = CH_SHRINKAGE;                                   // '>1' (Q12)
        else if (new_op == OP_LT && p_lt == 0) change = CH_EXPANSION;
        int pos = copy_span(0, 0, lo);
        nt[i] = retag;
        mp = copy_span(pos, hi, m);
        Q = p_rop * pt.w[new_oi] / Nt;
        int Lp, Tp; count_new(Lp, Tp);
        double new_p0 = (double)Lp / (4 * (Lp + 3));
        double fs = fstruc_span(pt, tk, dp, lo, hi);
        Qinv = 0.125 * (1 - new_p0) * exp(fs) * pt.w[last_oi] / mp;            // divides by m' (Q22)
      } else {                                                                 // b -> b  funcs.py:898-903
        copy_span(0, 0, m);
        nt[i] = retag;
        Q = pt.w[new_oi];
        Qinv = pt.w[last_oi];
      }
    }
  } else {                                                                     // reassignFeature funcs.py:907-917
    move = MV_RFEAT;
    int pod = dr.randint(0, T);
    int i = nth(pod, true);
    int fod = dr.randint(0, nf);
    copy_span(0, 0, m);
    nt[i] = make_tok(OP_LEAF, 0, fod);
    Q = Qinv = 1.0;
  }

  info.move = move;
  info.flags = 0;
  info.m_old = m;
  if (overflow) {   // tree would not fit the slot: counted reject, nothing else is needed
    info.flags = PF_CAPACITY;
    info.change = 0; info.Q = info.Qinv = 1.0; info.hratio = info.detjacob = 1.0;
    info.new_sigma = info.new_sa2 = info.new_sb2 = 1.0; info.ll_new = info.lp_new = info.fs_old = 0.0;
    info.m_new = 0; info.ndraws = dr.ndraws;
    *nn_out = 0;
    return;
  }

  const double new_sigma = dr.invgamma(4);                                     // funcs.py:1194-1195

  // ---- auxProp (funcs.py:935-1138); lt parameters are assigned by pre-order position ----
  uint8_t nl[BSR_MAXN];
  int Lp = 0;
#pragma unroll 1
  for (int j = 0; j < mp; ++j) if (tok_op(nt[j]) == OP_LT) nl[Lp++] = (uint8_t)j;
#pragma unroll 1
  for (int j = 0; j < mp; ++j) { ntok[j] = nt[j]; na[j] = 0.0; nb[j] = 0.0; }
  double new_sa2 = dr.invgamma(1), new_sb2 = dr.invgamma(1);                   // funcs.py:945-946
  double hratio = 1.0, detjacob = 1.0;
  const double LG1 = 0.0;   // lgamma(1)
  if (change == CH_SHRINKAGE) {                                                // funcs.py:950-1026
    int n_prsv = L - (changed_ln >= 0 ? 1 : 0);
    int n0 = n_prsv;
    if (Lp > n_prsv && changed_ln >= 0) n0 = n_prsv + 1;                       // top-up from the cut list (:964-966)
    double logh = log_ig_pdf(new_sa2, 1.0, LG1) + log_ig_pdf(new_sb2, 1.0, LG1);
    double loghstar = log_ig_pdf(sa, 1.0, LG1) + log_ig_pdf(sb, 1.0, LG1);
    const double sd_a = sqrt(new_sa2), sd_b = sqrt(new_sb2);
    int src = 0;
#pragma unroll 1
    for (int i = 0; i < n0; ++i) {
      int slot;
      if (i < n_prsv) { while (lts[src] == changed_ln) ++src; slot = lts[src++]; }
      else slot = changed_ln;
      double th_a = oa[slot], th_b = ob[slot];
      double ua = dr.normal(0.0, sd_a), ub = dr.normal(0.0, sd_b);
      logh += log_norm_pdf0(ua, new_sa2) + log_norm_pdf0(ub, new_sb2);
      loghstar += log_norm_pdf0(th_a - ua, sa) + log_norm_pdf0(th_b - ub, sb);
      if (i < Lp) { na[nl[i]] = th_a + ua; nb[nl[i]] = th_b + ub; }
    }
#pragma unroll 1
    for (int i = 0; i < L; ++i)                                                // all last_a appended to U* (Q10)
      loghstar += log_norm_pdf0(oa[lts[i]], sa) + log_norm_pdf0(ob[lts[i]], sb);
    hratio = exp(loghstar - logh);
    detjacob = exp2((double)(2 * n0));
  } else if (change == CH_EXPANSION) {                                         // funcs.py:1030-1110
    new_sa2 = dr.invgamma(1); new_sb2 = dr.invgamma(1);                        // second draw is the one used (Q11)
    const double sd_a = sqrt(new_sa2), sd_b = sqrt(new_sb2);
    double logh = log_ig_pdf(new_sa2, 1.0, LG1) + log_ig_pdf(new_sb2, 1.0, LG1);
    double loghstar = log_ig_pdf(sa, 1.0, LG1) + log_ig_pdf(sb, 1.0, LG1);
#pragma unroll 1
    for (int i = 0; i < L; ++i) {
      double th_a = oa[lts[i]], th_b = ob[lts[i]];
      double ua = dr.normal(0.0, sd_a), ub = dr.normal(0.0, sd_b);
      logh += log_norm_pdf0(ua, new_sa2) + log_norm_pdf0(ub, new_sb2);
      loghstar += log_norm_pdf0((th_a - ua) / 2, sa) + log_norm_pdf0((th_b - ub) / 2, sb);
      if (i < Lp) { na[nl[i]] = (th_a + ua) / 2; nb[nl[i]] = (th_b + ub) / 2; }
    }
    int nnew = Lp - L;
#pragma unroll 1
    for (int i = 0; i < nnew; ++i) {
      double ua = dr.normal(1.0, sd_a), ub = dr.normal(0.0, sd_b);
      int k = L + i;
      if (k < Lp) { na[nl[k]] = ua; nb[nl[k]] = ub; }
      if (k < nnew) logh += norm_pdf(ua, 1.0, new_sa2) + norm_pdf(ub, 0.0, new_sb2);   // pdf, range(L, nn) (Q9)
    }
    hratio = exp(loghstar - logh);
    detjacob = exp2(-(double)(2 * L));
  } else {                                                                     // funcs.py:1113-1138
    new_sa2 = dr.invgamma(1); new_sb2 = dr.invgamma(1);                        // redrawn at :1127-1128
    const double sd_a = sqrt(new_sa2), sd_b = sqrt(new_sb2);
#pragma unroll 1
    for (int i = 0; i < Lp; ++i) { na[nl[i]] = dr.normal(1.0, sd_a); nb[nl[i]] = dr.normal(0.0, sd_b); }
  }

  // ---- prior terms entering log_strucratio (funcs.py:1241-1245, 1265-1269, 1287-1289) ----
  double ll_o, lp_o, ll_n, lp_n;
  if (live_fs != nullptr) { ll_o = live_fs[0]; lp_o = live_fs[1]; }      // the live tree's, computed once per tree (same call, same bits)
  else fstruc_tree(pt, tk, m, oa, ob, sa, sb, ll_o, lp_o);
  fstruc_tree(pt, nt, mp, na, nb, new_sa2, new_sb2, ll_n, lp_n);
  info.change = change;
  info.Q = Q; info.Qinv = Qinv; info.hratio = hratio; info.detjacob = detjacob;
  info.new_sigma = new_sigma; info.new_sa2 = new_sa2; info.new_sb2 = new_sb2;
  info.fs_old = (change != CH_NONE) ? (ll_o + lp_o) : ll_o;
  info.ll_new = ll_n; info.lp_new = lp_n;
  info.m_new = mp;
  info.ndraws = dr.ndraws;
  if (MODE == 1 && dr.desync) info.flags |= PF_TAPE_DESYNC;
  *nn_out = mp;
}

// Prior initialisation of one tree (bsr_class.py:128-142): sigma_a, sigma_b ~ IG(1), then grow() from the root.
template <int MODE>
__device__ void init_tree(const PriorTables& pt, Draws<MODE>& dr, uint32_t* tok, double* pa, double* pb, int* nn_out,
                          double& sa, double& sb) {
  uint32_t nt[BSR_MAXN];
#pragma unroll 1
  for (int attempt = 0; attempt < 16; ++attempt) {
    sa = dr.invgamma(1);
    sb = dr.invgamma(1);
    // grow() draws the lt parameters of the initial tree itself (funcs.py:104-107); replicate the draws in order
    uint8_t pend[BSR_MAXN + 2];
    int sp = 0, pos = 0;
    bool overflow = false;
    pend[sp++] = 0;
    const double sd_a = sqrt(sa), sd_b = sqrt(sb);
    while (sp > 0) {
      int d = pend[--sp];
      if (pos >= BSR_MAXN) { overflow = true; break; }
      bool terminal = false;
      int oi = 0;
      if (d > 0) {
        double test = dr.uniform();
        if (test > pt.psplit[d]) { (void)dr.randint(0, pt.n_feature); terminal = true; }
        else oi = dr.choice(pt);
      } else oi = dr.choice(pt);
      if (terminal) {
        nt[pos] = make_tok(OP_LEAF, 0, dr.randint(0, pt.n_feature));
        pa[pos] = 0.0; pb[pos] = 0.0; ++pos;
      } else {
        int o = pt.ops[oi];
        double a = 0.0, b = 0.0;
        if (o == OP_LT) { a = dr.normal(1.0, sd_a); b = dr.normal(0.0, sd_b); }
        nt[pos] = make_tok(o, oi, 0);
        pa[pos] = a; pb[pos] = b; ++pos;
        pend[sp++] = (uint8_t)(d + 1);
        if (o >= OP_ADD) pend[sp++] = (uint8_t)(d + 1);
      }
    }
    if (!overflow) {
#pragma unroll 1
      for (int j = 0; j < pos; ++j) tok[j] = nt[j];
      *nn_out = pos;
      return;
    }
  }
  // (practically unreachable) fall back to the smallest legal tree: neg(x0)
  tok[0] = make_tok(pt.ops[0], 0, 0); tok[1] = make_tok(OP_LEAF, 0, 0);
  pa[0] = pb[0] = pa[1] = pb[1] = 0.0;
  if (pt.ops[0] == OP_LT) { pa[0] = 1.0; }
  if (pt.ops[0] >= OP_ADD) { tok[2] = make_tok(OP_LEAF, 0, 0); pa[2] = pb[2] = 0.0; *nn_out = 3; }
  else *nn_out = 2;
}
