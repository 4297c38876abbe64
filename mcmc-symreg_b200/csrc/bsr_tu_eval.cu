// Evaluation stage launch: k_eval<K> (fp32 tree evaluation + fp64 Gram; fp64 re-evaluation inside the kernel).
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include "bsr_handle.h"
#include "bsr_kernels.cuh"

// Scratch for row-split partial records, grown on demand (never while kernels that use it may be in flight: the
// size only depends on the launch geometry, which is fixed between bsr_set_data / bsr_set_launch_geometry calls).
static int ensure_part(bsr_handle* h, size_t doubles) {
  if (doubles <= h->part_cap) return 0;
  CK(cudaDeviceSynchronize());
  if (h->part) cudaFree(h->part);
  h->part = nullptr; h->part_cap = 0;
  CK(cudaMalloc((void**)&h->part, doubles * sizeof(double)));
  h->part_cap = doubles;
  return 0;
}

template <int PASS, int CM, bool LOADALL = false>
static int launch_pass(bsr_handle* h, cudaStream_t s, EvalCtx ec, int threads, int tpc, int n_splits) {
  const int K = h->cfg.K, P = 2 * K;
  ec.tpc = tpc;
  const int groups = threads / tpc;
  if (tpc <= 32) n_splits = 1;
  ec.n_splits = n_splits;
  if (n_splits > 1) {
    if (ensure_part(h, (size_t)h->cfg.n_chains * n_splits * (gram_n_sum(P) + P))) return 1;
    ec.part = h->part; ec.split_cnt = h->split_cnt;
  }
  const dim3 blocks((ec.cn + groups - 1) / groups, n_splits);
  const size_t smem = eval_smem_bytes(P, threads, tpc);
#define LAUNCH_K(KT)                                                                                        \
  do {                                                                                                      \
    CK(cudaFuncSetAttribute(k_eval<KT, PASS, CM, LOADALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    k_eval<KT, PASS, CM, LOADALL><<<blocks, threads, smem, s>>>(h->st, ec);                                              \
  } while (0)
  switch (K) {
    case 1: LAUNCH_K(1); break;
    case 2: LAUNCH_K(2); break;
    case 3: LAUNCH_K(3); break;
    case 4: LAUNCH_K(4); break;
    case 5: LAUNCH_K(5); break;
    default: LAUNCH_K(0); break;
  }
#undef LAUNCH_K
  CK(cudaGetLastError());
  return 0;
}

int bsr_launch_eval(bsr_handle* h, cudaStream_t s, int init_only, int c0, int cn) {
  const int K = h->cfg.K, P = 2 * K, C = h->cfg.n_chains;
  EvalCtx ec;
  ec.X32 = h->X32; ec.y32 = h->y32; ec.X64 = h->X64; ec.y64 = h->y64;
  ec.n = (uint32_t)h->n; ec.ld = (uint32_t)h->ld;
  ec.sums = h->gram; ec.maxs = h->gram + (size_t)C * gram_n_sum(P);
  ec.precision = h->cfg.precision; ec.init_only = init_only;
  ec.c0 = c0; ec.cn = cn;
  ec.fill_cache = init_only;
  ec.need64 = h->need64;
  // fp32 pass: a warp per chain (no block-level synchronisation, 4 chains per 128-thread block) unless a chain has so
  // many rows that a whole block should share them; fp64 pass: a fat block per chain (few chains, latency matters).
  int threads = h->threads_eval;
  int tpc = ((int64_t)h->n <= 8192 || (int64_t)cn * 32 >= 148 * 2048) ? 32 : threads;
  if (K > 5) tpc = threads;                          // generic path keeps its accumulators in local memory
  if (const char* e = getenv("BSR_EVAL_THREADS")) threads = atoi(e);
  if (const char* e = getenv("BSR_EVAL_TPC")) tpc = atoi(e);
  if (tpc > threads) tpc = threads;
  ec.part = nullptr; ec.split_cnt = nullptr;
  // row splits: when a launch would have too few blocks to fill 148 SMs (few chains x many rows), or for the fp64
  // pass (few flagged chains, each slow), several blocks share one chain's rows
  auto splits_for = [&](int vec_rows, int blk_threads, int chains_expected) {
    const int n_vec = (int)((h->n + vec_rows - 1) / vec_rows);
    const int max_s = std::max(1, n_vec / (2 * blk_threads));           // keep >= 2 passes per block
    const int want = (148 * 8 + chains_expected - 1) / std::max(1, chains_expected);
    return std::max(1, std::min(std::min(max_s, want), 64));
  };
  if (h->cfg.precision == 0) {
    const bool cache = K <= 5 && h->st.col[0] != nullptr;
    const int s32 = splits_for(4, threads, cn);
    int rc;
    const bool split_tg = cache && !getenv("BSR_FUSED_EVAL");   // trees kernel + Gram kernel instead of one fused kernel
    if (split_tg) {
      const int per_chain = ec.fill_cache ? 2 * K : K;
      const int items = cn * per_chain;
      k_trees<<<(items + 3) / 4, 128, 0, s>>>(h->st, ec);
      CK(cudaGetLastError());
      if (h->prof_inner) cudaEventRecord(h->ev[4], s);
      if (ec.fill_cache) rc = launch_pass<0, CM_FILL, true>(h, s, ec, threads, tpc, s32);
      else rc = launch_pass<0, CM_CACHED, true>(h, s, ec, threads, tpc, s32);
      if (h->prof_inner) cudaEventRecord(h->ev[5], s);
    } else if (!cache) rc = launch_pass<0, CM_PLAIN>(h, s, ec, threads, tpc, s32);
    else if (ec.fill_cache) rc = launch_pass<0, CM_FILL>(h, s, ec, threads, tpc, s32);
    else rc = launch_pass<0, CM_CACHED>(h, s, ec, threads, tpc, s32);
    if (rc) return 1;
    // fp64 pass of the flagged chains: with a column cache only the out-of-range columns are re-interpreted in double
    if (cache) return launch_pass<1, CM_MIXED>(h, s, ec, 256, 256, splits_for(2, 256, std::max(1, cn / 64)));
    return launch_pass<1, CM_PLAIN>(h, s, ec, 256, 256, splits_for(2, 256, std::max(1, cn / 64)));
  }
  const int t64 = K > 5 ? threads : tpc;
  return launch_pass<1, CM_PLAIN>(h, s, ec, threads, t64, splits_for(2, threads, cn));
}
