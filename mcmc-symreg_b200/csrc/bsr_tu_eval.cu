// Evaluation stage launch: k_eval<K> (fp32 tree evaluation + fp64 Gram; fp64 re-evaluation inside the kernel).
#include <cstdio>
#include <cstdlib>
#include "bsr_handle.h"
#include "bsr_kernels.cuh"

int bsr_launch_eval(bsr_handle* h, cudaStream_t s, int init_only, int c0, int cn) {
  const int K = h->cfg.K, P = 2 * K, C = h->cfg.n_chains;
  EvalCtx ec;
  ec.X32 = h->X32; ec.y32 = h->y32; ec.X64 = h->X64; ec.y64 = h->y64;
  ec.n = (uint32_t)h->n; ec.ld = (uint32_t)h->ld;
  ec.sums = h->gram; ec.maxs = h->gram + (size_t)C * gram_n_sum(P);
  ec.precision = h->cfg.precision; ec.init_only = init_only;
  ec.c0 = c0; ec.cn = cn;
  // a block per chain: the fp64 re-evaluation of an out-of-range chain is then shared by the whole block instead of
  // stalling one warp; only tiny row counts fall back to a warp per chain (measured on C2, profiles/README.md)
  int threads = h->threads_eval;
  ec.tpc = (h->n <= 256) ? 32 : threads;
  if (const char* e = getenv("BSR_EVAL_THREADS")) threads = atoi(e);
  if (const char* e = getenv("BSR_EVAL_TPC")) ec.tpc = atoi(e);
  if (ec.tpc > threads) ec.tpc = threads;
  const int groups = threads / ec.tpc;
  const int blocks = (cn + groups - 1) / groups;
  const size_t smem = eval_smem_bytes(P, threads, ec.tpc);
#define LAUNCH_K(KT)                                                                                    \
  do {                                                                                                  \
    CK(cudaFuncSetAttribute(k_eval<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    k_eval<KT><<<blocks, threads, smem, s>>>(h->st, ec);                                                \
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
