// Per-chain counter-based RNG (Philox4x32-10) and the value-level draw interface.
//
// The five draw kinds are exactly the calls the reference makes (SURVEY.md 8a "RNG tape"):
//   uniform()          np.random.uniform(0,1,1)                 codes/funcs.py:81,483,604,623,1299
//   randint(lo,hi)     np.random.randint(lo,hi)                 codes/funcs.py:83,99,508,544,558,585,687,794,912,914
//   choice()           np.random.choice(len(Ops), p=Op_weights) codes/funcs.py:86,92,689,803
//   normal(loc,scale)  scipy norm.rvs(loc,scale)                codes/funcs.py:105-106,499-500,974-975,...
//   invgamma(a)        scipy invgamma.rvs(a), a in {1,4}        codes/funcs.py:945-946,1195; bsr_class.py:123,131-132
// MODE 0: Philox, 1: replay a recorded tape (one double per call), 2: Philox + record what was drawn.
// Philox is keyed by (seed) and countered by (global chain id, proposal index, draw block, purpose), so a
// chain's stream does not depend on launch geometry or on how chains are sharded over GPUs.
#pragma once
#include "bsr_common.cuh"

struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ void round(uint32_t (&c)[4], uint32_t ka, uint32_t kb) const {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ ka, n1 = lo1, n2 = hi0 ^ c[3] ^ kb, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __device__ __forceinline__ void gen(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) const {
    uint32_t c[4] = {c0, c1, c2, c3};
    uint32_t ka = k0, kb = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round(c, ka, kb);
      ka += 0x9E3779B9u; kb += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
};

template <int MODE>
struct Draws {
  // Philox side
  Philox ph;
  uint32_t chain_lo, chain_hi_purpose, step, block;
  uint32_t buf[4];
  int have;   // number of unused 64-bit words in buf (0..2)
  // tape side
  const double* tape;  // MODE 1: values to consume;
  double* rec;         // MODE 2: where to record
  int pos, end;        // MODE 1: [pos,end) ; MODE 2: pos counts, end = capacity
  int ndraws;
  bool desync;

  __device__ void init_philox(uint64_t seed, uint64_t chain, uint32_t step_, uint32_t purpose) {
    ph.k0 = (uint32_t)seed; ph.k1 = (uint32_t)(seed >> 32);
    chain_lo = (uint32_t)chain; chain_hi_purpose = ((uint32_t)(chain >> 32) << 8) | (purpose & 0xffu);
    step = step_; block = 0; have = 0; ndraws = 0; desync = false;
    tape = nullptr; rec = nullptr; pos = 0; end = 0;
  }
  __device__ void init_tape(const double* t, int lo, int hi) { tape = t; pos = lo; end = hi; }
  __device__ void init_record(double* r, int capacity) { rec = r; pos = 0; end = capacity; }

  // u01 / u01_open0 (the only callers) are out of line on purpose (like normal / invgamma below): the proposal kernels call the generator from dozens of
  // divergent places, and with Philox and the fp64 math inlined at each of them the kernel is 250 KB of code that
  // thrashes the instruction cache (ncu: 9.5 no-instruction stall cycles per issued instruction)
  __device__ __forceinline__ uint64_t next_u64() {
    if (have == 0) { ph.gen(chain_lo, step, block++, chain_hi_purpose, buf); have = 2; }
    --have;
    return ((uint64_t)buf[2 * have + 1] << 32) | buf[2 * have];
  }
  // [0,1) with 53 bits
  __device__ __noinline__ double u01() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
  // (0,1]
  __device__ __noinline__ double u01_open0() { return (double)((next_u64() >> 11) + 1ull) * (1.0 / 9007199254740992.0); }

  __device__ __forceinline__ double tape_next() {
    ++ndraws;
    if (pos >= end) { desync = true; return 0.5; }
    return tape[pos++];
  }
  __device__ __forceinline__ double record(double v) {
    ++ndraws;
    if (MODE == 2 && rec != nullptr) { if (pos < end) rec[pos] = v; ++pos; }
    return v;
  }

  __device__ double uniform() {
    if (MODE == 1) return tape_next();
    return record(u01());
  }
  __device__ __noinline__ int randint(int lo, int hi) {
    if (MODE == 1) {
      int v = (int)tape_next();
      if (v < lo || v >= hi) { desync = true; v = lo; }
      return v;
    }
    int v = lo + (int)(u01() * (double)(hi - lo));
    if (v >= hi) v = hi - 1;
    return (int)record((double)v);
  }
  __device__ __noinline__ int choice(const PriorTables& pt) {
    if (MODE == 1) {
      int v = (int)tape_next();
      if (v < 0 || v >= pt.n_ops) { desync = true; v = 0; }
      return v;
    }
    double u = u01();
    int v = 0;
    while (v < pt.n_ops - 1 && !(pt.cdf[v] > u)) ++v;   // searchsorted(cdf, u, side='right')
    return (int)record((double)v);
  }
  __device__ __noinline__ double normal(double loc, double scale) {
    if (MODE == 1) return tape_next();
    double u1 = u01_open0(), u2 = u01();
    double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    return record(loc + scale * z);
  }
  // invgamma(1) = 1/Exp(1); invgamma(4) = 1/Gamma(4,1) with Gamma(4) = -log(u1 u2 u3 u4)
  __device__ __noinline__ double invgamma(int shape) {
    if (MODE == 1) return tape_next();
    double g = 0.0;
    if (shape == 1) {
      g = -log(u01_open0());
    } else {
      double p = 1.0;
      for (int i = 0; i < shape; ++i) p *= u01_open0();
      g = -log(p);
    }
    return record(1.0 / g);
  }
};
