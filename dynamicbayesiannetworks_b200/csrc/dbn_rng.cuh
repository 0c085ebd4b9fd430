// Philox4x32-10 counter-based generator and the FP64 distributions the chain step draws from.
//
// Replaces the reference's two host generators — Python `random.randint` (moves.py:36,150,434,544,651) and
// NumPy's legacy global RandomState (`np.random.gamma/choice/uniform/multivariate_normal`, samplers.py,
// utils.py, changepointMoves.py) — with one stateless stream per (chain, node, iteration):
// counter = (global chain id, node, iteration, draw index), key = seed.  Bit-exact parity with the reference's
// streams is only defined in replay mode; this generator is the free-running mode.
#pragma once
#include <cstdint>

namespace dbn {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

struct Rng {
  uint2 key;
  uint32_t chain, node, it, ctr;

  __device__ __forceinline__ void init(uint64_t seed, uint32_t chain_, uint32_t node_) {
    key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    chain = chain_;
    node = node_;
    it = 0;
    ctr = 0;
  }
  __device__ __forceinline__ void start_iteration(uint32_t it_) {
    it = it_;
    ctr = 0;
  }
  __device__ __forceinline__ uint4 block() { return philox4x32_10(make_uint4(chain, node, it, ctr++), key); }

  // two uniforms in [0,1) with 53 random bits each
  __device__ __forceinline__ void uniform2(double& a, double& b) {
    uint4 r = block();
    a = (double)((((uint64_t)r.x << 32) | r.y) >> 11) * 0x1.0p-53;
    b = (double)((((uint64_t)r.z << 32) | r.w) >> 11) * 0x1.0p-53;
  }
  __device__ __forceinline__ double uniform() {
    double a, b;
    uniform2(a, b);
    return a;
  }
  // index in [0, n): np.random.choice over n candidates / random.randint(0, n-1)
  __device__ __forceinline__ int below(int n) { return (int)__umulhi(block().x, (uint32_t)n); }

  // two independent N(0,1) (Box-Muller)
  __device__ __forceinline__ void normal2(double& z0, double& z1) {
    double u, v;
    uniform2(u, v);
    double r = sqrt(-2.0 * log(1.0 - u));  // 1-u in (0,1]
    double s, c;
    sincospi(2.0 * v, &s, &c);
    z0 = r * c;
    z1 = r * s;
  }

  // Gamma(shape, scale=1) by Marsaglia-Tsang; shape < 1 handled by the u^(1/shape) boost.
  // ONE exit: lanes of a warp leave the rejection loop after different numbers of tries; with a `return` inside the loop
  // each of them would return to the caller on its own and the warp would stay split there (ptxas treats a call as
  // convergent and emits no reconvergence point after it) -- see the note at gamma_draw_fast.
  __device__ double gamma(double shape) {
    double boost = 1.0;
    if (shape < 1.0) {
      boost = pow(1.0 - uniform(), 1.0 / shape);
      shape += 1.0;
    }
    const double d = shape - 1.0 / 3.0;
    const double c = rsqrt(9.0 * d);
    double out = d;  // unreachable in practice (acceptance > 0.95 per try)
    bool done = false;
    for (int tries = 0; tries < 64 && !done; ++tries) {
      double x, unused;
      normal2(x, unused);
      double v = 1.0 + c * x;
      const double u = 1.0 - uniform();
      if (v > 0.0) {
        v = v * v * v;
        if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) {
          out = d * v;
          done = true;
        }
      }
    }
    return boost * out;
  }
};

// ---- out-of-line variants for the h / nh kernel (dbn_chain_fast.cuh) ------------------------------------------------
// The chain step is one long program per thread; inlining Philox and the FP64 log / exp / sincospi expansions at
// every use made it ~31k SASS instructions and instruction-fetch bound (profiles/r1_v1).  These are called instead.
// They are stateless: every consumer owns a fixed range of the per-iteration counter word.
__device__ __noinline__ inline uint4 philox_block(uint32_t chain, uint32_t node, uint32_t it, uint32_t ctr, uint2 key) {
  return philox4x32_10(make_uint4(chain, node, it, ctr), key);
}
__device__ __noinline__ inline double dlog(double x) { return log(x); }
__device__ __noinline__ inline double dexp(double x) { return exp(x); }

// two independent N(0,1) from one Philox block (Box-Muller on two 53-bit uniforms)
__device__ __noinline__ inline double2 normal_pair(uint32_t chain, uint32_t node, uint32_t it, uint32_t ctr, uint2 key) {
  const uint4 r = philox4x32_10(make_uint4(chain, node, it, ctr), key);
  const double u = (double)((((uint64_t)r.x << 32) | r.y) >> 11) * 0x1.0p-53;
  const double v = (double)((((uint64_t)r.z << 32) | r.w) >> 11) * 0x1.0p-53;
  const double rad = sqrt(-2.0 * log(1.0 - u));
  double s, c;
  sincospi(2.0 * v, &s, &c);
  return make_double2(rad * c, rad * s);
}

// Gamma(shape, 1) by Marsaglia-Tsang on the counter range [ctr0, ctr0 + 192); one exit (see gamma_draw_fast)
__device__ __noinline__ inline double gamma_draw(uint32_t chain, uint32_t node, uint32_t it, uint32_t ctr0, uint2 key, double shape) {
  double boost = 1.0;
  uint32_t ctr = ctr0;
  if (shape < 1.0) {
    const uint4 r = philox4x32_10(make_uint4(chain, node, it, ctr++), key);
    const double u = (double)((((uint64_t)r.x << 32) | r.y) >> 11) * 0x1.0p-53;
    boost = pow(1.0 - u, 1.0 / shape);
    shape += 1.0;
  }
  const double d = shape - 1.0 / 3.0;
  const double c = rsqrt(9.0 * d);
  double out = d;  // unreachable in practice (acceptance > 0.95 per try)
  bool done = false;
  for (int tries = 0; tries < 64 && !done; ++tries) {
    const double2 z = normal_pair(chain, node, it, ctr++, key);
    const uint4 r = philox4x32_10(make_uint4(chain, node, it, ctr++), key);
    double v = 1.0 + c * z.x;
    if (v > 0.0) {
      v = v * v * v;
      const double u = 1.0 - (double)((((uint64_t)r.x << 32) | r.y) >> 11) * 0x1.0p-53;
      const double x2 = z.x * z.x;
      // Marsaglia-Tsang squeeze (inside the exact test's region), then the exact test
      if (u < 1.0 - 0.0331 * x2 * x2 || log(u) < 0.5 * x2 + d - d * v + d * log(v)) {
        out = d * v;
        done = true;
      }
    }
  }
  return boost * out;
}

// ---- fast free-running normals (h / nh kernel) -------------------------------------------------------------------
// Box-Muller with the transcendental part in single precision on the SFU (MUFU.LG2 / MUFU.SIN / MUFU.COS / MUFU.SQRT):
// FOUR N(0,1) variates per Philox block (two radius / angle pairs from four 32-bit words) instead of two from the FP64
// log / sincospi expansions, which were 17.6 % of the kernel's executed instructions (profiles/r1_v10).  The variates
// carry the resolution of their 32-bit uniforms: |z| <= 6.76 (radius word 0 -> u = 2^-33), absolute error of a variate
// below 1e-6 (MUFU accuracy ~2^-21.4 on sin / cos over [-pi, pi], 2^-22 on log2), far below the Monte-Carlo error of
// any chain length this library runs; the model arithmetic itself stays FP64.  Only the free-running mode draws them:
// replay mode injects the reference's draws, so every replay parity test is unaffected; the distribution is pinned by
// tests/test_gpu_pinning.py::test_free_running_gibbs_draw_moments and the free-running edge-posterior tests.
__device__ __forceinline__ float2 box_muller_f32(uint32_t a, uint32_t b) {
  const float u = fmaf(__uint2float_rn(a), 0x1.0p-32f, 0x1.0p-33f);       // (0, 1]
  const float ang = __int2float_rn((int)b) * (6.283185307179586f * 0x1.0p-32f);  // [-pi, pi]
  const float rad = sqrtf(fmaxf(0.0f, -1.3862943611198906f * __log2f(u)));  // sqrt(-2 ln u)
  float s, c;
  __sincosf(ang, &s, &c);
  return make_float2(rad * c, rad * s);
}
__device__ __noinline__ inline float4 normal4_fast(uint32_t chain, uint32_t node, uint32_t it, uint32_t ctr, uint2 key) {
  const uint4 r = philox4x32_10(make_uint4(chain, node, it, ctr), key);
  const float2 p = box_muller_f32(r.x, r.y), q = box_muller_f32(r.z, r.w);
  return make_float4(p.x, p.y, q.x, q.y);
}

// Gamma(shape, 1) by Marsaglia-Tsang, ONE Philox block per try: words 0, 1 -> the normal, words 2, 3 -> the 53-bit uniform.
// ONE exit.  Lanes leave the rejection loop after different numbers of tries.  With `return` statements inside the loop
// (round 1) every lane returned to the call site on its own, and because ptxas treats a call as convergent there is no
// reconvergence point after it: the warp stayed split until the next WARPSYNC, ran `__syncthreads()` in two parts, each
// counted as a whole warp, and the block's barrier phases slipped (the h-dbn instantiations have no WARPSYNC.ALL before
// their barriers at all).  Symptoms: garbage work counters of the h-dbn kernel (profiles/r1_v13/methods.jsonl), and the
// lost cooperative updates of profiles/r1_notes.md.  With a single exit the loop's own reconvergence point rejoins the
// lanes inside the function.
__device__ __noinline__ inline double gamma_draw_fast(uint32_t chain, uint32_t node, uint32_t it, uint32_t ctr0, uint2 key, double shape) {
  double boost = 1.0;
  uint32_t ctr = ctr0;
  if (shape < 1.0) {
    const uint4 r = philox4x32_10(make_uint4(chain, node, it, ctr++), key);
    const double u = (double)((((uint64_t)r.x << 32) | r.y) >> 11) * 0x1.0p-53;
    boost = pow(1.0 - u, 1.0 / shape);
    shape += 1.0;
  }
  const double d = shape - 1.0 / 3.0;
  const double c = rsqrt(9.0 * d);
#ifdef DBN_GAMMA_EARLY   // experiment: the round-1 shape (lanes return from inside the loop)
  for (int tries = 0; tries < 64; ++tries) {
    const uint4 r = philox4x32_10(make_uint4(chain, node, it, ctr++), key);
    const double x = (double)box_muller_f32(r.x, r.y).x;
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    const double u = 1.0 - (double)((((uint64_t)r.z << 32) | r.w) >> 11) * 0x1.0p-53;
    const double x2 = x * x;
    if (u < 1.0 - 0.0331 * x2 * x2) return boost * d * v;
    if (log(u) < 0.5 * x2 + d - d * v + d * log(v)) return boost * d * v;
  }
  return boost * d;
#else
  double out = d;  // unreachable in practice (acceptance > 0.95 per try)
  bool done = false;
  for (int tries = 0; tries < 64 && !done; ++tries) {
    const uint4 r = philox4x32_10(make_uint4(chain, node, it, ctr++), key);
    const double x = (double)box_muller_f32(r.x, r.y).x;
    double v = 1.0 + c * x;
    if (v > 0.0) {
      v = v * v * v;
      const double u = 1.0 - (double)((((uint64_t)r.z << 32) | r.w) >> 11) * 0x1.0p-53;
      const double x2 = x * x;
      // Marsaglia-Tsang squeeze (inside the exact test's region), then the exact test
      if (u < 1.0 - 0.0331 * x2 * x2 || log(u) < 0.5 * x2 + d - d * v + d * log(v)) {
        out = d * v;
        done = true;
      }
    }
  }
  return boost * out;
#endif
}

}  // namespace dbn
