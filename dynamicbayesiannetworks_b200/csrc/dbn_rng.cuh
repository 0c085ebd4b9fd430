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

  // Gamma(shape, scale=1) by Marsaglia-Tsang; shape < 1 handled by the u^(1/shape) boost
  __device__ double gamma(double shape) {
    double boost = 1.0;
    if (shape < 1.0) {
      boost = pow(1.0 - uniform(), 1.0 / shape);
      shape += 1.0;
    }
    const double d = shape - 1.0 / 3.0;
    const double c = rsqrt(9.0 * d);
    for (int tries = 0; tries < 64; ++tries) {
      double x, unused;
      normal2(x, unused);
      double v = 1.0 + c * x;
      if (v <= 0.0) continue;
      v = v * v * v;
      double u = 1.0 - uniform();
      if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return boost * d * v;
    }
    return boost * d;  // unreachable in practice (acceptance > 0.95 per try)
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

// Gamma(shape, 1) by Marsaglia-Tsang on the counter range [ctr0, ctr0 + 192)
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
  for (int tries = 0; tries < 64; ++tries) {
    const double2 z = normal_pair(chain, node, it, ctr++, key);
    const uint4 r = philox4x32_10(make_uint4(chain, node, it, ctr++), key);
    double v = 1.0 + c * z.x;
    if (v <= 0.0) continue;
    v = v * v * v;
    const double u = 1.0 - (double)((((uint64_t)r.x << 32) | r.y) >> 11) * 0x1.0p-53;
    const double x2 = z.x * z.x;
    if (u < 1.0 - 0.0331 * x2 * x2) return boost * d * v;  // Marsaglia-Tsang squeeze: inside the exact test's region
    if (log(u) < 0.5 * x2 + d - d * v + d * log(v)) return boost * d * v;
  }
  return boost * d;  // unreachable in practice (acceptance > 0.95 per try)
}

}  // namespace dbn
