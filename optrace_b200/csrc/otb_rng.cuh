// otb_rng.cuh — counter-based random numbers for on-device ray generation and HURB.
// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11): key = seed,
// counter = (global ray id, stream id) — results do not depend on grid shape or GPU count.
// The reference uses numpy's SFC64 generator plus global shuffles (random.py:5, 41-45), so generated
// bundles agree statistically, not bit-wise.
#pragma once
#include <stdint.h>

struct Philox4 { uint32_t v[4]; };

__host__ __device__ __forceinline__ void philox_round(uint32_t* c, uint32_t k0, uint32_t k1)
{
    const uint64_t p0 = (uint64_t)0xD2511F53u*c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u*c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint64_t counter, uint32_t stream, uint32_t sub, uint64_t key)
{
    uint32_t c[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), stream, sub};
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 r;
    r.v[0] = c[0]; r.v[1] = c[1]; r.v[2] = c[2]; r.v[3] = c[3];
    return r;
}

// uniform double in [0, 1) with 53 random bits
__host__ __device__ __forceinline__ double u01(uint32_t hi, uint32_t lo)
{
    uint64_t b = ((uint64_t)hi << 32) | lo;
    return (double)(b >> 11)*(1.0/9007199254740992.0);
}

// two standard normal deviates (Box–Muller) from one Philox block
__device__ __forceinline__ void normal2(const Philox4& r, double& z0, double& z1)
{
    double u = 1.0 - u01(r.v[0], r.v[1]);       // (0, 1]
    double v = u01(r.v[2], r.v[3]);
    double rad = sqrt(-2.0*log(u));
    double sn, cs;
    sincospi(2.0*v, &sn, &cs);
    z0 = rad*cs;
    z1 = rad*sn;
}

// Keyed bijection of [0, n): 4-round Feistel network on the smallest even-bit domain >= n with cycle
// walking.  Stands in for the global shuffle that follows the reference's stratified grids
// (random.py:41-45, 62-66): ray k gets stratum perm(k), every stratum is used exactly once.
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__host__ __device__ __forceinline__ int bits_for(uint64_t n)      // smallest b with 2^b >= n (n >= 2)
{
#ifdef __CUDA_ARCH__
    return 64 - __clzll((long long)(n - 1));
#else
    int bits = 0;
    while (((uint64_t)1 << bits) < n) ++bits;
    return bits;
#endif
}

// Keyed bijection of [0, n).  For n < 2^31: generalised Feistel network on Z_a x Z_b with a = ceil(sqrt(n)),
// b = ceil(n / a) (Black & Rogaway, "Ciphers with arbitrary finite domains", CT-RSA 2002, method fe[r, a, b]):
// the domain a*b exceeds n by less than a, so the cycle walk that brings the result back into [0, n) almost
// never iterates.  (A power-of-two domain rejects up to half of the values, and a warp pays the MAXIMUM
// iteration count of its 32 lanes: 6-7 walks of 4 rounds per permutation were 40 % of the generator.)
// Larger n: balanced power-of-two Feistel with cycle walking.
// moduli of the generalised Feistel network for n < 2^31: a = ceil(sqrt(n)), b = ceil(n / a).  They depend on n
// only; the generator computes them once per source, not once per ray and random variable.
__host__ __device__ inline void feistel_setup(uint64_t n, uint32_t& a, uint32_t& b)
{
    a = b = 0;
    if (n <= 1 || n >= 0x80000000ull) return;
    const uint32_t n32 = (uint32_t)n;
    a = (uint32_t)sqrtf((float)n32);
    while ((uint64_t)a*a < n32) ++a;                 // float rounding repaired
    while (a > 1 && (uint64_t)(a - 1)*(a - 1) >= n32) --a;
    b = (n32 + a - 1)/a;
}

__host__ __device__ inline uint64_t feistel_perm_ab(uint64_t i, uint64_t n, uint32_t a, uint32_t b, uint64_t key)
{
    if (n <= 1) return 0;
    const uint32_t ka = (uint32_t)key ^ (uint32_t)(key >> 32), kb = (uint32_t)(key >> 16) ^ (uint32_t)(key >> 32);
    if (n < 0x80000000ull) {
        const uint32_t n32 = (uint32_t)n;
        uint32_t x = (uint32_t)i;
        do {
            uint32_t R = x/a, L = x - R*a;               // L in Z_a, R in Z_b
#pragma unroll
            for (int rd = 0; rd < 4; ++rd) {
                const uint32_t m = (rd & 1) ? b : a;     // rounds alternate between the two moduli
                const uint32_t h = mix32(R ^ ((rd & 1) ? kb : ka) ^ (0x9E3779B9u*(rd + 1)));
                uint32_t t = L + (uint32_t)(((uint64_t)h*m) >> 32);
                if (t >= m) t -= m;
                L = R;
                R = t;
            }
            x = a*R + L;
        } while (x >= n32);
        return x;
    }
    int bits = bits_for(n);
    bits += bits & 1;                       // even number of bits
    const int half = bits >> 1;
    const uint32_t mask = (half >= 32) ? 0xffffffffu : (((uint32_t)1 << half) - 1);
    uint64_t x = i;
    do {
        uint32_t l = (uint32_t)(x >> half) & mask, r = (uint32_t)x & mask;
#pragma unroll
        for (int rd = 0; rd < 4; ++rd) {
            uint32_t f = mix32(r ^ ((rd & 1) ? kb : ka) ^ (0x9E3779B9u*(rd + 1))) & mask;
            uint32_t nl = r;
            r = l ^ f;
            l = nl;
        }
        x = ((uint64_t)l << half) | r;
    } while (x >= n);
    return x;
}

__host__ __device__ inline uint64_t feistel_perm(uint64_t i, uint64_t n, uint64_t key)
{
    uint32_t a, b;
    feistel_setup(n, a, b);
    return feistel_perm_ab(i, n, a, b, key);
}
