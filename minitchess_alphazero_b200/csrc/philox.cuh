// philox.cuh -- counter-based RNG for throughput mode (device only): Philox4x32-10, uniform / normal /
// gamma (Marsaglia-Tsang) draws.  Root Dirichlet noise = normalised gamma(alpha) draws, keyed by
// (seed, game, simulation counter, edge) so results do not depend on launch geometry.
#pragma once
#include <stdint.h>

namespace mcaz {

struct Philox {
    uint32_t key[2];
    uint32_t ctr[4];
    uint32_t out[4];
    int have;
    __device__ Philox(unsigned long long seed, uint32_t a, uint32_t b, uint32_t c) : have(0) {
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
        ctr[0] = 0; ctr[1] = a; ctr[2] = b; ctr[3] = c;
    }
    __device__ void round(uint32_t* c, const uint32_t* k) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    __device__ void refill() {
        uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k[2] = {key[0], key[1]};
#pragma unroll
        for (int r = 0; r < 10; ++r) { round(c, k); k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
        ctr[0] += 1;
        have = 4;
    }
    __device__ uint32_t next() { if (!have) refill(); return out[--have]; }
    __device__ double uniform() {  // (0,1)
        uint32_t a = next() >> 5, b = next() >> 6;
        return ((double)a * 67108864.0 + (double)b + 0.5) * (1.0 / 9007199254740992.0);
    }
    __device__ float uniformf() { return ((float)(next() >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0,1)
    __device__ float normalf() {
        const float u1 = uniformf(), u2 = uniformf();
        return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
    }
    // Marsaglia-Tsang in single precision (the noise only has to be a Dirichlet sample; nothing is compared
    // against numpy's stream here); shape < 1 via gamma(shape+1) * U^(1/shape).  The squeeze test accepts most
    // candidates without a logarithm.
    __device__ double gamma(double shape_d) {
        float shape = (float)shape_d, boost = 1.0f;
        if (shape < 1.0f) { boost = powf(uniformf(), 1.0f / shape); shape += 1.0f; }
        const float d = shape - 1.0f / 3.0f, c = rsqrtf(9.0f * d);
        for (int it = 0; it < 64; ++it) {
            const float x = normalf();
            float v = 1.0f + c * x;
            if (v <= 0.0f) continue;
            v = v * v * v;
            const float u = uniformf(), x2 = x * x;
            if (u < 1.0f - 0.0331f * x2 * x2 || logf(u) < 0.5f * x2 + d - d * v + d * logf(v)) {
                const float g = boost * d * v;
                return (double)(g > 1e-30f ? g : 1e-30f);     // never exactly zero: the draws are normalised by their sum
            }
        }
        return (double)(boost * d);
    }
};


}  // namespace mcaz
