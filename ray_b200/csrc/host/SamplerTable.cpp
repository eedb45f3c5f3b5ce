// SamplerTable.cpp -- the built-in 32 x 4096 x 2 sample table of the standalone host layer.
//
// Role in the reference: precomputed/__pmj02_samples.inl (a baked PMJ02 table, internal/Core.h:363-368) that
// Cpu::Renderer passes as `rand_seq`.  That table is data of the reference and is NOT copied here: inside the reference
// tree Cuda::Renderer uploads `__pmj02_samples` itself, and the parity tests upload it through
// Renderer::SetSamplerTable / rc_upload_tables.  Standalone, every dimension pair is an Owen-scrambled, index-shuffled
// copy of the base-2 Sobol' (0,2)-sequence -- the same stratification class as PMJ02 (every power-of-two prefix is a
// (0,m,2)-net), which is all the integrator's lookup (Owen-scrambling by dimension/pixel on top) relies on.
#include <cstdint>
#include <vector>

#include "../rt_types.h"
#include "RendererCuda.h"

namespace RayB200 {
namespace Cuda {

namespace {
uint32_t reverse_bits(uint32_t x) {
    x = ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
    x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2);
    x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8);
    return (x >> 16) | (x << 16);
}
uint32_t hash32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
// hash-based Owen scrambling (Laine-Karras style permutation applied in bit-reversed space)
uint32_t owen(uint32_t x, uint32_t seed) {
    x = reverse_bits(x);
    x += seed;
    x ^= x * 0x6c50b47cu;
    x ^= x * 0xb82f1e52u;
    x ^= x * 0xc7afe638u;
    x ^= x * 0x8d22f6e6u;
    return reverse_bits(x);
}
uint32_t sobol_dim2(uint32_t n) { // second dimension of the Sobol' sequence: direction numbers v_k = v_{k-1} ^ (v_{k-1} >> 1)
    uint32_t v = 1u << 31, r = 0;
    for (; n; n >>= 1) {
        if (n & 1u) {
            r ^= v;
        }
        v ^= v >> 1;
    }
    return r;
}
} // namespace

std::vector<uint32_t> GenerateSamplerTable() {
    const int dims = rt::kRandDims, samples = rt::kRandSamples;
    std::vector<uint32_t> t(size_t(dims) * samples * 2);
    for (int d = 0; d < dims; ++d) {
        const uint32_t s_idx = hash32(0x9e3779b9u * uint32_t(d + 1)), s_x = hash32(s_idx ^ 0x68bc21ebu),
                       s_y = hash32(s_idx ^ 0x02e5be93u);
        for (int i = 0; i < samples; ++i) {
            // shuffle the index with an Owen scramble too: keeps every power-of-two block a permutation of itself
            const uint32_t j = owen(uint32_t(i), s_idx) & uint32_t(samples - 1);
            const uint32_t x = reverse_bits(j); // van der Corput = first Sobol' dimension
            const uint32_t y = sobol_dim2(j);
            t[(size_t(d) * samples + i) * 2 + 0] = owen(x, s_x);
            t[(size_t(d) * samples + i) * 2 + 1] = owen(y, s_y);
        }
    }
    return t;
}

} // namespace Cuda
} // namespace RayB200
