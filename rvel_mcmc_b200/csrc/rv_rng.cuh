// rv_rng.cuh -- counter-based RNG for the device samplers: Philox-4x32-10 (Salmon et al. 2011).
// key = (seed_lo, seed_hi); counter = (id_lo, id_hi, step, stream).  Every draw of every walker at every
// step is a pure function of (seed, global walker id, step, stream), so chains do not depend on how the
// walkers are sharded over threads or GPUs.
#pragma once
#include <stdint.h>
#include "rv_core.cuh"

namespace rv {

struct U4 { uint32_t x, y, z, w; };

RV_HD uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }

RV_HD U4 philox4x32_10(uint64_t seed, uint64_t id, uint32_t step, uint32_t stream) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    U4 c = {(uint32_t)id, (uint32_t)(id >> 32), step, stream};
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        U4 n = {hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
        c = n;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

// 53-bit uniform in (0,1) from two words
RV_HD double u53(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}

// stream ids
enum : uint32_t { RNG_ACCEPT = 1u, RNG_SCHEDULE = 2u, RNG_STRETCH_Z = 0x10u, RNG_STRETCH_J = 0x20u, RNG_NORMAL = 0x100u };

// two standard normals (Box-Muller) for the pair index j of a walker at a step
RV_HD void normal_pair(uint64_t seed, uint64_t id, uint32_t step, uint32_t j, double& z0, double& z1) {
    const U4 r = philox4x32_10(seed, id, step, RNG_NORMAL + j);
    const double u1 = u53(r.x, r.y), u2 = u53(r.z, r.w);
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincos(2.0 * M_PI * u2, &s, &c);
    z0 = rad * c;
    z1 = rad * s;
}

}  // namespace rv
