// philox.h -- counter-based Philox-4x32-10 streams (Salmon et al., SC'11).
//
// The reference draws every random number from one sequential Random stream owned by the
// cloud (DSMC/clouds/dsmcCloud.H:153), which cannot be reproduced by a parallel engine;
// north_star asks for counter-based Philox instead.  A stream is addressed by
//   key     = 64-bit seed (dsmcProperties seedNumber)
//   counter = (entity, sub-entity, step, stream<<24 | block)
// and yields two 53-bit uniforms in [0,1) per block, so any draw of any parcel / cell /
// candidate can be regenerated independently of execution order.
#pragma once
#include <cstdint>

#include "vec3.h"

namespace dsmc {

enum PhiloxStream : uint32_t {
    STREAM_REMAINDER = 1,  // collisionSelectionRemainder initialisation
    STREAM_FILL = 2,       // dsmcMeshFill
    STREAM_COLLIDE = 3,    // noTimeCounter candidate selection + collision model
    STREAM_WALL = 4,       // wall-model draws of a parcel
    STREAM_INFLOW = 5,     // free-stream inflow (per face)
    STREAM_NEWPARCEL = 6,  // random step fraction of freshly inserted parcels
    STREAM_WEIGHT = 7      // dsmcAxisymmetric::axisymmetricWeighting (per sorted parcel)
};

#if defined(__CUDACC__)
#define DSMC_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define DSMC_HD_NOINLINE inline
#endif

// out of line on the device: ~40 call sites otherwise inline 10 rounds each and the collision kernel no longer
// fits the instruction cache
#ifdef DSMC_PHILOX_LOCAL_STATE
DSMC_HD_NOINLINE void philox4x32_10_to(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = uint64_t(M0) * c0;
        uint64_t p1 = uint64_t(M1) * c2;
        uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = uint32_t(p1);
        uint32_t n2 = uint32_t(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = uint32_t(p0);
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
#endif
struct PhiloxBlock { uint32_t x0, x1, x2, x3; };   // returned in registers: a pointer argument would pin the caller's buffer to local memory
DSMC_HD_NOINLINE PhiloxBlock philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = uint64_t(M0) * c0;
        uint64_t p1 = uint64_t(M1) * c2;
        uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = uint32_t(p1);
        uint32_t n2 = uint32_t(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = uint32_t(p0);
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    PhiloxBlock b;
    b.x0 = c0; b.x1 = c1; b.x2 = c2; b.x3 = c3;
    return b;
}

// DSMC_PHILOX_LOCAL_STATE (defined by kernels_move.cu): the block lives in an indexed array, i.e. in local memory.  The move kernel
// draws random numbers only in its rare wall path; keeping that state out of the register file is worth 7 % of the stage there
// (4.76 vs 5.11 ms at 31 M parcels), while the collide kernels gain 12 % from the state in registers (1.31 -> 1.15 ms).
struct Rng {
    uint32_t c0, c1, c2, c3base, k0, k1;
    uint32_t idx;      // next draw index
    uint32_t cached;   // block currently held in b
#ifdef DSMC_PHILOX_LOCAL_STATE
    uint32_t buf[4];
#else
    PhiloxBlock b;     // scalars selected by predicate, never indexed: the state stays in registers
#endif

    DSMC_HD void init(uint64_t seed, uint32_t entity, uint32_t sub, uint32_t step, uint32_t stream, uint32_t firstDraw = 0) {
        k0 = uint32_t(seed); k1 = uint32_t(seed >> 32);
        c0 = entity; c1 = sub; c2 = step; c3base = stream << 24;
        idx = firstDraw;
        cached = 0xFFFFFFFFu;
    }
    // uniform in [0,1): the analogue of Random::sample01<scalar>()
    DSMC_HD double sample01() {
        const uint32_t blk = (idx >> 1) & 0xFFFFFFu;
#ifdef DSMC_PHILOX_LOCAL_STATE
        if (blk != cached) { philox4x32_10_to(c0, c1, c2, c3base | blk, k0, k1, buf); cached = blk; }
        uint32_t lo = buf[(idx & 1u) * 2], hi = buf[(idx & 1u) * 2 + 1];
#else
        if (blk != cached) { b = philox4x32_10(c0, c1, c2, c3base | blk, k0, k1); cached = blk; }
        const bool second = (idx & 1u) != 0;
        uint32_t lo = second ? b.x2 : b.x0, hi = second ? b.x3 : b.x1;
#endif
        ++idx;
        uint64_t v = (uint64_t(hi) << 32) | lo;
        return double(v >> 11) * (1.0 / 9007199254740992.0);
    }
    // dsmcCloud::randomLabel (DSMC/clouds/dsmcCloud.C:1015-1040); sample01 < 1 so no redraw is needed
    DSMC_HD int32_t randomLabel(int32_t valOne, int32_t valTwo) {
        if (valOne == valTwo) return valOne;
        int32_t start = valOne < valTwo ? valOne : valTwo;
        int32_t end = valOne < valTwo ? valTwo : valOne;
        return start + int32_t(sample01() * double(end - start + 1));
    }
    // standard normal (Random::GaussNormal<scalar> analogue; Box-Muller instead of the polar method)
    DSMC_HD double gaussNormal() {
        double u1 = 1.0 - sample01();  // (0,1]
        double u2 = sample01();
        return sqrt(-2.0 * log(u1)) * cos(TWO_PI * u2);
    }
};

}  // namespace dsmc
