// device_models.cuh -- kinetic-theory samplers shared by the wall, inflow, fill and collision
// kernels.  Each function cites the dsmcCloud member it stands in for.
#pragma once
#include "engine.h"

namespace dsmc {

// dsmcCloud::equipartitionRotationalEnergy, DSMC/clouds/dsmcCloud.C:1181-1218
__device__ __forceinline__ double equipartitionRotationalEnergy(Rng& rng, double kB, double T, double rotDof) {
    double ERot = 0.0;
    if (rotDof < SMALL) return ERot;
    if (rotDof < 2.0 + SMALL && rotDof > 2.0 - SMALL) {
        ERot = -log(rng.sample01()) * kB * T;
    } else {
        double a = 0.5 * rotDof - 1;
        double energyRatio, P;
        do {
            energyRatio = 10 * rng.sample01();
            P = pow(energyRatio / a, a) * exp(a - energyRatio);
        } while (P < rng.sample01());
        ERot = energyRatio * kB * T;
    }
    return ERot;
}

// dsmcCloud::equipartitionVibrationalEnergyLevel, DSMC/clouds/dsmcCloud.C:1221-1244 (one mode)
__device__ __forceinline__ int32_t equipartitionVibrationalEnergyLevel(Rng& rng, double T, double thetaV) {
    return int32_t(-log(rng.sample01()) * T / thetaV);
}

// dsmcCloud::equipartitionElectronicLevel, DSMC/clouds/dsmcCloud.C:1247-1324
__device__ __forceinline__ int32_t equipartitionElectronicLevel(Rng& rng, double kB, double T, const DevSpecies& sp) {
    const double EMax = kB * T;
    const int jMax = sp.nElec - 1;
    int jDash = 0;
    if (jMax > 0 && T > SMALL) {
        double expSum = 0.0;
        for (int i = 0; i <= jMax; ++i) expSum += sp.gElec[i] * exp(-sp.eElec[i] / EMax);
        double boltzMax = 0.0;
        int jSelect = 0;
        for (int i = 0; i <= jMax; ++i) {
            double boltz = sp.gElec[i] * exp(-sp.eElec[i] / EMax) / expSum;
            if (boltzMax < boltz) { boltzMax = boltz; jSelect = i; }
        }
        const double expMax = sp.gElec[jSelect] * exp(-sp.eElec[jSelect] / EMax);
        double func;
        do {
            jDash = rng.randomLabel(0, jMax);
            func = sp.gElec[jDash] * exp(-sp.eElec[jDash] / EMax) / expMax;
        } while (func < rng.sample01());
    }
    return jDash;
}

// dsmcCloud::equipartitionLinearVelocity, DSMC/clouds/dsmcCloud.C:1043-1051
__device__ __forceinline__ V3 equipartitionLinearVelocity(Rng& rng, double kB, double T, double mass) {
    double s = sqrt(kB * T / mass);
    double gx = rng.gaussNormal(), gy = rng.gaussNormal(), gz = rng.gaussNormal();
    return s * mk(gx, gy, gz);
}

// tetrahedron::randomPoint (OpenFOAM v1706 tetrahedronI.H)
__device__ __forceinline__ V3 tetRandomPoint(Rng& rng, const V3& a, const V3& b, const V3& c, const V3& d) {
    double s = rng.sample01(), t = rng.sample01(), u = rng.sample01();
    if (s + t > 1.0) { s = 1.0 - s; t = 1.0 - t; }
    if (t + u > 1.0) {
        double tmp = u; u = 1.0 - s - t; t = 1.0 - tmp;
    } else if (s + t + u > 1.0) {
        double tmp = u; u = s + t + u - 1.0; s = 1.0 - t - tmp;
    }
    return (1 - s - t - u) * a + s * b + t * c + u * d;
}

// triangle::randomPoint (OpenFOAM v1706 triangleI.H)
__device__ __forceinline__ V3 triRandomPoint(Rng& rng, const V3& a, const V3& b, const V3& c) {
    double s = rng.sample01();
    double t = sqrt(rng.sample01());
    return (1 - t) * a + (1 - s) * t * b + s * t * c;
}

__device__ __forceinline__ double atomicAddDouble(double* addr, double v) { return atomicAdd(addr, v); }

}  // namespace dsmc
