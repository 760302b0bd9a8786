// vec3.h -- 3-vector with OpenFOAM's evaluation order (left-to-right sums,
// component-wise scalar division).  Shared by the host mesh baker and the CUDA
// kernels of libdsmcb200; compiled without FMA contraction on both sides so the
// baked tables and the device arithmetic round identically.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define DSMC_HD __host__ __device__ __forceinline__
#else
#define DSMC_HD inline
#endif

namespace dsmc {

constexpr double SMALL = 1.0e-15;
constexpr double VSMALL = 1.0e-300;
constexpr double ROOTVSMALL = 1.0e-150;
constexpr double GREAT = 1.0e15;
constexpr double VGREAT = 1.0e300;
constexpr double PI = 3.14159265358979323846;
constexpr double TWO_PI = 6.28318530717958647692;

struct V3 {
    double x, y, z;
};

DSMC_HD V3 mk(double x, double y, double z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
DSMC_HD V3 operator+(const V3& a, const V3& b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
DSMC_HD V3 operator-(const V3& a, const V3& b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
DSMC_HD V3 operator-(const V3& a) { return mk(-a.x, -a.y, -a.z); }
DSMC_HD V3 operator*(double s, const V3& a) { return mk(s * a.x, s * a.y, s * a.z); }
DSMC_HD V3 operator*(const V3& a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
DSMC_HD V3 operator/(const V3& a, double s) { return mk(a.x / s, a.y / s, a.z / s); }
DSMC_HD V3& operator+=(V3& a, const V3& b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
DSMC_HD V3& operator-=(V3& a, const V3& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
DSMC_HD V3& operator/=(V3& a, double s) { a.x /= s; a.y /= s; a.z /= s; return a; }
DSMC_HD V3& operator*=(V3& a, double s) { a.x *= s; a.y *= s; a.z *= s; return a; }
// OpenFOAM '&' (inner product) and '^' (cross product)
DSMC_HD double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DSMC_HD V3 cross(const V3& a, const V3& b) {
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
DSMC_HD double magSqr(const V3& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
DSMC_HD double mag(const V3& a) { return sqrt(magSqr(a)); }
DSMC_HD double comp(const V3& a, int d) { return d == 0 ? a.x : (d == 1 ? a.y : a.z); }
DSMC_HD void setComp(V3& a, int d, double v) { if (d == 0) a.x = v; else if (d == 1) a.y = v; else a.z = v; }

}  // namespace dsmc
