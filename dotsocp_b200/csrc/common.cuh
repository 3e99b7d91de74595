// common.cuh -- shared geometry / helper definitions for libdotsocp (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dsocp {

typedef long long i64;

// mexBFd.mexa64 .rodata @0x2000: the reference multiplies scaleBF by the decimal literal 0.707106781186548
// (bits 0x3fe6a09e667f3bd1), not by sqrt(1/2); keep the same constant so results match to the last bit.
#define DSOCP_INV_SQRT2_BITS 0x3fe6a09e667f3bd1ULL

// Grid geometry of one level (nodes), with the staggered-array sizes of SURVEY.md A.1.
struct Geo {
    int nt, nx, ny;
    i64 P;      // nx*ny          nodes per time level
    i64 PBX;    // (nx-1)*ny      bx edges per time level
    i64 PBY;    // nx*(ny-1)      by edges per time level
    i64 L;      // (nt-1)*P       cells  (= q0 entries)
    i64 NBX;    // nt*PBX
    i64 NBY;    // nt*PBY
    i64 Q;      // L+NBX+NBY
    i64 N;      // nt*P
};

inline Geo make_geo(int nt, int nx, int ny)
{
    Geo g;
    g.nt = nt; g.nx = nx; g.ny = ny;
    g.P = (i64)nx * ny;
    g.PBX = (i64)(nx - 1) * ny;
    g.PBY = (i64)nx * (ny - 1);
    g.L = (i64)(nt - 1) * g.P;
    g.NBX = (i64)nt * g.PBX;
    g.NBY = (i64)nt * g.PBY;
    g.Q = g.L + g.NBX + g.NBY;
    g.N = (i64)nt * g.P;
    return g;
}

// Scalars of the iteration (names follow solver_socp_inPALM.m:54-59, 96-97).
struct IterScal {
    double S;      // scaleBF = E/D
    double SF;     // 0.707106781186548 * S
    double DF;     // scaleD  = E/dScale
    double tau;
    double gt, gx, gy;   // D/ht, D/hx, D/hy  (entries of model.grad after InitialScaling)
    double dinv1;  // 1/(1+2 s^2)   diagQInv away from the first/last time level (oper_q.m:17-26)
    double dinv2;  // 1/(1+  s^2)   bx,by at t = first/last
    double s2x2;   // 2 s^2   (weighted: diagQ = w^2 + {2 s^2 | s^2}, wdot2d/utils/oper_q.m:15-28)
    double s2x1;   //   s^2
};

// exact (never contracted) double arithmetic where the operation order of the reference is mirrored
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

}  // namespace dsocp
