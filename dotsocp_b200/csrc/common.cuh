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
//
// Node arrays (phi, rhs, c) are always packed: (t, x, y) at t*P + x*ny + y.  The staggered arrays (q, alpha, q2, weight:
// [q0 | bx | by]) and the 10 columns of beta / z have a row PITCH: q0 / cell (t,x,y) at t*PC + x*py + y, bx at
// L + t*PBX + x*py + y, by at L + NBX + t*PBY + x*pyb + y.  make_geo() gives the reference's packed layout (py = ny,
// pyb = ny-1: what crosses the C ABI); make_geo_padded() rounds both pitches up to 32 doubles so that every row starts on a
// 256-byte boundary -- the layout of the device-resident sessions: with packed rows a warp's 256-byte store straddles two
// partially written 32-byte sectors, which costs the 21-read / 13-write march of k_mult a third of its bandwidth
// (tools/stream_pattern3.cu: 6.05 TB/s aligned, 4.29 TB/s with ny = 513).  Pad entries (y >= ny, resp. y >= ny-1 for by)
// are zero and stay zero.
struct Geo {
    int nt, nx, ny;
    int py;     // row pitch of q0 / bx / cell rows (>= ny)
    int pyb;    // row pitch of by rows (>= ny-1)
    i64 P;      // nx*ny          nodes per time level (packed)
    i64 PC;     // nx*py          q0 entries / cells per time level
    i64 PBX;    // (nx-1)*py      bx edges per time level
    i64 PBY;    // nx*pyb         by edges per time level
    i64 L;      // (nt-1)*PC      cells  (= q0 entries)
    i64 NBX;    // nt*PBX
    i64 NBY;    // nt*PBY
    i64 Q;      // L+NBX+NBY
    i64 N;      // nt*P
    bool packed() const { return py == ny && pyb == ny - 1; }
};

inline Geo make_geo_pitched(int nt, int nx, int ny, int py, int pyb)
{
    Geo g;
    g.nt = nt; g.nx = nx; g.ny = ny;
    g.py = py; g.pyb = pyb;
    g.P = (i64)nx * ny;
    g.PC = (i64)nx * py;
    g.PBX = (i64)(nx - 1) * py;
    g.PBY = (i64)nx * pyb;
    g.L = (i64)(nt - 1) * g.PC;
    g.NBX = (i64)nt * g.PBX;
    g.NBY = (i64)nt * g.PBY;
    g.Q = g.L + g.NBX + g.NBY;
    g.N = (i64)nt * g.P;
    return g;
}
inline Geo make_geo(int nt, int nx, int ny) { return make_geo_pitched(nt, nx, ny, ny, ny - 1); }
// rows of 32-double multiples (the 1-D variant, ny == 1, has no rows to align and stays packed)
inline Geo make_geo_padded(int nt, int nx, int ny)
{
    if (ny <= 1) return make_geo(nt, nx, ny);
    const int py = (ny + 31) / 32 * 32;
    return make_geo_pitched(nt, nx, ny, py, py);
}

// One thread per entry p of a cell plane (p < PC): its (x, y), whether it is a real column, and its node-plane index.
struct PlanePos {
    int x, y;
    bool ok;      // y < ny (pad columns of a pitched layout are skipped)
    i64 pn;       // x*ny + y : index inside a node plane
};
__host__ __device__ __forceinline__ PlanePos plane_pos(const Geo& g, i64 p)
{
    PlanePos r;
    r.x = (int)(p / g.py);
    r.y = (int)(p - (i64)r.x * g.py);
    r.ok = (p < g.PC) && (r.y < g.ny);
    r.pn = (i64)r.x * g.ny + r.y;
    return r;
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
