/* TEST INFRASTRUCTURE ONLY (oracle/) -- never linked into, imported by or shipped with the product.
 *
 * Plain-C restatement of the reference's native MEX kernels, which ship as source-less binaries
 * (socp/<variant>/utils/mex*.mexa64).  Semantics were recovered from `objdump -d -M intel`
 * of those binaries and are pinned BIT-FOR-BIT against the binaries themselves by
 * tests/test_oracle_kernels.py (when oracle/_ref is present) and against the golden vectors
 * in tests/golden/kernels_*.npz (generated from the binaries by tests/golden/make_golden.py).
 *
 * Memory layout everywhere = MATLAB column-major: y fastest, then x, then t.
 *   q  = [ q0 (nt-1,nx,ny) | bx (nt,nx-1,ny) | by (nt,nx,ny-1) ]           (C-order views)
 *   z  = L x 10 column-major (structure of arrays), L = (nt-1)*nx*ny; 1-D: L x 6, L=(nt-1)*nx
 *
 * No FMA contraction is allowed in this file (the binaries are scalar SSE2 mulsd/addsd):
 * build with -ffp-contract=off (see oracle/Makefile).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* mexBFd.mexa64 .rodata @0x2000 : 0x3fe6a09e667f3bd1 -- the decimal literal 0.707106781186548,
 * NOT sqrt(0.5) (= ...3bcd).  SCALEFACTOR = 0.707106781186548 * SCALE (mexFunction @0x15ac). */
static double inv_sqrt2_literal(void)
{
    const uint64_t bits = 0x3fe6a09e667f3bd1ULL;
    double v;
    memcpy(&v, &bits, sizeof v);
    return v;
}

/* ---------------------------------------------------------------------------------------------
 * mexBFd(z2, q, nt, nx, ny, scaleBF, scaleD)          mexBFd.mexa64 @0x14f0
 *   oper_BFd_a @0x1120 : col0 = DF - q0*S ; col9 = q0*S + DF
 *   oper_BFd_b @0x11a0 : bx -> cols 1..4     oper_BFd_c @0x1310 : by -> cols 5..8
 * Entries whose neighbour lies outside the domain are NOT written (they keep the caller's values,
 * zeros in the solver: solver_socp_inPALM.m:131).
 * ------------------------------------------------------------------------------------------- */
void oracle_BFd(double *z, const double *q, int nt, int nx, int ny, double S, double DF)
{
    const ptrdiff_t L = (ptrdiff_t)(nt - 1) * nx * ny;
    const double SF = inv_sqrt2_literal() * S;
    const double *q0 = q;
    const double *bx = q + L;
    const double *by = bx + (ptrdiff_t)nt * (nx - 1) * ny;
    ptrdiff_t i;
    int t, x, y;

    for (i = 0; i < L; i++) {
        double p = q0[i] * S;
        z[i] = DF - p;
    }
    for (i = 0; i < L; i++) {
        double p = q0[i] * S;
        z[9 * L + i] = p + DF;
    }
    for (t = 0; t < nt - 1; t++)
        for (x = 0; x < nx; x++)
            for (y = 0; y < ny; y++) {
                const ptrdiff_t c = ((ptrdiff_t)t * nx + x) * ny + y;
                if (x >= 1) {
                    z[1 * L + c] = bx[((ptrdiff_t)t * (nx - 1) + (x - 1)) * ny + y] * SF;
                    z[3 * L + c] = bx[((ptrdiff_t)(t + 1) * (nx - 1) + (x - 1)) * ny + y] * SF;
                }
                if (x <= nx - 2) {
                    z[2 * L + c] = bx[((ptrdiff_t)t * (nx - 1) + x) * ny + y] * SF;
                    z[4 * L + c] = bx[((ptrdiff_t)(t + 1) * (nx - 1) + x) * ny + y] * SF;
                }
                if (y >= 1) {
                    z[5 * L + c] = by[((ptrdiff_t)t * nx + x) * (ny - 1) + (y - 1)] * SF;
                    z[7 * L + c] = by[((ptrdiff_t)(t + 1) * nx + x) * (ny - 1) + (y - 1)] * SF;
                }
                if (y <= ny - 2) {
                    z[6 * L + c] = by[((ptrdiff_t)t * nx + x) * (ny - 1) + y] * SF;
                    z[8 * L + c] = by[((ptrdiff_t)(t + 1) * nx + x) * (ny - 1) + y] * SF;
                }
            }
}

/* ---------------------------------------------------------------------------------------------
 * mexBFdConj(q2, z, nt, nx, ny, scaleBF)              mexBFdConj.mexa64 @0x1610
 *   a @0x1120 : q0 = (z9 - z0) * S
 *   b @0x1160 : bx[t,x,y] = SF * sum over the (up to) four cells that read it in mexBFd
 *   c @0x1310 : by likewise.
 * Add order (addsd chain @0x1260): ((z1[t,x+1] + z2[t,x]) + z3[t-1,x+1]) + z4[t-1,x]; boundary time
 * levels use the two-term forms @0x11c0 / @0x12e8.
 * ------------------------------------------------------------------------------------------- */
void oracle_BFdConj(double *q, const double *z, int nt, int nx, int ny, double S)
{
    const ptrdiff_t L = (ptrdiff_t)(nt - 1) * nx * ny;
    const double SF = inv_sqrt2_literal() * S;
    double *q0 = q;
    double *bx = q + L;
    double *by = bx + (ptrdiff_t)nt * (nx - 1) * ny;
    ptrdiff_t i;
    int t, x, y;

    for (i = 0; i < L; i++)
        q0[i] = (z[9 * L + i] - z[i]) * S;

    for (t = 0; t < nt; t++)
        for (x = 0; x < nx - 1; x++)
            for (y = 0; y < ny; y++) {
                const ptrdiff_t cu = ((ptrdiff_t)t * nx + x) * ny + y;        /* cell (t,  x, y) */
                const ptrdiff_t cd = ((ptrdiff_t)(t - 1) * nx + x) * ny + y;  /* cell (t-1,x, y) */
                double s;
                if (t == 0)
                    s = z[1 * L + cu + ny] + z[2 * L + cu];
                else if (t == nt - 1)
                    s = z[3 * L + cd + ny] + z[4 * L + cd];
                else {
                    s = z[1 * L + cu + ny] + z[2 * L + cu];
                    s = s + z[3 * L + cd + ny];
                    s = s + z[4 * L + cd];
                }
                bx[((ptrdiff_t)t * (nx - 1) + x) * ny + y] = s * SF;
            }

    for (t = 0; t < nt; t++)
        for (x = 0; x < nx; x++)
            for (y = 0; y < ny - 1; y++) {
                const ptrdiff_t cu = ((ptrdiff_t)t * nx + x) * ny + y;
                const ptrdiff_t cd = ((ptrdiff_t)(t - 1) * nx + x) * ny + y;
                double s;
                if (t == 0)
                    s = z[5 * L + cu + 1] + z[6 * L + cu];
                else if (t == nt - 1)
                    s = z[7 * L + cd + 1] + z[8 * L + cd];
                else {
                    s = z[5 * L + cu + 1] + z[6 * L + cu];
                    s = s + z[7 * L + cd + 1];
                    s = s + z[8 * L + cd];
                }
                by[((ptrdiff_t)t * nx + x) * (ny - 1) + y] = s * SF;
            }
}

/* ---------------------------------------------------------------------------------------------
 * mexProjSoc(out, in)      M x N column-major, col 0 = t        mexProjSoc.mexa64 @0x1170
 *
 * Row norm = Eigen `in.rightCols(N-1).rowwise().norm()` (@0x1440).  Eigen evaluates it two rows
 * per SSE2 packet with a 4-way unrolled column loop:
 *     acc = s[0];  for k = 1, 5, 9, ... while k < ((N-2) & ~3):  acc += (s[k+3]+s[k+2]) + (s[k+1]+s[k]);
 *     remaining columns are added one at a time;                       (s[k] = in(i,k+1)^2)
 * and, when M is odd, the last row by a plain sequential loop (@0x1668).  Both orders are restated
 * so that the result is bit-identical to the binary (new[] buffers are 16-byte aligned => no peeled
 * head row).
 * coefficient (@0x1320): r = (v0/nrm + 1) * 0.5 ;
 *     r > 1  -> coef = 1, keep v0        0 > r -> coef = 0        else coef = r, keep v0 iff r == 1
 * out(i,j) = in(i,j) * coef (j>=1) ; out(i,0) = keep ? in(i,0) : coef * nrm.
 * v0 = 0 and nrm = 0 gives 0/0 = NaN in every column, exactly like the binary.
 * ------------------------------------------------------------------------------------------- */
static double rownorm_packet(const double *in, ptrdiff_t M, int N, ptrdiff_t i)
{
    const int n = N - 1; /* summed columns */
    double acc;
    int k, kend;
    if (n <= 0)
        return 0.0;
    acc = in[1 * M + i] * in[1 * M + i];
    kend = (n - 1) & ~3;
    k = 1;
    if (kend > 1) {
        for (; k < kend; k += 4) {
            const double s0 = in[(k + 1) * M + i] * in[(k + 1) * M + i];
            const double s1 = in[(k + 2) * M + i] * in[(k + 2) * M + i];
            const double s2 = in[(k + 3) * M + i] * in[(k + 3) * M + i];
            const double s3 = in[(k + 4) * M + i] * in[(k + 4) * M + i];
            const double hi = s3 + s2;
            const double lo = s1 + s0;
            acc = acc + (hi + lo);
        }
    }
    for (; k < n; k++) {
        const double s = in[(k + 1) * M + i] * in[(k + 1) * M + i];
        acc = acc + s;
    }
    return sqrt(acc);
}

static double rownorm_scalar(const double *in, ptrdiff_t M, int N, ptrdiff_t i)
{
    const int n = N - 1;
    double acc;
    int k;
    if (n <= 0)
        return 0.0;
    acc = in[1 * M + i] * in[1 * M + i];
    for (k = 1; k < n; k++) {
        const double s = in[(k + 1) * M + i] * in[(k + 1) * M + i];
        acc = acc + s;
    }
    return sqrt(acc);
}

void oracle_ProjSoc(double *out, const double *in, ptrdiff_t M, int N)
{
    const ptrdiff_t Mpair = M & ~(ptrdiff_t)1;
    ptrdiff_t i;
    int j;
    for (i = 0; i < M; i++) {
        const double nrm = (i < Mpair) ? rownorm_packet(in, M, N, i) : rownorm_scalar(in, M, N, i);
        const double v0 = in[i];
        double r = (v0 / nrm + 1.0) * 0.5;
        double coef;
        int keep;
        if (r > 1.0) {
            coef = 1.0;
            keep = 1;
        } else if (0.0 > r) {
            coef = 0.0;
            keep = 0;
        } else {
            coef = r;
            keep = (r == 1.0);
        }
        for (j = 1; j < N; j++)
            out[j * M + i] = in[j * M + i] * coef;
        out[i] = keep ? v0 : coef * nrm;
    }
}

/* ---------------------------------------------------------------------------------------------
 * 1-D kernels: mexBFd1d(z, q, nt, nx [,scale [,dFactor]])   mexBFd1d.mexa64 @0x1320
 *              mexBFdConj1d(q, z, nt, nx, scale)             mexBFdConj1d.mexa64 @0x1350
 * columns: c0 = DF - q0*S, c1 = bx[t,x-1]*SF, c2 = bx[t,x]*SF, c3 = bx[t+1,x-1]*SF, c4 = bx[t+1,x]*SF,
 *          c5 = q0*S + DF.   Identical to the 2-D kernels with ny = 1 and columns 5..8 dropped.
 * (The optional-argument statefulness of the binary -- omitted scale/dFactor keep the previous call's
 *  values -- is a property of its file-static globals; the oracle's Python wrapper models it.)
 * ------------------------------------------------------------------------------------------- */
void oracle_BFd1d(double *z, const double *q, int nt, int nx, double S, double DF)
{
    const ptrdiff_t L = (ptrdiff_t)(nt - 1) * nx;
    const double SF = inv_sqrt2_literal() * S;
    const double *q0 = q;
    const double *bx = q + L;
    ptrdiff_t i;
    int t, x;
    for (i = 0; i < L; i++) {
        double p = q0[i] * S;
        z[i] = DF - p;
    }
    for (i = 0; i < L; i++) {
        double p = q0[i] * S;
        z[5 * L + i] = p + DF;
    }
    for (t = 0; t < nt - 1; t++)
        for (x = 0; x < nx; x++) {
            const ptrdiff_t c = (ptrdiff_t)t * nx + x;
            if (x >= 1) {
                z[1 * L + c] = bx[(ptrdiff_t)t * (nx - 1) + (x - 1)] * SF;
                z[3 * L + c] = bx[(ptrdiff_t)(t + 1) * (nx - 1) + (x - 1)] * SF;
            }
            if (x <= nx - 2) {
                z[2 * L + c] = bx[(ptrdiff_t)t * (nx - 1) + x] * SF;
                z[4 * L + c] = bx[(ptrdiff_t)(t + 1) * (nx - 1) + x] * SF;
            }
        }
}

void oracle_BFdConj1d(double *q, const double *z, int nt, int nx, double S)
{
    const ptrdiff_t L = (ptrdiff_t)(nt - 1) * nx;
    const double SF = inv_sqrt2_literal() * S;
    double *q0 = q;
    double *bx = q + L;
    ptrdiff_t i;
    int t, x;
    for (i = 0; i < L; i++)
        q0[i] = (z[5 * L + i] - z[i]) * S;
    for (t = 0; t < nt; t++)
        for (x = 0; x < nx - 1; x++) {
            const ptrdiff_t cu = (ptrdiff_t)t * nx + x;
            const ptrdiff_t cd = (ptrdiff_t)(t - 1) * nx + x;
            double s;
            if (t == 0)
                s = z[1 * L + cu + 1] + z[2 * L + cu];
            else if (t == nt - 1)
                s = z[3 * L + cd + 1] + z[4 * L + cd];
            else {
                s = z[1 * L + cu + 1] + z[2 * L + cu];
                s = s + z[3 * L + cd + 1];
                s = s + z[4 * L + cd];
            }
            bx[(ptrdiff_t)t * (nx - 1) + x] = s * SF;
        }
}
