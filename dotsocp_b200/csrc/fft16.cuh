// fft16.cuh -- register-resident radix-16 building blocks of the power-of-two FFTs inside the Bluestein DCT (v2 kernel).
//
// An M = 256*Q2 point FFT (Q2 in {1,2,4,8,16}) runs on TP = M/16 threads, 16 complex values per thread, as
//   pass 1: radix 16 over elements  j + m*TP            (thread j)              twiddle W_M^(f*j)
//   pass 2: radix 16 over elements  b*TP + k + m*Q2      (thread b*Q2 + k)       twiddle W_M^(16*f*k)
//   pass 3: radix Q2 over 16/Q2 groups of the contiguous elements [16j, 16j+16)  no twiddle
// with one shared-memory exchange between consecutive passes.  The spectrum comes out digit-reversed
// (position m1*TP + m2*Q2 + m3  <->  frequency m1 + 16*m2 + 256*m3); the point-wise product with the chirp spectrum is
// done in that order in registers and the inverse transform runs the transposed graph (pass 3', 2', 1').
// Everything is __host__ __device__ so that the index algebra is unit-tested on the CPU (tests/host/fft16_host_test.cu).
#pragma once
#include <cuda_runtime.h>

namespace dsocp {

#define FHD __host__ __device__ __forceinline__

FHD double2 c_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
FHD double2 c_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
FHD double2 c_mul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
FHD double2 c_mulc(double2 a, double2 b) { return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }   // a*conj(b)
FHD double2 c_conj(double2 a) { return make_double2(a.x, -a.y); }
template <int SIGN> FHD double2 c_muli(double2 a) { return SIGN < 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x); }
// multiply by exp(SIGN * i * pi/4 * e), e = 1 or 3 (the sqrt(1/2) twiddles)
template <int SIGN, int E> FHD double2 c_mul8(double2 a)
{
    const double h = 0.70710678118654752440;
    if (E == 1) return SIGN < 0 ? make_double2((a.x + a.y) * h, (a.y - a.x) * h) : make_double2((a.x - a.y) * h, (a.y + a.x) * h);
    return SIGN < 0 ? make_double2((a.y - a.x) * h, (-a.x - a.y) * h) : make_double2((-a.x - a.y) * h, (a.x - a.y) * h);
}

template <int SIGN> FHD void r_dft2(double2& a, double2& b)
{
    const double2 t = c_sub(a, b);
    a = c_add(a, b);
    b = t;
}
template <int SIGN> FHD void r_dft4(double2& a0, double2& a1, double2& a2, double2& a3)
{
    const double2 t0 = c_add(a0, a2), t1 = c_sub(a0, a2), t2 = c_add(a1, a3), t3 = c_muli<SIGN>(c_sub(a1, a3));
    a0 = c_add(t0, t2);
    a1 = c_add(t1, t3);
    a2 = c_sub(t0, t2);
    a3 = c_sub(t1, t3);
}
template <int SIGN> FHD void r_dft8(double2* v)   // natural in, natural out
{
    double2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    r_dft4<SIGN>(e0, e1, e2, e3);
    r_dft4<SIGN>(o0, o1, o2, o3);
    const double2 t1 = c_mul8<SIGN, 1>(o1), t2 = c_muli<SIGN>(o2), t3 = c_mul8<SIGN, 3>(o3);
    v[0] = c_add(e0, o0); v[4] = c_sub(e0, o0);
    v[1] = c_add(e1, t1); v[5] = c_sub(e1, t1);
    v[2] = c_add(e2, t2); v[6] = c_sub(e2, t2);
    v[3] = c_add(e3, t3); v[7] = c_sub(e3, t3);
}
// 16-point DFT, natural in / natural out, as 4 x 4 with the w16 twiddles in between
template <int SIGN> FHD void r_dft16(double2* v)
{
    const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;   // cos(pi/8), sin(pi/8)
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) r_dft4<SIGN>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);   // -> index 4*k1 + n2
    // element (k1, n2) *= w16^(n2*k1)
    const double2 w1 = make_double2(c1, SIGN < 0 ? -s1 : s1), w3 = make_double2(s1, SIGN < 0 ? -c1 : c1);
    v[4 + 1] = c_mul(v[4 + 1], w1);                 // e = 1
    v[4 + 2] = c_mul8<SIGN, 1>(v[4 + 2]);           // e = 2
    v[4 + 3] = c_mul(v[4 + 3], w3);                 // e = 3
    v[8 + 1] = c_mul8<SIGN, 1>(v[8 + 1]);           // e = 2
    v[8 + 2] = c_muli<SIGN>(v[8 + 2]);              // e = 4
    v[8 + 3] = c_mul8<SIGN, 3>(v[8 + 3]);           // e = 6
    v[12 + 1] = c_mul(v[12 + 1], w3);               // e = 3
    v[12 + 2] = c_mul8<SIGN, 3>(v[12 + 2]);         // e = 6
    {                                               // e = 9: w16^9 = -w16^1
        const double2 t = c_mul(v[12 + 3], w1);
        v[12 + 3] = make_double2(-t.x, -t.y);
    }
    double2 o[16];
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) {
        r_dft4<SIGN>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);               // -> X[k1 + 4*k2] at 4*k1 + k2
#pragma unroll
        for (int k2 = 0; k2 < 4; k2++) o[k1 + 4 * k2] = v[4 * k1 + k2];
    }
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = o[i];
}
// radix-Q DFTs on the 16/Q contiguous groups of a thread's 16 registers (pass 3 / 3')
template <int SIGN, int Q> FHD void r_dft_groups(double2* v)
{
    if (Q == 2) {
#pragma unroll
        for (int g = 0; g < 8; g++) r_dft2<SIGN>(v[2 * g], v[2 * g + 1]);
    } else if (Q == 4) {
#pragma unroll
        for (int g = 0; g < 4; g++) r_dft4<SIGN>(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
    } else if (Q == 8) {
        r_dft8<SIGN>(v);
        r_dft8<SIGN>(v + 8);
    } else if (Q == 16) {
        r_dft16<SIGN>(v);
    }
}
// v[f] *= w^f (f = 1..15) for forward, conj for inverse, powers built by a multiplication ladder (depth <= 4)
template <bool CONJ> FHD void twiddle_powers(double2* v, double2 w1)
{
    if (CONJ) w1 = c_conj(w1);
    const double2 w2 = c_mul(w1, w1), w3 = c_mul(w2, w1), w4 = c_mul(w2, w2);
    const double2 w5 = c_mul(w4, w1), w6 = c_mul(w3, w3), w7 = c_mul(w4, w3), w8 = c_mul(w4, w4);
    v[1] = c_mul(v[1], w1); v[2] = c_mul(v[2], w2); v[3] = c_mul(v[3], w3); v[4] = c_mul(v[4], w4);
    v[5] = c_mul(v[5], w5); v[6] = c_mul(v[6], w6); v[7] = c_mul(v[7], w7); v[8] = c_mul(v[8], w8);
    v[9] = c_mul(v[9], c_mul(w8, w1)); v[10] = c_mul(v[10], c_mul(w8, w2)); v[11] = c_mul(v[11], c_mul(w8, w3));
    v[12] = c_mul(v[12], c_mul(w8, w4)); v[13] = c_mul(v[13], c_mul(w8, w5)); v[14] = c_mul(v[14], c_mul(w8, w6));
    v[15] = c_mul(v[15], c_mul(w8, w7));
}

#define PAD16(i) ((i) + ((i) >> 4))

// ---- the three index patterns --------------------------------------------------------------------------------------
template <int LOG2M> struct F16 {
    static constexpr int M = 1 << LOG2M, TP = M / 16, Q2 = M / 256;
    static_assert(LOG2M >= 8 && LOG2M <= 12, "M = 256 .. 4096");
    FHD static int pos1(int j, int m) { return j + m * TP; }                                  // pass 1 / 1'
    FHD static int pos2(int t, int m) { return (t / Q2) * TP + (t % Q2) + m * Q2; }            // pass 2 / 2'
    FHD static int pos3(int j, int e) { return 16 * j + e; }                                   // pass 3 / 3'
    template <class S> FHD static void load1(const S* s, int j, double2* v) { for (int m = 0; m < 16; m++) v[m] = s[PAD16(pos1(j, m))]; }
    template <class S> FHD static void store1(S* s, int j, const double2* v) { for (int m = 0; m < 16; m++) s[PAD16(pos1(j, m))] = v[m]; }
    template <class S> FHD static void load2(const S* s, int t, double2* v) { for (int m = 0; m < 16; m++) v[m] = s[PAD16(pos2(t, m))]; }
    template <class S> FHD static void store2(S* s, int t, const double2* v) { for (int m = 0; m < 16; m++) s[PAD16(pos2(t, m))] = v[m]; }
    template <class S> FHD static void load3(const S* s, int j, double2* v) { for (int e = 0; e < 16; e++) v[e] = s[PAD16(pos3(j, e))]; }
    template <class S> FHD static void store3(S* s, int j, const double2* v) { for (int e = 0; e < 16; e++) s[PAD16(pos3(j, e))] = v[e]; }
    // forward pass 1 on registers holding x[pos1(j, m)]: result y_f (twiddled) belongs at pos1(j, f)
    FHD static void fwd1(double2* v, int j, const double2* tw) { r_dft16<-1>(v); twiddle_powers<false>(v, tw[j]); }
    FHD static void fwd2(double2* v, int t, const double2* tw) { r_dft16<-1>(v); if (Q2 > 1) twiddle_powers<false>(v, tw[16 * (t % Q2)]); }
    FHD static void fwd3(double2* v) { if (Q2 > 1) r_dft_groups<-1, Q2>(v); }
    FHD static void inv3(double2* v) { if (Q2 > 1) r_dft_groups<1, Q2>(v); }
    FHD static void inv2(double2* v, int t, const double2* tw) { if (Q2 > 1) twiddle_powers<true>(v, tw[16 * (t % Q2)]); r_dft16<1>(v); }
    FHD static void inv1(double2* v, int j, const double2* tw) { twiddle_powers<true>(v, tw[j]); r_dft16<1>(v); }
    // frequency held at a position after the three forward passes
    FHD static int freq_of_pos(int p)
    {
        const int m1 = p / TP, r = p % TP, m2 = r / Q2, m3 = r % Q2;
        return m1 + 16 * m2 + 256 * m3;
    }
};

}  // namespace dsocp
