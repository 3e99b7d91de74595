// Host check of the index algebra in dotsocp_b200/csrc/fft16.cuh: the three register passes, run "thread" by "thread" on
// the CPU with an array standing in for shared memory, must give the DFT in the digit-reversed order freq_of_pos()
// describes, and the transposed passes must invert it.  Built and run by tests/test_native_host_cpu.py (no GPU needed).
#include <cmath>
#include <cstdio>
#include <vector>

#include "fft16.cuh"

using namespace dsocp;

template <int LOG2M> static int check()
{
    typedef F16<LOG2M> F;
    const int M = F::M, TP = F::TP;
    const double PI = 3.14159265358979323846;
    std::vector<double2> x(M), s(M + M / 16 + 1), tw(M);
    for (int i = 0; i < M; i++) {
        x[i] = make_double2(std::sin(0.37 * i) + 0.01 * i, std::cos(1.3 * i) - 0.5);
        tw[i] = make_double2(std::cos(2 * PI * i / M), -std::sin(2 * PI * i / M));
        s[PAD16(i)] = x[i];
    }
    double2 v[16];
    for (int j = 0; j < TP; j++) { F::load1(s.data(), j, v); F::fwd1(v, j, tw.data()); F::store1(s.data(), j, v); }
    for (int t = 0; t < TP; t++) { F::load2(s.data(), t, v); F::fwd2(v, t, tw.data()); F::store2(s.data(), t, v); }
    for (int j = 0; j < TP; j++) { F::load3(s.data(), j, v); F::fwd3(v); F::store3(s.data(), j, v); }
    double err = 0, scale = 0;
    std::vector<char> seen(M, 0);
    for (int p = 0; p < M; p++) {
        const int f = F::freq_of_pos(p);
        if (f < 0 || f >= M || seen[f]) { printf("LOG2M=%d: freq_of_pos is not a permutation\n", LOG2M); return 1; }
        seen[f] = 1;
        long double re = 0, im = 0;
        for (int n = 0; n < M; n++) {
            const long double a = -2.0L * PI * (long double)((long long)f * n % M) / M;
            re += x[n].x * cosl(a) - x[n].y * sinl(a);
            im += x[n].x * sinl(a) + x[n].y * cosl(a);
        }
        err = std::fmax(err, std::fmax(std::fabs((double)re - s[PAD16(p)].x), std::fabs((double)im - s[PAD16(p)].y)));
        scale = std::fmax(scale, std::fabs((double)re));
    }
    if (err > 1e-11 * scale) { printf("LOG2M=%d: forward error %.3e (scale %.3e)\n", LOG2M, err, scale); return 1; }
    for (int j = 0; j < TP; j++) { F::load3(s.data(), j, v); F::inv3(v); F::store3(s.data(), j, v); }
    for (int t = 0; t < TP; t++) { F::load2(s.data(), t, v); F::inv2(v, t, tw.data()); F::store2(s.data(), t, v); }
    for (int j = 0; j < TP; j++) { F::load1(s.data(), j, v); F::inv1(v, j, tw.data()); F::store1(s.data(), j, v); }
    double ierr = 0;
    for (int i = 0; i < M; i++)
        ierr = std::fmax(ierr, std::fmax(std::fabs(s[PAD16(i)].x / M - x[i].x), std::fabs(s[PAD16(i)].y / M - x[i].y)));
    if (ierr > 1e-12 * (1 + 0.01 * M)) { printf("LOG2M=%d: round-trip error %.3e\n", LOG2M, ierr); return 1; }
    printf("LOG2M=%d ok (forward %.2e, round trip %.2e)\n", LOG2M, err / scale, ierr);
    return 0;
}

int main()
{
    int bad = check<8>() + check<9>() + check<10>() + check<11>() + check<12>();
    if (!bad) printf("FFT16_HOST_OK\n");
    return bad;
}
