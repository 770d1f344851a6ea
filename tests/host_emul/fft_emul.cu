// CPU emulation of the cooperative FFT passes in csrc/fft_core.cuh: every "thread" b of every pass
// is run in sequence with a barrier = loop boundary.  Checks butterflies, twiddle indices, the padded
// shared-memory indexing, and the two-real-in-one-complex packing against a long double DFT.
// Built and run by tests/test_host_emulation.py (no GPU needed).
#include <cstdio>
#include <cmath>
#include <vector>
#include <cstdlib>
#include "../../3d_navier_stokes_b200/csrc/fft_core.cuh"

static double frand() { return (double)rand() / RAND_MAX - 0.5; }

template <class P, int DIR, int STRIDE> double run_plan() {
    const int N = P::N;
    std::vector<cplx> tw(N), x(N), out(N), sm((size_t)P::NPAD * STRIDE, mk(1e300, 1e300));
    for (int m = 0; m < N; ++m) {
        long double a = -2.0L * 3.14159265358979323846264338327950288L * m / N;
        tw[m] = mk((double)cosl(a), (double)sinl(a));
    }
    for (int i = 0; i < N; ++i) x[i] = mk(frand(), frand());
    cplx* s = sm.data() + (STRIDE > 1 ? 1 : 0);  // pencil p = 1 of an interleaved tile
    for (int b = 0; b < P::NB1; ++b) fft_pass1<P, DIR, STRIDE>(b, s, tw.data(), [&](int n) { return x[n]; });
    if constexpr (P::PASSES == 3)
        for (int b = 0; b < P::NB2; ++b) fft_pass2<P, DIR, STRIDE>(b, s, tw.data());
    for (int b = 0; b < P::NBL; ++b) {
        cplx v[P::RL];
        fft_pass_last<P, DIR, STRIDE>(b, s, v);
        for (int k2 = 0; k2 < P::RL; ++k2) out[b + k2 * P::NBL] = v[k2];
    }
    double err = 0, nrm = 0;
    for (int k = 0; k < N; ++k) {
        long double sr = 0, si = 0;
        for (int n = 0; n < N; ++n) {
            long double a = (long double)DIR * 2.0L * 3.14159265358979323846264338327950288L * ((long long)n * k % N) / N;
            long double c = cosl(a), sn = sinl(a);
            sr += x[n].x * c - x[n].y * sn;
            si += x[n].x * sn + x[n].y * c;
        }
        err = fmax(err, fmax(fabs(out[k].x - (double)sr), fabs(out[k].y - (double)si)));
        nrm = fmax(nrm, fmax(fabs((double)sr), fabs((double)si)));
    }
    return err / nrm;
}

template <class P> int check(const char* name) {
    double e1 = run_plan<P, FWD, 1>(), e2 = run_plan<P, INV, 1>(), e3 = run_plan<P, FWD, 8>(), e4 = run_plan<P, INV, 4>();
    double e = fmax(fmax(e1, e2), fmax(e3, e4));
    printf("%s N=%d (%d,%d,%d): max rel err %.3e\n", name, P::N, P::R1, P::R2, P::R3, e);
    return e < 5e-15 ? 0 : 1;
}

// pair packing: c2r then r2c of two random Hermitian half spectra
template <int N> int check_pack() {
    typedef typename ZPlan<N>::type P;
    std::vector<cplx> tw(N), A(N / 2 + 1), B(N / 2 + 1), z(N), sm(P::NPAD), Z(N);
    for (int m = 0; m < N; ++m) {
        long double a = -2.0L * 3.14159265358979323846264338327950288L * m / N;
        tw[m] = mk((double)cosl(a), (double)sinl(a));
    }
    for (int k = 0; k <= N / 2; ++k) { A[k] = mk(frand(), frand()); B[k] = mk(frand(), frand()); }
    // inverse
    for (int b = 0; b < P::NB1; ++b)
        fft_pass1<P, INV, 1>(b, sm.data(), tw.data(), [&](int n) { int k = n <= N / 2 ? n : N - n; return pack_hermitian<N>(n, A[k], B[k]); });
    if constexpr (P::PASSES == 3) for (int b = 0; b < P::NB2; ++b) fft_pass2<P, INV, 1>(b, sm.data(), tw.data());
    for (int b = 0; b < P::NBL; ++b) { cplx v[P::RL]; fft_pass_last<P, INV, 1>(b, sm.data(), v); for (int j = 0; j < P::RL; ++j) z[b + j * P::NBL] = v[j]; }
    // reference c2r for a: a(n) = sum_k Ahat(k) e^{+i...} with Hermitian extension, Im DC/Nyq dropped
    double err = 0;
    for (int n = 0; n < N; ++n) {
        long double sa = 0, sb = 0;
        for (int k = 0; k < N; ++k) {
            int kk = k <= N / 2 ? k : N - k;
            long double ar = A[kk].x, ai = (k <= N / 2 ? A[kk].y : -A[kk].y), br = B[kk].x, bi = (k <= N / 2 ? B[kk].y : -B[kk].y);
            if (k == 0 || k == N / 2) { ai = 0; bi = 0; }
            long double a = 2.0L * 3.14159265358979323846264338327950288L * ((long long)n * k % N) / N;
            sa += ar * cosl(a) - ai * sinl(a);
            sb += br * cosl(a) - bi * sinl(a);
        }
        err = fmax(err, fmax(fabs(z[n].x - (double)sa), fabs(z[n].y - (double)sb)));
    }
    // forward again and unpack: must return N * (A, B) with real DC/Nyquist
    for (int b = 0; b < P::NB1; ++b) fft_pass1<P, FWD, 1>(b, sm.data(), tw.data(), [&](int n) { return z[n]; });
    if constexpr (P::PASSES == 3) for (int b = 0; b < P::NB2; ++b) fft_pass2<P, FWD, 1>(b, sm.data(), tw.data());
    for (int b = 0; b < P::NBL; ++b) { cplx v[P::RL]; fft_pass_last<P, FWD, 1>(b, sm.data(), v); for (int j = 0; j < P::RL; ++j) Z[b + j * P::NBL] = v[j]; }
    double err2 = 0;
    for (int k = 0; k <= N / 2; ++k) {
        cplx a, b2;
        unpack_pair(Z[k], Z[(N - k) % N], a, b2);
        double ai = (k == 0 || k == N / 2) ? 0.0 : A[k].y, bi = (k == 0 || k == N / 2) ? 0.0 : B[k].y;
        err2 = fmax(err2, fmax(fabs(a.x / N - A[k].x), fabs(a.y / N - ai)));
        err2 = fmax(err2, fmax(fabs(b2.x / N - B[k].x), fabs(b2.y / N - bi)));
    }
    printf("pack N=%d: c2r err %.3e  round-trip err %.3e\n", N, err / N, err2);
    return (err / N < 5e-15 && err2 < 5e-15) ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += check<BigPlan<16>::type>("big");
    bad += check<BigPlan<32>::type>("big");
    bad += check<BigPlan<64>::type>("big");
    bad += check<BigPlan<128>::type>("big");
    bad += check<BigPlan<256>::type>("big");
    bad += check<BigPlan<512>::type>("big");
    bad += check<BigPlan<1024>::type>("big");
    bad += check<ZPlan<64>::type>("z");
    bad += check<ZPlan<128>::type>("z");
    bad += check<ZPlan<256>::type>("z");
    bad += check_pack<16>();
    bad += check_pack<64>();
    bad += check_pack<512>();
    printf(bad ? "FAIL\n" : "OK\n");
    return bad;
}
