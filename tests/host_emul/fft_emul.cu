// CPU emulation of the cooperative FFT passes in csrc/fft_core.cuh: every "thread" b of every pass
// is run in sequence with a barrier = loop boundary.  Checks butterflies, twiddle indices, the padded
// shared-memory indexing, and the two-real-in-one-complex packing against a long double DFT.
// Built and run by tests/test_host_emulation.py (no GPU needed).
#include <cstdio>
#include <cmath>
#include <vector>
#include <cstdlib>
#include "../../3d_navier_stokes_b200/csrc/fft_kernels.cuh"

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
    if constexpr (P::ROW == P::M1) {   // unpadded plan: the strided kernel lands the tile in natural order and runs pass 1 in place
        for (int i = 0; i < N; ++i) s[(size_t)i * STRIDE] = x[i];
        for (int b = 0; b < P::NB1; ++b) fft_pass1_inplace<P, DIR, STRIDE>(b, s, tw.data());
    } else {
        for (int b = 0; b < P::NB1; ++b) fft_pass1<P, DIR, STRIDE>(b, s, tw.data(), [&](int n) { return x[n]; });
    }
    if constexpr (P::PASSES >= 3)
        for (int b = 0; b < P::NB2; ++b) {
            if constexpr (P::PASSES == 4) fft_pass2_w1<P, DIR, STRIDE>(b, s, tw[P::R1 * (b % P::M2)]);
            else fft_pass2<P, DIR, STRIDE>(b, s, tw.data());
        }
    if constexpr (P::PASSES == 4)
        for (int b = 0; b < P::NB3; ++b) {
            if (b & 1) fft_pass3_w1<P, DIR, STRIDE>(b, s, tw[P::R1 * P::R2 * (b % P::R4)]);
            else fft_pass3<P, DIR, STRIDE>(b, s, tw.data());
        }
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

// Emulates k_z_fused's data flow for one pencil pair (teams, in-place rows, partner look-up) and checks
// it against the straightforward formulation: c2r of six half spectra, u x w, r2c of the product.
template <class P> int check_fused() {
    constexpr int N = P::N, TP = P::NB1, NP = P::NPAD;
    static_assert(P::R1 == P::RL, "balanced plan");
    std::vector<cplx> tw(N), sm(6 * NP, mk(1e300, 1e300));
    for (int m = 0; m < N; ++m) {
        long double a = -2.0L * 3.14159265358979323846264338327950288L * m / N;
        tw[m] = mk((double)cosl(a), (double)sinl(a));
    }
    std::vector<cplx> rows[6][2], outA[3], outB[3];
    for (int f = 0; f < 6; ++f) for (int r = 0; r < 2; ++r) { rows[f][r].resize(N / 2 + 1); for (auto& z : rows[f][r]) z = mk(frand(), frand()); }
    auto rowbase = [](int q) { return fft_row_base<P>(q); };
    // inverse
    for (int t = 0; t < 3; ++t) for (int ff = 0; ff < 2; ++ff) for (int q = 0; q < TP; ++q) {
        int f = 2 * t + ff;
        cplx w1[P::R1 - 1]; load_tw_pass1<P>(q, tw.data(), w1);
        fft_pass1_rw<P, INV, 1>(q, sm.data() + f * NP, w1, [&](int n) { int k = n <= N / 2 ? n : N - n; return pack_hermitian<N>(n, rows[f][0][k], rows[f][1][k]); });
    }
    constexpr bool RW2 = (P::PASSES == 3) && (P::NB2 == TP);
    if constexpr (P::PASSES >= 3)
        for (int f = 0; f < 6; ++f) for (int b = 0; b < P::NB2; ++b) {
            if constexpr (RW2) { cplx w2[P::R2 - 1]; load_tw_pass2<P>(b, tw.data(), w2); fft_pass2_rw<P, INV, 1>(b, sm.data() + f * NP, w2); }
            else if constexpr (P::PASSES == 4) fft_pass2_w1<P, INV, 1>(b, sm.data() + f * NP, tw[P::R1 * (b % P::M2)]);
            else fft_pass2<P, INV, 1>(b, sm.data() + f * NP, tw.data());
        }
    if constexpr (P::PASSES == 4)
        for (int f = 0; f < 6; ++f) for (int b = 0; b < P::NB3; ++b) fft_pass3_w1<P, INV, 1>(b, sm.data() + f * NP, tw[P::R1 * P::R2 * (b % P::R4)]);
    for (int f = 0; f < 6; ++f) for (int q = 0; q < TP; ++q) {
        cplx v[P::RL]; fft_pass_last<P, INV, 1>(q, sm.data() + f * NP, v);
        for (int j = 0; j < P::RL; ++j) sm[f * NP + rowbase(q) + j] = v[j];
    }
    // reference real-space fields via naive c2r
    std::vector<double> real[6][2];
    for (int f = 0; f < 6; ++f) for (int r = 0; r < 2; ++r) {
        real[f][r].resize(N);
        for (int n = 0; n < N; ++n) {
            long double sacc = 0;
            for (int k = 0; k < N; ++k) {
                int kk = k <= N / 2 ? k : N - k;
                long double ar = rows[f][r][kk].x, ai = (k <= N / 2 ? rows[f][r][kk].y : -rows[f][r][kk].y);
                if (k == 0 || k == N / 2) ai = 0;
                long double a = 2.0L * 3.14159265358979323846264338327950288L * ((long long)n * k % N) / N;
                sacc += ar * cosl(a) - ai * sinl(a);
            }
            real[f][r][n] = (double)sacc;
        }
    }
    // product + forward pass 1 in registers, then scatter
    std::vector<std::vector<cplx>> creg(3 * TP, std::vector<cplx>(P::R1));
    for (int t = 0; t < 3; ++t) for (int q = 0; q < TP; ++q) {
        int i1 = (t + 1) % 3, i2 = (t + 2) % 3;
        cplx c[P::R1];
        for (int j = 0; j < P::R1; ++j)
            c[j] = cross_comp(sm[i1 * NP + rowbase(q) + j], sm[(3 + i2) * NP + rowbase(q) + j], sm[i2 * NP + rowbase(q) + j], sm[(3 + i1) * NP + rowbase(q) + j]);
        { cplx w1[P::R1 - 1]; load_tw_pass1<P>(q, tw.data(), w1); fft_pass1_regs_rw<P, FWD>(c, w1); }
        for (int j = 0; j < P::R1; ++j) creg[t * TP + q][j] = c[j];
    }
    for (int t = 0; t < 3; ++t) for (int q = 0; q < TP; ++q) fft_pass1_scatter<P, 1>(q, sm.data() + t * NP, creg[t * TP + q].data());
    if constexpr (P::PASSES >= 3)
        for (int t = 0; t < 3; ++t) for (int b = 0; b < P::NB2; ++b) {
            if constexpr (RW2) { cplx w2[P::R2 - 1]; load_tw_pass2<P>(b, tw.data(), w2); fft_pass2_rw<P, FWD, 1>(b, sm.data() + t * NP, w2); }
            else if constexpr (P::PASSES == 4) fft_pass2_w1<P, FWD, 1>(b, sm.data() + t * NP, tw[P::R1 * (b % P::M2)]);
            else fft_pass2<P, FWD, 1>(b, sm.data() + t * NP, tw.data());
        }
    if constexpr (P::PASSES == 4)
        for (int t = 0; t < 3; ++t) for (int b = 0; b < P::NB3; ++b) fft_pass3_w1<P, FWD, 1>(b, sm.data() + t * NP, tw[P::R1 * P::R2 * (b % P::R4)]);
    std::vector<std::vector<cplx>> vreg(3 * TP, std::vector<cplx>(P::RL));
    for (int t = 0; t < 3; ++t) for (int q = 0; q < TP; ++q) {
        cplx v[P::RL]; fft_pass_last<P, FWD, 1>(q, sm.data() + t * NP, v);
        for (int j = 0; j < P::RL; ++j) vreg[t * TP + q][j] = v[j];
    }
    for (int t = 0; t < 3; ++t) for (int q = 0; q < TP; ++q) for (int j = 0; j < P::RL; ++j) sm[t * NP + rowbase(q) + j] = vreg[t * TP + q][j];
    for (int t = 0; t < 3; ++t) { outA[t].assign(N / 2 + 1, mk(0, 0)); outB[t].assign(N / 2 + 1, mk(0, 0)); }
    for (int t = 0; t < 3; ++t) for (int q = 0; q < TP; ++q) for (int j = 0; j <= P::RL / 2; ++j) {
        int k = q + j * P::NBL;
        if (k <= N / 2) {
            int m = (N - k) & (N - 1), qm = m % P::NBL, jm = m / P::NBL;
            cplx Zm = sm[t * NP + rowbase(qm) + jm];
            unpack_pair(vreg[t * TP + q][j], Zm, outA[t][k], outB[t][k]);
        }
    }
    // reference: r2c of the real-space product
    double err = 0, nrm = 0;
    for (int t = 0; t < 3; ++t) for (int r = 0; r < 2; ++r) {
        int i1 = (t + 1) % 3, i2 = (t + 2) % 3;
        std::vector<double> c(N);
        for (int n = 0; n < N; ++n) c[n] = real[i1][r][n] * real[3 + i2][r][n] - real[i2][r][n] * real[3 + i1][r][n];
        for (int k = 0; k <= N / 2; ++k) {
            long double sr = 0, si = 0;
            for (int n = 0; n < N; ++n) {
                long double a = -2.0L * 3.14159265358979323846264338327950288L * ((long long)n * k % N) / N;
                sr += c[n] * cosl(a); si += c[n] * sinl(a);
            }
            cplx got = r == 0 ? outA[t][k] : outB[t][k];
            err = fmax(err, fmax(fabs(got.x - (double)sr), fabs(got.y - (double)si)));
            nrm = fmax(nrm, fmax(fabs((double)sr), fabs((double)si)));
        }
    }
    printf("fused N=%d (%d,%d,%d): max rel err %.3e\n", N, P::R1, P::R2, P::R3, err / nrm);
    return err / nrm < 1e-13 ? 0 : 1;
}

// Emulates k_z_fused_w (one warp per transform, mirrored butterfly pairs per lane) for one pencil pair with the
// kernel's own lane helpers, and checks it against the straightforward formulation.
template <class P> int check_fused_warp() {
    constexpr int N = P::N, NP = P::NPAD;
    std::vector<cplx> tw(N), sm(6 * NP, mk(1e300, 1e300));
    for (int m = 0; m < N; ++m) {
        long double a = -2.0L * 3.14159265358979323846264338327950288L * m / N;
        tw[m] = mk((double)cosl(a), (double)sinl(a));
    }
    std::vector<cplx> rows[6][2], outA[3], outB[3];
    for (int f = 0; f < 6; ++f) for (int r = 0; r < 2; ++r) { rows[f][r].resize(N / 2 + 1); for (auto& z : rows[f][r]) z = mk(frand(), frand()); }
    for (int t = 0; t < 3; ++t) { outA[t].assign(N / 2 + 1, mk(1e300, 1e300)); outB[t].assign(N / 2 + 1, mk(1e300, 1e300)); }
    // inverse transforms, warp t: fields (t+1)%3 and 3 + (t+2)%3
    for (int t = 0; t < 3; ++t) for (int ff = 0; ff < 2; ++ff) {
        const int f = ff ? 3 + (t + 2) % 3 : (t + 1) % 3;
        cplx* buf = sm.data() + f * NP;
        for (int L = 0; L < 32; ++L) {
            cplx w1[7]; zw_load_tw1<P>(L, tw.data(), w1);
            zw_inv_pass1<P>(L, buf, w1, [&](int k, cplx& A, cplx& B) { A = rows[f][0][k]; B = rows[f][1][k]; });
        }
        for (int L = 0; L < 32; ++L) {
            cplx w2[P::R2 - 1]; load_tw_pass2<P>(L, tw.data(), w2);
            for (int i = 0; i < P::NB2 / 32; ++i) fft_pass2_rw<P, INV, 1>(L + 32 * i, buf, w2);
        }
        std::vector<cplx> keep(32 * 16);
        for (int L = 0; L < 32; ++L) zw_last_pair<P, INV>(L, buf, &keep[L * 16], &keep[L * 16 + 8]);
        for (int L = 0; L < 32; ++L) {
            int bA, bB; bool self; zw_lane_pair<P>(L, bA, bB, self);
            for (int j = 0; j < 8; ++j) { buf[fft_row_base<P>(bA) + j] = keep[L * 16 + j]; buf[fft_row_base<P>(bB) + j] = keep[L * 16 + 8 + j]; }
        }
    }
    std::vector<double> real[6][2];
    for (int f = 0; f < 6; ++f) for (int r = 0; r < 2; ++r) {
        real[f][r].resize(N);
        for (int n = 0; n < N; ++n) {
            long double sacc = 0;
            for (int k = 0; k < N; ++k) {
                int kk = k <= N / 2 ? k : N - k;
                long double ar = rows[f][r][kk].x, ai = (k <= N / 2 ? rows[f][r][kk].y : -rows[f][r][kk].y);
                if (k == 0 || k == N / 2) ai = 0;
                long double a = 2.0L * 3.14159265358979323846264338327950288L * ((long long)n * k % N) / N;
                sacc += ar * cosl(a) - ai * sinl(a);
            }
            real[f][r][n] = (double)sacc;
        }
    }
    // cross product + forward pass 1 in registers (before the second barrier), then scatter
    std::vector<cplx> creg(3 * 32 * 16);
    for (int t = 0; t < 3; ++t) for (int L = 0; L < 32; ++L) {
        int bA, bB; bool self; zw_lane_pair<P>(L, bA, bB, self);
        const int i1 = (t + 1) % 3, i2 = (t + 2) % 3, rbA = fft_row_base<P>(bA), rbB = fft_row_base<P>(bB);
        cplx* ca = &creg[(t * 32 + L) * 16]; cplx* cb = ca + 8;
        for (int j = 0; j < 8; ++j) {
            ca[j] = cross_comp(sm[i1 * NP + rbA + j], sm[(3 + i2) * NP + rbA + j], sm[i2 * NP + rbA + j], sm[(3 + i1) * NP + rbA + j]);
            cb[j] = cross_comp(sm[i1 * NP + rbB + j], sm[(3 + i2) * NP + rbB + j], sm[i2 * NP + rbB + j], sm[(3 + i1) * NP + rbB + j]);
        }
        cplx w1[7]; zw_load_tw1<P>(L, tw.data(), w1);
        zw_bfly_pair<FWD>(self, ca, cb, w1);
    }
    for (int t = 0; t < 3; ++t) {
        cplx* buf = sm.data() + ((t + 1) % 3) * NP;
        for (int L = 0; L < 32; ++L) {
            int bA, bB; bool self; zw_lane_pair<P>(L, bA, bB, self);
            zw_scatter_pair<P>(bA, bB, buf, &creg[(t * 32 + L) * 16], &creg[(t * 32 + L) * 16 + 8]);
        }
        for (int L = 0; L < 32; ++L) {
            cplx w2[P::R2 - 1]; load_tw_pass2<P>(L, tw.data(), w2);
            for (int i = 0; i < P::NB2 / 32; ++i) fft_pass2_rw<P, FWD, 1>(L + 32 * i, buf, w2);
        }
        for (int L = 0; L < 32; ++L) {
            cplx va[8], vb[8];
            zw_last_pair<P, FWD>(L, buf, va, vb);
            zw_unpack_store<P>(L, va, vb, [&](int k, cplx A, cplx B) { outA[t][k] = A; outB[t][k] = B; });
        }
    }
    double err = 0, nrm = 0;
    for (int t = 0; t < 3; ++t) for (int r = 0; r < 2; ++r) {
        int i1 = (t + 1) % 3, i2 = (t + 2) % 3;
        std::vector<double> c(N);
        for (int n = 0; n < N; ++n) c[n] = real[i1][r][n] * real[3 + i2][r][n] - real[i2][r][n] * real[3 + i1][r][n];
        for (int k = 0; k <= N / 2; ++k) {
            long double sr = 0, si = 0;
            for (int n = 0; n < N; ++n) {
                long double a = -2.0L * 3.14159265358979323846264338327950288L * ((long long)n * k % N) / N;
                sr += c[n] * cosl(a); si += c[n] * sinl(a);
            }
            cplx got = r == 0 ? outA[t][k] : outB[t][k];
            err = fmax(err, fmax(fabs(got.x - (double)sr), fabs(got.y - (double)si)));
            nrm = fmax(nrm, fmax(fabs((double)sr), fabs((double)si)));
        }
    }
    printf("fused (warp per transform) N=%d (%d,%d,%d): max rel err %.3e\n", N, P::R1, P::R2, P::R3, err / nrm);
    return err / nrm < 1e-13 ? 0 : 1;
}

// stand-alone warp z passes (k_z_c2r_w / k_z_r2c_w) with LPT = M1/2 lanes per transform: the lane program of one pencil
// pair, lanes emulated one after the other between the __syncwarp points.  c2r against a long double DFT, then r2c of the
// exact real rows against a long double DFT.
template <class P> int check_warp_passes() {
    constexpr int N = P::N, NP = P::NPAD, LPT = P::M1 / 2;
    std::vector<cplx> tw(N), buf(NP, mk(1e300, 1e300));
    for (int m = 0; m < N; ++m) {
        long double a = -2.0L * 3.14159265358979323846264338327950288L * m / N;
        tw[m] = mk((double)cosl(a), (double)sinl(a));
    }
    std::vector<cplx> rows[2];
    for (int r = 0; r < 2; ++r) { rows[r].resize(N / 2 + 1); for (auto& z : rows[r]) z = mk(frand(), frand()); }
    // ---- c2r
    for (int l = 0; l < LPT; ++l) {
        cplx w1[7]; zw_load_tw1<P>(l, tw.data(), w1);
        zw_inv_pass1<P>(l, buf.data(), w1, [&](int k, cplx& A, cplx& B) { A = rows[0][k]; B = rows[1][k]; });
    }
    for (int l = 0; l < LPT; ++l) zw_pass2<P, INV>(l, buf.data(), ZwTw2<P>(l, tw.data()));
    std::vector<double> got[2]; got[0].assign(N, 1e300); got[1].assign(N, 1e300);
    for (int l = 0; l < LPT; ++l) {
        int bA, bB; bool self; zw_lane_pair<P>(l, bA, bB, self);
        cplx va[8], vb[8];
        zw_last_pair<P, INV>(l, buf.data(), va, vb);
        for (int j = 0; j < 8; ++j) {
            got[0][bA + j * P::M1] = va[j].x; got[1][bA + j * P::M1] = va[j].y;
            got[0][bB + j * P::M1] = vb[j].x; got[1][bB + j * P::M1] = vb[j].y;
        }
    }
    std::vector<double> real[2];
    double err = 0, nrm = 0;
    for (int r = 0; r < 2; ++r) {
        real[r].resize(N);
        for (int n = 0; n < N; ++n) {
            long double sacc = 0;
            for (int k = 0; k < N; ++k) {
                int kk = k <= N / 2 ? k : N - k;
                long double ar = rows[r][kk].x, ai = (k <= N / 2 ? rows[r][kk].y : -rows[r][kk].y);
                if (k == 0 || k == N / 2) ai = 0;
                long double a = 2.0L * 3.14159265358979323846264338327950288L * ((long long)n * k % N) / N;
                sacc += ar * cosl(a) - ai * sinl(a);
            }
            real[r][n] = (double)sacc;
            err = fmax(err, fabs(got[r][n] - real[r][n])); nrm = fmax(nrm, fabs(real[r][n]));
        }
    }
    // ---- r2c of the exact real rows
    std::vector<cplx> outA(N / 2 + 1, mk(1e300, 1e300)), outB(N / 2 + 1, mk(1e300, 1e300));
    std::vector<cplx> creg(LPT * 16);
    for (int l = 0; l < LPT; ++l) {
        int bA, bB; bool self; zw_lane_pair<P>(l, bA, bB, self);
        cplx* ca = &creg[l * 16]; cplx* cb = ca + 8;
        for (int j = 0; j < 8; ++j) { ca[j] = mk(real[0][bA + j * P::M1], real[1][bA + j * P::M1]); cb[j] = mk(real[0][bB + j * P::M1], real[1][bB + j * P::M1]); }
        cplx w1[7]; zw_load_tw1<P>(l, tw.data(), w1);
        zw_bfly_pair<FWD>(self, ca, cb, w1);
    }
    for (int l = 0; l < LPT; ++l) {
        int bA, bB; bool self; zw_lane_pair<P>(l, bA, bB, self);
        zw_scatter_pair<P>(bA, bB, buf.data(), &creg[l * 16], &creg[l * 16 + 8]);
    }
    for (int l = 0; l < LPT; ++l) zw_pass2<P, FWD>(l, buf.data(), ZwTw2<P>(l, tw.data()));
    for (int l = 0; l < LPT; ++l) {
        cplx va[8], vb[8];
        zw_last_pair<P, FWD>(l, buf.data(), va, vb);
        zw_unpack_store<P>(l, va, vb, [&](int k, cplx A, cplx B) { outA[k] = A; outB[k] = B; });
    }
    double err2 = 0, nrm2 = 0;
    for (int r = 0; r < 2; ++r) for (int k = 0; k <= N / 2; ++k) {
        long double sr = 0, si = 0;
        for (int n = 0; n < N; ++n) {
            long double a = -2.0L * 3.14159265358979323846264338327950288L * ((long long)n * k % N) / N;
            sr += real[r][n] * cosl(a); si += real[r][n] * sinl(a);
        }
        cplx g = r == 0 ? outA[k] : outB[k];
        err2 = fmax(err2, fmax(fabs(g.x - (double)sr), fabs(g.y - (double)si)));
        nrm2 = fmax(nrm2, fmax(fabs((double)sr), fabs((double)si)));
    }
    printf("warp z passes N=%d (%d,%d,%d), %d lanes per transform: c2r max rel err %.3e, r2c %.3e\n", N, P::R1, P::R2, P::R3, LPT, err / nrm, err2 / nrm2);
    return (err / nrm < 1e-13 && err2 / nrm2 < 1e-13) ? 0 : 1;
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
    bad += check<FftPlan<1024, 8, 8, 16, 0>>("big3");
    bad += check<ZPlan<64>::type>("z");
    bad += check<ZPlan<128>::type>("z");
    bad += check<ZPlan<256>::type>("z");
    bad += check<ZFPlan<32>::type>("zf");
    bad += check<ZFPlan<128>::type>("zf");
    bad += check<ZFPlan<256>::type>("zf");
    bad += check<ZFPlan<1024>::type>("zf");
    bad += check_fused<ZFPlan<16>::type>();
    bad += check_fused<ZFPlan<32>::type>();
    bad += check_fused<ZFPlan<64>::type>();
    bad += check_fused<ZFPlan<128>::type>();
    bad += check_fused<ZFPlan<256>::type>();
    bad += check_fused<ZFPlan<512>::type>();
    bad += check<ZFPlan<1024>::type>("zf4");
    bad += check<FftPlan<256, 4, 4, 4, 1, 4>>("zf4");
    bad += check_fused<FftPlan<256, 4, 4, 4, 1, 4>>();
    bad += check_fused<ZFPlan<1024>::type>();
    bad += check_fused_warp<ZFPlan<512>::type>();
    bad += check_warp_passes<ZWPlan<512>::type>();
    bad += check_warp_passes<ZWPlan<256>::type>();
    bad += check_warp_passes<ZWPlan<128>::type>();
    bad += check_warp_passes<FftPlan<1024, 8, 16, 8>>();   // lane program of the two-warp kernels at 1024
    bad += check_pack<16>();
    bad += check_pack<64>();
    bad += check_pack<512>();
    printf(bad ? "FAIL\n" : "OK\n");
    return bad;
}
