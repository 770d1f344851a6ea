// FP64 cooperative FFT engine for sm_100a: radix-2/4/8/16 butterflies in registers, twiddles from an
// L1-resident table, digit exchange through shared memory.  Replaces the FFTW codelets behind
// fftw_mpi_execute_dft_r2c/c2r (reference call sites solver.c:656,658,683).
//
// A length-N transform is executed by TP threads in 2 or 3 radix passes (N = R1*R2[*R3]):
//   pass 1  thread b loads x[b + j*N/R1] (unit stride across threads -> coalesced), R1-point DFT,
//           twiddle W_N^{b*k1}, scatter to shared memory row k1
//   pass 2  (3-pass plans) in-place R2-point DFTs inside each row, twiddle W_{N/R1}^{m2*k1'}
//   last    thread b = k1 + R1*k1' gathers its R_last contiguous elements, DFT, and owns outputs
//           X[b + j*N/R_last] (again unit stride across threads)
// so input and output are both in natural order and both coalesced.  Shared-memory index i is
// padded to i + i/(N/R1) which makes every access pattern above bank-conflict free for 16-byte
// elements.  Everything is __host__ __device__ so tests/host_emul can run the passes on the CPU.
#pragma once
#include <cuda_runtime.h>

#define NSB_HD __host__ __device__ __forceinline__

typedef double2 cplx;

constexpr int FWD = -1;  // exp(-i ...)  r2c direction
constexpr int INV = +1;  // exp(+i ...)  c2r direction

NSB_HD cplx mk(double x, double y) { cplx r; r.x = x; r.y = y; return r; }
NSB_HD cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
NSB_HD cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
NSB_HD cplx cmul(cplx a, cplx b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
NSB_HD cplx cconj(cplx a) { return mk(a.x, -a.y); }
NSB_HD cplx cscale(cplx a, double s) { return mk(a.x * s, a.y * s); }
// multiply by -i (forward) or +i (inverse)
template <int DIR> NSB_HD cplx rot(cplx a) { return DIR == FWD ? mk(a.y, -a.x) : mk(-a.y, a.x); }
// multiply by exp(DIR * i * theta) given (c, s) = (cos theta, sin theta)
template <int DIR> NSB_HD cplx cmul_cs(cplx a, double c, double s) {
    return DIR == FWD ? mk(a.x * c + a.y * s, a.y * c - a.x * s) : mk(a.x * c - a.y * s, a.y * c + a.x * s);
}
template <int DIR> NSB_HD cplx cmul_dir(cplx a, cplx w) {   // a * w (FWD) or a * conj(w) (INV)
    return DIR == FWD ? mk(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x) : mk(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y);
}
// table holds exp(-2 pi i m / N); conjugate for the inverse direction
template <int DIR> NSB_HD cplx twid(const cplx* __restrict__ tw, int idx) {
    cplx w = tw[idx];
    if (DIR == INV) w.y = -w.y;
    return w;
}

template <int DIR> NSB_HD void dft2(cplx& a, cplx& b) {
    cplx t = csub(a, b);
    a = cadd(a, b);
    b = t;
}

template <int DIR> NSB_HD void dft4(cplx& a0, cplx& a1, cplx& a2, cplx& a3) {
    cplx t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = rot<DIR>(csub(a1, a3));
    a0 = cadd(t0, t2);
    a1 = cadd(t1, t3);
    a2 = csub(t0, t2);
    a3 = csub(t1, t3);
}

#define NSB_SQRT1_2 0.70710678118654752440
#define NSB_COS_PI_8 0.92387953251128675613
#define NSB_SIN_PI_8 0.38268343236508977173

template <int R, int DIR> struct Dft;

template <int DIR> struct Dft<2, DIR> {
    static NSB_HD void run(cplx* v) { dft2<DIR>(v[0], v[1]); }
};
template <int DIR> struct Dft<4, DIR> {
    static NSB_HD void run(cplx* v) { dft4<DIR>(v[0], v[1], v[2], v[3]); }
};
template <int DIR> struct Dft<8, DIR> {
    static NSB_HD void run(cplx* v) {
        // even / odd 4-point transforms
        dft4<DIR>(v[0], v[2], v[4], v[6]);
        dft4<DIR>(v[1], v[3], v[5], v[7]);
        // odd outputs O[k] live in v[1], v[3], v[5], v[7]; apply W8^k
        cplx o1 = cmul_cs<DIR>(v[3], NSB_SQRT1_2, NSB_SQRT1_2);
        cplx o2 = rot<DIR>(v[5]);
        cplx o3 = cmul_cs<DIR>(v[7], -NSB_SQRT1_2, NSB_SQRT1_2);
        cplx e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
        v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
        v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
        v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
        v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
    }
};
template <int DIR> struct Dft<16, DIR> {
    static NSB_HD void run(cplx* v) {
        // E_r[k] = DFT4 over v[r + 4m]; result k stored at v[r + 4k]
        dft4<DIR>(v[0], v[4], v[8], v[12]);
        dft4<DIR>(v[1], v[5], v[9], v[13]);
        dft4<DIR>(v[2], v[6], v[10], v[14]);
        dft4<DIR>(v[3], v[7], v[11], v[15]);
        // k = 0: no twiddles
        cplx a0 = v[0], a1 = v[1], a2 = v[2], a3 = v[3];
        dft4<DIR>(a0, a1, a2, a3);
        // k = 1: W16^1, W16^2, W16^3
        cplx b0 = v[4];
        cplx b1 = cmul_cs<DIR>(v[5], NSB_COS_PI_8, NSB_SIN_PI_8);
        cplx b2 = cmul_cs<DIR>(v[6], NSB_SQRT1_2, NSB_SQRT1_2);
        cplx b3 = cmul_cs<DIR>(v[7], NSB_SIN_PI_8, NSB_COS_PI_8);
        dft4<DIR>(b0, b1, b2, b3);
        // k = 2: W16^2, W16^4, W16^6
        cplx c0 = v[8];
        cplx c1 = cmul_cs<DIR>(v[9], NSB_SQRT1_2, NSB_SQRT1_2);
        cplx c2 = rot<DIR>(v[10]);
        cplx c3 = cmul_cs<DIR>(v[11], -NSB_SQRT1_2, NSB_SQRT1_2);
        dft4<DIR>(c0, c1, c2, c3);
        // k = 3: W16^3, W16^6, W16^9
        cplx d0 = v[12];
        cplx d1 = cmul_cs<DIR>(v[13], NSB_SIN_PI_8, NSB_COS_PI_8);
        cplx d2 = cmul_cs<DIR>(v[14], -NSB_SQRT1_2, NSB_SQRT1_2);
        cplx d3 = cmul_cs<DIR>(v[15], -NSB_COS_PI_8, -NSB_SIN_PI_8);
        dft4<DIR>(d0, d1, d2, d3);
        // X[k + 4m] = m-th output of group k
        v[0] = a0; v[4] = a1; v[8] = a2;  v[12] = a3;
        v[1] = b0; v[5] = b1; v[9] = b2;  v[13] = b3;
        v[2] = c0; v[6] = c1; v[10] = c2; v[14] = c3;
        v[3] = d0; v[7] = d1; v[11] = d2; v[15] = d3;
    }
};

// ---------------------------------------------------------------------------------------------
// Plans: N = R1 * R2 * R3 (R3 == 1 for two-pass plans)
// ---------------------------------------------------------------------------------------------
template <int N_, int R1_, int R2_, int R3_, int PAD_ = 1, int R4_ = 1> struct FftPlan {
    static constexpr int N = N_, R1 = R1_, R2 = R2_, R3 = R3_, R4 = R4_;
    static constexpr int PASSES = (R4_ > 1) ? 4 : ((R3_ > 1) ? 3 : 2);
    static constexpr int M1 = N_ / R1_;                  // row length after pass 1
    static constexpr int RL = (R4_ > 1) ? R4_ : ((R3_ > 1) ? R3_ : R2_);   // radix of the last pass
    static constexpr int M2 = N_ / (R1_ * R2_);          // sub-row length after pass 2
    static constexpr int NB3 = N_ / R3_;                 // butterflies of pass 3 (4-pass plans)
    static constexpr int NB1 = N_ / R1_;                 // butterflies per pass
    static constexpr int NB2 = N_ / R2_;
    static constexpr int NBL = N_ / RL;
    static constexpr int ROW = N_ / R1_ + PAD_;          // shared-memory row pitch (PAD_ = 0 for tile-interleaved buffers,
                                                         // which are conflict free without padding and can be filled in place)
    static constexpr int NPAD = R1_ * ROW;               // shared-memory elements per transform
    static_assert(R1_ * R2_ * R3_ * R4_ == N_, "bad plan");
};

// throughput plans (strided x / y passes): large radices, few passes
template <int N> struct BigPlan;
template <> struct BigPlan<16> { typedef FftPlan<16, 4, 4, 1, 0> type; };
template <> struct BigPlan<32> { typedef FftPlan<32, 4, 8, 1, 0> type; };
template <> struct BigPlan<64> { typedef FftPlan<64, 8, 8, 1, 0> type; };
template <> struct BigPlan<128> { typedef FftPlan<128, 8, 16, 1, 0> type; };
template <> struct BigPlan<256> { typedef FftPlan<256, 16, 16, 1, 0> type; };
template <> struct BigPlan<512> { typedef FftPlan<512, 8, 8, 8, 0> type; };
#ifndef NSB_BIG1024_4PASS
template <> struct BigPlan<1024> { typedef FftPlan<1024, 8, 8, 16, 0> type; };
#else   // measured slower (1024 threads per tile at 64 registers): y pass 48.6 vs 45.4 ms per step at 1024^3
template <> struct BigPlan<1024> { typedef FftPlan<1024, 8, 4, 4, 0, 8> type; };
#endif

// z-pencil plans: R1 is the smallest radix so pass 1 has exactly one butterfly per thread with
// TP = N/R1 threads per transform (later passes use a subset of the threads)
template <int N> struct ZPlan;
template <> struct ZPlan<16> { typedef FftPlan<16, 4, 4, 1> type; };
template <> struct ZPlan<32> { typedef FftPlan<32, 4, 8, 1> type; };
template <> struct ZPlan<64> { typedef FftPlan<64, 4, 4, 4> type; };
template <> struct ZPlan<128> { typedef FftPlan<128, 4, 4, 8> type; };
template <> struct ZPlan<256> { typedef FftPlan<256, 4, 8, 8> type; };
template <> struct ZPlan<512> { typedef FftPlan<512, 8, 8, 8> type; };
template <> struct ZPlan<1024> { typedef FftPlan<1024, 8, 8, 16> type; };

// fused z kernel plans: balanced (R1 == R_last), see k_z_fused
template <int N> struct ZFPlan;
template <> struct ZFPlan<16> { typedef FftPlan<16, 4, 4, 1> type; };
template <> struct ZFPlan<32> { typedef FftPlan<32, 4, 2, 4> type; };
template <> struct ZFPlan<64> { typedef FftPlan<64, 8, 8, 1> type; };
template <> struct ZFPlan<128> { typedef FftPlan<128, 4, 8, 4> type; };
template <> struct ZFPlan<256> { typedef FftPlan<256, 8, 4, 8> type; };
template <> struct ZFPlan<512> { typedef FftPlan<512, 8, 8, 8> type; };
template <> struct ZFPlan<1024> { typedef FftPlan<1024, 8, 4, 4, 1, 8> type; };   // 4 passes, radix <= 8 (the radix-16 plan needed 168 registers)

// plans of the stand-alone warp-per-pair z passes (k_z_c2r_w / k_z_r2c_w): 8 x R2 x 8, M1 = N/8 butterflies per pass
template <int N> struct ZWPlan { typedef void type; };
template <> struct ZWPlan<128> { typedef FftPlan<128, 8, 2, 8> type; };
template <> struct ZWPlan<256> { typedef FftPlan<256, 8, 4, 8> type; };
template <> struct ZWPlan<512> { typedef FftPlan<512, 8, 8, 8> type; };

template <class P> NSB_HD int padi(int i) { return (i / P::M1) * P::ROW + i % P::M1; }

// ---------------------------------------------------------------------------------------------
// Passes.  `sm` points at element 0 of this transform's shared-memory buffer, STRIDE is the distance
// (in cplx) between consecutive elements (T for tile-interleaved pencils, 1 for a private buffer).
// ---------------------------------------------------------------------------------------------
template <class P, int DIR, int STRIDE, class Ld>
NSB_HD void fft_pass1(int b, cplx* sm, const cplx* __restrict__ tw, Ld ld) {
    cplx v[P::R1];
#pragma unroll
    for (int j = 0; j < P::R1; ++j) v[j] = ld(b + j * P::M1);
    Dft<P::R1, DIR>::run(v);
#pragma unroll
    for (int k1 = 1; k1 < P::R1; ++k1) v[k1] = cmul(v[k1], twid<DIR>(tw, b * k1));
#pragma unroll
    for (int k1 = 0; k1 < P::R1; ++k1) sm[(k1 * P::ROW + b) * STRIDE] = v[k1];
}

// pass 1 on a transform already sitting in shared memory in natural order (unpadded plans): in place, no barrier
template <class P, int DIR, int STRIDE>
NSB_HD void fft_pass1_inplace(int b, cplx* sm, const cplx* __restrict__ tw) {
    static_assert(P::ROW == P::M1, "in-place pass 1 needs an unpadded plan");
    cplx v[P::R1];
#pragma unroll
    for (int j = 0; j < P::R1; ++j) v[j] = sm[(b + j * P::M1) * STRIDE];
    Dft<P::R1, DIR>::run(v);
#pragma unroll
    for (int k1 = 1; k1 < P::R1; ++k1) v[k1] = cmul(v[k1], twid<DIR>(tw, b * k1));
#pragma unroll
    for (int k1 = 0; k1 < P::R1; ++k1) sm[(k1 * P::ROW + b) * STRIDE] = v[k1];
}

// pass 1 on values already in registers (used when the loader needs a barrier before the scatter)
template <class P, int DIR> NSB_HD void fft_pass1_regs(int b, cplx* v, const cplx* __restrict__ tw) {
    Dft<P::R1, DIR>::run(v);
#pragma unroll
    for (int k1 = 1; k1 < P::R1; ++k1) v[k1] = cmul(v[k1], twid<DIR>(tw, b * k1));
}
template <class P, int STRIDE> NSB_HD void fft_pass1_scatter(int b, cplx* sm, const cplx* v) {
#pragma unroll
    for (int k1 = 0; k1 < P::R1; ++k1) sm[(k1 * P::ROW + b) * STRIDE] = v[k1];
}

template <class P, int DIR, int STRIDE>
NSB_HD void fft_pass2(int b, cplx* sm, const cplx* __restrict__ tw) {
    static_assert(P::PASSES >= 3, "pass2 only exists in 3- and 4-pass plans");
    const int k1 = b / P::M2, m2 = b % P::M2;
    const int base = k1 * P::ROW + m2;
    cplx v[P::R2];
#pragma unroll
    for (int j = 0; j < P::R2; ++j) v[j] = sm[(base + j * P::M2) * STRIDE];
    Dft<P::R2, DIR>::run(v);
#pragma unroll
    for (int kp = 1; kp < P::R2; ++kp) v[kp] = cmul(v[kp], twid<DIR>(tw, P::R1 * m2 * kp));
#pragma unroll
    for (int kp = 0; kp < P::R2; ++kp) sm[(base + kp * P::M2) * STRIDE] = v[kp];
}

// pass 3 of 4-pass plans: in-place R3-point DFTs inside each sub-row (k1, k1') of length M2 = R3*R4, twiddle
// W_{M2}^{m3 k''} = W_N^{R1 R2 m3 k''}
template <class P, int DIR, int STRIDE>
NSB_HD void fft_pass3(int b, cplx* sm, const cplx* __restrict__ tw) {
    static_assert(P::PASSES == 4, "pass3 only exists in 4-pass plans");
    const int m3 = b % P::R4, sub = b / P::R4;          // sub = k1 * R2 + k1'
    const int k1 = sub / P::R2, k1p = sub % P::R2;
    const int base = k1 * P::ROW + k1p * P::M2 + m3;
    cplx v[P::R3];
#pragma unroll
    for (int j = 0; j < P::R3; ++j) v[j] = sm[(base + j * P::R4) * STRIDE];
    Dft<P::R3, DIR>::run(v);
#pragma unroll
    for (int kq = 1; kq < P::R3; ++kq) v[kq] = cmul(v[kq], twid<DIR>(tw, P::R1 * P::R2 * m3 * kq));
#pragma unroll
    for (int kq = 0; kq < P::R3; ++kq) sm[(base + kq * P::R4) * STRIDE] = v[kq];
}

// ---- mid passes with ONE base twiddle per butterfly in registers; the higher powers are formed by multiplication
// (radix <= 4 mid passes of the 4-pass plans: two extra complex multiplies instead of per-lane table fetches)
template <int R, int DIR> NSB_HD void twiddle_powers(cplx* v, cplx w1) {
    cplx w = w1;
#pragma unroll
    for (int k = 1; k < R; ++k) {
        v[k] = cmul_dir<DIR>(v[k], w);
        if (k + 1 < R) w = cmul(w, w1);
    }
}
template <class P, int DIR, int STRIDE>
NSB_HD void fft_pass2_w1(int b, cplx* sm, cplx w1) {      // w1 = tw[R1 * (b % M2)]
    const int k1 = b / P::M2, m2 = b % P::M2;
    const int base = k1 * P::ROW + m2;
    cplx v[P::R2];
#pragma unroll
    for (int j = 0; j < P::R2; ++j) v[j] = sm[(base + j * P::M2) * STRIDE];
    Dft<P::R2, DIR>::run(v);
    twiddle_powers<P::R2, DIR>(v, w1);
#pragma unroll
    for (int kp = 0; kp < P::R2; ++kp) sm[(base + kp * P::M2) * STRIDE] = v[kp];
}
template <class P, int DIR, int STRIDE>
NSB_HD void fft_pass3_w1(int b, cplx* sm, cplx w1) {      // w1 = tw[R1 * R2 * (b % R4)]
    const int m3 = b % P::R4, sub = b / P::R4;
    const int k1 = sub / P::R2, k1p = sub % P::R2;
    const int base = k1 * P::ROW + k1p * P::M2 + m3;
    cplx v[P::R3];
#pragma unroll
    for (int j = 0; j < P::R3; ++j) v[j] = sm[(base + j * P::R4) * STRIDE];
    Dft<P::R3, DIR>::run(v);
    twiddle_powers<P::R3, DIR>(v, w1);
#pragma unroll
    for (int kq = 0; kq < P::R3; ++kq) sm[(base + kq * P::R4) * STRIDE] = v[kq];
}

// ---- passes with the (loop invariant) twiddles of this thread held in registers
// pass 1 twiddles of butterfly b: W_N^{b k1}, k1 = 1..R1-1  (stored forward; conjugated on use for INV)
template <class P> NSB_HD void load_tw_pass1(int b, const cplx* __restrict__ tw, cplx* w) {
#pragma unroll
    for (int k1 = 1; k1 < P::R1; ++k1) w[k1 - 1] = tw[b * k1];
}
// pass 2 twiddles of butterfly b: W_N^{R1 (b % R3) kp}, kp = 1..R2-1
template <class P> NSB_HD void load_tw_pass2(int b, const cplx* __restrict__ tw, cplx* w) {
#pragma unroll
    for (int kp = 1; kp < P::R2; ++kp) w[kp - 1] = tw[P::R1 * (b % P::M2) * kp];
}
template <class P, int DIR, int STRIDE, class Ld>
NSB_HD void fft_pass1_rw(int b, cplx* sm, const cplx* w, Ld ld) {
    cplx v[P::R1];
#pragma unroll
    for (int j = 0; j < P::R1; ++j) v[j] = ld(b + j * P::M1);
    Dft<P::R1, DIR>::run(v);
#pragma unroll
    for (int k1 = 1; k1 < P::R1; ++k1) v[k1] = cmul_dir<DIR>(v[k1], w[k1 - 1]);
#pragma unroll
    for (int k1 = 0; k1 < P::R1; ++k1) sm[(k1 * P::ROW + b) * STRIDE] = v[k1];
}
template <class P, int DIR> NSB_HD void fft_pass1_regs_rw(cplx* v, const cplx* w) {
    Dft<P::R1, DIR>::run(v);
#pragma unroll
    for (int k1 = 1; k1 < P::R1; ++k1) v[k1] = cmul_dir<DIR>(v[k1], w[k1 - 1]);
}
template <class P, int DIR, int STRIDE>
NSB_HD void fft_pass2_rw(int b, cplx* sm, const cplx* w) {
    static_assert(P::PASSES >= 3, "pass2 only exists in 3- and 4-pass plans");
    const int k1 = b / P::M2, m2 = b % P::M2;
    const int base = k1 * P::ROW + m2;
    cplx v[P::R2];
#pragma unroll
    for (int j = 0; j < P::R2; ++j) v[j] = sm[(base + j * P::M2) * STRIDE];
    Dft<P::R2, DIR>::run(v);
#pragma unroll
    for (int kp = 1; kp < P::R2; ++kp) v[kp] = cmul_dir<DIR>(v[kp], w[kp - 1]);
#pragma unroll
    for (int kp = 0; kp < P::R2; ++kp) sm[(base + kp * P::M2) * STRIDE] = v[kp];
}

// radix-8 mid pass with the twiddle powers formed on the fly from W^1, W^2, W^4 of the butterfly (12 registers
// instead of 28): w3 = w1 w2, w5 = w1 w4, w6 = w2 w4, w7 = w3 w4
template <class P, int DIR, int STRIDE>
NSB_HD void fft_pass2_r8_base(int b, cplx* sm, cplx wa, cplx wb, cplx wc) {
    static_assert(P::PASSES >= 3 && P::R2 == 8, "radix-8 pass 2");
    const int k1 = b / P::M2, m2 = b % P::M2;
    const int base = k1 * P::ROW + m2;
    cplx v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = sm[(base + j * P::M2) * STRIDE];
    Dft<8, DIR>::run(v);
    v[1] = cmul_dir<DIR>(v[1], wa);
    v[2] = cmul_dir<DIR>(v[2], wb);
    v[4] = cmul_dir<DIR>(v[4], wc);
    const cplx w3 = cmul(wa, wb);
    v[3] = cmul_dir<DIR>(v[3], w3);
    v[7] = cmul_dir<DIR>(v[7], cmul(w3, wc));
    v[5] = cmul_dir<DIR>(v[5], cmul(wa, wc));
    v[6] = cmul_dir<DIR>(v[6], cmul(wb, wc));
#pragma unroll
    for (int kp = 0; kp < 8; ++kp) sm[(base + kp * P::M2) * STRIDE] = v[kp];
}

// v[k] *= W^k (DIR) for k = 1..7 with W^3, W^5, W^6, W^7 formed on the fly from W^1, W^2, W^4
template <int DIR> NSB_HD void twiddle8_base(cplx* v, cplx wa, cplx wb, cplx wc) {
    v[1] = cmul_dir<DIR>(v[1], wa);
    v[2] = cmul_dir<DIR>(v[2], wb);
    v[4] = cmul_dir<DIR>(v[4], wc);
    const cplx w3 = cmul(wa, wb);
    v[3] = cmul_dir<DIR>(v[3], w3);
    v[7] = cmul_dir<DIR>(v[7], cmul(w3, wc));
    v[5] = cmul_dir<DIR>(v[5], cmul(wa, wc));
    v[6] = cmul_dir<DIR>(v[6], cmul(wb, wc));
}

// radix-2 mid pass
template <class P, int DIR, int STRIDE>
NSB_HD void fft_pass2_r2_base(int b, cplx* sm, cplx wa) {
    static_assert(P::PASSES >= 3 && P::R2 == 2, "radix-2 pass 2");
    const int k1 = b / P::M2, m2 = b % P::M2;
    const int base = k1 * P::ROW + m2;
    cplx v[2];
    v[0] = sm[base * STRIDE];
    v[1] = sm[(base + P::M2) * STRIDE];
    Dft<2, DIR>::run(v);
    v[1] = cmul_dir<DIR>(v[1], wa);
    sm[base * STRIDE] = v[0];
    sm[(base + P::M2) * STRIDE] = v[1];
}

// radix-4 mid pass, twiddles W^{R1 m2 kp} from the powers 1 and 2
template <class P, int DIR, int STRIDE>
NSB_HD void fft_pass2_r4_base(int b, cplx* sm, cplx wa, cplx wb) {
    static_assert(P::PASSES >= 3 && P::R2 == 4, "radix-4 pass 2");
    const int k1 = b / P::M2, m2 = b % P::M2;
    const int base = k1 * P::ROW + m2;
    cplx v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = sm[(base + j * P::M2) * STRIDE];
    Dft<4, DIR>::run(v);
    v[1] = cmul_dir<DIR>(v[1], wa);
    v[2] = cmul_dir<DIR>(v[2], wb);
    v[3] = cmul_dir<DIR>(v[3], cmul(wa, wb));
#pragma unroll
    for (int kp = 0; kp < 4; ++kp) sm[(base + kp * P::M2) * STRIDE] = v[kp];
}

// radix-16 mid pass with the fifteen twiddle powers formed on the fly from W^1, W^2, W^4, W^8 (16 registers instead of 60)
template <int DIR> NSB_HD void twiddle16_base(cplx* v, cplx w1, cplx w2, cplx w4, cplx w8) {
    const cplx w3 = cmul(w1, w2), w5 = cmul(w1, w4), w6 = cmul(w2, w4);
    const cplx w7 = cmul(w3, w4);
    v[1] = cmul_dir<DIR>(v[1], w1);   v[2] = cmul_dir<DIR>(v[2], w2);   v[3] = cmul_dir<DIR>(v[3], w3);
    v[4] = cmul_dir<DIR>(v[4], w4);   v[5] = cmul_dir<DIR>(v[5], w5);   v[6] = cmul_dir<DIR>(v[6], w6);
    v[7] = cmul_dir<DIR>(v[7], w7);   v[8] = cmul_dir<DIR>(v[8], w8);
    v[9] = cmul_dir<DIR>(v[9], cmul(w1, w8));    v[10] = cmul_dir<DIR>(v[10], cmul(w2, w8));
    v[11] = cmul_dir<DIR>(v[11], cmul(w3, w8));  v[12] = cmul_dir<DIR>(v[12], cmul(w4, w8));
    v[13] = cmul_dir<DIR>(v[13], cmul(w5, w8));  v[14] = cmul_dir<DIR>(v[14], cmul(w6, w8));
    v[15] = cmul_dir<DIR>(v[15], cmul(w7, w8));
}
template <class P, int DIR, int STRIDE>
NSB_HD void fft_pass2_r16_base(int b, cplx* sm, cplx w1, cplx w2, cplx w4, cplx w8) {
    static_assert(P::PASSES >= 3 && P::R2 == 16, "radix-16 pass 2");
    const int k1 = b / P::M2, m2 = b % P::M2;
    const int base = k1 * P::ROW + m2;
    cplx v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = sm[(base + j * P::M2) * STRIDE];
    Dft<16, DIR>::run(v);
    twiddle16_base<DIR>(v, w1, w2, w4, w8);
#pragma unroll
    for (int kp = 0; kp < 16; ++kp) sm[(base + kp * P::M2) * STRIDE] = v[kp];
}

// start of the RL contiguous elements the last pass of butterfly b works on (b = k1 + R1*k1' [+ R1*R2*k1''])
template <class P> NSB_HD int fft_row_base(int b) {
    if constexpr (P::PASSES == 4) {
        const int k1 = b % P::R1, r = b / P::R1;
        return k1 * P::ROW + (r % P::R2) * P::M2 + (r / P::R2) * P::RL;
    } else {
        return (b % P::R1) * P::ROW + (b / P::R1) * P::RL;
    }
}

// last pass: v[k2] is output element  b + k2 * P::NBL
template <class P, int DIR, int STRIDE>
NSB_HD void fft_pass_last(int b, const cplx* sm, cplx* v) {
    const int base = fft_row_base<P>(b);
#pragma unroll
    for (int m = 0; m < P::RL; ++m) v[m] = sm[(base + m) * STRIDE];
    Dft<P::RL, DIR>::run(v);
}

// ---------------------------------------------------------------------------------------------
// Two real transforms packed into one complex transform (z = a + i b).
// ---------------------------------------------------------------------------------------------
// c2r: element n of the full Hermitian spectrum built from the two half spectra A, B (k = 0..N/2).
// Like FFTW's c2r the imaginary parts of the k = 0 and k = N/2 entries are ignored.
template <int N> NSB_HD cplx pack_hermitian(int n, cplx A, cplx B) {
    // A, B are the half-spectrum entries at k = (n <= N/2 ? n : N - n)
    if (n == 0 || n == N / 2) { A.y = 0.0; B.y = 0.0; }
    if (n > N / 2) { A.y = -A.y; B.y = -B.y; }
    return mk(A.x - B.y, A.y + B.x);
}
// r2c: half-spectrum entries of the two real inputs from Z(k) and Z(N-k)
NSB_HD void unpack_pair(cplx Zk, cplx Zm, cplx& A, cplx& B) {
    A = mk(0.5 * (Zk.x + Zm.x), 0.5 * (Zk.y - Zm.y));
    B = mk(0.5 * (Zk.y + Zm.y), -0.5 * (Zk.x - Zm.x));
}
