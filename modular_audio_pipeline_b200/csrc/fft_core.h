// FFT cores shared by the STFT-1024 (denoise) and STFT-400 (Whisper) kernels.
//
// Every routine is a per-thread step function: the kernels call the steps with a __syncthreads() between them,
// tests/host_fft_check.cpp calls them in a loop over "thread ids" on the CPU to verify the index math against a
// direct DFT without a GPU.  FP32 throughout (FP32 STFT/iSTFT gives 136 dB vs float64).
//
//   fft512_*  : 512-point complex FFT as 8 x 8 x 8 (three radix-8 register butterflies, two shared-memory
//               exchanges), 64 threads per transform, 8 points per thread.
//   rfft1024  : 1024-point real FFT = 512-point complex FFT of (even + i odd) + split post-pass; and its inverse.
//   fft200_*  : 200-point complex FFT as 8 x 25 (radix-8, then 5 x 5 in registers) for the 400-point real FFT
//               Whisper needs (400 = 2^4 5^2: radix-2/4 alone cannot produce it).
#pragma once

#if defined(__CUDACC__)
#define MAF_HD __host__ __device__ __forceinline__
#else
#define MAF_HD inline
#endif

namespace maf {

struct c32 { float x, y; };

MAF_HD c32 mk(float x, float y) { c32 r; r.x = x; r.y = y; return r; }
MAF_HD c32 cadd(c32 a, c32 b) { return mk(a.x + b.x, a.y + b.y); }
MAF_HD c32 csub(c32 a, c32 b) { return mk(a.x - b.x, a.y - b.y); }
MAF_HD c32 cmul(c32 a, c32 b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
MAF_HD c32 cconj(c32 a) { return mk(a.x, -a.y); }
MAF_HD c32 mul_mi(c32 a) { return mk(a.y, -a.x); }      // a * (-i)
MAF_HD c32 mul_pi(c32 a) { return mk(-a.y, a.x); }      // a * (+i)

// 4-point forward DFT, natural order in and out.
MAF_HD void fft4(c32& c0, c32& c1, c32& c2, c32& c3) {
    c32 t0 = cadd(c0, c2), t1 = csub(c0, c2), t2 = cadd(c1, c3), t3 = mul_mi(csub(c1, c3));
    c0 = cadd(t0, t2); c2 = csub(t0, t2); c1 = cadd(t1, t3); c3 = csub(t1, t3);
}

// 8-point forward DFT (kernel e^{-2 pi i nk/8}), natural order in and out.
MAF_HD void fft8(c32 (&v)[8]) {
    const float h = 0.70710678118654752440f;
    c32 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
    c32 b0 = csub(v[0], v[4]), b1 = csub(v[1], v[5]), b2 = csub(v[2], v[6]), b3 = csub(v[3], v[7]);
    b1 = mk((b1.x + b1.y) * h, (b1.y - b1.x) * h);          // * W8^1
    b2 = mul_mi(b2);                                        // * W8^2
    b3 = mk((b3.y - b3.x) * h, -(b3.x + b3.y) * h);         // * W8^3
    fft4(a0, a1, a2, a3);
    fft4(b0, b1, b2, b3);
    v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
    v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

// 5-point forward DFT, natural order in and out.
MAF_HD void fft5(c32& v0, c32& v1, c32& v2, c32& v3, c32& v4) {
    const float c1 = 0.30901699437494742410f;   // cos(2 pi/5)
    const float c2 = -0.80901699437494742410f;  // cos(4 pi/5)
    const float s1 = 0.95105651629515357212f;   // sin(2 pi/5)
    const float s2 = 0.58778525229247312917f;   // sin(4 pi/5)
    c32 p14 = cadd(v1, v4), m14 = csub(v1, v4), p23 = cadd(v2, v3), m23 = csub(v2, v3);
    c32 r0 = mk(v0.x + p14.x + p23.x, v0.y + p14.y + p23.y);
    c32 ra = mk(v0.x + c1 * p14.x + c2 * p23.x, v0.y + c1 * p14.y + c2 * p23.y);
    c32 rb = mk(v0.x + c2 * p14.x + c1 * p23.x, v0.y + c2 * p14.y + c1 * p23.y);
    // -i * (s1 m14 + s2 m23) and -i * (s2 m14 - s1 m23)
    c32 ia = mk(s1 * m14.y + s2 * m23.y, -(s1 * m14.x + s2 * m23.x));
    c32 ib = mk(s2 * m14.y - s1 * m23.y, -(s2 * m14.x - s1 * m23.x));
    v0 = r0; v1 = cadd(ra, ia); v4 = csub(ra, ia); v2 = cadd(rb, ib); v3 = csub(rb, ib);
}

// ------------------------------------------------------------------------------------------------------------
// 512-point complex FFT, 64 threads, thread t holds 8 points.  Shared scratch: FFT512_SMEM c32 per transform.
//   in : v[m] = x[t + 64 m]
//   out: v[d] = X[t + 64 d]  (fft512_out_base(t) == t)
// Twiddle tables are laid out for conflict-free shared-memory reads (a stride-p walk through one linear table puts
// 2-8 lanes on the same bank):  tw1[p * 64 + t] = e^{-2 pi i t p / 512}  (step 1, lanes walk t)
//                                tw2[c * 8 + a]  = e^{-2 pi i 8 a c / 512} (step 2, lanes walk a; p is a broadcast)
// ------------------------------------------------------------------------------------------------------------
constexpr int FFT512_SMEM = 8 * 72;

MAF_HD int fft512_out_base(int t) { return t; }

// Exchange-2 layout.  Element (p, a, c) lives at ((p + 2a + 8c) & 15) + 16 a + 128 (c >> 1): a bijection onto 0..511
// whose low 4 bits (the 8-byte bank within a half-warp) are distinct both for the step-2 writers
// (half-warp = 2 p's x 8 a's, fixed c) and for the step-3 readers (half-warp = 8 p's x 2 c's, fixed a).
MAF_HD int fft512_x2(int p, int a, int c) { return ((p + 2 * a + 8 * c) & 15) + 16 * a + 128 * (c >> 1); }

constexpr int FFT512_TW1 = 8 * 64;
constexpr int FFT512_TW2 = 64;

MAF_HD void fft512_step1(c32 (&v)[8], c32* sm, const c32* tw1, int t) {
    fft8(v);
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        c32 w = tw1[p * 64 + t];
        sm[p * 72 + t] = (p == 0) ? v[0] : cmul(v[p], w);
    }
}
// Same step with the twiddles e^{-2 pi i t p / 512} = w^p built from the per-thread base w = e^{-2 pi i t / 512} by repeated
// multiplication (p <= 7: < 5e-7 relative) instead of seven shared-memory loads -- for kernels bound by shared-memory traffic.
MAF_HD void fft512_step1_base(c32 (&v)[8], c32* sm, c32 w1, int t) {
    fft8(v);
    c32 w = w1;
    sm[t] = v[0];
#pragma unroll
    for (int p = 1; p < 8; ++p) {
        sm[p * 72 + t] = cmul(v[p], w);
        if (p < 7) w = cmul(w, w1);
    }
}
// (barrier)
MAF_HD void fft512_step2_load(c32 (&v)[8], const c32* sm, int t) {
    const int p = t >> 3, a = t & 7;
#pragma unroll
    for (int b = 0; b < 8; ++b) v[b] = sm[p * 72 + a + 8 * b];
}
// (barrier)
MAF_HD void fft512_step2_store(c32 (&v)[8], c32* sm, const c32* tw2, int t) {
    const int p = t >> 3, a = t & 7;
    fft8(v);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        c32 w = tw2[c * 8 + a];
        sm[fft512_x2(p, a, c)] = (c == 0) ? v[0] : cmul(v[c], w);
    }
}
// twiddles e^{-2 pi i 8 a c / 512} = w^c from the per-thread base w = e^{-2 pi i a / 64}, a = t & 7
MAF_HD void fft512_step2_store_base(c32 (&v)[8], c32* sm, c32 w2, int t) {
    const int p = t >> 3, a = t & 7;
    fft8(v);
    c32 w = w2;
    sm[fft512_x2(p, a, 0)] = v[0];
#pragma unroll
    for (int c = 1; c < 8; ++c) {
        sm[fft512_x2(p, a, c)] = cmul(v[c], w);
        if (c < 7) w = cmul(w, w2);
    }
}
// (barrier)
// step 3: thread t handles (p = t & 7, c = t >> 3); its outputs are X[t + 64 d], d = 0..7 (natural order)
MAF_HD void fft512_step3(c32 (&v)[8], const c32* sm, int t) {
    const int p = t & 7, c = t >> 3;
#pragma unroll
    for (int a = 0; a < 8; ++a) v[a] = sm[fft512_x2(p, a, c)];
    fft8(v);
}

// ------------------------------------------------------------------------------------------------------------
// Real 1024-point FFT split pass.  Z = FFT512(x[2n] + i x[2n+1]); for one pair index k in 1..255 returns
// X[k] and X[512-k].  tw1024k = e^{-2 pi i k / 1024}.
// ------------------------------------------------------------------------------------------------------------
MAF_HD void rfft_split_pair(c32 zk, c32 znk, c32 tw1024k, c32& xk, c32& xnk) {
    // E = (Z[k] + conj Z[512-k]) / 2 ; O = -i (Z[k] - conj Z[512-k]) / 2 ; X[k] = E + W^k O ; X[512-k] = conj(E - W^k O)
    c32 e = mk(0.5f * (zk.x + znk.x), 0.5f * (zk.y - znk.y));
    c32 d = mk(0.5f * (zk.x - znk.x), 0.5f * (zk.y + znk.y));
    c32 o = mul_mi(d);
    c32 wo = cmul(tw1024k, o);
    xk = cadd(e, wo);
    xnk = cconj(csub(e, wo));
}

// Inverse: from Hermitian X[k], X[512-k] (k in 1..255) build 2*conj(Z[k]) and 2*conj(Z[512-k]) so that a FORWARD
// FFT512 of the result F gives x[2n] = Re F[n] / 2, x[2n+1] = -Im F[n] / 2 (unnormalised inverse; the 1/2 and the
// synthesis window are folded into the kernel's window table).
MAF_HD void irfft_merge_pair(c32 xk, c32 xnk, c32 tw1024k, c32& zck, c32& zcnk) {
    // 2E = X[k] + conj X[512-k] ; 2O = (X[k] - conj X[512-k]) * conj(W^k) ; Z[k] = E + iO ; Z[512-k] = conj(E) + i conj(O)
    c32 e2 = mk(xk.x + xnk.x, xk.y - xnk.y);
    c32 d2 = mk(xk.x - xnk.x, xk.y + xnk.y);
    c32 o2 = cmul(d2, cconj(tw1024k));
    c32 zk = cadd(e2, mul_pi(o2));
    c32 znk = cadd(cconj(e2), mul_pi(cconj(o2)));
    zck = cconj(zk);
    zcnk = cconj(znk);
}

// ------------------------------------------------------------------------------------------------------------
// 1024-point complex FFT by ONE WARP as 32 x 32: every lane runs two 32-point FFTs in registers with a single
// shared-memory transpose in between (one store + one load of the data per transform instead of the three round
// trips of the 8 x 8 x 8 scheme).  The STFT kernels feed it TWO real frames at once (frame A in the real part,
// frame B in the imaginary part); the split X_A[k] = (Z[k] + conj Z[1024-k]) / 2, X_B[k] = (Z[k] - conj Z[1024-k]) / 2i
// pairs lane l with lane 32 - l and is done with warp shuffles.
//   in : v[j] = x[lane + 32 j]        out: v[j] = X[32 j + lane]
// Scratch: FFT1024_SMEM c32 per warp, pitch 33 (both the row-wise stores and the column-wise 64-bit loads are
// conflict-free).  Twiddles: tww[j * 32 + lane] = e^{-2 pi i lane j / 1024}.
// ------------------------------------------------------------------------------------------------------------
// e^{-2 pi i e / 32}; after unrolling every index is a compile-time constant, so these fold into immediates
MAF_HD c32 w32(int e) {
    const float re[32] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654757f, 0.55557023301960229f, 0.38268343236508984f, 0.19509032201612833f,
                          0.0f, -0.19509032201612819f, -0.38268343236508973f, -0.55557023301960196f, -0.70710678118654746f, -0.83146961230254535f, -0.92387953251128674f, -0.98078528040323043f,
                          -1.0f, -0.98078528040323043f, -0.92387953251128685f, -0.83146961230254546f, -0.70710678118654768f, -0.55557023301960218f, -0.38268343236509034f, -0.19509032201612866f,
                          0.0f, 0.1950903220161283f, 0.38268343236509f, 0.55557023301960184f, 0.70710678118654735f, 0.83146961230254524f, 0.92387953251128652f, 0.98078528040323032f};
    const float im[32] = {-0.0f, -0.19509032201612825f, -0.38268343236508978f, -0.55557023301960218f, -0.70710678118654746f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f,
                          -1.0f, -0.98078528040323043f, -0.92387953251128674f, -0.83146961230254546f, -0.70710678118654757f, -0.55557023301960218f, -0.38268343236508989f, -0.19509032201612861f,
                          -0.0f, 0.19509032201612836f, 0.38268343236508967f, 0.55557023301960196f, 0.70710678118654746f, 0.83146961230254535f, 0.92387953251128652f, 0.98078528040323032f,
                          1.0f, 0.98078528040323043f, 0.92387953251128663f, 0.83146961230254546f, 0.70710678118654768f, 0.55557023301960229f, 0.38268343236509039f, 0.19509032201612872f};
    return mk(re[e], im[e]);
}

// 32-point forward DFT in registers (4 x fft8, twiddle, 8 x fft4), natural order in and out.
// n = 4 n1 + n2 (n1 < 8, n2 < 4), k = k1 + 8 k2:  X[k1 + 8 k2] = sum_n2 W32^{n2 k1} W4^{n2 k2} sum_n1 x[4 n1 + n2] W8^{n1 k1}
MAF_HD void fft32(c32 (&v)[32]) {
    c32 a[4][8];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
        c32 t[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) t[n1] = v[4 * n1 + n2];
        fft8(t);
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) a[n2][k1] = (n2 == 0 || k1 == 0) ? t[k1] : (n2 * k1 == 8 ? mul_mi(t[k1]) : cmul(t[k1], w32(n2 * k1)));
    }
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) {
        c32 b0 = a[0][k1], b1 = a[1][k1], b2 = a[2][k1], b3 = a[3][k1];
        fft4(b0, b1, b2, b3);
        v[k1] = b0; v[k1 + 8] = b1; v[k1 + 16] = b2; v[k1 + 24] = b3;
    }
}

constexpr int FFT1024_PITCH = 33;
constexpr int FFT1024_SMEM = 32 * FFT1024_PITCH;

MAF_HD void fft1024w_pass1(c32 (&v)[32], c32* __restrict__ sm, const c32* __restrict__ tww, int lane) {
    fft32(v);                                                         // over n2 (registers) -> k2
    // the 31 twiddle loads are software-pipelined TW_AHEAD deep: with the kernel at its register cap the compiler otherwise
    // funnels all of them through one register pair (load -> multiply -> store -> next load: every shared-memory latency exposed)
    constexpr int TW_AHEAD = 4;
    c32 tq[TW_AHEAD];
#pragma unroll
    for (int a = 0; a < TW_AHEAD; ++a) tq[a] = tww[(1 + a) * 32 + lane];
    sm[lane] = v[0];
#pragma unroll
    for (int j = 1; j < 32; ++j) {
        const c32 t = tq[(j - 1) % TW_AHEAD];
        if (j + TW_AHEAD < 32) tq[(j - 1) % TW_AHEAD] = tww[(j + TW_AHEAD) * 32 + lane];
        sm[j * FFT1024_PITCH + lane] = cmul(v[j], t);
    }
}
// (warp barrier)
MAF_HD void fft1024w_pass2(c32 (&v)[32], const c32* sm, int lane) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = sm[lane * FFT1024_PITCH + j];  // lane = k2, j = n1
    fft32(v);                                                         // over n1 -> k1: v[k1] = X[32 k1 + k2]
}

// ------------------------------------------------------------------------------------------------------------
// 200-point complex FFT as 8 x 25.  n = j + 25 m (j < 25, m < 8), k = p + 8 q (p < 8, q < 25):
//   X[p + 8q] = sum_j W25^{jq} [ W200^{jp} sum_m x[j + 25m] W8^{mp} ]
// step A: work item (j): 8-point FFT over m, twiddle W200^{jp}, store sm[p * 25 + j]        (25 items / transform)
// step B: work item (p): 25-point DFT over j (5 x 5 in registers) -> X[p + 8q], q = 0..24   ( 8 items / transform)
// Shared scratch: 200 c32 per transform.
// ------------------------------------------------------------------------------------------------------------
// twA[p * 25 + j] = e^{-2 pi i j p / 200} (lanes walk j: conflict-free)
MAF_HD void fft200_stepA(c32 (&v)[8], c32* sm, const c32* twA, int j) {
    fft8(v);
#pragma unroll
    for (int p = 0; p < 8; ++p) sm[p * 25 + j] = (p == 0) ? v[0] : cmul(v[p], twA[p * 25 + j]);
}

// e^{-2 pi i e / 25}: after unrolling every index is a compile-time constant, so these fold into immediates
MAF_HD c32 w25(int e) {
    const float re[25] = {1.0f, 0.96858316112863108f, 0.87630668004386358f, 0.72896862742141155f, 0.53582679497899655f, 0.30901699437494745f, 0.062790519529313527f, -0.1873813145857246f, -0.42577929156507272f, -0.63742398974868975f, -0.80901699437494734f, -0.92977648588825135f, -0.99211470131447776f, -0.99211470131447788f, -0.92977648588825146f, -0.80901699437494778f, -0.63742398974868952f, -0.42577929156507216f, -0.18738131458572463f, 0.062790519529312833f, 0.30901699437494723f, 0.53582679497899677f, 0.72896862742141122f, 0.87630668004386314f, 0.96858316112863097f};
    const float im[25] = {0.0f, -0.24868988716485479f, -0.48175367410171532f, -0.68454710592868862f, -0.84432792550201508f, -0.95105651629515353f, -0.99802672842827156f, -0.98228725072868872f, -0.90482705246601947f, -0.77051324277578925f, -0.58778525229247325f, -0.36812455268467814f, -0.12533323356430454f, 0.12533323356430429f, 0.36812455268467792f, 0.58778525229247269f, 0.77051324277578936f, 0.9048270524660198f, 0.98228725072868872f, 0.99802672842827156f, 0.95105651629515364f, 0.84432792550201496f, 0.68454710592868895f, 0.4817536741017161f, 0.24868988716485535f};
    return mk(re[e], im[e]);
}

// 25-point DFT of sm[p*25 + j], j = 0..24, result X[p + 8 q] written to out[q] (q = 0..24).
// j = j0 + 5 j1, q = q1 + 5 q0:  W25^{jq} = W25^{j0 q1} W5^{j0 q0} W5^{j1 q1}
MAF_HD void fft200_stepB(const c32* sm, int p, c32 (&out)[25]) {
    c32 y[25];
#pragma unroll
    for (int j0 = 0; j0 < 5; ++j0) {
        c32 a0 = sm[p * 25 + j0], a1 = sm[p * 25 + j0 + 5], a2 = sm[p * 25 + j0 + 10],
            a3 = sm[p * 25 + j0 + 15], a4 = sm[p * 25 + j0 + 20];
        fft5(a0, a1, a2, a3, a4);                 // over j1 -> q1
        // twiddle W25^{j0 q1}
        y[j0 * 5 + 0] = a0;
        y[j0 * 5 + 1] = (j0 == 0) ? a1 : cmul(a1, w25(j0 * 1));
        y[j0 * 5 + 2] = (j0 == 0) ? a2 : cmul(a2, w25(j0 * 2));
        y[j0 * 5 + 3] = (j0 == 0) ? a3 : cmul(a3, w25(j0 * 3));
        y[j0 * 5 + 4] = (j0 == 0) ? a4 : cmul(a4, w25(j0 * 4));
    }
#pragma unroll
    for (int q1 = 0; q1 < 5; ++q1) {
        c32 b0 = y[0 * 5 + q1], b1 = y[1 * 5 + q1], b2 = y[2 * 5 + q1], b3 = y[3 * 5 + q1], b4 = y[4 * 5 + q1];
        fft5(b0, b1, b2, b3, b4);                 // over j0 -> q0
        out[q1 + 0] = b0; out[q1 + 5] = b1; out[q1 + 10] = b2; out[q1 + 15] = b3; out[q1 + 20] = b4;
    }
}

// Real 400-point split: Z = FFT200(x[2n] + i x[2n+1]); X[k], k = 0..200, from Z[k], Z[200-k]; tw400k = e^{-2 pi i k/400}.
MAF_HD c32 rfft400_bin(c32 zk, c32 znk, c32 tw400k) {
    c32 e = mk(0.5f * (zk.x + znk.x), 0.5f * (zk.y - znk.y));
    c32 d = mk(0.5f * (zk.x - znk.x), 0.5f * (zk.y + znk.y));
    return cadd(e, cmul(tw400k, mul_mi(d)));
}

}  // namespace maf
