// CPU emulation of the per-thread FFT step functions in csrc/fft_core.h (no GPU needed): checks the index math of
// fft512 (8x8x8), the real-1024 split / merge passes and fft200 (8x25) + real-400 split against direct DFTs.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../modular_audio_pipeline_b200/csrc/fft_core.h"
using namespace maf;
typedef std::complex<double> cd;
static double frand() { return rand() / (double)RAND_MAX * 2.0 - 1.0; }

static void fft512_emul(const std::vector<c32>& x, std::vector<c32>& X, const std::vector<c32>& tw512) {
    std::vector<c32> tw1(FFT512_TW1), tw2(FFT512_TW2);
    for (int p = 0; p < 8; ++p) for (int t = 0; t < 64; ++t) tw1[p * 64 + t] = tw512[t * p];
    for (int c = 0; c < 8; ++c) for (int a = 0; a < 8; ++a) tw2[c * 8 + a] = tw512[8 * a * c];
    std::vector<c32> sm(FFT512_SMEM);
    c32 regs[64][8];
    for (int t = 0; t < 64; ++t) { for (int m = 0; m < 8; ++m) regs[t][m] = x[t + 64 * m]; fft512_step1(regs[t], sm.data(), tw1.data(), t); }
    for (int t = 0; t < 64; ++t) fft512_step2_load(regs[t], sm.data(), t);
    for (int t = 0; t < 64; ++t) fft512_step2_store(regs[t], sm.data(), tw2.data(), t);
    for (int t = 0; t < 64; ++t) { fft512_step3(regs[t], sm.data(), t); int u = fft512_out_base(t); for (int d = 0; d < 8; ++d) X[u + 64 * d] = regs[t][d]; }
}

// fft512 with the twiddles built from per-thread bases (fft512_step1_base / fft512_step2_store_base)
static void fft512_emul_base(const std::vector<c32>& x, std::vector<c32>& X) {
    const double PI = 3.14159265358979323846;
    std::vector<c32> sm(FFT512_SMEM);
    c32 regs[64][8];
    for (int t = 0; t < 64; ++t) {
        for (int m = 0; m < 8; ++m) regs[t][m] = x[t + 64 * m];
        fft512_step1_base(regs[t], sm.data(), mk((float)cos(2 * PI * t / 512), (float)-sin(2 * PI * t / 512)), t);
    }
    for (int t = 0; t < 64; ++t) fft512_step2_load(regs[t], sm.data(), t);
    for (int t = 0; t < 64; ++t) fft512_step2_store_base(regs[t], sm.data(), mk((float)cos(2 * PI * (t & 7) / 64), (float)-sin(2 * PI * (t & 7) / 64)), t);
    for (int t = 0; t < 64; ++t) { fft512_step3(regs[t], sm.data(), t); for (int d = 0; d < 8; ++d) X[t + 64 * d] = regs[t][d]; }
}

int main() {
    const double PI = 3.14159265358979323846;
    int fails = 0;
    std::vector<c32> tw512(512), tw1024(513), tw200(200), tw400(201);
    for (int e = 0; e < 512; ++e) tw512[e] = mk((float)cos(2 * PI * e / 512), (float)-sin(2 * PI * e / 512));
    for (int e = 0; e <= 512; ++e) tw1024[e] = mk((float)cos(2 * PI * e / 1024), (float)-sin(2 * PI * e / 1024));
    for (int e = 0; e < 200; ++e) tw200[e] = mk((float)cos(2 * PI * e / 200), (float)-sin(2 * PI * e / 200));
    for (int e = 0; e <= 200; ++e) tw400[e] = mk((float)cos(2 * PI * e / 400), (float)-sin(2 * PI * e / 400));
    // ---- fft512 vs direct
    {
        std::vector<c32> x(512), X(512);
        for (auto& v : x) v = mk((float)frand(), (float)frand());
        fft512_emul(x, X, tw512);
        double err = 0, nrm = 0;
        for (int k = 0; k < 512; ++k) {
            cd s = 0; for (int n = 0; n < 512; ++n) s += cd(x[n].x, x[n].y) * std::polar(1.0, -2 * PI * n * k / 512.0);
            err = fmax(err, std::abs(s - cd(X[k].x, X[k].y))); nrm = fmax(nrm, std::abs(s));
        }
        printf("fft512   max err %.3e (max |X| %.3f)\n", err, nrm); if (err > 1e-4) ++fails;
        std::vector<c32> Xb(512);
        fft512_emul_base(x, Xb);
        double errb = 0;
        for (int k = 0; k < 512; ++k) {
            cd s = 0; for (int n = 0; n < 512; ++n) s += cd(x[n].x, x[n].y) * std::polar(1.0, -2 * PI * n * k / 512.0);
            errb = fmax(errb, std::abs(s - cd(Xb[k].x, Xb[k].y)));
        }
        printf("fft512 (twiddles from bases) max err %.3e\n", errb); if (errb > 1e-4) ++fails;
    }
    // ---- rfft1024 forward and inverse
    {
        std::vector<float> x(1024); for (auto& v : x) v = (float)frand();
        std::vector<c32> z(512), Z(512), Xr(513);
        for (int n = 0; n < 512; ++n) z[n] = mk(x[2 * n], x[2 * n + 1]);
        fft512_emul(z, Z, tw512);
        Xr[0] = mk(Z[0].x + Z[0].y, 0); Xr[512] = mk(Z[0].x - Z[0].y, 0); Xr[256] = cconj(Z[256]);
        for (int k = 1; k < 256; ++k) rfft_split_pair(Z[k], Z[512 - k], tw1024[k], Xr[k], Xr[512 - k]);
        double err = 0;
        for (int k = 0; k <= 512; ++k) {
            cd s = 0; for (int n = 0; n < 1024; ++n) s += (double)x[n] * std::polar(1.0, -2 * PI * n * k / 1024.0);
            err = fmax(err, std::abs(s - cd(Xr[k].x, Xr[k].y)));
        }
        printf("rfft1024 max err %.3e\n", err); if (err > 2e-4) ++fails;
        // inverse
        std::vector<c32> zc(512), F(512);
        zc[0] = mk(Xr[0].x + Xr[512].x, -(Xr[0].x - Xr[512].x));     // 2 conj(Z0), Z0 = (X0+X512)/2 + i (X0-X512)/2
        zc[256] = mk(2 * Xr[256].x, 2 * Xr[256].y);                  // 2 conj(Z256) = 2 X256
        for (int k = 1; k < 256; ++k) irfft_merge_pair(Xr[k], Xr[512 - k], tw1024[k], zc[k], zc[512 - k]);
        fft512_emul(zc, F, tw512);
        double ierr = 0;
        for (int n = 0; n < 512; ++n) {
            ierr = fmax(ierr, fabs(F[n].x / 1024.0 - x[2 * n])); ierr = fmax(ierr, fabs(-F[n].y / 1024.0 - x[2 * n + 1]));
        }
        printf("irfft1024 round-trip max err %.3e\n", ierr); if (ierr > 1e-5) ++fails;
    }
    // ---- fft32 vs direct, warp-level fft1024 (32 x 32) vs direct, two-real-frames split / merge
    {
        c32 v[32]; std::vector<cd> x(32);
        for (int n = 0; n < 32; ++n) { v[n] = mk((float)frand(), (float)frand()); x[n] = cd(v[n].x, v[n].y); }
        fft32(v);
        double err = 0;
        for (int k = 0; k < 32; ++k) { cd s = 0; for (int n = 0; n < 32; ++n) s += x[n] * std::polar(1.0, -2 * PI * n * k / 32.0); err = fmax(err, std::abs(s - cd(v[k].x, v[k].y))); }
        printf("fft32    max err %.3e\n", err); if (err > 1e-5) ++fails;
    }
    {
        std::vector<c32> tww(1024), sm(FFT1024_SMEM);
        for (int j = 0; j < 32; ++j) for (int l = 0; l < 32; ++l) tww[j * 32 + l] = mk((float)cos(2 * PI * l * j / 1024), (float)-sin(2 * PI * l * j / 1024));
        std::vector<float> xa(1024), xb(1024);
        for (int n = 0; n < 1024; ++n) { xa[n] = (float)frand(); xb[n] = (float)frand(); }
        c32 regs[32][32];
        for (int l = 0; l < 32; ++l) { for (int j = 0; j < 32; ++j) regs[l][j] = mk(xa[l + 32 * j], xb[l + 32 * j]); fft1024w_pass1(regs[l], sm.data(), tww.data(), l); }
        for (int l = 0; l < 32; ++l) fft1024w_pass2(regs[l], sm.data(), l);
        std::vector<cd> Z(1024);
        for (int l = 0; l < 32; ++l) for (int j = 0; j < 32; ++j) Z[32 * j + l] = cd(regs[l][j].x, regs[l][j].y);
        double err = 0, erra = 0;
        for (int k = 0; k <= 512; ++k) {
            cd sa = 0, sb = 0;
            for (int n = 0; n < 1024; ++n) { cd w = std::polar(1.0, -2 * PI * n * k / 1024.0); sa += (double)xa[n] * w; sb += (double)xb[n] * w; }
            err = fmax(err, std::abs(sa + cd(0, 1) * sb - Z[k]));
            const cd zp = std::conj(Z[(1024 - k) % 1024]);
            erra = fmax(erra, std::abs((Z[k] + zp) * 0.5 - sa)); erra = fmax(erra, std::abs((Z[k] - zp) * cd(0, -0.5) - sb));
        }
        printf("fft1024w max err %.3e, two-frame split max err %.3e\n", err, erra); if (err > 4e-4 || erra > 4e-4) ++fails;
    }
    // ---- fft200 + rfft400
    {
        std::vector<float> x(400); for (auto& v : x) v = (float)frand();
        std::vector<c32> sm(200), Z(200), twA(200);
        for (int p = 0; p < 8; ++p) for (int j = 0; j < 25; ++j) twA[p * 25 + j] = tw200[j * p];
        for (int j = 0; j < 25; ++j) { c32 v[8]; for (int m = 0; m < 8; ++m) v[m] = mk(x[2 * (j + 25 * m)], x[2 * (j + 25 * m) + 1]); fft200_stepA(v, sm.data(), twA.data(), j); }
        for (int p = 0; p < 8; ++p) { c32 o[25]; fft200_stepB(sm.data(), p, o); for (int q = 0; q < 25; ++q) Z[p + 8 * q] = o[q]; }
        double err = 0;
        for (int k = 0; k <= 200; ++k) {
            c32 zk = Z[k % 200], znk = Z[(200 - k) % 200];
            c32 X = rfft400_bin(zk, znk, tw400[k]);
            cd s = 0; for (int n = 0; n < 400; ++n) s += (double)x[n] * std::polar(1.0, -2 * PI * n * k / 400.0);
            err = fmax(err, std::abs(s - cd(X.x, X.y)));
        }
        printf("rfft400  max err %.3e\n", err); if (err > 1e-4) ++fails;
    }
    printf(fails ? "FAIL\n" : "OK\n");
    return fails;
}
