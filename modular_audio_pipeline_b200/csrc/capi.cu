// C ABI (include/maf.h): context, batch plan (item / chunk tables, workspace layout) and the stage sequencing.
// No kernel lives here; see noise_detect.cu, stft_denoise.cu, loudness.cu, logmel.cu.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <stddef.h>
#include <stdint.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/maf.h"
#include "maf_common.cuh"

using namespace maf;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(x)                                                                                           \
    do {                                                                                                \
        cudaError_t e_ = (x);                                                                           \
        if (e_ != cudaSuccess) return fail(MAF_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); \
    } while (0)

static inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

// ---- context --------------------------------------------------------------------------------------------------
struct MelHost { std::vector<int> start, count, woff; std::vector<float> w; std::vector<int> qstart, qcount, qoff, qrow; std::vector<float> wq; };

struct maf_ctx {
    int device = 0;
    void* blob = nullptr;                // one allocation holding every constant table
    void* scratch = nullptr;             // 1 KB + SEL_SCRATCH_BYTES of device scratch for the single-recording helpers
    Consts c{};
    MelBank mel80{}, mel128{};
};

static double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = 1000.0 / f_sp, logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = 1000.0 / f_sp, logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}
// librosa.filters.mel(sr=16000, n_fft=400, n_mels, htk=False, norm='slaney') as sparse rows (whisper's mel_filters.npz)
static MelHost build_mel(int n_mels) {
    MelHost h;
    std::vector<double> mel_f(n_mels + 2);
    const double m0 = hz_to_mel(0.0), m1 = hz_to_mel(8000.0);
    for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz(m0 + (m1 - m0) * i / (n_mels + 1));
    for (int m = 0; m < n_mels; ++m) {
        const double enorm = 2.0 / (mel_f[m + 2] - mel_f[m]);
        int st = -1, cnt = 0;
        std::vector<float> row;
        for (int k = 0; k < W_NBIN; ++k) {
            const double fk = k * (16000.0 / W_NFFT);
            const double lower = (fk - mel_f[m]) / (mel_f[m + 1] - mel_f[m]);
            const double upper = (mel_f[m + 2] - fk) / (mel_f[m + 2] - mel_f[m + 1]);
            const double wv = fmax(0.0, fmin(lower, upper)) * enorm;
            const float wf = (float)wv;
            if (wf != 0.0f) {
                if (st < 0) st = k;
                while (st + (int)row.size() < k) row.push_back(0.0f);     // keep the run contiguous
                row.push_back(wf);
            }
        }
        cnt = (int)row.size();
        h.start.push_back(st < 0 ? 0 : st); h.count.push_back(cnt); h.woff.push_back((int)h.w.size());
        h.w.insert(h.w.end(), row.begin(), row.end());
    }
    // quads of adjacent rows: contiguous bin run covering the four rows, one float4 of weights per bin
    for (int q = 0; q < n_mels / 4; ++q) {
        int lo = W_NBIN, hi = 0;
        for (int j = 0; j < 4; ++j) {
            const int m = 4 * q + j;
            if (h.count[m] > 0) { lo = std::min(lo, h.start[m]); hi = std::max(hi, h.start[m] + h.count[m]); }
        }
        if (hi <= lo) { lo = 0; hi = 0; }
        const int padded = (hi - lo + 3) / 4 * 4;                // whole groups of four bins (zero weights past the run: the
        h.qstart.push_back(lo); h.qcount.push_back(padded); h.qoff.push_back((int)h.wq.size() / 4);   // kernel has no remainder loop)
        for (int k = lo; k < lo + padded; ++k)
            for (int j = 0; j < 4; ++j) {
                const int m = 4 * q + j;
                const bool in = k < hi && k >= h.start[m] && k < h.start[m] + h.count[m];
                h.wq.push_back(in ? h.w[h.woff[m] + k - h.start[m]] : 0.0f);
            }
    }
    // deal the quads to the kernel's 8 warps, longest first onto the least loaded warp (the runs grow from 4 bins at the
    // bottom of the scale to 34 at the top: round-robin leaves one warp with 40 % more work than the average)
    {
        const int nq = n_mels / 4;
        std::vector<int> order(nq), load(8, 0);
        std::vector<std::vector<int>> lists(8);
        for (int q = 0; q < nq; ++q) order[q] = q;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h.qcount[a] > h.qcount[b]; });
        for (int q : order) {
            int w = 0;
            for (int i = 1; i < 8; ++i) if (load[i] < load[w]) w = i;
            load[w] += h.qcount[q] + 8;                          // + the log / store epilogue of a quad
            lists[w].push_back(q);
        }
        size_t depth = 0;
        for (auto& l : lists) depth = std::max(depth, l.size());
        std::vector<int> qs(8 * depth, 0), qc(8 * depth, 0), qo(8 * depth, 0), qr(8 * depth, -1);
        for (int w = 0; w < 8; ++w)
            for (size_t i = 0; i < lists[w].size(); ++i) {
                const int q = lists[w][i], sl = (int)(8 * i) + w;
                qs[sl] = h.qstart[q]; qc[sl] = h.qcount[q]; qo[sl] = h.qoff[q]; qr[sl] = q;
            }
        h.qstart = qs; h.qcount = qc; h.qoff = qo; h.qrow = qr;
    }
    return h;
}

extern "C" int maf_version(void) { return 1; }
extern "C" const char* maf_last_error(void) { return g_err.c_str(); }

extern "C" int maf_create(int device, maf_ctx** out) {
    if (!out) return fail(MAF_ERR_INVALID, "maf_create: out is NULL");
    CU(cudaSetDevice(device));
    const double PI = 3.14159265358979323846;
    std::vector<unsigned char> host;
    auto put = [&](const void* p, size_t bytes) { size_t off = align_up((long long)host.size(), 256); host.resize(off + bytes); memcpy(host.data() + off, p, bytes); return off; };
    std::vector<float> win(NFFT), winv(NFFT), win4(W_NFFT);
    for (int i = 0; i < NFFT; ++i) { const double w = 0.5 - 0.5 * cos(2.0 * PI * i / NFFT); win[i] = (float)(w / 512.0); winv[i] = (float)(w * 0.5); }
    for (int i = 0; i < W_NFFT; ++i) win4[i] = (float)(0.5 - 0.5 * cos(2.0 * PI * i / W_NFFT));
    std::vector<c32> t512(FFT512_TW1), t2(FFT512_TW2), t1024(513), t200(200), t400(201);
    for (int p = 0; p < 8; ++p) for (int t = 0; t < 64; ++t) t512[p * 64 + t] = mk((float)cos(2 * PI * (t * p) / 512), (float)-sin(2 * PI * (t * p) / 512));
    for (int c = 0; c < 8; ++c) for (int a = 0; a < 8; ++a) t2[c * 8 + a] = mk((float)cos(2 * PI * (8 * a * c) / 512), (float)-sin(2 * PI * (8 * a * c) / 512));
    for (int e = 0; e <= 512; ++e) t1024[e] = mk((float)cos(2 * PI * e / 1024), (float)-sin(2 * PI * e / 1024));
    for (int p = 0; p < 8; ++p) for (int j = 0; j < 25; ++j) t200[p * 25 + j] = mk((float)cos(2 * PI * (j * p) / 200), (float)-sin(2 * PI * (j * p) / 200));
    for (int e = 0; e <= 200; ++e) t400[e] = mk((float)cos(2 * PI * e / 400), (float)-sin(2 * PI * e / 400));
    std::vector<c32> tww(1024);
    for (int j = 0; j < 32; ++j) for (int l = 0; l < 32; ++l) tww[j * 32 + l] = mk((float)cos(2 * PI * (l * j) / 1024), (float)-sin(2 * PI * (l * j) / 1024));
    MelHost m80 = build_mel(80), m128 = build_mel(128);
    size_t o_tww = put(tww.data(), tww.size() * 8);
    size_t o_win = put(win.data(), win.size() * 4), o_winv = put(winv.data(), winv.size() * 4), o_win4 = put(win4.data(), win4.size() * 4);
    size_t o_t512 = put(t512.data(), t512.size() * 8), o_t2 = put(t2.data(), t2.size() * 8), o_t1024 = put(t1024.data(), t1024.size() * 8);
    size_t o_t200 = put(t200.data(), t200.size() * 8), o_t400 = put(t400.data(), t400.size() * 8);
    if (m80.wq.size() / 4 > MEL_WQ_MAX || m128.wq.size() / 4 > MEL_WQ_MAX) return fail(MAF_ERR_UNSUPPORTED, "mel quad table exceeds MEL_WQ_MAX");
    if (m80.qrow.size() > MEL_SLOTS_MAX || m128.qrow.size() > MEL_SLOTS_MAX) return fail(MAF_ERR_UNSUPPORTED, "mel slot table exceeds MEL_SLOTS_MAX");
    auto put_mel = [&](const MelHost& m, size_t (&o)[5]) {
        const size_t ns = m.qrow.size() * 4;
        o[0] = put(m.qrow.data(), ns); o[1] = put(m.qstart.data(), ns); o[2] = put(m.qcount.data(), ns); o[3] = put(m.qoff.data(), ns);
        o[4] = put(m.wq.data(), m.wq.size() * 4);
    };
    size_t o80[5], o128[5];
    put_mel(m80, o80); put_mel(m128, o128);
    maf_ctx* ctx = new maf_ctx();
    ctx->device = device;
    cudaError_t e = cudaMalloc(&ctx->blob, host.size());
    if (e != cudaSuccess) { delete ctx; return fail(MAF_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
    e = cudaMemcpy(ctx->blob, host.data(), host.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(ctx->blob); delete ctx; return fail(MAF_ERR_CUDA, std::string("cudaMemcpy: ") + cudaGetErrorString(e)); }
    if (cudaMalloc(&ctx->scratch, 1024 + SEL_SCRATCH_BYTES) != cudaSuccess) { cudaFree(ctx->blob); delete ctx; return fail(MAF_ERR_CUDA, "cudaMalloc scratch"); }
    unsigned char* b = (unsigned char*)ctx->blob;
    ctx->c.win1024 = (const float*)(b + o_win); ctx->c.winv1024 = (const float*)(b + o_winv); ctx->c.win400 = (const float*)(b + o_win4);
    ctx->c.tw1 = (const c32*)(b + o_t512); ctx->c.tw2 = (const c32*)(b + o_t2); ctx->c.tw1024 = (const c32*)(b + o_t1024);
    ctx->c.twA = (const c32*)(b + o_t200); ctx->c.tw400 = (const c32*)(b + o_t400);
    ctx->c.tww = (const c32*)(b + o_tww);
    ctx->mel80 = MelBank{80, (int)m80.wq.size() / 4, (int)m80.qrow.size(), (const int*)(b + o80[0]), (const int*)(b + o80[1]), (const int*)(b + o80[2]),
                         (const int*)(b + o80[3]), (const float4*)(b + o80[4])};
    ctx->mel128 = MelBank{128, (int)m128.wq.size() / 4, (int)m128.qrow.size(), (const int*)(b + o128[0]), (const int*)(b + o128[1]), (const int*)(b + o128[2]),
                          (const int*)(b + o128[3]), (const float4*)(b + o128[4])};
    *out = ctx;
    return MAF_OK;
}

extern "C" void maf_destroy(maf_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->blob);
    cudaFree(ctx->scratch);
    delete ctx;
}

// ---- plan -----------------------------------------------------------------------------------------------------
struct WsBuf { long long off = 0, bytes = 0; };

struct maf_plan {
    maf_ctx* ctx = nullptr;
    int device = 0;
    maf_params p{};
    int n_items = 0, n_chunks = 0;
    std::vector<Item> items;           // den_off = workspace den16 layout
    std::vector<Chunk> chunks;
    std::vector<maf_layout> layout;
    std::vector<int> seg_fixed;        // (start, end) of the fixed noise clip per item
    void* tables = nullptr;            // device: Item[n] (ws layout) | Item[n] (io layout) | Chunk[] | foff | fcount | Clip[n] (fixed)
    Item* d_items = nullptr; Item* d_items_io = nullptr; Chunk* d_chunks = nullptr;
    long long* d_foff = nullptr; int* d_fcount = nullptr; Clip* d_clips_fixed = nullptr;
    long long total_pcm = 0, total_out = 0, total_mel = 0, total_rows = 0, total_sframes = 0, total_den = 0;
    int total_ktiles = 0, max_chunk_frames = 0, max_sframes = 0, max_ktiles = 0, max_mel_tiles = 0, max_mel_frames = 0;
    long long max_own = 0;
    int frame_len = 0, frame_hop = 0, tile_len = 0, n_slices = 1;
    bool too_short = false, sharded = false, have_noise_clips = false;
    std::string clip_err;              // why the stationary denoise of this plan must fail (noise clip <= noverlap)
    double target_peak = 0.0, inv_block = 0.0;
    KWeight kw{};
    SmoothTaps taps{};
    NsParams ns{}; void* d_ns = nullptr;
    KwGeom kgeo{}; KwScan kscan{}; void* d_kscan = nullptr; int max_kgroups = 0;
    long long total_tiles = 0;
    std::vector<WsBuf> ws_layout;        // every carved buffer in order (maf_plan_ws_layout)
    WsBuf S, M, P, Q, Yend, Zend, bits, rowflag, T2, T2hi, nmax, partial, E, ZC, clips, seg, den16, norm16, peak, pnorm, peak_norm, mel_max, mel_tmin, sub, gates, selws;
    long long ws_bytes = 0;
    long long launches = 0;
    // optional per-stage CUDA-event timing (maf_plan_profile): ring of event sets, one per maf_process call
    int prof_steps = 0; long long prof_calls = 0; std::vector<cudaEvent_t> prof_ev;   // [(step, boundary)]
    // host-path staging (maf_process_host)
    cudaStream_t hstream = nullptr; bool host_ready = false;
    void* h_ws = nullptr; int16_t* h_pcm = nullptr; int16_t* h_out = nullptr; float* h_mel = nullptr; double* h_lufs = nullptr;
    int* h_clip = nullptr; int16_t* h_noise = nullptr; long long h_noise_cap = 0;
};

static void design_kweight(double rate, KWeight& kw) {
    const double PI = 3.141592653589793;
    {   // high shelf: G = 4 dB, Q = 1/sqrt(2), fc = 1500 Hz
        const double G = 4.0, Q = 1.0 / sqrt(2.0), fc = 1500.0;
        const double A = pow(10.0, G / 40.0), w0 = 2.0 * PI * (fc / rate), alpha = sin(w0) / (2.0 * Q);
        const double b0 = A * ((A + 1) + (A - 1) * cos(w0) + 2 * sqrt(A) * alpha);
        const double b1 = -2 * A * ((A - 1) + (A + 1) * cos(w0));
        const double b2 = A * ((A + 1) + (A - 1) * cos(w0) - 2 * sqrt(A) * alpha);
        const double a0 = (A + 1) - (A - 1) * cos(w0) + 2 * sqrt(A) * alpha;
        const double a1 = 2 * ((A - 1) - (A + 1) * cos(w0));
        const double a2 = (A + 1) - (A - 1) * cos(w0) - 2 * sqrt(A) * alpha;
        kw.b1[0] = b0 / a0; kw.b1[1] = b1 / a0; kw.b1[2] = b2 / a0; kw.a1[0] = 1.0; kw.a1[1] = a1 / a0; kw.a1[2] = a2 / a0;
    }
    {   // high pass: G = 0, Q = 0.5, fc = 38 Hz
        const double Q = 0.5, fc = 38.0;
        const double w0 = 2.0 * PI * (fc / rate), alpha = sin(w0) / (2.0 * Q);
        const double b0 = (1 + cos(w0)) / 2, b1 = -(1 + cos(w0)), b2 = (1 + cos(w0)) / 2;
        const double a0 = 1 + alpha, a1 = -2 * cos(w0), a2 = 1 - alpha;
        kw.b2[0] = b0 / a0; kw.b2[1] = b1 / a0; kw.b2[2] = b2 / a0; kw.a2[0] = 1.0; kw.a2[1] = a1 / a0; kw.a2[2] = a2 / a0;
    }
}

// Segmentation of the loudness sub-blocks for kweight_energy_kernel and the powers of the homogeneous state transition.
static int design_kwscan(const KWeight& kw, int tile_len, KwGeom& g, KwScan& t) {
    int best = 0;
    for (int d = 16; d <= 128; ++d)
        if (tile_len % d == 0 && 2 * (tile_len / d) <= KW_MAX_THREADS && (!best || abs(d - 50) < abs(best - 50))) best = d;
    if (!best) return -1;
    g.seg = best; g.lpt = tile_len / best;
    g.tpc = std::min(15, std::min(KW_MAX_THREADS / g.lpt - 1, 26000 / tile_len - 1));
    if (g.tpc < 1) return -1;
    g.threads = ((g.tpc + 1) * g.lpt + 31) / 32 * 32;
    // one homogeneous step (x = 0) of the cascade on (z10, z11, z20, z21); see kw_step in loudness.cu
    typedef long double ld;
    memset(&t, 0, sizeof(t));
    ld A[16] = {0};
    A[0] = -(ld)kw.a1[1]; A[1] = 1;                                   // z10' = z11 - a11 z10
    A[4] = -(ld)kw.a1[2];                                             // z11' = -a12 z10
    A[8] = (ld)kw.b2[1] - (ld)kw.a2[1] * (ld)kw.b2[0]; A[10] = -(ld)kw.a2[1]; A[11] = 1;   // y1 = z10, y2 = b20 z10 + z20
    A[12] = (ld)kw.b2[2] - (ld)kw.a2[2] * (ld)kw.b2[0]; A[14] = -(ld)kw.a2[2];
    auto mul = [](const ld* x, const ld* y, ld* o) {
        ld r[16];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { ld a = 0; for (int k = 0; k < 4; ++k) a += x[4 * i + k] * y[4 * k + j]; r[4 * i + j] = a; }
        for (int i = 0; i < 16; ++i) o[i] = r[i];
    };
    ld P[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    for (int i = 0; i < g.seg; ++i) mul(A, P, P);                     // P = A^seg
    ld L[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    for (int l = 0; l < 32; ++l) { for (int i = 0; i < 16; ++i) t.Mlane[l][i] = (double)L[i]; mul(P, L, L); }
    for (int i = 0; i < 16; ++i) t.Mwarp[i] = (double)L[i];           // A^(32 seg)
    {   // state after one step from zero state with x = 1, then i homogeneous steps
        const ld b10 = kw.b1[0], b20 = kw.b2[0];
        ld v[4] = {(ld)kw.b1[1] - (ld)kw.a1[1] * b10, (ld)kw.b1[2] - (ld)kw.a1[2] * b10,
                   ((ld)kw.b2[1] - (ld)kw.a2[1] * b20) * b10, ((ld)kw.b2[2] - (ld)kw.a2[2] * b20) * b10};
        for (int i = g.seg - 1; i >= 0; --i) {
            for (int r = 0; r < 4; ++r) t.H[i][r] = (float)(v[r] / 32768.0L);
            ld nv[4];
            for (int r = 0; r < 4; ++r) nv[r] = A[4 * r] * v[0] + A[4 * r + 1] * v[1] + A[4 * r + 2] * v[2] + A[4 * r + 3] * v[3];
            for (int r = 0; r < 4; ++r) v[r] = nv[r];
        }
    }
    ld S[16]; for (int i = 0; i < 16; ++i) S[i] = P[i];
    for (int k = 0; k < 5; ++k) { for (int i = 0; i < 16; ++i) t.Mstep[k][i] = (double)S[i]; mul(S, S, S); }
    return 0;
}

static int design_taps(int sr, SmoothTaps& t) {
    // noisereduce: n_grad_freq = int(500 / (sr / (n_fft / 2))), n_grad_time = int(50 / ((hop / sr) * 1000))
    const int nf = (int)(500.0 / ((double)sr / (NFFT / 2)));
    const int nt = (int)(50.0 / (((double)HOP / sr) * 1000.0));
    if (nf > MAX_NF || nt > MAX_NT || nf < 0 || nt < 0) return -1;
    // concat(linspace(0, 1, n + 1, endpoint=False), linspace(1, 0, n + 2))[1:-1], normalised per axis
    auto tri = [](int n, float* out) {
        std::vector<double> v;
        for (int i = 1; i <= n; ++i) v.push_back((double)i / (n + 1));
        for (int i = 0; i <= n; ++i) v.push_back(1.0 - (double)i / (n + 1));
        double s = 0; for (double x : v) s += x;
        for (size_t i = 0; i < v.size(); ++i) out[i] = (float)(v[i] / s);
    };
    t.nf = nf; t.nt = nt;
    memset(t.vf, 0, sizeof(t.vf)); memset(t.vt, 0, sizeof(t.vt));
    tri(nf, t.vf); tri(nt, t.vt);
    return 0;
}

// Workspace carving.  MAF_DEBUG_REDZONE=<bytes> (read at plan creation) leaves that many unused bytes after every buffer: the
// bounds-check build of the parity suite fills the workspace with 0xFF, runs a step and verifies that every gap is intact and
// that the outputs do not depend on what the gaps hold (tests/test_gpu_parity.py::test_workspace_red_zones) -- compute-sanitizer
// is not available on the GPU pool.  `g_layout` collects the carved (offset, bytes) pairs of the plan being built.
static long long g_redzone = 0;
static std::vector<WsBuf>* g_layout = nullptr;
static WsBuf carve(long long& cursor, long long bytes) {
    WsBuf b; b.off = align_up(cursor, 256); b.bytes = bytes; cursor = b.off + bytes + g_redzone;
    if (g_layout) g_layout->push_back(b);
    return b;
}

extern "C" int maf_plan_create(maf_ctx* ctx, const maf_params* params, const maf_item_desc* descs, int32_t n_items, maf_plan** out) {
    if (!ctx || !params || !descs || !out || n_items <= 0) return fail(MAF_ERR_INVALID, "maf_plan_create: bad argument");
    const maf_params& p = *params;
    if (p.sample_rate < 4000 || p.sample_rate > 96000) return fail(MAF_ERR_INVALID, "sample_rate out of range");
    if (p.n_mels != 0 && p.n_mels != 80 && p.n_mels != 128) return fail(MAF_ERR_INVALID, "n_mels must be 0, 80 or 128");
    if (p.n_mels && p.sample_rate != 16000) return fail(MAF_ERR_UNSUPPORTED, "Whisper log-mel is defined at 16 kHz");
    if (p.mel_padding < 0 || (p.mel_padding > 0 && p.mel_padding < W_HOP)) return fail(MAF_ERR_INVALID, "mel_padding must be 0 or >= 160");
    // several kernels index items with blockIdx.y (noise statistics, frame statistics, the peak scan)
    if (n_items > 65535) return fail(MAF_ERR_UNSUPPORTED, "more than 65535 items in one plan: split the batch");
    CU(cudaSetDevice(ctx->device));
    maf_plan* pl = new maf_plan();
    pl->ctx = ctx; pl->device = ctx->device; pl->p = p; pl->n_items = n_items;
    if (p.stationary && !p.external_noise) {
        // scipy.signal.stft raises "noverlap must be less than nperseg" for a noise clip of <= 768 samples (preprocessor.py:207-208
        // turns that into AudioProcessingError); the shortest run the detector accepts is 100 ms, the fallback clip is
        // noise_sample_duration_s long (preprocessor.py:189-197).  Reported by the denoise entry points, not here: the peak /
        // loudness / log-mel stages of such a plan are fine.
        const int hop_s0 = (int)(p.sample_rate * 0.010);
        const int min_run = hop_s0 > 0 ? (int)ceil(0.1 * p.sample_rate / hop_s0) * hop_s0 : 0;
        if (p.auto_detect_noise && min_run <= NFFT - HOP)
            pl->clip_err = "sample rate too low: a detected 100 ms noise clip would be <= 768 samples (scipy.signal.stft: noverlap must be less than nperseg)";
        else if ((long long)(p.sample_rate * (double)p.noise_sample_duration_s) <= NFFT - HOP)
            pl->clip_err = "noise_sample_duration_s gives a noise clip of <= 768 samples (scipy.signal.stft: noverlap must be less than nperseg)";
    }
    if (design_taps(p.sample_rate, pl->taps)) { delete pl; return fail(MAF_ERR_UNSUPPORTED, "smoothing taps exceed kernel limits for this rate"); }
    design_kweight((double)p.sample_rate, pl->kw);
    {   // noisereduce non-stationary defaults: time_constant_s = 2.0, thresh_n_mult_nonstationary = 2, sigmoid_slope_nonstationary = 10
        const double t = 2.0 * p.sample_rate / HOP;
        NsParams& ns = pl->ns;
        ns.b = (sqrt(1.0 + 4.0 * t * t) - 1.0) / (2.0 * t * t);
        ns.a = 1.0 - ns.b; ns.inv_a = 1.0 / ns.a; ns.inv_1ma2 = 1.0 / (1.0 - ns.a * ns.a);
        for (int i = 0; i <= 2 * STFT_TILE; ++i) ns.apow[i] = pow(ns.a, (double)i);
        ns.thresh = 2.0f; ns.slope = 10.0f;
    }
    pl->target_peak = 32768.0 * pow(10.0, -0.1 / 20.0);
    pl->inv_block = 1.0 / (0.4 * (double)p.sample_rate);
    pl->frame_len = (int)(p.sample_rate * 0.025);
    pl->frame_hop = (int)(p.sample_rate * 0.010);
    if (pl->frame_len >= 2048) { delete pl; return fail(MAF_ERR_UNSUPPORTED, "frame length too large for the ZCR histogram"); }
    pl->tile_len = (int)(0.4 * (1 * 0.25) * (double)p.sample_rate);
    if (design_kwscan(pl->kw, pl->tile_len, pl->kgeo, pl->kscan)) { delete pl; return fail(MAF_ERR_UNSUPPORTED, "no loudness sub-block segmentation at this rate"); }
    std::vector<Clip> fixed(n_items);
    std::vector<long long> foff(n_items); std::vector<int> fcount(n_items);
    long long pcm_cur = 0, out_cur = 0, den_cur = 0, mel_cur = 0, row_cur = 0, sf_cur = 0, tile_cur = 0; int kt_cur = 0;
    for (int i = 0; i < n_items; ++i) {
        const maf_item_desc& d = descs[i];
        const bool whole = d.g0 == 0 && d.own_len == d.n_global;
        if (d.own_len <= 0 || d.n_global <= 0 || d.origin < 0 || d.origin + d.own_len > d.buf_len || d.g0 < 0 || d.g0 + d.own_len > d.n_global)
            { delete pl; return fail(MAF_ERR_INVALID, "maf_plan_create: inconsistent item descriptor"); }
        if (d.n_global > 2000000000LL) { delete pl; return fail(MAF_ERR_UNSUPPORTED, "recordings above 2^31 samples are not supported"); }
        if (p.stationary && !p.external_noise && d.n_global <= NFFT - HOP && pl->clip_err.empty())
            pl->clip_err = "recording of <= 768 samples: its noise clip is shorter than noverlap (scipy.signal.stft: noverlap must be less than nperseg)";
        if (!whole) {
            pl->sharded = true;
            if (d.n_global <= CHUNK || d.g0 % CHUNK != 0 || (d.g0 + d.own_len != d.n_global && d.own_len % CHUNK != 0))
                { delete pl; return fail(MAF_ERR_INVALID, "time shards must be whole 600000-sample chunks of a multi-chunk recording"); }
            if (d.g0 % pl->frame_hop != 0 || d.g0 % W_HOP != 0 || d.g0 % pl->tile_len != 0)
                { delete pl; return fail(MAF_ERR_UNSUPPORTED, "shard boundary is not aligned to the frame / block hops at this rate"); }
        }
        Item it{};
        it.pcm_off = pcm_cur; pcm_cur += align_up(d.buf_len, 8);
        it.buf_len = d.buf_len; it.origin = d.origin; it.own_len = d.own_len; it.g0 = d.g0; it.n_global = d.n_global;
        it.lo_avail = whole ? 0 : d.lo_avail; it.hi_avail = whole ? 0 : d.hi_avail;
        if (it.lo_avail > KW_WARM || it.hi_avail > KW_WARM || it.lo_avail < 0 || it.hi_avail < 0)
            { delete pl; return fail(MAF_ERR_INVALID, "halo larger than 2048 samples"); }
        it.den_off = den_cur + KW_WARM; den_cur += align_up(d.own_len + 2 * KW_WARM, 8);
        maf_layout lay{}; lay.pcm_off = it.pcm_off; lay.out_off = out_cur; out_cur += align_up(d.own_len, 8);
        it.out_off = lay.out_off;
        // denoise chunks
        it.chunk_first = (int)pl->chunks.size();
        const bool multi = d.n_global > CHUNK;
        const int nch = multi ? (int)((d.own_len + CHUNK - 1) / CHUNK) : 1;
        it.n_chunks = nch;
        for (int k = 0; k < nch; ++k) {
            Chunk c{};
            c.item = i;
            c.L = multi ? CHUNK + 2 * PADS : (int)d.n_global + 2 * PADS;
            c.n_frames = c.L / HOP + 1;
            c.keep = multi ? (int)std::min<long long>(CHUNK, d.own_len - (long long)k * CHUNK) : (int)d.n_global;
            c.s_local = d.origin + (long long)k * CHUNK;
            c.row0 = row_cur; row_cur += c.n_frames;
            c.out_off = (long long)k * CHUNK;
            c.tile0 = tile_cur; tile_cur += (c.n_frames + STFT_TILE - 1) / STFT_TILE;
            pl->chunks.push_back(c);
            pl->max_chunk_frames = std::max(pl->max_chunk_frames, c.n_frames);
        }
        // S0 frames
        {
            const long long Fg = d.n_global >= pl->frame_len ? (d.n_global - pl->frame_len) / pl->frame_hop + 1 : 0;
            const long long first = (d.g0 + pl->frame_hop - 1) / pl->frame_hop;
            const long long last = std::min<long long>(Fg, (d.g0 + d.own_len + pl->frame_hop - 1) / pl->frame_hop);
            it.n_sframes = (int)std::max<long long>(0, last - first);
            it.sframe_off = sf_cur; sf_cur += it.n_sframes;
            if (it.n_sframes > 0 && d.origin + (long long)(it.n_sframes - 1) * pl->frame_hop + pl->frame_len > d.buf_len)
                { delete pl; return fail(MAF_ERR_INVALID, "item buffer lacks the right halo needed by the noise detector"); }
            foff[i] = it.sframe_off; fcount[i] = it.n_sframes;
            pl->max_sframes = std::max(pl->max_sframes, it.n_sframes);
        }
        // fixed noise clip: first noise_sample_duration_s of the recording (preprocessor.py:189-197)
        {
            long long len = std::min<long long>((long long)(p.sample_rate * (double)p.noise_sample_duration_s), d.n_global);
            if (len > CHUNK) len = CHUNK;
            fixed[i].start = it.pcm_off + d.origin - d.g0;       // absolute offset in the pcm array of global sample 0
            fixed[i].len = (int)len; fixed[i].n_frames = clip_frames((int)len);
            pl->seg_fixed.push_back(0); pl->seg_fixed.push_back((int)len);
        }
        // loudness tiles / blocks
        {
            it.ktile_first = kt_cur;
            it.n_ktiles = (int)((d.own_len + pl->tile_len - 1) / pl->tile_len);
            kt_cur += it.n_ktiles;
            pl->max_ktiles = std::max(pl->max_ktiles, it.n_ktiles);
            {
                const long long t0 = d.g0 / pl->tile_len;
                const long long groups = (t0 + it.n_ktiles - 1) / pl->kgeo.tpc - t0 / pl->kgeo.tpc + 1;
                pl->max_kgroups = std::max<int>(pl->max_kgroups, (int)groups);
            }
            const double T = (double)d.n_global / (double)p.sample_rate;
            if ((double)d.n_global < 0.4 * (double)p.sample_rate) { pl->too_short = true; it.n_blocks = 0; }
            else {
                const long long nb = (long long)(nearbyint((T - 0.4) / (0.4 * 0.25)) + 1);
                it.n_blocks = (int)nb;
                for (long long j = 0; j < nb; ++j) {       // the tile decomposition needs pyloudnorm's bounds to be regular
                    const long long l = (long long)(0.4 * (j * 0.25) * (double)p.sample_rate);
                    const long long u = (long long)(0.4 * (j * 0.25 + 1) * (double)p.sample_rate);
                    if (l != j * pl->tile_len || u != l + 4LL * pl->tile_len)
                        { delete pl; return fail(MAF_ERR_UNSUPPORTED, "irregular loudness block grid at this rate / length"); }
                }
            }
        }
        // log-mel frames
        if (p.n_mels) {
            it.mel_frames_global = (int)((d.n_global + p.mel_padding) / W_HOP);
            it.mel_f0 = (int)(d.g0 / W_HOP);
            const bool last = d.g0 + d.own_len == d.n_global;
            it.mel_frames = last ? it.mel_frames_global - it.mel_f0 : (int)(d.own_len / W_HOP);
            if (d.n_global + p.mel_padding <= W_NFFT / 2) { delete pl; return fail(MAF_ERR_TOO_SHORT, "recording too short for reflect padding"); }
            it.mel_off = mel_cur; mel_cur += align_up((long long)p.n_mels * it.mel_frames, 4);
            pl->max_mel_frames = std::max(pl->max_mel_frames, it.mel_frames);
            pl->max_mel_tiles = std::max(pl->max_mel_tiles, (it.mel_frames + MEL_TILE - 1) / MEL_TILE);
        }
        lay.mel_off = it.mel_off; lay.mel_frames = it.mel_frames; lay.n_chunks = nch;
        pl->max_own = std::max<long long>(pl->max_own, d.own_len);
        pl->items.push_back(it); pl->layout.push_back(lay);
    }
    pl->n_chunks = (int)pl->chunks.size();
    pl->total_pcm = pcm_cur; pl->total_out = out_cur; pl->total_mel = mel_cur; pl->total_rows = row_cur;
    pl->total_sframes = sf_cur; pl->total_den = den_cur; pl->total_ktiles = kt_cur; pl->total_tiles = tile_cur;
    {   // the persistent kernels index (chunk, tile) / (item, tile) units with 32-bit arithmetic
        const long long st_units = (long long)pl->n_chunks * ((pl->max_chunk_frames + STFT_TILE - 1) / STFT_TILE);
        const long long mel_units = (long long)n_items * std::max(1, pl->max_mel_tiles);
        if (st_units >= (1LL << 31) - 4096 || mel_units >= (1LL << 31) - 4096)
            { delete pl; return fail(MAF_ERR_UNSUPPORTED, "plan too large: more than 2^31 tiles, split the batch"); }
    }
    pl->n_slices = std::max(1, std::min(32, (592 + n_items - 1) / n_items));
    // workspace layout
    long long cur = 0;
    {
        const char* rz = getenv("MAF_DEBUG_REDZONE");
        g_redzone = rz ? std::max<long long>(0, atoll(rz)) / 256 * 256 : 0;
        g_layout = &pl->ws_layout;
    }
    pl->S = carve(cur, pl->total_rows * S_PITCH * (long long)sizeof(c32));
    pl->M = carve(cur, pl->total_rows * M_PITCH * 4LL);
    {
        const long long tb = p.stationary ? 0 : pl->total_tiles * T_PITCH * 4LL;
        pl->P = carve(cur, tb); pl->Q = carve(cur, tb); pl->Yend = carve(cur, tb); pl->Zend = carve(cur, tb);
    }
    pl->bits = carve(cur, p.stationary ? pl->total_rows * B_PITCH * 4LL : 0);
    pl->rowflag = carve(cur, (long long)pl->n_chunks * M_PITCH);
    pl->T2 = carve(cur, (long long)n_items * T_PITCH * 4);
    pl->T2hi = carve(cur, (long long)n_items * T_PITCH * 4);
    pl->nmax = carve(cur, (long long)n_items * T_PITCH * 4);
    pl->partial = carve(cur, (long long)n_items * pl->n_slices * 4 * 2 * T_PITCH * 8);
    pl->E = carve(cur, std::max<long long>(1, pl->total_sframes) * 4);
    pl->ZC = carve(cur, std::max<long long>(1, pl->total_sframes) * 4);
    pl->clips = carve(cur, (long long)n_items * sizeof(Clip));
    pl->seg = carve(cur, (long long)n_items * 2 * 4);
    pl->den16 = carve(cur, pl->total_den * 2);
    pl->norm16 = carve(cur, pl->total_den * 2);       // peak-normalised samples (den16 layout), written by kweight_energy_kernel for logmel_kernel
    pl->peak = carve(cur, (long long)n_items * 4);
    pl->pnorm = carve(cur, (long long)n_items * sizeof(PeakNorm));
    pl->peak_norm = carve(cur, (long long)n_items * 4);
    pl->mel_max = carve(cur, (long long)n_items * 4);
    pl->mel_tmin = carve(cur, (long long)n_items * std::max(1, pl->max_mel_tiles) * 4);
    pl->sub = carve(cur, std::max(1, pl->total_ktiles) * 8LL);
    pl->gates = carve(cur, (long long)n_items * sizeof(Gate));
    pl->selws = carve(cur, SEL_SCRATCH_BYTES);
    pl->ws_bytes = align_up(cur, 256);
    g_layout = nullptr; g_redzone = 0;
    // device tables
    std::vector<Item> io = pl->items;
    for (int i = 0; i < n_items; ++i) { io[i].den_off = io[i].pcm_off + io[i].origin; io[i].lo_avail = 0; io[i].hi_avail = 0; }
    std::vector<unsigned char> host;
    auto put = [&](const void* ptr, size_t bytes) { size_t off = align_up((long long)host.size(), 256); host.resize(off + bytes); memcpy(host.data() + off, ptr, bytes); return off; };
    const size_t o_it = put(pl->items.data(), sizeof(Item) * n_items), o_io = put(io.data(), sizeof(Item) * n_items);
    const size_t o_ch = put(pl->chunks.data(), sizeof(Chunk) * pl->chunks.size());
    const size_t o_fo = put(foff.data(), 8 * n_items), o_fc = put(fcount.data(), 4 * n_items), o_cl = put(fixed.data(), sizeof(Clip) * n_items);
    const size_t o_ns = put(&pl->ns, sizeof(NsParams));
    const size_t o_ks = put(&pl->kscan, sizeof(KwScan));
    cudaError_t e = cudaMalloc(&pl->tables, host.size());
    if (e == cudaSuccess) e = cudaMemcpy(pl->tables, host.data(), host.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { if (pl->tables) cudaFree(pl->tables); delete pl; return fail(MAF_ERR_CUDA, std::string("plan tables: ") + cudaGetErrorString(e)); }
    unsigned char* b = (unsigned char*)pl->tables;
    pl->d_items = (Item*)(b + o_it); pl->d_items_io = (Item*)(b + o_io); pl->d_chunks = (Chunk*)(b + o_ch);
    pl->d_foff = (long long*)(b + o_fo); pl->d_fcount = (int*)(b + o_fc); pl->d_clips_fixed = (Clip*)(b + o_cl);
    pl->d_ns = (void*)(b + o_ns);
    pl->d_kscan = (void*)(b + o_ks);
    *out = pl;
    return MAF_OK;
}

extern "C" void maf_plan_destroy(maf_plan* pl) {
    if (!pl) return;
    cudaSetDevice(pl->device);
    if (pl->hstream) { cudaStreamSynchronize(pl->hstream); cudaStreamDestroy(pl->hstream); }
    for (cudaEvent_t e : pl->prof_ev) cudaEventDestroy(e);
    cudaFree(pl->tables); cudaFree(pl->h_ws); cudaFree(pl->h_pcm); cudaFree(pl->h_out); cudaFree(pl->h_mel); cudaFree(pl->h_lufs);
    cudaFree(pl->h_clip); cudaFree(pl->h_noise);
    delete pl;
}

extern "C" size_t maf_plan_workspace_bytes(const maf_plan* pl) { return pl ? (size_t)pl->ws_bytes : 0; }
extern "C" int maf_plan_layout(const maf_plan* pl, int32_t item, maf_layout* out) {
    if (!pl || !out || item < 0 || item >= pl->n_items) return fail(MAF_ERR_INVALID, "maf_plan_layout: bad argument");
    *out = pl->layout[item];
    return MAF_OK;
}
extern "C" int64_t maf_plan_total_pcm(const maf_plan* pl) { return pl ? pl->total_pcm : 0; }
extern "C" int64_t maf_plan_total_out(const maf_plan* pl) { return pl ? pl->total_out : 0; }
extern "C" int64_t maf_plan_total_mel(const maf_plan* pl) { return pl ? pl->total_mel : 0; }
extern "C" int64_t maf_plan_launches(const maf_plan* pl) { return pl ? pl->launches : 0; }
extern "C" int64_t maf_plan_den_offset(const maf_plan* pl, int32_t item) {
    return (pl && item >= 0 && item < pl->n_items) ? pl->items[item].den_off : -1;
}
extern "C" int maf_plan_buffer(const maf_plan* pl, const char* name, int64_t* off, int64_t* size) {
    if (!pl || !name || !off || !size) return fail(MAF_ERR_INVALID, "maf_plan_buffer: bad argument");
    const WsBuf* b = nullptr;
    if (!strcmp(name, "peak")) b = &pl->peak; else if (!strcmp(name, "peak_norm")) b = &pl->peak_norm;
    else if (!strcmp(name, "mel_max")) b = &pl->mel_max; else if (!strcmp(name, "sub_energy")) b = &pl->sub;
    else if (!strcmp(name, "den16")) b = &pl->den16; else if (!strcmp(name, "gates")) b = &pl->gates;
    else if (!strcmp(name, "T2")) b = &pl->T2; else if (!strcmp(name, "seg")) b = &pl->seg;
    else if (!strcmp(name, "mel_tmin")) b = &pl->mel_tmin;
    if (!b) return fail(MAF_ERR_INVALID, std::string("maf_plan_buffer: unknown buffer ") + name);
    *off = b->off; *size = b->bytes;
    return MAF_OK;
}

extern "C" int maf_plan_ws_layout(const maf_plan* pl, int64_t* offsets, int64_t* sizes, int32_t cap) {
    if (!pl) return 0;
    const int n = (int)pl->ws_layout.size();
    for (int i = 0; i < n && i < cap; ++i) { if (offsets) offsets[i] = pl->ws_layout[i].off; if (sizes) sizes[i] = pl->ws_layout[i].bytes; }
    return n;
}

extern "C" int maf_set_noise_clips(maf_plan* pl, const int64_t* offsets, const int32_t* lengths) {
    if (!pl || !offsets || !lengths) return fail(MAF_ERR_INVALID, "maf_set_noise_clips: bad argument");
    CU(cudaSetDevice(pl->ctx->device));
    std::vector<Clip> cl(pl->n_items);
    for (int i = 0; i < pl->n_items; ++i) {
        if (lengths[i] <= 0 || offsets[i] < 0) return fail(MAF_ERR_INVALID, "maf_set_noise_clips: empty clip");
        if (lengths[i] <= NFFT - HOP) return fail(MAF_ERR_INVALID, "maf_set_noise_clips: noise clip of <= 768 samples (scipy.signal.stft: noverlap must be less than nperseg)");
        const int len = std::min<int>(lengths[i], CHUNK);
        pl->seg_fixed[2 * i] = 0; pl->seg_fixed[2 * i + 1] = lengths[i];      // reported range: the whole external clip
        cl[i].start = offsets[i]; cl[i].len = len; cl[i].n_frames = clip_frames(len);
    }
    CU(cudaMemcpy(pl->d_clips_fixed, cl.data(), sizeof(Clip) * pl->n_items, cudaMemcpyHostToDevice));
    pl->have_noise_clips = true;
    return MAF_OK;
}

// ---- helpers ----------------------------------------------------------------------------------------------------
template <typename T> static T* ws_ptr(void* ws, const WsBuf& b) { return reinterpret_cast<T*>(static_cast<unsigned char*>(ws) + b.off); }

__global__ void fill_int_kernel(int* p, int n, int v) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }
__global__ void zero_bytes_kernel(unsigned char* p, size_t bytes) {
    const size_t n16 = bytes / 16, gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = gtid; i < n16; i += (size_t)gridDim.x * blockDim.x) reinterpret_cast<uint4*>(p)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (gtid < bytes - 16 * n16) p[16 * n16 + gtid] = 0;
}
__global__ void copy_bytes_kernel(unsigned char* dst, const unsigned char* src, size_t bytes) {
    const size_t n16 = bytes / 16, gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = gtid; i < n16; i += (size_t)gridDim.x * blockDim.x) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    if (gtid < bytes - 16 * n16) dst[16 * n16 + gtid] = src[16 * n16 + gtid];
}
namespace maf {
static int fill_grid(size_t bytes) { const size_t b = (bytes / 16 + 255) / 256; return (int)(b < 1 ? 1 : b > 1184 ? 1184 : b); }
void launch_zero_bytes(void* p, size_t bytes, cudaStream_t st) {
    if (bytes) zero_bytes_kernel<<<fill_grid(bytes), 256, 0, st>>>(static_cast<unsigned char*>(p), bytes);
}
void launch_copy_bytes(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (!bytes) return;
    if ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) { cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st); return; }
    copy_bytes_kernel<<<fill_grid(bytes), 256, 0, st>>>(static_cast<unsigned char*>(dst), static_cast<const unsigned char*>(src), bytes);
}
}  // namespace maf
__global__ void abs_max_kernel(const int16_t* __restrict__ x, const Item* __restrict__ items, int* __restrict__ peak) {
    const Item it = items[blockIdx.y];
    int pk = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < it.own_len; i += (long long)gridDim.x * blockDim.x) {
        const int s = x[it.den_off + i];
        pk = max(pk, s < 0 ? -s : s);
    }
    for (int o = 16; o > 0; o >>= 1) pk = max(pk, __shfl_xor_sync(0xffffffffu, pk, o));
    if ((threadIdx.x & 31) == 0 && pk > 0) atomicMax(&peak[blockIdx.y], pk);
}
__global__ void gather_lufs_kernel(const Gate* __restrict__ g, int n, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = g[i].lufs;
}

// Small host structs reach the device as a KERNEL PARAMETER (copied at launch / captured by value in a CUDA graph) instead of a
// cudaMemcpyAsync from a stack variable, which a graph would replay from a dead host address.
struct PutBlob { unsigned char b[256]; };
__global__ void put_blob_kernel(PutBlob v, unsigned char* dst, int n) { if ((int)threadIdx.x < n) dst[threadIdx.x] = v.b[threadIdx.x]; }
static void launch_put(void* dst, const void* src, size_t n, cudaStream_t st) {
    PutBlob v; memset(&v, 0, sizeof(v)); memcpy(v.b, src, n < sizeof(v.b) ? n : sizeof(v.b));
    put_blob_kernel<<<1, 256, 0, st>>>(v, static_cast<unsigned char*>(dst), (int)n);
}

static const MelBank& pick_mel(const maf_plan* pl) { return pl->p.n_mels == 128 ? pl->ctx->mel128 : pl->ctx->mel80; }


// ---- per-stage timing ---------------------------------------------------------------------------------------------
enum { ST_BEGIN = 0, ST_NOISE, ST_STFT, ST_NSMASK, ST_SMOOTH, ST_ISTFT, ST_KWEIGHT, ST_GATE, ST_LOGMEL, ST_MELFIN, ST_COUNT };
static inline void mark(maf_plan* pl, int boundary, cudaStream_t st) {
    if (!pl->prof_steps) return;
    const int step = (int)(pl->prof_calls % pl->prof_steps);
    cudaEventRecord(pl->prof_ev[(size_t)step * ST_COUNT + boundary], st);
}

extern "C" int maf_plan_profile(maf_plan* pl, int32_t max_steps) {
    if (!pl || max_steps < 0) return fail(MAF_ERR_INVALID, "maf_plan_profile: bad argument");
    CU(cudaSetDevice(pl->ctx->device));
    for (cudaEvent_t e : pl->prof_ev) cudaEventDestroy(e);
    pl->prof_ev.clear(); pl->prof_steps = 0; pl->prof_calls = 0;
    for (int i = 0; i < max_steps * ST_COUNT; ++i) { cudaEvent_t e; CU(cudaEventCreate(&e)); pl->prof_ev.push_back(e); }
    pl->prof_steps = max_steps;
    return MAF_OK;
}

// elapsed ms of stage `stage` (1..9: noise, stft, ns_mask, smooth, istft, kweight, gate, logmel, mel_finalize) in the
// `step`-th most recent profiled maf_process call window (caller must have synchronised the stream)
extern "C" int maf_plan_stage_ms(maf_plan* pl, int32_t step, int32_t stage, float* ms) {
    if (!pl || !ms || !pl->prof_steps || step < 0 || step >= pl->prof_steps || stage < 1 || stage >= ST_COUNT)
        return fail(MAF_ERR_INVALID, "maf_plan_stage_ms: bad argument");
    CU(cudaEventElapsedTime(ms, pl->prof_ev[(size_t)step * ST_COUNT + stage - 1], pl->prof_ev[(size_t)step * ST_COUNT + stage]));
    return MAF_OK;
}

// S0 + thresholds + K1..K3.  `den_dst` receives the truncated int16 signal (workspace den16 layout, or the caller's
// output layout when to_out), the per-item peak lands in the workspace.
static int run_denoise(maf_plan* pl, const int16_t* pcm, const int16_t* noise_pcm, void* ws, int16_t* den_dst, int to_out, float* out_f32,
                       cudaStream_t st) {
    const maf_params& p = pl->p;
    const int n = pl->n_items;
    if (!pl->clip_err.empty()) return fail(MAF_ERR_INVALID, pl->clip_err);
    Clip* clips = ws_ptr<Clip>(ws, pl->clips);
    mark(pl, ST_BEGIN, st);
    launch_zero_bytes(ws_ptr<int>(ws, pl->peak), 4LL * n, st);
    pl->launches += 1;
    if (p.stationary) {
        const int16_t* noise_src = pcm;
        const Clip* use_clips = clips;
        if (p.external_noise) {
            if (!pl->have_noise_clips || !noise_pcm) return fail(MAF_ERR_INVALID, "external_noise set but no noise clips / noise pcm supplied");
            noise_src = noise_pcm; use_clips = pl->d_clips_fixed;
            if (!pl->sharded) CU(cudaMemcpyAsync(ws_ptr<int>(ws, pl->seg), pl->seg_fixed.data(), 8LL * n, cudaMemcpyHostToDevice, st));
        } else if (p.auto_detect_noise) {
            if (pl->sharded) return fail(MAF_ERR_INVALID, "sharded plans need external noise clips (run maf_frame_stats / maf_select_clip first)");
            launch_frame_stats(pcm, pl->d_items, n, pl->max_sframes, pl->frame_len, pl->frame_hop, ws_ptr<float>(ws, pl->E), ws_ptr<int>(ws, pl->ZC), st);
            const int fallback = (int)(p.sample_rate * (double)p.noise_sample_duration_s);
            if (pl->max_sframes > SEL_GRID_MIN_FRAMES && n <= 64) {
                // long recordings: grid-wide co-operative selection per item (a single CTA would need ~1 ms per hour of audio)
                for (int i = 0; i < n; ++i) {
                    const Item& hi = pl->items[i];
                    if (hi.n_sframes > SEL_GRID_MIN_FRAMES)
                        launch_select_clip_grid(ws_ptr<float>(ws, pl->E) + hi.sframe_off, ws_ptr<int>(ws, pl->ZC) + hi.sframe_off, hi.n_sframes, hi,
                                                pl->frame_len, pl->frame_hop, p.sample_rate, fallback, ws_ptr<void>(ws, pl->selws), clips + i,
                                                ws_ptr<int>(ws, pl->seg) + 2 * i, st);
                    else
                        launch_select_clip(ws_ptr<float>(ws, pl->E), ws_ptr<int>(ws, pl->ZC), pl->d_foff + i, pl->d_fcount + i, pl->d_items + i, 1,
                                           pl->frame_len, pl->frame_hop, p.sample_rate, fallback, clips + i, ws_ptr<int>(ws, pl->seg) + 2 * i, st);
                    pl->launches += 1;
                }
                pl->launches += 1;
            } else {
                launch_select_clip(ws_ptr<float>(ws, pl->E), ws_ptr<int>(ws, pl->ZC), pl->d_foff, pl->d_fcount, pl->d_items, n, pl->frame_len,
                                   pl->frame_hop, p.sample_rate, fallback, clips, ws_ptr<int>(ws, pl->seg), st);
                pl->launches += 2;
            }
        } else {
            if (pl->sharded) return fail(MAF_ERR_INVALID, "sharded plans need external noise clips");
            use_clips = pl->d_clips_fixed;
            CU(cudaMemcpyAsync(ws_ptr<int>(ws, pl->seg), pl->seg_fixed.data(), 8LL * n, cudaMemcpyHostToDevice, st));
        }
        launch_noise_threshold(noise_src, use_clips, n, pl->n_slices, pl->ctx->c, ws_ptr<unsigned>(ws, pl->nmax),
                               ws_ptr<double>(ws, pl->partial), ws_ptr<float>(ws, pl->T2), ws_ptr<float>(ws, pl->T2hi), 1.5, st);
        pl->launches += 3;
        launch_zero_bytes(ws_ptr<uint8_t>(ws, pl->rowflag), pl->rowflag.bytes, st);
        pl->launches += 2;                                   // + the nmax clear inside launch_noise_threshold
    }
    mark(pl, ST_NOISE, st);
    if (launch_stft(pcm, pl->d_items, pl->d_chunks, pl->n_chunks, pl->max_chunk_frames, pl->ctx->c, p.stationary, ws_ptr<float>(ws, pl->T2),
                    ws_ptr<float>(ws, pl->T2hi), ws_ptr<c32>(ws, pl->S), ws_ptr<uint32_t>(ws, pl->bits), ws_ptr<uint8_t>(ws, pl->rowflag),
                    (const NsParams*)pl->d_ns, ws_ptr<float>(ws, pl->P), ws_ptr<float>(ws, pl->Q), st))
        return fail(MAF_ERR_UNSUPPORTED, "STFT tile count exceeds 2^31");
    pl->launches += 1;
    mark(pl, ST_STFT, st);
    if (!p.stationary) {
        launch_ns_mask(ws_ptr<c32>(ws, pl->S), pl->d_chunks, pl->n_chunks, pl->max_chunk_frames, pl->ns, ws_ptr<float>(ws, pl->P),
                       ws_ptr<float>(ws, pl->Q), ws_ptr<float>(ws, pl->Yend), ws_ptr<float>(ws, pl->Zend), p.prop_decrease, pl->taps,
                       ws_ptr<float>(ws, pl->M), st);
        pl->launches += 2;
    }
    mark(pl, ST_NSMASK, st);
    if (p.stationary) {
        launch_smooth_mask(ws_ptr<uint32_t>(ws, pl->bits), ws_ptr<uint8_t>(ws, pl->rowflag), pl->d_chunks, pl->n_chunks,
                           pl->max_chunk_frames, p.prop_decrease, pl->taps, ws_ptr<float>(ws, pl->M), st);
        pl->launches += 1;
    }
    mark(pl, ST_SMOOTH, st);
    launch_istft(ws_ptr<c32>(ws, pl->S), ws_ptr<float>(ws, pl->M), pl->d_items, pl->d_chunks, pl->n_chunks, pl->ctx->c,
                 den_dst, to_out, out_f32, ws_ptr<int>(ws, pl->peak), st);
    pl->launches += 1;
    mark(pl, ST_ISTFT, st);
    CU(cudaGetLastError());
    return MAF_OK;
}

static int run_loudness_stats(maf_plan* pl, const int16_t* src, const Item* items, int peak_normalize, void* ws, cudaStream_t st,
                              int16_t* norm16 = nullptr) {
    if (pl->too_short) return fail(MAF_ERR_TOO_SHORT, "Audio must have length greater than the block size");
    const int n = pl->n_items;
    launch_zero_bytes(ws_ptr<int>(ws, pl->peak_norm), 4LL * n, st);
    pl->launches += 1;
    if (peak_normalize) { launch_peak_factor(ws_ptr<int>(ws, pl->peak), n, pl->target_peak, ws_ptr<PeakNorm>(ws, pl->pnorm), st); pl->launches += 1; }
    launch_kweight_energy(src, items, n, pl->max_kgroups, pl->tile_len, pl->kgeo, pl->kw, (const KwScan*)pl->d_kscan, peak_normalize, ws_ptr<PeakNorm>(ws, pl->pnorm),
                          ws_ptr<double>(ws, pl->sub), ws_ptr<int>(ws, pl->peak_norm), peak_normalize ? norm16 : nullptr, st);
    pl->launches += 1;
    mark(pl, ST_KWEIGHT, st);
    CU(cudaGetLastError());
    return MAF_OK;
}

static int run_gate(maf_plan* pl, const Item* items, void* ws, cudaStream_t st) {
    int max_blocks = 0;
    for (const Item& it : pl->items) max_blocks = std::max(max_blocks, it.n_blocks);
    launch_gate(ws_ptr<double>(ws, pl->sub), items, pl->n_items, pl->inv_block, (double)pl->p.target_lufs, ws_ptr<int>(ws, pl->peak_norm),
                ws_ptr<Gate>(ws, pl->gates), max_blocks, ws_ptr<void>(ws, pl->selws), st);
    pl->launches += 1;
    mark(pl, ST_GATE, st);
    CU(cudaGetLastError());
    return MAF_OK;
}

static int run_logmel(maf_plan* pl, const int16_t* src, const Item* items, int peak_normalize, const Gate* gates, void* ws, int16_t* out16,
                      float* mel, bool finalize, cudaStream_t st, const int16_t* norm16 = nullptr) {
    const int n = pl->n_items;
    fill_int_kernel<<<(n + 255) / 256, 256, 0, st>>>(ws_ptr<int>(ws, pl->mel_max), n, (int)0x80000001);
    {
        const int nt = n * std::max(1, pl->max_mel_tiles);
        fill_int_kernel<<<(nt + 255) / 256, 256, 0, st>>>(ws_ptr<int>(ws, pl->mel_tmin), nt, (int)0x7fffffff);
    }
    if (launch_logmel(src, items, n, pl->max_mel_tiles, pl->ctx->c, pick_mel(pl), pl->p.mel_padding, peak_normalize,
                      ws_ptr<PeakNorm>(ws, pl->pnorm), peak_normalize ? norm16 : nullptr, gates, out16, mel, ws_ptr<int>(ws, pl->mel_max), ws_ptr<int>(ws, pl->mel_tmin), st))
        return fail(MAF_ERR_UNSUPPORTED, "log-mel tile count exceeds 2^31");
    pl->launches += 3;
    mark(pl, ST_LOGMEL, st);
    if (finalize) {
        launch_mel_finalize(mel, items, n, pl->p.n_mels, std::max(1, pl->max_mel_tiles), ws_ptr<int>(ws, pl->mel_max), ws_ptr<int>(ws, pl->mel_tmin), st);
        pl->launches += 1;
    }
    mark(pl, ST_MELFIN, st);
    CU(cudaGetLastError());
    return MAF_OK;
}

extern "C" int maf_process(maf_plan* pl, const int16_t* pcm, const int16_t* noise_pcm, void* ws, int16_t* out_pcm, float* mel,
                           double* lufs, int32_t* clip, void* stream) {
    if (!pl || !pcm || !ws || !out_pcm) return fail(MAF_ERR_INVALID, "maf_process: NULL argument");
    if (pl->p.n_mels && !mel) return fail(MAF_ERR_INVALID, "maf_process: mel_dev is NULL but n_mels > 0");
    if (pl->sharded) return fail(MAF_ERR_INVALID, "maf_process: sharded plans use the maf_sharded_* stages");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(pl->ctx->device));
    int16_t* den16 = ws_ptr<int16_t>(ws, pl->den16);
    int rc = run_denoise(pl, pcm, noise_pcm, ws, den16, 0, nullptr, st);
    if (rc) return rc;
    rc = run_loudness_stats(pl, den16, pl->d_items, pl->p.peak_normalize, ws, st, ws_ptr<int16_t>(ws, pl->norm16));
    if (rc) return rc;
    rc = run_gate(pl, pl->d_items, ws, st);
    if (rc) return rc;
    const Gate* gates = ws_ptr<Gate>(ws, pl->gates);
    if (pl->p.n_mels) {
        rc = run_logmel(pl, den16, pl->d_items, pl->p.peak_normalize, gates, ws, out_pcm, mel, true, st, ws_ptr<int16_t>(ws, pl->norm16));
        if (rc) return rc;
    } else {
        launch_apply_gain(den16, pl->d_items, pl->n_items, pl->max_own, pl->p.peak_normalize, ws_ptr<PeakNorm>(ws, pl->pnorm), gates, 1,
                          out_pcm, st);
        pl->launches += 1;
    }
    if (lufs) { gather_lufs_kernel<<<(pl->n_items + 255) / 256, 256, 0, st>>>(gates, pl->n_items, lufs); pl->launches += 1; }
    if (clip && pl->p.stationary) { launch_copy_bytes(clip, ws_ptr<int>(ws, pl->seg), 8LL * pl->n_items, st); pl->launches += 1; }
    CU(cudaGetLastError());
    if (pl->prof_steps) pl->prof_calls += 1;
    return MAF_OK;
}

extern "C" int maf_process_host(maf_plan* pl, const int16_t* pcm_host, const int16_t* noise_host, int64_t noise_len, int16_t* out_host,
                                float* mel_host, double* lufs_host, int32_t* clip_host) {
    if (!pl || !pcm_host || !out_host) return fail(MAF_ERR_INVALID, "maf_process_host: NULL argument");
    CU(cudaSetDevice(pl->ctx->device));
    if (!pl->host_ready) {
        // allocate into locals; the plan takes them only when every allocation succeeded (a partial set would leave NULL
        // buffers behind a "ready" flag and fault on the next call)
        cudaStream_t hs = nullptr; void* bufs[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        const size_t sizes[6] = {(size_t)pl->ws_bytes, (size_t)(2 * std::max<long long>(8, pl->total_pcm)), (size_t)(2 * std::max<long long>(8, pl->total_out)),
                                 (size_t)(4 * std::max<long long>(4, pl->total_mel)), (size_t)(8LL * pl->n_items), (size_t)(8LL * pl->n_items)};
        cudaError_t e = cudaStreamCreateWithFlags(&hs, cudaStreamNonBlocking);
        for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaMalloc(&bufs[i], sizes[i]);
        if (e == cudaSuccess) e = cudaMemsetAsync(bufs[5], 0, sizes[5], hs);
        if (e != cudaSuccess) {
            for (void* b : bufs) if (b) cudaFree(b);
            if (hs) cudaStreamDestroy(hs);
            return fail(MAF_ERR_CUDA, std::string("maf_process_host: staging allocation failed: ") + cudaGetErrorString(e));
        }
        pl->hstream = hs; pl->h_ws = bufs[0]; pl->h_pcm = (int16_t*)bufs[1]; pl->h_out = (int16_t*)bufs[2]; pl->h_mel = (float*)bufs[3];
        pl->h_lufs = (double*)bufs[4]; pl->h_clip = (int*)bufs[5];
        pl->host_ready = true;
    }
    cudaStream_t st = pl->hstream;
    if (noise_host && noise_len > 0) {
        if (noise_len > pl->h_noise_cap) { cudaFree(pl->h_noise); pl->h_noise = nullptr; CU(cudaMalloc((void**)&pl->h_noise, 2 * noise_len)); pl->h_noise_cap = noise_len; }
        CU(cudaMemcpyAsync(pl->h_noise, noise_host, 2 * noise_len, cudaMemcpyHostToDevice, st));
    }
    CU(cudaMemcpyAsync(pl->h_pcm, pcm_host, 2 * pl->total_pcm, cudaMemcpyHostToDevice, st));
    int rc = maf_process(pl, pl->h_pcm, (noise_host && noise_len > 0) ? pl->h_noise : nullptr, pl->h_ws, pl->h_out, pl->p.n_mels ? pl->h_mel : nullptr,
                         pl->h_lufs, pl->h_clip, st);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out_host, pl->h_out, 2 * pl->total_out, cudaMemcpyDeviceToHost, st));
    if (mel_host && pl->p.n_mels) CU(cudaMemcpyAsync(mel_host, pl->h_mel, 4 * pl->total_mel, cudaMemcpyDeviceToHost, st));
    if (lufs_host) CU(cudaMemcpyAsync(lufs_host, pl->h_lufs, 8LL * pl->n_items, cudaMemcpyDeviceToHost, st));
    if (clip_host) CU(cudaMemcpyAsync(clip_host, pl->h_clip, 8LL * pl->n_items, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return MAF_OK;
}

// ---- stage entry points ---------------------------------------------------------------------------------------
extern "C" int maf_denoise(maf_plan* pl, const int16_t* pcm, const int16_t* noise_pcm, void* ws, int16_t* out_pcm, float* out_f32,
                           int32_t* clip, void* stream) {
    if (!pl || !pcm || !ws || !out_pcm) return fail(MAF_ERR_INVALID, "maf_denoise: NULL argument");
    if (pl->sharded) return fail(MAF_ERR_INVALID, "maf_denoise: sharded plans use maf_sharded_denoise");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(pl->ctx->device));
    int rc = run_denoise(pl, pcm, noise_pcm, ws, out_pcm, 1, out_f32, st);
    if (rc) return rc;
    if (clip && pl->p.stationary) { launch_copy_bytes(clip, ws_ptr<int>(ws, pl->seg), 8LL * pl->n_items, st); pl->launches += 1; }
    return MAF_OK;
}

extern "C" int maf_peak_normalize(maf_plan* pl, const int16_t* pcm, void* ws, int16_t* out_pcm, void* stream) {
    if (!pl || !pcm || !ws || !out_pcm) return fail(MAF_ERR_INVALID, "maf_peak_normalize: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(pl->ctx->device));
    launch_zero_bytes(ws_ptr<int>(ws, pl->peak), 4LL * pl->n_items, st);
    long long bx = std::max<long long>(1, std::min<long long>(1024, (pl->max_own + 4095) / 4096));
    dim3 grid((unsigned)bx, pl->n_items);
    abs_max_kernel<<<grid, 256, 0, st>>>(pcm, pl->d_items_io, ws_ptr<int>(ws, pl->peak));
    launch_peak_factor(ws_ptr<int>(ws, pl->peak), pl->n_items, pl->target_peak, ws_ptr<PeakNorm>(ws, pl->pnorm), st);
    launch_apply_gain(pcm, pl->d_items_io, pl->n_items, pl->max_own, 1, ws_ptr<PeakNorm>(ws, pl->pnorm), nullptr, 0, out_pcm, st);
    pl->launches += 2;
    CU(cudaGetLastError());
    return MAF_OK;
}

extern "C" int maf_loudness(maf_plan* pl, const int16_t* pcm, void* ws, int16_t* out_pcm, double* lufs, void* stream) {
    if (!pl || !pcm || !ws || !out_pcm) return fail(MAF_ERR_INVALID, "maf_loudness: NULL argument");
    if (pl->sharded) return fail(MAF_ERR_INVALID, "maf_loudness: sharded plans use maf_sharded_energy / maf_gate");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(pl->ctx->device));
    int rc = run_loudness_stats(pl, pcm, pl->d_items_io, 0, ws, st);
    if (rc) return rc;
    rc = run_gate(pl, pl->d_items_io, ws, st);
    if (rc) return rc;
    const Gate* gates = ws_ptr<Gate>(ws, pl->gates);
    launch_apply_gain(pcm, pl->d_items_io, pl->n_items, pl->max_own, 0, nullptr, gates, 1, out_pcm, st);
    pl->launches += 1;
    if (lufs) { gather_lufs_kernel<<<(pl->n_items + 255) / 256, 256, 0, st>>>(gates, pl->n_items, lufs); pl->launches += 1; }
    CU(cudaGetLastError());
    return MAF_OK;
}

// Multi-channel normalize_loudness (preprocessor.py:256-282): channel mean in float32, BS.1770 loudness of the mean, gain /
// limiter / int16 of the MONO signal.  Workspace layout: Item | Gate | absmax, peak | sub-block energies | float mean.
static long long loud_il_layout(long long n_frames, int tile_len, long long* o_gate, long long* o_max, long long* o_sub, long long* o_m) {
    const long long n_tiles = (n_frames + tile_len - 1) / tile_len;
    long long cur = align_up((long long)sizeof(Item), 256);
    *o_gate = cur; cur = align_up(cur + (long long)sizeof(Gate), 256);
    *o_max = cur; cur = align_up(cur + 16, 256);
    *o_sub = cur; cur = align_up(cur + 8 * std::max<long long>(1, n_tiles), 256);
    *o_m = cur; cur = align_up(cur + 4 * n_frames, 256);
    return cur;
}

extern "C" int64_t maf_loudness_interleaved_ws(int64_t n_frames, int32_t sample_rate) {
    if (n_frames <= 0 || sample_rate <= 0) return 0;
    long long a, b, c, d;
    return loud_il_layout(n_frames, (int)(0.4 * 0.25 * (double)sample_rate), &a, &b, &c, &d);
}

extern "C" int maf_loudness_interleaved(maf_ctx* ctx, const int16_t* pcm, int64_t n_frames, int32_t channels, int32_t sr, float target_lufs,
                                        void* ws, int16_t* out, double* lufs, void* stream) {
    if (!ctx || !pcm || !ws || !out || n_frames <= 0) return fail(MAF_ERR_INVALID, "maf_loudness_interleaved: bad argument");
    if (channels < 1 || channels > 128) return fail(MAF_ERR_UNSUPPORTED, "maf_loudness_interleaved: 1..128 channels");
    if (sr < 4000 || sr > 96000) return fail(MAF_ERR_INVALID, "sample_rate out of range");
    if ((double)n_frames < 0.4 * (double)sr) return fail(MAF_ERR_TOO_SHORT, "Audio must have length greater than the block size");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int tile_len = (int)(0.4 * 0.25 * (double)sr);
    const double T = (double)n_frames / (double)sr;
    const long long nb = (long long)(nearbyint((T - 0.4) / (0.4 * 0.25)) + 1);
    for (long long j = 0; j < nb; ++j) {                           // the sub-block decomposition needs pyloudnorm's bounds to be regular
        const long long l = (long long)(0.4 * (j * 0.25) * (double)sr), u = (long long)(0.4 * (j * 0.25 + 1) * (double)sr);
        if (l != j * tile_len || u != l + 4LL * tile_len) return fail(MAF_ERR_UNSUPPORTED, "irregular loudness block grid at this rate / length");
    }
    long long o_gate, o_max, o_sub, o_m;
    loud_il_layout(n_frames, tile_len, &o_gate, &o_max, &o_sub, &o_m);
    unsigned char* b = static_cast<unsigned char*>(ws);
    Item it{};
    it.ktile_first = 0; it.n_ktiles = (int)((n_frames + tile_len - 1) / tile_len); it.n_blocks = (int)nb;
    CU(cudaMemcpyAsync(b, &it, sizeof(it), cudaMemcpyHostToDevice, st));
    KWeight kw; design_kweight((double)sr, kw);
    launch_loudness_interleaved(pcm, n_frames, channels, tile_len, it.n_ktiles, (const Item*)b, kw, 1.0 / (0.4 * (double)sr), (double)target_lufs,
                                (float*)(b + o_m), (double*)(b + o_sub), (unsigned*)(b + o_max), (int*)(b + o_max + 4), (Gate*)(b + o_gate), out, st);
    if (lufs) gather_lufs_kernel<<<1, 32, 0, st>>>((const Gate*)(b + o_gate), 1, lufs);
    CU(cudaGetLastError());
    return MAF_OK;
}

extern "C" int maf_logmel(maf_plan* pl, const int16_t* pcm, void* ws, float* mel, void* stream) {
    if (!pl || !pcm || !ws || !mel) return fail(MAF_ERR_INVALID, "maf_logmel: NULL argument");
    if (!pl->p.n_mels) return fail(MAF_ERR_INVALID, "maf_logmel: plan was created with n_mels = 0");
    if (pl->sharded) return fail(MAF_ERR_INVALID, "maf_logmel: sharded plans use maf_sharded_logmel");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(pl->ctx->device));
    return run_logmel(pl, pcm, pl->d_items_io, 0, nullptr, ws, nullptr, mel, true, st);
}

// ---- time-sharded pieces ----------------------------------------------------------------------------------------
extern "C" int maf_frame_stats(maf_plan* pl, const int16_t* pcm, float* energy, int32_t* zc, void* stream) {
    if (!pl || !pcm || !energy || !zc) return fail(MAF_ERR_INVALID, "maf_frame_stats: NULL argument");
    CU(cudaSetDevice(pl->ctx->device));
    launch_frame_stats(pcm, pl->d_items, pl->n_items, pl->max_sframes, pl->frame_len, pl->frame_hop, energy, zc, (cudaStream_t)stream);
    pl->launches += 1;
    CU(cudaGetLastError());
    return MAF_OK;
}

extern "C" int maf_select_clip(maf_ctx* ctx, const float* energy, const int32_t* zc, int64_t n_frames, int64_t n_global, int32_t sr,
                               float dur, int32_t* seg, void* stream) {
    if (!ctx || !energy || !zc || !seg) return fail(MAF_ERR_INVALID, "maf_select_clip: NULL argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    struct { Item it; long long foff; int fcount; int pad; Clip clip; } h{};
    h.it.n_global = n_global; h.it.own_len = n_global; h.it.buf_len = n_global;
    h.foff = 0; h.fcount = (int)n_frames;
    unsigned char* d = (unsigned char*)ctx->scratch;
    static_assert(sizeof(h) <= 256, "selection header exceeds the put blob");
    launch_put(d, &h, sizeof(h), st);
    const int frame_len = (int)(sr * 0.025), hop = (int)(sr * 0.010);
    if (n_frames > SEL_GRID_MIN_FRAMES)
        launch_select_clip_grid(energy, zc, (int)n_frames, h.it, frame_len, hop, sr, (int)(sr * (double)dur), d + 1024,
                                (Clip*)(d + offsetof(decltype(h), clip)), seg, st);
    else
        launch_select_clip(energy, zc, (const long long*)(d + offsetof(decltype(h), foff)), (const int*)(d + offsetof(decltype(h), fcount)),
                           (const Item*)d, 1, frame_len, hop, sr, (int)(sr * (double)dur), (Clip*)(d + offsetof(decltype(h), clip)), seg, st);
    CU(cudaGetLastError());
    return MAF_OK;
}

// Sharded noise clip without a host round trip: the selected range stays on the device (seg = {start, end} in global samples).
__global__ void clip_contrib_kernel(const int16_t* __restrict__ pcm, Item it, const int* __restrict__ seg, int16_t* __restrict__ clip) {
    const long long gs = seg[0];
    long long len = (long long)seg[1] - gs; if (len > CHUNK) len = CHUNK;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < CHUNK; i += gridDim.x * blockDim.x) {
        const long long g = gs + i;
        int v = 0;
        if (i < len && g >= it.g0 && g < it.g0 + it.own_len) v = pcm[it.pcm_off + it.origin + (g - it.g0)];
        clip[i] = (int16_t)v;
    }
}
__global__ void set_clip_dev_kernel(const int* __restrict__ seg, Clip* __restrict__ clips) {
    long long len = (long long)seg[1] - seg[0]; if (len > CHUNK) len = CHUNK; if (len < 0) len = 0;
    Clip c; c.start = 0; c.len = (int)len; c.n_frames = clip_frames((int)len);
    clips[0] = c;
}

extern "C" int maf_sharded_clip_contrib(maf_plan* pl, const int16_t* pcm, const int32_t* seg, int16_t* clip, void* stream) {
    if (!pl || !pcm || !seg || !clip || pl->n_items != 1) return fail(MAF_ERR_INVALID, "maf_sharded_clip_contrib: bad argument");
    CU(cudaSetDevice(pl->ctx->device));
    clip_contrib_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(pcm, pl->items[0], seg, clip);
    CU(cudaGetLastError());
    return MAF_OK;
}

extern "C" int maf_set_noise_clips_dev(maf_plan* pl, const int32_t* seg, void* stream) {
    if (!pl || !seg || pl->n_items != 1) return fail(MAF_ERR_INVALID, "maf_set_noise_clips_dev: bad argument");
    CU(cudaSetDevice(pl->ctx->device));
    set_clip_dev_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(seg, pl->d_clips_fixed);
    pl->have_noise_clips = true;
    CU(cudaGetLastError());
    return MAF_OK;
}

// ---- fused exchanges of the time-sharded mode: one all-gather per barrier instead of one collective per statistic -----------
// generic: dst = concat over ranks r of src[r * stride + offset : ... + counts[r]] (elements of elem_bytes = 4 or 8)
template <typename T>
__global__ void gather_compact_kernel(const T* __restrict__ src, long long stride, long long offset, const int* __restrict__ counts, T* __restrict__ dst) {
    const int r = blockIdx.y;
    long long pre = 0;
    for (int q = 0; q < r; ++q) pre += counts[q];
    const int cnt = counts[r];
    const T* s = src + (long long)r * stride + offset;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) dst[pre + i] = s[i];
}

extern "C" int maf_gather_compact(maf_ctx* ctx, const void* src, int32_t elem_bytes, int32_t world, int64_t stride, int64_t offset,
                                  const int32_t* counts_dev, void* dst, void* stream) {
    if (!ctx || !src || !dst || !counts_dev || world <= 0 || world > 65535 || (elem_bytes != 4 && elem_bytes != 8))
        return fail(MAF_ERR_INVALID, "maf_gather_compact: bad argument");
    CU(cudaSetDevice(ctx->device));
    dim3 grid(64, world);
    if (elem_bytes == 4) gather_compact_kernel<uint32_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint32_t*)src, stride, offset, counts_dev, (uint32_t*)dst);
    else gather_compact_kernel<unsigned long long><<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned long long*)src, stride, offset, counts_dev, (unsigned long long*)dst);
    CU(cudaGetLastError());
    return MAF_OK;
}

// edges: [head KW_WARM int16 | tail KW_WARM int16 | peak int32 | pad] = MAF_EDGE_BYTES per rank
constexpr int EDGE_BYTES = 4 * KW_WARM + 16;
__global__ void pack_edges_kernel(const int16_t* __restrict__ den, long long own0, long long n, const int* __restrict__ peak, unsigned char* __restrict__ out) {
    int16_t* o = reinterpret_cast<int16_t*>(out);
    const long long k = n < KW_WARM ? n : KW_WARM;
    for (int i = threadIdx.x; i < KW_WARM; i += blockDim.x) {
        o[i] = i < k ? den[own0 + i] : (int16_t)0;                                   // first own samples
        o[KW_WARM + i] = i >= KW_WARM - k ? den[own0 + n - KW_WARM + i] : (int16_t)0; // last own samples, right-aligned
    }
    if (threadIdx.x == 0) *reinterpret_cast<int*>(out + 4 * KW_WARM) = peak[0];
}
__global__ void unpack_edges_kernel(const unsigned char* __restrict__ gathered, int world, int rank, int16_t* __restrict__ den, long long own0, long long n,
                                    int* __restrict__ peak) {
    if (threadIdx.x == 0) {
        int pk = 0;
        for (int r = 0; r < world; ++r) pk = max(pk, *reinterpret_cast<const int*>(gathered + (size_t)r * EDGE_BYTES + 4 * KW_WARM));
        peak[0] = pk;
    }
    if (rank > 0) {                                                                  // left neighbour's tail in front of own sample 0
        const int16_t* t = reinterpret_cast<const int16_t*>(gathered + (size_t)(rank - 1) * EDGE_BYTES) + KW_WARM;
        for (int i = threadIdx.x; i < KW_WARM; i += blockDim.x) den[own0 - KW_WARM + i] = t[i];
    }
    if (rank < world - 1) {                                                          // right neighbour's head behind the last own sample
        const int16_t* h = reinterpret_cast<const int16_t*>(gathered + (size_t)(rank + 1) * EDGE_BYTES);
        for (int i = threadIdx.x; i < KW_WARM; i += blockDim.x) den[own0 + n + i] = h[i];
    }
}

extern "C" int maf_sharded_edge_bytes(void) { return EDGE_BYTES; }

extern "C" int maf_sharded_pack_edges(maf_plan* pl, void* ws, void* out, void* stream) {
    if (!pl || !ws || !out || pl->n_items != 1) return fail(MAF_ERR_INVALID, "maf_sharded_pack_edges: bad argument");
    CU(cudaSetDevice(pl->ctx->device));
    pack_edges_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(ws_ptr<int16_t>(ws, pl->den16), pl->items[0].den_off, pl->items[0].own_len,
                                                          ws_ptr<int>(ws, pl->peak), (unsigned char*)out);
    pl->launches += 1;
    CU(cudaGetLastError());
    return MAF_OK;
}

extern "C" int maf_sharded_unpack_edges(maf_plan* pl, void* ws, const void* gathered, int32_t world, int32_t rank, void* stream) {
    if (!pl || !ws || !gathered || pl->n_items != 1 || world <= 0 || rank < 0 || rank >= world)
        return fail(MAF_ERR_INVALID, "maf_sharded_unpack_edges: bad argument");
    CU(cudaSetDevice(pl->ctx->device));
    unpack_edges_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>((const unsigned char*)gathered, world, rank, ws_ptr<int16_t>(ws, pl->den16),
                                                            pl->items[0].den_off, pl->items[0].own_len, ws_ptr<int>(ws, pl->peak));
    pl->launches += 1;
    CU(cudaGetLastError());
    return MAF_OK;
}

// energies: [peak_norm as double | sub-block energies padded to max_tiles] per rank
__global__ void pack_energy_kernel(const double* __restrict__ sub, int n_tiles, int max_tiles, const int* __restrict__ peak_norm, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) out[0] = (double)peak_norm[0];
    if (i < max_tiles) out[1 + i] = i < n_tiles ? sub[i] : 0.0;
}
__global__ void unpack_peak_norm_kernel(const double* __restrict__ gathered, int world, long long stride, int* __restrict__ peak_norm) {
    int pk = 0;
    for (int r = 0; r < world; ++r) pk = max(pk, (int)gathered[(long long)r * stride]);
    peak_norm[0] = pk;
}

extern "C" int maf_sharded_pack_energy(maf_plan* pl, void* ws, double* out, int32_t max_tiles, void* stream) {
    if (!pl || !ws || !out || pl->n_items != 1 || max_tiles < pl->items[0].n_ktiles) return fail(MAF_ERR_INVALID, "maf_sharded_pack_energy: bad argument");
    CU(cudaSetDevice(pl->ctx->device));
    pack_energy_kernel<<<(max_tiles + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ws_ptr<double>(ws, pl->sub), pl->items[0].n_ktiles, max_tiles,
                                                                                 ws_ptr<int>(ws, pl->peak_norm), out);
    pl->launches += 1;
    CU(cudaGetLastError());
    return MAF_OK;
}

extern "C" int maf_sharded_unpack_energy(maf_plan* pl, void* ws, const double* gathered, int32_t world, int32_t max_tiles, const int32_t* counts_dev,
                                         double* sub_all, void* stream) {
    if (!pl || !ws || !gathered || !counts_dev || !sub_all || world <= 0) return fail(MAF_ERR_INVALID, "maf_sharded_unpack_energy: bad argument");
    CU(cudaSetDevice(pl->ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    unpack_peak_norm_kernel<<<1, 1, 0, st>>>(gathered, world, 1LL + max_tiles, ws_ptr<int>(ws, pl->peak_norm));
    dim3 grid(64, world);
    gather_compact_kernel<double><<<grid, 256, 0, st>>>(gathered, 1LL + max_tiles, 1, counts_dev, sub_all);
    pl->launches += 2;
    CU(cudaGetLastError());
    return MAF_OK;
}

extern "C" int maf_sharded_denoise(maf_plan* pl, const int16_t* pcm, const int16_t* noise_pcm, void* ws, void* stream) {
    if (!pl || !pcm || !ws) return fail(MAF_ERR_INVALID, "maf_sharded_denoise: NULL argument");
    CU(cudaSetDevice(pl->ctx->device));
    return run_denoise(pl, pcm, noise_pcm, ws, ws_ptr<int16_t>(ws, pl->den16), 0, nullptr, (cudaStream_t)stream);
}

extern "C" int maf_sharded_energy(maf_plan* pl, void* ws, void* stream) {
    if (!pl || !ws) return fail(MAF_ERR_INVALID, "maf_sharded_energy: NULL argument");
    CU(cudaSetDevice(pl->ctx->device));
    return run_loudness_stats(pl, ws_ptr<int16_t>(ws, pl->den16), pl->d_items, pl->p.peak_normalize, ws, (cudaStream_t)stream,
                              ws_ptr<int16_t>(ws, pl->norm16));
}

extern "C" int maf_gate(maf_ctx* ctx, const double* sub_energy, int64_t n_tiles, int64_t n_blocks, int32_t sr, float target_lufs,
                        const int32_t* peak_norm, void* gate, void* stream) {
    if (!ctx || !sub_energy || !peak_norm || !gate) return fail(MAF_ERR_INVALID, "maf_gate: NULL argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    Item it{};
    it.ktile_first = 0; it.n_ktiles = (int)n_tiles; it.n_blocks = (int)n_blocks;
    unsigned char* d = (unsigned char*)ctx->scratch + 512;
    static_assert(sizeof(Item) <= 256, "Item exceeds the put blob");
    launch_put(d, &it, sizeof(it), st);
    launch_gate(sub_energy, (const Item*)d, 1, 1.0 / (0.4 * (double)sr), (double)target_lufs, peak_norm, (Gate*)gate, (int)n_blocks,
                (unsigned char*)ctx->scratch + 1024, st);
    CU(cudaGetLastError());
    return MAF_OK;
}

extern "C" int maf_sharded_logmel(maf_plan* pl, void* ws, const void* gate, int16_t* out_pcm, float* mel, void* stream) {
    if (!pl || !ws || !gate || !mel) return fail(MAF_ERR_INVALID, "maf_sharded_logmel: NULL argument");
    if (pl->n_items != 1) return fail(MAF_ERR_INVALID, "maf_sharded_logmel: one shard per plan");
    CU(cudaSetDevice(pl->ctx->device));
    return run_logmel(pl, ws_ptr<int16_t>(ws, pl->den16), pl->d_items, pl->p.peak_normalize, (const Gate*)gate, ws, out_pcm, mel, false,
                      (cudaStream_t)stream, ws_ptr<int16_t>(ws, pl->norm16));
}

extern "C" int maf_mel_finalize(maf_plan* pl, void* ws, float* mel, void* stream) {
    if (!pl || !ws || !mel) return fail(MAF_ERR_INVALID, "maf_mel_finalize: NULL argument");
    CU(cudaSetDevice(pl->ctx->device));
    launch_mel_finalize(mel, pl->d_items, pl->n_items, pl->p.n_mels, std::max(1, pl->max_mel_tiles), ws_ptr<int>(ws, pl->mel_max), ws_ptr<int>(ws, pl->mel_tmin),
                        (cudaStream_t)stream);
    pl->launches += 1;
    CU(cudaGetLastError());
    return MAF_OK;
}

// ---- silence detection (next-row: remove_silence / detect_silence_segments) -------------------------------------------
extern "C" int maf_ms_energy(maf_ctx* ctx, const int16_t* pcm, int64_t n, int32_t sr, int64_t n_ms, uint64_t* bucket, uint64_t* total,
                             void* stream) {
    if (!ctx || !pcm || !bucket || !total || n <= 0 || n_ms < 0) return fail(MAF_ERR_INVALID, "maf_ms_energy: bad argument");
    CU(cudaSetDevice(ctx->device));
    launch_ms_energy(pcm, n, sr, n_ms, (unsigned long long*)bucket, (unsigned long long*)total, (cudaStream_t)stream);
    CU(cudaGetLastError());
    return MAF_OK;
}

extern "C" int maf_silence_layout(int64_t n_ms, int32_t window_ms, int64_t* ws_bytes, int64_t* result_bytes, int64_t* max_ranges,
                                  int64_t* off_silent, int64_t* off_map) {
    if (n_ms < 0 || window_ms <= 0) return fail(MAF_ERR_INVALID, "maf_silence_layout: bad argument");
    size_t a = 0, b = 0; long long cap = 0, os = 0, om = 0;
    silence_layout(n_ms, window_ms, &a, &b, &cap, &os, &om);
    if (ws_bytes) *ws_bytes = (int64_t)a;
    if (result_bytes) *result_bytes = (int64_t)b;
    if (max_ranges) *max_ranges = cap;
    if (off_silent) *off_silent = os;
    if (off_map) *off_map = om;
    return MAF_OK;
}

extern "C" int maf_silence(maf_ctx* ctx, const int16_t* pcm, int64_t n, int32_t sr, const uint64_t* bucket, int64_t n_ms, int32_t window_ms,
                           double thresh, int32_t margin_ms, void* ws, int16_t* out, int64_t out_cap, void* stream) {
    if (!ctx || !pcm || !bucket || !ws || n <= 0 || n_ms < 0 || window_ms <= 0 || margin_ms < 0 || sr <= 0)
        return fail(MAF_ERR_INVALID, "maf_silence: bad argument");
    if (n_ms > 2000000000LL) return fail(MAF_ERR_UNSUPPORTED, "maf_silence: recording longer than 2^31 ms");
    CU(cudaSetDevice(ctx->device));
    launch_silence(pcm, n, sr, (const unsigned long long*)bucket, n_ms, window_ms, thresh, margin_ms, ws, out, out_cap, (cudaStream_t)stream);
    CU(cudaGetLastError());
    return MAF_OK;
}

// ---- front door ------------------------------------------------------------------------------------------------
static long long gcd_ll(long long a, long long b) { while (b) { long long t = a % b; a = b; b = t; } return a; }

extern "C" int64_t maf_front_door_out_len(int64_t n_frames_in, int32_t in_rate, int32_t out_rate) {
    if (n_frames_in <= 0 || in_rate <= 0 || out_rate <= 0) return 0;
    if (in_rate == out_rate) return n_frames_in;
    const long long g = gcd_ll(in_rate, out_rate);
    return ((n_frames_in - 1) * (out_rate / g)) / (in_rate / g) + 1;
}

extern "C" int maf_front_door(maf_ctx* ctx, const int16_t* pcm, int64_t n_frames_in, int32_t channels, int32_t in_rate, int32_t out_rate,
                              int32_t peak_normalize, int16_t* out, void* stream) {
    if (!ctx || !pcm || !out || n_frames_in <= 0) return fail(MAF_ERR_INVALID, "maf_front_door: bad argument");
    if (channels != 1 && channels != 2) return fail(MAF_ERR_UNSUPPORTED, "maf_front_door: 1 or 2 channels");
    if (in_rate <= 0 || out_rate <= 0) return fail(MAF_ERR_INVALID, "maf_front_door: bad rate");
    CU(cudaSetDevice(ctx->device));
    const long long g = gcd_ll(in_rate, out_rate);
    launch_front_door(pcm, n_frames_in, channels, in_rate / g, out_rate / g, (int*)ctx->scratch, 32768.0 * pow(10.0, -0.1 / 20.0),
                      peak_normalize, out, maf_front_door_out_len(n_frames_in, in_rate, out_rate), (cudaStream_t)stream);
    CU(cudaGetLastError());
    return MAF_OK;
}

// ---- encoder windows ---------------------------------------------------------------------------------------------
extern "C" int maf_mel_windows(maf_ctx* ctx, const float* mel, int32_t n_mels, int64_t frames, const int32_t* seeks, const int32_t* lens,
                               int32_t n_windows, int32_t window_frames, int32_t out_fp16, void* out, void* stream) {
    if (!ctx || !mel || !seeks || !out) return fail(MAF_ERR_INVALID, "maf_mel_windows: bad argument");
    if (n_mels <= 0 || n_mels > 65535 || frames <= 0 || n_windows < 0 || window_frames <= 0)
        return fail(MAF_ERR_INVALID, "maf_mel_windows: bad shape");
    CU(cudaSetDevice(ctx->device));
    launch_mel_windows(mel, n_mels, frames, seeks, lens, n_windows, window_frames, out_fp16, out, (cudaStream_t)stream);
    CU(cudaGetLastError());
    return MAF_OK;
}
