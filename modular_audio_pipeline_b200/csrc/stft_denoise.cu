// S1-S4 -- spectral-gating denoise (reference call site: audio_pipeline/preprocessor.py:200-206 ->
// noisereduce.reduce_noise(stationary=True, prop_decrease=0.8); algorithm restated in oracle/reference_path.py).
//
//   noise_threshold_*  per-bin mean + 1.5 std of the noise clip's dB spectrogram  -> power thresholds T2 / T2hi
//   stft1024_kernel    K1: int16 PCM -> windowed real FFT 1024 / hop 256 on each chunk's own frame grid
//                          -> S (complex64, 4 KB per frame) + 513 raw-mask bits per frame (stationary)
//   ns_mask_kernel     K2a (non-stationary): forward-backward one-pole smoothing along time + sigmoid mask
//   smooth_mask_kernel K2: separable 7(time) x 33(freq) triangular smoothing of the raw mask -> M (fp32)
//   istft_kernel       K3: S * M -> inverse real FFT -> windowed overlap-add -> / 1.5 -> int16 (+ peak)
//
// The stationary gate "dB > thresh" is evaluated without a logarithm: 20 log10(|X| + eps) > t  <=>
// |X|^2 > (10^(t/20) - eps)^2 =: T2, and noisereduce's top_db=80 row clamp max(D, rowmax - 80) > t adds
// "or rowmax^2 > (10^((t+80)/20) - eps)^2 =: T2hi", recorded per (chunk, bin) in `rowflag`.
#include "maf_common.cuh"

namespace maf {

// 512-point FFT for one 64-thread group of a 256-thread CTA (group-local barriers inside).
// in: v[m] = x[t + 64 m]; out: v[d] = X[fft512_out_base(t) + 64 d].  `sm` is this group's FFT512_SMEM scratch.
// Barrier among the 64 threads (2 warps) of FFT group g only: named barriers 1..4, barrier 0 stays __syncthreads().
__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 64;" ::"r"(g + 1) : "memory"); }

__device__ __forceinline__ void group_fft512(c32 (&v)[8], c32* sm, const c32* tw1, const c32* tw2, int t, int g) {
    fft512_step1(v, sm, tw1, t);
    group_bar(g);
    fft512_step2_load(v, sm, t);
    group_bar(g);
    fft512_step2_store(v, sm, tw2, t);
    group_bar(g);
    fft512_step3(v, sm, t);
}

// ============================================================================================================
// K1  forward STFT
// ============================================================================================================
template <bool STATIONARY>
__global__ void __launch_bounds__(256) stft1024_kernel(const int16_t* __restrict__ pcm, const Item* __restrict__ items,
                                                       const Chunk* __restrict__ chunks, Consts cs,
                                                       const float* __restrict__ T2, const float* __restrict__ T2hi,
                                                       c32* __restrict__ S, uint32_t* __restrict__ bits,
                                                       uint8_t* __restrict__ rowflag) {
    constexpr int SPAN = (STFT_TILE + 3) * HOP;       // samples covered by the tile's frames
    __shared__ __align__(16) int16_t span[SPAN];
    __shared__ float win[NFFT];
    __shared__ c32 tw1[FFT512_TW1];
    __shared__ c32 tw2[FFT512_TW2];
    __shared__ c32 tw1024[257];
    __shared__ __align__(16) c32 fsm[4][FFT512_SMEM];
    __shared__ __align__(4) uint8_t flags[4][528];

    const Chunk ch = chunks[blockIdx.y];
    const int f0 = blockIdx.x * STFT_TILE;
    if (f0 >= ch.n_frames) return;
    const Item it = items[ch.item];
    const int tid = threadIdx.x;

    for (int i = tid; i < NFFT; i += 256) win[i] = cs.win1024[i];
    for (int i = tid; i < FFT512_TW1; i += 256) tw1[i] = cs.tw1[i];
    if (tid < FFT512_TW2) tw2[tid] = cs.tw2[tid];
    for (int i = tid; i < 257; i += 256) tw1024[i] = cs.tw1024[i];

    // stage the contiguous sample span once (frames overlap 75 %); zero outside the padded chunk and the recording
    {
        const long long j0 = (long long)f0 * HOP - NFFT / 2;          // padded-chunk coordinate of span[0]
        const long long li0 = ch.s_local - PADS + j0;                 // local buffer index of span[0]
        const int16_t* base = pcm + it.pcm_off;
        for (int g = tid; g < SPAN / 8; g += 256) {
            const long long j = j0 + 8 * g, li = li0 + 8 * g;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (j >= 0 && j + 8 <= ch.L && li >= 0 && li + 8 <= it.buf_len && ((it.pcm_off + li) & 7) == 0) {
                val = *reinterpret_cast<const uint4*>(base + li);
            } else {
                int16_t tmp[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const long long jj = j + e, ll = li + e;
                    tmp[e] = (jj >= 0 && jj < ch.L && ll >= 0 && ll < it.buf_len) ? base[ll] : (int16_t)0;
                }
                val = *reinterpret_cast<uint4*>(tmp);
            }
            *reinterpret_cast<uint4*>(&span[8 * g]) = val;
        }
    }
    __syncthreads();

    const int g = tid >> 6, t = tid & 63;
    const float* t2 = STATIONARY ? T2 + (long long)ch.item * T_PITCH : nullptr;
    const float* t2hi = STATIONARY ? T2hi + (long long)ch.item * T_PITCH : nullptr;
    uint8_t* rf = STATIONARY ? rowflag + (long long)blockIdx.y * M_PITCH : nullptr;

    // the four 64-thread groups run decoupled from here on: span / tables are read-only, fsm / flags are per group
    const short2* span2 = reinterpret_cast<const short2*>(span);
    const float2* win2 = reinterpret_cast<const float2*>(win);
    for (int r = 0; r < STFT_TILE / 4; ++r) {
        const int lf = r * 4 + g;
        const int f = f0 + lf;
        const bool active = f < ch.n_frames;
        c32 v[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int n = t + 64 * m;
            const short2 sv = span2[lf * (HOP / 2) + n];
            const float2 w = win2[n];
            v[m] = mk((float)sv.x * w.x, (float)sv.y * w.y);
        }
        group_fft512(v, fsm[g], tw1, tw2, t, g);
        group_bar(g);
#pragma unroll
        for (int d = 0; d < 8; ++d) fsm[g][t + 64 * d] = v[d];
        group_bar(g);
        if (active) {
            c32* row = S + (ch.row0 + f) * S_PITCH;
            const c32* z = fsm[g];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = t + 64 * q;
                if (k == 0) {
                    const c32 z0 = z[0], zm = z[256];
                    const float x0 = z0.x + z0.y, xn = z0.x - z0.y;
                    row[0] = mk(x0, xn);
                    const c32 xm = cconj(zm);
                    row[256] = xm;
                    if (STATIONARY) {
                        const float p0 = x0 * x0, pn = xn * xn, pm = xm.x * xm.x + xm.y * xm.y;
                        flags[g][0] = p0 > t2[0]; flags[g][512] = pn > t2[512]; flags[g][256] = pm > t2[256];
                        if (p0 > t2hi[0]) rf[0] = 1;
                        if (pn > t2hi[512]) rf[512] = 1;
                        if (pm > t2hi[256]) rf[256] = 1;
                    }
                } else {
                    c32 xk, xnk;
                    rfft_split_pair(z[k], z[512 - k], tw1024[k], xk, xnk);
                    row[k] = xk;
                    row[512 - k] = xnk;
                    if (STATIONARY) {
                        const float pk = xk.x * xk.x + xk.y * xk.y, pnk = xnk.x * xnk.x + xnk.y * xnk.y;
                        flags[g][k] = pk > t2[k]; flags[g][512 - k] = pnk > t2[512 - k];
                        if (pk > t2hi[k]) rf[k] = 1;
                        if (pnk > t2hi[512 - k]) rf[512 - k] = 1;
                    }
                }
            }
        }
        if (STATIONARY) {
            group_bar(g);
            if (active && t < 17) {
                const uint32_t* fw = reinterpret_cast<const uint32_t*>(&flags[g][32 * t]);
                uint32_t word = 0;
                const int nb = (t == 16) ? 1 : 32;
                for (int b = 0; b < nb; b += 4) {
                    const uint32_t w4 = fw[b >> 2];                      // four 0/1 bytes, lowest bin in the low byte
                    const uint32_t nib = (w4 & 1u) | ((w4 >> 7) & 2u) | ((w4 >> 14) & 4u) | ((w4 >> 21) & 8u);
                    word |= nib << b;
                }
                if (t == 16) word &= 1u;
                bits[(ch.row0 + f) * B_PITCH + t] = word;
            }
        }
        group_bar(g);
    }
}

void launch_stft(const int16_t* pcm, const Item* items, const Chunk* chunks, int n_chunks, int max_frames,
                 const Consts& c, int stationary, const float* T2, const float* T2hi, c32* S, uint32_t* bits,
                 uint8_t* rowflag, cudaStream_t st) {
    if (n_chunks <= 0) return;
    const int tiles = (max_frames + STFT_TILE - 1) / STFT_TILE;
    for (int y0 = 0; y0 < n_chunks; y0 += 32768) {
        const int ny = n_chunks - y0 < 32768 ? n_chunks - y0 : 32768;
        dim3 grid(tiles, ny);
        if (stationary)
            stft1024_kernel<true><<<grid, 256, 0, st>>>(pcm, items, chunks + y0, c, T2, T2hi, S, bits,
                                                        rowflag + (long long)y0 * M_PITCH);
        else
            stft1024_kernel<false><<<grid, 256, 0, st>>>(pcm, items, chunks + y0, c, nullptr, nullptr, S, nullptr, nullptr);
    }
}

// ============================================================================================================
// Noise threshold: STFT of the noise clip (zero boundary, padded=False -> len/256 + 1 frames), dB with the
// top_db=80 row clamp, per-bin mean and population std in float64, threshold = mean + n_std * std.
// Two passes over the clip's frames (row max first, then the clamped statistics), both recomputing the FFT;
// per-(slice, group) partial sums are combined in a fixed order so the thresholds are run-to-run identical.
// ============================================================================================================
template <int PASS>
__global__ void __launch_bounds__(256) noise_pass_kernel(const int16_t* __restrict__ pcm,
                                                         const Clip* __restrict__ clips, Consts cs, unsigned* __restrict__ nmax,
                                                         double* __restrict__ partial, int n_slices) {
    __shared__ float win[NFFT];
    __shared__ c32 tw1[FFT512_TW1];
    __shared__ c32 tw2[FFT512_TW2];
    __shared__ c32 tw1024[257];
    __shared__ __align__(16) c32 fsm[4][FFT512_SMEM];
    const int item = blockIdx.y;
    const Clip cl = clips[item];
    const int tid = threadIdx.x, g = tid >> 6, t = tid & 63;
    for (int i = tid; i < NFFT; i += 256) win[i] = cs.win1024[i];
    for (int i = tid; i < FFT512_TW1; i += 256) tw1[i] = cs.tw1[i];
    if (tid < FFT512_TW2) tw2[tid] = cs.tw2[tid];
    for (int i = tid; i < 257; i += 256) tw1024[i] = cs.tw1024[i];
    __syncthreads();
    const int16_t* base = pcm + cl.start;
    // bins owned by thread t: k = t + 64 q and 512 - k (q = 0..3); t == 0 additionally owns 0 / 256 / 512
    float pmax[9];
    double s1[9], s2[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { pmax[i] = 0.f; s1[i] = 0.0; s2[i] = 0.0; }
    const double EPS = 2.220446049250313e-16;
    double floor_db[9];
    if (PASS == 2) {
        const unsigned* nm = nmax + (long long)item * T_PITCH;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = t + 64 * q;
            if (k == 0) {
                floor_db[0] = 20.0 * log10(sqrt((double)__uint_as_float(nm[0])) + EPS) - 80.0;
                floor_db[1] = 20.0 * log10(sqrt((double)__uint_as_float(nm[512])) + EPS) - 80.0;
                floor_db[8] = 20.0 * log10(sqrt((double)__uint_as_float(nm[256])) + EPS) - 80.0;
            } else {
                floor_db[2 * q] = 20.0 * log10(sqrt((double)__uint_as_float(nm[k])) + EPS) - 80.0;
                floor_db[2 * q + 1] = 20.0 * log10(sqrt((double)__uint_as_float(nm[512 - k])) + EPS) - 80.0;
            }
        }
    }
    const int stride = 4 * n_slices;
    const int first = blockIdx.x * 4 + g;
    const int rounds = (cl.n_frames + stride - 1) / stride;
    for (int r = 0; r < rounds; ++r) {
        const int f = first + r * stride;
        const bool active = f < cl.n_frames;
        c32 v[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int n = t + 64 * m;
            const long long i = (long long)f * HOP - NFFT / 2 + 2 * n;     // clip-relative sample index
            float a = 0.f, b = 0.f;
            if (active) {
                if (i >= 0 && i < cl.len) a = (float)base[i];
                if (i + 1 >= 0 && i + 1 < cl.len) b = (float)base[i + 1];
            }
            v[m] = mk(a * win[2 * n], b * win[2 * n + 1]);
        }
        group_fft512(v, fsm[g], tw1, tw2, t, g);
        group_bar(g);
#pragma unroll
        for (int d = 0; d < 8; ++d) fsm[g][t + 64 * d] = v[d];
        group_bar(g);
        if (active) {
            const c32* z = fsm[g];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = t + 64 * q;
                float pa, pb, pc = 0.f;
                if (k == 0) {
                    const c32 z0 = z[0], zm = z[256];
                    const float x0 = z0.x + z0.y, xn = z0.x - z0.y;
                    pa = x0 * x0; pb = xn * xn; pc = zm.x * zm.x + zm.y * zm.y;
                } else {
                    c32 xk, xnk;
                    rfft_split_pair(z[k], z[512 - k], tw1024[k], xk, xnk);
                    pa = xk.x * xk.x + xk.y * xk.y; pb = xnk.x * xnk.x + xnk.y * xnk.y;
                }
                if (PASS == 1) {
                    pmax[2 * q] = fmaxf(pmax[2 * q], pa); pmax[2 * q + 1] = fmaxf(pmax[2 * q + 1], pb);
                    if (k == 0) pmax[8] = fmaxf(pmax[8], pc);
                } else {
                    double da = fmax(20.0 * log10(sqrt((double)pa) + EPS), floor_db[2 * q]);
                    double db = fmax(20.0 * log10(sqrt((double)pb) + EPS), floor_db[2 * q + 1]);
                    s1[2 * q] += da; s2[2 * q] += da * da; s1[2 * q + 1] += db; s2[2 * q + 1] += db * db;
                    if (k == 0) { double dc = fmax(20.0 * log10(sqrt((double)pc) + EPS), floor_db[8]); s1[8] += dc; s2[8] += dc * dc; }
                }
            }
        }
        group_bar(g);
    }
    // publish
    if (PASS == 1) {
        unsigned* nm = nmax + (long long)item * T_PITCH;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = t + 64 * q;
            if (k == 0) {
                atomicMax(&nm[0], __float_as_uint(pmax[0])); atomicMax(&nm[512], __float_as_uint(pmax[1]));
                atomicMax(&nm[256], __float_as_uint(pmax[8]));
            } else {
                atomicMax(&nm[k], __float_as_uint(pmax[2 * q])); atomicMax(&nm[512 - k], __float_as_uint(pmax[2 * q + 1]));
            }
        }
    } else {
        double* pr = partial + ((long long)item * stride + first) * (2 * T_PITCH);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = t + 64 * q;
            if (k == 0) {
                pr[0] = s1[0]; pr[T_PITCH + 0] = s2[0]; pr[512] = s1[1]; pr[T_PITCH + 512] = s2[1];
                pr[256] = s1[8]; pr[T_PITCH + 256] = s2[8];
            } else {
                pr[k] = s1[2 * q]; pr[T_PITCH + k] = s2[2 * q];
                pr[512 - k] = s1[2 * q + 1]; pr[T_PITCH + 512 - k] = s2[2 * q + 1];
            }
        }
    }
}

__global__ void noise_finalize_kernel(const Clip* __restrict__ clips, const double* __restrict__ partial, int n_parts,
                                      double n_std, float* __restrict__ T2, float* __restrict__ T2hi) {
    const int item = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= NBIN) return;
    const int n = clips[item].n_frames;
    double s1 = 0.0, s2 = 0.0;
    for (int p = 0; p < n_parts; ++p) {
        const double* pr = partial + ((long long)item * n_parts + p) * (2 * T_PITCH);
        s1 += pr[k]; s2 += pr[T_PITCH + k];
    }
    const double mean = s1 / n;
    double var = s2 / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const double thresh = mean + n_std * sqrt(var);
    const double EPS = 2.220446049250313e-16;
    const double ta = pow(10.0, thresh / 20.0) - EPS;
    const double tb = pow(10.0, (thresh + 80.0) / 20.0) - EPS;
    T2[(long long)item * T_PITCH + k] = ta < 0.0 ? -1.0f : (float)fmin(ta * ta, 3.0e38);
    T2hi[(long long)item * T_PITCH + k] = tb < 0.0 ? -1.0f : (float)fmin(tb * tb, 3.0e38);
}

void launch_noise_threshold(const int16_t* pcm, const Clip* clips, int n_items, int n_slices,
                            const Consts& c, unsigned* nmax, double* partial, float* T2, float* T2hi, double n_std,
                            cudaStream_t st) {
    if (n_items <= 0) return;
    cudaMemsetAsync(nmax, 0, sizeof(unsigned) * (size_t)n_items * T_PITCH, st);
    cudaMemsetAsync(partial, 0, sizeof(double) * (size_t)n_items * n_slices * 4 * 2 * T_PITCH, st);
    dim3 grid(n_slices, n_items);
    noise_pass_kernel<1><<<grid, 256, 0, st>>>(pcm, clips, c, nmax, partial, n_slices);
    noise_pass_kernel<2><<<grid, 256, 0, st>>>(pcm, clips, c, nmax, partial, n_slices);
    dim3 g2((NBIN + 127) / 128, n_items);
    noise_finalize_kernel<<<g2, 128, 0, st>>>(clips, partial, n_slices * 4, n_std, T2, T2hi);
}

// ============================================================================================================
// K2a  non-stationary raw mask: Sm = filtfilt([b], [1, b-1], |S|) along time, M0 = sigmoid(((|S|-Sm)/Sm - 2) * 10)
// One thread per (chunk, bin); consecutive threads read consecutive bins of a frame row (coalesced).  The
// recurrence runs in float64 (cheap on B200, keeps the 0.992-pole accumulation at 1e-16).
// Deviation (documented in DESIGN.md): where Sm == 0 (an all-zero frequency row) noisereduce yields NaN; here the
// relative excess is taken as 0 so the mask is sigmoid(-20) and the output is silence instead of NaN.
// ============================================================================================================
__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

constexpr int NS_BINS = 64;          // bins per CTA: one 512-byte segment of every S row, 256 bytes of every M0 row
constexpr int NS_TT = 32;            // frames per streamed tile
constexpr int NS_THREADS = 128;
constexpr int NS_PITCH = NS_BINS + 1;

// A CTA owns (chunk, 64 bins) -- plus bin 512, which shares slot 0 with bin 0, for group 0 -- and streams the chunk's
// frames through shared memory in tiles of 32: all threads load the tile with 16-byte coalesced accesses (the next tile
// is prefetched into registers while the current one is processed), lanes 0..64 run the serial one-pole recurrence over
// the tile from shared memory, all threads write the tile back.  Forward pass writes yf into M0, the backward pass
// re-reads S and yf in reverse and overwrites M0 with the sigmoid mask.  8 groups x chunks CTAs, all resident at once.
__global__ void __launch_bounds__(NS_THREADS, 6) ns_mask_kernel(const c32* __restrict__ S, const Chunk* __restrict__ chunks, double b,
                                                                float thresh_mult, float slope, float* __restrict__ M0) {
    __shared__ float A[NS_TT][NS_PITCH];
    __shared__ float Y[NS_TT][NS_PITCH];
    const Chunk ch = chunks[blockIdx.y];
    const int gi = blockIdx.x;
    const int F = ch.n_frames;
    const int tid = threadIdx.x;
    const bool has_nyq = gi == 0;
    const int nser = has_nyq ? NS_BINS + 1 : NS_BINS;
    const float4* s4 = reinterpret_cast<const float4*>(S + ch.row0 * S_PITCH + gi * NS_BINS);   // row pitch = 256 float4
    float* m0 = M0 + ch.row0 * M_PITCH + gi * NS_BINS;
    const double a = 1.0 - b;
    const int n_tiles = (F + NS_TT - 1) / NS_TT;

    float4 ps[8];                                   // prefetched S tile: element idx = tid + 128 i -> row idx/32, float4 idx%32
    auto load_s = [&](int f0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + NS_THREADS * i, r = idx >> 5, c = idx & 31;
            ps[i] = (f0 + r < F) ? __ldg(s4 + (long long)(f0 + r) * (S_PITCH / 2) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto put_abs = [&]() {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + NS_THREADS * i, r = idx >> 5, c = idx & 31;
            if (has_nyq && c == 0) { A[r][0] = fabsf(ps[i].x); A[r][NS_BINS] = fabsf(ps[i].y); }
            else A[r][2 * c] = fast_sqrt(ps[i].x * ps[i].x + ps[i].y * ps[i].y);
            A[r][2 * c + 1] = fast_sqrt(ps[i].z * ps[i].z + ps[i].w * ps[i].w);
        }
    };

    // ---------------- forward pass (steady-state start: y[-1] = x[0]) ----------------
    double y = 0.0;
    load_s(0);
    for (int tI = 0; tI < n_tiles; ++tI) {
        const int f0 = tI * NS_TT;
        const int rows = min(NS_TT, F - f0);
        put_abs();
        __syncthreads();
        if (tI + 1 < n_tiles) load_s(f0 + NS_TT);
        if (tid < nser) {
            if (tI == 0) y = (double)A[0][tid];
#pragma unroll 8
            for (int r = 0; r < rows; ++r) { y = b * (double)A[r][tid] + a * y; Y[r][tid] = (float)y; }
        }
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const int idx = tid + NS_THREADS * i, r = idx >> 6, c = idx & 63;
            if (r < rows) m0[(long long)(f0 + r) * M_PITCH + c] = Y[r][c];
        }
        if (has_nyq && tid < rows) M0[(ch.row0 + f0 + tid) * M_PITCH + 512] = Y[tid][NS_BINS];
        // the next iteration's put_abs only touches A, and Y is rewritten after the next barrier: no extra barrier needed
    }
    // ---------------- backward pass over yf, again started in steady state ----------------
    double z = (double)(float)y;                    // yf[F - 1] as stored
    float4 py[4];
    auto load_y = [&](int f0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + NS_THREADS * i, r = idx >> 4, c = idx & 15;
            py[i] = (f0 + r < F) ? *reinterpret_cast<const float4*>(m0 + (long long)(f0 + r) * M_PITCH + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    __syncthreads();
    load_s((n_tiles - 1) * NS_TT);
    load_y((n_tiles - 1) * NS_TT);
    float ynyq = (has_nyq && tid < NS_TT && (n_tiles - 1) * NS_TT + tid < F) ? M0[(ch.row0 + (n_tiles - 1) * NS_TT + tid) * M_PITCH + 512] : 0.f;
    for (int tI = n_tiles - 1; tI >= 0; --tI) {
        const int f0 = tI * NS_TT;
        const int rows = min(NS_TT, F - f0);
        put_abs();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + NS_THREADS * i, r = idx >> 4, c = idx & 15;
            Y[r][4 * c] = py[i].x; Y[r][4 * c + 1] = py[i].y; Y[r][4 * c + 2] = py[i].z; Y[r][4 * c + 3] = py[i].w;
        }
        if (has_nyq && tid < NS_TT) Y[tid][NS_BINS] = ynyq;
        __syncthreads();
        if (tI > 0) {
            load_s(f0 - NS_TT); load_y(f0 - NS_TT);
            if (has_nyq && tid < NS_TT) ynyq = M0[(ch.row0 + f0 - NS_TT + tid) * M_PITCH + 512];
        }
        if (tid < nser) {
#pragma unroll 8
            for (int r = rows - 1; r >= 0; --r) { z = b * (double)Y[r][tid] + a * z; Y[r][tid] = (float)z; }
        }
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const int idx = tid + NS_THREADS * i, r = idx >> 6, c = idx & 63;
            if (r < rows) {
                const float sm = Y[r][c], x = A[r][c];
                const float rr = sm > 0.f ? __fdividef(x - sm, sm) : 0.f;
                m0[(long long)(f0 + r) * M_PITCH + c] = __fdividef(1.0f, 1.0f + __expf(-(rr - thresh_mult) * slope));
            }
        }
        if (has_nyq && tid < rows) {
            const float sm = Y[tid][NS_BINS], x = A[tid][NS_BINS];
            const float rr = sm > 0.f ? __fdividef(x - sm, sm) : 0.f;
            M0[(ch.row0 + f0 + tid) * M_PITCH + 512] = __fdividef(1.0f, 1.0f + __expf(-(rr - thresh_mult) * slope));
        }
        __syncthreads();
    }
}

void launch_ns_mask(const c32* S, const Chunk* chunks, int n_chunks, double b, float thresh_mult, float slope,
                    float* M0, cudaStream_t st) {
    if (n_chunks <= 0) return;
    for (int y0 = 0; y0 < n_chunks; y0 += 32768) {
        const int ny = n_chunks - y0 < 32768 ? n_chunks - y0 : 32768;
        dim3 grid(S_PITCH / NS_BINS, ny);
        ns_mask_kernel<<<grid, NS_THREADS, 0, st>>>(S, chunks + y0, b, thresh_mult, slope, M0);
    }
}

// ============================================================================================================
// K2  separable triangular smoothing (scipy.signal.fftconvolve(mask, outer(vf, vt)/sum, 'same'), zero padded)
//   stationary    : raw = bit | rowflag ? 1 : 1 - prop          (blend, then smooth)
//   non-stationary: raw = M0 ; out = smooth(raw) * prop + 1 - prop   (smooth, then blend)
// Phase 1 applies the time taps while reading the raw rows from global/L2 (each row is re-read by at most
// 2 nt + 1 neighbouring output rows of the same CTA), writing rows into shared memory with MAX_NF zeros on both
// sides; phase 2 applies the frequency taps from shared memory, 8 outputs per thread in registers.
// ============================================================================================================
constexpr int SM_ROW = MAX_NF + 520 + MAX_NF;     // 584 floats, multiple of 4

template <bool STATIONARY>
__global__ void __launch_bounds__(256) smooth_mask_kernel(const uint32_t* __restrict__ bits, const uint8_t* __restrict__ rowflag,
                                                          const float* __restrict__ M0, const Chunk* __restrict__ chunks,
                                                          float prop, SmoothTaps taps, float* __restrict__ M) {
    __shared__ __align__(16) float rows[MASK_TILE][SM_ROW];
    __shared__ uint32_t rfbits[17];
    const Chunk ch = chunks[blockIdx.y];
    const int f0 = blockIdx.x * MASK_TILE;
    if (f0 >= ch.n_frames) return;
    const int tid = threadIdx.x;
    const int nt = taps.nt, nf = taps.nf;
    for (int i = tid; i < MASK_TILE * SM_ROW; i += 256) (&rows[0][0])[i] = 0.f;
    if (STATIONARY && tid < 17) {
        const uint8_t* rf = rowflag + (long long)blockIdx.y * M_PITCH;
        uint32_t w = 0;
        for (int b = 0; b < 32; ++b) { const int k = 32 * tid + b; if (k < NBIN && rf[k]) w |= 1u << b; }
        rfbits[tid] = w;
    }
    __syncthreads();
    if (STATIONARY) {
        const float lo = 1.0f - prop;
        for (int w = tid; w < MASK_TILE * 17; w += 256) {
            const int lf = w / 17, wi = w % 17;
            const int f = f0 + lf;
            if (f >= ch.n_frames) continue;
            float acc[32];
#pragma unroll
            for (int b = 0; b < 32; ++b) acc[b] = 0.f;
            const uint32_t rfw = rfbits[wi];
            for (int j = -nt; j <= nt; ++j) {
                const int ff = f + j;
                if (ff < 0 || ff >= ch.n_frames) continue;
                const uint32_t word = bits[(ch.row0 + ff) * B_PITCH + wi] | rfw;
                const float wt = taps.vt[j + nt];
#pragma unroll
                for (int b = 0; b < 32; ++b) acc[b] += wt * (((word >> b) & 1u) ? 1.0f : lo);
            }
            const int nb = (wi == 16) ? 1 : 32;
#pragma unroll
            for (int b = 0; b < 32; ++b) if (b < nb) rows[lf][MAX_NF + 32 * wi + b] = acc[b];
        }
    } else if (nt == 3) {
        // one work item = 4 consecutive bins x 8 output frames: 14 float4 row loads feed 8 outputs (sliding window)
        for (int w = tid; w < 2 * 129; w += 256) {
            const int q = w >> 1, h = w & 1;
            const int fa = f0 + 8 * h;                       // first output frame of this item
            const float* src = M0 + ch.row0 * M_PITCH + 4 * q;
            float4 x[14];
#pragma unroll
            for (int j = 0; j < 14; ++j) {
                const int ff = fa - 3 + j;
                x[j] = (ff >= 0 && ff < ch.n_frames) ? *reinterpret_cast<const float4*>(src + (long long)ff * M_PITCH) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const float t0 = taps.vt[0], t1 = taps.vt[1], t2 = taps.vt[2], t3 = taps.vt[3];
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                if (fa + o >= ch.n_frames) break;
                float4 r;
                r.x = t0 * (x[o].x + x[o + 6].x) + t1 * (x[o + 1].x + x[o + 5].x) + t2 * (x[o + 2].x + x[o + 4].x) + t3 * x[o + 3].x;
                r.y = t0 * (x[o].y + x[o + 6].y) + t1 * (x[o + 1].y + x[o + 5].y) + t2 * (x[o + 2].y + x[o + 4].y) + t3 * x[o + 3].y;
                r.z = t0 * (x[o].z + x[o + 6].z) + t1 * (x[o + 1].z + x[o + 5].z) + t2 * (x[o + 2].z + x[o + 4].z) + t3 * x[o + 3].z;
                r.w = t0 * (x[o].w + x[o + 6].w) + t1 * (x[o + 1].w + x[o + 5].w) + t2 * (x[o + 2].w + x[o + 4].w) + t3 * x[o + 3].w;
                if (q == 128) { r.y = 0.f; r.z = 0.f; r.w = 0.f; }           // bins 513..515 are row padding
                *reinterpret_cast<float4*>(&rows[8 * h + o][MAX_NF + 4 * q]) = r;
            }
        }
    } else {
        for (int e = tid; e < MASK_TILE * NBIN; e += 256) {
            const int lf = e / NBIN, k = e - lf * NBIN;
            const int f = f0 + lf;
            if (f >= ch.n_frames) continue;
            float acc = 0.f;
            for (int j = -nt; j <= nt; ++j) {
                const int ff = f + j;
                if (ff < 0 || ff >= ch.n_frames) continue;
                acc += taps.vt[j + nt] * M0[(ch.row0 + ff) * M_PITCH + k];
            }
            rows[lf][MAX_NF + k] = acc;
        }
    }
    __syncthreads();
    // phase 2: frequency taps, 8 consecutive outputs per work item
    for (int w = tid; w < MASK_TILE * 65; w += 256) {
        const int lf = w / 65, seg = w - lf * 65;
        const int f = f0 + lf;
        if (f >= ch.n_frames) continue;
        const int k0 = 8 * seg;
        float out[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) out[o] = 0.f;
        const float* src = &rows[lf][MAX_NF + k0];
        if (nf == 16) {
            // triangle(33) = box(17) * box(17) / 289: two running sums restarted per work item (<= 23 updates each)
            float in[40];
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                const float4 q = *reinterpret_cast<const float4*>(src - 16 + 4 * i);
                in[4 * i] = q.x; in[4 * i + 1] = q.y; in[4 * i + 2] = q.z; in[4 * i + 3] = q.w;
            }
            float b1[24];                                   // b1[i] = sum in[i .. i + 16]  (centre k0 - 8 + i)
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 17; ++i) acc += in[i];
            b1[0] = acc;
#pragma unroll
            for (int i = 1; i < 24; ++i) { acc += in[i + 16] - in[i - 1]; b1[i] = acc; }
            float a2 = 0.f;
#pragma unroll
            for (int i = 0; i < 17; ++i) a2 += b1[i];
            out[0] = a2;
#pragma unroll
            for (int o = 1; o < 8; ++o) { a2 += b1[o + 16] - b1[o - 1]; out[o] = a2; }
#pragma unroll
            for (int o = 0; o < 8; ++o) out[o] *= (1.0f / 289.0f);
        } else {
            for (int tp = 0; tp <= 2 * nf; ++tp) {
                const float wt = taps.vf[tp];
#pragma unroll
                for (int o = 0; o < 8; ++o) out[o] += wt * src[o + tp - nf];
            }
        }
        float* dst = M + (ch.row0 + f) * M_PITCH + k0;
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            if (k0 + o < NBIN) dst[o] = STATIONARY ? out[o] : out[o] * prop + (1.0f - prop);
        }
    }
}

void launch_smooth_mask(const uint32_t* bits, const uint8_t* rowflag, const float* M0, const Chunk* chunks,
                        int n_chunks, int max_frames, int stationary, float prop, const SmoothTaps& taps, float* M,
                        cudaStream_t st) {
    if (n_chunks <= 0) return;
    const int tiles = (max_frames + MASK_TILE - 1) / MASK_TILE;
    for (int y0 = 0; y0 < n_chunks; y0 += 32768) {
        const int ny = n_chunks - y0 < 32768 ? n_chunks - y0 : 32768;
        dim3 grid(tiles, ny);
        if (stationary)
            smooth_mask_kernel<true><<<grid, 256, 0, st>>>(bits, rowflag + (long long)y0 * M_PITCH, nullptr, chunks + y0, prop, taps, M);
        else
            smooth_mask_kernel<false><<<grid, 256, 0, st>>>(nullptr, nullptr, M0, chunks + y0, prop, taps, M);
    }
}

// ============================================================================================================
// K3  masked inverse STFT with window-squared normalised overlap-add (scipy.signal.istft), int16 truncation
// (preprocessor.py:211) and the per-recording peak that the following peak-normalise needs.
// A CTA owns ISTFT_SEG output segments of 256 samples and recomputes the 3 extra frames that overlap them, so
// every output sample is summed in ascending frame order inside one CTA (deterministic, no atomics, no second
// pass).  Four frames are transformed concurrently; they are 8 frames apart so their 1024-sample spans are disjoint.
// ============================================================================================================
constexpr int ISTFT_FR = ISTFT_SEG + 3;                 // 32 frames per CTA
constexpr int OLA_LEN = (ISTFT_FR + 3) * HOP;           // 8960 floats
constexpr size_t ISTFT_SMEM = sizeof(float) * OLA_LEN + sizeof(float) * NFFT + sizeof(c32) * (FFT512_TW1 + FFT512_TW2) + sizeof(c32) * 264 +
                              sizeof(c32) * 4 * FFT512_SMEM;

__global__ void __launch_bounds__(256) istft_kernel(const c32* __restrict__ S, const float* __restrict__ M,
                                                    const Item* __restrict__ items, const Chunk* __restrict__ chunks, Consts cs,
                                                    int16_t* __restrict__ den16, int to_out, float* __restrict__ den_f32, int* __restrict__ peak) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* ola = reinterpret_cast<float*>(smem_raw);
    float* winv = ola + OLA_LEN;
    c32* tw1 = reinterpret_cast<c32*>(winv + NFFT);
    c32* tw2 = tw1 + FFT512_TW1;
    c32* tw1024 = tw2 + FFT512_TW2;
    c32* fsm = tw1024 + 264;

    const Chunk ch = chunks[blockIdx.y];
    const int seg_lo = PADS / HOP;                                  // first segment holding kept samples
    const int seg_hi = (PADS + ch.keep - 1) / HOP;                  // last one
    const int g0 = seg_lo + blockIdx.x * ISTFT_SEG;
    if (ch.keep <= 0 || g0 > seg_hi) return;
    const Item it = items[ch.item];
    const int tid = threadIdx.x, g = tid >> 6, t = tid & 63;

    for (int i = tid; i < OLA_LEN; i += 256) ola[i] = 0.f;
    for (int i = tid; i < NFFT; i += 256) winv[i] = cs.winv1024[i];
    for (int i = tid; i < FFT512_TW1; i += 256) tw1[i] = cs.tw1[i];
    if (tid < FFT512_TW2) tw2[tid] = cs.tw2[tid];
    for (int i = tid; i < 257; i += 256) tw1024[i] = cs.tw1024[i];
    __syncthreads();

    const int fbase = g0 - 1;                                       // local frame 0
    for (int r = 0; r < ISTFT_FR / 4; ++r) {
        const int lf = r + (ISTFT_FR / 4) * g;
        const int f = fbase + lf;
        const bool active = f >= 0 && f < ch.n_frames;
        c32* z = fsm + g * FFT512_SMEM;
        if (active) {
            const c32* row = S + (ch.row0 + f) * S_PITCH;
            const float* mrow = M + (ch.row0 + f) * M_PITCH;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = t + 64 * q;
                if (k == 0) {
                    const c32 x0 = row[0];
                    const float a = x0.x * mrow[0], b = x0.y * mrow[512];
                    z[0] = mk(a + b, -(a - b));                     // 2 conj(Z0)
                    const c32 xm = row[256];
                    const float mm = mrow[256];
                    z[256] = mk(2.f * xm.x * mm, 2.f * xm.y * mm);  // 2 conj(Z256) = 2 X256
                } else {
                    const c32 xk = row[k], xnk = row[512 - k];
                    const float mk_ = mrow[k], mnk = mrow[512 - k];
                    c32 zk, znk;
                    irfft_merge_pair(mk(xk.x * mk_, xk.y * mk_), mk(xnk.x * mnk, xnk.y * mnk), tw1024[k], zk, znk);
                    z[k] = zk; z[512 - k] = znk;
                }
            }
        }
        group_bar(g);
        c32 v[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) v[m] = active ? z[t + 64 * m] : mk(0.f, 0.f);
        group_bar(g);
        group_fft512(v, z, tw1, tw2, t, g);
        if (active) {
            float2* o2 = reinterpret_cast<float2*>(ola + lf * HOP);
            const float2* w2 = reinterpret_cast<const float2*>(winv);
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const int n = t + 64 * d;
                float2 acc = o2[n];
                const float2 w = w2[n];
                acc.x += v[d].x * w.x;
                acc.y += -v[d].y * w.y;
                o2[n] = acc;
            }
        }
        __syncthreads();            // the overlap-add buffer is shared by all four groups: one CTA-wide barrier per round
    }
    // emit: segment g covers padded-chunk samples [256 g, 256 g + 256) = ola[768 + 256 (g - g0) ...]
    int pk = 0;
    const long long obase = (to_out ? it.out_off : it.den_off) + ch.out_off;
    const long long fbase_o = it.out_off + ch.out_off;
    for (int i = tid; i < ISTFT_SEG * HOP; i += 256) {
        const long long j = (long long)g0 * HOP + i;
        const long long o = j - PADS;
        if (o < 0 || o >= ch.keep) continue;
        const float val = ola[3 * HOP + i] / 1.5f;
        const int q = (int)(short)(int)val;                         // astype(int16): truncate, keep low 16 bits
        den16[obase + o] = (int16_t)q;
        if (den_f32) den_f32[fbase_o + o] = val;
        pk = max(pk, q < 0 ? -q : q);
    }
    pk = max(pk, __shfl_xor_sync(0xffffffffu, pk, 16));
    pk = max(pk, __shfl_xor_sync(0xffffffffu, pk, 8));
    pk = max(pk, __shfl_xor_sync(0xffffffffu, pk, 4));
    pk = max(pk, __shfl_xor_sync(0xffffffffu, pk, 2));
    pk = max(pk, __shfl_xor_sync(0xffffffffu, pk, 1));
    if ((tid & 31) == 0 && pk > 0) atomicMax(&peak[ch.item], pk);
}

void launch_istft(const c32* S, const float* M, const Item* items, const Chunk* chunks, int n_chunks,
                  const Consts& c, int16_t* den16, int to_out, float* den_f32, int* peak, cudaStream_t st) {
    if (n_chunks <= 0) return;
    cudaFuncSetAttribute(istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ISTFT_SMEM);
    const int segs = (PADS + CHUNK - 1) / HOP - PADS / HOP + 1;        // 2344 segments for a full chunk
    const int tiles = (segs + ISTFT_SEG - 1) / ISTFT_SEG;
    for (int y0 = 0; y0 < n_chunks; y0 += 32768) {
        const int ny = n_chunks - y0 < 32768 ? n_chunks - y0 : 32768;
        dim3 grid(tiles, ny);
        istft_kernel<<<grid, 256, ISTFT_SMEM, st>>>(S, M, items, chunks + y0, c, den16, to_out, den_f32, peak);
    }
}

}  // namespace maf
