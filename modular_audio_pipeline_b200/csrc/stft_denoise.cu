// S1-S4 -- spectral-gating denoise (reference call site: audio_pipeline/preprocessor.py:200-206 ->
// noisereduce.reduce_noise(stationary=True, prop_decrease=0.8); algorithm restated in oracle/reference_path.py).
//
//   noise_threshold_*  per-bin mean + 1.5 std of the noise clip's dB spectrogram  -> power thresholds T2 / T2hi
//   stft1024_kernel    K1: int16 PCM -> windowed real FFT 1024 / hop 256 on each chunk's own frame grid
//                          -> S (complex64, 4 KB per frame) + 513 raw-mask bits per frame (stationary)
//   ns_scan_kernel / ns_mask_smooth_kernel
//                      K2a (non-stationary): forward-backward one-pole smoothing along time as a tile scan, sigmoid
//                          mask and the separable smoothing, S read once -> M (fp32)
//   smooth_mask_kernel K2 (stationary): separable 7(time) x 33(freq) triangular smoothing of the raw bits -> M (fp32)
//   istft_kernel       K3: S * M -> inverse real FFT -> windowed overlap-add -> / 1.5 -> int16 (+ peak)
//
// The stationary gate "dB > thresh" is evaluated without a logarithm: 20 log10(|X| + eps) > t  <=>
// |X|^2 > (10^(t/20) - eps)^2 =: T2, and noisereduce's top_db=80 row clamp max(D, rowmax - 80) > t adds
// "or rowmax^2 > (10^((t+80)/20) - eps)^2 =: T2hi", recorded per (chunk, bin) in `rowflag`.
#include <algorithm>
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

// the same transform with the twiddles rebuilt from two per-thread bases (no twiddle tables in shared memory)
__device__ __forceinline__ void group_fft512_base(c32 (&v)[8], c32* sm, c32 w1, c32 w2, int t, int g) {
    fft512_step1_base(v, sm, w1, t);
    group_bar(g);
    fft512_step2_load(v, sm, t);
    group_bar(g);
    fft512_step2_store_base(v, sm, w2, t);
    group_bar(g);
    fft512_step3(v, sm, t);
}

// ============================================================================================================
// K1  forward STFT.  A CTA stages the contiguous int16 span of a tile of STFT_TILE frames once (75 % overlap); every warp
// transforms PAIRS of consecutive frames with the warp-level 1024-point complex FFT of fft_core.h (frame A in the real
// part, frame B in the imaginary part), separates them with one shuffle exchange between lanes l and 32 - l, and stores
// two 4 KB rows with 256-byte coalesced stores.  Lane l ends up with bins 32 k1 + l (k1 = 0..15) of both frames, lane 0
// additionally with bin 512.
//   stationary    : `|X|^2 > T2` bits straight from __ballot_sync (bit = lane), row flags for the top_db clamp
//   non-stationary: per (tile, bin) the two weighted sums of |X| from which the scan kernel rebuilds the forward /
//                   backward one-pole states at every tile boundary (see ns_scan_kernel):
//     P = sum_s b a^(n-1-s) |X_s|                            zero-state forward response at the end of the tile
//     Q = sum_s b^2 a^s (1 - a^(2(n-s))) / (1 - a^2) |X_s|   zero-state backward response (of the zero-state forward
//                                                            one) at the start of the tile;  n = frames in the tile
// ============================================================================================================
__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

constexpr int STFT_WARPS = 8;
constexpr int STFT_SPAN = (STFT_TILE + 3) * HOP;                          // samples covered by a tile's frames
constexpr size_t STFT_SMEM = sizeof(c32) * (STFT_WARPS * FFT1024_SMEM + 1024) + 2 * sizeof(int16_t) * STFT_SPAN;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}

// Stage the contiguous int16 span of tile (chunk c, frames f0 ..) into `span`: 16-byte cp.async where the group lies
// inside the padded chunk and the recording (asynchronous: no register staging, no stall), plain stores of the zero-
// filled samples at the edges.
__device__ __forceinline__ void stft_stage_span(int16_t* span, const int16_t* __restrict__ pcm, long long it_pcm_off, long long it_buf_len,
                                                const Chunk& ch, int f0, int tid) {
    const long long j0 = (long long)f0 * HOP - NFFT / 2;              // padded-chunk coordinate of span[0]
    const long long li0 = ch.s_local - PADS + j0;                     // local buffer index of span[0]
    const int16_t* base = pcm + it_pcm_off;
    if (j0 >= 0 && j0 + STFT_SPAN <= ch.L && li0 >= 0 && li0 + STFT_SPAN <= it_buf_len && ((it_pcm_off + li0) & 7) == 0) {
        // the whole span lies inside the padded chunk and the recording (all but a chunk's first / last tiles): no per-group tests
        const int16_t* src = base + li0;
#pragma unroll
        for (int g = tid; g < STFT_SPAN / 8; g += 32 * STFT_WARPS) cp_async16(&span[8 * g], src + 8 * g);
        return;
    }
    for (int g = tid; g < STFT_SPAN / 8; g += 32 * STFT_WARPS) {
        const long long j = j0 + 8 * g, li = li0 + 8 * g;
        if (j >= 0 && j + 8 <= ch.L && li >= 0 && li + 8 <= it_buf_len && ((it_pcm_off + li) & 7) == 0) {
            cp_async16(&span[8 * g], base + li);
        } else {
            int16_t tmp[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const long long jj = j + e, ll = li + e;
                tmp[e] = (jj >= 0 && jj < ch.L && ll >= 0 && ll < it_buf_len) ? base[ll] : (int16_t)0;
            }
            *reinterpret_cast<uint4*>(&span[8 * g]) = *reinterpret_cast<uint4*>(tmp);
        }
    }
}

// Persistent CTAs (two per SM) walk the (chunk, tile) list; the span of the NEXT tile is in flight (cp.async, double
// buffer) while the warps transform the current one, so the only exposed latency is the first tile's.
template <bool STATIONARY>
__global__ void __launch_bounds__(32 * STFT_WARPS, 2) stft1024_kernel(const int16_t* __restrict__ pcm, const Item* __restrict__ items,
                                                                     const Chunk* __restrict__ chunks, int n_chunks, int tiles_per_chunk,
                                                                     Consts cs, const float* __restrict__ T2, const float* __restrict__ T2hi,
                                                                     c32* __restrict__ S, uint32_t* __restrict__ bits,
                                                                     uint8_t* __restrict__ rowflag, const NsParams* __restrict__ nsp,
                                                                     float* __restrict__ P, float* __restrict__ Q) {
    extern __shared__ __align__(16) unsigned char stft_smem[];
    c32* scratch = reinterpret_cast<c32*>(stft_smem);                     // [STFT_WARPS][FFT1024_SMEM]
    c32* tww = scratch + STFT_WARPS * FFT1024_SMEM;                       // [1024]
    int16_t* span0 = reinterpret_cast<int16_t*>(tww + 1024);              // [2][STFT_SPAN]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    c32* sm = scratch + warp * FFT1024_SMEM;
    const int pl = (32 - lane) & 31;                                  // lane holding bins 1024 - k of this lane's bins k
    for (int i = tid; i < 1024; i += 32 * STFT_WARPS) tww[i] = cs.tww[i];
    __shared__ float wtab[2][STATIONARY ? 1 : STFT_TILE];             // scan weights of a full tile (n = STFT_TILE)
    if (!STATIONARY && tid < STFT_TILE) {
        wtab[0][tid] = (float)(nsp->b * nsp->apow[STFT_TILE - 1 - tid]);
        wtab[1][tid] = (float)(nsp->b * nsp->b * nsp->apow[tid] * (1.0 - nsp->apow[2 * (STFT_TILE - tid)]) * nsp->inv_1ma2);
    }
    // periodic Hann / 1024 at n = lane + 32 j:  h - h cos(2 pi lane / 1024 + 2 pi j / 32) = h + wc cos(2 pi j / 32) + ws sin(2 pi j / 32)
    float wc, ws;
    {
        float sl, cl;
        sincospif((float)lane * (1.0f / 512.0f), &sl, &cl);
        wc = -cl * (0.5f / 1024.0f); ws = sl * (0.5f / 1024.0f);
    }

    const unsigned total = (unsigned)n_chunks * (unsigned)tiles_per_chunk;    // < 2^31 (checked by the launcher): 32-bit index math
    // The descriptors of this CTA's first STFT_DESC tiles (grid-strided walk: a different chunk almost every step) are fetched
    // ONCE, all loads in flight together, instead of two dependent global loads at the top of every tile.
    constexpr int STFT_DESC = 32;
    __shared__ Chunk s_ch[STFT_DESC];
    __shared__ long long s_io[STFT_DESC][2];                              // the item's pcm_off, buf_len
    for (int i = tid; i < STFT_DESC; i += 32 * STFT_WARPS) {
        const unsigned tt = blockIdx.x + (unsigned)i * gridDim.x;
        if (tt < total) {
            const Chunk v = chunks[tt / (unsigned)tiles_per_chunk];
            s_ch[i] = v;
            s_io[i][0] = items[v.item].pcm_off; s_io[i][1] = items[v.item].buf_len;
        }
    }
    __syncthreads();
    auto tile_valid = [&](unsigned t, Chunk& ch, int& f0, long long& po, long long& bl) {
        const unsigned c = t / (unsigned)tiles_per_chunk;
        const unsigned i = (t - blockIdx.x) / gridDim.x;
        if (i < (unsigned)STFT_DESC) { ch = s_ch[i]; po = s_io[i][0]; bl = s_io[i][1]; }
        else { ch = chunks[c]; po = items[ch.item].pcm_off; bl = items[ch.item].buf_len; }
        f0 = (int)(t - c * (unsigned)tiles_per_chunk) * STFT_TILE;
        return f0 < ch.n_frames;
    };
    // first valid tile of this CTA
    unsigned t = blockIdx.x;
    Chunk ch; int f0 = 0; long long po = 0, bl = 0;
    while (t < total && !tile_valid(t, ch, f0, po, bl)) t += gridDim.x;
    if (t >= total) return;
    int buf = 0;
    stft_stage_span(span0, pcm, po, bl, ch, f0, tid);
    asm volatile("cp.async.commit_group;" ::: "memory");

    while (t < total) {
        // next valid tile: start its span load into the other buffer
        unsigned tn = t + gridDim.x;
        Chunk chn; int f0n = 0;
        while (tn < total && !tile_valid(tn, chn, f0n, po, bl)) tn += gridDim.x;
        if (tn < total) stft_stage_span(span0 + (buf ^ 1) * STFT_SPAN, pcm, po, bl, chn, f0n, tid);
        asm volatile("cp.async.commit_group;\n cp.async.wait_group 1;" ::: "memory");
        __syncthreads();                                              // current span landed; everyone is done with the previous tile's scratch

        const int16_t* span = span0 + buf * STFT_SPAN;
        const int c = (int)(t / (unsigned)tiles_per_chunk), tile = (int)(t - (unsigned)c * (unsigned)tiles_per_chunk);
        const int n_tile = min(STFT_TILE, ch.n_frames - f0);
        const float* t2 = STATIONARY ? T2 + (long long)ch.item * T_PITCH : nullptr;
        const float* t2hi = STATIONARY ? T2hi + (long long)ch.item * T_PITCH : nullptr;
        uint8_t* rf = STATIONARY ? rowflag + (long long)c * M_PITCH : nullptr;
        float accP[17], accQ[17];                                     // bins 32 k1 + lane -> k1, bin 512 (lane 0) -> 16
#pragma unroll
        for (int i = 0; i < 17; ++i) { accP[i] = 0.f; accQ[i] = 0.f; }

        for (int pr = warp; pr < STFT_TILE / 2; pr += STFT_WARPS) {
            const int lfA = 2 * pr, fA = f0 + lfA;
            if (fA >= ch.n_frames) break;                             // warp-uniform
            const bool actB = fA + 1 < ch.n_frames;
            c32 v[32];
            {
                // frame B is frame A shifted by 256 samples = 8 strides of 32: 40 loads serve both
                const int16_t* sp = span + lfA * HOP + lane;
                float s[40];
#pragma unroll
                for (int m = 0; m < 40; ++m) s[m] = int_to_float_exact((int)sp[32 * m]);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float w = fmaf(wc, w32(j).x, fmaf(ws, -w32(j).y, 0.5f / 1024.0f));
                    v[j] = mk(s[j] * w, actB ? s[j + 8] * w : 0.f);
                }
            }
            fft1024w_pass1(v, sm, tww, lane);
            __syncwarp();
            fft1024w_pass2(v, sm, lane);
            __syncwarp();
            // v[k1] = Z[32 k1 + lane], Z = X_A + i X_B (the window carries the 1/2):  X_A = Z[k] + conj Z[1024-k],  X_B = -i (Z[k] - conj Z[1024-k])
            c32* rowA = S + (ch.row0 + fA) * S_PITCH;
            c32* rowB = rowA + S_PITCH;
            float wpA = 0.f, wqA = 0.f, wpB = 0.f, wqB = 0.f;
            if (!STATIONARY && n_tile == STFT_TILE) {
                wpA = wtab[0][STATIONARY ? 0 : lfA]; wqA = wtab[1][STATIONARY ? 0 : lfA];
                wpB = wtab[0][STATIONARY ? 0 : lfA + 1]; wqB = wtab[1][STATIONARY ? 0 : lfA + 1];
            } else if (!STATIONARY) {
                wpA = (float)(nsp->b * nsp->apow[n_tile - 1 - lfA]);
                wqA = (float)(nsp->b * nsp->b * nsp->apow[lfA] * (1.0 - nsp->apow[2 * (n_tile - lfA)]) * nsp->inv_1ma2);
                if (actB) {
                    wpB = (float)(nsp->b * nsp->apow[n_tile - 2 - lfA]);
                    wqB = (float)(nsp->b * nsp->b * nsp->apow[lfA + 1] * (1.0 - nsp->apow[2 * (n_tile - lfA - 1)]) * nsp->inv_1ma2);
                }
            }
            uint32_t wordA = 0, wordB = 0;
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) {
                const int k = 32 * k1 + lane;
                c32 zp;
                zp.x = __shfl_sync(0xffffffffu, v[31 - k1].x, pl);
                zp.y = __shfl_sync(0xffffffffu, v[31 - k1].y, pl);
                if (lane == 0) zp = v[(32 - k1) & 31];                // 1024 - 32 k1 = 32 (32 - k1): lane 0 pairs with itself
                const c32 z = v[k1];
                c32 xa = mk(z.x + zp.x, z.y - zp.y), xb = mk(z.y + zp.y, zp.x - z.x);
                float magA2, magB2, a512 = 0.f, b512 = 0.f;
                if (k1 == 0 && lane == 0) {
                    // bins 0 and 512 are real and share slot 0: (re X_0, re X_512)
                    a512 = 2.f * v[16].x; b512 = 2.f * v[16].y;
                    xa = mk(xa.x, a512); xb = mk(xb.x, b512);
                    magA2 = xa.x * xa.x; magB2 = xb.x * xb.x;
                } else {
                    magA2 = xa.x * xa.x + xa.y * xa.y; magB2 = xb.x * xb.x + xb.y * xb.y;
                }
                rowA[k] = xa;
                if (actB) rowB[k] = xb;
                if (STATIONARY) {
                    const float th = t2[k], thh = t2hi[k];
                    const uint32_t ba = __ballot_sync(0xffffffffu, magA2 > th), bb = __ballot_sync(0xffffffffu, actB && magB2 > th);
                    if (lane == k1) { wordA = ba; wordB = bb; }
                    if (magA2 > thh || (actB && magB2 > thh)) rf[k] = 1;
                    if (k1 == 0) {
                        const float th5 = t2[512], thh5 = t2hi[512];
                        const uint32_t ca = __ballot_sync(0xffffffffu, lane == 0 && a512 * a512 > th5);
                        const uint32_t cb = __ballot_sync(0xffffffffu, lane == 0 && actB && b512 * b512 > th5);
                        if (lane == 16) { wordA = ca & 1u; wordB = cb & 1u; }
                        if (lane == 0 && (a512 * a512 > thh5 || (actB && b512 * b512 > thh5))) rf[512] = 1;
                    }
                } else {
                    const float ma = fast_sqrt(magA2), mb = fast_sqrt(magB2);
                    accP[k1] = fmaf(wpB, mb, fmaf(wpA, ma, accP[k1]));
                    accQ[k1] = fmaf(wqB, mb, fmaf(wqA, ma, accQ[k1]));
                    if (k1 == 0) {
                        accP[16] = fmaf(wpB, fabsf(b512), fmaf(wpA, fabsf(a512), accP[16]));
                        accQ[16] = fmaf(wqB, fabsf(b512), fmaf(wqA, fabsf(a512), accQ[16]));
                    }
                }
            }
            if (STATIONARY && lane < 17) {
                bits[(ch.row0 + fA) * B_PITCH + lane] = wordA;
                if (actB) bits[(ch.row0 + fA + 1) * B_PITCH + lane] = wordB;
            }
        }
        if (!STATIONARY) {
            // combine the warps' partial sums in a fixed order (a warp's scratch is its own until the barrier)
            float* gp = reinterpret_cast<float*>(sm);
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) { gp[32 * k1 + lane] = accP[k1]; gp[T_PITCH + 32 * k1 + lane] = accQ[k1]; }
            if (lane == 0) { gp[512] = accP[16]; gp[T_PITCH + 512] = accQ[16]; }
            __syncthreads();
            // threads 0..127 add the P rows, 128..255 the Q rows, four bins each (16-byte shared loads, one 16-byte store;
            // same warp order as a scalar loop, so the sums are bit-identical); bin 512 by the first thread of each half
            const long long trow = (ch.tile0 + tile) * T_PITCH;
            {
                const int half = tid >> 7, k4 = 4 * (tid & 127);
                float* dst = (half ? Q : P) + trow;
                float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
                float a512 = 0.f;
#pragma unroll
                for (int w = 0; w < STFT_WARPS; ++w) {
                    const float* g = reinterpret_cast<const float*>(scratch + w * FFT1024_SMEM) + half * T_PITCH;
                    const float4 x = *reinterpret_cast<const float4*>(g + k4);
                    acc4.x += x.x; acc4.y += x.y; acc4.z += x.z; acc4.w += x.w;
                    if (k4 == 0) a512 += g[512];
                }
                *reinterpret_cast<float4*>(dst + k4) = acc4;
                if (k4 == 0) dst[512] = a512;
            }
        }
        if (STATIONARY) __syncthreads();                              // every warp is done with this span before it is refilled
        t = tn; ch = chn; f0 = f0n; buf ^= 1;
    }
}

int launch_stft(const int16_t* pcm, const Item* items, const Chunk* chunks, int n_chunks, int max_frames,
                 const Consts& c, int stationary, const float* T2, const float* T2hi, c32* S, uint32_t* bits,
                 uint8_t* rowflag, const NsParams* ns, float* P, float* Q, cudaStream_t st) {
    if (n_chunks <= 0) return 0;
    const int tiles = (max_frames + STFT_TILE - 1) / STFT_TILE;
    static int n_sm = 0;
    if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
    const long long total = (long long)n_chunks * tiles;
    if (total >= (1LL << 31) - 4096) return -1;                   // maf_plan_create rejects such plans; never skip work silently
    const int grid = (int)std::min<long long>(total, 2LL * n_sm);
    cudaFuncSetAttribute(stft1024_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STFT_SMEM);
    cudaFuncSetAttribute(stft1024_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STFT_SMEM);
    if (stationary)
        stft1024_kernel<true><<<grid, 32 * STFT_WARPS, STFT_SMEM, st>>>(pcm, items, chunks, n_chunks, tiles, c, T2, T2hi, S, bits, rowflag, nullptr, nullptr, nullptr);
    else
        stft1024_kernel<false><<<grid, 32 * STFT_WARPS, STFT_SMEM, st>>>(pcm, items, chunks, n_chunks, tiles, c, nullptr, nullptr, S, nullptr, nullptr, ns, P, Q);
    return 0;
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
    // a clip shorter than n_fft gets scipy's shrunken window / hop (see clip_nperseg in maf_common.cuh)
    const int np = clip_nperseg(cl.len), step = np - (NFFT - HOP), half = np / 2;
    if (np == NFFT) {
        for (int i = tid; i < NFFT; i += 256) win[i] = cs.win1024[i];
    } else {
        for (int i = tid; i < NFFT; i += 256)                         // periodic Hann(np) / sum, zero-padded to n_fft at the end
            win[i] = i < np ? (float)((0.5 - 0.5 * cospi(2.0 * (double)i / (double)np)) / (0.5 * (double)np)) : 0.f;
    }
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
            const long long i = (long long)f * step - half + 2 * n;        // clip-relative sample index
            float a = 0.f, b = 0.f;
            if (active && 2 * n < np) {                                    // (the window is zero beyond np anyway; never read past the segment)
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

// 16 bins x 8 part-groups per CTA: the partial sums of a bin are split over 8 threads (loads in flight together), the
// 8 group sums folded in a fixed order -- the summation order is fixed, so the thresholds are reproducible run to run.
__global__ void __launch_bounds__(128) noise_finalize_kernel(const Clip* __restrict__ clips, const double* __restrict__ partial, int n_parts,
                                                             double n_std, float* __restrict__ T2, float* __restrict__ T2hi) {
    __shared__ double sh1[8][16], sh2[8][16];
    const int item = blockIdx.y;
    const int kl = threadIdx.x & 15, sub = threadIdx.x >> 4;
    const int k = blockIdx.x * 16 + kl;
    const int per = (n_parts + 7) / 8;
    const int p0 = min(n_parts, sub * per), p1 = min(n_parts, p0 + per);
    double s1 = 0.0, s2 = 0.0;
    if (k < NBIN) {
        const double* pr0 = partial + (long long)item * n_parts * (2 * T_PITCH) + k;
        int p = p0;
        for (; p + 8 <= p1; p += 8) {                  // loads first (independent), fixed summation order kept
            double a[8], b[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { a[u] = pr0[(long long)(p + u) * (2 * T_PITCH)]; b[u] = pr0[(long long)(p + u) * (2 * T_PITCH) + T_PITCH]; }
#pragma unroll
            for (int u = 0; u < 8; ++u) { s1 += a[u]; s2 += b[u]; }
        }
        for (; p < p1; ++p) { s1 += pr0[(long long)p * (2 * T_PITCH)]; s2 += pr0[(long long)p * (2 * T_PITCH) + T_PITCH]; }
    }
    sh1[sub][kl] = s1; sh2[sub][kl] = s2;
    __syncthreads();
    if (sub != 0 || k >= NBIN) return;
#pragma unroll
    for (int u = 1; u < 8; ++u) { s1 += sh1[u][kl]; s2 += sh2[u][kl]; }
    const int n = clips[item].n_frames;
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
    launch_zero_bytes(nmax, sizeof(unsigned) * (size_t)n_items * T_PITCH, st);
    // `partial` needs no clearing: pass 2 writes every (slice, group, bin) entry, zeros included
    dim3 grid(n_slices, n_items);
    noise_pass_kernel<1><<<grid, 256, 0, st>>>(pcm, clips, c, nmax, partial, n_slices);
    noise_pass_kernel<2><<<grid, 256, 0, st>>>(pcm, clips, c, nmax, partial, n_slices);
    dim3 g2((NBIN + 15) / 16, n_items);
    noise_finalize_kernel<<<g2, 128, 0, st>>>(clips, partial, n_slices * 4, n_std, T2, T2hi);
}

// ============================================================================================================
// K2a  non-stationary mask (noisereduce spectralgate/nonstationary.py): Sm = filtfilt([b], [1, b-1], |S|) along time
// (padtype=None: both passes start in steady state), raw = sigmoid(((|S| - Sm) / Sm - 2) * 10), then the same
// separable triangular smoothing as the stationary gate and the prop_decrease blend.
//
// The forward-backward one-pole recurrence spans the whole chunk (pole 0.992: no usable warm-up length), so it is
// evaluated as a two-level scan instead of a serial sweep over HBM:
//   level 1 (inside stft1024_kernel)  per 16-frame tile two weighted sums P, Q of |X|
//   level 2 (ns_scan_kernel)          one thread per (chunk, bin) walks the ~162 tiles twice in float64:
//                                       y_e(j) = a^n y_e(j-1) + P_j                     y_e(-1) = |X_0|
//                                       z_e(j-1) = Q_j + c_n y_e(j-1) + a^n z_e(j)      z_e(last) = y_e(last)
//                                     with c_n = b a (1 - a^2n) / (1 - a^2); y_e(j) = yf at the tile's last frame,
//                                     z_e(j) = backward state entering the tile from the right
//   level 3 (ns_mask_smooth_kernel)   a CTA owns 32 output frames; one thread per bin walks the 32 + 2 nt frames it
//                                     needs DOWNWARDS from a tile boundary: z = b yf + a z, yf' = (yf - b |X|) / a
//                                     (the forward recurrence run in reverse: error growth a^-38 = 1.36), sigmoid,
//                                     time taps in a register window, rows to shared memory; then the frequency
//                                     taps as in smooth_mask_kernel.  S is read once (+ 2 nt / 32 halo), M written once.
// Deviation (documented in DESIGN.md): where Sm == 0 (an all-zero frequency row) noisereduce yields NaN; here the
// relative excess is taken as 0 so the mask is sigmoid(-20) and the output is silence instead of NaN.
// ============================================================================================================
__global__ void __launch_bounds__(64) ns_scan_kernel(const c32* __restrict__ S, const Chunk* __restrict__ chunks, int n_chunks,
                                                      NsParams ns, const float* __restrict__ P, const float* __restrict__ Q,
                                                      float* __restrict__ Yend, float* __restrict__ Zend) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = (int)(idx / NBIN), k = (int)(idx - (long long)c * NBIN);
    if (c >= n_chunks) return;
    const Chunk ch = chunks[c];
    const int F = ch.n_frames;
    const int nt = (F + STFT_TILE - 1) / STFT_TILE;
    const int n_last = F - (nt - 1) * STFT_TILE;
    const c32 s0 = S[ch.row0 * S_PITCH + (k & 511)];
    const double a0 = (double)(k == 0 ? fabsf(s0.x) : k == 512 ? fabsf(s0.y) : fast_sqrt(s0.x * s0.x + s0.y * s0.y));
    const double aN = ns.apow[STFT_TILE], aL = ns.apow[n_last];
    const double cN = ns.b * ns.a * (1.0 - ns.apow[2 * STFT_TILE]) * ns.inv_1ma2;
    const double cL = ns.b * ns.a * (1.0 - ns.apow[2 * n_last]) * ns.inv_1ma2;
    const float* p = P + ch.tile0 * T_PITCH + k;
    const float* q = Q + ch.tile0 * T_PITCH + k;
    float* ye = Yend + ch.tile0 * T_PITCH + k;
    float* ze = Zend + ch.tile0 * T_PITCH + k;
    // Loads do not depend on the recurrence: they are issued SCAN_B tiles ahead (double-buffered batches) so that the
    // walk is bound by the FP64 chain, not by one memory latency per tile.
    constexpr int SCAN_B = 16;
    const int nfull = nt - 1;                                  // tiles of 16 frames; the last one has n_last
    double y = a0;
    {
        float cur[SCAN_B], nxt[SCAN_B];
#pragma unroll
        for (int u = 0; u < SCAN_B; ++u) cur[u] = u < nfull ? p[(long long)u * T_PITCH] : 0.f;
        for (int j = 0; j < nfull; j += SCAN_B) {
#pragma unroll
            for (int u = 0; u < SCAN_B; ++u) nxt[u] = j + SCAN_B + u < nfull ? p[(long long)(j + SCAN_B + u) * T_PITCH] : 0.f;
#pragma unroll
            for (int u = 0; u < SCAN_B; ++u) {
                if (j + u < nfull) { y = aN * y + (double)cur[u]; ye[(long long)(j + u) * T_PITCH] = (float)y; }
            }
#pragma unroll
            for (int u = 0; u < SCAN_B; ++u) cur[u] = nxt[u];
        }
    }
    const double y_before_last = y;
    y = aL * y + (double)p[(long long)(nt - 1) * T_PITCH];
    ye[(long long)(nt - 1) * T_PITCH] = (float)y;
    // backward: z enters tile j from the right
    double z = (double)(float)y;
    ze[(long long)(nt - 1) * T_PITCH] = (float)z;
    z = (double)q[(long long)(nt - 1) * T_PITCH] + cL * (nt > 1 ? (double)(float)y_before_last : a0) + aL * z;
    {
        // tile j = nfull - 1 - i (i ascending); y_e(j - 1) is re-read from the array written above (j = 0 uses |X_0|)
        float cq[SCAN_B], cy[SCAN_B], nq[SCAN_B], ny[SCAN_B];
#pragma unroll
        for (int u = 0; u < SCAN_B; ++u) {
            const int j = nfull - 1 - u;
            cq[u] = j >= 0 ? q[(long long)j * T_PITCH] : 0.f;
            cy[u] = j >= 1 ? ye[(long long)(j - 1) * T_PITCH] : (float)a0;
        }
        for (int i = 0; i < nfull; i += SCAN_B) {
#pragma unroll
            for (int u = 0; u < SCAN_B; ++u) {
                const int j = nfull - 1 - (i + SCAN_B + u);
                nq[u] = j >= 0 ? q[(long long)j * T_PITCH] : 0.f;
                ny[u] = j >= 1 ? ye[(long long)(j - 1) * T_PITCH] : (float)a0;
            }
#pragma unroll
            for (int u = 0; u < SCAN_B; ++u) {
                const int j = nfull - 1 - (i + u);
                if (j >= 0) { ze[(long long)j * T_PITCH] = (float)z; z = (double)cq[u] + cN * (double)cy[u] + aN * z; }
            }
#pragma unroll
            for (int u = 0; u < SCAN_B; ++u) { cq[u] = nq[u]; cy[u] = ny[u]; }
        }
    }
}

constexpr int SM_ROW = MAX_NF + 520 + MAX_NF;     // 584 floats, multiple of 4: a mask row with zero halos for the frequency taps
constexpr int NSM_THREADS = 544;                  // 17 warps: one thread per bin in the recurrence phase
constexpr size_t NSM_SMEM = sizeof(float) * NS_TILE * SM_ROW;

// Column swizzle of the mask rows in shared memory: in every other block of 32 floats adjacent float4 groups trade places.
// The frequency taps read float4s at a stride of two groups per lane -- unswizzled, the eight lanes of a quarter warp fall on
// four bank groups (a 2-way conflict on every load); swizzled they cover all eight.  Writes of consecutive columns stay
// conflict-free (the swizzle permutes inside a 32-float block).
__device__ __forceinline__ int swz(int c) { return c ^ (((c >> 5) & 1) << 2); }

// frequency taps on one shared-memory row (`row` = start of the row, outputs = bins k0 .. k0 + 7, i.e. columns MAX_NF + k0 ...)
__device__ __forceinline__ void freq_taps8(const float* row, int k0, int nf, const float* vf, float (&out)[8]) {
#pragma unroll
    for (int o = 0; o < 8; ++o) out[o] = 0.f;
    if (nf == 16) {
        // triangle(33) = box(17) * box(17) / 289: two running sums restarted per work item (<= 23 updates each)
        float in[40];
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const float4 q = *reinterpret_cast<const float4*>(row + swz(MAX_NF + k0 - 16 + 4 * i));
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
            const float wt = vf[tp];
#pragma unroll
            for (int o = 0; o < 8; ++o) out[o] += wt * row[swz(MAX_NF + k0 + o + tp - nf)];
        }
    }
}

// Recurrence phase of one bin: walks W = 32 + 2 NT frames downwards from frame base + 31, writes the time-smoothed raw
// mask of the 32 output frames into column k of `rows`.  INTERIOR = every walked frame lies inside the chunk.
template <int NT, bool INTERIOR>
__device__ __forceinline__ void ns_walk(const c32* __restrict__ sp, int k, int base, int F, float yend, float zend, const NsParams& ns,
                                        const SmoothTaps& taps, float* __restrict__ rows) {
    constexpr int W = NS_TILE + 2 * NT;
    constexpr int NB = 4;                                     // frames per load batch (double-buffered)
    // with u = b y:  z_r = u_r + a z_(r+1),  u_(r-1) = u_r / a - (b^2 / a) |X_r|.  The walk is at most 32 + 2 nt steps
    // from float32-rounded states, so float32 arithmetic adds ~4e-7 relative to Sm (mask error < 3e-6, -110 dB) and keeps
    // the quarter-rate conversion unit (F2F) off the per-bin path.
    const float a = (float)ns.a, ia = (float)ns.inv_a, b2ia = (float)(ns.b * ns.b * ns.inv_a);
    // sigmoid(((A - sm) / sm - thresh) * slope) = 1 / (1 + 2^(c0 - c1 A / sm))
    const float c1 = ns.slope * 1.4426950408889634f, c0 = ns.slope * (1.0f + ns.thresh) * 1.4426950408889634f;
    // bins 0 and 512 share slot 0 as (re X_0, re X_512): select the component, zero the other one
    const bool use_y = k == 512, real_only = k == 0 || k == 512;
    float w[2 * NT + 1];
#pragma unroll
    for (int i = 0; i < 2 * NT + 1; ++i) w[i] = 0.f;
    float vt[2 * NT + 1];
#pragma unroll
    for (int i = 0; i < 2 * NT + 1; ++i) vt[i] = taps.vt[i];
    c32 cur[NB], nxt[NB];
    auto load = [&](c32 (&dst)[NB], int b0) {
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            const int rr = base + NS_TILE - 1 - (b0 + q);
            if (b0 + q < W) dst[q] = (INTERIOR || (rr >= 0 && rr < F)) ? sp[(long long)rr * S_PITCH] : mk(0.f, 0.f);
            // pull the rows of three batches ahead into L2 (the walk streams S from DRAM exactly once)
            const int rp = rr - 3 * NB;
            if (b0 + q + 3 * NB < W && (INTERIOR || (rp >= 0 && rp < F))) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + (long long)rp * S_PITCH));
        }
    };
    if (INTERIOR || base + NS_TILE - 1 - 3 * NB >= 0) {
#pragma unroll
        for (int q = 0; q < 3 * NB; ++q) {
            const int rr = min(base + NS_TILE - 1, F - 1) - q;
            if (rr >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + (long long)rr * S_PITCH));
        }
    }
    load(cur, 0);
    // first use of the tile-end states: only now, with the L2 prefetches and the first batch of S loads already in flight (the
    // Yend / Zend loads of the caller used to be waited for before anything else was issued)
    float u = (float)(ns.b * (double)yend);
    float zf = zend;
#pragma unroll
    for (int b0 = 0; b0 < W; b0 += NB) {
        if (b0 + NB < W) load(nxt, b0 + NB);
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            const int i = b0 + q;
            if (i < W) {
                const int rr = base + NS_TILE - 1 - i;
                float R = 0.f;
                if (INTERIOR || (rr >= 0 && rr < F)) {
                    const float xr = use_y ? cur[q].y : cur[q].x, xi = real_only ? 0.f : cur[q].y;
                    const float A = fast_sqrt(fmaf(xr, xr, xi * xi));
                    zf = fmaf(a, zf, u);
                    u = fmaf(u, ia, -b2ia * A);
                    const float sm = zf;
                    float inv;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(sm));
                    const float ratio = sm > 0.f ? A * inv : 1.0f;
                    float ex;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fmaf(-c1, ratio, c0)));
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(R) : "f"(1.0f + ex));
                }
#pragma unroll
                for (int j = 2 * NT; j > 0; --j) w[j] = w[j - 1];
                w[0] = R;                                     // w[j] = raw mask of frame rr + j
                if (i >= 2 * NT) {
                    const int lr = NS_TILE - 1 + 2 * NT - i;  // output frame rr + NT = base - NT + lr
                    float acc = vt[0] * w[0];
#pragma unroll
                    for (int j = 1; j < 2 * NT + 1; ++j) acc = fmaf(vt[j], w[j], acc);
                    rows[lr * SM_ROW + swz(MAX_NF + k)] = acc;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < NB; ++q) cur[q] = nxt[q];
    }
}

template <int NT>
__global__ void __launch_bounds__(NSM_THREADS, 2) ns_mask_smooth_kernel(const c32* __restrict__ S, const Chunk* __restrict__ chunks,
                                                                        NsParams ns, const float* __restrict__ Yend,
                                                                        const float* __restrict__ Zend, float prop, SmoothTaps taps,
                                                                        float* __restrict__ M) {
    extern __shared__ __align__(16) float rows[];                 // [NS_TILE][SM_ROW]
    // chunks from the LAST to the first (tiles of a chunk ascending): the forward STFT wrote the high chunks last, so their S rows
    // are still in the 126 MB L2, and the iSTFT that follows starts at chunk 0, where this kernel's last M rows and S reads are
    // (A/B on one box: mask -0.5 %, iSTFT -0.5 %; box-to-box variation of this kernel is ten times that)
    const int by = gridDim.y - 1 - blockIdx.y, bx = blockIdx.x;
    const Chunk ch = chunks[by];
    const int F = ch.n_frames;
    const int base = NS_TILE * bx;                                // outputs [base - NT, base + 32 - NT), walked frames [base - 2 NT, base + 32)
    if (base - NT >= F) return;
    const int tid = threadIdx.x;
    // zero halos left and right of the 513 bins
    for (int i = tid; i < NS_TILE * (SM_ROW - NBIN); i += NSM_THREADS) {
        const int r = i / (SM_ROW - NBIN), c = i - r * (SM_ROW - NBIN);
        rows[r * SM_ROW + swz(c < MAX_NF ? c : c + NBIN)] = 0.f;
    }
    if (tid < NBIN) {
        const int k = tid;
        const int e = min(base + NS_TILE, F);                    // the walk starts at frame e - 1; states are known at tile ends
        const long long srow = (ch.tile0 + (e + STFT_TILE - 1) / STFT_TILE - 1) * T_PITCH + k;
        const float y = Yend[srow], z = Zend[srow];                // (converted at their first use inside the walk, after its loads are issued)
        const c32* sp = S + ch.row0 * S_PITCH + (k & 511);
        if (base - 2 * NT >= 0 && base + NS_TILE <= F) ns_walk<NT, true>(sp, k, base, F, y, z, ns, taps, rows);
        else ns_walk<NT, false>(sp, k, base, F, y, z, ns, taps, rows);
    }
    __syncthreads();
    // frequency taps + blend, 8 consecutive outputs per work item
    const int nf = taps.nf;
    const float lo = 1.0f - prop;
    for (int wi = tid; wi < NS_TILE * 65; wi += NSM_THREADS) {
        const int lr = wi / 65, seg = wi - lr * 65;
        const int f = base - NT + lr;
        if (f < 0 || f >= F) continue;
        const int k0 = 8 * seg;
        float out[8];
        freq_taps8(&rows[lr * SM_ROW], k0, nf, taps.vf, out);
        float* dst = M + (ch.row0 + f) * M_PITCH + k0;
        if (k0 + 8 <= NBIN) {
            *reinterpret_cast<float4*>(dst) = make_float4(fmaf(out[0], prop, lo), fmaf(out[1], prop, lo), fmaf(out[2], prop, lo), fmaf(out[3], prop, lo));
            *reinterpret_cast<float4*>(dst + 4) = make_float4(fmaf(out[4], prop, lo), fmaf(out[5], prop, lo), fmaf(out[6], prop, lo), fmaf(out[7], prop, lo));
        } else {
            dst[0] = fmaf(out[0], prop, lo);                      // bin 512
        }
    }
}

template <int NT>
static void launch_nsm(dim3 grid, cudaStream_t st, const c32* S, const Chunk* chunks, const NsParams& ns, const float* Yend,
                       const float* Zend, float prop, const SmoothTaps& taps, float* M) {
    cudaFuncSetAttribute(ns_mask_smooth_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NSM_SMEM);
    ns_mask_smooth_kernel<NT><<<grid, NSM_THREADS, NSM_SMEM, st>>>(S, chunks, ns, Yend, Zend, prop, taps, M);
}

void launch_ns_mask(const c32* S, const Chunk* chunks, int n_chunks, int max_frames, const NsParams& ns, const float* P,
                    const float* Q, float* Yend, float* Zend, float prop, const SmoothTaps& taps, float* M,
                    cudaStream_t st) {
    if (n_chunks <= 0) return;
    {
        const long long threads = (long long)n_chunks * NBIN;
        ns_scan_kernel<<<(unsigned)((threads + 63) / 64), 64, 0, st>>>(S, chunks, n_chunks, ns, P, Q, Yend, Zend);
    }
    const int tiles = (max_frames + taps.nt + NS_TILE - 1) / NS_TILE;
    for (int y0 = 0; y0 < n_chunks; y0 += 32768) {
        const int ny = n_chunks - y0 < 32768 ? n_chunks - y0 : 32768;
        dim3 grid(tiles, ny);
        switch (taps.nt) {
            case 1: launch_nsm<1>(grid, st, S, chunks + y0, ns, Yend, Zend, prop, taps, M); break;
            case 2: launch_nsm<2>(grid, st, S, chunks + y0, ns, Yend, Zend, prop, taps, M); break;
            case 3: launch_nsm<3>(grid, st, S, chunks + y0, ns, Yend, Zend, prop, taps, M); break;
            case 4: launch_nsm<4>(grid, st, S, chunks + y0, ns, Yend, Zend, prop, taps, M); break;
            case 5: launch_nsm<5>(grid, st, S, chunks + y0, ns, Yend, Zend, prop, taps, M); break;
            case 6: launch_nsm<6>(grid, st, S, chunks + y0, ns, Yend, Zend, prop, taps, M); break;
            case 7: launch_nsm<7>(grid, st, S, chunks + y0, ns, Yend, Zend, prop, taps, M); break;
            case 8: launch_nsm<8>(grid, st, S, chunks + y0, ns, Yend, Zend, prop, taps, M); break;
            default: launch_nsm<9>(grid, st, S, chunks + y0, ns, Yend, Zend, prop, taps, M); break;
        }
    }
}

// ============================================================================================================
// K2  separable triangular smoothing of the stationary gate (scipy.signal.fftconvolve(mask, outer(vf, vt)/sum, 'same'),
// zero padded): raw = bit | rowflag ? 1 : 1 - prop (blend, then smooth).  Same structure as the non-stationary mask
// kernel: a CTA owns 32 output frames, one thread per bin walks the 32 + 2 nt raw rows (bit words staged in shared memory,
// one broadcast load per row), applies the time taps in a register window and leaves the rows in shared memory for the
// frequency taps.
// ============================================================================================================
template <int NT>
__global__ void __launch_bounds__(NSM_THREADS, 2) smooth_mask_kernel(const uint32_t* __restrict__ bits, const uint8_t* __restrict__ rowflag,
                                                                     const Chunk* __restrict__ chunks, float prop, SmoothTaps taps,
                                                                     float* __restrict__ M) {
    extern __shared__ __align__(16) float rows[];                 // [NS_TILE][SM_ROW]
    constexpr int W = NS_TILE + 2 * NT;
    // last chunk / tile first: the iSTFT that follows starts at chunk 0 and finds this kernel's most recent M rows in the 126 MB L2
    const int by = gridDim.y - 1 - blockIdx.y, bx = gridDim.x - 1 - blockIdx.x;
    const Chunk ch = chunks[by];
    const int F = ch.n_frames;
    const int base = NS_TILE * bx;                                // outputs [base, base + 32), raw rows [base - NT, base + 32 + NT)
    if (base >= F) return;
    const int tid = threadIdx.x;
    // the tile's bit words (W rows x 17 words) go to shared memory with ONE round of loads per CTA: the per-bin walk below
    // used to wait for a global load every eight rows (46 % of the kernel's stall samples)
    constexpr int BW = (NBIN + 31) / 32;                          // 17 words per row
    __shared__ uint32_t bw[W * BW];
    for (int i = tid; i < W * BW; i += NSM_THREADS) {
        const int r = i / BW, c = i - r * BW;
        const int rr = base - NT + r;
        bw[i] = (rr >= 0 && rr < F) ? __ldg(bits + (ch.row0 + rr) * B_PITCH + c) : 0u;
    }
    for (int i = tid; i < NS_TILE * (SM_ROW - NBIN); i += NSM_THREADS) {
        const int r = i / (SM_ROW - NBIN), c = i - r * (SM_ROW - NBIN);
        rows[r * SM_ROW + swz(c < MAX_NF ? c : c + NBIN)] = 0.f;
    }
    const bool rf = tid < NBIN ? rowflag[(long long)by * M_PITCH + tid] != 0 : false;
    // Interior tiles with few time taps: the time-smoothed value of an output depends only on the 2 NT + 1 raw BITS around it, so
    // it is looked up: lut[pattern] is computed once per CTA with the very multiply-add chain of the generic walk (bit j of the
    // pattern = frame rr - NT + j, oldest first: bit-identical results), and a bin's walk collapses to gathering its column of
    // W bits into a 64-bit word and one table read per output.
    constexpr bool USE_LUT = (2 * NT + 1 <= 7) && (W <= 64);
    __shared__ float lut[USE_LUT ? (1 << (2 * NT + 1)) : 1];
    const bool interior = base - NT >= 0 && base + NS_TILE + NT <= F;
    if (USE_LUT && interior && tid < (1 << (2 * NT + 1))) {
        const float lo = 1.0f - prop;
        float acc = taps.vt[0] * ((tid & 1) ? 1.0f : lo);
#pragma unroll
        for (int j = 1; j < 2 * NT + 1; ++j) acc = fmaf(taps.vt[j], ((tid >> j) & 1) ? 1.0f : lo, acc);
        lut[tid] = acc;
    }
    __syncthreads();
    if (USE_LUT && interior) {
        if (tid < NBIN) {
            const int k = tid;
            const uint32_t* bp = bw + (k >> 5);
            const int sh = k & 31;
            unsigned long long col = 0ull;                        // bit i = raw bit of frame base - NT + i
            if (rf) col = ~0ull;
            else {
#pragma unroll
                for (int i = 0; i < W; ++i) col |= (unsigned long long)((bp[i * BW] >> sh) & 1u) << i;
            }
#pragma unroll
            for (int lr = 0; lr < NS_TILE; ++lr)
                rows[lr * SM_ROW + swz(MAX_NF + k)] = lut[(unsigned)(col >> lr) & ((1u << (2 * NT + 1)) - 1u)];
        }
    } else if (tid < NBIN) {
        const int k = tid;
        const float lo = 1.0f - prop;
        const uint32_t* bp = bw + (k >> 5);
        const int sh = k & 31;
        float vt[2 * NT + 1], w[2 * NT + 1];
#pragma unroll
        for (int i = 0; i < 2 * NT + 1; ++i) { vt[i] = taps.vt[i]; w[i] = 0.f; }
#pragma unroll
        for (int i = 0; i < W; ++i) {
            const int rr = base - NT + i;
            const float R = (rr >= 0 && rr < F) ? ((((bp[i * BW] >> sh) & 1u) | (rf ? 1u : 0u)) ? 1.0f : lo) : 0.f;
#pragma unroll
            for (int j = 0; j < 2 * NT; ++j) w[j] = w[j + 1];
            w[2 * NT] = R;                                        // w[j] = raw mask of frame rr - 2 NT + j
            if (i >= 2 * NT) {
                const int lr = i - 2 * NT;                        // output frame rr - NT = base + lr
                float acc = vt[0] * w[0];
#pragma unroll
                for (int j = 1; j < 2 * NT + 1; ++j) acc = fmaf(vt[j], w[j], acc);
                rows[lr * SM_ROW + swz(MAX_NF + k)] = acc;
            }
        }
    }
    __syncthreads();
    const int nf = taps.nf;
    for (int wi = tid; wi < NS_TILE * 65; wi += NSM_THREADS) {
        const int lr = wi / 65, seg = wi - lr * 65;
        const int f = base + lr;
        if (f >= F) continue;
        const int k0 = 8 * seg;
        float out[8];
        freq_taps8(&rows[lr * SM_ROW], k0, nf, taps.vf, out);
        float* dst = M + (ch.row0 + f) * M_PITCH + k0;
        if (k0 + 8 <= NBIN) {
            *reinterpret_cast<float4*>(dst) = make_float4(out[0], out[1], out[2], out[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(out[4], out[5], out[6], out[7]);
        } else {
            dst[0] = out[0];                                      // bin 512
        }
    }
}

template <int NT>
static void launch_sm(dim3 grid, cudaStream_t st, const uint32_t* bits, const uint8_t* rowflag, const Chunk* chunks, float prop,
                      const SmoothTaps& taps, float* M) {
    cudaFuncSetAttribute(smooth_mask_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NSM_SMEM);
    smooth_mask_kernel<NT><<<grid, NSM_THREADS, NSM_SMEM, st>>>(bits, rowflag, chunks, prop, taps, M);
}

void launch_smooth_mask(const uint32_t* bits, const uint8_t* rowflag, const Chunk* chunks, int n_chunks,
                        int max_frames, float prop, const SmoothTaps& taps, float* M, cudaStream_t st) {
    if (n_chunks <= 0) return;
    const int tiles = (max_frames + NS_TILE - 1) / NS_TILE;
    for (int y0 = 0; y0 < n_chunks; y0 += 32768) {
        const int ny = n_chunks - y0 < 32768 ? n_chunks - y0 : 32768;
        dim3 grid(tiles, ny);
        const uint8_t* rf = rowflag + (long long)y0 * M_PITCH;
        switch (taps.nt) {
            case 1: launch_sm<1>(grid, st, bits, rf, chunks + y0, prop, taps, M); break;
            case 2: launch_sm<2>(grid, st, bits, rf, chunks + y0, prop, taps, M); break;
            case 3: launch_sm<3>(grid, st, bits, rf, chunks + y0, prop, taps, M); break;
            case 4: launch_sm<4>(grid, st, bits, rf, chunks + y0, prop, taps, M); break;
            case 5: launch_sm<5>(grid, st, bits, rf, chunks + y0, prop, taps, M); break;
            case 6: launch_sm<6>(grid, st, bits, rf, chunks + y0, prop, taps, M); break;
            case 7: launch_sm<7>(grid, st, bits, rf, chunks + y0, prop, taps, M); break;
            case 8: launch_sm<8>(grid, st, bits, rf, chunks + y0, prop, taps, M); break;
            default: launch_sm<9>(grid, st, bits, rf, chunks + y0, prop, taps, M); break;
        }
    }
}

// ============================================================================================================
// K3  masked inverse STFT with window-squared normalised overlap-add (scipy.signal.istft), int16 truncation
// (preprocessor.py:211) and the per-recording peak that the following peak-normalise needs.
// A CTA owns ISTFT_SEG output segments of 256 samples and recomputes the 3 extra frames that overlap them, so
// every output sample is summed in ascending frame order inside one CTA (deterministic, no atomics, no second
// pass).  Four frames are transformed concurrently; they are 8 frames apart so their 1024-sample spans are disjoint.
// ============================================================================================================
// cos / sin of d pi / 4, d = 0..7 (compile-time after unrolling)
__device__ __forceinline__ float w8c(int d) { const float c[8] = {1.f, 0.70710678118654752f, 0.f, -0.70710678118654752f, -1.f, -0.70710678118654752f, 0.f, 0.70710678118654752f}; return c[d]; }
__device__ __forceinline__ float w8s(int d) { const float s[8] = {0.f, 0.70710678118654752f, 1.f, 0.70710678118654752f, 0.f, -0.70710678118654752f, -1.f, -0.70710678118654752f}; return s[d]; }

constexpr int ISTFT_FR = ISTFT_SEG + 3;                 // 64 frames per CTA, 16 per group
static_assert(ISTFT_FR % 4 == 0, "four groups share the frames of a CTA");
constexpr int GSPAN = (ISTFT_FR / 4 + 3) * HOP;          // samples covered by the frames of one group (16 frames: 4864)
constexpr int OLA_LEN = 4 * GSPAN;                      // four group spans (they overlap by 768 samples in time)
constexpr size_t ISTFT_SMEM = sizeof(float) * OLA_LEN + sizeof(c32) * 4 * FFT512_SMEM;

__global__ void __launch_bounds__(256, 2) istft_kernel(const c32* __restrict__ S, const float* __restrict__ M,
                                                    const Item* __restrict__ items, const Chunk* __restrict__ chunks, Consts cs,
                                                    int16_t* __restrict__ den16, int to_out, float* __restrict__ den_f32, int* __restrict__ peak) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* ola = reinterpret_cast<float*>(smem_raw);
    c32* fsm = reinterpret_cast<c32*>(ola + OLA_LEN);

    const Chunk ch = chunks[blockIdx.y];
    const int seg_lo = PADS / HOP;                                  // first segment holding kept samples
    const int seg_hi = (PADS + ch.keep - 1) / HOP;                  // last one
    const int g0 = seg_lo + blockIdx.x * ISTFT_SEG;
    if (ch.keep <= 0 || g0 > seg_hi) return;
    const Item it = items[ch.item];
    const int tid = threadIdx.x, g = tid >> 6, t = tid & 63;
    // hann(m) / 2 = 0.25 - 0.25 cos(2 pi m / 1024) at m = 2 t + 128 d (+ 1): 0.25 + wc cos(d pi / 4) + ws sin(d pi / 4)
    float wc0, ws0, wc1, ws1;
    {
        float sn, cs;
        sincospif((float)t * (1.0f / 256.0f), &sn, &cs); wc0 = -0.25f * cs; ws0 = 0.25f * sn;
        sincospif((float)(2 * t + 1) * (1.0f / 512.0f), &sn, &cs); wc1 = -0.25f * cs; ws1 = 0.25f * sn;
    }

    // twiddle bases of this thread (the kernel is bound by shared-memory traffic: no twiddle tables, products of bases instead):
    //   FFT step 1: e^{-2 pi i t p / 512} = w1^p;  step 2: e^{-2 pi i 8 a c / 512} = w2^c, a = t & 7;
    //   real merge: e^{-2 pi i (t + 64 q) / 1024} = wt e^{-2 pi i q / 16}
    c32 w1b, w2b, wtb;
    {
        float sn, cs_;
        sincospif((float)t * (1.0f / 256.0f), &sn, &cs_); w1b = mk(cs_, -sn);
        sincospif((float)(t & 7) * (1.0f / 32.0f), &sn, &cs_); w2b = mk(cs_, -sn);
        sincospif((float)t * (1.0f / 512.0f), &sn, &cs_); wtb = mk(cs_, -sn);
    }

    const int fbase = g0 - 1;                                       // local frame 0
    // Group g walks local frames 8 g .. 8 g + 7.  The masked spectrum of the NEXT frame is fetched into registers while the
    // current one is transformed (the global-load latency used to sit exposed in front of every frame).
    c32 pxk[4], pxn[4];                                             // S[k], S[512 - k], k = t + 64 q  (q = 0, t = 0: S[0], S[256])
    float pmk[4], pmn[4], pm512 = 0.f;                              // M[k], M[512 - k]                 (q = 0, t = 0: M[0], M[256]; M[512])
    auto fetch = [&](int lf) {
        const int f = fbase + lf;
        if (f >= 0 && f < ch.n_frames) {
            const c32* row = S + (ch.row0 + f) * S_PITCH;
            const float* mrow = M + (ch.row0 + f) * M_PITCH;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = t + 64 * q;
                pxk[q] = row[k]; pmk[q] = mrow[k];
                const int kn = k == 0 ? 256 : 512 - k;
                pxn[q] = row[kn]; pmn[q] = mrow[kn];
            }
            if (t == 0) pm512 = mrow[512];
        }
        // the rows two frames further on go to L2 meanwhile: one 128-byte line per thread (32 of S, 17 of M)
        const int fp = f + 2;
        if (lf + 2 < (ISTFT_FR / 4) * (g + 1) && fp >= 0 && fp < ch.n_frames) {
            const char* ps = reinterpret_cast<const char*>(S + (ch.row0 + fp) * S_PITCH);
            const char* pm = reinterpret_cast<const char*>(M + (ch.row0 + fp) * M_PITCH);
            if (t < 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(ps + 128 * t));
            else if (t < 49) asm volatile("prefetch.global.L2 [%0];" ::"l"(pm + 128 * (t - 32)));
        }
    };
    if (t < 32) {                                                   // the group's first two frames
        const int f1 = fbase + (ISTFT_FR / 4) * g + 1;
        if (f1 >= 0 && f1 < ch.n_frames) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(S + (ch.row0 + f1) * S_PITCH) + 128 * t));
    }
    fetch((ISTFT_FR / 4) * g);
    // Overlap-add in registers: the group's frames advance by 256 samples = two of the thread's 128-sample strides, so thread t
    // accumulates the SAME addresses across consecutive frames.  win[j] holds the group-span positions 128 (2 r + j) + 2 t (+1);
    // after frame r the two oldest entries are complete and go to the group's own span buffer (plain stores, no read-modify-
    // write); the four group spans overlap by 768 samples and are added in a fixed order when the samples are emitted.
    float2 win[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) win[j] = make_float2(0.f, 0.f);
    float2* gspan = reinterpret_cast<float2*>(ola + g * GSPAN);
    for (int r = 0; r < ISTFT_FR / 4; ++r) {
        const int lf = r + (ISTFT_FR / 4) * g;
        const int f = fbase + lf;
        const bool active = f >= 0 && f < ch.n_frames;
        c32* z = fsm + g * FFT512_SMEM;
        c32 v[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) v[m] = mk(0.f, 0.f);
        if (active) {
            // thread t produces z[t + 64 q] (kept: they are its own FFT inputs v[q]) and z[512 - t - 64 q], which are the inputs
            // v[7 - q] of thread 64 - t: only those four cross the group through shared memory
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = t + 64 * q;
                if (k == 0) {
                    const float a = pxk[0].x * pmk[0], b = pxk[0].y * pm512;
                    v[0] = mk(a + b, -(a - b));                     // 2 conj(Z0)
                    z[256] = mk(2.f * pxn[0].x * pmn[0], 2.f * pxn[0].y * pmn[0]);   // 2 conj(Z256) = 2 X256
                } else {
                    c32 zk, znk;
                    // e^{-2 pi i q / 16}, q = 1..3: (cos, -sin) of pi q / 8
                    const c32 wq = q == 0 ? mk(1.f, 0.f) : q == 1 ? mk(0.92387953251128674f, -0.38268343236508978f)
                                 : q == 2 ? mk(0.70710678118654752f, -0.70710678118654752f) : mk(0.38268343236508978f, -0.92387953251128674f);
                    const c32 twk = q == 0 ? wtb : cmul(wtb, wq);
                    irfft_merge_pair(mk(pxk[q].x * pmk[q], pxk[q].y * pmk[q]), mk(pxn[q].x * pmn[q], pxn[q].y * pmn[q]), twk, zk, znk);
                    v[q] = zk; z[512 - k] = znk;
                }
            }
        }
        if (r + 1 < ISTFT_FR / 4) fetch(lf + 1);
        group_bar(g);
        if (active) {
#pragma unroll
            for (int m = 4; m < 8; ++m) v[m] = z[t + 64 * m];
        }
        group_bar(g);
        group_fft512_base(v, z, w1b, w2b, t, g);
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            // synthesis window hann / 2 at samples 2 n, 2 n + 1 (n = t + 64 d) by angle addition (angle = alpha_t + d pi / 4): two
            // FMAs each instead of a shared-memory load -- this kernel is bound by shared-memory traffic
            const float cd = w8c(d), sd = w8s(d);
            const float w0 = fmaf(wc0, cd, fmaf(ws0, sd, 0.25f)), w1 = fmaf(wc1, cd, fmaf(ws1, sd, 0.25f));
            win[d].x += v[d].x * w0;
            win[d].y += -v[d].y * w1;
        }
        gspan[64 * (2 * r) + t] = win[0];
        gspan[64 * (2 * r + 1) + t] = win[1];
#pragma unroll
        for (int j = 0; j < 6; ++j) win[j] = win[j + 2];
        win[6] = make_float2(0.f, 0.f); win[7] = make_float2(0.f, 0.f);
        group_bar(g);                                               // the scratch z is rewritten by the next frame
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) gspan[64 * (2 * (ISTFT_FR / 4) + j) + t] = win[j];
    __syncthreads();
    // emit: segment g covers padded-chunk samples [256 g, 256 g + 256) = ola[768 + 256 (g - g0) ...]; eight consecutive
    // samples per thread (16-byte shared loads, one 16-byte int16 store)
    int pk = 0;
    const long long obase = (to_out ? it.out_off : it.den_off) + ch.out_off;
    const long long fbase_o = it.out_off + ch.out_off;
    const long long o_first = (long long)g0 * HOP - PADS;           // kept-sample index of the CTA's first output (multiple of 8)
    const bool vec16 = ((obase + o_first) & 7) == 0;
    const bool vecf = den_f32 && ((fbase_o + o_first) & 3) == 0;
    const DivConst d15 = div_const(1.5f);                           // scipy istft: / sum of squared windows (hann^2 at 75 % overlap)
    constexpr int GSTRIDE = (ISTFT_FR / 4) * HOP;                   // samples between the starts of two group spans
    for (int i = 8 * tid; i < ISTFT_SEG * HOP; i += 8 * 256) {
        const long long o = o_first + i;
        if (o + 8 <= 0 || o >= ch.keep) continue;
        // CTA position x = 768 + i lies in the span of group x / 4096 and, for the first 768 samples of that span, also in the
        // previous group's; fixed order: earlier group first
        const int x = 3 * HOP + i, ga = x / GSTRIDE, xo = x - ga * GSTRIDE;
        const float4 a0 = *reinterpret_cast<const float4*>(ola + ga * GSPAN + xo), a1 = *reinterpret_cast<const float4*>(ola + ga * GSPAN + xo + 4);
        float sum[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        if (ga > 0 && xo < 3 * HOP) {
            const float* pr = ola + (ga - 1) * GSPAN + GSTRIDE + xo;
            const float4 b0 = *reinterpret_cast<const float4*>(pr), b1 = *reinterpret_cast<const float4*>(pr + 4);
            sum[0] = b0.x + sum[0]; sum[1] = b0.y + sum[1]; sum[2] = b0.z + sum[2]; sum[3] = b0.w + sum[3];
            sum[4] = b1.x + sum[4]; sum[5] = b1.y + sum[5]; sum[6] = b1.z + sum[6]; sum[7] = b1.w + sum[7];
        }
        float val[8]; int q[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            val[e] = div_by(sum[e], d15);
            q[e] = (int)(short)(int)val[e];                         // astype(int16): truncate, keep low 16 bits
        }
        if (o >= 0 && o + 8 <= ch.keep) {
            if (vec16) {
                uint4 w;
                w.x = (uint32_t)(q[0] & 0xffff) | ((uint32_t)q[1] << 16); w.y = (uint32_t)(q[2] & 0xffff) | ((uint32_t)q[3] << 16);
                w.z = (uint32_t)(q[4] & 0xffff) | ((uint32_t)q[5] << 16); w.w = (uint32_t)(q[6] & 0xffff) | ((uint32_t)q[7] << 16);
                *reinterpret_cast<uint4*>(den16 + obase + o) = w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) den16[obase + o + e] = (int16_t)q[e];
            }
            if (vecf) {
                *reinterpret_cast<float4*>(den_f32 + fbase_o + o) = make_float4(val[0], val[1], val[2], val[3]);
                *reinterpret_cast<float4*>(den_f32 + fbase_o + o + 4) = make_float4(val[4], val[5], val[6], val[7]);
            } else if (den_f32) {
#pragma unroll
                for (int e = 0; e < 8; ++e) den_f32[fbase_o + o + e] = val[e];
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) pk = max(pk, q[e] < 0 ? -q[e] : q[e]);
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                if (o + e < 0 || o + e >= ch.keep) continue;
                den16[obase + o + e] = (int16_t)q[e];
                if (den_f32) den_f32[fbase_o + o + e] = val[e];
                pk = max(pk, q[e] < 0 ? -q[e] : q[e]);
            }
        }
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
