// S8-S9 -- Whisper log-mel front-end (reference call site: audio_pipeline/transcriber.py:229-238 ->
// whisper.audio.log_mel_spectrogram(audio, n_mels, padding); restated in oracle/reference_path.py).
//
//   logmel_kernel        K5: loudness gain / limiter / int16 truncation fused into the load (the final PCM is written
//                        on the way), reflect-centred STFT 400 / hop 160 with a periodic Hann window as a 200-point
//                        complex FFT (8 x 25: radix-8 then 5 x 5 in registers) + real split, power spectrum,
//                        sparse Slaney mel projection, log10(max(., 1e-10)); per-recording running maximum.
//                        The spectrum never leaves shared memory.
//   mel_finalize_kernel  max(L, gmax - 8), (L + 4) / 4 in place once the recording-wide maximum is known.
//   mel_windows_kernel   the encoder's 30 s windows: mel[:, seek : seek + len] zero-padded to N_FRAMES, float32 or half
//                        (whisper/transcribe.py seek loop: pad_or_trim(mel_segment, N_FRAMES).to(dtype)).
#include <cuda_fp16.h>
#include "maf_common.cuh"

namespace maf {

constexpr int MEL_SPAN = (MEL_TILE - 1) * W_HOP + W_NFFT;            // 5360 samples per tile
constexpr int Z_PITCH = 200;
constexpr int P_PITCH = 201;
constexpr size_t LOGMEL_SMEM = sizeof(float) * MEL_SPAN + sizeof(c32) * MEL_TILE * Z_PITCH + sizeof(float) * (MEL_TILE * P_PITCH + 4) +
                               sizeof(float) * W_NFFT + sizeof(c32) * 200 + sizeof(c32) * 208 + sizeof(int) * 4 * MEL_SLOTS_MAX + sizeof(float4) * MEL_WQ_MAX;

struct MelTile {            // geometry of one (item, 32-frame tile) unit
    int item, lf0;
    long long i0, l0, own_lo, own_hi, den_off;
    bool last_tile, interior;
};

// Persistent CTAs (two per SM) walk the (item, tile) list: the tables are loaded once per CTA, and the raw samples of
// the NEXT tile are fetched into registers before the current tile's FFT phases, so no global-load latency is exposed.
__global__ void __launch_bounds__(256) logmel_kernel(const int16_t* __restrict__ den16, const Item* __restrict__ items, int n_items,
                                                     int max_tiles, Consts cs, MelBank mel, int padding, int peak_normalize,
                                                     const PeakNorm* __restrict__ pnorm, const int16_t* __restrict__ norm16,
                                                     const Gate* __restrict__ gates, int16_t* __restrict__ out16, float* __restrict__ mel_out,
                                                     int* __restrict__ mel_max, int* __restrict__ tile_min) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* span = reinterpret_cast<float*>(smem_raw);
    c32* zb = reinterpret_cast<c32*>(span + MEL_SPAN);
    float* pw = reinterpret_cast<float*>(zb + MEL_TILE * Z_PITCH);
    float* win = pw + MEL_TILE * P_PITCH + 4;                            // four floats of slack: padded mel runs may read past bin 200
    c32* twA = reinterpret_cast<c32*>(win + W_NFFT);
    c32* tw400 = twA + 200;
    float4* m_wq = reinterpret_cast<float4*>(tw400 + 208);               // 16-byte aligned: every array before it is a multiple of 16 bytes
    int* m_qstart = reinterpret_cast<int*>(m_wq + MEL_WQ_MAX);
    int* m_qcount = m_qstart + MEL_SLOTS_MAX;
    int* m_qoff = m_qcount + MEL_SLOTS_MAX;
    int* m_qrow = m_qoff + MEL_SLOTS_MAX;
    const int tid = threadIdx.x;
    // descriptor / peak factor / gate of the item the CTA is working on, fetched once per ITEM instead of once per tile (a long
    // recording is thousands of tiles of the same item: the global-load latency of the descriptor sat at the top of every tile)
    __shared__ Item s_item;
    __shared__ PeakNorm s_pn;
    __shared__ Gate s_gt;
    int cached = -1;

    const unsigned total = (unsigned)n_items * (unsigned)max_tiles;            // < 2^31 (checked by the launcher): 32-bit index math
    auto geom = [&](unsigned t, MelTile& g) {
        g.item = (int)(t / (unsigned)max_tiles);
        g.lf0 = (int)(t - (unsigned)g.item * (unsigned)max_tiles) * MEL_TILE;
        const Item& it = g.item == cached ? s_item : items[g.item];
        g.den_off = it.den_off;
        if (g.lf0 >= it.mel_frames) return false;
        g.i0 = ((long long)it.mel_f0 + g.lf0) * W_HOP - W_NFFT / 2;
        g.last_tile = g.lf0 + MEL_TILE >= it.mel_frames;
        g.own_lo = g.i0 + W_NFFT / 2;
        g.own_hi = g.last_tile ? it.g0 + it.own_len : g.own_lo + (long long)MEL_TILE * W_HOP;
        g.l0 = g.i0 - it.g0;                                        // item-relative index of span[0]
        // (norm16, the peak-normalised copy kweight_energy_kernel leaves behind, covers the own range only: no halos)
        // (peak normalisation requested but no norm16 -- a stage-wise call without the loudness pass -- takes the generic path)
        g.interior = (norm16 || !peak_normalize) && g.i0 >= 0 && g.i0 + MEL_SPAN <= it.n_global && g.l0 >= (norm16 ? 0LL : -(long long)it.lo_avail) &&
                     g.l0 + MEL_SPAN <= it.own_len + (norm16 ? 0 : it.hi_avail) && ((it.den_off + g.l0) & 7) == 0 && !g.last_tile &&
                     (!out16 || (((it.out_off + g.l0) & 7) == 0 && g.l0 + W_NFFT / 2 >= 0 && g.l0 + W_NFFT / 2 + MEL_TILE * W_HOP <= it.own_len));
        return true;
    };
    constexpr int PRE = (MEL_SPAN / 8 + 255) / 256;                 // 16-byte groups per thread
    uint4 pre[PRE];
    auto fetch = [&](const MelTile& g) {                            // interior fast path only: raw samples -> registers
        if (!g.interior) return;
        const int16_t* x = (norm16 ? norm16 : den16) + g.den_off + g.l0;
#pragma unroll
        for (int u = 0; u < PRE; ++u) {
            const int gq = tid + 256 * u;
            if (gq < MEL_SPAN / 8) pre[u] = *reinterpret_cast<const uint4*>(x + 8 * gq);
        }
    };

    unsigned t = blockIdx.x;
    MelTile cur, nxt;
    while (t < total && !geom(t, cur)) t += gridDim.x;
    if (t >= total) return;
    fetch(cur);

    for (int i = tid; i < W_NFFT; i += 256) win[i] = cs.win400[i];
    for (int i = tid; i < 200; i += 256) twA[i] = cs.twA[i];
    if (tid < mel.n_slots) { m_qstart[tid] = mel.qstart[tid]; m_qcount[tid] = mel.qcount[tid]; m_qoff[tid] = mel.qoff[tid]; m_qrow[tid] = mel.qrow[tid]; }
    for (int i = tid; i < mel.n_wq; i += 256) m_wq[i] = mel.wq[i];
    for (int i = tid; i < 201; i += 256) tw400[i] = cs.tw400[i];
    if (tid < 4) pw[MEL_TILE * P_PITCH + tid] = 0.f;              // slack read (times a zero weight) by the last frame's padded mel runs

  while (t < total) {
    const int item = cur.item;
    if (item != cached) {                                           // CTA-uniform
        __syncthreads();                                            // nobody still reads the previous item's copy
        if (tid < (int)(sizeof(Item) / 4)) reinterpret_cast<int*>(&s_item)[tid] = reinterpret_cast<const int*>(items + item)[tid];
        if (tid == 64) { PeakNorm v; v.factor = 1.0; v.fh = 1.0f; v.fl = 0.f; v.active = 0; v.pad_ = 0; if (peak_normalize) v = pnorm[item]; s_pn = v; }
        if (tid == 96) { Gate v; v.lufs = 0.0; v.gain = 1.0f; v.limp = 1.0f; v.applied = 0; v.lim = 0; if (gates) v = gates[item]; s_gt = v; }
        cached = item;
        __syncthreads();
    }
    const Item& it = s_item;
    const int lf0 = cur.lf0;
    const PeakNorm pnv = s_pn;
    const bool pn = pnv.active != 0;
    const Gate gt = s_gt;
    const DivConst dv = limiter_div(gt);
    const float gain = gt.applied ? gt.gain : 1.0f;
    const int16_t* x = den16 + it.den_off;
    const long long np_global = it.n_global + padding;            // length of the zero-padded recording
    const long long i0 = cur.i0, l0 = cur.l0, own_lo = cur.own_lo, own_hi = cur.own_hi;
    __syncthreads();                                              // tables ready / previous tile's shared memory free

    // Stage the tile's samples: global index i = 160 f - 200 + e, reflected at the true ends of the zero-padded
    // recording.  The tile also writes the final PCM of the 32 hops it is centred on (the last tile up to the item's
    // end: with padding 0 the tail N mod 160 plus the last hop still lies inside its span).
    {
        if (cur.interior) {
            // fast path: no reflection, no padding, 16-byte loads and stores (i0, own_lo and the hop are multiples of 8)
#pragma unroll
            for (int u = 0; u < PRE; ++u) {
                const int gq = tid + 256 * u;
                if (gq >= MEL_SPAN / 8) break;
                const uint4 v = pre[u];
                const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
                float a[8];
                uint32_t ow[4];
                // the samples are final up to the loudness gain (peak normalisation done by kweight_energy_kernel, or off):
                // int16 -> float by byte permutation (0x4B00'0000 | (s ^ 0x8000) = 2^23 + 32768 + s), gain, limiter, clamp,
                // truncation; the low 16 bits of (q + 1.5 * 2^23) are q's two's complement, so the int16 pair is one permute
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const uint32_t b = wv[h] ^ 0x80008000u;
                    float t0 = __fadd_rn(__uint_as_float(__byte_perm(b, 0x4B000000u, 0x7410)), -8421376.0f);
                    float t1 = __fadd_rn(__uint_as_float(__byte_perm(b, 0x4B000000u, 0x7432)), -8421376.0f);
                    // (a pass-through gate runs the same chain with gain 1: 1 * t, the clamp and the truncation leave an int16-valued
                    // float unchanged, so no second code path -- the loop body has to stay small for the instruction cache)
                    t0 = __fmul_rn(gain, t0); t1 = __fmul_rn(gain, t1);
                    if (gt.lim) { t0 = div_by(t0, dv); t1 = div_by(t1, dv); }
                    t0 = truncf(fminf(fmaxf(t0, -32768.0f), 32767.0f));
                    t1 = truncf(fminf(fmaxf(t1, -32768.0f), 32767.0f));
                    ow[h] = __byte_perm(__float_as_uint(__fadd_rn(t0, MAGIC_F)), __float_as_uint(__fadd_rn(t1, MAGIC_F)), 0x5410);
                    a[2 * h] = t0 * (1.0f / 32768.0f); a[2 * h + 1] = t1 * (1.0f / 32768.0f);
                }
                *reinterpret_cast<float4*>(span + 8 * gq) = make_float4(a[0], a[1], a[2], a[3]);
                *reinterpret_cast<float4*>(span + 8 * gq + 4) = make_float4(a[4], a[5], a[6], a[7]);
                const int e0 = 8 * gq;
                if (out16 && e0 >= W_NFFT / 2 && e0 < W_NFFT / 2 + MEL_TILE * W_HOP) {
                    const uint4 o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                    *reinterpret_cast<uint4*>(out16 + it.out_off + l0 + e0) = o;
                }
            }
        } else {
            for (int e = tid; e < MEL_SPAN; e += 256) {
                const long long i = i0 + e;
                long long r = i;
                if (r < 0) r = -r;
                if (r >= np_global) r = 2 * (np_global - 1) - r;
                float a = 0.f;
                if (r >= 0 && r < it.n_global) {
                    const long long l = r - it.g0;                    // item-relative index into the denoised samples
                    if (l >= -(long long)it.lo_avail && l < it.own_len + it.hi_avail) {
                        int sv = x[l];
                        float qf;
                        const int q = norm_gain_sample(sv, pn, pnv, gt, dv, qf);
                        a = qf * (1.0f / 32768.0f);
                        if (out16 && r == i && i >= own_lo && i < own_hi && l >= 0 && l < it.own_len) out16[it.out_off + l] = (int16_t)q;
                    }
                }
                span[e] = a;
            }
        }
    }
    // next unit of this CTA: its raw samples travel while this tile is transformed
    unsigned tn = t + gridDim.x;
    while (tn < total && !geom(tn, nxt)) tn += gridDim.x;
    if (tn < total) fetch(nxt);
    __syncthreads();

    // stage A: (frame, j) -> 8-point FFT over m of z[j + 25 m], twiddle, store.  Thread = (j, frame group): j is FIXED per
    // thread, so its 16 window values and 7 twiddles are read once per tile instead of once per transform, and the four
    // transforms of a thread (frames fg, fg + 8, fg + 16, fg + 24) need only their 8 sample pairs from shared memory.
    if (tid < 200) {
        const int j = tid % 25, fg = tid / 25;
        float2 wn[8];
        c32 twj[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) wn[m] = *reinterpret_cast<const float2*>(win + 2 * (j + 25 * m));
#pragma unroll
        for (int p = 1; p < 8; ++p) twj[p] = twA[p * 25 + j];
#pragma unroll 1
        for (int r = 0; r < MEL_TILE / 8; ++r) {
            const int lf = fg + 8 * r;
            const float* sp0 = span + lf * W_HOP + 2 * j;
            c32 v[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const float2 sp = *reinterpret_cast<const float2*>(sp0 + 50 * m);        // 8-byte aligned: every term is even
                v[m] = mk(sp.x * wn[m].x, sp.y * wn[m].y);
            }
            fft8(v);
            c32* zo = zb + lf * Z_PITCH + j;
            zo[0] = v[0];
#pragma unroll
            for (int p = 1; p < 8; ++p) zo[p * 25] = cmul(v[p], twj[p]);
        }
    }
    __syncthreads();
    // stage B: (frame, p) -> 25-point DFT in registers; exactly one round, so the in-place write is safe after a barrier
    {
        const int lf = tid >> 3, p = tid & 7;
        c32 o[25];
        fft200_stepB(zb + lf * Z_PITCH, p, o);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 25; ++q) zb[lf * Z_PITCH + p + 8 * q] = o[q];
    }
    __syncthreads();
    // real split + power: one work item = (frame, k <= 100) produces bins k and 200 - k from the same two loads.  Thread =
    // (k, frame parity): k, its twiddle and every address offset are fixed per thread; the 16 frames of a thread are a
    // constant stride apart (no division, no per-item twiddle load).
    if (tid < 202) {
        const int k = tid % 101, fg = tid / 101;
        const c32 tw = tw400[k];
        const c32 twn = mk(-tw.x, tw.y);                                // e^{-2 pi i (200 - k) / 400} = -conj(e^{-2 pi i k / 400})
        const int kn = k == 0 ? 0 : 200 - k;
        const c32* z = zb + fg * Z_PITCH;
        float* po = pw + fg * P_PITCH;
#pragma unroll 2
        for (int r = 0; r < MEL_TILE / 2; ++r) {
            const c32 za = z[k], zn = z[kn];
            const c32 Xa = rfft400_bin(za, zn, tw);
            const c32 Xb = rfft400_bin(zn, za, twn);
            po[k] = Xa.x * Xa.x + Xa.y * Xa.y;
            po[200 - k] = Xb.x * Xb.x + Xb.y * Xb.y;
            z += 2 * Z_PITCH; po += 2 * P_PITCH;
        }
    }
    __syncthreads();
    // mel projection: warp-uniform quad of four adjacent mel rows, lane = frame (coalesced 128 B row segments of the
    // [n_mels, frames] output); each power bin of the quad's run is read once and feeds four accumulators
    {
        const int lane = tid & 31, warp = tid >> 5;
        const int lf = lane;
        const bool valid = lf0 + lf < it.mel_frames;
        float bmax = -INFINITY, bmin = INFINITY;
        float* orow = mel_out + it.mel_off + lf0 + lf;
        for (int sl = warp; sl < mel.n_slots; sl += 8) {
            const int q = m_qrow[sl];                                   // quads were dealt to the warps by cost on the host
            if (q < 0) continue;
            const int cnt = m_qcount[sl];
            const float4* wts = m_wq + m_qoff[sl];
            const float* pr = pw + lf * P_PITCH + m_qstart[sl];
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            // the host pads every run to a multiple of four bins with zero weights (the power rows have four floats of slack):
            // whole groups of four, all eight loads of a group issued before its sixteen FMAs, no remainder loop
#pragma unroll 1
            for (int i = 0; i < cnt; i += 4) {
                const float p0 = pr[i], p1 = pr[i + 1], p2 = pr[i + 2], p3 = pr[i + 3];
                const float4 w0 = wts[i], w1 = wts[i + 1], w2 = wts[i + 2], w3 = wts[i + 3];
                a0 = fmaf(w0.x, p0, a0); a1 = fmaf(w0.y, p0, a1); a2 = fmaf(w0.z, p0, a2); a3 = fmaf(w0.w, p0, a3);
                a0 = fmaf(w1.x, p1, a0); a1 = fmaf(w1.y, p1, a1); a2 = fmaf(w1.z, p1, a2); a3 = fmaf(w1.w, p1, a3);
                a0 = fmaf(w2.x, p2, a0); a1 = fmaf(w2.y, p2, a1); a2 = fmaf(w2.z, p2, a2); a3 = fmaf(w2.w, p2, a3);
                a0 = fmaf(w3.x, p3, a0); a1 = fmaf(w3.y, p3, a1); a2 = fmaf(w3.z, p3, a2); a3 = fmaf(w3.w, p3, a3);
            }
            // lg2.approx (|err| < 3e-6 in log10 units; the clamp keeps the argument normal, so the .ftz form needs no denormal
            // fix-up) and (log10 + 4) / 4 as ONE fused multiply-add: 1 + lg2 * (log10(2) / 4)
            auto sc = [](float a) {
                float l;
                asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fmaxf(a, 1e-10f)));
                return fmaf(l, 0.30102999566398120f * 0.25f, 1.0f);
            };
            const float l0 = sc(a0), l1 = sc(a1), l2 = sc(a2), l3 = sc(a3);
            if (valid) {
                float* o = orow + (long long)(4 * q) * it.mel_frames;
                // written already scaled; the max - 8 clamp is applied afterwards only to tiles that need it
                o[0] = l0; o[it.mel_frames] = l1; o[2LL * it.mel_frames] = l2; o[3LL * it.mel_frames] = l3;
                bmax = fmaxf(fmaxf(bmax, fmaxf(l0, l1)), fmaxf(l2, l3));
                bmin = fminf(fminf(bmin, fminf(l0, l1)), fminf(l2, l3));
            }
        }
        // back to log10 units for the recording-wide maximum and the tile minimum (scaling is monotonic)
        bmax = fmaf(bmax, 4.0f, -4.0f); bmin = fmaf(bmin, 4.0f, -4.0f);
        bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, 16));
        bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, 8));
        bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, 4));
        bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, 2));
        bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, 1));
        bmin = fminf(bmin, __shfl_xor_sync(0xffffffffu, bmin, 16));
        bmin = fminf(bmin, __shfl_xor_sync(0xffffffffu, bmin, 8));
        bmin = fminf(bmin, __shfl_xor_sync(0xffffffffu, bmin, 4));
        bmin = fminf(bmin, __shfl_xor_sync(0xffffffffu, bmin, 2));
        bmin = fminf(bmin, __shfl_xor_sync(0xffffffffu, bmin, 1));
        if (lane == 0 && bmax > -INFINITY) {
            atomicMax(&mel_max[item], float_to_ordered(bmax));
            atomicMin(&tile_min[(long long)item * max_tiles + lf0 / MEL_TILE], float_to_ordered(bmin));
        }
    }
    t = tn; cur = nxt;
  }
}

int launch_logmel(const int16_t* den16, const Item* items, int n_items, int max_tiles, const Consts& c,
                   const MelBank& mel, int padding, int peak_normalize, const PeakNorm* pnorm, const int16_t* norm16,
                   const Gate* gates, int16_t* out16, float* mel_out, int* mel_max, int* tile_min, cudaStream_t st) {
    if (n_items <= 0) return 0;
    cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LOGMEL_SMEM);
    if (max_tiles < 1) max_tiles = 1;
    static int n_sm = 0;
    if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
    const long long total = (long long)n_items * max_tiles;
    if (total >= (1LL << 31) - 4096) return -1;                    // maf_plan_create rejects such plans; never skip work silently
    const int grid = (int)(total < 2LL * n_sm ? total : 2LL * n_sm);
    logmel_kernel<<<grid, 256, LOGMEL_SMEM, st>>>(den16, items, n_items, max_tiles, c, mel, padding, peak_normalize, pnorm, norm16, gates, out16,
                                                  mel_out, mel_max, tile_min);
    return 0;
}

// whisper.audio.log_mel_spectrogram: log_spec = maximum(log_spec, log_spec.max() - 8.0); (log_spec + 4.0) / 4.0.
// logmel_kernel has already stored (L + 4) / 4 and the minimum L of every 32-frame tile; division by 4 and adding 4 are
// monotonic, so clamping the stored value at ((max - 8) + 4) / 4 is the same float32 arithmetic as the reference, and only
// tiles whose minimum lies below max - 8 (silence, the zero padding) have to be touched again.
__global__ void __launch_bounds__(256) mel_finalize_kernel(float* __restrict__ mel_out, const Item* __restrict__ items, int n_mels,
                                                           const int* __restrict__ mel_max, const int* __restrict__ tile_min) {
    const int item = blockIdx.y;
    const Item it = items[item];
    const int lf0 = blockIdx.x * MEL_TILE;
    if (lf0 >= it.mel_frames) return;
    const float floor_v = ordered_to_float(mel_max[item]) - 8.0f;
    if (ordered_to_float(tile_min[(long long)item * gridDim.x + blockIdx.x]) >= floor_v) return;
    const float floor_s = __fmul_rn(__fadd_rn(floor_v, 4.0f), 0.25f);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lf0 + lane >= it.mel_frames) return;
    float* p = mel_out + it.mel_off + lf0 + lane;
    for (int m = warp; m < n_mels; m += 8) {
        float* q = p + (long long)m * it.mel_frames;
        *q = fmaxf(*q, floor_s);
    }
}

void launch_mel_finalize(float* mel_out, const Item* items, int n_items, int n_mels, int max_tiles,
                         const int* mel_max, const int* tile_min, cudaStream_t st) {
    if (n_items <= 0 || max_tiles <= 0) return;
    for (int y0 = 0; y0 < n_items; y0 += 32768) {
        const int ny = n_items - y0 < 32768 ? n_items - y0 : 32768;
        dim3 grid(max_tiles, ny);
        mel_finalize_kernel<<<grid, 256, 0, st>>>(mel_out, items + y0, n_mels, mel_max + y0, tile_min + (long long)y0 * max_tiles);
    }
}

// out[w, m, f] = (f < len_w && seek_w + f < frames) ? mel[m, seek_w + f] : 0; one CTA per (mel row, window), two frames per
// thread and step, so a warp reads 256 contiguous bytes and writes 128 (half) or 256 (float) contiguous bytes.
template <typename OUT>
__global__ void __launch_bounds__(256) mel_windows_kernel(const float* __restrict__ mel, long long frames, const int* __restrict__ seeks,
                                                          const int* __restrict__ lens, int window_frames, OUT* __restrict__ out) {
    const int m = blockIdx.x, w = blockIdx.y;
    const long long seek = seeks[w];
    long long len = lens ? (long long)lens[w] : (long long)window_frames;
    len = min(len, frames - seek);                                          // <= 0: an all-zero window
    const float* row = mel + (long long)m * frames + seek;
    OUT* orow = out + ((long long)w * gridDim.x + m) * window_frames;
    for (int f = 2 * threadIdx.x; f < window_frames; f += 512) {
        const float a = (f < len && seek + f >= 0) ? row[f] : 0.f;
        const float b = (f + 1 < len && seek + f + 1 >= 0) ? row[f + 1] : 0.f;
        if constexpr (sizeof(OUT) == 2) {
            if (f + 1 < window_frames && (window_frames & 1) == 0) *reinterpret_cast<__half2*>(orow + f) = __floats2half2_rn(a, b);
            else { orow[f] = __float2half_rn(a); if (f + 1 < window_frames) orow[f + 1] = __float2half_rn(b); }
        } else {
            if (f + 1 < window_frames && (window_frames & 1) == 0) *reinterpret_cast<float2*>(orow + f) = make_float2(a, b);
            else { orow[f] = a; if (f + 1 < window_frames) orow[f + 1] = b; }
        }
    }
}

void launch_mel_windows(const float* mel, int n_mels, long long frames, const int* seeks, const int* lens, int n_windows,
                        int window_frames, int out_fp16, void* out, cudaStream_t st) {
    if (n_windows <= 0 || n_mels <= 0 || window_frames <= 0) return;
    for (int w0 = 0; w0 < n_windows; w0 += 32768) {
        const int nw = n_windows - w0 < 32768 ? n_windows - w0 : 32768;
        dim3 grid(n_mels, nw);
        const long long o = (long long)w0 * n_mels * window_frames;
        if (out_fp16) mel_windows_kernel<__half><<<grid, 256, 0, st>>>(mel, frames, seeks + w0, lens ? lens + w0 : nullptr, window_frames, (__half*)out + o);
        else mel_windows_kernel<float><<<grid, 256, 0, st>>>(mel, frames, seeks + w0, lens ? lens + w0 : nullptr, window_frames, (float*)out + o);
    }
}

}  // namespace maf
