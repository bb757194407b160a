// S0 -- noise-clip auto-detection on the device (reference: audio_pipeline/preprocessor.py:92-147, :179-197).
//
// Kernel frame_stats : one thread per 25 ms / 10 ms frame.  Energy is sqrt(mean(frame**2)) evaluated in float32
//                      with numpy's pairwise summation order (8 strided accumulators per <=128-element leaf,
//                      halving splits rounded down to multiples of 8), so it is bit-identical to the reference's
//                      np.mean on float32 and the percentile RANKS cannot differ.  ZCR is kept as an integer count.
// Kernel select_clip : one CTA per recording.  20th percentile of energy by 4-pass radix select on the float bits,
//                      median ZCR from a histogram, numpy 'linear' interpolation in float64, the
//                      (E < eT) & (ZCR > 0.5 zT) mask, and the longest run of >= 100 ms that is closed by a
//                      non-noise frame (an open run at the end is dropped, like the reference's loop).
#include <cooperative_groups.h>
#include "maf_common.cuh"
#include "run_merge.h"

namespace maf {

// ---- numpy pairwise_sum (loops_utils.h.src) over x[i]^2, float32, no FMA contraction ------------------------
__device__ float sq_at(const int16_t* p, int i) {
    float v = (float)p[i];
    return __fmul_rn(v, v);
}

__device__ float pairwise_leaf(const int16_t* p, int n) {
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, sq_at(p, i));
        return res;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = sq_at(p, j);
    int i = 8;
    const int lim = n - (n % 8);
    for (; i < lim; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], sq_at(p, i + j));
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, sq_at(p, i));
    return res;
}

// iterative form of the recursion: explicit stack of (offset, n) with post-order combination
__device__ float pairwise_sumsq(const int16_t* p, int n) {
    if (n <= 128) return pairwise_leaf(p, n);
    // depth <= 12 for any realistic frame length
    int off_stack[16], n_stack[16], state[16];
    float val[16];
    int sp = 0;
    off_stack[0] = 0; n_stack[0] = n; state[0] = 0;
    float ret = 0.f;
    while (sp >= 0) {
        int o = off_stack[sp], m = n_stack[sp];
        if (m <= 128) { ret = pairwise_leaf(p + o, m); --sp; continue; }
        int n2 = m / 2; n2 -= n2 % 8;
        if (state[sp] == 0) {            // descend left
            state[sp] = 1;
            ++sp; off_stack[sp] = o; n_stack[sp] = n2; state[sp] = 0;
        } else if (state[sp] == 1) {     // left done -> descend right
            val[sp] = ret; state[sp] = 2;
            ++sp; off_stack[sp] = o + n2; n_stack[sp] = m - n2; state[sp] = 0;
        } else {                         // both done
            ret = __fadd_rn(val[sp], ret); --sp;
        }
    }
    return ret;
}

__global__ void __launch_bounds__(128) frame_stats_kernel(const int16_t* __restrict__ pcm, const Item* __restrict__ items,
                                                          int frame_len, int hop, float* __restrict__ energies,
                                                          int* __restrict__ zc) {
    const Item it = items[blockIdx.y];
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= it.n_sframes) return;
    const int16_t* p = pcm + it.pcm_off + it.origin + (long long)f * hop;
    float s = pairwise_sumsq(p, frame_len);
    float mean = __fdiv_rn(s, (float)frame_len);
    energies[it.sframe_off + f] = __fsqrt_rn(mean);
    int c = 0;
    int prev = p[0] < 0;
    for (int i = 1; i < frame_len; ++i) { int cur = p[i] < 0; c += (cur != prev); prev = cur; }
    zc[it.sframe_off + f] = c;
}

// Fast path for the 16 kHz geometry (frame 400, hop 160): one WARP per frame, the tile's samples staged in shared memory
// ONCE as float32 squares plus one sign bit each (every sample belongs to 2.5 frames).
// numpy's pairwise sum of 400 float32 squares is ((leaf(0,96) + leaf(96,104)) + (leaf(200,96) + leaf(296,104))) where a leaf
// keeps 8 strided accumulators r[j] += x[8 i + j]^2 and folds them as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)): that is 32
// independent sequential chains -- one per lane -- followed by exactly the xor-1, 2, 4 (inside a leaf) and xor-8, 16
// (across leaves) butterfly, so the warp-parallel sum is bit-identical to the serial one (float addition commutes).
// The sign-change count of a frame is a population count over the tile's change bits c[i] = sign[i] ^ sign[i-1]:
// frames start on multiples of 160 = 5 words, so a frame is 12 whole words and half of a 13th.
constexpr int FS_FRAMES = 64;                                    // frames per CTA
constexpr int FS_SPAN = (FS_FRAMES - 1) * 160 + 400;             // 10480 samples
constexpr int FS_WORDS = (FS_SPAN + 31) / 32;                    // 328 words of sign bits
constexpr size_t FS_SMEM = sizeof(float) * (FS_SPAN + 8) + sizeof(unsigned) * 2 * (FS_WORDS + 4);

__global__ void __launch_bounds__(256) frame_stats400_kernel(const int16_t* __restrict__ pcm, const Item* __restrict__ items,
                                                            float* __restrict__ energies, int* __restrict__ zc) {
    extern __shared__ __align__(16) unsigned char fs_smem[];
    float* sq = reinterpret_cast<float*>(fs_smem);                               // [FS_SPAN + 8] squares
    unsigned* bw = reinterpret_cast<unsigned*>(sq + FS_SPAN + 8);                // [FS_WORDS + 4] sign bits
    unsigned* cw = bw + FS_WORDS + 4;                                            // [FS_WORDS + 4] change bits
    const Item it = items[blockIdx.y];
    const int f0 = blockIdx.x * FS_FRAMES;
    if (f0 >= it.n_sframes) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nfr = min(FS_FRAMES, it.n_sframes - f0);
    const int need = (nfr - 1) * 160 + 400;
    const long long g0 = it.pcm_off + it.origin + (long long)f0 * 160;          // first sample of the tile in the pcm array
    const int16_t* p = pcm + g0;
    const bool aligned = (g0 & 7) == 0;
    const int n_groups = (need + 7) / 8;                                         // 8 samples per thread and step; 4 threads = one sign word
    // all 16-byte loads of the thread are issued before the first one is used (one exposed load latency per CTA instead of six)
    constexpr int FS_ROUNDS = ((FS_SPAN + 7) / 8 + 31) / 32 * 32 / 256 + 1;
    uint4 pre[FS_ROUNDS];
#pragma unroll
    for (int u = 0; u < FS_ROUNDS; ++u) {
        const int g = tid + 256 * u;
        if (g < n_groups && aligned && 8 * g + 8 <= need) pre[u] = *reinterpret_cast<const uint4*>(p + 8 * g);
    }
#pragma unroll
    for (int u = 0; u < FS_ROUNDS; ++u) {                                        // whole warps: the sign words are assembled by shuffles
        const int g = tid + 256 * u;
        if (g >= ((n_groups + 31) & ~31)) break;
        int v[8];
        if (g < n_groups && aligned && 8 * g + 8 <= need) {
            const uint4 q = pre[u];
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int h = 0; h < 4; ++h) { v[2 * h] = (int)(short)(w[h] & 0xffffu); v[2 * h + 1] = (int)w[h] >> 16; }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (g < n_groups && 8 * g + e < need) ? (int)p[8 * g + e] : 0;
        }
        unsigned sb = 0;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float x = int_to_float_exact(v[e]);
            f[e] = __fmul_rn(x, x);
            sb |= (v[e] < 0 ? 1u : 0u) << e;
        }
        sb <<= 8 * (lane & 3);
        sb |= __shfl_xor_sync(0xffffffffu, sb, 1);
        sb |= __shfl_xor_sync(0xffffffffu, sb, 2);
        if (g < n_groups) {
            *reinterpret_cast<float4*>(sq + 8 * g) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(sq + 8 * g + 4) = make_float4(f[4], f[5], f[6], f[7]);
            if ((lane & 3) == 0) bw[g >> 2] = sb;
        }
    }
    __syncthreads();
    const int n_words = (need + 31) / 32;
    for (int w = tid; w < n_words; w += 256) {
        const unsigned b = bw[w];
        cw[w] = b ^ ((b << 1) | (w ? bw[w - 1] >> 31 : 0u));                     // bit j = change between samples 32 w + j - 1 and 32 w + j
    }
    __syncthreads();
    // lane -> (leaf, accumulator): leaf offsets 0 / 96 / 200 / 296, lengths 96 / 104 / 96 / 104
    const int leaf = lane >> 3, j = lane & 7;
    const int loff = leaf == 0 ? 0 : leaf == 1 ? 96 : leaf == 2 ? 200 : 296;
    const int lcnt = (leaf & 1) ? 13 : 12;
    for (int lf = warp; lf < nfr; lf += 8) {
        const float* x = sq + lf * 160 + loff + j;
        float r = x[0];
#pragma unroll
        for (int i = 1; i < 13; ++i) {
            if (i < lcnt) r = __fadd_rn(r, x[8 * i]);
        }
        r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
        r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
        r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
        r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 8));
        r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 16));
        // sign changes of the pairs (i - 1, i), i = 1 .. 399 of the frame: bits 1..31 of word 0, words 1..11, bits 0..15 of word 12
        int c = 0;
        if (lane < 13) {
            const unsigned w = cw[5 * lf + lane];
            c = __popc(lane == 0 ? (w & ~1u) : lane == 12 ? (w & 0xffffu) : w);
        }
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) {
            energies[it.sframe_off + f0 + lf] = __fsqrt_rn(__fdiv_rn(r, 400.0f));
            zc[it.sframe_off + f0 + lf] = c;
        }
    }
}

void launch_frame_stats(const int16_t* pcm, const Item* items, int n_items, int max_frames, int frame_len, int hop,
                        float* energies, int* zc, cudaStream_t st) {
    if (max_frames <= 0 || n_items <= 0) return;
    if (frame_len == 400 && hop == 160) {
        dim3 grid((max_frames + FS_FRAMES - 1) / FS_FRAMES, n_items);
        cudaFuncSetAttribute(frame_stats400_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FS_SMEM);
        frame_stats400_kernel<<<grid, 256, FS_SMEM, st>>>(pcm, items, energies, zc);
        return;
    }
    dim3 grid((max_frames + 127) / 128, n_items);
    frame_stats_kernel<<<grid, 128, 0, st>>>(pcm, items, frame_len, hop, energies, zc);
}

// ---- selection ------------------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 1024;

// numpy _lerp(a, b, t) in float64 without FMA contraction
__device__ double np_lerp(double a, double b, double t) {
    double d = __dsub_rn(b, a);
    if (t >= 0.5) return __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, t)));
    return __dadd_rn(a, __dmul_rn(d, t));
}

// Recordings of up to SEL_CACHE * SEL_THREADS frames (82 s) keep their statistics in registers -- element tid + 1024 j in
// slot j -- so the five passes over the energies and the two over the counts never go back to global memory.
constexpr int SEL_CACHE = 8;

template <typename FN>
__device__ __forceinline__ void for_each_energy(bool cached, const float (&eR)[SEL_CACHE], const float* __restrict__ v, int n, FN fn) {
    if (cached) {
#pragma unroll
        for (int j = 0; j < SEL_CACHE; ++j) { if ((int)threadIdx.x + j * SEL_THREADS < n) fn(__float_as_uint(eR[j])); }
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) fn(__float_as_uint(v[i]));
    }
}

// k-th smallest (0-based) of non-negative floats by radix select on their bit patterns; also returns via vk1 the
// (k+1)-th smallest.  All threads of the CTA call this; hist is 256 ints of shared memory, scratch 4 ints.
__device__ void radix_select2(bool cached, const float (&eR)[SEL_CACHE], const float* __restrict__ v, int n, int k, unsigned* hist,
                              unsigned* scratch, float& vk, float& vk1) {
    unsigned prefix = 0, mask = 0;
    int kk = k;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        {   // neighbouring frames mostly share a digit: aggregate runs locally before touching shared memory
            int cur = -1; unsigned run = 0;
            for_each_energy(cached, eR, v, n, [&](unsigned u) {
                if ((u & mask) != prefix) return;
                const int d = (u >> shift) & 255;
                if (d == cur) { ++run; return; }
                if (run) atomicAdd(&hist[cur], run);
                cur = d; run = 1;
            });
            if (run) atomicAdd(&hist[cur], run);
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // first digit whose cumulative count exceeds kk: every lane sums 8 bins, warp scan, the crossing lane finishes
            const int lane = threadIdx.x;
            unsigned h[8], tot = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { h[j] = hist[8 * lane + j]; tot += h[j]; }
            unsigned incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned v2 = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v2; }
            const unsigned excl = incl - tot;
            const unsigned hit = __ballot_sync(0xffffffffu, incl > (unsigned)kk);
            if (hit == 0) { if (lane == 31) { scratch[0] = 256; scratch[1] = incl; } }
            else if (lane == __ffs(hit) - 1) {
                unsigned acc = excl; int j = 0;
                for (; j < 8; ++j) { if (acc + h[j] > (unsigned)kk) break; acc += h[j]; }
                scratch[0] = 8 * lane + j; scratch[1] = acc;
            }
        }
        __syncthreads();
        prefix |= scratch[0] << shift; mask |= 255u << shift; kk -= (int)scratch[1];
        __syncthreads();
    }
    vk = __uint_as_float(prefix);
    // count of elements <= vk and the smallest element > vk
    if (threadIdx.x == 0) { scratch[2] = 0; scratch[3] = 0x7f800000u; }
    __syncthreads();
    unsigned cnt = 0, mn = 0x7f800000u;
    for_each_energy(cached, eR, v, n, [&](unsigned u) { if (u <= prefix) ++cnt; else mn = min(mn, u); });
    cnt = __reduce_add_sync(0xffffffffu, cnt); mn = __reduce_min_sync(0xffffffffu, mn);
    if ((threadIdx.x & 31) == 0) { if (cnt) atomicAdd(&scratch[2], cnt); atomicMin(&scratch[3], mn); }
    __syncthreads();
    vk1 = (scratch[2] > (unsigned)(k + 1)) ? vk : __uint_as_float(scratch[3]);
    __syncthreads();
}

__device__ __forceinline__ RunSeg shfl_down_seg(const RunSeg& s, int d) {
    RunSeg o;
    o.r.all = __shfl_down_sync(0xffffffffu, s.r.all, d); o.r.lead = __shfl_down_sync(0xffffffffu, s.r.lead, d);
    o.r.trail = __shfl_down_sync(0xffffffffu, s.r.trail, d); o.r.best_len = __shfl_down_sync(0xffffffffu, s.r.best_len, d);
    o.r.best_start = __shfl_down_sync(0xffffffffu, s.r.best_start, d);
    o.start = __shfl_down_sync(0xffffffffu, s.start, d); o.len = __shfl_down_sync(0xffffffffu, s.len, d);
    return o;
}

// Ordered (thread order = frame order) reduction of the CTA's segments with the associative merge: a shuffle tree per warp,
// then one over the warps' results.  wsm = 32 RunSeg of shared memory; the result is valid in thread 0.  Replaces a serial
// walk of thread 0 over 1024 summaries (~25 us).
__device__ RunSeg block_merge_segs(RunSeg s, RunSeg* wsm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const RunSeg o = shfl_down_seg(s, d);
        if (lane + d < 32) s = merge_seg(s, o);
    }
    if (lane == 0) wsm[warp] = s;
    __syncthreads();
    if (warp == 0) {
        s = lane < n_warps ? wsm[lane] : empty_seg();
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const RunSeg o = shfl_down_seg(s, d);
            if (lane + d < 32) s = merge_seg(s, o);
        }
    }
    return s;
}

// the duration test of the reference's loop on the longest closed run: (i - start) * hop / sr >= 0.1 (length ordering equals
// frame-count ordering)
__device__ __forceinline__ void pick_run(const RunSummary& acc, int hop, int sr, int& found, int& bs, int& bl) {
    pick_longest_closed(acc, bs, bl);
    found = bl > 0 && __ddiv_rn((double)((long long)bl * hop), (double)sr) >= 0.1;
}

__global__ void __launch_bounds__(SEL_THREADS) select_clip_kernel(const float* __restrict__ energies, const int* __restrict__ zc,
                                                                   const long long* __restrict__ foff, const int* __restrict__ fcount,
                                                                   const Item* __restrict__ items, int frame_len, int hop, int sr,
                                                                   int fallback_len, Clip* __restrict__ clips, int* __restrict__ seg_out) {
    __shared__ unsigned hist[2048];
    __shared__ unsigned scratch[4];
    __shared__ RunSeg wsm[32];
    __shared__ double thr[2];
    const int item = blockIdx.x;
    const Item it = items[item];
    const int F = fcount[item];
    const float* E = energies + foff[item];
    const int* Z = zc + foff[item];
    const int tid = threadIdx.x;
    int found = 0, best_s = 0, best_l = 0;
    if (F > 0) {
        const bool cached = F <= SEL_CACHE * SEL_THREADS;
        float eR[SEL_CACHE]; int zR[SEL_CACHE];
        if (cached) {
#pragma unroll
            for (int j = 0; j < SEL_CACHE; ++j) {
                const int i = tid + j * SEL_THREADS;
                eR[j] = i < F ? E[i] : 0.f; zR[j] = i < F ? Z[i] : 0;
            }
        } else {
#pragma unroll
            for (int j = 0; j < SEL_CACHE; ++j) { eR[j] = 0.f; zR[j] = 0; }
        }
        // --- energy 20th percentile
        {
            double vi = __dmul_rn((double)(F - 1), 0.2);        // (n - 1) * (20 / 100)
            double lo = floor(vi);
            int k = (int)lo;
            float a, b;
            radix_select2(cached, eR, E, F, k, hist, scratch, a, b);
            if (k + 1 > F - 1) b = a;
            if (tid == 0) thr[0] = np_lerp((double)a, (double)b, __dsub_rn(vi, lo));
        }
        // --- ZCR median from a histogram of the integer counts (count < frame_len <= 2047)
        {
            for (int i = tid; i < 2048; i += blockDim.x) hist[i] = 0;
            __syncthreads();
            if (cached) {
#pragma unroll
                for (int j = 0; j < SEL_CACHE; ++j) { if (tid + j * SEL_THREADS < F) atomicAdd(&hist[min(zR[j], 2047)], 1u); }
            } else {
                for (int i = tid; i < F; i += blockDim.x) atomicAdd(&hist[min(Z[i], 2047)], 1u);
            }
            __syncthreads();
            if (tid < 32) {
                // order statistics k and k + 1 of the counts: lanes sum 64 bins each, warp scan, the crossing lanes finish
                const int lane = tid;
                const double vi = __dmul_rn((double)(F - 1), 0.5);
                const double lo = floor(vi);
                const int k = (int)lo, k1 = min(k + 1, F - 1);
                unsigned tot = 0;
                for (int j = 0; j < 64; ++j) tot += hist[64 * lane + j];
                unsigned incl = tot;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                const unsigned excl = incl - tot;
                int ca = -1, cb = -1;
                if (excl <= (unsigned)k && incl > (unsigned)k) {
                    unsigned acc = excl;
                    for (int j = 0; j < 64; ++j) { acc += hist[64 * lane + j]; if (acc > (unsigned)k) { ca = 64 * lane + j; break; } }
                }
                if (excl <= (unsigned)k1 && incl > (unsigned)k1) {
                    unsigned acc = excl;
                    for (int j = 0; j < 64; ++j) { acc += hist[64 * lane + j]; if (acc > (unsigned)k1) { cb = 64 * lane + j; break; } }
                }
                ca = __reduce_max_sync(0xffffffffu, ca); cb = __reduce_max_sync(0xffffffffu, cb);
                if (lane == 0) {
                    double a = __ddiv_rn((double)ca, (double)frame_len), b = __ddiv_rn((double)cb, (double)frame_len);
                    thr[1] = __dmul_rn(np_lerp(a, b, __dsub_rn(vi, lo)), 0.5);
                }
            }
        }
        __syncthreads();
        const double eT = thr[0], zT = thr[1];
        // --- run detection: each thread summarises a contiguous segment; the segments are merged in frame order.  In cached
        // mode the (E < eT) & (ZCR > zT) flags are first written to shared memory (the histogram's 8192 bytes) by the threads
        // that hold the values.
        uint8_t* flags = reinterpret_cast<uint8_t*>(hist);
        if (cached) {
#pragma unroll
            for (int j = 0; j < SEL_CACHE; ++j) {
                const int i = tid + j * SEL_THREADS;
                if (i < F) flags[i] = (uint8_t)(((double)eR[j] < eT) && (__ddiv_rn((double)zR[j], (double)frame_len) > zT));
            }
            __syncthreads();
        }
        const int per = (F + SEL_THREADS - 1) / SEL_THREADS;
        const int s0 = min(F, tid * per), s1 = min(F, s0 + per);
        RunSeg seg; seg.start = s0; seg.len = s1 - s0;
        const RunSummary rs = summarize_range(s0, s1, [&](int i) {
            return cached ? flags[i] != 0 : (((double)E[i] < eT) && (__ddiv_rn((double)Z[i], (double)frame_len) > zT));
        });
        seg.r = rs;
        seg = block_merge_segs(seg, wsm);
        if (tid == 0) pick_run(seg.r, hop, sr, found, best_s, best_l);
    }
    if (tid == 0) {
        // sample range in GLOBAL coordinates of the recording; map to the item's local buffer
        long long gs, ge;
        if (found) { gs = (long long)best_s * hop; ge = (long long)(best_s + best_l) * hop; }
        else { gs = 0; ge = min((long long)fallback_len, it.n_global); }
        Clip c;
        c.start = it.pcm_off + gs - it.g0 + it.origin;
        long long len = ge - gs; if (len > CHUNK) len = CHUNK;
        c.len = (int)len; c.n_frames = clip_frames((int)len);
        clips[item] = c;
        if (seg_out) { seg_out[2 * item] = (int)gs; seg_out[2 * item + 1] = (int)ge; }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Long recordings (hundreds of thousands of frames: one hour has 360 000): the same selection, but by a whole GRID that
// co-operates through cooperative-groups grid barriers -- histograms in global memory, every CTA scanning its share of the
// frames, one warp of CTA 0 taking the (tiny) decisions.  One CTA needs ~0.8 ms for an hour and would dominate the
// stationary path; the grid needs ~20 barriers.  Results are bit-identical to select_clip_kernel (same order statistics,
// same float64 interpolation, same run rule); the run summaries are combined with an associative merge.
// ------------------------------------------------------------------------------------------------------------------
struct SelShared {                 // global scratch of one selection (SEL_SCRATCH_BYTES)
    unsigned hist[2048];
    unsigned prefix, mask, digit, acc, cnt_le, min_gt;
    int kk, pad_;
    double eT, zT;
    RunSummary sums[1024];         // one per CTA
};

__global__ void __launch_bounds__(SEL_THREADS) select_clip_grid_kernel(const float* __restrict__ E, const int* __restrict__ Z, int F,
                                                                        Item it, int frame_len, int hop, int sr, int fallback_len,
                                                                        SelShared* __restrict__ g, Clip* __restrict__ clip_out,
                                                                        int* __restrict__ seg_out) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ unsigned hist[2048];
    __shared__ RunSeg wsm[32];
    const int tid = threadIdx.x;
    const long long gtid = (long long)blockIdx.x * blockDim.x + tid, gsize = (long long)gridDim.x * blockDim.x;
    const bool leader = blockIdx.x == 0;
    // ---- energy: k-th and (k+1)-th smallest by radix select on the float bits (4 x 8 bits)
    const double vi_e = __dmul_rn((double)(F - 1), 0.2);
    const double lo_e = floor(vi_e);
    const int k_e = (int)lo_e;
    if (leader) {
        for (int i = tid; i < 256; i += blockDim.x) g->hist[i] = 0;
        if (tid == 0) { g->prefix = 0; g->mask = 0; g->kk = k_e; g->cnt_le = 0; g->min_gt = 0x7f800000u; }
    }
    grid.sync();
    for (int shift = 24; shift >= 0; shift -= 8) {
        const unsigned prefix = g->prefix, mask = g->mask;
        for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        {
            int cur = -1; unsigned run = 0;
            for (long long i = gtid; i < F; i += gsize) {
                const unsigned u = __float_as_uint(E[i]);
                if ((u & mask) != prefix) continue;
                const int d = (u >> shift) & 255;
                if (d == cur) { ++run; continue; }
                if (run) atomicAdd(&hist[cur], run);
                cur = d; run = 1;
            }
            if (run) atomicAdd(&hist[cur], run);
        }
        __syncthreads();
        for (int i = tid; i < 256; i += blockDim.x) if (hist[i]) atomicAdd(&g->hist[i], hist[i]);
        grid.sync();
        if (leader && tid < 32) {
            const int lane = tid, kk = g->kk;
            unsigned h[8], tot = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { h[j] = g->hist[8 * lane + j]; tot += h[j]; }
            unsigned incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            const unsigned excl = incl - tot;
            const unsigned hit = __ballot_sync(0xffffffffu, incl > (unsigned)kk);
            if (hit && lane == __ffs(hit) - 1) {
                unsigned acc = excl; int j = 0;
                for (; j < 8; ++j) { if (acc + h[j] > (unsigned)kk) break; acc += h[j]; }
                g->prefix = prefix | ((unsigned)(8 * lane + j) << shift); g->mask = mask | (255u << shift); g->kk = kk - (int)acc;
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) g->hist[8 * lane + j] = 0;                 // ready for the next pass
        }
        grid.sync();
    }
    {   // count of elements <= v_k and the smallest element above it
        const unsigned vk = g->prefix;
        unsigned cnt = 0, mn = 0x7f800000u;
        for (long long i = gtid; i < F; i += gsize) { const unsigned u = __float_as_uint(E[i]); if (u <= vk) ++cnt; else mn = min(mn, u); }
        cnt = __reduce_add_sync(0xffffffffu, cnt); mn = __reduce_min_sync(0xffffffffu, mn);
        if ((tid & 31) == 0) { if (cnt) atomicAdd(&g->cnt_le, cnt); atomicMin(&g->min_gt, mn); }
    }
    // ---- ZCR histogram (counts < frame_len <= 2047)
    if (leader) for (int i = tid; i < 2048; i += blockDim.x) g->hist[i] = 0;
    grid.sync();
    for (int i = tid; i < 2048; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (long long i = gtid; i < F; i += gsize) atomicAdd(&hist[min(Z[i], 2047)], 1u);
    __syncthreads();
    for (int i = tid; i < 2048; i += blockDim.x) if (hist[i]) atomicAdd(&g->hist[i], hist[i]);
    grid.sync();
    if (leader && tid < 32) {
        const int lane = tid;
        if (lane == 0) {
            const float a = __uint_as_float(g->prefix);
            float b = (g->cnt_le > (unsigned)(k_e + 1)) ? a : __uint_as_float(g->min_gt);
            if (k_e + 1 > F - 1) b = a;
            g->eT = np_lerp((double)a, (double)b, __dsub_rn(vi_e, lo_e));
        }
        const double vi = __dmul_rn((double)(F - 1), 0.5);
        const double lo = floor(vi);
        const int k = (int)lo, k1 = min(k + 1, F - 1);
        unsigned tot = 0;
        for (int j = 0; j < 64; ++j) tot += g->hist[64 * lane + j];
        unsigned incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const unsigned excl = incl - tot;
        int ca = -1, cb = -1;
        if (excl <= (unsigned)k && incl > (unsigned)k) {
            unsigned acc = excl;
            for (int j = 0; j < 64; ++j) { acc += g->hist[64 * lane + j]; if (acc > (unsigned)k) { ca = 64 * lane + j; break; } }
        }
        if (excl <= (unsigned)k1 && incl > (unsigned)k1) {
            unsigned acc = excl;
            for (int j = 0; j < 64; ++j) { acc += g->hist[64 * lane + j]; if (acc > (unsigned)k1) { cb = 64 * lane + j; break; } }
        }
        ca = __reduce_max_sync(0xffffffffu, ca); cb = __reduce_max_sync(0xffffffffu, cb);
        if (lane == 0) {
            const double a = __ddiv_rn((double)ca, (double)frame_len), b = __ddiv_rn((double)cb, (double)frame_len);
            g->zT = __dmul_rn(np_lerp(a, b, __dsub_rn(vi, lo)), 0.5);
        }
    }
    grid.sync();
    // ---- runs: CTA c owns frames [c per_cta, (c+1) per_cta), thread t a contiguous part of them
    const double eT = g->eT, zT = g->zT;
    const int per_cta = (F + gridDim.x - 1) / gridDim.x;
    const int per = (per_cta + SEL_THREADS - 1) / SEL_THREADS;
    const int c0 = min(F, blockIdx.x * per_cta), c1 = min(F, c0 + per_cta);
    {
        const int s0 = min(c1, c0 + tid * per), s1 = min(c1, s0 + per);
        const RunSummary rs = summarize_range(s0, s1, [&](int i) {
            return ((double)E[i] < eT) && (__ddiv_rn((double)Z[i], (double)frame_len) > zT);
        });
        RunSeg seg; seg.r = rs; seg.start = s0; seg.len = s1 - s0;
        seg = block_merge_segs(seg, wsm);
        if (tid == 0) g->sums[blockIdx.x] = seg.r;
    }
    grid.sync();
    if (leader) {
        RunSeg seg = empty_seg();
        if (tid < (int)gridDim.x) {
            const int a0 = min(F, tid * per_cta), a1 = min(F, a0 + per_cta);
            seg.r = g->sums[tid]; seg.start = a0; seg.len = a1 - a0;
        }
        seg = block_merge_segs(seg, wsm);
        if (tid == 0) {
            int found, bs, bl;
            pick_run(seg.r, hop, sr, found, bs, bl);
            long long gs, ge;
            if (found) { gs = (long long)bs * hop; ge = (long long)(bs + bl) * hop; }
            else { gs = 0; ge = min((long long)fallback_len, it.n_global); }
            Clip c;
            c.start = it.pcm_off + gs - it.g0 + it.origin;
            long long len2 = ge - gs; if (len2 > CHUNK) len2 = CHUNK;
            c.len = (int)len2; c.n_frames = clip_frames((int)len2);
            *clip_out = c;
            if (seg_out) { seg_out[0] = (int)gs; seg_out[1] = (int)ge; }
        }
    }
}

static_assert(sizeof(SelShared) <= SEL_SCRATCH_BYTES, "selection scratch too small");

void launch_select_clip(const float* energies, const int* zc, const long long* foff, const int* fcount,
                        const Item* items, int n_items, int frame_len, int hop, int sr, int fallback_len,
                        Clip* clips, int* seg_out, cudaStream_t st) {
    if (n_items <= 0) return;
    select_clip_kernel<<<n_items, SEL_THREADS, 0, st>>>(energies, zc, foff, fcount, items, frame_len, hop, sr,
                                                       fallback_len, clips, seg_out);
}

// one long recording: host-side item description, `scratch` = SEL_SCRATCH_BYTES of device memory owned by the caller
void launch_select_clip_grid(const float* E, const int* Z, int F, const Item& it, int frame_len, int hop, int sr, int fallback_len,
                             void* scratch, Clip* clip_out, int* seg_out, cudaStream_t st) {
    static int n_sm = 0;
    if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
    int grid = n_sm < 1024 ? n_sm : 1024;
    SelShared* g = (SelShared*)scratch;
    Item itc = it;
    void* args[] = {(void*)&E, (void*)&Z, (void*)&F, (void*)&itc, (void*)&frame_len, (void*)&hop, (void*)&sr, (void*)&fallback_len,
                    (void*)&g, (void*)&clip_out, (void*)&seg_out};
    cudaLaunchCooperativeKernel((const void*)select_clip_grid_kernel, dim3(grid), dim3(SEL_THREADS), args, 0, st);
}

}  // namespace maf
