// S5-S7 -- peak normalise + BS.1770 loudness normalisation.
// References: audio_pipeline/preprocessor.py:227-233 (pydub.effects.normalize -> audioop.max / audioop.mul) and
// :256-279 (pyloudnorm.Meter(rate).integrated_loudness, normalize.loudness, peak limiter, int16 quantisation).
//
//   kweight_energy_kernel  K4: peak-normalise on load, the two K-weighting biquads (float64 recurrences, float32
//                          re-quantisation between and after the stages exactly as lfilter's float32 round trips
//                          do), and the 100 ms sub-block energies.  The recurrence is parallelised as independent
//                          tiles: every thread owns one sub-block and first runs KW_WARM = 2048 warm-up samples
//                          from zero state (|pole| <= 0.9852 -> residual 5.6e-14).
//   gate_kernel            400 ms / 75 % overlap block energies = 4 consecutive sub-blocks, absolute (-70 LUFS) and
//                          relative (-10 LU) gates, integrated loudness, float32 gain and limiter peak per recording.
//   apply_gain_kernel      gain / limiter / int16 truncation when the log-mel kernel (which fuses it) is not run.
#include "maf_common.cuh"

namespace maf {

__global__ void __launch_bounds__(128) kweight_energy_kernel(const int16_t* __restrict__ den16, const Item* __restrict__ items,
                                                             int tile_len, KWeight kw, int peak_normalize, double target_peak,
                                                             const int* __restrict__ peak, double* __restrict__ sub_energy,
                                                             int* __restrict__ peak_norm) {
    const int item = blockIdx.y;
    const Item it = items[item];
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    int pk = 0;
    if (tile < it.n_ktiles) {
        const int pkin = peak_normalize ? peak[item] : 0;
        const bool pn = peak_normalize && pkin > 0;
        const double factor = pn ? peak_factor(pkin, target_peak) : 1.0;
        const int16_t* x = den16 + it.den_off;
        const long long start = (long long)tile * tile_len;
        long long begin = start - KW_WARM;
        if (begin < -(long long)it.lo_avail) begin = -(long long)it.lo_avail;
        long long end = start + tile_len;
        if (end > it.own_len) end = it.own_len;
        // direct form II transposed, like scipy.signal.lfilter.  Samples are fetched and converted eight at a time so
        // that only the four dependent DFMAs per sample sit on the recurrence's critical path.
        double z10 = 0.0, z11 = 0.0, z20 = 0.0, z21 = 0.0;
        double acc = 0.0;
        auto step = [&](float xin, int s, bool own) {
            const double xd = (double)xin;
            const double y1 = kw.b1[0] * xd + z10;
            z10 = (kw.b1[1] * xd + z11) - kw.a1[1] * y1;
            z11 = kw.b1[2] * xd - kw.a1[2] * y1;
            const double x2 = (double)(float)y1;
            const double y2 = kw.b2[0] * x2 + z20;
            z20 = (kw.b2[1] * x2 + z21) - kw.a2[1] * y2;
            z21 = kw.b2[2] * x2 - kw.a2[2] * y2;
            const float d = (float)y2;
            if (own) {
                acc += (double)__fmul_rn(d, d);
                pk = max(pk, s < 0 ? -s : s);
            }
        };
        long long i = begin;
        const bool aligned = ((it.den_off + begin) & 7) == 0;
        for (; i + 8 <= end; i += 8) {
            int sv[8];
            if (aligned) {
                const uint4 v = *reinterpret_cast<const uint4*>(x + i);
                sv[0] = (short)(v.x & 0xffff); sv[1] = (short)(v.x >> 16); sv[2] = (short)(v.y & 0xffff); sv[3] = (short)(v.y >> 16);
                sv[4] = (short)(v.z & 0xffff); sv[5] = (short)(v.z >> 16); sv[6] = (short)(v.w & 0xffff); sv[7] = (short)(v.w >> 16);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) sv[e] = x[i + e];
            }
            float xin[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                if (pn) sv[e] = peak_norm_sample(sv[e], factor);
                xin[e] = (float)sv[e] * (1.0f / 32768.0f);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) step(xin[e], sv[e], i + e >= start);
        }
        for (; i < end; ++i) {
            int sv = x[i];
            if (pn) sv = peak_norm_sample(sv, factor);
            step((float)sv * (1.0f / 32768.0f), sv, i >= start);
        }
        sub_energy[it.ktile_first + tile] = acc;
    }
    pk = max(pk, __shfl_xor_sync(0xffffffffu, pk, 16));
    pk = max(pk, __shfl_xor_sync(0xffffffffu, pk, 8));
    pk = max(pk, __shfl_xor_sync(0xffffffffu, pk, 4));
    pk = max(pk, __shfl_xor_sync(0xffffffffu, pk, 2));
    pk = max(pk, __shfl_xor_sync(0xffffffffu, pk, 1));
    if ((threadIdx.x & 31) == 0 && pk > 0) atomicMax(&peak_norm[item], pk);
}

void launch_kweight_energy(const int16_t* den16, const Item* items, int n_items, int max_tiles, int tile_len,
                           const KWeight& kw, int peak_normalize, double target_peak, const int* peak,
                           double* sub_energy, int* peak_norm, cudaStream_t st) {
    if (n_items <= 0 || max_tiles <= 0) return;
    for (int y0 = 0; y0 < n_items; y0 += 32768) {
        const int ny = n_items - y0 < 32768 ? n_items - y0 : 32768;
        dim3 grid((max_tiles + 127) / 128, ny);
        kweight_energy_kernel<<<grid, 128, 0, st>>>(den16, items + y0, tile_len, kw, peak_normalize, target_peak,
                                                    peak ? peak + y0 : nullptr, sub_energy, peak_norm + y0);
    }
}

// ---- gating -----------------------------------------------------------------------------------------------------
constexpr int GATE_THREADS = 1024;

__device__ double block_sum(double v, double* sh) {
    // deterministic tree reduction over the CTA
    const int tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (int s = GATE_THREADS / 2; s > 0; s >>= 1) {
        if (tid < s) sh[tid] += sh[tid + s];
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}

// The two gates are evaluated on the block energies themselves: l_j >= -70 <=> z_j >= 10^((-70 + 0.691)/10) and
// l_j > Gamma_r <=> z_j > mean(z_abs_gated) / 10 (BS.1770-4 eq. 5-7; pyloudnorm compares the equivalent dB values).
__global__ void __launch_bounds__(GATE_THREADS) gate_kernel(const double* __restrict__ sub_energy, const Item* __restrict__ items,
                                                            double inv_block, double target_lufs, const int* __restrict__ peak_norm,
                                                            Gate* __restrict__ gates) {
    __shared__ double sh[GATE_THREADS];
    const int item = blockIdx.x;
    const Item it = items[item];
    const double* sb = sub_energy + it.ktile_first;
    const int nb = it.n_blocks;
    const double zabs = pow(10.0, (-70.0 + 0.691) / 10.0);
    double s = 0.0, c = 0.0;
    for (int j = threadIdx.x; j < nb; j += GATE_THREADS) {
        double e = sb[j];
        if (j + 1 < it.n_ktiles) e += sb[j + 1];
        if (j + 2 < it.n_ktiles) e += sb[j + 2];
        if (j + 3 < it.n_ktiles) e += sb[j + 3];
        const double z = inv_block * e;
        if (z >= zabs) { s += z; c += 1.0; }
    }
    s = block_sum(s, sh); c = block_sum(c, sh);
    const double zrel = c > 0.0 ? (s / c) * 0.1 : NAN;                 // NaN threshold: nothing passes, like np.mean([])
    double s2 = 0.0, c2 = 0.0;
    for (int j = threadIdx.x; j < nb; j += GATE_THREADS) {
        double e = sb[j];
        if (j + 1 < it.n_ktiles) e += sb[j + 1];
        if (j + 2 < it.n_ktiles) e += sb[j + 2];
        if (j + 3 < it.n_ktiles) e += sb[j + 3];
        const double z = inv_block * e;
        if (z > zrel && z > zabs) { s2 += z; c2 += 1.0; }
    }
    s2 = block_sum(s2, sh); c2 = block_sum(c2, sh);
    if (threadIdx.x == 0) {
        const double zavg = c2 > 0.0 ? s2 / c2 : 0.0;                 // nan_to_num(mean([])) = 0
        const double lufs = -0.691 + 10.0 * log10(zavg);
        Gate gt;
        gt.lufs = lufs;
        gt.applied = (isfinite(lufs) && !(lufs < -70.0)) ? 1 : 0;
        gt.gain = 1.0f; gt.limp = 0.0f; gt.lim = 0;
        if (gt.applied) {
            gt.gain = (float)pow(10.0, (target_lufs - lufs) / 20.0);
            const float xpk = (float)peak_norm[item] * (1.0f / 32768.0f);
            gt.limp = __fmul_rn(gt.gain, xpk);
            gt.lim = gt.limp > 1.0f ? 1 : 0;
        }
        gates[item] = gt;
    }
}

void launch_gate(const double* sub_energy, const Item* items, int n_items, double inv_block, double target_lufs,
                 const int* peak_norm, Gate* gates, cudaStream_t st) {
    if (n_items <= 0) return;
    gate_kernel<<<n_items, GATE_THREADS, 0, st>>>(sub_energy, items, inv_block, target_lufs, peak_norm, gates);
}

// ---- standalone gain application ------------------------------------------------------------------------------------
// float32 chain of preprocessor.py:257, :271-279: x = s / 32768 ; y = gain * x ; y /= peak (if peak > 1) ;
// clip(y * 32768, -32768, 32767).astype(int16)
__device__ __forceinline__ int gain_sample(int s, const Gate& gt) {
    if (!gt.applied) return s;
    float y = __fmul_rn(gt.gain, (float)s * (1.0f / 32768.0f));
    if (gt.lim) y = __fdiv_rn(y, gt.limp);
    float v = __fmul_rn(y, 32768.0f);
    v = fminf(fmaxf(v, -32768.0f), 32767.0f);
    return (int)v;
}

__global__ void __launch_bounds__(256) apply_gain_kernel(const int16_t* __restrict__ den16, const Item* __restrict__ items,
                                                         int peak_normalize, double target_peak, const int* __restrict__ peak,
                                                         const Gate* __restrict__ gates, int use_gate, int16_t* __restrict__ out16) {
    const int item = blockIdx.y;
    const Item it = items[item];
    const int pkin = peak_normalize ? peak[item] : 0;
    const bool pn = peak_normalize && pkin > 0;
    const double factor = pn ? peak_factor(pkin, target_peak) : 1.0;
    Gate gt; gt.applied = 0;
    if (use_gate) gt = gates[item];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < it.own_len; i += (long long)gridDim.x * blockDim.x) {
        int s = den16[it.den_off + i];
        if (pn) s = peak_norm_sample(s, factor);
        if (use_gate) s = gain_sample(s, gt);
        out16[it.out_off + i] = (int16_t)s;
    }
}

void launch_apply_gain(const int16_t* den16, const Item* items, int n_items, long long max_len, int peak_normalize,
                       double target_peak, const int* peak, const Gate* gates, int use_gate, int16_t* out16,
                       cudaStream_t st) {
    if (n_items <= 0 || max_len <= 0) return;
    long long bx = (max_len + 256 * 8 - 1) / (256 * 8);
    if (bx > 4096) bx = 4096;
    for (int y0 = 0; y0 < n_items; y0 += 32768) {
        const int ny = n_items - y0 < 32768 ? n_items - y0 : 32768;
        dim3 grid((unsigned)bx, ny);
        apply_gain_kernel<<<grid, 256, 0, st>>>(den16, items + y0, peak_normalize, target_peak, peak ? peak + y0 : nullptr,
                                                gates ? gates + y0 : nullptr, use_gate, out16);
    }
}

}  // namespace maf
