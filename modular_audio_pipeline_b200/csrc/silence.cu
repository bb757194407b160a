// Silence detection for AudioPreprocessor.remove_silence / detect_silence_segments (reference
// audio_pipeline/preprocessor.py:290-405 -> pydub.silence.detect_silence over audioop.rms).
// pydub slides a min_silence_len window in 1 ms steps and calls audioop.rms on every slice (O(N * window)); here the
// sums of squares are exact integers per millisecond bucket and a window is the sum of its buckets.
#include "maf_common.cuh"

namespace maf {

// bucket i = sum of s^2 over samples [int(i * r), int((i + 1) * r)), r = rate / 1000.0 (pydub frame_count(ms))
__global__ void __launch_bounds__(256) ms_energy_kernel(const int16_t* __restrict__ x, long long n, double r, long long n_ms,
                                                        unsigned long long* __restrict__ bucket, unsigned long long* __restrict__ total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long acc = 0;
    if (i < n_ms) {
        long long lo = (long long)((double)i * r), hi = (long long)((double)(i + 1) * r);
        if (hi > n) hi = n;
        for (long long j = lo; j < hi; ++j) { const long long v = x[j]; acc += (unsigned long long)(v * v); }
        bucket[i] = acc;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(total, acc);
}

// flag[i] = audioop.rms(seg[i : i + win]) <= thresh for every start i in 0 .. n_ms - win
__global__ void __launch_bounds__(256) silence_flags_kernel(const unsigned long long* __restrict__ bucket, long long n, double r,
                                                            long long n_starts, int win, double thresh, uint8_t* __restrict__ flags) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_starts) return;
    unsigned long long acc = 0;
    for (int j = 0; j < win; ++j) acc += bucket[i + j];
    long long lo = (long long)((double)i * r), hi = (long long)((double)(i + win) * r);
    if (hi > n) hi = n;
    const long long cnt = hi - lo;
    unsigned int rms = 0;
    if (cnt > 0) rms = (unsigned int)sqrt((double)acc / (double)cnt);
    flags[i] = (double)rms <= thresh ? 1 : 0;
}

void launch_ms_energy(const int16_t* x, long long n, int rate, long long n_ms, unsigned long long* bucket, unsigned long long* total,
                      cudaStream_t st) {
    launch_zero_bytes(total, 8, st);
    if (n_ms <= 0) return;
    ms_energy_kernel<<<(unsigned)((n_ms + 255) / 256), 256, 0, st>>>(x, n, (double)rate / 1000.0, n_ms, bucket, total);
}

void launch_silence_flags(const unsigned long long* bucket, long long n, int rate, long long n_starts, int win, double thresh,
                          uint8_t* flags, cudaStream_t st) {
    if (n_starts <= 0) return;
    silence_flags_kernel<<<(unsigned)((n_starts + 255) / 256), 256, 0, st>>>(bucket, n, (double)rate / 1000.0, n_starts, win, thresh, flags);
}

}  // namespace maf
