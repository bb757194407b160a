// Front door of normalize_audio (reference: audio_pipeline/preprocessor.py:227-233 -> pydub normalize(seg)
// .set_frame_rate(sr).set_channels(1).set_sample_width(2), i.e. stdlib audioop.max / mul / ratecv / tomono in this order)
// for 16-bit input that is not yet mono at the target rate.  Restated in
// oracle/reference_path.py (ratecv / tomono / front_door), which tests/test_oracle.py pins against the real audioop.
//
// audioop.ratecv with the default weights is a sequential loop over a phase counter d, but the phase of output frame m is
// a closed form -- it reads input frame n = ceil(m in / out) with d = n out - m in (rates divided by their gcd) -- so every
// output frame is independent:  out = (int)((prev * d + cur * (out - d)) / out) >> 16  on the samples scaled to 32 bits,
// in float64 like the C code (correctly rounded division, truncation toward zero).  Peak normalisation (floor(s * factor))
// is applied to the samples as they are read, the stereo downmix floor((l + r) / 2) to the resampled pair: one pass.
#include "maf_common.cuh"

namespace maf {

__global__ void __launch_bounds__(256) abs_max_flat_kernel(const int16_t* __restrict__ x, long long n, int* __restrict__ peak) {
    int pk = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int s = x[i];
        pk = max(pk, s < 0 ? -s : s);                              // audioop.max: 32768 for -32768
    }
    pk = __reduce_max_sync(0xffffffffu, pk);
    if ((threadIdx.x & 31) == 0 && pk > 0) atomicMax(peak, pk);
}

__global__ void __launch_bounds__(256) front_door_kernel(const int16_t* __restrict__ x, long long n_in, int channels, long long inr,
                                                         long long outr, const int* __restrict__ peak, double target_peak,
                                                         int peak_normalize, int16_t* __restrict__ out, long long n_out) {
    __shared__ double s_factor;
    __shared__ int s_pn;
    if (threadIdx.x == 0) {
        const int pk = peak_normalize ? *peak : 0;
        s_pn = peak_normalize && pk > 0;
        s_factor = s_pn ? peak_factor(pk, target_peak) : 1.0;
    }
    __syncthreads();
    const bool pn = s_pn != 0;
    const double factor = s_factor;
    const double doutr = (double)outr;
    for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < n_out; m += (long long)gridDim.x * blockDim.x) {
        const long long n = (m * inr + outr - 1) / outr;           // ceil(m in / out)
        const double d = (double)(n * outr - m * inr);
        int v[2];
        for (int c = 0; c < channels && c < 2; ++c) {
            int cur = x[n * channels + c];
            if (pn) cur = peak_norm_sample(cur, factor);
            if (inr == outr) { v[c] = cur; continue; }
            int prev = 0;
            if (n > 0) { prev = x[(n - 1) * channels + c]; if (pn) prev = peak_norm_sample(prev, factor); }
            const double pv = (double)(prev * 65536), cv = (double)(cur * 65536);      // GETSAMPLE32: sample << 16
            const double val = (pv * d + cv * (doutr - d)) / doutr;
            v[c] = __double2int_rz(val) >> 16;                     // (int) cast, SETSAMPLE32 keeps the high 16 bits
        }
        int o = v[0];
        if (channels == 2) o = (int)floor(fmin(fmax((double)v[0] * 0.5 + (double)v[1] * 0.5, -32768.0), 32767.0));   // audioop.tomono
        out[m] = (int16_t)o;
    }
}

void launch_front_door(const int16_t* x, long long n_in, int channels, long long inr, long long outr, int* peak_scratch,
                       double target_peak, int peak_normalize, int16_t* out, long long n_out, cudaStream_t st) {
    if (n_in <= 0 || n_out <= 0) return;
    if (peak_normalize) {
        launch_zero_bytes(peak_scratch, sizeof(int), st);
        const long long n = n_in * channels;
        long long bx = (n + 256 * 16 - 1) / (256 * 16);
        if (bx > 2048) bx = 2048;
        abs_max_flat_kernel<<<(unsigned)bx, 256, 0, st>>>(x, n, peak_scratch);
    }
    long long bx = (n_out + 255) / 256;
    if (bx > 8192) bx = 8192;
    front_door_kernel<<<(unsigned)bx, 256, 0, st>>>(x, n_in, channels, inr, outr, peak_scratch, target_peak, peak_normalize, out, n_out);
}

}  // namespace maf
