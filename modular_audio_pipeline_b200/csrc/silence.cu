// Silence detection and removal for AudioPreprocessor.remove_silence / detect_silence_segments (reference
// audio_pipeline/preprocessor.py:290-405 -> pydub.silence.detect_silence / detect_nonsilent over audioop.rms, then the
// reference's own margin / cross-fade / position loop :330-358 over AudioSegment.__getitem__ / append).
//
// pydub slides a min_silence_len window in 1 ms steps and calls audioop.rms on every slice (O(N * window)), merges the silent
// starts into ranges in a Python loop, and the reference then concatenates the kept chunks one AudioSegment.append at a time.
// Here everything after the recording-wide RMS (one scalar the host turns into the threshold with the reference's own
// float64 expression) stays on the device:
//   ms_energy_kernel        exact integer sum of squares per millisecond bucket [int(i r), int((i+1) r)), r = rate / 1000
//   scans                   device-wide exclusive prefix sums (CTA-local scan + scan of the CTA totals): bucket energies ->
//                           window sums in O(1); silent flags -> "any silent start within W ms"; range firsts -> ranks
//   silence_ranges          pydub's merge rule in closed form: a silent start opens a range iff no silent start lies in the
//                           W ms before it, closes one iff none lies in the W ms after it -> [first, last + W]
//   silence_table_kernel    detect_nonsilent's complement, the reference's margin / duration / cross-fade / position
//                           bookkeeping (one short serial pass over the ranges for the running positions)
//   silence_compact_kernel  stream compaction of the kept chunks with pydub's cross-fade (per-frame audioop.mul fades,
//                           saturating audioop.add) applied in append order
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

void launch_ms_energy(const int16_t* x, long long n, int rate, long long n_ms, unsigned long long* bucket, unsigned long long* total,
                      cudaStream_t st) {
    launch_zero_bytes(total, 8, st);
    if (n_ms <= 0) return;
    ms_energy_kernel<<<(unsigned)((n_ms + 255) / 256), 256, 0, st>>>(x, n, (double)rate / 1000.0, n_ms, bucket, total);
}

// ---- device-wide exclusive scan: CTA-local scan (SC_TILE elements per CTA) + scan of the CTA totals --------------------------
constexpr int SC_T = 256, SC_I = 8, SC_TILE = SC_T * SC_I;

template <typename T, typename F>
__device__ __forceinline__ void cta_scan(long long n, F value, T* __restrict__ loc, T* __restrict__ tot) {
    __shared__ T wsum[SC_T / 32];
    const long long base = (long long)blockIdx.x * SC_TILE + (long long)threadIdx.x * SC_I;
    T x[SC_I];
    T run = 0;
#pragma unroll
    for (int k = 0; k < SC_I; ++k) { const T v = base + k < n ? value(base + k) : (T)0; x[k] = run; run += v; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const T t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    T woff = 0;
    for (int w = 0; w < warp; ++w) woff += wsum[w];
    const T off = woff + inc - run;
#pragma unroll
    for (int k = 0; k < SC_I; ++k) if (base + k < n) loc[base + k] = x[k] + off;
    if (threadIdx.x == SC_T - 1) tot[blockIdx.x] = woff + inc;
}

// exclusive scan of the nb CTA totals in place, by one CTA; grand[0] = sum of all
template <typename T>
__global__ void __launch_bounds__(1024) tot_scan_kernel(T* __restrict__ tot, int nb, T* __restrict__ grand) {
    __shared__ T sh[1024];
    __shared__ T carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const T v = i < nb ? tot[i] : (T)0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const T t = threadIdx.x >= o ? sh[threadIdx.x - o] : (T)0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        const T c = carry;
        if (i < nb) tot[i] = c + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0 && grand) grand[0] = carry;
}

template <typename T>
__device__ __forceinline__ T scan_at(const T* __restrict__ loc, const T* __restrict__ tot, long long i) { return loc[i] + tot[i / SC_TILE]; }

struct SilHeader {          // first 64 bytes of the silence workspace; copied to the host with the tables
    int status;             // 0 ok, 1 = "no non-silent segments" (reference returns the input), 2 = pydub ValueError (cross-fade
                            // longer than the audio accumulated so far), 3 = pydub TooManyMissingFrames
    int n_silent;           // silent ranges
    int n_chunks;           // kept chunks (TimestampMapping entries)
    int j_base;             // first chunk that is part of the output (pydub's `if len(cleaned) == 0: cleaned = chunk`)
    long long n_out;        // frames of the cleaned recording
    long long n_ms;
    long long pad_[4];
};

struct SilWs {              // device pointers into the caller's workspace
    SilHeader* hdr;
    unsigned long long *pb_loc, *pb_tot;     // bucket prefix
    unsigned *cf_loc, *cf_tot, *cf_grand;    // silent-flag prefix (n_starts + 1 entries)
    unsigned *cr_loc, *cr_tot, *cr_grand;    // range-first prefix
    int* sil;               // [cap][2] silent ranges in ms
    int* map;               // [cap + 1][4] pos_ms, dur_ms, s_ms, e_ms per kept chunk
    long long* piece;       // [cap + 1][4] f0, len, out_off, ncf per kept chunk
    long long* pend;        // [cap + 1] cumulative end E_j of the kept chunks
    uint8_t* flags;         // [n_starts + 1]
    long long cap;
};

__global__ void __launch_bounds__(SC_T) bucket_scan_kernel(const unsigned long long* __restrict__ bucket, long long n_ms, SilWs w) {
    cta_scan<unsigned long long>(n_ms + 1, [&](long long i) { return i < n_ms ? bucket[i] : 0ull; }, w.pb_loc, w.pb_tot);
}

// flag[i] = audioop.rms(seg[i : i + win]) <= thresh for every start i in 0 .. n_ms - win.  AudioSegment.__getitem__ pads a
// slice that runs past the last frame with silence up to the expected length, so the frame count of a window is
// int((i + win) r) - int(i r) even where that exceeds the recording by a fraction of a millisecond.
__global__ void __launch_bounds__(SC_T) flag_scan_kernel(long long n_starts, double r, int win, double thresh, SilWs w) {
    cta_scan<unsigned>(n_starts + 1, [&](long long i) {
        unsigned f = 0;
        if (i < n_starts) {
            const unsigned long long acc = scan_at(w.pb_loc, w.pb_tot, i + win) - scan_at(w.pb_loc, w.pb_tot, i);
            const long long cnt = (long long)((double)(i + win) * r) - (long long)((double)i * r);
            unsigned int rms = 0;
            if (cnt > 0) rms = (unsigned int)sqrt((double)acc / (double)cnt);
            f = (double)rms <= thresh ? 1u : 0u;
        }
        w.flags[i] = (uint8_t)f;
        return f;
    }, w.cf_loc, w.cf_tot);
}

__device__ __forceinline__ bool sil_first(const SilWs& w, long long s, int win) {
    if (!w.flags[s]) return false;
    const long long lo = s - win > 0 ? s - win : 0;
    return scan_at(w.cf_loc, w.cf_tot, s) - scan_at(w.cf_loc, w.cf_tot, lo) == 0;
}

__global__ void __launch_bounds__(SC_T) first_scan_kernel(long long n_starts, int win, SilWs w) {
    cta_scan<unsigned>(n_starts + 1, [&](long long s) { return s < n_starts && sil_first(w, s, win) ? 1u : 0u; }, w.cr_loc, w.cr_tot);
}

__global__ void __launch_bounds__(256) ranges_scatter_kernel(long long n_starts, int win, SilWs w) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_starts || !w.flags[s]) return;
    const bool first = sil_first(w, s, win);
    const unsigned rank = scan_at(w.cr_loc, w.cr_tot, s);           // range firsts before s
    if (first && rank < w.cap) w.sil[2 * rank] = (int)s;
    const long long hi = s + win + 1 < n_starts ? s + win + 1 : n_starts;
    const bool last = scan_at(w.cf_loc, w.cf_tot, hi) - scan_at(w.cf_loc, w.cf_tot, s + 1) == 0;
    const unsigned k = rank + (first ? 1u : 0u);                    // ranges opened up to and including s
    if (last && k >= 1 && k - 1 < w.cap) w.sil[2 * (k - 1) + 1] = (int)s + win;
}

__device__ __forceinline__ long long ms_len(long long frames, double sr) { return (long long)rint(1000.0 * ((double)frames / sr)); }   // pydub len()

// detect_nonsilent's complement of the silent ranges, then the loop of preprocessor.py:330-358.
__global__ void __launch_bounds__(1024) silence_table_kernel(long long n, double sr, double r, long long n_ms, int margin, int have_ranges, SilWs w) {
    __shared__ int s_nchunks, s_shift, s_status;
    const int K = have_ranges ? (int)min((long long)w.cr_grand[0], w.cap) : 0;
    if (threadIdx.x == 0) {
        int st = 0, nc, shift = 0;
        if (K == 0) nc = 1;
        else if (w.sil[0] == 0 && w.sil[1] == n_ms) { nc = 0; st = 1; }
        else {
            shift = w.sil[0] == 0 ? 1 : 0;                          // out[0] == [0, 0] is dropped
            nc = K + (w.sil[2 * (K - 1) + 1] != n_ms ? 1 : 0) - shift;
            if (nc <= 0) st = 1;
        }
        s_nchunks = nc; s_shift = shift; s_status = st;
    }
    __syncthreads();
    const int nc = s_nchunks, shift = s_shift;
    // per-chunk quantities (independent of the history)
    for (int j = threadIdx.x; j < nc; j += blockDim.x) {
        long long a, b;
        if (K == 0) { a = 0; b = n_ms; }
        else {
            const int c = j + shift;                                // candidate index: (0, a_0), (b_0, a_1), ..., (b_{K-1}, n_ms)
            a = c == 0 ? 0 : w.sil[2 * (c - 1) + 1];
            b = c < K ? w.sil[2 * c] : n_ms;
        }
        const long long s = a - margin > 0 ? a - margin : 0, e = b + margin < n_ms ? b + margin : n_ms;
        const long long f0 = (long long)((double)s * r), f1 = (long long)((double)e * r);     // AudioSegment._parse_position
        const long long len = f1 > f0 ? f1 - f0 : 0;
        const long long dur = ms_len(len, sr);
        const long long cf = dur / 4 < 20 ? dur / 4 : 20;
        w.map[4 * j + 1] = (int)dur; w.map[4 * j + 2] = (int)s; w.map[4 * j + 3] = (int)e;
        w.piece[4 * j] = f0; w.piece[4 * j + 1] = len; w.piece[4 * j + 3] = (long long)((double)cf * r);
        // frames missing at the end of the recording are padded with silence unless there are more than 2 ms of them
        const long long have = f0 < n ? (f1 < n ? f1 : n) - f0 : 0;
        if (len - (have > 0 ? have : 0) > (long long)(2.0 * r)) atomicMax(&s_status, 3);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long total = 0, pos = 0; int j_base = 0, st = s_status;
        for (int j = 0; j < nc; ++j) {
            const long long len = w.piece[4 * j + 1], dur = w.map[4 * j + 1];
            w.map[4 * j] = (int)pos;
            if (ms_len(total, sr) == 0) {                           // `if len(cleaned) == 0: cleaned = chunk`
                j_base = j; total = len;
                w.piece[4 * j + 2] = 0; w.piece[4 * j + 3] = 0;
            } else {
                const long long cf = dur / 4 < 20 ? dur / 4 : 20;
                long long ncf = w.piece[4 * j + 3];
                if (cf > ms_len(total, sr) && st == 0) st = 2;      // AudioSegment.append: ValueError
                if (cf == 0) ncf = 0;                               // append(crossfade=0): plain concatenation
                if (ncf > len) ncf = len;
                if (ncf > total) ncf = total;
                w.piece[4 * j + 3] = ncf;
                w.piece[4 * j + 2] = total - ncf;
                total += len - ncf;
                pos -= cf;
            }
            w.pend[j] = total;
            pos += dur;
        }
        SilHeader h; h.status = st; h.n_silent = K; h.n_chunks = nc; h.j_base = j_base; h.n_out = st == 1 ? 0 : total; h.n_ms = n_ms;
        h.pad_[0] = h.pad_[1] = h.pad_[2] = h.pad_[3] = 0;
        *w.hdr = h;
    }
}

// audioop.mul: floor(fbound(sample * factor)); audioop.add: floor(fbound(a + b)) (saturating)
__device__ __forceinline__ int aop_mul(int s, double f) { return (int)floor(fmin(fmax((double)s * f, -32768.0), 32767.0)); }

// out[o], o < n_out.  Chunk j occupies [off_j, E_j) of the output with E_j = off_j + len_j; its first ncf_j frames are the
// cross-fade with what the output held there before (seg1[-cf:].fade(to_gain=-120) * seg2[:cf].fade(from_gain=-120)).
// The base value of o comes from the chunk whose un-faded body [E_{j-1}, E_j) holds it; later chunks whose cross-fade region
// reaches back to o are applied in append order (more than one only when a chunk is shorter than the next cross-fade).
__global__ void __launch_bounds__(256) silence_compact_kernel(const int16_t* __restrict__ x, long long n, long long ncf_max, SilWs w,
                                                              int16_t* __restrict__ out) {
    const SilHeader h = *w.hdr;
    if (h.status != 0) return;
    const long long o0 = (long long)blockIdx.x * 1024;
    if (o0 >= h.n_out) return;
    __shared__ int j_cta;
    if (threadIdx.x == 0) {                                         // first chunk >= j_base with E_j > o0
        int lo = h.j_base, hi = h.n_chunks - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (w.pend[mid] > o0) hi = mid; else lo = mid + 1; }
        j_cta = lo;
    }
    __syncthreads();
    const double lo_gain = 1e-6;                                    // db_to_float(-120)
    for (int u = 0; u < 4; ++u) {
        const long long o = o0 + u * 256 + threadIdx.x;
        if (o >= h.n_out) break;
        int j = j_cta;
        while (j < h.n_chunks - 1 && w.pend[j] <= o) ++j;
        const long long off = w.piece[4 * j + 2];
        const long long src = w.piece[4 * j] + (o - off);
        int v = src < n ? (int)x[src] : 0;
        for (int q = j + 1; q < h.n_chunks; ++q) {
            if (w.pend[q - 1] > o + ncf_max) break;                 // cross-fade regions end at E_{q-1} and are at most ncf_max long
            const long long qoff = w.piece[4 * q + 2], qn = w.piece[4 * q + 3];
            if (o < qoff || o >= qoff + qn) continue;
            const long long i = o - qoff, qsrc = w.piece[4 * q] + i;
            const int head = qsrc < n ? (int)x[qsrc] : 0;
            const double g_out = 1.0 + ((lo_gain - 1.0) / (double)qn) * (double)i;
            const double g_in = lo_gain + ((1.0 - lo_gain) / (double)qn) * (double)i;
            const double sum = (double)aop_mul(v, g_out) + (double)aop_mul(head, g_in);
            v = (int)floor(fmin(fmax(sum, -32768.0), 32767.0));
        }
        out[o] = (int16_t)v;
    }
}

static SilWs carve_silence_ws(void* ws, long long n_ms, int win, size_t* total_bytes = nullptr, size_t* result_bytes = nullptr,
                              long long* off_sil = nullptr, long long* off_map = nullptr) {
    const long long n_starts = n_ms - win + 1 > 0 ? n_ms - win + 1 : 0;
    const long long cap = n_ms / (win > 0 ? win : 1) + 2;
    const uintptr_t base = reinterpret_cast<uintptr_t>(ws);
    uintptr_t p = base;
    auto take = [&](size_t bytes) { const uintptr_t q = p; p += (bytes + 255) / 256 * 256; return q; };
    SilWs w;
    w.cap = cap;
    w.hdr = (SilHeader*)take(sizeof(SilHeader));
    w.sil = (int*)take(8 * cap);
    w.map = (int*)take(16 * (cap + 1));
    if (off_sil) *off_sil = (long long)(reinterpret_cast<uintptr_t>(w.sil) - base);
    if (off_map) *off_map = (long long)(reinterpret_cast<uintptr_t>(w.map) - base);
    if (result_bytes) *result_bytes = (size_t)(p - base);           // header + ranges + mappings: what the host copies back
    w.piece = (long long*)take(32 * (cap + 1));
    w.pend = (long long*)take(8 * (cap + 1));
    const long long nb1 = (n_ms + 1 + SC_TILE - 1) / SC_TILE, nb2 = (n_starts + 1 + SC_TILE - 1) / SC_TILE;
    w.pb_loc = (unsigned long long*)take(8 * (n_ms + 1)); w.pb_tot = (unsigned long long*)take(8 * (nb1 + 1));
    w.cf_loc = (unsigned*)take(4 * (n_starts + 1)); w.cf_tot = (unsigned*)take(4 * (nb2 + 1)); w.cf_grand = (unsigned*)take(4);
    w.cr_loc = (unsigned*)take(4 * (n_starts + 1)); w.cr_tot = (unsigned*)take(4 * (nb2 + 1)); w.cr_grand = (unsigned*)take(4);
    w.flags = (uint8_t*)take(n_starts + 1);
    if (total_bytes) *total_bytes = (size_t)(p - base);
    return w;
}

void silence_layout(long long n_ms, int win, size_t* ws_bytes, size_t* result_bytes, long long* cap, long long* off_sil, long long* off_map) {
    const SilWs w = carve_silence_ws(nullptr, n_ms, win, ws_bytes, result_bytes, off_sil, off_map);
    if (cap) *cap = w.cap;
}

// silent ranges of the recording into the workspace (header.n_silent, sil table)
static void launch_silence_ranges(const unsigned long long* bucket, long long n_ms, int rate, int win, double thresh, const SilWs& w, cudaStream_t st) {
    const long long n_starts = n_ms - win + 1;
    const double r = (double)rate / 1000.0;
    const int nb1 = (int)((n_ms + 1 + SC_TILE - 1) / SC_TILE), nb2 = (int)((n_starts + 1 + SC_TILE - 1) / SC_TILE);
    bucket_scan_kernel<<<nb1, SC_T, 0, st>>>(bucket, n_ms, w);
    tot_scan_kernel<unsigned long long><<<1, 1024, 0, st>>>(w.pb_tot, nb1, nullptr);
    flag_scan_kernel<<<nb2, SC_T, 0, st>>>(n_starts, r, win, thresh, w);
    tot_scan_kernel<unsigned><<<1, 1024, 0, st>>>(w.cf_tot, nb2, w.cf_grand);
    first_scan_kernel<<<nb2, SC_T, 0, st>>>(n_starts, win, w);
    tot_scan_kernel<unsigned><<<1, 1024, 0, st>>>(w.cr_tot, nb2, w.cr_grand);
    ranges_scatter_kernel<<<(unsigned)((n_starts + 255) / 256), 256, 0, st>>>(n_starts, win, w);
}

int launch_silence(const int16_t* x, long long n, int rate, const unsigned long long* bucket, long long n_ms, int win, double thresh,
                   int margin, void* ws, int16_t* out, long long out_cap, cudaStream_t st) {
    const SilWs w = carve_silence_ws(ws, n_ms, win);
    const bool have = n_ms >= win;                                  // detect_silence: a clip shorter than the window has no silence
    if (have) launch_silence_ranges(bucket, n_ms, rate, win, thresh, w, st);
    const double r = (double)rate / 1000.0;
    silence_table_kernel<<<1, 1024, 0, st>>>(n, (double)rate, r, n_ms, margin, have ? 1 : 0, w);
    int launches = have ? 8 : 1;
    if (out && out_cap > 0) {
        const long long ncf_max = (long long)(20.0 * r) + 1;
        silence_compact_kernel<<<(unsigned)((out_cap + 1023) / 1024), 256, 0, st>>>(x, n, ncf_max, w, out);
        ++launches;
    }
    return launches;
}

}  // namespace maf
