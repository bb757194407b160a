// Longest-closed-run bookkeeping of the S0 noise-clip selection (reference: audio_pipeline/preprocessor.py:125-145), written so
// that frame ranges can be summarised independently and merged in any bracketing: the per-thread scan, the associative merge
// and the final rule.  Host-checkable (tests/host/host_runs_check.cpp runs the same code on the CPU against the literal loop).
#pragma once

#if defined(__CUDACC__)
#define MAF_RHD __host__ __device__ __forceinline__
#else
#define MAF_RHD inline
#endif

namespace maf {

// Summary of the noise flags of a frame range: all = every frame is noise; lead / trail = length of the run touching the
// range's first / last frame; best = longest run closed on BOTH sides inside the range (first one on ties).
struct RunSummary { int all, lead, trail, best_len, best_start; };

// A frame range with its summary; len == 0 is the identity of merge_seg.
struct RunSeg { RunSummary r; int start, len; };

// flags nz(i), i in [s0, s1) -> summary (s1 <= s0: the identity)
template <typename F>
MAF_RHD RunSummary summarize_range(int s0, int s1, F nz) {
    RunSummary rs; rs.all = 1; rs.lead = 0; rs.trail = 0; rs.best_len = 0; rs.best_start = 0;
    int run_start = -1; bool seen_false = false;
    for (int i = s0; i < s1; ++i) {
        if (nz(i)) {
            if (run_start < 0) run_start = i;
        } else {
            if (run_start >= 0) {
                if (!seen_false && run_start == s0) rs.lead = i - s0;
                else if (i - run_start > rs.best_len) { rs.best_len = i - run_start; rs.best_start = run_start; }
                run_start = -1;
            }
            seen_false = true; rs.all = 0;
        }
    }
    if (s1 <= s0) return rs;
    if (rs.all) { rs.lead = s1 - s0; rs.trail = s1 - s0; }
    else rs.trail = (run_start >= 0) ? s1 - run_start : 0;
    return rs;
}

// merge of the summaries of two adjacent frame ranges (L before R); lenL / lenR are their lengths, startR = first frame of R
MAF_RHD RunSummary merge_runs(const RunSummary& L, int lenL, const RunSummary& R, int lenR, int startR) {
    RunSummary o;
    if (lenL <= 0) return R;
    if (lenR <= 0) return L;
    o.all = L.all && R.all;
    o.lead = L.all ? lenL + R.lead : L.lead;
    o.trail = R.all ? lenR + L.trail : R.trail;
    o.best_len = L.best_len; o.best_start = L.best_start;
    if (!L.all && !R.all) {                                     // the run across the seam is closed on both sides
        const int mid = L.trail + R.lead;
        if (mid > o.best_len) { o.best_len = mid; o.best_start = startR - L.trail; }
    }
    if (R.best_len > o.best_len) { o.best_len = R.best_len; o.best_start = R.best_start; }
    if (o.all) { o.best_len = 0; o.best_start = 0; }
    return o;
}

MAF_RHD RunSeg merge_seg(const RunSeg& L, const RunSeg& R) {
    RunSeg o;
    o.r = merge_runs(L.r, L.len, R.r, R.len, R.start);
    o.start = L.len > 0 ? L.start : R.start;
    o.len = L.len + R.len;
    return o;
}

MAF_RHD RunSeg empty_seg() {
    RunSeg e; e.r.all = 1; e.r.lead = 0; e.r.trail = 0; e.r.best_len = 0; e.r.best_start = 0; e.start = 0; e.len = 0;
    return e;
}

// The reference's rule on the summary of the WHOLE recording: candidates in order of their start -- the leading run (closed
// unless everything is noise), then the best interior run; an open run at the end is dropped like in the reference's loop.
// Longest first-maximum run, in frames: (length, start); length 0 = none.
MAF_RHD void pick_longest_closed(const RunSummary& acc, int& bs, int& bl) {
    bl = 0; bs = 0;
    if (!acc.all && acc.lead > bl) { bl = acc.lead; bs = 0; }
    if (acc.best_len > bl) { bl = acc.best_len; bs = acc.best_start; }
}

}  // namespace maf
