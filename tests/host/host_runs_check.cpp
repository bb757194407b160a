// CPU check of csrc/run_merge.h: per-segment summaries merged in the kernels' tree order (32-wide shuffle-down trees, then a
// tree over the warps, then -- for the grid kernel -- over the CTAs) must give the run the literal loop of the reference
// (audio_pipeline/preprocessor.py:125-145) gives: the longest run of noise frames that is closed by a non-noise frame, the
// first one on ties; an open run at the end does not count.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../modular_audio_pipeline_b200/csrc/run_merge.h"
using namespace maf;

static void literal(const std::vector<unsigned char>& f, int& bs, int& bl) {
    bs = 0; bl = 0;
    bool in = false; int start = 0;
    for (int i = 0; i < (int)f.size(); ++i) {
        if (f[i] && !in) { start = i; in = true; }
        else if (!f[i] && in) { if (i - start > bl) { bl = i - start; bs = start; } in = false; }
    }
}

// shuffle-down tree over `width` lanes, result in lane 0 (what block_merge_segs does per warp)
static RunSeg tree(std::vector<RunSeg> v) {
    const int width = (int)v.size();
    for (int d = 1; d < width; d <<= 1) {
        std::vector<RunSeg> nv = v;
        for (int l = 0; l < width; ++l) if (l + d < width) nv[l] = merge_seg(v[l], v[l + d]);
        v = nv;
    }
    return v[0];
}

static RunSeg block_merge(const std::vector<RunSeg>& threads) {       // 1024 threads = 32 warps of 32 lanes
    std::vector<RunSeg> warps;
    for (size_t w = 0; w < threads.size(); w += 32) {
        std::vector<RunSeg> lanes(threads.begin() + w, threads.begin() + std::min(threads.size(), w + 32));
        lanes.resize(32, empty_seg());
        warps.push_back(tree(lanes));
    }
    warps.resize(32, empty_seg());
    return tree(warps);
}

int main() {
    int fails = 0;
    srand(12345);
    for (int trial = 0; trial < 1500; ++trial) {
        const int F = trial < 50 ? trial : 1 + rand() % (trial % 7 == 0 ? 70000 : 3000);
        const int density = rand() % 101;                      // 0 = no noise frame, 100 = all noise
        std::vector<unsigned char> f(F);
        int burst = 0; bool val = false;
        for (int i = 0; i < F; ++i) {
            if (burst == 0) { val = (rand() % 100) < density; burst = 1 + rand() % (1 + rand() % 40); }
            f[i] = val; --burst;
        }
        int ws, wl; literal(f, ws, wl);
        // (1) one CTA: 1024 threads, contiguous segments of `per` frames
        {
            const int T = 1024, per = (F + T - 1) / T;
            std::vector<RunSeg> th(T);
            for (int t = 0; t < T; ++t) {
                const int s0 = std::min(F, t * per), s1 = std::min(F, s0 + per);
                th[t].start = s0; th[t].len = s1 - s0;
                th[t].r = summarize_range(s0, s1, [&](int i) { return f[i] != 0; });
            }
            int bs, bl; pick_longest_closed(block_merge(th).r, bs, bl);
            if (bl != wl || (bl > 0 && bs != ws)) { if (fails < 5) printf("CTA  F=%d want (%d,%d) got (%d,%d)\n", F, ws, wl, bs, bl); ++fails; }
        }
        // (2) the grid: G CTAs (148 on a B200; others to vary the split), each 1024 threads, then a tree over the CTA results
        {
            const int G = trial % 3 == 0 ? 148 : 1 + rand() % 200;
            const int per_cta = (F + G - 1) / G, per = (per_cta + 1023) / 1024;
            std::vector<RunSeg> ctas(1024, empty_seg());
            for (int c = 0; c < G; ++c) {
                const int c0 = std::min(F, c * per_cta), c1 = std::min(F, c0 + per_cta);
                std::vector<RunSeg> th(1024);
                for (int t = 0; t < 1024; ++t) {
                    const int s0 = std::min(c1, c0 + t * per), s1 = std::min(c1, s0 + per);
                    th[t].start = s0; th[t].len = s1 - s0;
                    th[t].r = summarize_range(s0, s1, [&](int i) { return f[i] != 0; });
                }
                ctas[c].r = block_merge(th).r; ctas[c].start = c0; ctas[c].len = c1 - c0;
            }
            int bs, bl; pick_longest_closed(block_merge(ctas).r, bs, bl);
            if (bl != wl || (bl > 0 && bs != ws)) { if (fails < 5) printf("GRID F=%d G=%d want (%d,%d) got (%d,%d)\n", F, G, ws, wl, bs, bl); ++fails; }
        }
    }
    printf(fails ? "FAIL (%d)\n" : "OK\n", fails);
    return fails != 0;
}
