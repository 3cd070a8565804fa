// Host-side construction of a GemmLaunch (ld_types.h) from per-job tap lists: sorts the taps, merges taps that can share
// one smem load into groups, sizes the smem ring, chooses the pipeline shape and encodes the tap program the MMA warps
// execute.  Shared by the inference context (ld_api.cu) and the training network (ld_train.cu).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "ld_net.h"
#include "ld_types.h"

namespace ld {

GemmTuning gemm_tuning_from_env() {
    auto env_int = [](const char* name, int dflt) { const char* v = std::getenv(name); return v ? std::atoi(v) : dflt; };
    GemmTuning t;
    t.loader = env_int("LD_GEMM_LOADER", 0);
    t.group_span = env_int("LD_GEMM_SPAN", 2);
    if (t.loader == 1) t.group_span = std::min(t.group_span, kBoxPixels - kTileM);
    t.max_stages = env_int("LD_GEMM_STAGES", 16);
    t.tile_stage_cin = env_int("LD_GEMM_TILE_STAGE_CIN", 32);
    t.align_loads = env_int("LD_GEMM_ALIGN", 0);
    t.n_rings_max = env_int("LD_GEMM_RINGS", 2);
    return t;
}

// The caller presets L's header (weights, shift, cin, cout, n_wtaps, relu, wp, out_mode, wp2, hp, mode, stats, prof).
bool gemm_build_launch(GemmLaunch& L, const std::vector<HostJob>& jobs, const GemmTuning& tune, std::string& err) {
    if (jobs.empty() || jobs.size() > static_cast<size_t>(kMaxJobs)) { err = "bad job count"; return false; }
    L.n_jobs = static_cast<int>(jobs.size());
    L.loader = tune.loader;
    struct TapInfo { int group[kMaxTaps], off[kMaxTaps], wslab[kMaxTaps]; };
    std::vector<TapInfo> info(jobs.size());
    int ext_max = 0;
    for (size_t j = 0; j < jobs.size(); ++j) {
        GemmJob& job = L.jobs[j];
        std::memset(&job, 0, sizeof(job));
        std::vector<HostTap> taps = jobs[j].taps;
        std::sort(taps.begin(), taps.end(), [](const HostTap& a, const HostTap& b) {
            return a.src != b.src ? a.src < b.src : a.shift < b.shift;
        });
        if (taps.empty() || taps.size() > static_cast<size_t>(kMaxTaps)) { err = "bad tap count"; return false; }
        int g = -1, g_min = 0;
        const void* g_src = nullptr;
        int group_ext[kMaxGroups] = {0};
        for (size_t t = 0; t < taps.size(); ++t) {
            if (g < 0 || taps[t].src != g_src || taps[t].shift - g_min > tune.group_span + (tune.align_loads ? 7 : 0)) {
                if (++g >= kMaxGroups) { err = "too many load groups"; return false; }
                g_src = taps[t].src; g_min = taps[t].shift;
                // optionally start every copy on a 128-byte boundary of the plane (8 pixels)
                if (tune.align_loads) g_min -= ((g_min % 8) + 8) % 8;
                job.groups[g].src = static_cast<const __half*>(taps[t].src);
                job.groups[g].kc_stride = taps[t].kc_stride;
                job.groups[g].tmap = taps[t].tmap;
                job.groups[g].pixel0 = taps[t].pixel0;
                job.groups[g].shift = g_min;
                group_ext[g] = kTileM;
            }
            group_ext[g] = std::max(group_ext[g], kTileM + taps[t].shift - g_min);
            info[j].group[t] = g;
            info[j].off[t] = taps[t].shift - g_min;
            info[j].wslab[t] = taps[t].wslab;
        }
        job.n_groups = g + 1;
        job.n_taps = static_cast<int>(taps.size());
        for (int q = 0; q < job.n_groups; ++q) ext_max = std::max(ext_max, group_ext[q]);
        job.out0 = static_cast<__half*>(jobs[j].out0);
        job.out1 = static_cast<__half*>(jobs[j].out1);
        job.out_kc_stride = jobs[j].out_kc_stride;
    }
    if (tune.loader == 1) {
        if (ext_max > kBoxPixels) { err = "tap span exceeds the TMA box"; return false; }
        L.ext_alloc = kBoxPixels;
    } else {
        L.ext_alloc = (ext_max + 7) & ~7;
    }
    // small-K layers: one smem stage (one barrier round trip) per TILE instead of per group
    int max_groups = 1;
    for (int j = 0; j < L.n_jobs; ++j) max_groups = std::max(max_groups, L.jobs[j].n_groups);
    L.groups_per_stage = (tune.tile_stage_cin > 0 && L.cin <= tune.tile_stage_cin) ? max_groups : 1;
    L.wp_magic = static_cast<uint32_t>((1ull << 32) / static_cast<unsigned>(L.wp)) + 1u;
    L.n_stages = gemm_pick_stages(L.cin, L.cout, L.n_wtaps, L.ext_alloc, L.groups_per_stage, tune.max_stages);
    {   // two rings when half of the stages still hold two whole tiles (tile-stage: one stage each; per-group: max_groups);
        // per-group launches whose half ring would hold less measured slower with two rings
        const int need = L.groups_per_stage > 1 ? 1 : max_groups;
        if (L.n_stages < need || L.n_stages < 2) { err = "smem ring shorter than one tile"; return false; }
        L.n_rings = (tune.n_rings_max >= 2 && L.n_stages / 2 >= 2 * need) ? 2 : 1;
        if (L.n_rings == 2) L.n_stages &= ~1;
    }
    // the tap program the MMA warp executes (ld_types.h: kTapFirst / kTapLast / kTapPass)
    const uint32_t kchunks = static_cast<uint32_t>(L.cin / 8);
    for (int j = 0; j < L.n_jobs; ++j) {
        GemmJob& job = L.jobs[j];
        const TapInfo& ti = info[j];
        const uint32_t box16 = static_cast<uint32_t>(L.ext_alloc) * kchunks;
        const bool tile_stage = L.groups_per_stage > 1;
        int last_first = 0;
        for (int t = 0; t < job.n_taps; ++t) {
            const uint32_t a16 = (tile_stage ? ti.group[t] * box16 : 0u) + static_cast<uint32_t>(ti.off[t]);
            const uint32_t b16 = static_cast<uint32_t>(ti.wslab[t]) * kchunks * L.cout;
            const bool first = tile_stage ? t == 0 : (t == 0 || ti.group[t] != ti.group[t - 1]);
            const bool last = tile_stage ? t == job.n_taps - 1 : (t == job.n_taps - 1 || ti.group[t] != ti.group[t + 1]);
            if (a16 >= (1u << 14) || b16 >= (1u << 14)) { err = "tap offset overflow"; return false; }
            job.tapw[t] = a16 | (b16 << 14) | (first ? kTapFirst : 0u) | (last ? kTapLast : 0u);
            if (first) last_first = t;
        }
        job.tapw[last_first] |= kTapPass;
    }
    return true;
}

}  // namespace ld
