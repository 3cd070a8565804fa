// Host-side construction of a GemmLaunch (ld_types.h) from per-output tap lists: finds the chains of output planes that
// share input planes, packs them into jobs, merges the taps of one input operand that feed adjacent outputs into single
// wide-N MMAs, assigns the loads to smem stages, sizes the ring and encodes the tap program the MMA warps execute.
// Shared by the inference context (ld_api.cu) and the training network (ld_train.cu).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <tuple>

#include "ld_net.h"
#include "ld_types.h"

namespace ld {

GemmTuning gemm_tuning_from_env() {
    auto env_int = [](const char* name, int dflt) { const char* v = std::getenv(name); return v ? std::atoi(v) : dflt; };
    GemmTuning t;
    t.group_span = env_int("LD_GEMM_SPAN", 2);
    t.max_stages = env_int("LD_GEMM_STAGES", 16);
    t.stage_bytes = env_int("LD_GEMM_STAGE_BYTES", 18 * 1024);
    t.max_outs = env_int("LD_GEMM_MAX_OUTS", kMaxOuts);
    t.issuers_wide = env_int("LD_GEMM_ISSUERS_WIDE", 2);
    t.issuers_narrow = env_int("LD_GEMM_ISSUERS_NARROW", 4);
    t.n_rings_max = env_int("LD_GEMM_RINGS", 2);
    t.dual_narrow = env_int("LD_GEMM_DUAL", 1);
    t.issuers_dual = env_int("LD_GEMM_ISSUERS_DUAL", 2);
    return t;
}

namespace {

struct Tap {          // one (input operand, weight slab) product feeding output `out` of a job
    const void* src;
    long long kc_stride;
    int shift, wslab, out;
    int group, off;   // load group inside the job, pixel offset inside the group
    int half_k;
};

bool shares_input(const HostJob& a, const HostJob& b) {
    for (const auto& x : a.taps)
        for (const auto& y : b.taps)
            if (x.src == y.src && x.shift == y.shift) return true;
    return false;
}

}  // namespace

static bool build_with(GemmLaunch& L, const std::vector<HostJob>& outs, const GemmTuning& tune, int max_outs, bool pack_all, std::string& err);
static bool gemm_build_launch_impl(GemmLaunch& L, const std::vector<HostJob>& outs, const GemmTuning& tune, std::string& err);

// The caller presets L's header (weights, shift, cin, cout, n_wtaps, w_stack, relu, wp, out_mode, wp2, hp, mode, stats, prof).
bool gemm_build_launch(GemmLaunch& L, const std::vector<HostJob>& outs, const GemmTuning& tune, std::string& err) {
    if (outs.empty()) { err = "no outputs"; return false; }
    {   // knock-out timing experiments produce garbage results: honoured only together with the profiling switch
        const char* v = std::getenv("LD_GEMM_DBG");
        const char* p = std::getenv("LD_GEMM_PROF");
        L.dbg = (v && p && std::atoi(p)) ? std::atoi(v) : 0;
    }
    {
        const char* v = std::getenv("LD_GEMM_L2PF");
        L.l2_prefetch = v ? std::atoi(v) : 0;
    }
    if (L.w_stack && L.w_blocks < 1) L.w_blocks = 1;
    if (!L.w_stack) L.w_blocks = 0;
    if (L.w_stack && L.n_wtaps < 9 * L.w_blocks) { err = "stacked weights need a 3x3 kernel"; return false; }
    // cout >= 48: two issuers with 256 accumulator columns each (chains of up to 4 outputs); narrower layers are bound by
    // the issue latency of their small MMAs: four issuers with 128 columns each
    // narrow inference layers: two CTAs per SM with 256 accumulator columns and half of the shared memory each; if the weights
    // leave no room for a useful ring in 113 KB (split precision, 64 -> 32), fall back to one CTA per SM
    // (measured, profiles/r02/gemm_dual_cta_experiment.log: -3.4 ms per step on the cout = 16 layers, +2.5 ms on cout = 32 -> 16 only)
    if (L.mode == 0 && L.cout <= (tune.dual_narrow >= 2 ? 32 : 16) && tune.dual_narrow) {
        GemmLaunch trial = L;
        GemmTuning t2 = tune;
        t2.dual_narrow = 0;
        trial.tmem_cols = 256;
        trial.n_issuers = tune.issuers_dual;
        std::string e2;
        if ((trial.n_issuers == 2 || trial.n_issuers == 4) && gemm_build_launch_impl(trial, outs, t2, e2) && trial.n_stages >= 4) {
            L = trial;
            return true;
        }
    }
    L.tmem_cols = kTmemCols;
    L.n_issuers = L.cout >= 48 ? tune.issuers_wide : tune.issuers_narrow;
    return gemm_build_launch_impl(L, outs, tune, err);
}

bool gemm_build_launch_impl(GemmLaunch& L, const std::vector<HostJob>& outs, const GemmTuning& tune, std::string& err) {
    if (L.n_issuers != 2 && L.n_issuers != 4) { err = "2 or 4 MMA issuers"; return false; }
    // the longest chains whose loads and taps fit a job (stride-2 layers merge nothing: their chains stay short)
    // 1x1 convs (one or two taps per output) share nothing, but several outputs per job still amortise the per-tile costs
    bool pack_all = true;
    for (const auto& o : outs) pack_all = pack_all && o.taps.size() <= 2;
    const int max_outs = (L.w_stack || pack_all) ? std::max(1, std::min({tune.max_outs, kMaxOuts, L.tmem_cols / L.n_issuers / L.cout})) : 1;
    auto build = [&]() {
        for (int mo = max_outs; mo >= 1; --mo)
            if (build_with(L, outs, tune, mo, pack_all, err)) return true;
        return false;
    };
    if (!L.w_stack) return build();
    // two stacking orders of the weight rows: ky = 2,1,0 merges the three outputs a stride-1 input row feeds, ky = 2,0,1 the
    // two outputs (ky = 2 of row h, ky = 0 of row h + 1) an odd input row of a stride-2 conv feeds; keep the shorter program
    int n_taps[3] = {0, 1 << 30, 1 << 30};
    for (int order = 1; order <= 2; ++order) {
        L.w_stack = order;
        if (!build()) continue;
        n_taps[order] = 0;
        for (int j = 0; j < L.n_jobs; ++j) n_taps[order] += L.job_taps[j].n_taps;
    }
    L.w_stack = n_taps[2] < n_taps[1] ? 2 : 1;
    return build();
}

static bool build_with(GemmLaunch& L, const std::vector<HostJob>& outs, const GemmTuning& tune, const int max_outs, const bool pack_all,
                       std::string& err) {
    const int kchunks = L.cin / 8;
    const bool stack = L.w_stack != 0;

    // ---- 1. jobs: maximal runs of consecutive outputs that share an input operand, cut into near-equal parts that fit TMEM
    std::vector<std::pair<int, int>> parts;  // [first, last) output indices
    for (size_t a = 0; a < outs.size();) {
        size_t b = a + 1;
        while (max_outs > 1 && b < outs.size() && (pack_all || shares_input(outs[b - 1], outs[b]))) ++b;
        const int n = static_cast<int>(b - a), n_parts = (n + max_outs - 1) / max_outs;
        for (int i = 0; i < n_parts; ++i)
            parts.push_back({static_cast<int>(a) + i * n / n_parts, static_cast<int>(a) + (i + 1) * n / n_parts});
        a = b;
    }
    if (parts.size() > static_cast<size_t>(kMaxJobs)) { err = "too many jobs"; return false; }
    L.n_jobs = static_cast<int>(parts.size());

    // ---- 2. per job: load groups (taps of one plane within group_span pixels share a load), in chain order
    std::vector<std::vector<Tap>> job_taps(parts.size());
    int ext_max = 0, max_groups = 1;
    for (size_t j = 0; j < parts.size(); ++j) {
        GemmJob& job = L.jobs[j];
        std::memset(&job, 0, sizeof(job));
        job.dep_back[0] = job.dep_back[1] = job.dep_back[2] = kNoDep;
        std::vector<Tap>& taps = job_taps[j];
        for (int o = parts[j].first; o < parts[j].second; ++o) {
            if (outs[o].taps.empty()) { err = "output without taps"; return false; }
            for (const auto& t : outs[o].taps) taps.push_back({t.src, t.kc_stride, t.shift, t.wslab, o - parts[j].first, -1, 0, t.half_k});
            job.outs[o - parts[j].first] = {static_cast<__half*>(outs[o].out0), static_cast<__half*>(outs[o].out1)};
            if (outs[o].out_kc_stride != outs[parts[j].first].out_kc_stride) { err = "outputs of a job differ in layout"; return false; }
        }
        job.n_outs = parts[j].second - parts[j].first;
        job.out_kc_stride = outs[parts[j].first].out_kc_stride;
        std::sort(taps.begin(), taps.end(), [](const Tap& a, const Tap& b) {
            return std::tie(a.src, a.shift, a.out) < std::tie(b.src, b.shift, b.out);
        });
        struct Group { const void* src; long long kc_stride; int g_min, ext, order; };
        std::vector<Group> groups;
        for (auto& t : taps) {
            if (groups.empty() || groups.back().src != t.src || t.shift - groups.back().g_min > tune.group_span)
                groups.push_back({t.src, t.kc_stride, t.shift, kTileM, 1 << 30});
            Group& g = groups.back();
            g.ext = std::max(g.ext, kTileM + t.shift - g.g_min);
            // chain position of the operand: input row = output row + ky - 1 (identity / 1x1 taps sit at their output)
            const int n_stacked = 9 * L.w_blocks;
            const int ky = (stack && t.wslab < n_stacked) ? (t.wslab % 9) / 3 : 1;
            g.order = std::min(g.order, 4 * (t.out + ky - 1) + ((stack && t.wslab >= n_stacked) ? 1 : 0));
            t.group = static_cast<int>(groups.size()) - 1;
            t.off = t.shift - g.g_min;
        }
        if (groups.size() > static_cast<size_t>(kMaxGroups)) { err = "too many load groups"; return false; }
        std::vector<int> perm(groups.size()), rank(groups.size());
        for (size_t i = 0; i < perm.size(); ++i) perm[i] = static_cast<int>(i);
        std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return groups[a].order < groups[b].order; });
        for (size_t i = 0; i < perm.size(); ++i) {
            rank[perm[i]] = static_cast<int>(i);
            const Group& g = groups[perm[i]];
            job.groups[i] = {static_cast<const __half*>(g.src), g.kc_stride, g.g_min, 0};
            ext_max = std::max(ext_max, g.ext);
        }
        for (auto& t : taps) t.group = rank[t.group];
        job.n_groups = static_cast<int>(groups.size());
        max_groups = std::max(max_groups, job.n_groups);
    }
    L.ext_alloc = (ext_max + 7) & ~7;
    L.ext_copy = ext_max;
    L.wp_magic = static_cast<uint32_t>((1ull << 32) / static_cast<unsigned>(L.wp)) + 1u;

    // ---- 3. smem ring: a stage holds consecutive groups of a job up to ~stage_bytes (small-K layers: fewer barrier trips)
    const int box_bytes = L.ext_alloc * 16 * kchunks;
    L.groups_per_stage = std::max(1, std::min(max_groups, tune.stage_bytes / box_bytes));
    const unsigned cap = gemm_smem_cap(L.tmem_cols);
    L.n_stages = gemm_pick_stages(L.cin, L.cout, L.n_wtaps, L.n_jobs, L.ext_alloc, L.groups_per_stage, tune.max_stages, cap);
    while (L.n_stages < 4 && L.groups_per_stage > 1) {
        --L.groups_per_stage;
        L.n_stages = gemm_pick_stages(L.cin, L.cout, L.n_wtaps, L.n_jobs, L.ext_alloc, L.groups_per_stage, tune.max_stages, cap);
    }
    if (L.n_stages < 2) { err = "smem ring shorter than two stages"; return false; }
    L.n_rings = (tune.n_rings_max >= 2 && L.n_stages >= 6) ? 2 : 1;
    if (L.n_rings == 2) L.n_stages &= ~1;

    // ---- 4. the tap program: per group, per pixel offset, runs of adjacent outputs with adjacent weight rows -> one MMA
    const int gps = L.groups_per_stage;
    const uint32_t box16 = static_cast<uint32_t>(L.ext_alloc) * kchunks;
    int n_launch_taps = 0;
    for (size_t j = 0; j < parts.size(); ++j) {
        GemmJob& job = L.jobs[j];
        std::vector<Tap>& taps = job_taps[j];
        uint4* tapw = L.taps + n_launch_taps;
        // weight row block of a tap inside its stacked block (ky = 2, 1, 0 or ky = 2, 0, 1), or -1 for a stand-alone slab
        const int n_stacked = 9 * L.w_blocks;
        auto wrow = [&](const Tap& t) { return (stack && t.wslab < n_stacked) ? w_stack_row(L.w_stack, (t.wslab % 9) / 3) : -1; };
        // merge key: column kx of weight block b (taps of different blocks never merge); stand-alone slabs keep their index
        auto kx_of = [&](const Tap& t) { return (stack && t.wslab < n_stacked) ? (t.wslab / 9) * 3 + t.wslab % 3 : 64 + t.wslab; };
        std::sort(taps.begin(), taps.end(), [&](const Tap& a, const Tap& b) {
            return std::make_tuple(a.group, a.off, kx_of(a), a.out) < std::make_tuple(b.group, b.off, kx_of(b), b.out);
        });
        job.n_stages = (job.n_groups + gps - 1) / gps;
        int n = 0, last_first = -1;
        for (size_t i = 0; i < taps.size();) {
            size_t e = i + 1;
            if (wrow(taps[i]) >= 0)
                while (e < taps.size() && taps[e].group == taps[i].group && taps[e].off == taps[i].off &&
                       kx_of(taps[e]) == kx_of(taps[i]) && wrow(taps[e]) >= 0 && taps[e].half_k == taps[i].half_k &&
                       taps[e].out == taps[e - 1].out + 1 &&
                       wrow(taps[e]) == wrow(taps[e - 1]) + 1)
                    ++e;
            if (n >= kMaxTaps || n_launch_taps + n >= kMaxLaunchTaps) { err = "too many MMA taps in a job"; return false; }
            const Tap& t = taps[i];
            const int n_merged = static_cast<int>(e - i);
            const uint32_t a16 = static_cast<uint32_t>(t.group % gps) * box16 + static_cast<uint32_t>(t.off);
            uint32_t b16, lbo16;
            if (wrow(t) >= 0) {
                const uint32_t blk = static_cast<uint32_t>(t.wslab / 9), kx = static_cast<uint32_t>(t.wslab % 3);
                b16 = blk * 9u * kchunks * L.cout + kx * kchunks * 3u * L.cout + static_cast<uint32_t>(wrow(t)) * L.cout;
                lbo16 = 3u * L.cout;
            } else {
                b16 = static_cast<uint32_t>(t.wslab) * kchunks * L.cout;
                lbo16 = static_cast<uint32_t>(L.cout);
            }
            if (a16 >= (1u << 14) || b16 >= (1u << 13)) { err = "tap offset overflow"; return false; }
            const int stage_of = t.group / gps;
            const bool first = i == 0 || taps[i - 1].group / gps != stage_of;
            const bool last = e == taps.size() || taps[e].group / gps != stage_of;
            const uint32_t col = static_cast<uint32_t>(t.out) * L.cout, ncols = static_cast<uint32_t>(n_merged) * L.cout;
            if (col + ncols > static_cast<uint32_t>(L.tmem_cols / L.n_issuers)) { err = "accumulator overflow"; return false; }
            tapw[n].x = a16 | (first ? kTapFirst : 0u) | (last ? kTapLast : 0u) | (t.half_k ? kTapHalfK : 0u);
            tapw[n].y = b16 | (lbo16 << 16);
            tapw[n].z = col;
            tapw[n].w = (ncols >> 3) << 17;
            if (first) last_first = n;
            ++n;
            i = e;
        }
        tapw[last_first].x |= kTapPass;
        job.n_taps = n;
        L.job_taps[j] = {static_cast<uint16_t>(n_launch_taps), static_cast<uint16_t>(n), static_cast<uint16_t>(job.n_stages), 0};
        n_launch_taps += n;
    }
    return true;
}

}  // namespace ld
