// Host-side planner for streaming ResNetBigger inference (models.py:181-244, datasets.py:72-93 of the
// reference): the network is evaluated on EVERY frame's 100-row window, and consecutive windows share
// 99 rows.  A conv output row whose receptive field does not touch the window's zero padding has the
// same value in every window that contains it, so it is computed once per sequence row ("interior"
// plane, a dilated fully-convolutional pass).  Only the rows near the window's top/bottom edge are
// window-specific; each of those gets one plane indexed by the window start row b.  Every conv then
// becomes a list of shifted-plane GEMM taps (ld_gemm.cu), and results are bit-identical to evaluating
// each window densely with the same arithmetic.
#include <algorithm>
#include <cstdio>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>

#include "ld_types.h"

namespace ld {
namespace {

struct Level {
    int H = 0, W = 0, C = 0, res = 1;
    bool colsplit = false;   // stored as even/odd column planes in the consumer's geometry
    bool split = false;      // split precision: [hi | lo] planes
    int wp = 0;              // padded width of the stored planes
    std::set<int> spec;      // window-specific local rows
    std::set<int> needed;    // local rows some consumer reads
    std::map<int, int> spec_plane;  // local row -> plane id (colsplit: even plane, odd = id+1)
    int interior = -1;       // plane id (colsplit: even plane, odd = id+1)
    std::string tag;
};

struct ConvOp {
    std::string conv, bn;
    int in = -1, out = -1, res = -1;  // level indices
    int stride = 1, ksize = 3, relu = 1;
};

int out_size(int n, int stride) { return (n - 1) / stride + 1; }  // k=3,p=1 and k=1,p=0 alike

}  // namespace

Plan build_stream_plan(const NetConfig& cfg) {
    Plan plan;
    // pad columns per stored row: 1 = one zero column shared by consecutive rows (default), 2 = one on either side (the round-1
    // layout, kept for A/B measurements: LD_PLAN_PAD_COLS=2)
    const char* pad_env = std::getenv("LD_PLAN_PAD_COLS");
    const int pad_cols = (pad_env && std::atoi(pad_env) == 2) ? 2 : 1;
    plan.H = cfg.H;
    plan.W = cfg.W;
    std::vector<Level> lv;
    std::vector<ConvOp> ops;

    auto new_level = [&](int H, int W, int C, int res, const std::string& tag) {
        Level l;
        l.H = H; l.W = W; l.C = C; l.res = res; l.tag = tag; l.wp = W + pad_cols;   // one shared pad column per row (ld_types.h)
        lv.push_back(l);
        return static_cast<int>(lv.size()) - 1;
    };

    // ---- topology -------------------------------------------------------------------------------
    const int stem_out = new_level(cfg.H, cfg.W, 64, 1, "stem");
    lv[stem_out].spec = {0, cfg.H - 1};  // 3x3 stem conv on the (global) features: only edge rows differ
    if (cfg.H < 2) throw std::runtime_error("window too short");

    int cur = stem_out;
    int in_c = 64;
    for (int b = 0; b < 4; ++b) {
        const int out_c = cfg.filters[b];
        const int stride = (b == 0) ? 1 : 2;
        for (int r = 0; r < 2; ++r) {
            const int s = (r == 0) ? stride : 1;
            const int ic = (r == 0) ? in_c : out_c;
            const std::string pre = "block" + std::to_string(b + 1) + "." + std::to_string(r);
            const Level x = lv[cur];
            const int Ho = out_size(x.H, s), Wo = out_size(x.W, s);
            if (s == 2) {
                lv[cur].colsplit = true;
                lv[cur].wp = Wo + pad_cols;
            }
            const bool split = cfg.precision == 1 && b >= 1;   // blocks 2-4 (tools/precision_budget.py: block1 does not need it)
            const int h = new_level(Ho, Wo, out_c, x.res * s, pre + ".h");
            lv[h].split = split;
            ops.push_back({pre + ".conv1", pre + ".bn1", cur, h, -1, s, 3, 1});
            int res_level = cur;
            if (s != 1 || ic != out_c) {
                const int sc = new_level(Ho, Wo, out_c, x.res * s, pre + ".sc");
                lv[sc].split = split;
                ops.push_back({pre + ".shortcut.0", pre + ".shortcut.1", cur, sc, -1, s, 1, 0});
                res_level = sc;
            }
            const int y = new_level(Ho, Wo, out_c, x.res * s, pre + ".y");
            lv[y].split = split;
            ops.push_back({pre + ".conv2", pre + ".bn2", h, y, res_level, 1, 3, 1});
            cur = y;
        }
        in_c = out_c;
    }
    const int final_level = cur;

    // ---- forward: which local rows are window-specific ---------------------------------------------
    for (auto& op : ops) {
        const Level& in = lv[op.in];
        Level& out = lv[op.out];
        const int r = (op.ksize == 3) ? 1 : 0;
        for (int j = 0; j < out.H; ++j) {
            bool s = false;
            for (int dy = -r; dy <= r; ++dy) {
                const int i = op.stride * j + dy;
                if (i < 0 || i >= in.H || in.spec.count(i)) s = true;
            }
            if (op.res >= 0 && lv[op.res].spec.count(j)) s = true;
            if (s) out.spec.insert(j);
        }
    }
    // a shortcut level shares the row classification of the block output it feeds
    for (auto& op : ops)
        if (op.res >= 0 && lv[op.res].tag.size() > 3 &&
            lv[op.res].tag.compare(lv[op.res].tag.size() - 3, 3, ".sc") == 0)
            lv[op.res].spec = lv[op.out].spec;
    // (the conv2 that consumes a shortcut comes after it, so re-run once to settle)
    for (auto& op : ops) {
        const Level& in = lv[op.in];
        Level& out = lv[op.out];
        const int r = (op.ksize == 3) ? 1 : 0;
        for (int j = 0; j < out.H; ++j) {
            bool s = out.spec.count(j) > 0;
            for (int dy = -r; dy <= r; ++dy) {
                const int i = op.stride * j + dy;
                if (i < 0 || i >= in.H || in.spec.count(i)) s = true;
            }
            if (op.res >= 0 && lv[op.res].spec.count(j)) s = true;
            if (s) out.spec.insert(j);
        }
    }

    // ---- backward: which rows are consumed -------------------------------------------------------
    {
        Level& f = lv[final_level];
        const int groups = f.H / 4;  // AvgPool2d(4) drops the remainder rows/cols (models.py:229)
        plan.head_pool_groups = groups;
        plan.head_pool_w = f.W / 4;
        if (plan.head_pool_w != 1 || f.C * groups * plan.head_pool_w != cfg.linear_in)
            throw std::runtime_error("head geometry does not match linear_layer_size");
        for (int i = 0; i < groups * 4; ++i) f.needed.insert(i);
    }
    for (auto it = ops.rbegin(); it != ops.rend(); ++it) {
        const Level& out = lv[it->out];
        Level& in = lv[it->in];
        const int r = (it->ksize == 3) ? 1 : 0;
        for (int j : out.needed) {
            for (int dy = -r; dy <= r; ++dy) {
                const int i = it->stride * j + dy;
                if (i >= 0 && i < in.H) in.needed.insert(i);
            }
            if (it->res >= 0) lv[it->res].needed.insert(j);
        }
    }

    // ---- planes ------------------------------------------------------------------------------------
    bool cur_split = false;
    auto new_plane = [&](int C, int wp, const std::string& tag) {
        PlaneSpec p;
        p.id = static_cast<int>(plan.planes.size());
        p.C = C; p.wp = wp; p.tag = tag; p.split = cur_split;
        plan.planes.push_back(p);
        return p.id;
    };
    for (auto& l : lv) {
        bool any_interior = false;
        cur_split = l.split;
        for (int i : l.needed) {
            if (l.spec.count(i)) {
                l.spec_plane[i] = new_plane(l.C, l.wp, l.tag + ".r" + std::to_string(i) + (l.colsplit ? ".e" : ""));
                if (l.colsplit) new_plane(l.C, l.wp, l.tag + ".r" + std::to_string(i) + ".o");
            } else {
                any_interior = true;
            }
        }
        if (any_interior) {
            l.interior = new_plane(l.C, l.wp, l.tag + ".int" + (l.colsplit ? ".e" : ""));
            if (l.colsplit) new_plane(l.C, l.wp, l.tag + ".int.o");
        }
    }
    // resolve local row i of the window starting at plane row b -> (plane, row shift)
    auto resolve = [&](const Level& l, int i, int& plane, int& row_shift) {
        auto it = l.spec_plane.find(i);
        if (it != l.spec_plane.end()) { plane = it->second; row_shift = 0; return; }
        if (l.spec.count(i) || l.interior < 0)
            throw std::runtime_error("planner: row " + std::to_string(i) + " of " + l.tag + " was not materialised");
        plane = l.interior; row_shift = i * l.res;
    };

    // ---- stem ----------------------------------------------------------------------------------------
    {
        const Level& l = lv[stem_out];
        plan.stem_wp = l.wp;
        auto mask_for = [&](int j) {
            int m = 0;
            for (int ky = 0; ky < 3; ++ky) {
                const int i = j + ky - 1;
                if (i >= 0 && i < cfg.H) m |= 1 << ky;
            }
            return m;
        };
        for (auto& kv : l.spec_plane) plan.stem.push_back({kv.second, kv.first, mask_for(kv.first)});
        if (l.interior >= 0) plan.stem.push_back({l.interior, 0, 7});
        plan.macs_per_row += 9.0 * 64 * cfg.W * plan.stem.size();
    }

    // ---- convs -----------------------------------------------------------------------------------------
    for (auto& op : ops) {
        const Level& in = lv[op.in];
        const Level& out = lv[op.out];
        ConvLaunchSpec L;
        L.conv = op.conv; L.bn = op.bn;
        L.cin = in.C; L.cout = out.C; L.ksize = op.ksize; L.relu = op.relu;
        L.wp = out.W + pad_cols;  // geometry of the GEMM's pixel index
        L.w_real = out.W;
        L.out_mode = out.colsplit ? OUT_COLSPLIT : OUT_PLAIN;
        L.wp2 = out.wp;
        L.hp = 0;
        L.split_in = in.split; L.split_out = out.split; L.split_w = out.split;
        if (op.res >= 0 && lv[op.res].split != in.split) throw std::runtime_error("planner: residual/input precision mismatch at " + op.conv);
        if ((op.stride == 1 && (in.colsplit || in.wp != L.wp)) ||
            (op.stride == 2 && (!in.colsplit || in.wp != L.wp)))
            throw std::runtime_error("planner: storage geometry mismatch at " + op.conv);
        const int r = (op.ksize == 3) ? 1 : 0;

        auto add_taps = [&](JobSpec& job, int plane, int row_shift, int dy) {
            for (int dx = -r; dx <= r; ++dx) {
                TapSpec t;
                t.wtap = (op.ksize == 3) ? (dy + 1) * 3 + (dx + 1) : 0;
                if (op.stride == 1) {
                    t.plane = plane; t.shift = row_shift * L.wp + dx;
                } else if (dx == 0) {
                    t.plane = plane; t.shift = row_shift * L.wp;          // even columns
                } else {
                    t.plane = plane + 1; t.shift = row_shift * L.wp + (dx < 0 ? -1 : 0);  // odd columns
                }
                job.taps.push_back(t);
            }
        };
        auto set_out = [&](JobSpec& job, int plane) {
            job.out0 = plane;
            job.out1 = out.colsplit ? plane + 1 : -1;
        };

        bool need_interior = false;
        for (int j : out.needed) {
            if (!out.spec.count(j)) { need_interior = true; continue; }
            JobSpec job;
            job.tag = out.tag + ".r" + std::to_string(j);
            for (int dy = -r; dy <= r; ++dy) {
                const int i = op.stride * j + dy;
                if (i < 0 || i >= in.H) continue;  // the window's zero padding
                int plane, rs;
                resolve(in, i, plane, rs);
                add_taps(job, plane, rs, dy);
            }
            if (op.res >= 0) {
                int plane, rs;
                resolve(lv[op.res], j, plane, rs);
                job.res_plane = plane; job.res_shift = rs * L.wp;
            }
            set_out(job, out.spec_plane.at(j));
            L.jobs.push_back(job);
        }
        if (need_interior) {
            JobSpec job;
            job.tag = out.tag + ".int";
            if (in.interior < 0) throw std::runtime_error("planner: missing interior input at " + op.conv);
            for (int dy = -r; dy <= r; ++dy) add_taps(job, in.interior, dy * in.res, dy);
            if (op.res >= 0) {
                if (lv[op.res].interior < 0) throw std::runtime_error("planner: missing interior residual at " + op.conv);
                job.res_plane = lv[op.res].interior; job.res_shift = 0;
            }
            set_out(job, out.interior);
            L.jobs.push_back(job);
        }
        for (auto& job : L.jobs) {
            // split precision executes hi*hi + lo*hi + hi*lo (three products per tap; two when the input is plain fp16)
            const double m = static_cast<double>(job.taps.size()) * L.cin * L.cout * L.wp * (L.split_w ? (L.split_in ? 3 : 2) : 1);
            plan.macs_per_row += m;
            plan.gemm_macs_per_row += m;
        }
        // split launches that exceed the per-launch job table
        for (size_t o = 0; o < L.jobs.size(); o += kMaxJobs) {
            ConvLaunchSpec part = L;
            part.jobs.assign(L.jobs.begin() + o, L.jobs.begin() + std::min(L.jobs.size(), o + kMaxJobs));
            plan.convs.push_back(part);
        }
    }

    // ---- head ------------------------------------------------------------------------------------------
    {
        const Level& f = lv[final_level];
        plan.head_wp = f.wp; plan.head_C = f.C;
        for (int i = 0; i < plan.head_pool_groups * 4; ++i) {
            HeadRowSpec h;
            resolve(f, i, h.plane, h.row_shift);
            plan.head_rows.push_back(h);
        }
    }
    return plan;
}

std::string plan_to_json(const Plan& plan) {
    std::ostringstream o;
    o << "{\"H\":" << plan.H << ",\"W\":" << plan.W << ",\"guard_rows\":" << kGuardRows
      << ",\"macs_per_row\":" << static_cast<long long>(plan.macs_per_row) << ",\"planes\":[";
    for (size_t i = 0; i < plan.planes.size(); ++i) {
        const auto& p = plan.planes[i];
        o << (i ? "," : "") << "{\"id\":" << p.id << ",\"C\":" << p.C << ",\"wp\":" << p.wp << ",\"split\":" << (p.split ? 1 : 0)
          << ",\"tag\":\"" << p.tag << "\"}";
    }
    o << "],\"stem_wp\":" << plan.stem_wp << ",\"stem\":[";
    for (size_t i = 0; i < plan.stem.size(); ++i) {
        const auto& s = plan.stem[i];
        o << (i ? "," : "") << "{\"out\":" << s.out_plane << ",\"row_shift\":" << s.row_shift << ",\"mask\":" << s.mask << "}";
    }
    o << "],\"convs\":[";
    for (size_t i = 0; i < plan.convs.size(); ++i) {
        const auto& c = plan.convs[i];
        o << (i ? "," : "") << "{\"conv\":\"" << c.conv << "\",\"bn\":\"" << c.bn << "\",\"cin\":" << c.cin
          << ",\"cout\":" << c.cout << ",\"ksize\":" << c.ksize << ",\"relu\":" << c.relu << ",\"wp\":" << c.wp << ",\"w_real\":" << c.w_real
          << ",\"out_mode\":" << c.out_mode << ",\"wp2\":" << c.wp2 << ",\"hp\":" << c.hp << ",\"split_in\":" << c.split_in
          << ",\"split_out\":" << c.split_out << ",\"split_w\":" << c.split_w << ",\"jobs\":[";
        for (size_t j = 0; j < c.jobs.size(); ++j) {
            const auto& job = c.jobs[j];
            o << (j ? "," : "") << "{\"tag\":\"" << job.tag << "\",\"out0\":" << job.out0 << ",\"out1\":" << job.out1
              << ",\"res\":" << job.res_plane << ",\"res_shift\":" << job.res_shift << ",\"taps\":[";
            for (size_t t = 0; t < job.taps.size(); ++t)
                o << (t ? "," : "") << "[" << job.taps[t].plane << "," << job.taps[t].shift << "," << job.taps[t].wtap << "]";
            o << "]}";
        }
        o << "]}";
    }
    o << "],\"head\":{\"wp\":" << plan.head_wp << ",\"C\":" << plan.head_C << ",\"pool_groups\":" << plan.head_pool_groups
      << ",\"rows\":[";
    for (size_t i = 0; i < plan.head_rows.size(); ++i)
        o << (i ? "," : "") << "[" << plan.head_rows[i].plane << "," << plan.head_rows[i].row_shift << "]";
    o << "]}}";
    return o.str();
}

}  // namespace ld
