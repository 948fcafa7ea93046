// Task partition of the elimination tree and its host replay. See tasks.hpp.
#include "tasks.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <stdexcept>

namespace jgb {

TaskOptions task_options_from_env() {
    TaskOptions o;
    auto envi = [](const char* k, int d) { const char* v = getenv(k); return (v && *v) ? atoi(v) : d; };
    o.enabled = envi("JGB_TASKS", 0) == 1;      // off by default: see tasks.hpp
    o.maxnf = std::min(16, envi("JGB_TASK_MAXNF", o.maxnf));
    o.stack_lim = envi("JGB_TASK_STACK", o.stack_lim);
    o.meta_lim = envi("JGB_TASK_META", o.meta_lim);
    o.bundle = envi("JGB_TASK_BUNDLE", o.bundle);
    o.task_max = envi("JGB_TASK_MAXFRONTS", o.task_max);
    return o;
}

// A front is eligible when it and all its descendants have at most `maxnf` rows; eligible fronts are grouped bottom-up
// into subtrees as long as the shared-memory stack of update blocks (elements per scenario) and the task's index data
// stay within their caps — where they do not, the children with the largest footprint are cut off and become tasks of
// their own whose root block goes through HBM. Tasks of the same dependency level and size class are then packed,
// several per CTA, so that the start-up latency of a CTA (descriptor -> index data) is paid once per ~`bundle` fronts.
void partition_tasks(const Symbolic& sym, const TaskOptions& opt, TaskPlan& out) {
    out = TaskPlan();
    const int F = sym.nfronts;
    out.in_task.assign(F, 0);
    const int maxnf = opt.maxnf, stack_lim = opt.stack_lim, meta_lim = opt.meta_lim, bundle = opt.bundle,
              task_max = opt.task_max;
    if (!opt.enabled || maxnf < 4) return;
    auto updsz = [&](int f) { const int u = sym.f_nf[f] - sym.f_k[f]; return u * (u + 1); };
    auto metasz = [&](int f) {       // upper bound (TE = 8) of the blob ints of one front
        int m = kTaskRec + 9 + 2 * (sym.f_asmptr[f + 1] - sym.f_asmptr[f] + sym.f_k[f]);
        for (int ci = sym.f_childptr[f]; ci < sym.f_childptr[f + 1]; ++ci) {
            const int c = sym.f_children[ci];
            m += 4 + sym.f_nf[c] - sym.f_k[c];
        }
        return m;
    };
    std::vector<char> elig(F, 0), closed(F, 0);
    std::vector<int> peak_open(F, 0), peak_in(F, 0), meta_open(F, 0), cnt_open(F, 0);
    for (int f = 0; f < F; ++f) {
        bool ok = sym.f_nf[f] <= maxnf;
        for (int ci = sym.f_childptr[f]; ci < sym.f_childptr[f + 1] && ok; ++ci) ok = elig[sym.f_children[ci]];
        elig[f] = ok;
        if (!ok) {
            for (int ci = sym.f_childptr[f]; ci < sym.f_childptr[f + 1]; ++ci)
                if (elig[sym.f_children[ci]]) closed[sym.f_children[ci]] = 1;
            continue;
        }
        std::vector<int> open;
        for (int ci = sym.f_childptr[f]; ci < sym.f_childptr[f + 1]; ++ci) open.push_back(sym.f_children[ci]);
        const int own = metasz(f);
        for (;;) {
            int base = 0, pk = 0, meta = own, cnt = 1;
            for (int c : open) {
                pk = std::max(pk, base + peak_open[c]);
                base += updsz(c);
                meta += meta_open[c];
                cnt += cnt_open[c];
            }
            pk = std::max(pk, base);
            if ((pk <= stack_lim && meta <= meta_lim && cnt <= task_max) || open.empty()) {
                peak_in[f] = pk;
                peak_open[f] = std::max(pk, updsz(f));
                meta_open[f] = meta;
                cnt_open[f] = cnt;
                break;
            }
            // cut off the child with the largest footprint (stack need first, then index data)
            size_t worst = 0;
            for (size_t q = 1; q < open.size(); ++q) {
                const int a = open[q], b = open[worst];
                const long long ka = (long long)std::max(peak_open[a], updsz(a)) * 65536 + meta_open[a];
                const long long kb = (long long)std::max(peak_open[b], updsz(b)) * 65536 + meta_open[b];
                if (ka > kb) worst = q;
            }
            closed[open[worst]] = 1;
            open.erase(open.begin() + worst);
        }
        if (own > meta_lim) { elig[f] = 0; for (int c : open) closed[c] = 1; continue; }
        if (sym.f_parent[f] < 0) closed[f] = 1;
    }
    // a root whose parent is not eligible was closed above; an eligible front with an ineligible parent too
    std::vector<int> root_of(F, -1), level(F, 0);
    for (int f = F - 1; f >= 0; --f) {
        if (!elig[f]) continue;
        const int p = sym.f_parent[f];
        if (p >= 0 && !elig[p]) closed[f] = 1;
        root_of[f] = closed[f] ? f : root_of[p];
    }
    for (int f = 0; f < F; ++f) {
        if (!elig[f]) continue;
        const int r = root_of[f];
        for (int ci = sym.f_childptr[f]; ci < sym.f_childptr[f + 1]; ++ci) {
            const int c = sym.f_children[ci];
            if (root_of[c] != r) level[r] = std::max(level[r], level[root_of[c]] + 1);
        }
    }
    std::vector<std::vector<int>> members(F);
    std::vector<int> roots;
    for (int f = 0; f < F; ++f)
        if (elig[f]) {
            members[root_of[f]].push_back(f);
            if (root_of[f] == f) roots.push_back(f);
            out.in_task[f] = 1;
        }
    if (roots.empty()) return;
    out.task_fronts = 0;
    for (int r : roots) out.task_fronts += (int)members[r].size();
    out.task_count = (int)roots.size();

    struct Cls { int te, maxnf; };
    auto cls_of = [&](int mx) { return mx <= 8 ? Cls{4, 8} : mx <= 12 ? Cls{4, 12} : Cls{8, 16}; };
    struct Bundle { std::vector<int> roots; int level, te, maxnf, fronts, meta, stack, front_cap; double work; };
    std::vector<Bundle> bundles;
    int nlev = 0;
    for (int r : roots) nlev = std::max(nlev, level[r] + 1);
    for (int l = 0; l < nlev; ++l)
        for (int v = 0; v < 3; ++v) {
            const int vm = v == 0 ? 8 : v == 1 ? 12 : 16;
            Bundle cur{};
            auto flush = [&] { if (!cur.roots.empty()) bundles.push_back(cur); cur = Bundle{}; };
            for (int r : roots) {
                if (level[r] != l) continue;
                int mx = 0;
                for (int f : members[r]) mx = std::max(mx, sym.f_nf[f]);
                const Cls c = cls_of(mx);
                if (c.maxnf != vm) continue;
                const int nfr = (int)members[r].size();
                if (!cur.roots.empty() && (cur.fronts + nfr > bundle || cur.meta + meta_open[r] > meta_lim)) flush();
                if (cur.roots.empty()) { cur.level = l; cur.te = c.te; cur.maxnf = c.maxnf; cur.meta = 1; }
                cur.roots.push_back(r);
                cur.fronts += nfr;
                cur.meta += meta_open[r];
                cur.stack = std::max(cur.stack, peak_in[r]);
                cur.front_cap = std::max(cur.front_cap, mx * (mx + 1));
                for (int f : members[r]) cur.work += (double)sym.f_nf[f] * sym.f_nf[f] * (sym.f_k[f] + 4);
            }
            flush();
        }
    // heaviest CTAs first inside a launch (shorter tail); launches ordered by level, then class
    std::stable_sort(bundles.begin(), bundles.end(), [](const Bundle& a, const Bundle& b) {
        if (a.level != b.level) return a.level < b.level;
        if (a.maxnf != b.maxnf) return a.maxnf < b.maxnf;
        return a.work > b.work;
    });
    std::vector<int>& blob = out.blob;
    std::vector<int>& descs = out.descs;
    std::vector<int> stack_off(F, -1);
    int last_level = -1;
    for (const Bundle& b : bundles) {
        const int start = (int)blob.size();
        const int te = b.te;
        std::vector<int> fl;
        for (int r : b.roots) fl.insert(fl.end(), members[r].begin(), members[r].end());
        blob.push_back((int)fl.size());
        const int rec0 = (int)blob.size();
        blob.resize(blob.size() + fl.size() * kTaskRec, 0);
        int sp = 0, peak = 0;
        for (size_t q = 0; q < fl.size(); ++q) {
            const int f = fl[q], nf = sym.f_nf[f], k = sym.f_k[f], r = root_of[f];
            int* rec = nullptr;       // (re)taken after every push_back below
            // entry lists by owner warp
            std::vector<std::vector<std::pair<int, int>>> sub(te);
            for (int a = sym.f_asmptr[f]; a < sym.f_asmptr[f + 1]; ++a)
                sub[(sym.asm_dst[a] / nf) % te].push_back({sym.asm_src[a], sym.asm_dst[a]});
            const int* rows = &sym.f_rows[sym.f_rowptr[f]];
            for (int p = 0; p < k; ++p) sub[nf % te].push_back({-rows[p] - 1, p + nf * nf});
            const int asmoff = (int)blob.size() - start;
            int cum = 0;
            for (int w = 0; w < te; ++w) { blob.push_back(cum); cum += (int)sub[w].size(); }
            blob.push_back(cum);
            for (int w = 0; w < te; ++w)
                for (auto& e : sub[w]) { blob.push_back(e.first); blob.push_back(e.second); }
            const int childoff = (int)blob.size() - start;
            int base = -1, nchild = 0;
            for (int ci = sym.f_childptr[f]; ci < sym.f_childptr[f + 1]; ++ci) {
                const int c = sym.f_children[ci], uc = sym.f_nf[c] - sym.f_k[c];
                const bool inside = elig[c] && root_of[c] == r;
                if (inside && base < 0) base = stack_off[c];
                if (inside) out.upd_on_chip += (long long)uc * (uc + 1);
                blob.push_back(uc);
                blob.push_back(inside ? stack_off[c] : -1);
                out.off_pos.push_back((int)blob.size());
                out.off_front.push_back(c);
                blob.push_back((int)(sym.f_updoff[c] & 0xffffffffLL));
                blob.push_back((int)(sym.f_updoff[c] >> 32));
                for (int i = 0; i < uc; ++i) blob.push_back(sym.f_rel[sym.f_relptr[c] + i]);
                ++nchild;
            }
            if (base < 0) base = sp;
            if (f != r) {
                stack_off[f] = base;
                sp = base + updsz(f);
            } else {
                sp = base;
            }
            peak = std::max(peak, sp);
            rec = &blob[rec0 + q * kTaskRec];
            rec[0] = nf; rec[1] = k; rec[2] = asmoff; rec[3] = childoff; rec[4] = nchild;
            rec[5] = (f != r) ? stack_off[f] : -1;
            rec[6] = (int)(sym.f_uoff[f] & 0xffffffffLL); rec[7] = (int)(sym.f_uoff[f] >> 32);
            rec[8] = (int)(sym.f_updoff[f] & 0xffffffffLL); rec[9] = (int)(sym.f_updoff[f] >> 32);
            rec[10] = 32;
            out.uoff_pos.push_back((int)(rec0 + q * kTaskRec + 6));
            out.uoff_front.push_back(f);
            out.off_pos.push_back((int)(rec0 + q * kTaskRec + 8));
            out.off_front.push_back(f);
            out.wout_pos.push_back((int)(rec0 + q * kTaskRec + 10));
            out.wout_front.push_back(f);
        }
        if (sp != 0) throw std::logic_error("task stack not empty at the end of a task list");
        while ((blob.size() - start) % 4) blob.push_back(0);
        descs.push_back(start);
        descs.push_back((int)blob.size() - start);
        // launches: consecutive bundles of equal (level, class)
        const int stack_cap = std::max(peak, 1), meta_cap = (int)blob.size() - start;
        if (!out.launches.empty() && out.launches.back().te == b.te && out.launches.back().maxnf == b.maxnf &&
            last_level == b.level) {
            TaskLaunch& tl = out.launches.back();
            tl.count++;
            tl.front_cap = std::max(tl.front_cap, b.front_cap);
            tl.stack_cap = std::max(tl.stack_cap, stack_cap);
            tl.meta_cap = std::max(tl.meta_cap, meta_cap);
        } else {
            last_level = b.level;
            TaskLaunch tl{};
            tl.begin = (int)descs.size() / 2 - 1; tl.count = 1; tl.te = b.te; tl.maxnf = b.maxnf;
            tl.front_cap = b.front_cap; tl.stack_cap = stack_cap; tl.meta_cap = meta_cap;
            out.launches.push_back(tl);
        }
    }
    for (TaskLaunch& tl : out.launches) {
        tl.smem = (size_t)(tl.front_cap + 2 * tl.maxnf + tl.stack_cap) * 256 + (size_t)tl.meta_cap * 4;
    }
}

int host_task_factor_solve(const Symbolic& S, const TaskPlan& tp, const double* aval, const double* rhs, double* x) {
    std::vector<double> U(S.u_size), upd(S.upd_size), F, stack;
    auto eliminate = [&](int nf, int k, long long uoff) {
        for (int p = 0; p < k; ++p) {
            const double piv = F[p + (size_t)p * nf];
            if (piv == 0.0 || !std::isfinite(piv)) return -3;
            const double inv = 1.0 / piv;
            for (int j = p + 1; j <= nf; ++j) {
                const double upj = F[p + (size_t)j * nf];
                for (int i = p + 1; i < nf; ++i) F[i + (size_t)j * nf] -= F[i + (size_t)p * nf] * inv * upj;
            }
            double* Urow = &U[uoff + (long long)p * (nf + 1) - (long long)p * (p - 1) / 2];
            Urow[0] = inv;
            for (int j = p + 1; j <= nf; ++j) Urow[j - p] = F[p + (size_t)j * nf];
        }
        return 0;
    };
    for (const TaskLaunch& tl : tp.launches)
        for (int t = tl.begin; t < tl.begin + tl.count; ++t) {
            const int* meta = &tp.blob[tp.descs[2 * t]];
            stack.assign((size_t)tl.stack_cap, 0.0);
            for (int fi = 0; fi < meta[0]; ++fi) {
                const int* fr = meta + 1 + fi * kTaskRec;
                const int nf = fr[0], k = fr[1], u = nf - k;
                if (nf * (nf + 1) > tl.front_cap) throw std::logic_error("task replay: front exceeds the front area");
                F.assign((size_t)nf * (nf + 1), 0.0);
                const int* aw = meta + fr[2];
                const int* pairs = aw + tl.te + 1;
                for (int w = 0; w < tl.te; ++w)
                    for (int a = aw[w]; a < aw[w + 1]; ++a) {
                        const int src = pairs[2 * a], dst = pairs[2 * a + 1];
                        if ((dst / nf) % tl.te != w) throw std::logic_error("task replay: entry in the wrong sublist");
                        F[dst] = src >= 0 ? aval[src] : rhs[-src - 1];
                    }
                const int* cp = meta + fr[3];
                for (int ci = 0; ci < fr[4]; ++ci) {
                    const int uc = cp[0], soff = cp[1];
                    const int* rel = cp + 4;
                    const long long uo = ((long long)cp[3] << 32) | (unsigned)cp[2];
                    if (soff >= 0 && soff + uc * (uc + 1) > tl.stack_cap) throw std::logic_error("task replay: stack overflow");
                    const double* src = soff >= 0 ? &stack[soff] : &upd[uo];
                    for (int j = 0; j <= uc; ++j) {
                        const int C = j < uc ? rel[j] : nf;
                        for (int i = 0; i < uc; ++i) F[rel[i] + (size_t)C * nf] += src[i + (size_t)j * uc];
                    }
                    cp += 4 + uc;
                }
                const long long uoff = ((long long)fr[7] << 32) | (unsigned)fr[6];
                if (eliminate(nf, k, uoff)) return -3;
                double* C;
                if (fr[5] >= 0) {
                    if (fr[5] + u * (u + 1) > tl.stack_cap) throw std::logic_error("task replay: stack overflow");
                    C = &stack[fr[5]];
                } else {
                    C = &upd[((long long)fr[9] << 32) | (unsigned)fr[8]];
                }
                for (int j = 0; j <= u; ++j)
                    for (int i = 0; i < u; ++i) C[i + (size_t)j * u] = F[(k + i) + (size_t)(k + j) * nf];
            }
        }
    for (int f = 0; f < S.nfronts; ++f) {
        if (!tp.in_task.empty() && tp.in_task[f]) continue;
        const int k = S.f_k[f], nf = S.f_nf[f], u = nf - k;
        const int* rows = &S.f_rows[S.f_rowptr[f]];
        F.assign((size_t)nf * (nf + 1), 0.0);
        for (int a = S.f_asmptr[f]; a < S.f_asmptr[f + 1]; ++a) F[S.asm_dst[a]] += aval[S.asm_src[a]];
        for (int p = 0; p < k; ++p) F[p + (size_t)nf * nf] = rhs[rows[p]];
        for (int ci = S.f_childptr[f]; ci < S.f_childptr[f + 1]; ++ci) {
            const int c = S.f_children[ci], uc = S.f_nf[c] - S.f_k[c];
            const int* rel = &S.f_rel[S.f_relptr[c]];
            const double* src = &upd[S.f_updoff[c]];
            for (int j = 0; j <= uc; ++j) {
                const int C = j < uc ? rel[j] : nf;
                for (int i = 0; i < uc; ++i) F[rel[i] + (size_t)C * nf] += src[i + (size_t)j * uc];
            }
        }
        if (eliminate(nf, k, S.f_uoff[f])) return -3;
        double* C = &upd[S.f_updoff[f]];
        for (int j = 0; j <= u; ++j)
            for (int i = 0; i < u; ++i) C[i + (size_t)j * u] = F[(k + i) + (size_t)(k + j) * nf];
    }
    for (int f = S.nfronts - 1; f >= 0; --f) {
        const int k = S.f_k[f], nf = S.f_nf[f];
        const int* rows = &S.f_rows[S.f_rowptr[f]];
        for (int p = k - 1; p >= 0; --p) {
            const double* Urow = &U[S.f_uoff[f] + (long long)p * (nf + 1) - (long long)p * (p - 1) / 2];
            double acc = Urow[nf - p];
            for (int j = p + 1; j < nf; ++j) acc -= Urow[j - p] * x[rows[j]];
            x[rows[p]] = acc * Urow[0];
        }
    }
    return 0;
}

}  // namespace jgb
