// Symbolic analysis: minimum-degree ordering on the bus (group) graph, elimination tree, relaxed
// supernodes, multifrontal assembly maps and level schedules.  See symbolic.hpp.
#include "symbolic.hpp"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <set>
#include <string>
#include <tuple>
#include <stdexcept>

namespace jgb {
namespace {

// sorted-vector set union of a and b, dropping x and y
void merge_drop(const std::vector<int>& a, const std::vector<int>& b, int x, int y, std::vector<int>& out) {
    out.clear();
    out.reserve(a.size() + b.size());
    size_t i = 0, j = 0;
    while (i < a.size() || j < b.size()) {
        int v;
        if (j >= b.size() || (i < a.size() && a[i] < b[j])) v = a[i++];
        else if (i >= a.size() || b[j] < a[i]) v = b[j++];
        else { v = a[i]; ++i; ++j; }
        if (v != x && v != y) out.push_back(v);
    }
}

// Minimum (weighted external) degree on an explicit elimination graph. Graphs here are bus graphs of
// power grids (10^4..10^5 nodes, fill a few 10^5), for which the explicit clique update is cheap.
std::vector<int> min_degree(int ng, std::vector<std::vector<int>> adj, const std::vector<int>& weight) {
    std::vector<long> deg(ng, 0);
    std::set<std::pair<long, int>> heap;
    for (int v = 0; v < ng; ++v) {
        long d = 0;
        for (int u : adj[v]) d += weight[u];
        deg[v] = d;
        heap.insert({d, v});
    }
    std::vector<int> order;
    order.reserve(ng);
    std::vector<char> dead(ng, 0);
    std::vector<int> tmp;
    while (!heap.empty()) {
        int v = heap.begin()->second;
        heap.erase(heap.begin());
        dead[v] = 1;
        order.push_back(v);
        const std::vector<int> nb = adj[v];
        for (int u : nb) {
            heap.erase({deg[u], u});
            merge_drop(adj[u], nb, u, v, tmp);
            adj[u].swap(tmp);
            long d = 0;
            for (int w : adj[u]) d += weight[w];
            deg[u] = d;
            heap.insert({d, u});
        }
        std::vector<int>().swap(adj[v]);
    }
    return order;
}

// Minimum local fill (Tinney scheme 3) on the same explicit elimination graph: the next vertex is the one whose
// elimination creates the fewest (weighted) fill edges, ties broken by weighted degree, then by id. The fill count of a
// vertex changes only when its neighbourhood or an edge between two of its neighbours changes, i.e. for the
// neighbours of the eliminated vertex and their neighbours.
std::vector<int> min_fill(int ng, std::vector<std::vector<int>> adj, const std::vector<int>& weight,
                          size_t kExactDegree) {
    auto has = [&](int a, int b) { return std::binary_search(adj[a].begin(), adj[a].end(), b); };
    // exact counts only for moderate degrees: beyond that (the dense tail of the elimination, where every remaining
    // vertex sits in a few big cliques) the quadratic count per vertex dominates the run time and no longer changes the
    // order much, so an upper bound (all pairs missing) ranks those vertices by degree behind the exact ones
    auto fill_of = [&](int v) {
        long f = 0;
        const std::vector<int>& nb = adj[v];
        if (nb.size() > kExactDegree) {
            long d = 0;
            for (int u : nb) d += weight[u];
            return (1L << 40) + d * d;
        }
        for (size_t i = 0; i < nb.size(); ++i)
            for (size_t j = i + 1; j < nb.size(); ++j)
                if (!has(nb[i], nb[j])) f += (long)weight[nb[i]] * weight[nb[j]];
        return f;
    };
    auto deg_of = [&](int v) { long d = 0; for (int u : adj[v]) d += weight[u]; return d; };
    std::vector<long> fill(ng), deg(ng);
    std::set<std::tuple<long, long, int>> heap;
    for (int v = 0; v < ng; ++v) { fill[v] = fill_of(v); deg[v] = deg_of(v); heap.insert({fill[v], deg[v], v}); }
    std::vector<int> order;
    order.reserve(ng);
    std::vector<int> tmp, dirty;
    std::vector<char> mark(ng, 0);
    while (!heap.empty()) {
        const int v = std::get<2>(*heap.begin());
        heap.erase(heap.begin());
        order.push_back(v);
        const std::vector<int> nb = adj[v];
        dirty.clear();
        for (int u : nb) {
            if (!mark[u]) { mark[u] = 1; dirty.push_back(u); }
            for (int w : adj[u])
                if (w != v && !mark[w] && adj[w].size() <= kExactDegree) { mark[w] = 1; dirty.push_back(w); }
        }
        for (int u : dirty) heap.erase({fill[u], deg[u], u});
        for (int u : nb) {
            merge_drop(adj[u], nb, u, v, tmp);
            adj[u].swap(tmp);
        }
        std::vector<int>().swap(adj[v]);
        for (int u : dirty) {
            mark[u] = 0;
            fill[u] = fill_of(u);
            deg[u] = deg_of(u);
            heap.insert({fill[u], deg[u], u});
        }
    }
    return order;
}

struct Structs {
    std::vector<std::vector<int>> st;   // st[j]: rows > j of column j of L (sorted)
    std::vector<int> parent;
};

// Column structures of L for the ordering `perm` (position -> variable) on the symmetric graph `sadj`
void column_structures(int n, const std::vector<std::vector<int>>& sadj, const std::vector<int>& perm,
                       const std::vector<int>& iperm, Structs& out) {
    out.st.assign(n, {});
    out.parent.assign(n, -1);
    std::vector<std::vector<int>> children(n);
    std::vector<int> tmp;
    for (int j = 0; j < n; ++j) {
        std::vector<int>& s = out.st[j];
        for (int w : sadj[perm[j]]) {
            int p = iperm[w];
            if (p > j) s.push_back(p);
        }
        std::sort(s.begin(), s.end());
        s.erase(std::unique(s.begin(), s.end()), s.end());
        for (int c : children[j]) {
            merge_drop(s, out.st[c], j, j, tmp);
            s.swap(tmp);
        }
        if (!s.empty()) {
            out.parent[j] = s.front();
            children[s.front()].push_back(j);
        }
    }
}

}  // namespace

void analyse(int n, const int* colptr, const int* rowidx, const int* group, const unsigned char* /*skip*/,
             const SymbolicOptions& opt_in, Symbolic& S) {
    SymbolicOptions opt = opt_in;
    if (const char* env = getenv("JGB_RELAX")) {   // tuning experiments: small,mid,midfrac,big,bigfrac,anyfrac
        sscanf(env, "%d,%d,%lf,%d,%lf,%lf", &opt.relax_small, &opt.relax_mid, &opt.relax_mid_frac, &opt.relax_big,
               &opt.relax_big_frac, &opt.relax_any_frac);
    }
    S = Symbolic();
    S.n = n;
    // ---- symmetric scalar graph (no diagonal)
    std::vector<std::vector<int>> sadj(n);
    for (int c = 0; c < n; ++c)
        for (int q = colptr[c]; q < colptr[c + 1]; ++q) {
            int r = rowidx[q];
            if (r < 0 || r >= n) throw std::runtime_error("symbolic: row index out of range");
            if (r != c) { sadj[c].push_back(r); sadj[r].push_back(c); }
        }
    for (auto& a : sadj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }

    // ---- group graph
    std::vector<int> grp(n);
    int ng = 0;
    if (group) {
        std::vector<int> remap;
        int maxg = 0;
        for (int v = 0; v < n; ++v) maxg = std::max(maxg, group[v]);
        remap.assign(maxg + 1, -1);
        for (int v = 0; v < n; ++v) {
            if (remap[group[v]] < 0) remap[group[v]] = ng++;
            grp[v] = remap[group[v]];
        }
    } else {
        for (int v = 0; v < n; ++v) grp[v] = v;
        ng = n;
    }
    std::vector<std::vector<int>> members(ng);
    for (int v = 0; v < n; ++v) members[grp[v]].push_back(v);
    std::vector<std::vector<int>> gadj(ng);
    std::vector<int> weight(ng);
    for (int g = 0; g < ng; ++g) {
        weight[g] = (int)members[g].size();
        for (int v : members[g])
            for (int w : sadj[v])
                if (grp[w] != g) gadj[g].push_back(grp[w]);
        std::sort(gadj[g].begin(), gadj[g].end());
        gadj[g].erase(std::unique(gadj[g].begin(), gadj[g].end()), gadj[g].end());
    }
    // variables of one group are made mutually adjacent so that they form one supervariable
    for (int g = 0; g < ng; ++g)
        for (int v : members[g])
            for (int w : members[g])
                if (v != w) sadj[v].push_back(w);
    for (auto& a : sadj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }

    bool use_fill = opt.min_fill;
    if (const char* ord = getenv("JGB_ORDER")) use_fill = std::string(ord) == "fill";     // "degree" | "fill": tuning only
    size_t exact = (size_t)std::max(1, opt.fill_exact_degree);
    if (const char* k = getenv("JGB_FILL_K")) exact = (size_t)std::max(1, atoi(k));
    std::vector<int> gorder = use_fill ? min_fill(ng, gadj, weight, exact) : min_degree(ng, gadj, weight);

    std::vector<int> perm;
    perm.reserve(n);
    for (int g : gorder)
        for (int v : members[g]) perm.push_back(v);
    std::vector<int> iperm(n);
    for (int j = 0; j < n; ++j) iperm[perm[j]] = j;

    // ---- elimination tree, postorder (largest-structure child last), final structures
    Structs st;
    column_structures(n, sadj, perm, iperm, st);
    {
        std::vector<std::vector<int>> children(n);
        std::vector<int> roots;
        for (int j = 0; j < n; ++j) {
            if (st.parent[j] >= 0) children[st.parent[j]].push_back(j);
            else roots.push_back(j);
        }
        for (auto& ch : children)
            std::stable_sort(ch.begin(), ch.end(),
                             [&](int a, int b) { return st.st[a].size() < st.st[b].size(); });
        std::vector<int> post;
        post.reserve(n);
        std::vector<std::pair<int, size_t>> stack;
        for (int r : roots) {
            stack.push_back({r, 0});
            while (!stack.empty()) {
                auto& top = stack.back();
                if (top.second < children[top.first].size()) {
                    int c = children[top.first][top.second++];
                    stack.push_back({c, 0});
                } else {
                    post.push_back(top.first);
                    stack.pop_back();
                }
            }
        }
        std::vector<int> perm2(n);
        for (int j = 0; j < n; ++j) perm2[j] = perm[post[j]];
        perm.swap(perm2);
        for (int j = 0; j < n; ++j) iperm[perm[j]] = j;
        column_structures(n, sadj, perm, iperm, st);
    }
    S.perm = perm;
    S.iperm = iperm;

    // ---- supernodes: fundamental chains, then relaxed amalgamation of last children
    struct SN { int first, last; int u; double nnz; };   // nnz: true entries of the L trapezoid (incl. diagonal)
    std::vector<SN> sns;
    for (int j = 0; j < n; ++j) {
        int cnt = (int)st.st[j].size();
        bool join = false;
        if (!sns.empty()) {
            SN& t = sns.back();
            if (t.last == j - 1 && st.parent[j - 1] == j && (int)st.st[j - 1].size() == cnt + 1) join = true;
        }
        if (join) { sns.back().last = j; sns.back().u = cnt; sns.back().nnz += cnt + 1; }
        else sns.push_back({j, j, cnt, (double)cnt + 1});
    }
    {
        std::vector<SN> merged;
        for (const SN& cur0 : sns) {
            SN cur = cur0;
            while (!merged.empty()) {
                SN& prev = merged.back();
                // prev is the last child of cur iff parent[prev.last] == cur.first and they are adjacent
                if (prev.last + 1 != cur.first || st.parent[prev.last] != cur.first) break;
                int k2 = (cur.last - cur.first + 1) + (prev.last - prev.first + 1);
                int nf2 = k2 + cur.u;
                double trap = 0;
                for (int c = 0; c < k2; ++c) trap += nf2 - c;
                double truennz = prev.nnz + cur.nnz;
                double frac = (trap - truennz) / trap;
                bool ok = (k2 <= opt.relax_small) || (k2 <= opt.relax_mid && frac < opt.relax_mid_frac) ||
                          (k2 <= opt.relax_big && frac < opt.relax_big_frac) || (frac < opt.relax_any_frac);
                if (!ok) break;
                cur.first = prev.first;
                cur.nnz = truennz;
                merged.pop_back();
            }
            merged.push_back(cur);
        }
        sns.swap(merged);
    }

    // ---- fronts
    int nf_total = (int)sns.size();
    S.nfronts = nf_total;
    std::vector<int> sn_of(n);
    for (int f = 0; f < nf_total; ++f)
        for (int j = sns[f].first; j <= sns[f].last; ++j) sn_of[j] = f;
    S.f_k.resize(nf_total);
    S.f_nf.resize(nf_total);
    S.f_parent.assign(nf_total, -1);
    S.f_rowptr.assign(nf_total + 1, 0);
    for (int f = 0; f < nf_total; ++f) {
        int k = sns[f].last - sns[f].first + 1;
        int u = (int)st.st[sns[f].last].size();
        S.f_k[f] = k;
        S.f_nf[f] = k + u;
        S.f_rowptr[f + 1] = S.f_rowptr[f] + k + u;
        if (u > 0) S.f_parent[f] = sn_of[st.st[sns[f].last].front()];
        S.max_front = std::max(S.max_front, k + u);
    }
    S.f_rows.resize(S.f_rowptr[nf_total]);
    // local position (in elimination numbering) of x inside front f
    auto local = [&](int f, int x) -> int {
        int a = sns[f].first, b = sns[f].last;
        if (x >= a && x <= b) return x - a;
        const std::vector<int>& up = st.st[b];
        auto it = std::lower_bound(up.begin(), up.end(), x);
        if (it == up.end() || *it != x) return -1;
        return (b - a + 1) + (int)(it - up.begin());
    };
    for (int f = 0; f < nf_total; ++f) {
        int o = S.f_rowptr[f];
        for (int j = sns[f].first; j <= sns[f].last; ++j) S.f_rows[o++] = perm[j];
        for (int r : st.st[sns[f].last]) S.f_rows[o++] = perm[r];
    }
    // children, relative indices
    S.f_childptr.assign(nf_total + 1, 0);
    for (int f = 0; f < nf_total; ++f)
        if (S.f_parent[f] >= 0) S.f_childptr[S.f_parent[f] + 1]++;
    for (int f = 0; f < nf_total; ++f) S.f_childptr[f + 1] += S.f_childptr[f];
    S.f_children.resize(S.f_childptr[nf_total]);
    {
        std::vector<int> fill(S.f_childptr.begin(), S.f_childptr.end() - 1);
        for (int f = 0; f < nf_total; ++f)
            if (S.f_parent[f] >= 0) S.f_children[fill[S.f_parent[f]]++] = f;
    }
    S.f_relptr.assign(nf_total + 1, 0);
    for (int f = 0; f < nf_total; ++f) S.f_relptr[f + 1] = S.f_relptr[f] + (S.f_nf[f] - S.f_k[f]);
    S.f_rel.resize(S.f_relptr[nf_total]);
    for (int f = 0; f < nf_total; ++f) {
        int p = S.f_parent[f];
        int o = S.f_relptr[f];
        for (int r : st.st[sns[f].last]) {
            int l = local(p, r);
            if (l < 0) throw std::runtime_error("symbolic: update row missing from parent front");
            S.f_rel[o++] = l;
        }
    }
    // ---- assembly map of the original entries
    {
        std::vector<std::vector<std::pair<int, int>>> per(nf_total);
        for (int c = 0; c < n; ++c)
            for (int q = colptr[c]; q < colptr[c + 1]; ++q) {
                int pr = iperm[rowidx[q]], pc = iperm[c];
                int f = sn_of[std::min(pr, pc)];
                int lr = local(f, pr), lc = local(f, pc);
                if (lr < 0 || lc < 0) throw std::runtime_error("symbolic: matrix entry outside the front structure");
                per[f].push_back({q, lr + lc * S.f_nf[f]});
            }
        S.f_asmptr.assign(nf_total + 1, 0);
        for (int f = 0; f < nf_total; ++f) S.f_asmptr[f + 1] = S.f_asmptr[f] + (int)per[f].size();
        S.asm_src.resize(S.f_asmptr[nf_total]);
        S.asm_dst.resize(S.f_asmptr[nf_total]);
        for (int f = 0; f < nf_total; ++f) {
            std::sort(per[f].begin(), per[f].end());
            int o = S.f_asmptr[f];
            for (auto& e : per[f]) { S.asm_src[o] = e.first; S.asm_dst[o] = e.second; ++o; }
        }
    }
    // ---- storage offsets, stats
    S.f_uoff.resize(nf_total);
    S.f_updoff.resize(nf_total);
    int64_t uo = 0, po = 0;
    for (int f = 0; f < nf_total; ++f) {
        int k = S.f_k[f], nf = S.f_nf[f], u = nf - k;
        S.f_uoff[f] = uo;
        S.f_updoff[f] = po;
        for (int p = 0; p < k; ++p) {
            uo += nf + 1 - p;
            S.nnz_lu += 2 * (nf - p) - 1;
            S.flops += 2.0 * (nf - p - 1) * (nf - p) + 1;
        }
        po += (int64_t)u * (u + 1);
    }
    S.u_size = uo;
    S.upd_size = po;
    if (po >= (int64_t)1 << 31) throw std::runtime_error("symbolic: update storage exceeds 2^31 elements");
    // ---- extend-add gather lists, organised in rounds (see symbolic.hpp)
    {
        S.f_eaptr.assign(nf_total + 1, 0);
        S.ea_roundptr.assign(1, 0);
        std::vector<std::pair<int, int>> rec;   // (dst, src) in child order
        std::vector<int> seen;                  // number of sources already met per destination
        for (int f = 0; f < nf_total; ++f) {
            rec.clear();
            const int nf = S.f_nf[f];
            for (int ci = S.f_childptr[f]; ci < S.f_childptr[f + 1]; ++ci) {
                const int c = S.f_children[ci];
                const int uc = S.f_nf[c] - S.f_k[c];
                const int* rel = &S.f_rel[S.f_relptr[c]];
                for (int j = 0; j <= uc; ++j) {
                    const int dc = (j < uc) ? rel[j] : nf;
                    for (int i = 0; i < uc; ++i)
                        rec.push_back({rel[i] + dc * nf, (int)(S.f_updoff[c] + i + (int64_t)j * uc)});
                }
            }
            seen.assign((size_t)nf * (nf + 1), 0);
            std::vector<std::vector<std::pair<int, int>>> rounds;
            for (auto& pr : rec) {
                int r = seen[pr.first]++;
                if ((int)rounds.size() <= r) rounds.resize(r + 1);
                rounds[r].push_back(pr);
            }
            for (auto& rd : rounds) {
                std::sort(rd.begin(), rd.end());
                for (auto& pr : rd) { S.ea_pair.push_back(pr.first); S.ea_pair.push_back(pr.second); }
                S.ea_roundptr.push_back((int)(S.ea_pair.size() / 2));
            }
            S.f_eaptr[f + 1] = (int)S.ea_roundptr.size() - 1;
            // symmetric variant: lower triangle + rhs only, packed destinations
            if (f == 0) { S.f_eaptr_sym.assign(nf_total + 1, 0); S.ea_roundptr_sym.assign(1, 0); }
            const int tri = nf * (nf + 1) / 2;
            seen.assign((size_t)tri + nf, 0);
            rounds.clear();
            for (auto& pr : rec) {
                const int c = pr.first / nf, r = pr.first - c * nf;
                int dst;
                if (c == nf) dst = tri + r;
                else if (r >= c) dst = (c * (2 * nf - c + 1)) / 2 + r - c;
                else continue;
                int rr = seen[dst]++;
                if ((int)rounds.size() <= rr) rounds.resize(rr + 1);
                rounds[rr].push_back({dst, pr.second});
            }
            for (auto& rd : rounds) {
                std::sort(rd.begin(), rd.end());
                for (auto& pr : rd) { S.ea_pair_sym.push_back(pr.first); S.ea_pair_sym.push_back(pr.second); }
                S.ea_roundptr_sym.push_back((int)(S.ea_pair_sym.size() / 2));
            }
            S.f_eaptr_sym[f + 1] = (int)S.ea_roundptr_sym.size() - 1;
        }
        if (nf_total == 0) { S.f_eaptr_sym.assign(1, 0); S.ea_roundptr_sym.assign(1, 0); }
    }
    // ---- level schedules
    std::vector<int> height(nf_total, 0), depth(nf_total, 0);
    for (int f = 0; f < nf_total; ++f)   // children precede parents (postorder)
        if (S.f_parent[f] >= 0) height[S.f_parent[f]] = std::max(height[S.f_parent[f]], height[f] + 1);
    for (int f = nf_total - 1; f >= 0; --f)
        if (S.f_parent[f] >= 0) depth[f] = depth[S.f_parent[f]] + 1;
    auto schedule = [&](const std::vector<int>& key, int& nl, std::vector<int>& ptr, std::vector<int>& list) {
        nl = 0;
        for (int f = 0; f < nf_total; ++f) nl = std::max(nl, key[f] + 1);
        ptr.assign(nl + 1, 0);
        for (int f = 0; f < nf_total; ++f) ptr[key[f] + 1]++;
        for (int l = 0; l < nl; ++l) ptr[l + 1] += ptr[l];
        list.resize(nf_total);
        std::vector<int> fill(ptr.begin(), ptr.end() - 1);
        for (int f = 0; f < nf_total; ++f) list[fill[key[f]]++] = f;
        for (int l = 0; l < nl; ++l)
            std::stable_sort(list.begin() + ptr[l], list.begin() + ptr[l + 1],
                             [&](int a, int b) { return S.f_nf[a] > S.f_nf[b]; });
    };
    schedule(height, S.nlevels, S.levelptr, S.level_fronts);
    schedule(depth, S.ndepths, S.depthptr, S.depth_fronts);
}

int host_factor_solve(const Symbolic& S, const double* aval, const double* rhs, double* x) {
    std::vector<double> U(S.u_size), upd(S.upd_size), F;
    for (int f = 0; f < S.nfronts; ++f) {
        int k = S.f_k[f], nf = S.f_nf[f], u = nf - k;
        const int* rows = &S.f_rows[S.f_rowptr[f]];
        F.assign((size_t)nf * (nf + 1), 0.0);
        for (int a = S.f_asmptr[f]; a < S.f_asmptr[f + 1]; ++a) F[S.asm_dst[a]] += aval[S.asm_src[a]];
        for (int p = 0; p < k; ++p) F[p + (size_t)nf * nf] = rhs[rows[p]];
        for (int r = S.f_eaptr[f]; r < S.f_eaptr[f + 1]; ++r)
            for (int t = S.ea_roundptr[r]; t < S.ea_roundptr[r + 1]; ++t)
                F[S.ea_pair[2 * t]] += upd[S.ea_pair[2 * t + 1]];
        for (int p = 0; p < k; ++p) {
            double piv = F[p + (size_t)p * nf];
            if (piv == 0.0 || !std::isfinite(piv)) return -3;
            double inv = 1.0 / piv;
            for (int j = p + 1; j <= nf; ++j) {
                double upj = F[p + (size_t)j * nf];
                if (upj == 0.0) continue;
                for (int i = p + 1; i < nf; ++i) F[i + (size_t)j * nf] -= F[i + (size_t)p * nf] * inv * upj;
            }
            double* Urow = &U[S.f_uoff[f] + (int64_t)p * (nf + 1) - (int64_t)p * (p - 1) / 2];
            Urow[0] = inv;
            for (int j = p + 1; j <= nf; ++j) Urow[j - p] = F[p + (size_t)j * nf];
        }
        double* C = &upd[S.f_updoff[f]];
        for (int j = 0; j <= u; ++j)
            for (int i = 0; i < u; ++i) C[i + (size_t)j * u] = F[(k + i) + (size_t)(k + j) * nf];
    }
    for (int f = S.nfronts - 1; f >= 0; --f) {
        int k = S.f_k[f], nf = S.f_nf[f];
        const int* rows = &S.f_rows[S.f_rowptr[f]];
        for (int p = k - 1; p >= 0; --p) {
            const double* Urow = &U[S.f_uoff[f] + (int64_t)p * (nf + 1) - (int64_t)p * (p - 1) / 2];
            double acc = Urow[nf - p];
            for (int j = p + 1; j < nf; ++j) acc -= Urow[j - p] * x[rows[j]];
            x[rows[p]] = acc * Urow[0];
        }
    }
    return 0;
}

}  // namespace jgb
