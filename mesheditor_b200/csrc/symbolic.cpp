// Geometric nested dissection + supernodal symbolic factorisation (host). See symbolic.h.
//
// Ordering: recursive coordinate bisection of the node graph. A subdomain is cut by the plane through its median
// coordinate along its longest axis; the side of the cut that touches the other side through fewer nodes gives the
// vertex separator. Tet meshes carry their embedding, and a planar cut of a 3-D mesh is within a small factor of the
// best separator, so no graph partitioner is needed (none exists in this image anyway).
// Supernodes are read straight off the dissection tree: each leaf subdomain is one supernode, each separator is a
// chain of panels of at most PanelNodes nodes. Their diagonal blocks are treated as dense (a separator is a clique
// once both sides are eliminated), the below-diagonal structure is the union of the graph adjacency and of the
// children's structures (a superset of the true fill; the few explicit zeros buy dense kernels).
#include "symbolic.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <future>
#include <memory>
#include <new>
#include <numeric>
#include <thread>

namespace me {
namespace {
// Runs body(begin, end, worker) over [0, count) split into contiguous chunks on up to 16 host threads.
template<typename Body>
void ParallelChunks(size_t count, size_t min_per_thread, Body &&body) {
    const size_t hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const size_t workers = std::max<size_t>(1, std::min(hw, count / std::max<size_t>(1, min_per_thread)));
    if (workers <= 1) {
        body(size_t(0), count, size_t(0));
        return;
    }
    std::vector<std::thread> pool;
    for (size_t w = 0; w < workers; ++w) pool.emplace_back([&, w] { body(count * w / workers, count * (w + 1) / workers, w); });
    for (auto &t : pool) t.join();
}

double Now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Dissector {
    uint32_t N;
    const uint32_t *RowPtr, *Col;
    const float *Xyz;
    SymbolicOptions Opt;
    // Current subdomain stamp of each node. The two halves of a cut are dissected on different threads (down to kForkDepth): they
    // write disjoint nodes, but a boundary scan reads the stamps of neighbours in the other half, hence relaxed atomics (a stamp
    // is unique to one call, so a neighbour's stamp changing under the scan never compares equal to the one looked for).
    std::unique_ptr<std::atomic<uint32_t>[]> Dom;
    std::atomic<uint32_t> NextStamp{1};
    static constexpr int kForkDepth = 4;

    // The supernodes of one subtree in elimination order, numbered from 0 inside it (the whole tree once the recursion returns).
    struct Sub {
        std::vector<uint32_t> Perm;       // new -> old
        std::vector<uint32_t> SuperFirst; // first index into Perm of each supernode
        std::vector<uint32_t> Parent;     // UINT32_MAX: none (yet)
        std::vector<uint32_t> Chain;      // panels of one separator (or one leaf) share a chain id
        uint32_t Chains{0};
        std::vector<uint32_t> Tops;       // tops of the subtrees emitted so far
    };
    static void Append(Sub &into, Sub &&part) {
        const uint32_t node0 = uint32_t(into.Perm.size()), super0 = uint32_t(into.SuperFirst.size());
        into.Perm.insert(into.Perm.end(), part.Perm.begin(), part.Perm.end());
        for (const uint32_t f : part.SuperFirst) into.SuperFirst.push_back(f + node0);
        for (const uint32_t p : part.Parent) into.Parent.push_back(p == UINT32_MAX ? UINT32_MAX : p + super0);
        for (const uint32_t c : part.Chain) into.Chain.push_back(c + into.Chains);
        into.Chains += part.Chains;
        for (const uint32_t t : part.Tops) into.Tops.push_back(t + super0);
    }

    // Emits `nodes` as a chain of panels on top of the subtrees whose tops `sub` holds; the chain's top becomes the only top.
    static void EmitChain(Sub &sub, const std::vector<uint32_t> &nodes, uint32_t panel) {
        const uint32_t first_panel = uint32_t(sub.SuperFirst.size());
        for (size_t at = 0; at < nodes.size(); at += panel) {
            const size_t end = std::min(nodes.size(), at + panel);
            const uint32_t id = uint32_t(sub.SuperFirst.size());
            sub.SuperFirst.push_back(uint32_t(sub.Perm.size()));
            sub.Parent.push_back(UINT32_MAX);
            sub.Chain.push_back(sub.Chains);
            if (id > first_panel) sub.Parent[id - 1] = id;
            sub.Perm.insert(sub.Perm.end(), nodes.begin() + at, nodes.begin() + end);
        }
        ++sub.Chains;
        for (const uint32_t c : sub.Tops) sub.Parent[c] = first_panel;
        sub.Tops.assign(1, uint32_t(sub.SuperFirst.size()) - 1);
    }

    // Orders the subdomain `nodes`.
    Sub Dissect(std::vector<uint32_t> &nodes, int depth) {
        Sub out;
        if (nodes.empty()) return out;
        if (nodes.size() <= Opt.LeafNodes) {
            EmitChain(out, nodes, std::max(Opt.LeafNodes, Opt.PanelNodes));
            return out;
        }
        float lo[3]{1e30f, 1e30f, 1e30f}, hi[3]{-1e30f, -1e30f, -1e30f};
        for (uint32_t v : nodes)
            for (int a = 0; a < 3; ++a) {
                lo[a] = std::min(lo[a], Xyz[3 * v + a]);
                hi[a] = std::max(hi[a], Xyz[3 * v + a]);
            }
        int axis = 0;
        for (int a = 1; a < 3; ++a)
            if (hi[a] - lo[a] > hi[axis] - lo[axis]) axis = a;
        // Median coordinate, then split by VALUE so that a grid plane stays on one side.
        std::vector<float> coord(nodes.size());
        for (size_t i = 0; i < nodes.size(); ++i) coord[i] = Xyz[3 * nodes[i] + axis];
        std::vector<float> sorted = coord;
        std::nth_element(sorted.begin(), sorted.begin() + sorted.size() / 2, sorted.end());
        const float t = sorted[sorted.size() / 2];
        std::vector<uint32_t> left, right;
        for (size_t i = 0; i < nodes.size(); ++i) (coord[i] < t ? left : right).push_back(nodes[i]);
        if (left.empty() || right.empty()) {
            left.clear(), right.clear();
            for (size_t i = 0; i < nodes.size(); ++i) (coord[i] <= t ? left : right).push_back(nodes[i]);
        }
        if (left.empty() || right.empty()) { // every node on one plane along the longest axis: split by position in the list
            left.assign(nodes.begin(), nodes.begin() + nodes.size() / 2);
            right.assign(nodes.begin() + nodes.size() / 2, nodes.end());
        }
        const uint32_t stamp_l = NextStamp.fetch_add(3, std::memory_order_relaxed), stamp_r = stamp_l + 1, stamp_s = stamp_l + 2;
        for (uint32_t v : left) Dom[v].store(stamp_l, std::memory_order_relaxed);
        for (uint32_t v : right) Dom[v].store(stamp_r, std::memory_order_relaxed);
        auto boundary = [&](const std::vector<uint32_t> &side, uint32_t other) {
            std::vector<uint32_t> found;
            for (uint32_t v : side)
                for (uint32_t j = RowPtr[v]; j < RowPtr[v + 1]; ++j)
                    if (Dom[Col[j]].load(std::memory_order_relaxed) == other) {
                        found.push_back(v);
                        break;
                    }
            return found;
        };
        std::vector<uint32_t> sep_l = boundary(left, stamp_r), sep_r = boundary(right, stamp_l);
        const bool take_left = sep_l.size() < sep_r.size();
        std::vector<uint32_t> &sep = take_left ? sep_l : sep_r;
        std::vector<uint32_t> &cut_side = take_left ? left : right;
        for (uint32_t v : sep) Dom[v].store(stamp_s, std::memory_order_relaxed);
        std::vector<uint32_t> rest;
        rest.reserve(cut_side.size() - sep.size());
        for (uint32_t v : cut_side)
            if (Dom[v].load(std::memory_order_relaxed) != stamp_s) rest.push_back(v);
        cut_side.swap(rest);
        nodes.clear();
        nodes.shrink_to_fit();
        if (depth < kForkDepth && std::min(left.size(), right.size()) >= 4096) {
            auto other = std::async(std::launch::async, [&] { return Dissect(left, depth + 1); });
            Sub r = Dissect(right, depth + 1);
            out = other.get();
            Append(out, std::move(r));
        } else {
            out = Dissect(left, depth + 1);
            Append(out, Dissect(right, depth + 1));
        }
        if (!sep.empty()) EmitChain(out, sep, Opt.PanelNodes);
        return out;
    }
};
} // namespace

namespace {
void BuildSchedules(Symbolic &sym);
}

void Symbolic::WaitSchedules() {
    if (ScheduleThread.joinable()) ScheduleThread.join();
    if (ScheduleFailed) throw std::bad_alloc(); // the only way building the schedules fails
}

Symbolic Analyse(uint32_t n, const uint32_t *rowptr, const uint32_t *col, const float *xyz, const SymbolicOptions &opt) {
    Symbolic sym;
    AnalyseInto(sym, n, rowptr, col, xyz, opt, false);
    return sym;
}

void AnalyseInto(Symbolic &sym, uint32_t n, const uint32_t *rowptr, const uint32_t *col, const float *xyz, const SymbolicOptions &opt, bool schedules_in_background) {
    sym.WaitSchedules();
    sym = Symbolic{};
    sym.NodeCount = n;
    const double t0 = Now();
    Dissector d{n, rowptr, col, xyz, opt};
    d.Dom.reset(new std::atomic<uint32_t>[n]);
    for (uint32_t i = 0; i < n; ++i) d.Dom[i].store(0, std::memory_order_relaxed);
    Dissector::Sub tree;
    {
        std::vector<uint32_t> all(n);
        std::iota(all.begin(), all.end(), 0u);
        tree = d.Dissect(all, 0);
    }
    sym.Perm = std::move(tree.Perm);
    sym.InvPerm.assign(n, 0);
    for (uint32_t i = 0; i < n; ++i) sym.InvPerm[sym.Perm[i]] = i;
    const uint32_t ns = uint32_t(tree.SuperFirst.size());
    sym.NumSuper = ns;
    sym.SuperFirst = std::move(tree.SuperFirst);
    sym.SuperFirst.push_back(n);
    sym.Parent = std::move(tree.Parent);
    const std::vector<uint32_t> chain = std::move(tree.Chain);
    for (auto &p : sym.Parent)
        if (p == UINT32_MAX) p = ns;
    sym.NodeSuper.assign(n, 0);
    for (uint32_t s = 0; s < ns; ++s)
        for (uint32_t v = sym.SuperFirst[s]; v < sym.SuperFirst[s + 1]; ++v) sym.NodeSuper[v] = s;
    // Macro blocks of the panel sweeps: every chain of c > 1 panels is cut into ceil(c / MacroPanels) runs of near-equal length.
    sym.MacroBackward = opt.MacroBackward;
    sym.MacroFirst.resize(ns), sym.MacroLast.resize(ns);
    for (uint32_t s = 0; s < ns;) {
        uint32_t e = s + 1;
        while (e < ns && chain[e] == chain[s]) ++e;
        const uint32_t panels = e - s, per = std::max(1u, opt.MacroPanels), blocks = (panels + per - 1) / per;
        for (uint32_t b = 0; b < blocks; ++b) {
            const uint32_t first = s + uint32_t(uint64_t(panels) * b / blocks), last = s + uint32_t(uint64_t(panels) * (b + 1) / blocks) - 1;
            for (uint32_t i = first; i <= last; ++i) sym.MacroFirst[i] = first, sym.MacroLast[i] = last;
        }
        s = e;
    }
    const double t1 = Now();
    sym.OrderingSeconds = t1 - t0;

    // Structures, children before parents (supernode ids are in elimination order, so ascending works).
    std::vector<std::vector<uint32_t>> children(ns);
    for (uint32_t s = 0; s < ns; ++s)
        if (sym.Parent[s] < ns) children[sym.Parent[s]].push_back(s);
    // Levels (height above the leaves) first: the supernodes of one level only read the structures of lower levels, so
    // a level is processed by several host threads at once, each with its own marker array.
    sym.Level.assign(ns, 0);
    for (uint32_t s = 0; s < ns; ++s)
        if (sym.Parent[s] < ns) sym.Level[sym.Parent[s]] = std::max(sym.Level[sym.Parent[s]], sym.Level[s] + 1);
    sym.NumLevels = ns ? *std::max_element(sym.Level.begin(), sym.Level.end()) + 1 : 0;
    sym.LevelPtr.assign(size_t(sym.NumLevels) + 1, 0);
    for (uint32_t s = 0; s < ns; ++s) ++sym.LevelPtr[sym.Level[s] + 1];
    for (uint32_t l = 0; l < sym.NumLevels; ++l) sym.LevelPtr[l + 1] += sym.LevelPtr[l];
    sym.LevelOrder.assign(ns, 0);
    {
        std::vector<uint32_t> at(sym.LevelPtr.begin(), sym.LevelPtr.end() - 1);
        for (uint32_t s = 0; s < ns; ++s) sym.LevelOrder[at[sym.Level[s]]++] = s;
    }
    std::vector<std::vector<uint32_t>> rows_of(ns);
    {
        std::vector<std::vector<uint32_t>> marks(16);
        static const bool check_followers = std::getenv("ME_SYMBOLIC_CHECK") != nullptr; // (debug: every shortcut structure against the general rule)
        const auto structure = [&](uint32_t s, std::vector<uint32_t> &mark) {
            const uint32_t first = sym.SuperFirst[s], last = sym.SuperFirst[s + 1];
            std::vector<uint32_t> &r = rows_of[s];
            if (s > 0 && chain[s - 1] == chain[s] && !check_followers) {
                // A later panel of a chain has one child, the panel before it, whose sorted structure already holds the rest of the
                // chain and everything inherited from below: this panel's structure is that list past its own nodes, plus whatever
                // its own adjacency adds (rarely anything). No marks, no sort: on the big separators this was most of the analysis.
                const std::vector<uint32_t> &below = rows_of[s - 1];
                r.assign(std::lower_bound(below.begin(), below.end(), last), below.end());
                std::vector<uint32_t> extra;
                for (uint32_t v = first; v < last; ++v) {
                    const uint32_t old = sym.Perm[v];
                    for (uint32_t j = rowptr[old]; j < rowptr[old + 1]; ++j) {
                        const uint32_t u = sym.InvPerm[col[j]];
                        if (u >= last && !std::binary_search(r.begin(), r.end(), u)) extra.push_back(u);
                    }
                }
                if (!extra.empty()) {
                    std::sort(extra.begin(), extra.end());
                    extra.erase(std::unique(extra.begin(), extra.end()), extra.end());
                    std::vector<uint32_t> merged(r.size() + extra.size());
                    std::merge(r.begin(), r.end(), extra.begin(), extra.end(), merged.begin());
                    r.swap(merged);
                }
                return;
            }
            for (uint32_t v = first; v < last; ++v) {
                const uint32_t old = sym.Perm[v];
                for (uint32_t j = rowptr[old]; j < rowptr[old + 1]; ++j) {
                    const uint32_t u = sym.InvPerm[col[j]];
                    if (u >= last && mark[u] != s) {
                        mark[u] = s;
                        r.push_back(u);
                    }
                }
            }
            for (uint32_t c : children[s])
                for (const uint32_t u : rows_of[c])
                    if (u >= last && mark[u] != s) {
                        mark[u] = s;
                        r.push_back(u);
                    }
            // A panel of a chain is coupled to the rest of its separator: the separator is a clique after elimination.
            for (uint32_t e = s + 1; e < ns && chain[e] == chain[s]; ++e)
                for (uint32_t u = sym.SuperFirst[e]; u < sym.SuperFirst[e + 1]; ++u)
                    if (mark[u] != s) {
                        mark[u] = s;
                        r.push_back(u);
                    }
            std::sort(r.begin(), r.end());
        };
        // Supernode ids are in post-order, so a subtree is a contiguous id range processed in ascending order. The tree is
        // cut below its top separators into a few dozen subtrees that host threads take from a shared counter; the top
        // supernodes follow sequentially.
        std::vector<uint32_t> subtree(ns, 1);
        for (uint32_t s = 0; s < ns; ++s)
            if (sym.Parent[s] < ns) subtree[sym.Parent[s]] += subtree[s];
        std::vector<uint32_t> cut, top;
        for (uint32_t s = 0; s < ns; ++s)
            if (sym.Parent[s] >= ns) cut.push_back(s);
        while (cut.size() < 48) {
            size_t big = cut.size();
            for (size_t i = 0; i < cut.size(); ++i)
                if (subtree[cut[i]] > 128 && !children[cut[i]].empty() && (big == cut.size() || subtree[cut[i]] > subtree[cut[big]])) big = i;
            if (big == cut.size()) break;
            const uint32_t root = cut[big];
            cut.erase(cut.begin() + big);
            top.push_back(root);
            cut.insert(cut.end(), children[root].begin(), children[root].end());
        }
        std::sort(cut.begin(), cut.end(), [&](uint32_t a, uint32_t b) { return subtree[a] > subtree[b]; });
        std::sort(top.begin(), top.end());
        std::atomic<size_t> next{0};
        const size_t workers = std::max<size_t>(1, std::min<size_t>({16, std::thread::hardware_concurrency(), cut.size()}));
        const auto work = [&](size_t worker) {
            std::vector<uint32_t> &mark = marks[worker];
            mark.assign(n, UINT32_MAX);
            for (size_t i = next++; i < cut.size(); i = next++)
                for (uint32_t s = cut[i] + 1 - subtree[cut[i]]; s <= cut[i]; ++s) structure(s, mark);
        };
        if (workers <= 1 || ns < 512) {
            work(0);
        } else {
            std::vector<std::thread> pool;
            for (size_t w = 0; w < workers; ++w) pool.emplace_back(work, w);
            for (auto &t : pool) t.join();
        }
        if (marks[0].empty()) marks[0].assign(n, UINT32_MAX);
        const double t_subtrees = Now();
        for (const uint32_t s : top) structure(s, marks[0]);
        if (std::getenv("ME_SYMBOLIC_TIMING")) fprintf(stderr, "[me] symbolic rows: subtrees %.1f ms on %zu threads, %zu top supernodes %.1f ms\n", 1e3 * (t_subtrees - t1), workers, top.size(), 1e3 * (Now() - t_subtrees));
    }
    sym.RowPtr.assign(size_t(ns) + 1, 0);
    for (uint32_t s = 0; s < ns; ++s) sym.RowPtr[s + 1] = sym.RowPtr[s] + rows_of[s].size();
    sym.Rows.resize(sym.RowPtr[ns]);
    ParallelChunks(ns, 256, [&](size_t a, size_t b, size_t) {
        for (size_t s = a; s < b; ++s) std::copy(rows_of[s].begin(), rows_of[s].end(), sym.Rows.begin() + sym.RowPtr[s]);
    });
    rows_of.clear();

    const bool timing = std::getenv("ME_SYMBOLIC_TIMING") != nullptr;
    double tm = Now();
    const auto lap = [&](const char *what) {
        if (timing) fprintf(stderr, "[me] symbolic %-12s %.1f ms\n", what, 1e3 * (Now() - tm));
        tm = Now();
    };
    if (timing) fprintf(stderr, "[me] symbolic %-12s %.1f ms\n", "rows", 1e3 * (tm - t1));
    // Panel offsets, segments, work lists.
    sym.PanelOffset.assign(size_t(ns) + 1, 0);
    sym.InvOffset.assign(size_t(ns) + 1, 0);
    sym.SlabOffset.assign(size_t(ns) + 1, 0);
    sym.SegPtr.assign(size_t(ns) + 1, 0);
    for (uint32_t s = 0; s < ns; ++s) {
        const uint64_t k = 3ull * (sym.SuperFirst[s + 1] - sym.SuperFirst[s]), m = 3ull * (sym.RowPtr[s + 1] - sym.RowPtr[s]);
        sym.PanelOffset[s + 1] = sym.PanelOffset[s] + (k + m) * k;
        sym.InvOffset[s + 1] = sym.InvOffset[s] + k * k;
        sym.SlabOffset[s + 1] = sym.SlabOffset[s] + ((m + kSolveRows - 1) / kSolveRows) * kSolveRows * ((k + 7) / 8 * 8);
        sym.FactorFlops += double(k) * k * k / 3.0 + double(m) * k * k + double(m) * m * k;
        sym.MaxPanelColumns = std::max<uint32_t>(sym.MaxPanelColumns, uint32_t(k));
        sym.MaxPanelRows = std::max<uint32_t>(sym.MaxPanelRows, uint32_t(m));
        const uint64_t r0 = sym.RowPtr[s], r1 = sym.RowPtr[s + 1];
        for (uint64_t j = r0; j < r1;) {
            const uint32_t target = sym.NodeSuper[sym.Rows[j]];
            uint64_t e = j + 1;
            while (e < r1 && sym.NodeSuper[sym.Rows[e]] == target) ++e;
            sym.SegTarget.push_back(target);
            sym.SegBegin.push_back(uint32_t(j - r0));
            sym.SegEnd.push_back(uint32_t(e - r0));
            j = e;
        }
        sym.SegPtr[s + 1] = sym.SegTarget.size();
    }
    sym.FactorNonZeros = sym.PanelOffset[ns];
    sym.MacroOffset.assign(size_t(ns) + 1, 0);
    sym.MacroOffsetT.assign(size_t(ns) + 1, 0);
    for (uint32_t s = 0; s < ns; ++s) {
        uint64_t fwd = 0, bwd = 0;
        if (sym.MacroFirst[s] != sym.MacroLast[s]) {
            const uint64_t k = 3ull * (sym.SuperFirst[s + 1] - sym.SuperFirst[s]);
            fwd = k * 3ull * (sym.SuperFirst[s + 1] - sym.SuperFirst[sym.MacroFirst[s]]);
            bwd = k * 3ull * (sym.SuperFirst[sym.MacroLast[s] + 1] - sym.SuperFirst[s]);
            for (uint32_t slice = 0; slice * kTile < k; ++slice) sym.MacroJobs.push_back({sym.MacroFirst[s], sym.MacroLast[s], s, slice});
        }
        sym.MacroOffset[s + 1] = sym.MacroOffset[s] + fwd;
        sym.MacroOffsetT[s + 1] = sym.MacroOffsetT[s] + bwd;
    }
    // (the first columns of a macro block carry the longest chains of products: they go first)
    std::stable_sort(sym.MacroJobs.begin(), sym.MacroJobs.end(), [](const Symbolic::MacroJob &a, const Symbolic::MacroJob &b) { return a.Last - a.Column > b.Last - b.Column; });
    lap("segments");
    // The solve schedules read only what exists by now, and nothing below writes it: they are built on a thread of their own
    // beside the tile lists, and (in the background form) on past this function's return.
    sym.ScheduleThread = std::thread([&sym] {
        try {
            BuildSchedules(sym);
        } catch (...) {
            sym.ScheduleFailed = true;
        }
    });

    // Tile lists in (level, order inside the level) sequence: count per supernode, prefix, then fill on several threads.
    sym.PanelTilePtr.assign(size_t(sym.NumLevels) + 1, 0);
    sym.UpdateTilePtr.assign(size_t(sym.NumLevels) + 1, 0);
    {
        std::vector<uint64_t> panel_at(size_t(ns) + 1, 0), update_at(size_t(ns) + 1, 0);
        const auto tiles_of = [&](uint32_t s, PanelTile *panel_out, UpdateTile *update_out, uint64_t &panels, uint64_t &updates) {
            const uint32_t m = 3 * uint32_t(sym.RowPtr[s + 1] - sym.RowPtr[s]);
            panels = updates = 0;
            if (!panel_out) panels = (m + kTile - 1) / kTile;
            else
                for (uint32_t t = 0; t * kTile < m; ++t, ++panels) panel_out[panels] = {s, t};
            for (uint64_t g = sym.SegPtr[s]; g < sym.SegPtr[s + 1]; ++g) {
                const uint32_t c0 = 3 * sym.SegBegin[g], c1 = 3 * sym.SegEnd[g];
                const uint32_t col_tiles = (c1 - c0 + kTile - 1) / kTile, row_tiles = (m - c0 + kTile - 1) / kTile;
                for (uint32_t ct = 0; ct < col_tiles; ++ct) {
                    if (!update_out) { // (the counting pass needs no walk over the tiles)
                        updates += row_tiles > ct ? row_tiles - ct : 0;
                        continue;
                    }
                    for (uint32_t rt = ct; rt < row_tiles; ++rt, ++updates) update_out[updates] = {s, uint32_t(g), uint16_t(rt), uint16_t(ct)};
                }
            }
        };
        ParallelChunks(ns, 256, [&](size_t a, size_t b, size_t) {
            for (size_t i = a; i < b; ++i) tiles_of(sym.LevelOrder[i], nullptr, nullptr, panel_at[i + 1], update_at[i + 1]);
        });
        for (uint32_t i = 0; i < ns; ++i) panel_at[i + 1] += panel_at[i], update_at[i + 1] += update_at[i];
        sym.PanelTiles.resize(panel_at[ns]);
        sym.UpdateTiles.resize(update_at[ns]);
        ParallelChunks(ns, 256, [&](size_t a, size_t b, size_t) {
            uint64_t panels, updates;
            for (size_t i = a; i < b; ++i) tiles_of(sym.LevelOrder[i], sym.PanelTiles.data() + panel_at[i], sym.UpdateTiles.data() + update_at[i], panels, updates);
        });
        for (uint32_t l = 0; l < sym.NumLevels; ++l) sym.PanelTilePtr[l + 1] = panel_at[sym.LevelPtr[l + 1]], sym.UpdateTilePtr[l + 1] = update_at[sym.LevelPtr[l + 1]];
    }
    lap("tiles");
    if (!schedules_in_background) {
        sym.WaitSchedules();
        lap("sweep tasks");
        if (timing) fprintf(stderr, "[me] symbolic levels %u, panel-sweep levels %u, macro jobs %zu, macro inverse doubles %llu\n", sym.NumLevels, sym.SweepLevels, sym.MacroJobs.size(), (unsigned long long)sym.MacroOffset[ns]);
    }
    sym.StructureSeconds = Now() - t1;
}

namespace {
void BuildSchedules(Symbolic &sym) {
    const uint32_t ns = sym.NumSuper;
    // Dataflow schedules of the solves. Ticket order = level order (height above the leaves), NOT elimination order:
    // both are topological, but the post-order would walk one subtree's separator chains at a time, while the level
    // order keeps the chains of all subtrees of the same height in flight together.
    auto columns = [&](uint32_t s) { return 3 * (sym.SuperFirst[s + 1] - sym.SuperFirst[s]); };
    auto diag_slabs = [&](uint32_t s) { return (columns(s) + kSolveRows - 1) / kSolveRows; };
    auto base_task = [&](uint32_t s) {
        SweepTask t{};
        t.Super = s;
        t.K = columns(s);
        t.VecOffset = 3 * sym.SuperFirst[s];
        t.RowsBase = uint32_t(sym.RowPtr[s]);
        t.Count = 1;
        return t;
    };
    // run_of(panel slabs of a level) = slabs per panel task on that level; max_links bounds the arrivals one forward task may owe.
    // With `macros` the supernodes of a macro block (Symbolic::MacroFirst) share the level of the block's first panel: their diagonal
    // slabs become slabs of the block's inverse, and their panel slabs cover only the rows outside the block.
    auto make_schedules = [&](bool macros, bool forward_slabs, auto &&run_of, uint32_t max_links, std::vector<SweepTask> &fwd, std::vector<uint32_t> &fwd_links, std::vector<SweepTask> &bwd,
                              std::vector<uint32_t> &bwd_links, std::vector<uint32_t> &bwd_link_need, std::vector<uint32_t> &fwd_expected, std::vector<uint32_t> &bwd_expected, uint32_t *levels_used) {
        auto first_of = [&](uint32_t s) { return macros ? sym.MacroFirst[s] : s; };
        auto last_of = [&](uint32_t s) { return macros ? sym.MacroLast[s] : s; };
        // rows at the top of s's panel that belong to its own macro block (the chain's next panels: the first rows of the list)
        auto inside = [&](uint32_t s) { return 3 * (sym.SuperFirst[last_of(s) + 1] - sym.SuperFirst[s + 1]); };
        auto panel_rows = [&](uint32_t s) { return 3 * uint32_t(sym.RowPtr[s + 1] - sym.RowPtr[s]); };
        // Panel tasks cover whole 32-row slabs of the slab copy; the first one of a macro panel starts inside the block's rows.
        auto first_slab_of = [&](uint32_t s) { return inside(s) / kSolveRows; };
        auto panel_slabs = [&](uint32_t s) { return (panel_rows(s) + kSolveRows - 1) / kSolveRows - (panel_rows(s) > inside(s) ? first_slab_of(s) : (panel_rows(s) + kSolveRows - 1) / kSolveRows); };
        auto slab_doubles = [&](uint32_t s) { return uint64_t(kSolveRows) * ((columns(s) + 7) / 8 * 8); };
        // The supernodes owning slabs [first_slab, first_slab + count) of s's panel (counted from the slab holding the first row
        // outside the macro block; rows inside it are left out), ascending (rows are), without repeats.
        auto targets_of = [&](uint32_t s, uint32_t first_slab, uint32_t count, std::vector<uint32_t> &out, size_t from) {
            const uint64_t r0 = sym.RowPtr[s], nodes = sym.RowPtr[s + 1] - r0, skip = inside(s), slab0 = first_slab_of(s);
            const uint64_t first = std::max<uint64_t>(skip, (slab0 + first_slab) * kSolveRows) / 3, last = std::min<uint64_t>(nodes, ((slab0 + first_slab + count) * kSolveRows + 2) / 3);
            for (uint64_t j = first; j < last; ++j) {
                const uint32_t target = sym.NodeSuper[sym.Rows[r0 + j]];
                if (out.size() == from || out.back() != target) out.push_back(target);
            }
        };
        auto diag_task = [&](uint32_t s, uint32_t h, bool backward) {
            SweepTask t = base_task(s);
            t.Row0 = h * kSolveRows, t.Ld = t.K;
            if (first_of(s) == last_of(s)) {
                t.Kind = 0, t.Base = sym.InvOffset[s], t.Limit = t.K, t.LinkCount = 1;
            } else if (!backward) {
                t.Kind = 2, t.Base = sym.MacroOffset[s], t.DiagColumn = 3 * (sym.SuperFirst[s] - sym.SuperFirst[first_of(s)]), t.Limit = t.DiagColumn + t.K, t.LinkCount = s - first_of(s) + 1;
            } else {
                t.Kind = 2, t.Base = sym.MacroOffsetT[s], t.Limit = 3 * (sym.SuperFirst[last_of(s) + 1] - sym.SuperFirst[s]), t.LinkCount = last_of(s) - s + 1;
            }
            return t;
        };
        // Levels of this schedule: a supernode sits on the level of its macro block's first panel (ascending ids inside a level).
        std::vector<uint32_t> level_ptr(size_t(sym.NumLevels) + 1, 0), order(ns), level_slabs(sym.NumLevels, 0);
        for (uint32_t s = 0; s < ns; ++s) ++level_ptr[sym.Level[first_of(s)] + 1], level_slabs[sym.Level[first_of(s)]] += panel_slabs(s);
        for (uint32_t l = 0; l < sym.NumLevels; ++l) level_ptr[l + 1] += level_ptr[l];
        {
            std::vector<uint32_t> at(level_ptr.begin(), level_ptr.end() - 1);
            for (uint32_t s = 0; s < ns; ++s) order[at[sym.Level[first_of(s)]]++] = s;
        }
        if (levels_used) {
            *levels_used = 0;
            for (uint32_t l = 0; l < sym.NumLevels; ++l) *levels_used += level_ptr[l + 1] > level_ptr[l];
        }
        fwd_expected.assign(ns, 0), bwd_expected.assign(ns, 0);
        size_t tasks = 0;
        for (uint32_t s = 0; s < ns; ++s) tasks += diag_slabs(s) + panel_slabs(s);
        fwd.reserve(tasks), bwd.reserve(tasks);
        fwd_links.reserve(sym.Rows.size() / 2), bwd_links.reserve(sym.Rows.size() / 2), bwd_link_need.reserve(sym.Rows.size() / 2);
        for (uint32_t l = 0; l < sym.NumLevels; ++l) {
            const uint32_t run = std::max(1u, run_of(level_slabs[l]));
            for (uint32_t i = level_ptr[l]; i < level_ptr[l + 1]; ++i)
                for (uint32_t h = 0; h < diag_slabs(order[i]); ++h) fwd.push_back(diag_task(order[i], h, false));
            for (uint32_t i = level_ptr[l]; i < level_ptr[l + 1]; ++i) {
                const uint32_t s = order[i], m = panel_rows(s), slabs = panel_slabs(s), skip = inside(s), slab0 = first_slab_of(s);
                for (uint32_t h = 0; h < slabs;) {
                    SweepTask t = base_task(s);
                    t.Kind = 1, t.Limit = m, t.Row0 = (slab0 + h) * kSolveRows, t.RowMin = skip, t.Need = diag_slabs(s);
                    if (forward_slabs) t.Base = sym.SlabOffset[s] + (slab0 + h) * slab_doubles(s), t.Ld = t.K;
                    else t.Base = sym.PanelOffset[s] + t.K, t.Ld = t.K + m;
                    t.LinkBegin = uint32_t(fwd_links.size());
                    // as many slabs as the run allows while the arrivals owed still fit the publishing warp
                    uint32_t count = 0;
                    while (count < run && h + count < slabs) {
                        const size_t before = fwd_links.size();
                        targets_of(s, h + count, 1, fwd_links, t.LinkBegin);
                        if (count > 0 && fwd_links.size() - t.LinkBegin > max_links) {
                            fwd_links.resize(before);
                            break;
                        }
                        ++count;
                    }
                    t.Count = count;
                    t.LinkCount = uint32_t(fwd_links.size()) - t.LinkBegin;
                    for (uint32_t j = 0; j < t.LinkCount; ++j) ++fwd_expected[fwd_links[t.LinkBegin + j]];
                    fwd.push_back(t);
                    h += count;
                }
            }
        }
        for (auto &t : fwd)
            if (t.Kind == 0) t.Need = fwd_expected[t.Super];
        for (uint32_t l = sym.NumLevels; l-- > 0;) {
            const uint32_t run = std::max(1u, run_of(level_slabs[l]));
            for (uint32_t i = level_ptr[l]; i < level_ptr[l + 1]; ++i) {
                const uint32_t s = order[i], m = panel_rows(s), slabs = panel_slabs(s), skip = inside(s), slab0 = first_slab_of(s);
                for (uint32_t h = 0; h < slabs; h += run) {
                    SweepTask t = base_task(s);
                    t.Kind = 1, t.Base = sym.SlabOffset[s] + (slab0 + h) * slab_doubles(s), t.Limit = m, t.Ld = t.K, t.Row0 = (slab0 + h) * kSolveRows, t.RowMin = skip;
                    t.Count = std::min(run, slabs - h);
                    t.LinkBegin = uint32_t(bwd_links.size());
                    targets_of(s, h, t.Count, bwd_links, t.LinkBegin);
                    t.LinkCount = uint32_t(bwd_links.size()) - t.LinkBegin;
                    for (uint32_t j = 0; j < t.LinkCount; ++j) bwd_link_need.push_back(diag_slabs(bwd_links[t.LinkBegin + j]));
                    ++bwd_expected[s];
                    bwd.push_back(t);
                }
            }
            for (uint32_t i = level_ptr[l]; i < level_ptr[l + 1]; ++i)
                for (uint32_t h = 0; h < diag_slabs(order[i]); ++h) {
                    SweepTask t = diag_task(order[i], h, true);
                    t.Need = bwd_expected[t.Super];
                    bwd.push_back(t);
                }
        }
    };
    // Single-vector sweeps: one slab per task (on a thread of their own: the two pairs of schedules share nothing but their inputs).
    bool single_failed = false;
    std::thread single([&] {
        try {
            std::vector<uint32_t> fwd_need, bwd_need;
            make_schedules(false, false, [](uint32_t) { return 1u; }, UINT32_MAX, sym.FwdTasks, sym.FwdLinks, sym.BwdTasks, sym.BwdLinks, sym.BwdLinkNeed, fwd_need, bwd_need, nullptr);
        } catch (...) {
            single_failed = true;
        }
    });
    struct Join { // (an exception below must not leave the thread running)
        std::thread &T;
        ~Join() {
            if (T.joinable()) T.join();
        }
    } join{single};
    // Panel sweeps: macro blocks, and runs on the levels wide enough to still hand every resident CTA (2 per SM x 148 SMs) a run of its own.
    // The two directions are scheduled separately when only the forward sweep takes macro blocks (SymbolicOptions::MacroBackward).
    {
        constexpr uint32_t resident = 2 * 148;
        const auto run_of = [&](uint32_t slabs) { return std::min(kWideRun, slabs / resident); };
        if (sym.MacroBackward) {
            make_schedules(true, true, run_of, kWideRunLinks, sym.WideFwdTasks, sym.WideFwdLinks, sym.WideBwdTasks, sym.WideBwdLinks, sym.WideBwdLinkNeed, sym.WideFwdNeed, sym.WideBwdNeed, &sym.SweepLevels);
        } else {
            std::vector<SweepTask> unused_tasks;
            std::vector<uint32_t> unused_links, unused_link_need, unused_need, unused_fwd_links;
            make_schedules(true, true, run_of, kWideRunLinks, sym.WideFwdTasks, sym.WideFwdLinks, unused_tasks, unused_links, unused_link_need, sym.WideFwdNeed, unused_need, &sym.SweepLevels);
            unused_tasks.clear(), unused_need.clear();
            make_schedules(false, true, run_of, kWideRunLinks, unused_tasks, unused_fwd_links, sym.WideBwdTasks, sym.WideBwdLinks, sym.WideBwdLinkNeed, unused_need, sym.WideBwdNeed, nullptr);
        }
    }
    single.join();
    if (single_failed) throw std::bad_alloc();
}
} // namespace

} // namespace me
