// astar_core — the A* of goal planning on flat per-search scratch arrays, written once for the device (k_goal_astar runs
// it on lane 0 of the agent's warp, the other lanes help with the row minimum) and for the host (compiled as plain C++
// into libhostgoal.so, where tests/test_goal_planning.py checks it path for path against the host planner, the oracle
// and the reference's own Astar-3D).
//
// Replaces src/Astar-3D/isearch.cpp:48-284 + astar.cpp:18-30 with the options GridBasedPlanner::planAstar passes
// (src/grid_based_planner.cpp:266-303, environmentoptions.cpp:13-21: Euclidean heuristic, 6-connected, unit cost,
// hweight 1, g-max tie break). The reference keeps one std::unordered_map per grid row and breaks (F, g) ties inside a
// row by that container's ITERATION ORDER (isearch.cpp:209-242), so the search carries an explicit model of libstdc++'s
// hash table node order (one forward list threaded through `next`, every bucket a contiguous run, new nodes at the front
// of their bucket's run or of the list, a rehash re-threads the list in iteration order; keys hash to themselves).
// The bucket counts the container moves through (1 -> 13 -> 29 -> 59 -> ...) depend only on the element count: the
// engine records them once from libstdc++'s own growth policy (goal_bucket_sequence in engine.cu) and hands the
// sequence in. Same statement as host/grid_based_planner.hpp's HashOrderModel, without containers.
//
// Storage. Per cell: one byte (search state, occupancy, the move that reached it), g and the list link — F is never
// stored, it is g + the Euclidean distance of the cell to the goal, recomputed (one DSQRT) where the reference reads it.
// The index type I is uint16_t when the grid has fewer than 65 534 cells: 5 bytes per cell, and the whole search state of
// the shipped 10 x 10 x 2.5 m world at grid/resolution 0.25 (18 491 cells + 41 row containers of <= 541 buckets = 137 KB)
// fits one SM's shared memory, where a dependent access costs ~30 cycles instead of an L2 round trip; larger grids use
// int32 indices in global memory.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define ASTAR_HD __host__ __device__ __forceinline__
#else
#define ASTAR_HD inline
#include <cmath>
#endif

// section timers of one expansion (debug builds of the kernel: -DLSCGPU_GOAL_TIMERS)
#if defined(LSCGPU_GOAL_TIMERS) && defined(__CUDA_ARCH__)
#define ASTAR_TICK(k) do { const long long t__ = clock64(); c.tsec[k] += t__ - c.tlast; c.tlast = t__; } while (0)
#else
#define ASTAR_TICK(k) do { } while (0)
#endif

namespace lscgpu {

// cell byte: bits 0-1 search state (0 unseen, 1 open, 2 closed), bit 2 occupied, bits 3-5 move that reached the cell
// (0..5 = -i, -j, -z, +z, +j, +i; 7: none = the start)
constexpr int kCellOpen = 1, kCellClosed = 2, kCellStateMask = 3, kCellOccupied = 4, kCellDirShift = 3, kCellNoParent = 7;
constexpr int kAstarMaxLevels = 16;

template <typename I>
struct AstarCtx {
    static constexpr I kEnd = (I)-1;                    // list end / "before the first node" mark of a bucket
    static constexpr I kNone = (I)-2;                   // empty bucket
    uint8_t* cell;          // [H * W * A]
    I* g;                   // [H * W * A] unit edge costs: g is the step count
    I* next;                // [H * W * A] next node in the row container's iteration order (kEnd: none)
    I* bkt;                 // [H][bcap] bucket array of every row container: kNone, kEnd (= the run starts at the list head)
                            // or the node BEFORE the bucket's first node
    int bcap;
    // Order stamps (optional; null: off). The runs of the buckets lie in the list in the order in which the buckets became
    // non-empty, newest first (a node that enters an empty bucket goes to the list head; erasures and insertions into a
    // non-empty bucket move no run; a rehash re-threads in iteration order, again newest first). bstamp[row][bucket] = value
    // of the row's event counter when the bucket last became non-empty: of two nodes in different buckets the one with the
    // SMALLER stamp comes later in the iteration. This lets deleteMin's tie-break (the last of the minimal nodes in
    // iteration order) be decided without walking the list; only ties inside one bucket walk that bucket's short run.
    I* bstamp;              // [H][bcap]
    int* stamp;             // [H] event counter of the row
    // per grid row i: the container (head, element count, index into bkt_seq) and the cached row minimum (isearch.cpp's
    // `min` node per row)
    int *head, *count, *level, *min_cell, *min_g;
    double* min_f;
    int *jlo, *jhi;         // optional (null: off): range of j in which the row has ever had an open node — deleteMin's re-scan
                            // of the row's cells stops there
    int H, W, A;
    const int* bkt_seq;     // bucket counts the row containers move through: 1, 13, 29, 59, ... (kAstarMaxLevels entries)
    // Optional accelerators (0 / null: plain division, sqrt). Divisions by A, W and the bucket counts as one multiply-high
    // with ceil(2^32 / d) — exact while dividend * d < 2^32, i.e. for the 16-bit grids; h = sqrt(d2) from a table indexed
    // by the integer squared distance (the same correctly rounded doubles).
    unsigned magic_a, magic_w;
    const unsigned* bkt_magic;  // [kAstarMaxLevels]
    const double* sqrt_tab;     // [(H-1)^2 + (W-1)^2 + (A-1)^2 + 1]
    int gi, gj, gz;         // goal cell
    long long expansions;
    int open_size;
#ifdef LSCGPU_GOAL_TIMERS
    long long tsec[6], tlast;   // split, erase, rescan, neighbours, find-min, (spare)
#endif
};

ASTAR_HD unsigned astar_mulhi(unsigned a, unsigned b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (unsigned)(((unsigned long long)a * b) >> 32);
#endif
}
ASTAR_HD unsigned astar_div(unsigned p, unsigned d, unsigned magic) { return magic ? astar_mulhi(p, magic) : p / d; }
// ceil(2^32 / d), or 0 (= divide) when the product bound does not hold or d == 1
inline unsigned astar_magic(unsigned d, unsigned long long max_dividend) {
    if (d < 2 || max_dividend * d >= (1ull << 32)) return 0;
    return (unsigned)(((1ull << 32) + d - 1) / d);
}

template <typename I> ASTAR_HD int astar_cell(const AstarCtx<I>& c, int i, int j, int z) { return (i * c.W + j) * c.A + z; }
template <typename I> ASTAR_HD void astar_split(const AstarCtx<I>& c, int cell, int& i, int& j, int& z) {
    const int ij = (int)astar_div((unsigned)cell, (unsigned)c.A, c.magic_a);
    z = cell - ij * c.A;
    i = (int)astar_div((unsigned)ij, (unsigned)c.W, c.magic_w);
    j = ij - i * c.W;
}
// Node::get_id (node.cpp:12-14): the key of the row container
template <typename I> ASTAR_HD unsigned astar_key(const AstarCtx<I>& c, int cell) {
    int i, j, z;
    astar_split(c, cell, i, j, z);
    return (unsigned)(c.H * c.W * z + c.W * i + j);
}
// key % (bucket count of level lvl)
template <typename I> ASTAR_HD unsigned astar_bucket_of_key(const AstarCtx<I>& c, unsigned key, int lvl) {
    const unsigned n = (unsigned)c.bkt_seq[lvl];
    return key - astar_div(key, n, c.bkt_magic ? c.bkt_magic[lvl] : 0u) * n;
}
template <typename I> ASTAR_HD unsigned astar_bucket(const AstarCtx<I>& c, int cell, int lvl) {
    return astar_bucket_of_key(c, astar_key(c, cell), lvl);
}
template <typename I> ASTAR_HD double astar_h(const AstarCtx<I>& c, int d2) { return c.sqrt_tab ? c.sqrt_tab[d2] : sqrt((double)d2); }
template <typename I> ASTAR_HD double astar_heuristic(const AstarCtx<I>& c, int i, int j, int z) {
    const int di = c.gi - i, dj = c.gj - j, dz = c.gz - z;
    return astar_h(c, di * di + dj * dj + dz * dz);
}
// F of an open cell = g + h, the value the reference stores in the node
template <typename I> ASTAR_HD double astar_f(const AstarCtx<I>& c, int cell, int g) {
    int i, j, z;
    astar_split(c, cell, i, j, z);
    return (double)g + astar_heuristic(c, i, j, z);
}

template <typename I> ASTAR_HD void astar_rehash(AstarCtx<I>& c, int r, int lvl) {
    const int n = c.bkt_seq[lvl];
    constexpr I kEnd = AstarCtx<I>::kEnd, kNone = AstarCtx<I>::kNone;
    I* nb = c.bkt + (size_t)r * c.bcap;
    for (int b = 0; b < n; b++) nb[b] = kNone;
    int p = c.head[r];
    int head = -1;
    unsigned begin_bkt = 0;
    while (p >= 0) {
        const I nx = c.next[p];
        const unsigned b = astar_bucket(c, p, lvl);
        if (nb[b] == kNone) {
            c.next[p] = head < 0 ? kEnd : (I)head;
            if (head >= 0) nb[begin_bkt] = (I)p;
            head = p; nb[b] = kEnd;
            begin_bkt = b;
            if (c.bstamp) c.bstamp[(size_t)r * c.bcap + b] = (I)(++c.stamp[r]);
        } else if (nb[b] == kEnd) {
            c.next[p] = head < 0 ? kEnd : (I)head; head = p;
        } else {
            c.next[p] = c.next[nb[b]]; c.next[nb[b]] = (I)p;
        }
        p = nx == kEnd ? -1 : (int)nx;
    }
    c.head[r] = head;
}

// (j, z): the node's coordinates inside row r — the caller knows them, so the key needs no division
template <typename I> ASTAR_HD void astar_insert(AstarCtx<I>& c, int r, int nd, int j, int z) {
    constexpr I kEnd = AstarCtx<I>::kEnd, kNone = AstarCtx<I>::kNone;
    // _Prime_rehash_policy::_M_need_rehash with max_load_factor 1: grow when the new element count exceeds the bucket
    // count (the empty container has one bucket and always grows)
    const int lvl = c.level[r];
    const int cnt = c.count[r];
    if (cnt + 1 > (lvl == 0 ? 0 : c.bkt_seq[lvl])) {
        c.level[r] = lvl + 1;
        astar_rehash(c, r, lvl + 1);
    }
    const int cl = c.level[r];
    I* bk = c.bkt + (size_t)r * c.bcap;
    const unsigned b = astar_bucket_of_key(c, (unsigned)(c.H * c.W * z + c.W * r + j), cl);
    const I at = bk[b];
    const int old = c.head[r];
    if (c.jlo) { if (j < c.jlo[r]) c.jlo[r] = j; if (j > c.jhi[r]) c.jhi[r] = j; }
    if (at == kEnd) {
        c.next[nd] = old < 0 ? kEnd : (I)old; c.head[r] = nd;
    } else if (at != kNone) {
        c.next[nd] = c.next[at]; c.next[at] = (I)nd;
    } else {
        c.next[nd] = old < 0 ? kEnd : (I)old; c.head[r] = nd;
        if (old >= 0) bk[astar_bucket(c, old, cl)] = (I)nd;
        bk[b] = kEnd;
        if (c.bstamp) c.bstamp[(size_t)r * c.bcap + b] = (I)(++c.stamp[r]);
    }
    c.count[r] = cnt + 1;
}

template <typename I> ASTAR_HD void astar_erase(AstarCtx<I>& c, int r, int nd, int j, int z) {
    constexpr I kEnd = AstarCtx<I>::kEnd, kNone = AstarCtx<I>::kNone;
    const int cl = c.level[r];
    I* bk = c.bkt + (size_t)r * c.bcap;
    const unsigned b = astar_bucket_of_key(c, (unsigned)(c.H * c.W * z + c.W * r + j), cl);
    const I first_prev = bk[b];
    I prev = first_prev;                                    // kEnd: "before the list head"
    for (int p = prev == kEnd ? c.head[r] : (int)c.next[prev]; p != nd; p = (int)c.next[p]) prev = (I)p;
    const I nxt = c.next[nd];
    if (prev == first_prev) {                               // first node of its bucket
        const unsigned nb = nxt != kEnd ? astar_bucket(c, (int)nxt, cl) : 0u;
        if (nxt == kEnd || nb != b) {                       // the bucket becomes empty
            if (nxt != kEnd) bk[nb] = first_prev;
            bk[b] = kNone;
        }
    } else if (nxt != kEnd) {
        const unsigned nb = astar_bucket(c, (int)nxt, cl);
        if (nb != b) bk[nb] = prev;
    }
    if (prev == kEnd) c.head[r] = nxt == kEnd ? -1 : (int)nxt; else c.next[prev] = nxt;
    c.count[r]--;
}

// isearch.cpp:244-284 (addOpen); (j, z): the cell's coordinates inside row i
template <typename I> ASTAR_HD void astar_add_open(AstarCtx<I>& c, int i, int j, int z, int cell, double f, int g, int dir) {
    bool inserted = false;
    const uint8_t st = c.cell[cell];
    if ((st & kCellStateMask) == kCellOpen) {
        // the reference compares F = g + h with the stored F of the same cell: the same h on both sides and integer g's
        // at least one apart (far above an ulp of the sums), so the comparison of the doubles is the comparison of the g's
        if (g < (int)c.g[cell]) {
            c.g[cell] = (I)g;
            c.cell[cell] = (uint8_t)((st & ~(7 << kCellDirShift)) | (dir << kCellDirShift));
            inserted = true;
        }
    } else {
        c.g[cell] = (I)g;
        c.cell[cell] = (uint8_t)((st & kCellOccupied) | kCellOpen | (dir << kCellDirShift));
        astar_insert(c, i, cell, j, z);
        inserted = true;
        c.open_size++;
    }
    if (c.count[i] == 1) {
        // the row's only node: when it was not touched just now it is the cached minimum already
        if (inserted) { c.min_cell[i] = cell; c.min_f[i] = f; c.min_g[i] = g; }
    } else if (inserted && f <= c.min_f[i]) {
        if (f == c.min_f[i]) { if (g >= c.min_g[i]) { c.min_cell[i] = cell; c.min_g[i] = g; } }
        else { c.min_cell[i] = cell; c.min_f[i] = f; c.min_g[i] = g; }
    }
}

// findMin (isearch.cpp:177-207): rows ascending, a later row replaces the incumbent on equal F unless its g is smaller.
// Serial form; the kernel does the same reduction with one row slice per lane.
template <typename I> ASTAR_HD int astar_find_min(const AstarCtx<I>& c) {
    int cur = -1; double bf = 0; int bg = 0;
    for (int i = 0; i < c.H; i++) {
        if (c.count[i] == 0) continue;
        if (cur < 0 || c.min_f[i] < bf || (c.min_f[i] == bf && c.min_g[i] >= bg)) { cur = c.min_cell[i]; bf = c.min_f[i]; bg = c.min_g[i]; }
    }
    return cur;
}

// One expansion (isearch.cpp:62-105 with deleteMin :209-242) in three pieces, so that the kernel can spread the middle one
// over the warp: (1) close `cur` and erase it from its row container; (2) re-scan the row for its new minimum; (3) return 1
// when `cur` is the goal column (the altitude is ignored, :74), else open its free neighbours.
template <typename I> ASTAR_HD int astar_close(AstarCtx<I>& c, int cur, int& ci, int& cj, int& cz) {
    astar_split(c, cur, ci, cj, cz);
    c.cell[cur] = (uint8_t)((c.cell[cur] & ~kCellStateMask) | kCellClosed);
    c.expansions++;
    const int cur_g = (int)c.g[cur];
    ASTAR_TICK(0);
    astar_erase(c, ci, cur, cj, cz);
    ASTAR_TICK(1);
    return cur_g;
}

// deleteMin's re-scan in the container's iteration order: lowest F, then highest g, then the LAST such node. Every node of
// the row shares i, so h depends on (j, z) only. With `only_g >= 0` the minimum (only_f, only_g) is already known (the
// kernel's warp-wide scan of the row found it, and found it more than once) and just the last node carrying it is looked for.
template <typename I> ASTAR_HD void astar_rescan(AstarCtx<I>& c, int ci, int only_g = -1, double only_f = 0.0) {
    constexpr I kEnd = AstarCtx<I>::kEnd;
    int best = -1; double bf = only_f; int bg = only_g;
    const int di = c.gi - ci;
    for (int p = c.head[ci]; p >= 0;) {
        const int pg = (int)c.g[p];
        const I nx = c.next[p];
        if (only_g < 0 || pg == only_g) {
            int pi, pj, pz;
            astar_split(c, p, pi, pj, pz);
            const int dj = c.gj - pj, dz = c.gz - pz;
            const double pf = (double)pg + astar_h(c, di * di + dj * dj + dz * dz);
            if (only_g >= 0) { if (pf == only_f) best = p; }
            else if (best < 0 || pf < bf || (pf == bf && pg >= bg)) { best = p; bf = pf; bg = pg; }
        }
        p = nx == kEnd ? -1 : (int)nx;
    }
    if (best >= 0) { c.min_cell[ci] = best; c.min_f[ci] = bf; c.min_g[ci] = bg; }
    ASTAR_TICK(2);
}

// the order stamp of an open cell of row ci (smaller = later in the row container's iteration)
template <typename I> ASTAR_HD unsigned astar_stamp_of(const AstarCtx<I>& c, int ci, int cell) {
    return (unsigned)c.bstamp[(size_t)ci * c.bcap + astar_bucket(c, cell, c.level[ci])];
}
// the last node of bucket `of_cell`'s run that carries (f, g): the tie-break among minimal nodes sharing one bucket
template <typename I> ASTAR_HD int astar_last_in_bucket(const AstarCtx<I>& c, int ci, int of_cell, double f, int g) {
    constexpr I kEnd = AstarCtx<I>::kEnd;
    const int cl = c.level[ci];
    const unsigned b = astar_bucket(c, of_cell, cl);
    const I before = c.bkt[(size_t)ci * c.bcap + b];
    int last = -1;
    for (int p = before == kEnd ? c.head[ci] : (int)c.next[before]; p >= 0;) {
        if (astar_bucket(c, p, cl) != b) break;
        const int pg = (int)c.g[p];
        if (pg == g && astar_f(c, p, pg) == f) last = p;
        const I nx = c.next[p];
        p = nx == kEnd ? -1 : (int)nx;
    }
    return last;
}
// deleteMin's re-scan stated over the row's cells and the order stamps instead of the list (serial form of what the kernel
// does with one slice of the row per lane); returns the new row minimum or -1 for an empty row
template <typename I> ASTAR_HD int astar_rescan_by_stamp(const AstarCtx<I>& c, int ci) {
    const int row_cells = c.W * c.A, base = ci * row_cells, di = c.gi - ci;
    int best = -1, bg = 0; double bf = 0; unsigned bs = 0; bool dup = false;
    const int o_lo = c.jlo ? c.jlo[ci] * c.A : 0, o_hi = c.jlo ? (c.jhi[ci] + 1) * c.A : row_cells;
    for (int o = o_lo; o < o_hi; o++) {
        const int id = base + o;
        if ((c.cell[id] & kCellStateMask) != kCellOpen) continue;
        const int pg = (int)c.g[id];
        const int pj = (int)astar_div((unsigned)o, (unsigned)c.A, c.magic_a), pz = o - pj * c.A;
        const int dj = c.gj - pj, dz = c.gz - pz;
        const double pf = (double)pg + astar_h(c, di * di + dj * dj + dz * dz);
        if (best < 0 || pf < bf || (pf == bf && pg > bg)) { best = id; bf = pf; bg = pg; bs = astar_stamp_of(c, ci, id); dup = false; }
        else if (pf == bf && pg == bg) {
            const unsigned st = astar_stamp_of(c, ci, id);
            if (st < bs) { best = id; bs = st; dup = false; }
            else if (st == bs) dup = true;
        }
    }
    if (best >= 0 && dup) best = astar_last_in_bucket(c, ci, best, bf, bg);
    return best;
}

template <typename I> ASTAR_HD int astar_open_neighbours(AstarCtx<I>& c, int ci, int cj, int cz, int cur_g) {
    c.open_size--;
    if (ci == c.gi && cj == c.gj) return 1;
    const int g = cur_g + 1;
    for (int s = 0; s < 6; s++) {
        // move order of the reference's 6-connected neighbourhood: -i, -j, -z, +z, +j, +i
        const int mi = s == 0 ? -1 : (s == 5 ? 1 : 0), mj = s == 1 ? -1 : (s == 4 ? 1 : 0), mz = s == 2 ? -1 : (s == 3 ? 1 : 0);
        const int ni = ci + mi, nj = cj + mj, nz = cz + mz;
        if (ni < 0 || ni >= c.H || nj < 0 || nj >= c.W || nz < 0 || nz >= c.A) continue;
        const int nc = astar_cell(c, ni, nj, nz);
        const uint8_t st = c.cell[nc];
        if ((st & kCellOccupied) || (st & kCellStateMask) == kCellClosed) continue;
        astar_add_open(c, ni, nj, nz, nc, (double)g + astar_heuristic(c, ni, nj, nz), g, s);
    }
    ASTAR_TICK(3);
    return 0;
}

template <typename I> ASTAR_HD int astar_expand(AstarCtx<I>& c, int cur) {
    int ci, cj, cz;
    const int cur_g = astar_close(c, cur, ci, cj, cz);
    astar_rescan(c, ci);
    return astar_open_neighbours(c, ci, cj, cz, cur_g);
}

// the cell the move `dir` came from
template <typename I> ASTAR_HD int astar_parent(const AstarCtx<I>& c, int cell) {
    const int dir = (c.cell[cell] >> kCellDirShift) & 7;
    if (dir == kCellNoParent) return -1;
    const int WA = c.W * c.A;
    const int mi = dir == 0 ? -1 : (dir == 5 ? 1 : 0), mj = dir == 1 ? -1 : (dir == 4 ? 1 : 0), mz = dir == 2 ? -1 : (dir == 3 ? 1 : 0);
    return cell - mi * WA - mj * c.A - mz;
}

template <typename I> ASTAR_HD void astar_begin(AstarCtx<I>& c, int si, int sj, int sz) {
    c.expansions = 0; c.open_size = 0;
    const int s = astar_cell(c, si, sj, sz);
    astar_add_open(c, si, sj, sz, s, astar_heuristic(c, si, sj, sz), 0, kCellNoParent);
    c.open_size = 1;
}

}  // namespace lscgpu
