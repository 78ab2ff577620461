// pm_lap.cu — K4: exact rectangular linear sum assignment (min), nr <= nc, on the GPU.
//
// Replaces scipy.optimize.linear_sum_assignment at platymatch/_dock_widget.py:604-611 (third-party
// C++ in the reference: modified Jonker-Volgenant / Crouse 2016 shortest augmenting paths).
// Same optimum (equal total cost; identical assignment where the optimum is unique), different
// schedule, built for the GPU:
//
//  Phase 1 — parallel bidding with epsilon = 0 (all SMs).  Every free row scans its cost row against
//    the column prices v and bids for its best column j1 with increment gamma = w2 - w1 (second best
//    minus best reduced value).  Per column the largest increment wins (ties: lowest row), the price
//    drops by gamma and the previous owner is released.  With epsilon = 0 every assigned edge stays
//    an exact arg-min of its row (prices only fall), so complementary slackness holds EXACTLY and
//    (u, v, matching) is a valid warm start for phase 2 at any time; rows whose bid would be a
//    zero-increment steal simply wait for phase 2 (this is where epsilon = 0 auctions stall).
//  Phase 2 — shortest augmenting paths for the remaining free rows (one CTA per matrix, float64
//    duals, Dijkstra over reduced costs).  Each thread owns 8 columns whose tentative distance d and
//    price v live in registers; one step = one coalesced cost-row load + one block arg-min.
//    Unassigned columns keep v = 0 (rectangular dual feasibility), exactly as in Crouse's algorithm.
//
// Costs are the float32 matrix produced by K3; duals, distances and the total are float64, so the
// result is optimal for the matrix given (up to float64 rounding, like scipy on the same values).
#include <limits.h>
#include <cooperative_groups.h>
#include "pm_common.cuh"

namespace cg = cooperative_groups;

#define PM_LAP_BID_THREADS 1024
#define PM_LAP_CPT 8
#define PM_LAP_MAX_THREADS 1024

struct PmLapView {
    const float *cost;      // [nr][ldc]
    double *u, *v;          // [nr], [nc]
    int32_t *row4col;       // [nc]
    int32_t *col4row;       // [nr]  (caller's output buffer)
    int32_t *bid_col;       // [nr]
    double *bid_gamma;      // [nr]
    unsigned long long *colbest;  // [nc] winning bid key of the round (0 = no bid)
    int32_t *free_lists;    // [3][nr] rotating free-row lists
    int32_t *counters;      // [0..2] list counts, [3] bid rounds run, [4] status
    long long *stats;       // [PM_LAP_STATS] or null
    double *total;
};

static inline size_t pm_lap_align(size_t x) { return (x + 255) & ~(size_t)255; }

struct PmLapLayout {
    size_t u, v, row4col, bid_col, bid_gamma, colbest, free_lists, counters, per_item;
};

static PmLapLayout pm_lap_layout(int nr, int nc) {
    PmLapLayout L;
    size_t o = 0;
    L.u = o; o += pm_lap_align((size_t)nr * 8);
    L.v = o; o += pm_lap_align((size_t)nc * 8);
    L.row4col = o; o += pm_lap_align((size_t)nc * 4);
    L.bid_col = o; o += pm_lap_align((size_t)nr * 4);
    L.bid_gamma = o; o += pm_lap_align((size_t)nr * 8);
    L.colbest = o; o += pm_lap_align((size_t)nc * 8);
    L.free_lists = o; o += pm_lap_align((size_t)nr * 4 * 3);
    L.counters = o; o += pm_lap_align(16 * 4);
    L.per_item = o;
    return L;
}

struct PmLapBatch {   // passed by value to kernels
    const float *cost; size_t cost_stride; int nr, nc, ncp, ldc;
    char *ws; PmLapLayout L;
    int32_t *col4row; long long *stats; double *total;
    int32_t *progress;   // [2] assignments made in the current / previous bidding round
};

__device__ __forceinline__ PmLapView pm_lap_view(const PmLapBatch &B, int b) {
    PmLapView V;
    char *w = B.ws + (size_t)b * B.L.per_item;
    V.cost = B.cost + (size_t)b * B.cost_stride;
    V.u = (double *)(w + B.L.u); V.v = (double *)(w + B.L.v);
    V.row4col = (int32_t *)(w + B.L.row4col);
    V.col4row = B.col4row + (size_t)b * B.nr;
    V.bid_col = (int32_t *)(w + B.L.bid_col); V.bid_gamma = (double *)(w + B.L.bid_gamma);
    V.colbest = (unsigned long long *)(w + B.L.colbest);
    V.free_lists = (int32_t *)(w + B.L.free_lists);
    V.counters = (int32_t *)(w + B.L.counters);
    V.stats = B.stats ? B.stats + (size_t)b * PM_LAP_STATS : nullptr;
    V.total = B.total + b;
    return V;
}

// ------------------------------------------------------------------------------------- init
__global__ void pm_lap_init_kernel(PmLapBatch B) {
    const PmLapView V = pm_lap_view(B, blockIdx.y);
    const int t = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int i = t; i < B.nr; i += stride) { V.u[i] = 0.0; V.col4row[i] = -1; V.free_lists[i] = i; V.bid_col[i] = -1; }
    for (int j = t; j < B.ncp; j += stride) { V.v[j] = 0.0; V.row4col[j] = -1; V.colbest[j] = 0ull; }
    if (t == 0) {
        V.counters[0] = B.nr; V.counters[1] = 0; V.counters[2] = 0; V.counters[3] = 0; V.counters[4] = 0;
        B.progress[0] = 0; B.progress[1] = 0;
        if (V.stats) for (int k = 0; k < PM_LAP_STATS; ++k) V.stats[k] = 0;
    }
}

// ------------------------------------------------------------------------------------- phase 1
struct PmBid { double w1, w2; int j1; };

__device__ __forceinline__ void pm_bid_push(PmBid &a, double w, int j) {
    if (w < a.w1) { a.w2 = a.w1; a.w1 = w; a.j1 = j; }
    else if (w < a.w2) a.w2 = w;
}
__device__ __forceinline__ void pm_bid_merge(PmBid &a, double w1, double w2, int j1) {
    if (w1 < a.w1 || (w1 == a.w1 && j1 < a.j1)) { a.w2 = fmin(a.w1, w2); a.w1 = w1; a.j1 = j1; }
    else a.w2 = fmin(a.w2, w1);
}

// Winner selection key of a bid: float32 image of the increment (monotone in gamma) in the high
// word, ~row in the low word -> atomicMax picks the largest increment, ties to the lowest row.
// ANY bidder may win without hurting exactness: the winner lowers the price by its OWN float64
// gamma, which makes its edge tight against its own second-best column.
__device__ __forceinline__ unsigned long long pm_bid_key(double gamma, int row) {
    return ((unsigned long long)__float_as_uint(__double2float_rd(gamma)) << 32) |
           (unsigned long long)(0xffffffffu - (unsigned)row);
}

// Persistent cooperative kernel: all rounds of the parallel bidding for all matrices of the batch.
// Per round: (A) every free row is scanned by one CTA and bids; grid barrier; (B) winners take
// their columns, everybody else goes to the next free list; grid barrier.  Three rotating list
// counters (current / next / stale-to-zero) avoid any reset race.  Stops when no row is free, when a
// round assigns nothing (only zero-increment steals left), or after max_rounds.
__global__ void __launch_bounds__(PM_LAP_BID_THREADS) pm_lap_bid_persistent(PmLapBatch B, int batch, int max_rounds) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double s_w1[PM_LAP_BID_THREADS / 32], s_w2[PM_LAP_BID_THREADS / 32];
    __shared__ int s_j1[PM_LAP_BID_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    int round = 0;
    for (; round < max_rounds; ++round) {
        const int cur = round % 3, nxt = (round + 1) % 3, old = (round + 2) % 3;
        // ---- phase A: bids.  Work item = (matrix b, position idx in its free list)
        int total_free = 0;
        for (int b = 0; b < batch; ++b) {
            const PmLapView V = pm_lap_view(B, b);
            const int nfree = __ldcg(&V.counters[cur]);
            const int32_t *list = V.free_lists + (size_t)cur * B.nr;
            // CTA c takes items (c - total_free) mod gridDim of this matrix, so matrices interleave over CTAs
            int first = (int)blockIdx.x - (total_free % (int)gridDim.x);
            if (first < 0) first += gridDim.x;
            for (int idx = first; idx < nfree; idx += gridDim.x) {
                const int i = __ldcg(&list[idx]);
                const float *ci = V.cost + (size_t)i * B.ldc;
                PmBid bid = {INFINITY, INFINITY, INT_MAX};
                // two float4 sweeps per trip so that all loads of a thread are in flight together
                for (int j0 = threadIdx.x * 4; j0 < B.nc; j0 += PM_LAP_BID_THREADS * 8) {
                    const int j1 = j0 + PM_LAP_BID_THREADS * 4;
                    const bool second = j1 < B.nc;
                    const float4 c4 = *reinterpret_cast<const float4 *>(ci + j0);   // ldc % 4 == 0, pad readable
                    const double2 va = __ldcg(reinterpret_cast<const double2 *>(V.v + j0));
                    const double2 vb = __ldcg(reinterpret_cast<const double2 *>(V.v + j0 + 2));
                    float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    double2 wa = make_double2(0.0, 0.0), wb = make_double2(0.0, 0.0);
                    if (second) {
                        d4 = *reinterpret_cast<const float4 *>(ci + j1);
                        wa = __ldcg(reinterpret_cast<const double2 *>(V.v + j1));
                        wb = __ldcg(reinterpret_cast<const double2 *>(V.v + j1 + 2));
                    }
                    if (j0 + 0 < B.nc) pm_bid_push(bid, (double)c4.x - va.x, j0 + 0);
                    if (j0 + 1 < B.nc) pm_bid_push(bid, (double)c4.y - va.y, j0 + 1);
                    if (j0 + 2 < B.nc) pm_bid_push(bid, (double)c4.z - vb.x, j0 + 2);
                    if (j0 + 3 < B.nc) pm_bid_push(bid, (double)c4.w - vb.y, j0 + 3);
                    if (second) {
                        if (j1 + 0 < B.nc) pm_bid_push(bid, (double)d4.x - wa.x, j1 + 0);
                        if (j1 + 1 < B.nc) pm_bid_push(bid, (double)d4.y - wa.y, j1 + 1);
                        if (j1 + 2 < B.nc) pm_bid_push(bid, (double)d4.z - wb.x, j1 + 2);
                        if (j1 + 3 < B.nc) pm_bid_push(bid, (double)d4.w - wb.y, j1 + 3);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ow1 = __shfl_xor_sync(0xffffffffu, bid.w1, o), ow2 = __shfl_xor_sync(0xffffffffu, bid.w2, o);
                    const int oj = __shfl_xor_sync(0xffffffffu, bid.j1, o);
                    pm_bid_merge(bid, ow1, ow2, oj);
                }
                __syncthreads();
                if (lane == 0) { s_w1[warp] = bid.w1; s_w2[warp] = bid.w2; s_j1[warp] = bid.j1; }
                __syncthreads();
                if (threadIdx.x == 0) {
                    PmBid tot = {s_w1[0], s_w2[0], s_j1[0]};
                    for (int w = 1; w < PM_LAP_BID_THREADS / 32; ++w) pm_bid_merge(tot, s_w1[w], s_w2[w], s_j1[w]);
                    // nc == 1: no second best; an unowned single column is simply taken with gamma 0
                    double gamma = (tot.w2 == INFINITY) ? 0.0 : tot.w2 - tot.w1;
                    if (!(gamma > 0.0)) gamma = 0.0;
                    int j1 = tot.j1;
                    if (j1 == INT_MAX || tot.w1 == INFINITY) j1 = -1;                    // all-inf row: phase 2 reports it
                    else if (gamma == 0.0 && __ldcg(&V.row4col[j1]) >= 0) j1 = -1;       // zero-increment steal: wait
                    V.bid_col[i] = j1;
                    V.bid_gamma[i] = gamma;
                    if (j1 >= 0) atomicMax(&V.colbest[j1], pm_bid_key(gamma, i));
                }
            }
            total_free += nfree;
            if (gtid == 0) V.counters[old] = 0;        // consumed in the previous round
        }
        if (gtid == 0) B.progress[round & 1] = 0;
        if (total_free == 0) break;                    // uniform: every CTA read the same counters
        grid.sync();
        // ---- phase B: winners take their columns; everybody else (and displaced owners) -> next list
        for (int b = 0; b < batch; ++b) {
            const PmLapView V = pm_lap_view(B, b);
            const int nfree = __ldcg(&V.counters[cur]);
            const int32_t *list = V.free_lists + (size_t)cur * B.nr;
            int32_t *next = V.free_lists + (size_t)nxt * B.nr;
            int32_t *next_count = &V.counters[nxt];
            for (int idx = gtid; idx < nfree; idx += gthreads) {
                const int i = __ldcg(&list[idx]), j = __ldcg(&V.bid_col[i]);
                bool won = false;
                if (j >= 0) {
                    const double gamma = __ldcg(&V.bid_gamma[i]);
                    won = __ldcg(&V.colbest[j]) == pm_bid_key(gamma, i);
                    if (won) {
                        const int prev = __ldcg(&V.row4col[j]);
                        const double vj = __ldcg(&V.v[j]) - gamma;
                        V.v[j] = vj;
                        V.u[i] = (double)V.cost[(size_t)i * B.ldc + j] - vj;
                        V.row4col[j] = i;
                        V.col4row[i] = j;
                        V.colbest[j] = 0ull;
                        if (prev >= 0) { V.col4row[prev] = -1; next[atomicAdd(next_count, 1)] = prev; }
                        atomicAdd(&B.progress[round & 1], 1);
                    }
                }
                if (!won) next[atomicAdd(next_count, 1)] = i;
            }
        }
        grid.sync();
        // every bid-on column has exactly one winner, which cleared its key above: nothing left to reset
        if (__ldcg(&B.progress[round & 1]) == 0) { ++round; break; }     // stalled: only zero-increment steals left
    }
    if (gtid == 0)
        for (int b = 0; b < batch; ++b) pm_lap_view(B, b).counters[3] = round;
}

// ------------------------------------------------------------------------------------- phase 2
// Block arg-min over (value, tie) with lexicographic order; every thread receives the result.
__device__ __forceinline__ void pm_argmin_warp(double &val, int &tie) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, val, o);
        const int ot = __shfl_xor_sync(0xffffffffu, tie, o);
        if (ov < val || (ov == val && ot < tie)) { val = ov; tie = ot; }
    }
}

// V_IN_REGS: prices of the owned columns live in registers (nc <= 8 * 1024); otherwise they are read
// from global memory (L2-resident) every step and d stays in registers (CPT up to 24).
template <int CPT, bool V_IN_REGS>
__global__ void __launch_bounds__(PM_LAP_MAX_THREADS, 1) pm_lap_sap_kernel(PmLapBatch B, int final_parity) {
    const PmLapView V = pm_lap_view(B, blockIdx.x);
    extern __shared__ __align__(16) unsigned char pm_lap_smem[];
    const int nc = B.nc, nr = B.nr, ldc = B.ldc;
    const int nthreads = blockDim.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int group_stride = nthreads * 4;           // columns covered by one float4 sweep of the CTA
    int32_t *pred = reinterpret_cast<int32_t *>(pm_lap_smem);
    int32_t *r4c = pred + (size_t)(CPT / 4) * group_stride;
    __shared__ double s_val[2][32];
    __shared__ int s_tie[2][32];
    __shared__ int s_scan[33];
    __shared__ int s_nfree;

    // column owned by slot q of this thread
    auto col_of = [&](int q) { return (q >> 2) * group_stride + t * 4 + (q & 3); };

    double vreg[V_IN_REGS ? CPT : 1];
    double d[CPT];
    unsigned assigned = 0u;
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
        const int c = col_of(q);
        const int r = (c < nc) ? V.row4col[c] : -1;
        r4c[c] = r;
        if (r >= 0) assigned |= 1u << q;
        if (V_IN_REGS) vreg[q] = (c < nc) ? V.v[c] : 0.0;
    }
    // ordered list of free rows (ascending row index -> deterministic), into list 0
    int32_t *flist = V.free_lists;
    if (t == 0) s_nfree = 0;
    __syncthreads();
    for (int base = 0; base < nr; base += nthreads) {
        const int i = base + t;
        const bool is_free = (i < nr) && (V.col4row[i] < 0);
        const unsigned bal = __ballot_sync(0xffffffffu, is_free);
        if (lane == 0) s_scan[warp] = __popc(bal);
        __syncthreads();
        if (t == 0) {
            int acc = s_nfree;
            for (int w = 0; w < (nthreads >> 5); ++w) { const int c = s_scan[w]; s_scan[w] = acc; acc += c; }
            s_scan[32] = acc;
        }
        __syncthreads();
        if (is_free) flist[s_scan[warp] + __popc(bal & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (t == 0) s_nfree = s_scan[32];
        __syncthreads();
    }
    const int nfree = s_nfree;
    long long steps = 0;
    int status = 0;
    int parity = 0;

    for (int f = 0; f < nfree && status == 0; ++f) {
        const int cur = flist[f];
        double min_val = 0.0;
        int i = cur, sink = -1;
        unsigned scanned = 0u;
#pragma unroll
        for (int q = 0; q < CPT; ++q) d[q] = INFINITY;
        while (true) {
            const float *ci = V.cost + (size_t)i * ldc;
            const double ui = V.u[i];
            double best = INFINITY;
            int best_tie = INT_MAX;
#pragma unroll
            for (int g = 0; g < CPT / 4; ++g) {
                const int c0 = g * group_stride + t * 4;
                float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 < ldc) c4 = *reinterpret_cast<const float4 *>(ci + c0);
                double2 va = make_double2(0.0, 0.0), vb = make_double2(0.0, 0.0);
                if (!V_IN_REGS && c0 < nc) {   // v padded to a multiple of 4 doubles in the workspace
                    va = *reinterpret_cast<const double2 *>(V.v + c0);
                    vb = *reinterpret_cast<const double2 *>(V.v + c0 + 2);
                }
                const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
                const double vv[4] = {va.x, va.y, vb.x, vb.y};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int q = g * 4 + e, c = c0 + e;
                    if (c < nc && !((scanned >> q) & 1u)) {
                        const double vj = V_IN_REGS ? vreg[q] : vv[e];
                        const double r = ((min_val + (double)cc[e]) - ui) - vj;
                        if (r < d[q]) { d[q] = r; pred[c] = i; }
                        const int tie = (((assigned >> q) & 1u) ? (1 << 30) : 0) | c;   // prefer a free column
                        if (d[q] < best || (d[q] == best && tie < best_tie)) { best = d[q]; best_tie = tie; }
                    }
                }
            }
            pm_argmin_warp(best, best_tie);
            if (lane == 0) { s_val[parity][warp] = best; s_tie[parity][warp] = best_tie; }
            __syncthreads();
            best = (lane < (nthreads >> 5)) ? s_val[parity][lane] : INFINITY;
            best_tie = (lane < (nthreads >> 5)) ? s_tie[parity][lane] : INT_MAX;
            pm_argmin_warp(best, best_tie);
            parity ^= 1;
            ++steps;
            if (!(best < INFINITY)) { status = PM_ERR_INFEASIBLE; break; }
            min_val = best;
            const int jmin = best_tie & ((1 << 30) - 1);
            {   // owner marks the column as scanned
                const int g = jmin / group_stride, rem = jmin - g * group_stride;
                if ((rem >> 2) == t) scanned |= 1u << (g * 4 + (rem & 3));
            }
            if (!(best_tie >> 30)) { sink = jmin; break; }
            i = r4c[jmin];
        }
        if (status) break;
        // dual update (Crouse: u[cur] += min; scanned rows/cols shift by min - d)
#pragma unroll
        for (int q = 0; q < CPT; ++q) {
            if ((scanned >> q) & 1u) {
                const int c = col_of(q);
                const double delta = min_val - d[q];
                if (V_IN_REGS) vreg[q] -= delta; else V.v[c] -= delta;
                if (c != sink) V.u[r4c[c]] += delta;
            }
        }
        if (t == 0) V.u[cur] += min_val;
        __syncthreads();
        if (t == 0) {   // augment along the predecessor chain
            int j = sink;
            while (true) {
                const int r = pred[j];
                r4c[j] = r;
                const int jn = V.col4row[r];
                V.col4row[r] = j;
                j = jn;
                if (r == cur) break;
            }
        }
        __syncthreads();
        assigned = 0u;
#pragma unroll
        for (int q = 0; q < CPT; ++q)
            if (r4c[col_of(q)] >= 0) assigned |= 1u << q;
    }
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
        const int c = col_of(q);
        if (c < nc) {
            if (V_IN_REGS) V.v[c] = vreg[q];
            V.row4col[c] = r4c[c];
        }
    }
    // total cost, float64, fixed order
    __syncthreads();
    double tot = 0.0;
    for (int r = t; r < nr; r += nthreads) {
        const int c = V.col4row[r];
        if (c >= 0) tot += (double)V.cost[(size_t)r * ldc + c];
    }
    tot = pm_block_sum(tot, &s_val[0][0]);
    if (t == 0) {
        V.total[0] = status ? nan("") : tot;
        V.counters[4] = status;
        if (V.stats) {
            V.stats[PM_LAP_STAT_BID_ROUNDS] = V.counters[3];
            V.stats[PM_LAP_STAT_ROWS_AFTER_BIDDING] = nr - nfree;
            V.stats[PM_LAP_STAT_AUGMENTATIONS] = nfree;
            V.stats[PM_LAP_STAT_DIJKSTRA_STEPS] = steps;
            V.stats[PM_LAP_STAT_STATUS] = status;
        }
    }
    (void)final_parity;
}

// ------------------------------------------------------------------------------------- host
extern "C" size_t pm_lap_workspace_bytes(int batch, int nr, int nc) {
    if (batch < 1 || nr < 1 || nc < 1) return 0;
    const int ncp = (nc + 3) & ~3;
    return 256 + (size_t)batch * pm_lap_layout(nr, ncp).per_item;
}

template <int CPT, bool VR>
static int pm_lap_launch_sap(const PmLapBatch &B, int batch, int threads, cudaStream_t s) {
    const size_t smem = (size_t)(CPT / 4) * threads * 4 * 2 * sizeof(int32_t);
    PM_CUDA_TRY(cudaFuncSetAttribute(pm_lap_sap_kernel<CPT, VR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pm_lap_sap_kernel<CPT, VR><<<batch, threads, smem, s>>>(B, 0);
    PM_LAUNCH_CHECK();
    return PM_OK;
}

extern "C" int pm_lap_solve(const float *cost, int batch, int nr, int nc, int ldc, int max_bid_rounds,
                            int32_t *col4row, double *total, int64_t *stats, void *workspace,
                            size_t workspace_bytes, void *stream) {
    PM_REQUIRE(cost && col4row && total && workspace, "null pointer");
    PM_REQUIRE(batch >= 1 && nr >= 1 && nc >= 1, "empty problem");
    PM_REQUIRE(nr <= nc, "need nr <= nc (transpose on the host side, as scipy does)");
    PM_REQUIRE(ldc >= nc && ldc % 4 == 0, "ldc must be >= nc and a multiple of 4");
    PM_REQUIRE((reinterpret_cast<size_t>(cost) & 15) == 0, "cost must be 16-byte aligned");
    PM_REQUIRE(max_bid_rounds >= 0, "max_bid_rounds < 0");
    if (nc > 24 * PM_LAP_MAX_THREADS) {
        pm_set_error("pm_lap_solve: nc = %d > %d columns not supported by the single-CTA path", nc,
                     24 * PM_LAP_MAX_THREADS);
        return PM_ERR_UNSUPPORTED;
    }
    if (workspace_bytes < pm_lap_workspace_bytes(batch, nr, nc)) {
        pm_set_error("pm_lap_solve: workspace too small");
        return PM_ERR_WORKSPACE;
    }
    cudaStream_t s = pm_stream(stream);
    PmLapBatch B;
    B.cost = cost; B.cost_stride = (size_t)nr * ldc; B.nr = nr; B.nc = nc; B.ldc = ldc;
    B.ncp = (nc + 3) & ~3;
    B.L = pm_lap_layout(nr, B.ncp);
    B.progress = (int32_t *)workspace;
    B.ws = (char *)workspace + 256;
    B.col4row = col4row; B.stats = (long long *)stats; B.total = total;

    pm_lap_init_kernel<<<dim3(32, batch), 256, 0, s>>>(B);
    PM_LAUNCH_CHECK();
    if (max_bid_rounds > 0) {
        int dev = 0, sms = 0, per_sm = 0;
        PM_CUDA_TRY(cudaGetDevice(&dev));
        PM_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        PM_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pm_lap_bid_persistent, PM_LAP_BID_THREADS, 0));
        if (per_sm < 1) { pm_set_error("pm_lap_solve: bidding kernel does not fit on an SM"); return PM_ERR_CUDA; }
        if (per_sm > 1) per_sm = 1;          // one CTA per SM: the grid barrier scales with the CTA count
        int grid = sms * per_sm;
        void *args[] = {(void *)&B, (void *)&batch, (void *)&max_bid_rounds};
        PM_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)pm_lap_bid_persistent, dim3(grid), dim3(PM_LAP_BID_THREADS),
                                                args, 0, s));
        PM_LAUNCH_CHECK();
    }
    int rc;
    if (nc <= PM_LAP_CPT * PM_LAP_MAX_THREADS) {
        int threads = ((nc + PM_LAP_CPT - 1) / PM_LAP_CPT + 31) & ~31;
        if (threads < 64) threads = 64;
        if (threads > PM_LAP_MAX_THREADS) threads = PM_LAP_MAX_THREADS;
        rc = pm_lap_launch_sap<PM_LAP_CPT, true>(B, batch, threads, s);
    } else if (nc <= 16 * PM_LAP_MAX_THREADS) {
        rc = pm_lap_launch_sap<16, false>(B, batch, PM_LAP_MAX_THREADS, s);
    } else {
        rc = pm_lap_launch_sap<24, false>(B, batch, PM_LAP_MAX_THREADS, s);
    }
    return rc;
}
