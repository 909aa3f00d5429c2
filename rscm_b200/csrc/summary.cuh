// summary.cuh — across-member quantiles of the ensemble output, computed where the output lives.
//
// The ensemble kernel leaves out[row][run] (row = variable x time x region, run = s*M + m) in HBM: 41 GB at the headline
// configuration.  What notebooks built on the reference consume are percentile bands across members per scenario
// (docs/notebooks/scenario_pipeline.py:339-400 takes pandas quantiles over the member axis of looped Model::run results);
// computing them on the device replaces a 41 GB device->host copy by rows x S x n_q doubles.
//
// One CTA per (row, scenario) segment of M contiguous values.  Exact order statistics by most-significant-digit radix
// selection on order-preserving 64-bit keys, all requested ranks at once:
//   * every quantile needs two order statistics (numpy "linear" interpolation); their ranks form <= QT_MAX sorted targets;
//   * targets that share a key prefix form a group with one 2048-bin shared-memory histogram of the next 11 key bits;
//     a pass streams the segment once, finds each element's group by binary search over the (sorted) group prefixes and
//     adds to its histogram with warp-aggregated atomics (the sign/exponent digit is nearly constant across members);
//   * a group whose bin holds <= Q_CAP elements switches to collecting them into shared memory, where the wanted ranks
//     are picked by counting — for smooth data that is the third pass (11 + 11 bits narrow 262 144 members to ~100);
//   * NaNs are excluded (numpy.nanquantile semantics); an all-NaN segment gives NaN.
// Interpolation follows numpy's _lerp (lib/_function_base_impl.py) without FMA contraction, so results are bit-identical
// to numpy.nanquantile(..., method="linear").
#pragma once

namespace rscm_dev {

constexpr int Q_MAXQ = 5;              // quantiles per call
constexpr int QT_MAX = 2 * Q_MAXQ;     // target order statistics per segment
constexpr int Q_BITS = 11, Q_BINS = 1 << Q_BITS;
constexpr int Q_CAP = 512;             // collected candidates per group
constexpr int Q_THREADS = 1024;
constexpr int Q_UNROLL = 4;

struct QArgs {
    const double *data; // [rows][S*M]
    long long M, runs;
    int S, nq;
    double q[Q_MAXQ];
    double *result;     // [nq][rows][S]
    long long rows;
};

__device__ __forceinline__ unsigned long long q_key(double x)
{
    const unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(x));
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double q_unkey(unsigned long long k)
{
    const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double(static_cast<long long>(u));
}

struct QShared {
    unsigned hist[QT_MAX][Q_BINS];
    unsigned long long buf[QT_MAX][Q_CAP];
    // groups: targets [gbeg, gend) share the key prefix gprefix (top `bits` bits, right-aligned) carried by gcount elements
    unsigned long long gprefix[QT_MAX];
    unsigned gcount[QT_MAX], gfill[QT_MAX];
    int gbeg[QT_MAX], gend[QT_MAX];
    int ngroups, bits, any_hist, pending;
    // targets (order statistics), sorted by rank
    long long trank[QT_MAX];         // rank within the group's prefix
    unsigned long long tkey[QT_MAX]; // result key
    unsigned long long tprefix[QT_MAX];
    unsigned tcount[QT_MAX];
    int tdone[QT_MAX];
    int ntargets;
    int qlo[Q_MAXQ], qhi[Q_MAXQ];
    double qt[Q_MAXQ];
    long long n;
};

__global__ void __launch_bounds__(Q_THREADS, 1) member_quantiles_kernel(const QArgs a)
{
    extern __shared__ __align__(16) unsigned char q_smem[];
    QShared &sh = *reinterpret_cast<QShared *>(q_smem);
    const long long row = blockIdx.x / a.S;
    const int s = static_cast<int>(blockIdx.x % a.S);
    const double *seg = a.data + row * a.runs + static_cast<long long>(s) * a.M;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARPS = Q_THREADS / 32;

    // ---- pass 0: histogram of the top digit over all non-NaN values --------------------------------------------
    for (int i = tid; i < Q_BINS; i += Q_THREADS) sh.hist[0][i] = 0u;
    __syncthreads();
    for (long long base = tid; base < a.M; base += Q_UNROLL * Q_THREADS) {
        double xs[Q_UNROLL]; // independent loads first: memory-level parallelism for the streaming pass
#pragma unroll
        for (int u = 0; u < Q_UNROLL; ++u) {
            const long long i = base + static_cast<long long>(u) * Q_THREADS;
            xs[u] = i < a.M ? seg[i] : __longlong_as_double(0x7ff8000000000000LL);
        }
#pragma unroll
        for (int u = 0; u < Q_UNROLL; ++u) {
            const double x = xs[u];
            if (x == x) {
                const unsigned d = static_cast<unsigned>(q_key(x) >> (64 - Q_BITS));
                const unsigned peers = __match_any_sync(__activemask(), d);
                if (lane == __ffs(peers) - 1) atomicAdd(&sh.hist[0][d], static_cast<unsigned>(__popc(peers)));
            }
        }
    }
    __syncthreads();
    if (warp == 0) { // n = number of non-NaN values
        unsigned long long c = 0;
        for (int i = lane; i < Q_BINS; i += 32) c += sh.hist[0][i];
        for (int off = 16; off > 0; off >>= 1) c += __shfl_down_sync(0xffffffffu, c, off);
        if (lane == 0) sh.n = static_cast<long long>(c);
    }
    __syncthreads();
    const long long n = sh.n;
    if (n == 0) {
        if (tid < a.nq) a.result[(static_cast<long long>(tid) * a.rows + row) * a.S + s] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    if (tid == 0) {
        // numpy "linear": virtual index q*(n-1), neighbours floor and floor+1 (clamped), weight = fractional part.
        // The two ranks of every quantile go into one sorted, duplicate-free target list.
        long long lo_rank[Q_MAXQ], hi_rank[Q_MAXQ], ranks[QT_MAX];
        int nt = 0;
        for (int k = 0; k < a.nq; ++k) {
            const double v = a.q[k] * static_cast<double>(n - 1);
            long long j = static_cast<long long>(floor(v));
            j = j < 0 ? 0 : (j > n - 1 ? n - 1 : j);
            lo_rank[k] = j;
            hi_rank[k] = (j + 1 > n - 1) ? n - 1 : j + 1;
            sh.qt[k] = v - static_cast<double>(j);
            for (int w = 0; w < 2; ++w) {
                const long long want = w ? hi_rank[k] : lo_rank[k];
                int p = 0;
                while (p < nt && ranks[p] < want) ++p;
                if (p < nt && ranks[p] == want) continue;
                for (int m = nt; m > p; --m) ranks[m] = ranks[m - 1];
                ranks[p] = want;
                ++nt;
            }
        }
        for (int k = 0; k < a.nq; ++k)
            for (int t = 0; t < nt; ++t) {
                if (ranks[t] == lo_rank[k]) sh.qlo[k] = t;
                if (ranks[t] == hi_rank[k]) sh.qhi[k] = t;
            }
        sh.ntargets = nt;
        for (int t = 0; t < nt; ++t) { sh.trank[t] = ranks[t]; sh.tdone[t] = 0; }
        // one group (empty prefix) owning every target; its histogram of the first digit is hist[0]
        sh.ngroups = 1;
        sh.gbeg[0] = 0;
        sh.gend[0] = nt;
        sh.gprefix[0] = 0ull;
        sh.gcount[0] = 0xffffffffu; // histogram mode
        sh.bits = 0;
    }
    __syncthreads();

    for (;;) {
        // ---- resolve: every target of a histogram group moves into the bin that holds its rank --------------------
        const int ng = sh.ngroups, bits = sh.bits;
        const int width = (64 - bits < Q_BITS) ? 64 - bits : Q_BITS;
        for (int g = 0; g < ng; ++g) {
            if (sh.gcount[g] <= Q_CAP) continue; // collected group: its targets are done
            for (int t = sh.gbeg[g] + warp; t < sh.gend[g]; t += NWARPS) { // one warp per target
                const unsigned long long r = static_cast<unsigned long long>(sh.trank[t]);
                const int nb = 1 << width, per = (nb + 31) / 32;
                unsigned long long mine = 0;
                for (int b = lane * per; b < (lane + 1) * per && b < nb; ++b) mine += sh.hist[g][b];
                unsigned long long incl = mine;
                for (int off = 1; off < 32; off <<= 1) {
                    const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += o;
                }
                const unsigned long long excl = incl - mine;
                if (r >= excl && r < incl) {
                    unsigned long long before = excl;
                    int b = lane * per;
                    for (;; ++b) {
                        const unsigned c = sh.hist[g][b];
                        if (r < before + c) break;
                        before += c;
                    }
                    sh.trank[t] = static_cast<long long>(r - before);
                    sh.tprefix[t] = (sh.gprefix[g] << width) | static_cast<unsigned long long>(b);
                    sh.tcount[t] = sh.hist[g][b];
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            // regroup: maximal runs of unfinished targets with equal prefixes (targets are sorted by rank, hence by key)
            const int nbits = bits + width;
            int g2 = 0, pending = 0, any_hist = 0;
            for (int t = 0; t < sh.ntargets; ++t) {
                if (sh.tdone[t]) continue;
                if (nbits == 64) { // the prefix is the whole key
                    sh.tkey[t] = sh.tprefix[t];
                    sh.tdone[t] = 1;
                    continue;
                }
                ++pending;
                if (g2 > 0 && sh.gprefix[g2 - 1] == sh.tprefix[t] && sh.gend[g2 - 1] == t) {
                    sh.gend[g2 - 1] = t + 1;
                } else {
                    sh.gprefix[g2] = sh.tprefix[t];
                    sh.gcount[g2] = sh.tcount[t];
                    sh.gfill[g2] = 0u;
                    sh.gbeg[g2] = t;
                    sh.gend[g2] = t + 1;
                    any_hist |= sh.tcount[t] > Q_CAP;
                    ++g2;
                }
            }
            sh.ngroups = g2;
            sh.bits = nbits;
            sh.pending = pending;
            sh.any_hist = any_hist;
        }
        __syncthreads();
        if (sh.pending == 0) break;

        // ---- one pass over the segment: histogram the next digit of large groups, collect the small ones -----------
        const int ng2 = sh.ngroups, b2 = sh.bits;
        const int w2 = (64 - b2 < Q_BITS) ? 64 - b2 : Q_BITS;
        for (int g = 0; g < ng2; ++g)
            if (sh.gcount[g] > Q_CAP)
                for (int i = tid; i < Q_BINS; i += Q_THREADS) sh.hist[g][i] = 0u;
        __syncthreads();
        for (long long base = tid; base < a.M; base += Q_UNROLL * Q_THREADS) {
          double xs[Q_UNROLL];
#pragma unroll
          for (int u = 0; u < Q_UNROLL; ++u) {
              const long long i = base + static_cast<long long>(u) * Q_THREADS;
              xs[u] = i < a.M ? seg[i] : __longlong_as_double(0x7ff8000000000000LL);
          }
#pragma unroll
          for (int u = 0; u < Q_UNROLL; ++u) {
            const double x = xs[u];
            if (x != x) continue;
            const unsigned long long key = q_key(x), pre = key >> (64 - b2);
            int lo = 0, hi = ng2; // first group with prefix >= pre (group prefixes ascend)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (sh.gprefix[mid] < pre) lo = mid + 1;
                else hi = mid;
            }
            if (lo < ng2 && sh.gprefix[lo] == pre) {
                if (sh.gcount[lo] > Q_CAP) {
                    const unsigned d = static_cast<unsigned>((key >> (64 - b2 - w2)) & ((1ull << w2) - 1ull));
                    const unsigned slot = (static_cast<unsigned>(lo) << Q_BITS) | d;
                    const unsigned peers = __match_any_sync(__activemask(), slot);
                    if (lane == __ffs(peers) - 1) atomicAdd(&sh.hist[lo][d], static_cast<unsigned>(__popc(peers)));
                } else {
                    const unsigned p = atomicAdd(&sh.gfill[lo], 1u);
                    if (p < Q_CAP) sh.buf[lo][p] = key;
                }
            }
          }
        }
        __syncthreads();
        // collected groups: each target's order statistic by counting (one warp per target; ties give the same key)
        for (int g = 0; g < ng2; ++g) {
            if (sh.gcount[g] > Q_CAP) continue;
            const int c = static_cast<int>(sh.gcount[g]);
            for (int t = sh.gbeg[g] + warp; t < sh.gend[g]; t += NWARPS) {
                const long long r = sh.trank[t];
                for (int e = lane; e < c; e += 32) {
                    const unsigned long long ke = sh.buf[g][e];
                    int less = 0, leq = 0;
                    for (int j = 0; j < c; ++j) {
                        const unsigned long long kj = sh.buf[g][j];
                        less += kj < ke;
                        leq += kj <= ke;
                    }
                    if (less <= r && r < leq) sh.tkey[t] = ke;
                }
                if (lane == 0) sh.tdone[t] = 1;
            }
        }
        __syncthreads();
        if (!sh.any_hist) break; // every remaining group was small enough to collect
    }

    if (tid < a.nq) {
        const double lo = q_unkey(sh.tkey[sh.qlo[tid]]), hi = q_unkey(sh.tkey[sh.qhi[tid]]), t = sh.qt[tid];
        // numpy 2.x _lerp: a + (b - a) * t, or b - (b - a) * (1 - t) when t >= 0.5; no FMA contraction (infinite neighbours
        // give NaN there too)
        const double diff = __dsub_rn(hi, lo);
        double v = __dadd_rn(lo, __dmul_rn(diff, t));
        if (t >= 0.5) v = __dsub_rn(hi, __dmul_rn(diff, __dsub_rn(1.0, t)));
        a.result[(static_cast<long long>(tid) * a.rows + row) * a.S + s] = v;
    }
}

} // namespace rscm_dev
