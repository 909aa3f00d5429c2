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
//   * a first pass finds the count, minimum and maximum key: members of one ensemble share sign and most exponent bits,
//     so the digits start after the common prefix of min and max (registers and shuffles only, no atomics);
//   * targets that share a key prefix form a group with one 2048-bin shared-memory histogram of the next 11 key bits;
//     a pass streams the segment once, finds each element's group with one lookup in a 2048-entry table keyed by
//     the low bits of the prefix (most elements miss every group) and adds to its histogram with shared-memory atomics;
//   * a group whose bin holds <= Q_CAP elements switches to collecting them into shared memory, where the wanted ranks
//     are picked by counting — for smooth data that is the third pass (11 bits narrow 262 144 members to a few hundred);
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
constexpr int Q_UNROLL = 8;

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
    unsigned lut[Q_BINS]; // bit g set at [low 11 bits of group g's prefix]: most elements miss every group in one lookup
    // groups: targets [gbeg, gend) share the key prefix gprefix (top `bits` bits, right-aligned) carried by gcount elements
    unsigned long long gprefix[QT_MAX];
    unsigned gcount[QT_MAX], gfill[QT_MAX];
    int gbeg[QT_MAX], gend[QT_MAX];
    int ngroups, bits, pending;
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
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    // ---- pass 0: count, smallest and largest key of the non-NaN values (registers + shuffles only) ---------------
    // Members of one ensemble share sign and most exponent bits: the digits start after the common prefix of min and max.
    {
        unsigned long long kmin = ~0ull, kmax = 0ull;
        long long cnt = 0;
        for (long long base = tid; base < a.M; base += Q_UNROLL * Q_THREADS) {
            double xs[Q_UNROLL]; // independent loads first: memory-level parallelism for the streaming pass
#pragma unroll
            for (int u = 0; u < Q_UNROLL; ++u) {
                const long long i = base + static_cast<long long>(u) * Q_THREADS;
                xs[u] = i < a.M ? seg[i] : qnan;
            }
#pragma unroll
            for (int u = 0; u < Q_UNROLL; ++u) {
                if (xs[u] == xs[u]) {
                    const unsigned long long k = q_key(xs[u]);
                    kmin = k < kmin ? k : kmin;
                    kmax = k > kmax ? k : kmax;
                    ++cnt;
                }
            }
        }
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long omin = __shfl_down_sync(0xffffffffu, kmin, off), omax = __shfl_down_sync(0xffffffffu, kmax, off);
            kmin = omin < kmin ? omin : kmin;
            kmax = omax > kmax ? omax : kmax;
            cnt += __shfl_down_sync(0xffffffffu, cnt, off);
        }
        // per-warp partials through the (not yet used) candidate buffers
        if (lane == 0) { sh.buf[0][warp] = kmin; sh.buf[1][warp] = kmax; sh.buf[2][warp] = static_cast<unsigned long long>(cnt); }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < NWARPS; ++w) {
                kmin = sh.buf[0][w] < kmin ? sh.buf[0][w] : kmin;
                kmax = sh.buf[1][w] > kmax ? sh.buf[1][w] : kmax;
                cnt += static_cast<long long>(sh.buf[2][w]);
            }
            sh.n = cnt;
            sh.tprefix[0] = kmin;
            sh.tprefix[1] = kmax;
        }
        __syncthreads();
    }
    const long long n = sh.n;
    if (n == 0) {
        if (tid < a.nq) a.result[(static_cast<long long>(tid) * a.rows + row) * a.S + s] = qnan;
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
        const unsigned long long kmin = sh.tprefix[0], kmax = sh.tprefix[1];
        const int common = kmin == kmax ? 64 : __clzll(static_cast<long long>(kmin ^ kmax));
        for (int t = 0; t < nt; ++t) {
            sh.trank[t] = ranks[t];
            sh.tdone[t] = common == 64; // every value is the same
            sh.tkey[t] = kmin;
        }
        // one group owning every target: the common prefix of all keys
        sh.ngroups = 1;
        sh.gbeg[0] = 0;
        sh.gend[0] = nt;
        sh.gprefix[0] = common == 0 ? 0ull : (kmin >> (64 - common));
        sh.gcount[0] = n > 0xffffffffLL ? 0xffffffffu : static_cast<unsigned>(n);
        sh.gfill[0] = 0u;
        sh.bits = common;
        sh.pending = common == 64 ? 0 : nt;
    }
    __syncthreads();

    while (sh.pending != 0) {
        // ---- one pass over the segment: histogram the next digit of large groups, collect the small ones -----------
        const int ng = sh.ngroups, bits = sh.bits;
        const int width = (64 - bits < Q_BITS) ? 64 - bits : Q_BITS;
        for (int g = 0; g < ng; ++g)
            if (sh.gcount[g] > Q_CAP)
                for (int i = tid; i < Q_BINS; i += Q_THREADS) sh.hist[g][i] = 0u;
        // which groups histogram (bit g) instead of collecting; group 0's prefix for the single-group pass
        unsigned histmask = 0u;
#pragma unroll
        for (int g = 0; g < QT_MAX; ++g)
            if (g < ng && sh.gcount[g] > Q_CAP) histmask |= 1u << g;
        const unsigned long long gp0 = sh.gprefix[0];
        const bool no_prefix = bits == 0, single = ng == 1;
        const int dshift = 64 - bits - width;
        const unsigned long long dmask = (1ull << width) - 1ull;
        if (!single) { // lookup table keyed by the low bits of the prefix (the digit resolved last)
            for (int i = tid; i < Q_BINS; i += Q_THREADS) sh.lut[i] = 0u;
            __syncthreads();
            if (tid < ng) atomicOr(&sh.lut[static_cast<unsigned>(sh.gprefix[tid]) & (Q_BINS - 1)], 1u << tid);
        }
        __syncthreads();
        for (long long base = tid; base < a.M; base += Q_UNROLL * Q_THREADS) {
            double xs[Q_UNROLL];
#pragma unroll
            for (int u = 0; u < Q_UNROLL; ++u) {
                const long long i = base + static_cast<long long>(u) * Q_THREADS;
                xs[u] = i < a.M ? seg[i] : qnan;
            }
#pragma unroll
            for (int u = 0; u < Q_UNROLL; ++u) {
                const double x = xs[u];
                if (x != x) continue;
                const unsigned long long key = q_key(x), pre = no_prefix ? 0ull : key >> (64 - bits);
                int g = 0;
                bool hit;
                if (single) { // the first histogram pass: one group, (nearly) every element belongs to it
                    hit = pre == gp0;
                } else {
                    unsigned cand = sh.lut[static_cast<unsigned>(pre) & (Q_BINS - 1)];
                    hit = false;
                    while (cand) { // rarely more than one candidate: groups whose prefixes share their low 11 bits
                        const int j = __ffs(cand) - 1;
                        cand &= cand - 1;
                        if (sh.gprefix[j] == pre) { g = j; hit = true; break; }
                    }
                }
                if (!hit) continue;
                if ((histmask >> g) & 1u) {
                    // plain shared-memory atomics: after the common prefix the digits are spread over the bins
                    atomicAdd(&sh.hist[g][static_cast<unsigned>((key >> dshift) & dmask)], 1u);
                } else {
                    const unsigned p = atomicAdd(&sh.gfill[g], 1u);
                    if (p < Q_CAP) sh.buf[g][p] = key;
                }
            }
        }
        __syncthreads();
        // collected groups: each target's order statistic by counting (one warp per target; ties give the same key)
        for (int g = 0; g < ng; ++g) {
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
        // histogram groups: every target moves into the bin that holds its rank (one warp per target)
        for (int g = 0; g < ng; ++g) {
            if (sh.gcount[g] <= Q_CAP) continue;
            for (int t = sh.gbeg[g] + warp; t < sh.gend[g]; t += NWARPS) {
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
            int g2 = 0, pending = 0;
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
                    ++g2;
                }
            }
            sh.ngroups = g2;
            sh.bits = nbits;
            sh.pending = pending;
        }
        __syncthreads();
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
