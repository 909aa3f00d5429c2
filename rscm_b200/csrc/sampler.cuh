// sampler.cuh — device side of the affine-invariant ensemble sampler's stretch move.
//
// Reference: crates/rscm-calibrate/src/sampler/moves.rs:55-59 (sample_z: z = ((a-1)u + 1)^2 / a, i.e. g(z) ~ 1/sqrt(z) on
// [1/a, a]), :110-125 (propose: y = x_j + z (x_k - x_j) with j uniform over the complementary half), :76-92
// (acceptance_probability: min(1, z^(n-1) p(y)/p(x)), 0 when the proposal's log-probability is not finite) and
// sampler/ensemble.rs:489-546 (update_group: propose for the whole active half, one log_posterior_batch, accept/reject).
//
// The reference draws from a non-reproducible thread_rng (ensemble.rs:436), so only statistical parity is required of the
// random stream.  Here every draw is a pure function of (seed, walker, step, purpose) through Philox4x32-10, which makes
// the move reproducible, order-independent and — with the walker state replicated on every rank — lets all ranks take
// identical accept/reject decisions without exchanging anything but the log-posteriors.
#pragma once

namespace rscm_dev {

struct Philox4 {
    unsigned x, y, z, w;
};

__host__ __device__ __forceinline__ unsigned mulhi32(unsigned a, unsigned b)
{
    return static_cast<unsigned>((static_cast<unsigned long long>(a) * b) >> 32);
}

// Philox4x32-10 (Salmon et al. 2011): counter (c0..c3), key (k0, k1)
__host__ __device__ inline Philox4 philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1)
{
    constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = mulhi32(M0, c0), lo0 = M0 * c0, hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// 53-bit uniform in [0, 1) from two 32-bit words
__host__ __device__ __forceinline__ double u01(unsigned hi, unsigned lo)
{
    const unsigned long long v = (static_cast<unsigned long long>(hi) << 32) | lo;
    return static_cast<double>(v >> 11) * (1.0 / 9007199254740992.0);
}

enum { STRETCH_PROPOSE = 0, STRETCH_ACCEPT = 1 };

// positions are SoA: element (column c, walker w) at pos[c * ld + w]
__global__ void stretch_propose_kernel(const double *__restrict__ pos, long long ld, int n_cols, long long active0, long long n_active,
                                       long long comp0, long long n_comp, double a, unsigned long long seed, unsigned step,
                                       double *__restrict__ prop, long long ld_prop, double *__restrict__ zs,
                                       const unsigned *__restrict__ iteration)
{
    // `iteration` (device counter, may be null) makes the launch replayable from a CUDA graph: step += 2 * *iteration
    if (iteration) step += 2u * *iteration;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_active) return;
    const long long w = active0 + i;
    const Philox4 r = philox4x32_10(static_cast<unsigned>(w), static_cast<unsigned>(w >> 32), step, STRETCH_PROPOSE,
                                    static_cast<unsigned>(seed), static_cast<unsigned>(seed >> 32));
    const double t = (a - 1.0) * u01(r.x, r.y) + 1.0;
    const double z = t * t / a;
    long long j = static_cast<long long>(u01(r.z, r.w) * static_cast<double>(n_comp));
    if (j >= n_comp) j = n_comp - 1;
    const long long wj = comp0 + j;
    for (int c = 0; c < n_cols; ++c) {
        const double xj = pos[c * ld + wj], xk = pos[c * ld + w];
        prop[c * ld_prop + i] = xj + z * (xk - xj);
    }
    zs[i] = z;
}

__global__ void stretch_accept_kernel(double *__restrict__ pos, long long ld, int n_cols, long long active0, long long n_active,
                                      const double *__restrict__ prop, long long ld_prop, const double *__restrict__ zs,
                                      const double *__restrict__ lp_new, double *__restrict__ logp, unsigned long long seed, unsigned step,
                                      unsigned long long *__restrict__ n_accepted, const unsigned *__restrict__ iteration)
{
    if (iteration) step += 2u * *iteration;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    bool acc = false;
    if (i < n_active) {
        const long long w = active0 + i;
        const double lnew = lp_new[i], lold = logp[w];
        double p = 0.0;
        if (fabs(lnew) <= 1.7976931348623157e308) { // is_finite (moves.rs:82-84)
            const double log_ratio = static_cast<double>(n_cols - 1) * log(zs[i]) + (lnew - lold);
            p = fmin(exp(log_ratio), 1.0); // exp(+inf) = inf -> 1; NaN (inf - inf) -> fmin ignores it -> 1, as f64::min does
        }
        const Philox4 r = philox4x32_10(static_cast<unsigned>(w), static_cast<unsigned>(w >> 32), step, STRETCH_ACCEPT,
                                        static_cast<unsigned>(seed), static_cast<unsigned>(seed >> 32));
        acc = u01(r.x, r.y) < p;
        if (acc) {
            for (int c = 0; c < n_cols; ++c) pos[c * ld + w] = prop[c * ld_prop + i];
            logp[w] = lnew;
        }
    }
    if (n_accepted) {
        const unsigned ballot = __ballot_sync(0xffffffffu, acc);
        if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(n_accepted, static_cast<unsigned long long>(__popc(ballot)));
    }
}

// end of a sampler iteration inside a captured graph: advance the device-side iteration counter
__global__ void advance_iteration_kernel(unsigned *iteration) { *iteration += 1u; }

} // namespace rscm_dev
