// Micro-benchmark: dependent-issue latency of DFMA and MUFU.RCP64H(+Newton) on sm_100a, and DFMA throughput per SMSP as a
// function of independent chains per warp and warps per SMSP.   nvcc -arch=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void dfma_chain(double *out, int iters, double b, double c, long long *cycles)
{
    double a[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) a[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < CH; ++i) a[i] = fma(a[i], b, c);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += a[i];
    if (s == -1.0) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

__global__ void rcp_chain(double *out, int iters, double c, long long *cycles)
{
    double x = 1.5 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double r;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
            const double e = fma(-x, r, 1.0);
            const double t = fma(e, e, e);
            r = fma(r, t, r);
            x = r + c; // next pivot depends on the reciprocal
        }
    }
    long long t1 = clock64();
    if (x == -1.0) out[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

__global__ void mufu_only_chain(double *out, int iters, long long *cycles)
{
    double x = 1.5 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double r;
            asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
            x = r;
        }
    }
    long long t1 = clock64();
    if (x == -1.0) out[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

__global__ void lds_fma_chain(double *out, int iters, long long *cycles)
{
    __shared__ double sm[1024];
    sm[threadIdx.x] = 0.5;
    __syncthreads();
    double x = 0.25;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = (__double2int_rn(x) + threadIdx.x) & 1023; // address depends on x: LDS in the chain
            x = fma(sm[idx], x, 0.125);
        }
    }
    long long t1 = clock64();
    if (x == -1.0) out[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int CH> void run(const char *name, int threads, int blocks, double *d_out, long long *d_cyc)
{
    const int iters = 2000;
    dfma_chain<CH><<<blocks, threads>>>(d_out, iters, 0.999999, 1e-6, d_cyc);
    dfma_chain<CH><<<blocks, threads>>>(d_out, iters, 0.999999, 1e-6, d_cyc);
    long long c = 0;
    cudaMemcpy(&c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
    const double per = double(c) / (iters * 8.0 * CH);
    printf("%s: chains/warp %d, warps/CTA %d, CTAs %d: %.2f cycles per DFMA per warp (%.2f per dependent step)\n", name, CH, threads / 32, blocks, per,
           per * CH);
}

int main()
{
    double *d_out;
    long long *d_cyc;
    cudaMalloc(&d_out, 64);
    cudaMalloc(&d_cyc, 64);
    run<1>("1 warp on one SMSP", 32, 1, d_out, d_cyc);
    run<2>("1 warp on one SMSP", 32, 1, d_out, d_cyc);
    run<4>("1 warp on one SMSP", 32, 1, d_out, d_cyc);
    run<8>("1 warp on one SMSP", 32, 1, d_out, d_cyc);
    run<1>("4 warps = 1 per SMSP", 128, 1, d_out, d_cyc);
    run<1>("8 warps = 2 per SMSP", 256, 1, d_out, d_cyc);
    run<1>("12 warps = 3 per SMSP", 384, 1, d_out, d_cyc);
    run<2>("12 warps = 3 per SMSP", 384, 1, d_out, d_cyc);
    run<4>("12 warps = 3 per SMSP", 384, 1, d_out, d_cyc);
    run<1>("32 warps = 8 per SMSP", 1024, 1, d_out, d_cyc);
    const int iters = 2000;
    long long c = 0;
    rcp_chain<<<1, 32>>>(d_out, iters, 0.5, d_cyc); rcp_chain<<<1, 32>>>(d_out, iters, 0.5, d_cyc);
    cudaMemcpy(&c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
    printf("reciprocal chain (MUFU.RCP64H + 3 FMA + 1 ADD dependent): %.1f cycles per link\n", double(c) / (iters * 8.0));
    mufu_only_chain<<<1, 32>>>(d_out, iters, d_cyc); mufu_only_chain<<<1, 32>>>(d_out, iters, d_cyc);
    cudaMemcpy(&c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
    printf("MUFU.RCP64H dependent chain: %.1f cycles per link\n", double(c) / (iters * 8.0));
    lds_fma_chain<<<1, 32>>>(d_out, iters, d_cyc); lds_fma_chain<<<1, 32>>>(d_out, iters, d_cyc);
    cudaMemcpy(&c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
    printf("LDS.64 + DFMA + F2I + IADD dependent chain: %.1f cycles per link\n", double(c) / (iters * 8.0));
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
