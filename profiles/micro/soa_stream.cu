// profiles/micro/soa_stream.cu — what HBM bandwidth does the mover's ACCESS PATTERN allow, with no compute at all?
// (a) plain copy b = a (the MEASURED_PEAKS.json pattern); (b) in-place update of 4 SoA arrays, one CTA per
// 4096-particle chunk (the mover's pattern: 32 B read + 32 B written per particle); (c) same, grid-stride.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_copy(const double *__restrict__ a, double *__restrict__ b, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void __launch_bounds__(256, 4) k_inplace_chunk(double *x, double *y, double *vx, double *vy, long long n, int chunk) {
    const long long start = (long long)blockIdx.x * chunk;
    const int count = (int)min((long long)chunk, n - start);
    for (int k = threadIdx.x; k < count; k += blockDim.x) {
        const long long p = start + k;
        double a = x[p], b = y[p], c = vx[p], d = vy[p];
        c += 1e-3 * a; d += 1e-3 * b; a += 1e-3 * c; b += 1e-3 * d;
        x[p] = a; y[p] = b; vx[p] = c; vy[p] = d;
    }
}
__global__ void __launch_bounds__(256, 4) k_inplace_stride(double *x, double *y, double *vx, double *vy, long long n) {
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        double a = x[p], b = y[p], c = vx[p], d = vy[p];
        c += 1e-3 * a; d += 1e-3 * b; a += 1e-3 * c; b += 1e-3 * d;
        x[p] = a; y[p] = b; vx[p] = c; vy[p] = d;
    }
}
__global__ void __launch_bounds__(256, 4) k_inplace_v2(double2 *x, double2 *y, double2 *vx, double2 *vy, long long n2) {
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n2; p += (long long)gridDim.x * blockDim.x) {
        double2 a = x[p], b = y[p], c = vx[p], d = vy[p];
        c.x += 1e-3 * a.x; c.y += 1e-3 * a.y; d.x += 1e-3 * b.x; d.y += 1e-3 * b.y;
        a.x += 1e-3 * c.x; a.y += 1e-3 * c.y; b.x += 1e-3 * d.x; b.y += 1e-3 * d.y;
        x[p] = a; y[p] = b; vx[p] = c; vy[p] = d;
    }
}
template <class F> float timeit(F f, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    return best;
}
int main() {
    const long long n = 500000000ll;   // particles
    double *x, *y, *vx, *vy;
    cudaMalloc(&x, n * 8); cudaMalloc(&y, n * 8); cudaMalloc(&vx, n * 8); cudaMalloc(&vy, n * 8);
    cudaMemset(x, 0, n * 8); cudaMemset(y, 0, n * 8); cudaMemset(vx, 0, n * 8); cudaMemset(vy, 0, n * 8);
    float t;
    t = timeit([&] { k_copy<<<148 * 16, 256>>>(x, y, n); }, 5);
    printf("copy b=a (16 B/elem)                         : %8.3f ms  %7.1f GB/s\n", t, 16.0 * n / t / 1e6);
    for (int chunk : {2048, 4096, 16384}) {
        t = timeit([&] { k_inplace_chunk<<<(unsigned)((n + chunk - 1) / chunk), 256>>>(x, y, vx, vy, n, chunk); }, 5);
        printf("in-place 4xSoA, CTA per %5d-chunk (64 B/p)   : %8.3f ms  %7.1f GB/s\n", chunk, t, 64.0 * n / t / 1e6);
    }
    t = timeit([&] { k_inplace_stride<<<148 * 16, 256>>>(x, y, vx, vy, n); }, 5);
    printf("in-place 4xSoA, grid-stride (64 B/p)           : %8.3f ms  %7.1f GB/s\n", t, 64.0 * n / t / 1e6);
    t = timeit([&] { k_inplace_v2<<<148 * 16, 256>>>((double2 *)x, (double2 *)y, (double2 *)vx, (double2 *)vy, n / 2); }, 5);
    printf("in-place 4xSoA, grid-stride, 128-bit (64 B/p)  : %8.3f ms  %7.1f GB/s\n", t, 64.0 * n / t / 1e6);
    return 0;
}
