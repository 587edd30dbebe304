// profiles/micro/fp64_latency.cu — dependent-chain latencies on sm_100a (DFMA, DADD, LDS, bar.sync, st.cg+fence).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double a, double b) {
    __shared__ double s[64];
    double x = a;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; i++) x = fma(x, b, a);
    long long t1 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; i++) x = x + b;
    long long t2 = clock64();
    s[threadIdx.x] = x; __syncthreads();
#pragma unroll 1
    for (int i = 0; i < 1024; i++) { x = s[(threadIdx.x + (int)x) & 63]; }
    long long t3 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; i++) { __syncthreads(); }
    long long t4 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) { __stcg(out + 8 + threadIdx.x, x); __threadfence(); }
    long long t5 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; i++) { x = x / b; }
    long long t6 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) { x += *(volatile double *)(out + 100 + (i & 7)); }
    long long t7 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) { x += __ldcg(out + 200 + ((int)x & 7)); }
    long long t8 = clock64();
    if (threadIdx.x == 0) {
        cyc[0] = (t1 - t0) / 1024; cyc[1] = (t2 - t1) / 1024; cyc[2] = (t3 - t2) / 1024; cyc[3] = (t4 - t3) / 1024;
        cyc[4] = (t5 - t4) / 256; cyc[5] = (t6 - t5) / 1024; cyc[6] = (t7 - t6) / 256; cyc[7] = (t8 - t7) / 256;
    }
    out[threadIdx.x] = x;
}
int main() {
    double *out; long long *cyc, h[8];
    cudaMalloc(&out, 4096); cudaMemset(out, 0, 4096); cudaMalloc(&cyc, 64);
    for (int threads : {32, 64}) {
        k<<<1, threads>>>(out, cyc, 1.0000001, 0.9999999);
        cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("threads %d: DFMA %lld  DADD %lld  LDS(dependent) %lld  bar.sync %lld  st.cg+threadfence %lld  DDIV %lld  volatile-ld %lld  ld.cg(dependent) %lld cycles\n",
               threads, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    }
    return 0;
}
