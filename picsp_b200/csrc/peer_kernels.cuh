// picsp_b200/csrc/peer_kernels.cuh — the one exchange step of the sharded path as OUR OWN kernels over NVLink peer memory.
//
// Every rank deposits a full partial charge density; the sum over ranks is what the field solve needs on every rank
// (SURVEY 8e).  Round 1 did that with ncclAllReduce (87 us for 8.4 MB at 8 ranks: latency-bound).  Here the ranks of
// one NVSwitch node map each other's buffers (CUDA IPC) and the reduction is a reduce-scatter + all-gather in ONE
// kernel: rank r owns the nodes [nn*r/R, nn*(r+1)/R), reads that slice of every rank's partial rho with plain peer
// loads, adds the R values in rank order (a fixed order: every node is summed by exactly one rank, so all ranks
// receive bit-identical totals) and stores the total into every rank's rho with peer stores.  The two cross-GPU
// barriers around it are flag exchanges through the same mapped memory (release / acquire at system scope), bounded by
// a time-out that raises the context's error flag instead of hanging the GPU.
#pragma once
#include "ctx.cuh"

namespace picsp {

constexpr int PEER_MAX_RANKS = 8;
constexpr int ERR_BIT_PEER = 8;

struct PeerPtrs {
    double *part[PEER_MAX_RANKS];                 // partial rho of rank k (nn doubles)
    double *full[PEER_MAX_RANKS];                 // summed rho of rank k (nn doubles)
    unsigned long long *flags[PEER_MAX_RANKS];    // [2][PEER_MAX_RANKS] arrival epochs on rank k: flags[k][which*MAX + writer]
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// cross-GPU barrier: thread k tells rank k "rank `rank` has reached epoch", then waits for rank k's word to arrive here
__global__ void k_peer_barrier(PeerPtrs p, int rank, int nranks, unsigned long long epoch, int which, int *err) {
    const int k = threadIdx.x;
    if (k >= nranks) return;
    __threadfence_system();                                     // everything this GPU wrote before is visible to the peers
    st_release_sys(p.flags[k] + which * PEER_MAX_RANKS + rank, epoch);
    const unsigned long long *mine = p.flags[rank] + which * PEER_MAX_RANKS + k;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) < epoch) {
        if (clock64() - t0 > 4000000000ll) { atomicOr(err, ERR_BIT_PEER); break; }     // ~2 s: a peer is gone; report, do not hang
    }
}

// peer data is read and written at SYSTEM scope: coherent with what the other GPUs wrote before the barrier
__device__ __forceinline__ double ld_sys(const double *p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(double *p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

__global__ void __launch_bounds__(256)
k_peer_reduce(PeerPtrs p, int rank, int nranks, long long nn) {
    const long long lo = nn * rank / nranks, hi = nn * (rank + 1) / nranks;
    for (long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
        double v[PEER_MAX_RANKS];
#pragma unroll
        for (int k = 0; k < PEER_MAX_RANKS; k++) v[k] = k < nranks ? ld_sys(p.part[k] + i) : 0.0;     // all loads in flight
        double s = v[0];
#pragma unroll
        for (int k = 1; k < PEER_MAX_RANKS; k++) if (k < nranks) s += v[k];                           // rank order: one fixed sum
#pragma unroll
        for (int k = 0; k < PEER_MAX_RANKS; k++) if (k < nranks) st_sys(p.full[k] + i, s);
    }
}

}  // namespace picsp
