// gate_peer.cuh -- the path's one collective, the all-gather of the final waveform (north_star),
// as NVLink stores issued by a kernel instead of a library collective.
//
// Every rank keeps the gathered [world * C][N] result in a buffer that all peers have mapped (CUDA IPC / symmetric
// memory).  The gate kernels write a rank's own rows straight into its slice of the LOCAL buffer at full speed; as soon
// as a channel group is finished, k_peer_push -- a few CTAs on SMs the gate kernels leave free -- reads those rows once
// and stores them into the same slice of every peer's buffer (one 16-byte load, world - 1 16-byte remote stores per
// lane, 512 contiguous bytes per warp and peer), so the NVLink traffic of group g travels behind the kernels of group
// g + 1.  k_peer_barrier (one warp) ends the step: release-store this rank's epoch into every peer's flag array, then
// acquire-spin on the local array until every peer's epoch has arrived -- all of a rank's pushes precede its flag in
// stream order, so a rank that passes the barrier holds every peer's rows.
#pragma once
#include "cuda_compat.h"

namespace b200 {

constexpr int kMaxPeers = 8;

struct PushArgs {
    const uint4* src;              // local rows, 16-byte aligned
    uint4* dst[kMaxPeers];         // the same rows in each peer's buffer
    int n_dst;
    long long rows, vec_per_row;   // 16-byte vectors per row
    long long src_stride, dst_stride;   // row pitches in vectors
};

// One k_peer_push CTA owns a whole SM (1024 threads + a dynamic shared-memory request no other CTA fits beside): the
// hardware block scheduler otherwise spreads small copy CTAs over as many SMs as it can, and every SM holding one cannot
// take a gate CTA (k1d / k2d need an SM's whole register file) -- measured: any co-running small-CTA copy kernel, ours or
// NCCL's, doubled k1d's time.  R fat CTAs block exactly R SMs; the gate kernels' grids are sized for the rest.
constexpr int kPushThreads = 1024;
constexpr int kPushSmemBytes = 200 * 1024;
constexpr int kPushUnroll = 12;              // 16-byte vectors in flight per thread (128 bytes): the copy is latency bound

__global__ void __launch_bounds__(kPushThreads, 1) k_peer_push(const PushArgs a) {
    const long long step = (long long)gridDim.x * blockDim.x;
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long r = 0; r < a.rows; ++r) {
        const uint4* s = a.src + r * a.src_stride;
        const long long drow = r * a.dst_stride;
        for (long long c = first; c < a.vec_per_row; c += kPushUnroll * step) {
            uint4 v[kPushUnroll];
#pragma unroll
            for (int u = 0; u < kPushUnroll; ++u)
                if (c + u * step < a.vec_per_row) v[u] = s[c + u * step];
#pragma unroll
            for (int p = 0; p < kMaxPeers; ++p) {
                if (p < a.n_dst) {
                    uint4* d = a.dst[p] + drow + c;
#pragma unroll
                    for (int u = 0; u < kPushUnroll; ++u)
                        if (c + u * step < a.vec_per_row) d[u * step] = v[u];
                }
            }
        }
    }
}

struct BarrierArgs {
    unsigned* peer_flags[kMaxPeers];    // entry p: peer p's flag array (this rank writes slot [rank] of it)
    unsigned* local_flags;              // this rank's flag array (peer p writes slot [p])
    int rank, world;
    unsigned epoch;
};

__global__ void k_peer_barrier(const BarrierArgs a) {
    const int p = threadIdx.x;
    if (p >= a.world || p == a.rank) return;
#ifdef B200_CUSIM_BUILD
    a.peer_flags[p][a.rank] = a.epoch;
#else
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.peer_flags[p] + a.rank), "r"(a.epoch) : "memory");
    unsigned v;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a.local_flags + p) : "memory");
    } while ((int)(v - a.epoch) < 0);
#endif
}

}  // namespace b200
