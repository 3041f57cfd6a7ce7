// cf_comm.cuh -- the path's only exchange step, the sum over participants of the result vector
// [payoff sums, aggregate, table adjoints], done over peer memory (NVLink 5 / NVSwitch) instead of a collective call.
//
// The reference has no counterpart: its worker threads share one address space and add their risks in a loop
// (mcBase.h:737-746, multi: 976-984).  Here a participant is one GPU -- a device of a single-process context
// (cf_init with several devices, peer access) or the device of one process of a multi-process job (cf_comm_create /
// cf_comm_connect, CUDA IPC).  Every participant owns a receive block
//
//     double   rows[2][world][cap]     two epochs, one row per sender
//     uint32_t flags[world]            flags[p] = last epoch participant p has published to this block
//
// mapped into every other participant.  An exchange of epoch e: every participant PUSHES its vector into its row of
// slot e & 1 of every block (posted remote stores: nobody waits for a round trip), publishes e to its flag in every block
// once all its stores are fenced, waits until every flag of its own block has reached e and adds the rows it has
// received -- local memory -- in rank order, so that all participants end with bit-identical sums.  A participant can
// only be one epoch ahead of a peer that still reads (it needs that peer's flag of the previous epoch to get there),
// hence the two slots; rows have a fixed stride (cap) whatever the length of the vector exchanged.
// A peer that never arrives (about ten seconds) yields NaN results and raises the status word, checked by the host
// on its next call -- not a hang.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace cf {

constexpr int kMaxPeers = 16;

struct DPeers {
    int       world, rank;         // world <= 1: no exchange
    uint32_t  epoch;               // 1, 2, ...: the same on every participant for one exchange
    uint32_t  pad;
    size_t    cap;                 // doubles per row
    long long timeout;             // clocks a participant waits for its peers
    double*   buf[kMaxPeers];      // rows of participant r, as mapped here
    uint32_t* flag[kMaxPeers];     // flags of participant r
    uint32_t* ticket;              // local: blocks of the exchanging kernel that have pushed
    int*      status;              // local (mapped host memory): set to the epoch that timed out
};

// Wait (threads 0 .. world - 1 of the block, one peer each) until every participant has published `epoch` to this
// participant's flags; returns false to the whole block when one never did.
__device__ __forceinline__ bool peers_wait(const DPeers& peers)
{
    bool ok = true;
    if (int(threadIdx.x) < peers.world) {
        volatile uint32_t* here = peers.flag[peers.rank] + threadIdx.x;
        const long long t0 = clock64();
        while (int32_t(*here - peers.epoch) < 0)
            if (clock64() - t0 > peers.timeout) { ok = false; break; }
    }
    ok = __syncthreads_and(ok ? 1 : 0) != 0;
    if (!ok && threadIdx.x == 0 && blockIdx.x == 0 && peers.status) *reinterpret_cast<volatile int*>(peers.status) = int(peers.epoch);
    __threadfence_system();
    return ok;
}

// Publish `epoch` to this participant's flag word in every block (threads 0 .. world - 1); the caller has fenced
// (__threadfence_system) every store the flag covers.
__device__ __forceinline__ void peers_publish(const DPeers& peers)
{
    if (int(threadIdx.x) < peers.world)
        *reinterpret_cast<volatile uint32_t*>(peers.flag[threadIdx.x] + peers.rank) = peers.epoch;   // remote store over NVLink
}

// out[i] = sum over participants of their local[i], i < n (n <= cap).  Launched by every participant with the same
// epoch; grid <= number of SMs (every block spins on the flags: all of them must be resident), any block size >= world.
static __global__ void peer_exchange_kernel(const double* __restrict__ local, int n, double* __restrict__ out, const DPeers peers)
{
    const size_t slot = size_t(peers.epoch & 1u) * size_t(peers.world) * peers.cap;
    const size_t mine = slot + size_t(peers.rank) * peers.cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double v = local[i];
        for (int p = 0; p < peers.world; ++p) peers.buf[p][mine + i] = v;
    }
    __threadfence_system();                                   // this block's rows are on their way before its ticket counts
    __syncthreads();
    __shared__ int isLast;
    if (threadIdx.x == 0) isLast = atomicAdd(peers.ticket, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (isLast) {                                             // the last block to push closes the epoch for this participant
        if (threadIdx.x == 0) *peers.ticket = 0u;
        __threadfence_system();
        peers_publish(peers);
    }
    const bool ok = peers_wait(peers);                        // every block waits for all participants (its own included)
    const double* rows = peers.buf[peers.rank] + slot;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < peers.world; ++r) s += __ldcg(rows + size_t(r) * peers.cap + i);   // rank order: identical everywhere
        out[i] = ok ? s : __longlong_as_double(0x7ff8000000000000ll);
    }
}

}  // namespace cf
