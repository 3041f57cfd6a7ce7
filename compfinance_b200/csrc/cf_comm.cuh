// cf_comm.cuh -- the path's only exchange step, the sum over participants of the result vector
// [payoff sums, aggregate, table adjoints], done over peer memory (NVLink 5 / NVSwitch) instead of a collective call.
//
// The reference has no counterpart: its worker threads share one address space and add their risks in a loop
// (mcBase.h:737-746, multi: 976-984).  Here a participant is one GPU -- a device of a single-process context
// (cf_init with several devices, peer access) or the device of one process of a multi-process job (cf_comm_create /
// cf_comm_connect, CUDA IPC).  Every participant owns a receive block
//
//     Slot rows[2][world][cap]      two epochs, one row per sender, one 16-byte slot per double
//
// mapped into every other participant.  The protocol is latency-bound (8.7 KB per participant for the north star), so
// it is built like a low-latency collective: a double travels as TWO 8-byte words {low half, epoch} {high half, epoch}
// -- each word is delivered atomically, in any order -- and the receiver simply polls its OWN memory until both words
// of a slot carry the epoch it is waiting for.  No fences, no flags, no tickets, no round trip: an exchange of epoch e is
//   push   the warp that owns output k writes its sum into slot [e & 1][rank][k] of every block (posted remote stores),
//   poll   the same warp reads slots [e & 1][0 .. world)[k] of its own block as they arrive,
//   add    in rank order, so that all participants end with bit-identical sums.
// A participant can only be one epoch ahead of a peer that still reads (to finish epoch e + 1 it needs that peer's words
// of e + 1, which the peer sends after it has finished e), hence the two slots per (sender, output); rows have a fixed
// stride (cap) whatever the length of the vector exchanged.  Epochs start at 1 on zeroed memory.
// A peer that never arrives (about ten seconds) yields NaN results and raises the status word, checked by the host
// on its next call -- not a hang.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace cf {

constexpr int kMaxPeers = 16;

struct PeerSlot { unsigned long long lo, hi; };      // {bits 0-31 | epoch << 32}, {bits 32-63 | epoch << 32}

struct DPeers {
    int       world, rank;         // world <= 1: no exchange
    uint32_t  epoch;               // 1, 2, ...: the same on every participant for one exchange
    uint32_t  pad;
    size_t    cap;                 // slots per row
    long long timeout;             // clocks a participant waits for its peers
    PeerSlot* buf[kMaxPeers];      // rows of participant r, as mapped here
    int*      status;              // local (mapped host memory): set to the epoch that timed out
};

__device__ __forceinline__ size_t peer_slot(const DPeers& p, int sender, size_t k)
{
    return (size_t(p.epoch & 1u) * size_t(p.world) + size_t(sender)) * p.cap + k;
}

// posted store of one double, tagged with the epoch, into a (possibly remote) slot
__device__ __forceinline__ void peer_put(PeerSlot* slot, double v, uint32_t epoch)
{
    const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(v));
    const unsigned long long tag = static_cast<unsigned long long>(epoch) << 32;
    const unsigned long long lo = (bits & 0xffffffffull) | tag, hi = (bits >> 32) | tag;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" :: "l"(slot), "l"(lo), "l"(hi) : "memory");
}

// polls a local slot until both words carry `epoch`; false when they have not by the deadline
__device__ __forceinline__ bool peer_get(const PeerSlot* slot, uint32_t epoch, long long timeout, double& v)
{
    const long long t0 = clock64();
    for (;;) {
        unsigned long long lo, hi;
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(slot) : "memory");
        if (uint32_t(lo >> 32) == epoch && uint32_t(hi >> 32) == epoch) {
            v = __longlong_as_double(static_cast<long long>((lo & 0xffffffffull) | (hi << 32)));
            return true;
        }
        if (clock64() - t0 > timeout) { v = 0.0; return false; }
    }
}

// The warp that owns output k: every lane holds the local sum s.  Returns (in every lane) the sum over participants in
// rank order, NaN when a participant never arrived.
__device__ __forceinline__ double peer_warp_sum(const DPeers& peers, size_t k, double s, int lane)
{
    if (lane < peers.world) peer_put(peers.buf[lane] + peer_slot(peers, peers.rank, k), s, peers.epoch);   // lane p -> peer p
    double v = 0.0;
    bool ok = true;
    if (lane < peers.world) ok = peer_get(peers.buf[peers.rank] + peer_slot(peers, lane, k), peers.epoch, peers.timeout, v);
    ok = __all_sync(0xffffffffu, ok ? 1 : 0) != 0;
    double tot = 0.0;
    for (int r = 0; r < peers.world; ++r) tot += __shfl_sync(0xffffffffu, v, r);      // rank order: identical everywhere
    if (!ok) {
        if (lane == 0 && peers.status) *reinterpret_cast<volatile int*>(peers.status) = int(peers.epoch);
        tot = __longlong_as_double(0x7ff8000000000000ll);
    }
    return tot;
}

// out[i] = sum over participants of their local[i], i < n (n <= cap).  Launched by every participant with the same
// epoch and the same grid; one thread per element, all pushes of a thread before its polls.
static __global__ void peer_exchange_kernel(const double* __restrict__ local, int n, double* __restrict__ out, const DPeers peers)
{
    const int stride = gridDim.x * blockDim.x, first = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = first; i < n; i += stride) {
        const double v = local[i];
        for (int p = 0; p < peers.world; ++p) peer_put(peers.buf[p] + peer_slot(peers, peers.rank, size_t(i)), v, peers.epoch);
    }
    long long patience = peers.timeout;                       // after one time-out the rest is not waited for again
    for (int i = first; i < n; i += stride) {
        double s = 0.0;
        bool ok = true;
        for (int r = 0; r < peers.world; ++r) {
            double v;
            if (!peer_get(peers.buf[peers.rank] + peer_slot(peers, r, size_t(i)), peers.epoch, patience, v)) { ok = false; patience = 0; }
            s += v;
        }
        if (!ok && peers.status) *reinterpret_cast<volatile int*>(peers.status) = int(peers.epoch);
        out[i] = ok ? s : __longlong_as_double(0x7ff8000000000000ll);
    }
}

}  // namespace cf
