// exchange.cuh -- device side of one exchange site (protocol: include/mgb200.h, host side: comm.cu).  Shared by the
// stand-alone exchange kernel (comm.cu) and by the SELL kernels that carry a site as extra CTAs (sell_kernels.cu).
#pragma once
#include "common.cuh"

namespace mgb {

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Wire format ("LL", as in low-latency collectives): every double travels as two 8-byte words, each holding 4 data
// bytes and the low 32 bits of the program's epoch as a tag.  An aligned 8-byte store lands atomically, so the
// receiver simply polls each word until its tag matches: no fence, no separate flag, one NVLink traversal of latency.
// Staging slots written two programs ago (same parity buffer) carry an older tag and can never match.  Messages of
// length zero still shake hands through the flag word, which keeps the run-ahead argument of mgb200.h intact.
struct ExPeer {
    const int32_t *send_idx;
    int64_t send_off, send_cnt;
    ulonglong2 *peer_stage;             // in the PEER's arena: where my message lands (parity 0)
    unsigned long long *peer_flag;      // in the PEER's arena: flags[my rank][site]
    const ulonglong2 *my_stage;         // in MY arena: where the peer's message lands (parity 0)
    const unsigned long long *my_flag;  // in MY arena: flags[peer][site]
    const int32_t *recv_idx;
    int64_t recv_off, recv_cnt;
};
struct ExArgs {
    int npeers, ctas_per_peer;
    const double *src;
    double *dst;
    const unsigned long long *epoch;
    unsigned int *err;
    int64_t parity_stride;              // 16-byte packets between the two staging buffers of a region
    unsigned long long timeout_ns;
    unsigned int site;
    int dry;                            // warm-up launch: do nothing
    // fused into a compute kernel: the exchange CTAs count themselves in *done; the last one publishes
    // epoch * 65536 + site + 1 in *ready, which the compute CTAs that read halo columns wait for
    unsigned int *done;
    unsigned long long *ready;
    ExPeer p[MG_MAX_RANKS];
};

__device__ __forceinline__ void st_packet(ulonglong2 *p, unsigned long long w0, unsigned long long w1) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ void ld_packet(const ulonglong2 *p, unsigned long long &w0, unsigned long long &w1) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}


// one CTA (index bid of npeers * ctas_per_peer) of a site: push my chunk, poll and unpack the peer's chunk
__device__ __forceinline__ void exchange_role(const ExArgs &a, int bid) {
    const int p = bid / a.ctas_per_peer, chunk = bid % a.ctas_per_peer;
    const ExPeer &P = a.p[p];
    const unsigned long long epoch = *a.epoch;
    const unsigned long long tag = (epoch & 0xffffffffull) << 32;
    const int64_t par = (int64_t)(epoch & 1ull) * a.parity_stride;
    const int64_t stride = (int64_t)a.ctas_per_peer * kBlock;
    const int64_t first = (int64_t)chunk * kBlock + threadIdx.x;
    // ---- push: never waits for anybody
    {
        ulonglong2 *out = P.peer_stage + par;
        for (int64_t i = first; i < P.send_cnt; i += stride) {
            const unsigned long long bits =
                (unsigned long long)__double_as_longlong(P.send_idx ? a.src[P.send_idx[i]] : a.src[P.send_off + i]);
            st_packet(out + i, (bits & 0xffffffffull) | tag, (bits >> 32) | tag);
        }
    }
    if (P.send_cnt == 0 && first == 0) st_relaxed_sys(P.peer_flag, epoch);
    // ---- receive: poll every packet until both words carry this program's tag
    const ulonglong2 *in = P.my_stage + par;
    bool timed_out = false;
    for (int64_t i = first; i < P.recv_cnt; i += stride) {
        unsigned long long w0, w1;
        ld_packet(in + i, w0, w1);
        if ((w0 & 0xffffffff00000000ull) != tag || (w1 & 0xffffffff00000000ull) != tag) {
            const unsigned long long t0 = global_timer_ns();
            unsigned int spins = 0;
            for (;;) {
                ld_packet(in + i, w0, w1);
                if ((w0 & 0xffffffff00000000ull) == tag && (w1 & 0xffffffff00000000ull) == tag) break;
                if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > a.timeout_ns) { timed_out = true; break; }
            }
            if (timed_out) break;
        }
        const double v = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
        if (P.recv_idx) a.dst[P.recv_idx[i]] = v;
        else a.dst[P.recv_off + i] = v;
    }
    if (P.recv_cnt == 0 && first == 0) {
        const unsigned long long t0 = global_timer_ns();
        unsigned int spins = 0;
        while (ld_relaxed_sys(P.my_flag) < epoch) {
            __nanosleep(20);
            if ((++spins & 255u) == 0 && global_timer_ns() - t0 > a.timeout_ns) { timed_out = true; break; }
        }
    }
    if (timed_out) atomicCAS(a.err, 0u, a.site + 1u);
}

// an exchange site riding on a SELL launch (prepared by comm_prepare)
struct SellFuse {
    ExArgs ex;
    int nex;                       // exchange CTAs in front of the compute CTAs
    const unsigned char *mask;     // per slice of the matrix: reads halo columns (NULL: assume every slice does)
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// exchange CTA of a fused launch: do the site (unless dry), then count in; the last CTA publishes `ready`
__device__ __forceinline__ void fused_exchange_cta(const ExArgs &a, int bid) {
    if (a.dry) return;          // warm-up launch: nothing is published (the epoch stands still, see fused_wait_ready)
    exchange_role(a, bid);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int total = (unsigned)(a.npeers * a.ctas_per_peer);
        const unsigned int prev = atomicAdd(a.done, 1u);
        if (prev == total - 1u) {
            *a.done = 0u;
            __threadfence();
            st_release_gpu_u64(a.ready, *a.epoch * 65536ull + a.site + 1ull);
        }
    }
}

// compute CTA of a fused launch whose rows read halo columns: wait until the site has been unpacked.
// Scheduling assumption: the exchange CTAs have the LOWEST block indices of the launch and CUDA dispatches the blocks
// of a grid in index order, so they are resident (or already done) before any compute CTA can spin here; the wait is
// bounded by timeout_ns all the same, and a time-out is raised by the host (DistributedHierarchy.check, called at the
// end of every solve), never swallowed.
__device__ __forceinline__ void fused_wait_ready(const ExArgs &a) {
    if (a.dry) return;
    const unsigned long long want = *a.epoch * 65536ull + a.site + 1ull;
    if (ld_acquire_gpu_u64(a.ready) >= want) return;
    const unsigned long long t0 = global_timer_ns();
    unsigned int spins = 0;
    while (ld_acquire_gpu_u64(a.ready) < want) {
        __nanosleep(32);
        if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > a.timeout_ns) {
            atomicCAS(a.err, 0u, a.site + 1u);
            break;
        }
    }
}

}  // namespace mgb
