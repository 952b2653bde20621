// common.cuh -- sm_100a PTX wrappers and shared definitions for the tsasr_b200 kernels.
//
// Everything here is hand-written inline PTX for Blackwell (tcgen05 / TMEM / TMA / mbarrier).
// No CUTLASS, no Triton.  Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor"
// and "instruction descriptor" tables (cross-checked against cute/arch/mma_sm100_desc.hpp).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <cstring>

namespace tsasr {

// ------------------------------------------------------------------------------------------
// Lattice layout ("skewed"): cell (b, t, u) lives at ((b * D + (t + u)) * U + u), D = T + U - 1.
// One anti-diagonal d = t + u of an utterance is contiguous in u, so the wavefront DP reads and
// writes fully coalesced rows, and prefetching k diagonals ahead needs no per-thread phase.
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ size_t skew_index(int b, int t, int u, int T, int U) {
    return ((size_t)b * (size_t)(T + U - 1) + (size_t)(t + u)) * (size_t)U + (size_t)u;
}

// Lengths as every kernel of the library reads them: clamped to the padded lattice (T_b in [1, T], label count in
// [0, U-1]).  The fused path validates lengths on the host only AFTER queueing its kernels, and three compat entry
// points (Transducer.apply, TransducerLoss on dense logits, rnnt_loss(check_lengths=False)) do not validate at all, so
// an out-of-range length must never index outside the utterance's slab -- and the DP and the gradient kernels must
// agree on the rectangle they work on.
__device__ __forceinline__ void clamped_lengths(const int* __restrict__ logit_lengths, const int* __restrict__ target_lengths,
                                                int b, int T, int U, int& Tb, int& Ub) {
    Tb = min(max(logit_lengths[b], 1), T);
    Ub = min(max(target_lengths[b], 0), U - 1) + 1;
}

// ------------------------------------------------------------------------------------------
// Live cell tiles.  Tiles are numbered densely, tile = ((b * nTt + tt) * nTu + tu); those entirely outside the
// utterance's T_b x U_b rectangle ("dead": ~40 % of a ragged batch) carry no data.  The kernels work on the k-th
// LIVE tile, k = 0, 1, ...: per (b, tt) group either all ceil(U_b / tU) leading tiles are live or none, so a
// cursor walking the groups finds the k-th live tile without any list in memory.  seek() is O(1) amortised for
// non-decreasing k (it rewinds otherwise); every role of a kernel runs its own cursor and sees the same sequence.
// P needs: logit_lengths, target_lengths, T, U, tT_log2, nTt, nTu, tile_begin, tile_end (whole groups), and
// active_ids / active_count: when active_ids != nullptr the backward kernels work on that explicit, ordered list of
// dense tile ids instead (the live tiles whose lattice occupancy is not negligible, see tile_activity_kernel).
// ------------------------------------------------------------------------------------------
template <typename P>
struct LiveCursor {
    // current group (b, tt) = dense group g; base = live tiles before it; after a successful seek tu is the tile's
    // label-tile index inside the group.  dense = every tile of the range is live (full-length batch): the k-th live
    // tile is simply tile_begin + k and the cursor degenerates to stateless index arithmetic.
    int g, g_end, base, b, tt, tu, Tb, Ub, live_tt, n_u;
    bool dense;
    int n_list, pf_k, pf_tile, pf_tile1;  // list mode (n_list >= 0): number of ids, prefetched entries pf_k and pf_k + 1
    __device__ __forceinline__ void load_utterance(const P& p) {
        // lengths are clamped to the padded lattice (validated on the host after launching)
        const int tT = 1 << p.tT_log2, tU = 128 >> p.tT_log2;
        Tb = min(max(p.logit_lengths[b], 1), p.T);
        Ub = min(max(p.target_lengths[b], 0), p.U - 1) + 1;
        live_tt = (Tb + tT - 1) >> p.tT_log2;
        n_u = (Ub + tU - 1) / tU;
    }
    __device__ __forceinline__ void rewind(const P& p) {
        g = p.tile_begin / p.nTu;
        b = g / p.nTt;
        tt = g - b * p.nTt;
        base = 0;
        if (g < g_end) load_utterance(p);
    }
    // n_live: count_live_tiles_warp(p) of the same range (every role computes it once, as a converged warp)
    __device__ __forceinline__ LiveCursor(const P& p, int n_live)
        : g_end(p.tile_end / p.nTu), tu(0), dense(n_live == p.tile_end - p.tile_begin && p.active_ids == nullptr),
          n_list(p.active_ids ? n_live : -1), pf_k(-2), pf_tile(0), pf_tile1(0) {
        if (dense || n_list >= 0) { g = 0; base = 0; b = -1; tt = 0; Tb = 0; Ub = 0; live_tt = 0; n_u = 0; }
        else rewind(p);
    }
    // list mode: start fetching entry k now (the id is needed at the next seek(k); hides the load latency)
    __device__ __forceinline__ void hint(const P& p, int k) {
        if (n_list >= 0 && k < n_list) { pf_k = k; pf_tile = __ldg(p.active_ids + k); pf_tile1 = -1; }
    }
    // same for the entries k and k + 1 of a tile pair (k even: one 8-byte load)
    __device__ __forceinline__ void hint_pair(const P& p, int k) {
        if (n_list >= 0 && k + 1 < n_list) {
            const int2 v = __ldg(reinterpret_cast<const int2*>(p.active_ids + k));
            pf_k = k; pf_tile = v.x; pf_tile1 = v.y;
        } else {
            hint(p, k);
        }
    }
    __device__ __forceinline__ void set_tile(const P& p, int tile) {
        g = tile / p.nTu;
        tu = tile - g * p.nTu;
        const int bb = g / p.nTt;
        tt = g - bb * p.nTt;
        if (bb != b) { b = bb; load_utterance(p); }
    }
    // positions the cursor on the k-th live tile of [tile_begin, tile_end); false when there are fewer
    __device__ __forceinline__ bool seek(const P& p, int k) {
        if (n_list >= 0) {
            if (k >= n_list) return false;
            set_tile(p, k == pf_k ? pf_tile : (k == pf_k + 1 && pf_tile1 >= 0) ? pf_tile1 : __ldg(p.active_ids + k));
            return true;
        }
        if (dense) {
            const int tile = p.tile_begin + k;
            if (tile >= p.tile_end) return false;
            set_tile(p, tile);
            return true;
        }
        if (k < base) rewind(p);
        for (;;) {
            if (g >= g_end) return false;
            const int n = tt < live_tt ? n_u : 0;
            if (k < base + n) { tu = k - base; return true; }
            base += n;
            ++g;
            if (++tt == p.nTt) {
                tt = 0;
                ++b;
                if (g < g_end) load_utterance(p);
            }
        }
    }
    __device__ __forceinline__ int tile(const P& p) const { return g * p.nTu + tu; }  // dense tile id
};

// Number of live tiles of [tile_begin, tile_end), computed cooperatively by one converged warp (lane l takes
// utterances l, l+32, ...).  Roles that only need to know HOW MANY rounds there are (MMA issue, W stream, relays)
// use this once instead of walking a cursor on their critical path.
template <typename P>
__device__ __forceinline__ int count_live_tiles_warp_geometric(const P& p) {
    const int lane = threadIdx.x & 31;
    const int g0 = p.tile_begin / p.nTu, g1 = p.tile_end / p.nTu;
    const int b0 = g0 / p.nTt, b1 = (g1 + p.nTt - 1) / p.nTt;
    const int tT = 1 << p.tT_log2, tU = 128 >> p.tT_log2;
    int n = 0;
    for (int b = b0 + lane; b < b1; b += 32) {
        const int Tb = min(max(p.logit_lengths[b], 1), p.T), Ub = min(max(p.target_lengths[b], 0), p.U - 1) + 1;
        const int live_tt = (Tb + tT - 1) >> p.tT_log2, n_u = (Ub + tU - 1) / tU;
        const int lo = max(0, g0 - b * p.nTt), hi = min(min(p.nTt, g1 - b * p.nTt), live_tt);
        n += max(0, hi - lo) * n_u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    return n;
}
template <typename P>
__device__ __forceinline__ int count_live_tiles_warp(const P& p) {
    if (p.active_ids) return *p.active_count;  // explicit list (backward tile pruning)
    return count_live_tiles_warp_geometric(p);
}

enum ActKind : int { ACT_LEAKY_RELU = 0, ACT_RELU = 1, ACT_TANH = 2, ACT_IDENTITY = 3 };

__device__ __forceinline__ float act_apply(float x, int kind, float param) {
    switch (kind) {
        case ACT_LEAKY_RELU: return x >= 0.f ? x : x * param;
        case ACT_RELU: return fmaxf(x, 0.f);
        case ACT_TANH: return tanhf(x);
        default: return x;
    }
}
__device__ __forceinline__ float logaddexp_fast(float a, float b) {
    // max + log1p(exp(-|a-b|)); -inf safe (both -inf -> -inf)
    const float m = fmaxf(a, b);
    if (m == -INFINITY) return -INFINITY;
    return m + __logf(1.f + __expf(-fabsf(a - b)));
}

// ------------------------------------------------------------------------------------------
// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start (set up barriers, allocate TMEM, prefetch descriptors) while its predecessor in the stream is still
// draining; griddep_wait() blocks until the predecessor has completed and its writes are visible.  Every thread
// calls it before the first access to data the predecessor produced.  Without the launch attribute it is a no-op.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// host side: appends the launch attribute (TSASR_DEBUG_NO_PDL=1 switches it off for A/B runs); returns the new count
inline int pdl_launch_attr(cudaLaunchAttribute* attrs, int n) {
    static const bool on = getenv("TSASR_DEBUG_NO_PDL") == nullptr;
    if (!on) return n;
    attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[n].val.programmaticStreamSerializationAllowed = 1;
    return n + 1;
}
// <<<grid, block, smem, stream>>> with the attribute
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    cfg.attrs = attr;
    cfg.numAttrs = pdl_launch_attr(attr, 0);
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------------
// shared-memory address helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Fail-stop instead of a silent hang.  A wait of these kernels lasts microseconds; a protocol bug would spin forever and
// wedge the process (and the GPU box) until somebody resets the device.  The wait therefore carries a WALL-CLOCK bound
// (%globaltimer, checked every 4096 failed probes): after kMbarTimeoutNs without progress the CTA records the barrier
// id and traps, which surfaces as a CUDA error on the host.  The bound is time based -- not a probe count -- so that
// tools that slow a kernel down by orders of magnitude (ncu replay, compute-sanitizer, a debugger, MPS time slicing)
// stay far away from it: 30 s by default, 2 s in -DTSASR_DEBUG builds (development: fail fast), and compiled out
// entirely with -DTSASR_NO_WATCHDOG.
#if defined(TSASR_DEBUG)
static constexpr unsigned long long kMbarTimeoutNs = 2000000000ull;
#else
static constexpr unsigned long long kMbarTimeoutNs = 30000000000ull;
#endif
static __device__ unsigned int g_tsasr_hang_info[4];
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t tag = 0) {
#if defined(TSASR_NO_WATCHDOG)
    while (!mbar_try_wait(bar, parity)) {}
#else
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 4095u) == 0u) {  // a failed probe already suspends the warp for a while: this is rare
            const unsigned long long now = globaltimer_ns();
            if (t0 == 0) {
                t0 = now;
            } else if (now - t0 > kMbarTimeoutNs) {
                g_tsasr_hang_info[0] = 0xDEAD0000u | tag;
                g_tsasr_hang_info[1] = blockIdx.x;
                g_tsasr_hang_info[2] = threadIdx.x;
                g_tsasr_hang_info[3] = parity;
                __threadfence_system();
                __trap();
            }
        }
    }
#endif
}

// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor / cp.async.bulk)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
// 1-D bulk copy global -> shared (used for pre-swizzled operand images)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// 1-D bulk copy shared -> global
__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_group() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------------
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of a 2-CTA cluster issue ONE M=256 MMA; each supplies its own
// 128 A rows and half of the B rows, the leader (rank 0) issues, accumulators live in both TMEMs.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes)
                 : "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once all MMAs issued so far completed) on the barrier at this shared offset in EVERY CTA of cta_mask
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
// TMA load into THIS CTA's shared memory, completing bytes on the barrier at cluster address `mbar_cluster`
// (the pair leader's barrier)
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const void* tmap, uint32_t mbar_cluster, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(mbar_cluster), "r"(x), "r"(y)
        : "memory");
}

// Same, but the box is written into the shared memory (same offset) of every CTA in `cta_mask`; with
// cta_group::2 the completion is signalled on the barrier at `mbar_local`'s offset in the even CTA of each
// destination pair.  The pair leader uses it (mask 0b10) to fill its partner's half of a W stage itself,
// so a stage refill never waits on the partner's producer thread.
__device__ __forceinline__ void tma_load_2d_2cta_mcast(void* smem_dst, const void* tmap, uint64_t* mbar_local, uint16_t cta_mask,
                                                       int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%4, %5}], [%2], %3;"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(mbar_local) & 0xFEFFFFFFu), "h"(cta_mask), "r"(x), "r"(y)
        : "memory");
}

// ------------------------------------------------------------------------------------------
// "_e" variants: executed by ALL 32 lanes of a converged warp, one lane is elected inside the asm.
// Issuing from `if (lane == 0)` instead makes the compiler treat every operand as divergent and wrap each
// UTCHMMA / UTCBAR / UTMALDG in an elect + R2UR.BROADCAST loop (~75 cycles per instruction, measured);
// with warp-uniform operands the descriptors live in uniform registers and issue back-to-back.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void umma_bf16_e(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta_e(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_e(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta_e(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
        ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_e(uint64_t* bar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
        ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d_e(void* smem_dst, const void* tmap, uint64_t* bar, int x, int y) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta_mcast_e(void* smem_dst, const void* tmap, uint64_t* mbar_local, uint16_t cta_mask,
                                                         int x, int y) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%4, %5}], [%2], %3;\n\t}"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(mbar_local) & 0xFEFFFFFFu), "h"(cta_mask), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta_e(void* smem_dst, const void* tmap, uint32_t mbar_cluster, int x, int y) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(mbar_cluster), "r"(x), "r"(y) : "memory");
}

__device__ __forceinline__ void bulk_load_1d_e(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}"
        ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Hopper-style multicast inside a cluster (cta_group::1 MMAs): one CTA loads a box and writes it to the same
// shared offset of every CTA in cta_mask, completing bytes on each destination CTA's own barrier; an MMA
// commit arrives on the barrier at the same offset of every CTA in cta_mask.
__device__ __forceinline__ void tma_load_2d_mcast_e(void* smem_dst, const void* tmap, uint64_t* bar, uint16_t cta_mask, int x, int y) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%4, %5}], [%2], %3;\n\t}"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "h"(cta_mask), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void bulk_load_1d_mcast_e(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n\t}"
        ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mcast_e(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
        ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

// Shared-memory matrix descriptor (PTX ISA "Matrix Descriptor Format", sm_100 version field = 1).
//   bits [0,14)  start address >> 4        bits [16,30) leading-dim byte offset >> 4
//   bits [32,46) stride-dim byte offset >> 4   bits [46,48) version (1)
//   bits [61,64) swizzle: 0 none, 2 = 128B, 4 = 64B, 6 = 32B
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulation.
//   [4,6) D fmt (1 = f32)  [7,10) A fmt (1 = bf16)  [10,13) B fmt (1 = bf16)
//   [15] A major (0 = K, 1 = MN)  [16] B major  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Byte offset of element (row r, 16-byte chunk c) inside a [rows x 64 bf16] SWIZZLE_128B block
// (rows of 128 B, 8-row / 1024-B swizzle atoms, block base 1024-B aligned): chunk index is
// XOR-ed with (r & 7).  This is the image TMA writes with CU_TENSOR_MAP_SWIZZLE_128B and the
// image tcgen05.mma reads for both K-major and MN-major SW128 descriptors.
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t c) {
    return r * 128u + ((c ^ (r & 7u)) << 4);
}

// 2^x, single MUFU.EX2 (flush-to-zero); -inf -> 0
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Packed fp32x2 arithmetic (sm_100): one issue slot for two lanes of work in the epilogues.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(reinterpret_cast<unsigned long long&>(d))
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
          "l"(reinterpret_cast<unsigned long long&>(c)));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("add.rn.f32x2 %0, %1, %2;"
        : "=l"(reinterpret_cast<unsigned long long&>(d))
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    float2 d;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(reinterpret_cast<unsigned long long&>(d))
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return d;
}

// Four (or two) consecutive K=16 MMAs of one pipeline stage in ONE asm statement: the descriptors advance by
// 32 bytes (2 x 16-byte units) per K step inside the swizzled row; constants are broadcast to uniform
// registers once per stage instead of once per instruction.
__device__ __forceinline__ void umma_bf16_2cta_x4_e(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                    uint32_t accumulate_first) {
    asm volatile(
        "{\n\t.reg .pred p, q, t;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "add.u64 a1, %1, 2;\n\tadd.u64 a2, %1, 4;\n\tadd.u64 a3, %1, 6;\n\t"
        "add.u64 b1, %2, 2;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 b3, %2, 6;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %3, t;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %3, t;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %3, t;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate_first) : "memory");
}
__device__ __forceinline__ void umma_bf16_x2_e(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate_first) {
    asm volatile(
        "{\n\t.reg .pred p, q, t;\n\t.reg .b64 a1, b1;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "add.u64 a1, %1, 2;\n\tadd.u64 b1, %2, 2;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate_first) : "memory");
}

// one 256-bit global store (sm_100: STG.256); dst must be 32-byte aligned
__device__ __forceinline__ void st_global_v8(void* dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f,
                                             uint32_t g, uint32_t h) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e),
                 "r"(f), "r"(g), "r"(h)
                 : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

}  // namespace tsasr
