// joint_gemm.cuh -- the fused joint network GEMM for sm_100a (tcgen05 + TMEM + TMA).
//
//   logits[m, v] = sum_h J[m, h] * W[v, h] + bias[v],   J[m, :] = bf16(act(enc[b,t,:] + dec[b,u,:]))
//
// replaces Transducer_joint.forward (SB/nnet/transducer/transducer_joint.py:73-74,95), the head
// Linear (SB/nnet/linear.py:74) and the log-softmax that follows (SB/nnet/losses.py:72-79,84) without
// ever writing J ([B,T,U,H]) or the logits ([B,T,U,V]) to HBM.
//
// One persistent CTA per SM, 20 warps (640 threads; Roles<8>, the default form):
//   warp 16      TMA producer: streams W k-slices through a 4-stage ring (3 when the shared memory is short); the MMA lane
//                consumes them two at a time (8 MMAs per wait / commit group)
//   warp 19      allocates TMEM, then ONE lane issues tcgen05.mma (N<=256, K=16, bf16 -> fp32); highest warp id
//                = highest issue priority, because this lane is the serial resource of the kernel.
//                (pair partner: relays "accumulator released" from its epilogue warps to the leader)
//   warp 18      pair partner only: relays "A block written" from its producer warps to the leader
//   warps 8-15   A producers: build the 128-cell x H operand in shared memory in the canonical
//                K-major SWIZZLE_128B layout (broadcast add + activation + bf16 round fused here) from enc/dec
//                rows read straight from global memory (L1/L2 resident): one enc load + four dec loads per thread and
//                k-block (a thread's four rows share their frame), two k-blocks prefetched ahead;
//                the operand stays resident for all N tiles of the cell tile
//   warps 0-7    epilogue, two groups of four warps; both groups work on every accumulator buffer (2 x 256 TMEM
//                columns), group g on columns [128 g, 128 g + 128).  MODE_FWD: online
//                log-softmax keeping {lp_blank, lp_emit, logZ} (the groups merge their running
//                (max, sum) through shared memory at the end of a cell tile); MODE_GRAD: recompute the
//                softmax and emit bf16 dlogits tiles as pre-swizzled operand images for the backward
//                GEMMs; MODE_DEBUG: dump raw logits (tests only).
//
// What paces the kernel (TSASR_DEBUG_SKIP ablations + epilogue cycle counters, B200, config 2,
// profiles/r1_ablation_joint_epilogue.txt): a pair round is 20.5k cycles of MMAs, but with the MMA issue switched OFF the
// round still takes 32.5k cycles -- the epilogue needs ~6k cycles per accumulator (two in-order warps per SM
// sub-partition cannot hide the scale -> max tree -> exp2 -> sum chain) against 5.1k for the MMAs that fill it.  The
// rare per-chunk work (blank / label picks, vocabulary tail) therefore sits behind ONE warp-uniform branch per chunk.
//
// PAIR = true runs the same roles on CTA pairs (2-CTA clusters, tcgen05 cta_group::2): the pair works on two
// cell tiles at once with ONE M=256 MMA stream issued by the leader CTA; each CTA keeps its own A operand,
// loads only HALF of every W stage and owns the accumulator rows of its own tile.  That halves the
// L2->SM traffic of the W stream (the measured bottleneck of the single-CTA kernel) and doubles the
// tensor work per issued instruction.
//
// NPW = 16 is the form for NARROW vocabularies (one vocabulary tile of at most 128 columns per cell tile: the recipe as
// shipped has 29 characters): warps 0-3 epilogue (one column group), warps 4-19 A producers, 20 TMA, 22 relay, 23 MMA
// (Roles<16>, 768 threads); and when the CTA's share of W fits the W area (JointParams::w_resident) every k-slice of W is
// loaded ONCE and stays for the whole kernel.  With one small MMA group per k-block the W ring's refill latency and the
// A producers pace the kernel, not the tensor pipe (profiles/r2_narrow_vocabulary.txt: forward 796 -> 508 us at V = 29).
//
// A cell tile is tT consecutive frames x tU consecutive label positions of one utterance
// (tT * tU = 128, tT in {8,16,32}, row r = ui * tT + ti); tiles completely outside the utterance's
// T_b x U_b rectangle are skipped by every role.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace tsasr {

static constexpr int kTileM = 128;        // cells per tile (UMMA M)
static constexpr int kTileN = 256;        // vocabulary columns per accumulator (UMMA N max)
static constexpr int kABlockK = 64;       // A k-block: 64 bf16 = one 128-byte swizzle row
static constexpr int kABlockBytes = kTileM * kABlockK * 2;  // 16 KB
// W stage (16 KB either way): single CTA  [256 v x 32 h] SWIZZLE_64B  (2 MMAs of K=16 per stage)
//                            CTA pair    [128 v x 64 h] SWIZZLE_128B per CTA (4 pair-MMAs per stage)
static constexpr int kWStageK = 32;
static constexpr int kWStageKPair = 64;
static constexpr int kWStageBytes = kTileN * kWStageK * 2;  // 16 KB
static constexpr int kMaxKB = 10;         // H <= 640
static constexpr int kMaxWStages = 4;
// Warp roles.  The SM's warp arbiter serves the highest warp id first, and the single MMA-issuing lane is the serial
// resource of the kernel, so it gets the highest id; the TMA lanes come next, the bulk workers last.
//   NPW = 8  (default): warps 0-7 epilogue (two column groups), 8-15 A producers, 16 TMA, 18 relay, 19 MMA  (640 threads)
//   NPW = 16 (narrow vocabularies, V <= 128: ONE vocabulary tile per cell tile, so the A operand is built once per
//             32..128 columns of MMA work and its producers pace the kernel -- the recipe as shipped has V = 29):
//             warps 0-3 epilogue (one column group), 4-19 A producers, 20 TMA, 22 relay, 23 MMA            (768 threads)
template <int NPW>
struct Roles {
    static constexpr int kEpilogueWarps = NPW == 16 ? 4 : 8;  // TMEM lane quarter = warp & 3
    static constexpr int kFirstProducerWarp = kEpilogueWarps;
    static constexpr int kWarpTmaW = kEpilogueWarps + NPW;
    static constexpr int kWarpRelayA = kWarpTmaW + 2;  // partner CTA of a pair only
    static constexpr int kWarpMma = kWarpTmaW + 3;     // leader: MMA issue; partner: accumulator-release relay
    static constexpr int kThreads = (kWarpTmaW + 4) * 32;
    static constexpr int kRowsPerThread = 32 / NPW;    // A rows per producer thread: 4 (rows rg + 32 i) or 2 (rows rg + 64 i)
    static constexpr int kRowStride = 4 * NPW;
};
static constexpr float kLog2eF = 1.4426950408889634f;
static constexpr float kLn2F = 0.6931471805599453f;

enum JointMode : int { MODE_FWD = 0, MODE_GRAD = 1, MODE_DEBUG = 2, MODE_GRAD_CLAMP = 3 };
// MODE_GRAD_CLAMP: MODE_GRAD with torchaudio's `clamp` (rnnt_loss(clamp=c)): the gradient of the UNIT cost is clamped
// to [-c, c] and multiplied by the upstream factor dcost[b] afterwards (ComputeGradients, then `grad * dy`,
// torchaudio/functional/functional.py:1729-1734).  A separate instantiation, so the default path pays nothing.
template <int MODE>
struct ModeTraits { static constexpr bool grad = MODE == MODE_GRAD || MODE == MODE_GRAD_CLAMP; static constexpr bool clamp = MODE == MODE_GRAD_CLAMP; };

struct JointParams {
    const float* bias;          // [V]
    const int* targets;         // [B,U-1]
    const int* logit_lengths;   // [B]
    const int* target_lengths;  // [B]
    int B, T, U, H, V, blank;
    int act_kind;
    float act_param;
    int tT_log2;                // tile = (1 << tT_log2) frames x (128 >> tT_log2) labels
    int nTt, nTu;               // tiles per utterance along t and u
    int tile_begin, tile_end;   // tile id range processed by this launch
    int KB;                     // H / 64
    int NT;                     // ceil(V / 256)
    int n_last;                 // UMMA N of the last vocabulary tile (multiple of 16)
    int num_w_stages;
    int w_resident;             // pairs, one narrow vocabulary tile (NT == 1): all KB k-slices of W ([n_last / 2 rows x 64 h] per CTA) are loaded ONCE
                                // and stay in the W area for the whole kernel -- no W stream, no w_full / w_empty traffic per cell tile
    int w_stage_bytes;          // distance of consecutive W k-slices in shared memory (kWStageBytes unless w_resident)
    int dbg_skip;               // development ablations (bits): 1 epilogue only releases, 2 producers only arrive, 4 no W stream,
                                // 8 epilogue loads but no math, 16 epilogue math but no TMEM loads, 32 MMA issue skipped
    const __nv_bfloat16* enc;   // [B,T,H]
    const __nv_bfloat16* dec;   // [B,U,H]
    // MODE_FWD outputs (skewed lattice layout)
    float2* lat2;
    float* logz;
    // MODE_GRAD inputs / outputs
    const float2* lat2_in;
    const float* logz_in;
    const float* alpha;
    const float* beta;
    const float* cost;
    const float* dcost;
    float clamp;                // MODE_GRAD_CLAMP: bound on the unit-cost gradient (> 0)
    __nv_bfloat16* dY_img;      // [tile - tile_begin][NT*4][128 x 64] SWIZZLE_128B images
    __nv_bfloat16* J_img;       // [tile - tile_begin][KB][128 x 64] SWIZZLE_128B images
    // MODE_DEBUG output
    float* dbg_logits;          // [B,T,U,V]
    long long* prof;            // development: per-CTA cycle counters of the MMA lane (or nullptr)
    // MODE_GRAD with tile pruning: ordered list of the chunk's active tiles (nullptr: every live tile)
    const int* active_ids;
    const int* active_count;
};

struct TileCoord { int b, t0, u0, Tb, Ub; bool live; };

// Work distribution over the LIVE tiles (common.cuh, LiveCursor).  Single CTA: CTA c takes live tiles c, c+G, ...
// Pair: cluster c takes live-tile pairs (2c, 2c+1), (2c+2C, ...), rank r works on the pair's r-th tile.  Only the
// very last pair can be incomplete: the CTA without a tile still runs the barrier protocol ("dummy" round)
// because the pair's MMA stream is shared.
template <bool PAIR>
struct Rounds {
    int next_round, step, rank;
    LiveCursor<JointParams> cur;
    // n_live = count_live_tiles_warp(p), computed by the calling (converged) warp
    __device__ __forceinline__ Rounds(const JointParams& p, int n_live) : cur(p, n_live) {
        rank = PAIR ? (int)cluster_ctarank() : 0;
        next_round = PAIR ? (int)cluster_id_x() : (int)blockIdx.x;
        step = PAIR ? (int)num_clusters_x() : (int)gridDim.x;
    }
    // rounds of this unit when the number of live tiles is known (roles that need no tile coordinates)
    __device__ __forceinline__ int count_my_rounds(int n_live) const {
        const int total = PAIR ? (n_live + 1) >> 1 : n_live;
        return next_round < total ? (total - next_round + step - 1) / step : 0;
    }
    // advances to this unit's next round; false when the live tiles are exhausted.  tc describes this CTA's own tile
    // (tc.live = false, my_tile = -1: the missing partner tile of the last, odd pair).
    __device__ __forceinline__ bool next(const JointParams& p, int& my_tile, TileCoord& tc) {
        const int k0 = PAIR ? 2 * next_round : next_round;
        next_round += step;
        if (!cur.seek(p, k0)) return false;
        tc.live = true;
        if (PAIR && rank) tc.live = cur.seek(p, k0 + 1);
        // list mode (pruned backward): the ids of this unit's NEXT round are fetched now, off the critical path
        if (PAIR) cur.hint_pair(p, 2 * next_round);
        else cur.hint(p, next_round);
        if (tc.live) {
            my_tile = cur.tile(p);
            tc.b = cur.b;
            tc.t0 = cur.tt << p.tT_log2;
            tc.u0 = cur.tu * (kTileM >> p.tT_log2);
            tc.Tb = cur.Tb;
            tc.Ub = cur.Ub;
        } else {
            my_tile = -1;
            tc.b = 0; tc.t0 = 0; tc.u0 = 0; tc.Tb = 0; tc.Ub = 0;
        }
        return true;
    }
};

// shared memory carve-up (dynamic; base must be 1024-byte aligned)
struct SmemLayout {
    uint32_t a_off, w_off, bias_off, xchg_off, bar_off, tmem_off, total;
};
__host__ __device__ inline SmemLayout smem_layout(int KB, int num_w_stages) {
    SmemLayout l;
    l.a_off = 0;
    l.w_off = l.a_off + (uint32_t)KB * kABlockBytes;
    l.bias_off = l.w_off + (uint32_t)num_w_stages * kWStageBytes;
    l.xchg_off = l.bias_off + 2 * (kTileN / 2) * 4;  // one 128-float bias slot per epilogue group
    l.bar_off = l.xchg_off + kTileM * 8;              // 128 x float2, used twice per cell tile (MODE_FWD merge)
    // barriers: w_full[4] w_empty[4] (4 spare) a_full[10] a_empty[10] acc_full[2] acc_empty[2] = 36,
    // a_done[10] acc_done[2] (partner CTA only: local collection points relayed to the leader) = 48
    l.tmem_off = l.bar_off + 48 * 8;
    l.total = l.tmem_off + 16;
    return l;
}

// ---- activation, specialised at compile time inside the producer ----
template <int ACT>
__device__ __forceinline__ float act_t(float x, float param) {
    if (ACT == ACT_LEAKY_RELU) return x >= 0.f ? x : x * param;
    if (ACT == ACT_RELU) return fmaxf(x, 0.f);
    if (ACT == ACT_TANH) return tanhf(x);
    return x;
}

// A producers: J k-blocks straight from enc / dec in global memory (L1/L2 resident: every enc row of a tile is
// read by tU threads, every dec row by tT).  The loads run two k-blocks ahead of the block being processed,
// so their latency hides behind the arithmetic and the a_empty wait.
template <int MODE, int ACT, bool PAIR, int NPW>
__device__ __forceinline__ void produce_a(const JointParams& p, uint8_t* smem_a, const __nv_bfloat16* __restrict__ enc,
                                          const __nv_bfloat16* __restrict__ dec, uint64_t* a_full, uint64_t* a_empty) {
    using R = Roles<NPW>;
    constexpr int NR = R::kRowsPerThread, RS = R::kRowStride;
    Rounds<PAIR> rounds(p, count_live_tiles_warp(p));
    // The partner CTA's producers arrive on a LOCAL barrier (a_full points at a_done there); a relay warp
    // forwards it to the leader.  A remote arrive has cluster-scope release semantics (MEMBAR.GPU), which in
    // MODE_GRAD would make every producer warp wait for its J-image global stores to drain.
    const int lane = threadIdx.x & 31;
    const int ptid = threadIdx.x - R::kFirstProducerWarp * 32;  // 0 .. 32 NPW - 1
    const int c = ptid & 7;                                  // 16-byte chunk inside the 128-byte row
    const int rg = ptid >> 3;                                // rows rg + RS i, i < NR
    const int tT = 1 << p.tT_log2, tTm = tT - 1;
    const int KB = p.KB;
    uint32_t it = 0;
    const bool slope_le1 = p.act_param >= 0.f && p.act_param <= 1.f;
    int ti[NR], ui[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        const int r = rg + RS * i;
        ti[i] = r & tTm;
        ui[i] = r >> p.tT_log2;
    }
    for (;;) {
        int tile;
        TileCoord tc;
        if (!rounds.next(p, tile, tc)) break;
        if (!tc.live || (p.dbg_skip & 2)) {  // dummy round: keep the pair's barrier protocol going, produce nothing
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait(&a_empty[kb], (it & 1) ^ 1, 0x540 | kb);
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[kb]);
            }
            ++it;
            continue;
        }
        // The rows of a thread (rg + RS i; RS = 32 or 64 is a multiple of every tT) share their frame index ti, so they
        // read the SAME enc row: one enc load + NR dec loads per k-block instead of 2 NR loads.  The registers that saves
        // pay for a second k-block of prefetch (two register sets used alternately: set A holds the even k-blocks, set B
        // the odd ones; a set is refilled with k-block kb + 2 right after k-block kb has been consumed), which matters
        // where the producers pace the kernel (few vocabulary tiles per cell tile) and trims their LSU instructions.
        bool ok[NR];
        const uint4* dp[NR];
        const int t_row = min(tc.t0 + ti[0], p.T - 1);
        const uint4* ep = reinterpret_cast<const uint4*>(enc + ((size_t)tc.b * p.T + t_row) * p.H) + c;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            ok[i] = tc.t0 + ti[i] < tc.Tb && tc.u0 + ui[i] < tc.Ub;
            // rows outside the utterance are masked below; clamp them so that the loads stay inside the tensors
            const int u = min(tc.u0 + ui[i], p.U - 1);
            dp[i] = reinterpret_cast<const uint4*>(dec + ((size_t)tc.b * p.U + u) * p.H) + c;
        }
        uint4 evA, dvA[NR], evB = make_uint4(0, 0, 0, 0), dvB[NR];
        evA = __ldg(ep);
#pragma unroll
        for (int i = 0; i < NR; ++i) { dvA[i] = __ldg(dp[i]); dvB[i] = make_uint4(0, 0, 0, 0); }
        if (KB > 1) {  // next k-block: 64 h further = 8 chunks of 16 bytes
            evB = __ldg(ep + 8);
#pragma unroll
            for (int i = 0; i < NR; ++i) dvB[i] = __ldg(dp[i] + 8);
        }
        auto produce_block = [&](const int kb, uint4& ev, uint4 (&dv)[NR]) {
            mbar_wait(&a_empty[kb], (it & 1) ^ 1, 0x500 | kb);
            uint8_t* blk = smem_a + kb * kABlockBytes;
            const uint32_t* e = reinterpret_cast<const uint32_t*>(&ev);
#pragma unroll
            for (int i = 0; i < NR; ++i) {  // one row at a time: four live output registers instead of sixteen
                const uint32_t* d = reinterpret_cast<const uint32_t*>(&dv[i]);
                uint4 o;
                uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const float2 x = fadd2(make_float2(bf16_lo(e[w]), bf16_hi(e[w])), make_float2(bf16_lo(d[w]), bf16_hi(d[w])));
                    float2 a;
                    if (ACT == ACT_LEAKY_RELU && slope_le1) {  // max(x, slope*x) == leaky_relu(x) for 0 <= slope <= 1
                        const float2 sx = fmul2(x, make_float2(p.act_param, p.act_param));
                        a = make_float2(fmaxf(x.x, sx.x), fmaxf(x.y, sx.y));
                    } else {
                        a = make_float2(act_t<ACT>(x.x, p.act_param), act_t<ACT>(x.y, p.act_param));
                    }
                    op[w] = ok[i] ? pack_bf16x2(a.x, a.y) : 0u;
                }
                const uint32_t off = sw128_offset((uint32_t)(rg + RS * i), (uint32_t)c);
                *reinterpret_cast<uint4*>(blk + off) = o;
                if (ModeTraits<MODE>::grad) {
                    uint8_t* img = reinterpret_cast<uint8_t*>(p.J_img) + ((size_t)(tile - p.tile_begin) * KB + kb) * kABlockBytes;
                    *reinterpret_cast<uint4*>(img + off) = o;
                }
            }
            if (kb + 2 < KB) {  // refill this register set with the k-block after next
                ev = __ldg(ep + (kb + 2) * 8);
#pragma unroll
                for (int i = 0; i < NR; ++i) dv[i] = __ldg(dp[i] + (kb + 2) * 8);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[kb]);
        };
#pragma unroll 1
        for (int kb = 0; kb < KB; kb += 2) {
            produce_block(kb, evA, dvA);
            if (kb + 1 < KB) produce_block(kb + 1, evB, dvB);
        }
        ++it;
    }
}

template <int MODE, bool PAIR, int NPW>
__global__ void __launch_bounds__(Roles<NPW>::kThreads, 1)
joint_gemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const JointParams p) {
    using R = Roles<NPW>;
    constexpr int kWarpTmaW = R::kWarpTmaW, kWarpRelayA = R::kWarpRelayA, kWarpMma = R::kWarpMma;
    constexpr int kFirstProducerWarp = R::kFirstProducerWarp, kNumProducerWarps = NPW, kNumEpilogueWarps = R::kEpilogueWarps;
    constexpr bool kTwoGroups = kNumEpilogueWarps == 8;  // two column groups of four epilogue warps, or one (NPW = 16: V <= 128)
    static_assert(!(NPW == 16) || PAIR, "the 16-producer form exists for CTA pairs only");
    extern __shared__ __align__(1024) uint8_t smem[];
    griddep_wait();  // programmatic dependent launch: nothing below may read global memory before the predecessor is done
    const SmemLayout L = smem_layout(p.KB, p.num_w_stages);
    uint8_t* smem_a = smem + L.a_off;
    uint8_t* smem_w = smem + L.w_off;
    float* bias_s = reinterpret_cast<float*>(smem + L.bias_off);
    float2* xchg = reinterpret_cast<float2*>(smem + L.xchg_off);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* w_full = bars;
    uint64_t* w_empty = bars + 4;
    uint64_t* a_full = bars + 12;
    uint64_t* a_empty = bars + 22;
    uint64_t* acc_full = bars + 32;
    uint64_t* acc_empty = bars + 34;
    uint64_t* a_done = bars + 36;
    uint64_t* acc_done = bars + 46;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L.tmem_off);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int KB = p.KB, NT = p.NT, NS = p.num_w_stages;
    const int n_live = count_live_tiles_warp(p);  // every warp counts for itself (a few hundred cycles, once)
    Rounds<PAIR> rounds(p, n_live);  // per-thread cursor over the live tiles; every role walks the same sequence
    const bool leader = rounds.rank == 0;
    constexpr uint32_t kCtas = PAIR ? 2 : 1;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B atoms need 1024-byte alignment
        // barriers the MMA lane waits on live in the leader CTA and collect arrivals from both CTAs
        for (int i = 0; i < NS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        // a_full / acc_empty (leader): its own 8 warps + one relayed arrival for the partner's 8 warps
        for (int i = 0; i < KB; ++i) { mbar_init(&a_full[i], kNumProducerWarps + (PAIR ? 1 : 0)); mbar_init(&a_empty[i], 1); mbar_init(&a_done[i], kNumProducerWarps); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kNumEpilogueWarps + (PAIR ? 1 : 0)); mbar_init(&acc_done[i], kNumEpilogueWarps); }
        fence_barrier_init();
    }
    if (warp_idx == kWarpTmaW && lane == 0) tma_prefetch_desc(&tmap_w);
    if (warp_idx == kWarpMma) {
        if (PAIR) tmem_alloc_2cta<512>(tmem_ptr);
        else tmem_alloc<512>(tmem_ptr);
    }
    tcgen05_fence_before();
    if (PAIR) cluster_sync_all();  // barrier inits and TMEM allocations of both CTAs are visible
    else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);  // warp-uniform for the compiler
    if (warp_idx == kWarpTmaW) {
        // ===================== W producer (TMA) =====================
        // PAIR: the leader fills BOTH halves of a stage (its own and, by multicast to CTA 1, its partner's),
        // so the refill latency is commit -> leader wake-up -> TMA, with no detour through the partner.
        if (PAIR && p.w_resident) {
            // one narrow vocabulary tile: every k-slice of W (both CTAs' halves, by multicast) once, on ONE barrier
            if (leader && rounds.count_my_rounds(n_live) > 0) {
                const uint32_t bytes = (uint32_t)p.w_stage_bytes;
                mbar_arrive_expect_tx_e(&w_full[0], 2u * (uint32_t)KB * bytes);
                for (int kb = 0; kb < KB; ++kb) {
                    uint8_t* dst = smem_w + (uint32_t)kb * bytes;
                    tma_load_2d_2cta_mcast_e(dst, &tmap_w, &w_full[0], 1, kb * kWStageKPair, 0);
                    tma_load_2d_2cta_mcast_e(dst, &tmap_w, &w_full[0], 2, kb * kWStageKPair, p.n_last >> 1);
                }
            }
        } else if (leader && !(p.dbg_skip & 4)) {  // whole warp, converged; one lane is elected inside each issuing instruction
            uint32_t stage = 0, phase = 0;
            const int my_rounds = rounds.count_my_rounds(n_live);
            for (int round = 0; round < my_rounds; ++round) {
                for (int nt = 0; nt < NT; ++nt) {
                    const int vt = nt;
                    if (PAIR) {
                        // CTA r holds rows [vt*256 + r*N/2, +128) of the vocabulary tile
                        const int n_cur = vt == NT - 1 ? p.n_last : kTileN;
                        const int row0 = vt * kTileN, row1 = row0 + (n_cur >> 1);
                        for (int kb = 0; kb < KB; ++kb) {
                            mbar_wait(&w_empty[stage], phase ^ 1, 0x100 | stage);
                            __syncwarp();
                            mbar_arrive_expect_tx_e(&w_full[stage], 2 * kWStageBytes);
                            uint8_t* dst = smem_w + stage * kWStageBytes;
                            tma_load_2d_2cta_mcast_e(dst, &tmap_w, &w_full[stage], 1, kb * kWStageKPair, row0);
                            tma_load_2d_2cta_mcast_e(dst, &tmap_w, &w_full[stage], 2, kb * kWStageKPair, row1);
                            if (++stage == (uint32_t)NS) { stage = 0; phase ^= 1; }
                        }
                    } else {
                        for (int ks = 0; ks < 2 * KB; ++ks) {
                            mbar_wait(&w_empty[stage], phase ^ 1, 0x100 | stage);
                            __syncwarp();
                            mbar_arrive_expect_tx_e(&w_full[stage], kWStageBytes);
                            tma_load_2d_e(smem_w + stage * kWStageBytes, &tmap_w, &w_full[stage], ks * kWStageK, vt * kTileN);
                            if (++stage == (uint32_t)NS) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp_idx == kWarpMma) {
        // ===================== MMA issuer (leader CTA only in PAIR mode) =====================
        // The single issuing lane is the serial resource of the kernel: everything that does not
        // depend on the stage is hoisted, descriptors are advanced with 32-bit adds on their low word
        // (start-address field, 16-byte units) and the loop body is wait -> MMAs -> commit.
        if (leader) {  // whole warp, converged
            uint32_t stage = 0, phase = 0, acc_it = 0, it = 0;
            const int M = PAIR ? 2 * kTileM : kTileM;
            const uint32_t idesc_full = make_idesc_bf16(M, kTileN, 0, 0);
            const uint32_t idesc_last = make_idesc_bf16(M, p.n_last, 0, 0);
            // A: K-major SWIZZLE_128B, 8-row atoms 1024 B apart.
            // W: K-major; single: SWIZZLE_64B (atoms 512 B apart), pair: SWIZZLE_128B (atoms 1024 B apart).
            const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(smem_a), 0, 1024);
            const uint64_t w_desc0 = PAIR ? make_smem_desc_sw128(smem_u32(smem_w), 0, 1024)
                                          : ((make_smem_desc_sw128(smem_u32(smem_w), 0, 512) & ~((uint64_t)7 << 61)) | ((uint64_t)4 << 61));
            const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), w_hi = (uint32_t)(w_desc0 >> 32);
            const uint32_t a_lo0 = (uint32_t)a_desc0, w_lo0 = (uint32_t)w_desc0;
            long long t_acc = 0, t_a = 0, t_w = 0, tm = 0, t_commit[3] = {0, 0, 0}, l_sum = 0, l_cnt = 0;
            const long long t_begin = clock64();
            const int my_rounds = rounds.count_my_rounds(n_live);
            const bool w_res = PAIR && p.w_resident;
            if (w_res && my_rounds > 0) {  // the resident W slices have landed (in both CTAs: one barrier collects all bytes)
                mbar_wait(&w_full[0], 0, 0x400);
                tcgen05_fence_after();
            }
            for (int round = 0; round < my_rounds; ++round) {
                for (int nt = 0; nt < NT; ++nt, ++acc_it) {
                    const uint32_t buf = acc_it & 1, acc_phase = (acc_it >> 1) & 1;
                    if (p.prof) tm = clock64();
                    mbar_wait(&acc_empty[buf], acc_phase ^ 1, 0x200 | buf);
                    if (p.prof) t_acc += clock64() - tm;
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * kTileN;
                    const uint32_t idesc = nt == NT - 1 ? idesc_last : idesc_full;
                    const bool first_nt = nt == 0, last_nt = nt == NT - 1;
                    uint32_t a_lo = a_lo0;
                    if (PAIR && NS == 4 && !(KB & 1) && !(p.dbg_skip & 256) && !w_res) {
                        // Two k-blocks (8 MMAs) per wait / commit group: measured 5 % faster than one group per k-block (the
                        // per-stage wait -> 4 MMAs -> commit structure, not the W bytes, is what the W stream costs; bit 256 of
                        // TSASR_DEBUG_SKIP selects the per-stage loop for A/B runs)
                        for (int kb = 0; kb < KB; kb += 2, a_lo += 2 * (kABlockBytes >> 4)) {
                            if (first_nt) {
                                mbar_wait(&a_full[kb], it & 1, 0x300 | kb);
                                mbar_wait(&a_full[kb + 1], it & 1, 0x300 | (kb + 1));
                            }
                            mbar_wait(&w_full[stage], phase, 0x400 | stage);
                            mbar_wait(&w_full[stage + 1], phase, 0x400 | (stage + 1));
                            tcgen05_fence_after();
                            const uint32_t w_lo = w_lo0 + stage * (kWStageBytes >> 4);
                            umma_bf16_2cta_x4_e(d_tmem, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)w_hi << 32) | w_lo, idesc, kb != 0);
                            umma_bf16_2cta_x4_e(d_tmem, ((uint64_t)a_hi << 32) | (a_lo + (kABlockBytes >> 4)),
                                                ((uint64_t)w_hi << 32) | (w_lo + (kWStageBytes >> 4)), idesc, true);
                            umma_commit_2cta_e(&w_empty[stage], 1);
                            umma_commit_2cta_e(&w_empty[stage + 1], 1);
                            stage += 2;
                            if (stage == 4u) { stage = 0; phase ^= 1; }
                            if (last_nt) { umma_commit_2cta_e(&a_empty[kb], 3); umma_commit_2cta_e(&a_empty[kb + 1], 3); }
                        }
                    } else
                    for (int kb = 0; kb < KB; ++kb, a_lo += kABlockBytes >> 4) {
                        if (first_nt) {
                            if (p.prof) tm = clock64();
                            mbar_wait(&a_full[kb], it & 1, 0x300 | kb);
                            if (p.prof) t_a += clock64() - tm;
                            tcgen05_fence_after();
                        }
                        if (PAIR && w_res) {  // W slice kb sits at its fixed place: nothing to wait for, nothing to release
                            umma_bf16_2cta_x4_e(d_tmem, ((uint64_t)a_hi << 32) | a_lo,
                                                ((uint64_t)w_hi << 32) | (w_lo0 + (uint32_t)kb * ((uint32_t)p.w_stage_bytes >> 4)), idesc, kb != 0);
                            umma_commit_2cta_e(&a_empty[kb], 3);  // NT == 1: every vocabulary tile is the last one
                        } else if (PAIR) {
                            if (p.prof) tm = clock64();
                            if (!(p.dbg_skip & 4)) mbar_wait(&w_full[stage], phase, 0x400 | stage);
                            if (p.prof) t_w += clock64() - tm;
                            tcgen05_fence_after();
                            const uint32_t w_lo = w_lo0 + stage * (kWStageBytes >> 4);
                            if (p.prof) tm = clock64();
                            // four K = 16 MMAs: 32 bytes per step along the 128-byte swizzled rows
                            if (!(p.dbg_skip & 32)) umma_bf16_2cta_x4_e(d_tmem, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)w_hi << 32) | w_lo, idesc, kb != 0);
                            if (p.prof) { l_sum += clock64() - tm; ++l_cnt; }
                            if (p.prof) tm = clock64();
                            if (!(p.dbg_skip & 4)) umma_commit_2cta_e(&w_empty[stage], 1);  // only the leader refills
                            if (++stage == (uint32_t)NS) { stage = 0; phase ^= 1; }
                            if (last_nt) umma_commit_2cta_e(&a_empty[kb], 3);
                            if (p.prof) t_commit[0] += clock64() - tm;
                        } else {
#pragma unroll
                            for (int half = 0; half < 2; ++half) {
                                if (p.prof) tm = clock64();
                                mbar_wait(&w_full[stage], phase, 0x400 | stage);
                                if (p.prof) t_w += clock64() - tm;
                                tcgen05_fence_after();
                                const uint32_t w_lo = w_lo0 + stage * (kWStageBytes >> 4);
                                const uint32_t al = a_lo + half * 4;  // 64 bytes into the 128-byte row
                                umma_bf16_x2_e(d_tmem, ((uint64_t)a_hi << 32) | al, ((uint64_t)w_hi << 32) | w_lo, idesc, (kb | half) != 0);
                                umma_commit_e(&w_empty[stage]);
                                if (++stage == (uint32_t)NS) { stage = 0; phase ^= 1; }
                            }
                            if (last_nt) umma_commit_e(&a_empty[kb]);
                        }
                    }
                    if (PAIR) umma_commit_2cta_e(&acc_full[buf], 3);
                    else umma_commit_e(&acc_full[buf]);
                }
                ++it;
            }
            if (p.prof && lane == 0) {
                long long* o = p.prof + blockIdx.x * 16;
                o[0] = clock64() - t_begin; o[1] = t_acc; o[2] = t_a; o[3] = t_w; o[4] = it; o[5] = l_sum; o[6] = l_cnt; o[7] = t_commit[0];
            }
        } else if (PAIR) {
            // partner CTA: relay "all 8 epilogue warps released accumulator buffer b" to the leader's acc_empty
            const uint32_t acc_empty_leader = mapa_u32(smem_u32(&acc_empty[0]), 0);
            uint32_t acc_it = 0;
            const int my_rounds = rounds.count_my_rounds(n_live);
            for (int round = 0; round < my_rounds; ++round) {
                for (int nt = 0; nt < NT; ++nt, ++acc_it) {
                    const uint32_t buf = acc_it & 1;
                    mbar_wait(&acc_done[buf], (acc_it >> 1) & 1, 0x280 | buf);
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(acc_empty_leader + buf * 8);
                }
            }
        }
    } else if (PAIR && warp_idx == kWarpRelayA) {
        // partner CTA: relay "all 8 producer warps wrote A block kb" to the leader's a_full
        if (!leader) {
            const uint32_t a_full_leader = mapa_u32(smem_u32(&a_full[0]), 0);
            uint32_t it = 0;
            const int my_rounds = rounds.count_my_rounds(n_live);
            for (int round = 0; round < my_rounds; ++round) {
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&a_done[kb], it & 1, 0x380 | kb);
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(a_full_leader + kb * 8);
                }
                ++it;
            }
        }
    } else if (warp_idx >= kFirstProducerWarp && warp_idx < kFirstProducerWarp + kNumProducerWarps) {
        // ===================== A producers: J = bf16(act(enc + dec)) =====================
        uint64_t* a_arrive = (PAIR && !leader) ? a_done : a_full;
        switch (p.act_kind) {
            case ACT_LEAKY_RELU: produce_a<MODE, ACT_LEAKY_RELU, PAIR, NPW>(p, smem_a, p.enc, p.dec, a_arrive, a_empty); break;
            case ACT_RELU: produce_a<MODE, ACT_RELU, PAIR, NPW>(p, smem_a, p.enc, p.dec, a_arrive, a_empty); break;
            case ACT_TANH: produce_a<MODE, ACT_TANH, PAIR, NPW>(p, smem_a, p.enc, p.dec, a_arrive, a_empty); break;
            default: produce_a<MODE, ACT_IDENTITY, PAIR, NPW>(p, smem_a, p.enc, p.dec, a_arrive, a_empty); break;
        }
    } else if (warp_idx < kNumEpilogueWarps) {
        // ===================== epilogue =====================
        const int grp = kTwoGroups ? warp_idx >> 2 : 0;        // column half of every vocabulary tile owned by this group
        const int q = warp_idx & 3;                            // TMEM lane quarter owned by this warp
        const int row = q * 32 + lane;                         // tile row == TMEM lane
        const int gtid = threadIdx.x & 127;                    // 0..127 inside the group
        const int tTm = (1 << p.tT_log2) - 1;
        const int n_lab = max(1, 32 >> p.tT_log2);  // distinct label positions inside one warp
        const uint32_t tmem_row0 = tmem_base + ((uint32_t)(q * 32) << 16);
        float* bias_g = bias_s + grp * (kTileN / 2);  // this group's 128 staged bias values
        constexpr int kGrpCols = kTileN / 2;
        uint32_t acc_it = 0, it = 0;
#ifdef TSASR_EPI_PROF
        long long pe_wait = 0, pe_proc = 0, pe_bar = 0, pe_t = 0;  // development build: cycle split of this warp
#endif
        float nb0 = 0.f;  // prefetched bias value of the next vocabulary tile
        int nb_nt = -1;
        // the partner CTA releases accumulators on a local barrier that its relay warp forwards to the leader
        uint64_t* acc_release = (PAIR && !leader) ? acc_done : acc_empty;
        for (;;) {
            int tile;
            TileCoord tc;
            if (!rounds.next(p, tile, tc)) break;
            if (!tc.live || (p.dbg_skip & 1)) {  // dummy round: release the accumulators the pair's MMA stream wrote for us
                for (int nt = 0; nt < NT; ++nt, ++acc_it) {
                    const uint32_t buf = acc_it & 1;
                    mbar_wait(&acc_full[buf], (acc_it >> 1) & 1, 0x640 | buf);
                    tcgen05_fence_after();
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_release[buf]);
                }
                if (MODE == MODE_FWD && kTwoGroups) {  // the merge of a cell tile passes four barriers
#pragma unroll
                    for (int i = 0; i < 4; ++i) asm volatile("bar.sync 3, 256;" ::: "memory");
                }
                ++it;
                continue;
            }
            const int ti = row & tTm, ui = row >> p.tT_log2;
            const int t = tc.t0 + ti, u = tc.u0 + ui;
            const bool valid = t < tc.Tb && u < tc.Ub;
            const int label = (valid && u < tc.Ub - 1 && p.targets) ? p.targets[(size_t)tc.b * (p.U - 1) + u] : -1;
            const size_t cell_o = valid ? skew_index(tc.b, t, u, p.T, p.U) : 0;
            // labels of the (at most four) label positions covered by this warp: warp-uniform values
            int lab_w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) lab_w[k] = __shfl_sync(0xffffffffu, label, (k << p.tT_log2) & 31);

            // per-row state
            float run_m = -INFINITY, run_s = 0.f, y_blank = -INFINITY, y_label = -INFINITY;  // MODE_FWD (log2 domain)
            float nz2 = 0.f, occ = 0.f, ob = 0.f, oe = 0.f, p_blank = 0.f, p_label = 0.f;    // MODE_GRAD
            float dy_post = 1.f;  // MODE_GRAD_CLAMP: upstream factor applied AFTER the clamp (occ / ob / oe are unit-cost terms then)
            if (ModeTraits<MODE>::grad && valid) {
                const float Lp = -p.cost[tc.b];
                float dy = p.dcost ? p.dcost[tc.b] : 1.f;
                if (ModeTraits<MODE>::clamp) { dy_post = dy; dy = 1.f; }
                const float a = p.alpha[cell_o];
                const float2 lp = p.lat2_in[cell_o];
                occ = dy * __expf(a + p.beta[cell_o] - Lp);
                float bt1 = -INFINITY;
                if (t < tc.Tb - 1) bt1 = p.beta[cell_o + p.U];
                else if (u == tc.Ub - 1) bt1 = 0.f;
                ob = bt1 == -INFINITY ? 0.f : dy * __expf(a + lp.x + bt1 - Lp);
                oe = u < tc.Ub - 1 ? dy * __expf(a + lp.y + p.beta[cell_o + p.U + 1] - Lp) : 0.f;
                nz2 = -p.logz_in[cell_o] * kLog2eF;
                p_blank = __expf(lp.x);
                p_label = u < tc.Ub - 1 ? __expf(lp.y) : 0.f;
            }

            for (int nt = 0; nt < NT; ++nt, ++acc_it) {
                // both groups work on every accumulator: group g takes columns [g*128, g*128+128), which halves
                // the time a TMEM buffer stays busy after its last MMA
                const uint32_t buf = acc_it & 1, acc_phase = (acc_it >> 1) & 1;
                const uint32_t tmem_row = tmem_row0 + buf * kTileN;
                // stage bias * log2(e) for this vocabulary tile (one buffer per group).  The values were
                // loaded into registers while the group's previous tile was being processed, so the L2
                // latency of the load is off the critical path.
                const int vt = nt;
#ifdef TSASR_EPI_PROF
                if (p.prof) pe_t = clock64();
#endif
                {
                    if (nb_nt != vt) {  // first tile of the kernel
                        const int v0 = vt * kTileN + grp * kGrpCols + gtid;
                        nb0 = v0 < p.V ? __ldg(p.bias + v0) : 0.f;
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(4 + grp) : "memory");  // previous tile's reads are done
                    bias_g[gtid] = nb0 * kLog2eF;
                    asm volatile("bar.sync %0, 128;" ::"r"(4 + grp) : "memory");
                    nb_nt = nt + 1 < NT ? nt + 1 : 0;  // next vocabulary tile
                    const int v0 = nb_nt * kTileN + grp * kGrpCols + gtid;
                    nb0 = v0 < p.V ? __ldg(p.bias + v0) : 0.f;
                }
#ifdef TSASR_EPI_PROF
                if (p.prof) { const long long c = clock64(); pe_bar += c - pe_t; pe_t = c; }
#endif
                mbar_wait(&acc_full[buf], acc_phase, 0x600 | buf);
#ifdef TSASR_EPI_PROF
                if (p.prof) { const long long c = clock64(); pe_wait += c - pe_t; pe_t = c; }
#endif
                tcgen05_fence_after();
                const int n_all = vt == NT - 1 ? p.n_last : kTileN;
                const int c_begin = grp * kGrpCols, n_cols = min(n_all, c_begin + kGrpCols);  // this group's columns [c_begin, n_cols)
                // chunks (16 columns each) of this group's half tile that need the rare path of process(): bit i <-> cc = c_begin + 16 i
                uint32_t special_mask = 0;
                uint8_t* img_grp = nullptr;  // MODE_GRAD: first dY image of this group's half tile
                if (ModeTraits<MODE>::grad)
                    img_grp = reinterpret_cast<uint8_t*>(p.dY_img) + ((size_t)(tile - p.tile_begin) * (NT * 4) + vt * 4 + grp * 2) * kABlockBytes;
                if (MODE == MODE_FWD || ModeTraits<MODE>::grad) {
                    const int g0 = vt * kTileN + c_begin;  // first vocabulary column of this group's half tile
                    if ((unsigned)(p.blank - g0) < (unsigned)kGrpCols) special_mask |= 1u << ((p.blank - g0) >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < n_lab && (unsigned)(lab_w[k] - g0) < (unsigned)kGrpCols) special_mask |= 1u << ((lab_w[k] - g0) >> 4);
                    if (g0 + kGrpCols > p.V) {  // the vocabulary ends inside (or before) this half tile
                        const int first = max(0, (p.V - g0) >> 4);  // first chunk with col0 + 16 > V
                        special_mask |= first < 8 ? (0xffu << first) & 0xffu : 0u;
                    }
                }
                // TMEM loads are software-pipelined: chunk c+1 is in flight while chunk c is processed
                uint32_t raw0[16], raw1[16];
                auto process = [&](const uint32_t (&raw)[16], const int ci) {  // ci: chunk of 16 columns inside the group's half tile (0..7)
                    const int cc = c_begin + 16 * ci;
                    const int col0 = vt * kTileN + cc;
                    const bool special = (special_mask >> ci) & 1u;
                    // y = logit * log2(e) = acc * log2(e) + bias * log2(e), two columns per instruction
                    float2 y2[8];
                    const float4* b4 = reinterpret_cast<const float4*>(bias_g + 16 * ci);
                    const float2 l2e = make_float2(kLog2eF, kLog2eF);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const float4 bb = b4[jj];
                        y2[2 * jj] = ffma2(make_float2(__uint_as_float(raw[4 * jj]), __uint_as_float(raw[4 * jj + 1])), l2e,
                                           make_float2(bb.x, bb.y));
                        y2[2 * jj + 1] = ffma2(make_float2(__uint_as_float(raw[4 * jj + 2]), __uint_as_float(raw[4 * jj + 3])), l2e,
                                               make_float2(bb.z, bb.w));
                    }
                    float* y = reinterpret_cast<float*>(y2);
                    if (MODE == MODE_FWD) {
                        // Rare work first, behind ONE warp-uniform branch per chunk (special: this chunk holds the blank
                        // column, one of the warp's label columns, or the end of the vocabulary).  Measured: with the
                        // range checks inline (5 not-taken branches per chunk) the epilogue needed 8.5k cycles per
                        // vocabulary tile against 5.1k for the MMAs and paced the whole kernel.
                        if (special) {
                            // blank / label logits: the column index is warp-uniform per label position
                            if (p.blank >= col0 && p.blank < col0 + 16) {
                                const int idx = p.blank - col0;
#pragma unroll
                                for (int jj = 0; jj < 16; ++jj)
                                    if (jj == idx) y_blank = y[jj];
                            }
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (k < n_lab && lab_w[k] >= col0 && lab_w[k] < col0 + 16) {
                                    const int idx = lab_w[k] - col0;
                                    float sel = 0.f;
#pragma unroll
                                    for (int jj = 0; jj < 16; ++jj)
                                        if (jj == idx) sel = y[jj];
                                    if ((lane >> p.tT_log2) == k || n_lab == 1) y_label = sel;
                                }
                            }
                            if (col0 + 16 > p.V) {
#pragma unroll
                                for (int jj = 0; jj < 16; ++jj)
                                    if (col0 + jj >= p.V) y[jj] = -INFINITY;
                            }
                        }
                        // pairwise trees keep the dependency chains short
                        float m8[8], m4[4];
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) m8[jj] = fmaxf(y[2 * jj], y[2 * jj + 1]);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) m4[jj] = fmaxf(m8[2 * jj], m8[2 * jj + 1]);
                        const float cm = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                        const float mn = fmaxf(run_m, cm);
                        const float2 nm2 = make_float2(-mn, -mn);
                        float2 e2[8];
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            const float2 d = fadd2(y2[jj], nm2);
                            e2[jj] = make_float2(ex2_approx(d.x), ex2_approx(d.y));
                        }
                        const float2 s4a = fadd2(fadd2(e2[0], e2[1]), fadd2(e2[2], e2[3]));
                        const float2 s4b = fadd2(fadd2(e2[4], e2[5]), fadd2(e2[6], e2[7]));
                        const float2 s2 = fadd2(s4a, s4b);
                        run_s = fmaf(run_s, ex2_approx(run_m - mn), s2.x + s2.y);
                        run_m = mn;
                    } else if (ModeTraits<MODE>::grad) {
                        // dlogits = occ * softmax - [blank] ob - [label] oe, emitted as bf16 into the
                        // [128 x 64] SWIZZLE_128B image of this (tile, 64-column block)
                        uint32_t packed[8];
                        const float2 nz = make_float2(nz2, nz2), oc = make_float2(occ, occ);
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            const float2 d = fadd2(y2[jj], nz);
                            float2 g = fmul2(make_float2(ex2_approx(d.x), ex2_approx(d.y)), oc);
                            if (ModeTraits<MODE>::clamp) {  // occ * p >= 0: only the upper bound can bite; then the upstream factor
                                g = fmul2(make_float2(fminf(g.x, p.clamp), fminf(g.y, p.clamp)), make_float2(dy_post, dy_post));
                            }
                            packed[jj] = pack_bf16x2(g.x, g.y);
                        }
                        if (special && col0 + 16 > p.V) {  // end of the vocabulary: padded columns are exact zeros
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) {
                                if (col0 + 2 * jj >= p.V) packed[jj] &= 0xffff0000u;
                                if (col0 + 2 * jj + 1 >= p.V) packed[jj] &= 0x0000ffffu;
                            }
                        }
                        // image of this (tile, 64-column block): vb = col0 >> 6 = 4 vt + 2 grp + (ci >> 2); the two
                        // 16-byte chunks of these 16 columns share one 32-byte sector of the swizzled row (their chunk
                        // index (ci & 3) * 2 is even, the XOR with row & 7 only swaps them for odd rows): ONE 256-bit store
                        uint8_t* img = img_grp + (size_t)(ci >> 2) * kABlockBytes;
                        {
                            const bool swap = row & 1;
                            uint8_t* dst = img + (uint32_t)row * 128u + ((((uint32_t)(ci & 3)) ^ (((uint32_t)row & 7u) >> 1)) << 5);
                            st_global_v8(dst, swap ? packed[4] : packed[0], swap ? packed[5] : packed[1], swap ? packed[6] : packed[2],
                                         swap ? packed[7] : packed[3], swap ? packed[0] : packed[4], swap ? packed[1] : packed[5],
                                         swap ? packed[2] : packed[6], swap ? packed[3] : packed[7]);
                        }
                        // patch the two special columns from the saved lattice (same thread, program order)
                        if (special && valid) {
                            __nv_bfloat16* rowp = reinterpret_cast<__nv_bfloat16*>(img);
                            if (p.blank >= col0 && p.blank < col0 + 16) {
                                float g = occ * p_blank - ob;
                                if (label == p.blank) g -= oe;
                                if (ModeTraits<MODE>::clamp) g = fminf(fmaxf(g, -p.clamp), p.clamp) * dy_post;
                                const int cv = p.blank & 63;
                                rowp[(sw128_offset((uint32_t)row, (uint32_t)(cv >> 3)) >> 1) + (cv & 7)] = __float2bfloat16_rn(g);
                            }
                            if (label >= col0 && label < col0 + 16 && label != p.blank) {
                                float g = occ * p_label - oe;
                                if (ModeTraits<MODE>::clamp) g = fminf(fmaxf(g, -p.clamp), p.clamp) * dy_post;
                                const int cv = label & 63;
                                rowp[(sw128_offset((uint32_t)row, (uint32_t)(cv >> 3)) >> 1) + (cv & 7)] = __float2bfloat16_rn(g);
                            }
                        }
                    } else {  // MODE_DEBUG
                        if (valid) {
                            float* out = p.dbg_logits + (((size_t)tc.b * p.T + t) * p.U + u) * p.V;
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj)
                                if (col0 + jj < p.V) out[col0 + jj] = y[jj] * kLn2F;
                        }
                    }
                };
                const bool abl_nomath = p.dbg_skip & 8, abl_nold = p.dbg_skip & 16;  // development ablations
                if (abl_nold) {
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) { raw0[jj] = 0x3f800000u + jj; raw1[jj] = 0x3f900000u + jj; }
                }
                if (c_begin < n_cols && !abl_nold) tmem_ld_32x32b_x16(tmem_row + c_begin, raw0);
                for (int ci = 0; ci < kGrpCols / 16; ci += 2) {  // two chunks of 16 columns per iteration (code size: the
                    const int cc = c_begin + 16 * ci;            // unrolled form was slower, instruction fetch)
                    if (cc >= n_cols) break;
                    if (!abl_nold) tmem_ld_wait();
                    const bool more1 = cc + 16 < n_cols;
                    if (more1 && !abl_nold) tmem_ld_32x32b_x16(tmem_row + cc + 16, raw1);
                    if (!abl_nomath) process(raw0, ci);
                    if (more1) {
                        if (!abl_nold) tmem_ld_wait();
                        if (cc + 32 < n_cols && !abl_nold) tmem_ld_32x32b_x16(tmem_row + cc + 32, raw0);
                        if (!abl_nomath) process(raw1, ci + 1);
                    }
                }
                if (ModeTraits<MODE>::grad && n_all < kTileN) {
                    // zero-fill the 16-column chunks of the last tile the MMA did not produce, so the
                    // backward GEMMs read finite zeros for the padded vocabulary columns
                    const int cols_pad = min(((p.V + 63) / 64) * 64 - vt * kTileN, c_begin + kGrpCols);  // columns the images cover
                    for (int cc = max(n_all, c_begin); cc < cols_pad; cc += 16) {
                        const int col0 = vt * kTileN + cc;
                        uint8_t* img = reinterpret_cast<uint8_t*>(p.dY_img) +
                                       ((size_t)(tile - p.tile_begin) * (NT * 4) + (col0 >> 6)) * kABlockBytes;
                        const int chunk0 = (col0 & 63) >> 3;
#pragma unroll
                        for (int c2 = 0; c2 < 2; ++c2)
                            *reinterpret_cast<uint4*>(img + sw128_offset((uint32_t)row, (uint32_t)(chunk0 + c2))) =
                                make_uint4(0, 0, 0, 0);
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_release[buf]);
#ifdef TSASR_EPI_PROF
                if (p.prof) pe_proc += clock64() - pe_t;
#endif
            }
            if (MODE == MODE_FWD) {
                // merge the two groups' running (max, sum) and picked logits; group 0 writes the lattice.  The exchange
                // buffer is one float2 per row and is used twice ((max, sum), then the picked logits): the shared memory
                // saved buys the fourth W stage at H = 640.  Barriers: buffer free / (max, sum) written / read / picks written.
                float2 o_ms = make_float2(-INFINITY, 0.f), o_y = make_float2(-INFINITY, -INFINITY);
                if (kTwoGroups) {
                    asm volatile("bar.sync 3, 256;" ::: "memory");
                    if (grp == 1) xchg[row] = make_float2(run_m, run_s);
                    asm volatile("bar.sync 3, 256;" ::: "memory");
                    if (grp == 0) o_ms = xchg[row];
                    asm volatile("bar.sync 3, 256;" ::: "memory");
                    if (grp == 1) xchg[row] = make_float2(y_blank, y_label);
                    asm volatile("bar.sync 3, 256;" ::: "memory");
                    if (grp == 0 && valid) o_y = xchg[row];
                }
                if (grp == 0 && valid) {
                    const float mn = fmaxf(run_m, o_ms.x);
                    const float s = run_s * ex2_approx(run_m - mn) + o_ms.y * ex2_approx(o_ms.x - mn);
                    const float lz2 = mn + lg2_approx(s);
                    const float yb = fmaxf(y_blank, o_y.x), yl = fmaxf(y_label, o_y.y);
                    const float lpb = (yb - lz2) * kLn2F;
                    const float lpe = label >= 0 ? (yl - lz2) * kLn2F : -INFINITY;
                    p.lat2[cell_o] = make_float2(lpb, lpe);
                    p.logz[cell_o] = lz2 * kLn2F;
                }
            }
            ++it;
        }
#ifdef TSASR_EPI_PROF
        if (p.prof && (warp_idx & 3) == 0 && lane == 0) {
            long long* o = p.prof + blockIdx.x * 16 + 8 + grp * 4;
            o[0] = pe_wait; o[1] = pe_proc; o[2] = pe_bar; o[3] = acc_it;
        }
#endif
    }

    tcgen05_fence_before();
    if (PAIR) cluster_sync_all();  // no CTA may exit while its partner still signals its barriers / reads its smem
    else __syncthreads();
    if (warp_idx == kWarpMma) {
        tcgen05_fence_after();
        if (PAIR) tmem_dealloc_2cta<512>(tmem_base);
        else tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace tsasr
