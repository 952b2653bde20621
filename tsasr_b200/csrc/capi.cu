// capi.cu -- extern "C" entry points of libtsasr_b200.so (see include/tsasr_b200.h).
//
// Plain pointers and sizes only; no torch types cross this boundary.  Every function validates its
// arguments, launches on the caller's stream and returns a status code; nothing allocates, nothing
// synchronises, nothing throws.
#include "../../include/tsasr_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges cost nothing unless a profiler is attached

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "backward_gemm.cuh"
#include "common.cuh"
#include "joint_gemm.cuh"
#include "linear.cuh"
#include "predictor.cuh"

namespace tsasr {
// lattice.cu
cudaError_t launch_logits_to_lattice(const void*, int, const int*, const int*, const int*, int, int, int, int, int, int,
                                     float2*, float*, cudaStream_t);
cudaError_t launch_alpha_beta(const float2*, const int*, const int*, int, int, int, float*, float*, float*, float*,
                              float*, cudaStream_t);
cudaError_t launch_logits_grad(const void*, int, const int*, const int*, const int*, int, int, int, int, int,
                               const float2*, const float*, const float*, const float*, const float*, const float*,
                               float, void*, cudaStream_t);
cudaError_t launch_logprobs_grad(const int*, const int*, const int*, int, int, int, int, int, const float2*,
                                 const float*, const float*, const float*, const float*, float*, cudaStream_t);
// prep.cu
cudaError_t launch_lengths(const float*, const float*, const int*, const int*, int, int, int, int*, int*, int*, cudaStream_t);
cudaError_t launch_cast3(const float*, size_t, const float*, size_t, const float*, size_t, void*, void*, void*, int, cudaStream_t);
cudaError_t launch_prepare_inputs(const float*, size_t, const float*, size_t, const float*, size_t, void*, void*, void*, const long long*,
                                  size_t, int*, const float*, const float*, const int*, const int*, int, int, int, int*, int*, int*, int*,
                                  int, int, cudaStream_t);
// decode.cu
cudaError_t launch_joint_decode_step(const float*, const float*, long long, long long, const float*, const float*, int, int, int,
                                     int, float, float*, void*, cudaStream_t);
size_t joint_decode_workspace_bytes(int V);
}  // namespace tsasr

using namespace tsasr;

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
    return fail(TSASR_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

// ---- optional per-kernel timing (measurement aid, see tsasr_kernel_timing_enable) ----
// Every launch site is bracketed by two CUDA events on the launching stream; nothing synchronises until
// tsasr_kernel_timings() is called.  Off by default (one relaxed atomic load per launch site); while on, the record is
// guarded by a mutex so that calls from several host threads (one per device) stay consistent.
struct TimedLaunch { const char* name; cudaEvent_t e0, e1; };
static std::atomic<bool> g_timing_on{false};
static std::mutex g_timed_mu;
static TimedLaunch g_timed[4096];
static int g_n_timed = 0;
struct ScopedTiming {
    cudaEvent_t e1 = nullptr;
    cudaStream_t st;
    ScopedTiming(const char* name, cudaStream_t s) : st(s) {
        if (!g_timing_on.load(std::memory_order_relaxed)) return;
        cudaEvent_t e0 = nullptr;
        {
            std::lock_guard<std::mutex> lock(g_timed_mu);
            if (g_n_timed >= 4096) return;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            g_timed[g_n_timed++] = TimedLaunch{name, e0, e1};
        }
        cudaEventRecord(e0, st);
    }
    ~ScopedTiming() {
        if (e1) cudaEventRecord(e1, st);
    }
};

// NVTX range over one C-ABI call (fwd / DP / bwd show up as named ranges in nsys / ncu timelines)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

#define REQUIRE(cond, ...) \
    do {                   \
        if (!(cond)) return fail(TSASR_E_INVALID, __VA_ARGS__); \
    } while (0)

// widest lattice the DP implements: up to 1024 columns one thread per column (alpha_beta_kernel), beyond that several
// columns per thread with the diagonal double-buffered in shared memory (alpha_beta_wide_kernel: 2 * U floats)
static constexpr int kMaxLatticeWidth = 8192;

static int check_dims(int B, int T, int U, int V, int blank) {
    REQUIRE(B >= 1 && T >= 1 && U >= 1 && V >= 1, "B, T, U, V must be >= 1 (got %d %d %d %d)", B, T, U, V);
    REQUIRE(blank >= 0 && blank < V, "blank must be within [0, V) (got %d, V=%d)", blank, V);
    REQUIRE((long long)B * (T + U - 1) * U < (1ll << 31), "lattice too large for 32-bit indexing");
    return TSASR_OK;
}

// ---- TMA descriptor encode through the driver entry point (no link-time libcuda dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D bf16 row-major tensor [rows, cols] -> tensor map with box {box_cols, box_rows}
static int make_tmap_2d_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_cols,
                             uint32_t box_rows, CUtensorMapSwizzle swz) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(TSASR_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(TSASR_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return TSASR_OK;
}

// SM count and opt-in shared memory of the CURRENT device, cached per device ordinal (one process may drive several
// GPUs: SpeechBrain's data_parallel_backend, multi-device tests).  The cache is written once per device with the same
// values by whoever gets there first, so concurrent callers need no lock.
static int device_info(int* num_sms, int* max_smem) {
    static constexpr int kMaxDevices = 64;
    static std::atomic<int> sms_cache[kMaxDevices];   // zero-initialised; 0 = not queried yet
    static std::atomic<int> smem_cache[kMaxDevices];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    const bool cacheable = dev >= 0 && dev < kMaxDevices;
    int sms = cacheable ? sms_cache[dev].load(std::memory_order_acquire) : 0, smem = 0;
    if (sms > 0) {
        smem = smem_cache[dev].load(std::memory_order_relaxed);
    } else {
        int major = 0;
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        if (major != 10)
            return fail(TSASR_E_UNSUPPORTED, "tsasr_b200 kernels are built for sm_100a only (device %d is sm_%d0)", dev, major);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (sms <= 0) return fail(TSASR_E_CUDA, "cudaDeviceGetAttribute(MultiProcessorCount) returned %d", sms);
        if (cacheable) {
            smem_cache[dev].store(smem, std::memory_order_relaxed);
            sms_cache[dev].store(sms, std::memory_order_release);
        }
    }
    *num_sms = sms;
    *max_smem = smem;
    return TSASR_OK;
}

// ---- tile-shape selection: (tT, tU) with tT * tU = 128, tT in {8,16,32}, minimising padded cells ----
static void choose_tile(int T, int U, int* tT_log2) {
    long long best = -1;
    int best_l = 4;
    for (int l = 3; l <= 5; ++l) {
        const int tT = 1 << l, tU = 128 >> l;
        const long long padded = (long long)((T + tT - 1) / tT) * tT * (long long)((U + tU - 1) / tU) * tU;
        // prefer tT = 16 on ties (d_enc partial sums stay in registers across label tiles)
        if (best < 0 || padded < best || (padded == best && l == 4)) { best = padded; best_l = l; }
    }
    *tT_log2 = best_l;
}

// CTA pairs (cta_group::2) are the default; TSASR_DEBUG_NO_PAIR=1 selects the single-CTA kernel (A/B runs).
static bool use_pair() {
    static const bool v = getenv("TSASR_DEBUG_NO_PAIR") == nullptr;
    return v;
}

static int fill_joint_params(JointParams& p, const void* enc, const void* dec, const float* bias,
                             const int32_t* targets, const int32_t* ll, const int32_t* tl, int B, int T, int U, int H,
                             int V, int blank, int act_kind, float act_param, int max_smem) {
    REQUIRE(H % 64 == 0 && H >= 64 && H <= 64 * kMaxKB, "fused joint needs H %% 64 == 0 and 64 <= H <= 640 (got %d)", H);
    REQUIRE(V >= 2, "fused joint needs V >= 2");
    REQUIRE(act_kind >= 0 && act_kind <= 3, "unknown activation code %d", act_kind);
    REQUIRE((reinterpret_cast<uintptr_t>(enc) & 15) == 0 && (reinterpret_cast<uintptr_t>(dec) & 15) == 0,
            "enc/dec must be 16-byte aligned");
    memset(&p, 0, sizeof(p));
    p.bias = bias;
    p.targets = targets;
    p.logit_lengths = ll;
    p.target_lengths = tl;
    p.B = B; p.T = T; p.U = U; p.H = H; p.V = V; p.blank = blank;
    p.act_kind = act_kind;
    p.act_param = act_param;
    choose_tile(T, U, &p.tT_log2);
    const int tT = 1 << p.tT_log2, tU = 128 >> p.tT_log2;
    p.nTt = (T + tT - 1) / tT;
    p.nTu = (U + tU - 1) / tU;
    p.tile_begin = 0;
    p.tile_end = B * p.nTt * p.nTu;
    p.KB = H / 64;
    p.NT = (V + kTileN - 1) / kTileN;
    // UMMA N of the last vocabulary tile: a multiple of 16 per CTA (pairs split N between the two CTAs)
    const int n_gran = use_pair() ? 32 : 16;
    p.n_last = ((V - (p.NT - 1) * kTileN) + n_gran - 1) / n_gran * n_gran;
    int ns = kMaxWStages;
    if (const char* env = getenv("TSASR_DEBUG_W_STAGES")) {  // development knob (pipeline-depth experiments)
        const int v = atoi(env);
        if (v >= 2 && v <= kMaxWStages) ns = v;
    }
    while (ns > 2 && (int)smem_layout(p.KB, ns).total > max_smem) --ns;
    if ((int)smem_layout(p.KB, ns).total > max_smem)
        return fail(TSASR_E_UNSUPPORTED, "not enough shared memory (%d B) for H=%d", max_smem, H);
    p.num_w_stages = ns;
    // One narrow vocabulary tile on CTA pairs: W is small enough to stay in the W area for the whole kernel (see JointParams).
    p.w_resident = 0;
    p.w_stage_bytes = kWStageBytes;
    static const bool narrow_ok = getenv("TSASR_DEBUG_NO_NARROW") == nullptr;
    if (narrow_ok && use_pair() && p.NT == 1 && (size_t)(p.n_last / 2) * 128 * p.KB <= (size_t)ns * kWStageBytes) {
        p.w_resident = 1;
        p.w_stage_bytes = (p.n_last / 2) * 128;  // n_last is a multiple of 32: whole 8-row swizzle atoms
    }
    static const int dbg_skip = getenv("TSASR_DEBUG_SKIP") ? atoi(getenv("TSASR_DEBUG_SKIP")) : 0;
    p.dbg_skip = dbg_skip;
    p.enc = reinterpret_cast<const __nv_bfloat16*>(enc);
    p.dec = reinterpret_cast<const __nv_bfloat16*>(dec);
    return TSASR_OK;
}

struct JointMaps { CUtensorMap w; };

// W [V,H] as k-slices: pairs [128 v x 64 h] SWIZZLE_128B per CTA, single CTA [256 v x 32 h] SWIZZLE_64B
static int make_joint_maps(JointMaps* m, const JointParams& p, const void* enc, const void* dec, const void* W) {
    const int tT = 1 << p.tT_log2, tU = 128 >> p.tT_log2;
    if (use_pair()) {
        const uint32_t box_rows = p.w_resident ? (uint32_t)(p.n_last / 2) : (uint32_t)(kTileN / 2);
        if (int rc = make_tmap_2d_bf16(&m->w, W, (uint64_t)p.V, (uint64_t)p.H, kWStageKPair, box_rows, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    } else {
        if (int rc = make_tmap_2d_bf16(&m->w, W, (uint64_t)p.V, (uint64_t)p.H, kWStageK, kTileN, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
    }
    return TSASR_OK;
}

template <int MODE, bool PAIR, int NPW>
static int launch_joint_impl(const JointMaps& maps, const JointParams& p, int num_sms, cudaStream_t st) {
    const SmemLayout L = smem_layout(p.KB, p.num_w_stages);
    auto kern = joint_gemm_kernel<MODE, PAIR, NPW>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(joint_gemm_kernel)");
    const int tiles = p.tile_end - p.tile_begin;
    int grid;
    if (PAIR) {
        const int pairs = (tiles + 1) / 2, max_pairs = num_sms / 2;
        grid = 2 * (pairs < max_pairs ? pairs : max_pairs);
    } else {
        grid = tiles < num_sms ? tiles : num_sms;
    }
    if (grid <= 0) return TSASR_OK;
    static const bool prof_on = getenv("TSASR_DEBUG_PROF") != nullptr;  // development: MMA-lane wait breakdown
    JointParams pp = p;
    long long* d_prof = nullptr;
    if (prof_on) {
        cudaMalloc(&d_prof, sizeof(long long) * 16 * grid);
        cudaMemset(d_prof, 0, sizeof(long long) * 16 * grid);
        pp.prof = d_prof;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(Roles<NPW>::kThreads);
    cfg.dynamicSmemBytes = L.total;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_launch_attr(attr, 1);
    e = cudaLaunchKernelEx(&cfg, kern, maps.w, pp);
    ++g_launches;
    if (e != cudaSuccess) return cuda_fail(e, "joint_gemm_kernel launch");
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "joint_gemm_kernel launch");
    if (prof_on) {
        cudaStreamSynchronize(st);
        long long* h = new long long[16 * grid];
        cudaMemcpy(h, d_prof, sizeof(long long) * 16 * grid, cudaMemcpyDeviceToHost);
        double tot = 0, acc = 0, a = 0, w = 0, rounds = 0, lsum = 0, lcnt = 0, tcom = 0;
        int n = 0;
        for (int i = 0; i < grid; ++i)
            if (h[16 * i] > 0) { tot += h[16 * i]; acc += h[16 * i + 1]; a += h[16 * i + 2]; w += h[16 * i + 3]; rounds += h[16 * i + 4]; lsum += h[16 * i + 5]; lcnt += h[16 * i + 6]; tcom += h[16 * i + 7]; ++n; }
        if (lcnt > 0) fprintf(stderr, "[tsasr prof] MMA issue blocks: %.0f stages per cta, %.0f cycles to issue the 4 MMAs of a stage, %.0f cycles in commits per stage\n", lcnt / n, lsum / lcnt, tcom / lcnt);
        fprintf(stderr, "[tsasr prof] mode=%d pair=%d issuing ctas=%d rounds/cta=%.1f cycles/cta=%.0f wait: acc_empty=%.1f%% a_full=%.1f%% w_full=%.1f%% other=%.1f%%  cycles/round=%.0f\n",
                MODE, (int)PAIR, n, rounds / n, tot / n, 100 * acc / tot, 100 * a / tot, 100 * w / tot, 100 * (tot - acc - a - w) / tot, tot / rounds);
        for (int g = 0; g < 2; ++g) {  // epilogue warp 0 of each column group, averaged over all CTAs (builds with -DTSASR_EPI_PROF)
            double ew = 0, ep = 0, eb = 0, en = 0;
            for (int i = 0; i < grid; ++i) { ew += h[16 * i + 8 + 4 * g]; ep += h[16 * i + 9 + 4 * g]; eb += h[16 * i + 10 + 4 * g]; en += h[16 * i + 11 + 4 * g]; }
            if (en > 0) fprintf(stderr, "[tsasr prof] epilogue group %d, cycles per vocabulary tile: wait acc_full=%.0f process+release=%.0f bias barriers=%.0f\n", g, ew / en, ep / en, eb / en);
        }
        delete[] h;
        cudaFree(d_prof);
    }
    return TSASR_OK;
}

template <int MODE>
static int launch_joint(const JointMaps& maps, const JointParams& p, int num_sms, cudaStream_t st) {
    if (!use_pair()) return launch_joint_impl<MODE, false, 8>(maps, p, num_sms, st);
    // Narrow vocabularies (one vocabulary tile of at most 128 columns per cell tile, e.g. the 29 characters of the recipe as
    // shipped): the A operand is rebuilt for every 32..128 columns of MMA work, so the kernel is paced by its producers --
    // sixteen producer warps and one epilogue column group instead of eight and two.  TSASR_DEBUG_NO_NARROW=1: A/B runs.
    static const bool narrow_ok = getenv("TSASR_DEBUG_NO_NARROW") == nullptr;
    static const bool narrow_8 = getenv("TSASR_DEBUG_NARROW_8") != nullptr;  // development: resident W with the 8-producer layout
    if (narrow_ok && !narrow_8 && p.NT == 1 && p.n_last <= kTileN / 2) return launch_joint_impl<MODE, true, 16>(maps, p, num_sms, st);
    return launch_joint_impl<MODE, true, 8>(maps, p, num_sms, st);
}

extern "C" {

int tsasr_abi_version(void) { return TSASR_ABI_VERSION; }
const char* tsasr_last_error(void) { return g_err; }
long long tsasr_launch_count(void) { return g_launches.load(); }

int tsasr_kernel_timing_enable(int on) {
    g_timing_on.store(on != 0);
    return TSASR_OK;
}

int tsasr_kernel_timings(char* names, float* ms, int* counts, int max_n) {
    std::lock_guard<std::mutex> lock(g_timed_mu);
    int n = 0;
    for (int i = 0; i < g_n_timed; ++i) {
        float t = 0.f;
        cudaEventSynchronize(g_timed[i].e1);
        cudaEventElapsedTime(&t, g_timed[i].e0, g_timed[i].e1);
        cudaEventDestroy(g_timed[i].e0);
        cudaEventDestroy(g_timed[i].e1);
        int k = 0;
        for (; k < n; ++k)
            if (strncmp(names + 32 * k, g_timed[i].name, 31) == 0) break;
        if (k == n) {
            if (n >= max_n) continue;
            strncpy(names + 32 * n, g_timed[i].name, 31);
            names[32 * n + 31] = 0;
            ms[n] = 0.f;
            counts[n] = 0;
            ++n;
        }
        ms[k] += t;
        counts[k] += 1;
    }
    g_n_timed = 0;
    return n;
}

size_t tsasr_lattice_elems(int B, int T, int U) { return (size_t)B * (size_t)(T + U - 1) * (size_t)U; }

int tsasr_logits_to_lattice(const void* logits, int logits_dtype, const int32_t* targets, const int32_t* logit_lengths,
                            const int32_t* target_lengths, int B, int T, int U, int V, int blank, int normalized,
                            float* lat2, float* den, tsasr_stream_t stream) {
    NvtxRange nvtx_range("tsasr_logits_to_lattice");
    if (int rc = check_dims(B, T, U, V, blank)) return rc;
    REQUIRE(logits && logit_lengths && target_lengths && lat2 && den, "null pointer argument");
    REQUIRE(U == 1 || targets, "targets must not be null when U > 1");
    REQUIRE(logits_dtype >= 0 && logits_dtype <= 2, "unknown logits dtype %d", logits_dtype);
    ScopedTiming tm("logits_to_lattice_kernel", static_cast<cudaStream_t>(stream));
    cudaError_t e = launch_logits_to_lattice(logits, logits_dtype, targets, logit_lengths, target_lengths, B, T, U, V,
                                             blank, normalized, reinterpret_cast<float2*>(lat2), den,
                                             static_cast<cudaStream_t>(stream));
    ++g_launches;
    return e == cudaSuccess ? TSASR_OK : cuda_fail(e, "logits_to_lattice_kernel");
}

int tsasr_lattice_alpha_beta(const float* lat2, const int32_t* logit_lengths, const int32_t* target_lengths, int B,
                             int T, int U, float* alpha, float* beta, float* cost, float* ll_alpha, float* ll_beta,
                             tsasr_stream_t stream) {
    NvtxRange nvtx_range("tsasr_lattice_alpha_beta");
    if (int rc = check_dims(B, T, U, 1, 0)) return rc;
    REQUIRE(lat2 && logit_lengths && target_lengths && alpha && beta && cost && ll_alpha && ll_beta, "null pointer argument");
    if (U > kMaxLatticeWidth) return fail(TSASR_E_UNSUPPORTED, "lattice width U=%d > %d is not supported", U, kMaxLatticeWidth);
    ScopedTiming tm("alpha_beta_kernel", static_cast<cudaStream_t>(stream));
    cudaError_t e = launch_alpha_beta(reinterpret_cast<const float2*>(lat2), logit_lengths, target_lengths, B, T, U,
                                      alpha, beta, ll_alpha, ll_beta, cost, static_cast<cudaStream_t>(stream));
    g_launches += 2;
    return e == cudaSuccess ? TSASR_OK : cuda_fail(e, "alpha_beta_kernel");
}

int tsasr_logits_grad(const void* logits, int logits_dtype, const int32_t* targets, const int32_t* logit_lengths,
                      const int32_t* target_lengths, int B, int T, int U, int V, int blank, const float* lat2,
                      const float* den, const float* alpha, const float* beta, const float* cost, const float* dcost,
                      float clamp, void* dlogits, tsasr_stream_t stream) {
    NvtxRange nvtx_range("tsasr_logits_grad");
    if (int rc = check_dims(B, T, U, V, blank)) return rc;
    REQUIRE(logits && logit_lengths && target_lengths && lat2 && den && alpha && beta && cost && dlogits, "null pointer argument");
    REQUIRE(U == 1 || targets, "targets must not be null when U > 1");
    REQUIRE(logits_dtype >= 0 && logits_dtype <= 2, "unknown logits dtype %d", logits_dtype);
    ScopedTiming tm("logits_grad_kernel", static_cast<cudaStream_t>(stream));
    cudaError_t e = launch_logits_grad(logits, logits_dtype, targets, logit_lengths, target_lengths, B, T, U, V, blank,
                                       reinterpret_cast<const float2*>(lat2), den, alpha, beta, cost, dcost, clamp,
                                       dlogits, static_cast<cudaStream_t>(stream));
    ++g_launches;
    return e == cudaSuccess ? TSASR_OK : cuda_fail(e, "logits_grad_kernel");
}

int tsasr_logprobs_grad(const int32_t* targets, const int32_t* logit_lengths, const int32_t* target_lengths, int B,
                        int T, int U, int V, int blank, const float* lat2, const float* alpha, const float* beta,
                        const float* cost, const float* dcost, float* grads, tsasr_stream_t stream) {
    if (int rc = check_dims(B, T, U, V, blank)) return rc;
    REQUIRE(logit_lengths && target_lengths && lat2 && alpha && beta && cost && grads, "null pointer argument");
    REQUIRE(U == 1 || targets, "targets must not be null when U > 1");
    ScopedTiming tm("logprobs_grad_kernel", static_cast<cudaStream_t>(stream));
    cudaError_t e = launch_logprobs_grad(targets, logit_lengths, target_lengths, B, T, U, V, blank,
                                         reinterpret_cast<const float2*>(lat2), alpha, beta, cost, dcost, grads,
                                         static_cast<cudaStream_t>(stream));
    ++g_launches;
    return e == cudaSuccess ? TSASR_OK : cuda_fail(e, "logprobs_grad_kernel");
}

int tsasr_joint_fwd(const void* enc, const void* dec, const void* W, const float* bias, const int32_t* targets,
                    const int32_t* logit_lengths, const int32_t* target_lengths, int B, int T, int U, int H, int V,
                    int blank, int act_kind, float act_param, float* lat2, float* logz, tsasr_stream_t stream) {
    NvtxRange nvtx_range("tsasr_joint_fwd");
    if (int rc = check_dims(B, T, U, V, blank)) return rc;
    REQUIRE(enc && dec && W && bias && logit_lengths && target_lengths && lat2 && logz, "null pointer argument");
    REQUIRE(U == 1 || targets, "targets must not be null when U > 1");
    int sms, max_smem;
    if (int rc = device_info(&sms, &max_smem)) return rc;
    JointParams p;
    if (int rc = fill_joint_params(p, enc, dec, bias, targets, logit_lengths, target_lengths, B, T, U, H, V, blank,
                                   act_kind, act_param, max_smem))
        return rc;
    p.lat2 = reinterpret_cast<float2*>(lat2);
    p.logz = logz;
    JointMaps maps;
    if (int rc = make_joint_maps(&maps, p, enc, dec, W)) return rc;
    ScopedTiming tm("joint_gemm_kernel<FWD>", static_cast<cudaStream_t>(stream));
    return launch_joint<MODE_FWD>(maps, p, sms, static_cast<cudaStream_t>(stream));
}

// ---- the whole forward of the fused loss behind ONE call (host-path saver: five launches, no Python in between) ----
// scratch layout (byte offsets, every block 256-byte aligned): enc16 | dec16 | W16 | targets32 | ll | tl | stats[4]
static void fwd_scratch_layout(int B, int T, int U, int H, int V, size_t off[8]) {
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    off[0] = 0;
    off[1] = off[0] + up((size_t)B * T * H * 2);
    off[2] = off[1] + up((size_t)B * U * H * 2);
    off[3] = off[2] + up((size_t)V * H * 2);
    off[4] = off[3] + up((size_t)B * (U > 1 ? U - 1 : 1) * 4);
    off[5] = off[4] + up((size_t)B * 4);
    off[6] = off[5] + up((size_t)B * 4);
    off[7] = off[6] + 256;  // total
}

int tsasr_joint_loss_fwd_layout(int B, int T, int U, int H, int V, size_t* offsets8) {
    REQUIRE(offsets8 && B >= 1 && T >= 1 && U >= 1 && H >= 1 && V >= 1, "bad argument");
    fwd_scratch_layout(B, T, U, H, V, offsets8);
    return TSASR_OK;
}

int tsasr_joint_loss_fwd(const void* enc, const void* dec, const void* W, int operand_dtype, const float* bias, const void* targets,
                         int targets_i64, const float* rel_logit_lengths, const float* rel_target_lengths,
                         const int32_t* abs_logit_lengths, const int32_t* abs_target_lengths, int B, int T, int U, int H, int V,
                         int blank, int act_kind, float act_param, void* scratch, size_t scratch_bytes, int32_t* stats_host,
                         int stats_seq, float* lat2, float* logz, float* alpha, float* beta, float* cost3, const void* enc_bf16,
                         const void* dec_bf16, tsasr_stream_t stream) {
    NvtxRange nvtx_range("tsasr_joint_loss_fwd");
    if (int rc = check_dims(B, T, U, V, blank)) return rc;
    REQUIRE(enc && dec && W && bias && scratch && lat2 && logz && alpha && beta && cost3, "null pointer argument");
    REQUIRE(U == 1 || targets, "targets must not be null when U > 1");
    REQUIRE(operand_dtype == TSASR_F32 || operand_dtype == TSASR_BF16, "operands must be fp32 (converted here) or bf16 (used as they are)");
    REQUIRE((rel_logit_lengths || abs_logit_lengths) && (rel_target_lengths || abs_target_lengths), "null length argument");
    REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "scratch must be 256-byte aligned");
    if (U > kMaxLatticeWidth) return fail(TSASR_E_UNSUPPORTED, "lattice width U=%d > %d is not supported", U, kMaxLatticeWidth);
    size_t off[8];
    fwd_scratch_layout(B, T, U, H, V, off);
    if (scratch_bytes < off[7]) return fail(TSASR_E_WORKSPACE, "scratch too small: need %zu bytes, got %zu", off[7], scratch_bytes);
    int sms, max_smem;
    if (int rc = device_info(&sms, &max_smem)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* sc = static_cast<uint8_t*>(scratch);
    const bool cast = operand_dtype == TSASR_F32;
    REQUIRE(cast || (!enc_bf16 && !dec_bf16), "enc_bf16 / dec_bf16 accompany fp32 operands only");
    // an operand whose bf16 copy already exists (tsasr_linear_fwd's second output) is not converted again
    const bool cast_enc = cast && !enc_bf16, cast_dec = cast && !dec_bf16;
    const void* enc16 = cast_enc ? sc + off[0] : (cast ? enc_bf16 : enc);
    const void* dec16 = cast_dec ? sc + off[1] : (cast ? dec_bf16 : dec);
    const void* W16 = cast ? sc + off[2] : W;
    const int32_t* tg32 = targets_i64 ? reinterpret_cast<const int32_t*>(sc + off[3]) : static_cast<const int32_t*>(targets);
    int32_t* ll = reinterpret_cast<int32_t*>(sc + off[4]);
    int32_t* tl = reinterpret_cast<int32_t*>(sc + off[5]);
    int32_t* stats = reinterpret_cast<int32_t*>(sc + off[6]);
    if (cast) {
        REQUIRE(((size_t)B * T * H) % 4 == 0 && ((size_t)B * U * H) % 4 == 0 && ((size_t)V * H) % 4 == 0, "operand element counts must be multiples of 4");
        REQUIRE(((reinterpret_cast<uintptr_t>(enc) | reinterpret_cast<uintptr_t>(dec) | reinterpret_cast<uintptr_t>(W)) & 15) == 0,
                "fp32 operands must be 16-byte aligned");
    }
    int32_t* stats_dev = nullptr;  // device-side address of the caller's mapped pinned landing slot (or none)
    if (stats_host) {
        void* dptr = nullptr;
        cudaError_t e = cudaHostGetDevicePointer(&dptr, stats_host, 0);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(TSASR_E_INVALID, "stats_host is not mapped pinned host memory: %s", cudaGetErrorString(e)); }
        stats_dev = static_cast<int32_t*>(dptr);
    }
    JointParams p;
    if (int rc = fill_joint_params(p, enc16, dec16, bias, tg32, ll, tl, B, T, U, H, V, blank, act_kind, act_param, max_smem)) return rc;
    {
        ScopedTiming tm("prepare_inputs_kernel", st);
        cudaError_t e = launch_prepare_inputs(
            cast_enc ? static_cast<const float*>(enc) : nullptr, cast_enc ? (size_t)B * T * H : 0, cast_dec ? static_cast<const float*>(dec) : nullptr,
            cast_dec ? (size_t)B * U * H : 0, cast ? static_cast<const float*>(W) : nullptr, cast ? (size_t)V * H : 0, sc + off[0], sc + off[1],
            sc + off[2], targets_i64 ? static_cast<const long long*>(targets) : nullptr, (size_t)B * (U - 1),
            reinterpret_cast<int*>(sc + off[3]), rel_logit_lengths, rel_target_lengths, abs_logit_lengths, abs_target_lengths, B, T, U - 1,
            ll, tl, stats, stats_dev, stats_seq, sms, st);
        ++g_launches;
        if (e != cudaSuccess) return cuda_fail(e, "prepare_inputs_kernel");
    }
    p.lat2 = reinterpret_cast<float2*>(lat2);
    p.logz = logz;
    JointMaps maps;
    if (int rc = make_joint_maps(&maps, p, enc16, dec16, W16)) return rc;
    {
        ScopedTiming tm("joint_gemm_kernel<FWD>", st);
        if (int rc = launch_joint<MODE_FWD>(maps, p, sms, st)) return rc;
    }
    ScopedTiming tm("alpha_beta_kernel", st);
    cudaError_t e = launch_alpha_beta(reinterpret_cast<const float2*>(lat2), ll, tl, B, T, U, alpha, beta, cost3 + B, cost3 + 2 * B, cost3, st);
    g_launches += 2;
    return e == cudaSuccess ? TSASR_OK : cuda_fail(e, "alpha_beta_kernel");
}

int tsasr_prepare_lengths(const float* rel_logit_lengths, const float* rel_target_lengths, const int32_t* abs_logit_lengths,
                          const int32_t* abs_target_lengths, int B, int T, int n_targets, int32_t* logit_lengths_out,
                          int32_t* target_lengths_out, int32_t* stats_out, tsasr_stream_t stream) {
    REQUIRE(B >= 1 && T >= 1 && n_targets >= 0, "B, T must be >= 1 and n_targets >= 0");
    REQUIRE((rel_logit_lengths || abs_logit_lengths) && (rel_target_lengths || abs_target_lengths) && stats_out,
            "null pointer argument");
    cudaError_t e = launch_lengths(rel_logit_lengths, rel_target_lengths, abs_logit_lengths, abs_target_lengths, B, T, n_targets,
                                   logit_lengths_out, target_lengths_out, stats_out, static_cast<cudaStream_t>(stream));
    ++g_launches;
    return e == cudaSuccess ? TSASR_OK : cuda_fail(e, "lengths_kernel");
}

int tsasr_cast_operands_bf16(const float* enc, size_t n_enc, const float* dec, size_t n_dec, const float* W, size_t n_w,
                             void* enc16, void* dec16, void* W16, tsasr_stream_t stream) {
    REQUIRE(enc && dec && W && enc16 && dec16 && W16, "null pointer argument");
    REQUIRE(n_enc % 4 == 0 && n_dec % 4 == 0 && n_w % 4 == 0, "element counts must be multiples of 4");
    REQUIRE(((reinterpret_cast<uintptr_t>(enc) | reinterpret_cast<uintptr_t>(dec) | reinterpret_cast<uintptr_t>(W)) & 15) == 0 &&
                ((reinterpret_cast<uintptr_t>(enc16) | reinterpret_cast<uintptr_t>(dec16) | reinterpret_cast<uintptr_t>(W16)) & 7) == 0,
            "operands must be 16-byte (fp32) / 8-byte (bf16) aligned");
    int sms, max_smem;
    if (int rc = device_info(&sms, &max_smem)) return rc;
    cudaError_t e = launch_cast3(enc, n_enc, dec, n_dec, W, n_w, enc16, dec16, W16, sms, static_cast<cudaStream_t>(stream));
    ++g_launches;
    return e == cudaSuccess ? TSASR_OK : cuda_fail(e, "cast3_kernel");
}

size_t tsasr_joint_decode_workspace_bytes(int V) { return V >= 1 ? joint_decode_workspace_bytes(V) : 0; }

int tsasr_joint_decode_step(const float* enc_t, const float* dec, long long enc_row_stride, long long dec_row_stride,
                            const float* W, const float* bias, int B, int H, int V, int act_kind, float act_param,
                            float* log_probs, void* workspace, size_t workspace_bytes, tsasr_stream_t stream) {
    REQUIRE(B >= 1 && H >= 4 && V >= 1, "B, V must be >= 1 and H >= 4 (got %d %d %d)", B, H, V);
    REQUIRE(H % 4 == 0, "decode step needs H %% 4 == 0 (got %d)", H);
    REQUIRE(H <= 1536, "decode step supports H <= 1536 (got %d)", H);
    REQUIRE(enc_t && dec && W && log_probs && workspace, "null pointer argument");
    REQUIRE(act_kind >= 0 && act_kind <= 3, "unknown activation code %d", act_kind);
    REQUIRE((reinterpret_cast<uintptr_t>(W) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
            "W and workspace must be 16-byte aligned");
    if (workspace_bytes < joint_decode_workspace_bytes(V))
        return fail(TSASR_E_WORKSPACE, "workspace too small: need %zu bytes, got %zu", joint_decode_workspace_bytes(V), workspace_bytes);
    int sms, max_smem;
    if (int rc = device_info(&sms, &max_smem)) return rc;
    ScopedTiming tm("joint_decode_step", static_cast<cudaStream_t>(stream));
    REQUIRE(enc_row_stride >= 0 && dec_row_stride >= 0, "row strides must be >= 0 (0 = broadcast one row over B)");
    cudaError_t e = launch_joint_decode_step(enc_t, dec, enc_row_stride, dec_row_stride, W, bias, B, H, V, act_kind, act_param,
                                             log_probs, workspace, static_cast<cudaStream_t>(stream));
    g_launches += 2 * ((B + 31) / 32);
    return e == cudaSuccess ? TSASR_OK : cuda_fail(e, "joint_decode_step kernels");
}

int tsasr_joint_debug_logits(const void* enc, const void* dec, const void* W, const float* bias, int B, int T, int U,
                             int H, int V, int act_kind, float act_param, float* logits_out, tsasr_stream_t stream) {
    if (int rc = check_dims(B, T, U, V, 0)) return rc;
    REQUIRE(enc && dec && W && bias && logits_out, "null pointer argument");
    int sms, max_smem;
    if (int rc = device_info(&sms, &max_smem)) return rc;
    // full lengths: a device-side length array is required by the kernel; the caller-free variant
    // builds one in the first bytes of logits_out?  No -- keep it explicit: lengths are passed through
    // a small static device buffer filled here.
    static int32_t* d_len = nullptr;
    static int d_len_cap = 0;
    if (d_len_cap < 2 * B) {
        if (d_len) cudaFree(d_len);
        if (cudaMalloc(&d_len, sizeof(int32_t) * 2 * B) != cudaSuccess) return fail(TSASR_E_CUDA, "cudaMalloc (debug lengths)");
        d_len_cap = 2 * B;
    }
    int32_t* h = new int32_t[2 * B];
    for (int i = 0; i < B; ++i) { h[i] = T; h[B + i] = U - 1; }
    cudaError_t e = cudaMemcpyAsync(d_len, h, sizeof(int32_t) * 2 * B, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream));
    cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    delete[] h;
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync (debug lengths)");
    JointParams p;
    if (int rc = fill_joint_params(p, enc, dec, bias, nullptr, d_len, d_len + B, B, T, U, H, V, 0, act_kind, act_param, max_smem))
        return rc;
    p.dbg_logits = logits_out;
    JointMaps maps;
    if (int rc = make_joint_maps(&maps, p, enc, dec, W)) return rc;
    return launch_joint<MODE_DEBUG>(maps, p, sms, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

#include "backward_capi.inl"
#include "linear_capi.inl"
#include "predictor_capi.inl"
