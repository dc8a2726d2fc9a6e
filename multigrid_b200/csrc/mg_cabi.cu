// multigrid_b200 -- C ABI (include/multigrid_b200.h) over the sm_100a kernels in mg_kernels.cuh.
// Argument validation, launch geometry and kernel dispatch only; no torch, no global state
// besides a launch counter.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <unistd.h>

#include "mg_kernels.cuh"
#include "mg_static.cuh"

namespace {

std::atomic<int64_t> g_launches{0};
std::atomic<unsigned long long *> g_trace{nullptr};  // diagnostics only (mg_debug_set_trace)

// A prepared launch (mg_step_plan_*): what a step call would hand to cudaLaunchKernelEx, kept so that later steps
// skip validation, planning and the knob lookups.
struct LaunchRecord {
    mg::Params p;
    const void *func;
    dim3 grid, block;
    size_t smem;
    int pdl;
};
thread_local LaunchRecord *g_capture = nullptr;  // non-null while mg_step_plan_create runs the planning code

int launch_record(const LaunchRecord &r, cudaStream_t stream) {
    cudaLaunchConfig_t lc;
    std::memset(&lc, 0, sizeof(lc));
    lc.gridDim = r.grid; lc.blockDim = r.block; lc.dynamicSmemBytes = r.smem; lc.stream = stream;
    // Programmatic dependent launch: the grid may be scheduled while the previous kernel of the stream drains; the
    // kernels execute griddepcontrol.wait before their first global access, so stream-order semantics are unchanged
    // (MG_PDL=0 turns the attribute off).
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr; lc.numAttrs = r.pdl ? 1 : 0;
    void *args[1] = {const_cast<mg::Params *>(&r.p)};
    const cudaError_t err = cudaLaunchKernelExC(&lc, r.func, args);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)err;
}

int launch_or_record(const mg::Params &p, const void *func, dim3 grid, dim3 block, size_t smem, cudaStream_t stream) {
    LaunchRecord local;
    LaunchRecord &r = g_capture ? *g_capture : local;
    r.p = p; r.func = func; r.grid = grid; r.block = block; r.smem = smem; r.pdl = p.pdl;
    return g_capture ? 0 : launch_record(r, stream);
}

constexpr int kSmemPerBlock = 227 * 1024;  // B200 opt-in maximum per block
constexpr int kSmemPerSM = 228 * 1024;

// Tuning / test knobs are environment variables named MG_*, read on every call (tests flip them between
// launches). One pass over `environ` finds out whether any exists at all -- normally none does, and the
// seven lookups of a launch then cost nothing.
bool any_mg_knob() {
    for (char **e = environ; e && *e; ++e)
        if ((*e)[0] == 'M' && (*e)[1] == 'G' && (*e)[2] == '_') return true;
    return false;
}

int env_int(const char *name, int dflt, bool present = true) {
    if (!present) return dflt;
    const char *s = std::getenv(name);
    return (s && *s) ? std::atoi(s) : dflt;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int validate(const MgConfig *c, int64_t num_envs) {
    if (!c || num_envs < 0 || num_envs > (1ll << 30)) return MG_ERR_BAD_ARG;
    if (c->width < 1 || c->height < 1 || c->width > 127 || c->height > 127) return MG_ERR_BAD_ARG;
    if (c->num_agents < 1 || c->num_agents > MG_MAX_AGENTS) return MG_ERR_BAD_ARG;
    if (c->view_size < 3 || c->view_size > MG_MAX_VIEW || !(c->view_size & 1)) return MG_ERR_BAD_ARG;
    if (c->max_steps < 1) return MG_ERR_BAD_ARG;
    if (c->obs_agent_stride < 3 * c->view_size * c->view_size || (c->obs_agent_stride & 3)) return MG_ERR_BAD_ARG;
    if ((c->flags & MG_FLAG_AUTO_RESET) && (c->num_layouts < 1 || c->layout_stride < 0)) return MG_ERR_BAD_ARG;
    return 0;
}

// Tuning / test knobs (read per call): MG_GROUP = envs per warp (16|32), MG_WPB = warps per block,
// MG_NO_BULK=1 = plain loads/stores instead of TMA bulk copies, MG_PDL=0 = no programmatic dependent
// launch, MG_L2HINT = override of the MG_FLAG_STREAM_STATE cache policy (bit 0 loads, bit 1 obs stores).
int plan(mg::Params &p) {
    const bool k = any_mg_knob();
    p.use_bulk = env_int("MG_NO_BULK", 0, k) ? 0 : 1;
    p.generic_view = env_int("MG_GENERIC_VIEW", 0, k) ? 1 : 0;
    p.pdl = env_int("MG_PDL", 1, k) ? 1 : 0;
    if (env_int("MG_NO_DEDUP", 0, k)) p.pool_rep = nullptr;  // knob: always read every env's own grid
    p.l2hint = env_int("MG_L2HINT", (p.flags & MG_FLAG_STREAM_STATE) ? 3 : 0, k);
    static thread_local int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int n_sm = 148;
    if (dev >= 0 && dev < 64) {
        if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
        if (sms[dev] > 0) n_sm = sms[dev];
    }
    p.num_sms = env_int("MG_NO_ROOMY", 0, k) ? 0 : n_sm;  // (knob: 0 disables the uncapped instantiations)
    return mg::plan_launch(p, env_int("MG_GROUP", 0, k), env_int("MG_WPB", 0, k), kSmemPerBlock - 16 /* the claim counter */, kSmemPerSM, n_sm);
}

template <int VT, int MODE, bool MULTI = false, bool CHAIN = false, int NT = 0, int HK = -1, bool OH = false, bool ROOMY = false>
int launch(const mg::Params &p, cudaStream_t stream) {
    auto kernel = mg::step_obs_kernel<VT, MODE, MULTI, CHAIN, NT, HK, OH, ROOMY>;
    static thread_local bool configured_dev[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured_dev[dev]) {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemPerBlock);
        if (err != cudaSuccess) return (int)err;
        configured_dev[dev] = true;
    }
    const int groups = (p.num_envs + p.G - 1) / p.G;
    const int blocks = (groups + p.wpb - 1) / p.wpb;
    return launch_or_record(p, (const void *)kernel, dim3((unsigned)blocks), dim3((unsigned)(p.wpb * mg::LANES)),
                            (size_t)(p.wpb * p.warp_bytes) + 16 /* the block's claim counter */, stream);
}

template <int MODE, bool MULTI = false>
int dispatch(const mg::Params &p, cudaStream_t stream) {
    if constexpr (MODE == mg::MODE_STEP) {
        return launch<0, MODE>(p, stream);  // no observation phase: view size is irrelevant
    } else {
        if constexpr (MODE == mg::MODE_STEP_OBS && !MULTI) {
            if (p.one_hot) {  // the launch also writes the one-hot images (plain launches only: step_common)
                if (!p.generic_view) {
                    switch (p.V) {
                        case 3: return launch<3, MODE, false, false, 0, -1, true>(p, stream);
                        case 5: return launch<5, MODE, false, false, 0, -1, true>(p, stream);
                        case 7: return launch<7, MODE, false, false, 0, -1, true>(p, stream);
                        case 9: return launch<9, MODE, false, false, 0, -1, true>(p, stream);
                        default: break;
                    }
                }
                return launch<0, MODE, false, false, 0, -1, true>(p, stream);
            }
            if (p.chained) {
                if (!p.generic_view) {
                    switch (p.V) {
                        case 3: return launch<3, MODE, false, true>(p, stream);
                        case 5: return launch<5, MODE, false, true>(p, stream);
                        case 7: return launch<7, MODE, false, true>(p, stream);
                        case 9: return launch<9, MODE, false, true>(p, stream);
                        default: break;
                    }
                }
                return launch<0, MODE, false, true>(p, stream);
            }
        }
        if constexpr (MODE == mg::MODE_STEP_OBS && !MULTI) {
            // a launch whose blocks are all resident at 4 per SM takes the instantiation without the register cap
            const int groups = (p.num_envs + p.G - 1) / p.G, blocks = (groups + p.wpb - 1) / p.wpb;
            const bool roomy = p.wpb == 4 && blocks <= 4 * p.num_sms;
            // BASELINE configs[2] (BlockedUnlockPickup, 2 agents, view 7): agent count and hook compiled in
            if (!p.generic_view && p.V == 7 && p.n == 2 && p.hook == MG_HOOK_BLOCKED_UNLOCK_PICKUP)
                return roomy ? launch<7, MODE, false, false, 2, MG_HOOK_BLOCKED_UNLOCK_PICKUP, false, true>(p, stream)
                             : launch<7, MODE, false, false, 2, MG_HOOK_BLOCKED_UNLOCK_PICKUP>(p, stream);
            if (!p.generic_view && p.V == 7 && roomy) return launch<7, MODE, false, false, 0, -1, false, true>(p, stream);
        }
        if (!p.generic_view) {
            switch (p.V) {
                case 3: return launch<3, MODE, MULTI>(p, stream);
                case 5: return launch<5, MODE, MULTI>(p, stream);
                case 7: return launch<7, MODE, MULTI>(p, stream);
                case 9: return launch<9, MODE, MULTI>(p, stream);
                default: break;
            }
        }
        return launch<0, MODE, MULTI>(p, stream);
    }
}

// The static-grid path (mg_static.cuh): one warp per block; unrolled instantiations for the BASELINE shapes
// (V = 7 with 4 or 2 agents, V = 9 with 8) at 8 / 16 / 32 envs per warp, rolled loops otherwise.
template <typename K>
int launch_static(K kernel, const mg::Params &p, cudaStream_t stream) {
    const int groups = (p.num_envs + p.G - 1) / p.G;
    if (p.warp_bytes > 48 * 1024) {
        const cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemPerBlock);
        if (err != cudaSuccess) return (int)err;
    }
    // (carve_static: warp_bytes = the bytes of the whole block)
    return launch_or_record(p, (const void *)kernel, dim3((unsigned)groups), dim3((unsigned)(p.wpb * mg::LANES)),
                            (size_t)p.warp_bytes, stream);
}

int dispatch_static(const mg::Params &p, cudaStream_t stream) {
    if (mg::static_fast_shape(p)) {
#define MG_STATIC_CASE(V_, N_, G_, W_) \
        if (p.V == V_ && p.n == N_ && p.G == G_ && p.wpb == W_) \
            return launch_static(mg::static_fast_kernel<V_, N_, G_, W_>, p, stream);
        // (the instantiations listed in mg::STATIC_SHAPES)
        MG_STATIC_CASE(7, 4, 16, 1) MG_STATIC_CASE(7, 4, 32, 2) MG_STATIC_CASE(7, 4, 32, 1)
        MG_STATIC_CASE(7, 4, 16, 2) MG_STATIC_CASE(7, 4, 8, 1)
        MG_STATIC_CASE(7, 2, 32, 1) MG_STATIC_CASE(7, 2, 32, 2) MG_STATIC_CASE(7, 2, 16, 1)
        MG_STATIC_CASE(9, 8, 16, 4) MG_STATIC_CASE(9, 8, 32, 4) MG_STATIC_CASE(9, 8, 16, 2)
        MG_STATIC_CASE(9, 8, 8, 2) MG_STATIC_CASE(9, 8, 8, 1)
#undef MG_STATIC_CASE
        return MG_ERR_BAD_ARG;  // (plan_static only produces the shapes above)
    }
    return launch_static(mg::static_rolled_kernel, p, stream);
}

void fill_config(mg::Params &p, const MgConfig *c, int64_t num_envs) {
    std::memset(&p, 0, sizeof(p));
    p.W = c->width; p.H = c->height; p.n = c->num_agents; p.V = c->view_size;
    p.max_steps = c->max_steps; p.flags = c->flags; p.hook = c->hook; p.hook_param = c->hook_param;
    p.ostride = c->obs_agent_stride; p.K = c->num_layouts; p.lstride = c->layout_stride;
    p.num_envs = (int32_t)num_envs;
    p.T = 1;
    p.rcp_vv = mg::rcp32(p.V * p.V);
    p.trace = g_trace.load(std::memory_order_relaxed);
}

int fill_state(mg::Params &p, const MgState *s) {
    if (!s || !s->grid || !s->agents || !s->step_count) return MG_ERR_BAD_ARG;
    if (p.n > 1 && (!s->pcg_state || !s->pcg_inc)) return MG_ERR_BAD_ARG;
    if ((p.flags & MG_FLAG_AUTO_RESET) && (!s->layout_idx || !s->pool_grid || !s->pool_agents)) return MG_ERR_BAD_ARG;
    if (!aligned16(s->grid) || !aligned16(s->agents) || !aligned16(s->pool_agents)) return MG_ERR_ALIGNMENT;
    p.grid = s->grid; p.agents = s->agents; p.step_count = s->step_count;
    p.pcg_state = s->pcg_state; p.pcg_inc = s->pcg_inc; p.layout_idx = s->layout_idx;
    p.pool_grid = s->pool_grid; p.pool_agents = s->pool_agents;
    if (p.hook == MG_HOOK_LOCKED_HALLWAY && !s->hook_state) return MG_ERR_BAD_ARG;
    p.hook_state = s->hook_state;
    if (reinterpret_cast<uintptr_t>(s->pool_rep) & 15u) return MG_ERR_ALIGNMENT;
    p.pool_rep = (s->pool_rep && s->chain && p.K == 1) ? s->pool_rep : nullptr;
    if (reinterpret_cast<uintptr_t>(s->chain) & 15u) return MG_ERR_ALIGNMENT;
    p.chain = s->chain;
    if ((p.flags & MG_FLAG_CHAINED) && !s->chain) return MG_ERR_BAD_ARG;
    p.chained = (p.flags & MG_FLAG_CHAINED) ? ((p.flags & MG_FLAG_CHAIN_HEAD) ? 2 : 1) : 0;
    if (reinterpret_cast<uintptr_t>(s->static_obs) & 15u) return MG_ERR_ALIGNMENT;
    p.static_obs = reinterpret_cast<const uint8_t *>(s->static_obs);
    p.static_stride = mg::static_obs_stride(p.ostride);
    p.static_move = reinterpret_cast<const uint32_t *>(p.static_obs + (size_t)p.W * p.H * 4 * p.static_stride);
    return 0;
}

int fill_out(mg::Params &p, const MgStepOut *o, bool need_obs) {
    if (!o || !o->reward || !o->terminated || !o->truncated) return MG_ERR_BAD_ARG;
    if (need_obs && !o->obs) return MG_ERR_BAD_ARG;
    if (need_obs && !aligned16(o->obs)) return MG_ERR_ALIGNMENT;
    p.obs = o->obs; p.reward = o->reward; p.terminated = o->terminated; p.truncated = o->truncated;
    p.status = o->status;
    if (need_obs && o->one_hot && !aligned16(o->one_hot)) return MG_ERR_ALIGNMENT;
    p.one_hot = need_obs ? o->one_hot : nullptr;  // (mg_step produces no observations: ignored there)
    return 0;
}

template <int MODE, bool MULTI = false>
int step_common(const MgConfig *cfg, int64_t num_envs, const MgState *state, const int8_t *actions,
                const MgStepOut *out, void *stream, int32_t num_steps = 1, int8_t *direction = nullptr) {
    int rc = validate(cfg, num_envs);
    if (rc) return rc;
    if (num_steps < 0 || (int64_t)num_steps * num_envs > (1ll << 40)) return MG_ERR_BAD_ARG;
    if (num_envs == 0 || num_steps == 0) return 0;
    if (!actions) return MG_ERR_BAD_ARG;
    mg::Params p;
    fill_config(p, cfg, num_envs);
    if ((rc = fill_state(p, state))) return rc;
    if ((rc = fill_out(p, out, MODE == mg::MODE_STEP_OBS))) return rc;
    p.actions = actions;
    p.T = num_steps;
    p.direction = direction;
    if (MULTI || MODE != mg::MODE_STEP_OBS) p.one_hot = nullptr;  // (fill_out rejects it for mg_step; a rollout has none)
    if (p.one_hot && (p.flags & MG_FLAG_CHAINED)) return MG_ERR_BAD_ARG;  // the one-hot variants are plain launches
    if (MULTI || MODE != mg::MODE_STEP_OBS) p.chained = 0;  // only the fused single-step launch chains; a rollout
                                                            // or mg_step is a plain launch and leaves the tickets alone
    // MG_FLAG_STATIC_GRID: the fused single-step launch on a grid no action can change (mg_static.cuh)
    if (MODE == mg::MODE_STEP_OBS && !MULTI && (p.flags & MG_FLAG_STATIC_GRID) && !p.chained) {
        if (!p.static_obs || p.hook != MG_HOOK_NONE || !state->pool_grid || !state->pool_agents) return MG_ERR_BAD_ARG;
        if ((p.flags & MG_FLAG_AUTO_RESET) && p.K != 1) return MG_ERR_BAD_ARG;
        const bool k = any_mg_knob();
        if (!env_int("MG_NO_STATIC", 0, k)) {
            p.use_bulk = env_int("MG_NO_BULK", 0, k) ? 0 : 1;
            p.generic_view = env_int("MG_GENERIC_VIEW", 0, k) ? 1 : 0;
            p.pdl = env_int("MG_PDL", 1, k) ? 1 : 0;
            // (the observation stores always carry the L2 evict_first hint here: they are written once and, at
            // 40+ MB per launch, would otherwise push the state out of L2; measured 12.1 -> 11.1 us per launch)
            p.l2hint = env_int("MG_L2HINT", 3, k);
            // the unrolled kernels read an env's actions as one 2 / 4 / 8-byte word
            if (reinterpret_cast<uintptr_t>(p.actions) & 7u) p.generic_view = 1;
            static thread_local int sms[64] = {0};
            int dev = 0, n_sm = 148;
            cudaGetDevice(&dev);
            if (dev >= 0 && dev < 64) {
                if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
                if (sms[dev] > 0) n_sm = sms[dev];
            }
            p.no_lut = env_int("MG_NO_LUT", 0, k);
            if ((rc = mg::plan_static(p, env_int("MG_GROUP", 0, k), env_int("MG_WPB", 0, k), kSmemPerBlock, n_sm))) return rc;
            auto misaligned = [](const void *ptr, uintptr_t a) { return (reinterpret_cast<uintptr_t>(ptr) & (a - 1)) != 0; };
            if (misaligned(p.pcg_state, 16) || misaligned(p.pcg_inc, 16) || misaligned(p.reward, 8) ||
                misaligned(p.step_count, 4) || misaligned(p.terminated, 4) || misaligned(p.pool_grid, 4) ||
                misaligned(p.pool_agents, 4)) return MG_ERR_ALIGNMENT;
            return dispatch_static(p, (cudaStream_t)stream);
        }
    }
    if ((rc = plan(p))) return rc;
    // a rollout re-reads its cells from L2 every step: never mark those loads evict_first
    if (MULTI) p.l2hint &= ~1;
    if (p.chained && !p.pdl) p.chained = 2;  // without programmatic launch every launch waits like a chain head
    // the actions span of a group (G * n bytes) must be a multiple of 16 for the bulk copy: 8 envs x odd n is not
    if ((p.G * p.n) & 15) p.use_bulk = 0;
    // TMA spans of step t start at t * E * n (actions) and t * E * n * stride (obs) bytes
    if (MULTI && (((size_t)num_envs * p.n) & 15u)) p.use_bulk = 0;
    // natural alignment of the per-env scalars (16-byte PCG words, 8-byte rewards, 4-byte counters and
    // the packed 4-agent terminated word) is required; 16-byte alignment of everything enables TMA
    auto misaligned = [](const void *ptr, uintptr_t a) { return (reinterpret_cast<uintptr_t>(ptr) & (a - 1)) != 0; };
    if (misaligned(p.pcg_state, 16) || misaligned(p.pcg_inc, 16) || misaligned(p.reward, 8) ||
        misaligned(p.step_count, 4) || misaligned(p.layout_idx, 4) || misaligned(p.hook_state, 4) ||
        misaligned(p.terminated, 4) || misaligned(p.pool_grid, 4))
        return MG_ERR_ALIGNMENT;
    // TMA bulk copies need 16-byte aligned spans; the small arrays may fall back to plain copies
    if (!aligned16(p.actions) || !aligned16(p.step_count) || !aligned16(p.pcg_state) || !aligned16(p.pcg_inc) ||
        !aligned16(p.layout_idx) || !aligned16(p.reward) || !aligned16(p.terminated) || !aligned16(p.truncated))
        p.use_bulk = 0;
    return dispatch<MODE, MULTI>(p, (cudaStream_t)stream);
}

}  // namespace

extern "C" {

int mg_abi_version(void) { return MG_ABI_VERSION; }

const char *mg_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case MG_ERR_BAD_ARG: return "multigrid_b200: bad argument";
        case MG_ERR_ALIGNMENT: return "multigrid_b200: device pointers must be 16-byte aligned";
        case MG_ERR_TOO_LARGE: return "multigrid_b200: grid/view too large for one thread block's shared memory";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "multigrid_b200: unknown error";
    }
}

int32_t mg_obs_agent_stride(int32_t view_size) { return (3 * view_size * view_size + 3) & ~3; }

int64_t mg_cells_per_env(int32_t width, int32_t height) { return mg::cells_per_env(width, height); }

int mg_pack_grid(int32_t width, int32_t height, int64_t num_envs, const int8_t *grid3, uint32_t *cells,
                 void *stream) {
    if (width < 1 || height < 1 || width > 127 || height > 127 || num_envs < 0) return MG_ERR_BAD_ARG;
    if (num_envs == 0) return 0;
    if (!grid3 || !cells) return MG_ERR_BAD_ARG;
    const int64_t total = num_envs * mg::cells_per_env(width, height);
    mg::pack_grid_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(width, height, total, grid3, cells);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int mg_full_obs(int32_t width, int32_t height, int32_t num_agents, int64_t num_envs, const uint32_t *cells,
                const int8_t *agents, int8_t *out, void *stream) {
    if (width < 1 || height < 1 || width > 127 || height > 127 || num_envs < 0 || num_agents < 1 ||
        num_agents > MG_MAX_AGENTS) return MG_ERR_BAD_ARG;
    if (num_envs == 0) return 0;
    if (!cells || !agents || !out) return MG_ERR_BAD_ARG;
    const int64_t total = num_envs * width * height;
    mg::full_obs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        width, height, num_agents, total, cells, agents, out);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

// One-hot of `images` images of `cells` cells each (3 bytes per cell, `stride` bytes apart).
static int one_hot_cells(int64_t cells, int64_t images, int32_t stride, const int8_t *obs, uint8_t *out, void *stream) {
    if (cells < 1 || cells > 127 * 127 || images < 0 || stride < 3 * cells) return MG_ERR_BAD_ARG;
    if (images == 0) return 0;
    if (!obs || !out) return MG_ERR_BAD_ARG;
    if (reinterpret_cast<uintptr_t>(out) & 3u) return MG_ERR_ALIGNMENT;
    // the 16-byte kernel needs 16-byte aligned blocks of 16 images: cells * 21 * 16 bytes each, always a multiple of 16
    if ((reinterpret_cast<uintptr_t>(out) & 15u) == 0 && cells <= 1024 && !env_int("MG_ONE_HOT_W32", 0)) {
        if (!env_int("MG_ONE_HOT_V16", 0))  // (knob: the previous 16-bytes-per-thread kernel)
            mg::one_hot_tile_kernel<<<(unsigned)((images + 31) / 32), 256, 0, (cudaStream_t)stream>>>(
                (int)cells, images, stride, mg::rcp32((int)cells), obs, out);
        else
            mg::one_hot_kernel_v16<<<(unsigned)((images + 15) / 16), 256, 0, (cudaStream_t)stream>>>(
                (int)cells, images, stride, mg::rcp32((int)cells), obs, (uint4 *)out);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return (int)cudaGetLastError();
    }
    const int64_t words = (images * cells * 21 + 3) / 4;
    mg::one_hot_kernel<<<(unsigned)((words + 255) / 256), 256, 0, (cudaStream_t)stream>>>((int)cells, images, stride, obs, out);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int mg_one_hot(int32_t view_size, int64_t num_agents_total, int32_t obs_agent_stride, const int8_t *obs,
               uint8_t *out, void *stream) {
    if (view_size < 3 || view_size > MG_MAX_VIEW) return MG_ERR_BAD_ARG;
    return one_hot_cells((int64_t)view_size * view_size, num_agents_total, obs_agent_stride, obs, out, stream);
}

int mg_one_hot_cells(int64_t cells_per_image, int64_t num_images, int32_t image_stride, const int8_t *images,
                     uint8_t *out, void *stream) {
    return one_hot_cells(cells_per_image, num_images, image_stride, images, out, stream);
}

int mg_gen_layouts_empty_random(int32_t width, int32_t height, int32_t num_agents, int64_t num_layouts,
                                uint64_t *rng_state, const uint64_t *rng_inc, uint64_t *rng_buf, uint32_t *cells,
                                int8_t *agents, int32_t *status, void *stream) {
    if (width < 3 || height < 3 || width > 127 || height > 127 || num_layouts < 0 || num_agents < 1 ||
        num_agents > MG_MAX_AGENTS) return MG_ERR_BAD_ARG;
    if (num_agents > (width - 2) * (height - 2) - 1) return MG_ERR_BAD_ARG;  // more agents than free cells
    if (num_layouts == 0) return 0;
    if (!rng_state || !rng_inc || !cells || !agents) return MG_ERR_BAD_ARG;
    mg::gen_layouts_empty_random_kernel<<<(unsigned)((num_layouts + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        width, height, num_agents, num_layouts, rng_state, rng_inc, rng_buf, cells, agents, status);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int mg_gen_layouts_red_blue_doors(int32_t size, int32_t num_agents, int64_t num_layouts, uint64_t *rng_state,
                                  const uint64_t *rng_inc, uint64_t *rng_buf, uint32_t *cells, int8_t *agents,
                                  int32_t *status, void *stream) {
    if (size < 4 || size > 63 || num_layouts < 0 || num_agents < 1 || num_agents > MG_MAX_AGENTS) return MG_ERR_BAD_ARG;
    if (num_agents > (size - 2) * (size - 2)) return MG_ERR_BAD_ARG;  // more agents than cells in the room
    if (num_layouts == 0) return 0;
    if (!rng_state || !rng_inc || !cells || !agents) return MG_ERR_BAD_ARG;
    mg::gen_layouts_red_blue_doors_kernel<<<(unsigned)((num_layouts + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        size, num_agents, num_layouts, rng_state, rng_inc, rng_buf, cells, agents, status);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int mg_gen_layouts_locked_hallway(int32_t num_rooms, int32_t room_size, int32_t max_hallway_keys,
                                  int32_t max_keys_per_room, int32_t num_agents, int64_t num_layouts,
                                  uint64_t *rng_state, const uint64_t *rng_inc, uint64_t *rng_buf, uint32_t *cells,
                                  int8_t *agents, int32_t *status, void *stream) {
    if (num_rooms < 2 || num_rooms > 6 || (num_rooms & 1) || room_size < 4 || room_size > 40 || max_hallway_keys < 1 ||
        max_keys_per_room < 1 || num_layouts < 0 || num_agents < 1 || num_agents > MG_MAX_AGENTS) return MG_ERR_BAD_ARG;
    if (num_layouts == 0) return 0;
    if (!rng_state || !rng_inc || !cells || !agents) return MG_ERR_BAD_ARG;
    mg::gen_layouts_locked_hallway_kernel<<<(unsigned)((num_layouts + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        num_rooms, room_size, max_hallway_keys, max_keys_per_room, num_agents, num_layouts, rng_state, rng_inc,
        rng_buf, cells, agents, status);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int mg_gen_layouts_playground(int32_t room_size, int32_t num_rows, int32_t num_cols, int32_t num_agents,
                              int64_t num_layouts, uint64_t *rng_state, const uint64_t *rng_inc, uint64_t *rng_buf,
                              uint64_t *order_state, const uint64_t *order_inc, uint64_t *order_buf, uint32_t *cells,
                              int8_t *agents, int32_t *status, void *stream) {
    if (room_size < 4 || num_rows < 1 || num_cols < 1 || num_rows * num_cols > 16 || num_layouts < 0 ||
        num_cols * (room_size - 1) + 1 > 127 || num_rows * (room_size - 1) + 1 > 127 || num_agents < 1 ||
        num_agents > MG_MAX_AGENTS) return MG_ERR_BAD_ARG;
    if (num_layouts == 0) return 0;
    if (!rng_state || !rng_inc || !order_state || !order_inc || !cells || !agents) return MG_ERR_BAD_ARG;
    mg::gen_layouts_playground_kernel<<<(unsigned)((num_layouts + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        room_size, num_rows, num_cols, num_agents, num_layouts, rng_state, rng_inc, rng_buf, order_state, order_inc,
        order_buf, cells, agents, status);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int mg_gen_layouts_bup(int32_t room_size, int32_t num_agents, int64_t num_layouts, uint64_t *rng_state,
                       const uint64_t *rng_inc, uint64_t *rng_buf, uint64_t *order_state, const uint64_t *order_inc,
                       uint64_t *order_buf, uint32_t *cells, int8_t *agents, int32_t *info, int32_t *status,
                       void *stream) {
    if (room_size < 4 || room_size > 60 || num_layouts < 0 || num_agents < 1 || num_agents > MG_MAX_AGENTS)
        return MG_ERR_BAD_ARG;
    if (num_layouts == 0) return 0;
    if (!rng_state || !rng_inc || !order_state || !order_inc || !cells || !agents) return MG_ERR_BAD_ARG;
    mg::gen_layouts_bup_kernel<<<(unsigned)((num_layouts + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        room_size, num_agents, num_layouts, rng_state, rng_inc, rng_buf, order_state, order_inc, order_buf, cells,
        agents, info, status);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int mg_obs_features(int32_t view_size, int64_t num_agents_total, int32_t obs_agent_stride, const int8_t *obs,
                    const int8_t *direction, int32_t direction_stride, const float *dir_lut, float *out, void *stream) {
    if (view_size < 3 || view_size > MG_MAX_VIEW || num_agents_total < 0 || direction_stride < 1 ||
        obs_agent_stride < 3 * view_size * view_size) return MG_ERR_BAD_ARG;
    if (num_agents_total == 0) return 0;
    if (!obs || !direction || !dir_lut || !out) return MG_ERR_BAD_ARG;
    // 16 agents x V*V*23 floats is a multiple of 16 bytes, so every block's 16-byte stores are aligned
    if (reinterpret_cast<uintptr_t>(out) & 15u) return MG_ERR_ALIGNMENT;
    if (env_int("MG_FEATURES_DIRECT", 0, any_mg_knob())) {  // knob: the first version (one 16-byte store per thread)
        mg::obs_features_kernel<<<(unsigned)((num_agents_total + 15) / 16), 256, 0, (cudaStream_t)stream>>>(
            view_size, num_agents_total, obs_agent_stride, mg::rcp32(view_size * view_size * 23), obs, direction,
            direction_stride, dir_lut, (float4 *)out);
    } else {  // 32 agents per block: 32 * V*V * 23 floats is a multiple of 16 bytes as well
        mg::obs_features_tile_kernel<<<(unsigned)((num_agents_total + 31) / 32), 256, 0, (cudaStream_t)stream>>>(
            view_size, num_agents_total, obs_agent_stride, mg::rcp32(view_size * view_size), obs, direction,
            direction_stride, dir_lut, out);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int mg_unpack_grid(int32_t width, int32_t height, int64_t num_envs, const uint32_t *cells, int8_t *grid3,
                   void *stream) {
    if (width < 1 || height < 1 || width > 127 || height > 127 || num_envs < 0) return MG_ERR_BAD_ARG;
    if (num_envs == 0) return 0;
    if (!grid3 || !cells) return MG_ERR_BAD_ARG;
    const int64_t total = num_envs * width * height;
    mg::unpack_grid_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(width, height, total, cells, grid3);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int64_t mg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int32_t mg_static_obs_stride(int32_t obs_agent_stride) { return mg::static_obs_stride(obs_agent_stride); }

int64_t mg_static_obs_bytes(int32_t width, int32_t height, int32_t obs_agent_stride) {
    return mg::static_table_bytes(width, height, obs_agent_stride);
}

int mg_build_static_obs(const MgConfig *cfg, const uint32_t *layout_cells, int8_t *static_obs, void *stream) {
    int rc = validate(cfg, 1);
    if (rc) return rc;
    if (!layout_cells || !static_obs) return MG_ERR_BAD_ARG;
    if (!aligned16(static_obs)) return MG_ERR_ALIGNMENT;
    mg::Params p;
    fill_config(p, cfg, 1);
    p.G = 16;
    mg::carve_static(p, false, 1);  // derived geometry (Hp, cstride)
    const int entries = p.W * p.H * 4;
    mg::static_build_kernel<<<(unsigned)((entries + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        p, layout_cells, reinterpret_cast<uint8_t *>(static_obs));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

void mg_debug_set_trace(void *device_buffer) { g_trace.store((unsigned long long *)device_buffer, std::memory_order_relaxed); }

int mg_gen_obs(const MgConfig *cfg, int64_t num_envs, const uint32_t *grid, const int8_t *agents,
               int8_t *obs, void *stream) {
    int rc = validate(cfg, num_envs);
    if (rc) return rc;
    if (num_envs == 0) return 0;
    if (!grid || !agents || !obs) return MG_ERR_BAD_ARG;
    if (!aligned16(grid) || !aligned16(agents) || !aligned16(obs)) return MG_ERR_ALIGNMENT;
    mg::Params p;
    fill_config(p, cfg, num_envs);
    p.flags &= ~MG_FLAG_AUTO_RESET;
    p.grid = const_cast<uint32_t *>(grid);    // MODE_OBS never writes state
    p.agents = const_cast<int8_t *>(agents);
    p.obs = obs;
    if ((rc = plan(p))) return rc;
    return dispatch<mg::MODE_OBS>(p, (cudaStream_t)stream);
}

int mg_step(const MgConfig *cfg, int64_t num_envs, const MgState *state, const int8_t *actions,
            const MgStepOut *out, void *stream) {
    return step_common<mg::MODE_STEP>(cfg, num_envs, state, actions, out, stream);
}

int mg_step_obs(const MgConfig *cfg, int64_t num_envs, const MgState *state, const int8_t *actions,
                const MgStepOut *out, void *stream) {
    return step_common<mg::MODE_STEP_OBS>(cfg, num_envs, state, actions, out, stream);
}

int mg_rollout(const MgConfig *cfg, int64_t num_envs, int32_t num_steps, const MgState *state,
               const int8_t *actions, const MgRolloutOut *out, void *stream) {
    if (!out) return MG_ERR_BAD_ARG;
    const MgStepOut so = {out->obs, out->reward, out->terminated, out->truncated, out->status, nullptr};
    return step_common<mg::MODE_STEP_OBS, true>(cfg, num_envs, state, actions, &so, stream, num_steps,
                                                out->direction);
}

int mg_reset_where(const MgConfig *cfg, int64_t num_envs, const MgState *state, const uint8_t *mask, void *stream) {
    int rc = validate(cfg, num_envs);
    if (rc) return rc;
    if (num_envs == 0) return 0;
    if (!mask || !state || !state->layout_idx || !state->pool_grid || !state->pool_agents || cfg->num_layouts < 1 ||
        cfg->layout_stride < 0) return MG_ERR_BAD_ARG;
    mg::Params p;
    fill_config(p, cfg, num_envs);
    if ((rc = fill_state(p, state))) return rc;
    p.G = 16;
    mg::carve_smem(p);  // derived geometry (cstride)
    const int64_t threads = num_envs * 32;
    mg::reset_where_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, mask);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int mg_refresh_done_layouts(const MgConfig *cfg, int64_t num_envs, const MgState *state, const MgLayoutGen *gen,
                            int32_t *status, void *stream) {
    int rc = validate(cfg, num_envs);
    if (rc) return rc;
    if (num_envs == 0) return 0;
    if (!state || !gen || !gen->rng_state || !gen->rng_inc || !state->pool_grid || !state->pool_agents ||
        !state->agents || !state->step_count) return MG_ERR_BAD_ARG;
    if (gen->family < mg::LAYOUT_EMPTY_RANDOM || gen->family > mg::LAYOUT_PLAYGROUND) return MG_ERR_BAD_ARG;
    if (cfg->num_layouts != num_envs) return MG_ERR_BAD_ARG;  // one pool slot per env
    if ((gen->family == mg::LAYOUT_BUP || gen->family == mg::LAYOUT_PLAYGROUND) && (!state->pcg_state || !state->pcg_inc))
        return MG_ERR_BAD_ARG;
    mg::Params p;
    fill_config(p, cfg, num_envs);
    if ((rc = fill_state(p, state))) return rc;
    p.status = status;
    p.G = 16;
    mg::carve_smem(p);  // derived geometry (cstride)
    mg::LayoutGen lg;
    lg.family = gen->family; lg.a = gen->params[0]; lg.b = gen->params[1]; lg.c = gen->params[2]; lg.d = gen->params[3];
    lg.rng_state = gen->rng_state; lg.rng_inc = gen->rng_inc; lg.rng_buf = gen->rng_buf; lg.order_buf = gen->order_buf; lg.info = gen->info;
    mg::refresh_done_kernel<<<(unsigned)((num_envs + 63) / 64), 64, 0, (cudaStream_t)stream>>>(p, lg);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int mg_step_obs_host(const MgConfig *cfg, int64_t num_envs, const MgState *state,
                     const int8_t *h_actions, int8_t *d_actions, const MgStepOut *d_out,
                     const MgStepOut *h_out, void *stream) {
    int rc = validate(cfg, num_envs);
    if (rc) return rc;
    if (!h_actions || !d_actions || !d_out || !h_out || !h_out->obs || !h_out->reward ||
        !h_out->terminated || !h_out->truncated) return MG_ERR_BAD_ARG;
    if (num_envs == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t E = (size_t)num_envs, n = (size_t)cfg->num_agents;
    cudaError_t err = cudaMemcpyAsync(d_actions, h_actions, E * n, cudaMemcpyHostToDevice, s);
    if (err != cudaSuccess) return (int)err;
    rc = step_common<mg::MODE_STEP_OBS>(cfg, num_envs, state, d_actions, d_out, stream);
    if (rc) return rc;
    if ((err = cudaMemcpyAsync(h_out->obs, d_out->obs, E * n * cfg->obs_agent_stride, cudaMemcpyDeviceToHost, s))) return (int)err;
    if ((err = cudaMemcpyAsync(h_out->reward, d_out->reward, E * n * sizeof(double), cudaMemcpyDeviceToHost, s))) return (int)err;
    if ((err = cudaMemcpyAsync(h_out->terminated, d_out->terminated, E * n, cudaMemcpyDeviceToHost, s))) return (int)err;
    if ((err = cudaMemcpyAsync(h_out->truncated, d_out->truncated, E, cudaMemcpyDeviceToHost, s))) return (int)err;
    return 0;
}

int32_t mg_packed_obs_stride(int32_t view_size) { return mg::packed_obs_stride(view_size); }

int mg_pack_obs(int32_t view_size, int64_t num_agents_total, int32_t obs_agent_stride, const int8_t *obs,
                uint8_t *packed, void *stream) {
    if (view_size < 3 || view_size > MG_MAX_VIEW || num_agents_total < 0 || (obs_agent_stride & 3) ||
        obs_agent_stride < 3 * view_size * view_size) return MG_ERR_BAD_ARG;
    if (num_agents_total == 0) return 0;
    if (!obs || !packed) return MG_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(packed) & 7u) || (reinterpret_cast<uintptr_t>(obs) & 15u)) return MG_ERR_ALIGNMENT;
    const unsigned blocks = (unsigned)((num_agents_total + 127) / 128);
    cudaStream_t s = (cudaStream_t)stream;
    if (view_size == 7) mg::pack_obs_kernel<7><<<blocks, 128, 0, s>>>(7, num_agents_total, obs_agent_stride, obs, packed);
    else if (view_size == 9) mg::pack_obs_kernel<9><<<blocks, 128, 0, s>>>(9, num_agents_total, obs_agent_stride, obs, packed);
    else mg::pack_obs_kernel<0><<<blocks, 128, 0, s>>>(view_size, num_agents_total, obs_agent_stride, obs, packed);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int mg_step_obs_host_packed(const MgConfig *cfg, int64_t num_envs, const MgState *state, const int8_t *h_actions,
                            int8_t *d_actions, const MgStepOut *d_out, uint8_t *d_packed, const MgStepOut *h_out,
                            void *stream) {
    int rc = validate(cfg, num_envs);
    if (rc) return rc;
    if (!h_actions || !d_actions || !d_out || !d_packed || !h_out || !h_out->obs || !h_out->reward ||
        !h_out->terminated || !h_out->truncated) return MG_ERR_BAD_ARG;
    if (num_envs == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t E = (size_t)num_envs, n = (size_t)cfg->num_agents;
    cudaError_t err = cudaMemcpyAsync(d_actions, h_actions, E * n, cudaMemcpyHostToDevice, s);
    if (err != cudaSuccess) return (int)err;
    rc = step_common<mg::MODE_STEP_OBS>(cfg, num_envs, state, d_actions, d_out, stream);
    if (rc) return rc;
    if ((rc = mg_pack_obs(cfg->view_size, (int64_t)(E * n), cfg->obs_agent_stride, d_out->obs, d_packed, stream))) return rc;
    const size_t pbytes = E * n * (size_t)mg::packed_obs_stride(cfg->view_size);
    if ((err = cudaMemcpyAsync(h_out->obs, d_packed, pbytes, cudaMemcpyDeviceToHost, s))) return (int)err;
    if ((err = cudaMemcpyAsync(h_out->reward, d_out->reward, E * n * sizeof(double), cudaMemcpyDeviceToHost, s))) return (int)err;
    if ((err = cudaMemcpyAsync(h_out->terminated, d_out->terminated, E * n, cudaMemcpyDeviceToHost, s))) return (int)err;
    if ((err = cudaMemcpyAsync(h_out->truncated, d_out->truncated, E, cudaMemcpyDeviceToHost, s))) return (int)err;
    return 0;
}

int32_t mg_packed_obs_stride_bits(int32_t view_size, int32_t bits) { return mg::packed_obs_stride_bits(view_size, bits); }

int mg_pack_obs_palette(int32_t view_size, int64_t num_agents_total, int32_t obs_agent_stride, const int8_t *obs,
                        int32_t bits, const uint8_t *lut, uint8_t *packed, int32_t *status, void *stream) {
    if (view_size < 3 || view_size > MG_MAX_VIEW || num_agents_total < 0 || (obs_agent_stride & 3) ||
        obs_agent_stride < 3 * view_size * view_size || bits < 1 || bits > 8) return MG_ERR_BAD_ARG;
    if (num_agents_total == 0) return 0;
    if (!obs || !packed || !lut) return MG_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(packed) & 7u) || (reinterpret_cast<uintptr_t>(obs) & 15u) ||
        (reinterpret_cast<uintptr_t>(lut) & 3u)) return MG_ERR_ALIGNMENT;
    const unsigned blocks = (unsigned)((num_agents_total + 127) / 128);
    cudaStream_t s = (cudaStream_t)stream;
    if (view_size == 7) mg::pack_obs_palette_kernel<7><<<blocks, 128, 0, s>>>(7, num_agents_total, obs_agent_stride, obs, bits, lut, packed, status);
    else if (view_size == 9) mg::pack_obs_palette_kernel<9><<<blocks, 128, 0, s>>>(9, num_agents_total, obs_agent_stride, obs, bits, lut, packed, status);
    else mg::pack_obs_palette_kernel<0><<<blocks, 128, 0, s>>>(view_size, num_agents_total, obs_agent_stride, obs, bits, lut, packed, status);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int mg_step_obs_host_palette(const MgConfig *cfg, int64_t num_envs, const MgState *state, const int8_t *h_actions,
                             int8_t *d_actions, const MgStepOut *d_out, uint8_t *d_packed, int32_t bits,
                             const uint8_t *lut, const MgStepOut *h_out, void *stream) {
    int rc = validate(cfg, num_envs);
    if (rc) return rc;
    if (!h_actions || !d_actions || !d_out || !d_packed || !lut || !h_out || !h_out->obs || !h_out->reward ||
        !h_out->terminated || !h_out->truncated) return MG_ERR_BAD_ARG;
    if (num_envs == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t E = (size_t)num_envs, n = (size_t)cfg->num_agents;
    cudaError_t err = cudaMemcpyAsync(d_actions, h_actions, E * n, cudaMemcpyHostToDevice, s);
    if (err != cudaSuccess) return (int)err;
    rc = step_common<mg::MODE_STEP_OBS>(cfg, num_envs, state, d_actions, d_out, stream);
    if (rc) return rc;
    if ((rc = mg_pack_obs_palette(cfg->view_size, (int64_t)(E * n), cfg->obs_agent_stride, d_out->obs, bits, lut,
                                  d_packed, d_out->status, stream))) return rc;
    // the small arrays first: they are ready as soon as the step kernel is, the observations follow
    if ((err = cudaMemcpyAsync(h_out->reward, d_out->reward, E * n * sizeof(double), cudaMemcpyDeviceToHost, s))) return (int)err;
    if ((err = cudaMemcpyAsync(h_out->terminated, d_out->terminated, E * n, cudaMemcpyDeviceToHost, s))) return (int)err;
    if ((err = cudaMemcpyAsync(h_out->truncated, d_out->truncated, E, cudaMemcpyDeviceToHost, s))) return (int)err;
    const size_t pbytes = E * n * (size_t)mg::packed_obs_stride_bits(cfg->view_size, bits);
    if ((err = cudaMemcpyAsync(h_out->obs, d_packed, pbytes, cudaMemcpyDeviceToHost, s))) return (int)err;
    return 0;
}

int32_t mg_wire_record_bytes(int32_t num_agents) { return mg::wire_record_bytes(num_agents); }

int64_t mg_wire_obs_bytes(int32_t view_size, int32_t bits, int32_t num_agents, int64_t num_envs) {
    const int64_t b = num_envs * num_agents * (int64_t)mg::packed_obs_stride_bits(view_size, bits);
    return (b + 15) & ~(int64_t)15;
}

int64_t mg_wire_bytes(int32_t view_size, int32_t bits, int32_t num_agents, int64_t num_envs) {
    return mg_wire_obs_bytes(view_size, bits, num_agents, num_envs) + num_envs * (int64_t)mg::wire_record_bytes(num_agents);
}

}  // extern "C"

namespace {
// Side stream + events of the chunked host wire (one set per device, created on first use).
struct WirePipe { cudaStream_t side = nullptr; cudaEvent_t ready[8] = {}, done = nullptr; };
WirePipe *wire_pipe() {
    static WirePipe pipes[64];
    static std::atomic<int> made[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    WirePipe &w = pipes[dev];
    int expect = 0;
    if (made[dev].load(std::memory_order_acquire) == 2) return &w;
    if (made[dev].compare_exchange_strong(expect, 1)) {
        bool ok = cudaStreamCreateWithFlags(&w.side, cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; i < 8 && ok; i++) ok = cudaEventCreateWithFlags(&w.ready[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming) == cudaSuccess;
        made[dev].store(ok ? 2 : 3, std::memory_order_release);
        return ok ? &w : nullptr;
    }
    while (made[dev].load(std::memory_order_acquire) == 1) {}
    return made[dev].load(std::memory_order_acquire) == 2 ? &w : nullptr;
}
}  // namespace

extern "C" {

int mg_step_obs_host_wire(const MgConfig *cfg, int64_t num_envs, const MgState *state, const int8_t *h_actions,
                          int8_t *d_actions, const MgStepOut *d_out, uint8_t *d_wire, int32_t bits, const uint8_t *lut,
                          uint8_t *h_wire, void *stream) {
    int rc = validate(cfg, num_envs);
    if (rc) return rc;
    if (!h_actions || !d_actions || !d_out || !d_wire || !lut || !h_wire || !d_out->status || !state) return MG_ERR_BAD_ARG;
    if (cfg->num_agents > 31) return MG_ERR_BAD_ARG;  // (terminated mask + the truncated bit share one word)
    if (reinterpret_cast<uintptr_t>(d_wire) & 15u) return MG_ERR_ALIGNMENT;
    if (num_envs == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t E = (size_t)num_envs, n = (size_t)cfg->num_agents;
    cudaError_t err = cudaMemcpyAsync(d_actions, h_actions, E * n, cudaMemcpyHostToDevice, s);
    if (err != cudaSuccess) return (int)err;
    const size_t ps = (size_t)mg::packed_obs_stride_bits(cfg->view_size, bits), rb = (size_t)mg::wire_record_bytes((int)n);
    const size_t obs_bytes = (size_t)mg_wire_obs_bytes(cfg->view_size, bits, cfg->num_agents, num_envs);
    // Chunked pipeline: the envs are stepped and packed in C slices on the caller's stream while a side stream copies
    // the finished slices to the host, so only the first slice's kernels are not hidden behind the PCIe transfer.
    // (Slices of a multiple of 256 envs keep every per-env array 16-byte aligned. Big batches only; MG_WIRE_CHUNKS=1
    // turns it off.)
    const bool k = any_mg_knob();
    int C = env_int("MG_WIRE_CHUNKS", num_envs >= 32768 ? 4 : 1, k);
    C = C < 1 ? 1 : (C > 8 ? 8 : C);
    WirePipe *pipe = C > 1 ? wire_pipe() : nullptr;
    if (!pipe) C = 1;
    const size_t chunk = C == 1 ? E : (((E + C - 1) / C + 255) & ~(size_t)255);
    const size_t cs = (size_t)mg::cells_per_env(cfg->width, cfg->height);
    int used = 0;
    for (size_t e0 = 0; e0 < E; e0 += chunk, used++) {
        const size_t ne = E - e0 < chunk ? E - e0 : chunk;
        MgState st = *state;
        MgStepOut out = *d_out;
        st.grid = state->grid + e0 * cs; st.agents = state->agents + e0 * n * 8; st.step_count = state->step_count + e0;
        if (state->pcg_state) st.pcg_state = state->pcg_state + 2 * e0;
        if (state->pcg_inc) st.pcg_inc = state->pcg_inc + 2 * e0;
        if (state->layout_idx) st.layout_idx = state->layout_idx + e0;
        if (state->hook_state) st.hook_state = state->hook_state + e0;
        if (state->chain) st.chain = state->chain + 4 * e0;
        out.obs = d_out->obs + e0 * n * (size_t)cfg->obs_agent_stride;
        out.reward = d_out->reward + e0 * n; out.terminated = d_out->terminated + e0 * n; out.truncated = d_out->truncated + e0;
        if (d_out->one_hot) out.one_hot = d_out->one_hot + e0 * n * (size_t)cfg->view_size * cfg->view_size * 21;
        if ((rc = step_common<mg::MODE_STEP_OBS>(cfg, (int64_t)ne, &st, d_actions + e0 * n, &out, stream))) return rc;
        if ((rc = mg_pack_obs_palette(cfg->view_size, (int64_t)(ne * n), cfg->obs_agent_stride, out.obs, bits, lut,
                                      d_wire + e0 * n * ps, d_out->status, stream))) return rc;
        mg::wire_env_records_kernel<<<(unsigned)((ne + 127) / 128), 128, 0, s>>>(
            (int)n, (int64_t)ne, out.reward, out.terminated, out.truncated, d_wire + obs_bytes + e0 * rb, d_out->status);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if ((err = cudaGetLastError()) != cudaSuccess) return (int)err;
        if (C == 1) break;
        // this slice's observations go out while the next slice is computed
        if ((err = cudaEventRecord(pipe->ready[used], s))) return (int)err;
        if ((err = cudaStreamWaitEvent(pipe->side, pipe->ready[used], 0))) return (int)err;
        if ((err = cudaMemcpyAsync(h_wire + e0 * n * ps, d_wire + e0 * n * ps, ne * n * ps, cudaMemcpyDeviceToHost, pipe->side)))
            return (int)err;
    }
    if (C == 1) {  // ONE copy
        const size_t bytes = (size_t)mg_wire_bytes(cfg->view_size, bits, cfg->num_agents, num_envs);
        if ((err = cudaMemcpyAsync(h_wire, d_wire, bytes, cudaMemcpyDeviceToHost, s))) return (int)err;
        return 0;
    }
    // the env records of all slices in one copy behind the last slice, then the caller's stream waits for the side stream
    if ((err = cudaMemcpyAsync(h_wire + obs_bytes, d_wire + obs_bytes, E * rb, cudaMemcpyDeviceToHost, pipe->side))) return (int)err;
    if ((err = cudaEventRecord(pipe->done, pipe->side))) return (int)err;
    if ((err = cudaStreamWaitEvent(s, pipe->done, 0))) return (int)err;
    return 0;
}

struct MgStepPlan { LaunchRecord rec; };

int mg_step_plan_create(const MgConfig *cfg, int64_t num_envs, const MgState *state, const int8_t *actions,
                        const MgStepOut *out, MgStepPlan **plan_out) {
    if (!plan_out) return MG_ERR_BAD_ARG;
    *plan_out = nullptr;
    if (num_envs <= 0) return MG_ERR_BAD_ARG;
    MgStepPlan *plan = new MgStepPlan();
    plan->rec.func = nullptr;
    g_capture = &plan->rec;
    const int rc = step_common<mg::MODE_STEP_OBS>(cfg, num_envs, state, actions, out, nullptr);
    g_capture = nullptr;
    if (rc || !plan->rec.func) {
        delete plan;
        return rc ? rc : MG_ERR_BAD_ARG;
    }
    *plan_out = plan;
    return 0;
}

int mg_step_plan_run(MgStepPlan *plan, const int8_t *actions, void *stream) {
    if (!plan || !actions) return MG_ERR_BAD_ARG;
    if (reinterpret_cast<uintptr_t>(actions) & 15u) return MG_ERR_ALIGNMENT;  // (the plan assumed 16-byte alignment)
    plan->rec.p.actions = actions;
    plan->rec.p.trace = g_trace.load(std::memory_order_relaxed);
    return launch_record(plan->rec, (cudaStream_t)stream);
}

void mg_step_plan_destroy(MgStepPlan *plan) { delete plan; }

}  // extern "C"
