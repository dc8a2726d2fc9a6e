// TEST INFRASTRUCTURE: runs the phase functions of multigrid_b200/csrc/mg_kernels.cuh on the CPU,
// lane by lane and phase by phase (a phase boundary == __syncwarp), so the kernel logic can be
// checked against the oracle in a container without a GPU. The TMA bulk copies of the CUDA build
// are replaced by the kernel's own plain-copy path. Never loaded by the product.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../multigrid_b200/csrc/mg_kernels.cuh"
#include "../../multigrid_b200/csrc/mg_static.cuh"

namespace {

// One obs pass: every lane gathers into its registers first, then every lane packs (the stage may
// alias cells the pass has just read: see carve_smem).
template <int VT>
void obs_pass(const mg::Params &p, const mg::Group &g, int pass) {
    static uint32_t cr[mg::LANES][VT ? VT * VT : 1];
    uint8_t *stage = mg::stage_of(p, g, pass);
    mg::ObsTask t[mg::LANES];
    for (int l = 0; l < mg::LANES; l++) t[l] = mg::obs_task(p, g, pass, l);
    if constexpr (VT != 0) {
        for (int l = 0; l < mg::LANES; l++)
            if (t[l].valid) mg::obs_compute<VT>(p, t[l].cells, t[l].a0, t[l].a1, cr[l]);
        for (int l = 0; l < mg::LANES; l++)
            if (t[l].valid) mg::obs_pack_store<VT>(p, cr[l], stage + l * p.ostride);
    } else {
        for (int l = 0; l < mg::LANES; l++)
            if (t[l].valid) mg::obs_agent_generic(p, t[l].cells, t[l].a0, t[l].a1, stage + l * p.ostride);
    }
}

template <int VT, int MODE>
void run_groups(const mg::Params &p) {
    std::vector<uint8_t> smem_store(p.warp_bytes + 128);
    uint8_t *ws = smem_store.data();
    ws += (128 - (reinterpret_cast<uintptr_t>(ws) & 127)) & 127;
    const int groups = (p.num_envs + p.G - 1) / p.G;
    const int L = mg::LANES;
    for (int grp = 0; grp < groups; grp++) {
        std::memset(ws, 0xCD, p.warp_bytes);  // poison: catches reads of unwritten smem
        const mg::Group g = mg::group_view(p, ws, grp);
        mg::EnvRegs er[mg::LANES];
        int env[mg::LANES];
        for (int l = 0; l < L; l++) env[l] = mg::lane_env(p, g, l);
        for (int t = 0; t < p.T; t++) {  // same loop as the kernel's (T > 1: mg_rollout)
            const size_t tE = (size_t)t * (size_t)p.num_envs;
            for (int l = 0; l < L; l++) mg::phase_load_plain<MODE>(p, g, l, t);
            if (t == 0)
                for (int l = 0; l < L; l++) mg::env_load<MODE>(p, g, env[l], er[l]);
            mg::OrderDraw draw[mg::LANES];
            for (int l = 0; l < L; l++) draw[l] = mg::phase_draw<MODE>(p, g, env[l], er[l]);
            if (MODE != mg::MODE_OBS && (p.flags & MG_FLAG_AUTO_RESET)) {
                for (int l = 0; l < L; l++) mg::phase_reset(p, g, env[l], er[l]);
                const uint32_t pending = mg::reset_mask_host(g);
                for (int l = 0; l < L; l++) mg::phase_reset_grid(p, g, pending, l);
            }
            for (int l = 0; l < L; l++) mg::phase_step<MODE>(p, g, env[l], er[l], draw[l], tE);
            if (MODE != mg::MODE_STEP) {
                const int passes = mg::obs_passes(p, g);
                for (int pass = 0; pass < passes; pass++) {
                    obs_pass<VT>(p, g, pass);
                    for (int l = 0; l < L; l++) mg::phase_obs_store_plain(p, g, pass, l, tE);
                    if (p.one_hot && p.T == 1)
                        for (int l = 0; l < L; l++) mg::phase_one_hot(p, g, pass, l);
                }
            }
        }
        if (MODE != mg::MODE_OBS)
            for (int l = 0; l < L; l++) mg::phase_store_plain(p, g, l);
    }
}

// The static-grid kernels (mg_static.cuh), lane by lane: same phase order as static_rolled_kernel /
// static_fast_kernel (whose cooperative table copy is a plain per-entry copy here).
void run_groups_static_rolled(const mg::Params &p) {
    std::vector<uint8_t> smem_store(p.warp_bytes + 128);
    uint8_t *ws = smem_store.data();
    ws += (128 - (reinterpret_cast<uintptr_t>(ws) & 127)) & 127;
    const int groups = (p.num_envs + p.G - 1) / p.G;
    const int L = mg::LANES;
    for (int grp = 0; grp < groups; grp++) {
        std::memset(ws, 0xCD, p.warp_bytes);
        const mg::Group g = mg::group_view(p, ws, grp);
        mg::EnvRegs er[mg::LANES];
        mg::OrderDraw draw[mg::LANES];
        int env[mg::LANES];
        for (int l = 0; l < L; l++) env[l] = l < g.ne ? l : -1;
        for (int l = 0; l < L; l++) mg::static_load(p, g, l);
        for (int l = 0; l < L; l++) mg::static_env_load(p, g, env[l], er[l]);
        for (int l = 0; l < L; l++) draw[l] = mg::phase_draw<mg::MODE_STEP_OBS>(p, g, env[l], er[l]);
        for (int l = 0; l < L; l++) mg::static_env_step(p, g, env[l], er[l], draw[l]);
        for (int l = 0; l < L; l++) mg::static_store_agents(p, g, l);
        const int n = p.n, tasks = g.ne * n, passes = (tasks + L - 1) / L;
        for (int pass = 0; pass < passes; pass++) {
            for (int l = 0; l < L; l++) mg::static_obs_agent(p, g, pass, l, g.stage);
            const int left = tasks - pass * L, cnt = left < L ? left : L;
            for (int l = 0; l < L; l++)
                mg::warp_copy(p.obs + ((size_t)g.e0 * n + (size_t)pass * L) * p.ostride, g.stage, cnt * p.ostride, l);
            if (p.one_hot)
                for (int l = 0; l < L; l++)
                    mg::one_hot_emit_cold(p.one_hot, (uint32_t)(p.V * p.V), p.rcp_vv, g.stage, p.ostride,
                                          (size_t)g.e0 * n + (size_t)pass * L, cnt, l);
        }
    }
}

template <int VT, int NT>
void run_groups_static_fast(const mg::Params &p) {
    typedef mg::StaticCopy<VT> C;
    std::vector<uint8_t> smem_store(p.warp_bytes + 128);
    uint8_t *ws = smem_store.data();
    ws += (128 - (reinterpret_cast<uintptr_t>(ws) & 127)) & 127;
    const int groups = (p.num_envs + p.G - 1) / p.G;
    const int L = mg::LANES;
    uint32_t colors = 0;
    for (int j = 0; j < NT; j++) colors |= (uint32_t)(uint8_t)p.pool_agents[j * 8 + 7] << (4 * j);
    for (int grp = 0; grp < groups; grp++) {
        std::memset(ws, 0xCD, p.warp_bytes);
        const mg::Group g = mg::group_view(p, ws, grp);
        uint32_t *a0T = g.ag;
        for (int l = 0; l < L; l++) mg::static_fast_env<NT>(p, g, l, a0T, p.static_move);
        const int tasks = g.ne * NT, passes = (tasks + L - 1) / L;
        for (int pass = 0; pass < passes; pass++) {
            uint8_t *stage = g.stage;
            const int left = tasks - pass * L, cnt = left < L ? left : L;
            uint32_t a0[mg::LANES], ent[mg::LANES];
            for (int l = 0; l < L; l++) {
                const int id = pass * L + l, el = mg::static_task_env<NT>(id), k = id & (NT - 1);
                a0[l] = id < tasks ? a0T[k * L + el] : 0u;
                ent[l] = mg::static_entry_offset(p, a0[l]);
            }
            for (int A = 0; A < cnt; A++) std::memcpy(stage + A * C::OS, p.static_obs + ent[A], C::OS);
            for (int l = 0; l < cnt; l++) {
                const int id = pass * L + l, el = mg::static_task_env<NT>(id), k = id & (NT - 1);
                mg::static_overlay_fast<VT, NT>(a0[l], k, a0T + el, colors, stage + l * C::OS);
            }
            for (int l = 0; l < L; l++)
                mg::warp_copy(p.obs + ((size_t)g.e0 * NT + (size_t)pass * L) * C::OS, stage, cnt * C::OS, l);
            if (p.one_hot)
                for (int l = 0; l < L; l++)
                    mg::one_hot_emit_cold(p.one_hot, (uint32_t)(VT * VT), p.rcp_vv, stage, C::OS,
                                          (size_t)g.e0 * NT + (size_t)pass * L, cnt, l);
        }
    }
}

template <int MODE>
void dispatch(const mg::Params &p, int generic) {
    if (MODE == mg::MODE_STEP) return run_groups<0, MODE>(p);
    if (!generic) {
        switch (p.V) {
            case 3: return run_groups<3, MODE>(p);
            case 5: return run_groups<5, MODE>(p);
            case 7: return run_groups<7, MODE>(p);
            case 9: return run_groups<9, MODE>(p);
            default: break;
        }
    }
    run_groups<0, MODE>(p);
}

}  // namespace

extern "C" int sim_run(int mode, const MgConfig *c, int64_t num_envs, const MgState *s,
                       const int8_t *actions, const MgStepOut *o, int forced_group, int generic,
                       int num_steps, int8_t *direction) {
    mg::Params p;
    std::memset(&p, 0, sizeof(p));
    p.W = c->width; p.H = c->height; p.n = c->num_agents; p.V = c->view_size;
    p.max_steps = c->max_steps; p.flags = c->flags; p.hook = c->hook; p.hook_param = c->hook_param;
    p.ostride = c->obs_agent_stride; p.K = c->num_layouts; p.lstride = c->layout_stride;
    p.num_envs = (int32_t)num_envs;
    p.generic_view = generic & 1;
    p.T = num_steps; p.direction = direction;
    if (mode == mg::MODE_OBS) p.flags &= ~MG_FLAG_AUTO_RESET;
    p.grid = s->grid; p.agents = s->agents; p.step_count = s->step_count;
    p.pcg_state = s->pcg_state; p.pcg_inc = s->pcg_inc; p.layout_idx = s->layout_idx;
    p.pool_grid = s->pool_grid; p.pool_agents = s->pool_agents; p.hook_state = s->hook_state;
    p.actions = actions;
    p.obs = o->obs; p.reward = o->reward; p.terminated = o->terminated; p.truncated = o->truncated;
    p.status = o->status;
    p.one_hot = mode == mg::MODE_STEP_OBS && num_steps == 1 ? o->one_hot : nullptr;
    p.rcp_vv = mg::rcp32(p.V * p.V);
    if (mode == mg::MODE_STEP_OBS && num_steps == 1 && (p.flags & MG_FLAG_STATIC_GRID)) {
        if (!s->static_obs || p.hook != MG_HOOK_NONE) return MG_ERR_BAD_ARG;
        p.static_obs = reinterpret_cast<const uint8_t *>(s->static_obs);
        p.static_stride = mg::static_obs_stride(p.ostride);
        p.static_move = reinterpret_cast<const uint32_t *>(p.static_obs + (size_t)p.W * p.H * 4 * p.static_stride);
        int rc = mg::plan_static(p, forced_group, 0, 227 * 1024);
        if (rc) return rc;
        if (!mg::static_fast_shape(p)) run_groups_static_rolled(p);
        else if (p.V == 7 && p.n == 4) run_groups_static_fast<7, 4>(p);
        else if (p.V == 9 && p.n == 8) run_groups_static_fast<9, 8>(p);
        else run_groups_static_fast<7, 2>(p);
        return 0;
    }
    int rc = mg::plan_launch(p, forced_group, 0, 227 * 1024, 228 * 1024);
    if (rc) return rc;
    if (mode == mg::MODE_OBS) dispatch<mg::MODE_OBS>(p, generic);
    else if (mode == mg::MODE_STEP) dispatch<mg::MODE_STEP>(p, generic);
    else dispatch<mg::MODE_STEP_OBS>(p, generic);
    return 0;
}

// mg_refresh_done_layouts on the CPU: the same env_is_done() / refresh_slot() the CUDA kernel calls.
extern "C" int sim_refresh_done_layouts(const MgConfig *c, int64_t num_envs, const MgState *s, const MgLayoutGen *gen) {
    mg::Params p;
    std::memset(&p, 0, sizeof(p));
    p.W = c->width; p.H = c->height; p.n = c->num_agents; p.V = c->view_size; p.max_steps = c->max_steps;
    p.flags = c->flags; p.hook = c->hook; p.hook_param = c->hook_param; p.ostride = c->obs_agent_stride;
    p.num_envs = (int32_t)num_envs; p.G = 16;
    mg::carve_smem(p);
    p.agents = s->agents; p.step_count = s->step_count; p.pcg_state = s->pcg_state; p.pcg_inc = s->pcg_inc;
    p.pool_grid = s->pool_grid; p.pool_agents = s->pool_agents; p.hook_state = s->hook_state;
    mg::LayoutGen lg;
    lg.family = gen->family; lg.a = gen->params[0]; lg.b = gen->params[1]; lg.c = gen->params[2]; lg.d = gen->params[3];
    lg.rng_state = gen->rng_state; lg.rng_inc = gen->rng_inc; lg.rng_buf = gen->rng_buf; lg.order_buf = gen->order_buf; lg.info = gen->info;
    int bad = 0;
    for (int64_t e = 0; e < num_envs; e++)
        if (mg::env_is_done(p, (size_t)e) && !mg::refresh_slot(p, lg, (size_t)e)) bad = 1;
    return bad;
}

// mg_build_static_obs on the CPU: the same static_build_entry() the CUDA kernel calls.
extern "C" int sim_build_static_obs(const MgConfig *c, const uint32_t *layout, uint8_t *table) {
    mg::Params p;
    std::memset(&p, 0, sizeof(p));
    p.W = c->width; p.H = c->height; p.n = c->num_agents; p.V = c->view_size; p.flags = c->flags;
    p.ostride = c->obs_agent_stride; p.G = 16;
    mg::carve_static(p, false, 1);
    for (int idx = 0; idx < p.W * p.H * 4; idx++) {
        const int dir = idx & 3, xy = idx >> 2, x = xy / p.H, y = xy - x * p.H;
        const int ts = mg::static_obs_stride(p.ostride);
        mg::static_build_entry(p, layout, x, y, dir, table + (size_t)idx * ts);
        ((uint32_t *)(table + (size_t)p.W * p.H * 4 * ts))[idx] = mg::static_move_word(p, layout, x, y, dir);
    }
    return 0;
}

// mg_gen_layouts_empty_random on the CPU: the same gen_layout_empty_random() the CUDA kernel calls.
extern "C" int sim_gen_layouts_empty_random(int W, int H, int n, int64_t K, uint64_t *rng_state,
                                            const uint64_t *rng_inc, uint64_t *rng_buf, uint32_t *cells,
                                            int8_t *agents) {
    int bad = 0;
    for (int64_t k = 0; k < K; k++) {
        mg::LayoutRng g;
        g.lo = rng_state[2 * k]; g.hi = rng_state[2 * k + 1]; g.ilo = rng_inc[2 * k]; g.ihi = rng_inc[2 * k + 1];
        g.has32 = (uint32_t)(rng_buf[k] >> 32) & 1u; g.buf32 = (uint32_t)rng_buf[k];
        if (!mg::gen_layout_empty_random(W, H, n, g, cells + k * (int64_t)(W + 1) * (H + 1), agents + k * n * 8)) bad = 1;
        rng_state[2 * k] = g.lo; rng_state[2 * k + 1] = g.hi;
        rng_buf[k] = ((uint64_t)g.has32 << 32) | g.buf32;
    }
    return bad;
}

extern "C" int sim_gen_layouts_bup(int S, int n, int64_t K, uint64_t *rng_state, const uint64_t *rng_inc,
                                   uint64_t *rng_buf, uint64_t *order_state, const uint64_t *order_inc,
                                   uint64_t *order_buf, uint32_t *cells, int8_t *agents, int32_t *info) {
    int bad = 0;
    const int W = 2 * (S - 1) + 1;
    for (int64_t k = 0; k < K; k++) {
        mg::LayoutRng g, o;
        g.lo = rng_state[2 * k]; g.hi = rng_state[2 * k + 1]; g.ilo = rng_inc[2 * k]; g.ihi = rng_inc[2 * k + 1];
        g.has32 = (uint32_t)(rng_buf[k] >> 32) & 1u; g.buf32 = (uint32_t)rng_buf[k];
        o.lo = order_state[2 * k]; o.hi = order_state[2 * k + 1]; o.ilo = order_inc[2 * k]; o.ihi = order_inc[2 * k + 1];
        o.has32 = (uint32_t)(order_buf[k] >> 32) & 1u; o.buf32 = (uint32_t)order_buf[k];
        info[k] = mg::gen_layout_bup(S, n, g, o, cells + k * (int64_t)(W + 1) * (S + 1), agents + k * n * 8);
        if (info[k] < 0) bad = 1;
        rng_state[2 * k] = g.lo; rng_state[2 * k + 1] = g.hi;
        rng_buf[k] = ((uint64_t)g.has32 << 32) | g.buf32;
        order_state[2 * k] = o.lo; order_state[2 * k + 1] = o.hi;
        order_buf[k] = ((uint64_t)o.has32 << 32) | o.buf32;
    }
    return bad;
}

extern "C" int sim_gen_layouts_red_blue_doors(int size, int n, int64_t K, uint64_t *rng_state, const uint64_t *rng_inc,
                                              uint64_t *rng_buf, uint32_t *cells, int8_t *agents) {
    int bad = 0;
    for (int64_t k = 0; k < K; k++) {
        mg::LayoutRng g;
        g.lo = rng_state[2 * k]; g.hi = rng_state[2 * k + 1]; g.ilo = rng_inc[2 * k]; g.ihi = rng_inc[2 * k + 1];
        g.has32 = (uint32_t)(rng_buf[k] >> 32) & 1u; g.buf32 = (uint32_t)rng_buf[k];
        if (!mg::gen_layout_red_blue_doors(size, n, g, cells + k * (int64_t)(2 * size + 1) * (size + 1), agents + k * n * 8)) bad = 1;
        rng_state[2 * k] = g.lo; rng_state[2 * k + 1] = g.hi;
        rng_buf[k] = ((uint64_t)g.has32 << 32) | g.buf32;
    }
    return bad;
}

extern "C" int sim_gen_layouts_locked_hallway(int num_rooms, int S, int mhk, int mkpr, int n, int64_t K,
                                              uint64_t *rng_state, const uint64_t *rng_inc, uint64_t *rng_buf,
                                              uint32_t *cells, int8_t *agents) {
    int bad = 0;
    const int64_t cs = (int64_t)(3 * (S - 1) + 2) * ((num_rooms / 2) * (S - 1) + 2);
    for (int64_t k = 0; k < K; k++) {
        mg::LayoutRng g;
        g.lo = rng_state[2 * k]; g.hi = rng_state[2 * k + 1]; g.ilo = rng_inc[2 * k]; g.ihi = rng_inc[2 * k + 1];
        g.has32 = (uint32_t)(rng_buf[k] >> 32) & 1u; g.buf32 = (uint32_t)rng_buf[k];
        if (!mg::gen_layout_locked_hallway(num_rooms, S, mhk, mkpr, n, g, cells + k * cs, agents + k * n * 8)) bad = 1;
        rng_state[2 * k] = g.lo; rng_state[2 * k + 1] = g.hi;
        rng_buf[k] = ((uint64_t)g.has32 << 32) | g.buf32;
    }
    return bad;
}

extern "C" int sim_gen_layouts_playground(int S, int rows, int cols, int n, int64_t K, uint64_t *rng_state,
                                          const uint64_t *rng_inc, uint64_t *rng_buf, uint64_t *order_state,
                                          const uint64_t *order_inc, uint64_t *order_buf, uint32_t *cells,
                                          int8_t *agents) {
    int bad = 0;
    const int64_t cs = (int64_t)(cols * (S - 1) + 2) * (rows * (S - 1) + 2);
    for (int64_t k = 0; k < K; k++) {
        mg::LayoutRng g, o;
        g.lo = rng_state[2 * k]; g.hi = rng_state[2 * k + 1]; g.ilo = rng_inc[2 * k]; g.ihi = rng_inc[2 * k + 1];
        g.has32 = (uint32_t)(rng_buf[k] >> 32) & 1u; g.buf32 = (uint32_t)rng_buf[k];
        o.lo = order_state[2 * k]; o.hi = order_state[2 * k + 1]; o.ilo = order_inc[2 * k]; o.ihi = order_inc[2 * k + 1];
        o.has32 = (uint32_t)(order_buf[k] >> 32) & 1u; o.buf32 = (uint32_t)order_buf[k];
        if (!mg::gen_layout_playground(S, rows, cols, n, g, o, cells + k * cs, agents + k * n * 8)) bad = 1;
        rng_state[2 * k] = g.lo; rng_state[2 * k + 1] = g.hi;
        rng_buf[k] = ((uint64_t)g.has32 << 32) | g.buf32;
        order_state[2 * k] = o.lo; order_state[2 * k + 1] = o.hi;
        order_buf[k] = ((uint64_t)o.has32 << 32) | o.buf32;
    }
    return bad;
}

// mg_step_obs_host_wire's per-env record on the CPU: the same wire_env_record() the CUDA kernel calls.
extern "C" int sim_wire_env_records(int n, int64_t E, const double *reward, const uint8_t *terminated,
                                    const uint8_t *truncated, uint8_t *records) {
    int bad = 0;
    for (int64_t e = 0; e < E; e++)
        if (!mg::wire_env_record(n, reward + e * n, terminated + e * n, truncated[e], records + e * mg::wire_record_bytes(n)))
            bad = 1;
    return bad;
}
