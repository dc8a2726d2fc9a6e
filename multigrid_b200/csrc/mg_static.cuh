// multigrid_b200 -- the STATIC-GRID fast path of the fused step/observe launch (sm_100a).
//
// Every Empty-family env of the reference (envs/empty.py: wall ring + goal, optional lava/floor subclasses)
// has a grid that no action can change: there is nothing to pick up, drop, or toggle. For such a batch
// (MG_FLAG_STATIC_GRID, the caller's promise, see include/multigrid_b200.h) two things follow:
//
//   * handle_actions (base.py:378-476) reduces to left / right / forward against ONE read-only layout;
//     pickup / drop / toggle / done are no-ops and no cell is ever written;
//   * gen_obs (utils/obs.py:66-102) of an agent is a pure function of (x, y, dir) -- slice, rotate and the
//     visibility flood only look at the layout -- plus the other agents drawn on top (utils/obs.py:163-171;
//     agents are see-through, so they never change what is visible). The (x, y, dir) part is MEMOISED:
//     mg_build_static_obs runs the engine's own observation code once per (x, y, dir) of the layout and
//     keeps the W*H*4 packed views in HBM (40 KB for 8x8 V=7, 256 KB for 16x16 V=9: L1/L2 resident).
//
// One warp = one group of G consecutive envs (32 for big batches), one warp per block, no barrier.
//   env lane    : its env's agent records (two 16-byte loads for n = 4), actions and PCG64 words go straight
//                 into registers; the position words live in a TRANSPOSED scratch in shared memory
//                 (word j of env l at [j][l]: the serial agent loop indexes it by the drawn order without
//                 bank conflicts and touches only its own column, so no warp sync until the observations);
//                 per-env outputs and the updated records leave from registers with 16-byte stores.
//   agent lane  : the 32 table entries of a pass are copied COOPERATIVELY -- a 16-byte load per lane covers
//                 whole entries with consecutive lanes (3 entries per instruction for V = 7), so an instruction
//                 touches ~6 cache lines instead of 32 -- into the pass's stage; each lane then patches the
//                 other agents of its env into its own slot (3 byte stores per visible agent); one TMA bulk
//                 store (UBLKCP) moves the pass's contiguous 4 736-byte span to HBM. Two stages alternate.
// Results are bit-identical to the general kernel (tests/test_static_path.py: every Empty fixture recorded
// from the reference, random static layouts against the C oracle, the full-size batch against the general
// kernel).
#pragma once
#include "mg_kernels.cuh"

namespace mg {

MG_HD int static_obs_stride(int ostride) { return align16(ostride); }

#ifdef __CUDA_ARCH__
#define MG_LDG(ptr) __ldg(ptr)
#else
#define MG_LDG(ptr) (*(ptr))
#endif

// Shared memory of one warp on the static path. Fast path (NT > 0): [stage 0][stage 1?][a0T: n x 32 words].
// Rolled path (NT == 0): [stage][keys/order][ag][act] as in the general kernel, without cells.
inline int carve_static(Params &p, bool fast, int nstage) {
    const int G = p.G, n = p.n;
    p.Hp = p.H + 1;
    p.cstride = (p.W + 1) * p.Hp;
    p.rcp_n = rcp32(n);
    p.stage_bytes = align16(LANES * p.ostride);
    p.pass_cell_bytes = 0; p.alias = 0; p.stage_extra = 0;
    int off = 0;
    p.off_stage = off; off += nstage * p.stage_bytes;
    p.off_cells = off;
    if (fast) {
        p.off_keys = off;
        p.off_ag = off; off += n * LANES * 4;  // a0T
        p.off_act = off;                       // the block's copy of the move words (small layouts)
        p.lut_words = (p.W * p.H * 4 <= 1024 && !p.no_lut) ? align16(p.W * p.H * 16) / 4 : 0;
        off += p.lut_words * 4;
    } else {
        p.off_keys = off;  off += n > 4 ? align16(G * n * 8) + align16(G * n) : 0;
        p.off_ag = off;    off += align16(G * n * 8);
        p.off_act = off;   off += align16(G * n);
    }
    p.off_rk = off; p.off_mbar = off;
    p.warp_bytes = off;
    return off;
}

// The unrolled kernels want observation slots that ARE table entries: obs_agent_stride = 3*V*V rounded up to
// 16 bytes (160 for V = 7, 256 for V = 9), so that the copy is 16-byte loads and stores throughout.
inline bool static_fast_shape(const Params &p) {
    const bool entry_stride = p.ostride == ((3 * p.V * p.V + 15) & ~15);
    return !p.generic_view && entry_stride &&
           ((p.V == 7 && (p.n == 4 || p.n == 2)) || (p.V == 9 && p.n == 8));
}

// Launch geometry of the static kernels. Unrolled shapes: one block = G envs, p.wpb warps (one stage each);
// the per-env part runs on warp 0, the G*n/32 observation passes are split over the warps. Rolled shapes: one
// warp per block, 16 envs. STATIC_SHAPES lists the instantiated (V, n, G, warps) combinations, the first entry
// of a (V, n) being its default for big batches; smaller batches take fewer envs per block so that every SM
// still gets work.
struct StaticShape { int V, n, G, wpb; };
constexpr StaticShape STATIC_SHAPES[] = {
    {7, 4, 16, 1}, {7, 4, 32, 2}, {7, 4, 32, 1}, {7, 4, 16, 2}, {7, 4, 8, 1},
    {7, 2, 32, 1}, {7, 2, 32, 2}, {7, 2, 16, 1},
    {9, 8, 16, 4}, {9, 8, 32, 4}, {9, 8, 16, 2}, {9, 8, 8, 1}, {9, 8, 8, 2},  // (8 envs: one warp, 13 KB -> the 2 048
                                                                           // blocks of BASELINE configs[3] are one wave;
                                                                           // two warps, 21 KB: 1.4 waves, 13.4 vs 11.0 us)
};

inline bool static_shape_ok(int V, int n, int G, int wpb) {
    for (const StaticShape &s : STATIC_SHAPES)
        if (s.V == V && s.n == n && s.G == G && s.wpb == wpb) return true;
    return false;
}

inline int plan_static(Params &p, int forced_G, int forced_wpb, int smem_per_block, int num_sms = 148) {
    const bool fast = static_fast_shape(p);
    p.wpb = 1;
    if (!fast) {
        p.G = (forced_G == 32 || forced_G == 8) ? forced_G : 16;
    } else {
        int G = 0, wpb = 0;
        for (const StaticShape &s : STATIC_SHAPES)  // default: the first shape of (V, n), halved while the batch
            if (s.V == p.V && s.n == p.n && !G) { G = s.G; wpb = s.wpb; }                // leaves SMs without work
        while (G > 8 && (p.num_envs + G - 1) / G < num_sms * 20) {
            int g2 = G / 2, w2 = 0;
            for (const StaticShape &s : STATIC_SHAPES)
                if (s.V == p.V && s.n == p.n && s.G == g2 && !w2) w2 = s.wpb;
            if (!w2) break;
            G = g2; wpb = w2;
        }
        if (forced_G || forced_wpb) {  // knobs: only instantiated combinations
            const int fg = forced_G ? forced_G : G;
            int fw = forced_wpb;
            if (!fw)
                for (const StaticShape &s : STATIC_SHAPES)
                    if (s.V == p.V && s.n == p.n && s.G == fg && !fw) fw = s.wpb;
            if (static_shape_ok(p.V, p.n, fg, fw)) { G = fg; wpb = fw; }
        }
        p.G = G; p.wpb = wpb;
    }
    p.nstage = p.wpb;
    if (carve_static(p, fast, p.nstage) > smem_per_block) return MG_ERR_TOO_LARGE;
    return 0;
}

// One entry of the table: the view of a lone agent at (x, y) facing `dir` carrying nothing, exactly as the
// general observation code computes it (obs_agent_generic: utils/obs.py:131-273), padded with zeros to the
// 16-byte entry stride.
MG_HD void static_build_entry(const Params &p, const uint32_t *layout, int x, int y, int dir, uint8_t *out) {
    const uint32_t a0 = (uint32_t)dir | ((uint32_t)x << 8) | ((uint32_t)y << 16);
    obs_agent_generic(p, layout, a0, CELL_EMPTY, out);
    for (int q = p.ostride; q < static_obs_stride(p.ostride); q++) out[q] = 0;
}

// The memoised `forward` action (base.py:420-436) of an agent at (x, y) facing dir: its position word after the
// move -- dir | x' << 8 | y' << 16 with (x', y') the cell in front when that cell is inside the grid and can be
// walked on (can_overlap, core/world_object.py:197-201, 287, 314, 339: empty, floor, goal, lava), else (x, y) --
// plus MOVE_OK when it moves and MOVE_GOAL / MOVE_LAVA for what it then stands on.
constexpr uint32_t MOVE_GOAL = 1u << 24, MOVE_LAVA = 1u << 25, MOVE_OK = 1u << 26;
MG_HD uint32_t static_move_word(const Params &p, const uint32_t *layout, int x, int y, int dir) {
    const int fx = x + (dir == 0) - (dir == 2), fy = y + (dir == 1) - (dir == 3);  // constants.py:21-30
    uint32_t w = (uint32_t)dir | ((uint32_t)x << 8) | ((uint32_t)y << 16);
    if ((unsigned)fx >= (unsigned)p.W || (unsigned)fy >= (unsigned)p.H) return w;
    const uint32_t t = layout[fx * p.Hp + fy] & 0xff;
    if (!((t == T_EMPTY) | (t == T_FLOOR) | (t == T_GOAL) | (t == T_LAVA))) return w;
    w = (uint32_t)dir | ((uint32_t)fx << 8) | ((uint32_t)fy << 16) | MOVE_OK;
    if (t == T_GOAL) w |= MOVE_GOAL;
    if (t == T_LAVA) w |= MOVE_LAVA;
    return w;
}

// Bytes of the table blob: W*H*4 observation entries followed by W*H*4 move words.
MG_HD int64_t static_table_bytes(int W, int H, int ostride) {
    return (int64_t)W * H * 4 * (static_obs_stride(ostride) + 4);
}

// ======================================================================================================
// Rolled path (any n, any odd V): per-env records in shared memory like the general kernel.
// ======================================================================================================

// MultiGridEnv.handle_actions (base.py:378-476) on a grid without doors, keys, balls or boxes: only
// left / right / forward can do anything. Agent k's position word is ag[k * AS] (AS = 2: env-major records,
// AS = 32: transposed scratch); act_of(k) yields its action.
template <int NT, int AS, typename ActFn>
MG_HD void static_handle_actions(const Params &p, uint32_t *ag, uint32_t ord, const uint8_t *order_e, int G,
                                 ActFn act_of, uint32_t &rewarded) {
    const int n = NT ? NT : p.n;
    const bool packed_order = NT ? true : n <= 4;
    const uint32_t *layout = p.pool_grid;
    const uint32_t TERM = 1u << 24;
#pragma unroll (NT ? NT : 1)
    for (int r = 0; r < n; r++) {
        const int k = packed_order ? (int)(ord & 15u) : (int)order_e[r * G];
        ord >>= 4;
        const int act = act_of(k);
        const uint32_t a0 = ag[k * AS];
        if (act < 0) continue;            // id not in the action dict (base.py:403-404)
        if (a0 & 0xff000000u) continue;   // terminated (base.py:408-409)
        const uint32_t dir = a0 & 3u;
        if (act == ACT_LEFT)  { ag[k * AS] = (a0 & ~0xffu) | ((dir + 3u) & 3u); continue; }  // base.py:412-413
        if (act == ACT_RIGHT) { ag[k * AS] = (a0 & ~0xffu) | ((dir + 1u) & 3u); continue; }  // base.py:416-417
        if (act != ACT_FORWARD) {  // pickup / drop / toggle find nothing to act on; done is a no-op
            if (act > ACT_DONE) status_or(p.status, 1);  // reference: ValueError (base.py:473-474)
            continue;
        }
        const int dx = (dir == 0) - (dir == 2), dy = (dir == 1) - (dir == 3);  // constants.py:21-30
        const int fx = (int)((a0 >> 8) & 0xff) + dx, fy = (int)((a0 >> 16) & 0xff) + dy;
        if ((unsigned)fx >= (unsigned)p.W || (unsigned)fy >= (unsigned)p.H) continue;
        const uint32_t t = MG_LDG(layout + fx * p.Hp + fy) & 0xff;
        // can_overlap (core/world_object.py:197-201, 287, 314, 339): empty, floor, goal, lava
        if (!((t == T_EMPTY) | (t == T_FLOOR) | (t == T_GOAL) | (t == T_LAVA))) continue;
        const uint32_t fxy = (uint32_t)fx | ((uint32_t)fy << 8);
        if (!(p.flags & MG_FLAG_ALLOW_OVERLAP)) {  // base.py:425-429 (terminated agents count)
            bool hit = false;
#pragma unroll (NT ? NT : 1)
            for (int j = 0; j < n; j++) hit |= ((ag[j * AS] >> 8) & 0xffffu) == fxy;
            if (hit) continue;
        }
        ag[k * AS] = (a0 & 0xff0000ffu) | (fxy << 8);
        if (t == T_GOAL) {  // on_success, base.py:478-507
            if (p.flags & MG_FLAG_SUCCESS_ANY) {
#pragma unroll (NT ? NT : 1)
                for (int j = 0; j < n; j++) ag[j * AS] |= TERM;
            } else {
                ag[k * AS] |= TERM;
            }
            rewarded |= (p.flags & MG_FLAG_JOINT_REWARD) ? all_agents(p) : (1u << k);
        }
        if (t == T_LAVA) {  // on_failure, base.py:509-532
            if (p.flags & MG_FLAG_FAILURE_ANY) {
#pragma unroll (NT ? NT : 1)
                for (int j = 0; j < n; j++) ag[j * AS] |= TERM;
            } else {
                ag[k * AS] |= TERM;
            }
        }
    }
}

// Per-env outputs of one step, straight from the env lane's registers (same as phase_step without hooks).
template <int NT>
MG_HD void static_env_outputs(const Params &p, size_t e, size_t eo, const EnvRegs &r, uint32_t term_mask,
                              uint32_t rewarded, bool truncated) {
    const int n = NT ? NT : p.n;
    p.step_count[e] = r.sc;
    if (n > 1) { U128 s; s.lo = r.lo; s.hi = r.hi; *(U128 *)(p.pcg_state + 2 * e) = s; }
    p.truncated[eo] = (uint8_t)truncated;
    const double rv = rewarded ? reward_value(r.sc, p.max_steps) : 0.0;  // base.py:394, 598-602
    if (NT == 4) {
        *(uint32_t *)(p.terminated + eo * 4) = bits_to_bytes4(term_mask);
#ifdef __CUDA_ARCH__
        if (((uintptr_t)p.reward & 15u) == 0) {
            double2 *rw = (double2 *)(p.reward + eo * 4);
            rw[0] = make_double2((rewarded & 1u) ? rv : 0.0, (rewarded & 2u) ? rv : 0.0);
            rw[1] = make_double2((rewarded & 4u) ? rv : 0.0, (rewarded & 8u) ? rv : 0.0);
            return;
        }
#endif
        for (int j = 0; j < 4; j++) p.reward[eo * 4 + j] = ((rewarded >> j) & 1u) ? rv : 0.0;
    } else {
#pragma unroll (NT ? NT : 1)
        for (int j = 0; j < n; j++) p.terminated[eo * n + j] = (uint8_t)((term_mask >> j) & 1u);
#pragma unroll (NT ? NT : 1)
        for (int j = 0; j < n; j++) p.reward[eo * n + j] = ((rewarded >> j) & 1u) ? rv : 0.0;
    }
}

// The env's lane on the rolled path: auto-reset decision (is_done, base.py:534-539), transition, outputs.
// Same observable behaviour as phase_reset + phase_step of the general kernel with one pool layout and no
// hook: a reset env takes pool_agents[0], consumes no draw and reports reward 0 / not terminated / not
// truncated for this launch.
MG_HD void static_env_step(const Params &p, const Group &g, int i, EnvRegs &r, const OrderDraw &d, size_t tE = 0) {
    if (i < 0) return;
    const int n = p.n;
    uint32_t *ag = g.ag + i * n * 2;
    const size_t e = (size_t)(g.e0 + i), eo = e + tE;
    bool was_reset = false;
    if (p.flags & MG_FLAG_AUTO_RESET) {
        uint32_t all_term = 1;
        for (int j = 0; j < n; j++) all_term &= ((ag[j * 2] >> 24) & 0xff) != 0;
        if (all_term || r.sc >= p.max_steps) {
            was_reset = true;
            r.sc = 0;
            const uint32_t *src = (const uint32_t *)p.pool_agents;
            for (int j = 0; j < n * 2; j++) ag[j] = MG_LDG(src + j);
        }
    }
    uint32_t rewarded = 0;
    bool truncated = false;
    if (!was_reset) {
        r.sc += 1;  // base.py:333
        const int8_t *act_e = g.act + i * n;
        static_handle_actions<0, 2>(p, ag, d.ord, g.order + i, p.G, [&](int k) { return (int)act_e[k]; }, rewarded);
        truncated = r.sc >= p.max_steps;  // base.py:339
    } else {
        r.lo = d.lo0; r.hi = d.hi0;  // no step, no draw
    }
    static_env_outputs<0>(p, e, eo, r, terminated_mask(p, ag), rewarded, truncated);
}

// Drawing the other agents of the env into an agent's packed view (the lane's stage slot `sb`).
// gen_obs_grid (utils/obs.py:163-171, 199-207): non-terminated agents are drawn into the grid in ascending
// index (the highest wins); the viewer's own cell shows what it carries (nothing). Agent j appears at view
// cell (a, b) = (lat + V/2, V-1-fwd), (fwd, lat) = its offset in the viewer's frame, if that cell is visible
// (agents are see-through, so the table's mask is final: UNSEEN cells have type 0). word(j) = agent j's
// position word, color(j) = its colour byte.
template <int VT, int NT, typename WordFn, typename ColorFn>
MG_HD void static_overlay(const Params &p, uint32_t a0, uint8_t *sb, WordFn word, ColorFn color) {
    const int n = NT ? NT : p.n, V = VT ? VT : p.V, half = V >> 1;
    if (n <= 1) return;  // a lone agent is not drawn (utils/obs.py:172-173)
    const uint32_t dir = a0 & 3u;
    const int x = (a0 >> 8) & 0xff, y = (a0 >> 16) & 0xff;
    // viewer frame: f = DIR_TO_VEC[dir], r = (-f.y, f.x)
    const int Fx = (dir == 0u) - (dir == 2u), Fy = (dir == 1u) - (dir == 3u);
    constexpr int MAXN = NT ? NT : 1;
    if constexpr (NT != 0) {
        // (reads first, then writes: whether a cell is visible does not depend on what was drawn into it)
        uint32_t off[MAXN], w[MAXN];
        bool hit[MAXN];
#pragma unroll
        for (int j = 0; j < NT; j++) {
            w[j] = word(j);
            const int dx = (int)((w[j] >> 8) & 0xff) - x, dy = (int)((w[j] >> 16) & 0xff) - y;
            const int fwd = Fx * dx + Fy * dy, lat = Fx * dy - Fy * dx;
            const int a = lat + half;
            const bool in_view = !(w[j] & 0xff000000u) && (unsigned)fwd < (unsigned)V && (unsigned)a < (unsigned)V &&
                                 (fwd | lat) != 0;
            off[j] = in_view ? (uint32_t)((a * V + (V - 1 - fwd)) * 3) : 0u;
            hit[j] = in_view && sb[off[j]] != 0;
        }
#pragma unroll
        for (int j = 0; j < NT; j++)
            if (hit[j]) {
                uint8_t *c = sb + off[j];
                c[0] = (uint8_t)T_AGENT; c[1] = (uint8_t)color(j); c[2] = (uint8_t)(w[j] & 0xff);
            }
    } else {
        for (int j = 0; j < n; j++) {
            const uint32_t b0 = word(j);
            if (b0 & 0xff000000u) continue;
            const int dx = (int)((b0 >> 8) & 0xff) - x, dy = (int)((b0 >> 16) & 0xff) - y;
            const int fwd = Fx * dx + Fy * dy, lat = Fx * dy - Fy * dx;
            const int a = lat + half;
            if ((unsigned)fwd >= (unsigned)V || (unsigned)a >= (unsigned)V || (fwd | lat) == 0) continue;
            uint8_t *c = sb + (a * V + (V - 1 - fwd)) * 3;
            if (c[0] == 0) continue;
            c[0] = (uint8_t)T_AGENT; c[1] = (uint8_t)color(j); c[2] = (uint8_t)(b0 & 0xff);
        }
    }
}

// Byte offset of an agent's table entry; a position outside the grid (a broken promise the kernel can see)
// is flagged and clamped so that no load leaves the table.
MG_HD uint32_t static_entry_offset(const Params &p, uint32_t a0) {
    const uint32_t dir = a0 & 3u;
    uint32_t x = (a0 >> 8) & 0xff, y = (a0 >> 16) & 0xff;
    if (x >= (uint32_t)p.W || y >= (uint32_t)p.H) {
        status_or(p.status, 4);
        x = 0; y = 0;
    }
    return ((x * (uint32_t)p.H + y) * 4u + dir) * (uint32_t)p.static_stride;
}

// Rolled path: observation of agent task `pass*32 + lane` into the lane's stage slot.
MG_HD void static_obs_agent(const Params &p, const Group &g, int pass, int lane, uint8_t *stage) {
    const int n = p.n;
    const int id = pass * LANES + lane;
    if (id >= g.ne * n) return;
    const int el = (int)fastdiv((uint32_t)id, p.rcp_n);
    const uint32_t *agw = g.ag + el * n * 2;
    const uint32_t a0 = g.ag[id * 2], a1 = g.ag[id * 2 + 1];
    if ((a1 & 0xff) != T_EMPTY) status_or(p.status, 4);  // carrying: the promise does not hold
    const uint32_t *s32 = (const uint32_t *)(p.static_obs + static_entry_offset(p, a0));
    uint8_t *sb = stage + lane * p.ostride;
    uint32_t *slot = (uint32_t *)sb;
    for (int w = 0; w * 4 < p.ostride; w++) slot[w] = MG_LDG(s32 + w);
    static_overlay<0, 0>(p, a0, sb, [&](int j) { return agw[j * 2]; }, [&](int j) { return agw[j * 2 + 1] >> 24; });
}

// Plain warp-cooperative moves of the group's small spans (agents, actions) on the rolled path.
MG_HD void static_copy(void *dst, const void *src, int nbytes, int lane) {
#ifdef __CUDACC__
    typedef uint4 V16;
#else
    struct alignas(16) V16 { uint32_t v[4]; };
#endif
    const int nv = nbytes >> 4;
    for (int v = lane; v < nv; v += LANES) ((V16 *)dst)[v] = ((const V16 *)src)[v];
    for (int w = (nv << 2) + lane; w * 4 < nbytes; w += LANES) ((uint32_t *)dst)[w] = ((const uint32_t *)src)[w];
}

MG_HD void static_load(const Params &p, const Group &g, int lane, int t = 0) {
    const size_t e0 = (size_t)g.e0;
    const int n = p.n;
    // (a group starts at an even env, so its agents span is 16-byte aligned)
    if (t == 0) static_copy(g.ag, p.agents + e0 * n * 8, g.ne * n * 8, lane);
    const int8_t *act = p.actions + ((size_t)t * p.num_envs + e0) * n;
    if (((uintptr_t)act & 15u) == 0 && ((g.ne * n) & 3) == 0) static_copy(g.act, act, g.ne * n, lane);
    else warp_copy(g.act, act, g.ne * n, lane);
}

MG_HD void static_store_agents(const Params &p, const Group &g, int lane) {
    static_copy(p.agents + (size_t)g.e0 * p.n * 8, g.ag, g.ne * p.n * 8, lane);
}

MG_HD void static_env_load(const Params &p, const Group &g, int i, EnvRegs &r) {
    r.lo = r.hi = r.ilo = r.ihi = 0; r.sc = 0; r.lidx = 0; r.hs = 0;
    if (i < 0) return;
    const size_t e = (size_t)(g.e0 + i);
    r.sc = p.step_count[e];
    if (p.n > 1) {
        const U128 s = *(const U128 *)(p.pcg_state + 2 * e), c = *(const U128 *)(p.pcg_inc + 2 * e);
        r.lo = s.lo; r.hi = s.hi; r.ilo = c.lo; r.ihi = c.hi;
    }
}

// ======================================================================================================
// Fast path (NT = n in {2, 4, 8}): everything of an env in its lane's registers + one column of the
// transposed scratch a0T[j][lane]; no cross-lane traffic before the observation phase.
// ======================================================================================================

// base.py:399 for n <= 8 in registers: order packed 4 bits per rank.
template <int NT>
MG_HD uint32_t static_draw_order(EnvRegs &r) {
    if (NT == 1) return 0;
    uint64_t k[NT];
#pragma unroll
    for (int j = 0; j < NT; j++) k[j] = pcg64_next53(r.lo, r.hi, r.ilo, r.ihi);
    uint32_t rank[NT];
#pragma unroll
    for (int j = 0; j < NT; j++) rank[j] = 0;
#pragma unroll
    for (int q = 0; q < NT; q++)
#pragma unroll
        for (int j = q + 1; j < NT; j++) {  // stable ascending: q < j goes first on ties
            const uint32_t q_first = k[q] <= k[j];
            rank[j] += q_first; rank[q] += q_first ^ 1u;
        }
    uint32_t ord = 0;
#pragma unroll
    for (int j = 0; j < NT; j++) ord |= (uint32_t)j << (4 * rank[j]);
    return ord;
}

#ifdef __CUDACC__
typedef uint4 SV16;
typedef uint2 SV8;
#else
struct alignas(16) SV16 { uint32_t x, y, z, w; };
struct alignas(8) SV8 { uint32_t x, y; };
#endif

// PCG64 jump-ahead: the state after K steps of s' = s * A + inc (mod 2^128) is s * A^K + inc * (1 + A + ... + A^(K-1)).
// The two constants are compile-time 128-bit values (numpy's PCG64 multiplier, numpy/random/src/pcg64/pcg64.h).
constexpr unsigned __int128 PCG64_MULT = ((unsigned __int128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;
#ifdef __CUDACC__
#define MG_CONSTEXPR_HD __host__ __device__ constexpr
#else
#define MG_CONSTEXPR_HD constexpr
#endif
MG_CONSTEXPR_HD unsigned __int128 pcg64_pow(int k) {
    unsigned __int128 r = 1;
    for (int i = 0; i < k; i++) r *= PCG64_MULT;
    return r;
}
MG_CONSTEXPR_HD unsigned __int128 pcg64_geo(int k) {
    unsigned __int128 r = 0, q = 1;
    for (int i = 0; i < k; i++) { r += q; q *= PCG64_MULT; }
    return r;
}
template <int K>
MG_HD void pcg64_jump(uint64_t &lo, uint64_t &hi, uint64_t inc_lo, uint64_t inc_hi) {
    constexpr unsigned __int128 A = pcg64_pow(K), G = pcg64_geo(K);
    constexpr uint64_t A_LO = (uint64_t)A, A_HI = (uint64_t)(A >> 64), G_LO = (uint64_t)G, G_HI = (uint64_t)(G >> 64);
    const uint64_t rlo = lo * A_LO, rhi = mulhi64(lo, A_LO) + lo * A_HI + hi * A_LO;
    const uint64_t glo = inc_lo * G_LO, ghi = mulhi64(inc_lo, G_LO) + inc_lo * G_HI + inc_hi * G_LO;
    const uint64_t slo = rlo + glo;
    hi = rhi + ghi + (slo < rlo ? 1ull : 0ull);
    lo = slo;
}

// handle_actions on a static grid, unrolled and branch-free except for the rare goal / lava outcome: every
// agent of the drawn order selects between its rotated word and the memoised `forward` word of its
// (x, y, dir) (static_move_word: one 4-byte load from a 1 KB table instead of the front-cell arithmetic).
template <int NT>
MG_HD void static_transition_fast(const Params &p, const uint32_t *moves, uint32_t *col, uint32_t ord, uint32_t acts0,
                                  uint32_t acts1, uint32_t &rewarded) {
    const uint32_t TERM = 1u << 24;
    const bool no_overlap = !(p.flags & MG_FLAG_ALLOW_OVERLAP);
    const uint32_t last = (uint32_t)(p.W * p.H * 4 - 1);
#pragma unroll
    for (int r = 0; r < NT; r++) {
        const int k = (int)(ord & 15u);
        ord >>= 4;
        const int act = (int)(int8_t)((NT > 4 && k >= 4 ? acts1 : acts0) >> (8 * (k & 3)));
        const uint32_t a0 = col[k * LANES];
        const uint32_t dir = a0 & 3u;
        // id in the action dict and not terminated (base.py:403-409)
        const bool live = act >= 0 && a0 < TERM;
        uint32_t idx = ((((a0 >> 8) & 0xff) * (uint32_t)p.H + ((a0 >> 16) & 0xff)) << 2) | dir;
        idx = idx < last ? idx : last;  // (memory safety only: the promise keeps agents inside the grid)
        const uint32_t mv = moves[idx];
        const uint32_t turn = act == ACT_LEFT ? 3u : (act == ACT_RIGHT ? 1u : 0u);       // base.py:412-417
        uint32_t na = (a0 & ~0xffu) | ((dir + turn) & 3u);
        bool move = live && act == ACT_FORWARD && (mv & MOVE_OK);                        // base.py:420-423
        if (no_overlap) {  // base.py:425-429 (terminated agents count)
            bool hit = false;
#pragma unroll
            for (int j = 0; j < NT; j++) hit |= ((col[j * LANES] ^ mv) & 0x00ffff00u) == 0;
            move = move && !hit;
        }
        if (move) na = mv & 0x00ffffffu;
        if (live) col[k * LANES] = na;
        if (live && act > ACT_DONE) status_or(p.status, 1);  // reference: ValueError (base.py:473-474)
        if (move && (mv & (MOVE_GOAL | MOVE_LAVA))) {
            const uint32_t any = (mv & MOVE_GOAL) ? MG_FLAG_SUCCESS_ANY : MG_FLAG_FAILURE_ANY;  // base.py:478-532
            if (p.flags & any) {
#pragma unroll
                for (int j = 0; j < NT; j++) col[j * LANES] |= TERM;
            } else {
                col[k * LANES] |= TERM;
            }
            if (mv & MOVE_GOAL) rewarded |= (p.flags & MG_FLAG_JOINT_REWARD) ? ((1u << NT) - 1u) : (1u << k);
        }
    }
}

// The whole per-env part of a launch for env lane `lane` (< g.ne) of the group: load, draw, auto-reset,
// transition, outputs, records back to HBM. a0T = the warp's transposed scratch.
template <int NT>
MG_HD void static_fast_env(const Params &p, const Group &g, int lane, uint32_t *a0T, const uint32_t *moves,
                           size_t tE = 0) {
    if (lane >= g.ne) return;
    const size_t e = (size_t)(g.e0 + lane), eo = e + tE;
    uint32_t *col = a0T + lane;  // word j of this env at col[j * 32]
    // ---- load: records (16-byte vectors), actions, scalars
    uint32_t a1[NT];
    SV16 *rec = (SV16 *)(p.agents + e * NT * 8);
#pragma unroll
    for (int v = 0; v < NT / 2; v++) {
        const SV16 q = rec[v];
        col[(2 * v) * LANES] = q.x; a1[2 * v] = q.y;
        col[(2 * v + 1) * LANES] = q.z; a1[2 * v + 1] = q.w;
    }
    uint32_t acts0 = 0, acts1 = 0;
    const int8_t *ap = p.actions + eo * NT;
    if (NT == 2) acts0 = *(const uint16_t *)ap;
    else if (NT == 4) acts0 = *(const uint32_t *)ap;
    else { const SV8 q = *(const SV8 *)ap; acts0 = q.x; acts1 = q.y; }
    EnvRegs r;
    r.lidx = 0; r.hs = 0;
    r.sc = p.step_count[e];
    {
        const U128 s = *(const U128 *)(p.pcg_state + 2 * e), c = *(const U128 *)(p.pcg_inc + 2 * e);
        r.lo = s.lo; r.hi = s.hi; r.ilo = c.lo; r.ihi = c.hi;
    }
    const uint64_t lo0 = r.lo, hi0 = r.hi;
    // ---- auto-reset decision (is_done, base.py:534-539)
    bool was_reset = false;
    if (p.flags & MG_FLAG_AUTO_RESET) {
        uint32_t all_term = 1;
#pragma unroll
        for (int j = 0; j < NT; j++) all_term &= (col[j * LANES] & 0xff000000u) != 0;
        if (all_term || r.sc >= p.max_steps) {
            was_reset = true;
            r.sc = 0;
            const uint32_t *src = (const uint32_t *)p.pool_agents;
#pragma unroll
            for (int j = 0; j < NT; j++) { col[j * LANES] = MG_LDG(src + 2 * j); a1[j] = MG_LDG(src + 2 * j + 1); }
        }
    }
    uint32_t rewarded = 0;
    bool truncated = false;
    if (!was_reset) {
        r.sc += 1;  // base.py:333
        // Order-independent fast path. The drawn order (base.py:396-399) only matters when agents interact: through
        // the no-overlap test (base.py:425-429) or through a goal / lava outcome (termination of the others, rewards).
        // With overlap allowed (the reference's default) and no agent stepping onto goal or lava in this step -- some
        // 98 % of the env-steps of Empty-8x8 -- every agent's move is independent of the others': the n moves are
        // computed side by side in registers, the permutation is never materialised (no draws' outputs, no sort, no
        // serial loop through shared memory), and the env's PCG64 stream is advanced by its n draws with ONE
        // jump-ahead. Anything else takes the exact serial path below from the same state.
        const uint32_t TERM = 1u << 24;
        bool slow = !(p.flags & MG_FLAG_ALLOW_OVERLAP), bad = false;
        uint32_t na[NT];
        if (!slow) {
            const uint32_t last = (uint32_t)(p.W * p.H * 4 - 1);
#pragma unroll
            for (int j = 0; j < NT; j++) {
                const int act = (int)(int8_t)((NT > 4 && j >= 4 ? acts1 : acts0) >> (8 * (j & 3)));
                const uint32_t a0 = col[j * LANES], dir = a0 & 3u;
                const bool live = act >= 0 && a0 < TERM;  // id in the action dict and not terminated (base.py:403-409)
                uint32_t idx = ((((a0 >> 8) & 0xff) * (uint32_t)p.H + ((a0 >> 16) & 0xff)) << 2) | dir;
                idx = idx < last ? idx : last;
                const uint32_t mv = moves[idx];
                const uint32_t turn = act == ACT_LEFT ? 3u : (act == ACT_RIGHT ? 1u : 0u);  // base.py:412-417
                const bool move = live && act == ACT_FORWARD && (mv & MOVE_OK);             // base.py:420-423
                na[j] = !live ? a0 : (move ? (mv & 0x00ffffffu) : ((a0 & ~0xffu) | ((dir + turn) & 3u)));
                slow |= move && (mv & (MOVE_GOAL | MOVE_LAVA)) != 0;
                bad |= live && act > ACT_DONE;  // reference: ValueError (base.py:473-474)
            }
        }
        if (!slow) {
#pragma unroll
            for (int j = 0; j < NT; j++) col[j * LANES] = na[j];
            if (bad) status_or(p.status, 1);
            pcg64_jump<NT>(r.lo, r.hi, r.ilo, r.ihi);  // the n draws of np_random.random(size=n), base.py:399
        } else {
            const uint32_t ord = static_draw_order<NT>(r);
            static_transition_fast<NT>(p, moves, col, ord, acts0, acts1, rewarded);
        }
        truncated = r.sc >= p.max_steps;  // base.py:339
    } else {
        r.lo = lo0; r.hi = hi0;  // no step, no draw
    }
    // ---- outputs and records
    uint32_t w0[NT], term_mask = 0, carried = 0;
#pragma unroll
    for (int j = 0; j < NT; j++) {
        w0[j] = col[j * LANES];
        term_mask |= (uint32_t)((w0[j] & 0xff000000u) != 0) << j;
        carried |= a1[j] ^ (uint32_t)T_EMPTY;
    }
    static_env_outputs<NT>(p, e, eo, r, term_mask, rewarded, truncated);
#pragma unroll
    for (int v = 0; v < NT / 2; v++) {
        SV16 q;
        q.x = w0[2 * v]; q.y = a1[2 * v]; q.z = w0[2 * v + 1]; q.w = a1[2 * v + 1];
        rec[v] = q;
    }
    if (carried & 0xff) status_or(p.status, 4);  // an agent carries something: the promise does not hold
}

// The other agents of the env drawn into the viewer's slot (see static_overlay), unrolled: NT - 1 candidates in
// ascending index, reads before writes (whether a cell is visible does not depend on what was drawn into it).
template <int VT, int NT>
MG_HD void static_overlay_fast(uint32_t a0, int k, const uint32_t *a0T_el, uint32_t colors, uint8_t *sb) {
    constexpr int V = VT, half = VT / 2, C0 = 3 * (V * half + V - 1);
    const uint32_t dir = a0 & 3u;
    const int x = (a0 >> 8) & 0xff, y = (a0 >> 16) & 0xff;
    const int Fx = (dir == 0u) - (dir == 2u), Fy = (dir == 1u) - (dir == 3u);  // f = DIR_TO_VEC[dir], r = (-f.y, f.x)
    const int KX = -3 * (V * Fy + Fx), KY = 3 * (V * Fx - Fy);                 // byte offset of (dx, dy) in the view
    uint32_t off[NT - 1], w[NT - 1], cj[NT - 1];
    bool hit[NT - 1];
#pragma unroll
    for (int i = 0; i < NT - 1; i++) {
        const int j = i + (i >= k);
        w[i] = a0T_el[j * LANES];
        cj[i] = (colors >> (4 * j)) & 15u;
        const int dx = (int)((w[i] >> 8) & 0xff) - x, dy = (int)((w[i] >> 16) & 0xff) - y;
        const int fwd = Fx * dx + Fy * dy, lat = Fx * dy - Fy * dx;
        const bool in_view = w[i] < (1u << 24) && (unsigned)fwd < (unsigned)V && (unsigned)(lat + half) < (unsigned)V &&
                             (dx | dy) != 0;
        off[i] = in_view ? (uint32_t)(C0 + KX * dx + KY * dy) : 0u;
        hit[i] = in_view && sb[off[i]] != 0;
    }
#pragma unroll
    for (int i = 0; i < NT - 1; i++)
        if (hit[i]) {
            uint8_t *c = sb + off[i];
            c[0] = (uint8_t)T_AGENT; c[1] = (uint8_t)cj[i]; c[2] = (uint8_t)(w[i] & 0xff);
        }
}

// Agent task -> local env for the unrolled agent counts.
template <int NT> MG_HD int static_task_env(int id) { return NT == 8 ? id >> 3 : (NT == 4 ? id >> 2 : id >> 1); }

// Geometry of the cooperative table copy for view VT: slots and entries of TS bytes = PV 16-byte pieces, APR
// entries per warp-wide 16-byte load (PV consecutive lanes each), ROUNDS loads per pass of 32 entries.
template <int VT> struct StaticCopy {
    static constexpr int TS = (3 * VT * VT + 15) & ~15, OS = TS, PV = TS / 16;
    static constexpr int APR = PV >= LANES ? 1 : LANES / PV, ROUNDS = (LANES + APR - 1) / APR;
};

#ifdef __CUDACC__
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// One pass: lanes [PV*a, PV*a + PV) move the PV 16-byte pieces of one entry into its slot, APR entries per
// instruction. A load touches at most 2*APR cache lines (6 for V = 7) where one entry per LANE would touch 32.
// FULL = all 32 slots of the pass are in use.
template <int VT, bool FULL>
__device__ __forceinline__ void static_coop_copy(const Params &p, uint32_t my_ent, int cnt, uint8_t *stage, int lane) {
    typedef StaticCopy<VT> C;
    const int la = lane / C::PV, q = lane - la * C::PV;
    const bool lane_ok = la < C::APR;
    const uint8_t *base = p.static_obs + 16 * q;
    const uint32_t dst0 = smem_u32(stage) + (uint32_t)(la * C::TS + 16 * q);
    constexpr int CH = C::ROUNDS > 11 ? 8 : C::ROUNDS;  // loads in flight per lane (16-byte registers)
#pragma unroll
    for (int r0 = 0; r0 < C::ROUNDS; r0 += CH) {
        uint4 v[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) {
            const int r = r0 + c, A = r * C::APR + la;
            if (r < C::ROUNDS) {
                const uint32_t ent = __shfl_sync(0xffffffffu, my_ent, A & 31);
                const bool ok = lane_ok && (FULL ? A < LANES : A < cnt);
                v[c] = make_uint4(0, 0, 0, 0);
                if (ok) v[c] = __ldg((const uint4 *)(base + ent));
            }
        }
#pragma unroll
        for (int c = 0; c < CH; c++) {
            const int r = r0 + c, A = r * C::APR + la;
            if (r < C::ROUNDS) {
                const bool ok = lane_ok && (FULL ? A < LANES : A < cnt);
                if (ok) sts128(dst0 + (uint32_t)(r * C::APR * C::TS), v[c]);
            }
        }
    }
}

// One thread per table entry (x, y, dir).
__global__ void static_build_kernel(const __grid_constant__ Params p, const uint32_t *__restrict__ layout,
                                    uint8_t *__restrict__ table) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.W * p.H * 4) return;
    const int dir = idx & 3, xy = idx >> 2, x = xy / p.H, y = xy - x * p.H;
    const int ts = static_obs_stride(p.ostride);
    static_build_entry(p, layout, x, y, dir, table + (size_t)idx * ts);
    ((uint32_t *)(table + (size_t)p.W * p.H * 4 * ts))[idx] = static_move_word(p, layout, x, y, dir);
}

// The fused step + observe launch on a static grid, unrolled shapes. One block = GW consecutive envs and NWARP
// warps: warp 0 runs the per-env part at (up to) full lane width, then -- after the block's only barrier --
// every warp produces its share of the GW * NT / 32 observation passes from its own stage.
template <int VT, int NT, int GW, int NWARP>
__global__ void __launch_bounds__(32 * NWARP, NWARP == 1 ? 24 : (NWARP == 2 ? 14 : 7))
static_fast_kernel(const __grid_constant__ Params p) {
    typedef StaticCopy<VT> C;
    constexpr int PASSES = GW * NT / LANES, PPW = PASSES / NWARP;  // per full block / per warp
    static_assert(PASSES >= 1 && PPW * NWARP == PASSES, "observation passes must split evenly over the warps");
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, group = blockIdx.x;
    pdl_launch_dependents();
    const Group g = group_view(p, smem, group);
    uint32_t *a0T = g.ag;
    const bool full = g.ne == GW;
    const bool bulk = p.use_bulk && full;  // TMA bulk stores need 16-byte multiples: full groups only
    trace_mark(p, group * NWARP + warp, lane, 0);
    trace_mark(p, group * NWARP + warp, lane, 7);
    const uint64_t policy = (p.l2hint & 2) ? l2_policy_evict_first() : 0ull;
    uint32_t colors = 0;  // colour byte of agent j at bits 4j (colours are < 6; read-only pool data, so the
                          // loads may precede griddepcontrol.wait; they are consumed after the per-env part)
#pragma unroll
    for (int j = 0; j < NT; j++) colors |= (uint32_t)(uint8_t)__ldg(p.pool_agents + j * 8 + 7) << (4 * j);
    // the move words are constants: a small table is copied into shared memory before griddepcontrol.wait (under
    // the previous launch's tail), so the serial agent loop reads it with shared-memory latency
    const uint32_t *moves = p.static_move;
    if (p.lut_words) {
        uint32_t *lut = (uint32_t *)(smem + p.off_act);
        for (int v = threadIdx.x; v < p.lut_words / 4; v += 32 * NWARP)
            ((uint4 *)lut)[v] = __ldg((const uint4 *)p.static_move + v);
        moves = lut;
        if (NWARP == 1) __syncwarp();
        else __syncthreads();
    }
    pdl_wait();  // nothing of the previous launch is read or overwritten before this point
    if (warp == 0) static_fast_env<NT>(p, g, lane, a0T, moves);
    if (NWARP == 1) __syncwarp();
    else __syncthreads();
    trace_mark(p, group * NWARP + warp, lane, 1);
    trace_mark(p, group * NWARP + warp, lane, 2);
    uint8_t *stage = g.stage + warp * p.stage_bytes;
    const int tasks = g.ne * NT;
#pragma unroll
    for (int i = 0; i < PPW; i++) {
        const int pass = warp * PPW + i;
        const int id = pass * LANES + lane;
        if (!full && pass * LANES >= tasks) break;  // (whole warp)
        const int cnt = full ? LANES : (tasks - pass * LANES < LANES ? tasks - pass * LANES : LANES);
        const int el = static_task_env<NT>(id), k = id & (NT - 1);
        const bool valid = full || id < tasks;
        const uint32_t a0 = valid ? a0T[k * LANES + el] : 0u;
        const uint32_t my_ent = static_entry_offset(p, a0);
        if (i > 0) {  // the previous pass's store must be done reading the stage
            if (bulk && lane == 0) bulk_wait_read();
            __syncwarp();
        }
        if (full) static_coop_copy<VT, true>(p, my_ent, cnt, stage, lane);
        else static_coop_copy<VT, false>(p, my_ent, cnt, stage, lane);
        __syncwarp();
        if (valid) static_overlay_fast<VT, NT>(a0, k, a0T + el, colors, stage + lane * C::OS);
        int8_t *dst = p.obs + ((size_t)g.e0 * NT + (size_t)pass * LANES) * C::OS;
        if (bulk) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (p.l2hint & 2) bulk_s2g_hint(dst, stage, LANES * C::OS, policy);
                else bulk_s2g(dst, stage, LANES * C::OS);
                bulk_commit();
            }
        } else {
            __syncwarp();
            warp_copy(dst, stage, cnt * C::OS, lane);
            __syncwarp();
        }
        // fused one-hot image (MgStepOut.one_hot): expanded from the stage, under the TMA store
        if (p.one_hot)
            one_hot_emit_cold(p.one_hot, (uint32_t)(VT * VT), p.rcp_vv, stage, C::OS, (size_t)g.e0 * NT + (size_t)pass * LANES,
                              cnt, lane);
    }
    trace_mark(p, group * NWARP + warp, lane, 3);
    if (bulk && lane == 0) bulk_wait_read();  // smem must stay valid until the TMA store has read it
    trace_mark(p, group * NWARP + warp, lane, 4);
}

// Rolled shapes (any n <= 32, any odd V): one warp = one group of p.G envs, records in shared memory.
__global__ void __launch_bounds__(32, 24) static_rolled_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x, group = blockIdx.x;
    pdl_launch_dependents();
    const Group g = group_view(p, smem, group);
    const int n = p.n;
    const bool bulk = p.use_bulk && g.ne == p.G;
    const int env = lane < g.ne ? lane : -1;
    pdl_wait();
    static_load(p, g, lane);
    EnvRegs er;
    static_env_load(p, g, env, er);
    const OrderDraw draw = phase_draw<MODE_STEP_OBS>(p, g, env, er);
    __syncwarp();
    static_env_step(p, g, env, er, draw);
    __syncwarp();
    static_store_agents(p, g, lane);
    uint8_t *stage = g.stage;
    const int tasks = g.ne * n, passes = (tasks + LANES - 1) / LANES;
    for (int pass = 0; pass < passes; pass++) {
        if (pass > 0) {
            if (bulk && lane == 0) bulk_wait_read();
            __syncwarp();
        }
        static_obs_agent(p, g, pass, lane, stage);
        const int left = tasks - pass * LANES;
        const uint32_t cnt = left < LANES ? left : LANES;
        int8_t *dst = p.obs + ((size_t)g.e0 * n + (size_t)pass * LANES) * p.ostride;
        if (bulk) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(dst, stage, cnt * p.ostride);
                bulk_commit();
            }
        } else {
            __syncwarp();
            warp_copy(dst, stage, cnt * p.ostride, lane);
        }
        if (p.one_hot)
            one_hot_emit_cold(p.one_hot, (uint32_t)(p.V * p.V), p.rcp_vv, stage, p.ostride,
                              (size_t)g.e0 * n + (size_t)pass * LANES, (int)cnt, lane);
    }
    if (bulk && lane == 0) bulk_wait_read();
}
#endif

}  // namespace mg
