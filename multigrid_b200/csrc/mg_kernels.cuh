// multigrid_b200 -- device code of the batched MultiGrid step/observe engine (sm_100a).
//
// One WARP advances one group of G consecutive envs (G = 16 or 32) and never talks to another warp:
// no block-wide barrier anywhere, warps of a block drift apart and hide each other's latencies.
// A group's state is one contiguous HBM span per array (env-major layout); the spans are moved with
// 1-D TMA bulk copies (cp.async.bulk + mbarrier in, cp.async.bulk.bulk_group out), so loads and
// stores cost the warp a handful of instructions; the per-env work runs out of shared memory:
//
//   P0 load      TMA: cell words / agents / actions -> smem; pcg state+inc / step_count /
//                layout_idx -> registers of the env's lane (coalesced: lane <-> env)
//   P1 reset     auto-reset decision per env, pool layout fetch                     (1 lane / env)
//   P2 step      handle_actions: PCG64 draw, argsort, serial agent loop, rewards,
//                termination, dirty-cell write-through, env hook, agent stamping;
//                per-env outputs straight from registers to HBM                     (1 lane / env)
//   P3 observe   32 agents per pass: view gather (closed-form slice+rotate), row-bitmask
//                visibility scan, masking, 24->32 bit packing into the smem stage   (1 lane / agent)
//                -> TMA bulk store of the pass's contiguous obs span
//   P4 store     TMA: agents
//
// The grid lives in HBM as 32-bit CELL WORDS (type | color<<8 | state<<16 | opaque<<31) in a
// padded (W+1) x (H+1) array per env whose last row and column are WALL sentinels, i.e. exactly the
// layout the observation gather wants in shared memory: the TMA load is the whole "load" phase.
// mg_pack_grid / mg_unpack_grid convert from / to the reference's (W,H,3) byte layout.
//
// Reference semantics restated here (cited inline): multigrid/base.py:303-532,598-602,
// multigrid/utils/obs.py:46-316, multigrid/core/world_object.py:197-233,452-474,599-605.
//
// Variants of the one kernel (template parameters of step_obs_kernel): MODE (observe only / step only / fused),
// VT (unrolled view size or 0 = rolled loops), MULTI (mg_rollout: T steps per launch), CHAIN (MG_FLAG_CHAINED:
// per-env chain tickets instead of the kernel-boundary barrier). Every launch uses programmatic dependent
// launch. Further down: layout generation kernels (each env class's _gen_grid on in-kernel numpy generators),
// one-hot / feature / full-observation kernels, grid pack / unpack, reset_where.
//
// The file also compiles as plain C++ (no __CUDACC__): tests/hostsim runs the very same phase
// functions lane-by-lane on the CPU (a phase boundary == __syncwarp) to check the logic against the
// oracle without a GPU. That build is test infrastructure; the product only launches the CUDA kernels.
#pragma once
#include <stdint.h>
#include "multigrid_b200.h"

#ifdef __CUDACC__
#define MG_HD __host__ __device__ __forceinline__
#define MG_HD_COLD __host__ __device__ __noinline__  // rarely-run helpers: one copy, small I-cache footprint
#else
#define MG_HD inline
#define MG_HD_COLD inline
#endif

namespace mg {

enum : int { T_UNSEEN = 0, T_EMPTY, T_WALL, T_FLOOR, T_DOOR, T_KEY, T_BALL, T_BOX, T_GOAL, T_LAVA, T_AGENT };
enum : int { S_OPEN = 0, S_CLOSED, S_LOCKED };
enum : int { ACT_LEFT = 0, ACT_RIGHT, ACT_FORWARD, ACT_PICKUP, ACT_DROP, ACT_TOGGLE, ACT_DONE };
enum : int { MODE_OBS = 0, MODE_STEP = 1, MODE_STEP_OBS = 2 };

// 32-bit cell word: type | color<<8 | state<<16 | opaque<<31
constexpr uint32_t OPAQUE_BIT = 1u << 31;
// bit 24: the Door OBJECT of this cell is closed although the state byte (what grid.state shows and
// observations read) says open. Only RedBlueDoorsEnv.step creates this: it closes the blue door's
// object without grid.update() (envs/redbluedoors.py:185). Rules read the object, obs the array.
constexpr uint32_t DOOR_OBJ_CLOSED = 1u << 24;
constexpr uint32_t CELL_EMPTY = T_EMPTY;
constexpr uint32_t CELL_WALL = T_WALL | (5u << 8) | OPAQUE_BIT;  // WALL_ENCODING, utils/obs.py:14
constexpr int LANES = 32;

struct Params {
    // config
    int32_t W, H, n, V, max_steps;
    uint32_t flags;
    int32_t hook, hook_param, ostride, K, lstride;
    int32_t num_envs, G, wpb, use_bulk, generic_view;
    int32_t l2hint;  // bit 0: state TMA loads L2 evict_first; bit 1: obs TMA stores L2 evict_first
    // mg_rollout: T consecutive steps in ONE launch. Step t reads actions[t][E][n] and writes slice t of
    // every output array ([T][E]...); agents and the per-env scalars stay on chip between steps.
    int32_t T;
    int32_t pdl;    // host side only: launch with programmatic stream serialization
    int32_t num_sms;  // host side only: SMs of the launching device (picks the ROOMY instantiations)
    // Chained launches (MG_FLAG_CHAINED, MgState.chain_next / chain_done, see include/multigrid_b200.h):
    // per-env tickets order consecutive chained step launches on the same state env by env, so a launch need
    // not wait for the whole previous grid. chained: 0 = plain launch (tickets untouched), 1 = chained,
    // 2 = head of a chain (takes tickets AND waits for the whole previous grid of the stream).
    uint32_t *chain;  // [E][4] per-env record {next ticket, tickets done, grid dirty, reserved}: one 16-byte load
    int32_t chained;
    // Single-layout dedup (MgState.grid_dirty / pool_rep): with ONE pool layout (all deterministic env ids) an
    // env whose grid still equals it need not read its 324-byte copy from HBM: the group's cells come from a
    // small L2-resident buffer of 32 copies of the layout, unless one of the group's envs is marked dirty.
    // (the flag is word 2 of the env's chain record: 1 = the grid may differ from pool_grid[layout_idx])
    const uint32_t *pool_rep;  // [32][cstride] copies of pool layout 0, only set when K == 1 (else NULL)

    // Static-grid path (MG_FLAG_STATIC_GRID, mg_static.cuh): memoised per-(x, y, dir) views of the one layout
    const uint8_t *static_obs;  // [W*H*4][static_stride]
    const uint32_t *static_move;  // [W*H*4] memoised forward moves (behind the observation entries)
    int32_t static_stride, nstage;  // nstage: observation stages of a block (one per warp)
    int32_t lut_words;              // move words copied into the block's shared memory (0: read from global)
    int32_t no_lut;                 // host side only (knob MG_NO_LUT): never copy them

    int8_t *direction;  // [T][E][n] per-step 'direction' observation (rollout only; NULL otherwise)
    // state (device)
    uint32_t *grid; int8_t *agents; int32_t *step_count; uint64_t *pcg_state; const uint64_t *pcg_inc;
    int32_t *layout_idx; const uint32_t *pool_grid; const int8_t *pool_agents; int32_t *hook_state;
    const int8_t *actions;
    // outputs (device)
    int8_t *obs; double *reward; uint8_t *terminated; uint8_t *truncated; int32_t *status;
    uint8_t *one_hot;  // [E][n][V][V][21] fused OneHotObsWrapper image (MgStepOut.one_hot; NULL = not wanted)
    uint32_t rcp_vv;   // ceil(2^32 / (V*V)) (see fastdiv)
    unsigned long long *trace;  // diagnostics: 8 timestamps per warp (mg_debug_set_trace), normally NULL
    // derived geometry. The cell array of an env is (W+1) x (H+1) words, row stride Hp = H+1:
    // row x = W and column y = H hold WALL sentinels, every out-of-range coordinate maps there.
    int32_t Hp, cstride;
    uint32_t rcp_n;  // ceil(2^32 / n) (see fastdiv)
    // per-warp shared-memory carve-up (byte offsets, all multiples of 16)
    int32_t off_cells, off_stage, off_keys, off_ag, off_act, off_rk, off_mbar, warp_bytes;
    // obs stage aliased onto the cells of the passes already gathered (see carve_smem)
    int32_t alias, pass_cell_bytes, stage_bytes, stage_extra;
};

MG_HD int align16(int x) { return (x + 15) & ~15; }
inline uint32_t rcp32(int d) { return d <= 1 ? 0u : (uint32_t)((1ull << 32) / (uint32_t)d + 1ull); }

// Number of cell words per env in HBM and in shared memory.
inline int64_t cells_per_env(int W, int H) { return (int64_t)(W + 1) * (H + 1); }

// Fills the derived fields for group size p.G; returns the shared memory bytes of one warp.
inline int carve_smem(Params &p) {
    const int G = p.G, n = p.n;
    p.Hp = p.H + 1;
    p.cstride = (p.W + 1) * p.Hp;
    p.rcp_n = rcp32(n);
    const bool unrolled_view = !p.generic_view && (p.V == 3 || p.V == 5 || p.V == 7 || p.V == 9);
    // An obs pass (32 agent tasks = 32/n envs when n divides 32) only reads the cells of ITS envs,
    // and it has them in registers before it writes its packed result. So the stage of pass q can
    // live on top of the cells of passes <= q: the region is [extra][cells pass 0][cells pass 1]...
    // and stage q = the stage_bytes that END where the cells of pass q end.
    p.stage_bytes = LANES * p.ostride;
    p.pass_cell_bytes = (LANES / n) * p.cstride * 4;
    p.alias = unrolled_view && n <= LANES && LANES % n == 0 && (G * n) % LANES == 0 &&
              p.pass_cell_bytes % 16 == 0;
    p.stage_extra = 0;
    int off = 0;
    if (p.alias) {
        p.stage_extra = p.stage_bytes > p.pass_cell_bytes ? align16(p.stage_bytes - p.pass_cell_bytes) : 0;
        p.off_stage = off; off += p.stage_extra;
        p.off_cells = off; off += align16(G * p.cstride * 4);
    } else {
        p.off_cells = off; off += align16(G * p.cstride * 4);
        p.off_stage = off; off += align16(p.stage_bytes);
    }
    p.off_keys = off;  off += n > 4 ? align16(G * n * 8) + align16(G * n) : 0;  // sort keys + order
    p.off_ag = off;    off += align16(G * n * 8);
    p.off_act = off;   off += align16(G * n);
    p.off_rk = off;    off += align16(G * 4);
    p.off_mbar = off;  off += 16;
    p.warp_bytes = off;
    return off;
}

// Launch geometry: G envs per warp (16, or 32 when forced and it fits), wpb warps per block chosen
// to maximise resident warps per SM. Returns 0, or MG_ERR_TOO_LARGE when 16 envs do not fit.
inline int best_warps_per_sm(const Params &p, int smem_per_block, int smem_per_sm, int *wpb_out) {
    int best = 0, best_wpb = 1;
    for (int wpb = 4; wpb >= 1; wpb >>= 1) {
        const int bytes = wpb * p.warp_bytes;
        if (bytes > smem_per_block) continue;
        int blocks = smem_per_sm / (bytes + 16 + 1024);  // 16 B claim counter; 1 KB per block is reserved by the driver
        if (blocks > 32) blocks = 32;
        int warps = blocks * wpb;
        if (warps > 64) warps = 64;
        if (warps > best) { best = warps; best_wpb = wpb; }
    }
    *wpb_out = best_wpb;
    return best;
}

inline int plan_launch(Params &p, int forced_G, int forced_wpb, int smem_per_block, int smem_per_sm,
                       int num_sms = 148) {
    p.G = (forced_G == 32 || forced_G == 8) ? forced_G : 16;
    if (carve_smem(p) > smem_per_block) {
        p.G = 16;
        if (carve_smem(p) > smem_per_block) {
            p.G = 8;  // very large grids (up to ~83 x 83): 8 envs per warp, one warp per block
            if (carve_smem(p) > smem_per_block) return MG_ERR_TOO_LARGE;
            p.wpb = 1;
            return 0;
        }
    }
    int best_wpb = 1;
    const int warps16 = best_warps_per_sm(p, smem_per_block, smem_per_sm, &best_wpb);
    // Chained launches overlap two launches on the chip, so half the warps per launch (32 envs each, the
    // transition at full lane width: -20 % instructions) still fill it: 17.0 -> 16.7 us on the bench. Only for
    // batches of about a wave or more; smaller ones want the warps.
    if (forced_G == 0 && p.chained) {
        p.G = 32;
        int wpb32 = 1;
        const int warps32 = carve_smem(p) <= smem_per_block ? best_warps_per_sm(p, smem_per_block, smem_per_sm, &wpb32) : 0;
        if (warps32 >= 12 && (p.num_envs + 31) / 32 >= (3 * num_sms * warps32) / 4) {
            p.wpb = wpb32;
            if (forced_wpb > 0 && forced_wpb <= 4 && forced_wpb * p.warp_bytes <= smem_per_block) p.wpb = forced_wpb;
            return 0;
        }
        p.G = 16;
        carve_smem(p);
    }
    // Large grids / views (Empty-16x16 n=8 V=9: 19 KB per warp) leave few resident warps, and a modest
    // batch then fills only a fraction even of those: halve the group (twice the warps) as long as an
    // observation pass stays full (8 envs x n agents a multiple of 32). Measured 26.4 -> 24.4 us there.
    // The switch is made when the batch is less than one wave at G = 16 and still fits one wave at G = 8
    // (resident warps are also capped by registers: 128 per thread for the unrolled V = 9 view).
    const int reg_warps = (!p.generic_view && p.V == 9) ? 16 : 28;
    const int groups16 = (p.num_envs + 15) / 16;
    if (forced_G == 0 && p.G == 16 && (8 * p.n) % LANES == 0 &&
        groups16 < num_sms * (warps16 < reg_warps ? warps16 : reg_warps)) {
        p.G = 8;
        carve_smem(p);
        int wpb8 = 1;
        int warps8 = best_warps_per_sm(p, smem_per_block, smem_per_sm, &wpb8);
        if (warps8 > reg_warps) warps8 = reg_warps;
        if (2 * groups16 <= num_sms * warps8) {
            best_wpb = wpb8;
        } else {
            p.G = 16;
            carve_smem(p);
        }
    }
    p.wpb = best_wpb;
    if (forced_wpb > 0 && forced_wpb <= 4 && forced_wpb * p.warp_bytes <= smem_per_block) p.wpb = forced_wpb;
    return 0;
}

// ---- small portable intrinsics ----------------------------------------------------------------
MG_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef __CUDA_ARCH__
    return __byte_perm(a, b, sel);
#else
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
#endif
}

MG_HD uint32_t brev32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __brev(v);
#else
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
#endif
}

MG_HD uint32_t shl1_in(uint32_t acc, uint32_t w) {  // (acc << 1) | (w >> 31): one funnel shift
#ifdef __CUDA_ARCH__
    return __funnelshift_l(w, acc, 1);
#else
    return (acc << 1) | (w >> 31);
#endif
}

MG_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

MG_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// x / d for small x (x * d < 2^32) with rcp = rcp32(d); rcp == 0 encodes d == 1.
MG_HD uint32_t fastdiv(uint32_t x, uint32_t rcp) { return rcp ? mulhi32(x, rcp) : x; }

// numpy Generator(PCG64).random() (call site base.py:399): 128-bit LCG step + XSL-RR output;
// returns the top 53 bits (the double is key * 2^-53, so integer order == double order).
MG_HD uint64_t pcg64_next53(uint64_t &lo, uint64_t &hi, uint64_t inc_lo, uint64_t inc_hi) {
    const uint64_t M_HI = 0x2360ED051FC65DA4ull, M_LO = 0x4385DF649FCCF645ull;
    uint64_t nlo = lo * M_LO;
    uint64_t nhi = mulhi64(lo, M_LO) + lo * M_HI + hi * M_LO;
    uint64_t slo = nlo + inc_lo;
    nhi += inc_hi + (slo < nlo ? 1ull : 0ull);
    lo = slo; hi = nhi;
    uint64_t x = hi ^ lo;
    unsigned rot = (unsigned)(hi >> 58);
    uint64_t out = (x >> rot) | (x << ((64u - rot) & 63u));
    return out >> 11;
}

MG_HD uint32_t cell_word(uint32_t t, uint32_t c, uint32_t s) {
    // see_behind (utils/obs.py:47-63): walls and non-open doors block the view
    const uint32_t opaque = (uint32_t)(t == T_WALL) | (uint32_t)((t == T_DOOR) & (s != S_OPEN));
    return t | (c << 8) | (s << 16) | (opaque << 31);
}

MG_HD uint32_t cell_word24(uint32_t c) {  // c = type | color<<8 | state<<16
    const uint32_t t = c & 0xffu;
    const uint32_t opaque = (uint32_t)(t == T_WALL) | (uint32_t)((t == T_DOOR) & ((c >> 16) != S_OPEN));
    return c | (opaque << 31);
}

// ---- numpy Generator(PCG64).integers(low, high) as RandomMixin._rand_int calls it -----------------
// (utils/random.py:23-38 -> numpy/random/_bounded_integers.pyx _rand_int64 -> distributions.c
// random_bounded_uint64_fill: ranges below 2^32 take buffered_bounded_lemire_uint32 on the bit
// generator's 32-bit stream). numpy is a third-party dependency of the reference (pyproject.toml:30,
// unpinned; 2.3.5 here): restated from its published algorithm and pinned against numpy itself.
struct LayoutRng {
    uint64_t lo, hi, ilo, ihi;  // PCG64 state / increment
    uint32_t has32, buf32;      // pcg64_next32's buffered upper half (numpy/random/src/pcg64/pcg64.h)
};

MG_HD uint64_t pcg64_next64(LayoutRng &g) {
    const uint64_t M_HI = 0x2360ED051FC65DA4ull, M_LO = 0x4385DF649FCCF645ull;
    uint64_t nlo = g.lo * M_LO;
    uint64_t nhi = mulhi64(g.lo, M_LO) + g.lo * M_HI + g.hi * M_LO;
    uint64_t slo = nlo + g.ilo;
    nhi += g.ihi + (slo < nlo ? 1ull : 0ull);
    g.lo = slo; g.hi = nhi;
    const uint64_t x = g.hi ^ g.lo;
    const unsigned rot = (unsigned)(g.hi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}

MG_HD uint32_t pcg64_next32(LayoutRng &g) {  // low half first, the high half is kept for the next call
    if (g.has32) { g.has32 = 0; return g.buf32; }
    const uint64_t v = pcg64_next64(g);
    g.has32 = 1; g.buf32 = (uint32_t)(v >> 32);
    return (uint32_t)v;
}

// Generator.integers(low, high) for 0 < high - low <= 2^32 (Lemire's nearly divisionless bounded draw)
MG_HD int32_t rng_integers(LayoutRng &g, int32_t low, int32_t high) {
    const uint32_t rng = (uint32_t)(high - low - 1);
    if (rng == 0) return low;  // no draw (random_bounded_uint64_fill, rng == 0)
    if (rng == 0xffffffffu) return low + (int32_t)pcg64_next32(g);
    const uint32_t rng_excl = rng + 1u;
    uint64_t m = (uint64_t)pcg64_next32(g) * rng_excl;
    uint32_t leftover = (uint32_t)m;
    if (leftover < rng_excl) {
        const uint32_t threshold = (0xffffffffu - rng) % rng_excl;
        while (leftover < threshold) {
            m = (uint64_t)pcg64_next32(g) * rng_excl;
            leftover = (uint32_t)m;
        }
    }
    return low + (int32_t)(m >> 32);
}

// numpy's random_interval (distributions.c): uniform in [0, max] by masked rejection on the 32-bit stream;
// Generator.shuffle of a Python list (the untyped path of _generator.pyx) is Fisher-Yates with it:
// for i = n-1 .. 1: j = random_interval(i); swap(x[i], x[j]). RandomMixin._rand_perm (utils/random.py:77-85).
MG_HD uint32_t rng_interval(LayoutRng &g, uint32_t max) {
    if (max == 0) return 0;
    uint32_t mask = max;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    do { v = pcg64_next32(g) & mask; } while (v > max);
    return v;
}

MG_HD void rng_shuffle(LayoutRng &g, uint8_t *x, int n) {
    for (int i = n - 1; i >= 1; i--) {
        const uint32_t j = rng_interval(g, (uint32_t)i);
        const uint8_t t = x[i]; x[i] = x[j]; x[j] = t;
    }
}

// EmptyEnv._gen_grid with random agent placement (envs/empty.py:151-170): wall ring, goal at (W-2,H-2),
// then per agent place_agent (base.py:672-697) = place_obj's rejection sampling (base.py:604-655: a cell
// is taken when it is empty and NO agent record -- placed or not -- has that position) followed by
// dir = _rand_int(0, 4). Writes one layout: padded cell words and packed agent records.
// Returns false when the sampling gave up (the reference would loop forever).
MG_HD bool gen_layout_empty_random(int W, int H, int n, LayoutRng &g, uint32_t *cells, int8_t *agents) {
    const int Hp = H + 1;
    for (int x = 0; x <= W; x++)
        for (int y = 0; y <= H; y++) {
            const bool wall = x == 0 || y == 0 || x >= W - 1 || y >= H - 1;
            cells[x * Hp + y] = wall ? CELL_WALL : CELL_EMPTY;
        }
    if (W >= 3 && H >= 3) cells[(W - 2) * Hp + (H - 2)] = T_GOAL | (1u << 8);  // Goal() is green
    for (int j = 0; j < n; j++) {
        int8_t *a = agents + j * 8;
        a[0] = -1; a[1] = -1; a[2] = -1; a[3] = 0; a[4] = T_EMPTY; a[5] = 0; a[6] = 0; a[7] = (int8_t)(j % 6);
    }
    for (int j = 0; j < n; j++) {
        int x = -1, y = -1;
        for (int tries = 0;; tries++) {
            if (tries >= (1 << 16)) return false;
            x = rng_integers(g, 0, W);
            y = rng_integers(g, 0, H);
            if ((cells[x * Hp + y] & 0xffu) != T_EMPTY) continue;   // grid.get(*pos) is not None
            bool taken = false;
            for (int q = 0; q < n; q++) taken |= agents[q * 8 + 1] == x && agents[q * 8 + 2] == y;
            if (!taken) break;
        }
        agents[j * 8 + 1] = (int8_t)x; agents[j * 8 + 2] = (int8_t)y;
        agents[j * 8 + 0] = (int8_t)rng_integers(g, 0, 4);
    }
    return true;
}

// ---- RoomGrid pieces (core/roomgrid.py) for on-device layouts ---------------------------------------
// place_obj (base.py:604-655) inside the rectangle [tx, tx+sx) x [ty, ty+sy) clipped to the grid, with
// max_tries = 1000 as every RoomGrid caller passes (roomgrid.py:365, 398): -1 = gave up (the reference
// raises RecursionError). `next_to_agents` = reject_next_to (roomgrid.py:46-51): no agent within
// Euclidean distance 1 of the position. Returns x | y << 8.
MG_HD int place_in_rect(int W, int H, int n, LayoutRng &g, const uint32_t *cells, const int8_t *agents,
                        int tx, int ty, int sx, int sy, bool next_to_agents, int max_tries = 1000) {
    const int Hp = H + 1;
    const int hx = tx + sx < W ? tx + sx : W, hy = ty + sy < H ? ty + sy : H;
    for (int tries = 0;; tries++) {
        if (tries > max_tries) return -1;
        const int x = rng_integers(g, tx, hx), y = rng_integers(g, ty, hy);
        if ((cells[x * Hp + y] & 0xffu) != T_EMPTY) continue;
        bool bad = false;
        for (int q = 0; q < n; q++) {
            const int dx = agents[q * 8 + 1] - x, dy = agents[q * 8 + 2] - y;
            bad |= (dx == 0 && dy == 0) || (next_to_agents && dx * dx + dy * dy <= 1);
        }
        if (!bad) return x | (y << 8);
    }
}

// BlockedUnlockPickupEnv._gen_grid (envs/blockedunlockpickup.py:142-164) on a 1 x 2 RoomGrid
// (core/roomgrid.py:203-236: rooms share walls, agents start in the middle of room (1,0) facing right):
// box of a random colour in the right room, locked door of a random colour in the shared wall at a height
// drawn from the ORDER generator (env.np_random, roomgrid.py:324 -> :106-124), a ball of a random colour
// in front of it, the door's key in the left room, then every agent in the left room, not facing an object
// (place_in_room, roomgrid.py:387-404). Returns the box colour (the mission names it), -1 = gave up.
MG_HD int gen_layout_bup(int S, int n, LayoutRng &g, LayoutRng &order, uint32_t *cells, int8_t *agents) {
    const int W = 2 * (S - 1) + 1, H = S, Hp = H + 1, step = S - 1;
    for (int x = 0; x <= W; x++)
        for (int y = 0; y <= H; y++) {
            const bool wall = x >= W || y >= H || y == 0 || y == H - 1 || x % step == 0;
            cells[x * Hp + y] = wall ? CELL_WALL : CELL_EMPTY;
        }
    for (int j = 0; j < n; j++) {
        int8_t *a = agents + j * 8;
        a[0] = 0; a[1] = (int8_t)(step + S / 2); a[2] = (int8_t)(S / 2); a[3] = 0;
        a[4] = T_EMPTY; a[5] = 0; a[6] = 0; a[7] = (int8_t)(j % 6);
    }
    // add_object(1, 0, kind=box): colour, then position (roomgrid.py:338-368)
    const uint32_t box_color = (uint32_t)rng_integers(g, 0, 6);
    int pos = place_in_rect(W, H, n, g, cells, agents, step, 0, S, S, true);
    if (pos < 0) return -1;
    cells[(pos & 0xff) * Hp + (pos >> 8)] = cell_word(T_BOX, box_color, 0);
    // add_door(0, 0, right, locked=True): colour from the layout generator, height from the order generator
    const uint32_t door_color = (uint32_t)rng_integers(g, 0, 6);
    const int door_y = rng_integers(order, 1, S - 1);
    cells[step * Hp + door_y] = cell_word(T_DOOR, door_color, S_LOCKED);
    // a ball blocks the door (blockedunlockpickup.py:155)
    const uint32_t ball_color = (uint32_t)rng_integers(g, 0, 6);
    cells[(step - 1) * Hp + door_y] = cell_word(T_BALL, ball_color, 0);
    // add_object(0, 0, key, door colour)
    pos = place_in_rect(W, H, n, g, cells, agents, 0, 0, S, S, true);
    if (pos < 0) return -1;
    cells[(pos & 0xff) * Hp + (pos >> 8)] = cell_word(T_KEY, door_color, 0);
    // place_agent(top, size) until the agent does not face an object
    for (int j = 0; j < n; j++) {
        for (;;) {
            agents[j * 8 + 1] = -1; agents[j * 8 + 2] = -1;
            pos = place_in_rect(W, H, n, g, cells, agents, 0, 0, S, S, false);
            if (pos < 0) return -1;
            const int x = pos & 0xff, y = pos >> 8, dir = rng_integers(g, 0, 4);
            agents[j * 8 + 1] = (int8_t)x; agents[j * 8 + 2] = (int8_t)y; agents[j * 8] = (int8_t)dir;
            const int fx = x + (dir == 0) - (dir == 2), fy = y + (dir == 1) - (dir == 3);
            const uint32_t t = cells[fx * Hp + fy] & 0xffu;
            if (t == T_EMPTY || t == T_WALL) break;
        }
    }
    return (int)box_color;
}

// RedBlueDoorsEnv._gen_grid (envs/redbluedoors.py:142-168): a (2*size) x size grid, a walled room in the
// middle half, every agent placed in it (place_agent with the default unlimited tries; capped at 65 536
// here), then a closed red door in the room's left wall and a closed blue door in its right wall at
// heights drawn from the layout generator. false = a placement gave up.
MG_HD bool gen_layout_red_blue_doors(int size, int n, LayoutRng &g, uint32_t *cells, int8_t *agents) {
    const int W = 2 * size, H = size, Hp = H + 1, rx = W / 4, rw = W / 2;
    for (int x = 0; x <= W; x++)
        for (int y = 0; y <= H; y++) {
            const bool wall = x >= W - 1 || y >= H - 1 || x == 0 || y == 0 || x == rx || x == rx + rw - 1;
            cells[x * Hp + y] = wall ? CELL_WALL : CELL_EMPTY;
        }
    for (int j = 0; j < n; j++) {
        int8_t *a = agents + j * 8;
        a[0] = -1; a[1] = -1; a[2] = -1; a[3] = 0; a[4] = T_EMPTY; a[5] = 0; a[6] = 0; a[7] = (int8_t)(j % 6);
    }
    for (int j = 0; j < n; j++) {
        const int pos = place_in_rect(W, H, n, g, cells, agents, rx, 0, rw, H, false, 1 << 16);
        if (pos < 0) return false;
        agents[j * 8 + 1] = (int8_t)(pos & 0xff); agents[j * 8 + 2] = (int8_t)(pos >> 8);
        agents[j * 8] = (int8_t)rng_integers(g, 0, 4);
    }
    int y = rng_integers(g, 1, H - 1);
    cells[rx * Hp + y] = cell_word(T_DOOR, 0, S_CLOSED);             // red
    y = rng_integers(g, 1, H - 1);
    cells[(rx + rw - 1) * Hp + y] = cell_word(T_DOOR, 2, S_CLOSED);  // blue
    return true;
}

// LockedHallwayEnv._gen_grid (envs/locked_hallway.py:150-194) on a (num_rooms/2) x 3 RoomGrid: a shuffled
// colour sequence, the hallway column opened vertically, one locked door per side room (colours from a second
// shuffle, popped from the end; doors sit in the middle of the wall), 1..max_hallway_keys keys in the hallway,
// then for every further key 1..max_keys_per_room keys inside the room the previous key opens, and every agent
// in the hallway (MultiGridEnv.place_agent: unlimited tries; capped at 65 536 here). num_rooms even, <= 6.
MG_HD bool gen_layout_locked_hallway(int num_rooms, int S, int max_hallway_keys, int max_keys_per_room, int n,
                                     LayoutRng &g, uint32_t *cells, int8_t *agents) {
    const int step = S - 1, num_rows = num_rooms / 2, W = 3 * step + 1, H = num_rows * step + 1, Hp = H + 1;
    const int INF = 1 << 16;
    for (int x = 0; x <= W; x++)
        for (int y = 0; y <= H; y++) {
            const bool wall = x >= W || y >= H || x % step == 0 || y % step == 0;
            cells[x * Hp + y] = wall ? CELL_WALL : CELL_EMPTY;
        }
    for (int j = 0; j < n; j++) {  // RoomGrid: agents start in the middle room facing right (roomgrid.py:231-236)
        int8_t *a = agents + j * 8;
        a[0] = 0; a[1] = (int8_t)(step + S / 2); a[2] = (int8_t)((num_rows / 2) * step + S / 2); a[3] = 0;
        a[4] = T_EMPTY; a[5] = 0; a[6] = 0; a[7] = (int8_t)(j % 6);
    }
    uint8_t seq[6] = {0, 1, 2, 3, 4, 5}, doors[6];   // list(Color) * ceil(num_rooms / 6), num_rooms <= 6
    rng_shuffle(g, seq, 6);                            // color_sequence = _rand_perm(...)[:num_rooms]
    for (int row = 0; row + 1 < num_rows; row++)       // remove_wall(HALLWAY, row, down)
        for (int x = step + 1; x <= 2 * step - 1; x++) cells[x * Hp + (row + 1) * step] = CELL_EMPTY;
    for (int q = 0; q < num_rooms; q++) doors[q] = seq[q];
    rng_shuffle(g, doors, num_rooms);                  // door_colors = _rand_perm(color_sequence)
    int room_of_color[6] = {0, 0, 0, 0, 0, 0};         // colour -> (col, row) packed col | row << 4
    int left = num_rooms;
    for (int row = 0; row < num_rows; row++)
        for (int side = 0; side < 2; side++) {         // (LEFT, right) then (RIGHT, left)
            const uint32_t color = doors[--left];      // door_colors.pop()
            const int col = side == 0 ? 0 : 2;
            room_of_color[color] = col | (row << 4);
            const int dx = side == 0 ? step : 2 * step, dy = row * step + (S - 1) / 2;
            cells[dx * Hp + dy] = cell_word(T_DOOR, color, S_LOCKED);
        }
    const int hallway_keys = rng_integers(g, 1, max_hallway_keys + 1);
    for (int q = 0; q < hallway_keys && q < num_rooms; q++) {
        const int pos = place_in_rect(W, H, n, g, cells, agents, step, 0, S, H, false, INF);
        if (pos < 0) return false;
        cells[(pos & 0xff) * Hp + (pos >> 8)] = cell_word(T_KEY, seq[q], 0);
    }
    int key_index = hallway_keys;
    while (key_index < num_rooms) {
        const int room = room_of_color[seq[key_index - 1]], tx = (room & 15) * step, ty = (room >> 4) * step;
        const int room_keys = rng_integers(g, 1, max_keys_per_room + 1);
        for (int q = 0; q < room_keys && key_index < num_rooms; q++) {  // color_sequence[key_index : key_index + k]
            const int pos = place_in_rect(W, H, n, g, cells, agents, tx, ty, S, S, false, INF);
            if (pos < 0) return false;
            cells[(pos & 0xff) * Hp + (pos >> 8)] = cell_word(T_KEY, seq[key_index], 0);
            key_index++;
        }
    }
    for (int j = 0; j < n; j++) {
        agents[j * 8 + 1] = -1; agents[j * 8 + 2] = -1;
        const int pos = place_in_rect(W, H, n, g, cells, agents, step, 0, S, H, false, INF);
        if (pos < 0) return false;
        agents[j * 8 + 1] = (int8_t)(pos & 0xff); agents[j * 8 + 2] = (int8_t)(pos >> 8);
        agents[j * 8] = (int8_t)rng_integers(g, 0, 4);
    }
    return true;
}

// PlaygroundEnv._gen_grid (envs/playground.py:122-137) on a rows x cols RoomGrid (3 x 3 rooms of 7):
// connect_all (roomgrid.py:406-452: until a search from room (0,0) over rooms joined by doors reaches every
// room, draw a room and a direction and, if there is a neighbour without a door, add an unlocked door of a
// random colour at a position drawn from the ORDER generator), 12 x add_object in a random room with random
// kind (key, ball, box) and colour, then every agent in a random room, not facing an object.
// rows * cols <= 16. false = a placement or connect_all gave up.
MG_HD bool gen_layout_playground(int S, int rows, int cols, int n, LayoutRng &g, LayoutRng &order,
                                 uint32_t *cells, int8_t *agents) {
    const int step = S - 1, W = cols * step + 1, H = rows * step + 1, Hp = H + 1, total = rows * cols;
    for (int x = 0; x <= W; x++)
        for (int y = 0; y <= H; y++) {
            const bool wall = x >= W || y >= H || x % step == 0 || y % step == 0;
            cells[x * Hp + y] = wall ? CELL_WALL : CELL_EMPTY;
        }
    for (int j = 0; j < n; j++) {
        int8_t *a = agents + j * 8;
        a[0] = 0; a[1] = (int8_t)((cols / 2) * step + S / 2); a[2] = (int8_t)((rows / 2) * step + S / 2); a[3] = 0;
        a[4] = T_EMPTY; a[5] = 0; a[6] = 0; a[7] = (int8_t)(j % 6);
    }
    uint8_t doors[16];  // per room (index row * cols + col): bit d = a door in direction d (right, down, left, up)
    for (int q = 0; q < total; q++) doors[q] = 0;
    const int DC[4] = {1, 0, -1, 0}, DR[4] = {0, 1, 0, -1};
    bool connected = false;
    for (int it = 0; it < 5000; it++) {
        uint32_t seen = 1u, frontier = 1u;  // reachability from room (0,0) through doors
        while (frontier) {
            uint32_t next = 0;
            for (int q = 0; q < total; q++)
                if ((frontier >> q) & 1u)
                    for (int d = 0; d < 4; d++)
                        if ((doors[q] >> d) & 1u) next |= 1u << (q + DR[d] * cols + DC[d]);
            frontier = next & ~seen;
            seen |= next;
        }
        if (seen == (total >= 32 ? 0xffffffffu : (1u << total) - 1u)) { connected = true; break; }
        const int col = rng_integers(g, 0, cols), row = rng_integers(g, 0, rows), d = rng_integers(g, 0, 4);
        const int oc = col + DC[d], orow = row + DR[d];
        if (oc < 0 || oc >= cols || orow < 0 || orow >= rows) continue;  // no neighbour
        if ((doors[row * cols + col] >> d) & 1u) continue;                // door already there
        const uint32_t color = (uint32_t)rng_integers(g, 0, 6);           // rand_elem(door_colors)
        const int left = col * step, top = row * step, right = left + S - 1, bottom = top + S - 1;
        int dx, dy;                                                       // Room.set_door_pos, roomgrid.py:87-124
        if (d == 0) { dx = right; dy = rng_integers(order, top + 1, bottom); }
        else if (d == 1) { dx = rng_integers(order, left + 1, right); dy = bottom; }
        else if (d == 2) { dx = left; dy = rng_integers(order, top + 1, bottom); }
        else { dx = rng_integers(order, left + 1, right); dy = top; }
        cells[dx * Hp + dy] = cell_word(T_DOOR, color, S_CLOSED);
        doors[row * cols + col] |= (uint8_t)(1u << d);
        doors[orow * cols + oc] |= (uint8_t)(1u << ((d + 2) & 3));
    }
    if (!connected) return false;
    for (int q = 0; q < 12; q++) {  // playground.py:128-131
        const int col = rng_integers(g, 0, cols), row = rng_integers(g, 0, rows);
        const uint32_t kind = (uint32_t)(T_KEY + rng_integers(g, 0, 3)), color = (uint32_t)rng_integers(g, 0, 6);
        const int pos = place_in_rect(W, H, n, g, cells, agents, col * step, row * step, S, S, true);
        if (pos < 0) return false;
        cells[(pos & 0xff) * Hp + (pos >> 8)] = cell_word(kind, color, 0);
    }
    for (int j = 0; j < n; j++) {   // place_agent() -> place_in_room in a random room (roomgrid.py:370-404)
        const int col = rng_integers(g, 0, cols), row = rng_integers(g, 0, rows);
        for (;;) {
            agents[j * 8 + 1] = -1; agents[j * 8 + 2] = -1;
            const int pos = place_in_rect(W, H, n, g, cells, agents, col * step, row * step, S, S, false);
            if (pos < 0) return false;
            const int x = pos & 0xff, y = pos >> 8, dir = rng_integers(g, 0, 4);
            agents[j * 8 + 1] = (int8_t)x; agents[j * 8 + 2] = (int8_t)y; agents[j * 8] = (int8_t)dir;
            const uint32_t t = cells[(x + (dir == 0) - (dir == 2)) * Hp + y + (dir == 1) - (dir == 3)] & 0xffu;
            if (t == T_EMPTY || t == T_WALL) break;
        }
    }
    return true;
}

// ---- fresh layouts on auto-reset -------------------------------------------------------------------------
// The reference draws a NEW layout at every reset() from the env's own RandomMixin generator (base.py:250-301 ->
// _gen_grid). With one pool slot per env (K = E, layout_idx[e] = e, layout_stride = 0) the same happens here:
// after a step, every env that will auto-reset at its next step (the predicate of phase_reset) gets its slot
// regenerated from ITS generator, which advances; door positions come from a COPY of the env's order stream
// (roomgrid.py:324 draws them from env.np_random; a reset does not advance that stream in this engine, and the
// fixtures recorded from the reference restore it the same way, tests/golden/make_golden.py).
enum : int { LAYOUT_EMPTY_RANDOM = 1, LAYOUT_BUP = 2, LAYOUT_RED_BLUE_DOORS = 3, LAYOUT_LOCKED_HALLWAY = 4, LAYOUT_PLAYGROUND = 5 };

struct LayoutGen {
    int32_t family, a, b, c, d;  // family parameters: size | room_size | (num_rooms, room_size, max_hallway_keys,
                                 // max_keys_per_room) | (room_size, num_rows, num_cols)
    uint64_t *rng_state; const uint64_t *rng_inc; uint64_t *rng_buf;  // [E] the envs' layout generators
    const uint64_t *order_buf;   // [E] buffered 32-bit half of the envs' order streams (may be NULL = empty)
    int32_t *info;               // [E] BlockedUnlockPickup: box colour of the slot's layout (may be NULL)
};

// Will env e be reset by the next step launch? (phase_reset's predicate; is_done, base.py:534-539)
MG_HD bool env_is_done(const Params &p, size_t e) {
    const uint32_t *ag = (const uint32_t *)(p.agents + e * p.n * 8);
    uint32_t all_term = 1;
    for (int j = 0; j < p.n; j++) all_term &= ((ag[j * 2] >> 24) & 0xff) != 0;
    if (p.hook == MG_HOOK_LOCKED_HALLWAY && __builtin_popcount((unsigned)p.hook_state[e]) == p.hook_param) all_term = 1;
    return all_term || p.step_count[e] >= p.max_steps;
}

// Regenerates pool slot e from generator e. Returns false when a placement gave up.
MG_HD bool refresh_slot(const Params &p, const LayoutGen &lg, size_t e) {
    LayoutRng g, o;
    g.lo = lg.rng_state[2 * e]; g.hi = lg.rng_state[2 * e + 1]; g.ilo = lg.rng_inc[2 * e]; g.ihi = lg.rng_inc[2 * e + 1];
    const uint64_t b = lg.rng_buf ? lg.rng_buf[e] : 0ull;
    g.has32 = (uint32_t)(b >> 32) & 1u; g.buf32 = (uint32_t)b;
    o.lo = o.hi = o.ilo = o.ihi = 0; o.has32 = 0; o.buf32 = 0;
    if (p.pcg_state) {  // (a copy: the env's own stream is not advanced, neither is its buffered half)
        o.lo = p.pcg_state[2 * e]; o.hi = p.pcg_state[2 * e + 1]; o.ilo = p.pcg_inc[2 * e]; o.ihi = p.pcg_inc[2 * e + 1];
        const uint64_t ob = lg.order_buf ? lg.order_buf[e] : 0ull;
        o.has32 = (uint32_t)(ob >> 32) & 1u; o.buf32 = (uint32_t)ob;
    }
    uint32_t *cells = const_cast<uint32_t *>(p.pool_grid) + e * (size_t)p.cstride;
    int8_t *agents = const_cast<int8_t *>(p.pool_agents) + e * (size_t)p.n * 8;
    bool ok = true;
    if (lg.family == LAYOUT_EMPTY_RANDOM) ok = gen_layout_empty_random(p.W, p.H, p.n, g, cells, agents);
    else if (lg.family == LAYOUT_RED_BLUE_DOORS) ok = gen_layout_red_blue_doors(lg.a, p.n, g, cells, agents);
    else if (lg.family == LAYOUT_LOCKED_HALLWAY) ok = gen_layout_locked_hallway(lg.a, lg.b, lg.c, lg.d, p.n, g, cells, agents);
    else if (lg.family == LAYOUT_PLAYGROUND) ok = gen_layout_playground(lg.a, lg.b, lg.c, p.n, g, o, cells, agents);
    else if (lg.family == LAYOUT_BUP) {
        const int color = gen_layout_bup(lg.a, p.n, g, o, cells, agents);
        ok = color >= 0;
        if (lg.info) lg.info[e] = color;
    }
    lg.rng_state[2 * e] = g.lo; lg.rng_state[2 * e + 1] = g.hi;
    if (lg.rng_buf) lg.rng_buf[e] = ((uint64_t)g.has32 << 32) | g.buf32;
    return ok;
}

// base.py:598-602: `1 - 0.9 * (step_count / max_steps)` in float64, round-to-nearest at every
// operation, never contracted into an FMA.
MG_HD double reward_value(int32_t step_count, int32_t max_steps) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(1.0, __dmul_rn(0.9, __ddiv_rn((double)step_count, (double)max_steps)));
#else
    volatile double ratio = (double)step_count / (double)max_steps;
    volatile double scaled = 0.9 * ratio;
    return 1.0 - scaled;
#endif
}

MG_HD void status_or(int32_t *status, int32_t v) {
    if (!status) return;
#ifdef __CUDA_ARCH__
    atomicOr(status, v);
#else
    *status |= v;
#endif
}

// Warp-cooperative plain copy of a contiguous span (the non-TMA path: ragged tail groups, and
// everything under tests/hostsim). 16-byte vectors when both ends allow it, else 4-byte, else bytes.
MG_HD_COLD void warp_copy(void *dst, const void *src, int nbytes, int lane) {
    const uintptr_t both = (uintptr_t)dst | (uintptr_t)src;
    if ((both & 15u) == 0) {
        const int nv = nbytes >> 4;
#ifdef __CUDACC__
        uint4 *d = (uint4 *)dst; const uint4 *s = (const uint4 *)src;
#else
        struct alignas(16) V16 { uint32_t v[4]; };
        V16 *d = (V16 *)dst; const V16 *s = (const V16 *)src;
#endif
#pragma unroll 1
        for (int v = lane; v < nv; v += LANES) d[v] = s[v];
#pragma unroll 1
        for (int b = (nv << 4) + lane; b < nbytes; b += LANES) ((uint8_t *)dst)[b] = ((const uint8_t *)src)[b];
    } else if ((both & 3u) == 0) {
        const int nw = nbytes >> 2;
#pragma unroll 1
        for (int v = lane; v < nw; v += LANES) ((uint32_t *)dst)[v] = ((const uint32_t *)src)[v];
#pragma unroll 1
        for (int b = (nw << 2) + lane; b < nbytes; b += LANES) ((uint8_t *)dst)[b] = ((const uint8_t *)src)[b];
    } else {
#pragma unroll 1
        for (int b = lane; b < nbytes; b += LANES) ((uint8_t *)dst)[b] = ((const uint8_t *)src)[b];
    }
}

// 4 consecutive 3-byte cells (12 bytes = words w0,w1,w2) -> 4 cell words. The opaque test runs on
// all four cells at once, one byte lane per cell (all encoded values are < 0x80).
MG_HD void cell_words_x4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t out[4]) {
    const uint32_t T = byte_perm(byte_perm(w0, w1, 0x0630u), w2, 0x5210u);  // type bytes: 0,3,6,9
    const uint32_t S = byte_perm(byte_perm(w0, w1, 0x0052u), w2, 0x7410u);  // state bytes: 2,5,8,11
    // x + 0x7f sets bit 7 of a byte lane iff x != 0 (x < 0x80: no carry between lanes)
    const uint32_t K = 0x7f7f7f7fu;
    const uint32_t not_wall = (T ^ 0x02020202u) + K, not_door = (T ^ 0x04040404u) + K, not_open = S + K;
    const uint32_t O = (~not_wall | (~not_door & not_open)) & 0x80808080u;  // 0x80 per opaque cell
    out[0] = byte_perm(w0, O, 0x4210u);
    out[1] = byte_perm(byte_perm(w0, w1, 0x0543u), O, 0x5210u);
    out[2] = byte_perm(byte_perm(w1, w2, 0x0432u), O, 0x6210u);
    out[3] = byte_perm(w2, O, 0x7321u);
}

// One warp's view of its group: e0 = first global env, ne = number of valid envs.
struct Group {
    int e0, ne;
    uint32_t *cells; uint8_t *stage; uint32_t *ag; int8_t *act; int32_t *rk;
    uint64_t *keys; uint8_t *order;  // n > 4 only
};

MG_HD Group group_view(const Params &p, uint8_t *ws, int group) {
    Group g;
    g.e0 = group * p.G;
    g.ne = p.num_envs - g.e0 < p.G ? p.num_envs - g.e0 : p.G;
    g.cells = (uint32_t *)(ws + p.off_cells);
    g.stage = ws + p.off_stage;  // pass 0's (see stage_of)
    g.ag = (uint32_t *)(ws + p.off_ag);
    g.act = (int8_t *)(ws + p.off_act);
    g.rk = (int32_t *)(ws + p.off_rk);
    g.keys = (uint64_t *)(ws + p.off_keys);
    g.order = ws + p.off_keys + align16(p.G * p.n * 8);
    return g;
}

// The per-env phases (reset decision, transition) run one lane per env: lanes >= G (G = 16 or 8) sit them out
// and rejoin at the __syncwarp that ends the phase. (Round 1 let them SHADOW the lower lanes -- same work on the
// same shared-memory words -- which compute-sanitizer racecheck rightly reports as intra-warp hazards. Without
// the shadows the hardware keeps the two halves of a G = 16 warp apart for the rest of the kernel, i.e. the
// observation passes issue twice at half width: 16.5 -> 23.4 us on a 65 536-env 4-agent batch that takes this
// kernel, no change for 2-agent batches such as BlockedUnlockPickup. Neither an explicit bar.warp.sync nor hiding
// the branch condition from the compiler brings them back together; grids no action can change -- the Empty
// family -- take the static-grid kernel of mg_static.cuh, which has a single divergent region and reconverges.)
MG_HD int lane_env(const Params &p, const Group &g, int lane) {
    const int i = lane < p.G ? lane : p.G;
    return i < g.ne ? i : -1;
}

// Per-env scalars the env's lane keeps in registers from load to store.
struct alignas(16) U128 { uint64_t lo, hi; };
struct EnvRegs {
    uint64_t lo, hi, ilo, ihi;  // numpy PCG64 state / increment
    int32_t sc, lidx, hs;       // step_count, layout cursor, post-hook state
};

// ---- P0: load ------------------------------------------------------------------------------------
template <int MODE>
MG_HD void env_load(const Params &p, const Group &g, int i, EnvRegs &r) {
    r.lo = r.hi = r.ilo = r.ihi = 0; r.sc = 0; r.lidx = 0; r.hs = 0;
    if (MODE == MODE_OBS || i < 0) return;
    const size_t e = (size_t)(g.e0 + i);
    r.sc = p.step_count[e];
    if (p.n > 1) {
        const U128 s = *(const U128 *)(p.pcg_state + 2 * e), c = *(const U128 *)(p.pcg_inc + 2 * e);
        r.lo = s.lo; r.hi = s.hi; r.ilo = c.lo; r.ihi = c.hi;
    }
    if (p.flags & MG_FLAG_AUTO_RESET) r.lidx = p.layout_idx[e];
    if (p.hook == MG_HOOK_LOCKED_HALLWAY) r.hs = p.hook_state[e];
}

// Step t of a launch (t > 0 only in mg_rollout): the cells are re-read every step (the observation
// phase stamps agents into them and packs its stage on top of them), the agents only at t = 0.
template <int MODE>
MG_HD void phase_load_plain(const Params &p, const Group &g, int lane, int t = 0) {
    const size_t e0 = (size_t)g.e0;
    warp_copy(g.cells, p.grid + e0 * p.cstride, g.ne * p.cstride * 4, lane);
    if (t == 0) warp_copy(g.ag, p.agents + e0 * p.n * 8, g.ne * p.n * 8, lane);
    if (MODE != MODE_OBS) warp_copy(g.act, p.actions + ((size_t)t * p.num_envs + e0) * p.n, g.ne * p.n, lane);
}

// n == 4 (the headline configuration): the env's four agent records in registers -- two 16-byte
// shared-memory loads instead of one 4/8-byte load per agent in each of the epilogue's loops.
struct Ag4 { uint32_t a0[4], a1[4]; };

MG_HD Ag4 load_ag4(const uint32_t *ag) {
    Ag4 a;
#ifdef __CUDA_ARCH__
    const uint4 q0 = *(const uint4 *)ag, q1 = *(const uint4 *)(ag + 4);
    a.a0[0] = q0.x; a.a1[0] = q0.y; a.a0[1] = q0.z; a.a1[1] = q0.w;
    a.a0[2] = q1.x; a.a1[2] = q1.y; a.a0[3] = q1.z; a.a1[3] = q1.w;
#else
    for (int j = 0; j < 4; j++) { a.a0[j] = ag[2 * j]; a.a1[j] = ag[2 * j + 1]; }
#endif
    return a;
}

MG_HD uint32_t terminated_mask4(const Ag4 &a) {
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) m |= (uint32_t)((a.a0[j] & 0xff000000u) != 0) << j;
    return m;
}

// 4 mask bits -> 4 bytes of 0/1 (bit i lands on bit 8*i; no two partial products collide)
MG_HD uint32_t bits_to_bytes4(uint32_t bits) { return ((bits & 15u) * 0x00204081u) & 0x01010101u; }

MG_HD void stamp_agents4(const Params &p, uint32_t *cells, const Ag4 &a, uint32_t terminated) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t x = (a.a0[j] >> 8) & 0xff, y = (a.a0[j] >> 16) & 0xff;
        const bool on = !((terminated >> j) & 1u) && x < (uint32_t)p.W && y < (uint32_t)p.H;
        if (on) cells[x * p.Hp + y] = T_AGENT | ((a.a1[j] >> 24) << 8) | ((a.a0[j] & 0xff) << 16);
    }
}

// ---- P2: auto-reset decision ("next-step" mode; is_done = base.py:534-539) --------------------------
MG_HD void phase_reset(const Params &p, const Group &g, int i, EnvRegs &r) {
    if (i < 0) return;
    uint32_t all_term = 1;
    if (p.n == 4) {
        all_term = terminated_mask4(load_ag4(g.ag + i * 8)) == 15u;
    } else {
        for (int j = 0; j < p.n; j++) all_term &= ((g.ag[(i * p.n + j) * 2] >> 24) & 0xff) != 0;
    }
    int k = -1;
    // LockedHallway terminates in the returned dict only (all doors unlocked), never in agent state
    if (p.hook == MG_HOOK_LOCKED_HALLWAY && __builtin_popcount((unsigned)r.hs) == p.hook_param) all_term = 1;
    if (all_term || r.sc >= p.max_steps) {
        k = (int)(((uint32_t)r.lidx + (uint32_t)p.lstride) % (uint32_t)p.K);
        r.lidx = k;
        r.sc = 0;
        r.hs = 0;
        const uint32_t *src = (const uint32_t *)(p.pool_agents + (size_t)k * p.n * 8);
            for (int j = 0; j < p.n * 2; j++) g.ag[i * p.n * 2 + j] = src[j];
    }
    g.rk[i] = k;
}

// Envs that were reset take their cells from the layout pool: into shared memory and written
// through to the state in HBM. `pending` = bit per env of the group that was reset.
MG_HD void phase_reset_grid(const Params &p, const Group &g, uint32_t pending, int lane) {
    for (int i = 0; i < g.ne; i++) {
        if (!((pending >> i) & 1u)) continue;
        const uint32_t *src = p.pool_grid + (size_t)g.rk[i] * p.cstride;
        warp_copy(g.cells + i * p.cstride, src, p.cstride * 4, lane);
        warp_copy(p.grid + (size_t)(g.e0 + i) * p.cstride, src, p.cstride * 4, lane);
        if (p.chain && lane == 0) p.chain[4 * (size_t)(g.e0 + i) + 2] = 0;  // equal to its pool layout again
    }
}

MG_HD uint32_t reset_mask_host(const Group &g) {  // hostsim only; the kernel uses a ballot
    uint32_t m = 0;
    for (int i = 0; i < g.ne; i++) m |= (uint32_t)(g.rk[i] >= 0) << i;
    return m;
}

// ---- P4: transition --------------------------------------------------------------------------------
MG_HD void store_cell(const Params &p, int e, int idx, uint32_t w) {  // dirty-cell write-through
    p.grid[(size_t)e * p.cstride + idx] = w;
    if (p.chain) p.chain[4 * (size_t)e + 2] = 1;  // the grid no longer equals its pool layout
}

MG_HD uint32_t all_agents(const Params &p) { return p.n >= 32 ? 0xffffffffu : (1u << p.n) - 1u; }

// base.py:478-507. Every reward handed out in one step has the same value (same step_count), so
// the step only records WHO is rewarded; `rewarded` = bit per agent.
MG_HD void on_success(const Params &p, uint32_t *ag, uint32_t &rewarded, int k) {
    if (p.flags & MG_FLAG_SUCCESS_ANY) {
            for (int j = 0; j < p.n; j++) ag[j * 2] |= 1u << 24;
    } else {
        ag[k * 2] |= 1u << 24;
    }
    rewarded |= (p.flags & MG_FLAG_JOINT_REWARD) ? all_agents(p) : (1u << k);
}

MG_HD void on_failure(const Params &p, uint32_t *ag, int k) {  // base.py:509-532
    if (p.flags & MG_FLAG_FAILURE_ANY) {
            for (int j = 0; j < p.n; j++) ag[j * 2] |= 1u << 24;
    } else {
        ag[k * 2] |= 1u << 24;
    }
}

MG_HD bool agent_at(const Params &p, const uint32_t *ag, uint32_t xy) {  // xy = x | y<<8
    bool hit = false;
    for (int j = 0; j < p.n; j++) hit |= ((ag[j * 2] >> 8) & 0xffffu) == xy;
    return hit;
}

// base.py:399: order = np_random.random(size=n).argsort(). Returns the order packed 4 bits per rank
// (n <= 4, keys in registers) or writes g.order (n > 4, keys in the stage).
MG_HD uint32_t draw_order(const Params &p, const Group &g, int i, EnvRegs &r) {
    const int n = p.n, G = p.G;
    if (n == 1) return 0;  // base.py:396-397
    if (n <= 4) {
        uint64_t k[4];
#pragma unroll
        for (int j = 0; j < 4; j++) k[j] = j < n ? pcg64_next53(r.lo, r.hi, r.ilo, r.ihi) : ~0ull;
        uint32_t rank[4] = {0, 0, 0, 0};
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int j = q + 1; j < 4; j++) {  // stable ascending: q < j goes first on ties
                const uint32_t q_first = k[q] <= k[j];
                rank[j] += q_first; rank[q] += q_first ^ 1u;
            }
        uint32_t ord = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) ord |= (uint32_t)j << (4 * rank[j]);
        return ord;
    }
    for (int j = 0; j < n; j++) g.keys[j * G + i] = pcg64_next53(r.lo, r.hi, r.ilo, r.ihi);
    for (int j = 0; j < n; j++) {  // rank = position in the ascending (stable) order
        const uint64_t kj = g.keys[j * G + i];
        int rk = 0;
        for (int q = 0; q < n; q++) {
            const uint64_t kq = g.keys[q * G + i];
            rk += (kq < kj) | ((kq == kj) & (q < j));
        }
        g.order[rk * G + i] = (uint8_t)j;
    }
    return 0;
}

// MultiGridEnv.handle_actions (base.py:378-476) for local env i; `ag` = this env's agent words.
MG_HD void handle_actions(const Params &p, const Group &g, int i, uint32_t *cells, uint32_t *ag,
                          uint32_t ord, uint32_t &rewarded) {
    const int n = p.n, G = p.G, e = g.e0 + i;
    const int8_t *act_e = g.act + i * n;
    const bool packed_order = n <= 4;  // (hoisted: the order of <= 4 agents travels in a register)
    const uint8_t *order_e = g.order + i;
    for (int r = 0; r < n; r++) {
        const int k = packed_order ? (int)(ord & 15u) : (int)order_e[r * G];
        ord >>= 4;
        const int act = act_e[k];
        uint32_t a0 = ag[k * 2], a1 = ag[k * 2 + 1];
        if (act < 0) continue;            // id not in the action dict (base.py:403-404)
        if ((a0 >> 24) & 0xff) continue;  // terminated (base.py:408-409)
        uint32_t dir = a0 & 3u;
        if (act == ACT_LEFT)  { ag[k * 2] = (a0 & ~0xffu) | ((dir + 3u) & 3u); continue; }  // base.py:412-413
        if (act == ACT_RIGHT) { ag[k * 2] = (a0 & ~0xffu) | ((dir + 1u) & 3u); continue; }  // base.py:416-417
        if (act == ACT_DONE) continue;
        if (act > ACT_DONE) { status_or(p.status, 1); continue; }  // reference: ValueError
        const int dx = (dir == 0) - (dir == 2), dy = (dir == 1) - (dir == 3);  // constants.py:21-30
        const int fx = (int)((a0 >> 8) & 0xff) + dx, fy = (int)((a0 >> 16) & 0xff) + dy;
        if ((unsigned)fx >= (unsigned)p.W || (unsigned)fy >= (unsigned)p.H) continue;
        const int idx = fx * p.Hp + fy;
        const uint32_t cw = cells[idx];
        const uint32_t t = cw & 0xff, col = (cw >> 8) & 0xff;
        const uint32_t st = (cw & DOOR_OBJ_CLOSED) ? (uint32_t)S_CLOSED : ((cw >> 16) & 0xff);
        const uint32_t fxy = (uint32_t)fx | ((uint32_t)fy << 8);
        if (act == ACT_FORWARD) {  // base.py:420-436
            const bool can_overlap = (t == T_EMPTY) | (t == T_FLOOR) | (t == T_GOAL) | (t == T_LAVA) |
                                     ((t == T_DOOR) & (st == S_OPEN));
            if (!can_overlap) continue;
            if (!(p.flags & MG_FLAG_ALLOW_OVERLAP) && agent_at(p, ag, fxy)) continue;
            ag[k * 2] = (a0 & 0xff0000ffu) | (fxy << 8);
            if (t == T_GOAL) on_success(p, ag, rewarded, k);
            if (t == T_LAVA) on_failure(p, ag, k);
        } else if (act == ACT_PICKUP) {  // base.py:439-446
            if (((t == T_KEY) | (t == T_BALL) | (t == T_BOX)) && (a1 & 0xff) == T_EMPTY) {
                ag[k * 2 + 1] = (a1 & 0xff000000u) | (cw & 0x00ffffffu);
                cells[idx] = CELL_EMPTY;
                store_cell(p, e, idx, CELL_EMPTY);
            }
        } else if (act == ACT_DROP) {  // base.py:449-459
            if ((a1 & 0xff) != T_EMPTY && t == T_EMPTY && !agent_at(p, ag, fxy)) {
                const uint32_t w = cell_word(a1 & 0xff, (a1 >> 8) & 0xff, (a1 >> 16) & 0xff);
                cells[idx] = w;
                store_cell(p, e, idx, w);
                ag[k * 2 + 1] = (a1 & 0xff000000u) | CELL_EMPTY;
            }
        } else {  // toggle, base.py:462-467
            if (t == T_DOOR) {  // Door.toggle, core/world_object.py:458-474
                uint32_t ns = st;
                if (st == S_LOCKED) {
                    if ((a1 & 0xff) == T_KEY && ((a1 >> 8) & 0xff) == col) ns = S_OPEN;
                } else {
                    ns = (st == S_OPEN) ? S_CLOSED : S_OPEN;
                }
                if (ns != st) {
                    const uint32_t w = cell_word(t, col, ns);
                    cells[idx] = w;
                    store_cell(p, e, idx, w);
                }
            } else if (t == T_BOX) {  // Box.toggle, core/world_object.py:599-605 (contains is None)
                cells[idx] = CELL_EMPTY;
                store_cell(p, e, idx, CELL_EMPTY);
            }
        }
    }
}

// gen_obs_grid, utils/obs.py:163-171: stamp non-terminated agents, ascending index (highest wins)
// `terminated` = bit per agent, taken BEFORE the env post-hook ran (the reference builds the
// observation before the hook, base.py:337 vs envs/*.py step()).
MG_HD void stamp_agents(const Params &p, uint32_t *cells, const uint32_t *ag, uint32_t terminated) {
    if (p.n <= 1) return;  // utils/obs.py:172-173
    for (int j = 0; j < p.n; j++) {
        const uint32_t a0 = ag[j * 2], a1 = ag[j * 2 + 1];
        if ((terminated >> j) & 1u) continue;
        const int x = (a0 >> 8) & 0xff, y = (a0 >> 16) & 0xff;
        if ((unsigned)x >= (unsigned)p.W || (unsigned)y >= (unsigned)p.H) continue;
        cells[x * p.Hp + y] = T_AGENT | ((a1 >> 24) << 8) | ((a0 & 0xff) << 16);
    }
}

// The env's lane: transition, then every per-env output straight from registers to HBM
// (lane <-> env, env-major arrays: the warp's accesses are contiguous).
MG_HD uint32_t terminated_mask(const Params &p, const uint32_t *ag) {
    uint32_t m = 0;
    for (int j = 0; j < p.n; j++) m |= (uint32_t)(((ag[j * 2] >> 24) & 0xff) != 0) << j;
    return m;
}

// RedBlueDoorsEnv.step post-hook (envs/redbluedoors.py:170-187): for every agent whose action was
// `toggle` (terminated or not), in agent order: if the cell in front of it is the (open) blue
// door, then success if the red door is open, else failure and the blue door's OBJECT is closed
// again. Runs on the un-stamped cells (before stamping).
MG_HD void hook_red_blue_doors(const Params &p, const Group &g, int i, const uint32_t *cells, uint32_t *ag,
                               uint32_t &rewarded) {
    const int n = p.n, e = g.e0 + i;
    const uint32_t DOOR_BLUE = T_DOOR | (2u << 8), DOOR_RED = T_DOOR | (0u << 8);
    bool blue_closed = false;
    for (int k = 0; k < n; k++) {
        if (g.act[i * n + k] != ACT_TOGGLE) continue;
        const uint32_t a0 = ag[k * 2], dir = a0 & 3u;
        const int fx = (int)((a0 >> 8) & 0xff) + (dir == 0) - (dir == 2);
        const int fy = (int)((a0 >> 16) & 0xff) + (dir == 1) - (dir == 3);
        if ((unsigned)fx >= (unsigned)p.W || (unsigned)fy >= (unsigned)p.H) continue;
        const int idx = fx * p.Hp + fy;
        const uint32_t cw = cells[idx];
        if ((cw & 0xffffu) != DOOR_BLUE || ((cw >> 16) & 0xff) != S_OPEN || (cw & DOOR_OBJ_CLOSED) || blue_closed) continue;
        bool red_open = false;  // the env's only red door (x-major scan)
        for (int x = 0; x < p.W; x++)
            for (int y = 0; y < p.H; y++) {
                const uint32_t c = cells[x * p.Hp + y];
                if ((c & 0xffffu) == DOOR_RED) { red_open = ((c >> 16) & 0xff) == S_OPEN; x = p.W; break; }
            }
        if (red_open) {
            on_success(p, ag, rewarded, k);
        } else {
            on_failure(p, ag, k);
            blue_closed = true;  // self.blue_door.is_open = False, WITHOUT grid.update(): state byte stays open
            store_cell(p, e, idx, cw | DOOR_OBJ_CLOSED);
        }
    }
}

// LockedHallwayEnv.step post-hook (envs/locked_hallway.py:203-227): every agent whose action was
// `toggle`, in agent order: the door in front of it is not locked and was not counted before ->
// remember it (bit per door colour; colours are distinct for <= 6 rooms, :156-158) and ADD the
// step's reward to every agent (joint) or to this agent. `bonus_all` / `bonus` count the additions.
MG_HD void hook_locked_hallway(const Params &p, const Group &g, int i, const uint32_t *cells, const uint32_t *ag,
                               int32_t &hs, uint32_t &bonus_all, uint32_t &bonus) {
    const int n = p.n;
    for (int k = 0; k < n; k++) {
        if (g.act[i * n + k] != ACT_TOGGLE) continue;
        const uint32_t a0 = ag[k * 2], dir = a0 & 3u;
        const int fx = (int)((a0 >> 8) & 0xff) + (dir == 0) - (dir == 2);
        const int fy = (int)((a0 >> 16) & 0xff) + (dir == 1) - (dir == 3);
        if ((unsigned)fx >= (unsigned)p.W || (unsigned)fy >= (unsigned)p.H) continue;
        const uint32_t cw = cells[fx * p.Hp + fy];
        if ((cw & 0xff) != T_DOOR || ((cw >> 16) & 0xff) == S_LOCKED) continue;
        const int32_t bit = 1 << ((cw >> 8) & 0x1f);
        if (hs & bit) continue;
        hs |= bit;
        if (p.flags & MG_FLAG_JOINT_REWARD) bonus_all += 1;
        else bonus |= 1u << k;
    }
}

// The agent order of this step only needs the env's PCG64 registers, not the TMA-loaded state, so
// it is drawn while the load is still in flight. `r` is advanced speculatively: phase_step keeps
// the advanced state only if the env really steps (an env that auto-resets consumes no draw).
struct OrderDraw { uint32_t ord; uint64_t lo0, hi0; };

template <int MODE>
MG_HD OrderDraw phase_draw(const Params &p, const Group &g, int i, EnvRegs &r) {
    OrderDraw d;
    d.ord = 0; d.lo0 = r.lo; d.hi0 = r.hi;
    if (MODE != MODE_OBS && i >= 0) d.ord = draw_order(p, g, i, r);
    return d;
}

template <int MODE>
MG_HD void phase_step(const Params &p, const Group &g, int i, EnvRegs &r, const OrderDraw &d, size_t tE = 0) {
    if (i < 0) return;  // tE = t * num_envs: slice t of the per-step output arrays (mg_rollout)
    const int n = p.n;
    uint32_t *cells = g.cells + i * p.cstride;
    uint32_t *ag = g.ag + i * n * 2;
    if constexpr (MODE == MODE_OBS) {
        stamp_agents(p, cells, ag, terminated_mask(p, ag));
    } else {
        const size_t e = (size_t)(g.e0 + i), eo = e + tE;
        const bool was_reset = (p.flags & MG_FLAG_AUTO_RESET) && g.rk[i] >= 0;
        uint32_t rewarded = 0;
        bool truncated = false;
        if (!was_reset) {
            r.sc += 1;  // base.py:333
            handle_actions(p, g, i, cells, ag, d.ord, rewarded);
            truncated = r.sc >= p.max_steps;  // base.py:339
        } else {
            r.lo = d.lo0; r.hi = d.hi0;  // no step, no draw
        }
        Ag4 a4;
        if (n == 4) a4 = load_ag4(ag);
        const uint32_t pre_hook_terminated = n == 4 ? terminated_mask4(a4) : terminated_mask(p, ag);
        if (!was_reset && p.hook == MG_HOOK_BLOCKED_UNLOCK_PICKUP) {  // envs/blockedunlockpickup.py:166-175
            for (int k = 0; k < n; k++)
                if ((ag[k * 2 + 1] & 0xff) == T_BOX) on_success(p, ag, rewarded, k);
        }
        if (!was_reset && p.hook == MG_HOOK_RED_BLUE_DOORS) hook_red_blue_doors(p, g, i, cells, ag, rewarded);
        uint32_t bonus_all = 0, bonus = 0;
        bool dict_terminated = false;  // LockedHallway: all doors unlocked -> terminations dict only
        if (p.hook == MG_HOOK_LOCKED_HALLWAY) {
            if (!was_reset) hook_locked_hallway(p, g, i, cells, ag, r.hs, bonus_all, bonus);
            dict_terminated = !was_reset && __builtin_popcount((unsigned)r.hs) == p.hook_param;
            p.hook_state[e] = r.hs;
        }
        if (MODE == MODE_STEP_OBS) {  // obs sees pre-hook state (the hooks only touch the terminated bytes)
            if (n == 4) stamp_agents4(p, cells, a4, pre_hook_terminated);
            else stamp_agents(p, cells, ag, pre_hook_terminated);
        }
        p.step_count[e] = r.sc;
        if (n > 1) { U128 s; s.lo = r.lo; s.hi = r.hi; *(U128 *)(p.pcg_state + 2 * e) = s; }
        if (p.flags & MG_FLAG_AUTO_RESET) p.layout_idx[e] = r.lidx;
        p.truncated[eo] = (uint8_t)truncated;
        const double rv = (rewarded | bonus_all | bonus) ? reward_value(r.sc, p.max_steps) : 0.0;  // base.py:394, 598-602
        const uint32_t term_force = dict_terminated ? 0x01010101u : 0u;
        if (n == 4) {
            const uint32_t post = p.hook == MG_HOOK_NONE ? pre_hook_terminated : terminated_mask(p, ag);
            *(uint32_t *)(p.terminated + eo * 4) = bits_to_bytes4(post) | term_force;
            if (p.direction)  // rollout: the 'direction' observation of this step (base.py:371)
                *(uint32_t *)(p.direction + eo * 4) = (a4.a0[0] & 0xff) | ((a4.a0[1] & 0xff) << 8) |
                                                      ((a4.a0[2] & 0xff) << 16) | ((a4.a0[3] & 0xff) << 24);
#ifdef __CUDA_ARCH__
            if (((uintptr_t)p.reward & 15u) == 0) {  // (the ABI only requires 8-byte alignment)
                double2 *rw = (double2 *)(p.reward + eo * 4);
                rw[0] = make_double2((rewarded & 1u) ? rv : 0.0, (rewarded & 2u) ? rv : 0.0);
                rw[1] = make_double2((rewarded & 4u) ? rv : 0.0, (rewarded & 8u) ? rv : 0.0);
            } else
#endif
            {
                for (int j = 0; j < 4; j++) p.reward[eo * 4 + j] = ((rewarded >> j) & 1u) ? rv : 0.0;
            }
        } else {
            for (int j = 0; j < n; j++)
                p.terminated[eo * n + j] = (uint8_t)((term_force & 1u) | (((ag[j * 2] >> 24) & 0xff) != 0));
            if (p.direction)
                for (int j = 0; j < n; j++) p.direction[eo * n + j] = (int8_t)(ag[j * 2] & 0xff);
            for (int j = 0; j < n; j++) p.reward[eo * n + j] = ((rewarded >> j) & 1u) ? rv : 0.0;
        }
        if (bonus_all | bonus) {  // LockedHallway only: rewards[k] += self._reward(), once per new door
            for (int j = 0; j < n; j++) {
                double rj = ((rewarded >> j) & 1u) ? rv : 0.0;
                for (uint32_t c = bonus_all + ((bonus >> j) & 1u); c > 0; c--) rj = rj + rv;
                p.reward[eo * n + j] = rj;
            }
        }
    }
}

// ---- P5: observation of one agent -------------------------------------------------------------------
// Closed form of get_view_exts + the rotation loop (utils/obs.py:175-202, 276-316):
//   obs[a][b] = G'[pos + f*(V-1-b) + r*(a - V/2)],  f = DIR_TO_VEC[dir], r = (-f.y, f.x),
// out of bounds -> wall (the sentinel row/column of the padded cell array).
// Visibility (utils/obs.py:236-273) as one bitmask per view row b. With s = see-through mask of the
// row and v = cells already visible, the two serial sweeps are a rightward then a leftward flood of v
// through s (carry chains of an addition), and the spill into row b-1 is A | A<<1 | A>>1 with
// A = visible & see-through.
struct ViewGeom {
    int pf, sf, Lf, stf;  // forward axis: agent coordinate, sign, length, BYTE stride
    int pl, sl, Ll, stl;  // lateral axis
    uint32_t carry;       // carried object as a cell word (utils/obs.py:207)
};

MG_HD ViewGeom view_geom(const Params &p, uint32_t a0, uint32_t a1) {
    ViewGeom v;
    const uint32_t dir = a0 & 3u;
    const int px = (a0 >> 8) & 0xff, py = (a0 >> 16) & 0xff;
    const bool horiz = !(dir & 1u);                    // forward axis is x for right/left
    v.sf = (dir & 2u) ? -1 : 1;                        // forward sign
    v.sl = (dir == 0u || dir == 3u) ? 1 : -1;          // lateral sign (r = (-f.y, f.x))
    v.pf = horiz ? px : py; v.Lf = horiz ? p.W : p.H; v.stf = horiz ? 4 * p.Hp : 4;
    v.pl = horiz ? py : px; v.Ll = horiz ? p.H : p.W; v.stl = horiz ? 4 : 4 * p.Hp;
    v.carry = cell_word24(a1 & 0x00ffffffu);
    return v;
}

MG_HD void vis_row(uint32_t &vis, uint32_t see, uint32_t full, uint32_t &m_out) {
    uint32_t m = vis;
    m |= ((see + (m & see)) ^ see) & full;              // i = 0..V-2 ascending  (utils/obs.py:256-262)
    // bit-reversed (cell 0 at bit 31): a carry now runs towards lower cells; it falls off bit 31
    const uint32_t sr = brev32(see);
    uint32_t mr = brev32(m);
    mr |= (sr + (mr & sr)) ^ sr;                        // i = V-1..1 descending (utils/obs.py:264-270)
    m = brev32(mr);
    const uint32_t a = m & see;
    vis = (a | (a << 1) | (a >> 1)) & full;
    m_out = m;
}

MG_HD uint32_t keep_if(uint32_t c, uint32_t m, uint32_t bit) {  // c if (m & bit) else UNSEEN (0,0,0)
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm("{\n.reg .pred p;\n.reg .b32 t;\nand.b32 t, %1, %2;\nsetp.ne.u32 p, t, 0;\nselp.b32 %0, %3, 0, p;\n}"
        : "=r"(r) : "r"(m), "r"(bit), "r"(c));
    return r;
#else
    return (m & bit) ? c : 0u;
#endif
}

// Shared-memory cell addressing for the gather: 32-bit shared-window addresses on the GPU so that
// one integer add per cell forms the address (the compiler otherwise re-derives base + offset per
// cell), plain pointers in the host simulator.
#ifdef __CUDA_ARCH__
typedef uint32_t cell_addr_t;
__device__ __forceinline__ cell_addr_t cell_base(const uint32_t *cells) { return (uint32_t)__cvta_generic_to_shared(cells); }
template <bool FENCE>
__device__ __forceinline__ uint32_t cell_load(cell_addr_t a) {
    uint32_t v;
    if (FENCE) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    else asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
#else
typedef const uint8_t *cell_addr_t;
inline cell_addr_t cell_base(const uint32_t *cells) { return (const uint8_t *)cells; }
template <bool FENCE>
inline uint32_t cell_load(cell_addr_t a) { return *(const uint32_t *)a; }
#endif

template <int VT>
MG_HD void obs_compute(const Params &p, const uint32_t *cells, uint32_t a0, uint32_t a1, uint32_t (&cr)[VT * VT]) {
    constexpr int V = VT, half = VT / 2;
    const ViewGeom g = view_geom(p, a0, a1);
    const uint32_t full = (1u << V) - 1u;
    const bool stw = (p.flags & MG_FLAG_SEE_THROUGH_WALLS) != 0;

    cell_addr_t colp[V];  // view column a -> its cell in grid row / column 0 (one add per cell below)
#pragma unroll
    for (int a = 0; a < V; a++) {
        int c = g.pl + g.sl * (a - half);
        c = (unsigned)c < (unsigned)g.Ll ? c : g.Ll;
        colp[a] = cell_base(cells) + c * g.stl;
    }
    // Row by row, nearest first. A row is only gathered while something in it can still be visible:
    // once the visibility mask entering a row is empty (the view is walled off), that row and all
    // further rows are UNSEEN whatever they contain, so their loads are skipped -- in small rooms
    // that is ~40 % of the gathers, and the shared-memory wavefronts of the gather are what bounds
    // this phase.
    uint32_t vis = 1u << half;                          // vis_mask[V//2][V-1] = True (utils/obs.py:252)
#pragma unroll
    for (int b = V - 1; b >= 0; b--) {
        int r = g.pf + g.sf * (V - 1 - b);
        r = (unsigned)r < (unsigned)g.Lf ? r : g.Lf;
        const int rowoff = r * g.stf;
        if (stw || vis != 0) {
#pragma unroll
            for (int a = V - 1; a >= 0; a--) {
                if (b == V - 1 && a == half) cr[a * V + b] = g.carry;
                else if (a == V - 1) cr[a * V + b] = cell_load<true>(colp[a] + rowoff);
                else cr[a * V + b] = cell_load<false>(colp[a] + rowoff);
            }
        } else {
#pragma unroll
            for (int a = 0; a < V; a++) cr[a * V + b] = 0;
        }
        if (!stw) {
            uint32_t opq = 0;
#pragma unroll
            for (int a = V - 1; a >= 0; a--) opq = shl1_in(opq, cr[a * V + b]);
            uint32_t m;
            vis_row(vis, ~opq & full, full, m);
#pragma unroll
            for (int a = 0; a < V; a++) cr[a * V + b] = keep_if(cr[a * V + b], m, 1u << a);  // utils/obs.py:95-100
        }
    }
}

// 24-bit cells -> dense byte stream, written as 32-bit words (bit 31 is never selected)
template <int VT>
MG_HD void obs_pack_store(const Params &p, const uint32_t (&cr)[VT * VT], uint8_t *out) {
    constexpr int NC = VT * VT, NW = (3 * NC + 3) / 4;
    uint32_t *o32 = (uint32_t *)out;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        const int i0 = (4 * w) / 3, sh = 4 * w - 3 * i0;  // first cell, byte offset inside it
        const uint32_t c0 = cr[i0 < NC ? i0 : 0];
        const uint32_t c1 = (i0 + 1 < NC) ? cr[i0 + 1 < NC ? i0 + 1 : 0] : 0u;
        const uint32_t sel = sh == 0 ? 0x4210u : (sh == 1 ? 0x5421u : 0x6542u);
        o32[w] = byte_perm(c0, c1, sel);  // c1 == 0 past the last cell: padding bytes are zero
    }
    for (int w = NW; w * 4 < p.ostride; w++) o32[w] = 0;
}

// Any odd V up to MG_MAX_VIEW: same algorithm, rolled loops, bytes written straight to the stage.
MG_HD void obs_agent_generic(const Params &p, const uint32_t *cells, uint32_t a0, uint32_t a1, uint8_t *out) {
    const int V = p.V, half = V >> 1;
    const ViewGeom g = view_geom(p, a0, a1);
    const uint32_t full = (1u << V) - 1u;
    const bool stw = (p.flags & MG_FLAG_SEE_THROUGH_WALLS) != 0;
    const uint8_t *base = (const uint8_t *)cells;
    uint32_t vis = 1u << half;
    for (int b = V - 1; b >= 0; b--) {
        int r = g.pf + g.sf * (V - 1 - b);
        r = (unsigned)r < (unsigned)g.Lf ? r : g.Lf;
        const uint8_t *row = base + r * g.stf;
        uint32_t opq = 0;
        for (int a = V - 1; a >= 0; a--) {
            int c = g.pl + g.sl * (a - half);
            c = (unsigned)c < (unsigned)g.Ll ? c : g.Ll;
            uint32_t w = *(const uint32_t *)(row + c * g.stl);
            if (b == V - 1 && a == half) w = g.carry;
            opq = shl1_in(opq, w);
            uint8_t *o = out + (a * V + b) * 3;
            o[0] = (uint8_t)w; o[1] = (uint8_t)(w >> 8); o[2] = (uint8_t)(w >> 16);
        }
        uint32_t m = full;
        if (!stw) vis_row(vis, ~opq & full, full, m);
        for (int a = 0; a < V; a++)
            if (!((m >> a) & 1u)) {
                uint8_t *o = out + (a * V + b) * 3;
                o[0] = 0; o[1] = 0; o[2] = 0;
            }
    }
    for (int q = 3 * V * V; q < p.ostride; q++) out[q] = 0;
}

// One pass: lane handles agent task `pass*32 + lane` of the group (tasks are env-major, so a pass
// is a contiguous span of the obs array). Split in two so that the wait for the previous pass's
// TMA store (which reads the stage) sits between the register work and the stage writes.
struct ObsTask { const uint32_t *cells; uint32_t a0, a1; bool valid; };

MG_HD ObsTask obs_task(const Params &p, const Group &g, int pass, int lane) {
    ObsTask t;
    const int id = pass * LANES + lane;
    t.valid = id < g.ne * p.n;
    const int idc = t.valid ? id : 0;
    t.cells = g.cells + (int)fastdiv((uint32_t)idc, p.rcp_n) * p.cstride;
    t.a0 = g.ag[idc * 2]; t.a1 = g.ag[idc * 2 + 1];
    return t;
}

MG_HD int obs_passes(const Params &p, const Group &g) { return (g.ne * p.n + LANES - 1) / LANES; }

// Where pass `pass` packs its 32 observations (see carve_smem).
MG_HD uint8_t *stage_of(const Params &p, const Group &g, int pass) {
    return p.alias ? g.stage + (pass + 1) * p.pass_cell_bytes + p.stage_extra - p.stage_bytes : g.stage;
}

MG_HD void phase_obs_store_plain(const Params &p, const Group &g, int pass, int lane, size_t tE = 0) {
    const int cnt = g.ne * p.n - pass * LANES < LANES ? g.ne * p.n - pass * LANES : LANES;
    warp_copy(p.obs + ((tE + (size_t)g.e0) * p.n + (size_t)pass * LANES) * p.ostride, stage_of(p, g, pass), cnt * p.ostride, lane);
}

// ---- P6 (plain path): agents back to HBM --------------------------------------------------------------
MG_HD void phase_store_plain(const Params &p, const Group &g, int lane) {
    warp_copy(p.agents + (size_t)g.e0 * p.n * 8, g.ag, g.ne * p.n * 8, lane);
}

// ---- fused one-hot observations (MgStepOut.one_hot) ------------------------------------------------------
// OneHotObsWrapper.one_hot (multigrid/wrappers.py:158-190) of the observations a pass has just packed into its
// stage: image (type, colour, state) -> 21 channels = 11 type + 6 colour + 4 state/direction, uint8, written as
// the contiguous tensor [E][n][V][V][21] the wrapper returns -- straight from shared memory, so the 7x larger
// tensor costs its HBM writes and nothing else (no second launch, no re-read of the observations).
// Bytes [first * VV * 21, (first + cnt) * VV * 21) of `out` belong to the pass (first = index of its first
// agent); they are cut at ABSOLUTE 16-byte boundaries of the tensor (base 16-byte aligned): a lane produces
// whole 16-byte vectors, the ragged head / tail bytes of a span that does not start or end on a boundary
// (n * G not a multiple of 16, tail groups) go out bytewise. A vector covers parts of at most two cells; the
// <= 6 one-bytes of its window are placed with one shift each into a 16-bit mask that a multiply spreads.
MG_HD uint32_t shl1(uint32_t q) {  // 1 << q, 0 for q >= 32 (PTX shift semantics; "negative" q wrapped)
#ifdef __CUDA_ARCH__
    uint32_t r; asm("shl.b32 %0, 1, %1;" : "=r"(r) : "r"(q)); return r;
#else
    return q < 32u ? 1u << q : 0u;
#endif
}

MG_HD uint32_t one_hot_mask(const uint8_t *stage, int sstride, uint32_t VV, uint32_t rcp_vv, uint32_t rel,
                            uint32_t nbytes) {
    // bit q of the result <=> byte rel + q of the span is 1 (q < 16; higher bits are garbage to be masked off).
    // Branch-free: the lanes of a warp sit at different offsets inside their cells, and a branch on "does my window
    // reach the next cell" splits the warp for the rest of the (unrolled) loop body.
    const uint32_t c = mulhi32(rel, 204522253u), off = rel - c * 21u;  // rel / 21 (rel < 2^21)
    const uint32_t a = fastdiv(c, rcp_vv), cia = c - a * VV;
    const uint8_t *src = stage + a * (uint32_t)sstride + cia * 3u;
    // the next cell (the first of the next agent's slot after the last cell of this one); its bits start at 21 - off:
    // at 16 or beyond -- nothing of it is in the window -- when off <= 5. Past the end of the span: same cell, no bits.
    const bool more = rel + (21u - off) < nbytes, wrap = cia + 1u == VV;
    const uint8_t *nxt = !more ? src : (wrap ? src + (sstride - (int)(cia * 3u)) : src + 3);
    const uint32_t base = 0u - off, base2 = more ? 21u - off : 64u;
    return shl1(base + src[0]) | shl1(base + 11u + src[1]) | shl1(base + 17u + src[2]) |
           shl1(base2 + nxt[0]) | shl1(base2 + 11u + nxt[1]) | shl1(base2 + 17u + nxt[2]);
}

// (explicit scalars instead of `const Params &`: the out-of-line copies below must not force the kernel's
// parameter block into local memory)
MG_HD void one_hot_emit(uint8_t *one_hot, uint32_t VV, uint32_t rcp_vv, const uint8_t *stage, int sstride, size_t first,
                        int cnt, int lane) {
    const uint32_t nbytes = (uint32_t)cnt * VV * 21u;
    const size_t b0 = first * VV * 21u;
    uint8_t *out = one_hot + b0;
    const uint32_t head = (uint32_t)((16u - (uint32_t)(b0 & 15u)) & 15u);  // bytes before the first boundary
    if (head >= nbytes) {  // (tiny span inside one vector)
        for (uint32_t b = lane; b < nbytes; b += LANES) out[b] = (uint8_t)(one_hot_mask(stage, sstride, VV, rcp_vv, b, nbytes) & 1u);
        return;
    }
    const uint32_t nvec = (nbytes - head) >> 4, tail0 = head + (nvec << 4);
#pragma unroll 4
    for (uint32_t v = lane; v < nvec; v += LANES) {
        const uint32_t rel = head + (v << 4);
        const uint32_t mask = one_hot_mask(stage, sstride, VV, rcp_vv, rel, nbytes);
        // 4 mask bits -> 4 bytes of 0/1: the multiply puts bit i at bit 8*i (no two products collide)
        U128 w;
        w.lo = (uint64_t)(((mask & 15u) * 0x00204081u) & 0x01010101u) |
               ((uint64_t)((((mask >> 4) & 15u) * 0x00204081u) & 0x01010101u) << 32);
        w.hi = (uint64_t)((((mask >> 8) & 15u) * 0x00204081u) & 0x01010101u) |
               ((uint64_t)((((mask >> 12) & 15u) * 0x00204081u) & 0x01010101u) << 32);
        *(U128 *)(out + rel) = w;
    }
    // ragged head and tail (fewer than 16 bytes each)
    const uint32_t ragged = head + (nbytes - tail0);
    for (uint32_t q = lane; q < ragged; q += LANES) {
        const uint32_t b = q < head ? q : tail0 + (q - head);
        out[b] = (uint8_t)(one_hot_mask(stage, sstride, VV, rcp_vv, b, nbytes) & 1u);
    }
}

// Out of line (one copy per module; the plain launch's registers and instruction footprint stay what they are).
MG_HD_COLD void one_hot_emit_cold(uint8_t *one_hot, uint32_t VV, uint32_t rcp_vv, const uint8_t *stage, int sstride,
                                  size_t first, int cnt, int lane) {
    one_hot_emit(one_hot, VV, rcp_vv, stage, sstride, first, cnt, lane);
}

// The fused one-hot image of a pass of the general kernel, from the pass's complete stage.
MG_HD void phase_one_hot(const Params &p, const Group &g, int pass, int lane) {
    const int left = g.ne * p.n - pass * LANES;
    one_hot_emit(p.one_hot, (uint32_t)(p.V * p.V), p.rcp_vv, stage_of(p, g, pass), p.ostride,
                 (size_t)g.e0 * p.n + (size_t)pass * LANES, left < LANES ? left : LANES, lane);
}

#ifdef __CUDACC__
// ---- TMA bulk copies + mbarrier (sm_90+/sm_100a PTX) -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    // (a C++ loop around try_wait, not a branch inside the asm: with asm-level branches the compiler no longer
    // knows the control flow and stops marking the regions that follow as reconvergent)
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
// L2 eviction-priority variants (createpolicy + .L2::cache_hint)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void *dst_gmem, const void *src_smem, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes), "l"(pol) : "memory");
}
// Programmatic dependent launch (PDL): `launch_dependents` lets the next kernel of the stream be
// scheduled onto SM resources as this grid's blocks retire; `wait` blocks until the previous grid
// has completed and its writes are visible. Both are no-ops for a launch without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t *p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_acquire_gpu_v4(const uint32_t *p) {
    uint4 v;
    asm volatile("ld.acquire.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (TMA) reads that follow
// generic-proxy global writes (dirty cells, reset layouts) -> visible to later TMA loads
__device__ __forceinline__ void fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Compiled in only with -DMG_TRACE (python -m multigrid_b200.build --trace): the product build has
// no trace instructions.
__device__ __forceinline__ void trace_mark(const Params &p, int group, int lane, int slot) {
#ifdef MG_TRACE
    if (p.trace && lane == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (slot == 7) { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); t = smid; }
        p.trace[(size_t)group * 8 + slot] = t;
    }
#endif
}

template <int MODE>
// dedup: 0 = everything now; 1 = everything but the cells (their source is decided once the group's dirty
// flags have arrived); 2 = the cells from the 32-copy pool buffer; 3 = the cells from the group's own grids.
// The transaction count of all parts is announced by the first call.
__device__ __forceinline__ void load_bulk(const Params &p, const Group &g, uint64_t *bar, int t, int dedup = 0) {
    const size_t e0 = (size_t)g.e0;
    const uint32_t G = (uint32_t)p.G, n = (uint32_t)p.n;
    uint32_t total = G * p.cstride * 4;
    if (t == 0) total += G * n * 8;
    if (MODE != MODE_OBS) total += G * n;
    const int8_t *act = p.actions + ((size_t)t * p.num_envs + e0) * n;
    if (dedup >= 2) {  // the cells: G copies of the one pool layout from the L2-resident buffer (no evict_first:
                       // every warp re-reads it), or the group's own grids
        if (dedup == 2) bulk_g2s(g.cells, p.pool_rep, G * p.cstride * 4, bar);
        else if (p.l2hint & 1) bulk_g2s_hint(g.cells, p.grid + e0 * p.cstride, G * p.cstride * 4, bar, l2_policy_evict_first());
        else bulk_g2s(g.cells, p.grid + e0 * p.cstride, G * p.cstride * 4, bar);
        return;
    }
    mbar_expect_tx(bar, total);
    if (dedup == 1) {
        if (t == 0) bulk_g2s(g.ag, p.agents + e0 * n * 8, G * n * 8, bar);
        if (MODE != MODE_OBS) {
            if (p.l2hint & 1) bulk_g2s_hint(g.act, act, G * n, bar, l2_policy_evict_first());
            else bulk_g2s(g.act, act, G * n, bar);
        }
        return;
    }
    if (p.l2hint & 1) {
        const uint64_t pol = l2_policy_evict_first();
        bulk_g2s_hint(g.cells, p.grid + e0 * p.cstride, G * p.cstride * 4, bar, pol);
        if (t == 0) bulk_g2s(g.ag, p.agents + e0 * n * 8, G * n * 8, bar);  // stored back by this launch: keep
        if (MODE != MODE_OBS) bulk_g2s_hint(g.act, act, G * n, bar, pol);
        return;
    }
    bulk_g2s(g.cells, p.grid + e0 * p.cstride, G * p.cstride * 4, bar);
    if (t == 0) bulk_g2s(g.ag, p.agents + e0 * n * 8, G * n * 8, bar);
    if (MODE != MODE_OBS) bulk_g2s(g.act, act, G * n, bar);
}

// ---- layout conversion kernels (API utilities, not on the hot path) -----------------------------------
// (E,W,H,3) bytes as in Grid.state (core/grid.py:54)  <->  padded cell words. One thread per word.
__global__ void pack_grid_kernel(int W, int H, int64_t total, const int8_t *__restrict__ grid3,
                                 uint32_t *__restrict__ cells) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int Hp = H + 1, cs = (W + 1) * Hp;
    const int64_t e = idx / cs;
    const int c = (int)(idx - e * cs), x = c / Hp, y = c - x * Hp;
    uint32_t w = CELL_WALL;
    if (x < W && y < H) {
        const uint8_t *src = (const uint8_t *)grid3 + ((e * W + x) * H + y) * 3;
        w = cell_word(src[0], src[1], src[2]);
    }
    cells[idx] = w;
}

__global__ void unpack_grid_kernel(int W, int H, int64_t total, const uint32_t *__restrict__ cells,
                                   int8_t *__restrict__ grid3) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over E*W*H
    if (idx >= total) return;
    const int64_t e = idx / (W * H);
    const int c = (int)(idx - e * (W * H)), x = c / H, y = c - x * H;
    const uint32_t w = cells[e * (int64_t)(W + 1) * (H + 1) + x * (H + 1) + y];
    uint8_t *dst = (uint8_t *)grid3 + idx * 3;
    dst[0] = (uint8_t)w; dst[1] = (uint8_t)(w >> 8); dst[2] = (uint8_t)(w >> 16);
}

// One thread per layout; the generator state is advanced in place (rng_buf: bit 32 = has_uint32,
// low word = the buffered upper half). Not on the step path: it fills the reset-layout pool.
__global__ void gen_layouts_empty_random_kernel(int W, int H, int n, int64_t K, uint64_t *rng_state,
                                                const uint64_t *rng_inc, uint64_t *rng_buf, uint32_t *cells,
                                                int8_t *agents, int32_t *status) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    LayoutRng g;
    g.lo = rng_state[2 * k]; g.hi = rng_state[2 * k + 1]; g.ilo = rng_inc[2 * k]; g.ihi = rng_inc[2 * k + 1];
    const uint64_t b = rng_buf ? rng_buf[k] : 0ull;
    g.has32 = (uint32_t)(b >> 32) & 1u; g.buf32 = (uint32_t)b;
    const bool ok = gen_layout_empty_random(W, H, n, g, cells + k * (int64_t)(W + 1) * (H + 1), agents + k * n * 8);
    if (!ok) status_or(status, 2);
    rng_state[2 * k] = g.lo; rng_state[2 * k + 1] = g.hi;
    if (rng_buf) rng_buf[k] = ((uint64_t)g.has32 << 32) | g.buf32;
}

__global__ void gen_layouts_red_blue_doors_kernel(int size, int n, int64_t K, uint64_t *rng_state,
                                                  const uint64_t *rng_inc, uint64_t *rng_buf, uint32_t *cells,
                                                  int8_t *agents, int32_t *status) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    LayoutRng g;
    g.lo = rng_state[2 * k]; g.hi = rng_state[2 * k + 1]; g.ilo = rng_inc[2 * k]; g.ihi = rng_inc[2 * k + 1];
    const uint64_t b = rng_buf ? rng_buf[k] : 0ull;
    g.has32 = (uint32_t)(b >> 32) & 1u; g.buf32 = (uint32_t)b;
    const bool ok = gen_layout_red_blue_doors(size, n, g, cells + k * (int64_t)(2 * size + 1) * (size + 1), agents + k * n * 8);
    if (!ok) status_or(status, 2);
    rng_state[2 * k] = g.lo; rng_state[2 * k + 1] = g.hi;
    if (rng_buf) rng_buf[k] = ((uint64_t)g.has32 << 32) | g.buf32;
}

__global__ void gen_layouts_locked_hallway_kernel(int num_rooms, int S, int mhk, int mkpr, int n, int64_t K,
                                                  uint64_t *rng_state, const uint64_t *rng_inc, uint64_t *rng_buf,
                                                  uint32_t *cells, int8_t *agents, int32_t *status) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    LayoutRng g;
    g.lo = rng_state[2 * k]; g.hi = rng_state[2 * k + 1]; g.ilo = rng_inc[2 * k]; g.ihi = rng_inc[2 * k + 1];
    const uint64_t b = rng_buf ? rng_buf[k] : 0ull;
    g.has32 = (uint32_t)(b >> 32) & 1u; g.buf32 = (uint32_t)b;
    const int64_t cs = (int64_t)(3 * (S - 1) + 2) * ((num_rooms / 2) * (S - 1) + 2);
    const bool ok = gen_layout_locked_hallway(num_rooms, S, mhk, mkpr, n, g, cells + k * cs, agents + k * n * 8);
    if (!ok) status_or(status, 2);
    rng_state[2 * k] = g.lo; rng_state[2 * k + 1] = g.hi;
    if (rng_buf) rng_buf[k] = ((uint64_t)g.has32 << 32) | g.buf32;
}

// Playground layouts (two generators, like BUP below).
__global__ void gen_layouts_playground_kernel(int S, int rows, int cols, int n, int64_t K, uint64_t *rng_state,
                                              const uint64_t *rng_inc, uint64_t *rng_buf, uint64_t *order_state,
                                              const uint64_t *order_inc, uint64_t *order_buf, uint32_t *cells,
                                              int8_t *agents, int32_t *status) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    LayoutRng g, o;
    g.lo = rng_state[2 * k]; g.hi = rng_state[2 * k + 1]; g.ilo = rng_inc[2 * k]; g.ihi = rng_inc[2 * k + 1];
    const uint64_t b = rng_buf ? rng_buf[k] : 0ull;
    g.has32 = (uint32_t)(b >> 32) & 1u; g.buf32 = (uint32_t)b;
    o.lo = order_state[2 * k]; o.hi = order_state[2 * k + 1]; o.ilo = order_inc[2 * k]; o.ihi = order_inc[2 * k + 1];
    const uint64_t ob = order_buf ? order_buf[k] : 0ull;
    o.has32 = (uint32_t)(ob >> 32) & 1u; o.buf32 = (uint32_t)ob;
    const int64_t cs = (int64_t)(cols * (S - 1) + 2) * (rows * (S - 1) + 2);
    if (!gen_layout_playground(S, rows, cols, n, g, o, cells + k * cs, agents + k * n * 8)) status_or(status, 2);
    rng_state[2 * k] = g.lo; rng_state[2 * k + 1] = g.hi;
    if (rng_buf) rng_buf[k] = ((uint64_t)g.has32 << 32) | g.buf32;
    order_state[2 * k] = o.lo; order_state[2 * k + 1] = o.hi;
    if (order_buf) order_buf[k] = ((uint64_t)o.has32 << 32) | o.buf32;
}

// BlockedUnlockPickup layouts, one thread per layout. order_state/order_inc: the env's own generator
// (env.np_random), advanced by the door-height draw; order_buf = its buffered 32-bit half (integers() leaves the
// unused half of a 64-bit draw there and the NEXT reset's door draw consumes it). info[k] = box colour.
__global__ void gen_layouts_bup_kernel(int S, int n, int64_t K, uint64_t *rng_state, const uint64_t *rng_inc,
                                       uint64_t *rng_buf, uint64_t *order_state, const uint64_t *order_inc,
                                       uint64_t *order_buf, uint32_t *cells, int8_t *agents, int32_t *info,
                                       int32_t *status) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    LayoutRng g, o;
    g.lo = rng_state[2 * k]; g.hi = rng_state[2 * k + 1]; g.ilo = rng_inc[2 * k]; g.ihi = rng_inc[2 * k + 1];
    const uint64_t b = rng_buf ? rng_buf[k] : 0ull;
    g.has32 = (uint32_t)(b >> 32) & 1u; g.buf32 = (uint32_t)b;
    o.lo = order_state[2 * k]; o.hi = order_state[2 * k + 1]; o.ilo = order_inc[2 * k]; o.ihi = order_inc[2 * k + 1];
    const uint64_t ob = order_buf ? order_buf[k] : 0ull;
    o.has32 = (uint32_t)(ob >> 32) & 1u; o.buf32 = (uint32_t)ob;
    const int W = 2 * (S - 1) + 1;
    const int color = gen_layout_bup(S, n, g, o, cells + k * (int64_t)(W + 1) * (S + 1), agents + k * n * 8);
    if (color < 0) status_or(status, 2);
    if (info) info[k] = color;
    rng_state[2 * k] = g.lo; rng_state[2 * k + 1] = g.hi;
    if (rng_buf) rng_buf[k] = ((uint64_t)g.has32 << 32) | g.buf32;
    order_state[2 * k] = o.lo; order_state[2 * k + 1] = o.hi;
    if (order_buf) order_buf[k] = ((uint64_t)o.has32 << 32) | o.buf32;
}

// One thread per env: regenerate the slot of every env that is done (see refresh_slot).
__global__ void refresh_done_kernel(const __grid_constant__ Params p, const LayoutGen lg) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.num_envs || !env_is_done(p, (size_t)e)) return;
    if (!refresh_slot(p, lg, (size_t)e)) status_or(p.status, 2);
}

// Host-driven reset of selected envs from the layout pool, one warp per env: exactly what the step
// kernel's auto-reset does to an env (phase_reset / phase_reset_grid) -- next layout of the pool, step_count
// and hook state zeroed, the env's PCG64 stream untouched.
__global__ void reset_where_kernel(const __grid_constant__ Params p, const uint8_t *__restrict__ mask) {
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (e >= p.num_envs || !mask[e]) return;  // whole warp
    int k = 0;
    if (lane == 0) k = (int)(((uint32_t)p.layout_idx[e] + (uint32_t)p.lstride) % (uint32_t)p.K);
    k = __shfl_sync(0xffffffffu, k, 0);
    // (single layout: a clean env already holds it)
    const bool clean = p.pool_rep != nullptr && p.chain[4 * (size_t)e + 2] == 0;
    if (!clean) {
        const uint32_t *src = p.pool_grid + (size_t)k * p.cstride;
        uint32_t *dst = p.grid + (size_t)e * p.cstride;
        for (int w = lane; w < p.cstride; w += LANES) dst[w] = src[w];
    }
    __syncwarp();  // (every lane has read the flag before lane 0 clears it below)
    const uint32_t *asrc = (const uint32_t *)(p.pool_agents + (size_t)k * p.n * 8);
    uint32_t *adst = (uint32_t *)(p.agents + (size_t)e * p.n * 8);
    for (int w = lane; w < p.n * 2; w += LANES) adst[w] = asrc[w];
    if (lane == 0) {
        p.layout_idx[e] = k;
        p.step_count[e] = 0;
        if (p.hook_state) p.hook_state[e] = 0;
        if (p.chain) p.chain[4 * (size_t)e + 2] = 0;
    }
}

// <= 4 warps per block; register cap: 72 (7 blocks x 128 threads per SM) up to V = 7, 128 for V = 9
// FullyObsWrapper.observation (multigrid/wrappers.py:50-58): the whole grid as (W,H,3) bytes with
// EVERY agent (terminated or not) written over its cell as (agent, colour, dir), ascending agent
// index (the highest index wins). One thread per (env, cell).
__global__ void full_obs_kernel(int W, int H, int n, int64_t total, const uint32_t *__restrict__ cells,
                                const int8_t *__restrict__ agents, int8_t *__restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over E*W*H
    if (idx >= total) return;
    const int64_t e = idx / (W * H);
    const int c = (int)(idx - e * (W * H)), x = c / H, y = c - x * H;
    uint32_t w = cells[e * (int64_t)(W + 1) * (H + 1) + x * (H + 1) + y];
    const int8_t *ag = agents + e * n * 8;
    for (int j = 0; j < n; j++)
        if (ag[j * 8 + 1] == x && ag[j * 8 + 2] == y)
            w = T_AGENT | ((uint32_t)(uint8_t)ag[j * 8 + 7] << 8) | ((uint32_t)(uint8_t)ag[j * 8] << 16);
    uint8_t *dst = (uint8_t *)out + idx * 3;
    dst[0] = (uint8_t)w; dst[1] = (uint8_t)(w >> 8); dst[2] = (uint8_t)(w >> 16);
}

// OneHotObsWrapper.one_hot (multigrid/wrappers.py:158-190) over a whole observation batch:
// image (type,color,state) -> 21 channels = 11 type + 6 colour + 4 state/direction, uint8.
// The output [agents][V][V][21] is written as one flat stream of 32-bit words (4 channels each).
__global__ void one_hot_kernel(int cells, int64_t agents, int ostride, const int8_t *__restrict__ obs,
                               uint8_t *__restrict__ out) {
    const int64_t per_agent = (int64_t)cells * 21, total = agents * per_agent;
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, p0 = 4 * w;
    if (p0 >= total) return;
    uint32_t word = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const int64_t pos = p0 + b;
        if (pos >= total) break;
        const int64_t agent = pos / per_agent;
        const int rem = (int)(pos - agent * per_agent), cell = rem / 21, ch = rem - cell * 21;
        const int8_t *src = obs + agent * ostride + cell * 3;
        const int hit = ch < 11 ? (src[0] == ch) : ch < 17 ? (src[1] == ch - 11) : (src[2] == ch - 17);
        word |= (uint32_t)hit << (8 * b);
    }
    if (p0 + 4 <= total) {
        *(uint32_t *)(out + p0) = word;
    } else {
        for (int b = 0; p0 + b < total; b++) out[p0 + b] = (uint8_t)(word >> (8 * b));
    }
}

// Network input of the reference's training script (scripts/train.py:56-63 preprocess_batch over the
// OneHotObsWrapper image): float32 [A][V][V][23] = the 21 one-hot channels as 0.0 / 1.0 followed by
// cos and sin of 2*pi*direction/4, broadcast over the view. One pass from the 3-byte observation instead of
// uint8 one-hot (x7) -> concatenate -> .float() (x4.4 again). dir_lut[d] = {cos, sin} comes from the caller
// (torch's float32 cos / sin, so the values are the reference's). One block = 16 agents, 4 floats per thread
// and store; the (agent, cell, channel) position is divided out once per thread and then incremented.
__global__ void __launch_bounds__(256) obs_features_kernel(int V, int64_t agents, int ostride, uint32_t rcp_per_agent,
                                                           const int8_t *__restrict__ obs,
                                                           const int8_t *__restrict__ direction, int dir_stride,
                                                           const float *__restrict__ dir_lut, float4 *__restrict__ out) {
    const uint32_t VV = (uint32_t)(V * V), per_agent = VV * 23u;
    const int64_t a0 = (int64_t)blockIdx.x * 16;
    const uint32_t na = (uint32_t)(agents - a0 < 16 ? agents - a0 : 16);
    const uint32_t nfl = na * per_agent;  // floats of this block (<= 16 * 225 * 23 = 82 800)
    const uint8_t *obs_b = (const uint8_t *)obs + a0 * ostride;
    const int8_t *dir_b = direction + a0 * dir_stride;
    float *out_b = (float *)out + a0 * per_agent;
    for (uint32_t q0 = 4u * threadIdx.x; q0 < nfl; q0 += 4u * 256u) {
        const uint32_t a = fastdiv(q0, rcp_per_agent), rem = q0 - a * per_agent;
        const uint32_t cell = mulhi32(rem, 186737709u), ch = rem - cell * 23u;  // rem / 23, rem < 2^16
        // the 4 floats lie in this cell and possibly the next: both cells as one 46-bit channel mask
        // (bit k of a cell = channel k is hot; channels 21, 22 are the direction features)
        const uint8_t *src = obs_b + a * (uint32_t)ostride + cell * 3u;
        uint64_t bits = (1ull << src[0]) | (1ull << (11u + src[1])) | (1ull << (17u + src[2]));
        const uint32_t d0 = (uint32_t)dir_b[a * dir_stride] & 3u;
        if (ch > 19u && q0 + (23u - ch) < nfl) {  // (the next cell contributes its channels 0..2 at most)
            uint32_t a1 = a, c1 = cell + 1u;
            if (c1 == VV) { c1 = 0; a1++; }
            const uint8_t *s1 = obs_b + a1 * (uint32_t)ostride + c1 * 3u;
            bits |= ((1ull << s1[0]) | (1ull << (11u + s1[1])) | (1ull << (17u + s1[2]))) << 23;
        }
        const uint32_t w = (uint32_t)(bits >> ch);
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t pos = ch + (uint32_t)j;  // 0..25: 21, 22 = features of this cell; 23.. = next cell
            v[j] = ((w >> j) & 1u) ? 1.0f : 0.0f;
            if (pos == 21u || pos == 22u) v[j] = dir_lut[2u * d0 + (pos - 21u)];
        }
        if (q0 + 4u <= nfl) {
            *(float4 *)(out_b + q0) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            for (uint32_t j = 0; q0 + j < nfl; j++) out_b[q0 + j] = v[j];
        }
    }
}

#endif
// Compact wire format of an observation batch for HOST consumers (PCIe is what bounds the host-buffer path:
// 148 bytes per agent and step for V = 7). Every cell (type < 16, colour < 8, state < 4: core/constants.py:34-97)
// becomes the 9-bit code type | colour << 4 | state << 7; the V*V codes of an agent are packed little-endian,
// cell a*V+b at bits [9*(a*V+b), +9), into mg_packed_obs_stride(V) bytes (56 for V = 7, 96 for V = 9).
// Lossless: multigrid_b200.engine.unpack_obs restores image[V][V][3]. One thread per agent.
#ifdef __CUDACC__
// The same tensor, tile by tile through shared memory: a warp zero-fills a 32-cell tile (32 x 23 floats = 184 float4),
// each lane drops its cell's five non-zero values into its row (1.0 at type, 11 + colour, 17 + state; cos, sin at 21,
// 22) and the warp streams the tile out with 16-byte loads and stores -- ~8 instructions per 16 output bytes instead
// of ~55 for the direct version above, which was instruction-bound at 57 % of the HBM peak. One block = 32 agents
// (32 * V*V cells = V*V whole tiles, so every tile starts on a 16-byte boundary of the output).
__global__ void __launch_bounds__(256) obs_features_tile_kernel(int V, int64_t agents, int ostride, uint32_t rcp_vv,
                                                                const int8_t *__restrict__ obs,
                                                                const int8_t *__restrict__ direction, int dir_stride,
                                                                const float *__restrict__ dir_lut, float *__restrict__ out) {
    __shared__ __align__(16) float tiles[8][LANES * 23];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t VV = (uint32_t)(V * V);
    const int64_t a_base = (int64_t)blockIdx.x * 32;
    const uint32_t na = (uint32_t)(agents - a_base < 32 ? agents - a_base : 32), ncell = na * VV;
    const uint8_t *obs_b = (const uint8_t *)obs + a_base * ostride;
    const int8_t *dir_b = direction + a_base * dir_stride;
    float *out_b = out + a_base * (int64_t)VV * 23;
    float *t = tiles[warp];
    for (uint32_t tl = warp; tl * 32u < ncell; tl += 8u) {
        for (int v = lane; v < LANES * 23 / 4; v += LANES) ((float4 *)t)[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        const uint32_t cl = tl * 32u + (uint32_t)lane;
        if (cl < ncell) {
            const uint32_t a = fastdiv(cl, rcp_vv), cell = cl - a * VV;  // (cl * VV < 2^32: cl < 32 * 225)
            const uint8_t *src = obs_b + a * (uint32_t)ostride + cell * 3u;
            const uint32_t d = (uint32_t)dir_b[a * dir_stride] & 3u;
            float *row = t + lane * 23;
            row[src[0]] = 1.0f;
            row[11u + src[1]] = 1.0f;
            row[17u + src[2]] = 1.0f;
            row[21] = dir_lut[2u * d];
            row[22] = dir_lut[2u * d + 1u];
        }
        __syncwarp();
        const uint32_t left = (ncell - tl * 32u) * 23u, nvalid = left < (uint32_t)(LANES * 23) ? left : (uint32_t)(LANES * 23);
        float *dst = out_b + (size_t)tl * (LANES * 23);
        for (uint32_t v = lane; 4u * v + 4u <= nvalid; v += LANES) ((float4 *)dst)[v] = ((const float4 *)t)[v];
        for (uint32_t q = (nvalid & ~3u) + (uint32_t)lane; q < nvalid; q += LANES) dst[q] = t[q];
        __syncwarp();
    }
}
#endif

#ifdef __CUDACC__
// OneHotObsWrapper.one_hot the same way: a 32-cell tile is 32 x 21 = 672 bytes = 42 16-byte vectors; zero-fill, three
// byte stores per lane, stream out. One block = 32 images (32 * cells whole tiles, 16-byte aligned in the output).
__global__ void __launch_bounds__(256) one_hot_tile_kernel(int cells, int64_t images, int istride, uint32_t rcp_cells,
                                                           const int8_t *__restrict__ obs, uint8_t *__restrict__ out) {
    __shared__ __align__(16) uint8_t tiles[8][LANES * 21];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t VV = (uint32_t)cells;
    const int64_t i_base = (int64_t)blockIdx.x * 32;
    const uint32_t ni = (uint32_t)(images - i_base < 32 ? images - i_base : 32), ncell = ni * VV;
    const uint8_t *obs_b = (const uint8_t *)obs + i_base * istride;
    uint8_t *out_b = out + i_base * (int64_t)VV * 21;
    uint8_t *t = tiles[warp];
    for (uint32_t tl = warp; tl * 32u < ncell; tl += 8u) {
        for (int v = lane; v < LANES * 21 / 16; v += LANES) ((uint4 *)t)[v] = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        const uint32_t cl = tl * 32u + (uint32_t)lane;
        if (cl < ncell) {
            const uint32_t a = fastdiv(cl, rcp_cells), cell = cl - a * VV;  // (cl * cells < 2^32: cl < 32 * 1024)
            const uint8_t *src = obs_b + a * (uint32_t)istride + cell * 3u;
            uint8_t *row = t + lane * 21;
            row[src[0]] = 1;
            row[11u + src[1]] = 1;
            row[17u + src[2]] = 1;
        }
        __syncwarp();
        const uint32_t left = (ncell - tl * 32u) * 21u, nvalid = left < (uint32_t)(LANES * 21) ? left : (uint32_t)(LANES * 21);
        uint8_t *dst = out_b + (size_t)tl * (LANES * 21);
        for (uint32_t v = lane; 16u * v + 16u <= nvalid; v += LANES) ((uint4 *)dst)[v] = ((const uint4 *)t)[v];
        for (uint32_t q = (nvalid & ~15u) + (uint32_t)lane; q < nvalid; q += LANES) dst[q] = t[q];
        __syncwarp();
    }
}
#endif

MG_HD int packed_obs_stride(int V) { return ((9 * V * V + 63) / 64) * 8; }

#ifdef __CUDACC__
template <int VT>
__global__ void __launch_bounds__(128) pack_obs_kernel(int Vr, int64_t agents, int ostride, const int8_t *__restrict__ obs,
                                                       uint8_t *__restrict__ packed) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= agents) return;
    const int V = VT ? VT : Vr, NC = V * V, pstride = packed_obs_stride(V);
    const uint8_t *src = (const uint8_t *)obs + a * ostride;
    uint64_t *dst = (uint64_t *)(packed + a * pstride);
    uint64_t acc = 0;
    int bits = 0, w = 0;
    if constexpr (VT != 0) {
        constexpr int NW = (3 * VT * VT + 3) / 4;
        uint32_t r[NW];
        if ((ostride & 15) == 0) {  // 16-byte slots: vector loads
#pragma unroll
            for (int q = 0; q < (NW + 3) / 4; q++) {
                const uint4 v = *(const uint4 *)(src + 16 * q);
                if (4 * q + 0 < NW) r[4 * q + 0] = v.x;
                if (4 * q + 1 < NW) r[4 * q + 1] = v.y;
                if (4 * q + 2 < NW) r[4 * q + 2] = v.z;
                if (4 * q + 3 < NW) r[4 * q + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < NW; q++) r[q] = *(const uint32_t *)(src + 4 * q);
        }
#pragma unroll
        for (int c = 0; c < VT * VT; c++) {
            const int b0 = 3 * c, b1 = 3 * c + 1, b2 = 3 * c + 2;
            const uint32_t t = (r[b0 >> 2] >> (8 * (b0 & 3))) & 0xff, col = (r[b1 >> 2] >> (8 * (b1 & 3))) & 0xff,
                           st = (r[b2 >> 2] >> (8 * (b2 & 3))) & 0xff;
            acc |= (uint64_t)((t & 15u) | ((col & 7u) << 4) | ((st & 3u) << 7)) << bits;
            bits += 9;
            if (bits >= 64) {
                dst[w++] = acc;
                bits -= 64;
                acc = (uint64_t)((t & 15u) | ((col & 7u) << 4) | ((st & 3u) << 7)) >> (9 - bits);
            }
        }
    } else {
        for (int c = 0; c < NC; c++) {
            const uint32_t code = (src[3 * c] & 15u) | ((src[3 * c + 1] & 7u) << 4) | ((src[3 * c + 2] & 3u) << 7);
            acc |= (uint64_t)code << bits;
            bits += 9;
            if (bits >= 64) {
                dst[w++] = acc;
                bits -= 64;
                acc = (uint64_t)code >> (9 - bits);
            }
        }
    }
    if (bits > 0) dst[w++] = acc;
    for (; w * 8 < pstride; w++) dst[w] = 0;
}
#endif

// Palette wire format: `bits` (1..8) per cell = the index of the cell's 9-bit code in a caller-provided palette
// (lut[code] = index, 0xff = not in the palette -> status |= 8 and index 0). The cell values a batch can show are
// few (Empty-8x8 with 4 agents: unseen, empty, wall, goal and 16 agent encodings = 20 -> 5 bits, 32 instead of 56
// bytes per 7x7 view), and the host path is PCIe-bound.
MG_HD int packed_obs_stride_bits(int V, int bits) { return ((bits * V * V + 63) / 64) * 8; }

#ifdef __CUDACC__
template <int VT>
__global__ void __launch_bounds__(128) pack_obs_palette_kernel(int Vr, int64_t agents, int ostride, const int8_t *__restrict__ obs,
                                                               int pb, const uint8_t *__restrict__ lut,
                                                               uint8_t *__restrict__ packed, int32_t *status) {
    __shared__ uint8_t slut[512];
    for (int q = threadIdx.x; q < 128; q += blockDim.x) ((uint32_t *)slut)[q] = __ldg((const uint32_t *)lut + q);
    __syncthreads();
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= agents) return;
    const int V = VT ? VT : Vr, NC = V * V, pstride = packed_obs_stride_bits(V, pb);
    const uint8_t *src = (const uint8_t *)obs + a * ostride;
    uint64_t *dst = (uint64_t *)(packed + a * pstride);
    uint64_t acc = 0;
    int bits = 0, w = 0;
    uint32_t bad = 0;
    auto put = [&](uint32_t t, uint32_t col, uint32_t st) {
        uint32_t idx = slut[(t & 15u) | ((col & 7u) << 4) | ((st & 3u) << 7)];
        bad |= (uint32_t)(idx == 0xffu);  // the code is not in the palette (at most 255 entries)
        idx = idx == 0xffu ? 0u : idx;
        acc |= (uint64_t)idx << bits;
        bits += pb;
        if (bits >= 64) {
            dst[w++] = acc;
            bits -= 64;
            acc = bits ? (uint64_t)idx >> (pb - bits) : 0ull;
        }
    };
    if constexpr (VT != 0) {
        constexpr int NW = (3 * VT * VT + 3) / 4;
        uint32_t r[NW];
        if ((ostride & 15) == 0) {  // 16-byte slots: vector loads
#pragma unroll
            for (int q = 0; q < (NW + 3) / 4; q++) {
                const uint4 v = *(const uint4 *)(src + 16 * q);
                if (4 * q + 0 < NW) r[4 * q + 0] = v.x;
                if (4 * q + 1 < NW) r[4 * q + 1] = v.y;
                if (4 * q + 2 < NW) r[4 * q + 2] = v.z;
                if (4 * q + 3 < NW) r[4 * q + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < NW; q++) r[q] = *(const uint32_t *)(src + 4 * q);
        }
#pragma unroll
        for (int c = 0; c < VT * VT; c++) {
            const int b0 = 3 * c, b1 = 3 * c + 1, b2 = 3 * c + 2;
            put((r[b0 >> 2] >> (8 * (b0 & 3))) & 0xff, (r[b1 >> 2] >> (8 * (b1 & 3))) & 0xff, (r[b2 >> 2] >> (8 * (b2 & 3))) & 0xff);
        }
    } else {
        for (int c = 0; c < NC; c++) put(src[3 * c], src[3 * c + 1], src[3 * c + 2]);
    }
    if (bits > 0) dst[w++] = acc;
    for (; w * 8 < pstride; w++) dst[w] = 0;
    if (bad) status_or(status, 8);
}
#endif

// Compact per-env record of the host wire (mg_step_obs_host_wire): a step's rewards are 0 or sums of ONE value per
// env -- `1 - 0.9 * step_count / max_steps` (base.py:598-602), the same for every agent rewarded in that step; the
// LockedHallway hook adds it once per door unlocked (envs/locked_hallway.py:203-227) -- so n float64 travel as
//   { float64 value; uint32 terminated mask | truncated << 31; uint32 counts[ceil(n / 8)] (4 bits per agent) }
// with reward[j] = value added counts[j] times (0 -> 0.0). The kernel CHECKS that this reproduces every reward bit
// for bit and sets bit 4 (value 16) of the status word otherwise.
MG_HD int wire_record_bytes(int n) { return (8 + 4 + 4 * ((n + 7) / 8) + 7) & ~7; }

MG_HD bool wire_env_record(int n, const double *reward, const uint8_t *terminated, uint8_t truncated, uint8_t *rec) {
    double v = 0.0;
    for (int j = 0; j < n; j++)
        if (reward[j] != 0.0 && (v == 0.0 || (reward[j] < v) == (v > 0.0))) v = reward[j];  // the smallest magnitude
    uint32_t tmask = (uint32_t)(truncated != 0) << 31;
    bool ok = true;
    uint32_t *counts = (uint32_t *)(rec + 12);
    for (int w = 0; w < (n + 7) / 8; w++) counts[w] = 0;
    for (int j = 0; j < n; j++) {
        tmask |= (uint32_t)(terminated[j] != 0) << j;
        double acc = 0.0;
        uint32_t c = 0;
        while (acc != reward[j] && c < 15u) { acc += v; c++; }
        ok &= acc == reward[j];
        counts[j >> 3] |= c << (4 * (j & 7));
    }
    *(double *)rec = v;
    *(uint32_t *)(rec + 8) = tmask;
    return ok;
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(128) wire_env_records_kernel(int n, int64_t E, const double *__restrict__ reward,
                                                               const uint8_t *__restrict__ terminated,
                                                               const uint8_t *__restrict__ truncated,
                                                               uint8_t *__restrict__ records, int32_t *status) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    if (!wire_env_record(n, reward + e * n, terminated + e * n, truncated[e], records + e * wire_record_bytes(n)))
        status_or(status, 16);
}
#endif

#ifdef __CUDACC__
// MULTI = mg_rollout (p.T steps per launch); the single-step kernels compile with T == 1 and no loop.
// Same result, 16 output bytes per thread (one 128-bit store): the 16 bytes [p0, p0+16) of the flat
// [A][V][V][21] stream lie in at most two consecutive cells, i.e. hold at most six 1-bytes; each is
// placed with one shift into a 16-bit mask that a multiply spreads into bytes. The output is 7x the
// observation (270 MB for the 65 536 x 4 x 7 x 7 bench batch): HBM writes are the floor.
// One block = 16 consecutive agents = VV*21 16-byte vectors: every index below is a small 32-bit number.
__global__ void __launch_bounds__(256) one_hot_kernel_v16(int cells, int64_t agents, int ostride, uint32_t rcp_vv,
                                                          const int8_t *__restrict__ obs, uint4 *__restrict__ out) {
    const uint32_t VV = (uint32_t)cells, nvec = VV * 21u;            // vectors per full block
    const int64_t a0 = (int64_t)blockIdx.x * 16;
    const uint32_t na = (uint32_t)(agents - a0 < 16 ? agents - a0 : 16);
    const uint32_t nbytes = na * VV * 21u;                            // bytes of this block (tail block: fewer)
    const uint8_t *obs_b = (const uint8_t *)obs + a0 * ostride;
    uint4 *out_b = out + (int64_t)blockIdx.x * nvec;
    auto shl = [](uint32_t q) { uint32_t r; asm("shl.b32 %0, 1, %1;" : "=r"(r) : "r"(q)); return r; };
    for (uint32_t v = threadIdx.x; 16u * v < nbytes; v += 256u) {
        const uint32_t rel = 16u * v, c = mulhi32(rel, 204522253u), off = rel - c * 21u;  // rel / 21, rel < 2^21
        uint32_t a = fastdiv(c, rcp_vv), cia = c - a * VV;
        // bit q of `mask` <=> output byte q is 1; shifts by >= 32 (positions outside the window, including
        // "negative" ones that wrapped) give 0 in PTX, positions 16..31 are masked off by the spread below
        const uint8_t *src = obs_b + a * (uint32_t)ostride + cia * 3u;
        uint32_t base = 0u - off;
        uint32_t mask = shl(base + src[0]) | shl(base + 11u + src[1]) | shl(base + 17u + src[2]);
        if (off > 5u && rel + (21u - off) < nbytes) {  // the window reaches into the next cell
            if (++cia == VV) { cia = 0; a++; }
            src = obs_b + a * (uint32_t)ostride + cia * 3u;
            base = 21u - off;
            mask |= shl(base + src[0]) | shl(base + 11u + src[1]) | shl(base + 17u + src[2]);
        }
        if (rel + 16u <= nbytes) {
            // 4 mask bits -> 4 bytes of 0/1: the multiply puts bit i at bit 8*i (no two products collide)
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; k++) w[k] = (((mask >> (4 * k)) & 15u) * 0x00204081u) & 0x01010101u;
            out_b[v] = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            uint8_t *o = (uint8_t *)out_b + rel;
            for (uint32_t b = 0; rel + b < nbytes; b++) o[b] = (uint8_t)((mask >> b) & 1u);
        }
    }
}

// CHAIN = the launch takes part in the chain tickets (MG_FLAG_CHAINED); plain launches compile without them.
// NT / HK: the agent count and the post-hook as compile-time constants (0 / -1 = read them from Params). The phase
// functions take `p.n` and `p.hook` from a kernel-local copy of Params whose two fields are overwritten with the
// constants, so after inlining the agent loops unroll and the other env classes' hooks disappear from the hot path.
// OH: the launch also writes the one-hot images (MgStepOut.one_hot) -- separate instantiations, so that the plain
// launch's code, registers and spills are exactly what they are without the feature.
// ROOMY: compiled for 4 blocks per SM (up to 128 registers: no spills) instead of 7 (72 registers). For launches whose
// blocks are all resident at 4 per SM anyway -- BASELINE configs[2], BlockedUnlockPickup x 32 768, is half a wave --
// the register cap only costs: 9.7 -> 8.5 us per launch.
template <int VT, int MODE, bool MULTI = false, bool CHAIN = false, int NT = 0, int HK = -1, bool OH = false, bool ROOMY = false>
__global__ void __launch_bounds__(128, (VT >= 9 || ROOMY) ? 4 : 7) step_obs_kernel(const __grid_constant__ Params p_in) {
    Params p = p_in;
    if (NT > 0) { p.n = NT; p.rcp_n = NT <= 1 ? 0u : (uint32_t)((1ull << 32) / (uint32_t)NT + 1ull); }
    if (HK >= 0) p.hook = HK;
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int group = blockIdx.x * p.wpb + warp;
    constexpr bool tickets = CHAIN;
    const bool live = group * p.G < p.num_envs;  // (whole warp)
    // Chain tickets (p.chain[e] = {next ticket, tickets done, grid dirty, -}). Every chained launch takes, per env,
    // the next ticket BEFORE it lets its dependents launch, so tickets follow launch order: the dependent grid's
    // blocks only start once all blocks of this grid have passed launch_dependents. The trigger is per BLOCK (the
    // first thread that executes it counts), so it is executed by the LAST warp of the block to have claimed
    // (a shared counter). An env may be touched once `done` equals the ticket, i.e. once the previous launch has
    // finished THAT env (its stores complete, then a release store). A warp whose envs are free (the common case)
    // issues its loads first and claims while they are in flight.
    unsigned &claimed = *(unsigned *)(smem + p.wpb * p.warp_bytes);  // (16 bytes behind the warps' regions)
    if (tickets) {
        if (threadIdx.x == 0) claimed = 0;
        __syncthreads();  // (the only block barrier: every warp has just started)
    } else {
        pdl_launch_dependents();
    }
    if (!live) {
        if (tickets && lane == 0 && atomicAdd(&claimed, 1u) == blockDim.x / 32 - 1) pdl_launch_dependents();
        return;
    }
    uint8_t *ws = smem + warp * p.warp_bytes;
    const Group g = group_view(p, ws, group);
    uint64_t *bar = (uint64_t *)(ws + p.off_mbar);
    // TMA needs 16-byte multiples: full groups only (G % 16 == 0 makes every span aligned)
    const bool bulk = p.use_bulk && g.ne == p.G;
    const int env = lane_env(p, g, lane);
    trace_mark(p, group, lane, 0);
    trace_mark(p, group, lane, 7);
    if (bulk && lane == 0) mbar_init(bar, 1);
    if (lane == 0) *(uint32_t *)(ws + p.off_mbar + 8) = 0u;  // the join counter of the transition phase (below)
    uint32_t ticket = 0;
    bool rec_dirty = false;  // (tickets: the dirty flag arrives with the ticket)
    if (tickets) {
        if (p.chained == 2) pdl_wait();  // head of a chain: the whole previous grid first, like a plain launch
        uint32_t *slot = p.chain + 4 * (size_t)(g.e0 + (env >= 0 ? env : 0));
        uint4 rec = ld_acquire_gpu_v4(slot);
        ticket = rec.x;
        while (!__all_sync(0xffffffffu, env < 0 || rec.y == ticket)) {
            __nanosleep(64);
            rec = ld_acquire_gpu_v4(slot);  // (next is ours until we claim: only done / dirty can change)
        }
        rec_dirty = env >= 0 && rec.z != 0;
        fence_async_global();  // what the acquire made visible is also visible to the TMA loads below
        __syncwarp();
    } else {
        pdl_wait();  // plain launch: nothing of the previous launch is read or overwritten before this point
    }
    EnvRegs er;
    const int T = MULTI ? p.T : 1;
#pragma unroll 1
    for (int t = 0; t < T; t++) {
        const size_t tE = (size_t)t * (size_t)p.num_envs;
        trace_mark(p, group, lane, 5);
        const bool dedup = !MULTI && MODE != MODE_OBS && bulk && p.pool_rep != nullptr;
        uint32_t dirty_mask = 0xffffffffu;  // envs whose grid may differ from the (single) pool layout
        if (bulk) {
            if (lane == 0) {
                if (t > 0) bulk_wait_read();  // the last obs store has read its stage (which lies on the cells)
                load_bulk<MODE>(p, g, bar, t, dedup ? 1 : 0);
            }
        } else {
            phase_load_plain<MODE>(p, g, lane, t);
        }
        if (dedup) {  // one flag per env decides where the group's cells come from
            const bool env_dirty = tickets ? rec_dirty : (env >= 0 && p.chain[4 * (size_t)(g.e0 + env) + 2] != 0);
            dirty_mask = __ballot_sync(0xffffffffu, env_dirty);  // (bit i = env i)
            if (lane == 0) load_bulk<MODE>(p, g, bar, t, dirty_mask ? 3 : 2);
        }
        if (tickets && t == 0) {
            // claim (performed device-wide) while the loads are in flight; the last warp of the block to get
            // here lets the dependents launch
            if (env >= 0) p.chain[4 * (size_t)(g.e0 + env)] = ticket + 1u;
            __threadfence();
            __syncwarp();
            if (lane == 0 && atomicAdd(&claimed, 1u) == blockDim.x / 32 - 1) pdl_launch_dependents();
        }
        if (t == 0) env_load<MODE>(p, g, env, er);  // the env's scalars, straight into its lane's registers
        const OrderDraw draw = phase_draw<MODE>(p, g, env, er);
        __syncwarp();
        if (bulk) mbar_wait(bar, (uint32_t)(t & 1));
        trace_mark(p, group, lane, 1);
        if (MODE != MODE_OBS && (p.flags & MG_FLAG_AUTO_RESET)) {
            phase_reset(p, g, env, er);
            const uint32_t pending = __ballot_sync(0xffffffffu, env >= 0 && g.rk[env] >= 0) & (p.G == 32 ? 0xffffffffu : (1u << p.G) - 1u);
            // single layout: a clean env that resets already holds the layout (in shared memory and in HBM)
            const uint32_t to_copy = pending & dirty_mask;
            if (to_copy) {
                __syncwarp();
                phase_reset_grid(p, g, to_copy, lane);
            }
            __syncwarp();
        }
        if (MODE != MODE_OBS && p.G < LANES && p.n > 2) {
            // Fewer envs than lanes: the env lanes run the (long, serial) transition while the others have nothing
            // to do. If the idle lanes simply branched around it they would sit at the region's convergence barrier
            // for microseconds, and the hardware then lets the two halves continue as separate groups for the rest
            // of the kernel -- the whole observation phase at half width, twice the instructions (ncu: 16.0 active
            // threads per instruction, profiles/r02_summary.md). So the idle lanes poll a shared counter instead and
            // reach the barrier together with the last env lane. (With 2 agents the region is short enough for the
            // halves to merge on their own, and the polling only costs: measured +0.3 us on BlockedUnlockPickup.)
            uint32_t *join = (uint32_t *)(ws + p.off_mbar + 8);
            if (env >= 0) {
                phase_step<MODE>(p, g, env, er, draw, tE);
                atomicAdd(join, 1u);
            } else {
                const uint32_t want = (uint32_t)g.ne * (uint32_t)(t + 1);
                while (atomicAdd(join, 0u) < want) __nanosleep(64);
            }
        } else {
            phase_step<MODE>(p, g, env, er, draw, tE);
        }
        __syncwarp();
        trace_mark(p, group, lane, 2);
        if (MODE != MODE_STEP) {
            const int passes = obs_passes(p, g);
            for (int pass = 0; pass < passes; pass++) {
                const ObsTask tk = obs_task(p, g, pass, lane);
                uint8_t *stage = stage_of(p, g, pass), *out = stage + lane * p.ostride;
                if constexpr (VT != 0) {
                    uint32_t cr[VT ? VT * VT : 1];
                    if (tk.valid) obs_compute<VT>(p, tk.cells, tk.a0, tk.a1, cr);
                    // the previous pass's TMA store must be done reading its stage, and (aliased stage)
                    // every lane must be done gathering before the cells under the stage are overwritten
                    if (bulk && pass > 0 && lane == 0) bulk_wait_read();
                    __syncwarp();
                    if (tk.valid) obs_pack_store<VT>(p, cr, out);
                } else {
                    if (bulk && pass > 0) {
                        if (lane == 0) bulk_wait_read();
                        __syncwarp();
                    }
                    if (tk.valid) obs_agent_generic(p, tk.cells, tk.a0, tk.a1, out);
                }
                if (bulk) {
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        const int left = g.ne * p.n - pass * LANES;
                        const uint32_t cnt = left < LANES ? left : LANES;
                        int8_t *dst = p.obs + ((tE + (size_t)g.e0) * p.n + (size_t)pass * LANES) * p.ostride;
                        if (p.l2hint & 2) bulk_s2g_hint(dst, stage, cnt * p.ostride, l2_policy_evict_first());
                        else bulk_s2g(dst, stage, cnt * p.ostride);
                        bulk_commit();
                    }
                    if constexpr (OH) phase_one_hot(p, g, pass, lane);  // (under the TMA store: both only read the stage)
                } else {
                    __syncwarp();
                    phase_obs_store_plain(p, g, pass, lane, tE);
                    if constexpr (OH) phase_one_hot(p, g, pass, lane);
                    __syncwarp();
                }
            }
        }
        trace_mark(p, group, lane, 3);
        if (t + 1 < T) {
            // this step's writes to the grid in HBM (dirty cells, reset layouts: generic proxy) must be
            // visible to the next step's TMA load of the cells (async proxy)
            if (bulk) fence_async_global();
            __syncwarp();
            trace_mark(p, group, lane, 6);
        }
    }
    if (MODE != MODE_OBS) {
        if (bulk) {
            if (MODE == MODE_STEP) fence_async_smem();  // (the obs passes already fenced)
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(p.agents + (size_t)g.e0 * p.n * 8, g.ag, (uint32_t)(p.G * p.n * 8));
                bulk_commit();
            }
        } else {
            phase_store_plain(p, g, lane);
        }
    }
    if (tickets) {
        // publish: this launch is done with its envs once every store of the warp is complete -- the TMA stores
        // (full completion, not only their shared-memory reads) and the lanes' own global stores
        if (bulk && lane == 0) bulk_wait_all();
        __syncwarp();
        __threadfence();
        if (env >= 0) st_release_gpu(p.chain + 4 * (size_t)(g.e0 + env) + 1, ticket + 1u);
        // completion order: a chained launch did not wait for its predecessor when it started; it must not
        // COMPLETE before it either, or the next unchained operation of the stream could overtake that grid
        if (p.chained == 1) pdl_wait();
    } else if (bulk && lane == 0) {
        bulk_wait_read();  // smem must stay valid until the TMA stores have read it
    }
    trace_mark(p, group, lane, 4);
}
#endif

}  // namespace mg
