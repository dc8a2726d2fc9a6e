// multigrid_b200 -- device code of the batched MultiGrid step/observe engine (sm_100a).
//
// One thread block advances `epb` consecutive envs. All of a block's state is one contiguous HBM
// span per array (env-major layout), staged through shared memory with 16-byte accesses; the
// per-env work then runs out of shared memory in bulk-synchronous phases:
//
//   P0 load      block-wide copy: raw grid span + agent span -> smem            (all threads)
//   P0b reset    auto-reset decision per env, pool layout fetch                 (1 thread / env)
//   P1 convert   3-byte cells -> 32-bit cell words (+ "opaque" bit for the scan) (tpe threads / env)
//   P2 step      handle_actions: PCG64 draw, argsort, serial agent loop, rewards,
//                termination, dirty-cell write-through, agent stamping, env hook (1 thread / env)
//   P3 observe   per-agent view gather (closed-form slice+rotate), row-bitmask
//                visibility scan, masking, 24->32 bit packing into smem          (1 thread / agent)
//   P4 store     block-wide copy: obs span + agent span -> HBM                   (all threads)
//
// Reference semantics restated here (cited inline): multigrid/base.py:303-532,598-602,
// multigrid/utils/obs.py:46-316, multigrid/core/world_object.py:197-233,452-474,599-605.
//
// The file also compiles as plain C++ (no __CUDACC__): tests/hostsim runs the very same phase
// functions thread-by-thread on the CPU to check the logic against the oracle without a GPU.
// That build is test infrastructure; the product only ever launches the CUDA kernels.
#pragma once
#include <stdint.h>
#include "multigrid_b200.h"

#ifdef __CUDACC__
#define MG_HD __host__ __device__ __forceinline__
#else
#define MG_HD inline
#endif

namespace mg {

enum : int { T_UNSEEN = 0, T_EMPTY, T_WALL, T_FLOOR, T_DOOR, T_KEY, T_BALL, T_BOX, T_GOAL, T_LAVA, T_AGENT };
enum : int { S_OPEN = 0, S_CLOSED, S_LOCKED };
enum : int { ACT_LEFT = 0, ACT_RIGHT, ACT_FORWARD, ACT_PICKUP, ACT_DROP, ACT_TOGGLE, ACT_DONE };
enum : int { MODE_OBS = 0, MODE_STEP = 1, MODE_STEP_OBS = 2 };

// 32-bit cell word: type | color<<8 | state<<16 | opaque<<24
constexpr uint32_t CELL_EMPTY = T_EMPTY;
constexpr uint32_t CELL_WALL = T_WALL | (5u << 8) | (1u << 24);  // WALL_ENCODING, utils/obs.py:14
constexpr uint32_t OPAQUE_BIT = 1u << 24;

struct u4 { uint32_t x, y, z, w; };

struct Params {
    // config
    int32_t W, H, n, V, max_steps;
    uint32_t flags;
    int32_t hook, ostride, K, lstride;
    int32_t num_envs, epb, tpe;
    // state (device)
    int8_t *grid; int8_t *agents; int32_t *step_count; uint64_t *pcg_state; const uint64_t *pcg_inc;
    int32_t *layout_idx; const int8_t *pool_grid; const int8_t *pool_agents;
    const int8_t *actions;
    // outputs (device)
    int8_t *obs; double *reward; uint8_t *terminated; uint8_t *truncated; int32_t *status;
    // shared-memory carve-up (byte offsets, all multiples of 16)
    int32_t cstride;  // words per env in the cell array (W*H rounded up to odd)
    int32_t off_cells, off_stage, off_agents, off_keys, off_order, off_sc, off_rk, smem_bytes;
};

MG_HD int align16(int x) { return (x + 15) & ~15; }

// Fills the derived fields; returns the dynamic shared memory size for (epb, tpe).
inline int carve_smem(Params &p) {
    const int WH = p.W * p.H;
    p.cstride = WH | 1;
    int raw = p.epb * WH * 3, stage = p.epb * p.n * p.ostride;
    int off = 0;
    p.off_cells = off;  off += align16(p.epb * p.cstride * 4);
    p.off_stage = off;  off += align16(raw > stage ? raw : stage);
    p.off_agents = off; off += align16(p.epb * p.n * 8);
    p.off_keys = off;   off += align16(p.epb * p.n * 8);
    p.off_order = off;  off += align16(p.epb * p.n);
    p.off_sc = off;     off += align16(p.epb * 4);
    p.off_rk = off;     off += align16(p.epb * 4);
    p.smem_bytes = off;
    return off;
}

// Launch geometry: tpe threads per env (one per agent, at most 8), epb envs per block (a multiple
// of 16 so every block's spans start 16-byte aligned), epb*tpe <= max_threads.
// Returns 0, or MG_ERR_TOO_LARGE when even 16 envs do not fit `smem_budget`.
inline int plan_launch(Params &p, int forced_epb, int max_threads, int smem_budget) {
    p.tpe = p.n < 8 ? p.n : 8;
    int epb = 16 * (max_threads / (16 * p.tpe));
    if (epb > 128) epb = 128;
    if (forced_epb > 0 && forced_epb % 16 == 0 && forced_epb * p.tpe <= max_threads) epb = forced_epb;
    for (; epb >= 16; epb -= 16) {
        p.epb = epb;
        if (carve_smem(p) <= smem_budget) return 0;
    }
    return MG_ERR_TOO_LARGE;
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

MG_HD uint32_t bitrev(uint32_t v, int nbits) {  // reverse the low `nbits` bits
#ifdef __CUDA_ARCH__
    return __brev(v) >> (32 - nbits);
#else
    uint32_t r = 0;
    for (int i = 0; i < nbits; i++) r |= ((v >> i) & 1u) << (nbits - 1 - i);
    return r;
#endif
}

MG_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

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

// Block-wide copy of a contiguous span. 16-byte vectors when both ends allow it.
MG_HD void coop_copy(void *dst, const void *src, int nbytes, int tid, int nt) {
    if ((((uintptr_t)dst | (uintptr_t)src) & 15u) == 0) {
        const int nv = nbytes >> 4;
        u4 *d = (u4 *)dst; const u4 *s = (const u4 *)src;
        for (int v = tid; v < nv; v += nt) d[v] = s[v];
        for (int b = (nv << 4) + tid; b < nbytes; b += nt) ((uint8_t *)dst)[b] = ((const uint8_t *)src)[b];
    } else {
        for (int b = tid; b < nbytes; b += nt) ((uint8_t *)dst)[b] = ((const uint8_t *)src)[b];
    }
}

MG_HD uint32_t cell_word(uint32_t t, uint32_t c, uint32_t s) {
    // see_behind (utils/obs.py:47-63): walls and non-open doors block the view
    uint32_t opaque = (t == T_WALL) | ((t == T_DOOR) & (s != S_OPEN));
    return t | (c << 8) | (s << 16) | (opaque << 24);
}

struct Block {
    int e0, ne;  // first global env of this block, number of valid envs
    uint32_t *cells; uint8_t *stage; uint32_t *ag; uint64_t *keys; uint8_t *order; int32_t *sc; int32_t *rk;
};

MG_HD Block block_view(const Params &p, uint8_t *smem, int blk) {
    Block b;
    b.e0 = blk * p.epb;
    b.ne = p.num_envs - b.e0 < p.epb ? p.num_envs - b.e0 : p.epb;
    b.cells = (uint32_t *)(smem + p.off_cells);
    b.stage = smem + p.off_stage;
    b.ag = (uint32_t *)(smem + p.off_agents);
    b.keys = (uint64_t *)(smem + p.off_keys);
    b.order = smem + p.off_order;
    b.sc = (int32_t *)(smem + p.off_sc);
    b.rk = (int32_t *)(smem + p.off_rk);
    return b;
}

// ---- P0: load ------------------------------------------------------------------------------------
template <int MODE>
MG_HD void phase_load(const Params &p, uint8_t *smem, int blk, int tid, int nt) {
    Block b = block_view(p, smem, blk);
    const int WH3 = p.W * p.H * 3;
    coop_copy(b.stage, p.grid + (size_t)b.e0 * WH3, b.ne * WH3, tid, nt);
    coop_copy(b.ag, p.agents + (size_t)b.e0 * p.n * 8, b.ne * p.n * 8, tid, nt);
    for (int i = tid; i < b.ne; i += nt) {
        b.rk[i] = -1;
        if (MODE != MODE_OBS) b.sc[i] = p.step_count[b.e0 + i];
    }
}

// ---- P0b: auto-reset decision ("next-step" mode; is_done = base.py:534-539) -------------------------
MG_HD void phase_reset(const Params &p, uint8_t *smem, int blk, int tid, int nt) {
    Block b = block_view(p, smem, blk);
    for (int i = tid; i < b.ne; i += nt) {
        uint32_t all_term = 1;
        for (int j = 0; j < p.n; j++) all_term &= ((b.ag[(i * p.n + j) * 2] >> 24) & 0xff) != 0;
        if (all_term || b.sc[i] >= p.max_steps) {
            const int e = b.e0 + i;
            int k = (int)(((int64_t)p.layout_idx[e] + p.lstride) % p.K);
            p.layout_idx[e] = k;
            b.rk[i] = k;
            b.sc[i] = 0;
            const uint32_t *src = (const uint32_t *)(p.pool_agents + (size_t)k * p.n * 8);
            for (int j = 0; j < p.n * 2; j++) b.ag[i * p.n * 2 + j] = src[j];
        }
    }
}

// ---- P1: 3-byte cells -> cell words ----------------------------------------------------------------
MG_HD void phase_convert(const Params &p, uint8_t *smem, int blk, int tid, int nt) {
    Block b = block_view(p, smem, blk);
    const int WH = p.W * p.H;
    const int i = tid / p.tpe, lane = tid - i * p.tpe;
    if (i >= b.ne) return;
    const int k = b.rk[i];
    uint32_t *dst = b.cells + i * p.cstride;
    if (k < 0) {
        const uint8_t *src = b.stage + i * WH * 3;
        for (int c = lane; c < WH; c += p.tpe)
            dst[c] = cell_word(src[3 * c], src[3 * c + 1], src[3 * c + 2]);
    } else {  // env was reset: take the layout from the pool and write it through to the state
        const uint8_t *src = (const uint8_t *)p.pool_grid + (size_t)k * WH * 3;
        uint8_t *g = (uint8_t *)p.grid + (size_t)(b.e0 + i) * WH * 3;
        for (int c = lane; c < WH; c += p.tpe) {
            uint8_t t = src[3 * c], col = src[3 * c + 1], s = src[3 * c + 2];
            g[3 * c] = t; g[3 * c + 1] = col; g[3 * c + 2] = s;
            dst[c] = cell_word(t, col, s);
        }
    }
}

// ---- P2: transition --------------------------------------------------------------------------------
MG_HD void store_cell(const Params &p, int e, int idx, uint32_t w) {  // dirty-cell write-through
    uint8_t *g = (uint8_t *)p.grid + ((size_t)e * p.W * p.H + idx) * 3;
    g[0] = (uint8_t)w; g[1] = (uint8_t)(w >> 8); g[2] = (uint8_t)(w >> 16);
}

MG_HD void on_success(const Params &p, uint32_t *ag, int k, int e, int32_t sc) {  // base.py:478-507
    if (p.flags & MG_FLAG_SUCCESS_ANY) {
        for (int j = 0; j < p.n; j++) ag[j * 2] |= 1u << 24;
    } else {
        ag[k * 2] |= 1u << 24;
    }
    const double r = reward_value(sc, p.max_steps);
    if (p.flags & MG_FLAG_JOINT_REWARD) {
        for (int j = 0; j < p.n; j++) p.reward[(size_t)e * p.n + j] = r;
    } else {
        p.reward[(size_t)e * p.n + k] = r;
    }
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

// MultiGridEnv.handle_actions (base.py:378-476) for local env i; `ag` = this env's agent words.
MG_HD void handle_actions(const Params &p, const Block &b, int i, int e, uint32_t *cells,
                          uint32_t *ag, int32_t sc) {
    const int n = p.n, H = p.H, epb = p.epb;
    if (n > 1) {  // base.py:399: order = np_random.random(size=n).argsort()
        uint64_t lo = p.pcg_state[2 * (size_t)e], hi = p.pcg_state[2 * (size_t)e + 1];
        const uint64_t ilo = p.pcg_inc[2 * (size_t)e], ihi = p.pcg_inc[2 * (size_t)e + 1];
        for (int j = 0; j < n; j++) b.keys[j * epb + i] = pcg64_next53(lo, hi, ilo, ihi);
        p.pcg_state[2 * (size_t)e] = lo; p.pcg_state[2 * (size_t)e + 1] = hi;
        for (int j = 0; j < n; j++) {  // rank = position in the ascending (stable) order
            const uint64_t kj = b.keys[j * epb + i];
            int r = 0;
            for (int q = 0; q < n; q++) {
                const uint64_t kq = b.keys[q * epb + i];
                r += (kq < kj) | ((kq == kj) & (q < j));
            }
            b.order[r * epb + i] = (uint8_t)j;
        }
    } else {
        b.order[i] = 0;  // base.py:396-397
    }
    for (int r = 0; r < n; r++) {
        const int k = b.order[r * epb + i];
        const int act = p.actions[(size_t)e * n + k];
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
        if ((unsigned)fx >= (unsigned)p.W || (unsigned)fy >= (unsigned)H) continue;
        const int idx = fx * H + fy;
        const uint32_t cw = cells[idx];
        const uint32_t t = cw & 0xff, col = (cw >> 8) & 0xff, st = (cw >> 16) & 0xff;
        const uint32_t fxy = (uint32_t)fx | ((uint32_t)fy << 8);
        if (act == ACT_FORWARD) {  // base.py:420-436
            const bool can_overlap = (t == T_EMPTY) | (t == T_FLOOR) | (t == T_GOAL) | (t == T_LAVA) |
                                     ((t == T_DOOR) & (st == S_OPEN));
            if (!can_overlap) continue;
            if (!(p.flags & MG_FLAG_ALLOW_OVERLAP) && agent_at(p, ag, fxy)) continue;
            ag[k * 2] = (a0 & 0xff0000ffu) | (fxy << 8);
            if (t == T_GOAL) on_success(p, ag, k, e, sc);
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
MG_HD void stamp_agents(const Params &p, uint32_t *cells, const uint32_t *ag) {
    if (p.n <= 1) return;  // utils/obs.py:172-173
    for (int j = 0; j < p.n; j++) {
        const uint32_t a0 = ag[j * 2], a1 = ag[j * 2 + 1];
        if ((a0 >> 24) & 0xff) continue;
        const int x = (a0 >> 8) & 0xff, y = (a0 >> 16) & 0xff;
        if ((unsigned)x >= (unsigned)p.W || (unsigned)y >= (unsigned)p.H) continue;
        cells[x * p.H + y] = T_AGENT | ((a1 >> 24) << 8) | ((a0 & 0xff) << 16);
    }
}

template <int MODE>
MG_HD void phase_step(const Params &p, uint8_t *smem, int blk, int tid, int nt) {
    Block b = block_view(p, smem, blk);
    for (int i = tid; i < b.ne; i += nt) {
        const int e = b.e0 + i, n = p.n;
        uint32_t *cells = b.cells + i * p.cstride;
        uint32_t *ag = b.ag + i * n * 2;
        if constexpr (MODE == MODE_OBS) {
            stamp_agents(p, cells, ag);
        } else {
            for (int j = 0; j < n; j++) p.reward[(size_t)e * n + j] = 0.0;  // base.py:394
            bool truncated = false;
            if (b.rk[i] < 0) {
                const int32_t sc = b.sc[i] + 1;  // base.py:333
                b.sc[i] = sc;
                handle_actions(p, b, i, e, cells, ag, sc);
                truncated = sc >= p.max_steps;  // base.py:339
                if (MODE == MODE_STEP_OBS) stamp_agents(p, cells, ag);  // obs sees pre-hook termination
                if (p.hook == MG_HOOK_BLOCKED_UNLOCK_PICKUP) {  // envs/blockedunlockpickup.py:166-175
                    for (int k = 0; k < n; k++)
                        if ((ag[k * 2 + 1] & 0xff) == T_BOX) on_success(p, ag, k, e, sc);
                }
            } else if (MODE == MODE_STEP_OBS) {
                stamp_agents(p, cells, ag);
            }
            for (int j = 0; j < n; j++)
                p.terminated[(size_t)e * n + j] = (uint8_t)(((ag[j * 2] >> 24) & 0xff) != 0);
            p.truncated[e] = (uint8_t)truncated;
            p.step_count[e] = b.sc[i];
        }
    }
}

// ---- P3: observation of one agent -------------------------------------------------------------------
// Closed form of get_view_exts + the rotation loop (utils/obs.py:175-202, 276-316):
//   obs[a][b] = G'[pos + f*(V-1-b) + r*(a - V/2)],  f = DIR_TO_VEC[dir], r = (-f.y, f.x),
// out of bounds -> wall. Visibility (utils/obs.py:236-273) as one bitmask per view row b.
template <int VT>
MG_HD void obs_agent(const Params &p, const uint32_t *cells, uint32_t a0, uint32_t a1, uint8_t *out) {
    const int V = VT ? VT : p.V, half = V >> 1, W = p.W, H = p.H;
    const uint32_t dir = a0 & 3u;
    const int px = (a0 >> 8) & 0xff, py = (a0 >> 16) & 0xff;
    const bool horiz = !(dir & 1u);               // forward axis is x for right/left
    const int sf = (dir & 2u) ? -1 : 1;           // forward sign
    const int sl = (dir == 0u || dir == 3u) ? 1 : -1;  // lateral sign (r = (-f.y, f.x))
    const int pf = horiz ? px : py, Lf = horiz ? W : H;
    const int pl = horiz ? py : px, Ll = horiz ? H : W;
    const int df = sf * (horiz ? H : 1), dl = sl * (horiz ? 1 : H);
    const int idx0 = px * H + py - dl * half;
    const uint32_t full = (1u << V) - 1u;
    const bool stw = (p.flags & MG_FLAG_SEE_THROUGH_WALLS) != 0;
    const uint32_t carry = a1 & 0x00ffffffu;      // utils/obs.py:207

    uint32_t colok = 0;                           // lateral coordinate in range, per view column a
#pragma unroll
    for (int a = 0; a < V; a++) colok |= (uint32_t)((unsigned)(pl + sl * (a - half)) < (unsigned)Ll) << a;

    uint32_t cr[VT ? VT * VT : 1];
    uint32_t vis = 1u << half;                    // vis_mask[V//2][V-1] = True (utils/obs.py:252)
#pragma unroll
    for (int b = V - 1; b >= 0; b--) {
        const int d = V - 1 - b;
        const bool rowok = (unsigned)(pf + sf * d) < (unsigned)Lf;
        const int ridx = idx0 + df * d;
        uint32_t see = 0;
#pragma unroll
        for (int a = 0; a < V; a++) {
            uint32_t c = CELL_WALL;
            if (rowok && ((colok >> a) & 1u)) c = cells[ridx + dl * a];
            if (b == V - 1 && a == half) c = carry;
            see |= (((c >> 24) & 1u) ^ 1u) << a;
            if (VT) {
                cr[VT ? a * VT + b : 0] = c;
            } else {
                uint8_t *o = out + (a * V + b) * 3;
                o[0] = (uint8_t)c; o[1] = (uint8_t)(c >> 8); o[2] = (uint8_t)(c >> 16);
            }
        }
        uint32_t m = full;
        if (!stw) {  // get_vis_mask row b: forward sweep, backward sweep, spill into row b-1
            m = vis;
            m |= ((see + (m & see)) ^ see) & full;            // i = 0..V-2 ascending
            const uint32_t af = m & see & (full >> 1);
            uint32_t nxt = af | (af << 1);
            uint32_t mr = bitrev(m, V);
            const uint32_t sr = bitrev(see, V);
            mr |= ((sr + (mr & sr)) ^ sr) & full;             // i = V-1..1 descending
            m = bitrev(mr, V);
            const uint32_t ab = m & see & (full & ~1u);
            nxt |= ab | (ab >> 1);
            vis = nxt & full;
        }
        if (VT) {
#pragma unroll
            for (int a = 0; a < V; a++)
                if (!((m >> a) & 1u)) cr[VT ? a * VT + b : 0] = 0;  // UNSEEN, utils/obs.py:95-100
        } else {
            for (int a = 0; a < V; a++)
                if (!((m >> a) & 1u)) {
                    uint8_t *o = out + (a * V + b) * 3;
                    o[0] = 0; o[1] = 0; o[2] = 0;
                }
        }
    }
    if (VT) {  // 24-bit cells -> dense byte stream, written as 32-bit words
        constexpr int NC = VT ? VT * VT : 1;
        constexpr int NW = (3 * NC + 3) / 4;
        uint32_t *o32 = (uint32_t *)out;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const int i0 = (4 * w) / 3, sh = 4 * w - 3 * i0;  // first cell, byte offset inside it
            const uint32_t c0 = cr[i0 < NC ? i0 : 0];
            const uint32_t c1 = (i0 + 1 < NC) ? cr[i0 + 1 < NC ? i0 + 1 : 0] : 0u;
            const uint32_t sel = sh == 0 ? 0x4210u : (sh == 1 ? 0x5421u : 0x6542u);
            o32[w] = byte_perm(c0 & 0x00ffffffu, c1 & 0x00ffffffu, sel);
        }
        for (int w = NW; w * 4 < p.ostride; w++) o32[w] = 0;
    } else {
        for (int q = 3 * V * V; q < p.ostride; q++) out[q] = 0;
    }
}

template <int VT>
MG_HD void phase_obs(const Params &p, uint8_t *smem, int blk, int tid, int nt) {
    Block b = block_view(p, smem, blk);
    const int i = tid / p.tpe, lane = tid - i * p.tpe;
    if (i >= b.ne) return;
    const uint32_t *cells = b.cells + i * p.cstride;
    for (int k = lane; k < p.n; k += p.tpe) {
        const uint32_t a0 = b.ag[(i * p.n + k) * 2], a1 = b.ag[(i * p.n + k) * 2 + 1];
        obs_agent<VT>(p, cells, a0, a1, b.stage + (size_t)(i * p.n + k) * p.ostride);
    }
}

// ---- P4: store -------------------------------------------------------------------------------------
template <int MODE>
MG_HD void phase_store(const Params &p, uint8_t *smem, int blk, int tid, int nt) {
    Block b = block_view(p, smem, blk);
    if (MODE != MODE_STEP)
        coop_copy(p.obs + (size_t)b.e0 * p.n * p.ostride, b.stage, b.ne * p.n * p.ostride, tid, nt);
    if (MODE != MODE_OBS)
        coop_copy(p.agents + (size_t)b.e0 * p.n * 8, b.ag, b.ne * p.n * 8, tid, nt);
}

#ifdef __CUDACC__
template <int VT, int MODE>
__global__ void __launch_bounds__(256) step_obs_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int blk = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    phase_load<MODE>(p, smem, blk, tid, nt);
    __syncthreads();
    if (MODE != MODE_OBS && (p.flags & MG_FLAG_AUTO_RESET)) {
        phase_reset(p, smem, blk, tid, nt);
        __syncthreads();
    }
    phase_convert(p, smem, blk, tid, nt);
    __syncthreads();
    phase_step<MODE>(p, smem, blk, tid, nt);
    __syncthreads();
    if (MODE != MODE_STEP) {
        phase_obs<VT>(p, smem, blk, tid, nt);
        __syncthreads();
    }
    phase_store<MODE>(p, smem, blk, tid, nt);
}
#endif

}  // namespace mg
