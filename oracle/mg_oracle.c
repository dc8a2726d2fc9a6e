/*
 * CPU ORACLE (plain C) for the MultiGrid step/observe hot path -- TEST INFRASTRUCTURE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu-baseline / `--impl reference` legs may
 * load the library built from this file. The product (multigrid_b200/) never links or calls it.
 *
 * It is a scalar port of the reference's algorithm (ini/multigrid, /root/reference), looping the
 * way the reference loops (copy the grid, stamp agents, slice+rotate cell by cell, serial
 * visibility sweeps), batched over independent envs with OpenMP. Each function cites the
 * reference lines it follows. Parity status: PINNED -- tests/test_oracle_golden.py checks it
 * against fixtures recorded from the unmodified reference (tests/golden/make_golden.py) and
 * against oracle/mg_oracle.py.
 *
 * Packed state per env (same as the engine's HBM layout):
 *   grid   int8 [W][H][3]  x-major (core/grid.py:54)      agents int8 [n][8] =
 *   {dir,x,y,terminated,carry_type,carry_color,carry_state,color} (core/agent.py:222-232)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { T_UNSEEN, T_EMPTY, T_WALL, T_FLOOR, T_DOOR, T_KEY, T_BALL, T_BOX, T_GOAL, T_LAVA, T_AGENT };
enum { S_OPEN, S_CLOSED, S_LOCKED };
enum { ACT_LEFT, ACT_RIGHT, ACT_FORWARD, ACT_PICKUP, ACT_DROP, ACT_TOGGLE, ACT_DONE };
enum { A_DIR, A_X, A_Y, A_TERM, A_CT, A_CC, A_CS, A_COLOR, A_DIM };
enum { HOOK_NONE, HOOK_BUP, HOOK_RBD, HOOK_LH };

#define MGO_MAX_AGENTS 64
#define MGO_MAX_VIEW 31

typedef struct {
    int32_t W, H, n, V, max_steps;
    int32_t see_through_walls, allow_overlap, joint_reward, success_any, failure_any;
    int32_t hook, auto_reset, layout_stride, num_layouts;
    int32_t obs_agent_stride; /* bytes between agents in obs (>= 3*V*V) */
    int32_t hook_param;       /* LockedHallway: number of rooms (= doors) */
} mgo_config;

static const int DIR_DX[4] = {1, 0, -1, 0}; /* core/constants.py:21-30 */
static const int DIR_DY[4] = {0, 1, 0, -1};

/* numpy Generator(PCG64).random(): 128-bit LCG, XSL-RR output, top 53 bits (base.py:399) */
static inline uint64_t pcg64_next53(unsigned __int128 *state, unsigned __int128 inc) {
    const unsigned __int128 mult =
        ((unsigned __int128)0x2360ED051FC65DA4ULL << 64) | 0x4385DF649FCCF645ULL;
    *state = *state * mult + inc;
    uint64_t hi = (uint64_t)(*state >> 64), lo = (uint64_t)*state;
    uint64_t x = hi ^ lo;
    unsigned rot = (unsigned)(hi >> 58);
    uint64_t out = (x >> rot) | (x << ((64 - rot) & 63));
    return out >> 11;
}

/* base.py:598-602 -- float64, evaluated in the reference's order (no fused multiply-add) */
static double reward_value(int32_t step_count, int32_t max_steps) {
    volatile double ratio = (double)step_count / (double)max_steps;
    volatile double scaled = 0.9 * ratio;
    return 1.0 - scaled;
}

static void on_success(const mgo_config *c, int8_t *agents, int k, double *rew, uint8_t *term,
                       int32_t step_count) { /* base.py:478-507 */
    if (c->success_any) {
        for (int j = 0; j < c->n; j++) { agents[j * A_DIM + A_TERM] = 1; term[j] = 1; }
    } else {
        agents[k * A_DIM + A_TERM] = 1; term[k] = 1;
    }
    double r = reward_value(step_count, c->max_steps);
    if (c->joint_reward) for (int j = 0; j < c->n; j++) rew[j] = r;
    else rew[k] = r;
}

static void on_failure(const mgo_config *c, int8_t *agents, int k, uint8_t *term) { /* base.py:509-532 */
    if (c->failure_any) {
        for (int j = 0; j < c->n; j++) { agents[j * A_DIM + A_TERM] = 1; term[j] = 1; }
    } else {
        agents[k * A_DIM + A_TERM] = 1; term[k] = 1;
    }
}

static int agent_present(const mgo_config *c, const int8_t *agents, int x, int y) {
    for (int j = 0; j < c->n; j++)
        if (agents[j * A_DIM + A_X] == x && agents[j * A_DIM + A_Y] == y) return 1;
    return 0;
}

/* MultiGridEnv.handle_actions (base.py:378-476) */
/* cell_flags[x*H+y] & 1: the Door OBJECT there is closed although grid.state shows it open. The
 * reference keeps both in sync with grid.update(), except RedBlueDoorsEnv.step, which closes the
 * blue door's object without updating the array (envs/redbluedoors.py:185): the rules read the
 * object, the observation reads the array. */
static int handle_actions(const mgo_config *c, int8_t *grid, int8_t *agents, int32_t step_count,
                          uint64_t *pcg_state, const uint64_t *pcg_inc, const int8_t *actions,
                          double *rew, uint8_t *cell_flags) {
    int n = c->n, H = c->H, bad = 0;
    int order[MGO_MAX_AGENTS];
    uint8_t scratch[MGO_MAX_AGENTS];
    if (n == 1) {
        order[0] = 0;
    } else { /* np_random.random(size=n).argsort(): ascending, insertion sort (stable) */
        uint64_t key[MGO_MAX_AGENTS];
        unsigned __int128 s = ((unsigned __int128)pcg_state[1] << 64) | pcg_state[0];
        unsigned __int128 inc = ((unsigned __int128)pcg_inc[1] << 64) | pcg_inc[0];
        for (int j = 0; j < n; j++) key[j] = pcg64_next53(&s, inc);
        pcg_state[0] = (uint64_t)s; pcg_state[1] = (uint64_t)(s >> 64);
        for (int j = 0; j < n; j++) {
            int p = j;
            while (p > 0 && key[order[p - 1]] > key[j]) { order[p] = order[p - 1]; p--; }
            order[p] = j;
        }
    }
    for (int oi = 0; oi < n; oi++) {
        int k = order[oi];
        int a = actions[k];
        int8_t *ag = agents + k * A_DIM;
        if (a < 0) continue;           /* id absent from the dict (base.py:403-404) */
        if (ag[A_TERM]) continue;      /* base.py:408-409 */
        if (a == ACT_LEFT) { ag[A_DIR] = (int8_t)((ag[A_DIR] + 3) & 3); continue; }
        if (a == ACT_RIGHT) { ag[A_DIR] = (int8_t)((ag[A_DIR] + 1) & 3); continue; }
        if (a == ACT_DONE) continue;
        if (a > ACT_DONE) { bad = 1; continue; } /* reference raises ValueError (base.py:473) */
        int fx = ag[A_X] + DIR_DX[ag[A_DIR] & 3], fy = ag[A_Y] + DIR_DY[ag[A_DIR] & 3];
        if (fx < 0 || fx >= c->W || fy < 0 || fy >= H) continue;
        int8_t *cell = grid + (fx * H + fy) * 3;
        int t = cell[0], col = cell[1], st = cell[2];
        if (t == T_DOOR && (cell_flags[fx * H + fy] & 1)) st = S_CLOSED; /* what the object says */
        if (a == ACT_FORWARD) { /* base.py:420-436 */
            int can_overlap = t == T_EMPTY || t == T_FLOOR || t == T_GOAL || t == T_LAVA ||
                              (t == T_DOOR && st == S_OPEN);
            if (!can_overlap) continue;
            if (!c->allow_overlap && agent_present(c, agents, fx, fy)) continue;
            ag[A_X] = (int8_t)fx; ag[A_Y] = (int8_t)fy;
            if (t == T_GOAL) on_success(c, agents, k, rew, scratch, step_count);
            if (t == T_LAVA) on_failure(c, agents, k, scratch);
        } else if (a == ACT_PICKUP) { /* base.py:439-446 */
            if ((t == T_KEY || t == T_BALL || t == T_BOX) && ag[A_CT] == T_EMPTY) {
                ag[A_CT] = (int8_t)t; ag[A_CC] = (int8_t)col; ag[A_CS] = (int8_t)st;
                cell[0] = T_EMPTY; cell[1] = 0; cell[2] = 0;
            }
        } else if (a == ACT_DROP) { /* base.py:449-459 */
            if (ag[A_CT] != T_EMPTY && t == T_EMPTY && !agent_present(c, agents, fx, fy)) {
                cell[0] = ag[A_CT]; cell[1] = ag[A_CC]; cell[2] = ag[A_CS];
                ag[A_CT] = T_EMPTY; ag[A_CC] = 0; ag[A_CS] = 0;
            }
        } else { /* toggle, base.py:462-467 */
            if (t == T_DOOR) { /* core/world_object.py:458-474 */
                if (st == S_LOCKED) {
                    if (ag[A_CT] == T_KEY && ag[A_CC] == col) cell[2] = S_OPEN;
                } else {
                    cell[2] = (st == S_OPEN) ? S_CLOSED : S_OPEN;
                }
                cell_flags[fx * H + fy] &= 0xFE; /* Door.toggle ends with grid.update() */
            } else if (t == T_BOX) { /* core/world_object.py:599-605, contains == None */
                cell[0] = T_EMPTY; cell[1] = 0; cell[2] = 0;
            }
        }
    }
    return bad;
}

/* gen_obs_grid_encoding (utils/obs.py:66-102) for one env */
static void gen_obs_env(const mgo_config *c, const int8_t *grid, const int8_t *agents,
                        int8_t *obs, int8_t *scratch_grid) {
    int n = c->n, V = c->V, W = c->W, H = c->H, half = c->V / 2;
    const int8_t *g = grid;
    if (n > 1) { /* utils/obs.py:163-171 */
        memcpy(scratch_grid, grid, (size_t)W * H * 3);
        for (int j = 0; j < n; j++) {
            const int8_t *ag = agents + j * A_DIM;
            if (!ag[A_TERM]) {
                int8_t *cell = scratch_grid + (ag[A_X] * H + ag[A_Y]) * 3;
                cell[0] = T_AGENT; cell[1] = ag[A_COLOR]; cell[2] = ag[A_DIR];
            }
        }
        g = scratch_grid;
    }
    for (int k = 0; k < n; k++) {
        const int8_t *ag = agents + k * A_DIM;
        int8_t *o = obs + (size_t)k * c->obs_agent_stride;
        int d = ag[A_DIR], px = ag[A_X], py = ag[A_Y], tx = 0, ty = 0;
        /* get_view_exts (utils/obs.py:276-316) */
        if (d == 0) { tx = px; ty = py - half; }
        else if (d == 1) { tx = px - half; ty = py; }
        else if (d == 2) { tx = px - V + 1; ty = py - half; }
        else if (d == 3) { tx = px - half; ty = py - V + 1; }
        int rot = (d + 1) & 3;
        for (int i = 0; i < V; i++)
            for (int j = 0; j < V; j++) { /* utils/obs.py:181-202 */
                int x = tx + i, y = ty + j, ir, jr;
                if (rot == 0) { ir = i; jr = j; }
                else if (rot == 1) { ir = j; jr = V - i - 1; }
                else if (rot == 2) { ir = V - i - 1; jr = V - j - 1; }
                else { ir = V - j - 1; jr = i; }
                int8_t *dst = o + (ir * V + jr) * 3;
                if (x >= 0 && x < W && y >= 0 && y < H) {
                    const int8_t *src = g + (x * H + y) * 3;
                    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
                } else {
                    dst[0] = T_WALL; dst[1] = 5; dst[2] = 0;
                }
            }
        int8_t *self = o + (half * V + V - 1) * 3; /* utils/obs.py:207 */
        self[0] = ag[A_CT]; self[1] = ag[A_CC]; self[2] = ag[A_CS];
        if (c->see_through_walls) continue;

        uint8_t vis[MGO_MAX_VIEW][MGO_MAX_VIEW], see[MGO_MAX_VIEW][MGO_MAX_VIEW];
        for (int i = 0; i < V; i++)
            for (int j = 0; j < V; j++) { /* see_behind, utils/obs.py:47-63 */
                const int8_t *cell = o + (i * V + j) * 3;
                see[i][j] = !(cell[0] == T_WALL || (cell[0] == T_DOOR && cell[2] != S_OPEN));
                vis[i][j] = 0;
            }
        vis[half][V - 1] = 1;
        for (int j = V - 1; j >= 0; j--) { /* get_vis_mask, utils/obs.py:236-273 */
            for (int i = 0; i < V - 1; i++)
                if (vis[i][j] && see[i][j]) {
                    vis[i + 1][j] = 1;
                    if (j > 0) { vis[i + 1][j - 1] = 1; vis[i][j - 1] = 1; }
                }
            for (int i = V - 1; i > 0; i--)
                if (vis[i][j] && see[i][j]) {
                    vis[i - 1][j] = 1;
                    if (j > 0) { vis[i - 1][j - 1] = 1; vis[i][j - 1] = 1; }
                }
        }
        for (int i = 0; i < V; i++)
            for (int j = 0; j < V; j++)
                if (!vis[i][j]) { /* utils/obs.py:95-100 */
                    int8_t *cell = o + (i * V + j) * 3;
                    cell[0] = 0; cell[1] = 0; cell[2] = 0;
                }
    }
}

int mgo_gen_obs(const mgo_config *c, int64_t num_envs, const int8_t *grid, const int8_t *agents,
                int8_t *obs, int nthreads) {
    if (c->n > MGO_MAX_AGENTS || c->V > MGO_MAX_VIEW) return -1;
    size_t gsz = (size_t)c->W * c->H * 3, asz = (size_t)c->n * A_DIM;
    size_t osz = (size_t)c->n * c->obs_agent_stride;
    (void)nthreads;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        int8_t *scratch = (int8_t *)malloc(gsz);
#pragma omp for schedule(static)
        for (int64_t e = 0; e < num_envs; e++)
            gen_obs_env(c, grid + e * gsz, agents + e * asz, obs + e * osz, scratch);
        free(scratch);
    }
    return 0;
}

/* RedBlueDoorsEnv.step post-hook (envs/redbluedoors.py:170-187): every agent whose action was toggle,
 * in agent order, terminated or not: front cell is the open blue door -> success if the red door is
 * open, else failure and the blue door is closed again. Colors: red 0, blue 2 (constants.py:51-60). */
static void hook_red_blue_doors(const mgo_config *c, int8_t *grid, int8_t *agents, const int8_t *actions,
                                double *rew, uint8_t *term, int32_t step_count, uint8_t *cell_flags) {
    for (int k = 0; k < c->n; k++) {
        if (actions[k] != ACT_TOGGLE) continue;
        const int8_t *ag = agents + k * A_DIM;
        int fx = ag[A_X] + DIR_DX[ag[A_DIR] & 3], fy = ag[A_Y] + DIR_DY[ag[A_DIR] & 3];
        if (fx < 0 || fx >= c->W || fy < 0 || fy >= c->H) continue;
        int8_t *cell = grid + ((size_t)fx * c->H + fy) * 3;
        if (cell[0] != T_DOOR || cell[1] != 2 || cell[2] != S_OPEN || (cell_flags[fx * c->H + fy] & 1)) continue;
        int red_open = 0, found = 0;
        for (int x = 0; x < c->W && !found; x++)
            for (int y = 0; y < c->H; y++) {
                const int8_t *r = grid + ((size_t)x * c->H + y) * 3;
                if (r[0] == T_DOOR && r[1] == 0) { red_open = r[2] == S_OPEN; found = 1; break; }
            }
        if (red_open) {
            on_success(c, agents, k, rew, term, step_count);
        } else {
            on_failure(c, agents, k, term);
            cell_flags[fx * c->H + fy] |= 1; /* blue_door.is_open = False, WITHOUT grid.update() */
        }
    }
}

/* LockedHallwayEnv.step post-hook (envs/locked_hallway.py:203-227). *hook_state = bit per door COLOUR
 * already unlocked (distinct colours for num_rooms <= 6, :156-158) = the reference's unlocked_doors. */
static void hook_locked_hallway(const mgo_config *c, const int8_t *grid, const int8_t *agents,
                                const int8_t *actions, double *rew, uint8_t *term, int32_t step_count,
                                int32_t *hook_state) {
    for (int k = 0; k < c->n; k++) {
        if (actions[k] != ACT_TOGGLE) continue;
        const int8_t *ag = agents + k * A_DIM;
        int fx = ag[A_X] + DIR_DX[ag[A_DIR] & 3], fy = ag[A_Y] + DIR_DY[ag[A_DIR] & 3];
        if (fx < 0 || fx >= c->W || fy < 0 || fy >= c->H) continue;
        const int8_t *cell = grid + ((size_t)fx * c->H + fy) * 3;
        if (cell[0] != T_DOOR || cell[2] == S_LOCKED) continue;
        if ((*hook_state >> cell[1]) & 1) continue;
        *hook_state |= 1 << cell[1];
        double r = reward_value(step_count, c->max_steps);
        if (c->joint_reward) for (int j = 0; j < c->n; j++) rew[j] = rew[j] + r; /* rewards[k] += _reward() */
        else rew[k] = rew[k] + r;
    }
    if (__builtin_popcount((unsigned)*hook_state) == c->hook_param) /* returned dict only, not agent state */
        for (int j = 0; j < c->n; j++) term[j] = 1;
}

/* MultiGridEnv.step (base.py:303-346) + env post-hook, with the engine's "next-step" auto-reset */
int mgo_step_obs(const mgo_config *c, int64_t num_envs, int8_t *grid, int8_t *agents,
                 int32_t *step_count, uint64_t *pcg_state, const uint64_t *pcg_inc,
                 int32_t *layout_idx, const int8_t *pool_grid, const int8_t *pool_agents,
                 const int8_t *actions, int8_t *obs, double *reward, uint8_t *terminated,
                 uint8_t *truncated, uint8_t *cell_flags /* [E][W*H] */, int32_t *hook_state /* [E] */,
                 int nthreads) {
    if (c->n > MGO_MAX_AGENTS || c->V > MGO_MAX_VIEW) return -1;
    size_t gsz = (size_t)c->W * c->H * 3, asz = (size_t)c->n * A_DIM;
    size_t osz = (size_t)c->n * c->obs_agent_stride;
    int n = c->n, bad_any = 0;
    (void)nthreads;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1) reduction(| : bad_any)
    {
        int8_t *scratch = (int8_t *)malloc(gsz);
#pragma omp for schedule(static)
        for (int64_t e = 0; e < num_envs; e++) {
            int8_t *g = grid + e * gsz, *ag = agents + e * asz;
            double *rew = reward + e * n;
            uint8_t *term = terminated + e * n;
            for (int j = 0; j < n; j++) rew[j] = 0.0; /* base.py:394 */
            if (c->auto_reset) {
                int all_term = 1;
                for (int j = 0; j < n; j++) all_term &= (ag[j * A_DIM + A_TERM] != 0);
                if (c->hook == HOOK_LH && __builtin_popcount((unsigned)hook_state[e]) == c->hook_param)
                    all_term = 1; /* LockedHallway terminates in the returned dict only */
                if (all_term || step_count[e] >= c->max_steps) { /* is_done, base.py:534-539 */
                    int32_t k = (int32_t)(((int64_t)layout_idx[e] + c->layout_stride) % c->num_layouts);
                    layout_idx[e] = k;
                    memcpy(g, pool_grid + (size_t)k * gsz, gsz);
                    memcpy(ag, pool_agents + (size_t)k * asz, asz);
                    memset(cell_flags + e * (size_t)c->W * c->H, 0, (size_t)c->W * c->H);
                    hook_state[e] = 0;
                    step_count[e] = 0;
                    gen_obs_env(c, g, ag, obs + e * osz, scratch);
                    for (int j = 0; j < n; j++) term[j] = 0;
                    truncated[e] = 0;
                    continue;
                }
            }
            step_count[e] += 1; /* base.py:333 */
            uint8_t *cf = cell_flags + e * (size_t)c->W * c->H;
            bad_any |= handle_actions(c, g, ag, step_count[e], pcg_state + 2 * e, pcg_inc + 2 * e,
                                      actions + e * n, rew, cf);
            gen_obs_env(c, g, ag, obs + e * osz, scratch);          /* base.py:337 */
            for (int j = 0; j < n; j++) term[j] = ag[j * A_DIM + A_TERM] != 0; /* base.py:338 */
            truncated[e] = step_count[e] >= c->max_steps;              /* base.py:339 */
            if (c->hook == HOOK_BUP) /* envs/blockedunlockpickup.py:166-175 */
                for (int j = 0; j < n; j++)
                    if (ag[j * A_DIM + A_CT] == T_BOX) on_success(c, ag, j, rew, term, step_count[e]);
            if (c->hook == HOOK_RBD) /* envs/redbluedoors.py:170-187 */
                hook_red_blue_doors(c, g, ag, actions + e * n, rew, term, step_count[e], cf);
            if (c->hook == HOOK_LH)
                hook_locked_hallway(c, g, ag, actions + e * n, rew, term, step_count[e], hook_state + e);
        }
        free(scratch);
    }
    return bad_any ? 1 : 0;
}

int mgo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
