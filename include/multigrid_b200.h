/*
 * multigrid_b200 -- C ABI of the B200 (sm_100a) batched MultiGrid step/observe engine.
 *
 * This is the drop-in boundary for the ONE hot path of ini/multigrid that this library replaces:
 *
 *     MultiGridEnv.step()            multigrid/base.py:303-346
 *       -> handle_actions()          multigrid/base.py:378-476  (+ on_success/on_failure :478-532)
 *       -> gen_obs()                 multigrid/base.py:348-376
 *            -> gen_obs_grid_encoding  multigrid/utils/obs.py:66-102 (numba)
 *
 * The reference has no FFI of its own (it is Python + 8 numba functions); the seam a maintainer
 * would bind is the array-level call `gen_obs_grid_encoding(grid_state, agent_state, view_size,
 * see_through_walls)` and the `step()` method around it. INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - Plain C: pointers + sizes only. All `d_`/unprefixed data pointers are DEVICE pointers
 *     (16-byte aligned, e.g. torch CUDA tensors); `h_` pointers are (pinned) HOST pointers.
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default
 *     stream); no internal synchronisation, no global state, re-entrant across streams/devices.
 *   - Return value: 0 = ok, >0 = cudaError_t of the launch, <0 = MG_ERR_* argument error.
 *   - Batch axis first: env e owns slice e of every array ("env-major", each env's record is
 *     contiguous, so a thread block's group of envs is one contiguous HBM span per array).
 *
 * State layout per env (all int8 values are < 128):
 *   grid        uint32 [W+1][H+1] CELL WORDS, x-major like Grid.state (core/grid.py:54):
 *                                 type | color<<8 | state<<16 | opaque<<31, where opaque = wall or
 *                                 non-open door (see_behind, utils/obs.py:47-63). Row x = W and
 *                                 column y = H are WALL sentinels (out-of-bounds view cells read
 *                                 them). Bytes 0..2 of word (x,y) ARE Grid.state[x,y,:]; build it
 *                                 from / turn it into the reference's (W,H,3) bytes with
 *                                 mg_pack_grid / mg_unpack_grid.
 *   agents      int8  [n][8]      {dir,x,y,terminated,carry_type,carry_color,carry_state,color}
 *                                 = AgentState (core/agent.py:222-232) without the constant TYPE
 *   step_count  int32
 *   pcg_state   uint64[2]         {lo,hi} of numpy PCG64's 128-bit state (env.np_random)
 *   pcg_inc     uint64[2]         {lo,hi} of its increment (constant)
 *   layout_idx  int32             cursor into the reset-layout pool (auto-reset only)
 *   hook_state  int32             post-hook state (LockedHallway: unlocked-door colour bits)
 *   chain       uint32[4]         {next ticket, tickets done, grid dirty, -}: chained launches, single-layout dedup
 * Outputs per env:
 *   obs         int8  [n][obs_agent_stride]  first 3*V*V bytes of each agent slot = image[V][V][3]
 *   reward      float64 [n]       bit-exact `1 - 0.9*(step_count/max_steps)` (base.py:598-602)
 *   terminated  uint8 [n]
 *   truncated   uint8
 */
#ifndef MULTIGRID_B200_H
#define MULTIGRID_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MG_ABI_VERSION 11

/* MgConfig.flags */
#define MG_FLAG_SEE_THROUGH_WALLS 0x01u /* agents[0].see_through_walls, base.py:364-365 */
#define MG_FLAG_ALLOW_OVERLAP     0x02u /* allow_agent_overlap, base.py:95            */
#define MG_FLAG_JOINT_REWARD      0x04u /* joint_reward, base.py:96                   */
#define MG_FLAG_SUCCESS_ANY       0x08u /* success_termination_mode == 'any'          */
#define MG_FLAG_FAILURE_ANY       0x10u /* failure_termination_mode == 'any'          */
#define MG_FLAG_AUTO_RESET        0x20u /* engine extension: "next-step" auto reset   */
#define MG_FLAG_STREAM_STATE      0x40u /* cache policy only (results unchanged): grid/action loads and
                                           observation stores get L2 evict_first priority. For batches that
                                           are NOT re-stepped while still L2-resident (state larger than L2,
                                           or several engines interleaved); leave it off when one batch that
                                           fits L2 is stepped back to back. */

#define MG_FLAG_CHAINED           0x80u /* scheduling only (results unchanged), needs MgState.chain:
                                           this step launch is ordered after the previous CHAINED step launch on the
                                           same state env by env (tickets) instead of waiting for that whole grid,
                                           so its blocks load and compute while the previous launch drains. The
                                           caller promises that since that previous launch nothing else was enqueued
                                           on the stream that writes this launch's inputs (the actions!) or reads or
                                           writes the state / the outputs. Completion stays in stream order. */
#define MG_FLAG_CHAIN_HEAD        0x100u /* with MG_FLAG_CHAINED: first launch of a chain (the previous operation on
                                           the state was not a chained step launch): takes tickets and also waits for
                                           the whole previous grid of the stream, like a plain launch. Launches
                                           without MG_FLAG_CHAINED never touch the tickets. */

#define MG_FLAG_STATIC_GRID       0x200u /* the caller's PROMISE (results unchanged when it holds), needs MgState.static_obs:
                                           no action can change any env's grid and every env's grid IS pool layout 0 --
                                           num_layouts == 1, hook == MG_HOOK_NONE, the layout holds only empty / wall /
                                           floor / goal / lava cells (no door, key, ball, box: every Empty-family env of
                                           the reference, envs/empty.py), no agent carries anything and every agent stands
                                           inside the grid on a cell that is not a wall. mg_step_obs then takes the static
                                           path: observations come from the memoised per-(x, y, dir) views of
                                           mg_build_static_obs plus the other agents drawn on top, the transition is
                                           left / right / forward against the one layout, and `grid` is neither read nor
                                           written. A violated promise that the kernel can see (an agent outside the grid
                                           or carrying something) ORs 4 into MgStepOut.status. Ignored (general path) by
                                           mg_step, mg_rollout and chained launches, and under MG_NO_STATIC=1. */

/* MgConfig.hook: env-specific step() post-hooks */
#define MG_HOOK_NONE 0
#define MG_HOOK_BLOCKED_UNLOCK_PICKUP 1 /* envs/blockedunlockpickup.py:166-175 */
#define MG_HOOK_RED_BLUE_DOORS 2        /* envs/redbluedoors.py:170-187 */
#define MG_HOOK_LOCKED_HALLWAY 3        /* envs/locked_hallway.py:203-227; hook_param = number of rooms */

#define MG_ERR_BAD_ARG   (-1)
#define MG_ERR_ALIGNMENT (-2)
#define MG_ERR_TOO_LARGE (-3)

#define MG_MAX_VIEW 15     /* odd view sizes 3..15 */
#define MG_MAX_AGENTS 32

typedef struct MgConfig {
    int32_t width, height;     /* grid size W,H (grid is [W][H][3]) */
    int32_t num_agents;        /* n */
    int32_t view_size;         /* V, odd, 3..MG_MAX_VIEW (core/agent.py:78-79) */
    int32_t max_steps;
    uint32_t flags;            /* MG_FLAG_* */
    int32_t hook;              /* MG_HOOK_* */
    int32_t obs_agent_stride;  /* bytes between agents in `obs`; multiple of 4, >= 3*V*V */
    int32_t num_layouts;       /* K: size of the reset-layout pool (auto-reset) */
    int32_t layout_stride;     /* on reset: layout_idx = (layout_idx + layout_stride) % K */
    int32_t hook_param;        /* hook-specific constant (MG_HOOK_LOCKED_HALLWAY: number of rooms) */
} MgConfig;

typedef struct MgState {
    uint32_t *grid;            /* [E][W+1][H+1] cell words */
    int8_t *agents;            /* [E][n][8]    */
    int32_t *step_count;       /* [E]          */
    uint64_t *pcg_state;       /* [E][2]       */
    const uint64_t *pcg_inc;   /* [E][2]       */
    int32_t *layout_idx;       /* [E]      (may be NULL without MG_FLAG_AUTO_RESET) */
    const uint32_t *pool_grid; /* [K][W+1][H+1] cell words (may be NULL without MG_FLAG_AUTO_RESET) */
    const int8_t *pool_agents; /* [K][n][8]    (may be NULL without MG_FLAG_AUTO_RESET) */
    int32_t *hook_state;       /* [E] per-env state of the post-hook; only MG_HOOK_LOCKED_HALLWAY uses it
                                  (bit per door colour already unlocked); may be NULL otherwise */
    const uint32_t *pool_rep;  /* [32][W+1][H+1] 32 copies of pool layout 0, 16-byte aligned (may be NULL = no dedup);
                                  used only when num_layouts == 1 and `chain` is given: a group whose envs are all
                                  clean loads its cells from this L2-resident buffer instead of reading
                                  4*(W+1)*(H+1) bytes per env from HBM */
    uint32_t *chain;           /* [E][4] per-env record {next ticket, tickets done, grid dirty, reserved}, 16-byte
                                  aligned, zero-initialised except dirty = 1 (may be NULL without MG_FLAG_CHAINED and
                                  without dedup). A chained launch takes ticket next[e]++ and publishes done[e] =
                                  ticket + 1 when it is done with env e. dirty = 1: the env's grid may differ from
                                  pool_grid[layout_idx]; the engine sets it on every cell write-through and clears it
                                  on reset; the CALLER sets it whenever it writes `grid` itself. */
    const int8_t *static_obs;  /* mg_static_obs_bytes(W, H, obs_agent_stride) bytes filled by mg_build_static_obs,
                                  16-byte aligned; only read under MG_FLAG_STATIC_GRID (may be NULL otherwise) */
} MgState;

typedef struct MgStepOut {
    int8_t *obs;               /* [E][n][obs_agent_stride] (NULL allowed for mg_step) */
    double *reward;            /* [E][n] */
    uint8_t *terminated;       /* [E][n] */
    uint8_t *truncated;        /* [E]    */
    int32_t *status;           /* [1] device word, OR-ed with 1 when an action outside 0..6
                                  (and != -1) was seen: the reference raises ValueError there
                                  (base.py:473-474). May be NULL. */
    uint8_t *one_hot;          /* [E][n][V][V][21], 16-byte aligned, or NULL. mg_step_obs only: the fused kernel ALSO
                                  writes OneHotObsWrapper.one_hot(image) of every observation (wrappers.py:158-190:
                                  11 type + 6 colour + 4 state channels, uint8) straight from its shared-memory
                                  stage -- what RLlib consumers of the reference read (rllib/__init__.py:110-111) --
                                  instead of a second pass (mg_one_hot) over `obs`. Plain launches only
                                  (MG_ERR_BAD_ARG together with MG_FLAG_CHAINED). */
} MgStepOut;

int mg_abi_version(void);
const char *mg_error_string(int code);

/* Smallest legal obs_agent_stride for a view size: 3*V*V rounded up to a multiple of 4. */
int32_t mg_obs_agent_stride(int32_t view_size);

/* Cell words per env: (W+1)*(H+1). */
int64_t mg_cells_per_env(int32_t width, int32_t height);

/*
 * Layout conversion between the reference's Grid.state bytes, int8 [E][W][H][3]
 * (core/grid.py:54, Grid.encode :310-327), and the engine's cell words (adds the opaque bit and the
 * wall sentinels / drops them). Replaces nothing on the hot path: it is how state enters and
 * leaves the engine (reset, state injection, inspection).
 */
int mg_pack_grid(int32_t width, int32_t height, int64_t num_envs, const int8_t *grid3, uint32_t *cells,
                 void *stream);
int mg_unpack_grid(int32_t width, int32_t height, int64_t num_envs, const uint32_t *cells, int8_t *grid3,
                   void *stream);

/*
 * On-device layout generation for EmptyEnv with random agent placement (the 'MultiGrid-Empty-Random-*'
 * ids and any EmptyEnv with agent_start_pos/agent_start_dir = None): one layout per generator.
 * Replaces: EmptyEnv._gen_grid (envs/empty.py:151-170) -> MultiGridEnv.place_agent / place_obj
 * (base.py:604-697) -> RandomMixin._rand_int (utils/random.py:23-38) -> numpy Generator(PCG64).integers
 * (bounded Lemire draw on the buffered 32-bit stream), bit-exactly: given the generator the reference's
 * RandomMixin holds, the layout and the generator's state afterwards are the reference's.
 *   rng_state  uint64 [K][2] {lo,hi} PCG64 state, advanced in place;  rng_inc uint64 [K][2] (constant)
 *   rng_buf    uint64 [K]: bit 32 = has_uint32, low word = uinteger of numpy's pcg64 state (may be NULL:
 *              starts empty, the leftover half is dropped)
 *   cells      uint32 [K][W+1][H+1] cell words (out);  agents int8 [K][n][8] (out)
 *   status     device word OR-ed with 2 if a placement gave up after 65 536 tries (may be NULL)
 */
int mg_gen_layouts_empty_random(int32_t width, int32_t height, int32_t num_agents, int64_t num_layouts,
                                uint64_t *rng_state, const uint64_t *rng_inc, uint64_t *rng_buf, uint32_t *cells,
                                int8_t *agents, int32_t *status, void *stream);

/*
 * On-device layouts of RedBlueDoorsEnv (envs/redbluedoors.py:142-168): grid (2*size) x size, agents placed in
 * the middle room, red / blue door heights from the layout generator. Arguments as
 * mg_gen_layouts_empty_random.
 */
int mg_gen_layouts_red_blue_doors(int32_t size, int32_t num_agents, int64_t num_layouts, uint64_t *rng_state,
                                  const uint64_t *rng_inc, uint64_t *rng_buf, uint32_t *cells, int8_t *agents,
                                  int32_t *status, void *stream);

/*
 * On-device layouts of LockedHallwayEnv (envs/locked_hallway.py:150-194; RandomMixin._rand_perm =
 * Generator.shuffle of a list = Fisher-Yates on numpy's random_interval): grid (3*(room_size-1)+1) x
 * ((num_rooms/2)*(room_size-1)+1); num_rooms even, 2..6 (door identity is its colour). Other arguments as
 * mg_gen_layouts_empty_random.
 */
int mg_gen_layouts_locked_hallway(int32_t num_rooms, int32_t room_size, int32_t max_hallway_keys,
                                  int32_t max_keys_per_room, int32_t num_agents, int64_t num_layouts,
                                  uint64_t *rng_state, const uint64_t *rng_inc, uint64_t *rng_buf, uint32_t *cells,
                                  int8_t *agents, int32_t *status, void *stream);

/*
 * On-device layouts of PlaygroundEnv (envs/playground.py:122-137 over core/roomgrid.py: connect_all,
 * add_object, place_agent in a random room) on a num_rows x num_cols RoomGrid (at most 16 rooms). Generator
 * arguments as mg_gen_layouts_bup (door positions come from the ORDER generator, roomgrid.py:324).
 */
int mg_gen_layouts_playground(int32_t room_size, int32_t num_rows, int32_t num_cols, int32_t num_agents,
                              int64_t num_layouts, uint64_t *rng_state, const uint64_t *rng_inc, uint64_t *rng_buf,
                              uint64_t *order_state, const uint64_t *order_inc, uint64_t *order_buf, uint32_t *cells,
                              int8_t *agents, int32_t *status, void *stream);

/*
 * On-device layouts of BlockedUnlockPickupEnv (envs/blockedunlockpickup.py:142-164 over
 * core/roomgrid.py:203-404: add_object, add_door, place_in_room on a 1 x 2 RoomGrid of `room_size`), same
 * generator conventions as mg_gen_layouts_empty_random plus the ORDER generator of each layout
 * (env.np_random: the door height is drawn from it, roomgrid.py:324): order_state [K][2] is advanced in
 * place, order_inc [K][2]; order_buf [K] (may be NULL = empty, dropped) is the stream's buffered 32-bit half in
 * the rng_buf format, in and out: integers() leaves the unused half of a 64-bit draw there and the env's NEXT
 * reset consumes it first. info [K] (may be NULL) receives the box colour index (the mission names it).
 * Grid is (2*(room_size-1)+1) x room_size. status |= 2 when a placement exceeded the reference's
 * max_tries = 1000 (the reference raises RecursionError).
 */
int mg_gen_layouts_bup(int32_t room_size, int32_t num_agents, int64_t num_layouts, uint64_t *rng_state,
                       const uint64_t *rng_inc, uint64_t *rng_buf, uint64_t *order_state, const uint64_t *order_inc,
                       uint64_t *order_buf, uint32_t *cells, int8_t *agents, int32_t *info, int32_t *status,
                       void *stream);

/*
 * Fully observable image: out int8 [E][W][H][3] = Grid.state with every agent (terminated or not)
 * written over its cell as (agent, colour, dir), highest agent index last.
 * Replaces: FullyObsWrapper.observation (multigrid/wrappers.py:50-58).
 */
int mg_full_obs(int32_t width, int32_t height, int32_t num_agents, int64_t num_envs, const uint32_t *cells,
                const int8_t *agents, int8_t *out, void *stream);

/*
 * One-hot encoding of a batch of observations: obs int8 [A][obs_agent_stride] (A = all agents of all
 * envs, images in the first 3*V*V bytes) -> out uint8 [A][V][V][21], channels = 11 type + 6 colour +
 * 4 state/direction. Replaces: OneHotObsWrapper.one_hot (multigrid/wrappers.py:158-190), which the
 * reference's RLlib registration always applies (multigrid/rllib/__init__.py:110-111).
 */
#define MG_ONE_HOT_CHANNELS 21
int mg_one_hot(int32_t view_size, int64_t num_agents_total, int32_t obs_agent_stride, const int8_t *obs,
               uint8_t *out, void *stream);
/* The same encoding for images of any cell count (e.g. FullyObsWrapper's W*H-cell grids under OneHotObsWrapper):
 * images int8 [num_images][image_stride] with 3 bytes per cell -> out uint8 [num_images][cells_per_image][21]. */
int mg_one_hot_cells(int64_t cells_per_image, int64_t num_images, int32_t image_stride, const int8_t *images,
                     uint8_t *out, void *stream);

/*
 * Network input of the reference's training script in one pass from the observations: out float32
 * [A][V][V][23] = the 21 one-hot channels (0.0 / 1.0) + cos, sin of 2*pi*direction/4 broadcast over the view.
 * Replaces: OneHotObsWrapper.one_hot (multigrid/wrappers.py:158-190) followed by preprocess_batch
 * (scripts/train.py:56-63: concatenate the direction features, .float()).
 *   direction     int8, one per agent, `direction_stride` bytes apart (the engine's agent records: stride 8)
 *   dir_lut       float32 [4][2] = {cos, sin} of 2*pi*d/4, computed by the caller the way the reference does
 *                 (torch float32), so the two feature channels are the reference's values bit for bit
 *   out           16-byte aligned
 */
#define MG_FEATURE_CHANNELS 23
int mg_obs_features(int32_t view_size, int64_t num_agents_total, int32_t obs_agent_stride, const int8_t *obs,
                    const int8_t *direction, int32_t direction_stride, const float *dir_lut, float *out, void *stream);

/*
 * Memoised observations and moves of a static layout (MG_FLAG_STATIC_GRID). `static_obs` holds, for EVERY cell
 * (x, y) of the layout and direction dir, at entry ((x*H + y)*4 + dir):
 *   - the packed view image[V][V][3] (obs_agent_stride bytes, zero padded to the entry stride) of a lone agent
 *     there that carries nothing, computed by the engine's own observation code. Replaces: gen_obs_grid_encoding
 *     (utils/obs.py:66-102) evaluated once per (position, direction) instead of once per agent and step;
 *   - behind the W*H*4 views, one uint32 per entry: the agent's position word after `forward` (base.py:420-436:
 *     the cell in front if it can be walked on, else unchanged) with flags for goal / lava.
 * layout_cells: uint32 [W+1][H+1] cell words (e.g. pool_grid); uses cfg's width, height, view_size,
 * obs_agent_stride and MG_FLAG_SEE_THROUGH_WALLS. The unrolled kernels (V = 7 with 4 or 2 agents, V = 9 with 8)
 * need obs_agent_stride = 3*V*V rounded up to 16 (160 / 256); any other stride takes the rolled static kernel.
 */
int32_t mg_static_obs_stride(int32_t obs_agent_stride); /* bytes per entry: obs_agent_stride rounded up to 16 */
int64_t mg_static_obs_bytes(int32_t width, int32_t height, int32_t obs_agent_stride); /* size of `static_obs` */
int mg_build_static_obs(const MgConfig *cfg, const uint32_t *layout_cells, int8_t *static_obs, void *stream);

/* Diagnostics: when set to a device buffer of 8 uint64 per warp (= per group of envs), every
 * following launch records %globaltimer at its phase boundaries (slots 0..4) and the SM id (slot 7).
 * NULL (the default) disables it. Used by tools/trace_timeline.py; not part of the hot path. */
void mg_debug_set_trace(void *device_buffer);

/* Number of engine kernels launched by this process so far (for launch accounting). */
int64_t mg_launch_count(void);

/*
 * gen_obs for every agent of every env. Pure function of (grid, agents).
 * Replaces: gen_obs_grid_encoding (utils/obs.py:66-102) as called by MultiGridEnv.gen_obs
 * (base.py:348-376); used for reset() observations.
 */
int mg_gen_obs(const MgConfig *cfg, int64_t num_envs, const uint32_t *grid, const int8_t *agents,
               int8_t *obs, void *stream);

/*
 * One lockstep transition for every env WITHOUT observations.
 * Replaces: step_count bookkeeping + handle_actions + terminations/truncations of
 * MultiGridEnv.step (base.py:333-340, 378-532) and the env post-hook selected by cfg->hook.
 * actions: int8 [E][n], values 0..6 (core/actions.py:5-15) or -1 = agent id absent from the dict.
 */
int mg_step(const MgConfig *cfg, int64_t num_envs, const MgState *state, const int8_t *actions,
            const MgStepOut *out, void *stream);

/* The fused hot path: mg_step followed by mg_gen_obs in ONE kernel (state read once).
 * Replaces: MultiGridEnv.step (base.py:303-346). This is what BASELINE.json's metric measures. */
int mg_step_obs(const MgConfig *cfg, int64_t num_envs, const MgState *state,
                const int8_t *actions, const MgStepOut *out, void *stream);

/*
 * T consecutive fused steps in ONE launch (engine extension; the reference has no counterpart: its
 * rollouts are a Python loop over env.step, e.g. scripts/visualize.py). Bit-identical to T calls of
 * mg_step_obs with actions[t] whose outputs go to slice t of the output arrays, but the agents and
 * the per-env scalars stay on chip between steps, warps run ahead of each other (no launch
 * boundary per step) and the cells are re-read from L2 instead of HBM. For open-loop action
 * tapes: random-policy rollouts, scripted policies, replays.
 *   actions     int8 [T][E][n]
 *   out->obs    int8 [T][E][n][obs_agent_stride];  direction int8 [T][E][n] (may be NULL)
 *   reward      float64 [T][E][n];  terminated uint8 [T][E][n];  truncated uint8 [T][E]
 */
typedef struct MgRolloutOut {
    int8_t *obs;
    int8_t *direction;
    double *reward;
    uint8_t *terminated;
    uint8_t *truncated;
    int32_t *status;
} MgRolloutOut;

int mg_rollout(const MgConfig *cfg, int64_t num_envs, int32_t num_steps, const MgState *state,
               const int8_t *actions, const MgRolloutOut *out, void *stream);

/*
 * Host-driven reset of the envs whose mask byte is non-zero (mask: uint8 [E], device): each takes the NEXT
 * layout of the pool (layout_idx = (layout_idx + layout_stride) % num_layouts), step_count = 0, hook state = 0;
 * the env's PCG64 stream continues. Identical to what the step kernels do to an env under MG_FLAG_AUTO_RESET,
 * for callers that decide about resets themselves (RLlib resets sub-envs from outside).
 * Replaces: MultiGridEnv.reset (base.py:250-301) for selected envs of a batch.
 */
int mg_reset_where(const MgConfig *cfg, int64_t num_envs, const MgState *state, const uint8_t *mask, void *stream);

/*
 * Fresh layouts on auto-reset. The reference draws a NEW layout at every reset() from the env's own generator
 * (base.py:250-301 -> the env class's _gen_grid). With ONE POOL SLOT PER ENV -- num_layouts == num_envs,
 * layout_idx[e] == e, layout_stride == 0 -- this call gives the same behaviour to MG_FLAG_AUTO_RESET: enqueue it
 * after every step launch; every env that is done (all agents terminated, step limit reached, or -- LockedHallway --
 * all doors unlocked) and will therefore be reset by the NEXT step launch gets its slot (pool_grid[e], pool_agents[e])
 * regenerated from generator e, which advances. Door positions (RoomGrid families) are drawn from a copy of the env's
 * order stream (pcg_state / pcg_inc; roomgrid.py:324 draws them from env.np_random); that stream itself is not
 * advanced by a reset in this engine. One thread per env; envs that are not done cost one predicate evaluation.
 *   gen->family        MG_LAYOUT_*; gen->params: EMPTY_RANDOM {}, BUP {room_size}, RED_BLUE_DOORS {size},
 *                      LOCKED_HALLWAY {num_rooms, room_size, max_hallway_keys, max_keys_per_room},
 *                      PLAYGROUND {room_size, num_rows, num_cols}
 *   gen->rng_state/inc/buf   uint64 [E][2] / [E][2] / [E]: the envs' layout generators (as mg_gen_layouts_*)
 *   gen->order_buf     uint64 [E] (may be NULL = empty): buffered 32-bit half of the order streams, as left by
 *                      mg_gen_layouts_bup / _playground; read, not advanced (like the stream itself)
 *   gen->info          int32 [E] (may be NULL): BlockedUnlockPickup box colour of the regenerated slots
 *   status             device word OR-ed with 2 when a placement gave up (may be NULL)
 */
#define MG_LAYOUT_EMPTY_RANDOM 1
#define MG_LAYOUT_BUP 2
#define MG_LAYOUT_RED_BLUE_DOORS 3
#define MG_LAYOUT_LOCKED_HALLWAY 4
#define MG_LAYOUT_PLAYGROUND 5
typedef struct MgLayoutGen {
    int32_t family;
    int32_t params[4];
    uint64_t *rng_state;
    const uint64_t *rng_inc;
    uint64_t *rng_buf;
    const uint64_t *order_buf;
    int32_t *info;
} MgLayoutGen;
int mg_refresh_done_layouts(const MgConfig *cfg, int64_t num_envs, const MgState *state, const MgLayoutGen *gen,
                            int32_t *status, void *stream);

/*
 * Host-buffer variant of mg_step_obs (what a CPU-side caller of env.step() sees):
 * copies h_actions -> d_actions, runs the fused kernel, copies the outputs in `d_out` to the
 * matching pointers in `h_out` (obs, reward, terminated, truncated; status is not copied).
 * Host buffers should be pinned; everything is enqueued on `stream`, the caller synchronises.
 */
int mg_step_obs_host(const MgConfig *cfg, int64_t num_envs, const MgState *state,
                     const int8_t *h_actions, int8_t *d_actions, const MgStepOut *d_out,
                     const MgStepOut *h_out, void *stream);

/*
 * Palette wire format (ABI v11): `bits` (1..8) per cell instead of 9 -- the index of the cell's 9-bit code
 * (type | colour << 4 | state << 7) in a palette chosen by the caller. The cell values a batch can show are few
 * (Empty-8x8 with 4 agents: unseen, empty, wall, goal and 16 agent encodings = 20 -> 5 bits: 32 instead of 56 bytes
 * per 7x7 view); lut is a DEVICE array uint8 [512], lut[code] = index (< 255) or 0xff when the code is not in the palette:
 * such a cell is written as index 0 and bit 3 (value 8) of `status` is set, so a palette that turns out too small
 * is noticed, never silently wrong. Record layout as mg_pack_obs with `bits` in place of 9;
 * mg_packed_obs_stride_bits(V, bits) = bytes per agent. mg_step_obs_host_palette = mg_step_obs_host_packed in this
 * format. Host decoder: multigrid_b200.engine.unpack_obs(packed, V, bits, palette).
 */
int32_t mg_packed_obs_stride_bits(int32_t view_size, int32_t bits);
int mg_pack_obs_palette(int32_t view_size, int64_t num_agents_total, int32_t obs_agent_stride, const int8_t *obs,
                        int32_t bits, const uint8_t *lut, uint8_t *packed, int32_t *status, void *stream);
int mg_step_obs_host_palette(const MgConfig *cfg, int64_t num_envs, const MgState *state, const int8_t *h_actions,
                             int8_t *d_actions, const MgStepOut *d_out, uint8_t *d_packed, int32_t bits,
                             const uint8_t *lut, const MgStepOut *h_out, void *stream);

/*
 * The host wire (ABI v11): everything a CPU-side caller of env.step() needs from one step in ONE compact buffer
 * (one device-to-host copy; for batches of 32 768 envs or more the envs are stepped and packed in 4 slices on the
 * caller's stream while a side stream copies the finished slices, so the kernels hide behind the PCIe transfer; the
 * caller's stream waits for the side stream before the call's work is complete; MG_WIRE_CHUNKS=1..8 overrides).
 * Layout of the buffer (mg_wire_bytes bytes; 16-byte aligned on the device):
 *   [0, mg_wire_obs_bytes)   observations in the palette format (mg_pack_obs_palette), [E][n][stride_bits]
 *   then E records of mg_wire_record_bytes(n) bytes:
 *       float64 value; uint32 terminated mask (bit j = agent j) | truncated << 31; uint32 counts[ceil(n / 8)]
 *   with reward[e][j] = `value` added counts[j] times (4 bits per agent; 0 -> 0.0): a step's rewards are 0 or ONE
 *   value per env (base.py:598-602), added once per unlocked door by the LockedHallway hook. The kernel checks that
 *   the record reproduces every float64 reward bit for bit and sets bit 4 (value 16) of the status word otherwise
 *   (bit 3: a cell outside the palette). n <= 31. Host decoder: multigrid_b200.engine.unpack_wire.
 */
int32_t mg_wire_record_bytes(int32_t num_agents);
int64_t mg_wire_obs_bytes(int32_t view_size, int32_t bits, int32_t num_agents, int64_t num_envs);
int64_t mg_wire_bytes(int32_t view_size, int32_t bits, int32_t num_agents, int64_t num_envs);
int mg_step_obs_host_wire(const MgConfig *cfg, int64_t num_envs, const MgState *state, const int8_t *h_actions,
                          int8_t *d_actions, const MgStepOut *d_out, uint8_t *d_wire, int32_t bits, const uint8_t *lut,
                          uint8_t *h_wire, void *stream);

/*
 * Prepared steps: mg_step_obs split into "validate + plan once" and "launch". A plan fixes the configuration
 * (flags included: build one per MG_FLAG_* variant you use), the batch size, every pointer of `state` and `out`, and
 * the launch geometry -- the MG_* tuning knobs are read when the plan is created; only the actions pointer
 * (16-byte aligned, as the one given at creation) and the stream change per call. mg_step_plan_run launches exactly
 * the kernel mg_step_obs would have launched (one driver call, ~3 us of host time instead of ~8). Not thread-safe
 * per plan; destroy it before freeing the buffers it points to, and re-create it when they move.
 */
typedef struct MgStepPlan MgStepPlan;
int mg_step_plan_create(const MgConfig *cfg, int64_t num_envs, const MgState *state, const int8_t *actions,
                        const MgStepOut *out, MgStepPlan **plan_out);
int mg_step_plan_run(MgStepPlan *plan, const int8_t *actions, void *stream);
void mg_step_plan_destroy(MgStepPlan *plan);

/*
 * Compact wire format for host consumers. The host-buffer path is bound by PCIe, and 99 % of its bytes are
 * observations (148 of 157 bytes per agent and step for V = 7). Every cell -- type < 16, colour < 8, state < 4
 * (core/constants.py:34-97; an agent cell is (10, colour, dir)) -- becomes the 9-bit code
 * type | colour << 4 | state << 7; the V*V codes of an agent are packed little-endian, cell (a, b) of image[a][b]
 * at bits [9*(a*V + b), +9) of its mg_packed_obs_stride(V)-byte record (56 bytes for V = 7, 96 for V = 9; the
 * record is a whole number of 8-byte words, unused bits are zero). Lossless.
 *   mg_pack_obs              obs int8 [A][obs_agent_stride] (device, 16-byte aligned) -> packed uint8 [A][stride]
 *   mg_step_obs_host_packed  mg_step_obs_host whose h_out->obs receives the PACKED observations (d_packed: device
 *                            scratch of num_envs * n * mg_packed_obs_stride(V) bytes); reward / terminated / truncated
 *                            as in mg_step_obs_host.
 * Replaces nothing in the reference (its observations never leave the host); it is the transport encoding of
 * MultiGridEnv.step's `image` observations (base.py:370) for a CPU-side consumer of a GPU-resident batch.
 */
int32_t mg_packed_obs_stride(int32_t view_size);
int mg_pack_obs(int32_t view_size, int64_t num_agents_total, int32_t obs_agent_stride, const int8_t *obs,
                uint8_t *packed, void *stream);
int mg_step_obs_host_packed(const MgConfig *cfg, int64_t num_envs, const MgState *state, const int8_t *h_actions,
                            int8_t *d_actions, const MgStepOut *d_out, uint8_t *d_packed, const MgStepOut *h_out,
                            void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MULTIGRID_B200_H */
