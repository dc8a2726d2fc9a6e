// TEST INFRASTRUCTURE: runs the phase functions of multigrid_b200/csrc/mg_kernels.cuh on the CPU,
// "thread" by "thread" and phase by phase (a phase boundary == __syncthreads), so the kernel logic
// can be checked against the oracle in a container without a GPU. Never loaded by the product.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../multigrid_b200/csrc/mg_kernels.cuh"

namespace {

template <int VT, int MODE>
void run_blocks(const mg::Params &p) {
    std::vector<uint8_t> smem_store(p.smem_bytes + 16);
    uint8_t *smem = smem_store.data();
    smem += (16 - (reinterpret_cast<uintptr_t>(smem) & 15)) & 15;
    const int nt = p.epb * p.tpe;
    const int blocks = (p.num_envs + p.epb - 1) / p.epb;
    for (int blk = 0; blk < blocks; blk++) {
        std::memset(smem, 0xCD, p.smem_bytes);  // poison: catches reads of unwritten smem
        for (int t = 0; t < nt; t++) mg::phase_load<MODE>(p, smem, blk, t, nt);
        if (MODE != mg::MODE_OBS && (p.flags & MG_FLAG_AUTO_RESET))
            for (int t = 0; t < nt; t++) mg::phase_reset(p, smem, blk, t, nt);
        for (int t = 0; t < nt; t++) mg::phase_convert(p, smem, blk, t, nt);
        for (int t = 0; t < nt; t++) mg::phase_step<MODE>(p, smem, blk, t, nt);
        if (MODE != mg::MODE_STEP)
            for (int t = 0; t < nt; t++) mg::phase_obs<VT>(p, smem, blk, t, nt);
        for (int t = 0; t < nt; t++) mg::phase_store<MODE>(p, smem, blk, t, nt);
    }
}

template <int MODE>
void dispatch(const mg::Params &p, int generic) {
    if (!generic) {
        switch (p.V) {
            case 3: return run_blocks<3, MODE>(p);
            case 5: return run_blocks<5, MODE>(p);
            case 7: return run_blocks<7, MODE>(p);
            case 9: return run_blocks<9, MODE>(p);
            default: break;
        }
    }
    run_blocks<0, MODE>(p);
}

}  // namespace

extern "C" int sim_run(int mode, const MgConfig *c, int64_t num_envs, const MgState *s,
                       const int8_t *actions, const MgStepOut *o, int forced_epb, int generic) {
    mg::Params p;
    std::memset(&p, 0, sizeof(p));
    p.W = c->width; p.H = c->height; p.n = c->num_agents; p.V = c->view_size;
    p.max_steps = c->max_steps; p.flags = c->flags; p.hook = c->hook;
    p.ostride = c->obs_agent_stride; p.K = c->num_layouts; p.lstride = c->layout_stride;
    p.num_envs = (int32_t)num_envs;
    if (mode == mg::MODE_OBS) p.flags &= ~MG_FLAG_AUTO_RESET;
    p.grid = s->grid; p.agents = s->agents; p.step_count = s->step_count;
    p.pcg_state = s->pcg_state; p.pcg_inc = s->pcg_inc; p.layout_idx = s->layout_idx;
    p.pool_grid = s->pool_grid; p.pool_agents = s->pool_agents;
    p.actions = actions;
    p.obs = o->obs; p.reward = o->reward; p.terminated = o->terminated; p.truncated = o->truncated;
    p.status = o->status;
    int rc = mg::plan_launch(p, forced_epb, 256, 200 * 1024);
    if (rc) return rc;
    if (mode == mg::MODE_OBS) dispatch<mg::MODE_OBS>(p, generic);
    else if (mode == mg::MODE_STEP) dispatch<mg::MODE_STEP>(p, generic);
    else dispatch<mg::MODE_STEP_OBS>(p, generic);
    return 0;
}
