// TEST INFRASTRUCTURE (never part of the product): host build of the planned sparse kernel
// (grid2op_b200/csrc/b200pf_sparse.cuh compiled with B200PF_EMULATE, a "phase" = loop over the lanes) so that the
// plan builder and the kernel's numerics can be checked against the oracle on a machine without a GPU.
#define B200PF_EMULATE 1
#include "../../grid2op_b200/csrc/b200pf_sparse.cuh"
#include "../../include/b200pf.h"

#include <map>
#include <string>
#include <vector>

using namespace b200pf;

static HostGrid host_grid(const b200pf_grid_desc *gd) {
    HostGrid g;
    g.n_sub = gd->n_sub; g.n_busbar = gd->n_busbar; g.n_slot = gd->n_sub * gd->n_busbar; g.n_line = gd->n_line;
    g.n_gen = gd->n_gen; g.n_hidden = gd->n_hidden; g.n_unit = gd->n_gen + gd->n_hidden; g.n_load = gd->n_load;
    g.n_sto = gd->n_storage; g.n_shunt = gd->n_shunt; g.dim_topo = gd->dim_topo;
    g.n_topo_in = g.dim_topo + g.n_shunt + g.n_hidden; g.base_mva = gd->sn_mva;
#define CP(dst, src, n) g.dst.assign(gd->src, gd->src + (n))
    CP(line_or_sub, line_or_sub, g.n_line); CP(line_ex_sub, line_ex_sub, g.n_line); CP(line_or_pos, line_or_pos, g.n_line);
    CP(line_ex_pos, line_ex_pos, g.n_line); CP(line_y, line_y, 8 * g.n_line); CP(line_bdc, line_bdc, g.n_line);
    CP(line_pshift, line_pshift, g.n_line); CP(unit_sub, unit_sub, g.n_unit); CP(unit_pos, unit_pos, g.n_unit);
    CP(unit_is_ref, unit_is_ref, g.n_unit); CP(unit_qmin, unit_qmin, g.n_unit); CP(unit_qmax, unit_qmax, g.n_unit);
    CP(load_sub, load_sub, g.n_load); CP(load_pos, load_pos, g.n_load); CP(sto_sub, storage_sub, g.n_sto);
    CP(sto_pos, storage_pos, g.n_sto); CP(sh_sub, shunt_sub, g.n_shunt);
#undef CP
    return g;
}

extern "C" int sparse_emu_run(const b200pf_grid_desc *gd, int batch, const int8_t *topo, const double *inj, int is_dc, int max_iter,
                              double tol_mva, float *out, int32_t *status, int32_t *iters, double *busv, int n1_lines,
                              const float *th_lim, float *rho, int32_t *stats /* [8] of the last plan, may be NULL */) {
    HostGrid hg = host_grid(gd);
    PlanBuilder pb(hg);
    DevGrid g{};
    g.n_sub = hg.n_sub; g.n_busbar = hg.n_busbar; g.n_slot = hg.n_slot; g.n_line = hg.n_line; g.n_gen = hg.n_gen; g.n_hidden = hg.n_hidden;
    g.n_unit = hg.n_unit; g.n_load = hg.n_load; g.n_sto = hg.n_sto; g.n_shunt = hg.n_shunt; g.dim_topo = hg.dim_topo;
    g.n_topo_in = hg.n_topo_in;
    g.n_inj = g.n_gen + g.n_unit + 2 * g.n_load + g.n_sto + 2 * g.n_shunt;
    g.n_out = 10 * g.n_line + 4 * g.n_unit + 2 * g.n_load + g.n_sto + 3 * g.n_shunt;
    g.base_mva = hg.base_mva;
    g.line_y = gd->line_y; g.line_bdc = gd->line_bdc; g.line_pshift = gd->line_pshift; g.line_or_vn = gd->line_or_vn; g.line_ex_vn = gd->line_ex_vn;
    g.unit_is_ref = gd->unit_is_ref; g.unit_qmin = gd->unit_qmin; g.unit_qmax = gd->unit_qmax; g.unit_vn = gd->unit_vn;
    g.load_vn = gd->load_vn; g.sto_vn = gd->storage_vn; g.sh_vn = gd->shunt_vn; g.sto_q = gd->storage_q; g.sh_vratio = gd->shunt_vratio;
    RunArgs a{};
    a.batch = n1_lines > 0 ? batch * n1_lines : batch;
    a.inj = inj; a.is_dc = is_dc; a.max_iter = max_iter; a.tol_pu = tol_mva / hg.base_mva; a.out = out; a.status = status; a.iters = iters;
    a.busv = busv; a.n1_lines = n1_lines; a.th_lim = th_lim; a.rho = rho;
    std::map<std::string, std::vector<unsigned char>> cache;
    for (int inst = 0; inst < a.batch; ++inst) {
        const int src = n1_lines > 0 ? inst / n1_lines : inst, outage = n1_lines > 0 ? inst % n1_lines : -1;
        const int8_t *tv = topo + (size_t)src * hg.n_topo_in;
        std::string key((const char *)tv, hg.n_topo_in);
        key.push_back((char)(outage & 0xff)); key.push_back((char)((outage >> 8) & 0xff));
        auto it = cache.find(key);
        if (it == cache.end()) it = cache.emplace(key, pb.build(tv, outage)).first;
        const std::vector<unsigned char> &blob = it->second;
        const PlanHeader *H = (const PlanHeader *)blob.data();
        if (stats) { stats[0] = H->nb; stats[1] = H->d; stats[2] = H->nnzF; stats[3] = H->n_pass; stats[4] = H->n_op; stats[5] = H->pad[0];
                     stats[6] = H->smem_bytes; stats[7] = H->total_bytes; }
        std::vector<double> ws((size_t)H->smem_bytes / 8 + 4);
        PlanArgs pa{};
        int off0 = 0;
        pa.blobs = blob.data(); pa.plan_off = &off0; pa.inst_plan = nullptr;
        solve_sparse<32>(g, a, pa, inst, (unsigned char *)ws.data(), 0);
    }
    return 0;
}
