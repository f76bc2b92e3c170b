// b200pf_small.cuh — one WARP per grid instance, for grids whose every element class fits 32 lanes
// (rte_case5, the 14-substation family): <= 32 bus slots, lines, units, loads, storages, shunts, and
// instances with <= 17 active buses (Newton system <= 32 unknowns, DC system <= 16).
//
// Compared with the generic group kernel (b200pf_kernel.cuh) the per-instance dependent chain is cut:
//   * lanes change role per phase: lane = line (flows, Jacobian off-diagonals), lane = bus (injections,
//     mismatch, diagonals, state update), lane = matrix row (Gauss-Jordan), lane = unit/load (read-back);
//   * topology in registers: active-slot bitmask + popc/fns numbering, reachability by bitmask sweeps;
//   * bus state (|V|, angle, P/Q spec, Yii) lives in registers of the bus lane; V and branch currents
//     are exchanged through small shared-memory tables (one STS.128 / LDS.128 each);
//   * Gauss-Jordan with the matrix in registers, pivot row broadcast by SHFL, pivot choice by one REDUX
//     on (|a| bits, lane) keys — no shared-memory round trip, no divergent branch in the step.
// Numerics are those of the generic kernel: fp64 state / mismatch / flows, fp32 Jacobian and solve.
#pragma once
#include "b200pf_kernel.cuh"

namespace b200pf {

struct SmallLayout {
    int off_V;      // double2 [32]
    int off_cur;    // double2 [64]
    int off_x;      // float [32] / double [16]
    int off_adj;    // unsigned [32]
    int off_b;      // bytes: brf[32] brt[32] colth[32] colv[32]
    int off_mat;    // matrix region
    int pitch_j;    // floats per Jacobian row
    int pitch_d;    // doubles per DC row
    int total;
};

__host__ __device__ inline SmallLayout small_layout(int nb_cap) {
    SmallLayout L;
    int o = 0;
    L.off_V = o; o += 32 * 16;
    L.off_cur = o; o += 64 * 16;
    L.off_x = o; o += 128;
    L.off_adj = o; o += 128;
    L.off_b = o; o += 128;
    L.off_mat = o;
    int d_cap = 2 * nb_cap - 2; if (d_cap > 32) d_cap = 32; if (d_cap < 2) d_cap = 2;
    int n1_cap = nb_cap - 1; if (n1_cap > 16) n1_cap = 16; if (n1_cap < 1) n1_cap = 1;
    L.pitch_j = (d_cap + 1) | 1;
    L.pitch_d = (n1_cap + 1) | 1;
    int ac = d_cap * L.pitch_j * 4, dc = n1_cap * L.pitch_d * 8;
    o += ((ac > dc ? ac : dc) + 15) & ~15;
    L.total = o;
    return L;
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// Gauss-Jordan, rows in registers, pivot row broadcast with SHFL (see header).  M is only read.
template <int DMAX, typename S>
__device__ __forceinline__ bool gj_warp_shfl(const S *M, int n, int pitch, S *xs, int lane) {
    S a[DMAX + 1];
    const bool mine = lane < n;
    const S *Mr = M + lane * pitch;
#pragma unroll
    for (int c = 0; c < DMAX; ++c) a[c] = (mine && c < n) ? Mr[c] : S(0);
    a[DMAX] = mine ? Mr[n] : S(0);
    bool used = !mine;
    int mycol = -1;
    S pivval = S(1);
#pragma unroll
    for (int k = 0; k < DMAX; ++k) {
        if (k >= n) break;                                            // warp-uniform
        const float mag = fabsf((float)a[k]);
        unsigned key = (__float_as_uint(mag) & 0xffffffe0u) | (unsigned)(31 - lane);
        if (used || !(mag == mag)) key = 0u;
        const unsigned mx = __reduce_max_sync(0xffffffffu, key);
        if (mx < 0x00800000u) return false;                           // |pivot| below the smallest normal: singular / NaN
        const int p = 31 - (int)(mx & 31u);
        const S piv = __shfl_sync(0xffffffffu, a[k], p);
        S inv;
        if (sizeof(S) == 4) inv = (S)__fdividef(1.0f, (float)piv); else inv = S(1) / piv;
        const bool isp = lane == p;
        const S m = isp ? S(0) : a[k] * inv;
#pragma unroll
        for (int c = k + 1; c < DMAX + 1; ++c) {
            const S pr = __shfl_sync(0xffffffffu, a[c], p);
            a[c] -= m * pr;
        }
        if (isp) { used = true; mycol = k; pivval = a[k]; }
    }
    if (mycol >= 0) xs[mycol] = a[DMAX] / pivval;
    __syncwarp();
    return true;
}

__device__ __forceinline__ bool gj_small_f(const float *M, int n, int pitch, float *xs, int lane) {
    if (n <= 8) return gj_warp_shfl<8, float>(M, n, pitch, xs, lane);
    if (n <= 16) return gj_warp_shfl<16, float>(M, n, pitch, xs, lane);
    if (n <= 24) return gj_warp_shfl<24, float>(M, n, pitch, xs, lane);
    return gj_warp_shfl<32, float>(M, n, pitch, xs, lane);
}
__device__ __forceinline__ bool gj_small_d(const double *M, int n, int pitch, double *xs, int lane) {
    if (n <= 8) return gj_warp_shfl<8, double>(M, n, pitch, xs, lane);
    return gj_warp_shfl<16, double>(M, n, pitch, xs, lane);
}

__device__ __forceinline__ void small_fail(const DevGrid &g, const RunArgs &a, int inst, int status, int iters, int lane) {
    float *out = a.out ? a.out + (size_t)inst * g.n_out : nullptr;
    if (lane == 0) { a.status[inst] = status; a.iters[inst] = iters; if (a.done) a.done[inst] = 1; }
    if (a.disc && status == ST_DONE) for (int k = lane; k < g.n_line; k += 32) a.disc[(size_t)inst * g.n_line + k] = -1;
    if (out) for (int k = lane; k < g.n_out; k += 32) out[k] = qnanf();
    if (a.busv) for (int k = lane; k < 2 * g.n_slot; k += 32) a.busv[(size_t)inst * 2 * g.n_slot + k] = __longlong_as_double(0x7ff8000000000000LL);
    if (a.rho) for (int k = lane; k < g.n_line; k += 32) a.rho[(size_t)inst * g.n_line + k] = qnanf();
    __syncwarp();
}

template <bool PROT>
__device__ void solve_small(const DevGrid &g, const RunArgs &a, const SmallLayout &L, int inst, unsigned char *sm, int lane) {
    const unsigned FULL = 0xffffffffu;
    double2 *Vs = reinterpret_cast<double2 *>(sm + L.off_V);
    double2 *cur = reinterpret_cast<double2 *>(sm + L.off_cur);
    float *xs = reinterpret_cast<float *>(sm + L.off_x);
    double *xd = reinterpret_cast<double *>(sm + L.off_x);
    unsigned *adj = reinterpret_cast<unsigned *>(sm + L.off_adj);
    signed char *brf = reinterpret_cast<signed char *>(sm + L.off_b);
    signed char *brt = brf + 32, *colth_s = brf + 64, *colv_s = brf + 96;
    float *J = reinterpret_cast<float *>(sm + L.off_mat);
    double *Md = reinterpret_cast<double *>(sm + L.off_mat);
    const int src = a.n1_lines > 0 ? inst / a.n1_lines : inst;          // record the inputs come from
    const int outage = a.n1_lines > 0 ? inst % a.n1_lines : -1;         // line forced out of service (N-1 sweep)
    const int8_t *tv = a.topo + (size_t)src * g.n_topo_in;
    float *out = a.out ? a.out + (size_t)inst * g.n_out : nullptr;
    const double base = g.base_mva;
    const int nl = g.n_line, nu = g.n_unit, nh = g.n_hidden, ng = g.n_gen, nld = g.n_load, nst = g.n_sto, nsh = g.n_shunt;
    const unsigned lt = (1u << lane) - 1u;

    // ---- 0. this lane's injections (lane = unit / load / storage / shunt index) ---------------------
    double u_p = 0.0, u_vm = 1.0, l_p = 0.0, l_q = 0.0, s_p = 0.0, sh_p = 0.0, sh_q = 0.0;
    if (a.series) {
        const int sc = a.rows ? 0 : a.scen[src], trow = a.rows ? 0 : a.t[src];
        const float *row = a.rows ? a.rows + (size_t)src * (size_t)(2 * nld + 2 * ng)
                                  : a.chron + ((size_t)sc * a.n_rows + trow) * (size_t)(2 * nld + 2 * ng);
        const double *si = a.static_inj;
        if (lane < nld) { l_p = (double)row[lane]; l_q = (double)row[nld + lane]; }
        if (lane < nu) {
            if (lane >= nh) {
                u_p = (double)row[2 * nld + lane - nh];
                u_vm = (double)__fdiv_rn(row[2 * nld + ng + lane - nh], g.unit_vn[lane]);   // float32 / float32, pPB:927
            } else u_vm = si[ng + lane];
        }
        if (lane < nst) s_p = si[ng + nu + 2 * nld + lane];
        if (lane < nsh) { sh_p = si[ng + nu + 2 * nld + nst + lane]; sh_q = si[ng + nu + 2 * nld + nst + nsh + lane]; }
        __syncwarp();
        if (lane == 0 && !a.rows && a.n1_lines <= 0) a.t[inst] = (trow + 1 >= a.n_rows) ? 0 : trow + 1;
    } else {
        const double *sp = a.inj + (size_t)src * g.n_inj;
        if (lane < nu) { u_vm = sp[ng + lane]; if (lane >= nh) u_p = sp[lane - nh]; }
        if (lane < nld) { l_p = sp[ng + nu + lane]; l_q = sp[ng + nu + nld + lane]; }
        if (lane < nst) s_p = sp[ng + nu + 2 * nld + lane];
        if (lane < nsh) { sh_p = sp[ng + nu + 2 * nld + nst + lane]; sh_q = sp[ng + nu + 2 * nld + nst + nsh + lane]; }
    }

    // protections state (lane = line)
    if (PROT && a.done[inst]) { small_fail(g, a, inst, ST_DONE, 0, lane); return; }
    int pc_env = 0, pc = 0, disc_it = -1;
    bool forced_off = false, inc_done = false;
    if (PROT && lane < nl) { pc_env = a.pcount[(size_t)inst * nl + lane]; pc = pc_env; }
    for (int casc = 0;; ++casc) {
    // ---- 1. topology: active bus slots as a bit mask ------------------------------------------------
    int slot_o = -1, slot_e = -1, slot_u = -1, slot_k = -1, slot_s = -1, slot_h = -1;
    if (lane < nl && lane != outage && !(PROT && forced_off)) {
        const int bo = tv[g.line_or_pos[lane]], be = tv[g.line_ex_pos[lane]];
        if (bo > 0) slot_o = g.line_or_sub[lane] + (bo - 1) * g.n_sub;
        if (be > 0) slot_e = g.line_ex_sub[lane] + (be - 1) * g.n_sub;
    }
    if (lane < nu) { const int b = tv[g.unit_pos[lane]]; if (b > 0) slot_u = g.unit_sub[lane] + (b - 1) * g.n_sub; }
    if (lane < nld) { const int b = tv[g.load_pos[lane]]; if (b > 0) slot_k = g.load_sub[lane] + (b - 1) * g.n_sub; }
    if (lane < nst) { const int b = tv[g.sto_pos[lane]]; if (b > 0) slot_s = g.sto_sub[lane] + (b - 1) * g.n_sub; }
    if (lane < nsh) { const int b = tv[g.dim_topo + lane]; if (b > 0) slot_h = g.sh_sub[lane] + (b - 1) * g.n_sub; }
    unsigned m = 0;
    if (slot_o >= 0) m |= 1u << slot_o;
    if (slot_e >= 0) m |= 1u << slot_e;
    if (slot_u >= 0) m |= 1u << slot_u;
    if (slot_k >= 0) m |= 1u << slot_k;
    if (slot_s >= 0) m |= 1u << slot_s;
    if (slot_h >= 0) m |= 1u << slot_h;
    const unsigned amask = __reduce_or_sync(FULL, m);
    const int nb = __popc(amask);
    if (nb > a.nb_cap || nb > 17) { small_fail(g, a, inst, ST_LARGE, 0, lane); return; }
#define CIDX(s) (__popc(amask & ((1u << (s)) - 1u)))
    const int bf = (slot_o >= 0 && slot_e >= 0) ? CIDX(slot_o) : -1;
    const int bt = (slot_o >= 0 && slot_e >= 0) ? CIDX(slot_e) : -1;
    const int bu = slot_u >= 0 ? CIDX(slot_u) : -1;
    const int bk = slot_k >= 0 ? CIDX(slot_k) : -1;
    const int bs = slot_s >= 0 ? CIDX(slot_s) : -1;
    const int bh = slot_h >= 0 ? CIDX(slot_h) : -1;
    const bool isbus = lane < nb;
    const int myslot = isbus ? (int)__fns(amask, 0, lane + 1) : 0;
    const int mysub = myslot % g.n_sub;
    brf[lane] = (signed char)bf; brt[lane] = (signed char)bt;
    adj[lane] = 0u;
    __syncwarp();
    if (bf >= 0) { atomicOr(&adj[bf], 1u << bt); atomicOr(&adj[bt], 1u << bf); }

    // ---- 2. per-bus aggregates, in element order (deterministic; the later unit wins the set point) -
    int btype = BT_PQ, cnt = 0, nref = 0;
    double vm = 1.0, pg = 0.0, pd = 0.0, qd = 0.0, gsh = 0.0, bsh = 0.0, qmins = 0.0, qmaxs = 0.0, pnonref = 0.0;
    {
        const int isref = (lane < nu) ? g.unit_is_ref[lane] : 0;
        const double qmn = (lane < nu) ? g.unit_qmin[lane] : 0.0, qmx = (lane < nu) ? g.unit_qmax[lane] : 0.0;
        for (int u = 0; u < nu; ++u) {
            const int b = __shfl_sync(FULL, bu, u);
            const int r = __shfl_sync(FULL, isref, u);
            const double p = shfl_d(u_p, u), v = shfl_d(u_vm, u), q1 = shfl_d(qmn, u), q2 = shfl_d(qmx, u);
            if (b == lane && isbus) {
                vm = v; cnt++; qmins += q1; qmaxs += q2; pg += p;
                if (r) { btype = BT_REF; nref++; } else { if (btype != BT_REF) btype = BT_PV; pnonref += p; }
            }
        }
        for (int k = 0; k < nld; ++k) {
            const int b = __shfl_sync(FULL, bk, k);
            const double p = shfl_d(l_p, k), q = shfl_d(l_q, k);
            if (b == lane && isbus) { pd += p; qd += q; }
        }
        const double sq = (lane < nst) ? g.sto_q[lane] : 0.0;
        for (int k = 0; k < nst; ++k) {
            const int b = __shfl_sync(FULL, bs, k);
            const double p = shfl_d(s_p, k), q = shfl_d(sq, k);
            if (b == lane && isbus) { pd += p; qd += q; }
        }
        const double vr = (lane < nsh) ? g.sh_vratio[lane] : 0.0;
        for (int k = 0; k < nsh; ++k) {
            const int b = __shfl_sync(FULL, bh, k);
            const double p = shfl_d(sh_p * vr, k), q = shfl_d(sh_q * vr, k);
            if (b == lane && isbus) { gsh += p; bsh -= q; }
        }
    }
    // ---- 3. reachability from the reference buses -------------------------------------------------
    __syncwarp();
    const unsigned myadj = isbus ? adj[lane] : 0u;
    const unsigned refmask = __ballot_sync(FULL, isbus && btype == BT_REF);
    if (!refmask) { small_fail(g, a, inst, ST_NOREF, 0, lane); return; }
    unsigned reach = refmask;
    for (int sweep = 0; sweep < 32; ++sweep) {
        const unsigned nr = reach | __reduce_or_sync(FULL, ((reach >> lane) & 1u) ? myadj : 0u);
        if (nr == reach) break;
        reach = nr;
    }
    const unsigned allmask = (nb >= 32) ? FULL : ((1u << nb) - 1u);
    if ((reach & allmask) != allmask) { small_fail(g, a, inst, ST_UNSUP, 0, lane); return; }

    // ---- 4. unknown numbering ---------------------------------------------------------------------
    const unsigned nonref = __ballot_sync(FULL, isbus && btype != BT_REF);
    const unsigned pqm = __ballot_sync(FULL, isbus && btype == BT_PQ);
    const int n1 = __popc(nonref), d = n1 + __popc(pqm);
    const int colth = (isbus && btype != BT_REF) ? __popc(nonref & lt) : -1;
    const int colv = (isbus && btype == BT_PQ) ? n1 + __popc(pqm & lt) : -1;
    if (n1 > 16 || d > 32 || n1 * L.pitch_d * 8 > (L.total - L.off_mat) || (!a.is_dc && d * L.pitch_j * 4 > (L.total - L.off_mat))) {
        small_fail(g, a, inst, ST_LARGE, 0, lane); return;
    }
    colth_s[lane] = (signed char)colth; colv_s[lane] = (signed char)colv;
    const double pspec = (pg - pd) / base, qspec = -qd / base;
    __syncwarp();

    // ---- 5. Ybus diagonal and the DC system (bus lane = matrix row) ---------------------------------
    double gii = gsh / base, bii = bsh / base, va = 0.0;
    {
        const int pitch = (n1 + 1) | 1;
        double *Mr = Md + (colth >= 0 ? colth : 0) * pitch;
        if (colth >= 0) for (int c = 0; c <= n1; ++c) Mr[c] = 0.0;
        double rhs = pspec - gsh / base, dsum = 0.0;
        if (isbus) {
            for (int e = g.sub_end_ptr[mysub]; e < g.sub_end_ptr[mysub + 1]; ++e) {
                const int code = g.sub_end[e], l = code >> 1, side = code & 1;
                const int f = brf[l];
                if (f < 0) continue;
                const int t = brt[l];
                if ((side ? t : f) != lane) continue;
                const int j = side ? f : t;
                const double *y = g.line_y + l * 8 + (side ? 6 : 0);
                gii += y[0]; bii += y[1];
                const double b = g.line_bdc[l];
                dsum += b;
                const int cj = colth_s[j];
                if (colth >= 0 && cj >= 0) Mr[cj] -= b;                 // reference angle is 0 (pPB:473)
                rhs -= side ? -g.line_pshift[l] : g.line_pshift[l];
            }
            if (colth >= 0) { Mr[colth] = dsum; Mr[n1] = rhs; }
        }
        __syncwarp();
        if (!gj_small_d(Md, n1, pitch, xd, lane)) { small_fail(g, a, inst, ST_DIV, 0, lane); return; }
        if (colth >= 0) va = xd[colth];
        const int bad = isbus && !isfinite(va);
        if (__any_sync(FULL, bad)) { small_fail(g, a, inst, ST_DIV, 0, lane); return; }
        __syncwarp();
    }

    // ---- 6. Newton-Raphson ----------------------------------------------------------------------------
    const double RAD2DEG = 57.295779513082320877, SQRT3 = 1.7320508075688772935;
    double ve = 0.0, vf = 0.0, P = 0.0, Q = 0.0;
    int iters = 0;
    if (!a.is_dc) {
        { double s, c; sincos(va, &s, &c); ve = vm * c; vf = vm * s; }
        bool conv = false;
        const int pitch = (d + 1) | 1;
        for (int it = 0;; ++it) {
            if (isbus) Vs[lane] = make_double2(ve, vf);
            __syncwarp();
            double2 Vf = make_double2(0, 0), Vt = make_double2(0, 0);
            if (bf >= 0) {                                            // lane = line: currents at both ends
                Vf = Vs[bf]; Vt = Vs[bt];
                const double4 *y4 = reinterpret_cast<const double4 *>(g.line_y + lane * 8);
                const double4 ya = y4[0], yb = y4[1];                   // (yff, yft), (ytf, ytt)
                cur[2 * lane] = make_double2(ya.x * Vf.x - ya.y * Vf.y + ya.z * Vt.x - ya.w * Vt.y,
                                             ya.x * Vf.y + ya.y * Vf.x + ya.z * Vt.y + ya.w * Vt.x);
                cur[2 * lane + 1] = make_double2(yb.x * Vf.x - yb.y * Vf.y + yb.z * Vt.x - yb.w * Vt.y,
                                                 yb.x * Vf.y + yb.y * Vf.x + yb.z * Vt.y + yb.w * Vt.x);
            }
            __syncwarp();
            int viol = 0, wild = 0;
            double dP = 0.0, dQ = 0.0;
            if (isbus) {                                              // lane = bus: S = V conj(I)
                double ir = (gsh * ve - bsh * vf) / base, ii = (gsh * vf + bsh * ve) / base;
                for (int e = g.sub_end_ptr[mysub]; e < g.sub_end_ptr[mysub + 1]; ++e) {
                    const int code = g.sub_end[e], l = code >> 1;
                    const int f = brf[l];
                    if (f < 0 || ((code & 1) ? brt[l] : f) != lane) continue;
                    const double2 c2 = cur[code];
                    ir += c2.x; ii += c2.y;
                }
                P = ve * ir + vf * ii; Q = vf * ir - ve * ii;
                if (btype != BT_REF) { dP = P - pspec; const double m1 = fabs(dP); viol |= !(m1 < a.tol_pu); wild |= !(m1 < 1e200); }
                if (btype == BT_PQ) { dQ = Q - qspec; const double m2 = fabs(dQ); viol |= !(m2 < a.tol_pu); wild |= !(m2 < 1e200); }
            }
            if (!__any_sync(FULL, viol)) { conv = true; iters = it; break; }
            if (it >= a.max_iter || __any_sync(FULL, wild)) { iters = it; break; }
            // Jacobian (fp32): zero, diagonals by the bus lanes, off-diagonals by the line lanes
            if (lane < d) { float *Jr = J + lane * pitch; for (int c = 0; c <= d; ++c) Jr[c] = 0.f; }
            __syncwarp();
            if (colth >= 0) {
                const float vi2 = (float)(ve * ve + vf * vf), Pf = (float)P, Qf = (float)Q, gi = (float)gii, bi = (float)bii;
                float *JP = J + colth * pitch;
                JP[colth] = -Qf - bi * vi2; JP[d] = (float)(-dP);
                if (colv >= 0) {
                    float *JQ = J + colv * pitch;
                    JP[colv] = Pf + gi * vi2;
                    JQ[colth] = Pf - gi * vi2; JQ[colv] = Qf - bi * vi2; JQ[d] = (float)(-dQ);
                }
            }
            if (bf >= 0) {
                const double4 *y4 = reinterpret_cast<const double4 *>(g.line_y + lane * 8);
                const double4 ya = y4[0], yb = y4[1];
                const float ef = (float)Vf.x, ff = (float)Vf.y, et = (float)Vt.x, ft = (float)Vt.y;
                {   // rows of bus f, columns of bus t:  T = Vf conj(yft Vt)
                    const float yr = (float)ya.z, yi = (float)ya.w;
                    const float ar = yr * et - yi * ft, ai = yr * ft + yi * et;
                    const float tr = ef * ar + ff * ai, ti = ff * ar - ef * ai;
                    const int rp = colth_s[bf], rq = colv_s[bf], cth = colth_s[bt], cv = colv_s[bt];
                    if (rp >= 0) { if (cth >= 0) atomicAdd(&J[rp * pitch + cth], ti); if (cv >= 0) atomicAdd(&J[rp * pitch + cv], tr); }
                    if (rq >= 0) { if (cth >= 0) atomicAdd(&J[rq * pitch + cth], -tr); if (cv >= 0) atomicAdd(&J[rq * pitch + cv], ti); }
                }
                {   // rows of bus t, columns of bus f:  T = Vt conj(ytf Vf)
                    const float yr = (float)yb.x, yi = (float)yb.y;
                    const float ar = yr * ef - yi * ff, ai = yr * ff + yi * ef;
                    const float tr = et * ar + ft * ai, ti = ft * ar - et * ai;
                    const int rp = colth_s[bt], rq = colv_s[bt], cth = colth_s[bf], cv = colv_s[bf];
                    if (rp >= 0) { if (cth >= 0) atomicAdd(&J[rp * pitch + cth], ti); if (cv >= 0) atomicAdd(&J[rp * pitch + cv], tr); }
                    if (rq >= 0) { if (cth >= 0) atomicAdd(&J[rq * pitch + cth], -tr); if (cv >= 0) atomicAdd(&J[rq * pitch + cv], ti); }
                }
            }
            __syncwarp();
            if (!gj_small_f(J, d, pitch, xs, lane)) { iters = it + 1; break; }
            if (isbus) {
                if (colth >= 0) va += (double)xs[colth];
                if (colv >= 0) vm *= 1.0 + (double)xs[colv];
                if (vm < 0.0) { vm = -vm; va += 3.14159265358979323846; }
                double s, c; sincos(va, &s, &c); ve = vm * c; vf = vm * s;
            }
            __syncwarp();
        }
        if (!conv) { small_fail(g, a, inst, ST_DIV, iters, lane); return; }
    } else {
        // DC: publish angles, line lanes compute flows, bus lanes sum them for the slack share
        if (isbus) Vs[lane] = make_double2(va, vm);
        __syncwarp();
        if (bf >= 0) {
            const double pfl = g.line_bdc[lane] * (Vs[bf].x - Vs[bt].x) + g.line_pshift[lane];
            cur[2 * lane] = make_double2(pfl, 0.0); cur[2 * lane + 1] = make_double2(-pfl, 0.0);
        }
        __syncwarp();
        if (isbus) {
            double p = gsh / base;
            for (int e = g.sub_end_ptr[mysub]; e < g.sub_end_ptr[mysub + 1]; ++e) {
                const int code = g.sub_end[e], l = code >> 1;
                const int f = brf[l];
                if (f < 0 || ((code & 1) ? brt[l] : f) != lane) continue;
                p += cur[code].x;
            }
            P = p; Q = 0.0;
        }
    }

    // ---- 7. results -------------------------------------------------------------------------------------
    if (lane == 0) { a.status[inst] = ST_OK; a.iters[inst] = iters; }
    float r[10] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    {   // lines (lane = line): flows from the currents of the last mismatch evaluation
        const int sf_ = bf >= 0 ? bf : 0, st_ = bt >= 0 ? bt : 0;
        const double vmf = shfl_d(vm, sf_), vmt = shfl_d(vm, st_), vaf = shfl_d(va, sf_), vat = shfl_d(va, st_);
        if (lane < nl && bf >= 0) {
            double pf, qf, pt, qt, sf, st;
            if (a.is_dc) {
                pf = cur[2 * lane].x * base; pt = -pf; qf = 0.0; qt = 0.0; sf = fabs(pf); st = sf;
            } else {
                const double2 A = Vs[bf], B = Vs[bt], If = cur[2 * lane], It = cur[2 * lane + 1];
                pf = (A.x * If.x + A.y * If.y) * base; qf = (A.y * If.x - A.x * If.y) * base;
                pt = (B.x * It.x + B.y * It.y) * base; qt = (B.y * It.x - B.x * It.y) * base;
                sf = sqrt(pf * pf + qf * qf); st = sqrt(pt * pt + qt * qt);
            }
            const float vnf = g.line_or_vn[lane], vnt = g.line_ex_vn[lane];
            float a1 = (float)(sf / (SQRT3 * (vmf * (double)vnf)) * 1000.0), a2 = (float)(st / (SQRT3 * (vmt * (double)vnt)) * 1000.0);
            if (!isfinite(a1)) a1 = 0.f;
            if (!isfinite(a2)) a2 = 0.f;
            r[0] = (float)pf; r[1] = (float)qf; r[2] = __fmul_rn((float)vmf, vnf); r[3] = a1; r[4] = (float)(vaf * RAD2DEG);
            r[5] = (float)pt; r[6] = (float)qt; r[7] = __fmul_rn((float)vmt, vnt); r[8] = a2; r[9] = (float)(vat * RAD2DEG);
        }
    }
    if (PROT) {
        // cascading failure step (reference Backend.next_grid_state, backend.py:1466-1521), lane = line
        const bool on = lane < nl && bf >= 0;
        const float lim = lane < nl ? a.th_lim[lane] : 0.f;
        bool to_disc = on && (r[3] > __fmul_rn(a.hard_thr, lim));
        const bool mask_inc = !a.from_reset && on && (r[3] > __fmul_rn(a.soft_thr, lim)) && !inc_done;
        if (mask_inc) { pc += 1; inc_done = true; }
        if (on && pc > a.max_pc) to_disc = true;
        if (__any_sync(FULL, to_disc)) {
            if (to_disc) { forced_off = true; disc_it = casc; }
            __syncwarp();
            continue;                                   // re-solve with the lines removed
        }
        if (lane < nl) {
            int *pcp = a.pcount + (size_t)inst * nl, *tsp = a.ts_over + (size_t)inst * nl;
            // counters: baseEnv.py:3361-3370; Environment.reset() zeroes them after its step (baseEnv.py:3969-3970)
            pcp[lane] = (!a.from_reset && r[3] > __fmul_rn(a.soft_thr, lim)) ? pc_env + 1 : 0;
            tsp[lane] = (!a.from_reset && r[3] > lim) ? tsp[lane] + 1 : 0;
            a.disc[(size_t)inst * nl + lane] = disc_it;
            if (forced_off) {                           // persist the outage (_BackendAction.update_state)
                int8_t *tw = const_cast<int8_t *>(tv);
                tw[g.line_or_pos[lane]] = -1; tw[g.line_ex_pos[lane]] = -1;
            }
        }
    }
    if (lane < nl) {
        if (out) {
#pragma unroll
            for (int k = 0; k < 10; ++k) out[k * nl + lane] = r[k];
        }
        if (a.rho) a.rho[(size_t)inst * nl + lane] = r[3] / a.th_lim[lane];
    }
    {   // units (lane = unit), loads, storages, shunts: read their bus through shuffles
        const int su = bu >= 0 ? bu : 0;
        const double Pb = shfl_d(P, su), Qb = shfl_d(Q, su), pdb = shfl_d(pd, su), qdb = shfl_d(qd, su), pnr = shfl_d(pnonref, su);
        const double qmn = shfl_d(qmins, su), qmx = shfl_d(qmaxs, su), vmb = shfl_d(vm, su), vab = shfl_d(va, su);
        const int cb = __shfl_sync(FULL, cnt, su), nrb = __shfl_sync(FULL, nref, su);
        if (lane < nu && out) {
            float p = 0.f, q = 0.f, v = 0.f, th = 0.f;
            if (bu >= 0) {
                double pu = u_p;
                if (g.unit_is_ref[lane]) pu = (Pb * base + pdb - pnr) / (double)nrb;     // slack share (pandapower pfsoln)
                double qu = 0.0;
                if (!a.is_dc) {
                    const double qtot = Qb * base + qdb;
                    if (cb <= 1 || qmn == qmx) qu = qtot / (double)cb;
                    else qu = g.unit_qmin[lane] + (qtot - qmn) / (qmx - qmn + 2.220446049250313e-16) * (g.unit_qmax[lane] - g.unit_qmin[lane]);
                }
                p = (float)pu; q = (float)qu; v = __fmul_rn((float)vmb, g.unit_vn[lane]); th = (float)(vab * RAD2DEG);
            }
            float *o = out + 10 * nl;
            o[lane] = p; o[nu + lane] = q; o[2 * nu + lane] = v; o[3 * nu + lane] = th;
        }
        const int sk = bk >= 0 ? bk : 0;
        const double vmk = shfl_d(vm, sk), vak = shfl_d(va, sk);
        if (lane < nld && out) {
            float *o = out + 10 * nl + 4 * nu;
            o[lane] = bk >= 0 ? __fmul_rn((float)vmk, g.load_vn[lane]) : 0.f;
            o[nld + lane] = bk >= 0 ? (float)(vak * RAD2DEG) : 0.f;
        }
        const int ss = bs >= 0 ? bs : 0;
        const double vms = shfl_d(vm, ss);
        if (lane < nst && out) out[10 * nl + 4 * nu + 2 * nld + lane] = bs >= 0 ? __fmul_rn((float)vms, g.sto_vn[lane]) : 0.f;
        const int sh = bh >= 0 ? bh : 0;
        const double vmh = shfl_d(vm, sh);
        if (lane < nsh && out) {
            float *o = out + 10 * nl + 4 * nu + 2 * nld + nst;
            float p = 0.f, q = 0.f, v = 0.f;
            if (bh >= 0) {
                const double v2 = a.is_dc ? 1.0 : vmh * vmh;
                p = (float)(sh_p * g.sh_vratio[lane] * v2);
                q = a.is_dc ? 0.f : (float)(sh_q * g.sh_vratio[lane] * v2);
                v = __fmul_rn((float)vmh, g.sh_vn[lane]);
            }
            o[lane] = p; o[nsh + lane] = q; o[2 * nsh + lane] = v;
        }
    }
    if (a.busv) {
        double *bv = a.busv + (size_t)inst * 2 * g.n_slot;
        const double nanv = __longlong_as_double(0x7ff8000000000000LL);
        if (lane < g.n_slot && !((amask >> lane) & 1u)) { bv[lane] = nanv; bv[g.n_slot + lane] = nanv; }
        if (isbus) { bv[myslot] = vm; bv[g.n_slot + myslot] = va; }
    }
    __syncwarp();
    break;
    }   // cascade loop
#undef CIDX
}

template <bool PROT>
__global__ void __launch_bounds__(32, B200PF_MIN_WARP_CTAS)
pf_kernel_small(const DevGrid g, const RunArgs a, const SmallLayout L) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x;
    for (int inst = blockIdx.x; inst < a.batch; inst += gridDim.x) solve_small<PROT>(g, a, L, inst, smem, lane);
}

}  // namespace b200pf
