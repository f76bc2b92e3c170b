/*
 * ORACLE (test infrastructure / CPU baseline; never linked into or called from the product path).
 *
 * Plain-C fp64 restatement, at the packed-record level of include/b200pf.h, of the algorithm
 * oracle/pandapower_ref.py restates from pandapower (reference call sites
 * grid2op/Backend/pandaPowerBackend.py:1090, 1097-1105; results read at :1122-1218):
 *   topo_vect -> in-service buses -> bus types -> connectivity -> dense Ybus ->
 *   DC angles (PYPOWER dcpf) -> full Newton in polar form (PYPOWER newtonpf, dense LU with partial
 *   pivoting) -> branch flows, slack share and reactive sharing (PYPOWER pfsoln).
 * It is validated against the numpy oracle (tests/test_c_oracle.py), which is itself pinned to the
 * results real pandapower stored in the reference's grid.json files.
 * bench.py times it on the host cores ("cpu_baseline", kind "port") with OpenMP over instances.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/b200pf.h"

#define IDX(i, j, n) ((size_t)(i) * (n) + (j))

static int lu_solve(double *A, double *b, int n) { /* in place, partial pivoting; 0 ok */
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(A[IDX(k, k, n)]);
        for (int r = k + 1; r < n; ++r) {
            double a = fabs(A[IDX(r, k, n)]);
            if (a > best) { best = a; p = r; }
        }
        if (!(best > 1e-300)) return 1;
        if (p != k) {
            for (int c = 0; c < n; ++c) { double t = A[IDX(k, c, n)]; A[IDX(k, c, n)] = A[IDX(p, c, n)]; A[IDX(p, c, n)] = t; }
            double t = b[k]; b[k] = b[p]; b[p] = t;
        }
        const double inv = 1.0 / A[IDX(k, k, n)];
        for (int r = k + 1; r < n; ++r) {
            const double m = A[IDX(r, k, n)] * inv;
            if (m == 0.0) continue;
            for (int c = k + 1; c < n; ++c) A[IDX(r, c, n)] -= m * A[IDX(k, c, n)];
            b[r] -= m * b[k];
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        double s = b[k];
        for (int c = k + 1; c < n; ++c) s -= A[IDX(k, c, n)] * b[c];
        b[k] = s / A[IDX(k, k, n)];
    }
    return 0;
}

typedef struct {
    int nslot, nl;
    int *cidx, *brf, *brt, *btype, *reach, *colth, *colv, *cnt, *nref;
    double *vm, *va, *pg, *pd, *qd, *gs, *bs, *qmins, *qmaxs, *pnonref, *pcalc, *qcalc;
    double *Yr, *Yi, *J, *rhs, *B;
} work_t;

static void *xm(size_t n) { void *p = malloc(n ? n : 1); return p; }

static work_t *work_new(const b200pf_grid_desc *g) {
    work_t *w = (work_t *)calloc(1, sizeof(work_t));
    int ns = g->n_sub * g->n_busbar, nl = g->n_line;
    w->nslot = ns; w->nl = nl;
    w->cidx = xm(sizeof(int) * ns); w->brf = xm(sizeof(int) * nl); w->brt = xm(sizeof(int) * nl);
    w->btype = xm(sizeof(int) * ns); w->reach = xm(sizeof(int) * ns); w->colth = xm(sizeof(int) * ns); w->colv = xm(sizeof(int) * ns);
    w->cnt = xm(sizeof(int) * ns); w->nref = xm(sizeof(int) * ns);
    double **d[] = {&w->vm, &w->va, &w->pg, &w->pd, &w->qd, &w->gs, &w->bs, &w->qmins, &w->qmaxs, &w->pnonref, &w->pcalc, &w->qcalc};
    for (unsigned k = 0; k < sizeof(d) / sizeof(d[0]); ++k) *d[k] = xm(sizeof(double) * ns);
    w->Yr = xm(sizeof(double) * ns * ns); w->Yi = xm(sizeof(double) * ns * ns);
    w->J = xm(sizeof(double) * 4 * ns * ns); w->rhs = xm(sizeof(double) * 2 * ns); w->B = xm(sizeof(double) * ns * ns);
    return w;
}

static void work_free(work_t *w) {
    free(w->cidx); free(w->brf); free(w->brt); free(w->btype); free(w->reach); free(w->colth); free(w->colv); free(w->cnt); free(w->nref);
    free(w->vm); free(w->va); free(w->pg); free(w->pd); free(w->qd); free(w->gs); free(w->bs); free(w->qmins); free(w->qmaxs);
    free(w->pnonref); free(w->pcalc); free(w->qcalc); free(w->Yr); free(w->Yi); free(w->J); free(w->rhs); free(w->B);
    free(w);
}

static int solve_one(const b200pf_grid_desc *g, work_t *w, const int8_t *tv, const double *inj, int is_dc, int max_iter,
                     double tol_pu, float *out, int *iters_out) {
    const int nsub = g->n_sub, nl = g->n_line, ng = g->n_gen, nh = g->n_hidden, nu = ng + nh;
    const int nld = g->n_load, nst = g->n_storage, nsh = g->n_shunt, dt = g->dim_topo, ns = w->nslot;
    const double base = g->sn_mva;
    const double *gen_p = inj, *unit_vm = inj + ng, *load_p = unit_vm + nu, *load_q = load_p + nld;
    const double *sto_p = load_q + nld, *sh_p = sto_p + nst, *sh_q = sh_p + nsh;
    const int n_out = 10 * nl + 4 * nu + 2 * nld + nst + 3 * nsh;
    *iters_out = 0;
    for (int k = 0; k < n_out; ++k) out[k] = NAN;
    /* active slots */
    for (int s = 0; s < ns; ++s) w->cidx[s] = 0;
#define MARK(sub, b) do { if ((b) > 0) w->cidx[(sub) + ((b) - 1) * nsub] = 1; } while (0)
    for (int l = 0; l < nl; ++l) { MARK(g->line_or_sub[l], tv[g->line_or_pos[l]]); MARK(g->line_ex_sub[l], tv[g->line_ex_pos[l]]); }
    for (int u = 0; u < nu; ++u) MARK(g->unit_sub[u], tv[g->unit_pos[u]]);
    for (int k = 0; k < nld; ++k) MARK(g->load_sub[k], tv[g->load_pos[k]]);
    for (int k = 0; k < nst; ++k) MARK(g->storage_sub[k], tv[g->storage_pos[k]]);
    for (int k = 0; k < nsh; ++k) MARK(g->shunt_sub[k], tv[dt + k]);
    int nb = 0;
    for (int s = 0; s < ns; ++s) w->cidx[s] = w->cidx[s] ? nb++ : -1;
#define BUS(sub, b) (w->cidx[(sub) + ((b) - 1) * nsub])
    for (int i = 0; i < nb; ++i) {
        w->btype[i] = 1; w->reach[i] = 0; w->vm[i] = 1.0; w->va[i] = 0.0; w->pg[i] = 0; w->pd[i] = 0; w->qd[i] = 0;
        w->gs[i] = 0; w->bs[i] = 0; w->qmins[i] = 0; w->qmaxs[i] = 0; w->pnonref[i] = 0; w->cnt[i] = 0; w->nref[i] = 0;
    }
    for (int u = 0; u < nu; ++u) {
        int b = tv[g->unit_pos[u]];
        if (b <= 0) continue;
        int i = BUS(g->unit_sub[u], b);
        double pu = u >= nh ? gen_p[u - nh] : 0.0;
        w->vm[i] = unit_vm[u]; w->cnt[i]++; w->qmins[i] += g->unit_qmin[u]; w->qmaxs[i] += g->unit_qmax[u];
        if (g->unit_is_ref[u]) { w->btype[i] = 3; w->nref[i]++; }
        else { if (w->btype[i] != 3) w->btype[i] = 2; w->pnonref[i] += pu; }
        w->pg[i] += pu;
    }
    for (int k = 0; k < nld; ++k) { int b = tv[g->load_pos[k]]; if (b > 0) { int i = BUS(g->load_sub[k], b); w->pd[i] += load_p[k]; w->qd[i] += load_q[k]; } }
    for (int k = 0; k < nst; ++k) { int b = tv[g->storage_pos[k]]; if (b > 0) { int i = BUS(g->storage_sub[k], b); w->pd[i] += sto_p[k]; w->qd[i] += g->storage_q[k]; } }
    for (int k = 0; k < nsh; ++k) { int b = tv[dt + k]; if (b > 0) { int i = BUS(g->shunt_sub[k], b); w->gs[i] += sh_p[k] * g->shunt_vratio[k]; w->bs[i] -= sh_q[k] * g->shunt_vratio[k]; } }
    for (int l = 0; l < nl; ++l) {
        int bo = tv[g->line_or_pos[l]], be = tv[g->line_ex_pos[l]];
        if (bo > 0 && be > 0) { w->brf[l] = BUS(g->line_or_sub[l], bo); w->brt[l] = BUS(g->line_ex_sub[l], be); }
        else { w->brf[l] = -1; w->brt[l] = -1; }
    }
    int anyref = 0;
    for (int i = 0; i < nb; ++i) if (w->btype[i] == 3) { w->reach[i] = 1; anyref = 1; }
    if (!anyref) return B200PF_ST_NO_REF;
    for (int sweep = 0; sweep < nb; ++sweep) {
        int ch = 0;
        for (int l = 0; l < nl; ++l) {
            int f = w->brf[l], t = w->brt[l];
            if (f < 0) continue;
            if (w->reach[f] != w->reach[t]) { w->reach[f] = 1; w->reach[t] = 1; ch = 1; }
        }
        if (!ch) break;
    }
    for (int i = 0; i < nb; ++i) if (!w->reach[i]) return B200PF_ST_UNSUPPLIED;
    /* numbering */
    int n1 = 0;
    for (int i = 0; i < nb; ++i) w->colth[i] = (w->btype[i] != 3) ? n1++ : -1;
    int d = n1;
    for (int i = 0; i < nb; ++i) w->colv[i] = (w->btype[i] == 1) ? d++ : -1;
    /* DC angles */
    {
        double *B = w->B, *rhs = w->rhs;
        memset(B, 0, sizeof(double) * (size_t)n1 * n1);
        for (int i = 0; i < nb; ++i) if (w->colth[i] >= 0) rhs[w->colth[i]] = (w->pg[i] - w->pd[i]) / base - w->gs[i] / base;
        for (int l = 0; l < nl; ++l) {
            int f = w->brf[l], t = w->brt[l];
            if (f < 0) continue;
            double b = g->line_bdc[l], pf = g->line_pshift[l];
            int cf = w->colth[f], ct = w->colth[t];
            if (cf >= 0) { B[IDX(cf, cf, n1)] += b; if (ct >= 0) B[IDX(cf, ct, n1)] -= b; rhs[cf] -= pf; }
            if (ct >= 0) { B[IDX(ct, ct, n1)] += b; if (cf >= 0) B[IDX(ct, cf, n1)] -= b; rhs[ct] += pf; }
        }
        if (lu_solve(B, rhs, n1)) return B200PF_ST_DIVERGED;
        for (int i = 0; i < nb; ++i) if (w->colth[i] >= 0) { w->va[i] = rhs[w->colth[i]]; if (!isfinite(w->va[i])) return B200PF_ST_DIVERGED; }
    }
    double *Yr = w->Yr, *Yi = w->Yi;
    if (!is_dc) {
        memset(Yr, 0, sizeof(double) * (size_t)nb * nb); memset(Yi, 0, sizeof(double) * (size_t)nb * nb);
        for (int l = 0; l < nl; ++l) {
            int f = w->brf[l], t = w->brt[l];
            if (f < 0) continue;
            const double *y = g->line_y + (size_t)l * 8;
            Yr[IDX(f, f, nb)] += y[0]; Yi[IDX(f, f, nb)] += y[1]; Yr[IDX(f, t, nb)] += y[2]; Yi[IDX(f, t, nb)] += y[3];
            Yr[IDX(t, f, nb)] += y[4]; Yi[IDX(t, f, nb)] += y[5]; Yr[IDX(t, t, nb)] += y[6]; Yi[IDX(t, t, nb)] += y[7];
        }
        for (int i = 0; i < nb; ++i) { Yr[IDX(i, i, nb)] += w->gs[i] / base; Yi[IDX(i, i, nb)] += w->bs[i] / base; }
        int conv = 0, it = 0;
        for (;; ++it) {
            double fmx = 0.0;
            for (int i = 0; i < nb; ++i) {            /* S = V conj(Y V) */
                double ir = 0, ii = 0;
                for (int j = 0; j < nb; ++j) {
                    double yr = Yr[IDX(i, j, nb)], yi = Yi[IDX(i, j, nb)];
                    if (yr == 0.0 && yi == 0.0) continue;
                    double ej = w->vm[j] * cos(w->va[j]), fj = w->vm[j] * sin(w->va[j]);
                    ir += yr * ej - yi * fj; ii += yr * fj + yi * ej;
                }
                double ei = w->vm[i] * cos(w->va[i]), fi = w->vm[i] * sin(w->va[i]);
                w->pcalc[i] = ei * ir + fi * ii; w->qcalc[i] = fi * ir - ei * ii;
                if (w->btype[i] != 3) { double m = fabs(w->pcalc[i] - (w->pg[i] - w->pd[i]) / base); if (!(m == m)) m = 1e300; if (m > fmx) fmx = m; }
                if (w->btype[i] == 1) { double m = fabs(w->qcalc[i] + w->qd[i] / base); if (!(m == m)) m = 1e300; if (m > fmx) fmx = m; }
            }
            if (fmx < tol_pu) { conv = 1; break; }
            if (it >= max_iter || !(fmx < 1e200)) break;
            double *J = w->J, *rhs = w->rhs;
            memset(J, 0, sizeof(double) * (size_t)d * d);
            for (int i = 0; i < nb; ++i) {
                int rp = w->colth[i], rq = w->colv[i];
                if (rp < 0) continue;
                double vi = w->vm[i], P = w->pcalc[i], Q = w->qcalc[i];
                double gii = Yr[IDX(i, i, nb)], bii = Yi[IDX(i, i, nb)];
                for (int j = 0; j < nb; ++j) {
                    if (j == i) continue;
                    double yr = Yr[IDX(i, j, nb)], yi = Yi[IDX(i, j, nb)];
                    if (yr == 0.0 && yi == 0.0) continue;
                    double th = w->va[i] - w->va[j], c = cos(th), s = sin(th), vj = w->vm[j];
                    double a = yr * c + yi * s, bq = yr * s - yi * c;     /* G cos + B sin ; G sin - B cos */
                    int cth = w->colth[j], cv = w->colv[j];
                    if (cth >= 0) { J[IDX(rp, cth, d)] = vi * vj * bq; if (rq >= 0) J[IDX(rq, cth, d)] = -vi * vj * a; }
                    if (cv >= 0) { J[IDX(rp, cv, d)] = vi * a; if (rq >= 0) J[IDX(rq, cv, d)] = vi * bq; }
                }
                J[IDX(rp, rp, d)] = -Q - bii * vi * vi;
                if (rq >= 0) {
                    J[IDX(rp, rq, d)] = P / vi + gii * vi;
                    J[IDX(rq, rp, d)] = P - gii * vi * vi;
                    J[IDX(rq, rq, d)] = Q / vi - bii * vi;
                }
                rhs[rp] = -(P - (w->pg[i] - w->pd[i]) / base);
                if (rq >= 0) rhs[rq] = -(Q + w->qd[i] / base);
            }
            if (lu_solve(J, rhs, d)) { ++it; break; }
            for (int i = 0; i < nb; ++i) {
                if (w->colth[i] >= 0) w->va[i] += rhs[w->colth[i]];
                if (w->colv[i] >= 0) w->vm[i] += rhs[w->colv[i]];
                if (w->vm[i] < 0) { w->vm[i] = -w->vm[i]; w->va[i] += M_PI; }
            }
        }
        *iters_out = it;
        if (!conv) return B200PF_ST_DIVERGED;
    } else {
        for (int i = 0; i < nb; ++i) { w->pcalc[i] = w->gs[i] / base; w->qcalc[i] = 0; }
        for (int l = 0; l < nl; ++l) {
            int f = w->brf[l], t = w->brt[l];
            if (f < 0) continue;
            double pf = g->line_bdc[l] * (w->va[f] - w->va[t]) + g->line_pshift[l];
            w->pcalc[f] += pf; w->pcalc[t] -= pf;
        }
    }
    /* results (same float32 rounding rules as the reference's read-back, pPB:1159-1183) */
    const double R2D = 57.295779513082320877, S3 = 1.7320508075688772935;
    /* pandapower reports Va = angle(V) after the Newton solve: angles in (-180, 180] (pandapower newtonpf / pfsoln) */
    if (!is_dc) for (int i = 0; i < nb; ++i) w->va[i] = w->va[i] - 6.283185307179586477 * rint(w->va[i] * 0.15915494309189533577);
    float *o = out;
    for (int l = 0; l < nl; ++l) {
        int f = w->brf[l], t = w->brt[l];
        float v[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (f >= 0) {
            double pf, qf, pt, qt, sf, st;
            if (is_dc) { pf = (g->line_bdc[l] * (w->va[f] - w->va[t]) + g->line_pshift[l]) * base; pt = -pf; qf = qt = 0; sf = fabs(pf); st = fabs(pt); }
            else {
                const double *y = g->line_y + (size_t)l * 8;
                double ef = w->vm[f] * cos(w->va[f]), ff = w->vm[f] * sin(w->va[f]), et = w->vm[t] * cos(w->va[t]), ft = w->vm[t] * sin(w->va[t]);
                double ifr = y[0] * ef - y[1] * ff + y[2] * et - y[3] * ft, ifi = y[0] * ff + y[1] * ef + y[2] * ft + y[3] * et;
                double itr = y[4] * ef - y[5] * ff + y[6] * et - y[7] * ft, iti = y[4] * ff + y[5] * ef + y[6] * ft + y[7] * et;
                pf = (ef * ifr + ff * ifi) * base; qf = (ff * ifr - ef * ifi) * base; pt = (et * itr + ft * iti) * base; qt = (ft * itr - et * iti) * base;
                sf = sqrt(pf * pf + qf * qf); st = sqrt(pt * pt + qt * qt);
            }
            float a1 = (float)(sf / (S3 * (w->vm[f] * (double)g->line_or_vn[l])) * 1000.0), a2 = (float)(st / (S3 * (w->vm[t] * (double)g->line_ex_vn[l])) * 1000.0);
            if (!isfinite(a1)) a1 = 0; if (!isfinite(a2)) a2 = 0;
            v[0] = (float)pf; v[1] = (float)qf; v[2] = (float)w->vm[f] * g->line_or_vn[l]; v[3] = a1; v[4] = (float)(w->va[f] * R2D);
            v[5] = (float)pt; v[6] = (float)qt; v[7] = (float)w->vm[t] * g->line_ex_vn[l]; v[8] = a2; v[9] = (float)(w->va[t] * R2D);
        }
        for (int k = 0; k < 10; ++k) o[k * nl + l] = v[k];
    }
    o += 10 * nl;
    for (int u = 0; u < nu; ++u) {
        int b = tv[g->unit_pos[u]];
        float p = 0, q = 0, v = 0, th = 0;
        if (b > 0) {
            int i = BUS(g->unit_sub[u], b);
            double pu = u >= nh ? gen_p[u - nh] : 0.0;
            if (g->unit_is_ref[u]) pu = (w->pcalc[i] * base + w->pd[i] - w->pnonref[i]) / (double)w->nref[i];
            double qu = 0;
            if (!is_dc) {
                double qtot = w->qcalc[i] * base + w->qd[i];
                if (w->cnt[i] <= 1 || w->qmins[i] == w->qmaxs[i]) qu = qtot / w->cnt[i];
                else qu = g->unit_qmin[u] + (qtot - w->qmins[i]) / (w->qmaxs[i] - w->qmins[i] + 2.220446049250313e-16) * (g->unit_qmax[u] - g->unit_qmin[u]);
            }
            p = (float)pu; q = (float)qu; v = (float)w->vm[i] * g->unit_vn[u]; th = (float)(w->va[i] * R2D);
        }
        o[u] = p; o[nu + u] = q; o[2 * nu + u] = v; o[3 * nu + u] = th;
    }
    o += 4 * nu;
    for (int k = 0; k < nld; ++k) { int b = tv[g->load_pos[k]]; float v = 0, th = 0; if (b > 0) { int i = BUS(g->load_sub[k], b); v = (float)w->vm[i] * g->load_vn[k]; th = (float)(w->va[i] * R2D); } o[k] = v; o[nld + k] = th; }
    o += 2 * nld;
    for (int k = 0; k < nst; ++k) { int b = tv[g->storage_pos[k]]; o[k] = b > 0 ? (float)w->vm[BUS(g->storage_sub[k], b)] * g->storage_vn[k] : 0.0f; }
    o += nst;
    for (int k = 0; k < nsh; ++k) {
        int b = tv[dt + k]; float p = 0, q = 0, v = 0;
        if (b > 0) { int i = BUS(g->shunt_sub[k], b); double v2 = is_dc ? 1.0 : w->vm[i] * w->vm[i]; p = (float)(sh_p[k] * g->shunt_vratio[k] * v2); q = is_dc ? 0.0f : (float)(sh_q[k] * g->shunt_vratio[k] * v2); v = (float)w->vm[i] * g->shunt_vn[k]; }
        o[k] = p; o[nsh + k] = q; o[2 * nsh + k] = v;
    }
    return B200PF_ST_CONVERGED;
}

/* batch entry point: same records as b200pf_run_host (include/b200pf.h) */
int pf_oracle_run(const b200pf_grid_desc *g, int batch, const int8_t *topo, const double *inj, int is_dc, int max_iter,
                  double tol_mva, float *out, int32_t *status, int32_t *iters, int nthreads) {
    const int nu = g->n_gen + g->n_hidden;
    const int n_topo_in = g->dim_topo + g->n_shunt + g->n_hidden;
    const int n_inj = g->n_gen + nu + 2 * g->n_load + g->n_storage + 2 * g->n_shunt;
    const int n_out = 10 * g->n_line + 4 * nu + 2 * g->n_load + g->n_storage + 3 * g->n_shunt;
    const double tol_pu = tol_mva / g->sn_mva;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
#endif
    {
        work_t *w = work_new(g);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int b = 0; b < batch; ++b) {
            int it = 0;
            status[b] = solve_one(g, w, topo + (size_t)b * n_topo_in, inj + (size_t)b * n_inj, is_dc, max_iter, tol_pu,
                                  out + (size_t)b * n_out, &it);
            iters[b] = it;
        }
        work_free(w);
    }
    return 0;
}

int pf_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
