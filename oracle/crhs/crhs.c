/*
 * TEST ORACLE / CPU BASELINE -- C11 + OpenMP restatement of the reference's
 * CPU design for the flux-reconstruction RHS.  Test infrastructure: used by
 * tests/ (second opinion on the NumPy oracle) and by bench.py's
 * cpu_baseline / --impl reference legs only; never by the product path.
 *
 * The reference's OpenMP backend cannot run offline (its kernels are Mako
 * templates, its only mat-mul provider is libxsmm).  This file follows its
 * design instead:
 *   - blocked AoSoA storage, one block = csubsz elements
 *     (pyfr/backends/base/types.py:56-88, pyfr/backends/openmp/base.py:14-26);
 *   - element kernels are "block kernels" run block by block, the kernels
 *     of a fusion group back to back on one block with the group's private
 *     temporaries in thread-local scratch
 *     (pyfr/backends/openmp/kernels/run-kernels.mako:30-70,
 *      pyfr/backends/openmp/types.py:136-186), schedule(static);
 *   - pointwise loop nest block -> row -> SoA chunk -> simd lanes
 *     (pyfr/backends/openmp/generator.py:105-158);
 *   - operator multiplies as sparse (CSR) x dense-block kernels, the
 *     stand-in for libxsmm_fsspmdm_execute
 *     (pyfr/backends/openmp/kernels/batch-gemm.mako:14-19);
 *   - interface kernels as flat parallel loops over interface points with
 *     the view dereference rules of pyfr/backends/base/generator.py:171-226.
 * Kernel arithmetic restates, point for point:
 *   pyfr/solvers/euler/kernels/flux.mako:3-26,
 *   pyfr/solvers/euler/kernels/rsolvers/{rusanov,hllc}.mako,
 *   pyfr/solvers/navstokes/kernels/flux.mako:3-105,
 *   pyfr/solvers/navstokes/kernels/{tflux,intconu,mpiconu,intcflux,mpicflux}.mako,
 *   pyfr/solvers/euler/kernels/{tflux,intcflux,mpicflux}.mako,
 *   pyfr/solvers/baseadvec/kernels/{negdivconf,smats}.mako,
 *   pyfr/solvers/baseadvecdiff/kernels/{gradcoru,transform_grad}.mako.
 * Build: see oracle/Makefile (reference flags: -O3 -fopenmp -march=native
 * -ffast-math, pyfr/backends/openmp/compiler.py:76-96; a second build
 * without -ffast-math serves the parity tests).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD 3
#define MAXV 5
#define MAXVERT 8

enum { K_MUL = 0, K_TFLUX = 1, K_GRADCORU = 2, K_NEGDIVCONF = 3 };
enum { RS_RUSANOV = 0, RS_HLLC = 1 };

typedef struct
{
    int kind, ndims, nvars, npts, ld, ksoa, csub, viscous, nverts, neles;
    /* matrix arguments: pointer, block stride (elements), scratch slot */
    double *p[4];
    long bs[4];
    int slot[4];
    long off[4];        /* element offset inside the scratch slot */
    /* K_MUL: out[M x ld] = A @ b + beta*out, A in CSR */
    int M;
    const int *rowptr, *cols;
    const double *vals;
    double beta;
    /* physics */
    double gamma, mu, gamma_pr;
    /* reference coordinates of the points, [npts][ndims] */
    const double *pts;
} kdesc;

/* ---- physics ------------------------------------------------------------ */
static inline __attribute__((always_inline)) void
inviscid_flux(int nd, int nv, double gamma, const double *s,
              double f[MAXD][MAXV], double *p, double *v)
{
    double invrho = 1.0/s[0], E = s[nv - 1], rhov[MAXD], ke = 0;

    for (int i = 0; i < nd; i++)
    {
        rhov[i] = s[i + 1];
        v[i] = invrho*rhov[i];
        ke += rhov[i]*rhov[i];
    }

    *p = (gamma - 1)*(E - 0.5*invrho*ke);

    for (int i = 0; i < nd; i++)
    {
        f[i][0] = rhov[i];
        f[i][nv - 1] = (E + *p)*v[i];

        for (int j = 0; j < nd; j++)
            f[i][j + 1] = rhov[i]*v[j] + (i == j ? *p : 0.0);
    }
}

static inline __attribute__((always_inline)) void
viscous_flux_add(int nd, int nv, double gamma, double mu, double gamma_pr,
                 const double *u, double g[MAXD][MAXV], double f[MAXD][MAXV])
{
    double rho = u[0], E = u[nv - 1], rcprho = 1.0/rho;
    double vel[MAXD], dv[MAXD][MAXD], T_x[MAXD], t[MAXD][MAXD], div = 0;

    for (int i = 0; i < nd; i++)
        vel[i] = rcprho*u[i + 1];

    /* dv[i][d] = rho * d(v_i)/d(x_d) */
    for (int i = 0; i < nd; i++)
        for (int d = 0; d < nd; d++)
            dv[i][d] = g[d][i + 1] - vel[i]*g[d][0];

    for (int d = 0; d < nd; d++)
    {
        double acc = rcprho*g[d][0]*E;
        for (int i = 0; i < nd; i++)
            acc += vel[i]*dv[i][d];
        T_x[d] = rcprho*(g[d][nv - 1] - acc);
    }

    for (int i = 0; i < nd; i++)
        div += dv[i][i];

    for (int i = 0; i < nd; i++)
    {
        t[i][i] = -2*mu*rcprho*(dv[i][i] - (1.0/3.0)*div);
        for (int j = i + 1; j < nd; j++)
            t[i][j] = t[j][i] = -mu*rcprho*(dv[j][i] + dv[i][j]);
    }

    for (int d = 0; d < nd; d++)
    {
        double e = 0;
        for (int i = 0; i < nd; i++)
        {
            f[d][i + 1] += t[d][i];
            e += vel[i]*t[d][i];
        }
        f[d][nv - 1] += e + -mu*gamma_pr*T_x[d];
    }
}

static inline __attribute__((always_inline)) void
rsolve_rusanov(int nd, int nv, double gamma, const double *ul,
               const double *ur, const double *n, double *nf)
{
    double fl[MAXD][MAXV], fr[MAXD][MAXV], vl[MAXD], vr[MAXD], pl, pr, nvs = 0;

    inviscid_flux(nd, nv, gamma, ul, fl, &pl, vl);
    inviscid_flux(nd, nv, gamma, ur, fr, &pr, vr);

    for (int i = 0; i < nd; i++)
        nvs += n[i]*(vl[i] + vr[i]);

    double a = sqrt((0.25*gamma)*(pl + pr)/(ul[0] + ur[0])) + 0.25*fabs(nvs);

    for (int i = 0; i < nv; i++)
    {
        double acc = 0;
        for (int j = 0; j < nd; j++)
            acc += n[j]*(fl[j][i] + fr[j][i]);
        nf[i] = 0.5*acc + a*(ul[i] - ur[i]);
    }
}

static inline __attribute__((always_inline)) void
rsolve_hllc(int nd, int nv, double gamma, const double *ul, const double *ur,
            const double *n, double *nf)
{
    double fl[MAXD][MAXV], fr[MAXD][MAXV], vl[MAXD], vr[MAXD], va[MAXD];
    double usl[MAXV], usr[MAXV], pl, pr, nvl = 0, nvr = 0, qq = 0;

    inviscid_flux(nd, nv, gamma, ul, fl, &pl, vl);
    inviscid_flux(nd, nv, gamma, ur, fr, &pr, vr);

    for (int i = 0; i < nd; i++)
    {
        nvl += n[i]*vl[i];
        nvr += n[i]*vr[i];
    }

    double al = sqrt(gamma*pl/ul[0]), ar = sqrt(gamma*pr/ur[0]);
    double srl = sqrt(ul[0]), srr = sqrt(ur[0]);
    double nva = (srl*nvl + srr*nvr)/(srl + srr);
    double H = (srl*(pr + ur[nd + 1]) + srr*(pl + ul[nd + 1]))
             / (srl*ur[0] + srr*ul[0]);
    double inv_rar = 1/(srl + srr);

    for (int i = 0; i < nd; i++)
    {
        va[i] = (vl[i]*srl + vr[i]*srr)*inv_rar;
        qq += va[i]*va[i];
    }

    double a = sqrt((gamma - 1)*(H - 0.5*qq));
    double sl = fmin(nva - a, nvl - al), sr = fmax(nva + a, nvr + ar);
    double sstar = (pr - pl + ul[0]*nvl*(sl - nvl) - ur[0]*nvr*(sr - nvr))
                 / (ul[0]*(sl - nvl) - ur[0]*(sr - nvr));
    double ul_com = (sl - nvl)/(sl - sstar), ur_com = (sr - nvr)/(sr - sstar);

    usl[0] = ul_com*ul[0];
    usr[0] = ur_com*ur[0];
    for (int i = 0; i < nd; i++)
    {
        usl[i + 1] = usl[0]*(vl[i] + (sstar - nvl)*n[i]);
        usr[i + 1] = usr[0]*(vr[i] + (sstar - nvr)*n[i]);
    }
    usl[nv - 1] = ul_com*(ul[nv - 1] + (sstar - nvl)*(ul[0]*sstar + pl/(sl - nvl)));
    usr[nv - 1] = ur_com*(ur[nv - 1] + (sstar - nvr)*(ur[0]*sstar + pr/(sr - nvr)));

    for (int i = 0; i < nv; i++)
    {
        double nf_fl = 0, nf_fr = 0;
        for (int j = 0; j < nd; j++)
        {
            nf_fl += n[j]*fl[j][i];
            nf_fr += n[j]*fr[j][i];
        }
        double nf_fsl = nf_fl + sl*(usl[i] - ul[i]);
        double nf_fsr = nf_fr + sr*(usr[i] - ur[i]);

        nf[i] = (0 <= sl) ? nf_fl : (sl <= 0 && 0 <= sstar) ? nf_fsl :
                (sstar <= 0 && 0 <= sr) ? nf_fsr : nf_fr;
    }
}

/* Metric terms of a multilinear (quad/hex) element from its vertices
 * (vertex n has sign bit e of n along axis e, first axis fastest). */
static inline __attribute__((always_inline)) void
calc_smats_detj(int nd, int nverts, double V[MAXVERT][MAXD], const double *x,
                double s[MAXD][MAXD], double *djac)
{
    double j[MAXD][MAXD];

    for (int d = 0; d < nd; d++)
        for (int i = 0; i < nd; i++)
        {
            double acc = 0;
            for (int n = 0; n < nverts; n++)
            {
                double w = 1;
                for (int e = 0; e < nd; e++)
                {
                    double sg = ((n >> e) & 1) ? 1.0 : -1.0;
                    w *= (e == d) ? sg : (1 + sg*x[e]);
                }
                acc += w*V[n][i];
            }
            j[d][i] = acc/nverts;
        }

    if (nd == 2)
    {
        s[0][0] = j[1][1]; s[0][1] = -j[1][0];
        s[1][0] = -j[0][1]; s[1][1] = j[0][0];
        *djac = s[0][0]*s[1][1] - s[0][1]*s[1][0];
    }
    else
    {
        static const int ab[3][2] = {{1, 2}, {2, 0}, {0, 1}};
        for (int i = 0; i < 3; i++)
        {
            int a = ab[i][0], b = ab[i][1];
            s[i][0] = j[a][1]*j[b][2] - j[a][2]*j[b][1];
            s[i][1] = j[a][2]*j[b][0] - j[a][0]*j[b][2];
            s[i][2] = j[a][0]*j[b][1] - j[a][1]*j[b][0];
        }
        *djac = j[0][0]*s[0][0] + j[0][1]*s[0][1] + j[0][2]*s[0][2];
    }
}

/* ---- block kernels -------------------------------------------------------- */
#define COFF(e, v, nv, k) (((e)/(k))*((k)*(nv)) + (v)*(k) + (e) % (k))

static void
blk_mul(const kdesc *k, double *const *m, int nvalid)
{
    const double *restrict b = m[0];
    double *restrict out = m[1];
    const int ld = k->ld;

    for (int r = 0; r < k->M; r++)
    {
        double *restrict o = out + (long) r*ld;

        if (k->beta == 0)
            for (int c = 0; c < ld; c++)
                o[c] = 0;
        else if (k->beta != 1)
            for (int c = 0; c < ld; c++)
                o[c] *= k->beta;

        for (int q = k->rowptr[r]; q < k->rowptr[r + 1]; q++)
        {
            const double a = k->vals[q];
            const double *restrict x = b + (long) k->cols[q]*ld;

            #pragma omp simd
            for (int c = 0; c < ld; c++)
                o[c] += a*x[c];
        }
    }
}

static inline __attribute__((always_inline)) void
load_verts(const kdesc *k, const int nd, const double *verts, int e,
           double V[MAXVERT][MAXD])
{
    for (int n = 0; n < (1 << nd); n++)
        for (int i = 0; i < nd; i++)
            V[n][i] = verts[(long) n*nd*k->csub + COFF(e, i, nd, k->ksoa)];
}

/* tflux (linear elements): f holds the physical gradient on entry for the
 * viscous system (the un-fused form the reference uses when blocks = True,
 * pyfr/solvers/navstokes/elements.py:64-128) and the transformed flux on
 * exit */
static inline __attribute__((always_inline)) void
blk_tflux_t(const kdesc *k, double *const *m, int nvalid, const int nd,
            const int nv)
{
    const int ld = k->ld, np = k->npts;
    const double *restrict u = m[0], *restrict verts = m[2];
    double *restrict f = m[1];

    for (int p = 0; p < np; p++)
        #pragma omp simd
        for (int e = 0; e < nvalid; e++)
        {
            double V[MAXVERT][MAXD], s[MAXD][MAXD], djac, us[MAXV];
            double ft[MAXD][MAXV], g[MAXD][MAXV], pr, vel[MAXD];

            load_verts(k, nd, verts, e, V);
            calc_smats_detj(nd, 1 << nd, V, k->pts + p*nd, s, &djac);

            for (int v = 0; v < nv; v++)
                us[v] = u[(long) p*ld + COFF(e, v, nv, k->ksoa)];

            inviscid_flux(nd, nv, k->gamma, us, ft, &pr, vel);

            if (k->viscous)
            {
                for (int d = 0; d < nd; d++)
                    for (int v = 0; v < nv; v++)
                        g[d][v] = f[((long) d*np + p)*ld
                                    + COFF(e, v, nv, k->ksoa)];
                viscous_flux_add(nd, nv, k->gamma, k->mu, k->gamma_pr, us, g,
                                 ft);
            }

            for (int i = 0; i < nd; i++)
                for (int v = 0; v < nv; v++)
                {
                    double acc = 0;
                    for (int j = 0; j < nd; j++)
                        acc += s[i][j]*ft[j][v];
                    f[((long) i*np + p)*ld + COFF(e, v, nv, k->ksoa)] = acc;
                }
        }
}

static void
blk_tflux(const kdesc *k, double *const *m, int nvalid)
{
    if (k->ndims == 3)
        blk_tflux_t(k, m, nvalid, 3, 5);
    else
        blk_tflux_t(k, m, nvalid, 2, 4);
}

static inline __attribute__((always_inline)) void
blk_gradcoru_t(const kdesc *k, double *const *m, int nvalid, const int nd,
               const int nv)
{
    const int ld = k->ld, np = k->npts;
    double *restrict gr = m[0];
    const double *restrict verts = m[1];

    for (int p = 0; p < np; p++)
        #pragma omp simd
        for (int e = 0; e < nvalid; e++)
        {
            double V[MAXVERT][MAXD], s[MAXD][MAXD], djac;

            load_verts(k, nd, verts, e, V);
            calc_smats_detj(nd, 1 << nd, V, k->pts + p*nd, s, &djac);
            const double rcpdjac = 1.0/djac;

            for (int v = 0; v < nv; v++)
            {
                double t[MAXD];
                for (int d = 0; d < nd; d++)
                    t[d] = gr[((long) d*np + p)*ld + COFF(e, v, nv, k->ksoa)];

                for (int i = 0; i < nd; i++)
                {
                    double acc = 0;
                    for (int d = 0; d < nd; d++)
                        acc += s[d][i]*t[d];
                    gr[((long) i*np + p)*ld + COFF(e, v, nv, k->ksoa)] =
                        rcpdjac*acc;
                }
            }
        }
}

static void
blk_gradcoru(const kdesc *k, double *const *m, int nvalid)
{
    if (k->ndims == 3)
        blk_gradcoru_t(k, m, nvalid, 3, 5);
    else
        blk_gradcoru_t(k, m, nvalid, 2, 4);
}

static void
blk_negdivconf(const kdesc *k, double *const *m, int nvalid)
{
    const int nv = k->nvars, ld = k->ld;
    double *t = m[0];
    const double *rj = m[1];

    for (int p = 0; p < k->npts; p++)
        for (int v = 0; v < nv; v++)
            #pragma omp simd
            for (int e = 0; e < nvalid; e++)
                t[(long) p*ld + COFF(e, v, nv, k->ksoa)] *=
                    -rj[(long) p*k->csub + e];
}

/* Runs the kernels of one fusion group block by block; arguments with a
 * scratch slot live in thread-local buffers (one block's worth). */
void
crhs_run_blocks(int nk, const kdesc *ks, int nblocks, int nslots,
                const long *slot_elems)
{
    #pragma omp parallel
    {
        double *tls[8] = {0};
        for (int s = 0; s < nslots; s++)
            tls[s] = aligned_alloc(64, ((slot_elems[s]*sizeof(double) + 63)/64)*64);

        #pragma omp for schedule(static)
        for (int b = 0; b < nblocks; b++)
            for (int i = 0; i < nk; i++)
            {
                const kdesc *k = &ks[i];
                double *m[4];
                int nvalid = k->neles - b*k->csub;
                nvalid = nvalid < k->csub ? nvalid : k->csub;

                for (int a = 0; a < 4; a++)
                    m[a] = k->slot[a] >= 0 ? tls[k->slot[a]] + k->off[a]
                         : k->p[a] ? k->p[a] + (long) b*k->bs[a] : 0;

                switch (k->kind)
                {
                case K_MUL: blk_mul(k, m, nvalid); break;
                case K_TFLUX: blk_tflux(k, m, nvalid); break;
                case K_GRADCORU: blk_gradcoru(k, m, nvalid); break;
                case K_NEGDIVCONF: blk_negdivconf(k, m, nvalid); break;
                }
            }

        for (int s = 0; s < nslots; s++)
            free(tls[s]);
    }
}

/* ---- interface kernels ------------------------------------------------------ */
typedef struct
{
    int ndims, nvars, ksoa, rsolver, viscous, mpi;
    double gamma, mu, gamma_pr, beta, tau;
    long n;
    double *base;                 /* storage root of the solution views */
    const double *gbase;          /* storage root of the gradient views */
    const int *ul_map, *ur_map;   /* scal views (ur_map unused if mpi) */
    const double *ur_mpi;         /* [nvars][n] when mpi */
    const int *gl_map, *gl_str, *gr_map, *gr_str;
    const double *gr_mpi;         /* [ndims*nvars][n] when mpi */
    const double *nl;             /* [ndims][nl_ld] */
    long nl_ld;
} cflux_args;

static inline __attribute__((always_inline)) void
cflux_t(const cflux_args *a, const int nd, const int nv, const int rs,
        const int viscous)
{
    const int k = a->ksoa;
    const long n = a->n;

    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++)
    {
        double l[MAXV], r[MAXV], nrm[MAXD], fn[MAXV], mag2 = 0;
        double *pl = a->base + a->ul_map[i], *pr = 0;

        for (int v = 0; v < nv; v++)
            l[v] = pl[k*v];

        if (a->mpi)
            for (int v = 0; v < nv; v++)
                r[v] = a->ur_mpi[(long) v*n + i];
        else
        {
            pr = a->base + a->ur_map[i];
            for (int v = 0; v < nv; v++)
                r[v] = pr[k*v];
        }

        for (int d = 0; d < nd; d++)
        {
            nrm[d] = a->nl[(long) d*a->nl_ld + i];
            mag2 += nrm[d]*nrm[d];
        }

        const double mag = sqrt(mag2), rcpmag = 1.0/mag;
        for (int d = 0; d < nd; d++)
            nrm[d] *= rcpmag;

        if (rs == RS_RUSANOV)
            rsolve_rusanov(nd, nv, a->gamma, l, r, nrm, fn);
        else
            rsolve_hllc(nd, nv, a->gamma, l, r, nrm, fn);

        if (viscous)
        {
            double g[MAXD][MAXV], fvl[MAXD][MAXV] = {{0}}, fvr[MAXD][MAXV] = {{0}};
            const int need_l = a->beta != -0.5, need_r = a->beta != 0.5;

            if (need_l)
            {
                const double *gp = a->gbase + a->gl_map[i];
                const long st = a->gl_str[i];
                for (int d = 0; d < nd; d++)
                    for (int v = 0; v < nv; v++)
                        g[d][v] = gp[st*d + k*v];
                viscous_flux_add(nd, nv, a->gamma, a->mu, a->gamma_pr, l, g,
                                 fvl);
            }
            if (need_r)
            {
                if (a->mpi)
                    for (int d = 0; d < nd; d++)
                        for (int v = 0; v < nv; v++)
                            g[d][v] = a->gr_mpi[(long) (nv*d + v)*n + i];
                else
                {
                    const double *gp = a->gbase + a->gr_map[i];
                    const long st = a->gr_str[i];
                    for (int d = 0; d < nd; d++)
                        for (int v = 0; v < nv; v++)
                            g[d][v] = gp[st*d + k*v];
                }
                viscous_flux_add(nd, nv, a->gamma, a->mu, a->gamma_pr, r, g,
                                 fvr);
            }

            for (int v = 0; v < nv; v++)
            {
                double fl = 0, fr = 0, fv;
                for (int j = 0; j < nd; j++)
                {
                    fl += nrm[j]*fvl[j][v];
                    fr += nrm[j]*fvr[j][v];
                }
                fv = a->beta == -0.5 ? fr : a->beta == 0.5 ? fl
                   : (0.5 + a->beta)*fl + (0.5 - a->beta)*fr;
                if (a->tau != 0.0)
                    fv += a->tau*(l[v] - r[v]);
                fn[v] += fv;
            }
        }

        for (int v = 0; v < nv; v++)
        {
            const double fc = mag*fn[v];
            pl[k*v] = fc;
            if (!a->mpi)
                pr[k*v] = -fc;
        }
    }
}

void
crhs_cflux(const cflux_args *a)
{
#define CF(nd, nv, rs, vi) \
    if (a->ndims == nd && a->rsolver == rs && a->viscous == vi) \
    { cflux_t(a, nd, nv, rs, vi); return; }
    CF(3, 5, RS_RUSANOV, 1) CF(3, 5, RS_HLLC, 1)
    CF(3, 5, RS_RUSANOV, 0) CF(3, 5, RS_HLLC, 0)
    CF(2, 4, RS_RUSANOV, 1) CF(2, 4, RS_HLLC, 1)
    CF(2, 4, RS_RUSANOV, 0) CF(2, 4, RS_HLLC, 0)
#undef CF
}

typedef struct
{
    int nvars, ksoa, mpi;
    double beta;
    long n;
    const double *base;           /* storage root of the input views */
    double *obase;                /* storage root of the output views */
    const int *li_map, *ri_map, *lo_map, *ro_map;
    const double *ri_mpi;
} conu_args;

void
crhs_conu(const conu_args *a)
{
    const int nv = a->nvars, k = a->ksoa;
    const long n = a->n;

    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++)
        for (int v = 0; v < nv; v++)
        {
            const double l = a->base[a->li_map[i] + k*v];
            const double r = a->mpi ? a->ri_mpi[(long) v*n + i]
                                    : a->base[a->ri_map[i] + k*v];

            if (a->mpi)
                a->obase[a->lo_map[i] + k*v] =
                    a->beta == -0.5 ? l : a->beta == 0.5 ? r
                    : r*(0.5 + a->beta) + l*(0.5 - a->beta);
            else if (a->beta == -0.5)
                a->obase[a->ro_map[i] + k*v] = l;
            else if (a->beta == 0.5)
                a->obase[a->lo_map[i] + k*v] = r;
            else
            {
                const double com = r*(0.5 + a->beta) + l*(0.5 - a->beta);
                a->obase[a->lo_map[i] + k*v] = com;
                a->obase[a->ro_map[i] + k*v] = com;
            }
        }
}

void
crhs_pack(long n, int nrv, int ncv, int ksoa, const double *base,
          const int *map, const int *str, double *pmat)
{
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++)
        for (int r = 0; r < nrv; r++)
            for (int c = 0; c < ncv; c++)
                pmat[(long) (r*ncv + c)*n + i] =
                    base[map[i] + (nrv > 1 ? (long) str[i]*r : 0) + ksoa*c];
}

void
crhs_copy_rows(int nblocks, long nelem, double *dst, long dbs,
               const double *src, long sbs)
{
    #pragma omp parallel for schedule(static)
    for (int b = 0; b < nblocks; b++)
        memcpy(dst + b*dbs, src + b*sbs, nelem*sizeof(double));
}

void
crhs_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#endif
}

int
crhs_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
