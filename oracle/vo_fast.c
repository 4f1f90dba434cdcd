/* ORACLE -- TEST INFRASTRUCTURE ONLY (included by valence_oracle.c; never part of the product path).
 *
 * FAST ORACLE.  The literal restatement above costs O(orbital quartets x n^3) (a Givens determinant per
 * spin term, every AO quartet recomputed per orbital quartet) and stops being usable beyond ~4 water
 * molecules.  This file evaluates THE SAME SUM -- same task list, same three screens, same counters --
 * in the inverse ("Loewdin") form of /root/reference/doc/notes-vsvb-energy.tex:560-745 and SURVEY.md
 * appendix B:
 *   - first/second-order cofactors of vsvb_energy's density() calls (valence.F90:1535-1600, 1895-2056)
 *     are read from the inverses of the alpha / beta overlap blocks:
 *         C1(i->k)       = det(M) Minv[k][i]
 *         C2(i->k, j->l) = det(M) (Minv[k][i] Minv[l][j] - Minv[l][i] Minv[k][j])
 *     (exact determinants: identical to the reference when its Givens skip threshold dtol is tight,
 *     ntol_d >= 16; singular blocks are refused);
 *   - the orbital-level integrals int2e(io,jo,ko,lo) (valence.F90:3184-3438) come from blocks over entry
 *     groups: every AO shell quartet is generated once (McMurchie-Davidson, own Hermite code below, exact
 *     Boys values from vo_boys tabulated + 9-term Taylor) and transformed with the screened LCAO weights
 *     (per-shell weight screen sum c^2 > dtol, valence.F90:3296-3348);
 *   - every integral value then visits the tasks of the reference's 2e loop in which it is the direct
 *     or the exchanged integral (valence.F90:1163-1433): Schwarz screen :1189-1190, Schwarz shortcut
 *     :1213-1216, value screen :1286-1287, 16-way spin loop :1333-1414, x2 for ij<->kl :1420-1428.
 * No code is shared with valence_b200/ (the product uses Obara-Saika recurrences on the GPU).
 * Validation: tests/test_oracle_fast.py compares this path with the literal one above on every input
 * the literal one finishes (energies to 1e-11, counters exactly).
 * Spin-coupled wavefunctions (npair > 0): one inverse pair per determinant pair, f_cofactors_sc.
 * Limits: l <= 2, non-singular overlap blocks; first_order matrices for npair == 0 only.
 */
#include <pthread.h>
#include <unistd.h>

/* ---- a small dynamic-schedule thread pool (pthreads: available with every gcc of this image) ---------- */
typedef void (*f_body)(void *ctx, long long i, int tid);
typedef struct { f_body fn; void *ctx; long long n, chunk; long long *next; int tid; } f_job;
static int f_nthreads = 0;
static int f_threads(void)
{
    if (f_nthreads > 0) return f_nthreads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (n > 64 ? 64 : (int)n);
}
static void *f_worker(void *arg)
{
    f_job *j = (f_job *)arg;
    for (;;) {
        long long i0 = __atomic_fetch_add(j->next, j->chunk, __ATOMIC_RELAXED);
        if (i0 >= j->n) break;
        long long i1 = i0 + j->chunk < j->n ? i0 + j->chunk : j->n;
        for (long long i = i0; i < i1; ++i) j->fn(j->ctx, i, j->tid);
    }
    return NULL;
}
static void f_parfor(long long n, long long chunk, f_body fn, void *ctx)
{
    int nt = f_threads();
    if (nt > n) nt = n < 1 ? 1 : (int)n;
    long long next = 0;
    f_job jobs[64]; pthread_t th[64];
    for (int t = 0; t < nt; ++t) { jobs[t].fn = fn; jobs[t].ctx = ctx; jobs[t].n = n; jobs[t].chunk = chunk; jobs[t].next = &next; jobs[t].tid = t; }
    for (int t = 1; t < nt; ++t) pthread_create(&th[t], NULL, f_worker, &jobs[t]);
    f_worker(&jobs[0]);
    for (int t = 1; t < nt; ++t) pthread_join(th[t], NULL);
}

#define FL_LP 4                 /* Hermite order of a shell pair, l <= 2 per shell */
#define FL_LT 8                 /* of a quartet */
#define FL_NH4 35               /* # Hermite functions up to order 4 */
#define FL_NH8 165              /* up to order 8 */
#define FB_STEP 0.1
#define FB_TMAX 46.0            /* beyond: F_0 = sqrt(pi/T)/2, exp(-T) < 1.1e-20 dropped */
#define FB_ROWS 462
#define FB_COLS (FL_LT + 10)

static const double F_PI = 3.14159265358979323846264338327950288;
static double fb_tab[FB_ROWS][FB_COLS];
static int fh_idx[FL_LT + 1][FL_LT + 1][FL_LT + 1];   /* (t,u,v) -> index, ordered by t+u+v */
static int fh_t[FL_NH8], fh_u[FL_NH8], fh_v[FL_NH8], fh_first[FL_LT + 2];
static int fh_par1[FL_NH8], fh_par2[FL_NH8], fh_dir[FL_NH8], fh_mul[FL_NH8];
static int fh_sum[FL_NH4][FL_NH4];                   /* index of h + h' */
static double fh_sgn[FL_NH4];                        /* (-1)^(t+u+v) */
static int f_tables_ready = 0;

static int f_nh(int L) { return (L + 1) * (L + 2) * (L + 3) / 6; }

static void f_init_tables(void)
{
    if (f_tables_ready) return;
    for (int k = 0; k < FB_ROWS; ++k) vo_boys(FB_COLS - 1, k * FB_STEP, fb_tab[k]);
    int n = 0;
    for (int N = 0; N <= FL_LT; ++N) {
        fh_first[N] = n;
        for (int t = N; t >= 0; --t)
            for (int u = N - t; u >= 0; --u) { int v = N - t - u; fh_idx[t][u][v] = n; fh_t[n] = t; fh_u[n] = u; fh_v[n] = v; ++n; }
    }
    fh_first[FL_LT + 1] = n;
    for (int h = 1; h < n; ++h) {       /* R_{tuv} from lower orders: first non-zero direction */
        int t = fh_t[h], u = fh_u[h], v = fh_v[h];
        if (t > 0) { fh_dir[h] = 0; fh_par1[h] = fh_idx[t - 1][u][v]; fh_mul[h] = t - 1; fh_par2[h] = t > 1 ? fh_idx[t - 2][u][v] : 0; }
        else if (u > 0) { fh_dir[h] = 1; fh_par1[h] = fh_idx[t][u - 1][v]; fh_mul[h] = u - 1; fh_par2[h] = u > 1 ? fh_idx[t][u - 2][v] : 0; }
        else { fh_dir[h] = 2; fh_par1[h] = fh_idx[t][u][v - 1]; fh_mul[h] = v - 1; fh_par2[h] = v > 1 ? fh_idx[t][u][v - 2] : 0; }
    }
    for (int h = 0; h < FL_NH4; ++h) {
        fh_sgn[h] = ((fh_t[h] + fh_u[h] + fh_v[h]) & 1) ? -1.0 : 1.0;
        for (int r = 0; r < FL_NH4; ++r) fh_sum[h][r] = fh_idx[fh_t[h] + fh_t[r]][fh_u[h] + fh_u[r]][fh_v[h] + fh_v[r]];
    }
    f_tables_ready = 1;
}

/* F_0..F_m(T): Taylor series about the nearest grid point for the top order (dF_m/dT = -F_{m+1}),
 * downward recursion below it; asymptotic form beyond the table */
static void f_boys(int m, double T, double *F)
{
    if (T >= FB_TMAX) {
        double r = 1.0 / T;
        F[0] = 0.5 * sqrt(F_PI * r);
        for (int k = 0; k < m; ++k) F[k + 1] = (2.0 * k + 1.0) * 0.5 * r * F[k];
        return;
    }
    int k = (int)(T * (1.0 / FB_STEP) + 0.5);
    double d = k * FB_STEP - T;
    const double *r = fb_tab[k] + m;
    double f = r[9] * (1.0 / 362880.0);
    f = f * d + r[8] * (1.0 / 40320.0);
    f = f * d + r[7] * (1.0 / 5040.0);
    f = f * d + r[6] * (1.0 / 720.0);
    f = f * d + r[5] * (1.0 / 120.0);
    f = f * d + r[4] * (1.0 / 24.0);
    f = f * d + r[3] * (1.0 / 6.0);
    f = f * d + r[2] * 0.5;
    f = f * d + r[1];
    F[m] = f * d + r[0];
    if (m > 0) {
        double eT = exp(-T), t2 = 2.0 * T;
        for (int j = m; j > 0; --j) F[j - 1] = (t2 * F[j] + eT) / (2.0 * j - 1.0);
    }
}
double vo_fast_boys(int m, double T) { double F[FL_LT + 2]; f_init_tables(); f_boys(m, T, F); return F[m]; }

/* Hermite Coulomb integrals R_{tuv}(alpha, PQ), t+u+v <= L, indexed by fh_idx */
static void f_hermite_R(int L, double alpha, const double *PQ, double *R)
{
    double F[FL_LT + 1];
    f_boys(L, alpha * (PQ[0] * PQ[0] + PQ[1] * PQ[1] + PQ[2] * PQ[2]), F);
    if (L == 0) { R[0] = F[0]; return; }
    double Rn[FL_LT + 1][FL_NH8];
    double f = 1.0;
    for (int n = 0; n <= L; ++n) { Rn[n][0] = f * F[n]; f *= -2.0 * alpha; }
    for (int N = 1; N <= L; ++N)
        for (int n = 0; n <= L - N; ++n)
            for (int h = fh_first[N]; h < fh_first[N + 1]; ++h) {
                double val = PQ[fh_dir[h]] * Rn[n + 1][fh_par1[h]];
                if (fh_mul[h]) val += fh_mul[h] * Rn[n + 1][fh_par2[h]];
                Rn[n][h] = val;
            }
    int nh = f_nh(L);
    for (int h = 0; h < nh; ++h) R[h] = Rn[0][h];
}

/* 1D Hermite expansion coefficients E[i][j][t], i <= imax, j <= jmax (no exponential factor) */
static void f_hermite_E(int imax, int jmax, double p, double PA, double PB, double E[5][5][10])
{
    double h = 0.5 / p;
    for (int i = 0; i <= imax; ++i) for (int j = 0; j <= jmax; ++j) for (int t = 0; t < 10; ++t) E[i][j][t] = 0.0;
    E[0][0][0] = 1.0;
    for (int i = 0; i <= imax; ++i) {
        if (i > 0)
            for (int t = 0; t <= i; ++t) {
                double v = PA * E[i - 1][0][t] + (t + 1) * E[i - 1][0][t + 1];
                if (t > 0) v += h * E[i - 1][0][t - 1];
                E[i][0][t] = v;
            }
        for (int j = 1; j <= jmax; ++j)
            for (int t = 0; t <= i + j; ++t) {
                double v = PB * E[i][j - 1][t] + (t + 1) * E[i][j - 1][t + 1];
                if (t > 0) v += h * E[i][j - 1][t - 1];
                E[i][j][t] = v;
            }
    }
}

/* ---- data model ------------------------------------------------------------------------------------ */
typedef struct { int l, nprim, ao_off, atom; const double *exps, *coef; double r[3]; } f_shell;   /* real-atom shells */
typedef struct { int nsh; int *gsh; double *cf; } f_orb;     /* expanded orbital: shells + weights x angn, 6 per shell */
typedef struct { double p, P[3], w; } f_pp;
typedef struct {
    int la, lb, nab, Lp, nh, npp;
    f_pp *pp; double *E;          /* E[(k * nab + ab) * nh + h], K_ab folded in; primitives sorted by w (descending) */
    double *D;                    /* pair densities D[ab * np + p] */
    double wmax;
} f_sp;
typedef struct { int g, h, np; int *ps, *pt; int nsp; f_sp *sp; double smax; } f_pg;
typedef struct { int n; int *e; int nshB, nshK; int *shB, *shK; double *cB, *cK; /* [entry][shell][6] */ } f_grp;

/* cofactor data of one determinant pair of a spin-coupled wavefunction: weight = c_isc c_jsc det(alpha block) det(beta block) */
struct f_dp_s { int *posa_b, *posa_k, *posb_b, *posb_k; double *Ai, *Bi; double wt; };

typedef struct {
    vo_ctx *c;
    int nsh, nao; f_shell *sh; int *atom_first;
    int norb; f_orb *o1, *o2;                    /* all shells (1e) / weight-screened (2e) */
    int nnd, nso, sym, jorb;
    int *eB, *eK;                                /* entry -> bra / ket orbital id (1-based) */
    double *Se, *He;                             /* entry level, nso x nso: <B(s)|K(t)>, <B(s)|h|K(t)> */
    double *S, *H;                               /* AO level */
    int na, nb; int *posa_b, *posa_k, *posb_b, *posb_k;   /* spin-orbital -> position in alpha / beta lists (1-based, 0 = absent) */
    double *Ai, *Bi; double c0, la_det, lb_det;  /* inverses (row = ket position, col = bra position), det product */
    int ndp; struct f_dp_s *dp;                  /* spin-coupled wavefunctions: one record per determinant pair (then c0 = 1) */
    int ngrp; f_grp *grp; int *grp_of;
    int npg; f_pg *pg; int *pg_of;               /* pg_of[g * ngrp + h] */
    double thr;
    int nkeep_shells_max;
    int *nshell_kept;                            /* per orbital id: shells passing the weight screen (count_only logic of int2e) */
    double **cache; long long ncache;            /* first_order: E blocks free of the subject entry */
    int subject_grp;
} f_ctx;

static int f_real_shell(const f_ctx *F, int ia, int ii)
{
    const vo_ctx *c = F->c;
    int ra = c->atom_alias[ia];
    return F->atom_first[ra] + (ii - c->map_atom2shell[c->atom_t[ra]]);
}

static void f_build_shells(f_ctx *F)
{
    vo_ctx *c = F->c;
    F->atom_first = IARR(c->natom + 1);
    int n = 0;
    for (int a = 1; a <= c->natom; ++a) { F->atom_first[a] = n; n += c->num_shell_atom[c->atom_t[a]]; }
    F->nsh = n;
    F->sh = (f_shell *)xcalloc((size_t)n, sizeof(f_shell));
    int ao = 0;
    for (int a = 1; a <= c->natom; ++a) {
        int it = c->atom_t[a];
        for (int k = 0; k < c->num_shell_atom[it]; ++k) {
            int ii = c->map_atom2shell[it] + k;
            f_shell *s = &F->sh[F->atom_first[a] + k];
            s->l = c->ang_mom[ii]; s->nprim = c->map_shell2prim[ii + 1] - c->map_shell2prim[ii];
            s->exps = c->exponent + c->map_shell2prim[ii]; s->coef = c->con_coeff + c->map_shell2prim[ii];
            for (int d = 1; d <= 3; ++d) s->r[d - 1] = COORDS(c, d, a);
            s->ao_off = ao; s->atom = a;
            if (s->l > 2) { fprintf(stderr, "oracle(fast): l > 2 not supported\n"); exit(2); }
            ao += shell_size(s->l);
        }
    }
    F->nao = ao;
}

/* scatter of one orbital over its OBS (the reference's coeffi, valence.F90:3221-3237) -> shells with weights x angn;
 * screened != 0 keeps only shells with sum c^2 > dtol (valence.F90:3296-3300) */
static void f_expand(f_ctx *F, int io, int screened, f_orb *out)
{
    vo_ctx *c = F->c;
    double *cf = DARR(c->max_obs);
    scatter(c, io, cf);
    int cap = 0;
    for (int ia = 1; ia <= c->orbas_atnum[io]; ++ia) cap += c->num_shell_atom[c->atom_t[ATSET(c, ia, io)]];
    out->gsh = IARR(cap); out->cf = DARR(6 * cap); out->nsh = 0;
    int beg = 1;
    for (int ia = 1; ia <= c->orbas_atnum[io]; ++ia) {
        int ic = ATSET(c, ia, io), it = c->atom_t[ic];
        for (int ii = c->map_atom2shell[it]; ii <= c->map_atom2shell[it] + c->num_shell_atom[it] - 1; ++ii) {
            int ns = shell_size(c->ang_mom[ii]);
            double sum = 0.0;
            for (int i = beg; i <= beg + ns - 1; ++i) sum = sum + cf[i] * cf[i];
            if (!screened || sum > c->dtol) {
                int k = out->nsh++;
                out->gsh[k] = f_real_shell(F, ic, ii);
                for (int a = 0; a < ns; ++a) out->cf[6 * k + a] = cf[beg + a] * c->angn[funmin[c->ang_mom[ii] + 1] + a];
            }
            beg += ns;
        }
    }
    free(cf);
}

/* ---- one-electron AO matrices (overlap, kinetic + nuclear attraction with -Z) ------------------------- */
typedef struct { f_ctx *F; int nnuc; double *nuc; } f_ao_job;
static void f_ao_1e_row(void *ctx, long long Xl, int tid)
{
    f_ao_job *J = (f_ao_job *)ctx;
    f_ctx *F = J->F;
    vo_ctx *c = F->c;
    const int nao = F->nao, nnuc = J->nnuc, X = (int)Xl;
    const double *nuc = J->nuc;
    (void)tid;
    {
        double Ex[3][5][5][10], R[FL_NH4];
        for (int Y = 0; Y <= X; ++Y) {
            const f_shell *A = &F->sh[X], *B = &F->sh[Y];
            int na = shell_size(A->l), nb = shell_size(B->l), L = A->l + B->l, nh = f_nh(L);
            double AB2 = 0.0;
            for (int d = 0; d < 3; ++d) AB2 += (A->r[d] - B->r[d]) * (A->r[d] - B->r[d]);
            double s[36], h[36], eh[36][FL_NH4];
            for (int n = 0; n < na * nb; ++n) { s[n] = 0.0; h[n] = 0.0; }
            for (int ia = 0; ia < A->nprim; ++ia)
                for (int ib = 0; ib < B->nprim; ++ib) {
                    double a = A->exps[ia], b = B->exps[ib], p = a + b, mu = a * b / p;
                    double K = A->coef[ia] * B->coef[ib] * exp(-mu * AB2);
                    if (K == 0.0) continue;
                    double P[3];
                    for (int d = 0; d < 3; ++d) { P[d] = (a * A->r[d] + b * B->r[d]) / p; f_hermite_E(A->l, B->l + 2, p, P[d] - A->r[d], P[d] - B->r[d], Ex[d]); }
                    double ps = K * pow(F_PI / p, 1.5);
                    for (int i = 0; i < na; ++i) {
                        int al[3] = {NXYZ(c, 1, funmin[A->l + 1] + i), NXYZ(c, 2, funmin[A->l + 1] + i), NXYZ(c, 3, funmin[A->l + 1] + i)};
                        for (int j = 0; j < nb; ++j) {
                            int bl[3] = {NXYZ(c, 1, funmin[B->l + 1] + j), NXYZ(c, 2, funmin[B->l + 1] + j), NXYZ(c, 3, funmin[B->l + 1] + j)};
                            double Sd[3], Td[3];
                            for (int d = 0; d < 3; ++d) {
                                int ii = al[d], jj = bl[d];
                                Sd[d] = Ex[d][ii][jj][0];
                                Td[d] = -2.0 * b * b * Ex[d][ii][jj + 2][0] + b * (2.0 * jj + 1.0) * Ex[d][ii][jj][0];
                                if (jj >= 2) Td[d] -= 0.5 * jj * (jj - 1) * Ex[d][ii][jj - 2][0];
                            }
                            s[i * nb + j] += ps * Sd[0] * Sd[1] * Sd[2];
                            h[i * nb + j] += ps * (Td[0] * Sd[1] * Sd[2] + Sd[0] * Td[1] * Sd[2] + Sd[0] * Sd[1] * Td[2]);
                            for (int q = 0; q < nh; ++q) {
                                int t = fh_t[q], u = fh_u[q], v = fh_v[q];
                                eh[i * nb + j][q] = (t <= al[0] + bl[0] && u <= al[1] + bl[1] && v <= al[2] + bl[2]) ? Ex[0][al[0]][bl[0]][t] * Ex[1][al[1]][bl[1]][u] * Ex[2][al[2]][bl[2]][v] : 0.0;
                            }
                        }
                    }
                    if (fabs(K) < 1e-40) continue;     /* nothing representable left for the attraction sum */
                    double pv = 2.0 * F_PI / p * K;
                    for (int n = 0; n < nnuc; ++n) {
                        double PC[3] = {P[0] - nuc[4 * n], P[1] - nuc[4 * n + 1], P[2] - nuc[4 * n + 2]};
                        f_hermite_R(L, p, PC, R);
                        double z = -nuc[4 * n + 3] * pv;
                        for (int ij = 0; ij < na * nb; ++ij) {
                            double acc = 0.0;
                            for (int q = 0; q < nh; ++q) acc += eh[ij][q] * R[q];
                            h[ij] += z * acc;
                        }
                    }
                }
            for (int i = 0; i < na; ++i)
                for (int j = 0; j < nb; ++j) {
                    size_t r = (size_t)(A->ao_off + i), q = (size_t)(B->ao_off + j);
                    F->S[r * nao + q] = s[i * nb + j]; F->S[q * nao + r] = s[i * nb + j];
                    F->H[r * nao + q] = h[i * nb + j]; F->H[q * nao + r] = h[i * nb + j];
                }
        }
    }
}
static void f_ao_1e(f_ctx *F)
{
    vo_ctx *c = F->c;
    const int nao = F->nao;
    F->S = (double *)xcalloc((size_t)nao * nao, sizeof(double));
    F->H = (double *)xcalloc((size_t)nao * nao, sizeof(double));
    f_ao_job J; J.F = F; J.nnuc = 0;
    J.nuc = DARR(4 * c->natom);
    for (int a = 1; a <= c->natom; ++a) {
        double z = c->nuc_charge[c->atom_t[a]];
        if (fabs(z) > 1.0e-12) { for (int d = 0; d < 3; ++d) J.nuc[4 * J.nnuc + d] = COORDS(c, d + 1, a); J.nuc[4 * J.nnuc + 3] = z; ++J.nnuc; }   /* valence.F90:3147-3149 */
    }
    f_parfor(F->nsh, 4, f_ao_1e_row, &J);
    free(J.nuc);
}

static void f_orb_1e(const f_ctx *F, const f_orb *A, const f_orb *B, double *s, double *h)
{
    double ss = 0.0, hh = 0.0;
    const int nao = F->nao;
    for (int x = 0; x < A->nsh; ++x) {
        const f_shell *X = &F->sh[A->gsh[x]];
        for (int y = 0; y < B->nsh; ++y) {
            const f_shell *Y = &F->sh[B->gsh[y]];
            for (int i = 0; i < shell_size(X->l); ++i)
                for (int j = 0; j < shell_size(Y->l); ++j) {
                    double d = A->cf[6 * x + i] * B->cf[6 * y + j];
                    size_t k = (size_t)(X->ao_off + i) * nao + (Y->ao_off + j);
                    ss += d * F->S[k]; hh += d * F->H[k];
                }
        }
    }
    *s = ss; *h = hh;
}

/* ---- dense inverse + determinant, M row-major n x n; returns 0 when singular ------------------------------
 * Gauss-Jordan with partial pivoting in extended precision (x87 long double, 64-bit mantissa), then one
 * Newton-Schulz step X <- X (2I - M X) with the residual formed in long double: the inverse handed back in
 * double is correctly rounded to a few ulp, so that E = numerator / wfnorm of a 1000-orbital cluster is
 * meaningful at the 1e-11 Eh level (the elements of M^-1 enter E with weights of order |E| ~ 1e4 Eh). */
static int f_invert(int n, double *M, double *Minv, double *det, double *minratio)
{
    typedef long double ld;
    ld *A = (ld *)xcalloc((size_t)n * n + 1, sizeof(ld)), *X = (ld *)xcalloc((size_t)n * n + 1, sizeof(ld));
    ld d = 1.0L;
    double pmax = 0.0, pmin = 1e300;
    for (size_t i = 0; i < (size_t)n * n; ++i) A[i] = M[i];
    for (int i = 0; i < n; ++i) X[(size_t)i * n + i] = 1.0L;
    int ok = 1;
    for (int k = 0; k < n && ok; ++k) {
        int p = k; ld best = fabsl(A[(size_t)k * n + k]);
        for (int i = k + 1; i < n; ++i) if (fabsl(A[(size_t)i * n + k]) > best) { best = fabsl(A[(size_t)i * n + k]); p = i; }
        if (best == 0.0L) { ok = 0; break; }
        if (p != k) {
            for (int j = 0; j < n; ++j) { ld t = A[(size_t)k * n + j]; A[(size_t)k * n + j] = A[(size_t)p * n + j]; A[(size_t)p * n + j] = t;
                                          t = X[(size_t)k * n + j]; X[(size_t)k * n + j] = X[(size_t)p * n + j]; X[(size_t)p * n + j] = t; }
            d = -d;
        }
        ld pv = A[(size_t)k * n + k];
        d *= pv; if ((double)best > pmax) pmax = (double)best; if ((double)best < pmin) pmin = (double)best;
        ld ipv = 1.0L / pv;
        for (int j = 0; j < n; ++j) { A[(size_t)k * n + j] *= ipv; X[(size_t)k * n + j] *= ipv; }
        for (int i = 0; i < n; ++i) {
            if (i == k) continue;
            ld f = A[(size_t)i * n + k];
            if (f == 0.0L) continue;
            ld *ai = A + (size_t)i * n, *xi = X + (size_t)i * n; const ld *ak = A + (size_t)k * n, *xk = X + (size_t)k * n;
            for (int j = 0; j < n; ++j) { ai[j] -= f * ak[j]; xi[j] -= f * xk[j]; }
        }
    }
    if (ok) {
        /* R = I - M X (long double), X <- X + X R */
        ld *R = A;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                ld s = (i == j) ? 1.0L : 0.0L;
                for (int k = 0; k < n; ++k) s -= (ld)M[(size_t)i * n + k] * X[(size_t)k * n + j];
                R[(size_t)i * n + j] = s;
            }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                ld s = X[(size_t)i * n + j];
                for (int k = 0; k < n; ++k) s += X[(size_t)i * n + k] * R[(size_t)k * n + j];
                Minv[(size_t)i * n + j] = (double)s;
            }
    }
    free(A); free(X);
    if (!ok) { *det = 0.0; *minratio = 0.0; return 0; }
    *det = (double)d; *minratio = pmin / pmax;
    return 1;
}

/* entry of a spin-orbital slot (valence.F90:1220-1223) */
static int f_entry_of_slot(const f_ctx *F, int i) { return i <= F->nnd ? i : F->nnd + (i - F->nnd + 1) / 2; }
static int f_slot(const f_ctx *F, int o, int is) { return o <= F->nnd ? o : 2 * o - F->nnd + is - 2; }   /* valence.F90:1334 */

/* alpha / beta blocks exactly as set_up_unpaired_docc + build_abket fill them for npair == 0
 * (valence.F90:2444-2586): rows = bra spin-orbitals, columns = ket spin-orbitals */
static int f_cofactors(f_ctx *F)
{
    vo_ctx *c = F->c;
    set_up_unpaired_docc(c);           /* bra_a / ket_a / bra_b / ket_b (the wdet copies it makes are not used) */
    int na = c->nalpha, nb = c->nbeta, ne = c->nelec;
    F->na = na; F->nb = nb;
    F->posa_b = IARR(ne); F->posa_k = IARR(ne); F->posb_b = IARR(ne); F->posb_k = IARR(ne);
    for (int i = 1; i <= na; ++i) { F->posa_b[c->bra_a[i]] = i; F->posa_k[c->ket_a[i]] = i; }
    for (int i = 1; i <= nb; ++i) { F->posb_b[c->bra_b[i]] = i; F->posb_k[c->ket_b[i]] = i; }
    double *A = (double *)xcalloc((size_t)na * na + 1, sizeof(double)), *B = (double *)xcalloc((size_t)nb * nb + 1, sizeof(double));
    F->Ai = (double *)xcalloc((size_t)na * na + 1, sizeof(double)); F->Bi = (double *)xcalloc((size_t)nb * nb + 1, sizeof(double));
    for (int r = 1; r <= na; ++r)
        for (int k = 1; k <= na; ++k)
            A[(size_t)(r - 1) * na + (k - 1)] = F->Se[(size_t)(f_entry_of_slot(F, c->bra_a[r]) - 1) * F->nso + (f_entry_of_slot(F, c->ket_a[k]) - 1)];
    for (int r = 1; r <= nb; ++r)
        for (int k = 1; k <= nb; ++k)
            B[(size_t)(r - 1) * nb + (k - 1)] = F->Se[(size_t)(f_entry_of_slot(F, c->bra_b[r]) - 1) * F->nso + (f_entry_of_slot(F, c->ket_b[k]) - 1)];
    double da = 1.0, db = 1.0, ra = 1.0, rb = 1.0;
    int ok = 1;
    if (na > 0) ok = ok && f_invert(na, A, F->Ai, &da, &ra);
    if (nb > 0) ok = ok && f_invert(nb, B, F->Bi, &db, &rb);
    free(A); free(B);
    if (!ok || ra < 1e-12 || rb < 1e-12) return 0;
    F->la_det = da; F->lb_det = db; F->c0 = da * db;
    return 1;
}

/* Minv[ket position][bra position] of the alpha / beta block; the inverse of M (rows bra, cols ket) is indexed
 * [ket][bra], and f_invert returns M^-1 row-major, i.e. Minv[k][r] */

/* first-order cofactor / c0 of bra spin-orbital i -> ket spin-orbital k (0 when the spins differ) */
static double f_c1_one(int na, int nb, const int *posa_b, const int *posa_k, const int *posb_b, const int *posb_k, const double *Ai, const double *Bi, int i, int k)
{
    if (posa_b[i] && posa_k[k]) return Ai[(size_t)(posa_k[k] - 1) * na + (posa_b[i] - 1)];
    if (posb_b[i] && posb_k[k]) return Bi[(size_t)(posb_k[k] - 1) * nb + (posb_b[i] - 1)];
    return 0.0;
}
/* second-order cofactor / c0 for the pairings i -> k, j -> l (det(), valence.F90:1895-2056, in inverse form) */
static double f_c2_one(int na, int nb, const int *posa_b, const int *posa_k, const int *posb_b, const int *posb_k, const double *Ai, const double *Bi,
                       int i, int k, int j, int l)
{
#define A_(kk, rr) Ai[(size_t)((kk) - 1) * na + ((rr) - 1)]
#define B_(kk, rr) Bi[(size_t)((kk) - 1) * nb + ((rr) - 1)]
    int ia = posa_b[i] && posa_k[k], ib = posb_b[i] && posb_k[k];
    int ja = posa_b[j] && posa_k[l], jb = posb_b[j] && posb_k[l];
    if (!(ia || ib) || !(ja || jb)) return 0.0;
    if (ia && ja) {
        /* same spin: the crossed pairing needs i -> l and j -> k in the alpha lists as well (always true there) */
        return A_(posa_k[k], posa_b[i]) * A_(posa_k[l], posa_b[j]) - A_(posa_k[l], posa_b[i]) * A_(posa_k[k], posa_b[j]);
    }
    if (ib && jb) return B_(posb_k[k], posb_b[i]) * B_(posb_k[l], posb_b[j]) - B_(posb_k[l], posb_b[i]) * B_(posb_k[k], posb_b[j]);
    if (ia && jb) return A_(posa_k[k], posa_b[i]) * B_(posb_k[l], posb_b[j]);
    return B_(posb_k[k], posb_b[i]) * A_(posa_k[l], posa_b[j]);
#undef A_
#undef B_
}
static double f_c1(const f_ctx *F, int i, int k)
{
    if (!F->ndp) return f_c1_one(F->na, F->nb, F->posa_b, F->posa_k, F->posb_b, F->posb_k, F->Ai, F->Bi, i, k);
    long double s = 0.0L;
    for (int d = 0; d < F->ndp; ++d) {
        const struct f_dp_s *D = &F->dp[d];
        s += (long double)D->wt * f_c1_one(F->na, F->nb, D->posa_b, D->posa_k, D->posb_b, D->posb_k, D->Ai, D->Bi, i, k);
    }
    return (double)s;
}
static double f_c2(const f_ctx *F, int i, int k, int j, int l)
{
    if (!F->ndp) return f_c2_one(F->na, F->nb, F->posa_b, F->posa_k, F->posb_b, F->posb_k, F->Ai, F->Bi, i, k, j, l);
    long double s = 0.0L;
    for (int d = 0; d < F->ndp; ++d) {
        const struct f_dp_s *D = &F->dp[d];
        s += (long double)D->wt * f_c2_one(F->na, F->nb, D->posa_b, D->posa_k, D->posb_b, D->posb_k, D->Ai, D->Bi, i, k, j, l);
    }
    return (double)s;
}

/* Spin-coupled wavefunctions (npair > 0): every determinant pair of density_sc / dbra / dket (valence.F90:1612-1870) -- bra coupling
 * isc x ket coupling jsc x every subset of pairs with alpha and beta exchanged in the bra x every such subset in the ket -- with its
 * own alpha / beta blocks (rows = bra spin-orbitals, columns = ket spin-orbitals, build_abket :2524-2586), inverted in extended
 * precision, weight coeff_sc(isc) coeff_sc(jsc) det(A) det(B) (density(), :1576-1588).  c0 = 1 in this mode. */
static int f_cofactors_sc(f_ctx *F)
{
    vo_ctx *c = F->c;
    const int npair = c->npair, nsc = c->nspinc, ne = c->nelec;
    if (nsc < 1 || npair > 12) return 0;
    set_up_unpaired_docc(c);           /* positions npair+1.. of bra_a / ket_a / bra_b / ket_b: unpaired, DOCC alpha / DOCC beta */
    const int na = npair + c->nunpd + c->ndocc, nb = npair + c->ndocc;
    F->na = na; F->nb = nb;
    const long long nmask = 1LL << npair;
    F->ndp = 0;
    F->dp = (struct f_dp_s *)xcalloc((size_t)(nsc * nsc * nmask * nmask) + 1, sizeof(struct f_dp_s));
    double *A = (double *)xcalloc((size_t)na * na + 1, sizeof(double)), *B = (double *)xcalloc((size_t)nb * nb + 1, sizeof(double));
    int ok = 1;
    for (int isc = 1; isc <= nsc && ok; ++isc)
        for (int jsc = 1; jsc <= nsc && ok; ++jsc)
            for (long long bm = 0; bm < nmask && ok; ++bm)
                for (long long km = 0; km < nmask && ok; ++km) {
                    for (int k = 1; k <= npair; ++k) {
                        int b1 = PAIRSC(c, k, 1, isc), b2 = PAIRSC(c, k, 2, isc), k1 = PAIRSC(c, k, 1, jsc), k2 = PAIRSC(c, k, 2, jsc);
                        if ((bm >> (k - 1)) & 1) { int t = b1; b1 = b2; b2 = t; }
                        if ((km >> (k - 1)) & 1) { int t = k1; k1 = k2; k2 = t; }
                        c->bra_a[k] = b1; c->bra_b[k] = b2; c->ket_a[k] = k1; c->ket_b[k] = k2;
                    }
                    struct f_dp_s *D = &F->dp[F->ndp];
                    D->posa_b = IARR(ne); D->posa_k = IARR(ne); D->posb_b = IARR(ne); D->posb_k = IARR(ne);
                    for (int i = 1; i <= na; ++i) { D->posa_b[c->bra_a[i]] = i; D->posa_k[c->ket_a[i]] = i; }
                    for (int i = 1; i <= nb; ++i) { D->posb_b[c->bra_b[i]] = i; D->posb_k[c->ket_b[i]] = i; }
                    for (int r = 1; r <= na; ++r)
                        for (int k = 1; k <= na; ++k)
                            A[(size_t)(r - 1) * na + (k - 1)] = F->Se[(size_t)(f_entry_of_slot(F, c->bra_a[r]) - 1) * F->nso + (f_entry_of_slot(F, c->ket_a[k]) - 1)];
                    for (int r = 1; r <= nb; ++r)
                        for (int k = 1; k <= nb; ++k)
                            B[(size_t)(r - 1) * nb + (k - 1)] = F->Se[(size_t)(f_entry_of_slot(F, c->bra_b[r]) - 1) * F->nso + (f_entry_of_slot(F, c->ket_b[k]) - 1)];
                    D->Ai = (double *)xcalloc((size_t)na * na + 1, sizeof(double)); D->Bi = (double *)xcalloc((size_t)nb * nb + 1, sizeof(double));
                    double da = 1.0, db = 1.0, ra = 1.0, rb = 1.0;
                    if (na > 0) ok = ok && f_invert(na, A, D->Ai, &da, &ra);
                    if (nb > 0) ok = ok && f_invert(nb, B, D->Bi, &db, &rb);
                    if (!ok || ra < 1e-12 || rb < 1e-12) ok = 0;
                    D->wt = c->coeff_sc[isc] * c->coeff_sc[jsc] * da * db;
                    ++F->ndp;
                }
    free(A); free(B);
    F->c0 = 1.0; F->la_det = F->lb_det = 1.0;
    return ok;
}

/* ---- entry groups, pair groups, shell-pair tables ---------------------------------------------------- */
/* sorted set of the real atoms of orbitals o1 and o2; returns its size */
static int f_atom_set(const vo_ctx *c, int o1, int o2, int *out)
{
    int n = 0;
    const int orb[2] = {o1, o2};
    for (int q = 0; q < 2; ++q)
        for (int k = 1; k <= c->orbas_atnum[orb[q]]; ++k) {
            int a = c->atom_alias[ATSET(c, k, orb[q])], dup = 0;
            for (int z = 0; z < n; ++z) dup = dup || out[z] == a;
            if (!dup) out[n++] = a;
        }
    for (int i = 1; i < n; ++i) { int x = out[i], j = i; for (; j > 0 && out[j - 1] > x; --j) out[j] = out[j - 1]; out[j] = x; }
    return n;
}
static int f_subset(const int *a, int na, const int *b, int nb)   /* a subset of b (both sorted) */
{
    int j = 0;
    for (int i = 0; i < na; ++i) { while (j < nb && b[j] < a[i]) ++j; if (j == nb || b[j] != a[i]) return 0; }
    return 1;
}

static void f_add_shell(int *list, int *n, int s) { for (int k = 0; k < *n; ++k) if (list[k] == s) return; list[(*n)++] = s; }

static void f_build_groups(f_ctx *F, int isolate)
{
    vo_ctx *c = F->c;
    const int GMAX = 6;
    F->grp = (f_grp *)xcalloc((size_t)F->nso + 1, sizeof(f_grp));
    F->grp_of = IARR(F->nso);
    F->ngrp = 0; F->subject_grp = -1;
    /* consecutive entries whose atom sets are nested share a group (e.g. the five orbitals of one water molecule):
     * every AO quartet of a group quartet is then generated once for all of their orbital pairs */
    int *gat = IARR(2 * c->mxctr + 2), *sat = IARR(2 * c->mxctr + 2), ngat = 0;
    for (int s = 1; s <= F->nso; ++s) {
        int g = F->ngrp - 1, join = 0;
        int nsat = f_atom_set(c, F->eB[s], F->eK[s], sat);
        if (g >= 0 && s != isolate && g != F->subject_grp && F->grp[g].n < GMAX)
            join = f_subset(sat, nsat, gat, ngat) || f_subset(gat, ngat, sat, nsat);
        if (!join) { g = F->ngrp++; F->grp[g].e = IARR(GMAX); F->grp[g].n = 0; if (s == isolate) F->subject_grp = g; ngat = 0; }
        if (nsat > ngat) { for (int k = 0; k < nsat; ++k) gat[k] = sat[k]; ngat = nsat; }
        F->grp[g].e[F->grp[g].n++] = s;
        F->grp_of[s] = g;
    }
    free(gat); free(sat);
    for (int g = 0; g < F->ngrp; ++g) {
        f_grp *G = &F->grp[g];
        int cap = 0;
        for (int k = 0; k < G->n; ++k) cap += F->o2[F->eB[G->e[k]]].nsh + F->o2[F->eK[G->e[k]]].nsh;
        G->shB = IARR(cap); G->shK = IARR(cap); G->nshB = G->nshK = 0;
        for (int k = 0; k < G->n; ++k) {
            const f_orb *ob = &F->o2[F->eB[G->e[k]]], *ok = &F->o2[F->eK[G->e[k]]];
            for (int x = 0; x < ob->nsh; ++x) f_add_shell(G->shB, &G->nshB, ob->gsh[x]);
            for (int x = 0; x < ok->nsh; ++x) f_add_shell(G->shK, &G->nshK, ok->gsh[x]);
        }
        G->cB = DARR((size_t)G->n * G->nshB * 6); G->cK = DARR((size_t)G->n * G->nshK * 6);
        for (int k = 0; k < G->n; ++k) {
            const f_orb *ob = &F->o2[F->eB[G->e[k]]], *ok = &F->o2[F->eK[G->e[k]]];
            for (int x = 0; x < ob->nsh; ++x)
                for (int y = 0; y < G->nshB; ++y)
                    if (G->shB[y] == ob->gsh[x]) for (int a = 0; a < 6; ++a) G->cB[((size_t)k * G->nshB + y) * 6 + a] = ob->cf[6 * x + a];
            for (int x = 0; x < ok->nsh; ++x)
                for (int y = 0; y < G->nshK; ++y)
                    if (G->shK[y] == ok->gsh[x]) for (int a = 0; a < 6; ++a) G->cK[((size_t)k * G->nshK + y) * 6 + a] = ok->cf[6 * x + a];
        }
    }
}

static void f_free_pg(f_pg *P)
{
    for (int k = 0; k < P->nsp; ++k) { free(P->sp[k].pp); free(P->sp[k].E); free(P->sp[k].D); }
    free(P->sp); free(P->ps); free(P->pt);
    memset(P, 0, sizeof *P);
}

static int f_cmp_pp(const void *a, const void *b) { double x = ((const f_pp *)a)->w, y = ((const f_pp *)b)->w; return x < y ? 1 : (x > y ? -1 : 0); }
static int f_cmp_sp(const void *a, const void *b) { double x = ((const f_sp *)a)->wmax, y = ((const f_sp *)b)->wmax; return x < y ? 1 : (x > y ? -1 : 0); }

/* tables of pair group (g,h): entry pairs (s in g, t in h; s >= t only for the diagonal group of symmetric lists),
 * shell pairs (X in bra-support(g), Y in ket-support(h)) with primitive pairs, Hermite expansions, densities and
 * the rigorous magnitude bound w = sum_comp max_p |D| sqrt((comp|comp)) of every primitive pair */
static void f_build_pg(const f_ctx *F, int g, int h, f_pg *P)
{
    const vo_ctx *c = F->c;
    const f_grp *G = &F->grp[g], *Hh = &F->grp[h];
    memset(P, 0, sizeof *P);
    P->g = g; P->h = h;
    P->ps = IARR(G->n * Hh->n); P->pt = IARR(G->n * Hh->n);
    int *pe = IARR(G->n * Hh->n), *pf = IARR(G->n * Hh->n);
    for (int a = 0; a < G->n; ++a)
        for (int b = 0; b < Hh->n; ++b) {
            if (F->sym && g == h && G->e[a] < Hh->e[b]) continue;
            P->ps[P->np] = G->e[a]; P->pt[P->np] = Hh->e[b]; pe[P->np] = a; pf[P->np] = b; P->np++;
        }
    const int np = P->np;
    P->sp = (f_sp *)xcalloc((size_t)G->nshB * Hh->nshK + 1, sizeof(f_sp));
    double Ex[3][5][5][10], R[FL_NH8];
    for (int x = 0; x < G->nshB; ++x)
        for (int y = 0; y < Hh->nshK; ++y) {
            const f_shell *A = &F->sh[G->shB[x]], *B = &F->sh[Hh->shK[y]];
            f_sp sp; memset(&sp, 0, sizeof sp);
            sp.la = A->l; sp.lb = B->l; sp.Lp = A->l + B->l; sp.nh = f_nh(sp.Lp);
            const int na = shell_size(A->l), nb = shell_size(B->l), nab = na * nb;
            sp.nab = nab;
            sp.D = DARR((size_t)nab * np);
            double dmax[36]; int any = 0;
            for (int ab = 0; ab < nab; ++ab) dmax[ab] = 0.0;
            for (int p = 0; p < np; ++p)
                for (int a = 0; a < na; ++a)
                    for (int b = 0; b < nb; ++b) {
                        double d = G->cB[((size_t)pe[p] * G->nshB + x) * 6 + a] * Hh->cK[((size_t)pf[p] * Hh->nshK + y) * 6 + b];
                        sp.D[(size_t)(a * nb + b) * np + p] = d;
                        if (fabs(d) > dmax[a * nb + b]) dmax[a * nb + b] = fabs(d);
                        any = any || d != 0.0;
                    }
            if (!any) { free(sp.D); continue; }
            double AB2 = 0.0;
            for (int d = 0; d < 3; ++d) AB2 += (A->r[d] - B->r[d]) * (A->r[d] - B->r[d]);
            sp.pp = (f_pp *)xcalloc((size_t)A->nprim * B->nprim, sizeof(f_pp));
            double *Etmp = DARR((size_t)A->nprim * B->nprim * nab * sp.nh);
            int *order = IARR(A->nprim * B->nprim);
            int npp = 0;
            for (int ia = 0; ia < A->nprim; ++ia)
                for (int ib = 0; ib < B->nprim; ++ib) {
                    double a = A->exps[ia], b = B->exps[ib], p = a + b;
                    double K = A->coef[ia] * B->coef[ib] * exp(-a * b / p * AB2);
                    if (K == 0.0) continue;
                    f_pp *pp = &sp.pp[npp];
                    pp->p = p;
                    for (int d = 0; d < 3; ++d) { pp->P[d] = (a * A->r[d] + b * B->r[d]) / p; f_hermite_E(A->l, B->l, p, pp->P[d] - A->r[d], pp->P[d] - B->r[d], Ex[d]); }
                    double *E = Etmp + (size_t)npp * nab * sp.nh;
                    for (int i = 0; i < na; ++i) {
                        int ax = NXYZ(c, 1, funmin[A->l + 1] + i), ay = NXYZ(c, 2, funmin[A->l + 1] + i), az = NXYZ(c, 3, funmin[A->l + 1] + i);
                        for (int j = 0; j < nb; ++j) {
                            int bx = NXYZ(c, 1, funmin[B->l + 1] + j), by = NXYZ(c, 2, funmin[B->l + 1] + j), bz = NXYZ(c, 3, funmin[B->l + 1] + j);
                            for (int q = 0; q < sp.nh; ++q) {
                                int t = fh_t[q], u = fh_u[q], v = fh_v[q];
                                E[(size_t)(i * nb + j) * sp.nh + q] = (t <= ax + bx && u <= ay + by && v <= az + bz) ? K * Ex[0][ax][bx][t] * Ex[1][ay][by][u] * Ex[2][az][bz][v] : 0.0;
                            }
                        }
                    }
                    /* self-repulsion of every component: (ab|ab) = 2 pi^2.5 / (p p sqrt(2p)) sum_hh' E_h (-1)^h' E_h' R_{h+h'}(p/2, 0) */
                    double zero[3] = {0.0, 0.0, 0.0};
                    f_hermite_R(2 * sp.Lp, 0.5 * p, zero, R);
                    double pref = 2.0 * pow(F_PI, 2.5) / (p * p * sqrt(2.0 * p)), w = 0.0;
                    for (int ab = 0; ab < nab; ++ab) {
                        if (dmax[ab] == 0.0) continue;
                        const double *e = E + (size_t)ab * sp.nh;
                        double sii = 0.0;
                        for (int q = 0; q < sp.nh; ++q) {
                            if (e[q] == 0.0) continue;
                            for (int r = 0; r < sp.nh; ++r) {
                                if (e[r] == 0.0) continue;
                                double sg = ((fh_t[r] + fh_u[r] + fh_v[r]) & 1) ? -1.0 : 1.0;
                                sii += e[q] * sg * e[r] * R[fh_idx[fh_t[q] + fh_t[r]][fh_u[q] + fh_u[r]][fh_v[q] + fh_v[r]]];
                            }
                        }
                        w += dmax[ab] * sqrt(fabs(pref * sii));
                    }
                    pp->w = w;
                    if (w >= 1e3) { fprintf(stderr, "oracle(fast): primitive bound %g out of the assumed range\n", w); exit(2); }
                    if (w * 1e3 < F->thr) continue;    /* cannot reach thr against any partner (all bounds < 1e3) */
                    order[npp] = npp;
                    ++npp;
                }
            /* sort primitives by decreasing bound, carrying their Hermite tables along */
            if (npp > 0) {
                f_pp *sorted = (f_pp *)xcalloc((size_t)npp, sizeof(f_pp));
                for (int k = 0; k < npp; ++k) { sorted[k] = sp.pp[k]; sorted[k].P[0] = sp.pp[k].P[0]; }
                /* stable insertion sort on (w, original index) through an index array */
                for (int k = 1; k < npp; ++k) {
                    int o = order[k], j = k;
                    for (; j > 0 && sp.pp[order[j - 1]].w < sp.pp[o].w; --j) order[j] = order[j - 1];
                    order[j] = o;
                }
                sp.E = DARR((size_t)npp * nab * sp.nh);
                for (int k = 0; k < npp; ++k) {
                    sorted[k] = sp.pp[order[k]];
                    memcpy(sp.E + (size_t)k * nab * sp.nh, Etmp + (size_t)order[k] * nab * sp.nh, sizeof(double) * (size_t)nab * sp.nh);
                }
                free(sp.pp); sp.pp = sorted; sp.npp = npp; sp.wmax = sorted[0].w;
                P->sp[P->nsp++] = sp;
            } else { free(sp.pp); free(sp.D); }
            free(Etmp); free(order);
        }
    qsort(P->sp, (size_t)P->nsp, sizeof(f_sp), f_cmp_sp);
    free(pe); free(pf);
    (void)f_cmp_pp;
}

/* contracted AO block I[ab][cd] += (ab|cd) of two shell pairs, primitive quartets below thr skipped */
static void f_eri_sp(const f_sp *A, const f_sp *B, double thr, double *I, long long *npq)
{
    const int nab = A->nab, ncd = B->nab, nhA = A->nh, nhB = B->nh, L = A->Lp + B->Lp;
    const double c25 = 34.98683665524972497;    /* 2 pi^2.5 */
    double R[FL_NH8], M[FL_NH4 * FL_NH4], tmp[36 * FL_NH4];
    for (int i = 0; i < A->npp; ++i) {
        const f_pp *a = &A->pp[i];
        if (a->w * B->wmax < thr) break;
        const double *Ea = A->E + (size_t)i * nab * nhA;
        for (int j = 0; j < B->npp; ++j) {
            const f_pp *b = &B->pp[j];
            if (a->w * b->w < thr) break;
            const double *Eb = B->E + (size_t)j * ncd * nhB;
            double p = a->p, q = b->p, alpha = p * q / (p + q);
            double PQ[3] = {a->P[0] - b->P[0], a->P[1] - b->P[1], a->P[2] - b->P[2]};
            double pref = c25 / (p * q * sqrt(p + q));
            ++*npq;
            f_hermite_R(L, alpha, PQ, R);
            if (L == 0) { I[0] += pref * Ea[0] * Eb[0] * R[0]; continue; }
            for (int h = 0; h < nhA; ++h)
                for (int r = 0; r < nhB; ++r) M[h * nhB + r] = fh_sgn[r] * R[fh_sum[h][r]];
            for (int ab = 0; ab < nab; ++ab)
                for (int r = 0; r < nhB; ++r) {
                    double s = 0.0;
                    for (int h = 0; h < nhA; ++h) s += Ea[ab * nhA + h] * M[h * nhB + r];
                    tmp[ab * nhB + r] = s;
                }
            for (int ab = 0; ab < nab; ++ab)
                for (int cd = 0; cd < ncd; ++cd) {
                    double s = 0.0;
                    for (int r = 0; r < nhB; ++r) s += tmp[ab * nhB + r] * Eb[cd * nhB + r];
                    I[ab * ncd + cd] += pref * s;
                }
        }
    }
}

/* E[p * Q->np + q] = (B(s) K(t) | B(u) K(v)) for the entry pairs p = (s,t) of P, q = (u,v) of Q */
static void f_block(const f_pg *P, const f_pg *Q, double thr, double *E, long long *npq)
{
    const int npP = P->np, npQ = Q->np;
    double I[36 * 36], *Hh = (double *)malloc(sizeof(double) * 36 * (size_t)npQ);
    for (int k = 0; k < npP * npQ; ++k) E[k] = 0.0;
    for (int x = 0; x < P->nsp; ++x) {
        const f_sp *A = &P->sp[x];
        if (Q->nsp == 0 || A->wmax * Q->sp[0].wmax < thr) break;
        for (int k = 0; k < A->nab * npQ; ++k) Hh[k] = 0.0;
        for (int y = 0; y < Q->nsp; ++y) {
            const f_sp *B = &Q->sp[y];
            if (A->wmax * B->wmax < thr) break;
            for (int k = 0; k < A->nab * B->nab; ++k) I[k] = 0.0;
            f_eri_sp(A, B, thr, I, npq);
            for (int ab = 0; ab < A->nab; ++ab)
                for (int cd = 0; cd < B->nab; ++cd) {
                    double v = I[ab * B->nab + cd];
                    if (v == 0.0) continue;
                    const double *d = B->D + (size_t)cd * npQ;
                    double *hrow = Hh + (size_t)ab * npQ;
                    for (int q = 0; q < npQ; ++q) hrow[q] += v * d[q];
                }
        }
        for (int ab = 0; ab < A->nab; ++ab) {
            const double *hrow = Hh + (size_t)ab * npQ, *d = A->D + (size_t)ab * npP;
            for (int p = 0; p < npP; ++p) {
                if (d[p] == 0.0) continue;
                double *e = E + (size_t)p * npQ;
                for (int q = 0; q < npQ; ++q) e[q] += d[p] * hrow[q];
            }
        }
    }
    free(Hh);
}

/* the same two routines with every accumulation (primitive sums, both density transformations) in extended precision:
 * VO_FAST_XACC=1.  A block of a 256-molecule cluster sums ~1e5 terms; in plain double that costs ~1e-15 relative per block,
 * which is what separates the fixture from the GPU engine's (double-double assembled) energy at that size. */
static void f_eri_sp_x(const f_sp *A, const f_sp *B, double thr, long double *I, long long *npq)
{
    const int nab = A->nab, ncd = B->nab, nhA = A->nh, nhB = B->nh, L = A->Lp + B->Lp;
    const double c25 = 34.98683665524972497;    /* 2 pi^2.5 */
    double R[FL_NH8], M[FL_NH4 * FL_NH4];
    long double tmp[36 * FL_NH4];
    for (int i = 0; i < A->npp; ++i) {
        const f_pp *a = &A->pp[i];
        if (a->w * B->wmax < thr) break;
        const double *Ea = A->E + (size_t)i * nab * nhA;
        for (int j = 0; j < B->npp; ++j) {
            const f_pp *b = &B->pp[j];
            if (a->w * b->w < thr) break;
            const double *Eb = B->E + (size_t)j * ncd * nhB;
            double p = a->p, q = b->p, alpha = p * q / (p + q);
            double PQ[3] = {a->P[0] - b->P[0], a->P[1] - b->P[1], a->P[2] - b->P[2]};
            double pref = c25 / (p * q * sqrt(p + q));
            ++*npq;
            f_hermite_R(L, alpha, PQ, R);
            if (L == 0) { I[0] += (long double)pref * Ea[0] * Eb[0] * R[0]; continue; }
            for (int h = 0; h < nhA; ++h)
                for (int r = 0; r < nhB; ++r) M[h * nhB + r] = fh_sgn[r] * R[fh_sum[h][r]];
            for (int ab = 0; ab < nab; ++ab)
                for (int r = 0; r < nhB; ++r) {
                    long double s = 0.0L;
                    for (int h = 0; h < nhA; ++h) s += (long double)Ea[ab * nhA + h] * M[h * nhB + r];
                    tmp[ab * nhB + r] = s;
                }
            for (int ab = 0; ab < nab; ++ab)
                for (int cd = 0; cd < ncd; ++cd) {
                    long double s = 0.0L;
                    for (int r = 0; r < nhB; ++r) s += tmp[ab * nhB + r] * Eb[cd * nhB + r];
                    I[ab * ncd + cd] += (long double)pref * s;
                }
        }
    }
}

static void f_block_x(const f_pg *P, const f_pg *Q, double thr, double *E, long long *npq)
{
    const int npP = P->np, npQ = Q->np;
    long double I[36 * 36], *Hh = (long double *)malloc(sizeof(long double) * 36 * (size_t)npQ);
    long double *El = (long double *)malloc(sizeof(long double) * ((size_t)npP * npQ + 1));
    for (int k = 0; k < npP * npQ; ++k) El[k] = 0.0L;
    for (int x = 0; x < P->nsp; ++x) {
        const f_sp *A = &P->sp[x];
        if (Q->nsp == 0 || A->wmax * Q->sp[0].wmax < thr) break;
        for (int k = 0; k < A->nab * npQ; ++k) Hh[k] = 0.0;
        for (int y = 0; y < Q->nsp; ++y) {
            const f_sp *B = &Q->sp[y];
            if (A->wmax * B->wmax < thr) break;
            for (int k = 0; k < A->nab * B->nab; ++k) I[k] = 0.0;
            f_eri_sp_x(A, B, thr, I, npq);
            for (int ab = 0; ab < A->nab; ++ab)
                for (int cd = 0; cd < B->nab; ++cd) {
                    long double v = I[ab * B->nab + cd];
                    if (v == 0.0L) continue;
                    const double *d = B->D + (size_t)cd * npQ;
                    long double *hrow = Hh + (size_t)ab * npQ;
                    for (int q = 0; q < npQ; ++q) hrow[q] += v * d[q];
                }
        }
        for (int ab = 0; ab < A->nab; ++ab) {
            const long double *hrow = Hh + (size_t)ab * npQ;
            const double *d = A->D + (size_t)ab * npP;
            for (int p = 0; p < npP; ++p) {
                if (d[p] == 0.0) continue;
                long double *e = El + (size_t)p * npQ;
                for (int q = 0; q < npQ; ++q) e[q] += d[p] * hrow[q];
            }
        }
    }
    for (int k = 0; k < npP * npQ; ++k) E[k] = (double)El[k];
    free(Hh); free(El);
}

static int f_xacc = 0;      /* VO_FAST_XACC: extended-precision block sums */

/* ---- the reference's task bookkeeping for one integral value ------------------------------------------- */
typedef struct { long double e2; vo_counters cnt; } f_acc;   /* extended-precision accumulation of the task sum */

static long long f_tri(int a, int b) { return (long long)a * (a - 1) / 2 + b; }   /* xm_dtriang: ij = i(i-1)/2 + j, i >= j, 1-based */
#define F_SCH(F, a, b) ((F)->c->schwarz[indx(a, b)])

/* sum over the spin loop (valence.F90:1333-1414) of the second-order cofactor / c0 for entries
 * (bra i' , bra j') -> (ket k', ket l'): pairings i->k and j->l; exch != 0 pairs i->l and j->k instead */
static double f_spin_sum(const f_ctx *F, int io, int jo, int ko, int lo, int exch)
{
    const int nnd = F->nnd;
    const int iso = io > nnd ? 2 : 1, jso = jo > nnd ? 2 : 1, kso = ko > nnd ? 2 : 1, lso = lo > nnd ? 2 : 1;
    double sum = 0.0;
    for (int is = 1; is <= iso; ++is) {
        int i = f_slot(F, io, is);
        for (int js = 1; js <= jso; ++js) {
            int j = f_slot(F, jo, js);
            if (!(j < i)) continue;
            for (int ks = 1; ks <= kso; ++ks) {
                int k = f_slot(F, ko, ks);
                for (int ls = 1; ls <= lso; ++ls) {
                    int l = f_slot(F, lo, ls);
                    if (!(l < k)) continue;
                    sum += exch ? f_c2(F, i, l, j, k) : f_c2(F, i, k, j, l);
                }
            }
        }
    }
    return sum;
}

/* V = int2e(B(a), K(b), B(c), K(d)).  It is the direct integral of task (io=a, ko=b, jo=c, lo=d) and the exchanged
 * integral of task (io=a, lo=b, jo=c, ko=d) whenever those tasks exist (valence.F90:1167-1186). */
static void f_visit(const f_ctx *F, int a, int b, int c_, int d, double V, f_acc *acc)
{
    const vo_ctx *c = F->c;
    const int nnd = F->nnd, sym = F->sym, jorb = F->jorb;
    if (!(F_SCH(F, a, b) * F_SCH(F, c_, d) > c->itol)) return;     /* :1189-1190 -- the same product in both roles */
    for (int role = 0; role < 2; ++role) {
        int io = a, jo = c_, ko = role == 0 ? b : d, lo = role == 0 ? d : b;
        if (io < jo || ko < lo) continue;
        if ((io == jo && io <= nnd) || (ko == lo && ko <= nnd)) continue;                  /* :1182-1186 */
        if (sym && f_tri(io, jo) < f_tri(ko, lo)) continue;                                /* ijorb >= klorb */
        if (role == 0) acc->cnt.schwarz_erep++; else acc->cnt.schwarz_exch++;
        int nonsub = io != jorb && jo != jorb && ko != jorb && lo != jorb;
        double val;
        if (nonsub && io == jo && ko == lo) {                                              /* :1213-1216 */
            val = F_SCH(F, io, ko) * F_SCH(F, io, ko);
            if (role == 0) acc->cnt.shortcut++;
        } else {
            val = V;
            int counted = 1;
            if (role == 1) {
                /* the exchanged integral reuses the direct one when that was computed and the two ket orbitals
                 * coincide (:1261-1262) */
                int erep_sig = F_SCH(F, io, ko) * F_SCH(F, jo, lo) > c->itol;
                if (erep_sig && F->eK[lo] == F->eK[ko]) counted = 0;
            }
            if (counted) {
                acc->cnt.int2e_calls++;
                long long nq = (long long)F->nshell_kept[F->eB[a]] * F->nshell_kept[F->eK[b]] * F->nshell_kept[F->eB[c_]] * F->nshell_kept[F->eK[d]];
                acc->cnt.shell_quartets += nq; acc->cnt.shell_quartets_2e += nq;
            }
        }
        if (!(fabs(val) > c->itol)) continue;                                              /* :1286-1287 */
        if (role == 0) acc->cnt.value_erep++; else acc->cnt.value_exch++;
        long double s = (long double)f_spin_sum(F, io, jo, ko, lo, role) * val;
        if (sym && !(ko == io && lo == jo)) s = s * 2.0L;                                   /* :1420-1428 */
        acc->e2 += s;                           /* erep_sum - exchanged_erep_sum: the exchange sign sits in the cofactor */
    }
}

/* all tasks fed by the block (P,Q), P >= Q: every canonical element with each distinct image under the
 * permutational symmetry of the integral */
static void f_visit_block(const f_ctx *F, const f_pg *P, const f_pg *Q, int same, const double *E, f_acc *acc)
{
    for (int p = 0; p < P->np; ++p)
        for (int q = 0; q < Q->np; ++q) {
            if (same && q > p) continue;
            const int s = P->ps[p], t = P->pt[p], u = Q->ps[q], v = Q->pt[q];
            const double V = E[(size_t)p * Q->np + q];
            int im[8][4], nim = 0;
            const int base[2][4] = {{s, t, u, v}, {u, v, s, t}};
            for (int k = 0; k < 2; ++k) {
                const int a = base[k][0], b = base[k][1], c_ = base[k][2], d = base[k][3];
                int cand[4][4] = {{a, b, c_, d}, {b, a, c_, d}, {a, b, d, c_}, {b, a, d, c_}};
                for (int m = 0; m < (F->sym ? 4 : 1); ++m) {
                    int dup = 0;
                    for (int z = 0; z < nim; ++z) dup = dup || (im[z][0] == cand[m][0] && im[z][1] == cand[m][1] && im[z][2] == cand[m][2] && im[z][3] == cand[m][3]);
                    if (!dup) { for (int z = 0; z < 4; ++z) im[nim][z] = cand[m][z]; ++nim; }
                }
            }
            for (int z = 0; z < nim; ++z) f_visit(F, im[z][0], im[z][1], im[z][2], im[z][3], V, acc);
        }
}

/* ---- set-up / tear-down --------------------------------------------------------------------------------- */
static void f_free_orb(f_orb *o) { free(o->gsh); free(o->cf); }

static void f_release_lists(f_ctx *F)
{
    if (F->pg) { for (int k = 0; k < F->npg; ++k) f_free_pg(&F->pg[k]); free(F->pg); F->pg = NULL; }
    free(F->pg_of); F->pg_of = NULL;
    if (F->grp) { for (int g = 0; g < F->ngrp; ++g) { free(F->grp[g].e); free(F->grp[g].shB); free(F->grp[g].shK); free(F->grp[g].cB); free(F->grp[g].cK); } free(F->grp); F->grp = NULL; }
    free(F->grp_of); F->grp_of = NULL;
    free(F->Se); free(F->He); F->Se = F->He = NULL;
    free(F->posa_b); free(F->posa_k); free(F->posb_b); free(F->posb_k); F->posa_b = F->posa_k = F->posb_b = F->posb_k = NULL;
    free(F->Ai); free(F->Bi); F->Ai = F->Bi = NULL;
    if (F->dp) {
        for (int d = 0; d < F->ndp; ++d) { free(F->dp[d].posa_b); free(F->dp[d].posa_k); free(F->dp[d].posb_b); free(F->dp[d].posb_k); free(F->dp[d].Ai); free(F->dp[d].Bi); }
        free(F->dp); F->dp = NULL; F->ndp = 0;
    }
    free(F->eB); free(F->eK); F->eB = F->eK = NULL;
}

static void f_release(f_ctx *F)
{
    f_release_lists(F);
    if (F->o1) { for (int o = 1; o <= F->norb; ++o) { f_free_orb(&F->o1[o]); f_free_orb(&F->o2[o]); } free(F->o1); free(F->o2); }
    if (F->cache) { for (long long k = 0; k < F->ncache; ++k) free(F->cache[k]); free(F->cache); }
    free(F->nshell_kept); free(F->sh); free(F->atom_first); free(F->S); free(F->H);
    memset(F, 0, sizeof *F);
}

/* geometry / basis / orbital dependent part; norb_total = orbitals incl. the dummy ones of first_order_opt */
static void f_prepare(f_ctx *F, vo_ctx *c, int norb_total, double thr)
{
    memset(F, 0, sizeof *F);
    f_init_tables();
    F->c = c; F->thr = thr;
    f_build_shells(F);
    f_ao_1e(F);
    F->norb = norb_total;
    F->o1 = (f_orb *)xcalloc((size_t)norb_total + 1, sizeof(f_orb)); F->o2 = (f_orb *)xcalloc((size_t)norb_total + 1, sizeof(f_orb));
    F->nshell_kept = IARR(norb_total);
    for (int o = 1; o <= norb_total; ++o) { f_expand(F, o, 0, &F->o1[o]); f_expand(F, o, 1, &F->o2[o]); F->nshell_kept[o] = F->o2[o].nsh; }
}

static void f_se_row(void *ctx, long long sl, int tid)
{
    f_ctx *F = (f_ctx *)ctx;
    const int s = (int)sl + 1, nso = F->nso;
    (void)tid;
    for (int t = 1; t <= nso; ++t)
        f_orb_1e(F, &F->o1[F->eB[s]], &F->o1[F->eK[t]], &F->Se[(size_t)(s - 1) * nso + (t - 1)], &F->He[(size_t)(s - 1) * nso + (t - 1)]);
}
static void f_pg_one(void *ctx, long long gh, int tid)
{
    f_ctx *F = (f_ctx *)ctx;
    (void)tid;
    if (F->pg_of[gh] >= 0) f_build_pg(F, (int)(gh / F->ngrp), (int)(gh % F->ngrp), &F->pg[F->pg_of[gh]]);
}
/* everything that depends on the current bra / ket lists except the cofactors */
static void f_set_lists(f_ctx *F, int nnd, int nso, int sym, int jorb, int isolate)
{
    vo_ctx *c = F->c;
    f_release_lists(F);
    F->nnd = nnd; F->nso = nso; F->sym = sym; F->jorb = jorb;
    F->eB = IARR(nso); F->eK = IARR(nso);
    for (int s = 1; s <= nso; ++s) { int sl = s <= nnd ? s : 2 * s - nnd - 1; F->eB[s] = c->bra[sl]; F->eK[s] = c->ket[sl]; }
    F->Se = (double *)xcalloc((size_t)nso * nso, sizeof(double)); F->He = (double *)xcalloc((size_t)nso * nso, sizeof(double));
    f_parfor(nso, 4, f_se_row, F);
    f_build_groups(F, isolate);
    const int ng = F->ngrp;
    F->pg_of = (int *)xcalloc((size_t)ng * ng, sizeof(int));
    F->pg = (f_pg *)xcalloc((size_t)ng * ng + 1, sizeof(f_pg));
    int n = 0;
    for (int g = 0; g < ng; ++g)
        for (int h = 0; h < ng; ++h) { F->pg_of[(size_t)g * ng + h] = -1; if (!sym || h <= g) F->pg_of[(size_t)g * ng + h] = n++; }
    F->npg = n;
    f_parfor((long long)ng * ng, 4, f_pg_one, F);
}

/* schwarz_ints (valence.F90:1489-1523): schwarz(indx(i,j)) = sqrt((B(i) K(j) | B(i) K(j))), i >= j, from the diagonal
 * of the blocks (P,P).  Lists must be the symmetric ones (as at every reference call site). */
static void f_schwarz_one(void *ctx, long long k, int tid)
{
    f_ctx *F = (f_ctx *)ctx;
    vo_ctx *c = F->c;
    const f_pg *P = &F->pg[k];
    (void)tid;
    if (P->np == 0 || P->nsp == 0) return;
    if (!F->sym && P->h > P->g) return;
    double *E = DARR((size_t)P->np * P->np);
    long long npq = 0;
    if (f_xacc) f_block_x(P, P, 1e-32, E, &npq); else f_block(P, P, 1e-32, E, &npq);
    for (int p = 0; p < P->np; ++p) {
        int s = P->ps[p], t = P->pt[p];
        if (s >= t) c->schwarz[indx(s, t)] = sqrt(E[(size_t)p * P->np + p]);
    }
    free(E);
}
static void f_schwarz(f_ctx *F)
{
    vo_ctx *c = F->c;
    for (int i = 1; i <= F->nso; ++i) for (int j = 1; j <= i; ++j) c->schwarz[indx(i, j)] = 0.0;
    f_parfor(F->npg, 4, f_schwarz_one, F);
}

static void f_set_smax(f_ctx *F)
{
    for (int k = 0; k < F->npg; ++k) {
        f_pg *P = &F->pg[k];
        double m = 0.0;
        for (int p = 0; p < P->np; ++p) { double x = F_SCH(F, P->ps[p], P->pt[p]); if (x > m) m = x; }
        P->smax = m;
    }
}

typedef struct { f_ctx *F; long long *blk; long double *eblk; f_acc *accs; long long *npq; } f_blk_job;
static void f_blk_one(void *ctx, long long k, int tid)
{
    f_blk_job *J = (f_blk_job *)ctx;
    f_ctx *F = J->F;
    const long long npg = F->npg;
    const int sg = F->subject_grp;
    f_acc *acc = &J->accs[tid];
    double E[36 * 36];
    const f_pg *P = &F->pg[J->blk[k] / npg], *Q = &F->pg[J->blk[k] % npg];
    const double *Eu = E;
    int cacheable = F->cache && sg >= 0 && P->g != sg && P->h != sg && Q->g != sg && Q->h != sg;
    if (cacheable && F->cache[J->blk[k]]) Eu = F->cache[J->blk[k]];
    else {
        long long npq = 0;
        if (f_xacc) f_block_x(P, Q, F->thr, E, &npq); else f_block(P, Q, F->thr, E, &npq);
        J->npq[tid] += npq;
        if (cacheable) { double *keep = DARR((size_t)P->np * Q->np); memcpy(keep, E, sizeof(double) * (size_t)P->np * Q->np); F->cache[J->blk[k]] = keep; }
    }
    acc->e2 = 0.0L;
    f_visit_block(F, P, Q, P == Q, Eu, acc);
    J->eblk[k] = acc->e2;
}

/* vsvb_energy (valence.F90:1010-1434) for the current lists: numerator energy and wfnorm, counters into c->cnt */
static int f_vsvb_energy(f_ctx *F, double *energy_out, double *wfnorm_out, long long *npq_out)
{
    vo_ctx *c = F->c;
    if (!(c->npair > 0 ? f_cofactors_sc(F) : f_cofactors(F))) return -1;
    const int nnd = F->nnd, nelec = c->nelec, nso = F->nso;
    /* 1e part (:1072-1106) */
    long double e1 = 0.0L, wn = 0.0L;
    for (int i = 1; i <= nelec; ++i) {
        int i_is_docc = i > nnd;
        for (int j = 1; j <= nelec; ++j) {
            int j_is_docc = j > nnd;
            if (i_is_docc && j_is_docc && (i % 2) != (j % 2)) continue;
            double d1 = f_c1(F, i, j);
            size_t k = (size_t)(f_entry_of_slot(F, i) - 1) * nso + (f_entry_of_slot(F, j) - 1);
            wn += (long double)F->Se[k] * d1; e1 += (long double)F->He[k] * d1;
        }
    }
    wn = wn / (long double)nelec;
    /* 2e part: blocks (P,Q), P >= Q, that can hold an integral passing the Schwarz screen */
    f_set_smax(F);
    const long long npg = F->npg;
    long long nblk = 0;
    long long *blk = (long long *)xcalloc((size_t)(npg * (npg + 1) / 2) + 1, sizeof(long long));
    for (long long p = 0; p < npg; ++p)
        for (long long q = 0; q <= p; ++q)
            if (F->pg[p].np && F->pg[q].np && F->pg[p].smax * F->pg[q].smax > c->itol) blk[nblk++] = p * npg + q;
    long double *eblk = (long double *)xcalloc((size_t)nblk + 1, sizeof(long double));
    const int nthr = 64;
    f_acc *accs = (f_acc *)xcalloc((size_t)nthr, sizeof(f_acc));
    long long npq_tot = 0;
    f_blk_job J; J.F = F; J.blk = blk; J.eblk = eblk; J.accs = accs; J.npq = (long long *)xcalloc(64, sizeof(long long));
    f_parfor(nblk, 16, f_blk_one, &J);
    for (int t = 0; t < 64; ++t) npq_tot += J.npq[t];
    free(J.npq);
    /* fixed-order extended-precision sum of the block energies: reproducible for any thread count */
    long double e2 = 0.0L;
    for (long long k = 0; k < nblk; ++k) e2 += eblk[k];
    for (int t = 0; t < nthr; ++t) {
        c->cnt.schwarz_erep += accs[t].cnt.schwarz_erep; c->cnt.schwarz_exch += accs[t].cnt.schwarz_exch;
        c->cnt.shortcut += accs[t].cnt.shortcut; c->cnt.int2e_calls += accs[t].cnt.int2e_calls;
        c->cnt.value_erep += accs[t].cnt.value_erep; c->cnt.value_exch += accs[t].cnt.value_exch;
        c->cnt.shell_quartets += accs[t].cnt.shell_quartets; c->cnt.shell_quartets_2e += accs[t].cnt.shell_quartets_2e;
    }
    free(accs); free(eblk); free(blk);
    if (getenv("VO_FAST_DEBUG")) fprintf(stderr, "[fast] e1/c0 %.17Lg e2/c0 %.17Lg trace/nelec %.17Lg c0 %.17g\n", e1, e2, wn, F->c0);
    *energy_out = (double)(F->c0 * (e1 + e2));
    *wfnorm_out = (double)(F->c0 * wn);
    if (npq_out) *npq_out = npq_tot;
    return 0;
}

/* ---- public entry points ----------------------------------------------------------------------------------- */
typedef struct { double enucrep, energy, wfnorm, numerator; vo_counters cnt; long long prim_quartets, blocks; double seconds; } vo_fast_result;

/* guess_energy (valence.F90:309-345) through the fast path.  thr = primitive-quartet magnitude cut (rigorous
 * bound on what is dropped per orbital-level integral element); 0 -> 1e-24.  Returns -1 for inputs outside the
 * fast path's limits (npair > 0, singular overlap blocks). */
int vo_fast_guess_energy(vo_ctx *c, int nthreads, double thr, vo_fast_result *out)
{
    memset(out, 0, sizeof *out);
    f_nthreads = nthreads;
    f_xacc = getenv("VO_FAST_XACC") != NULL && atoi(getenv("VO_FAST_XACC")) != 0;
    double t0 = mono_now();
    setup_energy(c);
    memset(&c->cnt, 0, sizeof c->cnt);
    c->nrank = 1; c->irank = 0;
    int nnd = 2 * c->npair + c->nunpd, nso = nnd + c->ndocc;
    default_lists(c, nnd);
    f_ctx F;
    const int dbg = getenv("VO_FAST_DEBUG") != NULL;
    double t1 = mono_now();
    f_prepare(&F, c, c->norbs, thr > 0.0 ? thr : 1e-24);
    double t2 = mono_now();
    f_set_lists(&F, nnd, nso, 1, 0, 0);
    double t3 = mono_now();
    f_schwarz(&F);
    double t4 = mono_now();
    double e, w; long long npq = 0;
    int rc = f_vsvb_energy(&F, &e, &w, &npq);
    if (dbg) fprintf(stderr, "[fast] setup %.2f s, AO 1e + orbitals %.2f s, lists/pair groups %.2f s (%d groups, %d pair groups), schwarz %.2f s, energy %.2f s (%lld primitive quartets)\n",
                     t1 - t0, t2 - t1, t3 - t2, F.ngrp, F.npg, t4 - t3, mono_now() - t4, npq);
    if (rc == 0) {
        out->enucrep = c->enucrep; out->numerator = e; out->wfnorm = w; out->energy = e / w + c->enucrep;
        out->cnt = c->cnt; out->prim_quartets = npq;
    }
    f_release(&F);
    out->seconds = mono_now() - t0;
    return rc;
}

/* ham / ovl of first_order_opt (valence.F90:527-764) for orbital iorb through the fast path: the (ib,jb) loop of
 * first_order_matrices above with f_vsvb_energy in place of vsvb_energy; integral blocks that do not touch the
 * substituted entry are computed once (the reference's eribuf, :1227-1273). */
int vo_fast_first_order(vo_ctx *c, int iorb, int nthreads, double thr, double *ham, double *ovl, int *hdim_out, vo_counters *cnt)
{
    if (c->npair > 0) return -1;
    f_nthreads = nthreads;
    setup_energy(c);
    ensure_opt_arrays(c);
    memset(&c->cnt, 0, sizeof c->cnt);
    c->nrank = 1; c->irank = 0;
    int num_non_docc = 2 * c->npair + c->nunpd, num_spatial_orbs = num_non_docc + c->ndocc, eorb;
    if (iorb <= num_non_docc) { eorb = iorb; default_lists(c, num_non_docc); }
    else {
        for (int i = 1; i <= num_non_docc; ++i) { c->bra[i] = i; c->ket[i] = i; }
        c->bra[num_non_docc + 1] = iorb; c->ket[num_non_docc + 1] = iorb; c->bra[num_non_docc + 2] = iorb; c->ket[num_non_docc + 2] = iorb;
        int i = num_non_docc + 3;
        for (int idocc = num_non_docc + 1; idocc <= num_non_docc + c->ndocc; ++idocc)
            if (idocc != iorb) { c->bra[i] = idocc; c->ket[i] = idocc; c->bra[i + 1] = idocc; c->ket[i + 1] = idocc; i += 2; }
        eorb = 2 * c->npair + c->nunpd + 1;
        num_non_docc += 2; num_spatial_orbs += 1;
    }
    first_order_dummies(c, iorb);
    const int norbas = c->map_orbs[iorb + 1] - c->map_orbs[iorb];
    f_ctx F;
    f_prepare(&F, c, c->norbs + norbas, thr > 0.0 ? thr : 1e-24);
    /* Schwarz table of the unsubstituted lists (:666-667) */
    f_set_lists(&F, num_non_docc, num_spatial_orbs, 1, 0, 0);
    f_schwarz(&F);
    for (int i = 1; i <= c->hdim; ++i) for (int j = 1; j <= c->hdim; ++j) { HAM(c, i, j) = 0.0; OVL(c, i, j) = 0.0; }
    int rc = 0;
    for (int pass = 0; pass < 2 && rc == 0; ++pass) {
        if (pass == 1) {
            if (!(c->nunpd > 0 && iorb > 2 * c->npair + c->nunpd)) break;
            c->bra[eorb] = iorb; c->ket[eorb] = iorb;
            eorb = 2 * c->npair + c->nunpd + 2;
        }
        /* entry of the substituted slot; subject orbital index as vsvb_energy derives it (:1199-1206) */
        int jorb = iorb;
        if (iorb > 2 * c->npair + c->nunpd) jorb = 2 * c->npair + c->nunpd + 1 + (pass == 1);
        if (F.cache) { for (long long k = 0; k < F.ncache; ++k) free(F.cache[k]); free(F.cache); F.cache = NULL; }
        for (int ib = 1; ib <= norbas && rc == 0; ++ib) {
            int idf = c->xpset[c->map_orbs[iorb] + ib - 1];
            c->bra[eorb] = (idf < 1) ? c->norbs + idf : c->norbs + ib;
            for (int jb = 1; jb <= ib && rc == 0; ++jb) {
                idf = c->xpset[c->map_orbs[iorb] + jb - 1];
                c->ket[eorb] = (idf < 1) ? c->norbs + idf : c->norbs + jb;
                f_set_lists(&F, num_non_docc, num_spatial_orbs, 0, jorb, jorb);
                if (!F.cache) { F.ncache = (long long)F.npg * F.npg; F.cache = (double **)xcalloc((size_t)F.ncache, sizeof(double *)); }
                double energy, wfnorm;
                rc = f_vsvb_energy(&F, &energy, &wfnorm, NULL);
                if (pass == 0) { HAM(c, ib, jb) = energy; OVL(c, ib, jb) = wfnorm; }
                else { HAM(c, ib, jb) = HAM(c, ib, jb) + energy; OVL(c, ib, jb) = OVL(c, ib, jb) + wfnorm; }
            }
        }
    }
    for (int i = 1; i <= norbas; ++i)
        for (int j = 1; j <= i - 1; ++j) { HAM(c, j, i) = HAM(c, i, j); OVL(c, j, i) = OVL(c, i, j); }
    f_release(&F);
    if (ham) memcpy(ham, c->ham, sizeof(double) * (size_t)c->hdim * c->hdim);
    if (ovl) memcpy(ovl, c->ovl, sizeof(double) * (size_t)c->hdim * c->hdim);
    if (hdim_out) *hdim_out = c->hdim;
    if (cnt) *cnt = c->cnt;
    return rc == 0 ? norbas : -1;
}
