/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker / baseline.
 *
 * A literal CPU restatement, in plain C, of the reference's VSVB energy path
 * (/root/reference/src/valence.F90 and friends).  The reference itself cannot
 * be built here (no Fortran compiler, no MPI, SIMINT un-vendored), so this
 * file follows its loops statement by statement -- same task order, same
 * screens, same Givens determinants -- and every function cites the lines it
 * follows.  AO integrals come from vo_integrals.c (own McMurchie-Davidson
 * code standing in for SIMINT).  Arrays are 1-based like the Fortran so the
 * two can be read side by side.
 *
 * Parity pin: tests/test_oracle_golden.py checks this oracle against the
 * reference's own golden energies (examples/test_examples.py:63-151,
 * testing/testing.py:69-151) committed under tests/golden/.
 *
 * Optional AO-block memoisation (vo_set_memo) caches SIMINT-equivalent shell
 * blocks by (atom, shell) so the literal loops run fast enough for tests; it
 * changes no value.  The CPU baseline is timed with memoisation OFF, which is
 * what the reference does (it recomputes every shell quartet per orbital
 * quartet, valence.F90:3287-3437).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "vo_internal.h"

/* ------------------------------------------------------------------------ */
typedef struct {
    long long ntasks;          /* 2e tasks visited (valence.F90:1163)                   */
    long long same_orb_skip;   /* skipped at :1182-1186                                 */
    long long schwarz_pass;    /* erep_is_signif .or. exchange_is_signif at :1194        */
    long long schwarz_erep;    /* erep_is_signif at :1189                                */
    long long schwarz_exch;    /* exchange_is_signif at :1190                            */
    long long shortcut;        /* Schwarz shortcut taken at :1213-1216                   */
    long long int2e_calls;     /* int2e calls from the 2e loop (:1256, :1264)            */
    long long value_erep;      /* |erep_int| > itol at :1286                             */
    long long value_exch;      /* |exchanged_erep_int| > itol at :1287                   */
    long long density2;        /* density(2,...) calls at :1393                          */
    long long shell_quartets;  /* simint_compute_eri calls (:3398), all callers          */
    long long shell_quartets_2e; /* ... from the 2e loop of vsvb_energy only             */
    long long determinants;    /* givdr calls (:2067)                                    */
    long long eri_cached;      /* retrieved from eribuf (:1234-1237)                     */
} vo_counters;

typedef struct memo_s memo_t;

typedef struct vo_ctx {
    /* header (xm_module.F90:41-42) */
    int natom, natom_t, npair, nunpd, ndocc, totlen, xpmax, nspinc, num_sh, num_pr, nang, ndf, nset, nxorb, mxctr;
    int ntol_c, ntol_d, ntol_i, ntol_e_min, ntol_e_max, max_iter;
    double ptbnmax, feather;
    int nelec, norbs, nalpha, nbeta, hdim, nstore, max_obs;
    int *orbset;                   /* (2,nset) */
    int *atom_t;                   /* natom+xpmax */
    double *coords;                /* (3,natom+xpmax), bohr after vo_load */
    double *coords_angs;           /* (3,natom) as read */
    int *map_atom2shell, *num_shell_atom; /* natom_t+xpmax */
    int *map_shell2prim, *ang_mom;
    double *nuc_charge, *exponent, *con_coeff, *unnorm;
    int *orbas_atnum, *orbas_atset, *map_orbs, *xpset, *xorb, *root;
    double *coeff, *coeff_in;      /* coeff_in = weights as read (for re-runs) */
    int *nxyz; double *angn, *ashl, *ashi;
    double *dij, *dkl, *coeffi, *coeffj, *coeffk, *coeffl;
    int *atom_ndf, *ndf2orb, *xpnew;
    double *schwarz, *eribuf;
    double dtol, itol;
    int spinopt, store_eri, eri_stored, dem_gs;
    double sint, hint, gint, enucrep;
    double *ham, *ovl;
    /* densitywork */
    int dme_b[3], dme_k[3];
    int *pair_sc; double *coeff_sc;
    int *bra_a, *bra_b, *ket_a, *ket_b, *bexch, *kexch, *bra, *ket;
    double *wdet, *abra_npair, *bbra_npair, *abra_docc_un, *bbra_docc_un, *aket, *bket, *aket_docc_un, *bket_docc_un;
    int padded_size;
    /* xm */
    int nrank, irank;
    /* oracle extras */
    int *atom_alias;               /* dummy atom -> real atom (memo key) */
    int in2e;                      /* inside the 2e loop of vsvb_energy */
    int count_only;                /* vo_count_tasks: the 2e loop only counts what int2e would evaluate */
    vo_counters cnt;
    memo_t *memo, *memo2;
    int memo_on;
    long long task_limit;          /* >0: stop the 2e loop after this many tasks (baseline sample) */
    double deadline;               /* >0: wall-clock second (CLOCK_MONOTONIC) after which the timed sample stops */
    int expired;
    int quiet;
    double *integrals_store;
} vo_ctx;

/* 1-based accessors mirroring the Fortran array shapes */
#define COORDS(c, d, a) ((c)->coords[((a) - 1) * 3 + ((d) - 1)])
#define ATSET(c, k, o) ((c)->orbas_atset[((o) - 1) * (c)->mxctr + ((k) - 1)])
#define PAIRSC(c, k, s, isc) ((c)->pair_sc[(((isc) - 1) * 2 + ((s) - 1)) * (c)->npair + ((k) - 1)])
#define WDET(c, i, j) ((c)->wdet[((j) - 1) * (c)->nelec + ((i) - 1)])
#define AKET(c, i, j) ((c)->aket[((j) - 1) * (c)->padded_size + ((i) - 1)])
#define BKET(c, i, j) ((c)->bket[((j) - 1) * (c)->padded_size + ((i) - 1)])
#define HAM(c, i, j) ((c)->ham[((j) - 1) * (c)->hdim + ((i) - 1)])
#define OVL(c, i, j) ((c)->ovl[((j) - 1) * (c)->hdim + ((i) - 1)])
#define NXYZ(c, d, n) ((c)->nxyz[((n) - 1) * 3 + ((d) - 1)])

static int shell_size(int l) { return (l + 1) * (l + 2) / 2; } /* integrals_module.F90:104-109 */
static const int funmin[8] = {0, 1, 2, 5, 11, 21, 36, 57};     /* valence.F90:2903-2904 (1-based) */
static const int funmax[8] = {0, 1, 4, 10, 20, 35, 56, 84};

static void *xcalloc(size_t n, size_t sz) { void *p = calloc(n ? n : 1, sz); if (!p) { fprintf(stderr, "oracle: out of memory\n"); exit(2); } return p; }
#define IARR(n) ((int *)xcalloc((size_t)(n) + 2, sizeof(int)))
#define DARR(n) ((double *)xcalloc((size_t)(n) + 2, sizeof(double)))

/* ======================================================================== */
/* record-based list-directed reader (xm_module.F90:25-335)  */
/* ======================================================================== */
typedef struct { char *buf; size_t len, pos; char **tok; int ntok, itok; } reader_t;

static int next_record(reader_t *r)
{
    /* tokenise the next line into r->tok */
    if (r->pos >= r->len) return 0;
    size_t e = r->pos;
    while (e < r->len && r->buf[e] != '\n') ++e;
    r->ntok = 0; r->itok = 0;
    size_t i = r->pos;
    while (i < e) {
        while (i < e && (r->buf[i] == ' ' || r->buf[i] == '\t' || r->buf[i] == ',' || r->buf[i] == '\r')) ++i;
        if (i >= e) break;
        size_t s = i;
        while (i < e && !(r->buf[i] == ' ' || r->buf[i] == '\t' || r->buf[i] == ',' || r->buf[i] == '\r')) ++i;
        r->buf[i == e ? e : i] = (i == e) ? r->buf[e] : '\0';
        r->tok = (char **)realloc(r->tok, sizeof(char *) * (size_t)(r->ntok + 1));
        r->tok[r->ntok++] = r->buf + s;
        if (i < e) ++i;
    }
    if (e < r->len) r->buf[e] = '\0';
    r->pos = e + 1;
    return 1;
}
static void begin_read(reader_t *r) { r->ntok = 0; r->itok = 0; }  /* a READ starts on a fresh record */
static const char *next_tok(reader_t *r)
{
    while (r->itok >= r->ntok)
        if (!next_record(r)) { fprintf(stderr, "oracle: input ended inside a READ\n"); exit(2); }
    return r->tok[r->itok++];
}
static int rd_i(reader_t *r) { return (int)strtol(next_tok(r), NULL, 10); }
static double rd_d(reader_t *r)
{
    char tmp[128]; const char *t = next_tok(r); size_t n = strlen(t); if (n > 127) n = 127;
    for (size_t i = 0; i < n; ++i) tmp[i] = (t[i] == 'D' || t[i] == 'd') ? 'e' : t[i];
    tmp[n] = 0; return strtod(tmp, NULL);
}

/* valence_initialize_module.F90:56-113 + xm_getdims/xm_input */
vo_ctx *vo_load(const char *path)
{
    FILE *fh = fopen(path, "rb");
    if (!fh) return NULL;
    reader_t R; memset(&R, 0, sizeof R);
    fseek(fh, 0, SEEK_END); R.len = (size_t)ftell(fh); fseek(fh, 0, SEEK_SET);
    R.buf = (char *)xcalloc(R.len + 2, 1);
    if (fread(R.buf, 1, R.len, fh) != R.len) { fclose(fh); return NULL; }
    fclose(fh);
    vo_ctx *c = (vo_ctx *)xcalloc(1, sizeof(vo_ctx));
    c->nrank = 1; c->irank = 0; c->memo_on = 1;

    begin_read(&R);
    c->natom = rd_i(&R); c->natom_t = rd_i(&R); c->npair = rd_i(&R); c->nunpd = rd_i(&R); c->ndocc = rd_i(&R);
    c->totlen = rd_i(&R); c->xpmax = rd_i(&R); c->nspinc = rd_i(&R); c->num_sh = rd_i(&R); c->num_pr = rd_i(&R);
    c->nang = rd_i(&R); c->ndf = rd_i(&R); c->nset = rd_i(&R); c->nxorb = rd_i(&R); c->mxctr = rd_i(&R);
    if (c->npair > 0 && c->nspinc < 1) { fprintf(stderr, "no spin couplings\n"); exit(2); }

    int nx = c->xpmax;
    c->atom_t = IARR(c->natom + nx); c->coords = DARR(3 * (c->natom + nx)); c->coords_angs = DARR(3 * c->natom);
    c->atom_alias = IARR(c->natom + nx);
    c->map_atom2shell = IARR(c->natom_t + nx); c->num_shell_atom = IARR(c->natom_t + nx);
    c->map_shell2prim = IARR(c->num_sh + 1); c->ang_mom = IARR(c->num_sh);
    c->nuc_charge = DARR(c->natom_t); c->exponent = DARR(c->num_pr); c->con_coeff = DARR(c->num_pr); c->unnorm = DARR(c->num_pr);
    int nsc = c->nspinc > 0 ? c->nspinc : 1;
    c->pair_sc = IARR(2 * c->npair * nsc); c->coeff_sc = DARR(nsc);
    c->xorb = IARR(c->nxorb); c->root = IARR(c->nxorb); c->orbset = IARR(2 * c->nset);
    c->nelec = 2 * c->npair + 2 * c->ndocc + c->nunpd;
    c->norbs = 2 * c->npair + c->ndocc + c->nunpd + c->ndf;
    c->nalpha = c->npair + c->nunpd + c->ndocc;
    c->nbeta = c->npair + c->ndocc;
    c->map_orbs = IARR(c->norbs + nx + 1); c->orbas_atnum = IARR(c->norbs + nx);
    c->orbas_atset = IARR(c->mxctr * (c->norbs + nx));
    c->xpset = IARR(c->totlen + nx); c->coeff = DARR(c->totlen + nx); c->coeff_in = DARR(c->totlen + nx);

    begin_read(&R);
    c->ntol_c = rd_i(&R); c->ntol_d = rd_i(&R); c->ntol_i = rd_i(&R);
    c->ntol_e_min = rd_i(&R); c->ntol_e_max = rd_i(&R); c->max_iter = rd_i(&R);
    c->ptbnmax = rd_d(&R); c->feather = rd_d(&R);
    for (int i = 1; i <= c->nset; ++i) { c->orbset[2 * (i - 1)] = rd_i(&R); c->orbset[2 * (i - 1) + 1] = rd_i(&R); }

    for (int i = 1; i <= c->natom; ++i) {
        begin_read(&R);
        c->atom_t[i] = rd_i(&R);
        for (int j = 1; j <= 3; ++j) COORDS(c, j, i) = rd_d(&R);
        c->atom_alias[i] = i;
    }
    int ns = 1, np = 1;
    for (int i = 1; i <= c->natom_t; ++i) {
        c->map_atom2shell[i] = ns;
        begin_read(&R);
        c->nuc_charge[i] = rd_d(&R);
        int nshell = rd_i(&R);
        c->num_shell_atom[i] = nshell;
        for (int j = 1; j <= nshell; ++j) {
            c->map_shell2prim[ns] = np;
            begin_read(&R);
            c->ang_mom[ns] = rd_i(&R);
            int con_length = rd_i(&R);
            if (con_length == 1) {
                begin_read(&R);
                c->exponent[np] = rd_d(&R); c->unnorm[np] = 1.0; ++np;
            } else {
                for (int k = 1; k <= con_length; ++k) {
                    begin_read(&R);
                    c->exponent[np] = rd_d(&R); c->unnorm[np] = rd_d(&R); ++np;
                }
            }
            ++ns;
        }
    }
    c->map_shell2prim[ns] = np;

    c->coeff_sc[1] = 1.0;
    if (c->npair > 0) {
        begin_read(&R);
        if (c->nspinc == 1) {
            for (int i = 1; i <= c->npair; ++i) { PAIRSC(c, i, 1, 1) = rd_i(&R); PAIRSC(c, i, 2, 1) = rd_i(&R); }
        } else {
            for (int j = 1; j <= c->nspinc; ++j) {
                c->coeff_sc[j] = rd_d(&R);
                for (int i = 1; i <= c->npair; ++i) { PAIRSC(c, i, 1, j) = rd_i(&R); PAIRSC(c, i, 2, j) = rd_i(&R); }
            }
        }
    }
    if (c->nxorb > 0) {
        begin_read(&R);
        for (int i = 1; i <= c->nxorb; ++i) { c->xorb[i] = rd_i(&R); c->root[i] = rd_i(&R); }
    }
    int j = 1;
    for (int i = 1; i <= c->norbs; ++i) {
        begin_read(&R);
        c->orbas_atnum[i] = rd_i(&R);
        for (int k = 1; k <= c->orbas_atnum[i]; ++k) ATSET(c, k, i) = rd_i(&R);
        int n = rd_i(&R);
        c->map_orbs[i] = j;
        begin_read(&R);
        for (int k = j; k <= j + n - 1; ++k) { c->xpset[k] = rd_i(&R); c->coeff[k] = rd_d(&R); }
        j += n;
    }
    c->map_orbs[c->norbs + 1] = j;
    for (int k = 1; k <= c->totlen + nx; ++k) c->coeff_in[k] = c->coeff[k];

    c->dtol = pow(10.0, -c->ntol_d);
    c->itol = pow(10.0, -c->ntol_i);
    for (int i = 1; i <= c->natom; ++i)
        for (int d = 1; d <= 3; ++d) c->coords_angs[(i - 1) * 3 + d - 1] = COORDS(c, d, i);
    /* tools_module.F90:5-17 angs2bohr */
    for (int i = 1; i <= c->natom; ++i)
        for (int d = 1; d <= 3; ++d) COORDS(c, d, i) = COORDS(c, d, i) * 1.889725987722;
    free(R.tok); free(R.buf);
    return c;
}

/* valence_api.F90:57-63 + angs2bohr: overwrite the geometry (Angstrom) */
void vo_set_coords(vo_ctx *c, const double *x)
{
    int k = 0;
    for (int i = 1; i <= c->natom; ++i)
        for (int j = 1; j <= 3; ++j) { c->coords_angs[k] = x[k]; COORDS(c, j, i) = x[k] * 1.889725987722; ++k; }
}

void vo_reset_orbitals(vo_ctx *c) { for (int k = 1; k <= c->totlen + c->xpmax; ++k) c->coeff[k] = c->coeff_in[k]; }

/* ======================================================================== */
/* AO block access (stands in for the SIMINT calls), with optional memo      */
/* ======================================================================== */
struct memo_s { uint64_t *keys; size_t *offs; size_t cap, used; double *pool; size_t pool_used, pool_cap; };

static void memo_free(memo_t *m) { if (!m) return; free(m->keys); free(m->offs); free(m->pool); free(m); }
static memo_t *memo_new(void)
{
    memo_t *m = (memo_t *)xcalloc(1, sizeof *m);
    m->cap = 1u << 16; m->keys = (uint64_t *)xcalloc(m->cap, 8); m->offs = (size_t *)xcalloc(m->cap, sizeof(size_t));
    m->pool_cap = 1u << 20; m->pool = (double *)xcalloc(m->pool_cap, 8);
    return m;
}
static void memo_grow(memo_t *m)
{
    size_t oc = m->cap; uint64_t *ok = m->keys; size_t *oo = m->offs;
    m->cap *= 2; m->keys = (uint64_t *)xcalloc(m->cap, 8); m->offs = (size_t *)xcalloc(m->cap, sizeof(size_t));
    for (size_t i = 0; i < oc; ++i)
        if (ok[i]) {
            size_t h = (size_t)((ok[i] * 0x9E3779B97F4A7C15ull) >> 20) & (m->cap - 1);
            while (m->keys[h]) h = (h + 1) & (m->cap - 1);
            m->keys[h] = ok[i]; m->offs[h] = oo[i];
        }
    free(ok); free(oo);
}
/* returns pointer to n doubles; *fresh = 1 when the caller must fill them */
static double *memo_get(memo_t *m, uint64_t key, size_t n, int *fresh)
{
    if (m->used * 2 > m->cap) memo_grow(m);
    size_t h = (size_t)((key * 0x9E3779B97F4A7C15ull) >> 20) & (m->cap - 1);
    while (m->keys[h]) { if (m->keys[h] == key) { *fresh = 0; return m->pool + m->offs[h]; } h = (h + 1) & (m->cap - 1); }
    if (m->pool_used + n > m->pool_cap) {
        /* pool pointers are handed out only transiently, so realloc is safe */
        while (m->pool_used + n > m->pool_cap) m->pool_cap *= 2;
        m->pool = (double *)realloc(m->pool, m->pool_cap * 8);
        if (!m->pool) { fprintf(stderr, "oracle: memo out of memory\n"); exit(2); }
    }
    m->keys[h] = key; m->offs[h] = m->pool_used; m->used++;
    double *p = m->pool + m->pool_used; m->pool_used += n; *fresh = 1; return p;
}

void vo_set_memo(vo_ctx *c, int on) { c->memo_on = on; if (!on) { memo_free(c->memo); c->memo = NULL; memo_free(c->memo2); c->memo2 = NULL; } }

/* shell_map(ii-mnshi+1, ia): shell `ii` (basis index) placed on atom `ia`
 * (valence_simint_module.F90:37-50, valence.F90:649-654) */
static void make_shell(const vo_ctx *c, int ii, int ia, vo_shell *s)
{
    s->l = c->ang_mom[ii];
    s->nprim = c->map_shell2prim[ii + 1] - c->map_shell2prim[ii];
    s->exps = c->exponent + c->map_shell2prim[ii];
    s->coef = c->con_coeff + c->map_shell2prim[ii];
    for (int d = 1; d <= 3; ++d) s->r[d - 1] = COORDS(c, d, ia);
}
/* dense id of a shell instance: (real atom, basis shell); < 2^16 at test sizes */
static uint64_t skey(const vo_ctx *c, int ii, int ia)
{
    uint64_t id = (uint64_t)(c->atom_alias[ia] - 1) * (uint64_t)c->num_sh + (uint64_t)(ii - 1);
    if (id > 0xFFFFull) { fprintf(stderr, "oracle: system too large for the AO memo (disable it)\n"); exit(2); }
    return id;
}

static const double *ao_block1(vo_ctx *c, int kind, int ii, int ia, int jj, int ja, int ictr)
{
    vo_shell A, B; make_shell(c, ii, ia, &A); make_shell(c, jj, ja, &B);
    size_t n = (size_t)vo_ncart(A.l) * vo_ncart(B.l);
    double *out = c->integrals_store; int fresh = 1;
    if (c->memo_on) {
        if (!c->memo) c->memo = memo_new();
        uint64_t key = ((uint64_t)(kind + 1) << 60) | (skey(c, ii, ia) << 36) | (skey(c, jj, ja) << 20) | (uint64_t)ictr;
        out = memo_get(c->memo, key, n, &fresh);
    }
    if (fresh) {
        if (kind == 0) vo_overlap_block(&A, &B, out);
        else if (kind == 1) vo_kinetic_block(&A, &B, out);
        else { double C[3] = {COORDS(c, 1, ictr), COORDS(c, 2, ictr), COORDS(c, 3, ictr)};
               vo_potential_block(&A, &B, c->nuc_charge[c->atom_t[ictr]], C, out); }
    }
    return out;
}
static const double *ao_eri(vo_ctx *c, int ii, int ic, int jj, int jc, int kk, int kc, int ll, int lc)
{
    vo_shell A, B, C, D; make_shell(c, ii, ic, &A); make_shell(c, jj, jc, &B); make_shell(c, kk, kc, &C); make_shell(c, ll, lc, &D);
    size_t n = (size_t)vo_ncart(A.l) * vo_ncart(B.l) * vo_ncart(C.l) * vo_ncart(D.l);
    double *out = c->integrals_store; int fresh = 1;
    c->cnt.shell_quartets++;
    if (c->in2e) c->cnt.shell_quartets_2e++;
    if (c->memo_on) {
        if (!c->memo2) c->memo2 = memo_new();
        /* exact key: four 16-bit ids; +1 so that no key is 0 (= empty slot) */
        uint64_t key = (skey(c, ii, ic) | (skey(c, jj, jc) << 16) | (skey(c, kk, kc) << 32) | (skey(c, ll, lc) << 48)) + 1;
        out = memo_get(c->memo2, key, n, &fresh);
    }
    if (fresh) vo_eri_block(&A, &B, &C, &D, out);
    return out;
}

/* ======================================================================== */
/* setup: norm_prim, cartesian, setangn, nuclear repulsion                   */
/* ======================================================================== */
static double dblfac(int n) { double r = 1.0; for (int i = 3; i <= n; i += 2) r *= (double)i; return r; } /* valence.F90:2314-2322 */

/* valence.F90:2264-2305 */
static void norm_prim(int ang_mom, int con_length, const double *exponent, double *con_coeff, const double *unnorm)
{
    const double two = 2.0;
    double pi32 = pow(acos(-1.0), 1.5);
    double fac = pi32 * dblfac(2 * ang_mom - 1) * pow(two, (double)(-ang_mom));
    double fax = -1.5 - ang_mom;
    for (int ig = 1; ig <= con_length; ++ig) {
        double sovl = fac * pow(two * exponent[ig], fax);
        con_coeff[ig] = unnorm[ig] * pow(sovl, -0.5);
    }
    double sovl = 0.0;
    for (int ig = 1; ig <= con_length; ++ig)
        for (int jg = 1; jg <= con_length; ++jg)
            sovl = sovl + fac * con_coeff[ig] * con_coeff[jg] * pow(exponent[ig] + exponent[jg], fax);
    sovl = pow(sovl, -0.5);
    for (int ig = 1; ig <= con_length; ++ig) con_coeff[ig] = con_coeff[ig] * sovl;
}

/* valence.F90:2337-2380 with -DCCA_ORDER (Makefile:61) */
static void cartesian(vo_ctx *c)
{
    NXYZ(c, 1, 1) = 0; NXYZ(c, 2, 1) = 0; NXYZ(c, 3, 1) = 0;
    int n = 2;
    for (int l = 1; l <= c->nang; ++l)
        for (int i = 0; i <= l; ++i)
            for (int j = 0; j <= i; ++j) { NXYZ(c, 1, n) = l - i; NXYZ(c, 2, n) = i - j; NXYZ(c, 3, n) = j; ++n; }
}

/* valence.F90:2395-2421 */
static void setangn(vo_ctx *c)
{
    int ij = 0;
    for (int i = 0; i <= c->nang; ++i) {
        double fii = 1.0;
        for (int j = 1; j <= 2 * i - 1; j += 2) fii = fii * (double)j;
        fii = sqrt(fii);
        c->ashl[i] = fii; c->ashi[i] = 1.0 / fii;
        for (int j = 1; j <= shell_size(i); ++j) {
            ++ij;
            c->angn[ij] = c->ashl[i] * c->ashi[NXYZ(c, 1, ij)] * c->ashi[NXYZ(c, 2, ij)] * c->ashi[NXYZ(c, 3, ij)];
        }
    }
}

/* tools_module.F90:19-40 */
static double nuclear_repulsion(const vo_ctx *c)
{
    double nre = 0.0;
    for (int i = 2; i <= c->natom; ++i)
        for (int j = 1; j <= i - 1; ++j) {
            double zij = c->nuc_charge[c->atom_t[i]] * c->nuc_charge[c->atom_t[j]];
            if (fabs(zij) > 1.e-12) {
                double dx = COORDS(c, 1, i) - COORDS(c, 1, j), dy = COORDS(c, 2, i) - COORDS(c, 2, j), dz = COORDS(c, 3, i) - COORDS(c, 3, j);
                double rsq = dx * dx + dy * dy + dz * dz;
                if (rsq > 1.e-5) nre = nre + zij * pow(rsq, -0.5);
            }
        }
    return nre;
}

/* ======================================================================== */
/* orbital-level integrals: ndf2obs, ovint, int1e, int2e                     */
/* ======================================================================== */
/* valence.F90:2191-2247 */
static void ndf2obs(vo_ctx *c, int iorb, int indf)
{
    int ib = 0;
    for (int iat = 1; iat <= c->orbas_atnum[indf]; ++iat) {
        c->atom_ndf[2 * (iat - 1) + 0] = ib + 1;
        int it = c->atom_t[ATSET(c, iat, indf)];
        int mnshi = c->map_atom2shell[it], mxshi = mnshi + c->num_shell_atom[it] - 1;
        for (int ish = mnshi; ish <= mxshi; ++ish) ib = ib + shell_size(c->ang_mom[ish]);
        c->atom_ndf[2 * (iat - 1) + 1] = ib;
    }
    ib = 1;
    for (int iat = 1; iat <= c->orbas_atnum[iorb]; ++iat) {
        for (int jat = 1; jat <= c->orbas_atnum[indf]; ++jat)
            if (ATSET(c, jat, indf) == ATSET(c, iat, iorb)) c->ndf2orb[jat] = ib;
        int it = c->atom_t[ATSET(c, iat, iorb)];
        int mnshi = c->map_atom2shell[it], mxshi = mnshi + c->num_shell_atom[it] - 1;
        for (int ish = mnshi; ish <= mxshi; ++ish) ib = ib + shell_size(c->ang_mom[ish]);
    }
    int i = 0, j = 0;
    for (ib = c->map_orbs[indf]; ib <= c->map_orbs[indf + 1] - 1; ++ib) {
        for (int ia = 1; ia <= c->orbas_atnum[indf]; ++ia)
            if (c->xpset[ib] >= c->atom_ndf[2 * (ia - 1)] && c->xpset[ib] <= c->atom_ndf[2 * (ia - 1) + 1])
                j = c->ndf2orb[ia] - c->atom_ndf[2 * (ia - 1)];
        ++i;
        c->xpnew[i] = c->xpset[ib] + j;
    }
}

/* the weight scatter shared by ovint/int1e/int2e (valence.F90:2919-2932 etc.) */
static void scatter(vo_ctx *c, int io, double *cf)
{
    for (int i = 1; i <= c->max_obs; ++i) cf[i] = 0.0;
    for (int i = c->map_orbs[io]; i <= c->map_orbs[io + 1] - 1; ++i) {
        if (c->xpset[i] < 1) {
            int indf = c->xpset[i] + c->norbs;
            ndf2obs(c, io, indf);
            int k = 0;
            for (int j = c->map_orbs[indf]; j <= c->map_orbs[indf + 1] - 1; ++j) {
                ++k;
                cf[c->xpnew[k]] = cf[c->xpnew[k]] + c->coeff[j] * c->coeff[i];
            }
        } else {
            cf[c->xpset[i]] = cf[c->xpset[i]] + c->coeff[i];
        }
    }
}

/* valence.F90:2891-3012 (ovint) and :3022-3176 (int1e); kind 0 -> sint, 1 -> hint */
static void oneint(vo_ctx *c, int io, int jo, int kind)
{
    scatter(c, io, c->coeffi);
    scatter(c, jo, c->coeffj);
    double acc = 0.0;
    int loxi = 0;
    for (int ic = 1; ic <= c->orbas_atnum[io]; ++ic) {
        int ia = ATSET(c, ic, io), it = c->atom_t[ia];
        int mnshi = c->map_atom2shell[it], mxshi = mnshi + c->num_shell_atom[it] - 1;
        for (int ii = mnshi; ii <= mxshi; ++ii) {
            int lit = c->ang_mom[ii] + 1, mini = funmin[lit], maxi = funmax[lit];
            int loxj = 0;
            for (int jc = 1; jc <= c->orbas_atnum[jo]; ++jc) {
                int ja = ATSET(c, jc, jo), jt = c->atom_t[ja];
                int mnshj = c->map_atom2shell[jt], mxshj = mnshj + c->num_shell_atom[jt] - 1;
                for (int jj = mnshj; jj <= mxshj; ++jj) {
                    int ljt = c->ang_mom[jj] + 1, minj = funmin[ljt], maxj = funmax[ljt];
                    int iao = loxi, n = 0;
                    for (int i = mini; i <= maxi; ++i) {
                        ++iao;
                        int jao = loxj;
                        for (int j = minj; j <= maxj; ++j) {
                            ++n; ++jao;
                            c->dij[n] = c->angn[i] * c->angn[j] * c->coeffi[iao] * c->coeffj[jao];
                        }
                    }
                    if (kind == 0) {
                        const double *S = ao_block1(c, 0, ii, ia, jj, ja, 0);
                        for (int ij = 1; ij <= n; ++ij) acc = acc + c->dij[ij] * S[ij - 1];
                    } else {
                        const double *T = ao_block1(c, 1, ii, ia, jj, ja, 0);
                        for (int ij = 1; ij <= n; ++ij) acc = acc + c->dij[ij] * T[ij - 1];
                        for (int ictr = 1; ictr <= c->natom; ++ictr) {
                            double nuchrg = c->nuc_charge[c->atom_t[ictr]];
                            if (fabs(nuchrg) > 1.0e-12) {
                                const double *V = ao_block1(c, 2, ii, ia, jj, ja, ictr);
                                for (int ij = 1; ij <= n; ++ij) acc = acc + c->dij[ij] * V[ij - 1];
                            }
                        }
                    }
                    loxj = loxj + shell_size(c->ang_mom[jj]);
                }
            }
            loxi = loxi + shell_size(c->ang_mom[ii]);
        }
    }
    if (kind == 0) c->sint = acc; else c->hint = acc;
}
static void ovint(vo_ctx *c, int io, int jo) { oneint(c, io, jo, 0); }
static void int1e(vo_ctx *c, int io, int jo) { oneint(c, io, jo, 1); }

/* valence.F90:3184-3438 */
static double mono_now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static void int2e(vo_ctx *c, int io, int jo, int ko, int lo)
{
    if (c->deadline > 0.0 && !c->expired && mono_now() > c->deadline) c->expired = 1;
    if (c->expired) { c->gint = 0.0; return; }
    scatter(c, io, c->coeffi); scatter(c, jo, c->coeffj); scatter(c, ko, c->coeffk); scatter(c, lo, c->coeffl);
    if (c->count_only && c->in2e) {
        /* the shell quartets the four nested weight screens below would let through */
        long long prod = 1;
        const int orb[4] = {io, jo, ko, lo};
        const double *cf[4] = {c->coeffi, c->coeffj, c->coeffk, c->coeffl};
        for (int q = 0; q < 4; ++q) {
            long long n = 0;
            int beg = 1;
            for (int ia = 1; ia <= c->orbas_atnum[orb[q]]; ++ia) {
                int it = c->atom_t[ATSET(c, ia, orb[q])];
                for (int ii = c->map_atom2shell[it]; ii <= c->map_atom2shell[it] + c->num_shell_atom[it] - 1; ++ii) {
                    double sum = 0.0;
                    for (int i = beg; i <= beg + shell_size(c->ang_mom[ii]) - 1; ++i) sum = sum + cf[q][i] * cf[q][i];
                    if (sum > c->dtol) ++n;
                    beg = beg + shell_size(c->ang_mom[ii]);
                }
            }
            prod *= n;
        }
        c->cnt.shell_quartets += prod; c->cnt.shell_quartets_2e += prod;
        c->gint = 0.0;
        return;
    }
    double gint = 0.0;
    int ish_beg = 1;
    for (int ia = 1; ia <= c->orbas_atnum[io]; ++ia) {
        int ic = ATSET(c, ia, io), it = c->atom_t[ic];
        int mnshi = c->map_atom2shell[it], mxshi = mnshi + c->num_shell_atom[it] - 1;
        for (int ii = mnshi; ii <= mxshi; ++ii) {
            double sum = 0.0;
            for (int i = ish_beg; i <= ish_beg + shell_size(c->ang_mom[ii]) - 1; ++i) sum = sum + c->coeffi[i] * c->coeffi[i];
            if (sum > c->dtol) {
                int jsh_beg = 1;
                for (int ja = 1; ja <= c->orbas_atnum[jo]; ++ja) {
                    int jc = ATSET(c, ja, jo), jt = c->atom_t[jc];
                    int mnshj = c->map_atom2shell[jt], mxshj = mnshj + c->num_shell_atom[jt] - 1;
                    for (int jj = mnshj; jj <= mxshj; ++jj) {
                        sum = 0.0;
                        for (int i = jsh_beg; i <= jsh_beg + shell_size(c->ang_mom[jj]) - 1; ++i) sum = sum + c->coeffj[i] * c->coeffj[i];
                        if (sum > c->dtol) {
                            int ksh_beg = 1;
                            for (int ka = 1; ka <= c->orbas_atnum[ko]; ++ka) {
                                int kc = ATSET(c, ka, ko), kt = c->atom_t[kc];
                                int mnshk = c->map_atom2shell[kt], mxshk = mnshk + c->num_shell_atom[kt] - 1;
                                for (int kk = mnshk; kk <= mxshk; ++kk) {
                                    sum = 0.0;
                                    for (int i = ksh_beg; i <= ksh_beg + shell_size(c->ang_mom[kk]) - 1; ++i) sum = sum + c->coeffk[i] * c->coeffk[i];
                                    if (sum > c->dtol) {
                                        int lsh_beg = 1;
                                        for (int la = 1; la <= c->orbas_atnum[lo]; ++la) {
                                            int lc = ATSET(c, la, lo), lt = c->atom_t[lc];
                                            int mnshl = c->map_atom2shell[lt], mxshl = mnshl + c->num_shell_atom[lt] - 1;
                                            for (int ll = mnshl; ll <= mxshl; ++ll) {
                                                sum = 0.0;
                                                for (int i = lsh_beg; i <= lsh_beg + shell_size(c->ang_mom[ll]) - 1; ++i) sum = sum + c->coeffl[i] * c->coeffl[i];
                                                if (sum > c->dtol) {
                                                    int lkt = c->ang_mom[kk] + 1, mink = funmin[lkt], maxk = funmax[lkt];
                                                    int llt = c->ang_mom[ll] + 1, minl = funmin[llt], maxl = funmax[llt];
                                                    int lit = c->ang_mom[ii] + 1, mini = funmin[lit], maxi = funmax[lit];
                                                    int ljt = c->ang_mom[jj] + 1, minj = funmin[ljt], maxj = funmax[ljt];
                                                    int kl = 0, kao = 0;
                                                    for (int k = mink; k <= maxk; ++k) {
                                                        ++kao;
                                                        double faci = c->angn[k];
                                                        int lao = 0;
                                                        for (int l = minl; l <= maxl; ++l) {
                                                            ++lao; ++kl;
                                                            c->dkl[kl] = faci * c->angn[l] * c->coeffk[ksh_beg + kao - 1] * c->coeffl[lsh_beg + lao - 1];
                                                        }
                                                    }
                                                    int ij = 0, iao = 0;
                                                    for (int i = mini; i <= maxi; ++i) {
                                                        ++iao;
                                                        double faci = c->angn[i];
                                                        int jao = 0;
                                                        for (int j = minj; j <= maxj; ++j) {
                                                            ++jao; ++ij;
                                                            c->dij[ij] = faci * c->angn[j] * c->coeffi[ish_beg + iao - 1] * c->coeffj[jsh_beg + jao - 1];
                                                        }
                                                    }
                                                    const double *I = ao_eri(c, ii, ic, jj, jc, kk, kc, ll, lc);
                                                    int n = 0;
                                                    for (int i = 1; i <= ij; ++i)
                                                        for (int k = 1; k <= kl; ++k) {
                                                            ++n;
                                                            gint = gint + c->dkl[k] * c->dij[i] * I[n - 1];
                                                        }
                                                }
                                                lsh_beg = lsh_beg + shell_size(c->ang_mom[ll]);
                                            }
                                        }
                                    }
                                    ksh_beg = ksh_beg + shell_size(c->ang_mom[kk]);
                                }
                            }
                        }
                        jsh_beg = jsh_beg + shell_size(c->ang_mom[jj]);
                    }
                }
            }
            ish_beg = ish_beg + shell_size(c->ang_mom[ii]);
        }
    }
    c->gint = gint;
}

/* valence.F90:2157-2182 */
static void normal(vo_ctx *c, int ist, int ind)
{
    for (int i = ist; i <= ind; ++i) {
        ovint(c, i, i);
        c->sint = pow(c->sint, -0.5);
        for (int j = c->map_orbs[i]; j <= c->map_orbs[i + 1] - 1; ++j) c->coeff[j] = c->coeff[j] * c->sint;
    }
}

/* ======================================================================== */
/* determinants and cofactor densities                                       */
/* ======================================================================== */
/* givens.F90:230-262 */
static void givens_single(double *a, int lda, int n, double tol)
{
#define A_(i, j) a[((j) - 1) * lda + ((i) - 1)]
    for (int j = 1; j <= n - 1; ++j)
        for (int i = n; i >= j + 1; --i) {
            double cx = A_(j, i - 1), sx = A_(j, i);
            double r = cx * cx + sx * sx;
            if (fabs(r) > tol) {
                r = pow(r, -0.5);
                cx = cx * r; sx = sx * r;
                for (int k = 1; k <= n; ++k) {
                    double t1 = A_(k, i - 1), t2 = A_(k, i);
                    A_(k, i - 1) = t1 * cx + t2 * sx;
                    A_(k, i) = -t1 * sx + t2 * cx;
                }
            }
        }
#undef A_
}

/* valence.F90:2067-2144 (USE_FORTRAN_VERSION branch) */
static double givdr(vo_ctx *c, int max_n, int n, double *adet, double tol)
{
    c->cnt.determinants++;
    givens_single(adet, max_n, n, tol);
    double d = 1.0;
    for (int i = 1; i <= n; ++i) d = d * adet[(i - 1) * max_n + (i - 1)];
    return d;
}

/* valence.F90:2444-2517 */
static void set_up_unpaired_docc(vo_ctx *c)
{
    int npair = c->npair, nunpd = c->nunpd, ndocc = c->ndocc;
    int dima = npair + ndocc + nunpd, dimb = npair + ndocc, num_sc_elec = npair * 2;
    int nu = ndocc + nunpd;
    int j = 2 * npair + 1;
    for (int i = npair + 1; i <= npair + nunpd; ++i) { c->bra_a[i] = j; c->ket_a[i] = j; ++j; }
    j = 2 * npair + nunpd + 1;
    for (int i = npair + nunpd + 1; i <= npair + nunpd + ndocc; ++i) { c->bra_a[i] = j; c->ket_a[i] = j; j += 2; }
    j = 2 * npair + nunpd + 2;
    for (int i = npair + 1; i <= npair + ndocc; ++i) { c->bra_b[i] = j; c->ket_b[i] = j; j += 2; }
    for (j = 1; j <= num_sc_elec; ++j) {
        for (int i = 1; i <= dima - npair; ++i) c->abra_docc_un[(j - 1) * nu + (i - 1)] = WDET(c, c->bra_a[i + npair], j);
        for (int i = 1; i <= dimb - npair; ++i) c->bbra_docc_un[(j - 1) * nu + (i - 1)] = WDET(c, c->bra_b[i + npair], j);
    }
    for (int i = 1; i <= dima - npair; ++i)
        for (j = 1; j <= dima - npair; ++j) c->aket_docc_un[(i - 1) * nu + (j - 1)] = WDET(c, c->bra_a[j + npair], c->ket_a[i + npair]);
    for (int i = 1; i <= dimb - npair; ++i)
        for (j = 1; j <= dimb - npair; ++j) c->bket_docc_un[(i - 1) * nu + (j - 1)] = WDET(c, c->bra_b[j + npair], c->ket_b[i + npair]);
}

/* valence.F90:2524-2586 */
static void build_abket(vo_ctx *c, int dima, int dimb)
{
    int npair = c->npair, ndocc = c->ndocc, nunpd = c->nunpd, nu = ndocc + nunpd;
    if (npair > 0) {
        for (int k = 1; k <= npair; ++k) {
            for (int l = 1 + npair; l <= dima; ++l) AKET(c, l, k) = c->abra_docc_un[(c->ket_a[k] - 1) * nu + (l - npair - 1)];
            for (int l = 1 + npair; l <= dimb; ++l) BKET(c, l, k) = c->bbra_docc_un[(c->ket_b[k] - 1) * nu + (l - npair - 1)];
        }
        for (int k = 1; k <= dima; ++k)
            for (int l = 1; l <= npair; ++l) AKET(c, l, k) = c->abra_npair[(c->ket_a[k] - 1) * npair + (l - 1)];
        for (int k = 1; k <= dimb; ++k)
            for (int l = 1; l <= npair; ++l) BKET(c, l, k) = c->bbra_npair[(c->ket_b[k] - 1) * npair + (l - 1)];
    }
    for (int k = 1; k <= ndocc; ++k) {
        for (int l = 1 + npair; l <= npair + ndocc; ++l) {
            AKET(c, l, k + npair) = c->aket_docc_un[(k - 1) * nu + (l - npair - 1)];
            BKET(c, l, k + npair) = c->bket_docc_un[(k - 1) * nu + (l - npair - 1)];
        }
        for (int l = 1 + npair + ndocc; l <= npair + nunpd + ndocc; ++l) AKET(c, l, k + npair) = c->aket_docc_un[(k - 1) * nu + (l - npair - 1)];
    }
    for (int k = 1 + ndocc; k <= ndocc + nunpd; ++k)
        for (int l = 1 + npair; l <= npair + ndocc + nunpd; ++l) AKET(c, l, k + npair) = c->aket_docc_un[(k - 1) * nu + (l - npair - 1)];
}

/* valence.F90:2670-2714 */
static void check_spin_and_locate(const vo_ctx *c, int sob, int sok, int *lab, int *lak, int *lbb, int *lbk, int *both_a, int *both_b)
{
    int ba = 0, ka = 0, bb = 0, kb = 0;
    for (int i = 1; i <= c->nalpha; ++i) {
        if (sob == c->bra_a[i]) { ba = 1; *lab = i; }
        if (sok == c->ket_a[i]) { ka = 1; *lak = i; }
    }
    for (int i = 1; i <= c->nbeta; ++i) {
        if (sob == c->bra_b[i]) { bb = 1; *lbb = i; }
        if (sok == c->ket_b[i]) { kb = 1; *lbk = i; }
    }
    *both_a = ba && ka; *both_b = bb && kb;
}

/* valence.F90:1895-2056 */
static void det(vo_ctx *c, int dima, int dimb, int nord, double *density, double *exchanged_density, int calc_dens, int calc_exchange_dens)
{
    int erep_spin_is_nonzero = 0, exchange_spin_is_nonzero = 0;
    double ad = 0.0, bd = 0.0;
    int lab = 0, lak = 0, lbb = 0, lbk = 0, both_a = 0, both_b = 0;
    if (calc_dens) {
        for (int iord = 1; iord <= nord; ++iord) {
            check_spin_and_locate(c, c->dme_b[iord], c->dme_k[iord], &lab, &lak, &lbb, &lbk, &both_a, &both_b);
            if (both_a) {
                for (int i = 1; i <= c->nalpha; ++i) { AKET(c, lab, i) = 0.0; AKET(c, i, lak) = 0.0; }
                AKET(c, lab, lak) = 1.0;
            } else if (both_b) {
                for (int i = 1; i <= c->nbeta; ++i) { BKET(c, lbb, i) = 0.0; BKET(c, i, lbk) = 0.0; }
                BKET(c, lbb, lbk) = 1.0;
            } else break;
        }
        erep_spin_is_nonzero = both_b || both_a;
    }
    int setup_erep = calc_dens && erep_spin_is_nonzero;
    if (calc_exchange_dens) {
        for (int iord = 1; iord <= nord; ++iord) {
            check_spin_and_locate(c, c->dme_b[iord], c->dme_k[nord - iord + 1], &lab, &lak, &lbb, &lbk, &both_a, &both_b);
            if (both_a) {
                if (!setup_erep) {
                    for (int i = 1; i <= c->nalpha; ++i) { AKET(c, lab, i) = 0.0; AKET(c, i, lak) = 0.0; }
                    AKET(c, lab, lak) = 1.0;
                }
            } else if (both_b) {
                if (!setup_erep) {
                    for (int i = 1; i <= c->nbeta; ++i) { BKET(c, lbb, i) = 0.0; BKET(c, i, lbk) = 0.0; }
                    BKET(c, lbb, lbk) = 1.0;
                }
            } else break;
        }
        exchange_spin_is_nonzero = both_a || both_b;
    }
    int setup_exch = calc_exchange_dens && exchange_spin_is_nonzero;
    if (setup_erep || setup_exch) {
        if (dima == 0) ad = 1.0;
        else if (dima == 1) ad = AKET(c, 1, 1);
        else ad = givdr(c, c->padded_size, c->nalpha, c->aket, c->dtol);
        if (dimb == 0) bd = 1.0;
        else if (dimb == 1) bd = BKET(c, 1, 1);
        else bd = givdr(c, c->padded_size, c->nbeta, c->bket, c->dtol);
    }
    if (setup_erep) *density = *density + ad * bd;
    if (setup_exch && !setup_erep) *exchanged_density = *exchanged_density - ad * bd;
    if (setup_exch && setup_erep) *exchanged_density = *exchanged_density + ad * bd;
}

static void fill_npair(vo_ctx *c)
{
    int npair = c->npair;
    for (int l = 1; l <= c->nelec; ++l)
        for (int k = 1; k <= npair; ++k) {
            c->abra_npair[(l - 1) * npair + (k - 1)] = WDET(c, c->bra_a[k], l);
            c->bbra_npair[(l - 1) * npair + (k - 1)] = WDET(c, c->bra_b[k], l);
        }
}

/* valence.F90:1777-1870 */
static void dket(vo_ctx *c, int nord, int nexk, int dima, int dimb, double *ed, double *xd, int jsc, int calc_dens, int calc_x)
{
    int npair = c->npair;
    for (int k = 1; k <= npair; ++k) { c->ket_a[k] = PAIRSC(c, k, 1, jsc); c->ket_b[k] = PAIRSC(c, k, 2, jsc); }
    build_abket(c, dima, dimb);
    det(c, dima, dimb, nord, ed, xd, calc_dens, calc_x);
    for (int jo = 1; jo <= nexk; ++jo) {
        for (int k = 1; k <= jo; ++k) c->kexch[k] = k;
        for (int k = 1; k <= npair; ++k) { c->ket_a[k] = PAIRSC(c, k, 1, jsc); c->ket_b[k] = PAIRSC(c, k, 2, jsc); }
        for (int k = 1; k <= jo; ++k) { int tmp = c->ket_a[k]; c->ket_a[k] = c->ket_b[k]; c->ket_b[k] = tmp; }
        build_abket(c, dima, dimb);
        det(c, dima, dimb, nord, ed, xd, calc_dens, calc_x);
        int j = jo;
        while (c->kexch[1] < 1 + nexk - jo) {
            if (c->kexch[j] < j + nexk - jo) {
                c->kexch[j] = c->kexch[j] + 1;
                for (int k = j + 1; k <= jo; ++k) c->kexch[k] = c->kexch[k - 1] + 1;
                for (int k = 1; k <= npair; ++k) { c->ket_a[k] = PAIRSC(c, k, 1, jsc); c->ket_b[k] = PAIRSC(c, k, 2, jsc); }
                for (int k = 1; k <= jo; ++k) { int l = c->kexch[k]; int tmp = c->ket_a[l]; c->ket_a[l] = c->ket_b[l]; c->ket_b[l] = tmp; }
                build_abket(c, dima, dimb);
                det(c, dima, dimb, nord, ed, xd, calc_dens, calc_x);
                j = jo;
            } else {
                j = j - 1;
            }
        }
    }
}

/* valence.F90:1648-1761 */
static void dbra(vo_ctx *c, int nord, int nexb, int nexk, int dima, int dimb, double *ed, double *xd, int isc, int jsc, int calc_dens, int calc_x)
{
    int npair = c->npair;
    for (int k = 1; k <= npair; ++k) { c->bra_a[k] = PAIRSC(c, k, 1, isc); c->bra_b[k] = PAIRSC(c, k, 2, isc); }
    fill_npair(c);
    dket(c, nord, nexk, dima, dimb, ed, xd, jsc, calc_dens, calc_x);
    for (int io = 1; io <= nexb; ++io) {
        for (int k = 1; k <= io; ++k) c->bexch[k] = k;
        for (int k = 1; k <= npair; ++k) { c->bra_a[k] = PAIRSC(c, k, 1, isc); c->bra_b[k] = PAIRSC(c, k, 2, isc); }
        for (int k = 1; k <= io; ++k) { int tmp = c->bra_a[k]; c->bra_a[k] = c->bra_b[k]; c->bra_b[k] = tmp; }
        fill_npair(c);
        dket(c, nord, nexk, dima, dimb, ed, xd, jsc, calc_dens, calc_x);
        int i = io;
        while (c->bexch[1] < 1 + nexb - io) {
            if (c->bexch[i] < i + nexb - io) {
                c->bexch[i] = c->bexch[i] + 1;
                for (int k = i + 1; k <= io; ++k) c->bexch[k] = c->bexch[k - 1] + 1;
                for (int k = 1; k <= npair; ++k) { c->bra_a[k] = PAIRSC(c, k, 1, isc); c->bra_b[k] = PAIRSC(c, k, 2, isc); }
                for (int k = 1; k <= io; ++k) { int l = c->bexch[k]; int tmp = c->bra_a[l]; c->bra_a[l] = c->bra_b[l]; c->bra_b[l] = tmp; }
                fill_npair(c);
                dket(c, nord, nexk, dima, dimb, ed, xd, jsc, calc_dens, calc_x);
                i = io;
            } else {
                i = i - 1;
            }
        }
    }
}

/* valence.F90:1612-1633 */
static void density_sc(vo_ctx *c, int nord, int isc, int jsc, double *ed, double *xd, int calc_dens, int calc_x)
{
    *ed = 0.0; *xd = 0.0;
    int dima = c->npair + c->ndocc + c->nunpd, dimb = c->npair + c->ndocc;
    dbra(c, nord, c->npair, c->npair, dima, dimb, ed, xd, isc, jsc, calc_dens, calc_x);
}

/* valence.F90:1535-1600 */
static void density(vo_ctx *c, int nord, double *ed, double *xd, double erep_int, double exch_int, int calc_dens, int calc_x)
{
    double eds, xds;
    if (c->spinopt) {
        if (nord == 1) {
            for (int isc = 1; isc <= c->nspinc; ++isc)
                for (int jsc = 1; jsc <= isc; ++jsc) {
                    density_sc(c, nord, isc, jsc, &eds, &xds, calc_dens, calc_x);
                    HAM(c, isc, jsc) = HAM(c, isc, jsc) + eds * c->hint;
                    OVL(c, isc, jsc) = OVL(c, isc, jsc) + eds * c->sint;
                }
        } else if (nord == 2) {
            for (int isc = 1; isc <= c->nspinc; ++isc)
                for (int jsc = 1; jsc <= isc; ++jsc) {
                    density_sc(c, nord, isc, jsc, &eds, &xds, calc_dens, calc_x);
                    HAM(c, isc, jsc) = HAM(c, isc, jsc) + eds * erep_int - xds * exch_int;
                }
        }
    } else {
        if (c->nspinc > 0) {
            *ed = 0.0; *xd = 0.0;
            for (int isc = 1; isc <= c->nspinc; ++isc)
                for (int jsc = 1; jsc <= c->nspinc; ++jsc) {
                    density_sc(c, nord, isc, jsc, &eds, &xds, calc_dens, calc_x);
                    *ed = *ed + eds * c->coeff_sc[isc] * c->coeff_sc[jsc];
                    *xd = *xd + xds * c->coeff_sc[isc] * c->coeff_sc[jsc];
                }
        } else {
            int dima = c->npair + c->ndocc + c->nunpd, dimb = c->npair + c->ndocc;
            *ed = 0.0; *xd = 0.0;
            build_abket(c, dima, dimb);
            det(c, dima, dimb, nord, ed, xd, calc_dens, calc_x);
        }
    }
}

/* ======================================================================== */
/* wfndet, schwarz_ints, vsvb_energy                                         */
/* ======================================================================== */
static int indx(int i, int j) { int ix = i > j ? i : j, mn = i < j ? i : j; return (ix * ix - ix) / 2 + mn; } /* valence.F90:2427-2432 */

/* xm_module.F90:953-966 */
static void xm_dtriang(int ij, int *i, int *j)
{
    int ii = (int)sqrt((double)(2 * ij));
    int k = (ii * ii + ii) / 2;
    while (k < ij) { ii = ii + 1; k = (ii * ii + ii) / 2; }
    *i = ii; *j = ij - (ii * ii - ii) / 2;
}
/* xm_module.F90:969-991 */
static void xm_dtriang8(long long ij, int *i, int *j)
{
    long long i8 = (long long)sqrt((double)(2 * ij));
    long long k = (i8 * i8 + i8) / 2;
    while (k < ij) { i8 = i8 + 1; k = (i8 * i8 + i8) / 2; }
    *i = (int)i8; *j = (int)(ij - (i8 * i8 - i8) / 2);
}

/* valence.F90:1440-1480; the round-robin + all-reduce collapses to a full fill */
static void wfndet(vo_ctx *c)
{
    for (int i = 1; i <= c->nelec; ++i)
        for (int j = 1; j <= c->nelec; ++j) WDET(c, j, i) = 0.0;
    for (int i = 1; i <= c->nelec; ++i)
        for (int j = 1; j <= i; ++j) {
            ovint(c, c->bra[i], c->ket[j]);
            WDET(c, i, j) = c->sint; WDET(c, j, i) = c->sint;
        }
}

/* valence.F90:1489-1523 */
static void schwarz_ints(vo_ctx *c, int num_spatial_orbs, int num_non_docc)
{
    int ij = 0;
    for (int i = 1; i <= num_spatial_orbs; ++i)
        for (int j = 1; j <= i; ++j) { ++ij; c->schwarz[ij] = 0.0; }
    for (int i = 1; i <= num_spatial_orbs; ++i)
        for (int j = 1; j <= i; ++j) {
            int ie = i; if (i > num_non_docc) ie = 2 * i - num_non_docc;
            int je = j; if (j > num_non_docc) je = 2 * j - num_non_docc;
            int2e(c, c->bra[ie], c->ket[je], c->bra[ie], c->ket[je]);
            c->schwarz[indx(i, j)] = sqrt(c->gint);
        }
}

/* valence.F90:1010-1434 */
static void vsvb_energy(vo_ctx *c, int iorb, int num_non_docc, int num_spatial_orbs, double *energy_out, double *wfnorm_out, int spinav, int sym)
{
    const double zero = 0.0;
    long long nproc8 = c->nrank, task = 0;
    double energy = zero, wfnorm = zero, d1, dummy = 0.0;
    int nelec = c->nelec;
    set_up_unpaired_docc(c);

    for (int i = 1; i <= nelec && !c->count_only; ++i) {
        int i_is_docc = i > num_non_docc;
        for (int j = 1; j <= nelec; ++j) {
            int j_is_docc = j > num_non_docc;
            int zero_spin = i_is_docc && j_is_docc && (i % 2) != (j % 2);
            if (!zero_spin) {
                task = task + 1;
                if (task % nproc8 == c->irank) {
                    ovint(c, c->bra[i], c->ket[j]);
                    int1e(c, c->bra[i], c->ket[j]);
                    c->dme_b[1] = i; c->dme_k[1] = j;
                    density(c, 1, &d1, &dummy, dummy, dummy, 1, 0);
                    wfnorm = wfnorm + c->sint * d1;
                    energy = energy + c->hint * d1;
                }
            }
        }
    }
    wfnorm = wfnorm / (double)nelec;

    long long ntasks_ijo = ((long long)num_spatial_orbs * num_spatial_orbs + num_spatial_orbs) / 2;
    long long ntasks_klo = ntasks_ijo, ij8 = ntasks_ijo;
    long long ntasks = ntasks_ijo * ntasks_klo;
    if (sym) ntasks = (ntasks_klo * ntasks_klo + ntasks_klo) / 2;
    long long mytasks = ntasks / nproc8;
    if (c->irank < ntasks % nproc8) mytasks = mytasks + 1;
    if (c->task_limit > 0 && mytasks > c->task_limit) mytasks = c->task_limit;

    int ericount = 0;
    task = c->irank;
    c->in2e = 1;
    for (long long loctask = 1; loctask <= mytasks; ++loctask) {
        int ijorb, klorb, io, jo, ko, lo;
        if (c->expired) break;
        c->cnt.ntasks++;
        if (sym) xm_dtriang8(1 + task, &ijorb, &klorb);
        else { klorb = (int)(1 + task / ij8); ijorb = (int)(1 + task % ij8); }
        xm_dtriang(ijorb, &io, &jo);
        xm_dtriang(klorb, &ko, &lo);
        if ((io == jo && io <= num_non_docc) || (ko == lo && ko <= num_non_docc)) {
            c->cnt.same_orb_skip++;
            task = task + c->nrank;
            continue;
        }
        int erep_is_signif = c->schwarz[indx(io, ko)] * c->schwarz[indx(jo, lo)] > c->itol;
        int exchange_is_signif = c->schwarz[indx(io, lo)] * c->schwarz[indx(jo, ko)] > c->itol;
        if (erep_is_signif) c->cnt.schwarz_erep++;
        if (exchange_is_signif) c->cnt.schwarz_exch++;
        double exchanged_erep_int = zero, erep_int = zero;
        c->gint = zero;
        if (erep_is_signif || exchange_is_signif) {
            c->cnt.schwarz_pass++;
            int jorb = iorb;
            if (iorb > 2 * c->npair + c->nunpd && !c->dem_gs) {
                jorb = 2 * c->npair + c->nunpd + 1;
                if (spinav) jorb = 2 * c->npair + c->nunpd + 2;
            }
            int nonsub = io != jorb && jo != jorb && ko != jorb && lo != jorb;
            if (nonsub && io == jo && ko == lo) {
                erep_int = c->schwarz[indx(io, ko)] * c->schwarz[indx(io, ko)];
                exchanged_erep_int = erep_int;
                c->cnt.shortcut++;
            } else {
                int ie = io; if (io > num_non_docc) ie = 2 * io - num_non_docc - 1;
                int je = jo; if (jo > num_non_docc) je = 2 * jo - num_non_docc - 1;
                int ke = ko; if (ko > num_non_docc) ke = 2 * ko - num_non_docc - 1;
                int le = lo; if (lo > num_non_docc) le = 2 * lo - num_non_docc - 1;
                int calculate_integrals = 0;
                if (c->store_eri) {
                    if (nonsub) {
                        ericount = ericount + 2;
                        if (ericount <= c->nstore) {
                            if (c->eri_stored) {
                                calculate_integrals = 0;
                                exchanged_erep_int = c->eribuf[ericount];
                                erep_int = c->eribuf[ericount - 1];
                                c->cnt.eri_cached++;
                            } else calculate_integrals = 1;
                        } else calculate_integrals = 1;
                    } else calculate_integrals = 1;
                } else calculate_integrals = 1;
                if (calculate_integrals) {
                    if (erep_is_signif) {
                        int2e(c, c->bra[ie], c->ket[ke], c->bra[je], c->ket[le]);
                        c->cnt.int2e_calls++;
                        erep_int = c->gint;
                    }
                    if (exchange_is_signif) {
                        if (erep_is_signif && c->ket[le] == c->ket[ke]) {
                            exchanged_erep_int = c->gint;
                        } else {
                            int2e(c, c->bra[ie], c->ket[le], c->bra[je], c->ket[ke]);
                            c->cnt.int2e_calls++;
                            exchanged_erep_int = c->gint;
                        }
                    }
                    if (c->store_eri && nonsub && ericount <= c->nstore && !c->eri_stored) {
                        c->eribuf[ericount] = exchanged_erep_int;
                        c->eribuf[ericount - 1] = erep_int;
                    }
                }
            }
        }
        erep_is_signif = fabs(erep_int) > c->itol;
        exchange_is_signif = fabs(exchanged_erep_int) > c->itol;
        if (erep_is_signif) c->cnt.value_erep++;
        if (exchange_is_signif) c->cnt.value_exch++;
        double erep_sum = zero, exchanged_erep_sum = zero, erep_density = zero, exchanged_erep_density = zero;
        if (erep_is_signif || exchange_is_signif) {
            int i_is_docc = 0, j_is_docc = 0, k_is_docc = 0, l_is_docc = 0, iso = 1, jso = 1, kso = 1, lso = 1;
            if (io > num_non_docc) { i_is_docc = 1; iso = 2; }
            if (jo > num_non_docc) { j_is_docc = 1; jso = 2; }
            if (ko > num_non_docc) { k_is_docc = 1; kso = 2; }
            if (lo > num_non_docc) { l_is_docc = 1; lso = 2; }
            for (int is = 1; is <= iso; ++is) {
                int i = io; if (io > num_non_docc) i = 2 * io - num_non_docc + is - 2;
                for (int js = 1; js <= jso; ++js) {
                    int j = jo; if (jo > num_non_docc) j = 2 * jo - num_non_docc + js - 2;
                    if (j < i) {
                        for (int ks = 1; ks <= kso; ++ks) {
                            int k = ko; if (ko > num_non_docc) k = 2 * ko - num_non_docc + ks - 2;
                            for (int ls = 1; ls <= lso; ++ls) {
                                int l = lo; if (lo > num_non_docc) l = 2 * lo - num_non_docc + ls - 2;
                                if (l < k) {
                                    int ik0 = i_is_docc && k_is_docc && (i % 2) != (k % 2);
                                    int jl0 = j_is_docc && l_is_docc && (j % 2) != (l % 2);
                                    int compute_erep_term = !(ik0 || jl0);
                                    int il0 = i_is_docc && l_is_docc && (i % 2) != (l % 2);
                                    int jk0 = j_is_docc && k_is_docc && (j % 2) != (k % 2);
                                    int compute_exch_term = !(il0 || jk0);
                                    if ((compute_erep_term && erep_is_signif) || (compute_exch_term && exchange_is_signif)) {
                                        c->dme_b[1] = i; c->dme_k[1] = k; c->dme_b[2] = j; c->dme_k[2] = l;
                                        c->cnt.density2++;
                                        density(c, 2, &erep_density, &exchanged_erep_density, erep_int, exchanged_erep_int, erep_is_signif, exchange_is_signif);
                                        if (erep_is_signif) erep_sum = erep_sum + erep_density * erep_int;
                                        if (exchange_is_signif) exchanged_erep_sum = exchanged_erep_sum + exchanged_erep_density * exchanged_erep_int;
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        if (sym) {
            if (!(ko == io && lo == jo)) { erep_sum = erep_sum * 2.0; exchanged_erep_sum = exchanged_erep_sum * 2.0; }
        }
        energy = energy + erep_sum - exchanged_erep_sum;
        task = task + c->nrank;
    }
    c->in2e = 0;
    *energy_out = energy; *wfnorm_out = wfnorm;
}

/* ======================================================================== */
/* calculate_vsvb_energy set-up/tear-down, guess_energy                      */
/* ======================================================================== */
static void default_lists(vo_ctx *c, int num_non_docc)
{
    for (int i = 1; i <= num_non_docc; ++i) { c->bra[i] = i; c->ket[i] = i; }
    int i = num_non_docc + 1;
    for (int idocc = num_non_docc + 1; idocc <= num_non_docc + c->ndocc; ++idocc) {
        c->bra[i] = idocc; c->ket[i] = idocc; c->bra[i + 1] = idocc; c->ket[i + 1] = idocc; i += 2;
    }
}

static void free_work(vo_ctx *c)
{
    double **dp[] = {&c->angn, &c->ashl, &c->ashi, &c->dij, &c->dkl, &c->coeffi, &c->coeffj, &c->coeffk, &c->coeffl, &c->schwarz,
                     &c->eribuf, &c->ham, &c->ovl, &c->wdet, &c->abra_npair, &c->bbra_npair, &c->abra_docc_un, &c->bbra_docc_un,
                     &c->aket, &c->bket, &c->aket_docc_un, &c->bket_docc_un, &c->integrals_store};
    for (size_t i = 0; i < sizeof dp / sizeof dp[0]; ++i) { free(*dp[i]); *dp[i] = NULL; }
    int **ip[] = {&c->nxyz, &c->atom_ndf, &c->ndf2orb, &c->xpnew, &c->bra_a, &c->bra_b, &c->ket_a, &c->ket_b, &c->bexch, &c->kexch, &c->bra, &c->ket};
    for (size_t i = 0; i < sizeof ip / sizeof ip[0]; ++i) { free(*ip[i]); *ip[i] = NULL; }
}

/* valence.F90:71-182: normalisation + work arrays (everything before guess_energy) */
static void setup_energy(vo_ctx *c)
{
    free_work(c);
    memo_free(c->memo); c->memo = NULL; memo_free(c->memo2); c->memo2 = NULL;  /* geometry may have changed */
    for (int i = 1; i <= c->natom_t; ++i) {
        int mnshi = c->map_atom2shell[i], mxshi = mnshi + c->num_shell_atom[i] - 1;
        for (int j = mnshi; j <= mxshi; ++j) {
            int k = c->map_shell2prim[j];
            norm_prim(c->ang_mom[j], c->map_shell2prim[j + 1] - k, c->exponent + k - 1, c->con_coeff + k - 1, c->unnorm + k - 1);
        }
    }
    c->enucrep = nuclear_repulsion(c);
    int nao_type = 0;
    for (int i = 0; i <= c->nang; ++i) nao_type += ((i + 1) * (i + 2)) / 2;
    c->nxyz = IARR(3 * nao_type); c->angn = DARR(nao_type); c->ashl = DARR(c->nang + 1); c->ashi = DARR(c->nang + 1);
    cartesian(c); setangn(c);
    int ncmax = ((c->nang + 1) * (c->nang + 2)) / 2, mxcf2 = ncmax * ncmax;
    c->dij = DARR(mxcf2); c->dkl = DARR(mxcf2);
    c->integrals_store = DARR(mxcf2 * mxcf2);
    c->max_obs = 0;
    for (int i = 1; i <= c->norbs; ++i) {
        int n = 0;
        for (int j = 1; j <= c->orbas_atnum[i]; ++j) {
            int it = c->atom_t[ATSET(c, j, i)];
            int mnshi = c->map_atom2shell[it], mxshi = mnshi + c->num_shell_atom[it] - 1;
            for (int k = mnshi; k <= mxshi; ++k) n += shell_size(c->ang_mom[k]);
        }
        if (n > c->max_obs) c->max_obs = n;
    }
    c->coeffi = DARR(c->max_obs); c->coeffj = DARR(c->max_obs); c->coeffk = DARR(c->max_obs); c->coeffl = DARR(c->max_obs);
    c->atom_ndf = IARR(2 * c->mxctr); c->ndf2orb = IARR(c->mxctr); c->xpnew = IARR(c->xpmax);
    normal(c, 2 * c->npair + c->nunpd + c->ndocc + 1, c->norbs);
    normal(c, 1, c->norbs - c->ndf);
    c->bra_a = IARR(c->nalpha); c->ket_a = IARR(c->nalpha); c->bra_b = IARR(c->nbeta); c->ket_b = IARR(c->nbeta);
    c->bexch = IARR(c->npair); c->kexch = IARR(c->npair); c->bra = IARR(c->nelec); c->ket = IARR(c->nelec);
    c->wdet = DARR(c->nelec * c->nelec);
    c->padded_size = c->nalpha;
    c->aket = DARR(c->padded_size * c->nalpha); c->bket = DARR(c->padded_size * c->nalpha);
    c->abra_npair = DARR(c->npair * c->nelec); c->bbra_npair = DARR(c->npair * c->nelec);
    int nu = c->ndocc + c->nunpd;
    c->abra_docc_un = DARR(nu * c->npair * 2); c->bbra_docc_un = DARR(nu * c->npair * 2);
    c->aket_docc_un = DARR(nu * nu); c->bket_docc_un = DARR(nu * nu);
    int norbz = 2 * c->npair + c->ndocc + c->nunpd + 1;
    c->schwarz = DARR((norbz * norbz + norbz) / 2);
}

/* valence.F90:309-345 (rank partial sums, no division) */
static void guess_partial(vo_ctx *c, double *energy, double *wfnorm)
{
    int nnd = 2 * c->npair + c->nunpd;
    default_lists(c, nnd);
    wfndet(c);
    int nso = nnd + c->ndocc;
    schwarz_ints(c, nso, nnd);
    c->store_eri = 0;
    vsvb_energy(c, 0, nnd, nso, energy, wfnorm, 0, 1);
}

typedef struct { double enucrep, energy, wfnorm, numerator; vo_counters cnt; } vo_result;

/* guess_energy, summing the reference's round-robin rank partials serially:
 * nrank = 1 is the serial reference; nrank > 1 reproduces the MPI summation
 * order (xm_equalize_scalar, valence.F90:342-343). */
int vo_guess_energy(vo_ctx *c, int nrank, vo_result *out)
{
    setup_energy(c);
    memset(&c->cnt, 0, sizeof c->cnt);
    double esum = 0.0, wsum = 0.0;
    c->nrank = nrank > 0 ? nrank : 1;
    for (int r = 0; r < c->nrank; ++r) {
        double e, w; c->irank = r;
        guess_partial(c, &e, &w);
        esum += e; wsum += w;
    }
    c->nrank = 1; c->irank = 0;
    out->enucrep = c->enucrep; out->numerator = esum; out->wfnorm = wsum;
    out->energy = esum / wsum + c->enucrep;
    out->cnt = c->cnt;
    return 0;
}

/* Task bookkeeping of guess_energy without the integrals and determinants of the 2e loop: the
 * Schwarz table is computed for real (schwarz_ints), then the task loop of vsvb_energy runs with
 * int2e only counting the shell quartets it would evaluate.  Valid counters: ntasks, same_orb_skip,
 * schwarz_*, shortcut, int2e_calls, shell_quartets_2e (the value_* screens need the integrals).
 * Lets the tests check the engine's screening counters at sizes the full oracle cannot reach. */
int vo_count_tasks(vo_ctx *c, vo_counters *out)
{
    setup_energy(c);
    memset(&c->cnt, 0, sizeof c->cnt);
    c->nrank = 1; c->irank = 0;
    int nnd = 2 * c->npair + c->nunpd;
    default_lists(c, nnd);
    int nso = nnd + c->ndocc;
    schwarz_ints(c, nso, nnd);
    c->store_eri = 0;
    c->count_only = 1;
    double e, w;
    vsvb_energy(c, 0, nnd, nso, &e, &w, 0, 1);
    c->count_only = 0;
    *out = c->cnt;
    return 0;
}

/* one rank's partial sums of guess_energy under the reference's round-robin decomposition
 * (task = irank, irank + nrank, ...; valence.F90:1089,1162-1163); the caller sums over ranks
 * (xm_equalize_scalar) and forms energy/wfnorm + enucrep. */
int vo_guess_partial(vo_ctx *c, int irank, int nrank, double *energy, double *wfnorm, double *enucrep)
{
    setup_energy(c);
    memset(&c->cnt, 0, sizeof c->cnt);
    c->nrank = nrank; c->irank = irank;
    guess_partial(c, energy, wfnorm);
    c->nrank = 1; c->irank = 0;
    *enucrep = c->enucrep;
    return 0;
}

/* one rank's share of the 2e loop only, for the timed CPU baseline:
 * returns shell quartets evaluated (simint_compute_eri-equivalent calls) */
long long vo_baseline_sample(vo_ctx *c, int irank, int nrank, long long task_limit, double *energy_partial)
{
    setup_energy(c);
    memset(&c->cnt, 0, sizeof c->cnt);
    int nnd = 2 * c->npair + c->nunpd;
    default_lists(c, nnd);
    wfndet(c);
    schwarz_ints(c, nnd + c->ndocc, nnd);
    c->store_eri = 0;
    c->nrank = nrank; c->irank = irank; c->task_limit = task_limit;
    double e, w;
    long long before = c->cnt.shell_quartets;
    vsvb_energy(c, 0, nnd, nnd + c->ndocc, &e, &w, 0, 1);
    c->nrank = 1; c->irank = 0; c->task_limit = 0;
    if (energy_partial) *energy_partial = e;
    return c->cnt.shell_quartets - before;
}

/* Timed sample of the reference algorithm for the CPU baseline: this rank's share of
 * schwarz_ints and of the 2e loop (round-robin, valence.F90:1162-1163,1513) runs until
 * `seconds` of wall clock are used up.  Returns simint_compute_eri-equivalent calls made and
 * the time they took (set-up excluded).  Memoisation must be off for a faithful timing. */
long long vo_baseline_timed(vo_ctx *c, int irank, int nrank, double seconds, double *elapsed)
{
    setup_energy(c);
    memset(&c->cnt, 0, sizeof c->cnt);
    int nnd = 2 * c->npair + c->nunpd, nso = nnd + c->ndocc;
    default_lists(c, nnd);
    double t0 = mono_now();
    c->deadline = t0 + seconds; c->expired = 0;
    long long before = c->cnt.shell_quartets;
    /* schwarz_ints, this rank's share (valence.F90:1509-1520) */
    {
        int task = 0;
        for (int i = 1; i <= nso && !c->expired; ++i)
            for (int j = 1; j <= i && !c->expired; ++j) {
                task = task + 1;
                if (task % nrank == irank) {
                    int ie = i; if (i > nnd) ie = 2 * i - nnd;
                    int je = j; if (j > nnd) je = 2 * j - nnd;
                    int2e(c, c->bra[ie], c->ket[je], c->bra[ie], c->ket[je]);
                    c->schwarz[indx(i, j)] = sqrt(c->gint);
                }
            }
    }
    if (!c->expired) {
        /* the 2e loop needs the whole table: the other ranks' entries are filled here untimed */
        double pause = mono_now();
        c->deadline = 0.0;
        wfndet(c);   /* only the densities of the 2e loop need it (nelec^2 overlaps: untimed, and skipped when the
                        sample ends inside schwarz_ints, as it does for the large clusters) */
        int task = 0;
        for (int i = 1; i <= nso; ++i)
            for (int j = 1; j <= i; ++j) {
                task = task + 1;
                if (task % nrank != irank) {
                    int ie = i; if (i > nnd) ie = 2 * i - nnd;
                    int je = j; if (j > nnd) je = 2 * j - nnd;
                    long long keep = c->cnt.shell_quartets;
                    int memo = c->memo_on; c->memo_on = 1;
                    int2e(c, c->bra[ie], c->ket[je], c->bra[ie], c->ket[je]);
                    c->memo_on = memo; c->cnt.shell_quartets = keep;
                    c->schwarz[indx(i, j)] = sqrt(c->gint);
                }
            }
        double resume = mono_now();
        t0 += resume - pause;
        c->deadline = t0 + seconds;
        c->store_eri = 0; c->nrank = nrank; c->irank = irank;
        double e, w;
        vsvb_energy(c, 0, nnd, nso, &e, &w, 0, 1);
        c->nrank = 1; c->irank = 0;
    }
    *elapsed = mono_now() - t0;
    c->deadline = 0.0; c->expired = 0;
    return c->cnt.shell_quartets - before;
}

/* exported views for tests */
void vo_get_counters(const vo_ctx *c, vo_counters *out) { *out = c->cnt; }
int vo_nelec(const vo_ctx *c) { return c->nelec; }
int vo_natom(const vo_ctx *c) { return c->natom; }
int vo_norbs(const vo_ctx *c) { return c->norbs; }
int vo_npairs_schwarz(const vo_ctx *c) { int n = 2 * c->npair + c->nunpd + c->ndocc; return (n * n + n) / 2; }
void vo_get_wdet(const vo_ctx *c, double *out) { memcpy(out, c->wdet, sizeof(double) * (size_t)c->nelec * c->nelec); }
void vo_get_schwarz(const vo_ctx *c, double *out, int n) { for (int i = 0; i < n; ++i) out[i] = c->schwarz[i + 1]; }
void vo_get_coeff(const vo_ctx *c, double *out) { for (int i = 0; i < c->totlen; ++i) out[i] = c->coeff[i + 1]; }
double vo_orbital_eri(vo_ctx *c, int io, int jo, int ko, int lo) { int2e(c, io, jo, ko, lo); return c->gint; }
double vo_orbital_ovl(vo_ctx *c, int io, int jo) { ovint(c, io, jo); return c->sint; }
double vo_orbital_h(vo_ctx *c, int io, int jo) { int1e(c, io, jo); return c->hint; }

#include "vo_opt.inc"
#include "vo_fast.c"

void vo_free(vo_ctx *c)
{
    if (!c) return;
    free_work(c); memo_free(c->memo); memo_free(c->memo2);
    free(c->orbset); free(c->atom_t); free(c->coords); free(c->coords_angs); free(c->atom_alias);
    free(c->map_atom2shell); free(c->num_shell_atom); free(c->map_shell2prim); free(c->ang_mom);
    free(c->nuc_charge); free(c->exponent); free(c->con_coeff); free(c->unnorm);
    free(c->orbas_atnum); free(c->orbas_atset); free(c->map_orbs); free(c->xpset); free(c->xorb); free(c->root);
    free(c->coeff); free(c->coeff_in); free(c->pair_sc); free(c->coeff_sc);
    free(c);
}
