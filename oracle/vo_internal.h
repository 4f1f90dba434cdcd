/* ORACLE -- TEST INFRASTRUCTURE ONLY (see valence_oracle.c). */
#ifndef VO_INTERNAL_H
#define VO_INTERNAL_H

typedef struct {
    int l, nprim;
    const double *exps, *coef; /* coef = VALENCE's normalised con_coeff */
    double r[3];
} vo_shell;

void vo_boys(int mmax, double T, double *F);
int vo_ncart(int l);
void vo_cart(int l, int idx, int *lx, int *ly, int *lz);
void vo_overlap_block(const vo_shell *A, const vo_shell *B, double *out);
void vo_kinetic_block(const vo_shell *A, const vo_shell *B, double *out);
void vo_potential_block(const vo_shell *A, const vo_shell *B, double Z, const double C[3], double *out);
void vo_eri_block(const vo_shell *A, const vo_shell *B, const vo_shell *C, const vo_shell *D, double *out);

#endif
