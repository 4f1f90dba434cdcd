#include "vb_cofactor.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace vb {
namespace {

struct BlockCof {
    int n = 0, nz = 0;
    double dN = 1.0, piZ = 1.0, pwo[2] = {1.0, 1.0};
    std::vector<double> G;          // n*n, G[r*n+c]
    std::vector<double> uz[2], vz[2];
    double sig_ratio = 1.0;
};

double det_lu(int n, std::vector<double> a)   // row-major copy
{
    double d = 1.0;
    for (int k = 0; k < n; ++k) {
        int piv = k;
        for (int r = k + 1; r < n; ++r)
            if (std::fabs(a[r * n + k]) > std::fabs(a[piv * n + k])) piv = r;
        if (a[piv * n + k] == 0.0) return 0.0;
        if (piv != k) {
            for (int c = 0; c < n; ++c) std::swap(a[k * n + c], a[piv * n + c]);
            d = -d;
        }
        d *= a[k * n + k];
        for (int r = k + 1; r < n; ++r) {
            double f = a[r * n + k] / a[k * n + k];
            for (int c = k + 1; c < n; ++c) a[r * n + c] -= f * a[k * n + c];
        }
    }
    return d;
}

// one-sided Jacobi SVD of a small square matrix, with an orthonormal completion of the null space
void factor_block(int n, const std::vector<double>& M, BlockCof* out)
{
    BlockCof& B = *out;
    B = BlockCof();
    B.n = n;
    if (n == 0) return;
    std::vector<double> W(M), V((size_t)n * n, 0.0), U((size_t)n * n, 0.0), sig(n);
    for (int i = 0; i < n; ++i) V[i * n + i] = 1.0;
    for (int sweep = 0; sweep < 80; ++sweep) {
        bool conv = true;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double al = 0, be = 0, ga = 0;
                for (int r = 0; r < n; ++r) { al += W[r * n + p] * W[r * n + p]; be += W[r * n + q] * W[r * n + q]; ga += W[r * n + p] * W[r * n + q]; }
                if (ga == 0.0 || std::fabs(ga) <= 1e-16 * std::sqrt(al * be)) continue;
                conv = false;
                double zeta = (be - al) / (2.0 * ga);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (int r = 0; r < n; ++r) {
                    double wp = W[r * n + p], wq = W[r * n + q];
                    W[r * n + p] = c * wp - s * wq; W[r * n + q] = s * wp + c * wq;
                    double vp = V[r * n + p], vq = V[r * n + q];
                    V[r * n + p] = c * vp - s * vq; V[r * n + q] = s * vp + c * vq;
                }
            }
        if (conv) break;
    }
    double smax = 0.0;
    for (int p = 0; p < n; ++p) {
        double s2 = 0;
        for (int r = 0; r < n; ++r) s2 += W[r * n + p] * W[r * n + p];
        sig[p] = std::sqrt(s2);
        smax = std::max(smax, sig[p]);
    }
    std::vector<int> Z;
    for (int p = 0; p < n; ++p) {
        // overlaps are O(1) quantities: a singular value below 1e-12 (absolute or relative) is treated
        // through the explicit null-vector terms, which stay exact because sigma_z itself is kept
        if (sig[p] > 1e-12 * std::max(smax, 1.0) && sig[p] > 0.0) {
            for (int r = 0; r < n; ++r) U[r * n + p] = W[r * n + p] / sig[p];
        } else {
            Z.push_back(p);
        }
    }
    double smin = smax;
    for (int p = 0; p < n; ++p)
        if (std::find(Z.begin(), Z.end(), p) == Z.end()) smin = std::min(smin, sig[p]);
    B.sig_ratio = smax > 0 ? smin / smax : 0.0;
    B.nz = (int)Z.size();
    if (B.nz > 2) { B.dN = 0.0; B.piZ = 0.0; B.nz = 0; B.G.assign((size_t)n * n, 0.0); return; }   // every cofactor up to 2nd order vanishes
    // complete U on the null columns
    std::vector<int> done;
    for (int p = 0; p < n; ++p)
        if (std::find(Z.begin(), Z.end(), p) == Z.end()) done.push_back(p);
    for (int z : Z) {
        double best = -1.0;
        std::vector<double> bestv(n);
        for (int k = 0; k < n; ++k) {
            std::vector<double> v(n, 0.0);
            v[k] = 1.0;
            for (int pass = 0; pass < 2; ++pass)
                for (int p : done) {
                    double dot = 0;
                    for (int r = 0; r < n; ++r) dot += U[r * n + p] * v[r];
                    for (int r = 0; r < n; ++r) v[r] -= dot * U[r * n + p];
                }
            double nn = 0;
            for (int r = 0; r < n; ++r) nn += v[r] * v[r];
            if (nn > best) { best = nn; bestv = v; }
        }
        double inv = 1.0 / std::sqrt(best);
        for (int r = 0; r < n; ++r) U[r * n + z] = bestv[r] * inv;
        done.push_back(z);
    }
    double sgn = det_lu(n, U) * det_lu(n, V);
    sgn = sgn >= 0 ? 1.0 : -1.0;
    B.dN = sgn;
    for (int p = 0; p < n; ++p)
        if (std::find(Z.begin(), Z.end(), p) == Z.end()) B.dN *= sig[p];
    B.piZ = 1.0;
    for (int z : Z) B.piZ *= sig[z];
    if (B.nz == 1) { B.pwo[0] = 1.0; }
    if (B.nz == 2) { B.pwo[0] = sig[Z[1]]; B.pwo[1] = sig[Z[0]]; }
    B.G.assign((size_t)n * n, 0.0);
    for (int p = 0; p < n; ++p) {
        if (std::find(Z.begin(), Z.end(), p) != Z.end()) continue;
        double is = 1.0 / sig[p];
        for (int r = 0; r < n; ++r) {
            double u = U[r * n + p] * is;
            for (int c = 0; c < n; ++c) B.G[(size_t)r * n + c] += u * V[c * n + p];
        }
    }
    for (int k = 0; k < B.nz; ++k) {
        B.uz[k].resize(n); B.vz[k].resize(n);
        for (int r = 0; r < n; ++r) { B.uz[k][r] = U[r * n + Z[k]]; B.vz[k][r] = V[r * n + Z[k]]; }
    }
}

}  // namespace

void build_cofactors(const Input& in, const Wavefunction& wf, const std::vector<double>& Se, CofactorSet* out, int only_isc, int only_jsc)
{
    CofactorSet& cs = *out;
    cs = CofactorSet();
    const int nso = wf.nso, nelec = (int)wf.bra.size();
    cs.nso = nso;
    std::vector<int> entry_of_slot(nelec);
    for (int s = 0; s < nso; ++s)
        for (int k = 0; k < wf.nslots(s); ++k) entry_of_slot[wf.slot(s, k)] = s;
    const int npair = in.npair, nunpd = in.nunpd, ndocc = in.ndocc;
    const int nsc = npair > 0 ? std::max(1, in.nspinc) : 1;
    const size_t stride = cof_stride(nso);
    const long long nmask = 1LL << npair;
    // fixed parts of the spin lists (set_up_unpaired_docc, valence.F90:2461-2480), 0-based slots
    std::vector<int> a_fixed, b_fixed;
    for (int i = 0; i < nunpd; ++i) a_fixed.push_back(2 * npair + i);
    for (int d = 0; d < ndocc; ++d) { a_fixed.push_back(2 * npair + nunpd + 2 * d); b_fixed.push_back(2 * npair + nunpd + 2 * d + 1); }
    BlockCof Ba, Bb;
    for (int isc = 0; isc < nsc; ++isc)
        for (int jsc = 0; jsc < nsc; ++jsc) {
            if (only_isc >= 0 && (isc != only_isc || jsc != only_jsc)) continue;
            for (long long bm = 0; bm < nmask; ++bm)
                for (long long km = 0; km < nmask; ++km) {
                    std::vector<int> abra, bbra, aket, bket;
                    for (int k = 0; k < npair; ++k) {
                        int b1 = in.pair(isc, k, 0) - 1, b2 = in.pair(isc, k, 1) - 1;
                        int k1 = in.pair(jsc, k, 0) - 1, k2 = in.pair(jsc, k, 1) - 1;
                        if ((bm >> k) & 1) std::swap(b1, b2);
                        if ((km >> k) & 1) std::swap(k1, k2);
                        abra.push_back(b1); bbra.push_back(b2); aket.push_back(k1); bket.push_back(k2);
                    }
                    abra.insert(abra.end(), a_fixed.begin(), a_fixed.end()); aket.insert(aket.end(), a_fixed.begin(), a_fixed.end());
                    bbra.insert(bbra.end(), b_fixed.begin(), b_fixed.end()); bket.insert(bket.end(), b_fixed.begin(), b_fixed.end());
                    const int na = (int)abra.size(), nb = (int)bbra.size();
                    std::vector<double> Ma((size_t)na * na), Mb((size_t)nb * nb);
                    for (int r = 0; r < na; ++r)
                        for (int c = 0; c < na; ++c) Ma[(size_t)r * na + c] = Se[(size_t)entry_of_slot[abra[r]] * nso + entry_of_slot[aket[c]]];
                    for (int r = 0; r < nb; ++r)
                        for (int c = 0; c < nb; ++c) Mb[(size_t)r * nb + c] = Se[(size_t)entry_of_slot[bbra[r]] * nso + entry_of_slot[bket[c]]];
                    factor_block(na, Ma, &Ba);
                    factor_block(nb, Mb, &Bb);
                    cs.min_sigma_ratio = std::min(cs.min_sigma_ratio, std::min(Ba.sig_ratio, Bb.sig_ratio));
                    cs.singular_blocks += (Ba.nz > 0 || Ba.dN == 0.0) + (Bb.nz > 0 || Bb.dN == 0.0);
                    cs.data.resize(cs.data.size() + stride, 0.0);
                    double* D = cs.data.data() + (size_t)cs.ndp * stride;
                    const double w = (npair > 0 && in.nspinc > 0) ? in.coeff_sc[isc] * in.coeff_sc[jsc] : 1.0;
                    D[0] = w;
                    D[1] = Ba.dN; D[2] = Ba.piZ; D[3] = Ba.pwo[0]; D[4] = Ba.pwo[1]; D[5] = Ba.nz;
                    D[6] = Bb.dN; D[7] = Bb.piZ; D[8] = Bb.pwo[0]; D[9] = Bb.pwo[1]; D[10] = Bb.nz;
                    double* Ga = D + COF_HEADER;
                    double* Gb = Ga + (size_t)nso * nso;
                    double* uza = Gb + (size_t)nso * nso;
                    double* vza = uza + 2 * nso;
                    double* uzb = vza + 2 * nso;
                    double* vzb = uzb + 2 * nso;
                    for (int r = 0; r < na; ++r)
                        for (int c = 0; c < na; ++c) Ga[(size_t)entry_of_slot[abra[r]] * nso + entry_of_slot[aket[c]]] += Ba.G[(size_t)r * na + c];
                    for (int r = 0; r < nb; ++r)
                        for (int c = 0; c < nb; ++c) Gb[(size_t)entry_of_slot[bbra[r]] * nso + entry_of_slot[bket[c]]] += Bb.G[(size_t)r * nb + c];
                    for (int z = 0; z < Ba.nz; ++z)
                        for (int r = 0; r < na; ++r) { uza[z * nso + entry_of_slot[abra[r]]] += Ba.uz[z][r]; vza[z * nso + entry_of_slot[aket[r]]] += Ba.vz[z][r]; }
                    for (int z = 0; z < Bb.nz; ++z)
                        for (int r = 0; r < nb; ++r) { uzb[z * nso + entry_of_slot[bbra[r]]] += Bb.uz[z][r]; vzb[z * nso + entry_of_slot[bket[r]]] += Bb.vz[z][r]; }
                    if (only_isc >= 0) D[0] = 1.0;
                    cs.ndp++;
                }
        }
}

void one_electron_from_cofactors(const CofactorSet& cs, const std::vector<double>& Se, const std::vector<double>& He, int nelec,
                                 double* e1, double* wfnorm)
{
    const int nso = cs.nso;
    const size_t stride = cof_stride(nso);
    double e = 0.0, nrm = 0.0;
    for (int d = 0; d < cs.ndp; ++d) {
        const double* D = cs.data.data() + (size_t)d * stride;
        const double w = D[0], dNa = D[1], pZa = D[2], dNb = D[6], pZb = D[7];
        const int nza = (int)D[5], nzb = (int)D[10];
        const double* Ga = D + COF_HEADER;
        const double* Gb = Ga + (size_t)nso * nso;
        const double* uza = Gb + (size_t)nso * nso;
        const double* vza = uza + 2 * nso;
        const double* uzb = vza + 2 * nso;
        const double* vzb = uzb + 2 * nso;
        const double c0a = dNa * pZa, c0b = dNb * pZb;
        double es = 0.0, ns = 0.0;
        for (int s = 0; s < nso; ++s)
            for (int t = 0; t < nso; ++t) {
                double c1a = pZa * Ga[(size_t)s * nso + t], c1b = pZb * Gb[(size_t)s * nso + t];
                for (int z = 0; z < nza; ++z) c1a += D[3 + z] * uza[z * nso + s] * vza[z * nso + t];
                for (int z = 0; z < nzb; ++z) c1b += D[8 + z] * uzb[z * nso + s] * vzb[z * nso + t];
                double c1 = dNa * c1a * c0b + dNb * c1b * c0a;
                es += He[(size_t)s * nso + t] * c1;
                ns += Se[(size_t)s * nso + t] * c1;
            }
        e += w * es;
        nrm += w * ns;
    }
    *e1 = e;
    *wfnorm = nrm / (double)nelec;
}

}  // namespace vb
