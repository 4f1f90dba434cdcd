// Record-based reader for VALENCE input files (see vb_input.h for the grammar
// citations).  Fortran list-directed semantics: each READ starts on a fresh
// record, may continue over following records, and drops the rest of the last
// record it touched; blanks and commas separate items; D exponents and r*c
// repeats are accepted.
#include "vb_input.h"

#include <cstdlib>
#include <fstream>
#include <sstream>

namespace vb {
namespace {

class Records {
public:
    explicit Records(const std::string& text)
    {
        std::string line;
        std::istringstream is(text);
        while (std::getline(is, line)) lines_.push_back(line);
    }
    void begin() { toks_.clear(); it_ = 0; }   // start of one READ statement
    int geti() { return (int)std::strtol(next().c_str(), nullptr, 10); }
    double getd()
    {
        std::string t = next();
        for (auto& ch : t)
            if (ch == 'D' || ch == 'd') ch = 'e';
        return std::strtod(t.c_str(), nullptr);
    }

private:
    const std::string& next()
    {
        while (it_ >= toks_.size()) {
            if (pos_ >= lines_.size()) throw InputError("input ended inside a READ");
            split(lines_[pos_++]);
        }
        return toks_[it_++];
    }
    void split(const std::string& line)
    {
        size_t i = 0, n = line.size();
        auto sep = [](char c) { return c == ' ' || c == '\t' || c == ',' || c == '\r'; };
        while (i < n) {
            while (i < n && sep(line[i])) ++i;
            if (i >= n) break;
            size_t s = i;
            while (i < n && !sep(line[i])) ++i;
            std::string tok = line.substr(s, i - s);
            if (tok == "/") return;
            size_t star = tok.find('*');
            if (star != std::string::npos && star > 0) {
                int rep = std::atoi(tok.substr(0, star).c_str());
                for (int r = 0; r < rep; ++r) toks_.push_back(tok.substr(star + 1));
            } else {
                toks_.push_back(tok);
            }
        }
    }
    std::vector<std::string> lines_, toks_;
    size_t pos_ = 0, it_ = 0;
};

}  // namespace

Input parse_input_text(const std::string& text)
{
    Records r(text);
    Input in;
    r.begin();
    in.natom = r.geti(); in.natom_t = r.geti(); in.npair = r.geti(); in.nunpd = r.geti(); in.ndocc = r.geti();
    in.totlen = r.geti(); in.xpmax = r.geti(); in.nspinc = r.geti(); in.num_sh = r.geti(); in.num_pr = r.geti();
    in.nang = r.geti(); in.ndf = r.geti(); in.nset = r.geti(); in.nxorb = r.geti(); in.mxctr = r.geti();
    if (in.npair > 0 && in.nspinc < 1) throw InputError("no spin couplings");   // valence_initialize_module.F90:58
    if (in.natom < 1 || in.natom_t < 1 || in.npair < 0 || in.nunpd < 0 || in.ndocc < 0 || in.ndf < 0 || in.nset < 0 ||
        in.nxorb < 0)
        throw InputError("bad header");

    r.begin();
    in.ntol_c = r.geti(); in.ntol_d = r.geti(); in.ntol_i = r.geti();
    in.ntol_e_min = r.geti(); in.ntol_e_max = r.geti(); in.max_iter = r.geti();
    in.ptbnmax = r.getd(); in.feather = r.getd();
    for (int i = 0; i < 2 * in.nset; ++i) in.orbset.push_back(r.geti());

    for (int i = 0; i < in.natom; ++i) {
        r.begin();
        in.atom_t.push_back(r.geti());
        for (int d = 0; d < 3; ++d) in.coords.push_back(r.getd());
        if (in.atom_t.back() < 1 || in.atom_t.back() > in.natom_t) throw InputError("atom type out of range");
    }
    for (int t = 0; t < in.natom_t; ++t) {
        AtomTypeDef ty;
        r.begin();
        ty.charge = r.getd();
        int nshell = r.geti();
        for (int j = 0; j < nshell; ++j) {
            ShellDef sh;
            r.begin();
            sh.l = r.geti();
            int k = r.geti();
            if (k == 1) {
                r.begin();
                sh.exps.push_back(r.getd());
                sh.raw.push_back(1.0);
            } else {
                for (int g = 0; g < k; ++g) {
                    r.begin();
                    sh.exps.push_back(r.getd());
                    sh.raw.push_back(r.getd());
                }
            }
            ty.shells.push_back(sh);
        }
        in.types.push_back(ty);
    }
    in.coeff_sc.assign(1, 1.0);
    if (in.npair > 0) {
        r.begin();
        if (in.nspinc == 1) {
            for (int i = 0; i < 2 * in.npair; ++i) in.pair_sc.push_back(r.geti());
        } else {
            in.coeff_sc.clear();
            for (int j = 0; j < in.nspinc; ++j) {
                in.coeff_sc.push_back(r.getd());
                for (int i = 0; i < 2 * in.npair; ++i) in.pair_sc.push_back(r.geti());
            }
        }
    }
    if (in.nxorb > 0) {
        r.begin();
        for (int i = 0; i < in.nxorb; ++i) { in.xorb.push_back(r.geti()); in.root.push_back(r.geti()); }
    }
    for (int o = 0; o < in.norbs(); ++o) {
        OrbitalDef od;
        r.begin();
        int atnum = r.geti();
        for (int k = 0; k < atnum; ++k) od.atoms.push_back(r.geti());
        int n = r.geti();
        r.begin();
        for (int k = 0; k < n; ++k) { od.xp.push_back(r.geti()); od.coeff.push_back(r.getd()); }
        for (int a : od.atoms)
            if (a < 1 || a > in.natom) throw InputError("orbital atom out of range");
        in.orbitals.push_back(od);
    }
    return in;
}

Input parse_input_file(const std::string& path)
{
    std::ifstream fh(path);
    if (!fh) throw InputError("problems opening input file");   // valence_initialize_module.F90:53
    std::stringstream ss;
    ss << fh.rdbuf();
    return parse_input_text(ss.str());
}

}  // namespace vb
