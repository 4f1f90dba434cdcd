// C-ABI of libvalence_b200.so (see include/valence_b200.h for the contract and the reference
// lines each entry point replaces).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/valence_b200.h"
#include "vb_engine.h"
#include "vb_nccl.h"

struct vb_engine {
    std::unique_ptr<vb::Engine> eng;
};

namespace {

thread_local std::string g_err;

void to_c(const vb::EnergyResult& r, vb_energy_result* o)
{
    o->energy = r.energy; o->enucrep = r.enucrep; o->numerator = r.numerator; o->wfnorm = r.wfnorm; o->e1 = r.e1; o->e2 = r.e2;
    for (int i = 0; i < VB_CNT_N; ++i) o->counters[i] = r.counters[i];
    o->n_entries = r.n_entries; o->n_groups = r.n_groups; o->n_pairgroups = r.n_pairgroups; o->n_tiles = r.n_tiles;
    o->n_tiles_mine = r.n_tiles_mine; o->n_ao_quartets = r.n_ao_quartets; o->n_prim_quartets = r.n_prim_quartets;
    o->flops_model = r.flops_model; o->ref_shell_quartets = r.ref_shell_quartets;
    o->t_total_ms = r.t_total; o->t_host_setup_ms = r.t_host_setup; o->t_1e_ms = r.t_1e; o->t_density_ms = r.t_density;
    o->t_diag_ms = r.t_diag; o->t_tiles_ms = r.t_tiles; o->launches = r.launches; o->diag_launches = r.diag_launches;
    o->tile_launches = r.tile_launches; o->min_pivot_ratio = r.min_pivot_ratio; o->h2d_bytes = r.h2d_bytes; o->d2h_bytes = r.d2h_bytes;
    o->flops_transform = r.flops_transform;
}

template <class F>
int guarded(F&& f)
{
    try {
        f();
        return 0;
    } catch (const std::exception& ex) {
        g_err = ex.what();
        return 1;
    } catch (...) {
        g_err = "unknown error";
        return 1;
    }
}

// one live instance per process, like the reference's module globals (valence_finalize_module.F90:23-52)
std::unique_ptr<vb::Engine> g_api;

// xm_abort (xm_module.F90:942-950): `error` line from this rank, then stop
[[noreturn]] void api_abort(const std::string& msg)
{
    std::printf("%-40s from rank %8d\n", msg.c_str(), g_api ? g_api->comm_rank() : vb::launch_env().rank);
    std::fflush(stdout);
    std::exit(1);
}

std::string host_argv1()
{
    if (const char* env = std::getenv("VALENCE_INPUT")) return env;
    std::ifstream f("/proc/self/cmdline", std::ios::binary);
    std::string all((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    size_t z = all.find('\0');
    if (z == std::string::npos || z + 1 >= all.size()) return "";
    return std::string(all.c_str() + z + 1);
}

// Fortran edit descriptor 1e10.2 (0.dddE+xx, right-justified in 10 columns)
std::string fortran_e10_2(double x)
{
    char buf[32];
    if (x == 0.0) return "  0.00E+00";
    int ex = (int)std::floor(std::log10(std::fabs(x))) + 1;
    double m = std::fabs(x) / std::pow(10.0, ex);
    long d = std::lround(m * 100.0);
    if (d >= 100) { d = 10; ex += 1; }
    std::snprintf(buf, sizeof buf, "%s0.%02ldE%c%02d", x < 0 ? "-" : "", d, ex < 0 ? '-' : '+', std::abs(ex));
    std::string out(buf);
    while (out.size() < 10) out.insert(out.begin(), ' ');
    return out;
}

// xm_output (xm_module.F90:464-577): `orbitals` (formats 1, 2, 5, 7 at :571-577) and, for more than one spin
// coupling, `nelecwfn` (formats 3, 4; :479-505).  mode 'done' adds the convergence line (:561).
void write_orbitals_file(const vb::Input& in, const std::vector<std::vector<double>>* coeff, const std::vector<double>* coeff_sc,
                         double energy, bool done = false, double etol = 0.0)
{
    if (in.nspinc > 1) {
        if (FILE* fn = std::fopen("nelecwfn", "w")) {
            for (int j = 0; j < in.nspinc; ++j) {
                const double w = (coeff_sc && (int)coeff_sc->size() > j) ? (*coeff_sc)[j] : in.coeff_sc[j];
                std::fprintf(fn, " %13.8f", w);                                   // format 4: 1x,1f13.8,7(2i3,2x)
                for (int k = 0; k < in.npair; ++k) {
                    if (k > 0 && k % 7 == 0) std::fprintf(fn, "\n");              // format reversion: new record from 7(2i3,2x)
                    std::fprintf(fn, "%3d%3d  ", in.pair(j, k, 0), in.pair(j, k, 1));
                }
                std::fprintf(fn, "\n");
            }
            std::fprintf(fn, " ");                                                // format 3: 1x,7(2i3,2x)
            for (int i = 0; i < in.nxorb; ++i) {
                if (i > 0 && i % 7 == 0) std::fprintf(fn, "\n");
                std::fprintf(fn, "%3d%3d  ", in.xorb[i], in.root[i]);
            }
            std::fprintf(fn, "\n\n");
            std::fclose(fn);
        }
    }
    FILE* fh = std::fopen("orbitals", "w");
    if (!fh) return;
    const int nval = in.norbs() - in.ndf;
    for (int o = 0; o < in.norbs(); ++o) {
        if (o == nval) std::fprintf(fh, "\n");
        const vb::OrbitalDef& od = in.orbitals[o];
        std::fprintf(fh, "     %2d     ", (int)od.atoms.size());
        for (int a : od.atoms) std::fprintf(fh, "%4d", a);
        std::fprintf(fh, "%4d\n", (int)od.xp.size());
        for (size_t k = 0; k < od.xp.size(); ++k) {
            double c = coeff ? (*coeff)[o][k] : od.coeff[k];
            std::fprintf(fh, "%4d %13.8f", od.xp[k], c);
            if (k % 4 == 3 || k + 1 == od.xp.size()) std::fprintf(fh, "\n");
        }
    }
    if (in.ndf == 0) std::fprintf(fh, "\n");
    std::fprintf(fh, "\n total energy in atomic units %32.16f\n", energy);
    if (done) std::fprintf(fh, " converged to %s kCal/mol\n", fortran_e10_2(etol).c_str());
    std::fprintf(fh, "\n");
    std::fclose(fh);
}

}  // namespace

extern "C" {

const char* vb_last_error(void) { return g_err.c_str(); }

int vb_engine_create(const char* input_path, int device, vb_engine** out)
{
    *out = nullptr;
    return guarded([&] {
        vb::Input in = vb::parse_input_file(input_path);
        std::unique_ptr<vb_engine> h(new vb_engine);
        h->eng.reset(new vb::Engine(in, device));
        *out = h.release();
    });
}

void vb_engine_destroy(vb_engine* e) { delete e; }
int vb_engine_natom(const vb_engine* e) { return e->eng->input().natom; }
int vb_engine_norbas(const vb_engine* e, int iorb)
{
    const vb::Input& in = e->eng->input();
    return (iorb >= 1 && iorb <= in.norbs()) ? (int)in.orbitals[iorb - 1].xp.size() : -1;
}
int vb_engine_attach_comm(vb_engine* e, int rank, int nranks, const char* key)
{
    return guarded([&] { e->eng->attach_comm(rank, nranks, key && *key ? key : vb::launch_env().key); });
}
int vb_engine_attach_nccl(vb_engine* e, int rank, int nranks, void* nccl_comm)
{
    return guarded([&] { e->eng->attach_nccl(rank, nranks, nccl_comm); });
}
int vb_engine_nelec(const vb_engine* e) { return e->eng->input().nelec(); }

int vb_engine_set_coords(vb_engine* e, const double* x)
{
    return guarded([&] { e->eng->set_coords_angstrom(x); });
}

int vb_engine_energy(vb_engine* e, vb_energy_result* out)
{
    return guarded([&] {
        vb::EnergyResult r;
        e->eng->energy(&r);
        to_c(r, out);
    });
}

int vb_engine_energy_partial(vb_engine* e, int rank, int nranks, vb_energy_result* out)
{
    return guarded([&] {
        if (nranks < 1 || rank < 0 || rank >= nranks) throw std::runtime_error("bad rank / nranks");
        vb::EnergyResult r;
        e->eng->energy_partial(rank, nranks, &r);
        to_c(r, out);
    });
}

int vb_engine_shard_tables(vb_engine* e, int rank, int nranks, const char* prefix)
{
    return guarded([&] { e->eng->shard_tables(rank, nranks, prefix ? prefix : ""); });
}

int vb_engine_energy_finish(vb_engine* e, vb_energy_result* out)
{
    return guarded([&] {
        vb::EnergyResult r;
        // carry the partial's workload fields through
        r.n_entries = out->n_entries; r.n_groups = out->n_groups; r.n_pairgroups = out->n_pairgroups; r.n_tiles = out->n_tiles;
        r.n_tiles_mine = out->n_tiles_mine; r.n_ao_quartets = out->n_ao_quartets; r.n_prim_quartets = out->n_prim_quartets;
        r.flops_model = out->flops_model; r.flops_transform = out->flops_transform; r.t_host_setup = out->t_host_setup_ms; r.t_1e = out->t_1e_ms; r.t_density = out->t_density_ms;
        r.t_diag = out->t_diag_ms; r.t_tiles = out->t_tiles_ms; r.diag_launches = out->diag_launches; r.tile_launches = out->tile_launches;
        r.min_pivot_ratio = out->min_pivot_ratio; r.enucrep = out->enucrep; r.e1 = out->e1; r.wfnorm = out->wfnorm;
        e->eng->energy_finish(&r);
        to_c(r, out);
    });
}

int vb_engine_first_order(vb_engine* e, int iorb, double* ham, double* ovl, int cap, int* norbas, vb_energy_result* stats)
{
    return guarded([&] {
        std::vector<double> h, s;
        vb::EnergyResult r;
        const int nb = vb_engine_norbas(e, iorb);
        if (nb < 0) throw std::runtime_error("first_order: orbital index out of range");
        *norbas = nb;
        if (nb * nb > cap) throw std::runtime_error("first_order: output buffers too small (vb_engine_norbas gives the order)");
        int n = e->eng->first_order(iorb - 1, &h, &s, &r);
        std::copy(h.begin(), h.end(), ham);
        std::copy(s.begin(), s.end(), ovl);
        if (stats) to_c(r, stats);
    });
}

int vb_engine_first_order_sharded(vb_engine* e, int iorb, int rank, int nranks, double* ham, double* ovl, int cap, int* norbas,
                                  vb_energy_result* stats)
{
    return guarded([&] {
        if (nranks < 1 || rank < 0 || rank >= nranks) throw std::runtime_error("first_order: bad rank / nranks");
        std::vector<double> h, s;
        vb::EnergyResult r;
        const int nb = vb_engine_norbas(e, iorb);
        if (nb < 0) throw std::runtime_error("first_order: orbital index out of range");
        *norbas = nb;
        if (nb * nb > cap) throw std::runtime_error("first_order: output buffers too small (vb_engine_norbas gives the order)");
        int n = e->eng->first_order(iorb - 1, &h, &s, &r, rank, nranks);
        (void)n;
        std::copy(h.begin(), h.end(), ham);
        std::copy(s.begin(), s.end(), ovl);
        if (stats) to_c(r, stats);
    });
}

int vb_engine_run(vb_engine* e, int print, double* enucrep, double* guess_energy, double* total_energy, int* converged, int* iterations)
{
    return guarded([&] {
        vb::Engine::RunResult r;
        e->eng->run(&r, print != 0);
        *enucrep = r.enucrep; *guess_energy = r.guess_energy; *total_energy = r.total_energy;
        *converged = r.converged; *iterations = r.iterations;
    });
}

double* vb_engine_accum_device(const vb_engine* e) { return e->eng->accum_device(); }
int vb_engine_accum_len(const vb_engine* e) { return e->eng->accum_len(); }
void* vb_engine_stream(const vb_engine* e) { return e->eng->stream(); }

long long vb_engine_debug_tile_energies(const vb_engine* e, double* out, long long cap) { return e->eng->debug_tile_energies(out, cap); }

int vb_measure_fp64_peak(int device, double* tflops)
{
    return guarded([&] { *tflops = vb::measure_fp64_peak_tflops(device); });
}

/* ---- reference-compatible layer ------------------------------------------------------------- */
void valence_api_initialize_(int* info, int* call_mpi_init, int* comm)
{
    // The reference either initialises MPI itself or adopts the host's communicator (valence_api.F90:22-27,
    // xm_module.F90:727-732).  Here ranks are one process per GPU of a node: rank / size / local rank are taken from the
    // launcher's environment (torchrun, mpirun, srun) and the ranks form an NCCL communicator; an MPI handle in `comm`
    // cannot be used (there is no MPI in this library) and is only accepted.
    (void)call_mpi_init; (void)comm;
    std::string path = host_argv1();
    if (path.empty()) api_abort("must have one input file");            // valence_initialize_module.F90:50
    try {
        vb::Input in = vb::parse_input_file(path);
        const vb::LaunchEnv le = vb::launch_env();
        g_api.reset(new vb::Engine(in, le.local_rank));
        if (le.nranks > 1) {
            g_api->attach_comm(le.rank, le.nranks, le.key);
            if (le.rank == 0) std::printf(" %-32s  %8d\n", "number of processors", le.nranks);   // xm_propagate, xm_module.F90:744-750
        }
    } catch (const vb::InputError& ex) {
        api_abort(ex.what());
    } catch (const std::exception& ex) {
        api_abort(ex.what());
    }
    *info = 0;                                                           // valence_api.F90:31
}

void valence_api_calculate_energy_(double* x, double* v)
{
    if (!g_api) api_abort("valence_api_calculate_energy before valence_api_initialize");
    try {
        g_api->set_coords_angstrom(x);                                   // valence_api.F90:57-63
        // `orbitals` is rewritten after the guess energy, after every orbital / spin-coupling update ('save') and at
        // convergence ('done' with the last feathering tolerance): valence.F90:192,2823,2859,2882.  The weights written are
        // the normalised ones the reference holds after normal() (valence.F90:144-145,807).
        vb::Engine* eng = g_api.get();
        eng->on_save = [eng](double energy, bool done) {
            const std::vector<std::vector<double>> w = eng->normalized_weights();
            write_orbitals_file(eng->input(), &w, &eng->coupling_weights(), energy, done, std::pow(10.0, -(double)eng->input().ntol_e_max));
        };
        vb::Engine::RunResult r;
        g_api->run(&r, true);                                            // calculate_vsvb_energy, valence_api.F90:96-97
        *v = r.total_energy;
    } catch (const std::exception& ex) {
        api_abort(ex.what());
    }
}

// Host-only utility (no GPU): write `orbitals` (and `nelecwfn`) for the wavefunction of an input file as xm_output
// does, into the current directory.  converged != 0 adds the 'done' line.
int vb_write_wavefunction_files(const char* input_path, double energy, int converged)
{
    return guarded([&] {
        vb::Input in = vb::parse_input_file(input_path);
        write_orbitals_file(in, nullptr, nullptr, energy, converged != 0, std::pow(10.0, -(double)in.ntol_e_max));
    });
}

const double* vb_api_input_coords(void) { return g_api ? g_api->input().coords.data() : nullptr; }

void valence_api_finalize_(int* call_mpi_finalize)
{
    (void)call_mpi_finalize;
    g_api.reset();
}

void init_(int* info)
{
    int one = 1, dummy = 0;
    valence_api_initialize_(info, &one, &dummy);                         // valence_api_nitrogen.F90:4-9
}
void getn_(int* n)
{
    if (!g_api) api_abort("getn before init");
    *n = g_api->input().natom;                                           // valence_api_nitrogen.F90:11-21
}
void calcsurface_(double* x, double* v)
{
    valence_api_calculate_energy_(x, v);
    *v = *v * 219474.631;                                                // valence_api_nitrogen.F90:23-32
}
void finalize_(void)
{
    int one = 1;
    valence_api_finalize_(&one);                                         // valence_api_nitrogen.F90:34-36
}

}  // extern "C"
