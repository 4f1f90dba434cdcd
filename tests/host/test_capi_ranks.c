/* Two (or more) ranks, one process per GPU, driven through the reference-compatible C entry points only
 * (valence_api_initialize_ / valence_api_calculate_energy_ / valence_api_finalize_; /root/reference/src/valence_api.F90):
 * every rank must return the single-rank energy.  Build:  gcc -O2 -o test_capi_ranks test_capi_ranks.c -ldl
 * Run:    ./test_capi_ranks <libvalence_b200.so> <input> <nranks>       (forks nranks children, RANK/WORLD_SIZE/LOCAL_RANK set) */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>

typedef void (*init_t)(int *, int *, int *);
typedef void (*calc_t)(double *, double *);
typedef void (*fin_t)(int *);
typedef const double *(*coords_t)(void);
typedef void (*getn_t)(int *);

static double run_rank(const char *lib, int rank, int nranks)
{
    char buf[32];
    snprintf(buf, sizeof buf, "%d", rank); setenv("RANK", buf, 1); setenv("LOCAL_RANK", buf, 1);
    snprintf(buf, sizeof buf, "%d", nranks); setenv("WORLD_SIZE", buf, 1);
    void *h = dlopen(lib, RTLD_NOW | RTLD_GLOBAL);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); exit(3); }
    init_t init = (init_t)dlsym(h, "valence_api_initialize_");
    calc_t calc = (calc_t)dlsym(h, "valence_api_calculate_energy_");
    fin_t fin = (fin_t)dlsym(h, "valence_api_finalize_");
    coords_t coords = (coords_t)dlsym(h, "vb_api_input_coords");
    getn_t getn = (getn_t)dlsym(h, "getn_");
    int info = -1, one = 1, comm = 0, n = 0;
    init(&info, &one, &comm);
    getn(&n);
    double *x = (double *)malloc(sizeof(double) * 3 * n), v = 0.0;
    memcpy(x, coords(), sizeof(double) * 3 * n);
    calc(x, &v);
    fin(&one);
    free(x);
    return v;
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s lib input nranks\n", argv[0]); return 2; }
    const int nranks = atoi(argv[3]);
    setenv("VALENCE_INPUT", argv[2], 1);
    char key[64];
    snprintf(key, sizeof key, "capitest%d", (int)getpid());
    setenv("VB_NCCL_KEY", key, 1); setenv("VB_SHARD_KEY", key, 1);
    int fds[64][2];
    for (int r = 0; r < nranks; ++r) {
        if (pipe(fds[r]) != 0) return 4;
        pid_t pid = fork();
        if (pid == 0) {
            close(fds[r][0]);
            if (r > 0) { if (!freopen("/dev/null", "w", stdout)) return 5; }
            double v = run_rank(argv[1], r, nranks);
            if (write(fds[r][1], &v, sizeof v) != sizeof v) return 6;
            return 0;
        }
        close(fds[r][1]);
    }
    double e[64];
    int bad = 0;
    for (int r = 0; r < nranks; ++r) { if (read(fds[r][0], &e[r], sizeof(double)) != sizeof(double)) bad = 1; }
    for (int r = 0; r < nranks; ++r) { int st = 0; wait(&st); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad = 1; }
    for (int r = 0; r < nranks; ++r) printf("RANK %d ENERGY %.14f\n", r, e[r]);
    return bad;
}
