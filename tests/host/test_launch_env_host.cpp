// Host check of the launcher detection of the in-library NCCL layer (vb_nccl.cpp, launch_env): rank / size / local rank
// from torchrun, Open MPI, PMI (MPICH, Intel MPI) and Slurm variables, single-process default, rendezvous key from
// VB_NCCL_KEY | MASTER_PORT | parent pid; a communicator with an impossible rank is refused before anything is loaded.
// (what the reference gets from MPI_Comm_rank / MPI_Comm_size in xm_propagate, /root/reference/src/xm_module.F90:727-750)
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <string>

#include "vb_nccl.h"

static int fails = 0;
#define EXPECT(c) do { if (!(c)) { std::printf("FAILED line %d: %s\n", __LINE__, #c); ++fails; } } while (0)

static void clear_all()
{
    for (const char* n : {"RANK", "WORLD_SIZE", "LOCAL_RANK", "OMPI_COMM_WORLD_RANK", "OMPI_COMM_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_RANK",
                          "PMI_RANK", "PMI_SIZE", "MPI_LOCALRANKID", "SLURM_PROCID", "SLURM_NTASKS", "SLURM_LOCALID", "VB_NCCL_KEY", "MASTER_PORT"})
        unsetenv(n);
}

int main()
{
    clear_all();
    vb::LaunchEnv e = vb::launch_env();
    EXPECT(e.rank == 0 && e.nranks == 1 && e.local_rank == 0);
    EXPECT(e.key == "ppid" + std::to_string((long)getppid()));

    setenv("RANK", "3", 1); setenv("WORLD_SIZE", "8", 1); setenv("LOCAL_RANK", "3", 1); setenv("MASTER_PORT", "29533", 1);
    e = vb::launch_env();
    EXPECT(e.rank == 3 && e.nranks == 8 && e.local_rank == 3 && e.key == "port29533");
    setenv("VB_NCCL_KEY", "job42", 1);
    EXPECT(vb::launch_env().key == "job42");

    clear_all();
    setenv("OMPI_COMM_WORLD_RANK", "5", 1); setenv("OMPI_COMM_WORLD_SIZE", "6", 1); setenv("OMPI_COMM_WORLD_LOCAL_RANK", "1", 1);
    e = vb::launch_env();
    EXPECT(e.rank == 5 && e.nranks == 6 && e.local_rank == 1);

    clear_all();
    setenv("PMI_RANK", "1", 1); setenv("PMI_SIZE", "2", 1);              // no local-rank variable: one node, local = global
    e = vb::launch_env();
    EXPECT(e.rank == 1 && e.nranks == 2 && e.local_rank == 1);

    clear_all();
    setenv("SLURM_PROCID", "2", 1); setenv("SLURM_NTASKS", "4", 1); setenv("SLURM_LOCALID", "2", 1);
    e = vb::launch_env();
    EXPECT(e.rank == 2 && e.nranks == 4 && e.local_rank == 2);

    clear_all();
    setenv("RANK", "0", 1);                                              // a rank without a size is not a launch
    e = vb::launch_env();
    EXPECT(e.nranks == 1 && e.rank == 0);
    setenv("WORLD_SIZE", "0", 1);
    EXPECT(vb::launch_env().nranks == 1);

    for (int bad : {-1, 2}) {
        bool thrown = false;
        try { vb::Comm c(bad, 2, "never"); } catch (const std::exception&) { thrown = true; }
        EXPECT(thrown);
    }
    if (fails) return 1;
    std::printf("PASS\n");
    return 0;
}
