// NCCL inside the library: the replacement of the reference's MPI layer (xm_propagate / xm_equalize /
// xm_equalize_scalar / xm_end, /root/reference/src/xm_module.F90:711-910) for one process per GPU on one node.
//
// libnccl is resolved at run time (dlopen of libnccl.so.2: the copy a host such as PyTorch has already loaded is
// reused, otherwise the system one), so the library neither links against NCCL nor needs it for single-GPU use.
// Bootstrap without MPI: rank and size come from the launcher's environment (torchrun, Open MPI, PMI, Slurm), rank 0
// publishes the ncclUniqueId in a file under /dev/shm keyed by the job, the others read it.
#pragma once
#include <cuda_runtime.h>

#include <string>

namespace vb {

struct LaunchEnv { int rank = 0, nranks = 1, local_rank = 0; std::string key; };
// RANK / WORLD_SIZE / LOCAL_RANK (torchrun), OMPI_COMM_WORLD_*, PMI_*, SLURM_*; nranks = 1 when none is set
LaunchEnv launch_env();

class Comm {
public:
    // creates the communicator collectively (every rank of the job must call it); throws on failure
    Comm(int rank, int nranks, const std::string& key);
    // adopts an existing ncclComm_t (not destroyed by this object)
    Comm(int rank, int nranks, void* nccl_comm);
    ~Comm();
    int rank() const { return rank_; }
    int nranks() const { return nranks_; }
    void allreduce_sum(double* dev, size_t n, cudaStream_t st);          // in place, FP64 sum
    // in place: rank r's bytes sit at base + r * bytes_per_rank on entry, everybody's on return
    void allgather_bytes(void* base, size_t bytes_per_rank, cudaStream_t st);
    void barrier(cudaStream_t st);                                       // all-reduce of one double + stream sync
private:
    int rank_, nranks_;
    void* comm_ = nullptr;
    bool owned_ = false;
    double* scratch_ = nullptr;
    std::string id_file_;
};

}  // namespace vb
