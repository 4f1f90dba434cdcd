#include "vb_nccl.h"

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <stdexcept>
#include <thread>

namespace vb {

namespace {

// the part of the NCCL ABI used here (nccl.h: stable since 2.x)
struct UniqueId { char internal[128]; };
using CommT = void*;
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;
struct Api {
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(CommT*, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(CommT) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, CommT, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, CommT, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};

Api& api()
{
    static Api a;
    if (a.ok) return a;
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) throw std::runtime_error(std::string("valence_b200: cannot load libnccl (") + dlerror() + ")");
    auto sym = [&](const char* n) {
        void* p = dlsym(h, n);
        if (!p) throw std::runtime_error(std::string("valence_b200: libnccl lacks ") + n);
        return p;
    };
    a.GetUniqueId = reinterpret_cast<int (*)(UniqueId*)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<int (*)(CommT*, int, UniqueId, int)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<int (*)(CommT)>(sym("ncclCommDestroy"));
    a.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, CommT, cudaStream_t)>(sym("ncclAllReduce"));
    a.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, CommT, cudaStream_t)>(sym("ncclAllGather"));
    a.GetErrorString = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
    a.ok = true;
    return a;
}

void check(int rc, const char* what)
{
    if (rc != 0) throw std::runtime_error(std::string("valence_b200: NCCL ") + what + ": " + api().GetErrorString(rc));
}

int env_int(const char* name, int dflt)
{
    const char* v = std::getenv(name);
    return v && *v ? std::atoi(v) : dflt;
}

}  // namespace

LaunchEnv launch_env()
{
    LaunchEnv e;
    struct Names { const char *rank, *size, *local; };
    for (const Names& n : {Names{"RANK", "WORLD_SIZE", "LOCAL_RANK"}, Names{"OMPI_COMM_WORLD_RANK", "OMPI_COMM_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_RANK"},
                           Names{"PMI_RANK", "PMI_SIZE", "MPI_LOCALRANKID"}, Names{"SLURM_PROCID", "SLURM_NTASKS", "SLURM_LOCALID"}}) {
        if (std::getenv(n.rank) && std::getenv(n.size)) {
            e.rank = env_int(n.rank, 0); e.nranks = std::max(1, env_int(n.size, 1)); e.local_rank = env_int(n.local, e.rank);
            break;
        }
    }
    if (const char* k = std::getenv("VB_NCCL_KEY")) e.key = k;
    else if (const char* p = std::getenv("MASTER_PORT")) e.key = std::string("port") + p;
    else e.key = "ppid" + std::to_string((long)getppid());
    return e;
}

Comm::Comm(int rank, int nranks, const std::string& key) : rank_(rank), nranks_(nranks), owned_(true)
{
    if (nranks < 1 || rank < 0 || rank >= nranks) throw std::runtime_error("valence_b200: bad rank / size");
    Api& a = api();
    // One rendezvous file per communicator of the process: the ranks of a job create their communicators in the same order, so a
    // rank that runs ahead can never pick up the record of the previous communicator (which stays until rank 0 has closed it).
    static std::atomic<unsigned> n_created{0};
    id_file_ = "/dev/shm/valence_b200_nccl_" + key + "." + std::to_string(n_created.fetch_add(1));
    UniqueId id;
    std::memset(&id, 0, sizeof id);
    struct Rec { long long stamp; UniqueId id; } rec;
    if (rank == 0) {
        check(a.GetUniqueId(&id), "ncclGetUniqueId");
        rec.stamp = (long long)std::time(nullptr);
        rec.id = id;
        const std::string tmp = id_file_ + ".tmp";
        FILE* f = std::fopen(tmp.c_str(), "wb");
        if (!f || std::fwrite(&rec, sizeof rec, 1, f) != 1) throw std::runtime_error("valence_b200: cannot publish the NCCL id in " + tmp);
        std::fclose(f);
        if (std::rename(tmp.c_str(), id_file_.c_str()) != 0) throw std::runtime_error("valence_b200: cannot publish the NCCL id in " + id_file_);
    } else {
        // wait for a fresh record (a stale file of an earlier job with the same key is older than two minutes or gets replaced)
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            FILE* f = std::fopen(id_file_.c_str(), "rb");
            bool got = false;
            if (f) {
                got = std::fread(&rec, sizeof rec, 1, f) == 1 && std::llabs((long long)std::time(nullptr) - rec.stamp) < 120;
                std::fclose(f);
            }
            if (got) break;
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(300)) throw std::runtime_error("valence_b200: no NCCL id from rank 0 in " + id_file_);
            std::this_thread::sleep_for(std::chrono::milliseconds(20));
        }
        id = rec.id;
    }
    check(a.CommInitRank(&comm_, nranks, id, rank), "ncclCommInitRank");
    cudaMalloc((void**)&scratch_, sizeof(double));
}

Comm::Comm(int rank, int nranks, void* nccl_comm) : rank_(rank), nranks_(nranks), comm_(nccl_comm), owned_(false)
{
    api();
    cudaMalloc((void**)&scratch_, sizeof(double));
}

Comm::~Comm()
{
    if (scratch_) cudaFree(scratch_);
    if (owned_ && comm_) api().CommDestroy(comm_);
    if (owned_ && rank_ == 0 && !id_file_.empty()) std::remove(id_file_.c_str());
}

void Comm::allreduce_sum(double* dev, size_t n, cudaStream_t st)
{
    if (nranks_ == 1 || n == 0) return;
    check(api().AllReduce(dev, dev, n, NCCL_FLOAT64, NCCL_SUM, comm_, st), "ncclAllReduce");
}

void Comm::allgather_bytes(void* base, size_t bytes_per_rank, cudaStream_t st)
{
    if (nranks_ == 1 || bytes_per_rank == 0) return;
    check(api().AllGather(static_cast<const char*>(base) + (size_t)rank_ * bytes_per_rank, base, bytes_per_rank, 0 /* ncclInt8 */, comm_, st), "ncclAllGather");
}

void Comm::barrier(cudaStream_t st)
{
    if (nranks_ == 1) return;
    cudaMemsetAsync(scratch_, 0, sizeof(double), st);
    allreduce_sum(scratch_, 1, st);
    cudaStreamSynchronize(st);
}

}  // namespace vb
