// dist.cu -- multi-GPU plumbing for the feature-sharded histogram.
//
// The reference has no distributed code at all (single process, device 0).  Here every rank holds all rows,
// builds the histograms of its own feature tiles only, and one ncclAllReduce(sum, int64) per depth level
// makes the per-bin sums of every tile visible on every rank; slices are disjoint and the sums are integers,
// so the result is bit-identical on all ranks whatever ring/tree order NCCL picks.  Each rank then runs the
// same scan / arg-max / partition locally (no second exchange).
//
// NCCL is resolved with dlopen so that a single-GPU process never needs libnccl.
#include "engine.cuh"
#include <dlfcn.h>
#include <string.h>
#include <mutex>

namespace gb {

typedef struct { char internal[128]; } nccl_uid_t;
typedef void *nccl_comm_t;
typedef int (*fn_GetUniqueId)(nccl_uid_t *);
typedef int (*fn_CommInitRank)(nccl_comm_t *, int, nccl_uid_t, int);
typedef int (*fn_AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_CommDestroy)(nccl_comm_t);
typedef const char *(*fn_GetErrorString)(int);

static struct {
    void *lib = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_AllReduce AllReduce = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_GetErrorString GetErrorString = nullptr;
} nccl;

static void load_nccl() {
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (nccl.lib) break;
        }
        if (!nccl.lib) return;
        nccl.GetUniqueId = (fn_GetUniqueId)dlsym(nccl.lib, "ncclGetUniqueId");
        nccl.CommInitRank = (fn_CommInitRank)dlsym(nccl.lib, "ncclCommInitRank");
        nccl.AllReduce = (fn_AllReduce)dlsym(nccl.lib, "ncclAllReduce");
        nccl.CommDestroy = (fn_CommDestroy)dlsym(nccl.lib, "ncclCommDestroy");
        nccl.GetErrorString = (fn_GetErrorString)dlsym(nccl.lib, "ncclGetErrorString");
    });
    GB_CHECK(nccl.lib && nccl.GetUniqueId && nccl.CommInitRank && nccl.AllReduce, "libnccl.so.2 could not be loaded");
}

#define GB_NCCL(expr)                                                                              \
    do {                                                                                           \
        int _r = (expr);                                                                           \
        if (_r != 0) throw gb::Error(std::string("NCCL error: ") + (nccl.GetErrorString ? nccl.GetErrorString(_r) : "?")); \
    } while (0)

void dist_unique_id(uint8_t id[128]) {
    load_nccl();
    nccl_uid_t u;
    GB_NCCL(nccl.GetUniqueId(&u));
    memcpy(id, &u, 128);
}

void dist_init(Model &m, const uint8_t id[128], int rank, int world) {
    load_nccl();
    nccl_uid_t u;
    memcpy(&u, id, 128);
    nccl_comm_t comm = nullptr;
    GB_NCCL(nccl.CommInitRank(&comm, world, u, rank));
    m.nccl_comm = comm; m.rank = rank; m.world = world;
}

void dist_shutdown(Model &m) {
    if (m.nccl_comm && nccl.CommDestroy) nccl.CommDestroy((nccl_comm_t)m.nccl_comm);
    m.nccl_comm = nullptr; m.rank = 0; m.world = 1;
}

void dist_allreduce_hist(Model &m, long long *buf, size_t count, cudaStream_t s) {
    GB_CHECK(m.nccl_comm != nullptr, "distributed histogram requested without an initialised communicator");
    const int ncclInt64 = 4, ncclSum = 0;
    GB_NCCL(nccl.AllReduce(buf, buf, count, ncclInt64, ncclSum, (nccl_comm_t)m.nccl_comm, s));
}

}  // namespace gb
