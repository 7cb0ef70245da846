// ct_comm.h -- the collectives the time-sharded entry points need, behind one small interface with two implementations:
//   * NCCL (ncclAllGather / ncclAllReduce / ncclSend+ncclRecv / ncclBroadcast over NVLink / NVSwitch), bound at run time
//     with dlopen so that the library has no link-time dependency and uses the libnccl the host process already loaded
//     (torch ships its own); the communicator is either created here from a ncclUniqueId the caller distributes, or a
//     ncclComm_t the caller owns;
//   * an in-process group (N contexts on one or more GPUs of ONE process, one host thread per rank): the same stream-ordered
//     semantics with events and peer copies.  It exists so that every line of the sharded drivers can run on a single-GPU
//     test box; it is not a product transport.
// All calls are stream-ordered like NCCL's: they enqueue on `st` and return.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>

namespace ctc {

enum Red { SUM_U32 = 0, SUM_U64 = 1, MIN_I64 = 2, MAX_I64 = 3 };

struct Comm {
    virtual ~Comm() {}
    virtual int rank() const = 0;
    virtual int size() const = 0;
    // recv[r * bytes .. (r + 1) * bytes) = rank r's send
    virtual int allgather(const void* send, void* recv, size_t bytes, cudaStream_t st) = 0;
    virtual int allreduce(void* buf, size_t count, Red op, cudaStream_t st) = 0;        // in place
    // one exchange with the two neighbours: send to `dst` and receive from `src` (either may be -1 = none)
    virtual int sendrecv(const void* send, int dst, void* recv, int src, size_t bytes, cudaStream_t st) = 0;
    virtual int bcast(void* buf, size_t bytes, int root, cudaStream_t st) = 0;
    std::string err;                          // message of the last failure (calls return 0 or -1)
};

// rank 0 makes the id (128 bytes), the caller distributes it (MPI_Bcast, a file, torch.distributed ...), every rank joins
int nccl_unique_id(unsigned char id[128], std::string& err);
Comm* nccl_create(const unsigned char id[128], int rank, int nranks, std::string& err);
Comm* nccl_wrap(void* nccl_comm, int rank, int nranks, std::string& err);               // the caller keeps ownership

struct LocalGroup;
LocalGroup* local_group_create(int nranks);
void local_group_destroy(LocalGroup* g);
Comm* local_comm(LocalGroup* g, int rank);

}  // namespace ctc

struct ct_comm {
    ctc::Comm* impl = nullptr;
    ctc::LocalGroup* group = nullptr;          // set on rank 0's handle of an in-process group (owner)
};
