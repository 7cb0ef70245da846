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
#include <vector>

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
    // ---- peer windows: every rank exposes ONE device allocation to all the others (NVLink / NVSwitch peer memory), so that
    // a kernel can store straight into its peers' buffers.  Collective; blocks the host.  peers[q] = address, valid on THIS
    // rank's device, of rank q's window (peers[rank] = local).  0 = mapped everywhere; 1 = peer memory is not available on
    // some rank (agreed by all: nothing stays mapped, the caller uses the collectives); -1 = error.
    virtual int window_map(void* local, size_t bytes, void** peers, cudaStream_t st) = 0;
    // collective; when it returns nobody has the windows mapped any more and the owners may free them
    virtual int window_unmap(cudaStream_t st) = 0;
    // true: ranks are separate processes / devices and signal each other through flag words inside the windows;
    // false (in-process group): window_fence() orders the ranks' streams instead (everything enqueued on every rank's `st`
    // before the fence is visible to what every rank enqueues after it)
    virtual bool window_device_flags() const = 0;
    virtual int window_fence(cudaStream_t st) = 0;
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
    // ---- state of the table exchange of the sharded run.  It lives with the communicator, not with a context, because it
    // only changes in collective steps: every rank of the communicator holds the same values at the same call.
    long capC = 0, capP = 0, capS = 0;         // negotiated per-rank slot capacities (0 = not yet)
    int window_mode = 0;                       // 0 = not tried, 1 = peer windows mapped, 2 = unavailable (collectives)
    void* window = nullptr; size_t window_bytes = 0; int window_device = -1;
    std::vector<void*> peers;                  // [nranks] rank q's window as seen from this device
    unsigned long long epoch = 0;              // exchanges since the windows were mapped (parity = double buffer)
};
