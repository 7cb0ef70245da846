// ct_comm.cu -- see ct_comm.h: NCCL bound with dlopen, and the in-process group used by single-GPU tests.
#include "ct_comm.h"

#include <dlfcn.h>
#include <nccl.h>                              // types and prototypes only: every entry point is looked up at run time

#include <condition_variable>
#include <cstring>
#include <mutex>
#include <vector>

namespace ctc {

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// NCCL
// ---------------------------------------------------------------------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string err;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // the libnccl the process already uses (torch loads its bundled one), else the system library
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { api.err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "?"); return; }
        api.handle = h;
#define CT_SYM(name)                                                                 \
    api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "nccl" #name));         \
    if (!api.name) { api.err = "libnccl lacks nccl" #name; return; }
        CT_SYM(GetUniqueId) CT_SYM(CommInitRank) CT_SYM(CommDestroy) CT_SYM(AllGather) CT_SYM(AllReduce) CT_SYM(Broadcast)
        CT_SYM(Send) CT_SYM(Recv) CT_SYM(GroupStart) CT_SYM(GroupEnd) CT_SYM(GetErrorString) CT_SYM(GetVersion)
#undef CT_SYM
    });
    return &api;
}

struct NcclComm : Comm {
    NcclApi* api; ncclComm_t comm; int r, n; bool owned;
    int rank() const override { return r; }
    int size() const override { return n; }
    int check(ncclResult_t e, const char* what) {
        if (e == ncclSuccess) return 0;
        err = std::string(what) + ": " + api->GetErrorString(e);
        return -1;
    }
    int allgather(const void* send, void* recv, size_t bytes, cudaStream_t st) override {
        return check(api->AllGather(send, recv, bytes, ncclUint8, comm, st), "ncclAllGather");
    }
    int allreduce(void* buf, size_t count, Red op, cudaStream_t st) override {
        const ncclDataType_t dt = op == SUM_U32 ? ncclUint32 : op == SUM_U64 ? ncclUint64 : ncclInt64;
        const ncclRedOp_t ro = op == MIN_I64 ? ncclMin : op == MAX_I64 ? ncclMax : ncclSum;
        return check(api->AllReduce(buf, buf, count, dt, ro, comm, st), "ncclAllReduce");
    }
    int sendrecv(const void* send, int dst, void* recv, int src, size_t bytes, cudaStream_t st) override {
        if (dst < 0 && src < 0) return 0;
        if (check(api->GroupStart(), "ncclGroupStart")) return -1;
        ncclResult_t e = ncclSuccess;
        if (dst >= 0) e = api->Send(send, bytes, ncclUint8, dst, comm, st);
        if (e == ncclSuccess && src >= 0) e = api->Recv(recv, bytes, ncclUint8, src, comm, st);
        const ncclResult_t e2 = api->GroupEnd();
        if (check(e, "ncclSend/ncclRecv")) return -1;
        return check(e2, "ncclGroupEnd");
    }
    int bcast(void* buf, size_t bytes, int root, cudaStream_t st) override {
        return check(api->Broadcast(buf, buf, bytes, ncclUint8, root, comm, st), "ncclBroadcast");
    }
    // ---- peer windows through CUDA IPC (one process per GPU) ----
    struct WinRec { cudaIpcMemHandle_t handle; unsigned long long offset, ok; };
    std::vector<void*> opened;                 // bases returned by cudaIpcOpenMemHandle
    void* stage = nullptr;                     // device staging for the exchange of the records
    int cuda(cudaError_t e, const char* what) {
        if (e == cudaSuccess) return 0;
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return -1;
    }
    void close_opened() {
        for (void* p : opened) if (p) cudaIpcCloseMemHandle(p);
        opened.clear();
    }
    // MIN over the ranks of one host value (blocks)
    int agree_min(long long* v, cudaStream_t st) {
        if (cuda(cudaMemcpyAsync(stage, v, 8, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync")) return -1;
        if (allreduce(stage, 1, MIN_I64, st)) return -1;
        if (cuda(cudaMemcpyAsync(v, stage, 8, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync")) return -1;
        return cuda(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    }
    int window_map(void* local, size_t bytes, void** peers, cudaStream_t st) override {
        (void)bytes;
        static_assert(sizeof(WinRec) == 80, "64-byte IPC handle + offset + ok");
        if (!stage && cuda(cudaMalloc(&stage, (size_t)(n + 1) * sizeof(WinRec) + 64), "cudaMalloc")) return -1;
        WinRec mine;
        memset(&mine, 0, sizeof(mine));
        // the handle names the whole allocation the pointer lies in: ship the offset into it as well
        {
            using GetRange = int (*)(unsigned long long*, size_t*, unsigned long long);
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            unsigned long long base = 0; size_t sz = 0;
            if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) == cudaSuccess && fn &&
                reinterpret_cast<GetRange>(fn)(&base, &sz, (unsigned long long)(uintptr_t)local) == 0 && base)
                mine.offset = (unsigned long long)(uintptr_t)local - base;
            (void)cudaGetLastError();
        }
        mine.ok = cudaIpcGetMemHandle(&mine.handle, local) == cudaSuccess ? 1 : 0;
        (void)cudaGetLastError();
        std::vector<WinRec> all((size_t)n);
        WinRec* dev = static_cast<WinRec*>(stage);
        if (cuda(cudaMemcpyAsync(dev + n, &mine, sizeof(mine), cudaMemcpyHostToDevice, st), "cudaMemcpyAsync")) return -1;
        if (allgather(dev + n, dev, sizeof(WinRec), st)) return -1;
        if (cuda(cudaMemcpyAsync(all.data(), dev, (size_t)n * sizeof(WinRec), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync")) return -1;
        if (cuda(cudaStreamSynchronize(st), "cudaStreamSynchronize")) return -1;
        long long ok = 1;
        for (int q = 0; q < n; ++q) ok = ok && all[q].ok;
        close_opened();
        if (ok) {
            for (int q = 0; q < n && ok; ++q) {
                if (q == r) { peers[q] = local; continue; }
                void* base = nullptr;
                if (cudaIpcOpenMemHandle(&base, all[q].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    (void)cudaGetLastError();
                    ok = 0;
                    break;
                }
                opened.push_back(base);
                peers[q] = static_cast<char*>(base) + all[q].offset;
            }
        }
        if (agree_min(&ok, st)) return -1;
        if (!ok) { close_opened(); return 1; }
        return 0;
    }
    int window_unmap(cudaStream_t st) override {
        close_opened();
        if (!stage) return 0;
        long long one = 1;
        return agree_min(&one, st);              // (a barrier: everybody has closed)
    }
    bool window_device_flags() const override { return true; }
    int window_fence(cudaStream_t) override { return 0; }
    ~NcclComm() override {
        close_opened();
        if (stage) cudaFree(stage);
        if (owned && comm) api->CommDestroy(comm);
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// in-process group
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int OP>
__global__ void __launch_bounds__(256) k_local_reduce(const void* const* bufs, int n, size_t count, T* out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        T acc = static_cast<const T*>(bufs[0])[i];
        for (int q = 1; q < n; ++q) {
            const T v = static_cast<const T*>(bufs[q])[i];
            if (OP == 0) acc += v; else if (OP == 1) acc = v < acc ? v : acc; else acc = v > acc ? v : acc;
        }
        out[i] = acc;
    }
}

}  // namespace

struct LocalGroup {
    int n = 0;
    std::mutex m; std::condition_variable cv; int arrived = 0; long gen = 0;
    std::vector<const void*> ptr;
    long votes = 0;
    std::vector<cudaEvent_t> ready, done;
    void barrier() {
        std::unique_lock<std::mutex> lk(m);
        const long g = gen;
        if (++arrived == n) { arrived = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};

namespace {

struct LocalComm : Comm {
    LocalGroup* g; int r;
    void* tmp = nullptr; size_t tmp_cap = 0;
    const void** dev_ptrs = nullptr;
    int rank() const override { return r; }
    int size() const override { return g->n; }
    int cuda(cudaError_t e, const char* what) {
        if (e == cudaSuccess) return 0;
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return -1;
    }
    // phase 1: my buffer is ready on `st`; after the barrier everybody's pointer is visible
    int publish(const void* p, cudaStream_t st) {
        if (!g->ready[r]) {
            if (cuda(cudaEventCreateWithFlags(&g->ready[r], cudaEventDisableTiming), "cudaEventCreate")) return -1;
            if (cuda(cudaEventCreateWithFlags(&g->done[r], cudaEventDisableTiming), "cudaEventCreate")) return -1;
        }
        g->ptr[r] = p;
        if (cuda(cudaEventRecord(g->ready[r], st), "cudaEventRecord")) return -1;
        g->barrier();
        return 0;
    }
    // phase 3: I am done reading; `waiters` must not touch their buffer before the readers have finished
    int finish(cudaStream_t st, bool wait_all, int wait_one) {
        if (cuda(cudaEventRecord(g->done[r], st), "cudaEventRecord")) return -1;
        g->barrier();
        for (int q = 0; q < g->n; ++q)
            if (q != r && (wait_all || q == wait_one))
                if (cuda(cudaStreamWaitEvent(st, g->done[q], 0), "cudaStreamWaitEvent")) return -1;
        g->barrier();                             // nobody re-records an event somebody else is about to wait for
        return 0;
    }
    int allgather(const void* send, void* recv, size_t bytes, cudaStream_t st) override {
        if (publish(send, st)) return -1;
        for (int q = 0; q < g->n; ++q) {
            if (q != r && cuda(cudaStreamWaitEvent(st, g->ready[q], 0), "cudaStreamWaitEvent")) return -1;
            if (cuda(cudaMemcpyAsync(static_cast<char*>(recv) + (size_t)q * bytes, g->ptr[q], bytes, cudaMemcpyDefault, st),
                     "cudaMemcpyAsync")) return -1;
        }
        return finish(st, true, -1);
    }
    int allreduce(void* buf, size_t count, Red op, cudaStream_t st) override {
        const size_t elt = op == SUM_U32 ? 4 : 8;
        if (count * elt > tmp_cap) {
            if (tmp) cudaFree(tmp);
            tmp = nullptr; tmp_cap = 0;
            if (cuda(cudaMalloc(&tmp, count * elt + 256), "cudaMalloc")) return -1;
            tmp_cap = count * elt + 256;
        }
        if (!dev_ptrs && cuda(cudaMalloc(&dev_ptrs, 64 * sizeof(void*)), "cudaMalloc")) return -1;
        if (g->n > 64) { err = "in-process group limited to 64 ranks"; return -1; }
        if (publish(buf, st)) return -1;
        for (int q = 0; q < g->n; ++q)
            if (q != r && cuda(cudaStreamWaitEvent(st, g->ready[q], 0), "cudaStreamWaitEvent")) return -1;
        if (cuda(cudaMemcpyAsync(dev_ptrs, g->ptr.data(), (size_t)g->n * sizeof(void*), cudaMemcpyHostToDevice, st),
                 "cudaMemcpyAsync")) return -1;
        if (cuda(cudaStreamSynchronize(st), "cudaStreamSynchronize")) return -1;     // (g->ptr is rewritten by the next call)
        const unsigned blocks = (unsigned)std::min<size_t>((count + 255) / 256, 1184);
        if (count) {
            if (op == SUM_U32) k_local_reduce<uint32_t, 0><<<blocks, 256, 0, st>>>(dev_ptrs, g->n, count, (uint32_t*)tmp);
            else if (op == SUM_U64) k_local_reduce<unsigned long long, 0><<<blocks, 256, 0, st>>>(dev_ptrs, g->n, count, (unsigned long long*)tmp);
            else if (op == MIN_I64) k_local_reduce<long long, 1><<<blocks, 256, 0, st>>>(dev_ptrs, g->n, count, (long long*)tmp);
            else k_local_reduce<long long, 2><<<blocks, 256, 0, st>>>(dev_ptrs, g->n, count, (long long*)tmp);
            if (cuda(cudaGetLastError(), "k_local_reduce")) return -1;
        }
        if (finish(st, true, -1)) return -1;
        return cuda(cudaMemcpyAsync(buf, tmp, count * elt, cudaMemcpyDeviceToDevice, st), "cudaMemcpyAsync");
    }
    int sendrecv(const void* send, int dst, void* recv, int src, size_t bytes, cudaStream_t st) override {
        if (publish(send, st)) return -1;
        if (src >= 0) {
            if (cuda(cudaStreamWaitEvent(st, g->ready[src], 0), "cudaStreamWaitEvent")) return -1;
            if (cuda(cudaMemcpyAsync(recv, g->ptr[src], bytes, cudaMemcpyDefault, st), "cudaMemcpyAsync")) return -1;
        }
        return finish(st, false, dst);
    }
    int bcast(void* buf, size_t bytes, int root, cudaStream_t st) override {
        if (publish(buf, st)) return -1;
        if (r != root) {
            if (cuda(cudaStreamWaitEvent(st, g->ready[root], 0), "cudaStreamWaitEvent")) return -1;
            if (cuda(cudaMemcpyAsync(buf, g->ptr[root], bytes, cudaMemcpyDefault, st), "cudaMemcpyAsync")) return -1;
        }
        return finish(st, r == root, -1);
    }
    // peer windows: the ranks share one address space; pointers of the other ranks are valid as they are (the ranks of a
    // test group sit on one device; for several devices of one process peer access is switched on)
    int window_map(void* local, size_t, void** peers, cudaStream_t) override {
        g->ptr[r] = local;
        g->barrier();
        int mydev = 0;
        long bad = 0;
        if (cuda(cudaGetDevice(&mydev), "cudaGetDevice")) bad = 1;
        for (int q = 0; q < g->n && !bad; ++q) {
            peers[q] = const_cast<void*>(g->ptr[q]);
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, peers[q]) != cudaSuccess) { bad = 1; break; }
            if (at.device != mydev) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) bad = 1;
            }
        }
        (void)cudaGetLastError();
        {
            std::unique_lock<std::mutex> lk(g->m);
            g->votes += bad;
        }
        g->barrier();
        const long total = g->votes;
        g->barrier();
        if (r == 0) g->votes = 0;
        g->barrier();
        return total ? 1 : 0;
    }
    int window_unmap(cudaStream_t) override { g->barrier(); return 0; }
    bool window_device_flags() const override { return false; }
    int window_fence(cudaStream_t st) override {
        if (publish(nullptr, st)) return -1;
        for (int q = 0; q < g->n; ++q)
            if (q != r && cuda(cudaStreamWaitEvent(st, g->ready[q], 0), "cudaStreamWaitEvent")) return -1;
        g->barrier();                             // nobody re-records `ready` before everybody has enqueued its waits
        return 0;
    }
    ~LocalComm() override { if (tmp) cudaFree(tmp); if (dev_ptrs) cudaFree(dev_ptrs); }
};

}  // namespace

int nccl_unique_id(unsigned char id[128], std::string& err) {
    NcclApi* api = nccl_api();
    if (!api->handle) { err = api->err; return -1; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    const ncclResult_t e = api->GetUniqueId(&u);
    if (e != ncclSuccess) { err = std::string("ncclGetUniqueId: ") + api->GetErrorString(e); return -1; }
    memcpy(id, &u, 128);
    return 0;
}

Comm* nccl_create(const unsigned char id[128], int rank, int nranks, std::string& err) {
    NcclApi* api = nccl_api();
    if (!api->handle) { err = api->err; return nullptr; }
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclComm_t comm = nullptr;
    const ncclResult_t e = api->CommInitRank(&comm, nranks, u, rank);
    if (e != ncclSuccess) { err = std::string("ncclCommInitRank: ") + api->GetErrorString(e); return nullptr; }
    NcclComm* c = new NcclComm();
    c->api = api; c->comm = comm; c->r = rank; c->n = nranks; c->owned = true;
    return c;
}

Comm* nccl_wrap(void* nccl_comm, int rank, int nranks, std::string& err) {
    NcclApi* api = nccl_api();
    if (!api->handle) { err = api->err; return nullptr; }
    NcclComm* c = new NcclComm();
    c->api = api; c->comm = static_cast<ncclComm_t>(nccl_comm); c->r = rank; c->n = nranks; c->owned = false;
    return c;
}

LocalGroup* local_group_create(int nranks) {
    LocalGroup* g = new LocalGroup();
    g->n = nranks;
    g->ptr.assign(nranks, nullptr); g->ready.assign(nranks, nullptr); g->done.assign(nranks, nullptr);
    return g;
}

void local_group_destroy(LocalGroup* g) {
    if (!g) return;
    for (cudaEvent_t e : g->ready) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : g->done) if (e) cudaEventDestroy(e);
    delete g;
}

Comm* local_comm(LocalGroup* g, int rank) {
    LocalComm* c = new LocalComm();
    c->g = g; c->r = rank;
    return c;
}

}  // namespace ctc
