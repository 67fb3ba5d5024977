// dd_comm.hpp — transports of the slab decomposition (dd_kernels.cuh): a ring exchange with the two neighbouring
// slabs and small all-reduces (flags, energies, state gathers).
//
//   NcclComm  : one process per GPU (torchrun); ncclSend/ncclRecv pairs grouped per exchange, over NVLink/NVSwitch.
//               libnccl is opened at run time (dlopen), so the engine has no link-time dependency on it.
//   LocalComm : all ranks are threads of this process (one stream each, on one or several devices): device-to-device
//               copies ordered by events.  This is the single-process multi-GPU mode, and with every rank on one
//               device it is how the decomposition is checked bit for bit against a single-domain run on one GPU.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace ocm {

enum DDType : int { DD_I32 = 0, DD_I64 = 1, DD_F32 = 2, DD_F64 = 3 };

struct DDComm {
    int world = 1;
    virtual ~DDComm() {}
    // on stream s of rank `rank`: send_lo -> rank-1, send_hi -> rank+1; recv_lo <- (rank-1)'s send_hi,
    // recv_hi <- (rank+1)'s send_lo (periodic ring).  Buffers must stay untouched until the stream passes this call.
    virtual void ring_exchange(int rank, const void* send_lo, const void* send_hi, void* recv_lo, void* recv_hi, size_t bytes,
                               cudaStream_t s) = 0;
    // in-place all-reduce: max for DD_I32 flags (is_max) or sum
    virtual void allreduce(int rank, void* buf, size_t count, DDType type, bool is_max, cudaStream_t s) = 0;
    virtual const char* kind() const = 0;
    // peer memory (NVLink/NVSwitch P2P): map the device allocation `base` of the two ring neighbours into this rank's
    // address space, so that kernels here can store straight into it.  Collective.  false = not available with this
    // transport (the caller keeps using ring_exchange).  peer_unmap releases what peer_map opened.
    virtual bool peer_map(int rank, void* base, void** lo_base, void** hi_base, cudaStream_t s) { (void)rank; (void)base; (void)lo_base; (void)hi_base; (void)s; return false; }
    virtual void peer_unmap(int rank) { (void)rank; }
};

#define OCM_CUDA_OK(expr)                                                                                          \
    do {                                                                                                           \
        cudaError_t _e = (expr);                                                                                   \
        if (_e != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " (" #expr ")"); \
    } while (0)

// ---- NCCL ------------------------------------------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;

    static NcclApi& get() {
        static NcclApi api;
        static std::once_flag once;
        std::call_once(once, [] {
            // a libnccl already mapped into the process (e.g. the one torch.distributed brought) is reused by soname
            const char* names[] = {"libnccl.so.2", "libnccl.so"};
            for (const char* nm : names) { api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
            if (!api.handle) return;
            auto sym = [&](const char* n) { return dlsym(api.handle, n); };
            api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
            api.Send = (decltype(api.Send))sym("ncclSend");
            api.Recv = (decltype(api.Recv))sym("ncclRecv");
            api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
            api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
            api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        });
        if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.Send || !api.Recv || !api.AllReduce || !api.GroupStart || !api.GroupEnd)
            throw std::runtime_error("NCCL is not available (libnccl.so.2 could not be opened)");
        return api;
    }
    void ok(ncclResult_t r, const char* what) const {
        if (r != ncclSuccess) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + (GetErrorString ? GetErrorString(r) : "?"));
    }
};

struct NcclComm : DDComm {
    ncclComm_t comm = nullptr;
    int rank = 0, device = 0;
    NcclComm(const ncclUniqueId& id, int world_, int rank_, int device_) : rank(rank_), device(device_) {
        world = world_;
        NcclApi& A = NcclApi::get();
        OCM_CUDA_OK(cudaSetDevice(device));
        A.ok(A.CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
    }
    ~NcclComm() override { peer_unmap(rank); if (stage) cudaFree(stage); if (comm) NcclApi::get().CommDestroy(comm); }
    const char* kind() const override { return "nccl"; }
    void ring_exchange(int r, const void* send_lo, const void* send_hi, void* recv_lo, void* recv_hi, size_t bytes, cudaStream_t s) override {
        NcclApi& A = NcclApi::get();
        const int prev = (r + world - 1) % world, next = (r + 1) % world;
        // same-peer operations match in issue order (world == 2: prev == next): my first send pairs with the peer's
        // first receive, i.e. my send_lo lands in its recv_hi
        A.ok(A.GroupStart(), "ncclGroupStart");
        A.ok(A.Send(send_lo, bytes, ncclChar, prev, comm, s), "ncclSend");
        A.ok(A.Send(send_hi, bytes, ncclChar, next, comm, s), "ncclSend");
        A.ok(A.Recv(recv_hi, bytes, ncclChar, next, comm, s), "ncclRecv");
        A.ok(A.Recv(recv_lo, bytes, ncclChar, prev, comm, s), "ncclRecv");
        A.ok(A.GroupEnd(), "ncclGroupEnd");
    }
    void allreduce(int, void* buf, size_t count, DDType type, bool is_max, cudaStream_t s) override {
        NcclApi& A = NcclApi::get();
        const ncclDataType_t t = type == DD_I32 ? ncclInt32 : type == DD_I64 ? ncclInt64 : type == DD_F32 ? ncclFloat32 : ncclFloat64;
        A.ok(A.AllReduce(buf, buf, count, t, is_max ? ncclMax : ncclSum, comm, s), "ncclAllReduce");
    }

    // CUDA IPC: every rank exports its allocation, the 64-byte handles travel round the ring over NCCL, each rank opens
    // its two neighbours' handles (one process per GPU on one node: NVLink peer access)
    void* opened[2] = {nullptr, nullptr};
    char* stage = nullptr;  // device: [my handle | from prev | from next], 64 bytes each
    bool peer_map(int r, void* base, void** lo_base, void** hi_base, cudaStream_t s) override {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
        peer_unmap(r);
        cudaIpcMemHandle_t mine, from_prev, from_next;
        int ok = cudaIpcGetMemHandle(&mine, base) == cudaSuccess ? 1 : 0;
        if (!ok) { cudaGetLastError(); std::memset(&mine, 0, sizeof(mine)); }
        if (!stage) OCM_CUDA_OK(cudaMalloc(&stage, 4 * 64));
        OCM_CUDA_OK(cudaMemcpyAsync(stage, &mine, 64, cudaMemcpyHostToDevice, s));
        ring_exchange(r, stage, stage, stage + 64, stage + 128, 64, s);
        // everybody must have exported successfully, or nobody maps
        int* flag = reinterpret_cast<int*>(stage + 192);
        OCM_CUDA_OK(cudaMemcpyAsync(flag, &ok, sizeof(int), cudaMemcpyHostToDevice, s));
        NcclApi::get().ok(NcclApi::get().AllReduce(flag, flag, 1, ncclInt32, ncclMin, comm, s), "ncclAllReduce");
        OCM_CUDA_OK(cudaMemcpyAsync(&from_prev, stage + 64, 64, cudaMemcpyDeviceToHost, s));
        OCM_CUDA_OK(cudaMemcpyAsync(&from_next, stage + 128, 64, cudaMemcpyDeviceToHost, s));
        OCM_CUDA_OK(cudaMemcpyAsync(&ok, flag, sizeof(int), cudaMemcpyDeviceToHost, s));
        OCM_CUDA_OK(cudaStreamSynchronize(s));
        if (!ok) return false;
        const int prev = (r + world - 1) % world, next = (r + 1) % world;
        int opened_ok = 1;
        if (cudaIpcOpenMemHandle(&opened[0], from_prev, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); opened[0] = nullptr; opened_ok = 0; }
        if (prev == next) opened[1] = nullptr;  // the same allocation: opened once
        else if (cudaIpcOpenMemHandle(&opened[1], from_next, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); opened[1] = nullptr; opened_ok = 0; }
        OCM_CUDA_OK(cudaMemcpyAsync(flag, &opened_ok, sizeof(int), cudaMemcpyHostToDevice, s));
        NcclApi::get().ok(NcclApi::get().AllReduce(flag, flag, 1, ncclInt32, ncclMin, comm, s), "ncclAllReduce");
        OCM_CUDA_OK(cudaMemcpyAsync(&opened_ok, flag, sizeof(int), cudaMemcpyDeviceToHost, s));
        OCM_CUDA_OK(cudaStreamSynchronize(s));
        if (!opened_ok) { peer_unmap(r); return false; }
        *lo_base = opened[0];
        *hi_base = prev == next ? opened[0] : opened[1];
        return true;
    }
    void peer_unmap(int) override {
        for (void*& q : opened) { if (q) cudaIpcCloseMemHandle(q); q = nullptr; }
    }
};

// ---- in-process ------------------------------------------------------------------------------------
constexpr int kMaxLocalRanks = 16;
struct PtrList { const void* p[kMaxLocalRanks]; };

template <class T, bool MAX>
__global__ void dd_local_reduce_kernel(T* __restrict__ out, PtrList in, int nin, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    T acc = static_cast<const T*>(in.p[0])[i];
    for (int r = 1; r < nin; r++) { const T t = static_cast<const T*>(in.p[r])[i]; acc = MAX ? (t > acc ? t : acc) : acc + t; }
    out[i] = acc;
}

struct LocalComm : DDComm {
    struct Slot {
        const void* send_lo = nullptr; const void* send_hi = nullptr; void* red = nullptr;
        cudaEvent_t ready = nullptr, done = nullptr;
        int device = -1;
        void* tmp = nullptr; size_t tmp_bytes = 0;
    };
    std::vector<Slot> slot;
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    unsigned long generation = 0;

    explicit LocalComm(int world_) { world = world_; slot.resize(world_); }
    ~LocalComm() override {
        for (auto& s : slot) {
            if (s.device >= 0) cudaSetDevice(s.device);
            if (s.ready) cudaEventDestroy(s.ready);
            if (s.done) cudaEventDestroy(s.done);
            if (s.tmp) cudaFree(s.tmp);
        }
    }
    const char* kind() const override { return "local"; }
    // called once per rank (from the thread creating that rank's context, with its device current)
    void attach(int rank, int device) {
        std::lock_guard<std::mutex> lk(mu);
        Slot& s = slot[rank];
        if (s.device >= 0) throw std::runtime_error("slab decomposition: rank attached twice to the in-process communicator");
        s.device = device;
        OCM_CUDA_OK(cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
        OCM_CUDA_OK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        for (auto& o : slot) {
            if (o.device < 0 || o.device == device) continue;
            int can = 0;
            OCM_CUDA_OK(cudaDeviceCanAccessPeer(&can, device, o.device));
            if (can) { cudaError_t e = cudaDeviceEnablePeerAccess(o.device, 0); if (e != cudaSuccess) cudaGetLastError(); }
            OCM_CUDA_OK(cudaSetDevice(o.device));
            OCM_CUDA_OK(cudaDeviceCanAccessPeer(&can, o.device, device));
            if (can) { cudaError_t e = cudaDeviceEnablePeerAccess(device, 0); if (e != cudaSuccess) cudaGetLastError(); }
            OCM_CUDA_OK(cudaSetDevice(device));
        }
    }
    void barrier() {
        std::unique_lock<std::mutex> lk(mu);
        const unsigned long gen = generation;
        if (++arrived == world) { arrived = 0; generation++; cv.notify_all(); }
        else if (!cv.wait_for(lk, std::chrono::seconds(120), [&] { return generation != gen; })) {
            arrived--;
            throw std::runtime_error("slab decomposition: a rank did not reach the exchange (in-process barrier timed out)");
        }
    }
    void ring_exchange(int r, const void* send_lo, const void* send_hi, void* recv_lo, void* recv_hi, size_t bytes, cudaStream_t s) override {
        const int prev = (r + world - 1) % world, next = (r + 1) % world;
        slot[r].send_lo = send_lo; slot[r].send_hi = send_hi;
        OCM_CUDA_OK(cudaEventRecord(slot[r].ready, s));
        barrier();
        OCM_CUDA_OK(cudaStreamWaitEvent(s, slot[prev].ready, 0));
        OCM_CUDA_OK(cudaStreamWaitEvent(s, slot[next].ready, 0));
        OCM_CUDA_OK(cudaMemcpyAsync(recv_lo, slot[prev].send_hi, bytes, cudaMemcpyDefault, s));
        OCM_CUDA_OK(cudaMemcpyAsync(recv_hi, slot[next].send_lo, bytes, cudaMemcpyDefault, s));
        OCM_CUDA_OK(cudaEventRecord(slot[r].done, s));
        barrier();
        // my send buffers are free again once both neighbours have copied them
        OCM_CUDA_OK(cudaStreamWaitEvent(s, slot[prev].done, 0));
        OCM_CUDA_OK(cudaStreamWaitEvent(s, slot[next].done, 0));
    }
    template <class T, bool MAX>
    void reduce_launch(int r, void* tmp, size_t count, cudaStream_t s) {
        PtrList in{};
        for (int q = 0; q < world; q++) in.p[q] = slot[q].red;
        dd_local_reduce_kernel<T, MAX><<<(unsigned int)((count + 255) / 256), 256, 0, s>>>(static_cast<T*>(tmp), in, world, count);
    }
    void allreduce(int r, void* buf, size_t count, DDType type, bool is_max, cudaStream_t s) override {
        const size_t esz = (type == DD_I32 || type == DD_F32) ? 4 : 8;
        Slot& me = slot[r];
        if (me.tmp_bytes < count * esz) {
            if (me.tmp) OCM_CUDA_OK(cudaFree(me.tmp));
            OCM_CUDA_OK(cudaMalloc(&me.tmp, count * esz));
            me.tmp_bytes = count * esz;
        }
        me.red = buf;
        OCM_CUDA_OK(cudaEventRecord(me.ready, s));
        barrier();
        for (int q = 0; q < world; q++) if (q != r) OCM_CUDA_OK(cudaStreamWaitEvent(s, slot[q].ready, 0));
        if (type == DD_I32) { if (is_max) reduce_launch<int, true>(r, me.tmp, count, s); else reduce_launch<int, false>(r, me.tmp, count, s); }
        else if (type == DD_I64) reduce_launch<long long, false>(r, me.tmp, count, s);
        else if (type == DD_F32) reduce_launch<float, false>(r, me.tmp, count, s);
        else reduce_launch<double, false>(r, me.tmp, count, s);
        OCM_CUDA_OK(cudaEventRecord(me.done, s));
        barrier();
        // every rank has read every input before anyone overwrites its own
        for (int q = 0; q < world; q++) if (q != r) OCM_CUDA_OK(cudaStreamWaitEvent(s, slot[q].done, 0));
        OCM_CUDA_OK(cudaMemcpyAsync(buf, me.tmp, count * esz, cudaMemcpyDeviceToDevice, s));
    }
};

}  // namespace ocm

struct ocm_comm {
    ocm::DDComm* impl = nullptr;
    int rank = -1;  // NCCL: the rank this communicator was created for; local: -1 (shared by all ranks)
    ~ocm_comm() { delete impl; }
};
