// comm.cuh -- NCCL plumbing of a sharded run, called from the library's own round loops
// (bdr_slab_rounds / bdr_slab_refine): halo planes of the label / known arrays move
// straight between the ranks' window arrays with grouped ncclSend / ncclRecv on the
// handle's stream, counters meet in one ncclAllReduce per decision.  The ring of slabs
// is periodic: rank r sends its top owned planes to r+1's low halo and its bottom owned
// planes to r-1's high halo.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a process that already loaded
// torch's copy gets that one, the single-GPU library has no NCCL dependency at all.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace bdr {

struct NcclApi {
    void *lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

inline NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (lib) {
            api.lib = lib;
#define BDR_NCCL_SYM(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(lib, "nccl" #name))
            BDR_NCCL_SYM(GetUniqueId);
            BDR_NCCL_SYM(CommInitRank);
            BDR_NCCL_SYM(CommDestroy);
            BDR_NCCL_SYM(GroupStart);
            BDR_NCCL_SYM(GroupEnd);
            BDR_NCCL_SYM(Send);
            BDR_NCCL_SYM(Recv);
            BDR_NCCL_SYM(AllReduce);
            BDR_NCCL_SYM(GetErrorString);
#undef BDR_NCCL_SYM
            if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.GroupStart || !api.GroupEnd ||
                !api.Send || !api.Recv || !api.AllReduce || !api.GetErrorString)
                api.lib = nullptr;
        }
    }
    return api.lib ? &api : nullptr;
}

struct SlabComm {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0, prev = 0, next = 0;
    unsigned long long *d_red = nullptr;   // device scratch of the counter all-reduce
    unsigned long long *h_red = nullptr;   // pinned mirror
    int32_t *plane_lo = nullptr, *plane_hi = nullptr;   // label planes next to the slab before an exchange
};

#define NC(x)                                                                              \
    do {                                                                                   \
        ncclResult_t r_ = (x);                                                             \
        if (r_ != ncclSuccess)                                                             \
            return bdr::fail_msg(std::string(#x) + " failed: " + bdr::nccl_api()->GetErrorString(r_)); \
    } while (0)

// fill the halo planes of a window array [W][plane] of `elem` bytes per voxel from the owners
inline int comm_halo_exchange(bdr_ctx *c, SlabComm *sc, void *base, int elem) {
    const int64_t plane = (int64_t)c->g.ny * c->g.nz;
    const int H = c->halo, W = c->g.nx, n = W - 2 * H;
    char *b = static_cast<char *>(base);
    const size_t bytes = (size_t)H * plane * elem;
    char *send_up = b + (size_t)n * plane * elem;          // top owned planes [n, n+H) -> next's low halo
    char *send_down = b + (size_t)H * plane * elem;        // bottom owned planes [H, 2H) -> prev's high halo
    char *recv_lo = b, *recv_hi = b + (size_t)(W - H) * plane * elem;
    if (sc->world == 1) {
        CU(cudaMemcpyAsync(recv_lo, send_up, bytes, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(recv_hi, send_down, bytes, cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    NcclApi *api = nccl_api();
    NC(api->GroupStart());
    NC(api->Send(send_up, bytes, ncclInt8, sc->next, sc->comm, c->stream));
    NC(api->Send(send_down, bytes, ncclInt8, sc->prev, sc->comm, c->stream));
    NC(api->Recv(recv_lo, bytes, ncclInt8, sc->prev, sc->comm, c->stream));
    NC(api->Recv(recv_hi, bytes, ncclInt8, sc->next, sc->comm, c->stream));
    NC(api->GroupEnd());
    return 0;
}

// sum of n (<= 8) counters over the ranks; completes on a rank only after every rank has
// enqueued it behind its own kernels, so it is also the barrier the peer loads of the
// trace kernel need ("every rank's classification is complete")
inline int comm_allreduce(bdr_ctx *c, SlabComm *sc, int n, long long *vals) {
    if (sc->world == 1) {
        CU(cudaStreamSynchronize(c->stream));
        return 0;
    }
    for (int i = 0; i < n; ++i) sc->h_red[i] = (unsigned long long)vals[i];
    CU(cudaMemcpyAsync(sc->d_red, sc->h_red, (size_t)n * sizeof(unsigned long long), cudaMemcpyHostToDevice,
                       c->stream));
    NC(nccl_api()->AllReduce(sc->d_red, sc->d_red, (size_t)n, ncclUint64, ncclSum, sc->comm, c->stream));
    CU(cudaMemcpyAsync(sc->h_red, sc->d_red, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                       c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; ++i) vals[i] = (long long)sc->h_red[i];
    return 0;
}

}  // namespace bdr
