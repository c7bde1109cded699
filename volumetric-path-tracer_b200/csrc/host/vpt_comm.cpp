// vpt_comm.cpp -- the multi-GPU exchange step of the path, issued from the C++ host.
//
// One process per GPU.  The frame is sharded by interleaved row stripes (vpt_set_partition), the scene is replicated, and
// pixels are independent (every Philox stream is keyed by the GLOBAL pixel index), so a pass needs no data-path collective.
// The only exchange is the gather of the rank-local frames: one ncclAllGather of the accumulators (and optionally of the
// packed display words) per vpt_render_pass(es) call, followed by k_unpermute into the caller's full-frame buffer.  It is
// enqueued by vpt_render_passes itself, right behind the last resolve kernel:
//   * default: on the caller's stream (in order, nothing else to synchronise);
//   * option "gather_async" = 1: on the context's side stream behind an event, so that the NEXT call's generate / trace
//     kernels overlap it; the library makes the next call wait only where it would overwrite the accumulator, and
//     vpt_comm_wait() joins the side stream into the caller's stream.
//
// NCCL is bound at run time (dlopen of the copy the process already uses -- torch.distributed's -- else libnccl.so.2), so the
// library loads and renders on hosts without NCCL; only vpt_comm_* needs it.  The reference has no multi-GPU path: there
// is nothing to be compatible with here except "the gathered frame equals the single-GPU frame bit for bit" (tested).
#include "vpt_host.h"

#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include <string>

namespace {

typedef struct { char internal[128]; } nccl_unique_id;      // ncclUniqueId (nccl.h:37-38)
typedef void* nccl_comm_t;
typedef int nccl_result_t;                                    // ncclSuccess == 0
enum { NCCL_CHAR = 0, NCCL_FLOAT32 = 7 };                     // ncclDataType_t values (nccl.h)

struct NcclApi {
    void* handle = nullptr;
    nccl_result_t (*GetUniqueId)(nccl_unique_id*) = nullptr;
    nccl_result_t (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
    nccl_result_t (*CommDestroy)(nccl_comm_t) = nullptr;
    nccl_result_t (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(nccl_result_t) = nullptr;
    nccl_result_t (*GetVersion)(int*) = nullptr;
    std::string error;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = { "libnccl.so.2", "libnccl.so" };
        for (const char* n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); if (api.handle) break; }   // the copy already in the process
        for (const char* n : names) { if (api.handle) break; api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); }
        if (!api.handle) { api.error = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "?"); return; }
        auto sym = [&](const char* s) { void* p = dlsym(api.handle, s); if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + s; return p; };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    });
    return api;
}

int nccl_fail(vpt_context* c, const char* what, nccl_result_t r) {
    NcclApi& n = nccl();
    return vpt::fail_ctx(c, VPT_ERR_CUDA, std::string(what) + ": NCCL error " + std::to_string(r) + " (" + (n.GetErrorString ? n.GetErrorString(r) : "?") + ")");
}

} // namespace

#define VPT_CU(c, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return vpt::fail_ctx(c, VPT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

namespace vpt {

int comm_before_accum_write(vpt_context* c, cudaStream_t stream) {
    if (c->gather_pending) { VPT_CU(c, cudaStreamWaitEvent(stream, c->ev_gather_done, 0)); c->gather_pending = false; }
    return VPT_OK;
}

int comm_gather_frame(vpt_context* c, const FrameGeom& g, const void* d_local_accum, const void* d_local_display, cudaStream_t stream) {
    NcclApi& n = nccl();
    cudaStream_t s = stream;
    if (c->gather_async) {
        VPT_CU(c, cudaEventRecord(c->ev_render_done, stream));
        VPT_CU(c, cudaStreamWaitEvent(c->comm_stream, c->ev_render_done, 0));
        s = c->comm_stream;
    }
    const size_t px = (size_t)g.n_local;
    if (c->d_full_accum) {
        const size_t need = px * 12 * (size_t)g.n_ranks;
        if (need > c->cap_gathered) { cudaFree(c->d_gathered); c->d_gathered = nullptr; c->cap_gathered = 0; VPT_CU(c, cudaMalloc(&c->d_gathered, need)); c->cap_gathered = need; }
        nccl_result_t r = n.AllGather(d_local_accum, c->d_gathered, px * 3, NCCL_FLOAT32, (nccl_comm_t)c->nccl_comm, s);
        if (r != 0) return nccl_fail(c, "ncclAllGather(accum)", r);
        VPT_CU(c, launch_unpermute(c->d_gathered, c->d_full_accum, g, 12, s));
        c->launches++;
    }
    if (c->d_full_display) {
        const size_t need = px * 4 * (size_t)g.n_ranks;
        if (need > c->cap_gathered_disp) { cudaFree(c->d_gathered_disp); c->d_gathered_disp = nullptr; c->cap_gathered_disp = 0; VPT_CU(c, cudaMalloc(&c->d_gathered_disp, need)); c->cap_gathered_disp = need; }
        nccl_result_t r = n.AllGather(d_local_display, c->d_gathered_disp, px * 4, NCCL_CHAR, (nccl_comm_t)c->nccl_comm, s);
        if (r != 0) return nccl_fail(c, "ncclAllGather(display)", r);
        VPT_CU(c, launch_unpermute(c->d_gathered_disp, c->d_full_display, g, 4, s));
        c->launches++;
    }
    if (c->gather_async) { VPT_CU(c, cudaEventRecord(c->ev_gather_done, s)); c->gather_pending = true; }
    return VPT_OK;
}

// ---- peer-memory exchange ---------------------------------------------------------------------------------------------------------
static constexpr size_t kP2pHeader = 512;                      // ready[16] done[16] error[1] as u64
static size_t p2p_accum_bytes(const vpt_context* c) { return (((size_t)c->p2p_w * c->p2p_h * 12) + 255) & ~(size_t)255; }
static unsigned long long* p2p_flags(void* block) { return reinterpret_cast<unsigned long long*>(block); }

int comm_p2p_begin(vpt_context* c, cudaStream_t stream) {
    if (!c->p2p_on) return VPT_OK;
    ++c->p2p_epoch;
    unsigned long long* fl[kMaxPeers];
    for (int p = 0; p < kMaxPeers; ++p) fl[p] = p < c->p2p_n ? p2p_flags(c->p2p_peer[p]) : nullptr;
    VPT_CU(c, launch_peer_signal(fl, c->p2p_n, c->rank, 0, c->p2p_epoch, stream));     // "my frame of the previous call has been consumed"
    c->launches++;
    return VPT_OK;
}

int comm_p2p_peers(vpt_context* c, cudaStream_t stream, PeerFrames* out) {
    *out = PeerFrames{};
    if (!c->p2p_on) return VPT_OK;
    VPT_CU(c, launch_peer_wait(p2p_flags(c->p2p_block), c->p2p_n, 0, c->p2p_epoch, stream));   // every peer has entered this call
    c->launches++;
    out->n = c->p2p_n;
    for (int p = 0; p < c->p2p_n; ++p) {
        char* base = reinterpret_cast<char*>(c->p2p_peer[p]);
        out->accum[p] = reinterpret_cast<float*>(base + kP2pHeader);
        out->display[p] = c->p2p_display ? reinterpret_cast<unsigned int*>(base + kP2pHeader + p2p_accum_bytes(c)) : nullptr;
    }
    return VPT_OK;
}

int comm_p2p_end(vpt_context* c, cudaStream_t stream) {
    if (!c->p2p_on) return VPT_OK;
    unsigned long long* fl[kMaxPeers];
    for (int p = 0; p < kMaxPeers; ++p) fl[p] = p < c->p2p_n ? p2p_flags(c->p2p_peer[p]) : nullptr;
    VPT_CU(c, launch_peer_signal(fl, c->p2p_n, c->rank, 1, c->p2p_epoch, stream));     // my stripes are in every frame ...
    VPT_CU(c, launch_peer_wait(p2p_flags(c->p2p_block), c->p2p_n, 1, c->p2p_epoch, stream));   // ... and everybody's are in mine
    c->launches += 2;
    return VPT_OK;
}

} // namespace vpt

extern "C" {

int vpt_comm_p2p_export(vpt_context* c, int rank, int n_ranks, int stripe_rows, unsigned width, unsigned height, int with_display,
                        unsigned char handle_out[VPT_P2P_HANDLE_BYTES]) {
    if (!c || !handle_out || n_ranks < 1 || n_ranks > vpt::kMaxPeers || rank < 0 || rank >= n_ranks || stripe_rows < 1 || !width || !height)
        return vpt::fail_ctx(c, VPT_ERR_INVALID, "vpt_comm_p2p_export: bad arguments (at most 8 ranks: one NVSwitch domain)");
    if (c->p2p_block) return vpt::fail_ctx(c, VPT_ERR_INVALID, "vpt_comm_p2p_export: already exported");
    static_assert(sizeof(cudaIpcMemHandle_t) == VPT_P2P_HANDLE_BYTES, "IPC handle size");
    VPT_CU(c, cudaSetDevice(c->device));
    int rc = vpt_set_partition(c, rank, n_ranks, stripe_rows);
    if (rc != VPT_OK) return rc;
    c->p2p_w = (int)width; c->p2p_h = (int)height; c->p2p_display = with_display != 0; c->p2p_n = n_ranks;
    c->p2p_bytes = vpt::kP2pHeader + vpt::p2p_accum_bytes(c) + (with_display ? (size_t)width * height * 4 : 0);
    VPT_CU(c, cudaMalloc(&c->p2p_block, c->p2p_bytes));
    VPT_CU(c, cudaMemset(c->p2p_block, 0, c->p2p_bytes));
    VPT_CU(c, cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    VPT_CU(c, cudaIpcGetMemHandle(&h, c->p2p_block));
    memcpy(handle_out, &h, sizeof(h));
    return VPT_OK;
}

int vpt_comm_p2p_import(vpt_context* c, const unsigned char* handles) {
    if (!c || !handles || !c->p2p_block) return vpt::fail_ctx(c, VPT_ERR_INVALID, "vpt_comm_p2p_import: call vpt_comm_p2p_export first");
    for (int p = 0; p < c->p2p_n; ++p) {
        if (p == c->rank) { c->p2p_peer[p] = c->p2p_block; continue; }
        cudaIpcMemHandle_t h; memcpy(&h, handles + (size_t)p * VPT_P2P_HANDLE_BYTES, sizeof(h));
        VPT_CU(c, cudaIpcOpenMemHandle(&c->p2p_peer[p], h, cudaIpcMemLazyEnablePeerAccess));
    }
    c->p2p_on = true; c->p2p_epoch = 0;
    return VPT_OK;
}

int vpt_comm_p2p_import_local(vpt_context* c, const vpt_devptr_t* blocks) {
    if (!c || !blocks || !c->p2p_block) return vpt::fail_ctx(c, VPT_ERR_INVALID, "vpt_comm_p2p_import_local: call vpt_comm_p2p_export first");
    for (int p = 0; p < c->p2p_n; ++p) {
        if (p != c->rank && !blocks[p]) return vpt::fail_ctx(c, VPT_ERR_INVALID, "vpt_comm_p2p_import_local: null block");
        c->p2p_peer[p] = p == c->rank ? c->p2p_block : reinterpret_cast<void*>(blocks[p]);
    }
    c->p2p_local = true; c->p2p_on = true; c->p2p_epoch = 0;
    return VPT_OK;
}

int vpt_comm_p2p_block(vpt_context* c, vpt_devptr_t* d_block) {
    if (!c || !d_block || !c->p2p_block) return vpt::fail_ctx(c, VPT_ERR_INVALID, "vpt_comm_p2p_block: no exchange block");
    *d_block = (vpt_devptr_t)c->p2p_block;
    return VPT_OK;
}

int vpt_comm_p2p_frame(vpt_context* c, vpt_devptr_t* d_full_accum, vpt_devptr_t* d_full_display) {
    if (!c || !c->p2p_block) return vpt::fail_ctx(c, VPT_ERR_INVALID, "vpt_comm_p2p_frame: no exchange block");
    char* base = reinterpret_cast<char*>(c->p2p_block);
    if (d_full_accum) *d_full_accum = (vpt_devptr_t)(base + vpt::kP2pHeader);
    if (d_full_display) *d_full_display = c->p2p_display ? (vpt_devptr_t)(base + vpt::kP2pHeader + vpt::p2p_accum_bytes(c)) : 0;
    return VPT_OK;
}

int vpt_comm_p2p_enable(vpt_context* c, int on) {
    if (!c || !c->p2p_block || !c->p2p_peer[c->rank]) return vpt::fail_ctx(c, VPT_ERR_INVALID, "vpt_comm_p2p_enable: exchange not set up");
    c->p2p_on = on != 0;
    return VPT_OK;
}

int vpt_comm_p2p_status(vpt_context* c, unsigned long long* timeouts) {
    if (!c || !timeouts || !c->p2p_block) return vpt::fail_ctx(c, VPT_ERR_INVALID, "vpt_comm_p2p_status: no exchange block");
    VPT_CU(c, cudaMemcpy(timeouts, vpt::p2p_flags(c->p2p_block) + 2 * vpt::kPeerFlagStride, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return VPT_OK;
}

int vpt_comm_get_unique_id(unsigned char id_out[VPT_COMM_ID_BYTES]) {
    if (!id_out) return vpt::fail_global(VPT_ERR_INVALID, "vpt_comm_get_unique_id: null argument");
    NcclApi& n = nccl();
    if (!n.error.empty()) return vpt::fail_global(VPT_ERR_UNSUPPORTED, "vpt_comm_get_unique_id: " + n.error);
    nccl_unique_id id; memset(&id, 0, sizeof(id));
    nccl_result_t r = n.GetUniqueId(&id);
    if (r != 0) return vpt::fail_global(VPT_ERR_CUDA, std::string("ncclGetUniqueId: ") + n.GetErrorString(r));
    memcpy(id_out, id.internal, VPT_COMM_ID_BYTES);
    return VPT_OK;
}

int vpt_comm_init(vpt_context* c, const unsigned char id[VPT_COMM_ID_BYTES], int rank, int n_ranks, int stripe_rows) {
    if (!c || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks || stripe_rows < 1) return vpt::fail_ctx(c, VPT_ERR_INVALID, "vpt_comm_init: bad arguments");
    if (c->nccl_comm) return vpt::fail_ctx(c, VPT_ERR_INVALID, "vpt_comm_init: communicator already initialised");
    NcclApi& n = nccl();
    if (!n.error.empty()) return vpt::fail_ctx(c, VPT_ERR_UNSUPPORTED, "vpt_comm_init: " + n.error);
    VPT_CU(c, cudaSetDevice(c->device));
    nccl_unique_id uid; memcpy(uid.internal, id, VPT_COMM_ID_BYTES);
    nccl_comm_t comm = nullptr;
    nccl_result_t r = n.CommInitRank(&comm, n_ranks, uid, rank);
    if (r != 0) return nccl_fail(c, "ncclCommInitRank", r);
    c->nccl_comm = comm;
    VPT_CU(c, cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
    VPT_CU(c, cudaEventCreateWithFlags(&c->ev_render_done, cudaEventDisableTiming));
    VPT_CU(c, cudaEventCreateWithFlags(&c->ev_gather_done, cudaEventDisableTiming));
    return vpt_set_partition(c, rank, n_ranks, stripe_rows);
}

int vpt_comm_set_gather(vpt_context* c, void* d_full_accum, void* d_full_display) {
    if (!c) return VPT_ERR_INVALID;
    if ((d_full_accum || d_full_display) && !c->nccl_comm) return vpt::fail_ctx(c, VPT_ERR_INVALID, "vpt_comm_set_gather: call vpt_comm_init first");
    c->d_full_accum = d_full_accum; c->d_full_display = d_full_display;
    return VPT_OK;
}

int vpt_comm_wait(vpt_context* c, void* stream) {
    if (!c) return VPT_ERR_INVALID;
    return vpt::comm_before_accum_write(c, (cudaStream_t)stream);
}

int vpt_comm_info(vpt_context* c, int* nccl_version, int* rank, int* n_ranks) {
    if (!c) return VPT_ERR_INVALID;
    NcclApi& n = nccl();
    if (nccl_version) { int v = 0; if (n.GetVersion) n.GetVersion(&v); *nccl_version = v; }
    if (rank) *rank = c->rank;
    if (n_ranks) *n_ranks = c->nccl_comm ? c->n_ranks : 1;
    return VPT_OK;
}

int vpt_comm_destroy(vpt_context* c) {
    if (!c) return VPT_ERR_INVALID;
    if (c->nccl_comm) {
        if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
        nccl().CommDestroy((nccl_comm_t)c->nccl_comm); c->nccl_comm = nullptr;
    }
    if (c->comm_stream) { cudaStreamDestroy(c->comm_stream); c->comm_stream = nullptr; }
    if (c->ev_render_done) { cudaEventDestroy(c->ev_render_done); c->ev_render_done = nullptr; }
    if (c->ev_gather_done) { cudaEventDestroy(c->ev_gather_done); c->ev_gather_done = nullptr; }
    cudaFree(c->d_gathered); c->d_gathered = nullptr; c->cap_gathered = 0;
    cudaFree(c->d_gathered_disp); c->d_gathered_disp = nullptr; c->cap_gathered_disp = 0;
    c->d_full_accum = c->d_full_display = nullptr; c->gather_pending = false;
    if (c->p2p_block) {
        cudaDeviceSynchronize();
        if (!c->p2p_local) for (int p = 0; p < c->p2p_n; ++p) if (p != c->rank && c->p2p_peer[p]) cudaIpcCloseMemHandle(c->p2p_peer[p]);
        cudaFree(c->p2p_block); c->p2p_block = nullptr; c->p2p_on = false; c->p2p_n = 0;
        for (auto& q : c->p2p_peer) q = nullptr;
    }
    return VPT_OK;
}

} // extern "C"
