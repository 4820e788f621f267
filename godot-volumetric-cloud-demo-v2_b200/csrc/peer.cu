// Multi-GPU plumbing of the fused all-gather (include/cloudsky.h "multi-GPU").
//
// One process per GPU.  Each rank allocates its copy of the gathered output (and a small flag array) with plain cudaMalloc,
// exports it as a CUDA IPC handle, and maps every peer's copy into its own address space.  The march kernels then store each
// finished pixel into ALL copies (CloudLaunch::mirror, plain st.global on peer-mapped pointers — the stores travel over
// NVLink/NVSwitch while the SMs keep marching), so no collective kernel and no copy pass exists at all.  What is left of the
// all-gather is its completion barrier: one 32-thread kernel per rank that publishes "my frame k is complete in every copy"
// to each peer's flag array and waits until every peer has published the same.
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>

#include "cs_context.h"

using namespace cs;

namespace {

#define CU(call)                                                                                                          \
    do {                                                                                                                  \
        cudaError_t e__ = (call);                                                                                         \
        if (e__ != cudaSuccess) return ctx_fail(c, CS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));     \
    } while (0)

struct PeerFlags { unsigned* ptr[kMaxMirrors + 1]; };

// Thread k < world: (1) release-store `epoch` into slot `rank` of rank k's flag array — every store of this rank's preceding
// kernels (same stream) is complete by then, and the system-scope fence orders the flag after them as seen from any GPU;
// (2) spin on slot k of the own array until rank k has published `epoch` (wrap-safe compare).  A peer that never arrives
// trips the watchdog instead of hanging the GPU: *err is set and the kernel returns.
__global__ void peer_barrier_kernel(PeerFlags f, int rank, int world, unsigned epoch, unsigned* err, long long watchdog_cycles) {
    const int k = threadIdx.x;
    if (k >= world) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.ptr[k] + rank), "r"(epoch) : "memory");
    const unsigned* mine = f.ptr[rank] + k;
    const long long t0 = clock64();
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int)(v - epoch) >= 0) break;
        if (clock64() - t0 > watchdog_cycles) { atomicExch(err, 1u); break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

}  // namespace

extern "C" {

int cs_peer_alloc(cs_context* c, size_t bytes, void** dev_ptr, uint8_t handle[CS_IPC_HANDLE_BYTES]) {
    if (!c || !dev_ptr || !handle || bytes == 0) return CS_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == CS_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    CU(cudaSetDevice(c->device));
    void* p = nullptr;
    CU(cudaMalloc(&p, bytes));  // a plain allocation (not from a memory pool): exportable with cudaIpcGetMemHandle
    cudaError_t e = cudaMemsetAsync(p, 0, bytes, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return ctx_fail(c, CS_ERR_CUDA, std::string("cs_peer_alloc: ") + cudaGetErrorString(e));
    }
    memcpy(handle, &h, sizeof(h));
    *dev_ptr = p;
    return CS_OK;
}

int cs_peer_open(cs_context* c, const uint8_t handle[CS_IPC_HANDLE_BYTES], void** dev_ptr) {
    if (!c || !handle || !dev_ptr) return CS_ERR_INVALID;
    CU(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr = p;
    return CS_OK;
}

int cs_peer_close(cs_context* c, void* dev_ptr) {
    if (!c || !dev_ptr) return CS_ERR_INVALID;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaIpcCloseMemHandle(dev_ptr));
    return CS_OK;
}

int cs_peer_free(cs_context* c, void* dev_ptr) {
    if (!c || !dev_ptr) return CS_ERR_INVALID;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    if (c->mirror_base == dev_ptr) { c->mirror_base = nullptr; c->mirror_bytes = 0; c->n_mirrors = 0; }
    CU(cudaFree(dev_ptr));
    return CS_OK;
}

int cs_set_output_mirrors(cs_context* c, void* base, size_t bytes, int n_mirrors, void* const* mirror_bases) {
    if (!c || n_mirrors < 0 || n_mirrors > kMaxMirrors || (n_mirrors > 0 && (!base || !bytes || !mirror_bases)))
        return c ? ctx_fail(c, CS_ERR_INVALID, "cs_set_output_mirrors: 0..7 mirrors of a non-empty range") : CS_ERR_INVALID;
    for (int m = 0; m < n_mirrors; m++)
        if (!mirror_bases[m]) return ctx_fail(c, CS_ERR_INVALID, "cs_set_output_mirrors: null mirror");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));  // launches already queued keep the mirrors they were made with (passed by value)
    c->mirror_base = n_mirrors ? (uint8_t*)base : nullptr;
    c->mirror_bytes = n_mirrors ? bytes : 0;
    c->n_mirrors = n_mirrors;
    for (int m = 0; m < n_mirrors; m++) c->mirror_peer[m] = (uint8_t*)mirror_bases[m];
    return CS_OK;
}

int cs_peer_barrier(cs_context* c, int rank, int world, void* const* flag_arrays, uint32_t epoch) {
    if (!c || !flag_arrays || world < 1 || world > kMaxMirrors + 1 || rank < 0 || rank >= world) return CS_ERR_INVALID;
    CU(cudaSetDevice(c->device));
    if (!c->d_peer_err) {
        CU(cudaMalloc(&c->d_peer_err, sizeof(unsigned)));
        CU(cudaMemsetAsync(c->d_peer_err, 0, sizeof(unsigned), c->stream));
    }
    PeerFlags f{};
    for (int k = 0; k < world; k++) {
        if (!flag_arrays[k]) return ctx_fail(c, CS_ERR_INVALID, "cs_peer_barrier: null flag array");
        f.ptr[k] = (unsigned*)flag_arrays[k];
    }
    if (c->peer_watchdog_cycles == 0) {  // queried once: device-attribute calls take a driver-wide lock (milliseconds with 8 busy processes)
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
        long long ms = 20000;  // ~20 s of SM clocks; CLOUDSKY_PEER_WATCHDOG_MS overrides (tests use a short one)
        if (const char* e = getenv("CLOUDSKY_PEER_WATCHDOG_MS")) { long long v = atoll(e); if (v > 0) ms = v; }
        c->peer_watchdog_cycles = (long long)(khz > 0 ? khz : 2000000) * ms;
    }
    const long long watchdog = c->peer_watchdog_cycles;
    peer_barrier_kernel<<<1, 32, 0, c->stream>>>(f, rank, world, epoch, c->d_peer_err, watchdog);
    CU(cudaGetLastError());
    return CS_OK;
}

int cs_peer_check(cs_context* c) {
    if (!c) return CS_ERR_INVALID;
    if (!c->d_peer_err) return CS_OK;
    CU(cudaSetDevice(c->device));
    unsigned e = 0;
    CU(cudaMemcpyAsync(&e, c->d_peer_err, sizeof(e), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (e) {
        cudaMemsetAsync(c->d_peer_err, 0, sizeof(unsigned), c->stream);
        return ctx_fail(c, CS_ERR_NOT_READY, "cs_peer_barrier: a peer did not arrive within the watchdog time");
    }
    return CS_OK;
}

}  // extern "C"
