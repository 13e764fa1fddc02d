// libpyfr_b200.so -- C-ABI runtime of the B200 backend (see
// include/pyfr_b200.h).  The CUDA runtime is linked statically; the driver,
// NVRTC and NCCL are bound lazily with dlopen so the library loads (and its
// symbol table can be checked) on machines without a GPU.

#include "../../include/pyfr_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace {

thread_local std::string g_err;

int fail(const std::string &msg) {
    g_err = msg;
    return 1;
}

int check(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    return fail(std::string(what) + ": " + cudaGetErrorName(e) + " (" +
                cudaGetErrorString(e) + ")");
}

#define RT(call) do { if (int rc_ = check((call), #call)) return rc_; } while (0)

// ---- lazily bound driver API --------------------------------------------
typedef int CUresult;
typedef void *CUmodule;
typedef void *CUfunction;
typedef void *CUstream;

struct Driver {
    void *lib = nullptr;
    CUresult (*cuInit)(unsigned) = nullptr;
    CUresult (*cuGetErrorString)(CUresult, const char **) = nullptr;
    CUresult (*cuModuleLoadData)(CUmodule *, const void *) = nullptr;
    CUresult (*cuModuleUnload)(CUmodule) = nullptr;
    CUresult (*cuModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
    CUresult (*cuFuncSetAttribute)(CUfunction, int, int) = nullptr;
    CUresult (*cuFuncGetAttribute)(int *, int, CUfunction) = nullptr;
    CUresult (*cuLaunchKernel)(CUfunction, unsigned, unsigned, unsigned,
                               unsigned, unsigned, unsigned, unsigned,
                               CUstream, void **, void **) = nullptr;
} drv;

template <class F> bool bind(void *lib, F &fn, const char *name) {
    fn = reinterpret_cast<F>(dlsym(lib, name));
    return fn != nullptr;
}

int load_driver() {
    if (drv.lib) return 0;

    void *lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail(std::string("dlopen libcuda.so.1: ") + dlerror());

    bool ok = bind(lib, drv.cuInit, "cuInit") &&
              bind(lib, drv.cuGetErrorString, "cuGetErrorString") &&
              bind(lib, drv.cuModuleLoadData, "cuModuleLoadData") &&
              bind(lib, drv.cuModuleUnload, "cuModuleUnload") &&
              bind(lib, drv.cuModuleGetFunction, "cuModuleGetFunction") &&
              bind(lib, drv.cuFuncSetAttribute, "cuFuncSetAttribute") &&
              bind(lib, drv.cuFuncGetAttribute, "cuFuncGetAttribute") &&
              bind(lib, drv.cuLaunchKernel, "cuLaunchKernel");
    if (!ok) return fail("libcuda.so.1 lacks a required entry point");

    drv.lib = lib;
    return 0;
}

int dcheck(CUresult r, const char *what) {
    if (r == 0) return 0;
    const char *s = nullptr;
    if (drv.cuGetErrorString) drv.cuGetErrorString(r, &s);
    return fail(std::string(what) + ": CUDA driver error " +
                std::to_string(r) + " (" + (s ? s : "?") + ")");
}

#define DRV(call) do { if (int rc_ = dcheck((call), #call)) return rc_; } while (0)

// ---- lazily bound NVRTC ---------------------------------------------------
typedef void *nvrtcProgram;

struct Nvrtc {
    void *lib = nullptr;
    int (*create)(nvrtcProgram *, const char *, const char *, int,
                  const char *const *, const char *const *) = nullptr;
    int (*destroy)(nvrtcProgram *) = nullptr;
    int (*compile)(nvrtcProgram, int, const char *const *) = nullptr;
    int (*cubin_size)(nvrtcProgram, size_t *) = nullptr;
    int (*cubin)(nvrtcProgram, char *) = nullptr;
    int (*log_size)(nvrtcProgram, size_t *) = nullptr;
    int (*log)(nvrtcProgram, char *) = nullptr;
    const char *(*errstr)(int) = nullptr;
} rtc;

int load_nvrtc() {
    if (rtc.lib) return 0;

    void *lib = nullptr;
    for (const char *n : {"libnvrtc.so.12", "libnvrtc.so",
                          "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
        if ((lib = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
    }
    if (!lib) return fail(std::string("dlopen libnvrtc: ") + dlerror());

    bool ok = bind(lib, rtc.create, "nvrtcCreateProgram") &&
              bind(lib, rtc.destroy, "nvrtcDestroyProgram") &&
              bind(lib, rtc.compile, "nvrtcCompileProgram") &&
              bind(lib, rtc.cubin_size, "nvrtcGetCUBINSize") &&
              bind(lib, rtc.cubin, "nvrtcGetCUBIN") &&
              bind(lib, rtc.log_size, "nvrtcGetProgramLogSize") &&
              bind(lib, rtc.log, "nvrtcGetProgramLog") &&
              bind(lib, rtc.errstr, "nvrtcGetErrorString");
    if (!ok) return fail("libnvrtc lacks a required entry point");

    rtc.lib = lib;
    return 0;
}

// ---- lazily bound NCCL ------------------------------------------------------
typedef void *ncclComm_t;
struct ncclUniqueId { char internal[128]; };

struct Nccl {
    void *lib = nullptr;
    int (*get_unique_id)(ncclUniqueId *) = nullptr;
    int (*comm_init_rank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*comm_destroy)(ncclComm_t) = nullptr;
    int (*group_start)() = nullptr;
    int (*group_end)() = nullptr;
    int (*send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*allreduce)(const void *, void *, size_t, int, int, ncclComm_t,
                     cudaStream_t) = nullptr;
    const char *(*errstr)(int) = nullptr;
} nccl;

int load_nccl() {
    if (nccl.lib) return 0;

    void *lib = nullptr;
    const char *env = getenv("PYFR_B200_NCCL_LIBRARY");
    if (env) lib = dlopen(env, RTLD_NOW | RTLD_LOCAL);
    for (const char *n : {"libnccl.so.2", "libnccl.so"}) {
        if (lib) break;
        lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    }
    if (!lib) return fail(std::string("dlopen libnccl: ") + dlerror());

    bool ok = bind(lib, nccl.get_unique_id, "ncclGetUniqueId") &&
              bind(lib, nccl.comm_init_rank, "ncclCommInitRank") &&
              bind(lib, nccl.comm_destroy, "ncclCommDestroy") &&
              bind(lib, nccl.group_start, "ncclGroupStart") &&
              bind(lib, nccl.group_end, "ncclGroupEnd") &&
              bind(lib, nccl.send, "ncclSend") &&
              bind(lib, nccl.recv, "ncclRecv") &&
              bind(lib, nccl.allreduce, "ncclAllReduce") &&
              bind(lib, nccl.errstr, "ncclGetErrorString");
    if (!ok) return fail("libnccl lacks a required entry point");

    nccl.lib = lib;
    return 0;
}

int ncheck(int r, const char *what) {
    if (r == 0) return 0;
    return fail(std::string(what) + ": NCCL error " + std::to_string(r) +
                " (" + (nccl.errstr ? nccl.errstr(r) : "?") + ")");
}

#define NC(call) do { if (int rc_ = ncheck((call), #call)) return rc_; } while (0)

// ncclDataType_t: float32 = 7, float64 = 8; ncclRedOp_t: sum 0, max 2, min 3
int nccl_dtype(int dtype) { return dtype == 0 ? 7 : 8; }

cudaStream_t S(void *s) { return static_cast<cudaStream_t>(s); }
cudaEvent_t E(void *e) { return static_cast<cudaEvent_t>(e); }

}  // namespace

extern "C" {

const char *b200_last_error(void) { return g_err.c_str(); }

int b200_init(int device) {
    RT(cudaSetDevice(device));
    RT(cudaFree(nullptr));          // force context creation
    if (int rc = load_driver()) return rc;
    DRV(drv.cuInit(0));
    return 0;
}

int b200_device_info(int *sm_count, int *cc_major, int *cc_minor,
                     size_t *total_mem, size_t *free_mem, size_t *smem_optin) {
    int dev = 0, v = 0;
    RT(cudaGetDevice(&dev));
    RT(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    RT(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    RT(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    RT(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    *smem_optin = static_cast<size_t>(v);
    RT(cudaMemGetInfo(free_mem, total_mem));
    return 0;
}

// ---- memory ---------------------------------------------------------------
int b200_malloc(void **dptr, size_t nbytes) {
    RT(cudaMalloc(dptr, nbytes ? nbytes : 1));
    return 0;
}

int b200_free(void *dptr) { RT(cudaFree(dptr)); return 0; }

int b200_malloc_host(void **hptr, size_t nbytes) {
    RT(cudaMallocHost(hptr, nbytes ? nbytes : 1));
    return 0;
}

int b200_free_host(void *hptr) { RT(cudaFreeHost(hptr)); return 0; }

int b200_memset(void *dptr, int value, size_t nbytes, void *stream) {
    RT(cudaMemsetAsync(dptr, value, nbytes, S(stream)));
    return 0;
}

int b200_memcpy(void *dst, const void *src, size_t nbytes) {
    RT(cudaMemcpy(dst, src, nbytes, cudaMemcpyDefault));
    return 0;
}

int b200_memcpy_async(void *dst, const void *src, size_t nbytes, void *stream) {
    RT(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDefault, S(stream)));
    return 0;
}

int b200_memcpy2d_async(void *dst, size_t dpitch, const void *src,
                        size_t spitch, size_t width, size_t height,
                        void *stream) {
    RT(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height,
                         cudaMemcpyDefault, S(stream)));
    return 0;
}

// ---- streams & events -------------------------------------------------------
int b200_stream_create(void **stream) {
    cudaStream_t s;
    RT(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return 0;
}

int b200_stream_create_priority(void **stream, int high) {
    int lo = 0, hi = 0;
    RT(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    cudaStream_t s;
    RT(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high ? hi : lo));
    *stream = s;
    return 0;
}

int b200_stream_destroy(void *stream) { RT(cudaStreamDestroy(S(stream))); return 0; }
int b200_stream_sync(void *stream) { RT(cudaStreamSynchronize(S(stream))); return 0; }
int b200_device_sync(void) { RT(cudaDeviceSynchronize()); return 0; }

int b200_event_create(void **event) {
    cudaEvent_t e;
    RT(cudaEventCreate(&e));
    *event = e;
    return 0;
}

int b200_event_destroy(void *event) { RT(cudaEventDestroy(E(event))); return 0; }
int b200_event_record(void *event, void *stream) {
    RT(cudaEventRecord(E(event), S(stream)));
    return 0;
}
int b200_event_sync(void *event) { RT(cudaEventSynchronize(E(event))); return 0; }
int b200_event_elapsed_ms(float *ms, void *start, void *stop) {
    RT(cudaEventElapsedTime(ms, E(start), E(stop)));
    return 0;
}
int b200_stream_wait_event(void *stream, void *event) {
    RT(cudaStreamWaitEvent(S(stream), E(event), 0));
    return 0;
}

// ---- compilation & launch -----------------------------------------------------
int b200_nvrtc_compile(const char *src, const char *name,
                       const char *const *opts, int nopts,
                       void **image, size_t *image_size, char **log) {
    if (int rc = load_nvrtc()) return rc;

    *image = nullptr;
    *image_size = 0;
    if (log) *log = nullptr;

    nvrtcProgram prog;
    int r = rtc.create(&prog, src, name, 0, nullptr, nullptr);
    if (r) return fail(std::string("nvrtcCreateProgram: ") + rtc.errstr(r));

    int cr = rtc.compile(prog, nopts, opts);

    size_t ln = 0;
    rtc.log_size(prog, &ln);
    std::string lg(ln ? ln : 1, '\0');
    if (ln > 1) rtc.log(prog, &lg[0]);
    if (log && ln > 1) {
        *log = static_cast<char *>(malloc(ln + 1));
        memcpy(*log, lg.c_str(), ln);
        (*log)[ln] = 0;
    }

    if (cr) {
        rtc.destroy(&prog);
        return fail(std::string("nvrtcCompileProgram: ") + rtc.errstr(cr) +
                    "\n" + lg.c_str());
    }

    size_t n = 0;
    r = rtc.cubin_size(prog, &n);
    if (r || !n) {
        rtc.destroy(&prog);
        return fail("nvrtcGetCUBINSize failed (need a real -arch=sm_XXX)");
    }

    char *buf = static_cast<char *>(malloc(n));
    r = rtc.cubin(prog, buf);
    rtc.destroy(&prog);
    if (r) {
        free(buf);
        return fail(std::string("nvrtcGetCUBIN: ") + rtc.errstr(r));
    }

    *image = buf;
    *image_size = n;
    return 0;
}

int b200_buffer_free(void *buf) { free(buf); return 0; }

int b200_module_load(void **module, const void *image) {
    if (int rc = load_driver()) return rc;
    DRV(drv.cuModuleLoadData(module, image));
    return 0;
}

int b200_module_unload(void *module) {
    if (int rc = load_driver()) return rc;
    DRV(drv.cuModuleUnload(module));
    return 0;
}

int b200_module_get_function(void **func, void *module, const char *name) {
    if (int rc = load_driver()) return rc;
    DRV(drv.cuModuleGetFunction(func, module, name));
    return 0;
}

int b200_function_set_dynamic_smem(void *func, int nbytes) {
    if (int rc = load_driver()) return rc;
    // CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES = 8
    DRV(drv.cuFuncSetAttribute(func, 8, nbytes));
    return 0;
}

int b200_function_info(void *func, int *nregs, int *static_smem,
                       int *local_bytes, int *max_threads) {
    if (int rc = load_driver()) return rc;
    // MAX_THREADS_PER_BLOCK 0, SHARED_SIZE_BYTES 1, LOCAL_SIZE_BYTES 3,
    // NUM_REGS 4
    DRV(drv.cuFuncGetAttribute(max_threads, 0, func));
    DRV(drv.cuFuncGetAttribute(static_smem, 1, func));
    DRV(drv.cuFuncGetAttribute(local_bytes, 3, func));
    DRV(drv.cuFuncGetAttribute(nregs, 4, func));
    return 0;
}

int b200_launch(void *func, unsigned gx, unsigned gy, unsigned gz,
                unsigned bx, unsigned by, unsigned bz, unsigned smem_bytes,
                void *stream, void **args) {
    DRV(drv.cuLaunchKernel(func, gx, gy, gz, bx, by, bz, smem_bytes, stream,
                           args, nullptr));
    return 0;
}

// ---- graphs -------------------------------------------------------------------
int b200_capture_begin(void *stream) {
    RT(cudaStreamBeginCapture(S(stream), cudaStreamCaptureModeThreadLocal));
    return 0;
}

int b200_capture_end(void *stream, void **graph_exec) {
    cudaGraph_t g;
    RT(cudaStreamEndCapture(S(stream), &g));

    cudaGraphExec_t ge;
    cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphDestroy(g);
    RT(e);

    *graph_exec = ge;
    return 0;
}

int b200_graph_launch(void *graph_exec, void *stream) {
    RT(cudaGraphLaunch(static_cast<cudaGraphExec_t>(graph_exec), S(stream)));
    return 0;
}

int b200_graph_destroy(void *graph_exec) {
    RT(cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(graph_exec)));
    return 0;
}

// ---- NCCL ---------------------------------------------------------------------
int b200_nccl_unique_id(char id[B200_NCCL_ID_BYTES]) {
    if (int rc = load_nccl()) return rc;
    ncclUniqueId u;
    NC(nccl.get_unique_id(&u));
    memcpy(id, u.internal, sizeof(u.internal));
    return 0;
}

int b200_nccl_init(void **comm, int nranks, int rank,
                   const char id[B200_NCCL_ID_BYTES]) {
    if (int rc = load_nccl()) return rc;
    ncclUniqueId u;
    memcpy(u.internal, id, sizeof(u.internal));
    ncclComm_t c;
    NC(nccl.comm_init_rank(&c, nranks, u, rank));
    *comm = c;
    return 0;
}

int b200_nccl_destroy(void *comm) {
    if (int rc = load_nccl()) return rc;
    NC(nccl.comm_destroy(comm));
    return 0;
}

int b200_nccl_group_start(void) {
    if (int rc = load_nccl()) return rc;
    NC(nccl.group_start());
    return 0;
}

int b200_nccl_group_end(void) {
    if (int rc = load_nccl()) return rc;
    NC(nccl.group_end());
    return 0;
}

int b200_nccl_send(void *comm, const void *buf, size_t count, int dtype,
                   int peer, void *stream) {
    NC(nccl.send(buf, count, nccl_dtype(dtype), peer, comm, S(stream)));
    return 0;
}

int b200_nccl_recv(void *comm, void *buf, size_t count, int dtype, int peer,
                   void *stream) {
    NC(nccl.recv(buf, count, nccl_dtype(dtype), peer, comm, S(stream)));
    return 0;
}

int b200_nccl_allreduce(void *comm, const void *sendbuf, void *recvbuf,
                        size_t count, int dtype, int op, void *stream) {
    NC(nccl.allreduce(sendbuf, recvbuf, count, nccl_dtype(dtype), op, comm,
                      S(stream)));
    return 0;
}

}  // extern "C"
