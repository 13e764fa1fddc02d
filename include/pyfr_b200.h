/*
 * pyfr_b200.h -- C ABI of libpyfr_b200.so, the B200 (sm_100a) execution
 * runtime behind the `b200` PyFR backend.
 *
 * Every entry point takes plain pointers / integers and returns an int
 * status (0 = success; on failure b200_last_error() describes it).  The
 * library keeps one CUDA context per process and is driven by a single
 * host thread, like the reference's backends (one stream per backend,
 * pyfr/backends/cuda/base.py:92).
 *
 * Each group cites the reference interface it replaces; the reference
 * reaches the same services through ctypes wrappers over libcuda, NVRTC
 * and MPI rather than through a library of its own.
 */
#ifndef PYFR_B200_H
#define PYFR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* -- context & errors  (cuInit / primary context: pyfr/backends/cuda/
 *    driver.py:143-153; error-code-to-exception table :107-141;
 *    CUDABackend.__init__ pyfr/backends/cuda/base.py:20-49) ------------- */
int b200_init(int device);
const char *b200_last_error(void);
int b200_device_info(int *sm_count, int *cc_major, int *cc_minor,
                     size_t *total_mem, size_t *free_mem, size_t *smem_optin);

/* -- memory  (cuMemAlloc, cuMemFree, cuMemAllocHost, cuMemcpy[Async],
 *    cuMemsetD8[Async]: driver.py:154-161; CUDADevAlloc / CUDAHostAlloc
 *    :235-255) ------------------------------------------------------------ */
int b200_malloc(void **dptr, size_t nbytes);
int b200_free(void *dptr);
int b200_malloc_host(void **hptr, size_t nbytes);
int b200_free_host(void *hptr);
int b200_memset(void *dptr, int value, size_t nbytes, void *stream);
int b200_memcpy(void *dst, const void *src, size_t nbytes);
int b200_memcpy_async(void *dst, const void *src, size_t nbytes, void *stream);
int b200_memcpy2d_async(void *dst, size_t dpitch, const void *src,
                        size_t spitch, size_t width, size_t height,
                        void *stream);

/* -- streams & events  (driver.py:162-171; CUDAStream :258-277,
 *    CUDAEvent :280-304) -------------------------------------------------- */
int b200_stream_create(void **stream);
/* high != 0: greatest priority, so that exchange kernels launched beside a
 * long-running grid are dispatched as soon as SM resources free up */
int b200_stream_create_priority(void **stream, int high);
int b200_stream_destroy(void *stream);
int b200_stream_sync(void *stream);
int b200_device_sync(void);
int b200_event_create(void **event);
int b200_event_destroy(void *event);
int b200_event_record(void *event, void *stream);
int b200_event_sync(void *event);
int b200_event_elapsed_ms(float *ms, void *start, void *stop);
int b200_stream_wait_event(void *stream, void *event);

/* -- run-time compilation and kernel launch  (NVRTC wrapper
 *    pyfr/backends/cuda/compiler.py:22-158; cuModuleLoadDataEx,
 *    cuModuleGetFunction, cuLaunchKernel, cuFunc{Get,Set}Attribute:
 *    driver.py:172-179; CUDAModule / CUDAFunction :307-360) --------------- */
int b200_nvrtc_compile(const char *src, const char *name,
                       const char *const *opts, int nopts,
                       void **image, size_t *image_size, char **log);
int b200_buffer_free(void *buf);
int b200_module_load(void **module, const void *image);
int b200_module_unload(void *module);
int b200_module_get_function(void **func, void *module, const char *name);
int b200_function_set_dynamic_smem(void *func, int nbytes);
int b200_function_info(void *func, int *nregs, int *static_smem,
                       int *local_bytes, int *max_threads);
int b200_launch(void *func, unsigned gx, unsigned gy, unsigned gz,
                unsigned bx, unsigned by, unsigned bz, unsigned smem_bytes,
                void *stream, void **args);

/* -- CUDA graphs via stream capture  (cuStreamBegin/EndCapture and the
 *    cuGraph* node builders: driver.py:164-165,180-199; CUDAGraph /
 *    CUDAExecGraph :363-470; backend graph pyfr/backends/cuda/types.py:
 *    79-116) -------------------------------------------------------------- */
int b200_capture_begin(void *stream);
int b200_capture_end(void *stream, void **graph_exec);
int b200_graph_launch(void *graph_exec, void *stream);
int b200_graph_destroy(void *graph_exec);

/* -- inter-partition exchange over NCCL  (replaces the persistent MPI
 *    requests of XchgMatrix.sendreq/recvreq, pyfr/backends/base/types.py:
 *    250-257, started from CUDAGraph.run, cuda/types.py:99-116) --------- */
#define B200_NCCL_ID_BYTES 128
int b200_nccl_unique_id(char id[B200_NCCL_ID_BYTES]);
int b200_nccl_init(void **comm, int nranks, int rank,
                   const char id[B200_NCCL_ID_BYTES]);
int b200_nccl_destroy(void *comm);
int b200_nccl_group_start(void);
int b200_nccl_group_end(void);
/* dtype: 0 = float32, 1 = float64 */
int b200_nccl_send(void *comm, const void *buf, size_t count, int dtype,
                   int peer, void *stream);
int b200_nccl_recv(void *comm, void *buf, size_t count, int dtype, int peer,
                   void *stream);
int b200_nccl_allreduce(void *comm, const void *sendbuf, void *recvbuf,
                        size_t count, int dtype, int op, void *stream);

#ifdef __cplusplus
}
#endif

#endif /* PYFR_B200_H */
