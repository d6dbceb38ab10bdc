// Library-level entry points: error text, device info, engine switch, launch counter.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <set>
#include <utility>

#include "common.cuh"

namespace gnngp {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};
static std::atomic<int64_t> g_tc_launches{0};
std::atomic<int> g_gemm_engine{2};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void count_tc_launch() { g_tc_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

cudaError_t ensure_dynamic_smem(const void *kernel, int bytes)
{
    static std::mutex mu;
    static std::set<std::pair<const void *, int>> done;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    std::lock_guard<std::mutex> lock(mu);
    if (done.count({kernel, dev})) return cudaSuccess;
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (err == cudaSuccess) done.insert({kernel, dev});
    return err;
}

}  // namespace gnngp

extern "C" {

int gnngp_abi_version(void) { return 1; }

const char *gnngp_last_error(void) { return gnngp::g_error; }

int gnngp_device_info(int *host_sm_count, int *host_cc_major, int *host_cc_minor)
{
    int dev = 0;
    GNNGP_CUDA(cudaGetDevice(&dev));
    int sms = 0, major = 0, minor = 0;
    GNNGP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GNNGP_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    GNNGP_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (host_sm_count) *host_sm_count = sms;
    if (host_cc_major) *host_cc_major = major;
    if (host_cc_minor) *host_cc_minor = minor;
    return GNNGP_OK;
}

int gnngp_set_gemm_engine(int engine)
{
    if (engine < 0 || engine > 2) engine = 2;
    return gnngp::g_gemm_engine.exchange(engine);
}

int64_t gnngp_launch_count(void) { return gnngp::g_launches.load(); }

int64_t gnngp_tc_launch_count(void) { return gnngp::g_tc_launches.load(); }


// ---- peer exchange buffers (CUDA IPC), see include/gnngp_b200.h
int64_t gnngp_exchange_alloc(size_t bytes, void *handle_out)
{
    void *ptr = nullptr;
    if (!handle_out || bytes == 0 || cudaMalloc(&ptr, bytes) != cudaSuccess) {
        gnngp::set_error("exchange_alloc: cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    cudaIpcMemHandle_t handle;
    if (cudaIpcGetMemHandle(&handle, ptr) != cudaSuccess) {
        gnngp::set_error("exchange_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(ptr);
        return 0;
    }
    memcpy(handle_out, &handle, sizeof(handle));
    return (int64_t)reinterpret_cast<uintptr_t>(ptr);
}

int64_t gnngp_exchange_open(const void *handle_in)
{
    cudaIpcMemHandle_t handle;
    memcpy(&handle, handle_in, sizeof(handle));
    void *ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        gnngp::set_error("exchange_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    return (int64_t)reinterpret_cast<uintptr_t>(ptr);
}

int gnngp_exchange_close(int64_t peer_ptr)
{
    GNNGP_CUDA(cudaIpcCloseMemHandle(reinterpret_cast<void *>((uintptr_t)peer_ptr)));
    return GNNGP_OK;
}

int gnngp_exchange_free(int64_t ptr)
{
    GNNGP_CUDA(cudaFree(reinterpret_cast<void *>((uintptr_t)ptr)));
    return GNNGP_OK;
}

int gnngp_copy2d_f32(const float *src, int64_t lds, int64_t rows, int64_t cols, int64_t dst_ptr, int64_t ldd, gnngp_stream_t stream)
{
    GNNGP_REQUIRE(src && dst_ptr && rows >= 0 && cols >= 0 && lds >= cols && ldd >= cols, "copy2d: bad arguments");
    if (rows == 0 || cols == 0) return GNNGP_OK;
    GNNGP_CUDA(cudaMemcpy2DAsync(reinterpret_cast<void *>((uintptr_t)dst_ptr), (size_t)ldd * sizeof(float), src, (size_t)lds * sizeof(float),
                                 (size_t)cols * sizeof(float), (size_t)rows, cudaMemcpyDeviceToDevice, gnngp::as_stream(stream)));
    return GNNGP_OK;
}

}  // extern "C"
