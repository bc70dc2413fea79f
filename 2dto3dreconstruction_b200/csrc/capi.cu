// Library-wide state of the C ABI: status strings, CUDA error capture, launch counter.
#include "common.cuh"
#include <mutex>
#include <string>

namespace r3d {
std::atomic<uint64_t> g_launches{0};
static std::mutex g_err_mutex;
static std::string g_last_error;

void set_cuda_error(cudaError_t e, const char* where) {
    std::lock_guard<std::mutex> lock(g_err_mutex);
    g_last_error = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e) + " at " + where;
}
}  // namespace r3d

extern "C" int r3d_version(void) { return 100; }

extern "C" const char* r3d_status_string(int status) {
    switch (status) {
        case R3D_OK: return "ok";
        case R3D_ERR_INVALID: return "invalid argument";
        case R3D_ERR_WORKSPACE: return "workspace too small";
        case R3D_ERR_CUDA: return "CUDA runtime error";
        case R3D_ERR_NO_DEVICE: return "no sm_100 CUDA device";
        case R3D_ERR_NUMERIC: return "non-finite value";
        default: return "unknown status";
    }
}

extern "C" const char* r3d_last_cuda_error(void) {
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lock(r3d::g_err_mutex);
    copy = r3d::g_last_error;
    return copy.c_str();
}

extern "C" int r3d_check_device(void) {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return R3D_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { cudaGetLastError(); return R3D_ERR_NO_DEVICE; }
    return (prop.major == 10) ? R3D_OK : R3D_ERR_NO_DEVICE;
}

extern "C" uint64_t r3d_launch_count(void) { return r3d::g_launches.load(); }
