// ctx.cu -- context lifecycle and error text for libsba_b200.so.
#include <cstdarg>

#include "common.cuh"

namespace sba {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace sba

extern "C" {

int sba_version(void) { return SBA_B200_VERSION; }

const char* sba_last_error(void) { return sba::g_err; }

int sba_ctx_create(int device, void* stream, sba_ctx** out)
{
    SBA_CHECK_ARG(out != nullptr);
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        sba::set_error("no CUDA device available (%s); libsba_b200 has no CPU fallback",
                       e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return SBA_ERR_NO_DEVICE;
    }
    SBA_CHECK_ARG(device >= 0 && device < n);
    SBA_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SBA_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        sba::set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return SBA_ERR_UNSUPPORTED;
    }
    sba_ctx* c = new sba_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (stream) {
        c->stream = (cudaStream_t)stream;
        c->own_stream = false;
    } else {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete c;
            sba::set_error("cudaStreamCreate: %s", cudaGetErrorString(e));
            return SBA_ERR_CUDA;
        }
        c->own_stream = true;
    }
    c->pdl = getenv("SBA_PDL") != nullptr;
    c->pdl_small = getenv("SBA_NO_PDL_SMALL") == nullptr;   // measurement switch
    e = cudaMallocHost((void**)&c->pinned_i32, 64 * sizeof(int));
    if (e != cudaSuccess) {
        if (c->own_stream) cudaStreamDestroy(c->stream);
        delete c;
        sba::set_error("cudaMallocHost: %s", cudaGetErrorString(e));
        return SBA_ERR_NOMEM;
    }
    *out = c;
    return SBA_OK;
}

int sba_ctx_destroy(sba_ctx* c)
{
    if (!c) return SBA_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& b : c->scratch) b.release();
    c->cache.release_all();
    for (auto& kv : c->plans) {
        if (kv.second.lut) cudaFree(kv.second.lut);
        sba::free_tiled_plan(&kv.second.tiled);
    }
    for (auto& kv : c->band_tiled) sba::free_tiled_plan(&kv.second);
    for (auto& kv : c->crop_plans)
        if (kv.second.lut) cudaFree(kv.second.lut);
    for (auto& kv : c->band_plans)
        if (kv.second) cudaFree(kv.second);
    for (auto& kv : c->tc_spans) cudaFree(kv.second.second);
    if (c->pinned_i32) cudaFreeHost(c->pinned_i32);
    for (int k = 0; k < 3; k++) {
        if (c->prof_e0[k]) cudaEventDestroy(c->prof_e0[k]);
        if (c->prof_e1[k]) cudaEventDestroy(c->prof_e1[k]);
    }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (int k = 0; k < 2; k++)
        if (c->copy_ev[k]) cudaEventDestroy(c->copy_ev[k]);
    if (c->main_ev) cudaEventDestroy(c->main_ev);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return SBA_OK;
}

int sba_ctx_set_matcher_ctas(sba_ctx* c, int n_ctas)
{
    SBA_CHECK_ARG(c && n_ctas >= 0);
    c->matcher_ctas = n_ctas;
    return SBA_OK;
}

int sba_ctx_set_dependent_launch(sba_ctx* c, int enable)
{
    SBA_CHECK_ARG(c != nullptr);
    c->pdl = enable != 0;
    return SBA_OK;
}

int sba_ctx_set_stream(sba_ctx* c, void* stream)
{
    SBA_CHECK_ARG(c != nullptr);
    if ((cudaStream_t)stream == c->stream) return SBA_OK;
    SBA_CUDA(cudaStreamSynchronize(c->stream));
    if (c->own_stream) {
        cudaStreamDestroy(c->stream);
        c->own_stream = false;
    }
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        SBA_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    return SBA_OK;
}

void* sba_ctx_get_stream(sba_ctx* c) { return c ? (void*)c->stream : nullptr; }

int sba_ctx_synchronize(sba_ctx* c)
{
    SBA_CHECK_ARG(c != nullptr);
    SBA_CUDA(cudaStreamSynchronize(c->stream));
    return SBA_OK;
}

int64_t sba_ctx_launch_count(sba_ctx* c) { return c ? c->launches : 0; }

int sba_ctx_set_profiling(sba_ctx* c, int enable)
{
    SBA_CHECK_ARG(c != nullptr);
    SBA_CUDA(cudaSetDevice(c->device));
    if (enable && !c->prof_e0[0]) {
        for (int k = 0; k < 3; k++) {
            SBA_CUDA(cudaEventCreate(&c->prof_e0[k]));
            SBA_CUDA(cudaEventCreate(&c->prof_e1[k]));
        }
    }
    c->profiling = enable != 0;
    return SBA_OK;
}

int sba_ctx_kernel_ms(sba_ctx* c, int id, float* ms)
{
    SBA_CHECK_ARG(c && ms && id >= 0 && id < 3);
    if (!c->prof_valid[id]) {
        sba::set_error("kernel %d has not been launched with profiling enabled", id);
        return SBA_ERR_INVALID;
    }
    SBA_CUDA(cudaEventSynchronize(c->prof_e1[id]));
    SBA_CUDA(cudaEventElapsedTime(ms, c->prof_e0[id], c->prof_e1[id]));
    return SBA_OK;
}

}  // extern "C"
