// C-ABI of the GDKVM memory op (see include/gdkvm_gdr.h).  Validation + dispatch only; no torch
// types, no allocation of user-visible memory, no host synchronisation.
#include <atomic>
#include <mutex>

#include "gdr_common.cuh"

namespace gdkvm {

static std::atomic<uint64_t> g_launches{0};
static thread_local int tl_last_cuda_error = 0;

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

namespace {

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Shape / dtype / alignment checks shared by plan() and fwd().  No GPU needed.
int validate(const GdkvmGdrParams* p) {
    if (p == nullptr) return GDKVM_ERR_NULL;
    if (p->struct_size != sizeof(GdkvmGdrParams)) return GDKVM_ERR_ABI;
    if (p->io_dtype != GDKVM_F32 && p->io_dtype != GDKVM_BF16) return GDKVM_ERR_DTYPE;
    if (p->gate_dtype != GDKVM_F32 && p->gate_dtype != GDKVM_BF16) return GDKVM_ERR_DTYPE;
    if (p->B <= 0 || p->H <= 0 || p->T < 0 || p->V <= 0) return GDKVM_ERR_SHAPE;
    if (p->K != 32 && p->K != 64 && p->K != 128) return GDKVM_ERR_SHAPE;
    if ((int64_t)p->B * p->H > 0x7fffffff) return GDKVM_ERR_SHAPE;
    if (p->frame_tokens < 0) return GDKVM_ERR_SHAPE;
    if (p->frame_tokens > 0 && p->T % p->frame_tokens != 0) return GDKVM_ERR_SHAPE;
    if (p->T > 0 && (!p->q || !p->k || !p->v || !p->g || !p->beta || !p->o)) return GDKVM_ERR_NULL;
    if ((p->flags & GDKVM_FLAG_FORCE_RECURRENT) && (p->flags & GDKVM_FLAG_FORCE_CHUNKED)) return GDKVM_ERR_UNSUPPORTED;
    // q/k rows are fetched with 16-byte vector loads (both kernels); TMA needs the same of v/o.
    const int64_t es = p->io_dtype == GDKVM_BF16 ? 2 : 4;
    if (p->T > 0) {
        if (!aligned16(p->q) || !aligned16(p->k)) return GDKVM_ERR_ALIGN;
        for (int i = 0; i < 3; ++i)
            if ((p->q_stride[i] * es) % 16 != 0 || (p->k_stride[i] * es) % 16 != 0) return GDKVM_ERR_ALIGN;
    }
    if ((p->initial_state && (reinterpret_cast<uintptr_t>(p->initial_state) & 3u)) ||
        (p->final_state && (reinterpret_cast<uintptr_t>(p->final_state) & 3u)))
        return GDKVM_ERR_ALIGN;
    return GDKVM_OK;
}

int pick(const GdkvmGdrParams* p) {
    if (p->flags & GDKVM_FLAG_FORCE_RECURRENT) return 0;
    const bool ok = chunked_supports(*p);
    if (p->flags & GDKVM_FLAG_FORCE_CHUNKED) return ok ? 1 : GDKVM_ERR_UNSUPPORTED;
    return ok ? 1 : 0;
}

// compute capability of the current device, cached per device ordinal
int device_is_sm100(bool* ok) {
    static std::mutex mu;
    static int cached[64];  // 0 unknown, 1 yes, 2 no
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64) { *ok = false; return 0; }
    std::lock_guard<std::mutex> lk(mu);
    if (cached[dev] == 0) {
        int major = 0;
        e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        if (e != cudaSuccess) return (int)e;
        cached[dev] = major == 10 ? 1 : 2;
    }
    *ok = cached[dev] == 1;
    return 0;
}

}  // namespace
}  // namespace gdkvm

extern "C" {

int gdkvm_abi_version(void) { return GDKVM_ABI_VERSION; }

const char* gdkvm_strerror(int status) {
    switch (status) {
        case GDKVM_OK: return "ok";
        case GDKVM_ERR_NULL: return "a required pointer is NULL";
        case GDKVM_ERR_ABI: return "GdkvmGdrParams.struct_size does not match this library (ABI mismatch)";
        case GDKVM_ERR_SHAPE: return "unsupported shape (need B,H,V>=1, T>=0, K in {32,64,128}, T % frame_tokens == 0)";
        case GDKVM_ERR_DTYPE: return "unknown dtype enum";
        case GDKVM_ERR_ALIGN: return "q/k pointers and strides must be 16-byte aligned; states 4-byte aligned";
        case GDKVM_ERR_ARCH: return "device is not sm_100 (B200); this library has no fallback path";
        case GDKVM_ERR_CUDA: return "CUDA call failed (see gdkvm_last_cuda_error)";
        case GDKVM_ERR_UNSUPPORTED: return "the forced kernel path does not support this problem";
        default: return "unknown gdkvm status";
    }
}

int gdkvm_last_cuda_error(void) { return gdkvm::tl_last_cuda_error; }

uint64_t gdkvm_launch_count(void) { return gdkvm::g_launches.load(std::memory_order_relaxed); }

int gdkvm_gdr_plan(const GdkvmGdrParams* params) {
    const int rc = gdkvm::validate(params);
    if (rc != GDKVM_OK) return rc;
    return gdkvm::pick(params);
}

const char* gdkvm_gdr_plan_reason(const GdkvmGdrParams* params) {
    const int rc = gdkvm::validate(params);
    if (rc != GDKVM_OK) return gdkvm_strerror(rc);
    if (params->flags & GDKVM_FLAG_FORCE_RECURRENT) return "GDKVM_FLAG_FORCE_RECURRENT is set";
    return gdkvm::chunked_unsupported_reason(*params);
}

int gdkvm_gdr_plan_segments(const GdkvmGdrParams* params, int sm_count) {
    const int rc = gdkvm::validate(params);
    if (rc != GDKVM_OK) return rc;
    if (params->T == 0 || gdkvm::pick(params) != 1) return 1;
    return gdkvm::chunked_segments(*params, sm_count);
}

int gdkvm_gdr_plan_units(const GdkvmGdrParams* params, int sm_count, int32_t out[4]) {
    const int rc = gdkvm::validate(params);
    if (rc != GDKVM_OK) return rc;
    if (out == nullptr) return GDKVM_ERR_NULL;
    if (params->T == 0 || gdkvm::pick(params) != 1) {          // not the chunk kernel: one unit per (clip, head) chain
        out[0] = params->B * params->H; out[1] = params->B; out[2] = 0; out[3] = 1;
        return 0;
    }
    int o[4];
    const int mixed = gdkvm::chunked_plan_units(*params, sm_count, o);
    for (int i = 0; i < 4; ++i) out[i] = o[i];
    return mixed;
}

int gdkvm_l2norm_fwd(const void* x, void* y, int64_t rows, int32_t D, int64_t x_row_stride, int64_t y_row_stride,
                     int32_t dtype, float eps, void* cuda_stream) {
    if (rows < 0 || (D != 32 && D != 64 && D != 128 && D != 256)) return GDKVM_ERR_SHAPE;
    if (dtype != GDKVM_F32 && dtype != GDKVM_BF16) return GDKVM_ERR_DTYPE;
    if (rows == 0) return GDKVM_OK;
    if (x == nullptr || y == nullptr) return GDKVM_ERR_NULL;
    const int64_t es = dtype == GDKVM_BF16 ? 2 : 4;
    if ((reinterpret_cast<uintptr_t>(x) & 15u) || (reinterpret_cast<uintptr_t>(y) & 15u) || (x_row_stride * es) % 16 != 0 ||
        (y_row_stride * es) % 16 != 0 || x_row_stride < D || y_row_stride < D)
        return GDKVM_ERR_ALIGN;
    bool sm100 = false;
    int ce = gdkvm::device_is_sm100(&sm100);
    if (ce != 0) { gdkvm::tl_last_cuda_error = ce; return GDKVM_ERR_CUDA; }
    if (!sm100) return GDKVM_ERR_ARCH;
    ce = gdkvm::launch_l2norm(x, y, rows, D, x_row_stride, y_row_stride, dtype, eps, reinterpret_cast<cudaStream_t>(cuda_stream));
    if (ce != 0) { gdkvm::tl_last_cuda_error = ce; return GDKVM_ERR_CUDA; }
    return GDKVM_OK;
}

int gdkvm_gdr_fwd_varlen(const GdkvmGdrParams* params, const void* cu_seqlens, int32_t cu_seqlens_bytes, int32_t n_seqs,
                         void* cuda_stream) {
    int rc = gdkvm::validate(params);
    if (rc != GDKVM_OK) return rc;
    if (params->B != 1 || n_seqs < 1 || (cu_seqlens_bytes != 4 && cu_seqlens_bytes != 8)) return GDKVM_ERR_SHAPE;
    if (cu_seqlens == nullptr) return GDKVM_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(cu_seqlens) % (uintptr_t)cu_seqlens_bytes != 0) return GDKVM_ERR_ALIGN;
    if ((int64_t)n_seqs * params->H > 0x3fffffff) return GDKVM_ERR_SHAPE;
    const int path = gdkvm::pick(params);
    if (path < 0) return path;
    bool sm100 = false;
    int ce = gdkvm::device_is_sm100(&sm100);
    if (ce != 0) { gdkvm::tl_last_cuda_error = ce; return GDKVM_ERR_CUDA; }
    if (!sm100) return GDKVM_ERR_ARCH;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    if (params->T == 0) {        // every clip is empty: the initial states pass through (zeros without one)
        if (params->final_state != nullptr) {
            const size_t bytes = (size_t)n_seqs * params->H * params->K * params->V * sizeof(float);
            const cudaError_t e = params->initial_state != nullptr
                ? cudaMemcpyAsync(params->final_state, params->initial_state, bytes, cudaMemcpyDeviceToDevice, stream)
                : cudaMemsetAsync(params->final_state, 0, bytes, stream);
            if (e != cudaSuccess) { gdkvm::tl_last_cuda_error = (int)e; return GDKVM_ERR_CUDA; }
        }
        return GDKVM_OK;
    }
    ce = path == 1 ? gdkvm::launch_chunked_varlen(*params, cu_seqlens, cu_seqlens_bytes, n_seqs, stream)
                   : gdkvm::launch_recurrent_varlen(*params, cu_seqlens, cu_seqlens_bytes, n_seqs, stream);
    if (ce != 0) { gdkvm::tl_last_cuda_error = ce; return GDKVM_ERR_CUDA; }
    return GDKVM_OK;
}

int64_t gdkvm_gdr_chunk_states_bytes(int32_t B, int32_t T, int32_t H, int32_t K, int32_t V) {
    if (B <= 0 || T < 0 || H <= 0 || K <= 0 || V <= 0) return 0;
    return (int64_t)B * H * ((T + 63) / 64) * V * K * 2;
}

int gdkvm_gdr_fwd_train(const GdkvmGdrParams* params, void* chunk_states, void* cuda_stream) {
    int rc = gdkvm::validate(params);
    if (rc != GDKVM_OK) return rc;
    if (params->flags & GDKVM_FLAG_FORCE_RECURRENT) return GDKVM_ERR_UNSUPPORTED;
    if (!gdkvm::chunked_supports(*params)) return params->T == 0 ? GDKVM_ERR_SHAPE : GDKVM_ERR_UNSUPPORTED;
    if (chunk_states == nullptr) return GDKVM_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(chunk_states) & 31u) return GDKVM_ERR_ALIGN;          // written with 256-bit stores
    bool sm100 = false;
    int ce = gdkvm::device_is_sm100(&sm100);
    if (ce != 0) { gdkvm::tl_last_cuda_error = ce; return GDKVM_ERR_CUDA; }
    if (!sm100) return GDKVM_ERR_ARCH;
    ce = gdkvm::launch_chunked(*params, reinterpret_cast<cudaStream_t>(cuda_stream), chunk_states);
    if (ce != 0) { gdkvm::tl_last_cuda_error = ce; return GDKVM_ERR_CUDA; }
    return GDKVM_OK;
}

int64_t gdkvm_gdr_chunk_states_bytes_varlen(int32_t T, int32_t n_seqs, int32_t H, int32_t K, int32_t V) {
    if (T < 0 || n_seqs <= 0 || H <= 0 || K <= 0 || V <= 0) return 0;
    return ((int64_t)T / 64 + n_seqs + 1) * H * V * K * 2;
}

int gdkvm_gdr_fwd_train_varlen(const GdkvmGdrParams* params, const void* cu_seqlens, int32_t cu_seqlens_bytes, int32_t n_seqs,
                               void* chunk_states, void* cuda_stream) {
    int rc = gdkvm::validate(params);
    if (rc != GDKVM_OK) return rc;
    if (params->B != 1 || n_seqs < 1 || (cu_seqlens_bytes != 4 && cu_seqlens_bytes != 8)) return GDKVM_ERR_SHAPE;
    if (cu_seqlens == nullptr || chunk_states == nullptr) return GDKVM_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(cu_seqlens) % (uintptr_t)cu_seqlens_bytes != 0 || (reinterpret_cast<uintptr_t>(chunk_states) & 31u)) return GDKVM_ERR_ALIGN;
    if ((int64_t)n_seqs * params->H > 0x3fffffff) return GDKVM_ERR_SHAPE;
    if (params->flags & GDKVM_FLAG_FORCE_RECURRENT) return GDKVM_ERR_UNSUPPORTED;
    if (params->T == 0) return GDKVM_ERR_SHAPE;
    if (!gdkvm::chunked_supports(*params)) return GDKVM_ERR_UNSUPPORTED;
    bool sm100 = false;
    int ce = gdkvm::device_is_sm100(&sm100);
    if (ce != 0) { gdkvm::tl_last_cuda_error = ce; return GDKVM_ERR_CUDA; }
    if (!sm100) return GDKVM_ERR_ARCH;
    ce = gdkvm::launch_chunked_varlen(*params, cu_seqlens, cu_seqlens_bytes, n_seqs, reinterpret_cast<cudaStream_t>(cuda_stream), chunk_states);
    if (ce != 0) { gdkvm::tl_last_cuda_error = ce; return GDKVM_ERR_CUDA; }
    return GDKVM_OK;
}

int gdkvm_gdr_bwd(const GdkvmGdrBwdParams* p, void* cuda_stream) {
    if (p == nullptr) return GDKVM_ERR_NULL;
    if (p->struct_size != sizeof(GdkvmGdrBwdParams)) return GDKVM_ERR_ABI;
    if (p->io_dtype != GDKVM_F32 && p->io_dtype != GDKVM_BF16) return GDKVM_ERR_DTYPE;
    if (p->gate_dtype != GDKVM_F32 && p->gate_dtype != GDKVM_BF16) return GDKVM_ERR_DTYPE;
    if (p->B <= 0 || p->H <= 0 || p->T <= 0 || p->V <= 0 || p->K <= 0) return GDKVM_ERR_SHAPE;
    if (!p->q || !p->k || !p->v || !p->g || !p->beta || !p->d_o || !p->chunk_states || !p->dq || !p->dk || !p->dv || !p->dg || !p->dbeta)
        return GDKVM_ERR_NULL;
    if (gdkvm::bwd_unsupported_reason(*p)[0] != '\0') return GDKVM_ERR_UNSUPPORTED;
    if (p->cu_seqlens != nullptr && (p->B != 1 || p->n_seqs < 1 || (p->cu_seqlens_bytes != 4 && p->cu_seqlens_bytes != 8))) return GDKVM_ERR_SHAPE;
    bool sm100 = false;
    int ce = gdkvm::device_is_sm100(&sm100);
    if (ce != 0) { gdkvm::tl_last_cuda_error = ce; return GDKVM_ERR_CUDA; }
    if (!sm100) return GDKVM_ERR_ARCH;
    ce = gdkvm::launch_bwd(*p, reinterpret_cast<cudaStream_t>(cuda_stream));
    if (ce != 0) { gdkvm::tl_last_cuda_error = ce; return GDKVM_ERR_CUDA; }
    return GDKVM_OK;
}

int gdkvm_qkvgb_project_fwd(const GdkvmProjParams* p, void* cuda_stream) {
    if (p == nullptr) return GDKVM_ERR_NULL;
    if (p->struct_size != sizeof(GdkvmProjParams)) return GDKVM_ERR_ABI;
    if (p->R < 0 || p->D <= 0 || p->H <= 0 || p->K <= 0 || p->V <= 0) return GDKVM_ERR_SHAPE;
    if (p->R == 0) return GDKVM_OK;
    if (!p->x || !p->w || !p->q || !p->k || !p->v || !p->g || !p->beta) return GDKVM_ERR_NULL;
    if (gdkvm::proj_unsupported_reason(*p)[0] != '\0') return GDKVM_ERR_UNSUPPORTED;
    bool sm100 = false;
    int ce = gdkvm::device_is_sm100(&sm100);
    if (ce != 0) { gdkvm::tl_last_cuda_error = ce; return GDKVM_ERR_CUDA; }
    if (!sm100) return GDKVM_ERR_ARCH;
    ce = gdkvm::launch_proj(*p, reinterpret_cast<cudaStream_t>(cuda_stream));
    if (ce != 0) { gdkvm::tl_last_cuda_error = ce; return GDKVM_ERR_CUDA; }
    return GDKVM_OK;
}

int gdkvm_gdr_fwd(const GdkvmGdrParams* params, void* cuda_stream) {
    int rc = gdkvm::validate(params);
    if (rc != GDKVM_OK) return rc;
    const int path = gdkvm::pick(params);
    if (path < 0) return path;
    bool sm100 = false;
    int ce = gdkvm::device_is_sm100(&sm100);
    if (ce != 0) { gdkvm::tl_last_cuda_error = ce; return GDKVM_ERR_CUDA; }
    if (!sm100) return GDKVM_ERR_ARCH;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    ce = path == 1 ? gdkvm::launch_chunked(*params, stream) : gdkvm::launch_recurrent(*params, stream);
    if (ce != 0) { gdkvm::tl_last_cuda_error = ce; return GDKVM_ERR_CUDA; }
    return GDKVM_OK;
}

}  // extern "C"
