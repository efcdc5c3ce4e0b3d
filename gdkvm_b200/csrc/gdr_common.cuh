// Shared declarations of the GDKVM memory-op kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gdkvm_gdr.h"

namespace gdkvm {

// Launchers (defined in gdr_recurrent.cu / gdr_chunked_sm100.cu). Return a cudaError_t as int.
int launch_recurrent(const GdkvmGdrParams& p, cudaStream_t stream);
// chunk_states != nullptr: training forward, also writes the bf16 chunk-start states [B*H][ceil(T/64)][V][K] (flat 64-token chunks)
int launch_chunked(const GdkvmGdrParams& p, cudaStream_t stream, void* chunk_states = nullptr);
// Packed variable-length sequences (q,k,v,o [1, T, H, *]; device-resident offsets cu[0..nseq], cu_bytes = 4 | 8; states [nseq, H, K, V])
int launch_recurrent_varlen(const GdkvmGdrParams& p, const void* cu, int cu_bytes, int nseq, cudaStream_t stream);
// chunk_states != nullptr: training forward, [T / 64 + nseq + 1 slots][H][V][K] bf16, slot of chunk c of clip n = cu[n] / 64 + n + c
int launch_chunked_varlen(const GdkvmGdrParams& p, const void* cu, int cu_bytes, int nseq, cudaStream_t stream, void* chunk_states = nullptr);
int launch_l2norm(const void* x, void* y, int64_t rows, int D, int64_t xs, int64_t ys, int dtype, float eps, cudaStream_t stream);
// Host-side eligibility test of the chunked tcgen05 kernel (no GPU needed).
bool chunked_supports(const GdkvmGdrParams& p);
// ... and the reason when it is not eligible ("" when it is)
const char* chunked_unsupported_reason(const GdkvmGdrParams& p);
// Time segments per chain the chunked kernel would use on a device with `sms` SMs (host-side schedule simulation).
int chunked_segments(const GdkvmGdrParams& p, int sms);

// Shared host helpers (gdr_chunked_sm100.cu): the schedule simulation behind the time segments, and a stream-ordered scratch
// allocation from the library's private per-device pool (*sms receives the device's SM count).  Return cudaError_t as int.
int plan_time_segments(int chains, int chunks, int sms);
int chunked_plan_units(const GdkvmGdrParams& p, int sms, int out[4]);
int library_scratch_alloc(void** ws, size_t bytes, cudaStream_t stream, int* sms, bool* mempools);

// Backward pass (gdr_bwd_sm100.cu)
int launch_bwd(const GdkvmGdrBwdParams& p, cudaStream_t stream);
const char* bwd_unsupported_reason(const GdkvmGdrBwdParams& p);

// Fused projection prologue (gdr_proj_sm100.cu)
int launch_proj(const GdkvmProjParams& p, cudaStream_t stream);
const char* proj_unsupported_reason(const GdkvmProjParams& p);

void count_launch();

__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }
__device__ __forceinline__ void from_f32(float& d, float x) { d = x; }
__device__ __forceinline__ void from_f32(__nv_bfloat16& d, float x) { d = __float2bfloat16_rn(x); }

// gate / beta scalar loads with a runtime dtype (they are 2 of ~650 values per token-head)
__device__ __forceinline__ float load_gate(const void* base, int64_t idx, int dtype) {
    return dtype == GDKVM_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx])
                               : reinterpret_cast<const float*>(base)[idx];
}

}  // namespace gdkvm
