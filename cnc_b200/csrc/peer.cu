// peer.cu -- the gradient exchange of the data-parallel training step over NVLink peer memory (sm_100a, one box).
//
// What is exchanged (cnc_b200/dp.py, ShardedTableAdam): every rank holds the full gradient of the four latent tables
// (161 MB fp32, written by the K2 scatter-add), needs the rank-average of the rows it owns (1/N of them) and hands back
// two bits per latent.  Over NCCL that is a reduce-scatter plus an all-gather: ring kernels that pass every byte through
// N - 1 hops and hold SMs while they wait.  On an NVSwitch box every GPU can load from every other GPU's HBM directly, so
// the same exchange is three small kernels on buffers that are mapped into all ranks (CUDA IPC):
//
//   cnc_peer_barrier      1 CTA, one thread per peer: release-store of the step's epoch into the peer's signal pad, then an
//                         acquire-spin on the own pad ("every rank's table gradient is complete")
//   cnc_peer_reduce       out[i] = scale * sum_k grad_k[lo + i]: 16-byte loads straight from the N gradient buffers (the own
//                         one and N - 1 peers), all N loads of a thread in flight together, summed in rank order (the
//                         result does not depend on which rank computes it); NVLink moves (N-1)/N * 161 MB per rank, once
//   cnc_peer_push         the owned words of the bit planes are stored into every peer's plane arena (2 bits per latent)
//
// Memory: cnc_peer_alloc = cudaMalloc (IPC handles need a plain allocation, not a pool / VMM block), cnc_peer_export /
// cnc_peer_import = cudaIpcGetMemHandle / cudaIpcOpenMemHandle.  The 64-byte handles travel through torch.distributed.
// A barrier that waits longer than `timeout_ms` traps (the step fails loudly instead of hanging the box).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace cnc {
namespace peer {

constexpr int MAX_WORLD = 8;
constexpr int PAD_SLOTS = 16;                    // signal pad = PAD_SLOTS x MAX_WORLD uint32 epochs

struct Pads { uint32_t *p[MAX_WORLD]; };
struct Srcs { const float *p[MAX_WORLD]; };
struct Dsts { uint32_t *p[MAX_WORLD]; };
struct Segs { int64_t off[8]; int64_t words[8]; int32_t n; };

__device__ __forceinline__ void st_release_sys(uint32_t *a, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *a) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__global__ void barrier_kernel(Pads pads, int rank, int world, int slot, uint32_t epoch, uint64_t timeout_ns) {
    const int k = threadIdx.x;
    if (k >= world) return;
    __threadfence_system();
    st_release_sys(pads.p[k] + slot * MAX_WORLD + rank, epoch);
    const uint32_t *mine = pads.p[rank] + slot * MAX_WORLD + k;
    const uint64_t t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
        __nanosleep(64);
        if (globaltimer_ns() - t0 > timeout_ns) {
            printf("cnc_peer_barrier: rank %d waited %llu ms for rank %d (slot %d, epoch %u): giving up\n", rank,
                   (unsigned long long)(timeout_ns / 1000000ull), k, slot, epoch);
            __trap();
        }
    }
}

// every rank contributes one number and learns the minimum over the ranks (the skip vote of the training step: the sample
// count of the rank's batch).  value words by epoch parity (a rank can be at most one vote ahead of another), epoch words
// released after them.
__global__ void min_kernel(Pads pads, int rank, int world, int slot, uint32_t epoch, uint32_t value, uint32_t *out,
                           uint64_t timeout_ns) {
    const int k = threadIdx.x;
    uint32_t v = 0xFFFFFFFFu;
    if (k < world) {
        const int vslot = slot + 1 + (int)(epoch & 1u);
        *reinterpret_cast<volatile uint32_t *>(pads.p[k] + vslot * MAX_WORLD + rank) = value;
        __threadfence_system();
        st_release_sys(pads.p[k] + slot * MAX_WORLD + rank, epoch);
        const uint32_t *mine = pads.p[rank] + slot * MAX_WORLD + k;
        const uint64_t t0 = globaltimer_ns();
        while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
            __nanosleep(64);
            if (globaltimer_ns() - t0 > timeout_ns) {
                printf("cnc_peer_min: rank %d waited %llu ms for rank %d (epoch %u): giving up\n", rank,
                       (unsigned long long)(timeout_ns / 1000000ull), k, epoch);
                __trap();
            }
        }
        v = *reinterpret_cast<volatile uint32_t *>(pads.p[rank] + vslot * MAX_WORLD + k);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const uint32_t o = __shfl_xor_sync(0xffffffffu, v, d);
        v = o < v ? o : v;
    }
    if (k == 0) *out = v;
}

template <int W>
__global__ void __launch_bounds__(512) reduce_kernel(Srcs src, int64_t lo, int64_t n4, float scale, float4 *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v[W];
#pragma unroll
        for (int k = 0; k < W; k++) v[k] = __ldcv(reinterpret_cast<const float4 *>(src.p[k] + lo) + i);
        float4 s = v[0];
#pragma unroll
        for (int k = 1; k < W; k++) {
            s.x = __fadd_rn(s.x, v[k].x); s.y = __fadd_rn(s.y, v[k].y); s.z = __fadd_rn(s.z, v[k].z); s.w = __fadd_rn(s.w, v[k].w);
        }
        out[i] = make_float4(__fmul_rn(s.x, scale), __fmul_rn(s.y, scale), __fmul_rn(s.z, scale), __fmul_rn(s.w, scale));
    }
}

__global__ void __launch_bounds__(256) push_kernel(Dsts dst, int rank, int world, Segs segs) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int s = 0; s < segs.n; s++) {
        const uint32_t *mine = dst.p[rank] + segs.off[s];
        for (int64_t i = t0; i < segs.words[s]; i += stride) {
            const uint32_t v = mine[i];
            for (int k = 0; k < world; k++)
                if (k != rank) dst.p[k][segs.off[s] + i] = v;
        }
    }
}

}  // namespace peer
}  // namespace cnc

using namespace cnc;

extern "C" {

int cnc_peer_alloc(uint64_t bytes, void **out) {
    if (!out || bytes == 0) { set_error("peer_alloc: bad argument"); return CNC_EINVAL; }
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess || cudaMemset(p, 0, bytes) != cudaSuccess) {
        set_error("peer_alloc: cudaMalloc(%llu) failed: %s", (unsigned long long)bytes, cudaGetErrorString(cudaGetLastError()));
        return CNC_ECUDA;
    }
    *out = p;
    return CNC_OK;
}

int cnc_peer_free(void *p) {
    if (p && cudaFree(p) != cudaSuccess) { set_error("peer_free: %s", cudaGetErrorString(cudaGetLastError())); return CNC_ECUDA; }
    return CNC_OK;
}

int cnc_peer_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int cnc_peer_export(void *p, uint8_t *handle) {
    if (!p || !handle) { set_error("peer_export: null pointer"); return CNC_EINVAL; }
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
        set_error("peer_export: cudaIpcGetMemHandle: %s", cudaGetErrorString(cudaGetLastError()));
        return CNC_ECUDA;
    }
    memcpy(handle, &h, sizeof(h));
    return CNC_OK;
}

int cnc_peer_import(const uint8_t *handle, void **out) {
    if (!handle || !out) { set_error("peer_import: null pointer"); return CNC_EINVAL; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        set_error("peer_import: cudaIpcOpenMemHandle: %s (peer access between the two GPUs, same container?)",
                  cudaGetErrorString(cudaGetLastError()));
        return CNC_ECUDA;
    }
    *out = p;
    return CNC_OK;
}

int cnc_peer_unmap(void *p) {
    if (p && cudaIpcCloseMemHandle(p) != cudaSuccess) { set_error("peer_unmap: %s", cudaGetErrorString(cudaGetLastError())); return CNC_ECUDA; }
    return CNC_OK;
}

int cnc_peer_pad_bytes(void) { return peer::PAD_SLOTS * peer::MAX_WORLD * (int)sizeof(uint32_t); }

int cnc_peer_barrier(void *const *pads, int32_t rank, int32_t world, int32_t slot, uint32_t epoch, uint32_t timeout_ms, cnc_stream_t stream) {
    if (!pads || world < 1 || world > peer::MAX_WORLD || rank < 0 || rank >= world || slot < 0 || slot >= peer::PAD_SLOTS) {
        set_error("peer_barrier: bad argument (world <= %d, slot < %d)", peer::MAX_WORLD, peer::PAD_SLOTS);
        return CNC_EINVAL;
    }
    peer::Pads a{};
    for (int k = 0; k < world; k++) {
        if (!pads[k]) { set_error("peer_barrier: null pad"); return CNC_EINVAL; }
        a.p[k] = static_cast<uint32_t *>(pads[k]);
    }
    peer::barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(a, rank, world, slot, epoch, (uint64_t)timeout_ms * 1000000ull);
    return check_launch("peer_barrier");
}

int cnc_peer_min(void *const *pads, int32_t rank, int32_t world, int32_t slot, uint32_t epoch, uint32_t value, uint32_t *out,
                 uint32_t timeout_ms, cnc_stream_t stream) {
    if (!pads || !out || world < 1 || world > peer::MAX_WORLD || rank < 0 || rank >= world || slot < 0 || slot + 2 >= peer::PAD_SLOTS) {
        set_error("peer_min: bad argument (uses slots slot .. slot + 2)");
        return CNC_EINVAL;
    }
    peer::Pads a{};
    for (int k = 0; k < world; k++) {
        if (!pads[k]) { set_error("peer_min: null pad"); return CNC_EINVAL; }
        a.p[k] = static_cast<uint32_t *>(pads[k]);
    }
    peer::min_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(a, rank, world, slot, epoch, value, out, (uint64_t)timeout_ms * 1000000ull);
    return check_launch("peer_min");
}

int cnc_peer_reduce(const void *const *srcs, int32_t world, int64_t lo, int64_t count, float scale, float *out, int32_t blocks,
                    cnc_stream_t stream) {
    if (count == 0) return CNC_OK;
    if (!srcs || !out || world < 1 || world > peer::MAX_WORLD || lo < 0 || count < 0) { set_error("peer_reduce: bad argument"); return CNC_EINVAL; }
    if ((lo & 3) || (count & 3) || (reinterpret_cast<uintptr_t>(out) & 15u)) {
        set_error("peer_reduce: lo, count must be multiples of 4 floats and out 16-byte aligned");
        return CNC_EINVAL;
    }
    peer::Srcs a{};
    for (int k = 0; k < world; k++) {
        if (!srcs[k] || (reinterpret_cast<uintptr_t>(srcs[k]) & 15u)) { set_error("peer_reduce: null / misaligned source"); return CNC_EINVAL; }
        a.p[k] = static_cast<const float *>(srcs[k]);
    }
    const int64_t n4 = count / 4;
    int64_t g = (n4 + 511) / 512;
    const int64_t cap = blocks > 0 ? blocks : 148;
    if (g > cap) g = cap;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float4 *o = reinterpret_cast<float4 *>(out);
    switch (world) {
        case 1: peer::reduce_kernel<1><<<(unsigned)g, 512, 0, s>>>(a, lo, n4, scale, o); break;
        case 2: peer::reduce_kernel<2><<<(unsigned)g, 512, 0, s>>>(a, lo, n4, scale, o); break;
        case 3: peer::reduce_kernel<3><<<(unsigned)g, 512, 0, s>>>(a, lo, n4, scale, o); break;
        case 4: peer::reduce_kernel<4><<<(unsigned)g, 512, 0, s>>>(a, lo, n4, scale, o); break;
        case 5: peer::reduce_kernel<5><<<(unsigned)g, 512, 0, s>>>(a, lo, n4, scale, o); break;
        case 6: peer::reduce_kernel<6><<<(unsigned)g, 512, 0, s>>>(a, lo, n4, scale, o); break;
        case 7: peer::reduce_kernel<7><<<(unsigned)g, 512, 0, s>>>(a, lo, n4, scale, o); break;
        default: peer::reduce_kernel<8><<<(unsigned)g, 512, 0, s>>>(a, lo, n4, scale, o); break;
    }
    return check_launch("peer_reduce");
}

int cnc_peer_push(void *const *arenas, int32_t rank, int32_t world, const int64_t *seg_off_words, const int64_t *seg_words, int32_t n_seg,
                  cnc_stream_t stream) {
    if (!arenas || !seg_off_words || !seg_words || world < 1 || world > peer::MAX_WORLD || rank < 0 || rank >= world || n_seg < 0 || n_seg > 8) {
        set_error("peer_push: bad argument (at most 8 segments)");
        return CNC_EINVAL;
    }
    if (n_seg == 0 || world == 1) return CNC_OK;
    peer::Dsts d{};
    for (int k = 0; k < world; k++) {
        if (!arenas[k]) { set_error("peer_push: null arena"); return CNC_EINVAL; }
        d.p[k] = static_cast<uint32_t *>(arenas[k]);
    }
    peer::Segs sg{};
    sg.n = n_seg;
    int64_t most = 0;
    for (int s = 0; s < n_seg; s++) {
        if (seg_off_words[s] < 0 || seg_words[s] < 0) { set_error("peer_push: negative segment"); return CNC_EINVAL; }
        sg.off[s] = seg_off_words[s];
        sg.words[s] = seg_words[s];
        if (seg_words[s] > most) most = seg_words[s];
    }
    int64_t g = (most + 255) / 256;
    if (g > 148) g = 148;
    if (g < 1) g = 1;
    peer::push_kernel<<<(unsigned)g, 256, 0, static_cast<cudaStream_t>(stream)>>>(d, rank, world, sg);
    return check_launch("peer_push");
}

}  // extern "C"
