// field_ops.cu -- small per-sample ops of the radiance field that are not the grid gather:
// spherical-harmonics direction encoding (tcnn replacement) and the sinusoidal position embedding.
//
// Reference behaviour restated: examples/radiance_fields/ngp.py:412-425,540-541 (tcnn
// SphericalHarmonics degree 4 on (dir+1)/2, fp16 output -- SURVEY Appendix C) and
// ngp.py:569-617 (Embedder: [x, sin(2^k x), cos(2^k x)], k=0..9).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace cnc {

__device__ __forceinline__ void sh16_eval(float x, float y, float z, float (&o)[16]) {
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    o[0] = 0.28209479177387814f;
    o[1] = -0.48860251190291987f * y;
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    o[10] = 2.8906114426405538f * xy * z;
    o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    o[14] = 1.4453057213202769f * z * (x2 - y2);
    o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

// d01 [n,3] in [0,1] -> out [n,16]; fp16_round emulates tcnn's half output
__global__ void __launch_bounds__(256)
sh16_kernel(const float *__restrict__ d01, float *__restrict__ out, uint64_t n, int fp16_round) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = __ldg(d01 + i * 3) * 2.f - 1.f, y = __ldg(d01 + i * 3 + 1) * 2.f - 1.f,
                z = __ldg(d01 + i * 3 + 2) * 2.f - 1.f;
    float o[16];
    sh16_eval(x, y, z, o);
    if (fp16_round) {
#pragma unroll
        for (int k = 0; k < 16; k++) o[k] = __half2float(__float2half_rn(o[k]));
    }
    float4 *dst = reinterpret_cast<float4 *>(out + i * 16);
#pragma unroll
    for (int k = 0; k < 4; k++) dst[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
}

// x [n,3] -> out [n, 3 + 6*n_freq]: x, then for k: sin(2^k x) (3), cos(2^k x) (3)
__global__ void __launch_bounds__(256)
freq_embed_kernel(const float *__restrict__ x, float *__restrict__ out, uint64_t n, int n_freq) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int W = 3 + 6 * n_freq;
    float v[3] = {__ldg(x + i * 3), __ldg(x + i * 3 + 1), __ldg(x + i * 3 + 2)};
    float *o = out + i * W;
    o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
    float f = 1.f;
    for (int k = 0; k < n_freq; k++) {
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const float a = __fmul_rn(v[d], f);
            o[3 + 6 * k + d] = sinf(a);
            o[3 + 6 * k + 3 + d] = cosf(a);
        }
        f *= 2.f;
    }
}

}  // namespace cnc

using namespace cnc;

extern "C" {

int cnc_sh16(const float *d01, float *out, uint64_t n, int fp16_round, cnc_stream_t stream) {
    if (n == 0) return CNC_OK;
    if (!d01 || !out) { set_error("sh16: null pointer"); return CNC_EINVAL; }
    if (reinterpret_cast<uintptr_t>(out) & 15u) { set_error("sh16: out must be 16-byte aligned"); return CNC_EINVAL; }
    sh16_kernel<<<div_up(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(d01, out, n, fp16_round);
    return check_launch("sh16");
}

int cnc_freq_embed(const float *x, float *out, uint64_t n, int n_freq, cnc_stream_t stream) {
    if (n == 0) return CNC_OK;
    if (!x || !out || n_freq < 0 || n_freq > 16) { set_error("freq_embed: bad argument"); return CNC_EINVAL; }
    freq_embed_kernel<<<div_up(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, n, n_freq);
    return check_launch("freq_embed");
}

}  // extern "C"
