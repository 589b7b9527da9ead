// coder.cu -- Bernoulli CDF quantiser and a torchac-compatible 32-bit binary range coder.
//
// Reference behaviour restated: examples/utils_bpp_acc.py:77-110 builds the float CDF
// [0, 1-p, 1] on the CPU and hands it to torchac==0.9.3 (third party, not in the reference
// tree; algorithm restated in SURVEY Appendix B and oracle/cnc_oracle.c).  The bytes produced
// here are identical to that coder's for equal (c1, symbol) inputs.
//
// Range coding is sequential inside a stream (low/high carry from symbol to symbol).  The
// B200 mapping is therefore one warp per independent stream (the product emits 33 of them):
//   * all 32 lanes stage the stream's (c1,sym) pairs through shared memory with coalesced
//     loads, one 1024-symbol tile ahead of the coder;
//   * lane 0 carries (low, high, pending) and runs the dependent chain -- one 32x16-bit
//     multiply, a select, and a renormalisation done in O(1) with clz instead of the
//     reference's bit-at-a-time loop;
//   * output bits are packed MSB-first in a 64-bit register and leave as 32-bit stores.
// The decoder mirrors it and replaces torchac's 64-bit division by the equivalent comparison
// c1*span <= ((value-low+1)<<16)-1.
#include <cuda_runtime.h>

#include "common.cuh"

namespace cnc {

__global__ void cdf_from_p_kernel(const float *__restrict__ p, uint16_t *__restrict__ c1, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float pu = __fsub_rn(1.0f, __ldg(p + i));          // utils_bpp_acc.py:81
        const float v = rintf(__fmul_rn(pu, 65534.0f));          // torchac: mul(2^16 - 2).round()
        c1[i] = (uint16_t)((int32_t)v + 1);                      // int16 wrap + arange(3)[1]
    }
}

constexpr int TILE = 1024;  // symbols staged per step (32 lanes x 32)

struct BitWriter {
    uint64_t acc = 0;   // bits, MSB first
    uint32_t nb = 0;    // valid bits in acc (< 32 between calls)
    uint32_t *dst;      // word cursor
    uint64_t words = 0, cap_words;
    __device__ __forceinline__ void put(uint32_t v, uint32_t n) {  // 0 <= n <= 32
        if (n == 0) return;
        acc |= ((uint64_t)v << (32 - n)) << (32 - nb);
        nb += n;
        if (nb >= 32) {
            const uint32_t w = (uint32_t)(acc >> 32);
            if (words < cap_words) dst[words] = __byte_perm(w, 0, 0x0123);
            words++;
            acc <<= 32;
            nb -= 32;
        }
    }
    __device__ __forceinline__ void put_run(uint32_t bit, uint64_t count) {
        const uint32_t fill = bit ? 0xFFFFFFFFu : 0u;
        while (count >= 32) { put(fill, 32); count -= 32; }
        if (count) put(fill >> (32 - (uint32_t)count), (uint32_t)count);
    }
};

// one warp (= one CTA of 32 threads) per stream
__global__ void __launch_bounds__(32)
ac_encode_kernel(const uint16_t *__restrict__ c1, const uint8_t *__restrict__ sym,
                 const int64_t *__restrict__ sym_off, uint8_t *__restrict__ out,
                 const int64_t *__restrict__ out_off, int64_t *__restrict__ out_len) {
    __shared__ uint32_t tile[2][TILE];
    const int k = blockIdx.x, lane = threadIdx.x;
    const int64_t s0 = sym_off[k], n = sym_off[k + 1] - s0;
    const int64_t o0 = out_off[k], cap = out_off[k + 1] - o0;
    c1 += s0;
    sym += s0;

    BitWriter bw;
    bw.dst = reinterpret_cast<uint32_t *>(out + o0);
    bw.cap_words = (uint64_t)(cap / 4);
    uint32_t low = 0, high = 0xFFFFFFFFu;
    uint64_t pending = 0;

    auto stage = [&](int buf, int64_t base) {
#pragma unroll 8
        for (int j = lane; j < TILE; j += 32) {
            const int64_t i = base + j;
            tile[buf][j] = (i < n) ? ((uint32_t)__ldg(c1 + i) | ((uint32_t)__ldg(sym + i) << 16)) : 0u;
        }
    };
    if (n > 0) stage(0, 0);
    __syncwarp();
    int buf = 0;
    for (int64_t base = 0; base < n; base += TILE) {
        if (base + TILE < n) stage(buf ^ 1, base + TILE);  // next tile in flight while lane 0 codes
        if (lane == 0) {
            const int m = (int)((n - base) < TILE ? (n - base) : TILE);
            for (int j = 0; j < m; j++) {
                const uint32_t pk = tile[buf][j];
                const uint32_t c = pk & 0xFFFFu, s = pk >> 16;
                const uint32_t r = high - low;                       // span - 1
                const uint32_t t = (uint32_t)(((uint64_t)r * c + c) >> 16);  // (span*c1) >> 16
                if (s) low += t; else high = low + t - 1u;
                // E1/E2: shift out the matching leading bits in one go
                const uint32_t nm = __clz(low ^ high);
                if (nm) {
                    const uint32_t lead = low >> (32 - nm);          // the nm matched bits
                    const uint32_t b0 = lead >> (nm - 1);
                    bw.put(b0, 1);
                    if (pending) { bw.put_run(b0 ^ 1u, pending); pending = 0; }
                    if (nm > 1) bw.put(lead & ((1u << (nm - 1)) - 1u), nm - 1);
                    low = (nm == 32) ? 0u : (low << nm);
                    high = (nm == 32) ? 0xFFFFFFFFu : ((high << nm) | ((1u << nm) - 1u));
                }
                // E3: low = 01.., high = 10.. -> count the underflow positions
                uint32_t ku = __clz(((~low) << 1) | (high << 1));
                ku = ku > 31u ? 31u : ku;
                if (ku) {
                    pending += ku;
                    low = (low << ku) & 0x7FFFFFFFu;
                    high = (high << ku) | 0x80000000u | ((1u << ku) - 1u);
                }
            }
        }
        __syncwarp();
        buf ^= 1;
    }
    if (lane == 0) {
        pending += 1;
        const uint32_t b = low < 0x40000000u ? 0u : 1u;
        bw.put(b, 1);
        bw.put_run(b ^ 1u, pending);
        // flush: whole words are out; the remaining nb (<32) bits go byte by byte, zero padded
        uint64_t bytes = bw.words * 4;
        uint8_t *tail = out + o0;
        uint32_t rem = bw.nb;
        uint64_t acc = bw.acc;
        while (rem > 0) {
            if ((int64_t)bytes < cap) tail[bytes] = (uint8_t)(acc >> 56);
            bytes++;
            acc <<= 8;
            rem = rem > 8 ? rem - 8 : 0;
        }
        out_len[k] = (int64_t)bytes;
    }
}

struct BitReader {
    const uint8_t *src;
    int64_t nbytes, pos = 0;  // pos = next byte to fetch
    uint64_t res = 0;         // upcoming bits, MSB first
    uint32_t navail = 0;
    __device__ __forceinline__ uint32_t fetch_word() {
        uint32_t w = 0;
        if (pos + 4 <= nbytes) {
            w = __byte_perm(*reinterpret_cast<const uint32_t *>(src + pos), 0, 0x0123);
        } else {
            for (int b = 0; b < 4; b++) {
                const int64_t q = pos + b;
                w = (w << 8) | (q < nbytes ? (uint32_t)src[q] : 0u);  // zeros past the end (torchac)
            }
        }
        pos += 4;
        return w;
    }
    __device__ __forceinline__ void refill() {
        if (navail <= 32) {
            res |= (uint64_t)fetch_word() << (32 - navail);
            navail += 32;
        }
    }
    __device__ __forceinline__ uint32_t get(uint32_t n) {  // 0 <= n <= 32
        refill();
        const uint32_t v = n ? (uint32_t)(res >> (64 - n)) : 0u;
        res <<= n;  // n <= 32 < 64
        navail -= n;
        return v;
    }
};

__global__ void __launch_bounds__(32)
ac_decode_kernel(const uint16_t *__restrict__ c1, const int64_t *__restrict__ sym_off,
                 const uint8_t *__restrict__ in, const int64_t *__restrict__ in_off,
                 const int64_t *__restrict__ in_len, uint8_t *__restrict__ sym) {
    __shared__ uint16_t ctile[2][TILE];
    __shared__ uint8_t stile[TILE];
    const int k = blockIdx.x, lane = threadIdx.x;
    const int64_t s0 = sym_off[k], n = sym_off[k + 1] - s0;
    c1 += s0;
    sym += s0;

    BitReader br;
    br.src = in + in_off[k];
    br.nbytes = in_len[k];
    uint32_t low = 0, high = 0xFFFFFFFFu, value = 0;
    if (lane == 0) value = br.get(32);

    auto stage = [&](int buf, int64_t base) {
#pragma unroll 8
        for (int j = lane; j < TILE; j += 32) {
            const int64_t i = base + j;
            ctile[buf][j] = (i < n) ? __ldg(c1 + i) : (uint16_t)1;
        }
    };
    if (n > 0) stage(0, 0);
    __syncwarp();
    int buf = 0;
    for (int64_t base = 0; base < n; base += TILE) {
        if (base + TILE < n) stage(buf ^ 1, base + TILE);
        const int m = (int)((n - base) < TILE ? (n - base) : TILE);
        if (lane == 0) {
            for (int j = 0; j < m; j++) {
                const uint32_t c = ctile[buf][j];
                const uint32_t r = high - low;
                const uint64_t lhs = (uint64_t)r * c + c;                       // c1 * span
                const uint64_t rhs = (((uint64_t)(value - low) + 1) << 16) - 1;  // (value-low+1)*2^16 - 1
                const uint32_t s = lhs <= rhs;                                  // c1 <= count
                stile[j] = (uint8_t)s;
                const uint32_t t = (uint32_t)(lhs >> 16);
                if (s) low += t; else high = low + t - 1u;
                const uint32_t nm = __clz(low ^ high);
                if (nm) {
                    const uint32_t bits = br.get(nm);
                    low = (nm == 32) ? 0u : (low << nm);
                    high = (nm == 32) ? 0xFFFFFFFFu : ((high << nm) | ((1u << nm) - 1u));
                    value = (nm == 32) ? bits : ((value << nm) | bits);
                }
                uint32_t ku = __clz(((~low) << 1) | (high << 1));
                ku = ku > 31u ? 31u : ku;
                if (ku) {
                    const uint32_t bits = br.get(ku);
                    low = (low << ku) & 0x7FFFFFFFu;
                    high = (high << ku) | 0x80000000u | ((1u << ku) - 1u);
                    value = ((value << ku) ^ 0x80000000u) + bits;  // k x {value -= 2^30; shift in a bit}
                }
            }
        }
        __syncwarp();
        for (int j = lane; j < m; j += 32) sym[base + j] = stile[j];
        __syncwarp();
        buf ^= 1;
    }
}

}  // namespace cnc

using namespace cnc;

extern "C" {

int cnc_cdf_from_p(const float *p, uint16_t *c1, uint64_t n, cnc_stream_t stream) {
    if (n == 0) return CNC_OK;
    if (!p || !c1) { set_error("cdf_from_p: null pointer"); return CNC_EINVAL; }
    const uint64_t b = (n + 255) / 256;
    cdf_from_p_kernel<<<(uint32_t)(b < 148 * 16 ? b : 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, c1, n);
    return check_launch("cdf_from_p");
}

int cnc_ac_encode(const uint16_t *c1, const uint8_t *sym, const int64_t *sym_off, uint8_t *out,
                  const int64_t *out_off, int64_t *out_len, int32_t n_streams, cnc_stream_t stream) {
    if (n_streams <= 0) return CNC_OK;
    if (!c1 || !sym || !sym_off || !out || !out_off || !out_len) { set_error("ac_encode: null pointer"); return CNC_EINVAL; }
    if (reinterpret_cast<uintptr_t>(out) & 3u) { set_error("ac_encode: out must be 4-byte aligned"); return CNC_EINVAL; }
    ac_encode_kernel<<<n_streams, 32, 0, static_cast<cudaStream_t>(stream)>>>(c1, sym, sym_off, out, out_off, out_len);
    return check_launch("ac_encode");
}

int cnc_ac_decode(const uint16_t *c1, const int64_t *sym_off, const uint8_t *in, const int64_t *in_off,
                  const int64_t *in_len, uint8_t *sym, int32_t n_streams, cnc_stream_t stream) {
    if (n_streams <= 0) return CNC_OK;
    if (!c1 || !sym_off || !in || !in_off || !in_len || !sym) { set_error("ac_decode: null pointer"); return CNC_EINVAL; }
    if (reinterpret_cast<uintptr_t>(in) & 3u) { set_error("ac_decode: in must be 4-byte aligned"); return CNC_EINVAL; }
    ac_decode_kernel<<<n_streams, 32, 0, static_cast<cudaStream_t>(stream)>>>(c1, sym_off, in, in_off, in_len, sym);
    return check_launch("ac_decode");
}

}  // extern "C"
