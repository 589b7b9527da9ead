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

constexpr int TILE = 2048;  // symbols staged per step

// The serial part of a stream is one dependent chain per symbol:
//   r = high-low -> t = (r*c + c) >> 16 -> select -> x = low^high -> nm = clz(x) -> ku = clz of the underflow run
//   -> one combined shift by S = nm + ku.
// E1/E2 (nm matching leading bits) and E3 (ku underflow positions: low = 0 1^ku.., high = 1 0^ku..) of the
// reference's bit-at-a-time loop are resolved together: after the update span >= 2^14 always holds for
// c1 in [1, 65535] (span > 2^30 before the symbol), so S <= 18 and a single 32-bit shift suffices; the S == 32
// corner of degenerate CDFs (c1 == 0) is kept correct by the clamped funnel shifts.
// Everything that is not on that chain (symbol fetch, bit packing, stores) is written so that it can issue in
// the chain's stall slots.
struct CoderState {
    uint32_t low = 0u, high = 0xFFFFFFFFu;
};

// returns nm (common prefix length) and ku (underflow run); updates the state.  Precondition c in [1, 65535]
// (what cnc_cdf_from_p / torchac's quantiser produce): then span >= 2^14 after the update, nm + ku <= 18.
__device__ __forceinline__ void coder_step(CoderState &st, uint32_t c, uint32_t s, uint32_t &low2, uint32_t &nm, uint32_t &ku) {
    const uint32_t r = st.high - st.low;                                   // span - 1
    const uint32_t t = (uint32_t)(((uint64_t)r * c + c) >> 16);            // (span * c1) >> 16
    const uint32_t nl = st.low + t;
    low2 = s ? nl : st.low;
    const uint32_t high2 = s ? st.high : nl - 1u;
    nm = __clz(low2 ^ high2);
    const uint32_t w = ~low2 | high2;                                      // 0 exactly where (low, high) = (1, 0)
    ku = __clz(((w << nm) << 1) | 0x2000u);                                // run after the first differing bit (bounded by span >= 2^14)
    const uint32_t S = nm + ku;
    st.low = (low2 << S) & 0x7FFFFFFFu;
    st.high = __funnelshift_l(0xFFFFFFFFu, high2, S) | 0x80000000u;
}

struct BitWriter {
    uint64_t acc = 0;   // bits, MSB first
    uint32_t nb = 0;    // valid bits in acc (< 32 between calls)
    uint32_t *dst;      // word cursor
    uint64_t words = 0, cap_words;
    __device__ __forceinline__ void put(uint32_t v, uint32_t n) {  // 0 <= n <= 32, v < 2^n
        acc |= ((uint64_t)v << (32u - n)) << (32u - nb);  // v's n bits go right after the nb valid ones
        nb += n;
        if (nb >= 32u) {
            const uint32_t wv = (uint32_t)(acc >> 32);
            if (words < cap_words) dst[words] = __byte_perm(wv, 0, 0x0123);
            words++;
            acc <<= 32;
            nb -= 32u;
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
    __shared__ __align__(16) uint32_t tile[2][TILE];
    const int k = blockIdx.x, lane = threadIdx.x;
    const int64_t s0 = sym_off[k], n = sym_off[k + 1] - s0;
    const int64_t o0 = out_off[k], cap = out_off[k + 1] - o0;
    c1 += s0;
    sym += s0;

    BitWriter bw;
    bw.dst = reinterpret_cast<uint32_t *>(out + o0);
    bw.cap_words = (uint64_t)(cap / 4);
    CoderState st;
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
            for (int j0 = 0; j0 < m; j0 += 4) {
                const uint4 q = *reinterpret_cast<const uint4 *>(&tile[buf][j0]);
                const uint32_t pk[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (j0 + u < m) {
                        uint32_t low2, nm, ku;
                        coder_step(st, pk[u] & 0xFFFFu, pk[u] >> 16, low2, nm, ku);
                        if (nm) {
                            const uint32_t lead = low2 >> (32u - nm);  // the nm matched bits
                            if (pending) {
                                const uint32_t b0 = lead >> (nm - 1u);
                                bw.put(b0, 1);
                                bw.put_run(b0 ^ 1u, pending);
                                if (nm > 1u) bw.put(lead & ((1u << (nm - 1u)) - 1u), nm - 1u);
                            } else {
                                bw.put(lead, nm);
                            }
                            pending = ku;
                        } else {
                            pending += ku;
                        }
                    }
                }
            }
        }
        __syncwarp();
        buf ^= 1;
    }
    if (lane == 0) {
        pending += 1;
        const uint32_t b = st.low < 0x40000000u ? 0u : 1u;
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

// input bits: the stream's bytes are staged (big-endian words) through shared memory by the whole warp, lane 0
// keeps a 64-bit MSB-aligned reservoir
__global__ void __launch_bounds__(32)
ac_decode_kernel(const uint16_t *__restrict__ c1, const int64_t *__restrict__ sym_off,
                 const uint8_t *__restrict__ in, const int64_t *__restrict__ in_off,
                 const int64_t *__restrict__ in_len, uint8_t *__restrict__ sym) {
    __shared__ __align__(16) uint16_t ctile[2][TILE];
    __shared__ __align__(16) uint8_t stile[TILE];
    constexpr int WT = 1024;                       // input words per staging tile
    __shared__ uint32_t wtile[2][WT];
    const int k = blockIdx.x, lane = threadIdx.x;
    const int64_t s0 = sym_off[k], n = sym_off[k + 1] - s0;
    c1 += s0;
    sym += s0;
    const uint8_t *src = in + in_off[k];
    const int64_t nbytes = in_len[k];

    auto stage_words = [&](int wb, int64_t wbase) {   // words [wbase, wbase+WT), zeros past the end (torchac)
        for (int j = lane; j < WT; j += 32) {
            const int64_t p = (wbase + j) * 4;
            uint32_t w = 0;
            if (p + 4 <= nbytes) w = __byte_perm(*reinterpret_cast<const uint32_t *>(src + p), 0, 0x0123);
            else if (p < nbytes) {
                for (int b = 0; b < 4; b++) w = (w << 8) | (p + b < nbytes ? (uint32_t)src[p + b] : 0u);
            }
            wtile[wb][j] = w;
        }
    };
    auto stage = [&](int buf, int64_t base) {
#pragma unroll 8
        for (int j = lane; j < TILE; j += 32) {
            const int64_t i = base + j;
            ctile[buf][j] = (i < n) ? __ldg(c1 + i) : (uint16_t)1;
        }
    };
    stage_words(0, 0);
    stage_words(1, WT);
    if (n > 0) stage(0, 0);
    __syncwarp();

    // reservoir
    uint64_t res = 0;
    uint32_t navail = 0;
    int64_t wpos = 0;          // next word index to consume
    int64_t staged_hi = 2 * WT; // words [0, staged_hi) have been staged
    auto refill = [&]() {      // lane 0 only: navail < 32 -> append one word
        const uint32_t w = wtile[(wpos / WT) & 1][wpos % WT];
        res |= (uint64_t)w << (32u - navail);
        navail += 32u;
        wpos++;
    };
    CoderState st;
    uint32_t value = 0;
    if (lane == 0) {
        refill();
        value = (uint32_t)(res >> 32);
        res <<= 32;
        navail -= 32u;
        refill();
    }
    int buf = 0;
    for (int64_t base = 0; base < n; base += TILE) {
        if (base + TILE < n) stage(buf ^ 1, base + TILE);
        const int m = (int)((n - base) < TILE ? (n - base) : TILE);
        // a tile of TILE symbols consumes at most TILE * 18 bits < WT words only if TILE*18/32 <= WT: 2048*18/32 = 1152 > 1024,
        // so the word tiles are topped up in two halves of the symbol tile
        for (int half = 0; half < 2; half++) {
            const int jb = half * (TILE / 2), je = min(m, jb + TILE / 2);
            if (lane == 0) {
                for (int j = jb; j < je; j++) {
                    const uint32_t c = ctile[buf][j];
                    // symbol: largest s with cdf[s] <= count  <=>  c1 * span <= ((value - low + 1) << 16) - 1
                    const uint32_t r = st.high - st.low;
                    const uint64_t lhs = (uint64_t)r * c + c;
                    const uint64_t rhs = (((uint64_t)(value - st.low) + 1ull) << 16) - 1ull;
                    const uint32_t s = lhs <= rhs;
                    stile[j] = (uint8_t)s;
                    uint32_t low2, nm, ku;
                    coder_step(st, c, s, low2, nm, ku);
                    const uint32_t S = nm + ku;
                    if (S) {
                        const uint32_t bits = (uint32_t)(res >> (64u - S));
                        res <<= S;
                        navail -= S;
                        value = (value << S) | bits;
                        if (ku) value ^= 0x80000000u;   // ku x {value -= 2^30; shift in a bit}
                        if (navail < 32u) refill();
                    }
                }
            }
            __syncwarp();
            // keep two word tiles ahead of the consumer
            const int64_t wp = __shfl_sync(0xFFFFFFFFu, wpos, 0);
            while (staged_hi - wp <= WT) {
                stage_words((int)((staged_hi / WT) & 1), staged_hi);
                staged_hi += WT;
            }
            __syncwarp();
        }
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
