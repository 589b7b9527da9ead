// coder.cu -- Bernoulli CDF quantiser and a torchac-compatible 32-bit binary range coder.
//
// Reference behaviour restated: examples/utils_bpp_acc.py:77-110 builds the float CDF
// [0, 1-p, 1] on the CPU and hands it to torchac==0.9.3 (third party, not in the reference
// tree; algorithm restated in SURVEY Appendix B and oracle/cnc_oracle.c).  The bytes produced
// here are identical to that coder's for equal (c1, symbol) inputs.
//
// Range coding is sequential inside a stream (the interval carries from symbol to symbol), so the
// B200 mapping is stream-parallel (the product emits 33 independent streams) and everything that is
// not the carried dependency is taken off the lane that runs it:
//   * encoder: one CTA of two warps per stream.  Warp 0 stages (c1, sym) tiles with coalesced loads
//     and its lane 0 runs the chain -- one 32x16-bit multiply, selects, ONE find-leading-one and a
//     comparison (closed form of the reference's bit-at-a-time renormalisation loop, see below) --
//     leaving (low, high) per symbol in shared memory.  Warp 1 converts those records into bits with
//     warp scans (matched-prefix lengths, pending-bit counts, bit offsets), ORs them into a shared
//     ring and stores finished words coalesced.
//   * decoder: one warp per stream, state (low, range, value - low), torchac's 64-bit division
//     replaced by the equivalent comparison value - low >= t, input bits from a three-word register
//     window that is advanced with predicated moves (no branch in the loop body).
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

// ------------------------------------------------------------------------------------------
// The serial part of a stream is one dependent chain per symbol.  State: low and r = high - low (both 32 bit;
// the decoder adds d = value - low).  With t = (span * c1) >> 16 = (r * c1 + c1) >> 16:
//     s = 1: low2 = low + t, r2 = r - t          s = 0: low2 = low, r2 = t - 1            high2 = low2 + r2
// The reference renormalises bit by bit (E1/E2: equal top bits are shifted out; E3: low = 01.., high = 10.. drops
// the second bit).  Every one of those steps maps (low, high, value) -> (2 low, 2 high + 1, 2 value + bit) minus
// a multiple of 2^31, and the loop stops at the largest S for which the window still fits, i.e. for which the
// top S+1 bits of high2 and low2 differ by at most one.  Because 2^(31-c0) <= r2 < 2^(32-c0) with c0 = clz(r2),
// that S is either c0 - 1 (always feasible) or c0, decided by one comparison:
//     S = c0 - 1 + [ (high2 >> (31 - c0)) - (low2 >> (31 - c0)) <= 1 ]
// and the new state is low = (low2 << S) & 0x7FFFFFFF, r = ~(~r2 << S), d = (d2 << S) | next S input bits.
// (oracle/cnc_oracle.c runs the bit loop; tests/test_oracle_golden.py checks this closed form against it.)
// One find-leading-one and no branch per symbol, against two and a data-dependent loop before.  c1 in
// [1, 65535] (what the quantiser produces) keeps r2 >= 2^14 - 1, so S <= 18.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t renorm_shift(uint32_t low2, uint32_t high2, uint32_t r2) {
    const uint32_t c0 = __clz(r2), sh = 31u - c0;
    return c0 - 1u + (((high2 >> sh) - (low2 >> sh)) <= 1u ? 1u : 0u);
}


// ---- the chain as it is executed: renormalisation deferred into the next symbol --------------------------------
// Carried state: R = r2, L2 = low2, H2 = high2 of the previous symbol (before its renormalisation).  With
// f = 31 - clz(R) and diff = (H2 >> f) - (L2 >> f) in {1, 2} the shift is S = 32 - g, g = f + diff in [1, 32], and
// everything symbol i needs from the renormalised state is a 64-bit quantity shifted right by g:
//     t      = ((R + 1) << S) * c >> 16      = ((R + 1) * (c << 16)) >> g
//     ~r     = ~R << S                         = (~R : 0) >> g
//     low    = (L2 << S) & 0x7FFFFFFF          = (L2 : 0) >> g, masked
//     d      = (D2 << S) | top S input bits    = (D2 : window) >> g
// (clamped funnel shifts: g = 32 returns the high word).  The 32x16 multiply does not wait for the shift any more
// -- it runs beside the find-leading-one, the longest-latency instruction of the chain -- and the dependent path
// per symbol is FLO -> 2 shifts -> add -> funnel shift -> [compare -> select].  A warp instruction holds its
// 16-lane pipe for two cycles whatever the number of active lanes, so the body is also kept short (~20
// instructions for the encoder, ~32 for the decoder).
__device__ __forceinline__ uint32_t shr64c(uint32_t hi, uint32_t lo, uint32_t g) { return __funnelshift_rc(lo, hi, g); }   // g <= 32
// (R + 1) * c16 and floor(log2 R) as one asm block, multiply first: both only depend on R, and the multiply has to be
// in flight while the find-leading-one (variable latency, ~19 cycles) completes -- left to the scheduler it is
// sunk behind the first consumer of f and ends up serialised after it
__device__ __forceinline__ void mul_and_flo(uint32_t R, uint32_t c16, uint32_t &qhi, uint32_t &qlo, uint32_t &f) {
    asm volatile("{\n\t.reg .u64 q, a;\n\t"
                 "mov.b64 a, {%4, %5};\n\t"
                 "mad.wide.u32 q, %3, %4, a;\n\t"
                 "bfind.u32 %2, %3;\n\t"
                 "mov.b64 {%0, %1}, q;\n\t}" : "=r"(qlo), "=r"(qhi), "=r"(f) : "r"(R), "r"(c16), "r"(0u));
}
// low after the renormalisation of the last symbol (encoder termination)
__device__ __forceinline__ uint32_t renorm_low(uint32_t R, uint32_t L2, uint32_t H2) {
    return (L2 << renorm_shift(L2, H2, R)) & 0x7FFFFFFFu;
}

constexpr int ET = 1024;            // symbols per tile
constexpr int RING_WORDS = 2048;    // output ring of the bit packer (power of two)
constexpr uint32_t SLOW_PEND = 1024;  // pending runs longer than this take the sequential emit path

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Bit packer state of one stream (held warp-uniformly by the 32 lanes of the packer warp).  Bits are MSB first;
// finished 32-bit words leave the shared ring as coalesced big-endian stores.
struct Packer {
    uint32_t *ring;          // [RING_WORDS], zero outside the unflushed window
    uint32_t *dst;           // output words
    uint8_t *dst8;
    uint64_t cap_bytes;
    uint64_t gbit = 0;       // bits emitted so far
    uint64_t flushed = 0;    // words already stored
    uint64_t pend = 0;       // pending underflow count

    // n <= 32 bits of v at absolute bit position pos (executed by the calling lane only)
    __device__ __forceinline__ void or_bits(uint64_t pos, uint32_t v, uint32_t n) {
        if (n == 0) return;
        const uint32_t o = (uint32_t)pos & 31u;
        const uint64_t V = ((uint64_t)v << (32u - n)) << (32u - o);
        const uint32_t w = (uint32_t)(pos >> 5);
        const uint32_t hi = (uint32_t)(V >> 32), lo = (uint32_t)V;
        if (hi) atomicOr(&ring[w & (RING_WORDS - 1)], hi);
        if (lo) atomicOr(&ring[(w + 1) & (RING_WORDS - 1)], lo);
    }
    // whole warp: store every complete word while at least `keep` of them are waiting
    __device__ __forceinline__ void flush(int lane, uint64_t min_words) {
        __syncwarp();
        const uint64_t complete = gbit >> 5;
        while (complete - flushed >= min_words && complete > flushed) {
            const uint64_t w = flushed + (uint64_t)lane;
            if (w < complete) {
                const uint32_t v = ring[w & (RING_WORDS - 1)];
                ring[w & (RING_WORDS - 1)] = 0u;
                if (w * 4 + 4 <= cap_bytes) dst[w] = __byte_perm(v, 0, 0x0123);
            }
            flushed = (complete - flushed > 32) ? flushed + 32 : complete;
        }
        __syncwarp();
    }
    // whole warp, sequential semantics: `count` copies of `bit`
    __device__ __forceinline__ void emit_run(int lane, uint32_t bit, uint64_t count) {
        while (count > 0) {
            const uint32_t chunk = count > 1024 ? 1024u : (uint32_t)count;
            if (bit) {
                const uint32_t my = (uint32_t)lane * 32u;
                if (my < chunk) {
                    const uint32_t n = chunk - my >= 32u ? 32u : chunk - my;
                    or_bits(gbit + my, n == 32u ? 0xFFFFFFFFu : ((1u << n) - 1u), n);
                }
            }
            gbit += chunk;
            count -= chunk;
            flush(lane, 32);
        }
    }
    __device__ __forceinline__ void emit_bits(int lane, uint32_t v, uint32_t n) {   // n <= 32
        if (lane == 0) or_bits(gbit, v, n);
        gbit += n;
    }
    // reference bo_bit_and_pending for the nm matched bits `lead` of one symbol
    __device__ __forceinline__ void emit_symbol_seq(int lane, uint32_t lead, uint32_t nm) {
        const uint32_t b0 = lead >> (nm - 1u);
        emit_bits(lane, b0, 1);
        emit_run(lane, b0 ^ 1u, pend);
        pend = 0;
        emit_bits(lane, lead & ((1u << (nm - 1u)) - 1u), nm - 1u);
        flush(lane, 32);
    }
};

// One CTA of two warps per stream.  Warp 0: all lanes stage (c1, symbol) tiles, lane 0 runs the dependent chain and
// leaves (low2, high2) per symbol in shared memory.  Warp 1 turns those records into bits with warp scans (matched
// prefix lengths, pending counts, bit offsets) one tile behind, so nothing but the chain is on the chain's lane.
__global__ void __launch_bounds__(64)
ac_encode_kernel(const uint16_t *__restrict__ c1, const uint8_t *__restrict__ sym,
                 const int64_t *__restrict__ sym_off, uint8_t *__restrict__ out,
                 const int64_t *__restrict__ out_off, int64_t *__restrict__ out_len) {
    __shared__ __align__(16) uint32_t in_tile[2][ET];
    __shared__ __align__(16) uint2 rec[2][ET];
    __shared__ uint32_t ring[RING_WORDS];
    __shared__ uint32_t final_low;
    const int k = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t s0 = sym_off[k], n = sym_off[k + 1] - s0;
    const int64_t ntiles = (n + ET - 1) / ET;
    constexpr int BAR_FULL = 1, BAR_EMPTY = 3;   // + buffer index

    if (warp == 0) {
        c1 += s0;
        sym += s0;
        auto stage = [&](int buf, int64_t base) {
#pragma unroll 8
            for (int j = lane; j < ET; j += 32) {
                const int64_t i = base + j;
                in_tile[buf][j] = (i < n) ? ((uint32_t)__ldg(c1 + i) | ((uint32_t)__ldg(sym + i) << 16)) : 0u;
            }
        };
        if (n > 0) stage(0, 0);
        __syncwarp();
        uint32_t R = 0xFFFFFFFFu, L2 = 0u, H2 = 0xFFFFFFFFu;   // "previous symbol" of the initial state: S = 0
        for (int64_t t = 0; t < ntiles; t++) {
            const int buf = (int)(t & 1);
            if (t + 1 < ntiles) stage(buf ^ 1, (t + 1) * ET);   // in flight while lane 0 codes
            if (t >= 2) named_bar_sync(BAR_EMPTY + buf, 64);    // the packer is done with rec[buf]
            if (lane == 0) {
                const int m = (int)((n - t * ET) < ET ? (n - t * ET) : ET);
                const uint32_t *src = in_tile[buf];
                uint2 *dstrec = rec[buf];
                auto step = [&](uint32_t pk, int j) {
                    const uint32_t c16 = pk << 16;                  // c1 << 16 (the symbol bit sits above it)
                    const bool s = (pk >> 16) != 0u;
                    uint32_t qhi, qlo, f;
                    mul_and_flo(R, c16, qhi, qlo, f);               // (R + 1) * c1 << 16, beside the FLO
                    const uint32_t g = f + (H2 >> f) - (L2 >> f);
                    const uint32_t t = shr64c(qhi, qlo, g);
                    const uint32_t nr = shr64c(~R, 0u, g);
                    const uint32_t low = shr64c(L2, 0u, g) & 0x7FFFFFFFu;
                    R = s ? ~nr - t : t - 1u;
                    L2 = s ? low + t : low;
                    H2 = L2 + R;
                    dstrec[j] = make_uint2(L2, H2);
                };
                int j = 0;
                uint4 q = *reinterpret_cast<const uint4 *>(src);
                for (; j + 4 <= m; j += 4) {
                    const uint4 cur = q;
                    q = *reinterpret_cast<const uint4 *>(src + ((j + 4) & (ET - 1)));   // next group, off the chain
                    step(cur.x, j);
                    step(cur.y, j + 1);
                    step(cur.z, j + 2);
                    step(cur.w, j + 3);
                }
                for (; j < m; j++) step(src[j], j);
                if (t == ntiles - 1) final_low = renorm_low(R, L2, H2);
            }
            __syncwarp();
            __threadfence_block();
            named_bar_arrive(BAR_FULL + buf, 64);
        }
        return;
    }

    // ------------------------------------------------------------------ packer warp
    for (int j = lane; j < RING_WORDS; j += 32) ring[j] = 0u;
    Packer pk;
    pk.ring = ring;
    pk.dst8 = out + out_off[k];
    pk.dst = reinterpret_cast<uint32_t *>(pk.dst8);
    pk.cap_bytes = (uint64_t)(out_off[k + 1] - out_off[k]);
    __syncwarp();
    for (int64_t t = 0; t < ntiles; t++) {
        const int buf = (int)(t & 1);
        named_bar_sync(BAR_FULL + buf, 64);
        const int m = (int)((n - t * ET) < ET ? (n - t * ET) : ET);
        for (int g = 0; g < m; g += 32) {
            const int j = g + lane;
            uint32_t nm = 0, ku = 0, lead = 0;
            if (j < m) {
                const uint2 lh = rec[buf][j];
                nm = __clz(lh.x ^ lh.y);
                ku = renorm_shift(lh.x, lh.y, lh.y - lh.x) - nm;
                lead = nm ? lh.x >> (32u - nm) : 0u;
            }
            // P = inclusive sum of ku; key = (sum before this symbol) + 1 at symbols that emit
            uint32_t P = ku;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, P, d);
                if (lane >= d) P += v;
            }
            const uint32_t Pex = P - ku;
            const uint32_t key = nm ? Pex + 1u : 0u;
            uint32_t kmax = key;          // inclusive prefix max
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, kmax, d);
                if (lane >= d) kmax = max(kmax, v);
            }
            uint32_t B = __shfl_up_sync(0xFFFFFFFFu, kmax, 1);
            if (lane == 0) B = 0u;        // exclusive: last emitter strictly before this symbol
            const uint64_t pend_i = B ? (uint64_t)(Pex - (B - 1u)) : pk.pend + Pex;
            const uint32_t Ptot = __shfl_sync(0xFFFFFFFFu, P, 31), klast = __shfl_sync(0xFFFFFFFFu, kmax, 31);
            const bool slow = __any_sync(0xFFFFFFFFu, nm && pend_i > SLOW_PEND);
            if (!slow) {
                const uint32_t pe = (uint32_t)pend_i;
                const uint32_t len = nm ? nm + pe : 0u;
                uint32_t off = len;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, off, d);
                    if (lane >= d) off += v;
                }
                const uint32_t total = __shfl_sync(0xFFFFFFFFu, off, 31);
                off -= len;
                if (nm) {
                    const uint64_t pos = pk.gbit + off;
                    const uint32_t b0 = lead >> (nm - 1u), rest = lead & ((1u << (nm - 1u)) - 1u);
                    if (len <= 32u) {
                        const uint32_t run = b0 ? 0u : (((1u << pe) - 1u) << (nm - 1u));   // pe <= 31 here
                        pk.or_bits(pos, (b0 << (len - 1u)) | run | rest, len);
                    } else {
                        pk.or_bits(pos, b0, 1);
                        if (!b0) {
                            uint64_t p2 = pos + 1;
                            for (uint32_t left = pe; left > 0;) {
                                const uint32_t cnk = left >= 32u ? 32u : left;
                                pk.or_bits(p2, cnk == 32u ? 0xFFFFFFFFu : ((1u << cnk) - 1u), cnk);
                                p2 += cnk;
                                left -= cnk;
                            }
                        }
                        pk.or_bits(pos + 1 + pe, rest, nm - 1u);
                    }
                }
                pk.gbit += total;
                pk.pend = klast ? (uint64_t)(Ptot - (klast - 1u)) : pk.pend + Ptot;
                pk.flush(lane, 32);
            } else {   // sequential semantics, symbol by symbol (astronomically rare: > SLOW_PEND pending bits)
                for (int i = 0; i < 32; i++) {
                    const uint32_t nmi = __shfl_sync(0xFFFFFFFFu, nm, i), kui = __shfl_sync(0xFFFFFFFFu, ku, i);
                    const uint32_t li = __shfl_sync(0xFFFFFFFFu, lead, i);
                    if (nmi) pk.emit_symbol_seq(lane, li, nmi);
                    pk.pend += kui;
                }
            }
        }
        if (t + 2 < ntiles) named_bar_arrive(BAR_EMPTY + buf, 64);
    }
    // termination: one more pending bit, then the quarter bit and its complements (reference tail)
    {
        const uint32_t lowf = n > 0 ? final_low : 0u;
        pk.pend += 1;
        const uint32_t b = lowf < 0x40000000u ? 0u : 1u;
        pk.emit_bits(lane, b, 1);
        pk.emit_run(lane, b ^ 1u, pk.pend);
        pk.flush(lane, 1);
        // the last partial word leaves byte by byte, zero padded
        const uint64_t bytes = (pk.gbit + 7) >> 3;
        const uint64_t done = pk.flushed * 4;
        if (lane == 0) {
            const uint32_t v = ring[pk.flushed & (RING_WORDS - 1)];
            for (uint64_t b8 = done; b8 < bytes; b8++)
                if (b8 < pk.cap_bytes) pk.dst8[b8] = (uint8_t)(v >> (24u - 8u * (uint32_t)(b8 - done)));
            out_len[k] = (int64_t)bytes;
        }
    }
}

// Decoder: one warp per stream.  All lanes stage the CDF entries and the input words (big-endian) through shared
// memory; lane 0 runs the chain on (low, r, d = value - low).  The input cursor is three prefetched words and a
// bit offset, advanced with predicated moves, so the loop body has no branch and one shared load (off the chain).
constexpr int TILE = 2048;  // symbols staged per step
__global__ void __launch_bounds__(32)
ac_decode_kernel(const uint16_t *__restrict__ c1, const int64_t *__restrict__ sym_off,
                 const uint8_t *__restrict__ in, const int64_t *__restrict__ in_off,
                 const int64_t *__restrict__ in_len, uint8_t *__restrict__ sym) {
    __shared__ __align__(16) uint16_t ctile[2][TILE];
    __shared__ __align__(16) uint8_t stile[TILE];
    constexpr int WT = 1024;                       // input words per staging tile
    __shared__ __align__(16) uint4 wtrip[2 * WT];  // (word[i], word[i+1], word[i+2], -): one aligned load covers a cursor step
    const int k = blockIdx.x, lane = threadIdx.x;
    const int64_t s0 = sym_off[k], n = sym_off[k + 1] - s0;
    c1 += s0;
    sym += s0;
    const uint8_t *src = in + in_off[k];
    const int64_t nbytes = in_len[k];

    auto load_word = [&](int64_t widx) -> uint32_t {   // big-endian word of the stream, zeros past the end (torchac)
        const int64_t p = widx * 4;
        uint32_t w = 0;
        if (p + 4 <= nbytes) w = __byte_perm(*reinterpret_cast<const uint32_t *>(src + p), 0, 0x0123);
        else if (p < nbytes) {
            for (int b = 0; b < 4; b++) w = (w << 8) | (p + b < nbytes ? (uint32_t)src[p + b] : 0u);
        }
        return w;
    };
    auto stage_words = [&](int64_t wbase) {   // triples [wbase, wbase+WT)
        for (int j = lane; j < WT; j += 32) {
            const uint32_t w0 = load_word(wbase + j);
            uint32_t w1 = __shfl_down_sync(0xFFFFFFFFu, w0, 1), w2 = __shfl_down_sync(0xFFFFFFFFu, w0, 2);
            if (lane == 31) w1 = load_word(wbase + j + 1);
            if (lane >= 30) w2 = load_word(wbase + j + 2);
            wtrip[(wbase + j) & (2 * WT - 1)] = make_uint4(w0, w1, w2, 0u);
        }
    };
    auto stage = [&](int buf, int64_t base) {
#pragma unroll 8
        for (int j = lane; j < TILE; j += 32) {
            const int64_t i = base + j;
            ctile[buf][j] = (i < n) ? __ldg(c1 + i) : (uint16_t)1;
        }
    };
    stage_words(0);
    stage_words(WT);
    if (n > 0) stage(0, 0);
    __syncwarp();

    int64_t staged_hi = 2 * WT;   // triples [0, staged_hi) have been staged
    uint32_t R = 0xFFFFFFFFu, L2 = 0u, H2 = 0xFFFFFFFFu, D2 = 0u;   // state before the (deferred) renormalisation
    uint32_t win = 0u;            // the 32 input bits at the cursor
    uint64_t bp = 32;             // input cursor in bits
    if (lane == 0) {
        D2 = wtrip[0].x;
        win = wtrip[0].y;
    }
    int buf = 0;
    for (int64_t base = 0; base < n; base += TILE) {
        if (base + TILE < n) stage(buf ^ 1, base + TILE);
        const int m = (int)((n - base) < TILE ? (n - base) : TILE);
        // a half tile consumes at most 1024 * 18 bits = 576 words < WT: the word ring is topped up in between
        for (int half = 0; half < 2; half++) {
            const int jb = half * (TILE / 2), je = min(m, jb + TILE / 2);
            if (lane == 0) {
                const uint16_t *cs = ctile[buf];
                uint32_t bpl = (uint32_t)bp;   // the low 32 bits of the cursor are enough inside a half tile
                uint4 w = wtrip[(bpl >> 5) & (2 * WT - 1)];   // the three words at the cursor, loaded one symbol ahead
                auto step = [&](uint32_t c, int j) {
                    const uint32_t c16 = c << 16;
                    uint32_t qhi, qlo, f;
                    mul_and_flo(R, c16, qhi, qlo, f);
                    const uint32_t g = f + (H2 >> f) - (L2 >> f);
                    const uint32_t t = shr64c(qhi, qlo, g);
                    const uint32_t d = shr64c(D2, win, g);            // value - low after the shift
                    const uint32_t nr = shr64c(~R, 0u, g);
                    const uint32_t low = shr64c(L2, 0u, g) & 0x7FFFFFFFu;
                    const bool s = d >= t;                            // <=> c1 <= ((value - low + 1) * 2^16 - 1) / span
                    R = s ? ~nr - t : t - 1u;
                    L2 = s ? low + t : low;
                    D2 = s ? d - t : d;
                    H2 = L2 + R;
                    stile[j] = (uint8_t)s;
                    const uint32_t bpn = bpl + 32u - g;               // the cursor advances by the shift just applied (<= 18 bits)
                    const bool cross = ((bpn ^ bpl) & 32u) != 0u;
                    win = __funnelshift_l(cross ? w.z : w.y, cross ? w.y : w.x, bpn);
                    bpl = bpn;
                    w = wtrip[(bpl >> 5) & (2 * WT - 1)];
                };
                int j = jb;
                for (; j + 4 <= je; j += 4) {
                    const uint2 q = *reinterpret_cast<const uint2 *>(cs + j);
                    step(q.x & 0xFFFFu, j);
                    step(q.x >> 16, j + 1);
                    step(q.y & 0xFFFFu, j + 2);
                    step(q.y >> 16, j + 3);
                }
                for (; j < je; j++) step(cs[j], j);
                bp += (uint64_t)(bpl - (uint32_t)bp);
            }
            __syncwarp();
            // keep more than one word tile ahead of the consumer
            const int64_t wp = (int64_t)(__shfl_sync(0xFFFFFFFFu, bp, 0) >> 5);
            while (staged_hi - wp <= WT) {
                stage_words(staged_hi);
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
    ac_encode_kernel<<<n_streams, 64, 0, static_cast<cudaStream_t>(stream)>>>(c1, sym, sym_off, out, out_off, out_len);
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
