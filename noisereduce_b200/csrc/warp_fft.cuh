// warp_fft.cuh -- a 1024-point complex FFT held entirely by ONE warp (32 lanes x 32 points).
//
// Layout.  A lane owns the 32 points  n = lane + 32*r  (r = register slot), so every global/shared
// access of "slot r across the warp" is one coalesced 128-byte row.  The transform is two radix-32
// passes (1024 = 32 x 32, Stockham auto-sort):
//
//   pass 1:  V[lane][q] = DFT32_r( x[lane + 32 r] )                     (registers only)
//   twiddle: V[lane][q] *= exp(-2 pi i * lane * q / 1024)               (table tw[q][lane], smem)
//   exchange: a 32x32 transpose through a per-warp 32x33 float tile     (the ONLY data movement)
//   pass 2:  X[lane + 32 q] = DFT32_r( V[r][lane] )                     (registers only)
//
// Only __syncwarp() is needed: a warp never waits for another warp.  The radix-32 butterfly is a
// fully unrolled radix-2 DIF network whose 32-point twiddles are compile-time immediates.
//
// Bit-reversal is never executed: the DIF network leaves frequency q in register slot brev5(q) and all
// callers index slots through brev5() with compile-time q.  Stages 2..5 of the network run on packed
// f32x2 registers (dft32_packed); dft32_scalar is the reference form of the same network.
//
// Inverse transform: ifft(z) = swap(fft(swap(z))) / N with swap(a + ib) = b + ia, i.e. simply call
// warp_fft1024(im, re) -- the scaling is folded into the synthesis window by the callers.
#pragma once
#include "cuda_compat.h"

namespace b200 {

constexpr int kFftN = 1024;
constexpr int kExchFloats = 32 * 33;   // per-warp exchange tile (one plane at a time)

__host__ __device__ __forceinline__ constexpr int brev5(int i) {
    return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// cos / sin of 2*pi*m/32, m = 0..15 (folds to an immediate once the loops are unrolled)
__host__ __device__ __forceinline__ constexpr float cos32(int m) {
    return m == 0 ? 1.0f : m == 1 ? 0.98078528040323044913f : m == 2 ? 0.92387953251128675613f
         : m == 3 ? 0.83146961230254523708f : m == 4 ? 0.70710678118654752440f
         : m == 5 ? 0.55557023301960222474f : m == 6 ? 0.38268343236508977173f
         : m == 7 ? 0.19509032201612826785f : m == 8 ? 0.0f
         : m == 9 ? -0.19509032201612826785f : m == 10 ? -0.38268343236508977173f
         : m == 11 ? -0.55557023301960222474f : m == 12 ? -0.70710678118654752440f
         : m == 13 ? -0.83146961230254523708f : m == 14 ? -0.92387953251128675613f
         : -0.98078528040323044913f;
}
__host__ __device__ __forceinline__ constexpr float sin32(int m) {
    return m == 0 ? 0.0f : m == 1 ? 0.19509032201612826785f : m == 2 ? 0.38268343236508977173f
         : m == 3 ? 0.55557023301960222474f : m == 4 ? 0.70710678118654752440f
         : m == 5 ? 0.83146961230254523708f : m == 6 ? 0.92387953251128675613f
         : m == 7 ? 0.98078528040323044913f : m == 8 ? 1.0f
         : m == 9 ? 0.98078528040323044913f : m == 10 ? 0.92387953251128675613f
         : m == 11 ? 0.83146961230254523708f : m == 12 ? 0.70710678118654752440f
         : m == 13 ? 0.55557023301960222474f : m == 14 ? 0.38268343236508977173f
         : 0.19509032201612826785f;
}

// ---- packed FP32 pairs (Blackwell add/sub/mul/fma .f32x2: one instruction, two lanes of a 64-bit register) ----
// The FP32 pipe already runs scalar FADD/FFMA at ~124 of 128 lanes/clk/SM (profiles/r01_g_f32x2_microbench.txt),
// so packing does not raise FLOP/s; it halves the ISSUE slots (and instruction bytes) of the butterfly network,
// which is what these instruction-issue / instruction-fetch bound kernels are short of.
#ifdef B200_CUSIM_BUILD
struct f2 { float lo, hi; };
__device__ __forceinline__ f2 f2_pack(float a, float b) { return f2{a, b}; }
__device__ __forceinline__ void f2_unpack(f2 v, float& a, float& b) { a = v.lo; b = v.hi; }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) { return f2{a.lo + b.lo, a.hi + b.hi}; }
__device__ __forceinline__ f2 f2_sub(f2 a, f2 b) { return f2{a.lo - b.lo, a.hi - b.hi}; }
__device__ __forceinline__ f2 f2_mul_s(f2 a, float c) { return f2{a.lo * c, a.hi * c}; }
__device__ __forceinline__ f2 f2_fma_s(f2 a, float c, f2 acc) { return f2{fmaf(a.lo, c, acc.lo), fmaf(a.hi, c, acc.hi)}; }
#else
typedef unsigned long long f2;
__device__ __forceinline__ f2 f2_pack(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f2_unpack(f2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 f2_sub(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 f2_mul_s(f2 a, float c) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(f2_pack(c, c))); return r; }
__device__ __forceinline__ f2 f2_fma_s(f2 a, float c, f2 acc) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(f2_pack(c, c)), "l"(acc)); return r; }
#endif

#ifdef B200_CUSIM_BUILD
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) { return f2{a.lo * b.lo, a.hi * b.hi}; }
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) { return f2{fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)}; }
#else
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
#endif

#ifndef B200_PACKED_DFT
#define B200_PACKED_DFT 1
#endif

// 32-point DFT over the register slots of one thread (radix-2 decimation in frequency).
//   BREV_IN == false: logical input i in slot i          -> frequency q in slot brev5(q)
//   BREV_IN == true : logical input i in slot brev5(i)   -> frequency q in slot q
template <bool BREV_IN>
__device__ __forceinline__ void dft32_scalar(float (&re)[32], float (&im)[32]) {
#pragma unroll
    for (int len = 32; len >= 2; len >>= 1) {
        const int half = len >> 1;
        const int step = 32 / len;
#pragma unroll
        for (int blk = 0; blk < 32; blk += len) {
#pragma unroll
            for (int i = 0; i < half; ++i) {
                const int ia = BREV_IN ? brev5(blk + i) : (blk + i);
                const int ib = BREV_IN ? brev5(blk + i + half) : (blk + i + half);
                const int m = i * step;                       // twiddle W32^m = cos32(m) - i sin32(m)
                const float ar = re[ia], ai = im[ia], br = re[ib], bi = im[ib];
                re[ia] = ar + br;
                im[ia] = ai + bi;
                const float dr = ar - br, di = ai - bi;
                if (m == 0) {
                    re[ib] = dr;
                    im[ib] = di;
                } else if (m == 8) {                           // * (-i)
                    re[ib] = di;
                    im[ib] = -dr;
                } else if (m == 4) {                           // * (1 - i)/sqrt2
                    re[ib] = (dr + di) * 0.70710678118654752440f;
                    im[ib] = (di - dr) * 0.70710678118654752440f;
                } else if (m == 12) {                          // * (-1 - i)/sqrt2
                    re[ib] = (di - dr) * 0.70710678118654752440f;
                    im[ib] = -(dr + di) * 0.70710678118654752440f;
                } else {
                    const float c = cos32(m), s = sin32(m);
                    re[ib] = fmaf(di, s, dr * c);              // (dr + i di)(c - i s)
                    im[ib] = fmaf(-dr, s, di * c);
                }
            }
        }
    }
}

// Same network, natural slots in -> brev5 slots out.  Stage 1 (partners i, i+16) is scalar; its outputs are
// paired (slot i, slot i+16) so that stages 2..5, which treat both halves identically (same partner offsets, same
// twiddles), run on packed registers: 4 x 8 packed butterflies instead of 4 x 16 scalar ones.
__device__ __forceinline__ void dft32_packed_core(const float (&re)[32], const float (&im)[32], f2 (&Pr)[16], f2 (&Pi)[16]) {
    constexpr float kH = 0.70710678118654752440f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float ar = re[i], ai = im[i], br = re[i + 16], bi = im[i + 16];
        const float sr = ar + br, si = ai + bi, dr = ar - br, di = ai - bi;
        float tr, ti;
        if (i == 0) { tr = dr; ti = di; }
        else if (i == 8) { tr = di; ti = -dr; }
        else if (i == 4) { tr = (dr + di) * kH; ti = (di - dr) * kH; }
        else if (i == 12) { tr = (di - dr) * kH; ti = -(dr + di) * kH; }
        else { const float c = cos32(i), s = sin32(i); tr = fmaf(di, s, dr * c); ti = fmaf(-dr, s, di * c); }
        Pr[i] = f2_pack(sr, tr);
        Pi[i] = f2_pack(si, ti);
    }
#pragma unroll
    for (int len = 16; len >= 2; len >>= 1) {
        const int half = len >> 1, step = 32 / len;
#pragma unroll
        for (int blk = 0; blk < 16; blk += len) {
#pragma unroll
            for (int i = 0; i < half; ++i) {
                const int ia = blk + i, ib = blk + i + half, m = i * step;
                const f2 ar = Pr[ia], ai = Pi[ia], br = Pr[ib], bi = Pi[ib];
                Pr[ia] = f2_add(ar, br);
                Pi[ia] = f2_add(ai, bi);
                if (m == 0) {
                    Pr[ib] = f2_sub(ar, br);
                    Pi[ib] = f2_sub(ai, bi);
                } else if (m == 8) {                           // * (-i): (di, -dr)
                    Pr[ib] = f2_sub(ai, bi);
                    Pi[ib] = f2_sub(br, ar);
                } else if (m == 4) {
                    const f2 dr = f2_sub(ar, br), di = f2_sub(ai, bi);
                    Pr[ib] = f2_mul_s(f2_add(dr, di), kH);
                    Pi[ib] = f2_mul_s(f2_sub(di, dr), kH);
                } else if (m == 12) {
                    const f2 dr = f2_sub(ar, br), di = f2_sub(ai, bi);
                    Pr[ib] = f2_mul_s(f2_sub(di, dr), kH);
                    Pi[ib] = f2_mul_s(f2_add(dr, di), -kH);
                } else {
                    const f2 dr = f2_sub(ar, br), di = f2_sub(ai, bi);
                    const float c = cos32(m), s = sin32(m);
                    Pr[ib] = f2_fma_s(di, s, f2_mul_s(dr, c));
                    Pi[ib] = f2_fma_s(dr, -s, f2_mul_s(di, c));
                }
            }
        }
    }
}
__device__ __forceinline__ void dft32_packed(float (&re)[32], float (&im)[32]) {
    f2 Pr[16], Pi[16];
    dft32_packed_core(re, im, Pr, Pi);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        f2_unpack(Pr[i], re[i], re[i + 16]);
        f2_unpack(Pi[i], im[i], im[i + 16]);
    }
}

template <bool BREV_IN>
__device__ __forceinline__ void dft32(float (&re)[32], float (&im)[32]) {
#if B200_PACKED_DFT
    if (!BREV_IN) { dft32_packed(re, im); return; }
#endif
    dft32_scalar<BREV_IN>(re, im);
}

// Twiddle by exp(-2 pi i lane q / 1024) and transpose lanes <-> slots through shared memory.
//   SLOTS_BREV: the element for index q currently sits in slot brev5(q) (true after dft32<false>).
// On return slot r holds element (r, lane) of the twiddled matrix, natural order.
// tw: [32][32] float2, tw[q*32 + lane] = (cos, -sin)(2 pi lane q / 1024)   (shared memory)
// tile: this warp's private 32x33 float tile.
template <bool SLOTS_BREV>
__device__ __forceinline__ void twiddle_transpose(float (&re)[32], float (&im)[32], float* __restrict__ tile,
                                                  const float2* __restrict__ tw, int lane) {
#pragma unroll
    for (int q = 1; q < 32; ++q) {
        const int s = SLOTS_BREV ? brev5(q) : q;
        const float2 w = tw[q * 32 + lane];
        const float a = re[s], b = im[s];
        re[s] = fmaf(-b, w.y, a * w.x);
        im[s] = fmaf(a, w.y, b * w.x);
    }
    float* row = tile + lane * 33;
#pragma unroll
    for (int q = 0; q < 32; ++q) row[q] = re[SLOTS_BREV ? brev5(q) : q];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; ++r) re[r] = tile[r * 33 + lane];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 32; ++q) row[q] = im[SLOTS_BREV ? brev5(q) : q];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; ++r) im[r] = tile[r * 33 + lane];
    __syncwarp();
}

// Forward 1024-point FFT: input point lane+32r in slot r, output X[lane + 32 q] in slot brev5(q).
// Both radix-32 passes run through ONE copy of the butterfly network (a 2-trip runtime loop): the
// kernels are instruction-fetch sensitive (each warp walks its own instruction stream), so code
// size matters more than the two saved branches.
__device__ __forceinline__ void warp_fft1024(float (&re)[32], float (&im)[32], float* __restrict__ tile,
                                             const float2* __restrict__ tw, int lane) {
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        dft32<false>(re, im);                                         // natural slots -> brev slots
        if (pass == 0) twiddle_transpose<true>(re, im, tile, tw, lane);   // brev slots -> natural slots
    }
}


// ---- variant with the inter-pass twiddles applied on the packed pairs -----------------------------------------
// After dft32_packed_core, register pair i holds slots (i, i+16) = frequencies (brev5(i), brev5(i) + 1).  The table
// tw4[i*32 + lane] = (cos_lo, cos_hi, -sin_lo, -sin_hi) of those two frequencies' twiddles exp(-2 pi i lane q / 1024)
// feeds packed multiplies: 5 packed instructions per pair instead of 8 scalar ones, 16 LDS.128 instead of 31 LDS.64.
struct __align__(16) tw4_t { f2 wx, wy; };

__device__ __forceinline__ void build_tw4(tw4_t* __restrict__ s_tw4, const float2* __restrict__ tw_global, int tid, int nthreads) {
    for (int j = tid; j < 16 * 32; j += nthreads) {
        const int i = j >> 5, l = j & 31;
        const float2 a = tw_global[brev5(i) * 32 + l], b = tw_global[brev5(i + 16) * 32 + l];
        tw4_t t;
        t.wx = f2_pack(a.x, b.x);
        t.wy = f2_pack(a.y, b.y);
        s_tw4[j] = t;
    }
}

__device__ __forceinline__ void warp_fft1024_ptw(float (&re)[32], float (&im)[32], float* __restrict__ tile,
                                                 const tw4_t* __restrict__ tw4, int lane) {
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        f2 Pr[16], Pi[16];
        dft32_packed_core(re, im, Pr, Pi);
        if (pass == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const tw4_t w = tw4[i * 32 + lane];
                const f2 a = Pr[i], b = Pi[i];
                Pr[i] = f2_sub(f2_mul(a, w.wx), f2_mul(b, w.wy));
                Pi[i] = f2_fma(a, w.wy, f2_mul(b, w.wx));
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            f2_unpack(Pr[i], re[i], re[i + 16]);
            f2_unpack(Pi[i], im[i], im[i + 16]);
        }
        if (pass == 0) {                       // brev slots -> natural slots through the 32x33 tile
            float* row = tile + lane * 33;
#pragma unroll
            for (int q = 0; q < 32; ++q) row[q] = re[brev5(q)];
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 32; ++r) re[r] = tile[r * 33 + lane];
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 32; ++q) row[q] = im[brev5(q)];
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 32; ++r) im[r] = tile[r * 33 + lane];
            __syncwarp();
        }
    }
}

// =============================================================================================================
// Dual transform: ONE warp runs TWO independent 1024-point FFTs, FFT A in the low half and FFT B in the high half
// of every 64-bit register pair.  Every floating-point instruction of the butterfly network, the inter-pass
// twiddles and the exchange is then a packed f32x2 / 64-bit one: the instruction count per transform drops by a
// third against the single-transform network above (which can pack only stages 2..5), the two dependency chains
// interleave, and values are born packed (64-bit loads, packed arithmetic), so no pack / unpack moves appear.
// Constants need no duplication: the packed instructions take a scalar register broadcast to both halves (SASS
// "FFMA2 R, R.F32x2, R.F32, R.F32"), so the single-transform twiddle table serves; the exchange tile holds 32 x 33
// 64-bit elements.
// =============================================================================================================
constexpr int kExchDual = 32 * 33;              // f2 elements of one warp's dual exchange tile (8448 bytes)

// natural slots in -> frequency q in slot brev5(q)
__device__ __forceinline__ void dft32_dual(f2 (&re)[32], f2 (&im)[32]) {
    constexpr float kH = 0.70710678118654752440f;
#pragma unroll
    for (int len = 32; len >= 2; len >>= 1) {
        const int half = len >> 1, step = 32 / len;
#pragma unroll
        for (int blk = 0; blk < 32; blk += len) {
#pragma unroll
            for (int i = 0; i < half; ++i) {
                const int ia = blk + i, ib = blk + i + half, m = i * step;
                const f2 ar = re[ia], ai = im[ia], br = re[ib], bi = im[ib];
                re[ia] = f2_add(ar, br);
                im[ia] = f2_add(ai, bi);
                if (m == 0) {
                    re[ib] = f2_sub(ar, br);
                    im[ib] = f2_sub(ai, bi);
                } else if (m == 8) {                           // * (-i): (di, -dr)
                    re[ib] = f2_sub(ai, bi);
                    im[ib] = f2_sub(br, ar);
                } else if (m == 4) {
                    const f2 dr = f2_sub(ar, br), di = f2_sub(ai, bi);
                    re[ib] = f2_mul_s(f2_add(dr, di), kH);
                    im[ib] = f2_mul_s(f2_sub(di, dr), kH);
                } else if (m == 12) {
                    const f2 dr = f2_sub(ar, br), di = f2_sub(ai, bi);
                    re[ib] = f2_mul_s(f2_sub(di, dr), kH);
                    im[ib] = f2_mul_s(f2_add(dr, di), -kH);
                } else {
                    const f2 dr = f2_sub(ar, br), di = f2_sub(ai, bi);
                    const float c = cos32(m), s = sin32(m);
                    re[ib] = f2_fma_s(di, s, f2_mul_s(dr, c));
                    im[ib] = f2_fma_s(dr, -s, f2_mul_s(di, c));
                }
            }
        }
    }
}

// Forward dual FFT: input point lane + 32 r in slot r, output X[lane + 32 q] in slot brev5(q).  One code instance
// serves both radix-32 passes (2-trip runtime loop), as in warp_fft1024.
__device__ __forceinline__ void warp_fft1024_dual(f2 (&re)[32], f2 (&im)[32], f2* __restrict__ tile,
                                                  const float2* __restrict__ tw, int lane) {
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        dft32_dual(re, im);
        if (pass == 0) {
#pragma unroll
            for (int q = 1; q < 32; ++q) {
                const int s = brev5(q);
                const float2 w = tw[q * 32 + lane];
                const f2 a = re[s], b = im[s];
                re[s] = f2_fma_s(b, -w.y, f2_mul_s(a, w.x));          // same operation order as twiddle_transpose: the dual and
                im[s] = f2_fma_s(a, w.y, f2_mul_s(b, w.x));           // single-unit kernels give bit-identical results
            }
            f2* row = tile + lane * 33;
#pragma unroll
            for (int q = 0; q < 32; ++q) row[q] = re[brev5(q)];
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 32; ++r) re[r] = tile[r * 33 + lane];
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 32; ++q) row[q] = im[brev5(q)];
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 32; ++r) im[r] = tile[r * 33 + lane];
            __syncwarp();
        }
    }
}

}  // namespace b200
