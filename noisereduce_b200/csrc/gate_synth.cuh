// gate_synth.cuh -- k2c_synthesize: the synthesis half of the gate when the analysis kernel kept the spectra.
//
// Reference semantics (paths relative to /root/reference): mask apply  X * mask  at
// noisereduce/spectralgate/stationary.py:117 / nonstationary.py:87, scipy.signal.istft as called at
// stationary.py:120-126 (window, overlap-add, sum-of-squares normalisation, centre crop), chunk centre
// written back at base.py:150.
//
// One warp owns a run of output hops of one (chunk, channel) unit and walks its frame pairs.  Per pair:
//   1. the pair's packed spectrum Z = X_a + i X_b (8 KB, written by k1 / k1n) and its two mask rows arrive in
//      this warp's shared-memory buffers by 1-D bulk asynchronous copies (cp.async.bulk -> TMA engine, SASS
//      UBLKCP) signalled on a per-warp mbarrier; the copies of pair t+2 are issued as soon as pair t has been
//      read, so they travel behind a whole inverse transform;
//   2. masks are applied on the packed spectrum,  Z'[k] = s Z[k] + d conj(Z[N-k]),  s, d = (m_a +- m_b)/2.
//      The mirrored bin is read from the same shared buffer by its FLAT index (N - k) & (N - 1), which is what
//      makes lane 0 (bins 0, 32, ... whose mirrors live in other SLOTS of the same lane) an ordinary lane: no
//      shuffles, no per-lane selects.  A lane computes both members of its 16 lower pairs; the upper members
//      travel to their owners through the (idle) FFT exchange tile;
//   3. ONE inverse 1024-point FFT (ifft(z) = swap(fft(swap z)): the apply step writes real/imaginary parts
//      exchanged), synthesis window, overlap-add in registers, normalise, store the chunk centre.
#pragma once
#include "gate_kernels.cuh"

namespace b200 {

// ---- mbarrier + bulk asynchronous copy (global -> shared through the TMA engine) ------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
#ifdef B200_CUSIM_BUILD
    *bar = 0; (void)count;
#else
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
#endif
}
__device__ __forceinline__ void mbar_init_fence() {
#ifndef B200_CUSIM_BUILD
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
#ifdef B200_CUSIM_BUILD
    (void)bar; (void)bytes;
#else
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
#endif
}
// dst, src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
#ifdef B200_CUSIM_BUILD
    memcpy(smem_dst, gmem_src, bytes); (void)bar;
#else
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gmem_src), "r"(bytes), "r"(b) : "memory");
#endif
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
#ifdef B200_CUSIM_BUILD
    (void)bar; (void)parity;
    __syncwarp();                  // simulator: the issuing lane's (synchronous) copy precedes its own wait
#else
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}"
        ::"r"(a), "r"(parity) : "memory");
#endif
}
// order this thread's earlier generic-proxy accesses of shared memory before later async-proxy (bulk copy) writes
__device__ __forceinline__ void fence_proxy_async_smem() {
#ifndef B200_CUSIM_BUILD
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
// shared -> global bulk copy (TMA store engine) of `bytes` (16-byte multiple, both sides 16-byte aligned); completion is
// tracked by bulk groups: commit, then wait_group.read before the shared source is overwritten
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes) {
#ifdef B200_CUSIM_BUILD
    memcpy(gmem_dst, smem_src, bytes);
#else
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_src);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(s), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void bulk_wait_read_all() {
#ifndef B200_CUSIM_BUILD
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel) {
#ifdef B200_CUSIM_BUILD
    const unsigned long long v = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFFu) << (8 * i);
    return r;
#else
    return __byte_perm(a, b, sel);
#endif
}

enum { kMaskU16 = 0, kMaskU16Blend = 1, kMaskF32 = 2 };
#ifndef B200_K2C_PTW
#define B200_K2C_PTW 0              // inter-pass twiddles on packed register pairs (warp_fft1024_ptw)
#endif


constexpr int kK2cTableBytes = kN * 4 + kN * 8 + kFPad * 4;          // synthesis window, twiddles, frequency edge factors
constexpr int kK2cTileBytes = kExchFloats * 4;                        // 4224: FFT exchange tile, also the 513 x float2 mirror buffer
template <int MK> __host__ __device__ constexpr int k2c_mask_bytes() { return MK == kMaskF32 ? 2 * kFPad * 4 : 2 * kFPad * 2; }
template <int MK> __host__ __device__ constexpr int k2c_warp_bytes() { return kN * 8 + k2c_mask_bytes<MK>() + kK2cTileBytes + 16; }
template <int MK> __host__ __device__ constexpr int k2c_smem_bytes() { return kK2cTableBytes + kWarps * k2c_warp_bytes<MK>(); }
static_assert(kK2cTileBytes >= 513 * 8, "mirror buffer must fit the exchange tile");

#ifndef B200_K2C_MINBLOCKS
#define B200_K2C_MINBLOCKS 3
#endif
template <int HR, int MK, typename T>
__global__ void __launch_bounds__(kThreads, B200_K2C_MINBLOCKS) k2c_synthesize(const K2Args a) {
    constexpr int NH = 32 / HR;             // frames overlapping one hop (win / hop)
    constexpr int MB = k2c_mask_bytes<MK>();
    constexpr int WB = k2c_warp_bytes<MK>();
    B200_DYN_SMEM(unsigned char, smraw);
    const Geom& g = a.g;
    const int H = g.H;
    float* s_ws = reinterpret_cast<float*>(smraw);
#if B200_K2C_PTW
    tw4_t* s_tw4 = reinterpret_cast<tw4_t*>(smraw + kN * 4);
#else
    float2* s_tw = reinterpret_cast<float2*>(smraw + kN * 4);
#endif
    float* s_ef = reinterpret_cast<float*>(smraw + kN * 12);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbase = smraw + kK2cTableBytes + warp * WB;
    float2* zbuf = reinterpret_cast<float2*>(wbase);
    unsigned char* mbuf = wbase + kN * 8;
    float* tile = reinterpret_cast<float*>(wbase + kN * 8 + MB);
    float2* qbuf = reinterpret_cast<float2*>(tile);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(wbase + kN * 8 + MB + kK2cTileBytes);
    for (int i = threadIdx.x; i < kN; i += kThreads) {
        s_ws[i] = a.tb.ws[i];
#if !B200_K2C_PTW
        s_tw[i] = a.tb.tw[i];
#endif
    }
#if B200_K2C_PTW
    build_tw4(s_tw4, a.tb.tw, threadIdx.x, kThreads);
#endif
    for (int i = threadIdx.x; i < kFPad; i += kThreads) s_ef[i] = a.tb.ef[i];
    if (lane == 0) mbar_init(bar, 1);
    mbar_init_fence();
    __syncthreads();
    unsigned phase = 0;

    float invn[HR];                          // interior overlap-add norm of this lane's HR rows of a hop
#pragma unroll
    for (int r = 0; r < HR; ++r) invn[r] = a.tb.invn[r * 32 + lane];

    const long long n_items = (long long)g.n_units * a.n_runs;
    for (long long item = (long long)blockIdx.x * kWarps + warp; item < n_items;
         item += (long long)gridDim.x * kWarps) {
        const int ul = (int)(item / a.n_runs);
        const int run = (int)(item - (long long)ul * a.n_runs);
        const int u = g.u0 + ul;
        const int ic = u / g.C, c = u - ic * g.C;
        const long long i1 = (long long)ic * g.step - g.pad;
        long long out_len = g.n_total - (long long)ic * g.step;
        if (out_len > g.step) out_len = g.step;
        long long jp_hi = g.pad + out_len;
        const long long sig_len = (long long)(g.T - 1) * H;          // istft output length
        if (jp_hi > sig_len) jp_hi = sig_len;                        // beyond it the reference leaves zeros
        if (jp_hi <= g.pad) continue;
        const long long jlo = g.pad + kN / 2, jhi = jp_hi + kN / 2;
        const int h_lo = (int)(jlo / H), h_hi = (int)((jhi + H - 1) / H);
        const int hs = h_lo + run * a.run;
        const int he = min(hs + a.run, h_hi);
        if (hs >= he) continue;
        const int t_start = max(0, hs - (NH - 1)) & ~1;              // k1 packed frames (2j, 2j+1): walk the same pairs
        const int t_last = min(he - 1, g.T - 1);
        T* yrow = static_cast<T*>(a.y) + (long long)c * g.out_stride;
        const float2* zunit = a.zcache + (long long)ul * a.zpairs * 1024;
        const unsigned short* munit = (MK == kMaskF32) ? nullptr : a.num + (long long)ul * a.zpairs * (2 * kFPad);
        const float* funit = (MK == kMaskF32) ? a.fmask + (long long)ul * g.T * kFPad : nullptr;

        float acc[32 + HR];
#pragma unroll
        for (int r = 0; r < 32 + HR; ++r) acc[r] = 0.f;

        // request pair (tt, tt+1): spectrum + mask rows.  All lanes have finished reading the buffers.
        auto stage = [&](int tt) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                const bool vb2 = tt + 1 <= t_last;
                const unsigned mbytes = (MK == kMaskF32 && !vb2) ? (unsigned)(kFPad * 4) : (unsigned)MB;
                mbar_expect_tx(bar, (unsigned)(kN * 8) + mbytes);
                bulk_g2s(zbuf, zunit + (long long)(tt >> 1) * 1024, kN * 8, bar);
                if (MK == kMaskF32) bulk_g2s(mbuf, funit + (long long)tt * kFPad, mbytes, bar);
                else bulk_g2s(mbuf, munit + (long long)(tt >> 1) * (2 * kFPad), mbytes, bar);
            }
        };
        if (t_start <= t_last) stage(t_start);

        for (int t = t_start; t < he; t += 2) {
            const bool va = (t <= t_last), vb = (t + 1 <= t_last);
            if (va) {                                   // vb implies va
                float re[32], im[32];                   // what the FFT call sees: re = Im Z', im = Re Z'
                mbar_wait(bar, phase);
                phase ^= 1u;
                // ---- mask scalars of this pair --------------------------------------------------------------
                float pa = 0.f, ca = 0.f, pb = 0.f, cb = 0.f, eta = 0.f, etb = 0.f;
                const float* mA = reinterpret_cast<const float*>(mbuf);
                const float* mB = vb ? mA + kFPad : mA;
                const unsigned* m32 = reinterpret_cast<const unsigned*>(mbuf);
                if (MK == kMaskF32) {
                    pa = 0.5f;
                    pb = vb ? 0.5f : 0.f;
                } else {
                    pa = 0.5f * a.pD;                   // masks are formed already halved: s, d = m_a/2 +- m_b/2
                    ca = -8388608.0f * pa;              // the 2^23 of the integer -> float bit pattern, folded into the FMA
                    pb = vb ? pa : 0.f;
                    cb = vb ? ca : 0.f;
                    if (MK == kMaskU16Blend) {
                        eta = 0.5f * a.one_minus_p * time_edge(t, g.T, a.nt);
                        etb = vb ? 0.5f * a.one_minus_p * time_edge(t + 1, g.T, a.nt) : 0.f;
                    }
                }
                auto half_masks = [&](int kk, float& ma, float& mb) {      // m_a / 2, m_b / 2 at bin kk
                    if (MK == kMaskF32) {
                        ma = pa * mA[kk];
                        mb = pb * mB[kk];
                    } else {
                        const unsigned pk = m32[kk];
                        const float fa = __uint_as_float(prmt(pk, 0x4B00u, 0x5410u));     // 2^23 + n_a
                        const float fb = __uint_as_float(prmt(pk, 0x4B00u, 0x5432u));     // 2^23 + n_b
                        if (MK == kMaskU16Blend) {
                            const float e = s_ef[kk];
                            ma = fmaf(fa - 8388608.0f, pa, eta * e);
                            mb = fmaf(fb - 8388608.0f, pb, etb * e);
                        } else {
                            ma = fmaf(fa, pa, ca);
                            mb = fmaf(fb, pb, cb);
                        }
                    }
                };
                if (a.dbg.ul == ul) {                        // parity tap (tests): the masks this pair applies
#pragma unroll 1
                    for (int k = lane; k < kF; k += 32) {
                        float ma, mb;
                        half_masks(k, ma, mb);
                        a.dbg.mask[(long long)t * kF + k] = 2.0f * ma;
                        if (vb) a.dbg.mask[(long long)(t + 1) * kF + k] = 2.0f * mb;
                    }
                }
                // ---- lower pairs k = lane + 32 q < 512: both members -------------------------------------------
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int k = 32 * q + lane;
                    const float2 own = zbuf[k];
                    const float2 par = zbuf[q == 0 ? ((kN - lane) & (kN - 1)) : (kN - k)];
                    float ma, mb;
                    half_masks(k, ma, mb);
                    const float s = ma + mb, d = ma - mb;
                    // Z'[k] = s Z[k] + d conj(Z[N-k]);  Z'[N-k] = s Z[N-k] + d conj(Z[k])
                    im[q] = fmaf(d, par.x, s * own.x);
                    re[q] = fmaf(-d, par.y, s * own.y);
                    const float qr = fmaf(d, own.x, s * par.x), qi = fmaf(-d, own.y, s * par.y);
                    qbuf[512 - k] = make_float2(qi, qr);                 // slot of bin N - k, stored (re, im) as the FFT wants them
                }
                {   // slot 16: bins 512 + lane, computed directly (lane 0 = the self-mirrored bin N/2)
                    const float2 own = zbuf[512 + lane];
                    const float2 par = zbuf[512 - lane];
                    float ma, mb;
                    half_masks(512 - lane, ma, mb);
                    const float s = ma + mb, d = ma - mb;
                    im[16] = fmaf(d, par.x, s * own.x);
                    re[16] = fmaf(-d, par.y, s * own.y);
                }
                __syncwarp();
#pragma unroll
                for (int q = 17; q < 32; ++q) {
                    const float2 v = qbuf[32 * (q - 16) + lane];
                    re[q] = v.x;
                    im[q] = v.y;
                }
                __syncwarp();
                if (t + 2 <= t_last) stage(t + 2);           // next pair streams in behind this pair's transform
#if B200_K2C_PTW
                warp_fft1024_ptw(re, im, tile, s_tw4, lane);
#else
                warp_fft1024(re, im, tile, s_tw, lane);
#endif
                // now im = N a'[n], re = N b'[n] (n = lane + 32 q) at slot brev5(q)
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const float w = s_ws[lane + 32 * q];
                    acc[q] = fmaf(im[brev5(q)], w, acc[q]);
                    acc[q + HR] = fmaf(re[brev5(q)], w, acc[q + HR]);
                }
            }
            // hops t and t+1 are now complete (all frames <= t+1 have been added)
            {
                const long long jp0 = (long long)t * H - kN / 2;             // chunk-local index of row 0, lane 0
                if (t >= hs && t + 1 < he && t >= NH - 1 && t + 1 <= g.T - 1 && jp0 >= g.pad &&
                    jp0 + 2 * H <= jp_hi) {
                    // both hops interior and fully inside the chunk centre: 2*HR coalesced row stores
                    T* dst = yrow + i1 + jp0 + lane;
#pragma unroll
                    for (int r = 0; r < 2 * HR; ++r) dst[32 * r] = st_cast<T>(acc[r] * invn[r % HR]);
                } else {
#pragma unroll 1
                    for (int r = 0; r < 2 * HR; ++r) {
                        float v = 0.f;
#pragma unroll
                        for (int rr = 0; rr < 2 * HR; ++rr)
                            if (rr == r) v = acc[rr];
                        const int hop = t + r / HR;
                        if (hop < hs || hop >= he) continue;
                        const int ro = (r % HR) * 32 + lane;
                        const long long jp = (long long)hop * H + ro - kN / 2;   // chunk-local output index
                        if (jp < g.pad || jp >= jp_hi) continue;
                        float inv;
                        if (hop >= NH - 1 && hop <= g.T - 1) {
                            inv = a.tb.invn[ro];
                        } else {                                                  // first / last hops
                            float nrm = 0.f;
                            for (int i = 0; i < NH; ++i) {
                                const int tf = hop - i;
                                if (tf >= 0 && tf <= g.T - 1) {
                                    const float w = s_ws[i * H + ro] * a.tb.ws_to_w;
                                    nrm = fmaf(w, w, nrm);
                                }
                            }
                            inv = nrm > 1e-10f ? 1.0f / nrm : 1.0f;
                        }
                        yrow[i1 + jp] = st_cast<T>(v * inv);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 32 - HR; ++r) acc[r] = acc[r + 2 * HR];
#pragma unroll
            for (int r = 32 - HR; r < 32 + HR; ++r) acc[r] = 0.f;
        }
    }
}

}  // namespace b200
