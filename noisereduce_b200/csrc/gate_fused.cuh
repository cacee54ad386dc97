// gate_fused.cuh -- single-pass stationary gate (n_fft = 1024): each STFT frame is forward-transformed ONCE.
//
// The two-pass path (k1_analyze -> k_rowfloor -> k_smooth -> k2_synthesize) transforms every frame twice
// because the smoothed mask of frame t needs the decisions of frames t-nt .. t+nt.  Here one warp owns a
// run of output hops and walks it in two phases:
//
//   phase A  frames [tA, tEnd]: load -> FFT -> mask decisions (same FP32 compare + FP64 re-decision as
//            k1_analyze) -> the packed spectrum Z (8 KB per frame pair) and the 68-byte decision row go to a
//            per-warp scratch in global memory (written once, read once; L2 / HBM, not shared memory);
//   phase B  frames [t_start, he): in-warp mask smoothing (time recurrence on byte lanes, frequency taps by
//            dp4a over a per-warp shared row -- the arithmetic of k_smooth_packed), reload Z, apply,
//            inverse FFT, register overlap-add, store (the arithmetic of k2_synthesize).
//
// The extra traffic (16 KB per frame pair) is affordable because the kernels are instruction-bound with HBM
// mostly idle; what disappears is k1's FFT of every frame, the k_smooth pass and the uint16 mask array.
//
// top_db row floor (spectralgate/utils.py:16): a row is lifted only when some |X[f,t]| reaches
// 10^((thresh[f]+top_db)/20), and |X[f,t]| <= max|x| (the analysis window sums to 1).  The kernel tracks
// max|x|; if it could reach the smallest floor the host discards the result and runs the exact two-pass
// path instead (practically never: the floor sits 80 dB above the noise threshold).
//
// Mask decisions and the integer smoothed masks are identical to the two-pass path; the waveform agrees to
// FP32 rounding (frames may be paired differently inside the complex FFT).
#pragma once
#include "gate_kernels.cuh"

namespace b200 {

__device__ __forceinline__ unsigned ld_cg(const unsigned* p) {
#ifdef B200_CUSIM_BUILD
    return *p;
#else
    return __ldcg(p);
#endif
}

struct KFArgs {
    Geom g;
    Tables tb;
    const void* x;
    void* y;
    float2* zscratch;          // [workers][max_pairs][32 slots][32 lanes]
    unsigned* bitscratch;      // [workers][max_rows][kFusedRowWords]
    int max_pairs, max_rows;
    unsigned* maxabs;          // max |x| seen, as float bits
    Counters* cnt;
    float pD, one_minus_p;
    int nf, nt;
    unsigned taps[3];          // 2 nf + 1 <= 12 triangle taps as packed bytes
    int run, n_runs;           // output hops per work item
    DebugTap dbg;
    unsigned* dbg_bits;        // [T][kFW] raw decisions of the tapped unit (or null)
};

constexpr int kFusedRowWords = kFW + 1;            // decision row padded so a 20-bit field may straddle
constexpr int kFusedCRow = 672;                    // bytes: 32 halo + 560 data + zero tail (multiple of 4)
constexpr int kfused_smem_floats() {
    return 2 * kN + 2 * kN + 256 + 3 * kFPad + kWarps * (kExchFloats + 2 * kFW + 2 * kFusedCRow / 4) + 8;
}

template <int HR, typename T>
__global__ void __launch_bounds__(kThreads, 3) k_fused(const KFArgs a) {
    constexpr int NH = 32 / HR;
    B200_DYN_SMEM(float, smem);
    const Geom& g = a.g;
    const int H = g.H;
    float* s_wa = smem;
    float* s_ws = smem + kN;
    float2* s_tw = reinterpret_cast<float2*>(smem + 2 * kN);
    float* s_invn = smem + 4 * kN;
    float* s_ef = s_invn + 256;
    float* s_thr4 = s_ef + kFPad;
    float* s_gco = s_thr4 + kFPad;
    float* s_warp = s_gco + kFPad;
    for (int i = threadIdx.x; i < kN; i += kThreads) {
        s_wa[i] = a.tb.wa[i];
        s_ws[i] = a.tb.ws[i];
        s_tw[i] = a.tb.tw[i];
    }
    for (int i = threadIdx.x; i < H; i += kThreads) s_invn[i] = a.tb.invn[i];
    for (int i = threadIdx.x; i < kFPad; i += kThreads) {
        s_ef[i] = a.tb.ef[i];
        s_thr4[i] = a.tb.thr4[i];
        s_gco[i] = a.tb.gco[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kPerWarp = kExchFloats + 2 * kFW + 2 * kFusedCRow / 4;
    float* tile = s_warp + warp * kPerWarp;
    unsigned* s_amb = reinterpret_cast<unsigned*>(tile + kExchFloats);
    unsigned* s_crow = s_amb + 2 * kFW;                       // two rows of kFusedCRow bytes (frames a, b)
    for (int i = lane; i < 2 * kFusedCRow / 4; i += 32) s_crow[i] = 0u;
    __syncthreads();

    const int pl = (32 - lane) & 31;
    const long long n_items = (long long)g.n_units * a.n_runs;
    const bool blend = (a.one_minus_p != 0.f);
    const int nt = a.nt, nf = a.nf, aa = nt + 1;
    const int ntE = (nt + 1) & ~1;                            // halo rounded up to whole frame pairs
    const long long worker = (long long)blockIdx.x * kWarps + warp;
    float2* zs = a.zscratch + worker * a.max_pairs * 1024;
    unsigned* brows = a.bitscratch + worker * a.max_rows * kFusedRowWords;
    float mxabs = 0.f;

    for (long long item = worker; item < n_items; item += (long long)gridDim.x * kWarps) {
        const int ul = (int)(item / a.n_runs);
        const int run = (int)(item - (long long)ul * a.n_runs);
        const int u = g.u0 + ul;
        const int ic = u / g.C, c = u - ic * g.C;
        const long long i1 = (long long)ic * g.step - g.pad;
        long long out_len = g.n_total - (long long)ic * g.step;
        if (out_len > g.step) out_len = g.step;
        long long jp_hi = g.pad + out_len;
        const long long sig_len = (long long)(g.T - 1) * H;
        if (jp_hi > sig_len) jp_hi = sig_len;
        if (jp_hi <= g.pad) continue;
        const long long jlo = g.pad + kN / 2, jhi = jp_hi + kN / 2;
        const int h_lo = (int)(jlo / H), h_hi = (int)((jhi + H - 1) / H);
        const int hs = h_lo + run * a.run;
        const int he = min(hs + a.run, h_hi);
        if (hs >= he) continue;
        const int t_start = max(0, hs - (NH - 1)) & ~1;      // even, so phase A / phase B pairs coincide
        const int t_last = min(he - 1, g.T - 1);
        const int tA = max(0, t_start - ntE);                 // even
        const int tEnd = min(g.T - 1, t_last + nt);
        const T* xrow = static_cast<const T*>(a.x) + (long long)c * g.in_stride;
        T* yrow = static_cast<T*>(a.y) + (long long)c * g.out_stride;
        const bool tap = (a.dbg.ul == ul);

        // ------------------------------------------------------------------ phase A: analysis
        for (int t = tA; t <= tEnd; t += 2) {
            const bool vb = (t + 1 <= tEnd);
            const long long base = (long long)t * H - kN / 2;
            float re[32], im[32];
            float e;
            {
                float xr[32 + HR];
                if (pair_window_interior<HR>(base, i1, g.Lp, g.n_total)) {
                    const T* p = xrow + i1 + base + lane;
#pragma unroll
                    for (int r = 0; r < 32 + HR; ++r) xr[r] = ld_sample(p + 32 * r);
                } else {
#pragma unroll
                    for (int r = 0; r < 32 + HR; ++r) xr[r] = chunk_sample(xrow, base + lane + 32 * r, i1, g.Lp, g.n_total);
                }
                e = 0.f;
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    const float w = s_wa[lane + 32 * r];
                    re[r] = xr[r] * w;
                    im[r] = xr[r + HR] * w;
                    e = fmaf(re[r], re[r], e);
                }
#pragma unroll
                for (int r = 0; r < 32 + HR; ++r) mxabs = fmaxf(mxabs, fabsf(xr[r]));
                if (vb) {
#pragma unroll
                    for (int r = 0; r < 32; ++r) e = fmaf(im[r], im[r], e);
                } else {
#pragma unroll
                    for (int r = 0; r < 32; ++r) im[r] = 0.f;
                }
            }
            const float S = sqrtf(warp_sum(e));
            warp_fft1024(re, im, tile, s_tw, lane);

            // keep the packed spectrum for phase B
            {
                float2* zp = zs + (long long)((t - tA) >> 1) * 1024 + lane;
#pragma unroll
                for (int q = 0; q < 32; ++q) zp[32 * q] = make_float2(re[brev5(q)], im[brev5(q)]);
            }
            unsigned wordA = 0u, wordB = 0u;
            unsigned anyamb = 0u;
#pragma unroll
            for (int q = 0; q < kFW; ++q) {
                const int sA = brev5(q), sP = brev5(31 - q), s0 = brev5((32 - q) & 31);
                const float zr = re[sA], zi = im[sA];
                float pr = __shfl_sync(0xffffffffu, re[sP], pl);
                float pi = __shfl_sync(0xffffffffu, im[sP], pl);
                if (lane == 0) { pr = re[s0]; pi = im[s0]; }
                const float ar = zr + pr, ai = zi - pi;
                const float br = zi + pi, bi = pr - zr;
                const float PA = fmaf(ar, ar, ai * ai);
                const float PB = fmaf(br, br, bi * bi);
                const int k = lane + 32 * q;
                const bool valid = (q < 16) || (lane == 0);
                const float th = s_thr4[k];
                const float gg = fmaf(s_gco[k], S, th * 8.0e-7f);
                const float dA = PA - th, dB = PB - th;
                const unsigned wA = __ballot_sync(0xffffffffu, valid && (dA > 0.f));
                const unsigned wB = __ballot_sync(0xffffffffu, valid && vb && (dB > 0.f));
                const bool amA = valid && fabsf(dA) <= gg, amB = valid && vb && fabsf(dB) <= gg;
                if (lane == q) { wordA = wA; wordB = wB; }
                if (__any_sync(0xffffffffu, amA || amB)) {
                    const unsigned mA = __ballot_sync(0xffffffffu, amA);
                    const unsigned mB = __ballot_sync(0xffffffffu, amB);
                    anyamb |= 1u << q;
                    if (lane == 0) { s_amb[2 * q] = mA; s_amb[2 * q + 1] = mB; }
                }
                if (tap && valid && k < kF) {
                    float* sp = a.dbg.spec + ((long long)t * kF + k) * 2;
                    sp[0] = 0.5f * ar; sp[1] = 0.5f * ai;
                    if (vb) { sp[2 * kF] = 0.5f * br; sp[2 * kF + 1] = 0.5f * bi; }
                }
            }
            if (anyamb) {                             // warp-uniform, rare: redo those bins in float64
                __syncwarp();
                unsigned nre = 0, nun = 0;
                for (int q = 0; q < kFW; ++q) {
                    if (!((anyamb >> q) & 1u)) continue;
                    for (int fr = 0; fr < 2; ++fr) {
                        unsigned m = s_amb[2 * q + fr];
                        while (m) {
                            const int src = __ffs((int)m) - 1;
                            m &= m - 1;
                            const int r = recheck_bin_fp64(xrow, base + (long long)fr * H, i1, g.Lp, g.n_total,
                                                           src + 32 * q, a.tb, lane);
                            ++nre;
                            if (r == 0) { ++nun; continue; }
                            if (lane == q) {
                                unsigned& wd = fr ? wordB : wordA;
                                wd = (r == 2) ? (wd | (1u << src)) : (wd & ~(1u << src));
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    atomicAdd(&a.cnt->rechecked, (unsigned long long)nre);
                    if (nun) atomicAdd(&a.cnt->unresolved, (unsigned long long)nun);
                }
            }
            {
                unsigned* br0 = brows + (long long)(t - tA) * kFusedRowWords;
                if (lane < kFusedRowWords) {
                    br0[lane] = (lane < kFW) ? wordA : 0u;
                    br0[kFusedRowWords + lane] = (lane < kFW && vb) ? wordB : 0u;
                }
                if (tap && a.dbg_bits && lane < kFW) {
                    a.dbg_bits[(long long)t * kFW + lane] = wordA;
                    if (vb) a.dbg_bits[(long long)(t + 1) * kFW + lane] = wordB;
                }
            }
        }
        __syncwarp();
#ifndef B200_CUSIM_BUILD
        __threadfence_block();                        // this warp's scratch stores before its own reloads
#endif

        // ------------------------------------------------------------------ phase B: smoothing + synthesis
        // decision nibble field of this lane's 20 bins in frame tau (0 outside the stream / the chunk)
        const int fld_w = (20 * lane) >> 5, fld_s = (20 * lane) & 31;
        const int tau_s = max(tA, t_start - nt);      // the recurrence treats everything before tau_s as zero
        auto field = [&](int tau) -> unsigned {
            if (lane >= 28 || tau < tau_s || tau > tEnd) return 0u;
            const unsigned* r = brows + (long long)(tau - tA) * kFusedRowWords + fld_w;
            return __funnelshift_r(ld_cg(r), ld_cg(r + 1), fld_s) & 0xFFFFFu;   // rows were written by other lanes
        };
        unsigned d1b[5], s2[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) { d1b[j] = 0x10101010u; s2[j] = 0u; }
        auto advance = [&](int tau) {
            const unsigned fa = field(tau), fb = field(tau - aa), fc = field(tau - 2 * aa);
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const unsigned ea = (((fa >> (4 * j)) & 0xFu) * 0x00204081u) & 0x01010101u;
                const unsigned eb = (((fb >> (4 * j)) & 0xFu) * 0x00204081u) & 0x01010101u;
                const unsigned ec = (((fc >> (4 * j)) & 0xFu) * 0x00204081u) & 0x01010101u;
                d1b[j] = d1b[j] + ea + ec - 2u * eb;
                s2[j] = s2[j] + d1b[j] - 0x10101010u;
            }
        };
        int tau_next = t_start - nt;                  // first frame fed (frames before it count as zero)
        if (tau_next < tau_s) tau_next = tau_s;       // (negative frames do not exist: same thing)
        // the stream is exact from tau_next + 2 nt = t_start + nt on, i.e. for every output frame >= t_start

        float acc[32 + HR];
#pragma unroll
        for (int r = 0; r < 32 + HR; ++r) acc[r] = 0.f;

        for (int t = t_start; t < he; t += 2) {
            const bool va = (t <= t_last), vb = (t + 1 <= t_last);
            if (va) {
                // ---- masks of frames t and t+1
                unsigned nmA[kFW], nmB[kFW];
#pragma unroll 1
                for (int fr = 0; fr < 2; ++fr) {
                    const int tt = t + fr;
                    for (; tau_next <= tt + nt; ++tau_next) advance(tau_next);
                    unsigned* crow = s_crow + fr * (kFusedCRow / 4);
                    if (lane < 28) {
#pragma unroll
                        for (int j = 0; j < 5; ++j) crow[8 + 5 * lane + j] = s2[j];       // bytes 32 + 20*lane ..
                    }
                }
                __syncwarp();
#pragma unroll
                for (int q = 0; q < kFW; ++q) {
                    const int k = lane + 32 * q;
                    const int b0 = 32 + k - nf;                          // first byte of the tap window
                    const int w0 = b0 >> 2, sh = (b0 & 3) * 8;
                    const unsigned* ca = s_crow + w0;
                    const unsigned* cb = ca + kFusedCRow / 4;
                    unsigned accA = 0u, accB = 0u;
#pragma unroll
                    for (int m = 0; m < 3; ++m) {
                        accA = dp4a_u(__funnelshift_r(ca[m], ca[m + 1], sh), a.taps[m], accA);
                        accB = dp4a_u(__funnelshift_r(cb[m], cb[m + 1], sh), a.taps[m], accB);
                    }
                    nmA[q] = accA;
                    nmB[q] = accB;
                }
                __syncwarp();
                float eta = 0.f, etb = 0.f;
                if (blend) {
                    eta = a.one_minus_p * time_edge(t, g.T, a.nt);
                    etb = a.one_minus_p * time_edge(t + 1, g.T, a.nt);
                }
                if (tap) {
#pragma unroll
                    for (int q = 0; q < kFW; ++q) {
                        const int k = lane + 32 * q;
                        if (k < kF && ((q < 16) || lane == 0)) {
                            a.dbg.mask[(long long)t * kF + k] = fmaf((float)nmA[q], a.pD, eta * s_ef[k]);
                            if (vb) a.dbg.mask[(long long)(t + 1) * kF + k] = fmaf((float)nmB[q], a.pD, etb * s_ef[k]);
                        }
                    }
                }
                // ---- reload the spectrum, apply, inverse FFT
                float re[32], im[32];
                {
                    const float2* zp = zs + (long long)((t - tA) >> 1) * 1024 + lane;
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const float2 v = zp[32 * q];
                        re[brev5(q)] = v.x;
                        im[brev5(q)] = v.y;
                    }
                }
#pragma unroll
                for (int q = 0; q < kFW; ++q) {
                    const int sA = brev5(q), sP = brev5(31 - q), s0 = brev5((32 - q) & 31);
                    const int k = lane + 32 * q;
                    float ma = fmaf((float)nmA[q], a.pD, eta * s_ef[k]);
                    float mb = fmaf((float)nmB[q], a.pD, etb * s_ef[k]);
                    if (!vb) mb = 0.f;
                    const float s = 0.5f * (ma + mb), d = 0.5f * (ma - mb);
                    const float zr = re[sA], zi = im[sA];
                    if (q < 16) {
                        float pr = __shfl_sync(0xffffffffu, re[sP], pl);
                        float pi = __shfl_sync(0xffffffffu, im[sP], pl);
                        if (lane == 0) { pr = re[s0]; pi = im[s0]; }
                        const float own_r = fmaf(d, pr, s * zr), own_i = fmaf(-d, pi, s * zi);
                        const float oth_r = fmaf(d, zr, s * pr), oth_i = fmaf(-d, zi, s * pi);
                        const float nr = __shfl_sync(0xffffffffu, oth_r, pl);
                        const float ni = __shfl_sync(0xffffffffu, oth_i, pl);
                        re[sA] = own_r;
                        im[sA] = own_i;
                        if (lane != 0) { re[sP] = nr; im[sP] = ni; }
                        else if (q != 0) { re[s0] = oth_r; im[s0] = oth_i; }
                    } else if (lane == 0) {
                        re[sA] = fmaf(d, zr, s * zr);
                        im[sA] = fmaf(-d, zi, s * zi);
                    }
                }
#pragma unroll
                for (int q = 0; q < 32; ++q) {                           // brev -> natural slots, re <-> im
                    const int b = brev5(q);
                    if (b == q) {
                        const float tr = re[q];
                        re[q] = im[q];
                        im[q] = tr;
                    } else if (q < b) {
                        const float t1 = re[q], t2 = im[q];
                        re[q] = im[b];
                        im[q] = re[b];
                        re[b] = t2;
                        im[b] = t1;
                    }
                }
                warp_fft1024(re, im, tile, s_tw, lane);
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const float w = s_ws[lane + 32 * q];
                    acc[q] = fmaf(im[brev5(q)], w, acc[q]);
                    acc[q + HR] = fmaf(re[brev5(q)], w, acc[q + HR]);
                }
            }
            // ---- hops t and t+1 are complete
            {
                const long long jp0 = (long long)t * H - kN / 2;
                if (t >= hs && t + 1 < he && t >= NH - 1 && t + 1 <= g.T - 1 && jp0 >= g.pad && jp0 + 2 * H <= jp_hi) {
                    T* dst = yrow + i1 + jp0 + lane;
#pragma unroll
                    for (int r = 0; r < 2 * HR; ++r) dst[32 * r] = st_cast<T>(acc[r] * s_invn[(r % HR) * 32 + lane]);
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
                        const long long jp = (long long)hop * H + ro - kN / 2;
                        if (jp < g.pad || jp >= jp_hi) continue;
                        float inv;
                        if (hop >= NH - 1 && hop <= g.T - 1) {
                            inv = s_invn[ro];
                        } else {
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
        __syncwarp();
    }
    mxabs = fmaxf(mxabs, __shfl_xor_sync(0xffffffffu, mxabs, 16));
    mxabs = fmaxf(mxabs, __shfl_xor_sync(0xffffffffu, mxabs, 8));
    mxabs = fmaxf(mxabs, __shfl_xor_sync(0xffffffffu, mxabs, 4));
    mxabs = fmaxf(mxabs, __shfl_xor_sync(0xffffffffu, mxabs, 2));
    mxabs = fmaxf(mxabs, __shfl_xor_sync(0xffffffffu, mxabs, 1));
    if (lane == 0) atomicMax(a.maxabs, __float_as_uint(mxabs));
}

}  // namespace b200
