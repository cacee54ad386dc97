// gate_dual_2k.cuh -- k1nd_magnitude_2k: analysis of the n_fft = 2048 family (BASELINE.json config 3) with TWO frames
// per warp -- frames t and t+1 of the same (chunk, channel) unit, frame t in the low half and frame t+1 in the high
// half of every 64-bit register pair (warp_fft.cuh, "Dual transform"; gate_dual.cuh does the same with two channels).
// Both frames share the control flow, so the window, the butterfly network, the inter-pass twiddles and the
// half-length post-processing are packed f32x2 instructions, and the mirrored element Z[1024-k] is read from shared
// memory by its flat index instead of being shuffled: about three quarters of the instructions per frame of
// k1n_magnitude_2k.
//
// Reference semantics (paths relative to /root/reference): scipy.signal.stft as called at
// noisereduce/spectralgate/nonstationary.py:58-64, abs at :65.  Outputs are those of k1n_magnitude_2k: |X| rows and
// (for the frames the synthesis kernel will need) the half-length spectra Z.
#pragma once
#include "gate_dual.cuh"
#include "gate_kernels_2k.cuh"

namespace b200 {

struct K1nd2Args {
    Geom g;
    Tables2 tb;
    const float* x;
    float* mag;                // [n_units][T][FPad2]
    float2* zcache;            // [n_units][T][1024] half-length spectra of frames [z_lo, z_hi), or null
    int z_lo, z_hi;
    DebugTap dbg;              // spec: [T][F2][2]
    int run, n_runs;           // frames per work item (even)
};

constexpr int kK1nd2Warps = 8;                                                       // one CTA per SM
constexpr int kK1nd2TableBytes = 1024 * 8 + 1024 * 8 + kFPad2 * 8;                   // wa2, tw, w2k
constexpr int kK1nd2WarpBytes = 1024 * 16 + kExchDual * 8;                           // dual spectrum, exchange tile
constexpr int k1nd2_smem_bytes() { return kK1nd2TableBytes + kK1nd2Warps * kK1nd2WarpBytes; }

__global__ void __launch_bounds__(kK1nd2Warps * 32, 1) k1nd_magnitude_2k(const K1nd2Args a) {
    constexpr int HR2 = 8;                              // a hop is 8 rows of 32 sample pairs
    B200_DYN_SMEM(unsigned char, smraw);
    float2* s_wa2 = reinterpret_cast<float2*>(smraw);
    float2* s_tw = reinterpret_cast<float2*>(smraw + 1024 * 8);
    float2* s_w2k = reinterpret_cast<float2*>(smraw + 2048 * 8);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbase = smraw + kK1nd2TableBytes + warp * kK1nd2WarpBytes;
    f2x2* zs = reinterpret_cast<f2x2*>(wbase);                       // [1024] (re pair, im pair) of the two frames
    f2* tile = reinterpret_cast<f2*>(wbase + 1024 * 16);
    const int nthr = kK1nd2Warps * 32;
    for (int i = threadIdx.x; i < 1024; i += nthr) {
        s_wa2[i] = a.tb.wa2[i];
        s_tw[i] = a.tb.tw[i];
    }
    for (int i = threadIdx.x; i < kFPad2; i += nthr) s_w2k[i] = (i < kF2) ? a.tb.w2k[i] : make_float2(0.f, 0.f);
    __syncthreads();

    const Geom& g = a.g;
    const int H = g.H;                                   // 512
    const long long n_items = (long long)g.n_units * a.n_runs;
    for (long long item = (long long)blockIdx.x * kK1nd2Warps + warp; item < n_items;
         item += (long long)gridDim.x * kK1nd2Warps) {
        const int ul = (int)(item / a.n_runs);
        const int run = (int)(item - (long long)ul * a.n_runs);
        const int u = g.u0 + ul;
        const int ic = u / g.C, c = u - ic * g.C;
        const long long i1 = (long long)ic * g.step - g.pad;
        const float* xrow = a.x + (long long)c * g.in_stride;
        const int t0 = run * a.run;
        const int t1 = min(t0 + a.run, g.T);
        for (int t = t0; t < t1; t += 2) {
            const bool vb = (t + 1 < t1);
            const long long base = (long long)t * H - kN2 / 2;
            {
                f2 re[32], im[32];
                {
                    // rows of 32 sample pairs: frame t is rows 0..31, frame t+1 rows 8..39 of the same 40-row window
                    float2 xr[32 + HR2];
                    const long long g0 = i1 + base;
                    const float* p0 = xrow + g0;
                    const int span = 64 * (32 + HR2);
                    if (base >= 0 && base + span <= g.Lp && g0 >= 0 && g0 + span <= g.n_total &&
                        ((reinterpret_cast<uintptr_t>(p0) & 7) == 0)) {
                        const float2* p = reinterpret_cast<const float2*>(p0) + lane;
#pragma unroll
                        for (int r = 0; r < 32 + HR2; ++r) xr[r] = __ldg(p + 32 * r);
                    } else {                                   // chunk / recording edges: zero-extended samples
#pragma unroll
                        for (int r = 0; r < 32 + HR2; ++r) {
                            const long long j = base + 2 * (lane + 32 * r);
                            xr[r] = make_float2(chunk_sample(xrow, j, i1, g.Lp, g.n_total),
                                                chunk_sample(xrow, j + 1, i1, g.Lp, g.n_total));
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const float2 w = s_wa2[lane + 32 * r];
                        // (a missing second frame transforms zeros; its outputs are not stored)
                        re[r] = f2_pack(xr[r].x * w.x, vb ? xr[r + HR2].x * w.x : 0.f);
                        im[r] = f2_pack(xr[r].y * w.y, vb ? xr[r + HR2].y * w.y : 0.f);
                    }
                }
                warp_fft1024_dual(re, im, tile, s_tw, lane);
                __syncwarp();                                  // (the previous pair's readers of zs are done)
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    f2x2 v;
                    v.a = re[brev5(q)];
                    v.b = im[brev5(q)];
                    zs[32 * q + lane] = v;
                }
                if (a.zcache) {                                // k2c_synthesize_2k loads Z instead of transforming again
                    float2* zpA = a.zcache + ((long long)ul * g.T + t) * 1024 + lane;
                    if (t >= a.z_lo && t < a.z_hi) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) zpA[32 * q] = make_float2(f2_lo(re[brev5(q)]), f2_lo(im[brev5(q)]));
                    }
                    if (vb && t + 1 >= a.z_lo && t + 1 < a.z_hi) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) zpA[1024 + 32 * q] = make_float2(f2_hi(re[brev5(q)]), f2_hi(im[brev5(q)]));
                    }
                }
            }
            __syncwarp();
            float* dA = a.mag + ((long long)ul * g.T + t) * kFPad2;
            float* dB = dA + kFPad2;
            const bool tap = (a.dbg.ul == ul);
            auto bins = [&](int k, bool valid) {
                const f2x2 own = zs[k];
                const f2x2 par = zs[(1024 - k) & 1023];
                const f2 Er = f2_mul_s(f2_add(own.a, par.a), 0.5f), Ei = f2_mul_s(f2_sub(own.b, par.b), 0.5f);
                const f2 Or = f2_mul_s(f2_add(own.b, par.b), 0.5f), Oi = f2_mul_s(f2_sub(par.a, own.a), 0.5f);
                const float2 W = s_w2k[k];
                const f2 Tr = f2_fma_s(Oi, -W.y, f2_mul_s(Or, W.x)), Ti = f2_fma_s(Or, W.y, f2_mul_s(Oi, W.x));
                const f2 Xr = f2_add(Er, Tr), Xi = f2_add(Ei, Ti);              // X[k]
                const f2 Yr = f2_sub(Er, Tr), Yi = f2_sub(Ti, Ei);              // X[1024-k] = conj(E - T)
                const f2 PX = f2_fma(Xr, Xr, f2_mul(Xi, Xi)), PY = f2_fma(Yr, Yr, f2_mul(Yi, Yi));
                if (valid) {
                    dA[k] = sqrtf(f2_lo(PX));
                    if (k != 512) dA[1024 - k] = sqrtf(f2_lo(PY));
                    if (vb) {
                        dB[k] = sqrtf(f2_hi(PX));
                        if (k != 512) dB[1024 - k] = sqrtf(f2_hi(PY));
                    }
                    if (tap) {
                        float* sp = a.dbg.spec + ((long long)t * kF2) * 2;
                        sp[2 * k] = f2_lo(Xr); sp[2 * k + 1] = f2_lo(Xi);
                        sp[2 * (1024 - k)] = f2_lo(Yr); sp[2 * (1024 - k) + 1] = f2_lo(Yi);
                        if (vb) {
                            sp += 2 * kF2;
                            sp[2 * k] = f2_hi(Xr); sp[2 * k + 1] = f2_hi(Xi);
                            sp[2 * (1024 - k)] = f2_hi(Yr); sp[2 * (1024 - k) + 1] = f2_hi(Yi);
                        }
                    }
                }
            };
#pragma unroll 4
            for (int q = 0; q < 16; ++q) bins(32 * q + lane, true);
            bins(512 + lane, lane == 0);                       // bin 512 (lane 0); the other lanes' slots hold mirrored bins
        }
    }
}

}  // namespace b200
