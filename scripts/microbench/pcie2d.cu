// PCIe copy shapes: one 1-D copy vs cudaMemcpy2DAsync of 64 rows vs 64 separate row copies (pinned host memory).
#include <cstdio>
#include <cuda_runtime.h>
#include <chrono>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    const size_t rows = 64, N = 28800000, es = 4;
    char *h, *d;
    cudaMallocHost(&h, rows * N * es);
    cudaMalloc(&d, rows * 8000000 * es);
    cudaStream_t s; cudaStreamCreate(&s);
    for (size_t w : {600000ul, 1800000ul, 4800000ul}) {
        const size_t wb = w * es;
        for (int dir = 0; dir < 2; ++dir) {
            auto kind = dir ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
            double t[3];
            for (int mode = 0; mode < 3; ++mode) {
                for (int rep = 0; rep < 3; ++rep) {
                    cudaStreamSynchronize(s);
                    double t0 = now();
                    if (mode == 0) { if (dir) cudaMemcpyAsync(h, d, rows * wb, kind, s); else cudaMemcpyAsync(d, h, rows * wb, kind, s); }
                    else if (mode == 1) { if (dir) cudaMemcpy2DAsync(h, N * es, d, wb, wb, rows, kind, s); else cudaMemcpy2DAsync(d, wb, h, N * es, wb, rows, kind, s); }
                    else for (size_t r = 0; r < rows; ++r) { if (dir) cudaMemcpyAsync(h + r * N * es, d + r * wb, wb, kind, s); else cudaMemcpyAsync(d + r * wb, h + r * N * es, wb, kind, s); }
                    cudaStreamSynchronize(s);
                    t[mode] = now() - t0;
                }
            }
            printf("%s width %zu samples x 64 rows (%.1f MB): 1D %.2f ms (%.1f GB/s)  2D %.2f ms (%.1f GB/s)  64x1D %.2f ms (%.1f GB/s)\n",
                   dir ? "D2H" : "H2D", w, rows * wb / 1e6, t[0] * 1e3, rows * wb / t[0] / 1e9, t[1] * 1e3, rows * wb / t[1] / 1e9,
                   t[2] * 1e3, rows * wb / t[2] / 1e9);
        }
    }
    return 0;
}
