// Development probe: is the hardware JPEG engine usable here, and how fast does it decode 4K frames straight
// into the pipeline's [B][H][W][3] BGR layout?   nvcc -O2 -o nvjpeg_probe nvjpeg_probe.cu -lnvjpeg ; ./nvjpeg_probe frame.jpg 32
#include <cuda_runtime.h>
#include <nvjpeg.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { int _s = (int)(x); if (_s) { printf("%s failed: %d (line %d)\n", #x, _s, __LINE__); return 1; } } while (0)
static int run(nvjpegBackend_t backend, const char *name, const std::vector<unsigned char> &jpg, int B)
{
    nvjpegHandle_t h; nvjpegJpegState_t st;
    nvjpegStatus_t s = nvjpegCreateEx(backend, nullptr, nullptr, 0, &h);
    if (s) { printf("%s: nvjpegCreateEx -> %d (backend not available)\n", name, (int)s); return 0; }
    CK(nvjpegJpegStateCreate(h, &st));
    int nc, ws[4], hs[4]; nvjpegChromaSubsampling_t ss;
    CK(nvjpegGetImageInfo(h, jpg.data(), jpg.size(), &nc, &ss, ws, hs));
    const int W = ws[0], H = hs[0];
    unsigned char *out; CK(cudaMalloc(&out, (size_t)B * W * H * 3));
    CK(nvjpegDecodeBatchedInitialize(h, st, B, 1, NVJPEG_OUTPUT_BGRI));
    std::vector<const unsigned char *> ptrs(B, jpg.data()); std::vector<size_t> lens(B, jpg.size()); std::vector<nvjpegImage_t> imgs(B);
    for (int i = 0; i < B; i++) { memset(&imgs[i], 0, sizeof(nvjpegImage_t)); imgs[i].channel[0] = out + (size_t)i * W * H * 3; imgs[i].pitch[0] = (size_t)W * 3; }
    cudaStream_t stream; CK(cudaStreamCreate(&stream));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    s = nvjpegDecodeBatched(h, st, ptrs.data(), lens.data(), imgs.data(), stream);
    if (s) { printf("%s: nvjpegDecodeBatched -> %d\n", name, (int)s); return 0; }
    CK(cudaStreamSynchronize(stream));
    const int reps = 5;
    cudaEventRecord(e0, stream);
    for (int r = 0; r < reps; r++) CK(nvjpegDecodeBatched(h, st, ptrs.data(), lens.data(), imgs.data(), stream));
    cudaEventRecord(e1, stream); CK(cudaStreamSynchronize(stream));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<unsigned char> px(16); cudaMemcpy(px.data(), out + ((size_t)(H / 2) * W + W / 2) * 3, 12, cudaMemcpyDeviceToHost);
    printf("%s: %dx%d subsampling %d, batch %d: %.1f frames/s (%.2f ms per frame), jpeg %.2f MB, centre pixels %d %d %d | %d %d %d\n", name, W, H, (int)ss, B,
           1e3 * reps * B / ms, ms / (reps * B), jpg.size() / 1e6, px[0], px[1], px[2], px[3], px[4], px[5]);
    cudaFree(out); nvjpegJpegStateDestroy(st); nvjpegDestroy(h);
    return 0;
}
int main(int argc, char **argv)
{
    FILE *f = fopen(argv[1], "rb"); if (!f) { printf("no file\n"); return 1; }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<unsigned char> jpg(n); if (fread(jpg.data(), 1, n, f) != (size_t)n) return 1; fclose(f);
    int B = argc > 2 ? atoi(argv[2]) : 32;
    run(NVJPEG_BACKEND_HARDWARE, "hardware", jpg, B);
    run(NVJPEG_BACKEND_GPU_HYBRID, "gpu_hybrid", jpg, B);
    run(NVJPEG_BACKEND_DEFAULT, "default", jpg, B);
    return 0;
}
