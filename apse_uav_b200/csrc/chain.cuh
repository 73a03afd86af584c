// chain.cuh -- per-pixel device code shared by the preprocess kernels (preprocess.cu) and the on-demand gray of the sparse
// evaluation (decode.cu): the Q5 / Q15 bilinear tap set of cv2.remap (aruco_detect.py:252) and the integer colour chain
// RGB2LAB -> LUT(L) -> LAB2RGB -> BGR2GRAY (aruco_detect.py:255-257,592) on the composed tables (P2Tables).
#pragma once
#include "common.cuh"

#define XZ_MAGIC 551553470                    // ceil(108 * 2^32 / 841)

__device__ __forceinline__ int gray_px(int c0, int c1, int c2) { return (c0 * 3735 + c1 * 19235 + c2 * 9798 + 16384) >> 15; }

// Q5 fixed-point source coordinate of the dependency's remap: rint(map * 32) evaluated in float32
__device__ __forceinline__ int q5(float m) { return __float2int_rn(__fmul_rn(m, 32.f)); }

// bilinear tap set of one output pixel
struct Taps {
    int off00;      // byte offset of tap (ix,iy) channel 0 in the source frame
    int w00, w01, w10, w11;
    unsigned mask;  // bit k set = tap k inside the image (k = 0:(ix,iy) 1:(ix+1,iy) 2:(ix,iy+1) 3:(ix+1,iy+1))
};

__device__ __forceinline__ Taps make_taps(float mx, float my, int sw, int sh, int cn)
{
    Taps t;
    int sx = q5(mx), sy = q5(my);
    int ix = sx >> 5, iy = sy >> 5, fx = sx & 31, fy = sy & 31;
    t.w00 = min(32767, (32 - fy) * (32 - fx) * 32);
    t.w01 = (32 - fy) * fx * 32;
    t.w10 = fy * (32 - fx) * 32;
    t.w11 = fy * fx * 32;
    bool x0 = ix >= 0 && ix < sw, x1 = ix + 1 >= 0 && ix + 1 < sw;
    bool y0 = iy >= 0 && iy < sh, y1 = iy + 1 >= 0 && iy + 1 < sh;
    t.mask = (x0 && y0 ? 1u : 0u) | (x1 && y0 ? 2u : 0u) | (x0 && y1 ? 4u : 0u) | (x1 && y1 ? 8u : 0u);
    t.off00 = (iy * sw + ix) * cn;
    return t;
}

__device__ __forceinline__ int sample(const uint8_t *__restrict__ src, const Taps &t, int rowbytes, int cn, int c)
{
    const uint8_t *p = src + t.off00 + c;
    int acc = 16384;
    if (t.mask == 15u) {
        acc += t.w00 * __ldg(p) + t.w01 * __ldg(p + cn) + t.w10 * __ldg(p + rowbytes) + t.w11 * __ldg(p + rowbytes + cn);
    } else {
        if (t.mask & 1u) acc += t.w00 * __ldg(p);
        if (t.mask & 2u) acc += t.w01 * __ldg(p + cn);
        if (t.mask & 4u) acc += t.w10 * __ldg(p + rowbytes);
        if (t.mask & 8u) acc += t.w11 * __ldg(p + rowbytes + cn);
    }
    return acc >> 15;
}

__device__ __forceinline__ int xz_px(int v)
{
    // v <= 3390: trunc(v*108/841) - 290 (signed high product + 1 for negative v); else floor(floor(v^2/2^14) v / 2^14)
    int lo = __mulhi(v, XZ_MAGIC) + (int)((unsigned)v >> 31) - 290;
    int hi = (((v * v) >> 14) * v) >> 14;
    return v <= 3390 ? lo : hi;
}

// colour chain of one pixel on the composed tables: (c0,c1,c2) -> corrected (o0,o1,o2) and gray.
// Shared-memory wavefronts are what the kernel runs out of first (84 % of the pipe), so the chain spends a few integer
// instructions where that saves look-ups with scattered indices: L comes from fY arithmetically and indexes a 256-entry
// {y, f} table (narrow index spread, ~1 wavefront) instead of an 8-byte entry per idxY (5.5 wavefronts), and the two
// chroma shifts are computed (clamp + multiply + shift) instead of being read from tables.
__device__ __forceinline__ int chain_px(const P2Tables *T, int c0, int c1, int c2, int &o0, int &o1, int &o2)
{
    int R = T->gamma[c0], G = T->gamma[c1], B = T->gamma[c2];
    int iX = (R * 1777 + G * 1541 + B * 778 + 2048) >> 12;
    int iY = (R * 871 + G * 2929 + B * 296 + 2048) >> 12;
    int iZ = (R * 73 + G * 448 + B * 3575 + 2048) >> 12;
    int fX = T->cb[iX], fY = T->cb[iY], fZ = T->cb[iZ];
    const int L = __vimin_s32_relu((296 * fY - 1336934 + 16384) >> 15, 255);
    const uint32_t yf = T->yf[L];
    const int y = (int)(yf & 0xffffu), f = (int)(yf >> 16);
    const int a = __vimin_s32_relu((500 * (fX - fY) + 128 * 32768 + 16384) >> 15, 255);
    const int b = __vimin_s32_relu((200 * (fY - fZ) + 128 * 32768 + 16384) >> 15, 255);
    const int adiv = ((a * (5 * 53687) + 128) >> 13) - 4194, bdiv = ((b * 41943 + 16) >> 9) - 10485 + 1;
    int X = xz_px(f + adiv), Z = xz_px(f - bdiv);
    int r0 = (12615 * X - 6296 * y - 2223 * Z + 8192) >> 14;
    int r1 = (-3773 * X + 7684 * y + 185 * Z + 8192) >> 14;
    int r2 = (217 * X - 836 * y + 4715 * Z + 8192) >> 14;
    o0 = T->invgamma[__vimin_s32_relu(r0, 4095)];   // clamp to [0, 4095] in one instruction
    o1 = T->invgamma[__vimin_s32_relu(r1, 4095)];
    o2 = T->invgamma[__vimin_s32_relu(r2, 4095)];
    return gray_px(o0, o1, o2);
}


// exact gray of output pixel (x, y) of one frame straight from the source frame (direct gather; tables may live in global
// or shared memory): what K1t / k_preprocess_fused write for that pixel
__device__ __forceinline__ int exact_gray_px(const uint8_t *__restrict__ src, const float *__restrict__ mapx, const float *__restrict__ mapy,
                                             const P2Tables *T, int w, int h, int x, int y)
{
    const size_t o = (size_t)y * w + x;
    const Taps t = make_taps(__ldg(mapx + o), __ldg(mapy + o), w, h, 3);
    const int c0 = sample(src, t, w * 3, 3, 0), c1 = sample(src, t, w * 3, 3, 1), c2 = sample(src, t, w * 3, 3, 2);
    int o0, o1, o2;
    return chain_px(T, c0, c1, c2, o0, o1, o2);
}
