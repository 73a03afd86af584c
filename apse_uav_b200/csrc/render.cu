// render.cu -- annotated frames of aruco_detect.py (drawMarkers :614-616, drawBoundingBox :421-425, drawLinesOnImage :494-500,
// drawPoints :378-380): overlay primitives drawn into device-resident BGR frames, so that `saveImages` output can be
// produced without moving the frames to the host first.  Text (cv2.putText) stays a host call.
//
// cv2.line of thickness t > 1 fills the rectangle of half-width t / 2 (+ 0.5 for odd t) around the segment, outlines it with
// 1-pixel lines (which adds about half a pixel) and closes both ends with a filled circle of radius floor(t / 2 + 0.5); a filled
// cv2.circle of radius r is the set of pixels with dx^2 + dy^2 <= r^2.  The kernel paints exactly those sets from their
// analytic description; cv2 rasterises the rectangle through its polygon filler in 16.16 fixed point, so single edge pixels
// can differ (tests/test_gpu_render.py: IoU >= 0.93 per frame, discs identical).
#include "common.cuh"

// one CTA per frame walks the frame's primitives in list order (later primitives overwrite earlier ones, as successive cv2
// calls do); the pixels of a primitive are spread over the CTA: a band around the segment, not its bounding box
__global__ void __launch_bounds__(256) k_draw_overlay(uint8_t *__restrict__ bgr, int w, int h, int batch, const apse_overlay_prim *__restrict__ prims,
                                                      int n_prims)
{
    const int f = blockIdx.x;
    // first primitive of this frame (the list is sorted by frame): binary search
    int lo = 0, hi = n_prims;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (prims[mid].frame < f) lo = mid + 1; else hi = mid; }
    uint8_t *img = bgr + (size_t)f * w * h * 3;
    for (int pi = lo; pi < n_prims && prims[pi].frame == f; pi++) {
        const apse_overlay_prim P = prims[pi];
        const int t_ = P.thickness;
        const float hw = 0.5f * (float)t_ + ((t_ & 1) ? 0.5f : 0.f) + 0.5f;          // body half-width (see the header comment)
        const float cap = P.kind == 1 ? (float)t_ : floorf(0.5f * (float)t_ + 0.5f);   // radius of the end caps / of the disc
        const float r = fmaxf(hw, cap);
        const int rr = (int)ceilf(r * 1.4143f) + 1;   // half-width of the band along the minor axis (a 45-degree segment is sqrt(2) wider there)
        const int x1 = P.kind == 1 ? P.x0 : P.x1, y1 = P.kind == 1 ? P.y0 : P.y1;
        const int dx = x1 - P.x0, dy = y1 - P.y0;
        const bool steep = abs(dy) > abs(dx);
        const int len = max(abs(dx), abs(dy)), band = 2 * rr + 1;
        const float fdx = (float)dx, fdy = (float)dy, l2 = fdx * fdx + fdy * fdy, inv = len > 0 ? 1.f / (float)len : 0.f;
        // walk the major axis (extended by the cap radius at both ends), and across it a band of +-rr pixels
        const long long total = (long long)(len + 1 + 2 * rr) * band;
        for (long long i = threadIdx.x; i < total; i += blockDim.x) {
            const int m = (int)(i / band) - rr, o = (int)(i % band) - rr;
            const float t = (float)m * inv;
            const int px = (int)lrintf((float)P.x0 + t * fdx) + (steep ? o : 0), py = (int)lrintf((float)P.y0 + t * fdy) + (steep ? 0 : o);
            if (px < 0 || px >= w || py < 0 || py >= h) continue;
            const float vx = (float)(px - P.x0), vy = (float)(py - P.y0);
            bool on = vx * vx + vy * vy <= cap * cap;                                  // cap at the first end point / the disc
            if (!on && l2 > 0.f) {
                const float wx = (float)(px - x1), wy = (float)(py - y1);
                on = wx * wx + wy * wy <= cap * cap;                                   // cap at the second end point
                if (!on) {
                    const float l = sqrtf(l2), along = (vx * fdx + vy * fdy) / l, perp = fabsf(vx * fdy - vy * fdx) / l;
                    on = along >= -0.5f && along <= l + 0.5f && perp <= hw;            // body
                }
            }
            if (!on) continue;
            uint8_t *q = img + ((size_t)py * w + px) * 3;
            q[0] = P.bgr[0]; q[1] = P.bgr[1]; q[2] = P.bgr[2];
        }
        __syncthreads();   // primitives of a frame are drawn in order
    }
}

extern "C" int apse_draw_overlay(apse_ctx *ctx, uint8_t *bgr, int w, int h, int batch, const apse_overlay_prim *prims, int n_prims, void *stream)
{
    if (!ctx || !bgr || w <= 0 || h <= 0 || batch <= 0 || n_prims < 0 || (n_prims > 0 && !prims)) CTX_FAIL(ctx, APSE_ERR_INVALID_ARG, "draw_overlay: bad argument");
    if (n_prims == 0) return APSE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    KLAUNCH(ctx, KID_DRAW, st, k_draw_overlay<<<batch, 256, 0, st>>>(bgr, w, h, batch, prims, n_prims));
    return APSE_OK;
}
