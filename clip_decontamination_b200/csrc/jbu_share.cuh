// Sharing the JBU kernel generation across overlapping crops (DESIGN.md "JBU kernels shared across crops").
//
// guidance pooling -> range projection -> range kernel -> kernel fix-up -> composite (bicubic-folded) kernel depend on
// the IMAGE, not on the crop, for every pixel that is far enough from the crop border: reflect padding (radius R) and
// the bicubic border clamps only reach R + 6 pixels into a crop.  With stride 112 / crop 224 every image pixel lies in
// up to four (nine at 1024) crops, so these tensors are computed ONCE per image pixel ("image level", canvas
// coordinates) plus, per crop, for a border frame: the top / bottom `fb` rows and the left / right 16 columns.
// The frame is stored compactly, crop after crop:
//   rows [0, fb*gw)                top strip, row-major            rows [fb*gw, 2*fb*gw)   bottom strip
//   rows [2*fb*gw, ...)            one 32-pixel group per interior row: x in [0,16) then x in [gw-16, gw)
// Valid when every window is a full crop whose origin is a multiple of 16 image pixels (all BASELINE configs; snapped
// last windows such as 1300 - 224 = 1076 are not, and take the per-crop path).
#pragma once
#include <stdint.h>

#ifndef __host__
#define __host__
#define __device__
#endif

struct ShareGeom {            // device-side view of cseg_jbu_share + the region size
  const int32_t* wins;        // [n_crops][4] = {y1, x1, h, w} in canvas pixels
  int shift;                  // log2(image pixels per stage pixel)
  int pitch;                  // stage pixels per image-level row
};

__host__ __device__ inline int border_rows(int gh, int gw, int fb) { return 2 * fb * gw + (gh - 2 * fb) * 32; }
__host__ __device__ inline bool border_interior(int y, int x, int gh, int gw, int fb) {
  return y >= fb && y < gh - fb && x >= 16 && x < gw - 16;
}
// compact row of border pixel (y, x) (must not be interior)
__host__ __device__ inline int border_index(int y, int x, int gh, int gw, int fb) {
  if (y < fb) return y * gw + x;
  if (y >= gh - fb) return fb * gw + (y - (gh - fb)) * gw + x;
  return 2 * fb * gw + (y - fb) * 32 + (x < 16 ? x : 16 + x - (gw - 16));
}
// inverse: compact row r -> (y, x)
__host__ __device__ inline void border_coords(int r, int gh, int gw, int fb, int& y, int& x) {
  if (r < fb * gw) { y = r / gw; x = r - y * gw; return; }
  r -= fb * gw;
  if (r < fb * gw) { const int yy = r / gw; y = gh - fb + yy; x = r - yy * gw; return; }
  r -= fb * gw;
  y = fb + (r >> 5);
  const int c = r & 31;
  x = c < 16 ? c : gw - 32 + c;
}
