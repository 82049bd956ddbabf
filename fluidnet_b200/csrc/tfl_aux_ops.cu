// Operators of tfluids/init.lua that sit around the simulation step ("next"):
// volumetricUpSamplingNearestForward, rectangularBlur, signedDistanceField.  Compiled with
// -fmad=false; each kernel restates the reference's float arithmetic in the reference's order
// (torch/tfluids/generic/tfluids.cc:509-557, 641-760, 766-822), so results are bit-identical.
#include <cuda_runtime.h>

#include "tfl_kernels.h"

namespace tfl {

namespace {

// out[bf][z][y][x] = in[bf][z / r][y / r][x / r]
__global__ void k_upsample_nearest(const float* __restrict__ in, float* __restrict__ out, int nz, int ny, int nx,
                                   int ratio, long long total) {
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= total) return;
  const int ox = nx * ratio, oy = ny * ratio, oz = nz * ratio;
  const int x = (int)(o % ox);
  const int y = (int)((o / ox) % oy);
  const int z = (int)((o / ((long long)ox * oy)) % oz);
  const long long bf = o / ((long long)ox * oy * oz);
  out[o] = __ldg(in + ((bf * nz + z / ratio) * ny + y / ratio) * (long long)nx + x / ratio);
}

// One thread per line: the running box sum of DoRectangularBlurAlongAxis is sequential along the line
// (that order IS the result in floating point); the lines are independent.
__global__ void k_blur_axis(const float* __restrict__ src, float* __restrict__ dst, int nz, int ny, int nx, int axis,
                            int rad, long long lines) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= lines) return;
  const long long sy = nx, sz = (long long)nx * ny, sf = sz * nz;
  int size;
  long long stride, base;
  if (axis == 2) {            // lines along z: l -> (bf, y, x)
    size = nz; stride = sz;
    base = (l / sz) * sf + (l % sz);
  } else if (axis == 1) {     // along y: l -> (bf, z, x), x fastest
    size = ny; stride = sy;
    const long long x = l % nx, zz = (l / nx) % nz, bf = l / ((long long)nx * nz);
    base = bf * sf + zz * sz + x;
  } else {                    // along x: l -> (bf, z, y)
    size = nx; stride = 1;
    base = l * (long long)nx;
  }
  const float* s = src + base;
  float* d = dst + base;
  float val = s[0] * (float)(rad + 1);
  for (int i = 0; i < size && i < rad; i++) val += s[i * stride];
  const float mul_const = 1.0f / (float)(rad * 2 + 1);
  for (int i = 0; i < size; i++) {
    const int iminus = i - rad - 1 > 0 ? i - rad - 1 : 0;
    val -= s[iminus * stride];
    const int iplus = i + rad < size - 1 ? i + rad : size - 1;
    val += s[iplus * stride];
    d[i * stride] = val * mul_const;
  }
}

__global__ void k_signed_distance_field(const float* __restrict__ flags, float* __restrict__ dst, int nz, int ny,
                                        int nx, int rad, long long total) {
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= total) return;
  const long long n = (long long)nz * ny * nx;
  const float* fb = flags + (o / n) * n;
  const long long c = o % n;
  const int x = (int)(c % nx), y = (int)((c / nx) % ny), z = (int)(c / ((long long)nx * ny));
  if (((int)__ldg(fb + c)) & 2) { dst[o] = 0.0f; return; }
  float dist_sq = (float)(rad * rad);
  const int zmin = max(0, z - rad), zmax = min(nz - 1, z + rad);
  const int ymin = max(0, y - rad), ymax = min(ny - 1, y + rad);
  const int xmin = max(0, x - rad), xmax = min(nx - 1, x + rad);
  for (int zs = zmin; zs <= zmax; zs++)
    for (int ys = ymin; ys <= ymax; ys++)
      for (int xs = xmin; xs <= xmax; xs++)
        if (((int)__ldg(fb + ((long long)zs * ny + ys) * nx + xs)) & 2) {
          const float cur = (float)((z - zs) * (z - zs) + (y - ys) * (y - ys) + (x - xs) * (x - xs));
          if (dist_sq > cur) dist_sq = cur;
        }
  dst[o] = sqrtf(dist_sq);
}

inline unsigned blocks(long long n) { return (unsigned)((n + 255) / 256); }

}  // namespace

void launch_upsample_nearest(const float* in, float* out, int nbf, int nz, int ny, int nx, int ratio, cudaStream_t st) {
  const long long total = (long long)nbf * nz * ny * nx * ratio * ratio * ratio;
  k_upsample_nearest<<<blocks(total), 256, 0, st>>>(in, out, nz, ny, nx, ratio, total);
}
void launch_blur_axis(const float* src, float* dst, int nbf, int nz, int ny, int nx, int axis, int rad, cudaStream_t st) {
  const long long cells = (long long)nbf * nz * ny * nx;
  const long long lines = cells / (axis == 2 ? nz : axis == 1 ? ny : nx);
  k_blur_axis<<<blocks(lines), 256, 0, st>>>(src, dst, nz, ny, nx, axis, rad, lines);
}
void launch_signed_distance_field(const float* flags, float* dst, int nb, int nz, int ny, int nx, int rad,
                                  cudaStream_t st) {
  const long long total = (long long)nb * nz * ny * nx;
  k_signed_distance_field<<<blocks(total), 256, 0, st>>>(flags, dst, nz, ny, nx, rad, total);
}

}  // namespace tfl
