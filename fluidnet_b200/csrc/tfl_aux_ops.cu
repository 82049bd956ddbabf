// Operators of tfluids/init.lua that sit around the simulation step ("next"):
// volumetricUpSamplingNearestForward / Backward, rectangularBlur, signedDistanceField, and the backward
// passes of the velocity-divergence and velocity-update modules.  Compiled with
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

// ---- backward operators (generic/tfluids.cc:49-134, 216-345, 563-635) as gathers ----------------
__device__ __forceinline__ bool interior_cell(int nz, int ny, int nx, int is3d, int k, int j, int i) {
  return !(i < 1 || i > nx - 2 || j < 1 || j > ny - 2 || (is3d && (k < 1 || k > nz - 2)));
}

// grad_u_c(X) = [X contributes] go(X) - [X - e_c contributes] go(X - e_c); a cell contributes when it is an
// interior fluid cell.  Two terms at most: the reference's atomic order cannot change the bits.
__global__ void k_velocity_divergence_backward(const float* __restrict__ flags, const float* __restrict__ go,
                                               float* __restrict__ grad_u, int nz, int ny, int nx, int is3d,
                                               long long total) {
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= total) return;
  const long long n = (long long)nz * ny * nx;
  const int b = (int)(o / n);
  const long long c = o % n;
  const int i = (int)(c % nx), j = (int)((c / nx) % ny), k = (int)(c / ((long long)nx * ny));
  const float* fb = flags + b * n;
  const float* gb = go + b * n;
  const int nc = is3d ? 3 : 2;
  const bool here = interior_cell(nz, ny, nx, is3d, k, j, i) && (((int)__ldg(fb + c)) & 1);
  const long long st[3] = {1, nx, (long long)nx * ny};
  const int ijk[3] = {i, j, k};
  for (int a = 0; a < nc; a++) {
    float g = 0.0f;
    if (here) g += __ldg(gb + c);
    if (ijk[a] > 0) {
      const int ii = i - (a == 0), jj = j - (a == 1), kk = k - (a == 2);
      if (interior_cell(nz, ny, nx, is3d, kk, jj, ii) && (((int)__ldg(fb + c - st[a])) & 1)) g -= __ldg(gb + c - st[a]);
    }
    grad_u[((long long)b * nc + a) * n + c] = g;
  }
}

// grad_p(X): minus this cell's own face gradients (fluid or empty -neighbour), plus the face gradients of the
// +x / +y / +z neighbours whose update read p(X).  Fixed summation order (the reference's is unspecified).
__global__ void k_velocity_update_backward(const float* __restrict__ flags, const float* __restrict__ go,
                                           float* __restrict__ grad_p, int nz, int ny, int nx, int is3d,
                                           long long total) {
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= total) return;
  const long long n = (long long)nz * ny * nx;
  const int b = (int)(o / n);
  const long long c = o % n;
  const int i = (int)(c % nx), j = (int)((c / nx) % ny), k = (int)(c / ((long long)nx * ny));
  const float* fb = flags + b * n;
  const int nc = is3d ? 3 : 2;
  const float* gb = go + (long long)b * nc * n;
  const long long st[3] = {1, nx, (long long)nx * ny};
  const int xf = ((int)__ldg(fb + c)) & 1;
  float g = 0.0f;
  if (xf && interior_cell(nz, ny, nx, is3d, k, j, i)) {
    for (int a = 0; a < nc; a++) if (((int)__ldg(fb + c - st[a])) & 1) g -= __ldg(gb + a * n + c);
    for (int a = 0; a < nc; a++) if (((int)__ldg(fb + c - st[a])) & 4) g -= __ldg(gb + a * n + c);
  }
  if (xf) {
    const int lim[3] = {nx, ny, nz}, ijk[3] = {i, j, k};
    for (int a = 0; a < nc; a++) {
      if (ijk[a] + 1 >= lim[a]) continue;
      const int ii = i + (a == 0), jj = j + (a == 1), kk = k + (a == 2);
      if (!interior_cell(nz, ny, nx, is3d, kk, jj, ii)) continue;
      const int fy = (int)__ldg(fb + c + st[a]);
      if ((fy & 1) || ((fy & 4) && !(fy & 16))) g += __ldg(gb + a * n + c + st[a]);
    }
  }
  grad_p[o] = g;
}

__global__ void k_upsample_nearest_backward(const float* __restrict__ go, float* __restrict__ gi, int nz, int ny,
                                            int nx, int ratio, long long total) {
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= total) return;
  const int x = (int)(o % nx), y = (int)((o / nx) % ny), z = (int)((o / ((long long)nx * ny)) % nz);
  const long long bf = o / ((long long)nx * ny * nz);
  const long long oy = (long long)ny * ratio, ox = (long long)nx * ratio, oz = (long long)nz * ratio;
  float sum = 0;
  for (int zu = 0; zu < ratio; zu++)
    for (int yu = 0; yu < ratio; yu++)
      for (int xu = 0; xu < ratio; xu++)
        sum += __ldg(go + ((bf * oz + (long long)z * ratio + zu) * oy + (long long)y * ratio + yu) * ox + (long long)x * ratio + xu);
  gi[o] = sum;
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

void launch_velocity_divergence_backward(const float* flags, const float* go, float* grad_u, int nb, int nz, int ny,
                                         int nx, int is3d, cudaStream_t st) {
  const long long total = (long long)nb * nz * ny * nx;
  k_velocity_divergence_backward<<<blocks(total), 256, 0, st>>>(flags, go, grad_u, nz, ny, nx, is3d, total);
}
void launch_velocity_update_backward(const float* flags, const float* go, float* grad_p, int nb, int nz, int ny, int nx,
                                     int is3d, cudaStream_t st) {
  const long long total = (long long)nb * nz * ny * nx;
  k_velocity_update_backward<<<blocks(total), 256, 0, st>>>(flags, go, grad_p, nz, ny, nx, is3d, total);
}
void launch_upsample_nearest_backward(const float* go, float* gi, int nbf, int nz, int ny, int nx, int ratio,
                                      cudaStream_t st) {
  const long long total = (long long)nbf * nz * ny * nx;
  k_upsample_nearest_backward<<<blocks(total), 256, 0, st>>>(go, gi, nz, ny, nx, ratio, total);
}

}  // namespace tfl
