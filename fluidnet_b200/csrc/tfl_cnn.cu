// Convolution stack of the pressure-projection network (torch/lib/model.lua:262-364:
// stride-1 zero-padded cross-correlation + bias (+ReLU), cudnn.Volumetric/Spatial
// Convolution in the reference, torch/lib/model_utils.lua:74-116).
//
// This file holds the fp32 FMA path: one thread per output voxel computing all output
// channels from shared-memory weights.  It is the numerically tight (1e-5 class)
// implementation and the parity anchor for the tensor-core path (tfl_cnn_tc.cu).
#include "tfl_device.cuh"
#include "tfl_kernels.h"

namespace tfl {

// Non-linearity between layers (torch.addNonlinearity, lib/model_utils.lua): 0 none, 1 ReLU, 2 sigmoid.
__device__ __forceinline__ float activate(float r, int act) {
  if (act == 1) return r < 0.0f ? 0.0f : r;
  if (act == 2) return 1.0f / (1.0f + expf(-r));
  return r;
}

// weights in smem as [cin][tap][COUT]; taps ordered (dz, dy, dx).
template <int COUT, int KS, bool IS3D>
__global__ void __launch_bounds__(256)
k_conv_direct(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ w,
              const float* __restrict__ bias, int cin, int act, Geo g) {
  extern __shared__ float sw[];
  constexpr int KZ = IS3D ? KS : 1;
  constexpr int TAPS = KZ * KS * KS;
  const int nthreads = blockDim.x * blockDim.y * blockDim.z;
  const int tid = (threadIdx.z * blockDim.y + threadIdx.y) * blockDim.x + threadIdx.x;
  for (int t = tid; t < cin * TAPS * COUT; t += nthreads) sw[t] = w[t];
  __syncthreads();

  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int zz = blockIdx.z * blockDim.z + threadIdx.z;
  const int nzr = g.zhi - g.zlo;
  const int b = zz / nzr;
  const int k = g.zlo + (zz - b * nzr);
  if (i >= g.nx || j >= g.ny || b >= g.nb) return;

  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; o++) acc[o] = __ldg(bias + o);

  constexpr int P = (KS - 1) / 2;
  constexpr int PZ = (KZ - 1) / 2;
  const int kg = k + g.zoff;
  for (int c = 0; c < cin; c++) {
    const float* ib = in + ((long long)b * cin + c) * g.n;
    const float* wc = sw + c * TAPS * COUT;
#pragma unroll
    for (int dz = 0; dz < KZ; dz++) {
      const int zg = kg + dz - PZ;
      if (zg < 0 || zg >= g.gnz) continue;           // zero padding at the GLOBAL boundary
      const int zl = zg - g.zoff;
#pragma unroll
      for (int dy = 0; dy < KS; dy++) {
        const int yy = j + dy - P;
        if (yy < 0 || yy >= g.ny) continue;
#pragma unroll
        for (int dx = 0; dx < KS; dx++) {
          const int xx = i + dx - P;
          if (xx < 0 || xx >= g.nx) continue;
          const float v = __ldg(ib + ((long long)zl * g.ny + yy) * g.nx + xx);
          const float* wt = wc + ((dz * KS + dy) * KS + dx) * COUT;
#pragma unroll
          for (int o = 0; o < COUT; o++) acc[o] = fmaf(v, wt[o], acc[o]);
        }
      }
    }
  }
  const long long c0 = cell(g, k, j, i);
#pragma unroll
  for (int o = 0; o < COUT; o++) {
    out[((long long)b * COUT + o) * g.n + c0] = activate(acc[o], act);
  }
}


// Any (cout, k): one thread per output value, weights [cin][tap][cout] read through the cache.  The
// fallback for layer shapes outside the specialised table (e.g. the 256-channel 1x1x1 convolution
// inside a VolumetricConvolutionUpsample); same accumulation order as k_conv_direct.
__global__ void k_conv_any(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ w,
                           const float* __restrict__ bias, int cin, int cout, int ks, int act, Geo g) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)g.nb * cout * g.n;
  if (t >= total) return;
  const long long c0 = t % g.n;
  const int o = (int)((t / g.n) % cout);
  const int b = (int)(t / (g.n * cout));
  const int i = (int)(c0 % g.nx), j = (int)((c0 / g.nx) % g.ny), k = (int)(c0 / ((long long)g.nx * g.ny));
  const int kz = g.is3d ? ks : 1;
  const int P = (ks - 1) / 2, PZ = (kz - 1) / 2;
  const int taps = kz * ks * ks;
  float acc = __ldg(bias + o);
  for (int c = 0; c < cin; c++) {
    const float* ib = in + ((long long)b * cin + c) * g.n;
    const float* wc = w + (long long)c * taps * cout;
    for (int dz = 0; dz < kz; dz++) {
      const int zz = k + dz - PZ;
      if (zz < 0 || zz >= g.nz) continue;
      for (int dy = 0; dy < ks; dy++) {
        const int yy = j + dy - P;
        if (yy < 0 || yy >= g.ny) continue;
        for (int dx = 0; dx < ks; dx++) {
          const int xx = i + dx - P;
          if (xx < 0 || xx >= g.nx) continue;
          acc = fmaf(__ldg(ib + ((long long)zz * g.ny + yy) * g.nx + xx), __ldg(wc + ((dz * ks + dy) * ks + dx) * cout + o), acc);
        }
      }
    }
  }
  out[t] = activate(acc, act);
}

// cudnn.{Spatial,Volumetric}{Average,Max}Pooling(p, p[, p], p, p[, p]) (lib/model_utils.lua:184-209):
// window p^d, stride p, no padding.  in [bc][nz][ny][nx] -> out [bc][nz/pz][ny/p][nx/p], pz = p in 3-D else 1.
__global__ void k_pool(const float* __restrict__ in, float* __restrict__ out, int nz, int ny, int nx, int p,
                       int is3d, int is_max, long long total) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int pz = is3d ? p : 1;
  const int ox = nx / p, oy = ny / p, oz = nz / pz;
  const int x = (int)(t % ox), y = (int)((t / ox) % oy), z = (int)((t / ((long long)ox * oy)) % oz);
  const long long bc = t / ((long long)ox * oy * oz);
  const float* ib = in + bc * (long long)nz * ny * nx;
  float acc = is_max ? -INFINITY : 0.0f;
  for (int dz = 0; dz < pz; dz++)
    for (int dy = 0; dy < p; dy++)
      for (int dx = 0; dx < p; dx++) {
        const float v = __ldg(ib + ((long long)(z * pz + dz) * ny + (y * p + dy)) * nx + (x * p + dx));
        acc = is_max ? fmaxf(acc, v) : acc + v;
      }
  out[t] = is_max ? acc : acc / (float)(pz * p * p);
}

// The view / permute / copy of nn.{Spatial,Volumetric}ConvolutionUpsample:updateOutput
// (lib/modules/*_convolution_upsample.lua): in [b][nO * sT * sH * sW][d][h][w] -> out [b][nO][d sT][h sH][w sW],
// out(b, o, z sT + st, y sH + sh, x sW + sw) = in(b, ((o sT + st) sH + sh) sW + sw, z, y, x); sT = 1 in 2-D.
__global__ void k_pixel_shuffle(const float* __restrict__ in, float* __restrict__ out, int n_out, int nz, int ny,
                                int nx, int s, int is3d, long long total) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int st_ = is3d ? s : 1;
  const int ox = nx * s, oy = ny * s, oz = nz * st_;
  const int X = (int)(t % ox), Y = (int)((t / ox) % oy), Z = (int)((t / ((long long)ox * oy)) % oz);
  const int o = (int)((t / ((long long)ox * oy * oz)) % n_out);
  const long long b = t / ((long long)ox * oy * oz * n_out);
  const int x = X / s, sw = X % s, y = Y / s, sh = Y % s, z = Z / st_, sz = Z % st_;
  const long long ch = ((long long)(o * st_ + sz) * s + sh) * s + sw;
  const long long cin_total = (long long)n_out * st_ * s * s;
  out[t] = __ldg(in + ((b * cin_total + ch) * nz + z) * (long long)ny * nx + (long long)y * nx + x);
}

template <int COUT, int KS, bool IS3D>
static bool conv_launch(const float* in, float* out, const float* w, const float* b, int cin, int act,
                        const Geo& g, cudaStream_t st) {
  const int nzr = g.zhi - g.zlo;
  dim3 block = IS3D ? dim3(32, 4, 2) : dim3(32, 8, 1);
  dim3 grid((g.nx + block.x - 1) / block.x, (g.ny + block.y - 1) / block.y,
            ((long long)g.nb * nzr + block.z - 1) / block.z);
  constexpr int KZ = IS3D ? KS : 1;
  const size_t smem = sizeof(float) * cin * KZ * KS * KS * COUT;
  if (smem > 200 * 1024) return false;             // weights do not fit shared memory: generic kernel
  if (smem > 48 * 1024) {
    static size_t allowed[64];                     // per instantiation and device (function attributes are per device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > allowed[dev & 63]) {
      if (cudaFuncSetAttribute(k_conv_direct<COUT, KS, IS3D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return false;
      }
      allowed[dev & 63] = smem;
    }
  }
  k_conv_direct<COUT, KS, IS3D><<<grid, block, smem, st>>>(in, out, w, b, cin, act, g);
  return true;
}

int launch_conv_direct(const float* in, float* out, const float* wdev, const float* bdev, int cin, int cout,
                       int ksize, int act, const Geo& g, cudaStream_t st) {
#define TFL_CONV_CASE(CO, KS_)                                                        \
  if (cout == CO && ksize == KS_) {                                                   \
    const bool ok_ = g.is3d ? conv_launch<CO, KS_, true>(in, out, wdev, bdev, cin, act, g, st)    \
                            : conv_launch<CO, KS_, false>(in, out, wdev, bdev, cin, act, g, st);  \
    if (ok_) return 1;                                                                \
  }
  TFL_CONV_CASE(8, 3) TFL_CONV_CASE(8, 1) TFL_CONV_CASE(1, 1) TFL_CONV_CASE(16, 3) TFL_CONV_CASE(16, 1)
  TFL_CONV_CASE(1, 3) TFL_CONV_CASE(6, 3) TFL_CONV_CASE(6, 1) TFL_CONV_CASE(32, 1) TFL_CONV_CASE(16, 5)
  TFL_CONV_CASE(32, 5) TFL_CONV_CASE(64, 5) TFL_CONV_CASE(64, 1)
#undef TFL_CONV_CASE
  if (g.zlo != 0 || g.zhi != g.nz || g.zoff != 0) return -1;      // the generic kernel works on whole grids only
  const long long total = (long long)g.nb * cout * g.n;
  k_conv_any<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, wdev, bdev, cin, cout, ksize, act, g);
  return 1;
}

void launch_pool(const float* in, float* out, int nbc, int nz, int ny, int nx, int p, int is3d, int is_max,
                 cudaStream_t st) {
  const int pz = is3d ? p : 1;
  const long long total = (long long)nbc * (nz / pz) * (ny / p) * (nx / p);
  k_pool<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, nz, ny, nx, p, is3d, is_max, total);
}
void launch_pixel_shuffle(const float* in, float* out, int nb, int n_out, int nz, int ny, int nx, int s, int is3d,
                          cudaStream_t st) {
  const long long total = (long long)nb * n_out * nz * (is3d ? s : 1) * ny * s * nx * s;
  k_pixel_shuffle<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, n_out, nz, ny, nx, s, is3d, total);
}

}  // namespace tfl
