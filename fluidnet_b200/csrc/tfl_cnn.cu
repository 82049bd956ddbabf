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

// weights in smem as [cin][tap][COUT]; taps ordered (dz, dy, dx).
template <int COUT, int KS, bool IS3D>
__global__ void __launch_bounds__(256)
k_conv_direct(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ w,
              const float* __restrict__ bias, int cin, int relu, Geo g) {
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
    float r = acc[o];
    if (relu && r < 0.0f) r = 0.0f;
    out[((long long)b * COUT + o) * g.n + c0] = r;
  }
}

template <int COUT, int KS, bool IS3D>
static void conv_launch(const float* in, float* out, const float* w, const float* b, int cin, int relu,
                        const Geo& g, cudaStream_t st) {
  const int nzr = g.zhi - g.zlo;
  dim3 block = IS3D ? dim3(32, 4, 2) : dim3(32, 8, 1);
  dim3 grid((g.nx + block.x - 1) / block.x, (g.ny + block.y - 1) / block.y,
            ((long long)g.nb * nzr + block.z - 1) / block.z);
  constexpr int KZ = IS3D ? KS : 1;
  const size_t smem = sizeof(float) * cin * KZ * KS * KS * COUT;
  k_conv_direct<COUT, KS, IS3D><<<grid, block, smem, st>>>(in, out, w, b, cin, relu, g);
}

int launch_conv_direct(const float* in, float* out, const float* wdev, const float* bdev, int cin, int cout,
                       int ksize, int relu, const Geo& g, cudaStream_t st) {
#define TFL_CONV_CASE(CO, KS_)                                                        \
  if (cout == CO && ksize == KS_) {                                                   \
    if (g.is3d) conv_launch<CO, KS_, true>(in, out, wdev, bdev, cin, relu, g, st);    \
    else conv_launch<CO, KS_, false>(in, out, wdev, bdev, cin, relu, g, st);          \
    return 1;                                                                         \
  }
  TFL_CONV_CASE(8, 3) TFL_CONV_CASE(8, 1) TFL_CONV_CASE(1, 1) TFL_CONV_CASE(16, 3) TFL_CONV_CASE(16, 1)
  TFL_CONV_CASE(1, 3) TFL_CONV_CASE(6, 3) TFL_CONV_CASE(6, 1) TFL_CONV_CASE(32, 1) TFL_CONV_CASE(16, 5)
  TFL_CONV_CASE(32, 5) TFL_CONV_CASE(64, 5) TFL_CONV_CASE(64, 1)
#undef TFL_CONV_CASE
  return -1;
}

}  // namespace tfl
