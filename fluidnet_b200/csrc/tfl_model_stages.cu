// Non-convolutional nodes of the projection graph (torch/lib/model.lua:27-401), compiled
// with -fmad=false so they match the float arithmetic of the reference modules:
//   tfluids.SetWallBcs      tfluids/set_wall_bcs.lua:29-48   (U * {0,1} mask)
//   tfluids.VelocityDivergence, nn.StandardDeviation (lib/modules/variance.lua:44-76),
//   nn.Clamp, nn.ApplyScale (lib/modules/apply_scale.lua), tfluids.FlagsToOccupancy,
//   tfluids.VelocityUpdate  tfluids/velocity_update.lua:28-38
#include "tfl_device.cuh"
#include "tfl_kernels.h"

namespace tfl {

// U1 = U * wallmask; sums[b] += (sum U1, sum U1^2) over the launch range (double).
template <bool IS3D, typename FT>
__global__ void k_cnn_mask_stats(const float* __restrict__ U, const FT* __restrict__ flags,
                                 float* __restrict__ U1, double* __restrict__ sums, int own_lo, int own_hi, Geo gin) {
  const Geo g = static_geo<IS3D>(gin);
  int b, k, j, i;
  const bool live = thread_cell(g, b, k, j, i);
  double s = 0.0, ss = 0.0;
  if (live) {
    bool z[3];
    wall_bc_zero_mask(flags + b * g.n, g, k, j, i, z);
    const long long c = (long long)b * g.nc * g.n + cell(g, k, j, i);
    for (int a = 0; a < g.nc; a++) {
      float u = U[c + a * g.n];
      if (z[a]) u = u * 0.0f;
      U1[c + a * g.n] = u;
      if (k >= own_lo && k < own_hi) {      // the statistics cover owned planes only (z-slabs)
        const float sq = u * u;
        s += (double)u;
        ss += (double)sq;
      }
    }
  }
  // One block never straddles two batch entries unless nb > 1 and the z range is odd; in that
  // (rare) case fall back to per-thread atomics.
  const int nzr = g.zhi - g.zlo;
  const bool block_uniform = (g.nb == 1) || (nzr % (int)blockDim.z == 0);
  if (block_uniform) {
    const int zz0 = blockIdx.z * blockDim.z;
    block_accumulate(s, ss, sums + 2 * (zz0 / nzr < g.nb ? zz0 / nzr : 0));
  } else if (live) {
    atomicAdd(sums + 2 * b, s);
    atomicAdd(sums + 2 * b + 1, ss);
  }
}

// nn.StandardDeviation + nn.Clamp(threshold, inf): float tensor ops on the two sums.
__global__ void k_cnn_scale(const double* __restrict__ sums, float* __restrict__ scale, int nb,
                            long long n, float threshold) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  scale[b] = scale_from_sums(sums, b, n, threshold);
}

// x0 = [pDiv / scale, div(U1) / scale, occupancy(flags)]
template <bool IS3D, typename FT>
__global__ void k_cnn_inputs(const float* __restrict__ p_div, const float* __restrict__ U1,
                             const FT* __restrict__ flags, const float* __restrict__ scale,
                             float* __restrict__ x0, Geo gin) {
  const Geo g = static_geo<IS3D>(gin);
  int b, k, j, i;
  if (!thread_cell(g, b, k, j, i)) return;
  const int c = cell(g, k, j, i);
  const float sc = __ldg(scale + b);
  const float* ub = U1 + (long long)b * g.nc * g.n;
  const int f = flag_i(flags + b * g.n, g, k, j, i);
  float dv = 0.0f;
  if (!on_border(g, k, j, i) && (f & kFluid)) {
    dv = __ldg(ub + c) - __ldg(ub + c + 1) + __ldg(ub + g.n + c) - __ldg(ub + g.n + c + g.nx);
    if (g.is3d) dv += (__ldg(ub + 2 * g.n + c) - __ldg(ub + 2 * g.n + c + (long long)g.nx * g.ny));
  }
  float* xb = x0 + (long long)b * 3 * g.n + c;
  xb[0] = __ldg(p_div + b * g.n + c) / sc;
  xb[g.n] = dv / sc;
  xb[2 * g.n] = (f == kFluid) ? 0.0f : ((f == kObstacle) ? 1.0f : -1.0f);
}

// Same three input channels written as the first channels-last float4 plane of the padded
// activation layout the tensor-core convolution reads (tfl_cnn_tc.cu): (pDiv/s, div/s, occ, 0).
template <bool IS3D, typename FT>
__global__ void k_cnn_inputs_padded(const float* __restrict__ p_div, const float* __restrict__ U1,
                                    const FT* __restrict__ flags, const float* __restrict__ scale,
                                    float4* __restrict__ x0, int px, int py, Geo gin) {
  const Geo g = static_geo<IS3D>(gin);
  int b, k, j, i;
  if (!thread_cell(g, b, k, j, i)) return;
  const int c = cell(g, k, j, i);
  const float sc = __ldg(scale + b);
  const float* ub = U1 + (long long)b * g.nc * g.n;
  const int f = flag_i(flags + b * g.n, g, k, j, i);
  float dv = 0.0f;
  if (!on_border(g, k, j, i) && (f & kFluid)) {
    dv = __ldg(ub + c) - __ldg(ub + c + 1) + __ldg(ub + g.n + c) - __ldg(ub + g.n + c + g.nx);
    if (g.is3d) dv += (__ldg(ub + 2 * g.n + c) - __ldg(ub + 2 * g.n + c + (long long)g.nx * g.ny));
  }
  const long long plane = (long long)(g.nz + 2) * py * px;
  const long long o = (long long)b * 2 * plane + ((long long)(k + 1) * py + (j + 1)) * px + (i + 1);
  x0[o] = make_float4(__ldg(p_div + b * g.n + c) / sc, dv / sc,
                      (f == kFluid) ? 0.0f : ((f == kObstacle) ? 1.0f : -1.0f), 0.0f);
}

// U = setWallBcs(velocityUpdate(U1 / scale, p_net) * scale);  p = p_net * scale.
template <bool IS3D, typename FT>
__global__ void k_cnn_finish(const float* __restrict__ p_net, const float* __restrict__ U1,
                             const FT* __restrict__ flags, const float* __restrict__ scale,
                             float* __restrict__ p_out, float* __restrict__ U_out, Geo gin) {
  const Geo g = static_geo<IS3D>(gin);
  int b, k, j, i;
  if (!thread_cell(g, b, k, j, i)) return;
  const int c = cell(g, k, j, i);
  const float sc = __ldg(scale + b);
  const FT* fl = flags + b * g.n;
  const float* pb = p_net + b * g.n;
  const float* ub = U1 + (long long)b * g.nc * g.n;
  const float pc = __ldg(pb + c);
  float u[3];
  for (int a = 0; a < g.nc; a++) u[a] = __ldg(ub + a * g.n + c) / sc;
  if (!on_border(g, k, j, i)) {
    const int st[3] = {1, g.nx, g.nx * g.ny};
    const int fc = flag_i(fl, g, k, j, i);
    int fn[3];
    fn[0] = flag_i(fl, g, k, j, i - 1);
    fn[1] = flag_i(fl, g, k, j - 1, i);
    fn[2] = g.is3d ? flag_i(fl, g, k - 1, j, i) : 0;
    if (fc & kFluid) {
      for (int a = 0; a < g.nc; a++) {
        if (fn[a] & kFluid) u[a] -= (pc - __ldg(pb + c - st[a]));
        if (fn[a] & kEmpty) u[a] -= pc;
      }
    } else if ((fc & kEmpty) && !(fc & kOutflow)) {
      for (int a = 0; a < g.nc; a++) {
        if (fn[a] & kFluid) u[a] += __ldg(pb + c - st[a]);
        else u[a] = 0.0f;
      }
    }
  }
  bool z[3];
  wall_bc_zero_mask(fl, g, k, j, i, z);
  float* uo = U_out + (long long)b * g.nc * g.n + c;
  for (int a = 0; a < g.nc; a++) {
    float v = u[a] * sc;
    if (z[a]) v = v * 0.0f;
    uo[a * g.n] = v;
  }
  p_out[b * g.n + c] = pc * sc;
}

#define TFL_LAUNCH3B(kernel, g, st, ...)                                        \
  do {                                                                          \
    dim3 grid_, block_;                                                         \
    launch_dims(g, grid_, block_);                                              \
    if ((g).is3d) kernel<true, float><<<grid_, block_, 0, st>>>(__VA_ARGS__);   \
    else kernel<false, float><<<grid_, block_, 0, st>>>(__VA_ARGS__);           \
  } while (0)

void launch_cnn_mask_stats(const float* U, const float* flags, float* U1, double* sums, int own_lo, int own_hi,
                           const Geo& g, cudaStream_t st) {
  TFL_LAUNCH3B(k_cnn_mask_stats, g, st, U, flags, U1, sums, own_lo, own_hi, g);
}
void launch_cnn_scale(const double* sums, float* scale, int nb, long long n_per_batch, float threshold,
                      cudaStream_t st) {
  k_cnn_scale<<<(nb + 31) / 32, 32, 0, st>>>(sums, scale, nb, n_per_batch, threshold);
}
void launch_cnn_inputs(const float* p_div, const float* U1, const float* flags, const float* scale,
                       float* x0, const Geo& g, cudaStream_t st) {
  TFL_LAUNCH3B(k_cnn_inputs, g, st, p_div, U1, flags, scale, x0, g);
}
void launch_cnn_inputs_padded(const float* p_div, const float* U1, const float* flags, const float* scale,
                              float* x0, int px, int py, const Geo& g, cudaStream_t st) {
  TFL_LAUNCH3B(k_cnn_inputs_padded, g, st, p_div, U1, flags, scale, (float4*)x0, px, py, g);
}
void launch_cnn_finish(const float* p_net, const float* U1, const float* flags, const float* scale,
                       float* p_out, float* U_out, const Geo& g, cudaStream_t st) {
  TFL_LAUNCH3B(k_cnn_finish, g, st, p_net, U1, flags, scale, p_out, U_out, g);
}

}  // namespace tfl
