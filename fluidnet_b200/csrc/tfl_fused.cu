// Fused variants of the point-wise stages of tfluids.simulate (torch/lib/simulate.lua:175-327)
// used by tfl_simulate_step on the convnet path.  Each kernel performs, per cell and in the
// SAME floating-point order, what a run of separate operators / cutorch calls does in the
// reference loop, so results are bit-identical to the operator-by-operator sequence while
// the velocity / density fields cross HBM once instead of once per operator.
// Compiled with -fmad=false.
//
//   k_post_advect        s:copy(tmp) ; U:copy(tmp)            (tfluids/init.lua:145-148, :215-218)
//                        setConstVals (U, density)            (lib/simulate.lua:202, :130-160)
//                        tfluids.addBuoyancy                  (lib/simulate.lua:216-226)
//   k_vort_bc_mask       vorticityConfinement's AddForceField (third_party/tfluids.cc:1312-1339)
//                        setConstVals (U)                     (lib/simulate.lua:252)
//                        tfluids.SetWallBcs mask multiply + the sums behind nn.StandardDeviation
//                                                             (lib/model.lua:81-117)
//   k_cnn_inputs_fused   scale, ApplyScale(pDiv, div), VelocityDivergence, FlagsToOccupancy
//                                                             (lib/model.lua:86-150)
//   k_cnn_finish_fused   VelocityUpdate, ApplyScale, SetWallBcs (lib/model.lua:380-390)
//                        setConstVals (U) ; U:clamp(-1e6, 1e6) (lib/simulate.lua:321, :326)
#include "tfl_device.cuh"
#include "tfl_kernels.h"

namespace tfl {

struct BcPtrs {
  const float* u_inv; const float* u_bc;        // may be null (no velocity BC)
  const float* d_inv; const float* d_bc;        // may be null (no density BC)
};

__device__ __forceinline__ float bc_apply(float x, const float* __restrict__ inv, const float* __restrict__ bc,
                                          long long o) {
  if (inv == nullptr) return x;
  const float t = x * __ldg(inv + o);
  return t + __ldg(bc + o);
}

// density = BC(tmp_s); U = BC(tmp_U) (+ buoyancy with the NEW density).
template <bool IS3D, typename FT>
__global__ void k_post_advect(const float* __restrict__ tmp_s, const float* __restrict__ tmp_u,
                              const FT* __restrict__ flags, float* __restrict__ density,
                              float* __restrict__ U, BcPtrs bc, int do_buoy, float sx, float sy_, float sz_,
                              Geo gin) {
  const Geo g = static_geo<IS3D>(gin);
  int b, k, j, i;
  if (!thread_cell(g, b, k, j, i)) return;
  const int c = cell(g, k, j, i);
  const long long sb = b * g.n, ub = (long long)b * g.nc * g.n;
  float rc = 0.0f;
  if (density) {
    rc = bc_apply(__ldg(tmp_s + sb + c), bc.d_inv, bc.d_bc, sb + c);
    density[sb + c] = rc;
  }
  float u[3];
  for (int a = 0; a < g.nc; a++) u[a] = bc_apply(__ldg(tmp_u + ub + a * g.n + c), bc.u_inv, bc.u_bc, ub + a * g.n + c);
  if (do_buoy && density && !on_border(g, k, j, i)) {
    const FT* fl = flags + sb;
    if (flag_i(fl, g, k, j, i) & kFluid) {
      const int st[3] = {1, g.nx, g.nx * g.ny};
      const float str[3] = {sx, sy_, sz_};
      int fn[3];
      fn[0] = flag_i(fl, g, k, j, i - 1);
      fn[1] = flag_i(fl, g, k, j - 1, i);
      fn[2] = g.is3d ? flag_i(fl, g, k - 1, j, i) : 0;
      for (int a = 0; a < g.nc; a++) {
        if (fn[a] & kFluid) {
          const float rn = bc_apply(__ldg(tmp_s + sb + c - st[a]), bc.d_inv, bc.d_bc, sb + c - st[a]);
          u[a] += (0.5f * str[a] * (rc + rn));
        }
      }
    }
  }
  for (int a = 0; a < g.nc; a++) U[ub + a * g.n + c] = u[a];
}

// U += confinement force; U = BC(U); then (mask_mode 1) U *= wall mask and accumulate the
// sums for the input scale.
template <bool IS3D, typename FT>
__global__ void k_vort_bc_mask(float* __restrict__ U, const FT* __restrict__ flags,
                               const float* __restrict__ force, int do_vort, BcPtrs bc, int mask_mode,
                               double* __restrict__ sums, Geo gin) {
  const Geo g = static_geo<IS3D>(gin);
  int b, k, j, i;
  const bool live = thread_cell(g, b, k, j, i);
  double s = 0.0, ss = 0.0;
  if (live) {
    const int c = cell(g, k, j, i);
    const long long ub = (long long)b * g.nc * g.n;
    const FT* fl = flags + b * g.n;
    float u[3];
    for (int a = 0; a < g.nc; a++) u[a] = U[ub + a * g.n + c];
    if (do_vort && !on_border(g, k, j, i)) {
      const int fc = flag_i(fl, g, k, j, i);
      const bool cf = fc & kFluid, ce = fc & kEmpty;
      if (cf || ce) {
        const float* fb = force + (long long)b * 3 * g.n;
        int f = flag_i(fl, g, k, j, i - 1);
        if ((f & kFluid) || (cf && (f & kEmpty))) u[0] += (0.5f * (__ldg(fb + c - 1) + __ldg(fb + c)));
        f = flag_i(fl, g, k, j - 1, i);
        if ((f & kFluid) || (cf && (f & kEmpty))) u[1] += (0.5f * (__ldg(fb + g.n + c - g.nx) + __ldg(fb + g.n + c)));
        if (g.is3d) {
          f = flag_i(fl, g, k - 1, j, i);
          if ((f & kFluid) || (cf && (f & kEmpty)))
            u[2] += (0.5f * (__ldg(fb + 2 * g.n + c - g.nx * g.ny) + __ldg(fb + 2 * g.n + c)));
        }
      }
    }
    for (int a = 0; a < g.nc; a++) u[a] = bc_apply(u[a], bc.u_inv, bc.u_bc, ub + a * g.n + c);
    if (mask_mode) {
      bool z[3];
      wall_bc_zero_mask(fl, g, k, j, i, z);
      for (int a = 0; a < g.nc; a++) {
        if (z[a]) u[a] = u[a] * 0.0f;
        const float sq = u[a] * u[a];
        s += (double)u[a];
        ss += (double)sq;
      }
    }
    for (int a = 0; a < g.nc; a++) U[ub + a * g.n + c] = u[a];
  }
  if (mask_mode) {
    // nb == 1 in the fused step (one grid per context); a block never straddles batches then.
    block_accumulate(s, ss, sums + 2 * (live ? b : 0));
  }
}

// First channels-last plane of the conv input: (pDiv/s, div(U1)/s, occupancy, 0); U1 is the
// already masked velocity left in U by k_vort_bc_mask.
template <bool IS3D, typename FT>
__global__ void k_cnn_inputs_fused(const float* __restrict__ p_div, const float* __restrict__ U1,
                                   const FT* __restrict__ flags, const double* __restrict__ sums,
                                   float threshold, float* __restrict__ scale_out, float4* __restrict__ x0,
                                   int px, int py, Geo gin) {
  const Geo g = static_geo<IS3D>(gin);
  int b, k, j, i;
  if (!thread_cell(g, b, k, j, i)) return;
  const int c = cell(g, k, j, i);
  const float sc = scale_from_sums(sums, b, (long long)g.nc * g.n, threshold);
  if (c == 0) scale_out[b] = sc;
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

// U = clamp(BC(setWallBcs(velocityUpdate(U1 / s, p_net) * s)));  p = p_net * s.  In place on U.
template <bool IS3D, typename FT>
__global__ void k_cnn_finish_fused(const float* __restrict__ p_net, float* __restrict__ U,
                                   const FT* __restrict__ flags, const float* __restrict__ scale,
                                   float* __restrict__ p_out, BcPtrs bc, float lo, float hi, Geo gin) {
  const Geo g = static_geo<IS3D>(gin);
  int b, k, j, i;
  if (!thread_cell(g, b, k, j, i)) return;
  const int c = cell(g, k, j, i);
  const float sc = __ldg(scale + b);
  const FT* fl = flags + b * g.n;
  const float* pb = p_net + b * g.n;
  const long long ub = (long long)b * g.nc * g.n;
  const float pc = __ldg(pb + c);
  float u[3];
  for (int a = 0; a < g.nc; a++) u[a] = U[ub + a * g.n + c] / sc;
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
  for (int a = 0; a < g.nc; a++) {
    float v = u[a] * sc;
    if (z[a]) v = v * 0.0f;
    v = bc_apply(v, bc.u_inv, bc.u_bc, ub + a * g.n + c);
    v = (v < lo) ? lo : ((v > hi) ? hi : v);
    U[ub + a * g.n + c] = v;
  }
  p_out[b * g.n + c] = pc * sc;
}

#define TFL_LAUNCH3F(kernel, g, st, ...)                                                \
  do {                                                                                  \
    dim3 grid_, block_;                                                                 \
    launch_dims(g, grid_, block_);                                                      \
    if ((g).is3d) kernel<true, unsigned char><<<grid_, block_, 0, st>>>(__VA_ARGS__);   \
    else kernel<false, unsigned char><<<grid_, block_, 0, st>>>(__VA_ARGS__);           \
  } while (0)

void launch_post_advect(const float* tmp_s, const float* tmp_u, const unsigned char* flags, float* density, float* U,
                        const float* u_inv, const float* u_bc, const float* d_inv, const float* d_bc,
                        int do_buoy, const float s[3], const Geo& g, cudaStream_t st) {
  BcPtrs bc{u_inv, u_bc, d_inv, d_bc};
  TFL_LAUNCH3F(k_post_advect, g, st, tmp_s, tmp_u, flags, density, U, bc, do_buoy, s[0], s[1], s[2], g);
}
void launch_vort_bc_mask(float* U, const unsigned char* flags, const float* force, int do_vort,
                         const float* u_inv, const float* u_bc, int mask_mode, double* sums,
                         const Geo& g, cudaStream_t st) {
  BcPtrs bc{u_inv, u_bc, nullptr, nullptr};
  TFL_LAUNCH3F(k_vort_bc_mask, g, st, U, flags, force, do_vort, bc, mask_mode, sums, g);
}
void launch_cnn_inputs_fused(const float* p_div, const float* U1, const unsigned char* flags, const double* sums,
                             float threshold, float* scale_out, float* x0, int px, int py, const Geo& g,
                             cudaStream_t st) {
  TFL_LAUNCH3F(k_cnn_inputs_fused, g, st, p_div, U1, flags, sums, threshold, scale_out, (float4*)x0, px, py, g);
}
void launch_cnn_finish_fused(const float* p_net, float* U, const unsigned char* flags, const float* scale, float* p_out,
                             const float* u_inv, const float* u_bc, float lo, float hi, const Geo& g,
                             cudaStream_t st) {
  BcPtrs bc{u_inv, u_bc, nullptr, nullptr};
  TFL_LAUNCH3F(k_cnn_finish_fused, g, st, p_net, U, flags, scale, p_out, bc, lo, hi, g);
}

}  // namespace tfl
