// Device functions shared by the two implementations of advectVel (tfl_stencils.cu: one kernel per
// MacCormack pass on global memory; tfl_advect_tile.cu: one kernel over shared-memory tiles).
// Included only by translation units compiled with -fmad=false.
#pragma once
#include "tfl_device.cuh"

namespace tfl {

// vel[c]: velocity at the centre of face c of cell (i, j, k) (third_party/grid.cc:374-417).
static __device__ __forceinline__ void mac_face_velocities(const float* __restrict__ ub, const Geo& g, int k, int j, int i,
                                                    V3 (&vel)[3]) {
  vel[0] = mac_at_x(ub, g, k, j, i);
  vel[1] = mac_at_y(ub, g, k, j, i);
  vel[2] = g.is3d ? mac_at_z(ub, g, k, j, i) : V3{0.0f, 0.0f, 0.0f};
}


// MacCormackClampMAC for one component (third_party/tfluids.cc:701-774): min / max of the original
// field over the two 2x2x2 boxes around (i, j, k) -+ vel, `fwd` when a box leaves the grid.
static __device__ __noinline__ float clamp_component_mac(const float* __restrict__ orig_c, const Geo& g,
                                                  float val, float fwd, int kglob, int j, int i, V3 vel) {
  const float fi = (float)i, fj = (float)j, fk = (float)kglob;
  float lo = FLT_MAX, hi = -FLT_MAX;
  for (int l = 0; l < 2; l++) {
    const int px = l == 0 ? (int)(fi - vel.x) : (int)(fi + vel.x);
    const int py = l == 0 ? (int)(fj - vel.y) : (int)(fj + vel.y);
    const int pz = l == 0 ? (int)(fk - vel.z) : (int)(fk + vel.z);
    const int i0 = clamp_i(px, 0, g.nx - 2), j0 = clamp_i(py, 0, g.ny - 2);
    const int k0 = clamp_i(pz, 0, g.is3d ? g.gnz - 2 : 1);
    const int i1 = i0 + 1, j1 = j0 + 1, k1 = g.is3d ? k0 + 1 : k0;
    bool inb = i0 >= 0 && j0 >= 0 && i1 < g.nx && j1 < g.ny;
    if (g.is3d) inb = inb && k0 >= 0 && k1 < g.gnz; else inb = inb && k0 == 0 && k1 == 0;
    if (!inb) return fwd;
    const int kl0 = local_z(g, k0), kl1 = g.is3d ? local_z(g, k1) : kl0;
// The eight corners sit at fixed offsets from the first one (same visiting order as the reference).
    const float* a0 = orig_c + cell(g, kl0, j0, i0);
    const float* a1 = a0 + g.nx;
#define TFL_MM(ptr, off) { const float t = __ldg((ptr) + (off)); if (t < lo) lo = t; if (t > hi) hi = t; }
    TFL_MM(a0, 0) TFL_MM(a0, 1) TFL_MM(a1, 0) TFL_MM(a1, 1)
    if (g.is3d) {
      const float* b0 = a0 + (kl1 - kl0) * g.ny * g.nx;
      const float* b1 = b0 + g.nx;
      TFL_MM(b0, 0) TFL_MM(b0, 1) TFL_MM(b1, 0) TFL_MM(b1, 1)
    }
#undef TFL_MM
  }
  return clamp_f(val, lo, hi);
}


// MacCormackClamp of the "Ours" scalar advection (third_party/tfluids.cc:331-413): clamp v to the min / max
// of the source field over the (fluid, unless sampleOutsideFluid) cells of the 3x3x3 neighbourhood of the
// cell that holds the forward trace's end point (px, py, pz); fw when no cell qualifies.  cl: clearance field
// of the grid or nullptr.
template <typename FT>
__device__ __forceinline__ float clamp_scalar_ours(const float* __restrict__ sb, const FT* __restrict__ fl,
                                                   const unsigned char* __restrict__ cl, const Geo& g, float v,
                                                   float fw, float px, float py, float pz, bool outside) {
    const int i0 = clamp_i((int)px, 0, g.nx - 1), j0 = clamp_i((int)py, 0, g.ny - 1);
    const int k0 = g.is3d ? clamp_i((int)pz, 0, g.gnz - 1) : 0;
    float lo = INFINITY, hi = -INFINITY;
    int found = 0;
    const int kl0 = k0 - g.zoff;
    const bool interior = i0 >= 1 && i0 <= g.nx - 2 && j0 >= 1 && j0 <= g.ny - 2 &&
                          (!g.is3d || (k0 >= 1 && k0 <= g.gnz - 2 && kl0 >= 1 && kl0 <= g.nz - 2));
    if (interior) {
      // Same cells in the same order as the general loop below, without its bounds tests: the
      // 27 (9 in 2-D) neighbours sit at fixed offsets from the centre.
      const int ctr = cell(g, g.is3d ? kl0 : 0, j0, i0);
      const int sy = g.nx, sz = g.nx * g.ny;
      // clearance >= 2 at the centre of the neighbourhood: all of it is fluid, no flag is read
      const bool all_fluid = outside || (cl && __ldg(cl + ctr) > 1);
      if (all_fluid) {
        found = 1;
#pragma unroll
        for (int dz = -1; dz <= 1; dz++) {
          if (!g.is3d && dz != 0) continue;
#pragma unroll
          for (int dy = -1; dy <= 1; dy++) {
            const float* srow = sb + (ctr + dz * sz + dy * sy);
#pragma unroll
            for (int dx = -1; dx <= 1; dx++) {
              const float t = __ldg(srow + dx);
              lo = (t < lo) ? t : lo;
              hi = (t > hi) ? t : hi;
            }
          }
        }
      } else {
#pragma unroll
      for (int dz = -1; dz <= 1; dz++) {
        if (!g.is3d && dz != 0) continue;
#pragma unroll
        for (int dy = -1; dy <= 1; dy++) {
          // one pointer pair per row: the three cells of a row are immediate offsets -1, 0, +1
          const float* srow = sb + (ctr + dz * sz + dy * sy);
          const FT* frow = fl + (ctr + dz * sz + dy * sy);
#pragma unroll
          for (int dx = -1; dx <= 1; dx++) {
            const float t = __ldg(srow + dx);
            const bool use = (flag_at(frow, dx) & kFluid);
            lo = (use && t < lo) ? t : lo;
            hi = (use && t > hi) ? t : hi;
            found |= use ? 1 : 0;
          }
        }
      }
      }
    } else {
    for (int kk = k0 - 1; kk <= k0 + 1; kk++) {
      if (kk < 0 || kk >= g.gnz) continue;
      const int kl = local_z(g, kk);
      for (int jj = j0 - 1; jj <= j0 + 1; jj++) {
        if (jj < 0 || jj >= g.ny) continue;
        for (int ii = i0 - 1; ii <= i0 + 1; ii++) {
          if (ii < 0 || ii >= g.nx) continue;
          if (outside || (flag_i(fl, g, kl, jj, ii) & kFluid)) {
            const float t = __ldg(sb + cell(g, kl, jj, ii));
            if (t < lo) lo = t;
            if (t > hi) hi = t;
            found++;
          }
        }
      }
    }
    }
    return (found < 1) ? fw : clamp_f(v, lo, hi);
}

}  // namespace tfl
