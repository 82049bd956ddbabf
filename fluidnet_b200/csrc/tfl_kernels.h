// Host-callable launchers of the CUDA kernels (internal to libtfl; the public surface is
// include/tfl.h).
#pragma once
#include <cuda_runtime.h>
#include "../../include/tfl.h"
#include "tfl_device.cuh"

namespace tfl {

// ---- tfl_stencils.cu (-fmad=false) ----
void launch_empty_domain(float* flags, const Geo& g, int bnd, cudaStream_t st);
void launch_flags_to_occupancy(const float* flags, float* occ, long long n, unsigned long long* bad,
                               cudaStream_t st);
template <typename FT>
void launch_set_wall_bcs(float* U, const FT* flags, const Geo& g, int as_mask, cudaStream_t st);
template <typename FT>
void launch_divergence(const float* U, const FT* flags, float* div, const Geo& g, cudaStream_t st);
template <typename FT>
void launch_velocity_update(float* U, const FT* flags, const float* p, const Geo& g, cudaStream_t st);
template <typename FT>
void launch_add_buoyancy(float* U, const FT* flags, const float* rho, const float s[3], const Geo& g,
                         cudaStream_t st);
template <typename FT>
void launch_add_gravity(float* U, const FT* flags, const float f[3], const Geo& g, cudaStream_t st);
template <typename FT>
int launch_vorticity(float* U, const FT* flags, float strength, float* curl, float* cnorm, float* force,
                     const Geo& g, cudaStream_t st);
// g: range of the result; g_fwd: (wider, in slab mode) range of the forward pass.
// clear: clearance field of the same grid (launch_clearance) or nullptr (general code everywhere).
template <typename FT>
int launch_advect_scalar(float dt, const float* s, const float* U, const FT* flags, const unsigned char* clear,
                         int method, int outside, float strength, float* dst, float* fwd, float* fwd_pos,
                         const Geo& g, const Geo& g_fwd, cudaStream_t st);
template <typename FT>
int launch_advect_vel(float dt, const float* U, const FT* flags, const unsigned char* clear, int method,
                      float strength, float* dst, float* fwd, const Geo& g, const Geo& g_fwd, cudaStream_t st);
// ---- tfl_advect_tile.cu (-fmad=false): maccormackOurs advectVel as one kernel over shared-memory tiles ----
// hf: halo of the forward field in cells (1: traces < ~0.5 cell are served from the tile, 2: < ~1.5; longer ones
// take the general code).  longest: optional device word, atomicMax of the longest trace (float bits).
// Returns false when the grid does not qualify (2-D, batch, slab, nx % 4 != 0): run launch_advect_vel then.
bool launch_advect_vel_tile(float dt, const float* U, const unsigned char* flags, const unsigned char* clear,
                            float strength, float* dst, const Geo& g, int hf, int variant, unsigned int* longest,
                            cudaStream_t st);
// advectScalar('maccormackOurs') on the same tiles.
bool launch_advect_scalar_tile(float dt, const float* src, const float* U, const unsigned char* flags,
                               const unsigned char* clear, int outside, float strength, float* dst, const Geo& g, int hf,
                               int variant, cudaStream_t st);
// Clearance of every cell of the local storage (advection fast path, tfl_device.cuh); tmp: scratch of the
// same size; gate: optional device word, the kernels do nothing when it is 0.  Returns the launch count.
template <typename FT>
int launch_clearance(const FT* flags, unsigned char* clear, unsigned char* tmp, const Geo& g, const int* gate,
                     cudaStream_t st);
template <typename FT>
void launch_jacobi_mask(const FT* flags, unsigned char* mask, const Geo& g, cudaStream_t st);
void launch_jacobi_iter(const unsigned char* mask, const float* div, const float* prev, float* cur,
                        const Geo& g, cudaStream_t st);
bool launch_jacobi_sweeps(const unsigned char* mask, const float* div, float* pa, float* pb, const Geo& g, int sweeps,
                          cudaStream_t st);
void launch_sqdiff(const float* a, const float* b, long long n, int nb, double* out, cudaStream_t st);
// changed: optional device word that is OR-ed with 1 when a byte differs from what `o` held before.
void launch_flags_to_u8(const float* f, unsigned char* o, long long n, int* changed, cudaStream_t st);
void launch_apply_bc(float* x, const float* inv, const float* bc, long long n, cudaStream_t st);
void launch_clamp(float* x, float lo, float hi, long long n, cudaStream_t st);

// CNN pre/post stages (exact float semantics of lib/model.lua's non-conv nodes).
void launch_cnn_mask_stats(const float* U, const float* flags, float* U1, double* sums, int own_lo, int own_hi,
                           const Geo& g, cudaStream_t st);
void launch_cnn_scale(const double* sums, float* scale, int nb, long long n_per_batch, float threshold,
                      cudaStream_t st);
void launch_cnn_inputs(const float* p_div, const float* U1, const float* flags, const float* scale,
                       float* x0, const Geo& g, cudaStream_t st);
void launch_cnn_inputs_padded(const float* p_div, const float* U1, const float* flags, const float* scale,
                              float* x0, int px, int py, const Geo& g, cudaStream_t st);
void launch_cnn_finish(const float* p_net, const float* U1, const float* flags, const float* scale,
                       float* p_out, float* U_out, const Geo& g, cudaStream_t st);

// ---- tfl_fused.cu (-fmad=false): fused point-wise stages of the convnet step ----
// qmask (may be null): BcPtrs::qmask of tfl_fused.cu, filled by launch_bc_quad_mask for the same step.
bool launch_bc_quad_mask(const float* u_inv, const float* u_bc, const float* d_inv, const float* d_bc,
                         unsigned char* qmask, const Geo& g, cudaStream_t st);
void launch_post_advect(const float* tmp_s, const float* tmp_u, const unsigned char* flags, float* density, float* U,
                        const float* u_inv, const float* u_bc, const float* d_inv, const float* d_bc,
                        const unsigned char* qmask, int do_buoy, const float s[3], const Geo& g, cudaStream_t st);
void launch_vort_curl(const float* U, float* curl, float* cnorm, float* force, float strength, const Geo& g,
                      cudaStream_t st);
bool launch_vort_curl_quad(const float* U, float* curl, float* cnorm, float* force, float strength, const Geo& g,
                           cudaStream_t st);
void launch_vort_bc_mask(float* U, const unsigned char* flags, const float* force, int do_vort, const float* u_inv, const float* u_bc, const unsigned char* qmask, int mask_mode, double* sums,
                         const Geo& g, cudaStream_t st);
void launch_cnn_inputs_fused(const float* p_div, const float* U1, const unsigned char* flags, const double* sums,
                             float threshold, float* scale_out, float* x0, int px, int py, const Geo& g,
                             cudaStream_t st);
void launch_cnn_finish_fused(const float* p_net, float* U, const unsigned char* flags, const float* scale, float* p_out,
                             const float* u_inv, const float* u_bc, const unsigned char* qmask, float lo, float hi,
                             const Geo& g, cudaStream_t st);

// ---- tfl_cnn.cu ----
// Generic direct convolution (fp32 FMA): in [b][cin][z][y][x] -> out [b][cout][z][y][x].
// wdev: device weights re-laid out as [cin][tap][cout_pad], bias [cout].
// act: 0 none, 1 ReLU, 2 sigmoid.
int launch_conv_direct(const float* in, float* out, const float* wdev, const float* bdev, int cin, int cout,
                       int ksize, int act, const Geo& g, cudaStream_t st);
void launch_pool(const float* in, float* out, int nbc, int nz, int ny, int nx, int p, int is3d, int is_max,
                 cudaStream_t st);
void launch_pixel_shuffle(const float* in, float* out, int nb, int n_out, int nz, int ny, int nx, int s, int is3d,
                          cudaStream_t st);

// ---- tfl_pcg.cu: matrix-free PCG pressure solve ----
struct PcgScratch {            // owned by the context, grow-only
  void* comp_buf = nullptr;    // per-component CG scalars
  size_t comp_cap = 0;
  unsigned long long* prog = nullptr;   // progress words of the triangular-sweep pipeline
  size_t prog_cap = 0;
  unsigned long long epoch = 0;
  int* host = nullptr;         // pinned read-back words
  int sm_count = 0;
  void* debug_timing = nullptr;   // debug: device buffer [chunks][4] of sweep timestamps
  int groups_override = 0;     // debug: planes per CTA of the sweep kernel (0 = as many as fit)
};
size_t pcg_workspace_bytes(int nb, int nz, int ny, int nx);
const char* pcg_status_string(int rc);
// precond: 0 none, 1 ilu0, 2 ic0.  Returns 0 or a status for pcg_status_string.  Synchronises `st`.
int pcg_solve(PcgScratch& sc, void* workspace, float* p, const float* flags, const float* div, int nb, int nz, int ny,
              int nx, int is3d, int precond, float tol, int max_iter, float* residual, int* iterations,
              long long* launches, cudaStream_t st);
void pcg_release(PcgScratch& sc);
int normalize_pressure_mean(void* workspace, float* p, const float* flags, int nb, int nz, int ny, int nx, int is3d,
                            long long* launches, cudaStream_t st);

// ---- tfl_aux_ops.cu (-fmad=false): operators of tfluids/init.lua around the step ----
void launch_upsample_nearest(const float* in, float* out, int nbf, int nz, int ny, int nx, int ratio, cudaStream_t st);
// axis: 0 = x, 1 = y, 2 = z of a [nbf][nz][ny][nx] array.
void launch_blur_axis(const float* src, float* dst, int nbf, int nz, int ny, int nx, int axis, int rad, cudaStream_t st);
void launch_signed_distance_field(const float* flags, float* dst, int nb, int nz, int ny, int nx, int rad,
                                  cudaStream_t st);
void launch_velocity_divergence_backward(const float* flags, const float* go, float* grad_u, int nb, int nz, int ny,
                                         int nx, int is3d, cudaStream_t st);
void launch_velocity_update_backward(const float* flags, const float* go, float* grad_p, int nb, int nz, int ny, int nx,
                                     int is3d, cudaStream_t st);
void launch_upsample_nearest_backward(const float* go, float* gi, int nbf, int nz, int ny, int nx, int ratio,
                                      cudaStream_t st);

}  // namespace tfl
