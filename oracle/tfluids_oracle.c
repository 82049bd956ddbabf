/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the FluidNet `tfluids.simulate` hot path.
 *
 * This file is a plain-C restatement (written from scratch, float32, IEEE, no FMA
 * contraction) of the algorithm implemented by the reference's CPU operators.  It is
 * used ONLY by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs as the checker.  The product (libtfl.so, CUDA) never links,
 * imports or calls anything in oracle/.
 *
 * Pinning: there are no golden vectors in the reference tree
 * (torch/tfluids/test_data/ holds only a .gitignore), so this restatement is pinned
 * against (1) the reference's own CPU code compiled in place (oracle/_ref, see
 * oracle/Makefile) on seeded random inputs -- bit-exact, tests/test_oracle_vs_ref.py --
 * (2) fixtures generated from that compiled reference and committed under
 * tests/golden/ (tests/golden/make_golden.py), and (3) the reference's data-free
 * known-answer tests (emptyDomain / flagsToOccupancy, test_tfluids.lua:675-753, and
 * the line-trace geometry cases of generic/CalcLineTraceTest.m:101-151).
 * Jacobi has no CPU reference (generic/tfluids.cc:836-839 errors out); it is restated
 * from the CUDA kernel + host loop generic/tfluids.cu:1765-1927.  The conv stack lives
 * in un-vendored cuDNN/cudnn.torch (README.md:35-52): textbook cross-correlation here,
 * "parity unpinned" for that part (the model-graph semantics around it are pinned on
 * torch.nn.functional, tests/test_oracle_model_graph.py).  The PCG solver is CUDA-only in the
 * reference too (generic/tfluids.cu:1245-1759): its algorithm is restated, its connected-component
 * labelling is pinned on the reference's CPU code, its iterates are "parity unpinned" and checked by the
 * reference test's own criteria.  The operators around the step (blur, signed distance, up-sampling,
 * pressure-mean removal) and the three backward operators are pinned on the compiled reference like the
 * step operators (tests/test_oracle_aux_ops.py, tests/test_oracle_edge_sizes.py).
 *
 * Layout everywhere: 5-D [b][c][z][y][x], x fastest, float32.
 * All file:line citations are relative to /root/reference/torch/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Cell-type bits: tfluids/third_party/cell_type.h:22-33. */
enum {
  CELL_FLUID = 1, CELL_OBSTACLE = 2, CELL_EMPTY = 4, CELL_INFLOW = 8,
  CELL_OUTFLOW = 16, CELL_OPEN = 32, CELL_STICK = 128
};

/* Advection methods: tfluids/generic/advect_type.h:21-28, advect_type.cc:19-38. */
enum {
  ADV_EULER_MANTA = 0, ADV_MACCORMACK_MANTA = 1, ADV_EULER_OURS = 2,
  ADV_RK2_OURS = 3, ADV_RK3_OURS = 4, ADV_MACCORMACK_OURS = 5
};

typedef struct { int nb, nz, ny, nx, is3d; } orc_dims;
typedef struct { float x, y, z; } v3;

/* Number of line traces that hit a condition the reference treats as a hard error
 * (THError in generic/calc_line_trace.cc).  Tests assert it stays 0. */
static long g_trace_faults = 0;
long orc_trace_faults(void) { return g_trace_faults; }
void orc_reset_trace_faults(void) { g_trace_faults = 0; }
int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------
 * Grid views (third_party/grid.h:26-262, grid.cc:58-78 index5d).
 * ---------------------------------------------------------------------------------- */
static inline long cells(const orc_dims* d) { return (long)d->nz * d->ny * d->nx; }
static inline long at(const orc_dims* d, int nc, int b, int c, int k, int j, int i) {
  return ((((long)b * nc + c) * d->nz + k) * d->ny + j) * d->nx + i;
}
static inline int flag_at(const float* fl, const orc_dims* d, int b, int k, int j, int i) {
  return (int)fl[at(d, 1, b, 0, k, j, i)];   /* static_cast<int>, grid.h:104 */
}
static inline int is_fluid(const float* fl, const orc_dims* d, int b, int k, int j, int i) {
  return (flag_at(fl, d, b, k, j, i) & CELL_FLUID) != 0;
}
static inline int is_obstacle(const float* fl, const orc_dims* d, int b, int k, int j, int i) {
  return (flag_at(fl, d, b, k, j, i) & CELL_OBSTACLE) != 0;
}
static inline int is_empty(const float* fl, const orc_dims* d, int b, int k, int j, int i) {
  return (flag_at(fl, d, b, k, j, i) & CELL_EMPTY) != 0;
}
static inline int is_stick(const float* fl, const orc_dims* d, int b, int k, int j, int i) {
  return (flag_at(fl, d, b, k, j, i) & CELL_STICK) != 0;
}
static inline int is_outflow(const float* fl, const orc_dims* d, int b, int k, int j, int i) {
  return (flag_at(fl, d, b, k, j, i) & CELL_OUTFLOW) != 0;
}
static inline int on_border(const orc_dims* d, int k, int j, int i, int bnd) {
  return i < bnd || i > d->nx - 1 - bnd || j < bnd || j > d->ny - 1 - bnd ||
         (d->is3d && (k < bnd || k > d->nz - 1 - bnd));
}
static inline int nchan_vel(const orc_dims* d) { return d->is3d ? 3 : 2; }

/* getDx, grid.cc:37-40. */
float orc_get_dx(const orc_dims* d) {
  int m = d->nx > d->ny ? d->nx : d->ny;
  if (d->nz > m) m = d->nz;
  return 1.0f / (float)m;
}

/* std::min<real>/std::max<real> semantics (NaN behaviour included). */
static inline float std_minf(float a, float b) { return (b < a) ? b : a; }
static inline float std_maxf(float a, float b) { return (a < b) ? b : a; }
static inline int clampi(int x, int lo, int hi) {          /* init.cu:33-35 */
  int m = x < hi ? x : hi;
  return m > lo ? m : lo;
}
static inline float clampf(float v, float lo, float hi) {  /* third_party/tfluids.cc:245-247 */
  return std_minf(hi, std_maxf(lo, v));
}

/* vec3::norm with the squared-length cutoff, generic/vec3.h:119-127 (float kEpsilon 1e-6). */
static inline float v3_norm(v3 a) {
  const float len_sq = a.x * a.x + a.y * a.y + a.z * a.z;
  return (len_sq > 1e-6f) ? sqrtf(len_sq) : 0.0f;
}

/* ------------------------------------------------------------------------------------
 * Interpolation: buildIndex grid.cc:82-130, interpol :182-202, interpolComponent
 * :435-456, interpolWithFluid :204-332.
 * ---------------------------------------------------------------------------------- */
typedef struct { int xi, yi, zi; float s0, s1, t0, t1, f0, f1; } lerp_idx;

static inline lerp_idx build_index(const orc_dims* d, v3 pos) {
  lerp_idx q;
  const float px = pos.x - 0.5f, py = pos.y - 0.5f, pz = pos.z - 0.5f;
  q.xi = (int)px; q.yi = (int)py; q.zi = (int)pz;           /* truncation, not floor */
  q.s1 = px - (float)q.xi; q.s0 = 1.0f - q.s1;
  q.t1 = py - (float)q.yi; q.t0 = 1.0f - q.t1;
  q.f1 = pz - (float)q.zi; q.f0 = 1.0f - q.f1;
  if (px < 0.0f) { q.xi = 0; q.s0 = 1.0f; q.s1 = 0.0f; }
  if (py < 0.0f) { q.yi = 0; q.t0 = 1.0f; q.t1 = 0.0f; }
  if (pz < 0.0f) { q.zi = 0; q.f0 = 1.0f; q.f1 = 0.0f; }
  if (q.xi >= d->nx - 1) { q.xi = d->nx - 2; q.s0 = 0.0f; q.s1 = 1.0f; }
  if (q.yi >= d->ny - 1) { q.yi = d->ny - 2; q.t0 = 0.0f; q.t1 = 1.0f; }
  if (d->nz > 1 && q.zi >= d->nz - 1) { q.zi = d->nz - 2; q.f0 = 0.0f; q.f1 = 1.0f; }
  return q;
}

/* g points at the [z][y][x] block of one (batch, channel). */
static inline float lerp_block(const float* g, const orc_dims* d, v3 pos) {
  const lerp_idx q = build_index(d, pos);
  const long sy = d->nx, sz = (long)d->nx * d->ny;
  if (d->is3d) {
    const float* a = g + (long)q.zi * sz + (long)q.yi * sy + q.xi;
    const float lo = ((a[0] * q.t0 + a[sy] * q.t1) * q.s0 +
                      (a[1] * q.t0 + a[sy + 1] * q.t1) * q.s1) * q.f0;
    const float hi = ((a[sz] * q.t0 + a[sz + sy] * q.t1) * q.s0 +
                      (a[sz + 1] * q.t0 + a[sz + sy + 1] * q.t1) * q.s1) * q.f1;
    return lo + hi;
  } else {
    const float* a = g + (long)q.yi * sy + q.xi;
    return (a[0] * q.t0 + a[sy] * q.t1) * q.s0 + (a[1] * q.t0 + a[sy + 1] * q.t1) * q.s1;
  }
}

typedef struct { float v; int ok; } fluid_val;
static inline fluid_val pair_fluid(fluid_val a, fluid_val b, float ta, float tb) {
  fluid_val r;                                   /* interpol1DWithFluid, grid.cc:204-222 */
  if (!a.ok && !b.ok) { r.v = 0.0f; r.ok = 0; }
  else if (!a.ok) { r.v = b.v; r.ok = 1; }
  else if (!b.ok) { r.v = a.v; r.ok = 1; }
  else { r.v = a.v * ta + b.v * tb; r.ok = 1; }
  return r;
}

static inline float lerp_block_fluid(const float* g, const float* flb, const orc_dims* d, v3 pos) {
  const lerp_idx q = build_index(d, pos);
  const long sy = d->nx, sz = (long)d->nx * d->ny;
#define FV(off) ((fluid_val){ g[(off)], (((int)flb[(off)]) & CELL_FLUID) != 0 })
  if (d->is3d) {
    const long o = (long)q.zi * sz + (long)q.yi * sy + q.xi;
    const fluid_val ab = pair_fluid(FV(o), FV(o + sy), q.t0, q.t1);
    const fluid_val cd = pair_fluid(FV(o + 1), FV(o + sy + 1), q.t0, q.t1);
    const fluid_val ef = pair_fluid(FV(o + sz), FV(o + sz + sy), q.t0, q.t1);
    const fluid_val gh = pair_fluid(FV(o + sz + 1), FV(o + sz + sy + 1), q.t0, q.t1);
    const fluid_val abcd = pair_fluid(ab, cd, q.s0, q.s1);
    const fluid_val efgh = pair_fluid(ef, gh, q.s0, q.s1);
    const fluid_val all = pair_fluid(abcd, efgh, q.f0, q.f1);
    return all.ok ? all.v : lerp_block(g, d, pos);
  } else {
    const long o = (long)q.yi * sy + q.xi;
    const fluid_val ab = pair_fluid(FV(o), FV(o + sy), q.t0, q.t1);
    const fluid_val cd = pair_fluid(FV(o + 1), FV(o + sy + 1), q.t0, q.t1);
    const fluid_val all = pair_fluid(ab, cd, q.s0, q.s1);
    return all.ok ? all.v : lerp_block(g, d, pos);
  }
#undef FV
}

/* MAC helpers: getCentered grid.cc:346-356, getAtMACX/Y/Z :374-417. */
static inline v3 mac_centered(const float* U, const orc_dims* d, int b, int k, int j, int i) {
  const int nc = nchan_vel(d);
  v3 r;
  r.x = 0.5f * (U[at(d, nc, b, 0, k, j, i)] + U[at(d, nc, b, 0, k, j, i + 1)]);
  r.y = 0.5f * (U[at(d, nc, b, 1, k, j, i)] + U[at(d, nc, b, 1, k, j + 1, i)]);
  r.z = d->is3d ? 0.5f * (U[at(d, nc, b, 2, k, j, i)] + U[at(d, nc, b, 2, k + 1, j, i)]) : 0.0f;
  return r;
}
static inline v3 mac_at_x(const float* U, const orc_dims* d, int b, int k, int j, int i) {
  const int nc = nchan_vel(d);
  v3 r;
  r.x = U[at(d, nc, b, 0, k, j, i)];
  r.y = 0.25f * (U[at(d, nc, b, 1, k, j, i)] + U[at(d, nc, b, 1, k, j, i - 1)] +
                 U[at(d, nc, b, 1, k, j + 1, i)] + U[at(d, nc, b, 1, k, j + 1, i - 1)]);
  r.z = d->is3d ? 0.25f * (U[at(d, nc, b, 2, k, j, i)] + U[at(d, nc, b, 2, k, j, i - 1)] +
                           U[at(d, nc, b, 2, k + 1, j, i)] + U[at(d, nc, b, 2, k + 1, j, i - 1)])
                : 0.0f;
  return r;
}
static inline v3 mac_at_y(const float* U, const orc_dims* d, int b, int k, int j, int i) {
  const int nc = nchan_vel(d);
  v3 r;
  r.x = 0.25f * (U[at(d, nc, b, 0, k, j, i)] + U[at(d, nc, b, 0, k, j - 1, i)] +
                 U[at(d, nc, b, 0, k, j, i + 1)] + U[at(d, nc, b, 0, k, j - 1, i + 1)]);
  r.y = U[at(d, nc, b, 1, k, j, i)];
  r.z = d->is3d ? 0.25f * (U[at(d, nc, b, 2, k, j, i)] + U[at(d, nc, b, 2, k, j - 1, i)] +
                           U[at(d, nc, b, 2, k + 1, j, i)] + U[at(d, nc, b, 2, k + 1, j - 1, i)])
                : 0.0f;
  return r;
}
static inline v3 mac_at_z(const float* U, const orc_dims* d, int b, int k, int j, int i) {
  const int nc = nchan_vel(d);      /* only called for 3-D grids */
  v3 r;
  r.x = 0.25f * (U[at(d, nc, b, 0, k, j, i)] + U[at(d, nc, b, 0, k - 1, j, i)] +
                 U[at(d, nc, b, 0, k, j, i + 1)] + U[at(d, nc, b, 0, k - 1, j, i + 1)]);
  r.y = 0.25f * (U[at(d, nc, b, 1, k, j, i)] + U[at(d, nc, b, 1, k - 1, j, i)] +
                 U[at(d, nc, b, 1, k, j + 1, i)] + U[at(d, nc, b, 1, k - 1, j + 1, i)]);
  r.z = U[at(d, nc, b, 2, k, j, i)];
  return r;
}
static inline v3 v3_scale(v3 a, float s) { v3 r = { a.x * s, a.y * s, a.z * s }; return r; }

/* ------------------------------------------------------------------------------------
 * Line trace: generic/calc_line_trace.cc.
 * ---------------------------------------------------------------------------------- */
#define HIT_MARGIN 1e-5f      /* calc_line_trace.cc:22 */
#define TRACE_EPS 1e-12f      /* calc_line_trace.cc:23 */

static inline int out_of_domain(v3 p, const orc_dims* d) {          /* :44-52 */
  return p.x <= 0.0f || p.x >= (float)d->nx || p.y <= 0.0f || p.y >= (float)d->ny ||
         p.z <= 0.0f || p.z >= (float)d->nz;
}
static inline int blocked_at(const float* fl, const orc_dims* d, int b, v3 p) { /* :54-63,85-90 */
  return !is_fluid(fl, d, b, (int)p.z, (int)p.y, (int)p.x);
}

/* Graphics-Gems ray/box with the reference's modified tolerance, :101-171. */
static int ray_hits_box(const float lo[3], const float hi[3], const float org[3],
                        const float dir[3], float out[3]) {
  int inside = 1, side[3];       /* 0 right, 1 left, 2 middle (generic/quadrants.h) */
  float plane[3] = {0, 0, 0}, tmax[3];
  for (int a = 0; a < 3; a++) {
    if (org[a] < lo[a]) { side[a] = 1; plane[a] = lo[a]; inside = 0; }
    else if (org[a] > hi[a]) { side[a] = 0; plane[a] = hi[a]; inside = 0; }
    else side[a] = 2;
  }
  if (inside) { out[0] = org[0]; out[1] = org[1]; out[2] = org[2]; return 1; }
  for (int a = 0; a < 3; a++)
    tmax[a] = (side[a] != 2 && dir[a] != 0.0f) ? (plane[a] - org[a]) / dir[a] : -1.0f;
  int w = 0;
  for (int a = 1; a < 3; a++) if (tmax[w] < tmax[a]) w = a;
  if (tmax[w] < 0.0f) return 0;
  const float tol = 1e-6f;                                        /* :158 */
  for (int a = 0; a < 3; a++) {
    if (a != w) {
      out[a] = org[a] + tmax[w] * dir[a];
      if (out[a] < (lo[a] - tol) || out[a] > (hi[a] + tol)) return 0;
    } else {
      out[a] = plane[a];
    }
  }
  return 1;
}

static int ray_unit_box(v3 pos, v3 dir, v3 ctr, float margin, v3* ip) {   /* :176-196 */
  const float lo[3] = { ctr.x - 0.5f - margin, ctr.y - 0.5f - margin, ctr.z - 0.5f - margin };
  const float hi[3] = { ctr.x + 0.5f + margin, ctr.y + 0.5f + margin, ctr.z + 0.5f + margin };
  const float o[3] = { pos.x, pos.y, pos.z }, dr[3] = { dir.x, dir.y, dir.z };
  float out[3] = { ip->x, ip->y, ip->z };
  const int hit = ray_hits_box(lo, hi, o, dr, out);
  ip->x = out[0]; ip->y = out[1]; ip->z = out[2];
  return hit;
}

/* Minimum parametric step that brings the segment pos->next to a face offset by margin,
 * :205-286. */
static int ray_border(v3 pos, v3 next, const orc_dims* d, float margin, v3* ip) {
  float step = FLT_MAX;
  const float p[3] = { pos.x, pos.y, pos.z }, n[3] = { next.x, next.y, next.z };
  const float ext[3] = { (float)d->nx, (float)d->ny, (float)d->nz };
  for (int a = 0; a < 3; a++) {                 /* low faces x, y, z */
    if (n[a] <= margin) {
      const float dl = n[a] - p[a];
      if (fabsf(dl) >= TRACE_EPS) step = std_minf(step, (margin - p[a]) / dl);
    }
  }
  for (int a = 0; a < 3; a++) {                 /* high faces x, y, z */
    if (n[a] >= (ext[a] - margin)) {
      const float dl = n[a] - p[a];
      if (fabsf(dl) >= TRACE_EPS) step = std_minf(step, (ext[a] - margin - p[a]) / dl);
    }
  }
  if (step < 0.0f || step >= FLT_MAX) return 0;
  ip->x = step * (next.x - pos.x) + pos.x;
  ip->y = step * (next.y - pos.y) + pos.y;
  ip->z = step * (next.z - pos.z) + pos.z;
  return 1;
}

/* calcLineTrace, :313-503 (do_line_trace == true, the only mode the operators use). */
static int line_trace(v3 pos, v3 delta, const float* fl, const orc_dims* d, int b, v3* out) {
  *out = pos;
  const float length = v3_norm(delta);
  if (length <= TRACE_EPS) return 0;
  const v3 dir = { delta.x / length, delta.y / length, delta.z / length };
  float travelled = 0.0f;
  while (travelled < (length - HIT_MARGIN)) {
    const float step = std_minf(length - travelled, 1.0f);
    v3 next = { out->x + dir.x * step, out->y + dir.y * step, out->z + dir.z * step };
    if (out_of_domain(next, d)) {
      v3 ip;
      if (!ray_border(*out, next, d, HIT_MARGIN, &ip)) {
        ip = next;                                                     /* :381-382, :72-81 */
        ip.x = std_minf(std_maxf(ip.x, HIT_MARGIN), (float)d->nx - HIT_MARGIN);
        ip.y = std_minf(std_maxf(ip.y, HIT_MARGIN), (float)d->ny - HIT_MARGIN);
        ip.z = std_minf(std_maxf(ip.z, HIT_MARGIN), (float)d->nz - HIT_MARGIN);
      }
      if (out_of_domain(ip, d)) {         /* reference: THError (:388-390) */
#pragma omp atomic
        g_trace_faults++;
        return 1;
      }
      if (!blocked_at(fl, d, b, ip)) { *out = ip; return 1; }
      next = ip;
    }
    if (blocked_at(fl, d, b, next)) {
      for (int tries = 0; tries <= 4; tries++) {                       /* :412-464 */
        if (!blocked_at(fl, d, b, next)) break;
        if (tries == 4) {                 /* reference: THError (:420-422) */
#pragma omp atomic
          g_trace_faults++;
          return 1;
        }
        const v3 ctr = { (float)((int)next.x) + 0.5f, (float)((int)next.y) + 0.5f,
                         (float)((int)next.z) + 0.5f };
        v3 ip = { 0, 0, 0 };
        if (!ray_unit_box(*out, dir, ctr, HIT_MARGIN, &ip)) return 1;   /* :444-453 */
        next = ip;
      }
      *out = next;
      return 1;
    }
    *out = next;
    travelled += step;
  }
  return 0;
}

int orc_calc_line_trace(const float* pos, const float* delta, const float* flags,
                        const orc_dims* d, float* new_pos) {
  v3 p = { pos[0], pos[1], pos[2] }, dl = { delta[0], delta[1], delta[2] }, o;
  const int hit = line_trace(p, dl, flags, d, 0, &o);
  new_pos[0] = o.x; new_pos[1] = o.y; new_pos[2] = o.z;
  return hit;
}

/* ------------------------------------------------------------------------------------
 * emptyDomain / flagsToOccupancy: generic/tfluids.cc:136-210.
 * ---------------------------------------------------------------------------------- */
void orc_empty_domain(float* flags, const orc_dims* d, int bnd) {
  for (int b = 0; b < d->nb; b++)
    for (int k = 0; k < d->nz; k++)
      for (int j = 0; j < d->ny; j++)
        for (int i = 0; i < d->nx; i++)
          flags[at(d, 1, b, 0, k, j, i)] =
              on_border(d, k, j, i, bnd) ? (float)CELL_OBSTACLE : (float)CELL_FLUID;
}

/* Returns the number of cells that are neither exactly Fluid nor exactly Obstacle (the
 * reference raises an error if any, generic/tfluids.cc:194-207; those cells are left
 * untouched). */
long orc_flags_to_occupancy(const float* flags, float* occ, long n) {
  long bad = 0;
  for (long i = 0; i < n; i++) {
    const int f = (int)flags[i];
    if (f == CELL_FLUID) occ[i] = 0.0f;
    else if (f == CELL_OBSTACLE) occ[i] = 1.0f;
    else bad++;
  }
  return bad;
}

/* ------------------------------------------------------------------------------------
 * advectScalar: third_party/tfluids.cc:23-588.
 * ---------------------------------------------------------------------------------- */
static inline float sample_scalar(const float* src_b, const float* fl_b, const orc_dims* d,
                                  v3 pos, int sample_outside) {
  return sample_outside ? lerp_block(src_b, d, pos) : lerp_block_fluid(src_b, fl_b, d, pos);
}
static inline v3 sample_vel(const float* U, const orc_dims* d, int b, v3 pos) {
  const int nc = nchan_vel(d);
  const long n = cells(d);
  const float* base = U + (long)b * nc * n;
  v3 r;
  r.x = lerp_block(base, d, pos);
  r.y = lerp_block(base + n, d, pos);
  r.z = d->is3d ? lerp_block(base + 2 * n, d, pos) : 0.0f;
  return r;
}

/* One forward (or, with -dt and src=fwd, backward) pass for every method. */
static float advect_cell_scalar(int method, const float* fl, const float* U, const float* src,
                                const orc_dims* d, float dt, int b, int k, int j, int i,
                                int sample_outside, float* pos_out /* [3] strided by cells */,
                                long pos_stride) {
  const long n = cells(d);
  const float* src_b = src + (long)b * n;
  const float* fl_b = fl + (long)b * n;
  const v3 start = { (float)i + 0.5f, (float)j + 0.5f, (float)k + 0.5f };
  if (method == ADV_EULER_MANTA || method == ADV_MACCORMACK_MANTA) {
    const v3 c = mac_centered(U, d, b, k, j, i);                      /* :211-220 */
    const v3 p = { start.x - c.x * dt, start.y - c.y * dt, start.z - c.z * dt };
    return lerp_block(src_b, d, p);
  }
  if (!is_fluid(fl, d, b, k, j, i)) {
    if (pos_out) {                                                    /* :157-161 */
      pos_out[0] = (float)i + 0.5f;
      pos_out[pos_stride] = (float)j + 0.5f;
      if (d->is3d) pos_out[2 * pos_stride] = (float)k + 0.5f;
    }
    return src_b[((long)k * d->ny + j) * d->nx + i];
  }
  const v3 c = mac_centered(U, d, b, k, j, i);
  if (method == ADV_EULER_OURS || method == ADV_MACCORMACK_OURS) {    /* :152-209 */
    v3 back;
    line_trace(start, v3_scale(c, -dt), fl, d, b, &back);
    if (pos_out) {
      pos_out[0] = back.x;
      pos_out[pos_stride] = back.y;
      if (d->is3d) pos_out[2 * pos_stride] = back.z;
    }
    return sample_scalar(src_b, fl_b, d, back, sample_outside);
  }
  if (method == ADV_RK2_OURS) {                                       /* :23-76 */
    v3 half;
    if (line_trace(start, v3_scale(c, -dt * 0.5f), fl, d, b, &half))
      return sample_scalar(src_b, fl_b, d, half, sample_outside);
    const v3 v = sample_vel(U, d, b, half);
    v3 back;
    line_trace(start, v3_scale(v, -dt), fl, d, b, &back);
    return sample_scalar(src_b, fl_b, d, back, sample_outside);
  }
  /* ADV_RK3_OURS, :78-147 (CPU behaviour: a 3rd-stage hit samples at k3_pos). */
  v3 p2, p3, back;
  if (line_trace(start, v3_scale(c, -dt * 0.5f), fl, d, b, &p2))
    return sample_scalar(src_b, fl_b, d, p2, sample_outside);
  const v3 k2 = sample_vel(U, d, b, p2);
  if (line_trace(start, v3_scale(k2, -dt * 0.75f), fl, d, b, &p3))
    return sample_scalar(src_b, fl_b, d, p3, sample_outside);
  const v3 k3 = sample_vel(U, d, b, p3);
  const float w1 = -dt * (float)(2.0 / 9.0), w2 = -dt * (float)(3.0 / 9.0),
              w3 = -dt * (float)(4.0 / 9.0);
  const v3 a1 = v3_scale(c, w1), a2 = v3_scale(k2, w2), a3 = v3_scale(k3, w3);
  const v3 disp = { (a1.x + a2.x) + a3.x, (a1.y + a2.y) + a3.y, (a1.z + a2.z) + a3.z };
  line_trace(start, disp, fl, d, b, &back);
  return sample_scalar(src_b, fl_b, d, back, sample_outside);
}

/* fwd, bwd: [b][1][z][y][x]; fwd_pos, bwd_pos: [b][2|3][z][y][x] scratch (may be NULL for
 * the non-MacCormack methods).  dst must not alias s. */
void orc_advect_scalar(float dt, const float* s, const float* U, const float* flags,
                       const orc_dims* d, int method, int sample_outside, float strength,
                       float* dst, float* fwd, float* bwd, float* fwd_pos, float* bwd_pos) {
  const long n = cells(d);
  const int nc = nchan_vel(d);
  const int two_pass = (method == ADV_MACCORMACK_MANTA || method == ADV_MACCORMACK_OURS);
  const int save_pos = (method == ADV_MACCORMACK_OURS);
  for (int b = 0; b < d->nb; b++) {
    float* first = two_pass ? fwd : dst;
    int k, j, i;
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      const long c0 = ((long)k * d->ny + j) * d->nx + i;
      float* pp = fwd_pos ? fwd_pos + (long)b * nc * n + c0 : NULL;
      if (on_border(d, k, j, i, 1)) {                                  /* :477-484 */
        first[(long)b * n + c0] = 0.0f;
        if (pp) { pp[0] = (float)i + 0.5f; pp[n] = (float)j + 0.5f; if (d->is3d) pp[2 * n] = (float)k + 0.5f; }
        continue;
      }
      first[(long)b * n + c0] = advect_cell_scalar(method, flags, U, s, d, dt, b, k, j, i,
                                                   sample_outside, save_pos ? pp : NULL, n);
    }
    if (!two_pass) continue;
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      const long c0 = ((long)k * d->ny + j) * d->nx + i;
      float* pp = bwd_pos ? bwd_pos + (long)b * nc * n + c0 : NULL;
      if (on_border(d, k, j, i, 1)) {                                  /* :529-535 */
        bwd[(long)b * n + c0] = 0.0f;
        if (pp) { pp[0] = (float)i + 0.5f; pp[n] = (float)j + 0.5f; if (d->is3d) pp[2 * n] = (float)k + 0.5f; }
        continue;
      }
      bwd[(long)b * n + c0] = advect_cell_scalar(method, flags, U, fwd, d, -dt, b, k, j, i,
                                                 sample_outside, save_pos ? pp : NULL, n);
    }
    /* MacCormackCorrect, :222-234: `strength * 0.5` is a double expression in the
     * reference, so the update is evaluated in double and rounded once to float. */
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      const long c = (long)b * n + ((long)k * d->ny + j) * d->nx + i;
      float v = fwd[c];
      if (is_fluid(flags, d, b, k, j, i)) {
        const float diff = s[c] - bwd[c];
        v = (float)((double)v + ((double)strength * 0.5) * (double)diff);
      }
      dst[c] = v;
    }
    /* Clamp, :562-583. */
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      if (on_border(d, k, j, i, 1)) continue;
      const long c0 = ((long)k * d->ny + j) * d->nx + i;
      const long c = (long)b * n + c0;
      if (method == ADV_MACCORMACK_MANTA) {                            /* :249-325 */
        const v3 vel = v3_scale(mac_centered(U, d, b, k, j, i), dt);
        const float fi = (float)i, fj = (float)j, fk = (float)k;
        float lo = FLT_MAX, hi = -FLT_MAX, val = dst[c];
        int bail = 0;
        for (int l = 0; l < 2 && !bail; l++) {
          const int px = l == 0 ? (int)(fi - vel.x) : (int)(fi + vel.x);
          const int py = l == 0 ? (int)(fj - vel.y) : (int)(fj + vel.y);
          const int pz = l == 0 ? (int)(fk - vel.z) : (int)(fk + vel.z);
          const int i0 = clampi(px, 0, d->nx - 2), j0 = clampi(py, 0, d->ny - 2);
          const int k0 = clampi(pz, 0, d->is3d ? d->nz - 2 : 1);
          const int i1 = i0 + 1, j1 = j0 + 1, k1 = d->is3d ? k0 + 1 : k0;
          /* isInBounds(.., 0), grid.cc:42-51 */
          int inb = i0 >= 0 && j0 >= 0 && i0 < d->nx && j0 < d->ny &&
                    i1 >= 0 && j1 >= 0 && i1 < d->nx && j1 < d->ny;
          if (d->is3d) inb = inb && k0 >= 0 && k0 < d->nz && k1 >= 0 && k1 < d->nz;
          else inb = inb && k0 == 0 && k1 == 0;
          if (!inb) { bail = 1; break; }
          const float* sb = s + (long)b * n;
#define MM(kk, jj, ii) { const float t = sb[((long)(kk) * d->ny + (jj)) * d->nx + (ii)]; \
                         if (t < lo) lo = t; if (t > hi) hi = t; }
          MM(k0, j0, i0) MM(k0, j0, i1) MM(k0, j1, i0) MM(k0, j1, i1)
          if (d->is3d) { MM(k1, j0, i0) MM(k1, j0, i1) MM(k1, j1, i0) MM(k1, j1, i1) }
#undef MM
        }
        val = bail ? fwd[c] : clampf(val, lo, hi);
        const int fx = (int)((fi + 0.5f) - vel.x), fy = (int)((fj + 0.5f) - vel.y),
                  fz = (int)((fk + 0.5f) - vel.z);
        const int bx = (int)((fi + 0.5f) + vel.x), by = (int)((fj + 0.5f) + vel.y),
                  bz = (int)((fk + 0.5f) + vel.z);
        const int ux = d->nx - 1, uy = d->ny - 1, uz = d->nz - 1;
        if (fx < 0 || fy < 0 || fz < 0 || bx < 0 || by < 0 || bz < 0 ||
            fx > ux || fy > uy || (fz > uz && d->is3d) ||
            bx > ux || by > uy || (bz > uz && d->is3d) ||
            is_obstacle(flags, d, b, fz, fy, fx) || is_obstacle(flags, d, b, bz, by, bx)) {
          val = fwd[c];
        }
        dst[c] = val;
      } else {                                                         /* :331-413 */
        const float* pp = fwd_pos + (long)b * nc * n + c0;
        const float px = pp[0], py = pp[n], pz = d->is3d ? pp[2 * n] : 0.0f;
        const int i0 = clampi((int)px, 0, d->nx - 1), j0 = clampi((int)py, 0, d->ny - 1);
        const int k0 = d->is3d ? clampi((int)pz, 0, d->nz - 1) : 0;
        float lo = INFINITY, hi = -INFINITY;
        int found = 0;
        for (int kk = k0 - 1; kk <= k0 + 1; kk++) for (int jj = j0 - 1; jj <= j0 + 1; jj++)
          for (int ii = i0 - 1; ii <= i0 + 1; ii++) {
            if (kk < 0 || kk >= d->nz || jj < 0 || jj >= d->ny || ii < 0 || ii >= d->nx) continue;
            if (sample_outside || is_fluid(flags, d, b, kk, jj, ii)) {
              const float t = s[(long)b * n + ((long)kk * d->ny + jj) * d->nx + ii];
              if (t < lo) lo = t;
              if (t > hi) hi = t;
              found++;
            }
          }
        dst[c] = (found < 1) ? fwd[c] : clampf(dst[c], lo, hi);
      }
    }
  }
}

/* ------------------------------------------------------------------------------------
 * advectVel: third_party/tfluids.cc:594-920.
 * ---------------------------------------------------------------------------------- */
static v3 advect_cell_mac(int ours, const float* fl, const float* U, const float* src,
                          const orc_dims* d, float dt, int b, int k, int j, int i) {
  const int nc = nchan_vel(d);
  const long n = cells(d);
  const float* sb = src + (long)b * nc * n;
  v3 r;
  if (ours && !is_fluid(fl, d, b, k, j, i)) {                          /* :598-601 */
    const long c0 = ((long)k * d->ny + j) * d->nx + i;
    r.x = sb[c0]; r.y = sb[n + c0]; r.z = d->is3d ? sb[2 * n + c0] : 0.0f;
    return r;
  }
  const v3 start = { (float)i + 0.5f, (float)j + 0.5f, (float)k + 0.5f };
  v3 p;
  if (ours) {                                                          /* :611-631 */
    line_trace(start, v3_scale(mac_at_x(U, d, b, k, j, i), -dt), fl, d, b, &p);
    r.x = lerp_block(sb, d, p);
    line_trace(start, v3_scale(mac_at_y(U, d, b, k, j, i), -dt), fl, d, b, &p);
    r.y = lerp_block(sb + n, d, p);
    if (d->is3d) {
      line_trace(start, v3_scale(mac_at_z(U, d, b, k, j, i), -dt), fl, d, b, &p);
      r.z = lerp_block(sb + 2 * n, d, p);
    } else r.z = 0.0f;
  } else {                                                             /* :634-658 */
    v3 v = v3_scale(mac_at_x(U, d, b, k, j, i), dt);
    p.x = start.x - v.x; p.y = start.y - v.y; p.z = start.z - v.z;
    r.x = lerp_block(sb, d, p);
    v = v3_scale(mac_at_y(U, d, b, k, j, i), dt);
    p.x = start.x - v.x; p.y = start.y - v.y; p.z = start.z - v.z;
    r.y = lerp_block(sb + n, d, p);
    if (d->is3d) {
      v = v3_scale(mac_at_z(U, d, b, k, j, i), dt);
      p.x = start.x - v.x; p.y = start.y - v.y; p.z = start.z - v.z;
      r.z = lerp_block(sb + 2 * n, d, p);
    } else r.z = 0.0f;
  }
  return r;
}

static float clamp_component_mac(const float* orig_c /* channel block */, const orc_dims* d,
                                 float val, float fwd, int k, int j, int i, v3 vel) {
  /* doClampComponentMAC, :701-746; pos = (i, j, k) WITHOUT the +0.5 (a Manta quirk). */
  const float fi = (float)i, fj = (float)j, fk = (float)k;
  float lo = FLT_MAX, hi = -FLT_MAX;
  for (int l = 0; l < 2; l++) {
    const int px = l == 0 ? (int)(fi - vel.x) : (int)(fi + vel.x);
    const int py = l == 0 ? (int)(fj - vel.y) : (int)(fj + vel.y);
    const int pz = l == 0 ? (int)(fk - vel.z) : (int)(fk + vel.z);
    const int i0 = clampi(px, 0, d->nx - 2), j0 = clampi(py, 0, d->ny - 2);
    const int k0 = clampi(pz, 0, d->is3d ? d->nz - 2 : 1);
    const int i1 = i0 + 1, j1 = j0 + 1, k1 = d->is3d ? k0 + 1 : k0;
    int inb = i0 >= 0 && j0 >= 0 && i0 < d->nx && j0 < d->ny && i1 < d->nx && j1 < d->ny;
    if (d->is3d) inb = inb && k0 >= 0 && k0 < d->nz && k1 < d->nz;
    else inb = inb && k0 == 0 && k1 == 0;
    if (!inb) return fwd;
#define MM(kk, jj, ii) { const float t = orig_c[((long)(kk) * d->ny + (jj)) * d->nx + (ii)]; \
                         if (t < lo) lo = t; if (t > hi) hi = t; }
    MM(k0, j0, i0) MM(k0, j0, i1) MM(k0, j1, i0) MM(k0, j1, i1)
    if (d->is3d) { MM(k1, j0, i0) MM(k1, j0, i1) MM(k1, j1, i0) MM(k1, j1, i1) }
#undef MM
  }
  return clampf(val, lo, hi);
}

/* fwd, bwd: [b][2|3][z][y][x] scratch. dst must not alias U. */
void orc_advect_vel(float dt, const float* U, const float* flags, const orc_dims* d,
                    int method, float strength, float* dst, float* fwd, float* bwd) {
  if (method == ADV_RK2_OURS || method == ADV_RK3_OURS) method = ADV_MACCORMACK_OURS; /* :799-802 */
  const int nc = nchan_vel(d);
  const long n = cells(d);
  const int two_pass = (method == ADV_MACCORMACK_MANTA || method == ADV_MACCORMACK_OURS);
  const int ours = (method == ADV_EULER_OURS || method == ADV_MACCORMACK_OURS);
  for (int b = 0; b < d->nb; b++) {
    float* first = (two_pass ? fwd : dst) + (long)b * nc * n;
    int k, j, i;
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      const long c0 = ((long)k * d->ny + j) * d->nx + i;
      v3 v = { 0.0f, 0.0f, 0.0f };
      if (!on_border(d, k, j, i, 1)) v = advect_cell_mac(ours, flags, U, U, d, dt, b, k, j, i);
      first[c0] = v.x; first[n + c0] = v.y; if (d->is3d) first[2 * n + c0] = v.z;
    }
    if (!two_pass) continue;
    float* bw = bwd + (long)b * nc * n;
    const float* fw = fwd + (long)b * nc * n;
    const float* ob = U + (long)b * nc * n;
    float* db = dst + (long)b * nc * n;
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      const long c0 = ((long)k * d->ny + j) * d->nx + i;
      v3 v = { 0.0f, 0.0f, 0.0f };
      if (!on_border(d, k, j, i, 1)) v = advect_cell_mac(ours, flags, U, fwd, d, -dt, b, k, j, i);
      bw[c0] = v.x; bw[n + c0] = v.y; if (d->is3d) bw[2 * n + c0] = v.z;
    }
    /* MacCormackCorrectMAC, :660-699 (double expression, see orc_advect_scalar). */
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      const long c0 = ((long)k * d->ny + j) * d->nx + i;
      const int cf = is_fluid(flags, d, b, k, j, i);
      int skip[3] = { !cf, !cf, !cf };
      if (i > 0 && !is_fluid(flags, d, b, k, j, i - 1)) skip[0] = 1;
      if (j > 0 && !is_fluid(flags, d, b, k, j - 1, i)) skip[1] = 1;
      if (d->is3d && k > 0 && !is_fluid(flags, d, b, k - 1, j, i)) skip[2] = 1;
      for (int c = 0; c < nc; c++) {
        const long o = (long)c * n + c0;
        float v = fw[o];
        if (!skip[c]) {
          const float diff = ob[o] - bw[o];
          v = (float)((double)v + ((double)strength * 0.5) * (double)diff);
        }
        db[o] = v;
      }
    }
    /* MacCormackClampMAC, :748-774. */
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      if (on_border(d, k, j, i, 1)) continue;
      const long c0 = ((long)k * d->ny + j) * d->nx + i;
      db[c0] = clamp_component_mac(ob, d, db[c0], fw[c0], k, j, i,
                                   v3_scale(mac_at_x(U, d, b, k, j, i), dt));
      db[n + c0] = clamp_component_mac(ob + n, d, db[n + c0], fw[n + c0], k, j, i,
                                       v3_scale(mac_at_y(U, d, b, k, j, i), dt));
      if (d->is3d)
        db[2 * n + c0] = clamp_component_mac(ob + 2 * n, d, db[2 * n + c0], fw[2 * n + c0], k, j, i,
                                             v3_scale(mac_at_z(U, d, b, k, j, i), dt));
    }
  }
}

/* ------------------------------------------------------------------------------------
 * setWallBcsForward: third_party/tfluids.cc:926-1002.
 * `zero` selects what a zeroed component becomes: the operator writes literal 0; the
 * nn wrapper tfluids/set_wall_bcs.lua:29-48 multiplies by a {0,1} mask (U * 0).
 * ---------------------------------------------------------------------------------- */
void orc_set_wall_bcs(float* U, const float* flags, const orc_dims* d, int as_mask_multiply) {
  const int nc = nchan_vel(d);
  const long n = cells(d);
  for (int b = 0; b < d->nb; b++) {
    float* ub = U + (long)b * nc * n;
    int k, j, i;
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      const int cf = is_fluid(flags, d, b, k, j, i), co = is_obstacle(flags, d, b, k, j, i);
      if (!cf && !co) continue;
      int z[3] = { 0, 0, 0 };
      if (i > 0 && is_obstacle(flags, d, b, k, j, i - 1)) z[0] = 1;
      if (i > 0 && co && is_fluid(flags, d, b, k, j, i - 1)) z[0] = 1;
      if (j > 0 && is_obstacle(flags, d, b, k, j - 1, i)) z[1] = 1;
      if (j > 0 && co && is_fluid(flags, d, b, k, j - 1, i)) z[1] = 1;
      /* The reference indexes channel 2 here even for 2-D grids (k > 0 is never true
       * when zsize == 1, so the access never happens), :967-972. */
      if (k > 0 && is_obstacle(flags, d, b, k - 1, j, i)) z[2] = 1;
      if (k > 0 && co && is_fluid(flags, d, b, k - 1, j, i)) z[2] = 1;
      if (cf) {
        if ((i > 0 && is_stick(flags, d, b, k, j, i - 1)) ||
            (i < d->nx - 1 && is_stick(flags, d, b, k, j, i + 1))) { z[1] = 1; if (d->is3d) z[2] = 1; }
        if ((j > 0 && is_stick(flags, d, b, k, j - 1, i)) ||
            (j < d->ny - 1 && is_stick(flags, d, b, k, j + 1, i))) { z[0] = 1; if (d->is3d) z[2] = 1; }
        if (d->is3d && ((k > 0 && is_stick(flags, d, b, k - 1, j, i)) ||
                        (k < d->nz - 1 && is_stick(flags, d, b, k + 1, j, i)))) { z[0] = 1; z[1] = 1; }
      }
      const long c0 = ((long)k * d->ny + j) * d->nx + i;
      for (int c = 0; c < nc; c++)
        if (z[c]) ub[(long)c * n + c0] = as_mask_multiply ? ub[(long)c * n + c0] * 0.0f : 0.0f;
    }
  }
}

/* velocityDivergenceForward: third_party/tfluids.cc:1008-1066 (note: = -div). */
void orc_velocity_divergence(const float* U, const float* flags, float* div, const orc_dims* d) {
  const int nc = nchan_vel(d);
  const long n = cells(d);
  for (int b = 0; b < d->nb; b++) {
    const float* ub = U + (long)b * nc * n;
    int k, j, i;
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      const long c0 = ((long)k * d->ny + j) * d->nx + i;
      float r = 0.0f;
      if (!on_border(d, k, j, i, 1) && is_fluid(flags, d, b, k, j, i)) {
        r = ub[c0] - ub[c0 + 1] + ub[n + c0] - ub[n + c0 + d->nx];
        if (d->is3d) r += (ub[2 * n + c0] - ub[2 * n + c0 + (long)d->nx * d->ny]);
      }
      div[(long)b * n + c0] = r;
    }
  }
}

/* velocityUpdateForward: third_party/tfluids.cc:1072-1156. */
void orc_velocity_update(float* U, const float* flags, const float* p, const orc_dims* d) {
  const int nc = nchan_vel(d);
  const long n = cells(d);
  const long sx = 1, sy = d->nx, sz = (long)d->nx * d->ny;
  for (int b = 0; b < d->nb; b++) {
    float* ub = U + (long)b * nc * n;
    const float* pb = p + (long)b * n;
    int k, j, i;
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      if (on_border(d, k, j, i, 1)) continue;
      const long c0 = ((long)k * d->ny + j) * d->nx + i;
      const long st[3] = { sx, sy, sz };
      const int nbf[3] = { is_fluid(flags, d, b, k, j, i - 1), is_fluid(flags, d, b, k, j - 1, i),
                           d->is3d ? is_fluid(flags, d, b, k - 1, j, i) : 0 };
      const int nbe[3] = { is_empty(flags, d, b, k, j, i - 1), is_empty(flags, d, b, k, j - 1, i),
                           d->is3d ? is_empty(flags, d, b, k - 1, j, i) : 0 };
      if (is_fluid(flags, d, b, k, j, i)) {
        for (int c = 0; c < nc; c++) if (nbf[c]) ub[(long)c * n + c0] -= (pb[c0] - pb[c0 - st[c]]);
        for (int c = 0; c < nc; c++) if (nbe[c]) ub[(long)c * n + c0] -= pb[c0];
      } else if (is_empty(flags, d, b, k, j, i) && !is_outflow(flags, d, b, k, j, i)) {
        for (int c = 0; c < nc; c++) {
          if (nbf[c]) ub[(long)c * n + c0] += pb[c0 - st[c]];
          else ub[(long)c * n + c0] = 0.0f;
        }
      }
    }
  }
}

/* addBuoyancy: third_party/tfluids.cc:1162-1233. */
void orc_add_buoyancy(float* U, const float* flags, const float* density, const float* gravity,
                      float dt, const orc_dims* d) {
  const int nc = nchan_vel(d);
  const long n = cells(d);
  const float scale = dt / orc_get_dx(d);
  const float str[3] = { (-gravity[0]) * scale, (-gravity[1]) * scale, (-gravity[2]) * scale };
  const long st[3] = { 1, d->nx, (long)d->nx * d->ny };
  for (int b = 0; b < d->nb; b++) {
    float* ub = U + (long)b * nc * n;
    const float* rb = density + (long)b * n;
    int k, j, i;
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      if (on_border(d, k, j, i, 1) || !is_fluid(flags, d, b, k, j, i)) continue;
      const long c0 = ((long)k * d->ny + j) * d->nx + i;
      const int nbf[3] = { is_fluid(flags, d, b, k, j, i - 1), is_fluid(flags, d, b, k, j - 1, i),
                           d->is3d ? is_fluid(flags, d, b, k - 1, j, i) : 0 };
      for (int c = 0; c < nc; c++)
        if (nbf[c]) ub[(long)c * n + c0] += (0.5f * str[c] * (rb[c0] + rb[c0 - st[c]]));
    }
  }
}

/* addGravity: third_party/tfluids.cc:1239-1306. */
void orc_add_gravity(float* U, const float* flags, const float* gravity, float dt,
                     const orc_dims* d) {
  const int nc = nchan_vel(d);
  const long n = cells(d);
  const float scale = dt / orc_get_dx(d);
  const float f[3] = { gravity[0] * scale, gravity[1] * scale, gravity[2] * scale };
  for (int b = 0; b < d->nb; b++) {
    float* ub = U + (long)b * nc * n;
    int k, j, i;
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      if (on_border(d, k, j, i, 1)) continue;
      const int cf = is_fluid(flags, d, b, k, j, i), ce = is_empty(flags, d, b, k, j, i);
      if (!cf && !ce) continue;
      const long c0 = ((long)k * d->ny + j) * d->nx + i;
      if (is_fluid(flags, d, b, k, j, i - 1) || (cf && is_empty(flags, d, b, k, j, i - 1))) ub[c0] += f[0];
      if (is_fluid(flags, d, b, k, j - 1, i) || (cf && is_empty(flags, d, b, k, j - 1, i))) ub[n + c0] += f[1];
      if (d->is3d && (is_fluid(flags, d, b, k - 1, j, i) || (cf && is_empty(flags, d, b, k - 1, j, i))))
        ub[2 * n + c0] += f[2];
    }
  }
}

/* vorticityConfinement: third_party/tfluids.cc:1312-1458 (4 passes; curl is always
 * stored with 3 channels, grid.cc:497-515).  Scratch: centered [b][2|3], curl [b][3],
 * curl_norm [b][1], force [b][2|3]. */
void orc_vorticity_confinement(float* U, const float* flags, float strength, const orc_dims* d,
                               float* centered, float* curl, float* curl_norm, float* force) {
  const int nc = nchan_vel(d);
  const long n = cells(d);
  const long sy = d->nx, sz = (long)d->nx * d->ny;
  for (int b = 0; b < d->nb; b++) {
    float* ub = U + (long)b * nc * n;
    float* ce = centered + (long)b * nc * n;
    float* cu = curl + (long)b * 3 * n;
    float* cn = curl_norm + (long)b * n;
    float* fo = force + (long)b * nc * n;
    int k, j, i;
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      const long c0 = ((long)k * d->ny + j) * d->nx + i;
      v3 v = { 0.0f, 0.0f, 0.0f };
      if (!on_border(d, k, j, i, 1)) v = mac_centered(U, d, b, k, j, i);
      ce[c0] = v.x; ce[n + c0] = v.y; if (d->is3d) ce[2 * n + c0] = v.z;
    }
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      const long c0 = ((long)k * d->ny + j) * d->nx + i;
      v3 w = { 0.0f, 0.0f, 0.0f };
      float nrm = 0.0f;
      if (!on_border(d, k, j, i, 1)) {
        w.z = 0.5f * ((ce[n + c0 + 1] - ce[n + c0 - 1]) - (ce[c0 + sy] - ce[c0 - sy]));
        if (d->is3d) {
          w.x = 0.5f * ((ce[2 * n + c0 + sy] - ce[2 * n + c0 - sy]) - (ce[n + c0 + sz] - ce[n + c0 - sz]));
          w.y = 0.5f * ((ce[c0 + sz] - ce[c0 - sz]) - (ce[2 * n + c0 + 1] - ce[2 * n + c0 - 1]));
        }
        nrm = v3_norm(w);
      }
      cu[c0] = w.x; cu[n + c0] = w.y; cu[2 * n + c0] = w.z;
      cn[c0] = nrm;
    }
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      const long c0 = ((long)k * d->ny + j) * d->nx + i;
      v3 f = { 0.0f, 0.0f, 0.0f };
      if (!on_border(d, k, j, i, 1)) {
        v3 g = { 0.0f, 0.0f, 0.0f };
        g.x = 0.5f * (cn[c0 + 1] - cn[c0 - 1]);
        g.y = 0.5f * (cn[c0 + sy] - cn[c0 - sy]);
        if (d->is3d) g.z = 0.5f * (cn[c0 + sz] - cn[c0 - sz]);
        const float gn = v3_norm(g);                                   /* normalize, vec3.h:129-141 */
        if (gn > 1e-6f) { g.x /= gn; g.y /= gn; g.z /= gn; } else { g.x = g.y = g.z = 0.0f; }
        const v3 w = { cu[c0], cu[n + c0], cu[2 * n + c0] };
        f.x = ((g.y * w.z) - (g.z * w.y)) * strength;
        f.y = ((g.z * w.x) - (g.x * w.z)) * strength;
        f.z = ((g.x * w.y) - (g.y * w.x)) * strength;
      }
      fo[c0] = f.x; fo[n + c0] = f.y; if (d->is3d) fo[2 * n + c0] = f.z;
    }
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      if (on_border(d, k, j, i, 1)) continue;                          /* CPU caller guard :1446-1451 */
      const int cf = is_fluid(flags, d, b, k, j, i), cem = is_empty(flags, d, b, k, j, i);
      if (!cf && !cem) continue;
      const long c0 = ((long)k * d->ny + j) * d->nx + i;
      if (is_fluid(flags, d, b, k, j, i - 1) || (cf && is_empty(flags, d, b, k, j, i - 1)))
        ub[c0] += (0.5f * (fo[c0 - 1] + fo[c0]));
      if (is_fluid(flags, d, b, k, j - 1, i) || (cf && is_empty(flags, d, b, k, j - 1, i)))
        ub[n + c0] += (0.5f * (fo[n + c0 - sy] + fo[n + c0]));
      if (d->is3d && (is_fluid(flags, d, b, k - 1, j, i) || (cf && is_empty(flags, d, b, k - 1, j, i))))
        ub[2 * n + c0] += (0.5f * (fo[2 * n + c0 - sz] + fo[2 * n + c0]));
    }
  }
}

/* ------------------------------------------------------------------------------------
 * solveLinearSystemJacobi: restated from the CUDA kernel + host loop
 * tfluids/generic/tfluids.cu:1765-1927 in IEEE float (no CPU reference exists).
 * p_prev is scratch of the same size as p.  Returns the residual
 * max_b ||p - p_prev||_2 of the last executed iteration (accumulated in double).
 * ---------------------------------------------------------------------------------- */
float orc_jacobi(float* p, const float* flags, const float* div, const orc_dims* d,
                 float p_tol, int max_iter, float* p_prev, int* iters_done) {
  const long n = cells(d);
  const long sy = d->nx, sz = (long)d->nx * d->ny;
  memset(p, 0, sizeof(float) * n * d->nb);                             /* :1854-1855 */
  memset(p_prev, 0, sizeof(float) * n * d->nb);
  float* cur = p;
  float* prev = p_prev;
  const float denom = d->is3d ? 6.0f : 4.0f;
  float residual = 0.0f;
  int iter = 0;
  for (;;) {
    for (int b = 0; b < d->nb; b++) {
      int k, j, i;
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
      for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
        const long c = (long)b * n + ((long)k * d->ny + j) * d->nx + i;
        if (on_border(d, k, j, i, 1) || is_obstacle(flags, d, b, k, j, i)) { cur[c] = 0.0f; continue; }
        const float pc = prev[c];
        float p1 = prev[c - 1], p2 = prev[c + 1], p3 = prev[c - sy], p4 = prev[c + sy];
        float p5 = d->is3d ? prev[c - sz] : 0.0f, p6 = d->is3d ? prev[c + sz] : 0.0f;
        if (is_obstacle(flags, d, b, k, j, i - 1)) p1 = pc;
        if (is_obstacle(flags, d, b, k, j, i + 1)) p2 = pc;
        if (is_obstacle(flags, d, b, k, j - 1, i)) p3 = pc;
        if (is_obstacle(flags, d, b, k, j + 1, i)) p4 = pc;
        if (d->is3d && is_obstacle(flags, d, b, k - 1, j, i)) p5 = pc;
        if (d->is3d && is_obstacle(flags, d, b, k + 1, j, i)) p6 = pc;
        cur[c] = (p1 + p2 + p3 + p4 + p5 + p6 + div[c]) / denom;
      }
    }
    double worst = 0.0;                                                /* :1879-1886 */
    for (int b = 0; b < d->nb; b++) {
      double acc = 0.0;
      long c;
#pragma omp parallel for reduction(+ : acc) schedule(static)
      for (c = 0; c < n; c++) {
        const float df = p[(long)b * n + c] - p_prev[(long)b * n + c];
        acc += (double)df * (double)df;
      }
      const double nr = sqrt(acc);
      if (nr > worst) worst = nr;
    }
    residual = (float)worst;
    if (residual < p_tol) break;                                       /* :1894-1900 */
    iter++;
    if (iter >= max_iter) break;                                       /* :1902-1909 */
    float* t = cur; cur = prev; prev = t;
  }
  if (cur == p_prev) memcpy(p, p_prev, sizeof(float) * n * d->nb);      /* :1919-1921 */
  if (iters_done) *iters_done = iter;
  return residual;
}

/* ------------------------------------------------------------------------------------
 * PCG pressure solve (a13).  The reference implementation is CUDA-only and sits on cuSPARSE /
 * cuBLAS calls that CUDA 12 no longer has (generic/tfluids.cu:1245-1759), so this is a
 * restatement of its ALGORITHM: per batch, per connected component of fluid cells
 * (find_connected_fluid_components.cc:17-82; the labelling below follows that flood fill and is
 * pinned against the compiled reference), the 7-point system of setupLaplacian (:909-1095: diag =
 * number of non-obstacle neighbours, -1 towards fluid neighbours), Golub & Van Loan PCG with
 * x0 = 0 (:1560-1725), preconditioner none / ilu0 / ic0 in the system's lexicographic order,
 * components of one cell skipped, fewer than 5 cells un-preconditioned (:1386-1404), mean of the
 * solution removed per component (:1735-1740).  For a symmetric matrix ILU(0) = L D L^T and
 * IC(0) = (L D^1/2)(L D^1/2)^T are the same operator, so both names share one factor.
 * Vectors are float, dot products accumulate in double (cuBLAS's order is unspecified).
 * Parity for this function is by property, as in the reference's own test
 * (test_tfluids.lua:836-905): residual < 2 tol, max |div| after the velocity update < 1e-4.
 * ---------------------------------------------------------------------------------- */
/* components: [nz][ny][nx] int32 for batch b, -1 for non-fluid; sizes: up to max_sizes entries.
 * Returns the number of components.  Same numbering as the reference: components are numbered
 * in the order a lexicographic scan (k, j, i) first meets them. */
int orc_find_components(const float* flags, const orc_dims* d, int b, int* comp, int* sizes, int max_sizes) {
  const long n = cells(d);
  for (long c = 0; c < n; c++) comp[c] = -1;
  long* stack = (long*)malloc(sizeof(long) * (size_t)(n > 0 ? n : 1));
  int cur = 0;
  const long sy = d->nx, sz = (long)d->nx * d->ny;
  for (int k = 0; k < d->nz; k++) for (int j = 0; j < d->ny; j++) for (int i = 0; i < d->nx; i++) {
    const long c0 = ((long)k * d->ny + j) * d->nx + i;
    if (comp[c0] != -1 || !is_fluid(flags, d, b, k, j, i)) continue;
    long top = 0, count = 0;
    stack[top++] = c0;
    comp[c0] = cur;
    while (top > 0) {
      const long c = stack[--top];
      count++;
      const int ci = (int)(c % d->nx), cj = (int)((c / d->nx) % d->ny), ck = (int)(c / sz);
      const int di[6] = {-1, 1, 0, 0, 0, 0}, dj[6] = {0, 0, -1, 1, 0, 0}, dk[6] = {0, 0, 0, 0, -1, 1};
      const int nn = d->is3d ? 6 : 4;
      for (int q = 0; q < nn; q++) {
        const int ii = ci + di[q], jj = cj + dj[q], kk = ck + dk[q];
        if (ii < 0 || ii >= d->nx || jj < 0 || jj >= d->ny || kk < 0 || kk >= d->nz) continue;
        const long cn = c + di[q] + dj[q] * sy + dk[q] * sz;
        if (comp[cn] == -1 && is_fluid(flags, d, b, kk, jj, ii)) { comp[cn] = cur; stack[top++] = cn; }
      }
    }
    if (cur < max_sizes) sizes[cur] = (int)count;
    cur++;
  }
  free(stack);
  return cur;
}

static inline double clamp_eps(double v) {              /* clampToEpsilon, generic/tfluids.cu:1153-1163 */
  const double eps = 1.17549435e-38;                    /* std::numeric_limits<float>::min() */
  if (fabs(v) < eps) return v < 0 ? -eps : eps;
  return v;
}

/* precond: 0 none, 1 ilu0, 2 ic0.  Returns 0, or 1 when a fluid cell sits on the domain border
 * (the reference raises "Non fluid cell found in a connected component", :1083-1090).
 * residual: max over batches / components of ||r||_2 (-inf when nothing was solved, as :1343);
 * iters: largest iteration count of any component. */
int orc_pcg(float* p, const float* flags, const float* div, const orc_dims* d, int precond, float tol,
            int max_iter, float* residual, int* iters) {
  const long n = cells(d);
  const long sy = d->nx, sz = (long)d->nx * d->ny;
  memset(p, 0, sizeof(float) * n * d->nb);                                   /* :1337 */
  int* comp = (int*)malloc(sizeof(int) * (size_t)n);
  int* sizes = (int*)malloc(sizeof(int) * (size_t)(n / 1 + 1));
  long* cell_of = (long*)malloc(sizeof(long) * (size_t)n);
  int* sys = (int*)malloc(sizeof(int) * (size_t)n);
  float *x = (float*)malloc(4 * (size_t)n), *r = (float*)malloc(4 * (size_t)n), *z = (float*)malloc(4 * (size_t)n),
        *pp = (float*)malloc(4 * (size_t)n), *w = (float*)malloc(4 * (size_t)n), *pre = (float*)malloc(4 * (size_t)n),
        *diag = (float*)malloc(4 * (size_t)n);
  int (*nb6)[6] = (int (*)[6])malloc(sizeof(int) * 6 * (size_t)n);           /* xm, ym, zm, xp, yp, zp system index or -1 */
  float worst = -INFINITY;
  int worst_it = 0, rc = 0;
  for (int b = 0; b < d->nb && !rc; b++) {
    const int ncomp = orc_find_components(flags, d, b, comp, sizes, (int)n);
    for (int ic = 0; ic < ncomp && !rc; ic++) {
      if (sizes[ic] == 1) continue;                                           /* :1386-1392 */
      const int use_pre = precond != 0 && sizes[ic] >= 5;                     /* :1399-1401 */
      /* createReducedSystemIndices (:864-905): lexicographic numbering of the component. */
      long m = 0;
      for (long c = 0; c < n; c++) { sys[c] = -1; if (comp[c] == ic) { sys[c] = (int)m; cell_of[m++] = c; } }
      for (long q = 0; q < m && !rc; q++) {
        const long c = cell_of[q];
        const int i = (int)(c % d->nx), j = (int)((c / d->nx) % d->ny), k = (int)(c / sz);
        if (on_border(d, k, j, i, 1)) { rc = 1; break; }
        float dg = 0.0f;
        if (!is_obstacle(flags, d, b, k, j, i - 1)) dg += 1;
        if (!is_obstacle(flags, d, b, k, j, i + 1)) dg += 1;
        if (!is_obstacle(flags, d, b, k, j - 1, i)) dg += 1;
        if (!is_obstacle(flags, d, b, k, j + 1, i)) dg += 1;
        if (d->is3d && !is_obstacle(flags, d, b, k - 1, j, i)) dg += 1;
        if (d->is3d && !is_obstacle(flags, d, b, k + 1, j, i)) dg += 1;
        diag[q] = dg;
        nb6[q][0] = is_fluid(flags, d, b, k, j, i - 1) ? sys[c - 1] : -1;
        nb6[q][1] = is_fluid(flags, d, b, k, j - 1, i) ? sys[c - sy] : -1;
        nb6[q][2] = (d->is3d && is_fluid(flags, d, b, k - 1, j, i)) ? sys[c - sz] : -1;
        nb6[q][3] = is_fluid(flags, d, b, k, j, i + 1) ? sys[c + 1] : -1;
        nb6[q][4] = is_fluid(flags, d, b, k, j + 1, i) ? sys[c + sy] : -1;
        nb6[q][5] = (d->is3d && is_fluid(flags, d, b, k + 1, j, i)) ? sys[c + sz] : -1;
        x[q] = 0.0f;
        r[q] = div[(long)b * n + c];                                          /* copyDivergenceToSystem */
      }
      if (rc) break;
      if (use_pre) {
        /* IC(0) of the 7-point matrix: R_ii = sqrt(A_ii - sum_{k lower} R_ki^2), R_ij = -1 / R_ii; the
         * level-0 pattern has no other update terms.  pre = 1 / R_ii. */
        for (long q = 0; q < m; q++) {
          float e = diag[q];
          for (int t = 0; t < 3; t++) if (nb6[q][t] >= 0) { const float pk = pre[nb6[q][t]]; e = e - pk * pk; }
          if (!(e > 1e-6f * diag[q])) e = diag[q];                           /* vanishing pivot guard */
          pre[q] = 1.0f / sqrtf(e);
        }
      }
      double rr = 0.0, rr0 = 0.0, rz = 0.0, rz_old = 0.0;
      for (long q = 0; q < m; q++) rr += (double)r[q] * (double)r[q];
      int iter = 0;
      const double tol2 = (double)tol * (double)tol;
      while (rr > tol2 && iter <= max_iter) {                                 /* :1588 */
        if (use_pre) {
          for (long q = 0; q < m; q++) {                                      /* R^T y = r */
            float acc = r[q];
            for (int t = 0; t < 3; t++) if (nb6[q][t] >= 0) acc = acc + pre[nb6[q][t]] * z[nb6[q][t]];
            z[q] = acc * pre[q];
          }
          for (long q = m - 1; q >= 0; q--) {                                 /* R z = y */
            float acc = 0.0f;
            for (int t = 3; t < 6; t++) if (nb6[q][t] >= 0) acc = acc + z[nb6[q][t]];
            z[q] = (z[q] + pre[q] * acc) * pre[q];
          }
          rz_old = rz;
          rz = 0.0;
          for (long q = 0; q < m; q++) rz += (double)r[q] * (double)z[q];
        } else {
          for (long q = 0; q < m; q++) z[q] = r[q];
          rz_old = rr0;
          rz = rr;
        }
        iter++;
        if (iter == 1) {
          for (long q = 0; q < m; q++) pp[q] = z[q];
        } else {
          const float beta = (float)(rz / clamp_eps(rz_old));
          for (long q = 0; q < m; q++) pp[q] = z[q] + beta * pp[q];
        }
        double pw = 0.0;
        for (long q = 0; q < m; q++) {
          float acc = diag[q] * pp[q];
          for (int t = 0; t < 6; t++) if (nb6[q][t] >= 0) acc = acc - pp[nb6[q][t]];
          w[q] = acc;
          pw += (double)pp[q] * (double)acc;
        }
        const float alpha = (float)(rz / clamp_eps(pw));
        rr0 = rr;
        rr = 0.0;
        for (long q = 0; q < m; q++) {
          x[q] = x[q] + alpha * pp[q];
          r[q] = r[q] - alpha * w[q];
          rr += (double)r[q] * (double)r[q];
        }
      }
      const float res = (float)sqrt(rr);
      if (res > worst) worst = res;
      if (iter > worst_it) worst_it = iter;
      double sum = 0.0;
      for (long q = 0; q < m; q++) sum += (double)x[q];
      const float mean = (float)(sum / (double)m);
      for (long q = 0; q < m; q++) p[(long)b * n + cell_of[q]] = x[q] - mean;  /* copyPressureFromSystem */
    }
  }
  free(comp); free(sizes); free(cell_of); free(sys); free(x); free(r); free(z); free(pp); free(w); free(pre);
  free(diag); free(nb6);
  if (residual) *residual = worst;
  if (iters) *iters = worst_it;
  return rc;
}

/* ------------------------------------------------------------------------------------
 * Operators of tfluids/init.lua around the step ("next" rows of the scope table): all four have
 * CPU code in the reference (generic/tfluids.cc), so the restatements below are pinned bit-exactly
 * (the mean removal: to float rounding, its accumulation order is an OpenMP atomic).
 * ---------------------------------------------------------------------------------- */
/* volumetricUpSamplingNearestForward, generic/tfluids.cc:509-557: out[z][y][x] = in[z/r][y/r][x/r].
 * in: [nb][nf][nz][ny][nx], out: [nb][nf][nz*r][ny*r][nx*r]. */
void orc_upsample_nearest(const float* in, float* out, int nb, int nf, int nz, int ny, int nx, int ratio) {
  const long oz = (long)nz * ratio, oy = (long)ny * ratio, ox = (long)nx * ratio;
  long bf;
#pragma omp parallel for schedule(static)
  for (bf = 0; bf < (long)nb * nf; bf++)
    for (long z = 0; z < oz; z++) for (long y = 0; y < oy; y++) for (long x = 0; x < ox; x++)
      out[((bf * oz + z) * oy + y) * ox + x] = in[((bf * nz + z / ratio) * ny + y / ratio) * nx + x / ratio];
}

/* DoRectangularBlurAlongAxis, generic/tfluids.cc:641-668: running box sum with clamped edges. */
static void blur_axis(const float* src, int size, long stride, int rad, float* dst) {
  float val = src[0] * (float)(rad + 1);
  for (int i = 0; i < size && i < rad; i++) val += src[i * stride];
  const float mul_const = 1.0f / (float)(rad * 2 + 1);
  for (int i = 0; i < size; i++) {
    const int iminus = i - rad - 1 > 0 ? i - rad - 1 : 0;
    val -= src[iminus * stride];
    const int iplus = i + rad < size - 1 ? i + rad : size - 1;
    val += src[iplus * stride];
    dst[i * stride] = val * mul_const;
  }
}
/* rectangularBlur, generic/tfluids.cc:670-760: z (3-D only), then y, then x; tmp is scratch. */
void orc_rectangular_blur(const float* src, int rad, int is3d, float* dst, float* tmp, int nb, int nf, int nz,
                          int ny, int nx) {
  const long sy = nx, sz = (long)nx * ny, sf = sz * nz;
  const float* cur_src = src;
  float* cur_dst = is3d ? dst : tmp;
  long bf;
  if (is3d) {
#pragma omp parallel for schedule(static)
    for (bf = 0; bf < (long)nb * nf; bf++)
      for (long y = 0; y < ny; y++) for (long x = 0; x < nx; x++)
        blur_axis(cur_src + bf * sf + y * sy + x, nz, sz, rad, cur_dst + bf * sf + y * sy + x);
    cur_src = dst;
    cur_dst = tmp;
  }
#pragma omp parallel for schedule(static)
  for (bf = 0; bf < (long)nb * nf; bf++)
    for (long z = 0; z < nz; z++) for (long x = 0; x < nx; x++)
      blur_axis(cur_src + bf * sf + z * sz + x, ny, sy, rad, cur_dst + bf * sf + z * sz + x);
  cur_src = tmp;
  cur_dst = dst;
#pragma omp parallel for schedule(static)
  for (bf = 0; bf < (long)nb * nf; bf++)
    for (long z = 0; z < nz; z++) for (long y = 0; y < ny; y++)
      blur_axis(cur_src + bf * sf + z * sz + y * sy, nx, 1, rad, cur_dst + bf * sf + z * sz + y * sy);
}

/* signedDistanceField, generic/tfluids.cc:766-822: 0 in obstacles, else the distance to the nearest
 * obstacle cell inside a (2 r + 1)^3 window, capped at r. */
void orc_signed_distance_field(const float* flags, int search_rad, float* dst, const orc_dims* d) {
  const long n = cells(d);
  long bc;
#pragma omp parallel for schedule(static)
  for (bc = 0; bc < n * d->nb; bc++) {
    const int b = (int)(bc / n);
    const long c = bc % n;
    const int x = (int)(c % d->nx), y = (int)((c / d->nx) % d->ny), z = (int)(c / ((long)d->nx * d->ny));
    if (is_obstacle(flags, d, b, z, y, x)) { dst[bc] = 0.0f; continue; }
    float dist_sq = (float)(search_rad * search_rad);
    const int zmin = z - search_rad > 0 ? z - search_rad : 0, zmax = z + search_rad < d->nz - 1 ? z + search_rad : d->nz - 1;
    const int ymin = y - search_rad > 0 ? y - search_rad : 0, ymax = y + search_rad < d->ny - 1 ? y + search_rad : d->ny - 1;
    const int xmin = x - search_rad > 0 ? x - search_rad : 0, xmax = x + search_rad < d->nx - 1 ? x + search_rad : d->nx - 1;
    for (int zs = zmin; zs <= zmax; zs++) for (int ys = ymin; ys <= ymax; ys++) for (int xs = xmin; xs <= xmax; xs++)
      if (is_obstacle(flags, d, b, zs, ys, xs)) {
        const float cur = (float)((z - zs) * (z - zs) + (y - ys) * (y - ys) + (x - xs) * (x - xs));
        if (dist_sq > cur) dist_sq = cur;
      }
    dst[bc] = sqrtf(dist_sq);
  }
}

/* normalizePressureMean, generic/tfluids.cc:845-921: subtract from every fluid cell the mean of p over
 * its connected fluid component (components of one cell included). */
void orc_normalize_pressure_mean(float* p, const float* flags, const orc_dims* d) {
  const long n = cells(d);
  int* comp = (int*)malloc(sizeof(int) * (size_t)n);
  int* sizes = (int*)malloc(sizeof(int) * (size_t)(n + 1));
  for (int b = 0; b < d->nb; b++) {
    const int nc = orc_find_components(flags, d, b, comp, sizes, (int)n);
    float* mean = (float*)calloc((size_t)(nc > 0 ? nc : 1), sizeof(float));
    for (long c = 0; c < n; c++) if (comp[c] >= 0) mean[comp[c]] += p[(long)b * n + c];
    for (int q = 0; q < nc; q++) mean[q] = mean[q] / (float)sizes[q];
    for (long c = 0; c < n; c++) if (comp[c] >= 0) p[(long)b * n + c] = p[(long)b * n + c] - mean[comp[c]];
    free(mean);
  }
  free(comp);
  free(sizes);
}

/* ------------------------------------------------------------------------------------
 * Backward operators of the two tfluids modules and of the nearest up-sampling (generic/tfluids.cc:49-134,
 * 216-345, 563-635).  The reference scatters with OpenMP atomics; these are the equivalent gathers.  The
 * divergence gradient has at most two terms per face (a - b is order-independent): bit-exact.  The
 * velocity-update gradient sums up to nine terms in an unspecified order in the reference: compared at
 * float-rounding tolerance.
 * ---------------------------------------------------------------------------------- */
static inline int interior(const orc_dims* d, int k, int j, int i) { return !on_border(d, k, j, i, 1); }

void orc_velocity_divergence_backward(const float* flags, const float* grad_output, float* grad_u, const orc_dims* d) {
  const long n = cells(d);
  const int nc = nchan_vel(d);
  for (int b = 0; b < d->nb; b++) {
    int k, j, i;
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      const long c = ((long)k * d->ny + j) * d->nx + i;
      const int here = interior(d, k, j, i) && is_fluid(flags, d, b, k, j, i);
      const int di[3] = {1, 0, 0}, dj[3] = {0, 1, 0}, dk[3] = {0, 0, 1};
      for (int a = 0; a < nc; a++) {
        float g = 0.0f;
        if (here) g += grad_output[(long)b * n + c];
        const int ii = i - di[a], jj = j - dj[a], kk = k - dk[a];
        if (ii >= 0 && jj >= 0 && kk >= 0 && interior(d, kk, jj, ii) && is_fluid(flags, d, b, kk, jj, ii))
          g -= grad_output[(long)b * n + ((long)kk * d->ny + jj) * d->nx + ii];
        grad_u[((long)b * nc + a) * n + c] = g;
      }
    }
  }
}

void orc_velocity_update_backward(const float* flags, const float* grad_output, float* grad_p, const orc_dims* d) {
  const long n = cells(d);
  const int nc = nchan_vel(d);
  for (int b = 0; b < d->nb; b++) {
    int k, j, i;
#pragma omp parallel for collapse(3) private(k, j, i) schedule(static)
    for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) for (i = 0; i < d->nx; i++) {
      const long c = ((long)k * d->ny + j) * d->nx + i;
      const int di[3] = {1, 0, 0}, dj[3] = {0, 1, 0}, dk[3] = {0, 0, 1};
      const int xf = is_fluid(flags, d, b, k, j, i);
      float g = 0.0f;
      if (interior(d, k, j, i) && xf) {                      /* this cell's own faces */
        for (int a = 0; a < nc; a++)
          if (is_fluid(flags, d, b, k - dk[a], j - dj[a], i - di[a])) g -= grad_output[((long)b * nc + a) * n + c];
        for (int a = 0; a < nc; a++)
          if (is_empty(flags, d, b, k - dk[a], j - dj[a], i - di[a])) g -= grad_output[((long)b * nc + a) * n + c];
      }
      if (xf) {                                              /* faces of the +x / +y / +z neighbours */
        for (int a = 0; a < nc; a++) {
          const int ii = i + di[a], jj = j + dj[a], kk = k + dk[a];
          if (ii >= d->nx || jj >= d->ny || kk >= d->nz || !interior(d, kk, jj, ii)) continue;
          const int yf = is_fluid(flags, d, b, kk, jj, ii);
          const int ye = is_empty(flags, d, b, kk, jj, ii) && !is_outflow(flags, d, b, kk, jj, ii);
          if (yf || ye) g += grad_output[((long)b * nc + a) * n + ((long)kk * d->ny + jj) * d->nx + ii];
        }
      }
      grad_p[(long)b * n + c] = g;
    }
  }
}

/* volumetricUpSamplingNearestBackward, generic/tfluids.cc:563-635: sum of the ratio^3 window, z, y, x order. */
void orc_upsample_nearest_backward(const float* grad_out, float* grad_in, int nb, int nf, int nz, int ny, int nx,
                                   int ratio) {
  const long oz = (long)nz * ratio, oy = (long)ny * ratio, ox = (long)nx * ratio;
  long bf;
#pragma omp parallel for schedule(static)
  for (bf = 0; bf < (long)nb * nf; bf++)
    for (long z = 0; z < nz; z++) for (long y = 0; y < ny; y++) for (long x = 0; x < nx; x++) {
      float sum = 0;
      for (int zu = 0; zu < ratio; zu++) for (int yu = 0; yu < ratio; yu++) for (int xu = 0; xu < ratio; xu++)
        sum += grad_out[((bf * oz + z * ratio + zu) * oy + y * ratio + yu) * ox + x * ratio + xu];
      grad_in[((bf * nz + z) * ny + y) * nx + x] = sum;
    }
}

/* ------------------------------------------------------------------------------------
 * Lua-side pieces of the loop (lib/simulate.lua).
 * ---------------------------------------------------------------------------------- */
/* x = x * invMask + bc  (setConstVals, lib/simulate.lua:136-158: cmul then add). */
void orc_apply_bc(float* x, const float* inv_mask, const float* bc, long n) {
  long c;
#pragma omp parallel for schedule(static)
  for (c = 0; c < n; c++) { const float t = x[c] * inv_mask[c]; x[c] = t + bc[c]; }
}
/* U:clamp(-1e6, 1e6), lib/simulate.lua:326 (THTensor clamp: (x < lo) ? lo : (x > hi ? hi : x)). */
void orc_clamp(float* x, float lo, float hi, long n) {
  long c;
#pragma omp parallel for schedule(static)
  for (c = 0; c < n; c++) { const float v = x[c]; x[c] = (v < lo) ? lo : ((v > hi) ? hi : v); }
}

/* ------------------------------------------------------------------------------------
 * CNN projection (lib/model.lua:27-401).  The conv arithmetic itself is un-vendored
 * (cudnn.torch R7 / cuDNN 7.6.4): textbook cross-correlation, zero padding (k-1)/2,
 * stride 1, bias, optional ReLU.  "Parity unpinned" for this part (see header).
 * weights: [cout][cin][kz][ky][kx] (kz == 1 for 2-D), accumulation in double.
 * ---------------------------------------------------------------------------------- */
void orc_conv(const float* in, float* out, const float* w, const float* bias, const orc_dims* d,
              int cin, int cout, int ksize, int relu) {
  const long n = cells(d);
  const int kz = d->is3d ? ksize : 1, pz = (kz - 1) / 2, pk = (ksize - 1) / 2;
  for (int b = 0; b < d->nb; b++) {
    int o, k, j;
#pragma omp parallel for collapse(3) private(o, k, j) schedule(static)
    for (o = 0; o < cout; o++) for (k = 0; k < d->nz; k++) for (j = 0; j < d->ny; j++) {
      for (int i = 0; i < d->nx; i++) {
        double acc = (double)bias[o];
        for (int c = 0; c < cin; c++) {
          const float* ib = in + ((long)b * cin + c) * n;
          const float* wb = w + (((long)o * cin + c) * kz) * ksize * ksize;
          for (int dz = 0; dz < kz; dz++) {
            const int zz = k + dz - pz;
            if (zz < 0 || zz >= d->nz) continue;
            for (int dy = 0; dy < ksize; dy++) {
              const int yy = j + dy - pk;
              if (yy < 0 || yy >= d->ny) continue;
              for (int dx = 0; dx < ksize; dx++) {
                const int xx = i + dx - pk;
                if (xx < 0 || xx >= d->nx) continue;
                acc += (double)ib[((long)zz * d->ny + yy) * d->nx + xx] *
                       (double)wb[((long)dz * ksize + dy) * ksize + dx];
              }
            }
          }
        }
        float r = (float)acc;
        if (relu && r < 0.0f) r = 0.0f;
        out[((long)b * cout + o) * n + ((long)k * d->ny + j) * d->nx + i] = r;
      }
    }
  }
}

/* Sample standard deviation of one batch element, lib/modules/variance.lua:44-76 +
 * standard_deviation.lua: var = (n*sum(x^2) - (sum x)^2) / (n (n-1)), float tensor ops on
 * sums whose accumulation order is unspecified in the reference (accumulated in double
 * here, compared at 1e-5 relative). */
float orc_sample_std(const float* x, long n) {
  double s = 0.0, ss = 0.0;
  long c;
#pragma omp parallel for reduction(+ : s, ss) schedule(static)
  for (c = 0; c < n; c++) { const float sq = x[c] * x[c]; s += (double)x[c]; ss += (double)sq; }
  const float sum = (float)s, sumsq = (float)ss;
  float out = sumsq * (float)n;
  out = out + (-1.0f) * (sum * sum);
  out = out / (float)((double)n * (double)(n - 1));
  return sqrtf(out);
}
