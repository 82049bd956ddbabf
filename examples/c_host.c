/* The C ABI from plain C (what a LuaJIT ffi.cdef, a cgo binding or any C host sees): one fluid step on device
 * buffers obtained through the library itself -- no torch, no C++.  Not part of the product; tests/test_abi.py
 * compiles it with `gcc -std=c99 -fsyntax-only` to keep include/tfl.h valid C, and it runs on a GPU box as
 *   gcc -std=c99 -Iinclude examples/c_host.c -o c_host -ldl && ./c_host fluidnet_b200/libtfl.so
 * (the library is loaded with dlopen so that the example needs no link-time dependency either). */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tfl.h"

#define LOAD(name) name##_fn = (name##_t)dlsym(lib, #name); if (!name##_fn) { fprintf(stderr, "missing %s\n", #name); return 2; }

typedef int (*tfl_create_t)(tfl_ctx**, int);
typedef void (*tfl_destroy_t)(tfl_ctx*);
typedef const char* (*tfl_last_error_t)(const tfl_ctx*);
typedef int (*tfl_alloc_t)(tfl_ctx*, size_t, void**);
typedef int (*tfl_free_t)(tfl_ctx*, void*);
typedef int (*tfl_memcpy_h2d_t)(tfl_ctx*, void*, const void*, size_t);
typedef int (*tfl_memcpy_d2h_t)(tfl_ctx*, void*, const void*, size_t);
typedef int (*tfl_sync_t)(tfl_ctx*);
typedef int (*tfl_empty_domain_t)(tfl_ctx*, const tfl_grid*, int, int);
typedef int (*tfl_simulate_step_t)(tfl_ctx*, const tfl_state*, const tfl_mconf*, tfl_cnn*);

int main(int argc, char** argv) {
  void* lib = dlopen(argc > 1 ? argv[1] : "libtfl.so", RTLD_NOW);
  tfl_create_t tfl_create_fn; tfl_destroy_t tfl_destroy_fn; tfl_last_error_t tfl_last_error_fn;
  tfl_alloc_t tfl_alloc_fn; tfl_free_t tfl_free_fn; tfl_memcpy_h2d_t tfl_memcpy_h2d_fn; tfl_memcpy_d2h_t tfl_memcpy_d2h_fn;
  tfl_sync_t tfl_sync_fn; tfl_empty_domain_t tfl_empty_domain_fn; tfl_simulate_step_t tfl_simulate_step_fn;
  tfl_ctx* ctx = NULL;
  tfl_state st;
  tfl_mconf mc;
  const int n = 32;
  const size_t cells = (size_t)n * n * n;
  float* host;
  size_t i;
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  LOAD(tfl_create) LOAD(tfl_destroy) LOAD(tfl_last_error) LOAD(tfl_alloc) LOAD(tfl_free) LOAD(tfl_memcpy_h2d)
  LOAD(tfl_memcpy_d2h) LOAD(tfl_sync) LOAD(tfl_empty_domain) LOAD(tfl_simulate_step)
  if (tfl_create_fn(&ctx, 0)) { fprintf(stderr, "tfl_create failed (no CUDA device?)\n"); return 1; }
  memset(&st, 0, sizeof(st));
  memset(&mc, 0, sizeof(mc));
#define GRID(g, channels)                                                           \
  do {                                                                              \
    void* p_ = NULL;                                                                \
    (g).nb = 1; (g).nc = (channels); (g).nz = n; (g).ny = n; (g).nx = n;            \
    if (tfl_alloc_fn(ctx, cells * (channels) * sizeof(float), &p_)) goto fail;      \
    (g).data = (float*)p_;                                                          \
  } while (0)
  GRID(st.p, 1); GRID(st.U, 3); GRID(st.flags, 1); GRID(st.density, 1); GRID(st.div, 1);
  host = (float*)calloc(cells * 3, sizeof(float));
  for (i = 0; i < cells; i++) host[cells + i] = 0.5f;                       /* a uniform upward draught */
  if (tfl_memcpy_h2d_fn(ctx, st.U.data, host, cells * 3 * sizeof(float))) goto fail;
  memset(host, 0, cells * sizeof(float));
  if (tfl_memcpy_h2d_fn(ctx, st.p.data, host, cells * sizeof(float))) goto fail;
  for (i = 0; i < cells; i++) host[i] = (float)(i % 7) / 7.0f;
  if (tfl_memcpy_h2d_fn(ctx, st.density.data, host, cells * sizeof(float))) goto fail;
  if (tfl_empty_domain_fn(ctx, &st.flags, 1, 1)) goto fail;                 /* tfluids.emptyDomain(flags, true, 1) */
  mc.dt = 0.1f;
  mc.advection_method = TFL_ADVECT_MACCORMACK_OURS;
  mc.maccormack_strength = 0.6f;
  mc.buoyancy_scale = 0.5;
  mc.gravity[1] = 1.0f;
  mc.vorticity_confinement_amp = 3.0;
  mc.sim_method = TFL_SIM_JACOBI;                                           /* no network needed for the example */
  mc.max_iter = 20;
  if (tfl_simulate_step_fn(ctx, &st, &mc, NULL)) goto fail;                 /* == tfluids.simulate(conf, mconf, batch) */
  if (tfl_memcpy_d2h_fn(ctx, host, st.density.data, cells * sizeof(float)) || tfl_sync_fn(ctx)) goto fail;
  printf("one %d^3 step done; density[centre] = %g\n", n, host[(cells + n * n + n) / 2]);
  tfl_free_fn(ctx, st.p.data); tfl_free_fn(ctx, st.U.data); tfl_free_fn(ctx, st.flags.data);
  tfl_free_fn(ctx, st.density.data); tfl_free_fn(ctx, st.div.data);
  free(host);
  tfl_destroy_fn(ctx);
  return 0;
fail:
  fprintf(stderr, "libtfl: %s\n", tfl_last_error_fn(ctx));
  return 1;
}
