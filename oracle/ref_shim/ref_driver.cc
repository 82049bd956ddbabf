// TEST INFRASTRUCTURE ONLY (oracle/_ref build). Not part of the product.
//
// Builds the reference's OWN CPU implementation of the hot path, unmodified and
// in place from /root/reference/torch/tfluids (never copied into this repo), by
// mirroring the include order of torch/tfluids/init.cu:15-72 with the fake
// TH/luaT layer in this directory, and exposes one generic C entry point that
// fills the fake Lua stack and invokes a registered lua_CFunction by name
// (table tfluids_FloatMain__ / tfluids_DoubleMain__,
// torch/tfluids/generic/tfluids.cc:927-952).
//
// Output: oracle/_ref/libtfluids_ref.so (git-ignored; shipped to the GPU box).
#include <assert.h>
#include <algorithm>
#include <iostream>
#include <cstring>

#include <TH.h>
#include <luaT.h>
#include "third_party/cell_type.h"
#include "generic/stack_trace.cc"
#include "generic/int3.cu.h"
#include "generic/advect_type.h"
#include "generic/advect_type.cc"

inline int32_t clamp(const int32_t x, const int32_t low, const int32_t high) {
  return std::max<int32_t>(std::min<int32_t>(x, high), low);
}

#define torch_(NAME) TH_CONCAT_3(torch_, Real, NAME)
#define torch_Tensor TH_CONCAT_STRING_3(torch., Real, Tensor)
#define tfluids_(NAME) TH_CONCAT_3(tfluids_, Real, NAME)

#define real float
#define accreal double
#define Real Float
#define THInf FLT_MAX
#define TH_REAL_IS_FLOAT
#include "generic/vec3.cc"
#include "third_party/grid.cc"
#include "generic/find_connected_fluid_components.cc"
#include "generic/tfluids.cc"
#undef accreal
#undef real
#undef Real
#undef THInf
#undef TH_REAL_IS_FLOAT

#define real double
#define accreal double
#define Real Double
#define THInf DBL_MAX
#define TH_REAL_IS_DOUBLE
#include "generic/vec3.cc"
#include "third_party/grid.cc"
#include "generic/find_connected_fluid_components.cc"
#include "generic/tfluids.cc"
#undef accreal
#undef real
#undef Real
#undef THInf
#undef TH_REAL_IS_DOUBLE

// ---------------------------------------------------------------------------
// C entry points (ctypes).
// ---------------------------------------------------------------------------
extern "C" {

// kinds: 0 = number / boolean, 1 = float tensor, 2 = string, 3 = int tensor,
//        4 = double tensor.
// Tensor arguments are described by (data pointer, ndim, sizes[5]); they are
// viewed as contiguous.
struct ref_tensor_desc {
  void* data;
  int32_t ndim;
  int64_t size[5];
};

static void set_err(char* err, int errlen, const char* msg) {
  if (err && errlen > 0) {
    strncpy(err, msg, errlen - 1);
    err[errlen - 1] = 0;
  }
}

int ref_call(const char* fn_name, int use_double, int nargs, const int* kinds,
             const double* nums, const ref_tensor_desc* tensors,
             const char* const* strs, double* ret, char* err, int errlen) {
  const luaL_Reg* table = use_double ? tfluids_DoubleMain__ : tfluids_FloatMain__;
  lua_CFunction fn = NULL;
  for (const luaL_Reg* r = table; r->name != NULL; ++r) {
    if (strcmp(r->name, fn_name) == 0) {
      fn = r->func;
      break;
    }
  }
  if (!fn) {
    set_err(err, errlen, "ref_call: unknown function");
    return 2;
  }
  lua_State L;
  L.args.resize(nargs);
  std::vector<THFloatTensor> ft(nargs);
  std::vector<THDoubleTensor> dt(nargs);
  std::vector<THIntTensor> it(nargs);
  for (int a = 0; a < nargs; ++a) {
    long sz[5] = {1, 1, 1, 1, 1};
    if (kinds[a] == 1 || kinds[a] == 3 || kinds[a] == 4) {
      for (int d = 0; d < tensors[a].ndim; ++d) sz[d] = (long)tensors[a].size[d];
    }
    switch (kinds[a]) {
      case 0:
        L.args[a].num = nums[a];
        break;
      case 1:
        memset(&ft[a], 0, sizeof(ft[a]));
        ft[a].data = (float*)tensors[a].data;
        refshim_set_contiguous(&ft[a], tensors[a].ndim, sz);
        L.args[a].ptr = &ft[a];
        break;
      case 2:
        L.args[a].str = strs[a];
        break;
      case 3:
        memset(&it[a], 0, sizeof(it[a]));
        it[a].data = (int*)tensors[a].data;
        refshim_set_contiguous(&it[a], tensors[a].ndim, sz);
        L.args[a].ptr = &it[a];
        break;
      case 4:
        memset(&dt[a], 0, sizeof(dt[a]));
        dt[a].data = (double*)tensors[a].data;
        refshim_set_contiguous(&dt[a], tensors[a].ndim, sz);
        L.args[a].ptr = &dt[a];
        break;
      default:
        set_err(err, errlen, "ref_call: bad arg kind");
        return 2;
    }
  }
  try {
    const int nret = fn(&L);
    if (ret) *ret = (nret > 0) ? L.ret : 0.0;
  } catch (const std::exception& e) {
    set_err(err, errlen, e.what());
    return 1;
  }
  return 0;
}

// Direct access to the reference line trace (float instantiation),
// torch/tfluids/generic/calc_line_trace.cc:313-503. flags is [1][1][z][y][x].
// Returns 0/1 = hit flag, -1 on a reference hard error (THError).
int ref_calc_line_trace(const float* pos, const float* delta, float* flags,
                        int zsize, int ysize, int xsize, int is_3d,
                        float* new_pos, char* err, int errlen) {
  THFloatTensor tf;
  memset(&tf, 0, sizeof(tf));
  long sz[5] = {1, 1, zsize, ysize, xsize};
  tf.data = flags;
  refshim_set_contiguous(&tf, 5, sz);
  try {
    tfluids_FloatFlagGrid fg(&tf, is_3d != 0);
    tfluids_Floatvec3 p(pos[0], pos[1], pos[2]);
    tfluids_Floatvec3 d(delta[0], delta[1], delta[2]);
    tfluids_Floatvec3 np;
    const bool hit = calcLineTrace(p, d, fg, 0, &np, true);
    new_pos[0] = np.x;
    new_pos[1] = np.y;
    new_pos[2] = np.z;
    return hit ? 1 : 0;
  } catch (const std::exception& e) {
    set_err(err, errlen, e.what());
    return -1;
  }
}


// The reference's connected-component labelling of fluid cells (CPU code used by its CUDA PCG,
// torch/tfluids/generic/find_connected_fluid_components.cc:17-82).  flags is [nb][1][z][y][x];
// components is [z][y][x] int32 for batch `ibatch`.  Returns the component count, sizes[] filled
// up to max_sizes, -1 on a reference hard error.
int ref_find_connected_fluid_components(float* flags, int nbatch, int zsize, int ysize, int xsize, int is_3d,
                                        int ibatch, int* components, int* sizes, int max_sizes,
                                        char* err, int errlen) {
  THFloatTensor tf;
  memset(&tf, 0, sizeof(tf));
  long sz[5] = {nbatch, 1, zsize, ysize, xsize};
  tf.data = flags;
  refshim_set_contiguous(&tf, 5, sz);
  THIntTensor ti;
  memset(&ti, 0, sizeof(ti));
  long sz3[3] = {zsize, ysize, xsize};
  ti.data = components;
  refshim_set_contiguous(&ti, 3, sz3);
  try {
    tfluids_FloatFlagGrid fg(&tf, is_3d != 0);
    std::vector<int32_t> csz;
    const int n = findConnectedFluidComponents(fg, &ti, ibatch, &csz);
    for (int i = 0; i < n && i < max_sizes; i++) sizes[i] = csz[i];
    return n;
  } catch (const std::exception& e) {
    set_err(err, errlen, e.what());
    return -1;
  }
}

int ref_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
