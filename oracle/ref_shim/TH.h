// TEST INFRASTRUCTURE ONLY (oracle/_ref build). Not part of the product.
//
// Minimal stand-in for Torch7's TH.h so that the reference's own CPU sources
// (/root/reference/torch/tfluids/{generic,third_party}/*.cc) compile, in place
// and unmodified, without Torch7. Same trick as the reference's MATLAB MEX
// (torch/tfluids/generic/CalcLineTrace.cc:24-60, fake THTensor at :46-54).
// Only what those sources touch is provided.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <stdio.h>
#include <float.h>
#include <stdexcept>
#include <string>

#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif

#define TH_CONCAT_STRING_3(x, y, z) TH_CONCAT_STRING_3_EXPAND(x, y, z)
#define TH_CONCAT_STRING_3_EXPAND(x, y, z) #x #y #z
#define TH_CONCAT_3(x, y, z) TH_CONCAT_3_EXPAND(x, y, z)
#define TH_CONCAT_3_EXPAND(x, y, z) x##y##z
#define TH_CONCAT_4(x, y, z, w) TH_CONCAT_4_EXPAND(x, y, z, w)
#define TH_CONCAT_4_EXPAND(x, y, z, w) x##y##z##w

#define THTensor TH_CONCAT_3(TH, Real, Tensor)
#define THTensor_(NAME) TH_CONCAT_4(TH, Real, Tensor_, NAME)

struct RefShimError : public std::runtime_error {
  explicit RefShimError(const std::string& s) : std::runtime_error(s) {}
};

[[noreturn]] inline void THError(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw RefShimError(buf);
}

template <typename T>
struct RefShimTensor {
  long size[5];
  long stride[5];
  int nDimension;
  T* data;
  bool owns;
};

typedef RefShimTensor<float> THFloatTensor;
typedef RefShimTensor<double> THDoubleTensor;
typedef RefShimTensor<int> THIntTensor;

template <typename T>
inline void refshim_set_contiguous(RefShimTensor<T>* t, int nd, const long* sz) {
  t->nDimension = nd;
  long st = 1;
  for (int d = nd - 1; d >= 0; --d) {
    t->size[d] = sz[d];
    t->stride[d] = st;
    st *= sz[d];
  }
}
template <typename T>
inline long refshim_numel(const RefShimTensor<T>* t) {
  long n = 1;
  for (int d = 0; d < t->nDimension; ++d) n *= t->size[d];
  return t->nDimension == 0 ? 0 : n;
}
template <typename T>
inline bool refshim_is_contiguous(const RefShimTensor<T>* t) {
  long st = 1;
  for (int d = t->nDimension - 1; d >= 0; --d) {
    if (t->size[d] != 1 && t->stride[d] != st) return false;
    st *= t->size[d];
  }
  return true;
}
template <typename T>
inline RefShimTensor<T>* refshim_new() {
  RefShimTensor<T>* t = new RefShimTensor<T>();
  memset(t, 0, sizeof(*t));
  t->owns = true;
  return t;
}
template <typename T>
inline void refshim_free(RefShimTensor<T>* t) {
  if (t->owns && t->data) free(t->data);
  delete t;
}
template <typename T>
inline void refshim_resize(RefShimTensor<T>* t, int nd, const long* sz) {
  long n = 1;
  for (int d = 0; d < nd; ++d) n *= sz[d];
  if (refshim_numel(t) < n || t->data == NULL) {
    if (!t->owns && t->data != NULL) {
      throw RefShimError("refshim: cannot grow a borrowed tensor");
    }
    t->data = (T*)realloc(t->data, sizeof(T) * n);
    t->owns = true;
  }
  refshim_set_contiguous(t, nd, sz);
}

#define REFSHIM_DEFINE_REAL_API(PFX, T)                                        \
  inline T* PFX##_data(const RefShimTensor<T>* t) { return t->data; }          \
  inline long PFX##_numel(const RefShimTensor<T>* t) { return refshim_numel(t); } \
  inline int PFX##_isContiguous(const RefShimTensor<T>* t) {                   \
    return refshim_is_contiguous(t);                                           \
  }                                                                            \
  inline RefShimTensor<T>* PFX##_new() { return refshim_new<T>(); }            \
  inline void PFX##_free(RefShimTensor<T>* t) { refshim_free(t); }             \
  inline void PFX##_resize1d(RefShimTensor<T>* t, long s0) {                   \
    long sz[1] = {s0};                                                         \
    refshim_resize(t, 1, sz);                                                  \
  }                                                                            \
  inline void PFX##_resize4d(RefShimTensor<T>* t, long s0, long s1, long s2,   \
                             long s3) {                                        \
    long sz[4] = {s0, s1, s2, s3};                                             \
    refshim_resize(t, 4, sz);                                                  \
  }                                                                            \
  inline void PFX##_fill(RefShimTensor<T>* t, T v) {                           \
    long n = refshim_numel(t);                                                 \
    for (long i = 0; i < n; ++i) t->data[i] = v;                               \
  }

REFSHIM_DEFINE_REAL_API(THFloatTensor, float)
REFSHIM_DEFINE_REAL_API(THDoubleTensor, double)
REFSHIM_DEFINE_REAL_API(THIntTensor, int)

inline int THIntTensor_get3d(const THIntTensor* t, long a, long b, long c) {
  return t->data[a * t->stride[0] + b * t->stride[1] + c * t->stride[2]];
}
inline void THIntTensor_set3d(THIntTensor* t, long a, long b, long c, int v) {
  t->data[a * t->stride[0] + b * t->stride[1] + c * t->stride[2]] = v;
}
inline int THIntTensor_get4d(const THIntTensor* t, long a, long b, long c,
                             long d) {
  return t->data[a * t->stride[0] + b * t->stride[1] + c * t->stride[2] +
                 d * t->stride[3]];
}
inline THIntTensor* THIntTensor_newSelect(THIntTensor* t, int dim, long idx) {
  THIntTensor* r = new THIntTensor();
  memset(r, 0, sizeof(*r));
  r->owns = false;
  r->data = t->data + idx * t->stride[dim];
  int o = 0;
  for (int d = 0; d < t->nDimension; ++d) {
    if (d == dim) continue;
    r->size[o] = t->size[d];
    r->stride[o] = t->stride[d];
    ++o;
  }
  r->nDimension = t->nDimension - 1;
  return r;
}
