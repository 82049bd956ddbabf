// TEST INFRASTRUCTURE ONLY (oracle/_ref build). Not part of the product.
//
// Stand-in for lua.h / luaT.h: the reference's lua_CFunctions
// (torch/tfluids/generic/tfluids.cc:927-952 registers them) read positional
// arguments off the Lua stack. Here a "lua_State" is just a positional argument
// vector filled by ref_driver.cc, so the functions run unmodified.
#pragma once
#include <vector>
#include <string>
#include "TH.h"

struct RefShimArg {
  double num;
  void* ptr;
  std::string str;
  RefShimArg() : num(0), ptr(NULL) {}
};

struct lua_State {
  std::vector<RefShimArg> args;  // 1-based on the Lua side.
  double ret;
  lua_State() : ret(0) {}
};

typedef int (*lua_CFunction)(lua_State*);
struct luaL_Reg {
  const char* name;
  lua_CFunction func;
};

inline RefShimArg& refshim_arg(lua_State* L, int i) {
  if (i < 1 || i > (int)L->args.size()) throw RefShimError("bad lua stack index");
  return L->args[i - 1];
}
inline double lua_tonumber(lua_State* L, int i) { return refshim_arg(L, i).num; }
inline long lua_tointeger(lua_State* L, int i) {
  return (long)refshim_arg(L, i).num;
}
inline long luaL_checkinteger(lua_State* L, int i) { return lua_tointeger(L, i); }
inline int lua_toboolean(lua_State* L, int i) {
  return refshim_arg(L, i).num != 0.0;
}
inline int lua_isboolean(lua_State*, int) { return 1; }
inline const char* lua_tostring(lua_State* L, int i) {
  return refshim_arg(L, i).str.c_str();
}
inline void* luaT_checkudata(lua_State* L, int i, const char*) {
  void* p = refshim_arg(L, i).ptr;
  if (!p) throw RefShimError("luaT_checkudata: nil tensor");
  return p;
}
inline void lua_pushnumber(lua_State* L, double v) { L->ret = v; }
[[noreturn]] inline int luaL_error(lua_State*, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw RefShimError(buf);
}
inline void luaT_pushmetatable(lua_State*, const char*) {}
inline void luaT_registeratname(lua_State*, const luaL_Reg*, const char*) {}
