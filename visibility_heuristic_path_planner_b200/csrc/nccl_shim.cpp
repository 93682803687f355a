// nccl_shim.cpp -- see nccl_shim.h
#include "nccl_shim.h"

#include <dlfcn.h>

#include <cstdlib>
#include <mutex>
#include <type_traits>

namespace {
VhpNccl g_nccl;
bool g_ok = false;
std::string g_why;
std::once_flag g_once;

void load() {
  void *h = nullptr;
  std::string tried;
  auto attempt = [&](const char *name, int flags) {
    if (h || !name || !*name) return;
    h = dlopen(name, flags);
    if (h) g_nccl.path = name;
    else tried += std::string(tried.empty() ? "" : "; ") + name;
  };
  attempt(std::getenv("VHP_NCCL_LIB"), RTLD_NOW | RTLD_GLOBAL);
  attempt("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD); // the copy this process already uses (torch)
  attempt("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  attempt("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    g_why = "NCCL not found (tried " + tried + ")";
    return;
  }
  bool all = true;
  auto sym = [&](auto &fn, const char *name) {
    fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(h, name));
    if (!fn) { all = false; g_why += std::string(" missing ") + name; }
  };
  sym(g_nccl.GetVersion, "ncclGetVersion");
  sym(g_nccl.GetUniqueId, "ncclGetUniqueId");
  sym(g_nccl.CommInitRank, "ncclCommInitRank");
  sym(g_nccl.CommSplit, "ncclCommSplit");
  sym(g_nccl.CommDestroy, "ncclCommDestroy");
  sym(g_nccl.Send, "ncclSend");
  sym(g_nccl.Recv, "ncclRecv");
  sym(g_nccl.AllGather, "ncclAllGather");
  sym(g_nccl.AllReduce, "ncclAllReduce");
  sym(g_nccl.GetErrorString, "ncclGetErrorString");
  g_ok = all;
}
} // namespace

const VhpNccl *vhp_nccl(std::string *why) {
  std::call_once(g_once, load);
  if (!g_ok && why) *why = g_why;
  return g_ok ? &g_nccl : nullptr;
}
