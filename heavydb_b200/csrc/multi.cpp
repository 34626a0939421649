/*
 * multi.cpp — run-time binding to NCCL for the cross-GPU merge of partial aggregate tables.
 *
 * Replaces the HOST-side merge of per-device result sets (Executor::reduceMultiDeviceResults -> ResultSetManager::reduce ->
 * ResultSetStorage::reduce, QueryEngine/Execute.cpp:1696,1772-1792; ResultSetReduction.cpp:203-396): the per-device tables
 * never leave HBM, they are merged by collectives over NVLink on the stream that produced them.
 */
#include "multi.h"

#include <dlfcn.h>
#include <stdlib.h>

#include <mutex>

namespace b2q {

const NcclApi* nccl_api(std::string* why) {
  static NcclApi api;
  static bool ok = false;
  static std::string err;
  static std::once_flag once;
  std::call_once(once, []() {
    /* Prefer the copy already in the process: a second library with the same SONAME cannot be loaded next to it, and the
     * first one loaded wins for everybody — a host that bundles its own NCCL (torch does) must therefore have loaded it before
     * the first b2q_comm_* call (heavydb_b200/executor.py imports torch first for that reason).  B2Q_NCCL_LIB names a file to use. */
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_LOCAL);
    if (!h) { const char* env = getenv("B2Q_NCCL_LIB"); if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_LOCAL); }
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) { const char* e = dlerror(); err = std::string("cannot load libnccl.so.2: ") + (e ? e : "?"); return; }
    auto sym = [&](const char* n) -> void* {
      void* p = dlsym(h, n);
      if (!p && err.empty()) err = std::string("libnccl lacks ") + n;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    ok = err.empty();
  });
  if (!ok && why) *why = err;
  return ok ? &api : nullptr;
}

}  // namespace b2q
