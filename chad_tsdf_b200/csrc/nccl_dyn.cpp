// see nccl_dyn.h
#include "nccl_dyn.h"

#include <dlfcn.h>

#include <cstdlib>

#include <mutex>
#include <string>

namespace chadgpu {

namespace {
NcclApi g_api{};
bool g_ok = false;
std::string g_why;
std::once_flag g_once;

void load() {
    // a libnccl.so.2 that is already mapped (PyTorch's) is found by its soname; otherwise the loader's search path (system NCCL).
    // CHAD_NCCL_LIB overrides.
    void* h = nullptr;
    if (const char* path = std::getenv("CHAD_NCCL_LIB")) h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) { g_why = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
    bool ok = true;
    auto sym = [&](const char* name) { void* p = dlsym(h, name); if (!p) { ok = false; g_why = std::string("libnccl.so.2 lacks ") + name; } return p; };
    g_api.GetVersion = reinterpret_cast<decltype(g_api.GetVersion)>(sym("ncclGetVersion"));
    g_api.GetUniqueId = reinterpret_cast<decltype(g_api.GetUniqueId)>(sym("ncclGetUniqueId"));
    g_api.CommInitRank = reinterpret_cast<decltype(g_api.CommInitRank)>(sym("ncclCommInitRank"));
    g_api.CommDestroy = reinterpret_cast<decltype(g_api.CommDestroy)>(sym("ncclCommDestroy"));
    g_api.CommAbort = reinterpret_cast<decltype(g_api.CommAbort)>(sym("ncclCommAbort"));
    g_api.GroupStart = reinterpret_cast<decltype(g_api.GroupStart)>(sym("ncclGroupStart"));
    g_api.GroupEnd = reinterpret_cast<decltype(g_api.GroupEnd)>(sym("ncclGroupEnd"));
    g_api.Send = reinterpret_cast<decltype(g_api.Send)>(sym("ncclSend"));
    g_api.Recv = reinterpret_cast<decltype(g_api.Recv)>(sym("ncclRecv"));
    g_api.AllGather = reinterpret_cast<decltype(g_api.AllGather)>(sym("ncclAllGather"));
    g_api.Broadcast = reinterpret_cast<decltype(g_api.Broadcast)>(sym("ncclBroadcast"));
    g_api.GetErrorString = reinterpret_cast<decltype(g_api.GetErrorString)>(sym("ncclGetErrorString"));
    g_ok = ok;
}
}  // namespace

const NcclApi* nccl_api(const char** why) {
    std::call_once(g_once, load);
    if (!g_ok && why) *why = g_why.c_str();
    return g_ok ? &g_api : nullptr;
}

}  // namespace chadgpu
