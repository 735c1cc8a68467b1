#include "nccl_dl.hpp"

#include <dlfcn.h>

#include <cstdlib>
#include <mutex>

namespace mbfea { namespace nccl {

static Api g_api;
static bool g_ok = false;
static std::string g_msg;
static std::once_flag g_once;

template <class T>
static bool sym(void *h, const char *name, T &out)
{
    out = reinterpret_cast<T>(dlsym(h, name));
    if (!out) { g_msg = std::string("NCCL symbol missing: ") + name; return false; }
    return true;
}

static void do_load()
{
    void *h = nullptr;
    if (const char *p = std::getenv("MBFEA_NCCL_LIB")) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
    // the copy already mapped into the process (e.g. torch's bundled NCCL) wins
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { g_msg = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
    g_ok = sym(h, "ncclGetUniqueId", g_api.GetUniqueId) && sym(h, "ncclCommInitRank", g_api.CommInitRank) &&
           sym(h, "ncclCommDestroy", g_api.CommDestroy) && sym(h, "ncclGetErrorString", g_api.GetErrorString) &&
           sym(h, "ncclAllGather", g_api.AllGather) && sym(h, "ncclAllReduce", g_api.AllReduce) &&
           sym(h, "ncclBroadcast", g_api.Broadcast) && sym(h, "ncclSend", g_api.Send) &&
           sym(h, "ncclRecv", g_api.Recv) && sym(h, "ncclGroupStart", g_api.GroupStart) &&
           sym(h, "ncclGroupEnd", g_api.GroupEnd) && sym(h, "ncclGetVersion", g_api.GetVersion);
}

const Api *load(std::string &msg)
{
    std::call_once(g_once, do_load);
    if (!g_ok) { msg = g_msg; return nullptr; }
    return &g_api;
}

}}  // namespace mbfea::nccl
