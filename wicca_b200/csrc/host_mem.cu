// host_mem.cu - page-locked host memory for the ingest path (wicca_host_alloc*, wicca_host_register).
//
// The end-to-end rate of the path is set by the host side of the PCIe links (159 MB per image in, 53 MB of
// icons out), so where the pages live matters: wicca_host_alloc_near places them on the NUMA node the GPU
// hangs off.  When the process may run on that node's CPUs, first touch under a scoped affinity does it; when the
// cpuset excludes them (a VM that only exposes one socket's cores), the pages are mmap'ed, bound with mbind()
// and then registered with CUDA - memory policy does not need a CPU on the node.  Everything falls back to
// plain cudaHostAlloc when the node is unknown or the policy call is refused.
#include <errno.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <ctype.h>
#include <mutex>
#include <unordered_map>

#include "host_common.h"

using namespace wicca;

namespace {

struct HostBlock { size_t bytes; int kind; };      // kind 0: cudaHostAlloc, 1: mmap + register, 2: caller's memory, registered
std::mutex g_host_mu;
std::unordered_map<void*, HostBlock> g_host_blocks;

int device_numa_node(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char* q = bus; *q; ++q) *q = (char)tolower(*q);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

int online_node_count() {
    FILE* f = fopen("/sys/devices/system/node/online", "r");
    if (!f) return 1;
    char text[256] = {0};
    const size_t got = fread(text, 1, sizeof text - 1, f);
    fclose(f);
    if (!got) return 1;
    int n = 0;
    for (const char* p = text; *p;) {
        while (*p == ',' || *p == ' ' || *p == '\n') ++p;
        if (!*p) break;
        char* end = nullptr;
        long a = strtol(p, &end, 10);
        if (end == p) break;
        long b = a;
        p = end;
        if (*p == '-') { b = strtol(p + 1, &end, 10); p = end; }
        n += (int)(b - a + 1);
    }
    return n > 0 ? n : 1;
}

// mmap + mbind(node) + touch + cudaHostRegister; nullptr when any step is refused (caller falls back)
void* alloc_bound(size_t bytes, int node) {
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t len = (bytes + page - 1) / page * page;
    void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
    unsigned long mask[16] = {0};
    mask[node / 64] |= 1ul << (node % 64);
    if (node >= 1024 || syscall(SYS_mbind, p, len, 2 /* MPOL_BIND */, mask, sizeof(mask) * 8, 0) != 0) { munmap(p, len); return nullptr; }
    madvise(p, len, MADV_HUGEPAGE);
    memset(p, 0, len);                                   // fault the pages in under the policy
    if (cudaHostRegister(p, len, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); munmap(p, len); return nullptr; }
    std::lock_guard<std::mutex> lk(g_host_mu);
    g_host_blocks[p] = {len, 1};
    return p;
}

}  // namespace

extern "C" int wicca_host_alloc_near(void** ptr, size_t bytes, int device) {
    if (!ptr) return fail(WICCA_EINVAL, "ptr is NULL");
    *ptr = nullptr;
    int rc = check_device(device < 0 ? 0 : device);
    if (rc) return rc;
    if (bytes == 0) bytes = 1;
    const char* mode = getenv("WICCA_HOST_ALLOC");       // "cuda": never bind; "bind": bind whenever the node is known
    const bool never = mode && !strcmp(mode, "cuda");
    if (device >= 0 && !never) {
        const int node = device_numa_node(device);
        if (node >= 0 && (online_node_count() > 1 || (mode && !strcmp(mode, "bind")))) {
            WICCA_CUDA(cudaSetDevice(device));
            if (void* p = alloc_bound(bytes, node)) { *ptr = p; return 0; }
        }
    }
    // first touch by this thread while it sits on the GPU's own CPUs (no-op when they are outside the cpuset)
    ScopedAffinity bind(device);
    WICCA_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
    std::lock_guard<std::mutex> lk(g_host_mu);
    g_host_blocks[*ptr] = {bytes, 0};
    return 0;
}

extern "C" int wicca_host_alloc(void** ptr, size_t bytes) { return wicca_host_alloc_near(ptr, bytes, -1); }

extern "C" int wicca_host_register(void* ptr, size_t bytes) {
    if (!ptr || bytes == 0) return fail(WICCA_EINVAL, "null pointer / zero size");
    int rc = check_device(0);
    if (rc) return rc;
    WICCA_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    std::lock_guard<std::mutex> lk(g_host_mu);
    g_host_blocks[ptr] = {bytes, 2};
    return 0;
}

extern "C" int wicca_host_unregister(void* ptr) {
    if (!ptr) return 0;
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        auto it = g_host_blocks.find(ptr);
        if (it == g_host_blocks.end() || it->second.kind != 2) return fail(WICCA_ESTATE, "pointer was not registered with wicca_host_register");
        g_host_blocks.erase(it);
    }
    WICCA_CUDA(cudaHostUnregister(ptr));
    return 0;
}

extern "C" int wicca_host_free(void* ptr) {
    if (!ptr) return 0;
    HostBlock blk{0, 0};
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        auto it = g_host_blocks.find(ptr);
        if (it != g_host_blocks.end()) { blk = it->second; g_host_blocks.erase(it); }
    }
    if (blk.kind == 2) return fail(WICCA_ESTATE, "registered caller memory: use wicca_host_unregister");
    if (blk.kind == 1) {
        cudaError_t e = cudaHostUnregister(ptr);
        munmap(ptr, blk.bytes);
        if (e != cudaSuccess) return cuda_fail(e, "cudaHostUnregister");
        return 0;
    }
    WICCA_CUDA(cudaFreeHost(ptr));
    return 0;
}
