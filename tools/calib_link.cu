// calib_link.cu - what the HOST side of this box gives N GPUs at once ("print the host's
// measured limit next to the result").  One thread per GPU, page-locked buffers, plain large copies:
//   * topology: PCI bus id, NUMA node and local CPUs of every GPU; the CPUs / memory nodes this process may use
//   * H2D, D2H and both at once, for every GPU alone, for pairs (0, k) and for the sets {0..1}, {0..3}, {4..7}, {0..7}
//   * the same with the source placed differently: cudaHostAlloc (default), write-combined, and
//     mmap + mbind(the GPU's node) + cudaHostRegister (wicca_host_alloc_near's placement)
//   * the ingest copy as the library issues it (2-D: 24,852-byte rows into a 24,960-byte pitch) next to 1-D
// Build: nvcc -O2 -std=c++17 -o tools/_build/calib_link tools/calib_link.cu -lpthread
// Run:   tools/_build/calib_link [seconds per measurement, default 0.6]
#include <cuda_runtime.h>
#include <ctype.h>
#include <errno.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

static const size_t kH = 6393, kW = 8284, kRowB = kW * 3, kPitch = (kRowB + 127) / 128 * 128;
static const size_t kImage = kRowB * kH;              // 158.9 MB
static const size_t kIcons = kImage / 3;              // what flows back per image at depths 1-6 (1/3 of the input)

enum MemKind { kDefault = 0, kWriteCombined = 1, kBound = 2 };
static const char* kind_name(int k) { return k == 0 ? "cudaHostAlloc" : k == 1 ? "write-combined" : "mmap+mbind+register"; }

static std::string read_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return "?";
    char buf[4096] = {0};
    size_t n = fread(buf, 1, sizeof buf - 1, f);
    fclose(f);
    while (n && (buf[n - 1] == '\n' || buf[n - 1] == ' ')) buf[--n] = 0;
    return buf;
}

static int device_node(int dev) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, dev) != cudaSuccess) return -1;
    for (char* q = bus; *q; ++q) *q = (char)tolower(*q);
    std::string s = read_file(std::string("/sys/bus/pci/devices/") + bus + "/numa_node");
    return s == "?" ? -1 : atoi(s.c_str());
}

static void* alloc_host(size_t bytes, int kind, int node, bool* bound) {
    *bound = false;
    void* p = nullptr;
    if (kind == kDefault) { CK(cudaHostAlloc(&p, bytes, cudaHostAllocPortable)); memset(p, 1, bytes); return p; }
    if (kind == kWriteCombined) { CK(cudaHostAlloc(&p, bytes, cudaHostAllocPortable | cudaHostAllocWriteCombined)); memset(p, 1, bytes); return p; }
    p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) { perror("mmap"); exit(1); }
    if (node >= 0) {
        unsigned long mask[16] = {0};
        mask[node / 64] |= 1ul << (node % 64);
        long rc = syscall(SYS_mbind, p, bytes, 2 /* MPOL_BIND */, mask, sizeof(mask) * 8, 0);
        *bound = rc == 0;
        if (rc != 0) printf("    (mbind to node %d refused: %s)\n", node, strerror(errno));
    }
    madvise(p, bytes, MADV_HUGEPAGE);
    memset(p, 1, bytes);
    CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return p;
}
static void free_host(void* p, size_t bytes, int kind) {
    if (kind == kBound) { cudaHostUnregister(p); munmap(p, bytes); }
    else cudaFreeHost(p);
}

struct Lane {            // one GPU
    int dev = 0, node = -1;
    unsigned char* d_img = nullptr;
    unsigned char* d_icons = nullptr;
    void* h_src[3] = {nullptr, nullptr, nullptr};    // by MemKind
    void* h_dst = nullptr;
    bool bound = false;
    cudaStream_t s_up = nullptr, s_down = nullptr;
    cudaEvent_t e[4] = {};
};

enum Mode { kUp = 1, kDown = 2, kDuplex = 3 };

// every lane of `set` copies for `seconds`; returns per-lane GB/s (up, down)
static void run_set(std::vector<Lane>& lanes, const std::vector<int>& set, int mode, int kind, bool two_d, double seconds,
                    std::vector<double>& up, std::vector<double>& down) {
    up.assign(set.size(), 0); down.assign(set.size(), 0);
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    std::vector<std::thread> th;
    for (size_t k = 0; k < set.size(); ++k)
        th.emplace_back([&, k] {
            Lane& L = lanes[set[k]];
            CK(cudaSetDevice(L.dev));
            const void* src = L.h_src[kind];
            auto copy_up = [&] {
                if (two_d) CK(cudaMemcpy2DAsync(L.d_img, kPitch, src, kRowB, kRowB, kH, cudaMemcpyHostToDevice, L.s_up));
                else CK(cudaMemcpyAsync(L.d_img, src, kImage, cudaMemcpyHostToDevice, L.s_up));
            };
            auto copy_down = [&] { CK(cudaMemcpyAsync(L.h_dst, L.d_icons, kIcons, cudaMemcpyDeviceToHost, L.s_down)); };
            if (mode & kUp) copy_up();
            if (mode & kDown) copy_down();
            CK(cudaDeviceSynchronize());
            ready.fetch_add(1);
            while (!go.load()) std::this_thread::yield();
            auto t0 = std::chrono::steady_clock::now();
            long n_up = 0, n_down = 0;
            if (mode & kUp) CK(cudaEventRecord(L.e[0], L.s_up));
            if (mode & kDown) CK(cudaEventRecord(L.e[2], L.s_down));
            // keep two copies in flight per direction, until the clock runs out
            for (;;) {
                if (mode & kUp) { copy_up(); ++n_up; }
                if (mode & kDown) { copy_down(); ++n_down; if (mode == kDuplex) { /* 1:3 byte ratio is kept by the sizes */ } }
                if (mode & kUp) { CK(cudaEventRecord(L.e[1], L.s_up)); }
                if (mode & kDown) { CK(cudaEventRecord(L.e[3], L.s_down)); }
                if ((n_up + n_down) % 2 == 0) {
                    if (mode & kUp) CK(cudaEventSynchronize(L.e[1]));
                    if (mode & kDown) CK(cudaEventSynchronize(L.e[3]));
                }
                double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                if (el > seconds) break;
            }
            CK(cudaDeviceSynchronize());
            float ms = 0;
            if (mode & kUp) { CK(cudaEventElapsedTime(&ms, L.e[0], L.e[1])); up[k] = n_up * (double)kImage / ms / 1e6; }
            if (mode & kDown) { CK(cudaEventElapsedTime(&ms, L.e[2], L.e[3])); down[k] = n_down * (double)kIcons / ms / 1e6; }
        });
    while (ready.load() < (int)set.size()) std::this_thread::yield();
    go.store(true);
    for (auto& t : th) t.join();
}

static void report(const char* what, const std::vector<int>& set, const std::vector<double>& up, const std::vector<double>& down) {
    double su = 0, sd = 0;
    std::string per;
    char buf[64];
    for (size_t k = 0; k < set.size(); ++k) {
        su += up[k]; sd += down[k];
        snprintf(buf, sizeof buf, " g%d:%.1f/%.1f", set[k], up[k], down[k]);
        per += buf;
    }
    std::string name;
    for (int g : set) { snprintf(buf, sizeof buf, "%d", g); name += buf; }
    printf("%-34s gpus {%s}: H2D %.1f GB/s, D2H %.1f GB/s  (per GPU up/down:%s)\n", what, name.c_str(), su, sd, per.c_str());
    fflush(stdout);
}

int main(int argc, char** argv) {
    const double seconds = argc > 1 ? atof(argv[1]) : 0.6;
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    printf("== topology ==\n");
    printf("Cpus_allowed_list / Mems_allowed_list:\n");
    {
        FILE* f = fopen("/proc/self/status", "r");
        char line[512];
        while (f && fgets(line, sizeof line, f))
            if (!strncmp(line, "Cpus_allowed_list", 17) || !strncmp(line, "Mems_allowed_list", 17)) printf("  %s", line);
        if (f) fclose(f);
    }
    printf("NUMA nodes online: %s; possible: %s\n", read_file("/sys/devices/system/node/online").c_str(),
           read_file("/sys/devices/system/node/possible").c_str());
    std::vector<Lane> lanes(n);
    for (int i = 0; i < n; ++i) {
        Lane& L = lanes[i];
        L.dev = i; L.node = device_node(i);
        char bus[32] = {0};
        cudaDeviceGetPCIBusId(bus, sizeof bus, i);
        for (char* q = bus; *q; ++q) *q = (char)tolower(*q);
        printf("gpu %d: %s numa_node %d local_cpulist %s\n", i, bus, L.node,
               read_file(std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist").c_str());
    }
    fflush(stdout);
    if (system("nvidia-smi topo -m 2>/dev/null | head -24") != 0) printf("(nvidia-smi topo unavailable)\n");
    for (int i = 0; i < n; ++i) {
        Lane& L = lanes[i];
        CK(cudaSetDevice(i));
        CK(cudaMalloc((void**)&L.d_img, kPitch * kH + 256));
        CK(cudaMalloc((void**)&L.d_icons, kIcons));
        for (int k = 0; k < 3; ++k) L.h_src[k] = alloc_host(kImage, k, L.node, &L.bound);
        bool dummy;
        L.h_dst = alloc_host(kIcons, kDefault, L.node, &dummy);
        CK(cudaStreamCreateWithFlags(&L.s_up, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&L.s_down, cudaStreamNonBlocking));
        for (auto& e : L.e) CK(cudaEventCreate(&e));
        printf("gpu %d: buffers ready (mbind to node %d %s)\n", i, L.node, L.bound ? "ok" : "not applied");
    }
    std::vector<double> up, down;
    printf("== every GPU alone (cudaHostAlloc, 1-D) ==\n");
    for (int i = 0; i < n; ++i) {
        run_set(lanes, {i}, kUp, kDefault, false, seconds, up, down); report("H2D only", {i}, up, down);
        run_set(lanes, {i}, kDown, kDefault, false, seconds, up, down); report("D2H only", {i}, up, down);
        run_set(lanes, {i}, kDuplex, kDefault, false, seconds, up, down); report("both directions", {i}, up, down);
    }
    printf("== copy shape and source placement, GPU 0 and the last GPU alone ==\n");
    for (int g : {0, n - 1}) {
        run_set(lanes, {g}, kUp, kDefault, true, seconds, up, down); report("H2D 2-D pitched, cudaHostAlloc", {g}, up, down);
        for (int kind = 1; kind < 3; ++kind) {
            run_set(lanes, {g}, kUp, kind, false, seconds, up, down);
            report((std::string("H2D 1-D, ") + kind_name(kind)).c_str(), {g}, up, down);
        }
        if (n == 1) break;
    }
    if (n > 1) {
        printf("== pairs with GPU 0 (H2D only, cudaHostAlloc) ==\n");
        for (int k = 1; k < n; ++k) { run_set(lanes, {0, k}, kUp, kDefault, false, seconds, up, down); report("H2D only", {0, k}, up, down); }
        printf("== growing sets ==\n");
        std::vector<std::vector<int>> sets;
        for (int m = 2; m <= n; m *= 2) { std::vector<int> s; for (int i = 0; i < m; ++i) s.push_back(i); sets.push_back(s); }
        if (n == 8) sets.push_back({4, 5, 6, 7});
        if (n >= 4) sets.push_back({0, n / 2});
        if (n == 8) sets.push_back({0, 2, 4, 6});
        for (auto& s : sets)
            for (int kind = 0; kind < 3; ++kind) {
                run_set(lanes, s, kUp, kind, false, seconds, up, down); report((std::string("H2D only, ") + kind_name(kind)).c_str(), s, up, down);
                if (kind == 1) continue;
                run_set(lanes, s, kDown, kind, false, seconds, up, down); report((std::string("D2H only, ") + kind_name(kind)).c_str(), s, up, down);
                run_set(lanes, s, kDuplex, kind, false, seconds, up, down); report((std::string("both, ") + kind_name(kind)).c_str(), s, up, down);
            }
    }
    for (int i = 0; i < n; ++i) {
        Lane& L = lanes[i];
        cudaSetDevice(i);
        for (int k = 0; k < 3; ++k) free_host(L.h_src[k], kImage, k);
        free_host(L.h_dst, kIcons, kDefault);
        cudaFree(L.d_img); cudaFree(L.d_icons);
    }
    return 0;
}
