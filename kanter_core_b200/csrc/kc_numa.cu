// kc_numa.cu — NUMA placement of the pinned staging memory and of the threads that fill it.
//
// The reference has no counterpart (its planes are Vec<f32>s wherever malloc puts them).  Here every byte that
// crosses PCIe comes out of, or lands in, page-locked host memory, and on an 8-GPU box the GPUs hang off two
// sockets: eight ranks each pulling ~50 GB/s out of buffers that all sit in ONE socket's DRAM is what flattened
// round 1's end-to-end curve (170 GB/s aggregate).  So pinned buffers are placed on the NUMA node the GPU is
// attached to: mmap + mbind(MPOL_PREFERRED, node) + cudaHostRegister, and a rank may move its host thread there too.
// Everything is best effort: without NUMA information (or permission) the plain cudaHostAlloc path is taken.
#include <dirent.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cctype>
#include <cstring>
#include <fstream>
#include <mutex>
#include <string>
#include <unordered_map>

#include "kc_internal.h"

namespace {

std::mutex g_mu;
std::unordered_map<void*, size_t> g_mapped;   // buffers made by mmap + cudaHostRegister: base -> length

int read_int_file(const std::string& path, int fallback) {
    std::ifstream f(path);
    int v = fallback;
    if (f && (f >> v)) return v;
    return fallback;
}

// "0-3,8,10-11" -> cpu ids
std::vector<int> parse_cpu_list(const std::string& s) {
    std::vector<int> out;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && !isdigit((unsigned char)s[i])) ++i;
        if (i >= s.size()) break;
        int a = 0;
        while (i < s.size() && isdigit((unsigned char)s[i])) a = a * 10 + (s[i++] - '0');
        int b = a;
        if (i < s.size() && s[i] == '-') {
            ++i;
            b = 0;
            while (i < s.size() && isdigit((unsigned char)s[i])) b = b * 10 + (s[i++] - '0');
        }
        for (int c = a; c <= b && c < 4096; ++c) out.push_back(c);
    }
    return out;
}

int online_nodes() {
    int n = 0;
    if (DIR* d = opendir("/sys/devices/system/node")) {
        while (dirent* e = readdir(d))
            if (strncmp(e->d_name, "node", 4) == 0 && isdigit((unsigned char)e->d_name[4])) ++n;
        closedir(d);
    }
    return n;
}

}  // namespace

// NUMA node of a CUDA device from its PCI address (-1: unknown, or a single-node machine says -1 itself)
int kc_device_numa_node(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    std::string id(bus);
    for (char& c : id) c = (char)tolower((unsigned char)c);
    return read_int_file("/sys/bus/pci/devices/" + id + "/numa_node", -1);
}

int32_t kc_host_alloc_on_node(int node, size_t bytes, void** out) {
    if (!out) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "out is NULL");
    if (bytes == 0) bytes = 1;
    static const bool off = getenv("KC_NO_NUMA") != nullptr;
    if (node >= 0 && !off && online_nodes() > 1) {
        const size_t page = 2u << 20;                       // whole 2 MiB pages: transparent huge pages where the host offers them
        const size_t len = (bytes + page - 1) / page * page;
        void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p != MAP_FAILED) {
            unsigned long mask[16] = {0};
            if (node < (int)(sizeof mask * 8)) mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
            // MPOL_PREFERRED (1): pages come from `node` while it has any, never a failure; a cpuset that forbids the node
            // makes the call fail, and the buffer simply keeps the default policy
            (void)syscall(SYS_mbind, p, len, 1 /* MPOL_PREFERRED */, mask, sizeof mask * 8, 0u);
            (void)madvise(p, len, MADV_HUGEPAGE);
            if (cudaHostRegister(p, len, cudaHostRegisterPortable) == cudaSuccess) {   // pins, and thereby faults the pages in under the policy
                std::lock_guard<std::mutex> lk(g_mu);
                g_mapped[p] = len;
                *out = p;
                return KC_OK;
            }
            cudaGetLastError();
            munmap(p, len);
        }
    }
    KC_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return KC_OK;
}

int32_t kc_host_release(void* p) {
    if (!p) return KC_OK;
    size_t len = 0;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_mapped.find(p);
        if (it != g_mapped.end()) {
            len = it->second;
            g_mapped.erase(it);
        }
    }
    if (len) {
        cudaError_t e = cudaHostUnregister(p);
        munmap(p, len);
        if (e != cudaSuccess) { cudaGetLastError(); }
        return KC_OK;
    }
    KC_CUDA(cudaFreeHost(p));
    return KC_OK;
}

extern "C" int32_t kc_host_alloc_near_device(int32_t device, size_t bytes, void** out) try {
    // page-locked host memory on the NUMA node the device is attached to (plain cudaHostAlloc where that is unknown)
    return kc_host_alloc_on_node(kc_device_numa_node(device), bytes, out);
} KC_ABI_CATCH

extern "C" int32_t kc_numa_info(int32_t device, int32_t* device_node, int32_t* thread_node, int32_t* nodes) try {
    // where the device hangs, where the calling thread runs right now, how many nodes the host has (-1: unknown)
    if (device_node) *device_node = kc_device_numa_node(device);
    if (thread_node) {
        unsigned cpu = 0, node = 0;
        *thread_node = syscall(SYS_getcpu, &cpu, &node, nullptr) == 0 ? (int32_t)node : -1;
    }
    if (nodes) *nodes = online_nodes();
    return KC_OK;
} KC_ABI_CATCH

extern "C" int32_t kc_bind_thread_near_device(int32_t device, int32_t* bound) try {
    // Move the CALLING thread onto the CPUs of the device's NUMA node -- those of them the process is allowed to use.
    // *bound: 1 moved, 0 left alone (node unknown, or none of its CPUs is in the process's affinity mask).
    if (bound) *bound = 0;
    const int node = kc_device_numa_node(device);
    if (node < 0) return KC_OK;
    std::ifstream f("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
    std::string list;
    if (!f || !std::getline(f, list)) return KC_OK;
    cpu_set_t allowed, want;
    CPU_ZERO(&allowed);
    CPU_ZERO(&want);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return KC_OK;
    int n = 0;
    for (int c : parse_cpu_list(list))
        if (c < CPU_SETSIZE && CPU_ISSET(c, &allowed)) { CPU_SET(c, &want); ++n; }
    if (n == 0) return KC_OK;
    if (sched_setaffinity(0, sizeof want, &want) == 0 && bound) *bound = 1;
    return KC_OK;
} KC_ABI_CATCH
