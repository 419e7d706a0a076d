// d2h_bw.cu -- what the host side of the box gives the end-to-end path: device-to-host copy bandwidth into pinned
// memory from 1..N GPUs at once (one thread + one stream per GPU, 415 MB per copy = one step's RGBA of 1024 CIF
// pictures), for default pinned memory, write-combined pinned memory, and buffers first touched on the GPU's own NUMA
// node (sched_setaffinity to the node's CPUs before cudaHostAlloc, as bench.py's shard.bind_rank_to_gpu_node does).
//   nvcc -O2 -o tools/d2h_bw.bin tools/d2h_bw.cu -lpthread && tools/d2h_bw.bin
#include <cuda_runtime.h>
#include <sched.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static std::vector<int> node_cpus(int node) {
    std::vector<int> out;
    char path[128];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = fopen(path, "r");
    if (!f) return out;
    char buf[4096] = {0};
    if (!fgets(buf, sizeof buf, f)) buf[0] = 0;
    fclose(f);
    for (char* tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a, b;
        if (sscanf(tok, "%d-%d", &a, &b) == 2)
            for (int c = a; c <= b; c++) out.push_back(c);
        else if (sscanf(tok, "%d", &a) == 1)
            out.push_back(a);
    }
    return out;
}

static int gpu_node(int dev) {
    char bus[64];
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, dev) != cudaSuccess) return -1;
    for (char* p = bus; *p; p++) *p = (char)tolower(*p);
    std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return -1;
    int n = -1;
    if (fscanf(f, "%d", &n) != 1) n = -1;
    fclose(f);
    return n;
}

int main() {
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    const size_t bytes = (size_t)1024 * 352 * 288 * 4;
    printf("devices %d, host cpus %ld, copy size %.1f MB\n", ndev, sysconf(_SC_NPROCESSORS_ONLN), bytes / 1e6);
    for (int d = 0; d < ndev; d++) printf("  gpu %d numa node %d\n", d, gpu_node(d));
    const char* modes[3] = {"pinned default", "pinned write-combined", "pinned, allocated on the GPU's NUMA node"};
    for (int mode = 0; mode < 3; mode++) {
        for (int n = 1; n <= ndev; n *= 2) {
            std::vector<double> gbs(n, 0.0);
            std::atomic<int> ready{0};
            std::atomic<bool> go{false};
            std::vector<std::thread> th;
            for (int d = 0; d < n; d++)
                th.emplace_back([&, d] {
                    cudaSetDevice(d);
                    if (mode == 2) {
                        std::vector<int> cpus = node_cpus(gpu_node(d));
                        if (!cpus.empty()) {
                            cpu_set_t set;
                            CPU_ZERO(&set);
                            for (int c : cpus) CPU_SET(c, &set);
                            sched_setaffinity(0, sizeof set, &set);
                        }
                    }
                    void *dv = nullptr, *h = nullptr;
                    cudaMalloc(&dv, bytes);
                    cudaMemset(dv, 1, bytes);
                    cudaHostAlloc(&h, bytes, mode == 1 ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
                    if (mode != 1) memset(h, 0, bytes);  // first touch on this thread's node
                    cudaStream_t s;
                    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
                    cudaMemcpyAsync(h, dv, bytes, cudaMemcpyDeviceToHost, s);
                    cudaStreamSynchronize(s);
                    ready++;
                    while (!go.load()) std::this_thread::yield();
                    const int reps = 10;
                    auto t0 = std::chrono::steady_clock::now();
                    for (int r = 0; r < reps; r++) cudaMemcpyAsync(h, dv, bytes, cudaMemcpyDeviceToHost, s);
                    cudaStreamSynchronize(s);
                    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                    gbs[d] = reps * bytes / dt / 1e9;
                    cudaFreeHost(h);
                    cudaFree(dv);
                    cudaStreamDestroy(s);
                });
            while (ready.load() < n) std::this_thread::yield();
            go = true;
            for (auto& t : th) t.join();
            double sum = 0;
            for (double g : gbs) sum += g;
            printf("%-42s %d GPU(s): aggregate %.1f GB/s = %.1f GP/s of RGBA (per GPU:", modes[mode], n, sum, sum / 4);
            for (double g : gbs) printf(" %.1f", g);
            printf(")\n");
        }
    }
    return 0;
}
