// Host <-> device copy ceiling of the box: plain cudaMemcpyAsync between page-locked host memory and HBM, one copy per call,
// 1 .. N GPUs concurrently (one host thread + one stream + one pinned buffer per GPU, common start).  This is the ceiling of
// the end-to-end (`e2e`) figure of bench.py, whose step ends with a device-to-host copy of 16 808 B per spectrum.
//   nvcc -O3 -o tools/micro/bin/d2h_probe tools/d2h_probe.cu -lpthread ;  tools/micro/bin/d2h_probe [max_gpus] [MiB per copy] [copies]
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Lane {
    int dev;
    void* host = nullptr;
    void* devp = nullptr;
    cudaStream_t st;
    double gbs_d2h = 0, gbs_h2d = 0;
};

static std::atomic<int> g_arrived{0};
static std::atomic<int> g_phase{0};

static void barrier(int n, int phase) {          // all lanes start a measurement together
    g_arrived.fetch_add(1);
    if (g_arrived.load() == n * (phase + 1)) g_phase.store(phase + 1);
    while (g_phase.load() <= phase) std::this_thread::yield();
}

static void worker(Lane* l, int n, size_t bytes, int copies) {
    CK(cudaSetDevice(l->dev));
    for (int dir = 0; dir < 2; ++dir) {
        // warm-up copy, then `copies` timed copies back to back on the lane's stream
        if (dir == 0) CK(cudaMemcpyAsync(l->host, l->devp, bytes, cudaMemcpyDeviceToHost, l->st));
        else CK(cudaMemcpyAsync(l->devp, l->host, bytes, cudaMemcpyHostToDevice, l->st));
        CK(cudaStreamSynchronize(l->st));
        barrier(n, dir);
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < copies; ++i) {
            if (dir == 0) CK(cudaMemcpyAsync(l->host, l->devp, bytes, cudaMemcpyDeviceToHost, l->st));
            else CK(cudaMemcpyAsync(l->devp, l->host, bytes, cudaMemcpyHostToDevice, l->st));
        }
        CK(cudaStreamSynchronize(l->st));
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        (dir == 0 ? l->gbs_d2h : l->gbs_h2d) = (double)bytes * copies / s / 1e9;
    }
}

int main(int argc, char** argv) {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    const int maxg = argc > 1 ? std::min(atoi(argv[1]), ndev) : ndev;
    const size_t bytes = (size_t)(argc > 2 ? atoi(argv[2]) : 1024) << 20;
    const int copies = argc > 3 ? atoi(argv[3]) : 8;
    printf("# cudaMemcpyAsync, pinned host memory, %zu MiB per copy, %d copies per lane; GB/s = 1e9 B/s\n", bytes >> 20, copies);
    for (int n = 1; n <= maxg; n *= 2) {
        std::vector<Lane> lanes(n);
        for (int i = 0; i < n; ++i) {
            lanes[i].dev = i;
            CK(cudaSetDevice(i));
            CK(cudaHostAlloc(&lanes[i].host, bytes, cudaHostAllocPortable));
            CK(cudaMalloc(&lanes[i].devp, bytes));
            CK(cudaMemset(lanes[i].devp, 1, bytes));
            CK(cudaStreamCreateWithFlags(&lanes[i].st, cudaStreamNonBlocking));
            CK(cudaDeviceSynchronize());
        }
        g_arrived = 0; g_phase = 0;
        std::vector<std::thread> th;
        for (int i = 0; i < n; ++i) th.emplace_back(worker, &lanes[i], n, bytes, copies);
        for (auto& t : th) t.join();
        double sd = 0, sh = 0, mind = 1e30;
        for (auto& l : lanes) { sd += l.gbs_d2h; sh += l.gbs_h2d; mind = std::min(mind, l.gbs_d2h); }
        printf("%d GPU(s) concurrently: D2H %.1f GB/s aggregate (%.1f per GPU, slowest %.1f)   H2D %.1f GB/s aggregate (%.1f per GPU)\n",
               n, sd, sd / n, mind, sh, sh / n);
        for (auto& l : lanes) {
            CK(cudaSetDevice(l.dev));
            CK(cudaFree(l.devp)); CK(cudaFreeHost(l.host)); CK(cudaStreamDestroy(l.st));
        }
    }
    return 0;
}
