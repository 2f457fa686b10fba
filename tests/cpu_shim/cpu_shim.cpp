// cpu_shim.cpp -- TEST INFRASTRUCTURE ONLY.  Builds the product's real kernel + C-ABI sources against
// cuda_emul.h so the GPU-less container can run them thread by thread (tests/test_cpu_shim.py).
// The resulting libnvb_cpu_shim.so exports the same nvb_* symbols as libnvorbis_b200.so; it is only
// ever loaded by tests through an explicit path.
#include "cuda_emul.h"

thread_local cuemu::Idx threadIdx, blockIdx;
cuemu::Idx blockDim, gridDim;
namespace cuemu {
Launch* g_launch = nullptr;
alignas(128) unsigned char g_dyn_smem[232448];

void run(int grid, int block, size_t smem, const std::function<void()>& body) {
    if (smem > sizeof(g_dyn_smem)) std::abort();
    Launch L;
    L.nthreads = block;
    L.cta = std::make_unique<std::barrier<>>(block);
    const int nw = (block + 31) / 32;
    for (int w = 0; w < nw; w++) { int lanes = block - 32 * w; if (lanes > 32) lanes = 32; L.warp.push_back(std::make_unique<std::barrier<>>(lanes)); }
    L.shfl.assign((size_t)nw * 32, 0);
    g_launch = &L;
    blockDim.x = (unsigned)block; gridDim.x = (unsigned)grid;
    std::vector<std::thread> th;
    for (int t = 0; t < block; t++)
        th.emplace_back([&, t]() {
            threadIdx.x = (unsigned)t;
            for (int b = 0; b < grid; b++) {
                blockIdx.x = (unsigned)b;
                body();
                L.cta->arrive_and_wait();           // CTAs run one after the other; static "shared" memory is reused
            }
        });
    for (auto& x : th) x.join();
    g_launch = nullptr;
}
}  // namespace cuemu

#include "../../nvorbis_b200/csrc/nvb_host.cpp"
#include "../../nvorbis_b200/csrc/nvb_kernels.cu"
#include "../../nvorbis_b200/csrc/nvb_fused.cu"
#include "../../nvorbis_b200/csrc/nvb_unpack.cu"
#include "../../nvorbis_b200/csrc/nvb_api.cu"
