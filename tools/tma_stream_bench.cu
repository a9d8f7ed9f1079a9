// tma_stream_bench.cu — stand-alone microbenchmark (NOT part of libcgvec_b200.so, not used by tests or bench.py).
//
// Question it answers (NOTES_ROUND2.md): the tensor-core scan streams 128-byte row segments with 2-D TMA boxes and tops
// out near 6.4 TB/s with the MMAs and the epilogue switched off, while the exact scan's full-row bulk copies reach
// 7.2 TB/s.  Is that the shape of the boxes?  This program streams an n x d fp16 matrix through a shared-memory ring
// with NO compute, once per load shape, and prints GB/s for each:
//   bulk   contiguous full rows with cp.async.bulk (what scan_exact_kernel does)
//   t2     2-D map, box {64 halves, R rows}, SWIZZLE_128B (what tc_scan_kernel does; one K-block per instruction)
//   t3     3-D map {64, K-blocks, rows}, box {64, KB, R}: one instruction brings KB K-blocks (layout [row][kb][128 B])
//   t4     4-D map {64, 8 rows, K-blocks, rows/8}, box {64, 8, KB, R/8}: the same bytes laid out as KB consecutive
//          K-major SWIZZLE_128B atoms per 8-row group, i.e. directly usable by tcgen05.mma with SBO = KB * 1024
//   t2w    2-D map without swizzle, box {256 halves, R rows} (512 contiguous bytes per row)
//
// Second part ("pipe" lines): the same ring feeding the tensor core the way tc_scan_kernel does, switched on step by step, to
// see which step costs bandwidth:   c0 consumer just frees the slot   c1 slot freed by tcgen05.commit (no MMA)
//   c2 four tcgen05.mma per K-block into TMEM + commit (resident query block of N columns)   c4 = c2 plus four warps that
//   read every accumulator tile back with tcgen05.ld (two TMEM buffers, full/empty handshake) and fold it into a checksum.
// Rows and queries hold small integers, so the checksum is exact: it must be equal for the t2 and t4 layouts (t4 feeds the
// MMA with SBO = KB * 1024) and, for <= 4096 rows, equal to the host's.
//
// Build and run (one GPU):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/tma_stream tools/tma_stream_bench.cu
//   gpurun_out/tma_stream [rows=2097152] [dim=1024] > gpurun_out/tma_stream.txt
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../codegraph-rust_b200/csrc/scan_tc.cuh"

using namespace cgv;

enum Mode : uint32_t { BULK = 0, T2 = 1, T3 = 2, T4 = 3, T2W = 4 };

struct StreamParams {
    const uint8_t* base;
    uint64_t n_tiles;        // row tiles of rows_per_load rows
    uint32_t row_bytes;
    uint32_t mode;
    uint32_t rows_per_load;  // R
    uint32_t kb_per_load;    // KB (K-blocks of 128 bytes brought by one instruction)
    uint32_t groups;         // loads per row tile = K-blocks / KB  (1 for bulk)
    uint32_t stage_bytes;
    uint32_t stages;
    uint32_t chunked;        // 0: CTA b takes tiles b, b+grid, ...   1: CTA b takes one contiguous run of tiles
};

__device__ __forceinline__ void tma2(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma4(void* dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

constexpr uint32_t kMaxStages = 16;

__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap map, StreamParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full[kMaxStages];
    __shared__ __align__(8) uint64_t empty[kMaxStages];
    uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // swizzled boxes need 1024-byte alignment

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        fence_mbar_init();
    }
    __syncthreads();

    uint64_t t_begin, t_end, t_step;
    if (p.chunked) {
        const uint64_t per = (p.n_tiles + gridDim.x - 1) / gridDim.x;
        t_begin = per * blockIdx.x;
        t_end = t_begin + per < p.n_tiles ? t_begin + per : p.n_tiles;
        t_step = 1;
    } else {
        t_begin = blockIdx.x; t_end = p.n_tiles; t_step = gridDim.x;
    }

    if (threadIdx.x == 0) {                       // producer
        uint32_t it = 0;
        for (uint64_t t = t_begin; t < t_end; t += t_step) {
            const uint64_t row0 = t * p.rows_per_load;
            for (uint32_t g = 0; g < p.groups; ++g, ++it) {
                const uint32_t s = it % p.stages, ph = (it / p.stages) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full[s], p.stage_bytes);
                uint8_t* dst = ring + (size_t)s * p.stage_bytes;
                const int kb0 = (int)(g * p.kb_per_load);
                switch (p.mode) {
                    case BULK: bulk_g2s(dst, p.base + row0 * p.row_bytes, p.stage_bytes, &full[s]); break;
                    case T2:   tma2(dst, &map, kb0 * 64, (int)row0, &full[s]); break;
                    case T2W:  tma2(dst, &map, kb0 * 64, (int)row0, &full[s]); break;
                    case T3:   tma3(dst, &map, 0, kb0, (int)row0, &full[s]); break;
                    default:   tma4(dst, &map, 0, 0, kb0, (int)(row0 / 8), &full[s]); break;
                }
            }
        }
    } else if (threadIdx.x == 32) {               // consumer: data landed -> slot free again (nothing is read)
        uint32_t it = 0;
        for (uint64_t t = t_begin; t < t_end; t += t_step)
            for (uint32_t g = 0; g < p.groups; ++g, ++it) {
                const uint32_t s = it % p.stages, ph = (it / p.stages) & 1u;
                mbar_wait(&full[s], ph);
                mbar_arrive(&empty[s]);
            }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// pipe_kernel: ring -> tcgen05.mma -> (optional) TMEM read-back.  Roles as in tc_scan_kernel: warps 0-3 read TMEM (one
// lane quarter each), warp 4 lane 0 = TMA producer, warp 5 lane 0 = MMA issuer, warp 6 allocates TMEM.
// ---------------------------------------------------------------------------------------------------------------------
struct PipeParams {
    uint64_t n_tiles;        // 128-row tiles
    uint32_t mode;           // T2 or T4
    uint32_t kb_per_load;    // 1 for T2
    uint32_t groups;         // loads per tile
    uint32_t stage_bytes, stages;
    uint32_t N, nkb, tmem_cols;
    uint32_t consume;        // 0, 1, 2, 4 (see the header comment)
    unsigned long long* checksum;
};

__device__ __forceinline__ uint64_t umma_desc_sw128_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

constexpr int kPipeThreads = 224;

__global__ void __launch_bounds__(kPipeThreads, 1) pipe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, PipeParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full[kMaxStages];
    __shared__ __align__(8) uint64_t empty[kMaxStages];
    __shared__ __align__(8) uint64_t b_bar, done_bar, tfull[2], tempty[2];
    __shared__ uint32_t s_tmem;
    uint8_t* sB = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ring = sB + (size_t)p.nkb * p.N * 128;                    // multiple of 1024 (N % 8 == 0)

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&b_bar, 1); mbar_init(&done_bar, 1);
        mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
        mbar_init(&tempty[0], 4); mbar_init(&tempty[1], 4);
        fence_mbar_init();
    }
    if (warp == 6) tmem_alloc(&s_tmem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const uint64_t my_tiles = p.n_tiles > blockIdx.x ? (p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const bool readback = p.consume == 4;

    if (warp == 4) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&b_bar, p.nkb * p.N * 128);
            for (uint32_t kb = 0; kb < p.nkb; ++kb) tma2(sB + (size_t)kb * p.N * 128, &tmB, kb * 64, 0, &b_bar);
            uint32_t it = 0;
            for (uint64_t t = 0; t < my_tiles; ++t) {
                const uint64_t row0 = (blockIdx.x + t * gridDim.x) * 128;
                for (uint32_t g = 0; g < p.groups; ++g, ++it) {
                    const uint32_t s = it % p.stages, ph = (it / p.stages) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    mbar_arrive_expect_tx(&full[s], p.stage_bytes);
                    uint8_t* dst = ring + (size_t)s * p.stage_bytes;
                    if (p.mode == T2) tma2(dst, &tmA, (int)g * 64, (int)row0, &full[s]);
                    else tma4(dst, &tmA, 0, 0, (int)(g * p.kb_per_load), (int)(row0 / 8), &full[s]);
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, p.N);
            const uint32_t sbo = p.mode == T2 ? 1024u : p.kb_per_load * 1024u;
            mbar_wait(&b_bar, 0);
            tc_fence_after();
            uint32_t it = 0;
            for (uint64_t t = 0; t < my_tiles; ++t) {
                const uint32_t buf = readback ? (uint32_t)(t & 1) : 0u;
                if (readback) { mbar_wait(&tempty[buf], ((t >> 1) & 1) ^ 1); tc_fence_after(); }
                const uint32_t d_tmem = tmem_base + buf * p.N;
                for (uint32_t g = 0; g < p.groups; ++g, ++it) {
                    const uint32_t s = it % p.stages;
                    mbar_wait(&full[s], (it / p.stages) & 1);
                    tc_fence_after();
                    if (p.consume == 0) { mbar_arrive(&empty[s]); continue; }
                    if (p.consume >= 2) {
                        const uint32_t a_stage = smem_u32(ring + (size_t)s * p.stage_bytes);
                        for (uint32_t kbi = 0; kbi < p.kb_per_load; ++kbi) {
                            const uint32_t kb = g * p.kb_per_load + kbi;
                            const uint32_t a_addr = a_stage + kbi * 1024u;
                            const uint32_t b_addr = smem_u32(sB + (size_t)kb * p.N * 128);
#pragma unroll
                            for (uint32_t k = 0; k < 4; ++k)
                                umma_ss<false, false>(d_tmem, umma_desc_sw128_sbo(a_addr + k * 32, sbo), umma_desc_sw128(b_addr + k * 32), idesc, (kb | k) != 0);
                        }
                    }
                    umma_commit(&empty[s]);
                }
                if (readback) umma_commit(&tfull[buf]);
            }
            if (p.consume != 0) { umma_commit(&done_bar); mbar_wait(&done_bar, 0); }
        }
    } else if (warp < 4 && readback) {
        long long sum = 0;
        for (uint64_t t = 0; t < my_tiles; ++t) {
            const uint32_t buf = t & 1;
            mbar_wait(&tfull[buf], (t >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + buf * p.N + ((warp * 32u) << 16);
            for (uint32_t c = 0; c < p.N; c += 16) {
                uint32_t r[16];
                tmem_ld_x16(taddr + c, r);
                tmem_ld_wait(r);
#pragma unroll
                for (int i = 0; i < 16; ++i) sum += (long long)__uint_as_float(r[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
        }
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) atomicAdd(p.checksum, (unsigned long long)sum);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 6) tmem_dealloc(tmem_base, p.tmem_cols);
}

__global__ void fill_pattern_kernel(__half* a, uint64_t rows, uint32_t dim, int mul_r, int mul_c, int mod, int off) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * dim) return;
    const uint64_t r = i / dim, c = i % dim;
    a[i] = __float2half((float)((int)((r * mul_r + c * mul_c) % mod) - off));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Config {
    const char* name;
    uint32_t mode, R, KB, stages, l2promo, chunked;
};

int main(int argc, char** argv) {
    const uint64_t rows = argc > 1 ? strtoull(argv[1], nullptr, 10) : (1ull << 21);
    const uint32_t dim = argc > 2 ? (uint32_t)atoi(argv[2]) : 1024;
    if (rows % 256 || dim % 64) { fprintf(stderr, "rows %% 256 == 0 and dim %% 64 == 0 required\n"); return 1; }
    const uint32_t row_bytes = dim * 2, nkb = dim / 64;
    const size_t bytes = (size_t)rows * row_bytes;

    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("# %s, %d SMs, matrix %llu x %u f16 = %.2f GB\n", prop.name, sms, (unsigned long long)rows, dim, bytes / 1e9);

    void* p_fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p_fn, cudaEnableDefault, &qres));
    EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(p_fn);
    if (!encode) { fprintf(stderr, "cuTensorMapEncodeTiled unavailable\n"); return 1; }

    uint8_t* d = nullptr;
    CK(cudaMalloc(&d, bytes));
    CK(cudaMemset(d, 0x3c, bytes));
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));

    std::vector<Config> cfgs;
    auto add = [&](const char* name, uint32_t mode, uint32_t R, uint32_t KB, uint32_t inflight_kb, uint32_t l2 = 1, uint32_t chunked = 0) {
        if (mode != BULK && nkb % KB) return;
        const uint32_t sb = mode == BULK ? R * row_bytes : R * KB * 128;
        uint32_t st = inflight_kb * 1024 / sb;
        if (st < 2) return;
        if (st > kMaxStages) st = kMaxStages;
        cfgs.push_back({name, mode, R, KB, st, l2, chunked});
    };
    for (uint32_t infl : {48u, 64u, 96u, 128u, 192u}) {
        add("bulk  full rows", BULK, 16, nkb, infl);
        add("bulk  full rows", BULK, 8, nkb, infl);
        add("t2    box 64x128", T2, 128, 1, infl);
        add("t2    box 64x256", T2, 256, 1, infl);
        add("t2    box 64x64", T2, 64, 1, infl);
        add("t3    box 64xKBx128", T3, 128, 2, infl);
        add("t3    box 64xKBx128", T3, 128, 4, infl);
        add("t3    box 64xKBx64", T3, 64, 4, infl);
        add("t3    box 64xKBx32", T3, 32, 8, infl);
        add("t3    full rows", T3, 16, nkb, infl);
        add("t4    box 64x8xKBx16", T4, 128, 2, infl);
        add("t4    box 64x8xKBx16", T4, 128, 4, infl);
        add("t4    box 64x8xKBx8", T4, 64, 4, infl);
        add("t4    box 64x8xKBx4", T4, 32, 8, infl);
        add("t4    full rows", T4, 16, nkb, infl);
        add("t2w   box 256x64 noswz", T2W, 64, 4, infl);
        add("t2w   box 256x32 noswz", T2W, 32, 4, infl);
    }
    add("t2    box 64x128 L2promo none", T2, 128, 1, 192, 0);
    add("t2    box 64x128 L2promo 256B", T2, 128, 1, 192, 2);
    add("t2    box 64x128 chunked tiles", T2, 128, 1, 192, 1, 1);
    add("t4    box 64x8xKBx16 chunked tiles", T4, 128, 4, 192, 1, 1);
    add("bulk  full rows chunked tiles", BULK, 16, nkb, 192, 1, 1);

    printf("%-36s %4s %3s %6s %8s %9s %9s %9s\n", "# shape", "R", "KB", "stages", "stage_KB", "best_ms", "med_ms", "best_GB/s");
    for (const Config& c : cfgs) {
        CUtensorMap map;
        memset(&map, 0, sizeof(map));
        CUresult r = CUDA_SUCCESS;
        const CUtensorMapL2promotion l2 = c.l2promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : c.l2promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
        if (c.mode == T2 || c.mode == T2W) {
            cuuint64_t gd[2] = {dim, rows};
            cuuint64_t gs[1] = {row_bytes};
            cuuint32_t box[2] = {c.mode == T2 ? 64u : 64u * c.KB, c.R};
            cuuint32_t es[2] = {1, 1};
            r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       c.mode == T2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else if (c.mode == T3) {
            cuuint64_t gd[3] = {64, nkb, rows};
            cuuint64_t gs[2] = {128, row_bytes};
            cuuint32_t box[3] = {64, c.KB, c.R};
            cuuint32_t es[3] = {1, 1, 1};
            r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else if (c.mode == T4) {
            cuuint64_t gd[4] = {64, 8, nkb, rows / 8};
            cuuint64_t gs[3] = {row_bytes, 128, 8ull * row_bytes};
            cuuint32_t box[4] = {64, 8, c.KB, c.R / 8};
            cuuint32_t es[4] = {1, 1, 1, 1};
            r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) { printf("%-36s %4u %3u  tensor map rejected (CUresult %d)\n", c.name, c.R, c.KB, (int)r); continue; }

        StreamParams p{};
        p.base = d;
        p.n_tiles = rows / c.R;
        p.row_bytes = row_bytes;
        p.mode = c.mode;
        p.rows_per_load = c.R;
        p.kb_per_load = c.KB;
        p.groups = c.mode == BULK ? 1 : nkb / c.KB;
        p.stage_bytes = c.mode == BULK ? c.R * row_bytes : c.R * c.KB * 128;
        p.stages = c.stages;
        p.chunked = c.chunked;
        const size_t smem = (size_t)p.stage_bytes * p.stages + 1024;
        std::vector<float> ms;
        bool ok = true;
        for (int rep = 0; rep < 7 && ok; ++rep) {
            CK(cudaEventRecord(e0));
            stream_kernel<<<sms, 64, smem>>>(map, p);
            CK(cudaEventRecord(e1));
            cudaError_t e = cudaEventSynchronize(e1);
            if (e == cudaSuccess) e = cudaGetLastError();
            if (e != cudaSuccess) { printf("%-36s %4u %3u  kernel failed: %s\n", c.name, c.R, c.KB, cudaGetErrorString(e)); ok = false; break; }
            float t;
            CK(cudaEventElapsedTime(&t, e0, e1));
            if (rep >= 2) ms.push_back(t);
        }
        if (!ok) return 2;                        // a sticky error poisons the context: stop here
        std::sort(ms.begin(), ms.end());
        printf("%-36s %4u %3u %6u %8.1f %9.4f %9.4f %9.1f\n", c.name, c.R, c.KB, c.stages, p.stage_bytes / 1024.0, ms.front(), ms[ms.size() / 2],
               bytes / 1e6 / ms.front());
        fflush(stdout);
    }

    // ------------------------------------------------------------------ ring -> tensor core, step by step
    {
        const uint32_t maxN = 128;
        __half* dB = nullptr;
        unsigned long long* d_sum = nullptr;
        CK(cudaMalloc(&dB, (size_t)maxN * dim * 2));
        CK(cudaMalloc(&d_sum, 8));
        fill_pattern_kernel<<<(unsigned)((rows * dim + 255) / 256), 256>>>(reinterpret_cast<__half*>(d), rows, dim, 131, 7, 17, 8);
        fill_pattern_kernel<<<(unsigned)(((size_t)maxN * dim + 255) / 256), 256>>>(dB, maxN, dim, 5, 3, 13, 6);
        CK(cudaDeviceSynchronize());
        const size_t smem_max = 232448 - 1024;
        CK(cudaFuncSetAttribute(pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        long long host_sum = 0;
        const bool host_check = rows <= 4096;
        printf("# pipe: %-28s %3s %3s %6s %7s %7s %9s %9s  %s\n", "shape", "N", "KB", "stages", "ring_KB", "consume", "best_ms", "best_GB/s", "checksum");
        for (uint32_t N : {16u, 64u, 128u}) {
            if (host_check) {
                host_sum = 0;
                for (uint64_t r = 0; r < rows; ++r)
                    for (uint32_t q = 0; q < N; ++q) {
                        long long acc = 0;
                        for (uint32_t c = 0; c < dim; ++c) acc += (long long)((int)((r * 131 + c * 7) % 17) - 8) * ((int)((q * 5ull + c * 3ull) % 13) - 6);
                        host_sum += acc;
                    }
            }
            CUtensorMap tmB;
            {
                cuuint64_t gd[2] = {dim, N};
                cuuint64_t gs[1] = {row_bytes};
                cuuint32_t box[2] = {64, N};
                cuuint32_t es[2] = {1, 1};
                CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { printf("# pipe: query map rejected (CUresult %d)\n", (int)r); continue; }
            }
            struct Shape { uint32_t mode, KB; };
            for (Shape sh : {Shape{T2, 1}, Shape{T4, 2}, Shape{T4, 4}}) {
                if (nkb % sh.KB) continue;
                CUtensorMap tmA;
                CUresult r;
                if (sh.mode == T2) {
                    cuuint64_t gd[2] = {dim, rows};
                    cuuint64_t gs[1] = {row_bytes};
                    cuuint32_t box[2] = {64, 128};
                    cuuint32_t es[2] = {1, 1};
                    r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                } else {
                    cuuint64_t gd[4] = {64, 8, nkb, rows / 8};
                    cuuint64_t gs[3] = {row_bytes, 128, 8ull * row_bytes};
                    cuuint32_t box[4] = {64, 8, sh.KB, 16};
                    cuuint32_t es[4] = {1, 1, 1, 1};
                    r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                }
                if (r != CUDA_SUCCESS) { printf("# pipe: row map rejected (CUresult %d)\n", (int)r); continue; }
                const uint32_t stage_bytes = 128 * sh.KB * 128;
                const size_t b_bytes = (size_t)nkb * N * 128;
                for (uint32_t ring_kb : {64u, 96u, 128u, 160u, 192u}) {
                    uint32_t stages = ring_kb * 1024 / stage_bytes;
                    if (stages < 2 || stages > kMaxStages) continue;
                    const size_t smem = b_bytes + (size_t)stages * stage_bytes + 1024;
                    if (smem > smem_max) continue;
                    for (uint32_t consume : {0u, 1u, 2u, 4u}) {
                        PipeParams pp{};
                        pp.n_tiles = rows / 128;
                        pp.mode = sh.mode;
                        pp.kb_per_load = sh.KB;
                        pp.groups = nkb / sh.KB;
                        pp.stage_bytes = stage_bytes;
                        pp.stages = stages;
                        pp.N = N;
                        pp.nkb = nkb;
                        pp.tmem_cols = 32;
                        while (pp.tmem_cols < 2 * N) pp.tmem_cols *= 2;
                        pp.consume = consume;
                        pp.checksum = d_sum;
                        std::vector<float> ms;
                        unsigned long long sum = 0;
                        for (int rep = 0; rep < 5; ++rep) {
                            CK(cudaMemset(d_sum, 0, 8));
                            CK(cudaEventRecord(e0));
                            pipe_kernel<<<sms, kPipeThreads, smem>>>(tmA, tmB, pp);
                            CK(cudaEventRecord(e1));
                            cudaError_t e = cudaEventSynchronize(e1);
                            if (e == cudaSuccess) e = cudaGetLastError();
                            if (e != cudaSuccess) { printf("# pipe: kernel failed: %s\n", cudaGetErrorString(e)); return 2; }
                            float t;
                            CK(cudaEventElapsedTime(&t, e0, e1));
                            if (rep >= 1) ms.push_back(t);
                            CK(cudaMemcpy(&sum, d_sum, 8, cudaMemcpyDeviceToHost));
                        }
                        std::sort(ms.begin(), ms.end());
                        char note[64] = "";
                        if (consume == 4 && host_check) snprintf(note, sizeof(note), (long long)sum == host_sum ? " == host" : " != host %lld", host_sum);
                        printf("  pipe: %-28s %3u %3u %6u %7u %7u %9.4f %9.1f  %lld%s\n", sh.mode == T2 ? "t2 box 64x128" : "t4 box 64x8xKBx16", N, sh.KB, stages,
                               stages * stage_bytes / 1024, consume, ms.front(), bytes / 1e6 / ms.front(), (long long)sum, note);
                        fflush(stdout);
                    }
                }
            }
        }
        cudaFree(dB);
        cudaFree(d_sum);
    }
    cudaFree(d);
    return 0;
}
