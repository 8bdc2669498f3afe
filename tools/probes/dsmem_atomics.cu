// probe: 64-bit atomicMax / atomicAdd / 32-bit atomicOr on distributed shared memory (cluster of 8), generic addressing
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(256)
k(unsigned long long* out_max, unsigned long long* out_add, unsigned* out_or, long long* cyc)
{
    __shared__ unsigned long long zb[64];
    __shared__ unsigned long long acc[64];
    __shared__ unsigned bits[64];
    cg::cluster_group cl = cg::this_cluster();
    const int r = cl.block_rank();
    if (threadIdx.x < 64) { zb[threadIdx.x] = 0ull; acc[threadIdx.x] = 0ull; bits[threadIdx.x] = 0u; }
    cl.sync();
    long long t0 = clock64();
    // every thread of every CTA hits every owner's slot (threadIdx & 63)
    for (int o = 0; o < 8; ++o) {
        unsigned long long* z = cl.map_shared_rank(zb, o);
        unsigned long long* a = cl.map_shared_rank(acc, o);
        unsigned* b = cl.map_shared_rank(bits, o);
        const int s = threadIdx.x & 63;
        atomicMax(z + s, ((unsigned long long)(r * 256 + threadIdx.x) << 32) | (unsigned)(~threadIdx.x));
        atomicAdd(a + s, (1ull << 16) + 1ull);
        atomicOr(b + s, 1u << (r * 4 + (threadIdx.x >> 6)));
    }
    long long t1 = clock64();
    cl.sync();
    if (threadIdx.x < 64) {
        const int g = (blockIdx.x) * 64 + threadIdx.x;
        out_max[g] = zb[threadIdx.x]; out_add[g] = acc[threadIdx.x]; out_or[g] = bits[threadIdx.x];
    }
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ unsigned mapa(const void* p, unsigned rank) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(256)
k2(unsigned long long* out_max, unsigned long long* out_add, unsigned* out_or, long long* cyc)
{
    __shared__ unsigned long long zb[64];
    __shared__ unsigned long long acc[64];
    __shared__ unsigned bits[64];
    cg::cluster_group cl = cg::this_cluster();
    const int r = cl.block_rank();
    if (threadIdx.x < 64) { zb[threadIdx.x] = 0ull; acc[threadIdx.x] = 0ull; bits[threadIdx.x] = 0u; }
    cl.sync();
    long long t0 = clock64();
    for (int o = 0; o < 8; ++o) {
        const int s = threadIdx.x & 63;
        const unsigned z = mapa(zb + s, o), a = mapa(acc + s, o), b = mapa(bits + s, o);
        const unsigned long long v = ((unsigned long long)(r * 256 + threadIdx.x) << 32) | (unsigned)(~threadIdx.x);
        unsigned long long old;
        asm volatile("atom.shared::cluster.max.u64 %0, [%1], %2;" : "=l"(old) : "r"(z), "l"(v) : "memory");
        asm volatile("atom.shared::cluster.add.u64 %0, [%1], %2;" : "=l"(old) : "r"(a), "l"((1ull << 16) + 1ull) : "memory");
        unsigned o32;
        asm volatile("atom.shared::cluster.or.b32 %0, [%1], %2;" : "=r"(o32) : "r"(b), "r"(1u << (r * 4 + (threadIdx.x >> 6))) : "memory");
    }
    long long t1 = clock64();
    cl.sync();
    if (threadIdx.x < 64) {
        const int g = (blockIdx.x) * 64 + threadIdx.x;
        out_max[g] = zb[threadIdx.x]; out_add[g] = acc[threadIdx.x]; out_or[g] = bits[threadIdx.x];
    }
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int run(int which) {
    unsigned long long *dm, *da; unsigned* dor; long long* dc;
    const int nblk = 8 * 4;
    cudaMalloc(&dm, nblk * 64 * 8); cudaMalloc(&da, nblk * 64 * 8); cudaMalloc(&dor, nblk * 64 * 4); cudaMalloc(&dc, nblk * 8);
    if (which == 0) k<<<nblk, 256>>>(dm, da, dor, dc); else k2<<<nblk, 256>>>(dm, da, dor, dc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    unsigned long long hm[64 * 32], ha[64 * 32]; unsigned ho[64 * 32]; long long hc[32];
    cudaMemcpy(hm, dm, sizeof(hm), cudaMemcpyDeviceToHost); cudaMemcpy(ha, da, sizeof(ha), cudaMemcpyDeviceToHost);
    cudaMemcpy(ho, dor, sizeof(ho), cudaMemcpyDeviceToHost); cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost);
    int bad = 0, bm = 0, ba = 0, bo = 0;
    for (int g = 0; g < 64 * 32; ++g) {
        const int s = g & 63;
        // max: rank 7, thread 192 + s  (largest r*256+t with t&63 == s)
        const unsigned long long want = ((unsigned long long)(7 * 256 + 192 + s) << 32) | (unsigned)(~(192 + s));
        if (hm[g] != want) { ++bad; if (bm++ < 4) printf("max g=%d got %llx want %llx\n", g, hm[g], want); }
        if (ha[g] != 32ull * ((1ull << 16) + 1ull)) { ++bad; if (ba++ < 4) printf("add g=%d got %llx want %llx\n", g, ha[g], 32ull * ((1ull << 16) + 1ull)); }
        if (ho[g] != 0xffffffffu) { ++bad; if (bo++ < 4) printf("or g=%d got %x\n", g, ho[g]); }
    }
    printf("bad max %d add %d or %d\n", bm, ba, bo);
    printf("variant %d dsmem atomics: %s (bad=%d); cycles per CTA for 24 atomics/thread: %lld\n", which, bad ? "FAIL" : "OK", bad, hc[0]);
    return bad != 0;
}
int main() { int a = run(0); int b = run(1); return a | b; }
