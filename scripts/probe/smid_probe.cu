// How does the block scheduler spread a grid over the SMs?  Prints, for several grid sizes, the
// range of %smid values and the histogram of CTAs per SM (CTAs stay resident ~20 us).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128) probe(unsigned int *count, int spin)
{
    __shared__ char pad[23 * 1024];
    unsigned int sm;
    asm("mov.u32 %0, %%smid;" : "=r"(sm));
    if (threadIdx.x == 0) { atomicAdd(&count[sm], 1u); pad[0] = 1; }
    long long t0 = clock64();
    while (clock64() - t0 < spin) { }
    if (pad[threadIdx.x] == 77) count[255] = 1;
}
int main()
{
    unsigned int *d, h[256];
    cudaMalloc(&d, 256 * 4);
    unsigned int nsmid; 
    int grids[] = {148, 296, 592, 1036, 1332, 2049};
    for (int g : grids) {
        cudaMemset(d, 0, 256 * 4);
        probe<<<g, 128>>>(d, 40000);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, 256 * 4, cudaMemcpyDeviceToHost);
        int lo = 1 << 30, hi = 0, used = 0, maxid = 0, hist[64] = {0};
        for (int i = 0; i < 255; i++) if (h[i]) { used++; maxid = i; if ((int)h[i] < lo) lo = h[i]; if ((int)h[i] > hi) hi = h[i]; hist[h[i] < 63 ? h[i] : 63]++; }
        printf("grid %5d: SMs used %d, max smid %d, CTAs/SM min %d max %d | hist:", g, used, maxid, lo, hi);
        for (int i = 0; i < 64; i++) if (hist[i]) printf(" %dx%d", hist[i], i);
        printf("\n");
    }
    return 0;
}
