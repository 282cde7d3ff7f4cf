// How long does __nanosleep(t) really suspend a warp on B200?  (cycles measured with clock64, 1 warp)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long *out, unsigned t, int reps)
{
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) __nanosleep(t);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / reps;
}
int main()
{
    long long *d, h; cudaMalloc(&d, 8);
    for (unsigned t : {0u, 20u, 50u, 100u, 200u, 400u, 800u, 1000u, 1500u, 2000u, 4000u, 10000u}) {
        k<<<1, 32>>>(d, t, 200); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("nanosleep(%5u) -> %6lld cycles (%.0f ns at 1.965 GHz)\n", t, h, h / 1.965);
    }
    return 0;
}
