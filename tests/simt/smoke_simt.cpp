#define TSGPU_SIMT 1
#include "simt.h"
static void k(int* out) {
    __shared__ int sm[64];
    int t = threadIdx.x;
    int lane = t & 31;
    int v = t + 1;
    for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
    sm[t] = v;
    __syncthreads();
    unsigned b = __ballot_sync(0xffffffffu, t % 3 == 0);
    out[blockIdx.x * 64 + t] = sm[63 - t] + (int)__popc(b);
}
int main() {
    std::vector<int> out(128);
    simt::launch(dim3(2), dim3(64), 0, [&] { k(out.data()); });
    for (int i = 0; i < 8; i++) printf("%d ", out[i]);
    printf("\n%zu collectives\n", simt::g_collectives);
}
