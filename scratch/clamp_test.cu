#include <cstdio>
__global__ void k(const int* a, int* out, const int* len) {
    int x = a[threadIdx.x];
    x = x < 0 ? 0 : (x > 2 ? 2 : x);
    int r;
    int l = len[threadIdx.x];
    if (x == 0) { r = 10 + l; }
    else if (x == 2) { r = 20 + 3 * l; }
    else r = 30 - l;
    out[threadIdx.x] = r * 100 + x;
}
int main() {
    int h[8] = {-5, -1, 0, 1, 2, 3, 7, 0}, l[8] = {0, 0, 0, 0, 0, 0, 0, 0}, o[8];
    int *d, *dl, *dout;
    cudaMalloc(&d, 32); cudaMalloc(&dl, 32); cudaMalloc(&dout, 32);
    cudaMemcpy(d, h, 32, cudaMemcpyHostToDevice); cudaMemcpy(dl, l, 32, cudaMemcpyHostToDevice);
    k<<<1, 8>>>(d, dout, dl);
    cudaMemcpy(o, dout, 32, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 8; ++i) printf("a=%d -> %d (expect %d)\n", h[i], o[i], (h[i] <= 0 ? 1000 : h[i] >= 2 ? 2002 : 3001));
    return 0;
}
