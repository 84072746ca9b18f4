#include <cstdio>
// Characterise the ptxas 12.9 / sm_100a VIMNMX predicate-output miscompile.
__global__ void k(const int* a, const int* b, int* out) {
    const int i = threadIdx.x;
    int x = a[i], y = b[i];
    int r = 0;
    { int m = x < 2 ? x : 2; if (m == 2) r |= 1; }                  // T1: min(x,2)==2          expect x>=2
    { int m = x > 0 ? x : 0; if (m == 0) r |= 2; }                  // T2: max(x,0)==0          expect x<=0
    { int m = x < y ? x : y; if (m == x) r |= 4; }                  // T3: min(x,y)==x          expect x<=y
    { int m = x < y ? x : y; if (m == y) r |= 8; }                  // T4: min(x,y)==y          expect y<=x
    { int m = x < 0 ? 0 : (x > 2 ? 2 : x); if (m == 0) r |= 16; }   // T5: clamp==0             expect x<=0
    { int m = x < 0 ? 0 : (x > 2 ? 2 : x); if (m == 2) r |= 32; }   // T6: clamp==2             expect x>=2
    { int m = x < 0 ? 0 : (x > 2 ? 2 : x); if (m == 1) r |= 64; }   // T7: clamp==1             expect x==1
    { int m = x < 0 ? 0 : (x > 2 ? 2 : x); if (m != 0) r |= 128; }  // T8: clamp!=0             expect x>0
    out[i] = r;
}
int main() {
    const int n = 9;
    int ha[n] = {-5, -1, 0, 1, 2, 3, 7, 0, 2}, hb[n] = {0, -1, 5, 1, 1, 9, 7, -3, 2}, o[n];
    int *da, *db, *dout;
    cudaMalloc(&da, 4 * n); cudaMalloc(&db, 4 * n); cudaMalloc(&dout, 4 * n);
    cudaMemcpy(da, ha, 4 * n, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, 4 * n, cudaMemcpyHostToDevice);
    k<<<1, n>>>(da, db, dout);
    cudaMemcpy(o, dout, 4 * n, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < n; ++i) {
        const int x = ha[i], y = hb[i];
        const int c = x < 0 ? 0 : (x > 2 ? 2 : x);
        const int e = (x >= 2 ? 1 : 0) | (x <= 0 ? 2 : 0) | (x <= y ? 4 : 0) | (y <= x ? 8 : 0) | (c == 0 ? 16 : 0) | (c == 2 ? 32 : 0) |
                      (c == 1 ? 64 : 0) | (c != 0 ? 128 : 0);
        printf("x=%d y=%d got %3d expect %3d diff-bits %d\n", x, y, o[i], e, o[i] ^ e);
        bad |= o[i] ^ e;
    }
    printf("miscompiled tests mask: %d\n", bad);
    return 0;
}
