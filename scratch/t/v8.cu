#include <cuda_runtime.h>
__global__ void k(const float* __restrict__ p, float* o, const int* idx) {
    int i = idx[threadIdx.x];
    float a,b,c,d,e,f,g,h;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a),"=f"(b),"=f"(c),"=f"(d),"=f"(e),"=f"(f),"=f"(g),"=f"(h) : "l"(p + (size_t)i*8));
    o[threadIdx.x] = a+b+c+d+e+f+g+h;
}
