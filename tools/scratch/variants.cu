// Standalone harness: variants of "XYZZ<Fq2> -> affine std" to localise a miscompile.
#include <cstdio>
#include <cuda_runtime.h>
#include "ec.cuh"
using namespace zkr;
typedef XYZZ<Fq2> P2;

__device__ P2 make_point() {
    // G2 generator (Montgomery) doubled 5 times via inline dbl => non-trivial zz
    const uint32_t gx0[8] = {0x8e83b5d1u,0x02a3e31fu,0x3b1f3f2eu,0x3fd2b2a4u,0x0b2e2bd8u,0x4c2dd7b5u,0xb1ac7d2du,0x19573841u};
    Affine<Fq2> g;
    // build from standard-form constants at runtime: x = (10857046999023057135944570762232829481370756359578518086990519993285655852781, 11559732032986387107991004021392285783925812861821192530917403151452391805634)
    const uint32_t x0[8] = {0xd992f6edu,0x46debd5cu,0xf75edaddu,0x674322d4u,0x5e5c4479u,0x426a0066u,0x121f1e76u,0x1800deefu};
    const uint32_t x1[8] = {0xaef312c2u,0x97e485b7u,0x35a9e712u,0xf1aa4933u,0x31fb5d25u,0x7260bfb7u,0x920d483au,0x198e9393u};
    const uint32_t y0[8] = {0x66fa7daau,0x4ce6cc01u,0x0c43d37bu,0xe3d1e769u,0x8dcb408fu,0x4aab7180u,0xdb8c6debu,0x12c85ea5u};
    const uint32_t y1[8] = {0xd122975bu,0x55acdadcu,0x70b38ef3u,0xbc4b3133u,0x690c3395u,0xec9e99adu,0x585ff075u,0x090689d0u};
    Fq a,b,c,d;
    for (int i=0;i<8;i++){a.v[i]=x0[i];b.v[i]=x1[i];c.v[i]=y0[i];d.v[i]=y1[i];}
    g.x = {a.to_mont(), b.to_mont()};
    g.y = {c.to_mont(), d.to_mont()};
    (void)gx0;
    P2 p = P2::from_affine(g);
    for (int i=0;i<5;i++) p = p.dbl();
    return p;
}
__global__ void k_make(P2* out) { make_point().store(out); }

// variant A: as in msm.cuh (noinline by value, then from_mont)
__global__ void vA(const P2* in, char* out) {
    Affine<Fq2> a = xyzz_to_affine_cold(P2::load(in));
    a.x.from_mont().store(out);
    a.y.from_mont().store(out + 64);
}
// variant B: fully inline
__global__ void vB(const P2* in, char* out) {
    Affine<Fq2> a = P2::load(in).to_affine();
    a.x.from_mont().store(out);
    a.y.from_mont().store(out + 64);
}
// variant C: noinline, store Montgomery; second kernel converts
__global__ void vC1(const P2* in, char* out) { xyzz_to_affine_cold(P2::load(in)).store(out); }
__global__ void vC2(const char* in, char* out) {
    int i = threadIdx.x;
    Fq::load(in + 32*i).from_mont().store(out + 32*i);
}
// variant D: noinline returning affine, then from_mont via to-one multiplication done differently
__global__ void vD(const P2* in, char* out) {
    Affine<Fq2> a = xyzz_to_affine_cold(P2::load(in));
    Fq one_std = Fq::zero(); one_std.v[0] = 1;
    (a.x.c0 * one_std).store(out);
    (a.x.c1 * one_std).store(out + 32);
    (a.y.c0 * one_std).store(out + 64);
    (a.y.c1 * one_std).store(out + 96);
}
// variant E: memory-to-memory out-of-line helper on GLOBAL pointers
__device__ __noinline__ void to_affine_mem(const P2* in, Affine<Fq2>* out) { P2::load(in).to_affine().store(out); }
__global__ void vE1(const P2* in, char* out) { to_affine_mem(in, (Affine<Fq2>*)out); }
// variant F: same through SHARED memory
__device__ __noinline__ void add_mem(P2* dst, const P2* src) { P2 a = *dst; a.add(*src); *dst = a; }
__device__ __noinline__ void dbl_mem(P2* dst) { P2 a = *dst; *dst = a.dbl(); }
__global__ void vF(const P2* in, P2* out) {
    __shared__ P2 sp[4];
    if (threadIdx.x == 0) { sp[0] = P2::load(in); sp[1] = sp[0]; }
    __syncthreads();
    if (threadIdx.x == 0) { dbl_mem(&sp[1]); add_mem(&sp[0], &sp[1]); add_mem(&sp[0], &sp[1]); sp[0].store(out); }   // 5P
}
__global__ void vFref(const P2* in, P2* out) {
    P2 a = P2::load(in); P2 d = a.dbl(); a.add(d); a.add(d); a.store(out);
}
int main() {
    P2* dp; char* dout; cudaMalloc(&dp, 256); cudaMalloc(&dout, 4096);
    cudaDeviceSetLimit(cudaLimitStackSize, 16384);
    k_make<<<1,1>>>(dp);
    uint32_t h[6][32];
    vA<<<1,1>>>(dp, dout); cudaMemcpy(h[0], dout, 128, cudaMemcpyDeviceToHost);
    vB<<<1,1>>>(dp, dout); cudaMemcpy(h[1], dout, 128, cudaMemcpyDeviceToHost);
    vC1<<<1,1>>>(dp, dout + 1024); vC2<<<1,4>>>(dout + 1024, dout); cudaMemcpy(h[2], dout, 128, cudaMemcpyDeviceToHost);
    vD<<<1,1>>>(dp, dout); cudaMemcpy(h[3], dout, 128, cudaMemcpyDeviceToHost);
    vE1<<<1,1>>>(dp, dout + 1024); vC2<<<1,4>>>(dout + 1024, dout); cudaMemcpy(h[4], dout, 128, cudaMemcpyDeviceToHost);
    { P2* o2; cudaMalloc(&o2, 512); uint32_t f[2][64];
      vF<<<1,32>>>(dp, o2); vFref<<<1,1>>>(dp, o2 + 1); cudaMemcpy(f, o2, 512, cudaMemcpyDeviceToHost);
      // compare affine of both via vB
      uint32_t g[2][32];
      vB<<<1,1>>>(o2, dout); cudaMemcpy(g[0], dout, 128, cudaMemcpyDeviceToHost);
      vB<<<1,1>>>(o2 + 1, dout); cudaMemcpy(g[1], dout, 128, cudaMemcpyDeviceToHost);
      int same = 1; for (int i = 0; i < 32; i++) same &= g[0][i] == g[1][i];
      printf("F(shared mem-to-mem)==inline: %d\n", same); }
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    const char* nm[5] = {"A","B","C","D","E"};
    for (int v=0; v<0; v++) { printf("%s:", nm[v]); for (int i=0;i<32;i++) printf(" %08x", h[v][i]); printf("\n"); }
    for (int v=0; v<5; v++) { int same=1; for (int i=0;i<32;i++) same &= (h[v][i]==h[1][i]); printf("%s==B: %d\n", nm[v], same); }
    return 0;
}
