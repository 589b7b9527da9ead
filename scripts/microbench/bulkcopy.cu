// Microbenchmark (profiling aid, not part of the library): throughput / latency of cp.async.bulk global->shared
// from an L2-resident buffer, per SM, as a function of copy size and copies in flight.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
__device__ __forceinline__ uint32_t s32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c){ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"::"r"(b),"r"(c)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity){
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"::"r"(bar),"r"(parity):"memory"); }
__device__ __forceinline__ void copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar){
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(bar),"r"(bytes):"memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"::"r"(dst),"l"(src),"r"(bytes),"r"(bar):"memory"); }
// one thread per CTA issues `n` copies of `bytes` with `depth` in flight (ring of depth stages)
__global__ void k(const uint8_t* src, uint32_t src_bytes, uint32_t bytes, int n, int depth, unsigned long long* out){
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[8];
  if (threadIdx.x==0){ for(int i=0;i<8;i++) mbar_init(s32(&bars[i]),1); asm volatile("fence.mbarrier_init.release.cluster;":::"memory"); }
  __syncthreads();
  if (threadIdx.x==0){
    unsigned long long t0=clock64();
    uint32_t off=(blockIdx.x*4096u)%src_bytes;
    for(int i=0;i<n+depth;i++){
      if(i>=depth){ int j=i-depth; mbar_wait(s32(&bars[j%depth]), (j/depth)&1); }
      if(i<n){ copy(s32(smem)+(i%depth)*bytes, src+off, bytes, s32(&bars[i%depth])); off+=bytes; if(off+bytes>src_bytes) off=0; }
    }
    unsigned long long t1=clock64();
    out[blockIdx.x]=t1-t0;
  }
}
int main(int argc,char**argv){
  const uint32_t SRC=768*1024; uint8_t* src; cudaMalloc(&src,SRC); cudaMemset(src,1,SRC);
  unsigned long long* out; cudaMalloc(&out,148*8); unsigned long long h[148];
  cudaFuncSetAttribute(k,cudaFuncAttributeMaxDynamicSharedMemorySize,200*1024);
  int grids[]={1,148};
  uint32_t sizes[]={4096,20480,40960};
  for(int gi=0;gi<2;gi++) for(int si=0;si<3;si++) for(int depth=1;depth<=4;depth++){
    uint32_t bytes=sizes[si]; if((size_t)bytes*depth>200*1024) continue; int n=64;
    for(int rep=0;rep<2;rep++){ k<<<grids[gi],32,bytes*depth>>>(src,SRC,bytes,n,depth,out); cudaDeviceSynchronize(); }
    cudaMemcpy(h,out,grids[gi]*8,cudaMemcpyDeviceToHost);
    double mx=0,av=0; for(int i=0;i<grids[gi];i++){ av+=h[i]; if(h[i]>mx) mx=h[i]; } av/=grids[gi];
    printf("grid %3d bytes %6u depth %d: %.0f clk per copy (max-CTA %.0f), %.1f B/clk/SM\n",grids[gi],bytes,depth,av/n,mx/n,(double)bytes*n/av);
  }
  cudaError_t e=cudaGetLastError(); printf("status %s\n",cudaGetErrorString(e));
  return 0;
}
