// Micro-benchmark: throughput of the accumulation primitives a scatter kernel can be built from on sm_100a.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

__device__ __forceinline__ uint32_t hash32(uint32_t x){ x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// MODE 0: smem u32 ATOMS.ADD random; 1: smem f32 atomicAdd (CAS loop) random; 2: smem plain RMW (racy, upper bound)
// 3: smem u32 atomics, 8 consecutive-ish addresses per "particle" (CIC-like pattern)
template<int MODE>
__global__ void k_smem(unsigned* out, int iters, int tile){
  extern __shared__ unsigned s[];
  for(int i=threadIdx.x;i<tile;i+=blockDim.x) s[i]=0;
  __syncthreads();
  uint32_t r = hash32(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int it=0; it<iters; it++){
    r = hash32(r+it);
    if(MODE==3){
      int T = 16; // tile 16^3 with halo 17 -> treat as 17^3 = 4913
      int x = r % 16, y = (r>>8)%16, z=(r>>16)%16;
      #pragma unroll
      for(int a=0;a<2;a++) for(int b=0;b<2;b++) for(int c=0;c<2;c++)
        atomicAdd(&s[((x+a)*17 + (y+b))*17 + z + c], r & 255);
      (void)T;
    } else {
      int idx = r % tile;
      if(MODE==0) atomicAdd(&s[idx], r & 255);
      else if(MODE==1) atomicAdd((float*)&s[idx], 1.0f);
      else s[idx] += r & 255;
    }
  }
  __syncthreads();
  unsigned acc=0;
  for(int i=threadIdx.x;i<tile;i+=blockDim.x) acc+=s[i];
  if(acc==0xdeadbeef) out[0]=acc;
}

// global REDG: MODE 0 f32, 1 f64, 2 f32x2, 3 u64 ; addresses random within `span` elements
template<int MODE>
__global__ void k_glob(void* buf, size_t span, int iters){
  uint32_t r = hash32(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int it=0; it<iters; it++){
    r = hash32(r+it);
    uint64_t rr = ((uint64_t)r << 16) ^ hash32(r);
    size_t idx = rr % span;
    if(MODE==0) atomicAdd(((float*)buf)+idx, 1.0f);
    else if(MODE==1) atomicAdd(((double*)buf)+idx, 1.0);
    else if(MODE==2) atomicAdd(((float2*)buf)+idx, make_float2(1.f,1.f));
    else atomicAdd(((unsigned long long*)buf)+idx, 1ull);
  }
}
// CIC-like global pattern: 8 REDs around a random cell of an N^3 mesh
template<typename T>
__global__ void k_glob_cic(T* buf, int N, int iters, int coherent){
  uint32_t r = hash32(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int it=0; it<iters; it++){
    r = hash32(r+it);
    int x,y,z;
    if(coherent){ // particles of a warp share a 16^3 neighbourhood that moves with the block
      uint32_t b = hash32(blockIdx.x*977+it/64);
      x = (b%N + (r&15))%N; y=((b>>10)%N + ((r>>4)&15))%N; z=((b>>20)%N + ((r>>8)&15))%N;
    } else { x=r%N; y=(r>>10)%N; z=(r>>20)%N; }
    #pragma unroll
    for(int a=0;a<2;a++) for(int b=0;b<2;b++) for(int c=0;c<2;c++){
      size_t idx = ((size_t)((x+a)%N)*N + (y+b)%N)*N + (z+c)%N;
      atomicAdd(buf+idx, (T)0.125);
    }
  }
}

template<typename F> float timeit(F f){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaDeviceSynchronize(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); return ms; }

int main(){
  int grid=148*8, block=256, iters=2000;
  unsigned* out; CK(cudaMalloc(&out, 4));
  double nops = (double)grid*block*iters;
  { float ms=timeit([&]{k_smem<0><<<grid,block,4096*4>>>(out,iters,4096);}); printf("smem u32 ATOMS random(4096): %.3f ms  %.2f Gop/s\n", ms, nops/ms/1e6);}
  { float ms=timeit([&]{k_smem<1><<<grid,block,4096*4>>>(out,iters,4096);}); printf("smem f32 CAS   random(4096): %.3f ms  %.2f Gop/s\n", ms, nops/ms/1e6);}
  { float ms=timeit([&]{k_smem<2><<<grid,block,4096*4>>>(out,iters,4096);}); printf("smem plain RMW random(4096): %.3f ms  %.2f Gop/s\n", ms, nops/ms/1e6);}
  { float ms=timeit([&]{k_smem<3><<<grid,block,4913*4>>>(out,iters/8,4913);}); printf("smem u32 ATOMS CIC-8 pattern: %.3f ms  %.2f Gop/s (x8 atomics per particle => %.2f Gpart/s)\n", ms, nops/ms/1e6, nops/8/ms/1e6);}
  CK(cudaGetLastError());
  size_t big = (size_t)1<<30; // bytes
  void* buf; CK(cudaMalloc(&buf, big)); CK(cudaMemset(buf,0,big));
  int giters=200; double gops=(double)grid*block*giters;
  for(size_t spanB : {(size_t)16<<20, (size_t)1<<30}){
    { float ms=timeit([&]{k_glob<0><<<grid,block>>>(buf, spanB/4, giters);}); printf("REDG f32   random span %4zu MB: %.3f ms %.2f Gop/s\n", spanB>>20, ms, gops/ms/1e6);}
    { float ms=timeit([&]{k_glob<1><<<grid,block>>>(buf, spanB/8, giters);}); printf("REDG f64   random span %4zu MB: %.3f ms %.2f Gop/s\n", spanB>>20, ms, gops/ms/1e6);}
    { float ms=timeit([&]{k_glob<2><<<grid,block>>>(buf, spanB/8, giters);}); printf("REDG f32x2 random span %4zu MB: %.3f ms %.2f Gop/s\n", spanB>>20, ms, gops/ms/1e6);}
    { float ms=timeit([&]{k_glob<3><<<grid,block>>>(buf, spanB/8, giters);}); printf("REDG u64   random span %4zu MB: %.3f ms %.2f Gop/s\n", spanB>>20, ms, gops/ms/1e6);}
  }
  for(int coh=0; coh<2; coh++){
    { float ms=timeit([&]{k_glob_cic<float><<<grid,block>>>((float*)buf, 512, giters, coh);}); printf("REDG f32 CIC 512^3 coherent=%d: %.3f ms %.2f Gpart/s\n", coh, ms, gops/ms/1e6);}
    { float ms=timeit([&]{k_glob_cic<double><<<grid,block>>>((double*)buf, 512, giters, coh);}); printf("REDG f64 CIC 512^3 coherent=%d: %.3f ms %.2f Gpart/s\n", coh, ms, gops/ms/1e6);}
  }
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  return 0;
}
