// microbench.cu -- instruction-throughput probes that inform the interaction-kernel design (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu ; run on the B200 box.
// Reports warp-instructions per clock per SM for the pipes the pair loop uses: FFMA, packed FFMA2 (f32x2),
// FADD, MUFU (rcp/rsqrt/ex2/lg2), broadcast LDS.128, per-lane LDS.128, predicated STS, SHFL.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 4096
#define CHECK(x) do{ cudaError_t e=(x); if(e!=cudaSuccess){ printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1; } }while(0)

template<int MODE> __global__ void __launch_bounds__(256) k_bench(float* out, float seed, unsigned long long* cycles){
  extern __shared__ float4 smem[];
  const int tid=threadIdx.x;
  for(int i=tid;i<2048;i+=blockDim.x)smem[i]=make_float4(seed+i,seed*i,seed-i,1.f+i);
  __syncthreads();
  float a0=seed+tid,a1=seed*2+tid,a2=seed*3,a3=seed*4,a4=seed*5,a5=seed*6,a6=seed*7,a7=seed*8;
  const float b=1.0001f,c=0.0001f;
  unsigned long long t0=clock64();
  if(MODE==0){ // FFMA x8 independent
    #pragma unroll 16
    for(int i=0;i<ITERS;i++){ a0=fmaf(a0,b,c); a1=fmaf(a1,b,c); a2=fmaf(a2,b,c); a3=fmaf(a3,b,c); a4=fmaf(a4,b,c); a5=fmaf(a5,b,c); a6=fmaf(a6,b,c); a7=fmaf(a7,b,c); }
  }
  else if(MODE==1){ // FFMA2 (fma.rn.f32x2) x4 independent pairs = 8 FMAs
    unsigned long long p0,p1,p2,p3,pb,pc;
    asm("mov.b64 %0,{%1,%2};":"=l"(p0):"f"(a0),"f"(a1)); asm("mov.b64 %0,{%1,%2};":"=l"(p1):"f"(a2),"f"(a3));
    asm("mov.b64 %0,{%1,%2};":"=l"(p2):"f"(a4),"f"(a5)); asm("mov.b64 %0,{%1,%2};":"=l"(p3):"f"(a6),"f"(a7));
    asm("mov.b64 %0,{%1,%2};":"=l"(pb):"f"(b),"f"(b)); asm("mov.b64 %0,{%1,%2};":"=l"(pc):"f"(c),"f"(c));
    #pragma unroll 16
    for(int i=0;i<ITERS;i++){
      asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p0):"l"(pb),"l"(pc));
      asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p1):"l"(pb),"l"(pc));
      asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p2):"l"(pb),"l"(pc));
      asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p3):"l"(pb),"l"(pc));
    }
    asm("mov.b64 {%0,%1},%2;":"=f"(a0),"=f"(a1):"l"(p0)); asm("mov.b64 {%0,%1},%2;":"=f"(a2),"=f"(a3):"l"(p1));
    asm("mov.b64 {%0,%1},%2;":"=f"(a4),"=f"(a5):"l"(p2)); asm("mov.b64 {%0,%1},%2;":"=f"(a6),"=f"(a7):"l"(p3));
  }
  else if(MODE==2){ // FADD x8
    #pragma unroll 16
    for(int i=0;i<ITERS;i++){ a0+=c; a1+=c; a2+=c; a3+=c; a4+=c; a5+=c; a6+=c; a7+=c; }
  }
  else if(MODE==3){ // MUFU.RCP x8
    #pragma unroll 16
    for(int i=0;i<ITERS;i++){
      asm volatile("rcp.approx.ftz.f32 %0,%0;":"+f"(a0)); asm volatile("rcp.approx.ftz.f32 %0,%0;":"+f"(a1));
      asm volatile("rcp.approx.ftz.f32 %0,%0;":"+f"(a2)); asm volatile("rcp.approx.ftz.f32 %0,%0;":"+f"(a3));
      asm volatile("rcp.approx.ftz.f32 %0,%0;":"+f"(a4)); asm volatile("rcp.approx.ftz.f32 %0,%0;":"+f"(a5));
      asm volatile("rcp.approx.ftz.f32 %0,%0;":"+f"(a6)); asm volatile("rcp.approx.ftz.f32 %0,%0;":"+f"(a7));
    }
  }
  else if(MODE==4){ // MUFU.RSQ x8
    #pragma unroll 16
    for(int i=0;i<ITERS;i++){
      asm volatile("rsqrt.approx.ftz.f32 %0,%0;":"+f"(a0)); asm volatile("rsqrt.approx.ftz.f32 %0,%0;":"+f"(a1));
      asm volatile("rsqrt.approx.ftz.f32 %0,%0;":"+f"(a2)); asm volatile("rsqrt.approx.ftz.f32 %0,%0;":"+f"(a3));
      asm volatile("rsqrt.approx.ftz.f32 %0,%0;":"+f"(a4)); asm volatile("rsqrt.approx.ftz.f32 %0,%0;":"+f"(a5));
      asm volatile("rsqrt.approx.ftz.f32 %0,%0;":"+f"(a6)); asm volatile("rsqrt.approx.ftz.f32 %0,%0;":"+f"(a7));
    }
  }
  else if(MODE==5){ // broadcast LDS.128 (all lanes same address) + 1 FADD per load to keep it live
    #pragma unroll 16
    for(int i=0;i<ITERS;i++){
      const float4 v0=smem[(i*2+0)&2047], v1=smem[(i*2+1)&2047];
      a0+=v0.x; a1+=v0.y; a2+=v0.z; a3+=v0.w; a4+=v1.x; a5+=v1.y; a6+=v1.z; a7+=v1.w;
    }
  }
  else if(MODE==6){ // per-lane LDS.128, consecutive addresses (conflict free, 4 wavefronts)
    #pragma unroll 16
    for(int i=0;i<ITERS;i++){
      const float4 v0=smem[(i*2+0+tid)&2047], v1=smem[(i*2+1+tid)&2047];
      a0+=v0.x; a1+=v0.y; a2+=v0.z; a3+=v0.w; a4+=v1.x; a5+=v1.y; a6+=v1.z; a7+=v1.w;
    }
  }
  else if(MODE==7){ // per-lane LDS.128 at pseudo-random addresses (gather with bank conflicts)
    unsigned h=tid*2654435761u;
    #pragma unroll 8
    for(int i=0;i<ITERS;i++){
      h=h*1664525u+1013904223u; const float4 v0=smem[(h>>8)&2047];
      h=h*1664525u+1013904223u; const float4 v1=smem[(h>>8)&2047];
      a0+=v0.x; a1+=v0.y; a2+=v0.z; a3+=v0.w; a4+=v1.x; a5+=v1.y; a6+=v1.z; a7+=v1.w;
    }
  }
  else if(MODE==8){ // pair-test body: broadcast LDS.128 + 3 FADD + FMUL + 2 FFMA + compare + predicated counter
    int cnt=0; const float xi=a0,yi=a1,zi=a2,r2max=a3;
    #pragma unroll 16
    for(int i=0;i<ITERS*8;i++){
      const float4 v=smem[i&2047];
      const float dx=xi-v.x, dy=yi-v.y, dz=zi-v.z;
      const float r2=fmaf(dz,dz,fmaf(dy,dy,dx*dx));
      cnt+=(r2<=r2max && r2>=1e-18f)? 1: 0;
    }
    a0=float(cnt);
  }
  else if(MODE==9){ // MUFU.EX2+LG2 pairs x4
    #pragma unroll 16
    for(int i=0;i<ITERS;i++){
      asm volatile("lg2.approx.ftz.f32 %0,%0;":"+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0,%0;":"+f"(a0));
      asm volatile("lg2.approx.ftz.f32 %0,%0;":"+f"(a1)); asm volatile("ex2.approx.ftz.f32 %0,%0;":"+f"(a1));
      asm volatile("lg2.approx.ftz.f32 %0,%0;":"+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0,%0;":"+f"(a2));
      asm volatile("lg2.approx.ftz.f32 %0,%0;":"+f"(a3)); asm volatile("ex2.approx.ftz.f32 %0,%0;":"+f"(a3));
    }
  }
  else if(MODE==10){ // FFMA x6 + MUFU.RCP x1 interleaved (pair-body like mix)
    #pragma unroll 16
    for(int i=0;i<ITERS;i++){
      a0=fmaf(a0,b,c); a1=fmaf(a1,b,c); a2=fmaf(a2,b,c); asm volatile("rcp.approx.ftz.f32 %0,%0;":"+f"(a6));
      a3=fmaf(a3,b,c); a4=fmaf(a4,b,c); a5=fmaf(a5,b,c); asm volatile("rsqrt.approx.ftz.f32 %0,%0;":"+f"(a7));
    }
  }
  unsigned long long t1=clock64();
  out[blockIdx.x*blockDim.x+tid]=a0+a1+a2+a3+a4+a5+a6+a7;
  if(tid==0)cycles[blockIdx.x]=t1-t0;
}

struct Res { const char* name; double winst_per_iter; };

template<int MODE> static int run(const char* name,double winst_per_iter,int nsm,float* dout,unsigned long long* dcyc){
  const int threads=256, blocks=nsm*4;   // 4 CTAs x 8 warps per SM = 32 warps/SM
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_bench<MODE><<<blocks,threads,2048*16>>>(dout,1.0f,dcyc);
  CHECK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  k_bench<MODE><<<blocks,threads,2048*16>>>(dout,1.0f,dcyc);
  cudaEventRecord(e1);
  CHECK(cudaDeviceSynchronize());
  float ms=0; cudaEventElapsedTime(&ms,e0,e1);
  unsigned long long hc[4096]; cudaMemcpy(hc,dcyc,sizeof(unsigned long long)*blocks,cudaMemcpyDeviceToHost);
  double cyc=0; for(int i=0;i<blocks;i++)cyc+=double(hc[i]); cyc/=blocks;
  const double winst=double(ITERS)*winst_per_iter*(threads/32)*4;   // warp-instructions per SM (4 CTAs)
  printf("%-34s %8.3f ms  %10.0f cyc/CTA  %6.3f warp-inst/clk/SM  (%.2f GHz eff)\n",name,ms,cyc,winst/cyc,cyc/(ms*1e6));
  return 0;
}

int main(){
  int dev=0; cudaDeviceProp pr; CHECK(cudaGetDeviceProperties(&pr,dev));
  printf("device: %s  SMs=%d  clock=%d kHz\n",pr.name,pr.multiProcessorCount,pr.clockRate);
  const int nsm=pr.multiProcessorCount;
  float* dout; unsigned long long* dcyc;
  CHECK(cudaMalloc(&dout,sizeof(float)*nsm*4*256)); CHECK(cudaMalloc(&dcyc,sizeof(unsigned long long)*4096));
  run<0>("FFMA (8 indep)",8,nsm,dout,dcyc);
  run<1>("FFMA2 f32x2 (4 indep = 8 FMA)",4,nsm,dout,dcyc);
  run<2>("FADD (8 indep)",8,nsm,dout,dcyc);
  run<3>("MUFU.RCP (8 indep)",8,nsm,dout,dcyc);
  run<4>("MUFU.RSQ (8 indep)",8,nsm,dout,dcyc);
  run<9>("MUFU LG2+EX2 (4 pairs)",8,nsm,dout,dcyc);
  run<5>("LDS.128 broadcast: 2 LDS+8 FADD",10,nsm,dout,dcyc);
  run<6>("LDS.128 consecutive: 2 LDS+8 FADD",10,nsm,dout,dcyc);
  run<7>("LDS.128 random: 2 LDS+8 FADD+LCG",16,nsm,dout,dcyc);
  run<8>("pair test (per candidate ~10 inst)",8*10,nsm,dout,dcyc);
  run<10>("6 FFMA + 2 MUFU mix",8,nsm,dout,dcyc);
  return 0;
}
